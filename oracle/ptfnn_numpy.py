"""TEST INFRASTRUCTURE ONLY (oracle) -- float64 NumPy restatement of the reference hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (ptnn_b200 + libptfnn.so) never does.

It restates, in a deterministic single-process form that takes every random draw as an INPUT,
what the reference does across R forked processes:

  R: = /root/reference/multicore-pt-regression/pt_timeseries_regression.py
  C: = /root/reference/multicore-pt-classification/pt_classification.py

Arithmetic is deliberately done the way the reference does it (row-by-row ``dot``/``exp`` on
tiny vectors, float64) so that (a) agreement with the live reference is at round-off level --
pinned by tests/test_oracle_golden.py against fixtures produced by oracle/gen_golden.py from the
unmodified reference -- and (b) its cost profile is the reference's, which is what makes it a
fair ``cpu_baseline`` ("port") on a box where /root/reference does not exist.

Parity status: PINNED against outputs of the reference itself run in the build container
(tests/golden/*.npz, generator committed).  The reference has no tests of its own.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

REGRESSION, CLASSIFICATION = 0, 1


def num_params(topology) -> int:
    i, h, o = topology
    return i * h + h * o + h + o            # R:494, C:504


# --------------------------------------------------------------------------------------
# a1-a5: Network
# --------------------------------------------------------------------------------------
class Network:
    """R:27-134 / C:26-153.  Weight layout (a1): [W1 (I x H row-major), W2 (H x O), B1 (H), B2 (O)]."""

    def __init__(self, topology, learn_rate, task):
        self.Top = list(topology)
        self.lrate = learn_rate
        self.task = task

    @staticmethod
    def sigmoid(x):                          # R:43-44
        return 1 / (1 + np.exp(-x))

    def decode(self, w):                     # R:80-90 (views into w: BackwardPass mutates w)
        I, H, O = self.Top
        a, b = I * H, I * H + H * O
        self.W1 = w[0:a].reshape(I, H)
        self.W2 = w[a:b].reshape(H, O)
        self.B1 = w[b:b + H]
        self.B2 = w[b + H:b + H + O]

    def forward(self, x):                    # R:51-55 -- bias is SUBTRACTED, sigmoid on both layers
        z1 = x.dot(self.W1) - self.B1
        self.hidout = self.sigmoid(z1)
        z2 = self.hidout.dot(self.W2) - self.B2
        self.out = self.sigmoid(z2)

    def backward(self, x, desired):          # R:57-78 / C:72-82 -- one online-SGD step on 1/2|y-out|^2
        if self.task == CLASSIFICATION:      # C:73-75 one-hot of the integer label
            onehot = np.zeros(self.Top[2])
            onehot[int(desired[0])] = 1
            desired = onehot
        out_delta = (desired - self.out) * (self.out * (1 - self.out))
        hid_delta = out_delta.dot(self.W2.T) * (self.hidout * (1 - self.hidout))   # pre-update W2
        lr = self.lrate
        if self.task == REGRESSION:
            # R:66-78 updates element by element in Python loops; kept that way so that this port
            # has the reference's cost profile when it serves as the CPU baseline (same values).
            I, H, O = self.Top
            for a in range(H):
                for b in range(O):
                    self.W2[a, b] += lr * out_delta[b] * self.hidout[a]
            for b in range(O):
                self.B2[b] += -1 * lr * out_delta[b]
            for a in range(I):
                for b in range(H):
                    self.W1[a, b] += lr * hid_delta[b] * x[a]
            for b in range(H):
                self.B1[b] += -1 * lr * hid_delta[b]
        else:                                                   # C:78-82 is vectorised in the reference too
            self.W2 += np.outer(self.hidout, lr * out_delta)
            self.B2 += -1 * lr * out_delta
            self.W1 += np.outer(x, lr * hid_delta)
            self.B1 += -1 * lr * hid_delta

    def langevin_gradient(self, data, w, depth=1):             # R:99-118 / C:114-132
        w = np.array(w, dtype=np.float64, copy=True)
        self.decode(w)
        I = self.Top[0]
        for _ in range(depth):
            for r in range(data.shape[0]):
                x = data[r, 0:I]
                self.forward(x)
                self.backward(x, data[r, I:])
        return w                                                # encode() of views == w itself

    def evaluate_proposal(self, data, w):                      # R:120-134 / C:134-153
        w = np.asarray(w, dtype=np.float64)
        self.decode(w)
        I, _, O = self.Top
        n = data.shape[0]
        fx = np.zeros(n)
        prob = np.zeros((n, O))
        for r in range(n):
            self.forward(data[r, 0:I])
            if self.task == REGRESSION:
                fx[r] = self.out[0]                             # R:132 (O must be 1)
            else:
                fx[r] = np.argmax(self.out)                     # C:55, C:147
                e = np.exp(self.out)                            # C:108-110 softmax OF the sigmoids
                prob[r] = e / np.sum(e)
        if self.task == REGRESSION:
            return fx
        return fx, prob


# --------------------------------------------------------------------------------------
# a6-a8: likelihood / prior
# --------------------------------------------------------------------------------------
def rmse(pred, actual):                                         # R:179-181
    return float(np.sqrt(((pred - actual) ** 2).mean()))


def accuracy(pred, actual):                                     # C:200-207
    return 100.0 * (float(np.count_nonzero(pred == actual)) / pred.shape[0])


def likelihood_regression(net, data, w, tau_sq, adapttemp):    # R:200-205
    y = data[:, net.Top[0]]
    fx = net.evaluate_proposal(data, w)
    loss = np.sum(-0.5 * np.log(2 * math.pi * tau_sq) - 0.5 * np.square(y - fx) / tau_sq)
    return float(loss) / adapttemp, fx, rmse(fx, y)


def likelihood_classification(net, data, w, adapttemp):        # C:209-222
    y = data[:, net.Top[0]]
    fx, prob = net.evaluate_proposal(data, w)
    lhood = 0.0
    for i in range(data.shape[0]):
        lhood += math.log(prob[i, int(y[i])])                   # z[i,j]=1 iff j==y[i]
    return lhood / adapttemp, fx, rmse(fx, y)


def prior_regression(sigma_squared, nu_1, nu_2, w, tausq, topology):     # R:215-221
    d, h = topology[0], topology[1]
    part1 = -1 * ((d * h + h + 2) / 2) * np.log(sigma_squared)
    part2 = 1 / (2 * sigma_squared) * (sum(np.square(w)))
    return float(part1 - part2 - (1 + nu_1) * np.log(tausq) - (nu_2 / tausq))


def prior_classification(sigma_squared, nu_1, nu_2, w, topology):         # C:224-230
    d, h, o = topology
    part1 = -1 * ((d * h + h + o + h * o) / 2) * np.log(sigma_squared)
    part2 = 1 / (2 * sigma_squared) * (sum(np.square(w)))
    return float(part1 - part2)


# --------------------------------------------------------------------------------------
# a14: ladder
# --------------------------------------------------------------------------------------
def geometric_ladder(num_chains: int, maxtemp) -> np.ndarray:
    """R:529-636: betas = logspace(0, -log10(Tmax), ntemps); T = 1/beta (so T_k = Tmax^(k/(R-1)))."""
    betas = np.logspace(0, -np.log10(maxtemp), num_chains)
    return 1.0 / betas


# --------------------------------------------------------------------------------------
# a13: swap
# --------------------------------------------------------------------------------------
def swap_probability(lhood1, lhood2):                           # R:674
    try:
        return min(1, 0.5 * np.exp(min(709, lhood2 - lhood1)))
    except OverflowError:
        return 1


def swap_sweep(lhood, u_row):
    """Sequential sweep over pairs (0,1),(1,2),... (R:741-748): returns ``src`` such that slot k
    ends up with the vector that was in slot ``src[k]``, plus the per-pair decisions."""
    lhood = list(lhood)
    src = list(range(len(lhood)))
    swapped = []
    for k in range(len(lhood) - 1):
        p = swap_probability(lhood[k], lhood[k + 1])
        s = bool(u_row[k] < p)                                  # R:677-679
        if s:
            lhood[k], lhood[k + 1] = lhood[k + 1], lhood[k]
            src[k], src[k + 1] = src[k + 1], src[k]
        swapped.append(s)
    return src, swapped


def swap_sweep_ratio_temperature(lhood, u_row, temperatures):
    """The same sweep under the swap rule of the reference's drafts (Misc/ldpt_fnn_multi_fixed.py:505-531):
    swap_proposal = (lhood1 / [1 if lhood2 == 0 else lhood2]) * (1/T1 * 1/T2), swap when u < swap_proposal; the
    temperature field travels with its vector (param[num_param + 2]).  No published result uses it."""
    lhood, temps = list(lhood), list(temperatures)
    src = list(range(len(lhood)))
    swapped = []
    for k in range(len(lhood) - 1):
        l1, l2, t1, t2 = lhood[k], lhood[k + 1], temps[k], temps[k + 1]
        with np.errstate(all="ignore"):
            p = (np.float64(l1) / (1.0 if l2 == 0 else np.float64(l2))) * (1.0 / t1 * 1.0 / t2)
        s = bool(u_row[k] < p)
        if s:
            lhood[k], lhood[k + 1] = l2, l1
            temps[k], temps[k + 1] = t2, t1
            src[k], src[k + 1] = src[k + 1], src[k]
        swapped.append(s)
    return src, swapped


# --------------------------------------------------------------------------------------
# configuration, draws, traces
# --------------------------------------------------------------------------------------
@dataclass
class PTConfig:
    task: int
    topology: tuple
    samples: int                 # S = int(NumSample / num_chains)  (R:506)
    swap_interval: int
    use_langevin_gradients: bool = True
    l_prob: float = 0.5          # R:174 ctor arg; C:192 fixed 0.5
    learn_rate: float = 0.1
    step_w: float = 0.025        # R:258
    step_eta: float = 0.2        # R:260
    sigma_squared: float = 25.0  # R:273
    nu_1: float = 0.0
    nu_2: float = 0.0
    pt_fraction: float = 0.6     # R:301

    @property
    def P(self):
        return num_params(self.topology)

    def swap_due(self, i: int) -> bool:
        if self.task == REGRESSION:                             # R:427
            return i % self.swap_interval == 0 and i != 0
        return (i + 1) % self.swap_interval == 0                # C:438

    def main_rounds(self) -> int:                               # R:719
        return int(self.samples / self.swap_interval)

    def inloop_rounds(self) -> int:
        return sum(1 for i in range(self.samples - 1) if self.swap_due(i))

    def total_rounds(self) -> int:
        """Rounds the coordinator really executes: every in-loop hand-shake, plus -- when the
        coordinator's ``int(S/s)`` loop has one round left over -- ONE round on the final-state
        vectors the replicas put at exit (R:442-444, R:485).  A second left-over round
        would find every replica dead and ``break`` (R:721-727)."""
        n_r, n_m = self.inloop_rounds(), self.main_rounds()
        return n_r + (1 if n_m > n_r else 0)


@dataclass
class Draws:
    lx: np.ndarray       # [R, S-1]        R:327  uniform
    z: np.ndarray        # [R, S-1, P]     R:331/353 standard normals (legacy normal = loc+scale*z)
    z_eta: np.ndarray    # [R, S-1]        R:355  (regression only; zeros otherwise)
    u: np.ndarray        # [R, S-1]        R:387  Python random.uniform
    u_swap: np.ndarray   # [rounds, R-1]   R:677  coordinator uniforms


@dataclass
class Traces:
    pos_w: np.ndarray          # [R, S, P]   R:240, R:408, R:417
    lik_prop: np.ndarray       # [R, S]      likeh_list[:,0]  R:391 / C:404   (row 0 = -100)
    lik_prop_t: np.ndarray     # [R, S]      likelihood_proposal as returned (tempered), R:360
    prior_prop: np.ndarray     # [R, S]      (not a reference output; recorded for parity tests)
    diff_prop: np.ndarray      # [R, S]
    rmse_train: np.ndarray     # [R, S]
    rmse_test: np.ndarray
    acc_train: np.ndarray
    acc_test: np.ndarray
    accept_list: np.ndarray    # [R, S]      R:380
    accepted: np.ndarray       # [R, S] bool (decision of step i stored at i+1)
    mh_prob: np.ndarray        # [R, S]
    swapped: np.ndarray        # [rounds, R-1] bool
    swap_src: np.ndarray       # [rounds, R]
    num_swap: int = 0
    total_swap_proposals: int = 0
    state_w: np.ndarray = None     # [R, S, P] chain state AFTER step i-1 incl. swaps (row i)
    state_eta: np.ndarray = None   # [R, S]
    state_lik: np.ndarray = None   # [R, S]  current (tempered) likelihood after step
    state_prior: np.ndarray = None
    state_tau: np.ndarray = None   # [R, S]  last PROPOSED tau^2 (R:356; Q11)
    final_w: np.ndarray = None
    final_eta: np.ndarray = None
    extra: dict = field(default_factory=dict)


class Replica:
    """One tempered MH chain: the body of ``ptReplica.run`` (R:223-444 / C:232-456)."""

    def __init__(self, cfg: PTConfig, train, test, temperature, w0):
        self.cfg, self.train, self.test = cfg, train, test
        self.temperature = float(temperature)
        self.adapttemp = float(temperature)                     # R:150
        self.net = Network(cfg.topology, cfg.learn_rate, cfg.task)
        self.w = np.array(w0, dtype=np.float64, copy=True)
        c = cfg
        I = c.topology[0]
        self.y_train, self.y_test = train[:, I], test[:, I]
        if c.task == REGRESSION:
            pred_train = self.net.evaluate_proposal(train, self.w)          # R:266
            self.eta = float(np.log(np.var(pred_train - self.y_train)))     # R:270
            self.tau_pro = float(np.exp(self.eta))                          # R:271
            self.prior_current = prior_regression(c.sigma_squared, c.nu_1, c.nu_2, self.w,
                                                  self.tau_pro, c.topology)  # R:280
            self.likelihood, _, self.rmsetrain = likelihood_regression(
                self.net, train, self.w, self.tau_pro, self.adapttemp)      # R:284
            _, _, self.rmsetest = likelihood_regression(self.net, test, self.w, self.tau_pro,
                                                        self.adapttemp)     # R:285
        else:
            self.eta = 0.0                                                  # C:263 junk variable
            self.tau_pro = 1.0
            self.prior_current = prior_classification(c.sigma_squared, c.nu_1, c.nu_2, self.w,
                                                      c.topology)           # C:281
            self.likelihood, self.pred_train, self.rmsetrain = likelihood_classification(
                self.net, train, self.w, self.adapttemp)                    # C:283
            _, self.pred_test, self.rmsetest = likelihood_classification(
                self.net, test, self.w, self.adapttemp)                     # C:284
        self.num_accepted = 0
        self.langevin_count = 0
        self.pt_samples = c.samples * c.pt_fraction                          # R:301 (a float)
        self.init_count = 0

    def _lik(self, data, w, tau):
        if self.cfg.task == REGRESSION:
            return likelihood_regression(self.net, data, w, tau, self.adapttemp)
        return likelihood_classification(self.net, data, w, self.adapttemp)

    def step(self, i, lx, z, z_eta, u, tr: Traces, r: int):
        c = self.cfg
        if i < self.pt_samples:                                              # R:317
            self.adapttemp = self.temperature
        if i == self.pt_samples and self.init_count == 0:                    # R:320-324 (Q11)
            self.adapttemp = 1
            self.likelihood, _, self.rmsetrain = self._lik(self.train, self.w, self.tau_pro)
            _, _, self.rmsetest = self._lik(self.test, self.w, self.tau_pro)
            self.init_count = 1
        w = self.w
        if c.use_langevin_gradients and lx < c.l_prob:                       # R:329
            w_gd = self.net.langevin_gradient(self.train, w.copy(), 1)       # R:330
            w_proposal = w_gd + c.step_w * z                                 # R:331
            w_prop_gd = self.net.langevin_gradient(self.train, w_proposal.copy(), 1)   # R:332
            wc_delta = w - w_prop_gd
            wp_delta = w_proposal - w_gd
            sigma_sq = c.step_w * c.step_w
            first = -0.5 * np.sum(wc_delta * wc_delta) / sigma_sq            # R:341
            second = -0.5 * np.sum(wp_delta * wp_delta) / sigma_sq
            diff_prop = float((first - second) / self.adapttemp)             # R:345-346 (Q4)
            self.langevin_count += 1
        else:
            diff_prop = 0.0
            w_proposal = w + c.step_w * z                                    # R:353
        if c.task == REGRESSION:
            eta_pro = self.eta + c.step_eta * z_eta                          # R:355
            self.tau_pro = math.exp(eta_pro)                                 # R:356
        else:
            eta_pro = self.eta
        lik_prop, pred_train, rmsetrain = self._lik(self.train, w_proposal, self.tau_pro)  # R:360
        _, pred_test, rmsetest = self._lik(self.test, w_proposal, self.tau_pro)            # R:362
        if c.task == REGRESSION:
            prior_prop = prior_regression(c.sigma_squared, c.nu_1, c.nu_2, w_proposal,
                                          self.tau_pro, c.topology)          # R:364
        else:
            prior_prop = prior_classification(c.sigma_squared, c.nu_1, c.nu_2, w_proposal,
                                              c.topology)                    # C:378
        diff_prior = prior_prop - self.prior_current
        diff_likelihood = lik_prop - self.likelihood
        try:
            mh_prob = min(1, math.exp(diff_likelihood + diff_prior + diff_prop))   # R:373
        except OverflowError:
            mh_prob = 1
        tr.accept_list[r, i + 1] = self.num_accepted                         # R:380 (count BEFORE)
        tr.lik_prop[r, i + 1] = lik_prop if c.task == REGRESSION else lik_prop * self.adapttemp
        tr.lik_prop_t[r, i + 1] = lik_prop
        tr.prior_prop[r, i + 1] = prior_prop
        tr.diff_prop[r, i + 1] = diff_prop
        tr.mh_prob[r, i + 1] = mh_prob
        if u < mh_prob:                                                      # R:395
            self.num_accepted += 1
            self.likelihood = lik_prop
            self.prior_current = prior_prop
            self.w = w_proposal
            self.eta = eta_pro
            tr.accepted[r, i + 1] = True
            if c.task == CLASSIFICATION:                                     # C:414-415 (Q13)
                tr.acc_train[r, i + 1] = accuracy(pred_train, self.y_train)
                tr.acc_test[r, i + 1] = accuracy(pred_test, self.y_test)
            tr.pos_w[r, i + 1] = w_proposal
            tr.rmse_train[r, i + 1] = rmsetrain
            tr.rmse_test[r, i + 1] = rmsetest
        else:                                                                # R:416-423
            tr.pos_w[r, i + 1] = tr.pos_w[r, i]
            tr.rmse_train[r, i + 1] = tr.rmse_train[r, i]
            tr.rmse_test[r, i + 1] = tr.rmse_test[r, i]
            tr.acc_train[r, i + 1] = tr.acc_train[r, i]
            tr.acc_test[r, i + 1] = tr.acc_test[r, i]

    def swap_field(self):
        if self.cfg.task == REGRESSION:
            return self.likelihood * self.temperature                        # R:430 (Q8)
        return self.likelihood                                               # C:439


def new_traces(R, S, P, rounds) -> Traces:
    z = lambda *s: np.zeros(s)                                               # noqa: E731
    tr = Traces(pos_w=np.ones((R, S, P)), lik_prop=z(R, S), lik_prop_t=z(R, S), prior_prop=z(R, S), diff_prop=z(R, S),
                rmse_train=z(R, S), rmse_test=z(R, S), acc_train=z(R, S), acc_test=z(R, S),
                accept_list=z(R, S), accepted=np.zeros((R, S), dtype=bool), mh_prob=z(R, S),
                swapped=np.zeros((rounds, max(R - 1, 0)), dtype=bool),
                swap_src=np.zeros((rounds, R), dtype=np.int64))
    tr.lik_prop[:, 0] = -100.0                                               # R:293
    tr.state_w = z(R, S, P)
    tr.state_eta = z(R, S)
    tr.state_lik = z(R, S)
    tr.state_prior = z(R, S)
    tr.state_tau = z(R, S)
    return tr


def run_pt(cfg: PTConfig, train, test, temperatures, w0, draws: Draws, n_steps=None) -> Traces:
    """Deterministic restatement of ``ParallelTempering.run_chains`` (R:694-771) + R replicas."""
    R, S, P = len(temperatures), cfg.samples, cfg.P
    rounds = cfg.total_rounds()
    tr = new_traces(R, S, P, rounds)
    reps = [Replica(cfg, train, test, temperatures[k], w0[k]) for k in range(R)]
    for k, rep in enumerate(reps):
        tr.state_w[k, 0], tr.state_eta[k, 0] = rep.w, rep.eta
        tr.state_lik[k, 0], tr.state_prior[k, 0] = rep.likelihood, rep.prior_current
        tr.state_tau[k, 0] = rep.tau_pro
    tr.extra["init_eta"] = np.array([rep.eta for rep in reps])
    tr.extra["init_lik"] = np.array([rep.likelihood for rep in reps])
    tr.extra["init_prior"] = np.array([rep.prior_current for rep in reps])
    tr.extra["init_rmse_train"] = np.array([rep.rmsetrain for rep in reps])
    tr.extra["init_rmse_test"] = np.array([rep.rmsetest for rep in reps])
    rnd = 0
    last = S - 1 if n_steps is None else min(S - 1, n_steps)
    for i in range(last):
        for k, rep in enumerate(reps):
            rep.step(i, draws.lx[k, i], draws.z[k, i], draws.z_eta[k, i], draws.u[k, i], tr, k)
        if cfg.swap_due(i) and R > 1:
            src, sw = swap_sweep([rep.swap_field() for rep in reps], draws.u_swap[rnd])
            new = [(reps[s].w, reps[s].eta) for s in src]
            for k, rep in enumerate(reps):                                   # R:435-437 (Q7):
                rep.w, rep.eta = new[k]                                      # only w and eta move
            tr.swapped[rnd], tr.swap_src[rnd] = sw, src
            tr.num_swap += sum(sw)
            tr.total_swap_proposals += R - 1
            rnd += 1
        for k, rep in enumerate(reps):
            tr.state_w[k, i + 1], tr.state_eta[k, i + 1] = rep.w, rep.eta
            tr.state_lik[k, i + 1], tr.state_prior[k, i + 1] = rep.likelihood, rep.prior_current
            tr.state_tau[k, i + 1] = rep.tau_pro
    if last == S - 1 and rnd < rounds and R > 1:
        # left-over coordinator round on the exit vectors [w, eta, likelihood, ...] (R:442; Q9):
        # the lhood field is the tempered likelihood itself, for both tasks.
        src, sw = swap_sweep([rep.likelihood for rep in reps], draws.u_swap[rnd])
        tr.swapped[rnd], tr.swap_src[rnd] = sw, src
        tr.num_swap += sum(sw)
        tr.total_swap_proposals += R - 1
    tr.final_w = np.stack([rep.w for rep in reps])
    tr.final_eta = np.array([rep.eta for rep in reps])
    tr.extra["langevin_count"] = np.array([rep.langevin_count for rep in reps])
    return tr


# --------------------------------------------------------------------------------------
# draw helpers
# --------------------------------------------------------------------------------------
def random_draws(cfg: PTConfig, R: int, seed: int, common_random_numbers: bool = True) -> Draws:
    """float32-representable draws.  ``common_random_numbers`` mirrors SURVEY Q10: after fork all
    replicas share the NumPy stream (same lx / z / z_eta), only ``u`` differs."""
    rs = np.random.RandomState(seed)
    S, P = cfg.samples, cfg.P
    f32 = lambda a: a.astype(np.float32).astype(np.float64)                  # noqa: E731
    uni = lambda *s: rs.randint(0, 1 << 24, size=s).astype(np.float64) / float(1 << 24)  # noqa: E731
    if common_random_numbers:
        lx = np.repeat(uni(1, S - 1), R, axis=0)
        z = np.repeat(f32(rs.standard_normal((1, S - 1, P))), R, axis=0)
        z_eta = np.repeat(f32(rs.standard_normal((1, S - 1))), R, axis=0)
    else:
        lx, z, z_eta = uni(R, S - 1), f32(rs.standard_normal((R, S - 1, P))), f32(rs.standard_normal((R, S - 1)))
    if cfg.task == CLASSIFICATION:
        z_eta = np.zeros_like(z_eta)
    return Draws(lx=lx, z=z, z_eta=z_eta, u=uni(R, S - 1),
                 u_swap=uni(max(cfg.total_rounds(), 1), max(R - 1, 1)))
