"""CPU: the C-ABI library loads and exports every symbol include/ptfnn.h declares; the product
refuses to compute without a CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import __graft_entry__ as ge
from ptnn_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(capi.LIB_PATH):
        ge.build_cuda()
    return capi.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ptfnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptfnn_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.ptfnn_abi_version() == capi.ABI_VERSION
    info = capi.build_info()
    assert "sm_100a" in info and "reg[4,5,1]" in info and "cls[16,256,10]" in info


def test_config_struct_layout_matches_header(lib):
    c = capi.default_config()
    assert c.abi_version == capi.ABI_VERSION
    assert (c.step_w, c.step_eta, c.sigma_squared, c.pt_fraction) == (0.025, 0.2, 25.0, 0.6)   # R:258-275, R:301
    assert (c.l_prob, c.learn_rate, c.nu_1, c.nu_2) == (0.5, 0.1, 0.0, 0.0)
    assert c.swap_rule == capi.SWAP_RULE_AUTO and c.use_langevin_gradients == 1


def test_every_topology_has_its_four_kernels():
    topo = ge._topologies()
    assert len(topo) >= 8
    import subprocess
    out = subprocess.run(["cuobjdump", "-elf", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    names = out.stdout
    for kern in ("chain_kernel", "init_kernel", "op_forward_kernel", "op_sgd_kernel"):
        assert names.count(kern) >= len(topo), kern
    assert "sm_100a" in names


def test_no_cpu_fallback(lib):
    """Without a GPU the product path must fail loudly, not compute on the host."""
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.PtfnnError) as e:
        capi.op_prior(capi.TASK_REGRESSION, (4, 5, 1), np.zeros(31), tausq=1.0)
    assert e.value.code == capi.E_CUDA and "no CPU path" in str(e.value)
    from ptnn_b200.sampler import Sampler
    with pytest.raises(capi.PtfnnError) as e:
        Sampler(capi.TASK_REGRESSION, (4, 5, 1), [1.0, 2.0], 10, 5)
    assert e.value.code == capi.E_CUDA


def test_argument_validation_happens_before_cuda(lib):
    from ptnn_b200.sampler import Sampler
    with pytest.raises(capi.PtfnnError) as e:
        Sampler(capi.TASK_REGRESSION, (4, 5, 2), [1.0, 2.0], 10, 5)           # regression needs O == 1 (R:132)
    assert e.value.code == capi.E_UNSUPPORTED
    with pytest.raises(capi.PtfnnError) as e:
        Sampler(capi.TASK_REGRESSION, (4, 600, 1), [1.0, 2.0], 10, 5)         # hidden layers up to capi.MAX_HIDDEN units
    assert e.value.code == capi.E_UNSUPPORTED

    with pytest.raises(capi.PtfnnError) as e:
        Sampler(capi.TASK_REGRESSION, (4, 5, 1), [1.0, 2.0], 1, 5)            # samples < 2
    assert e.value.code == capi.E_INVALID
    for temps, kw in (([1.0, 0.0], {}), ([1.0, float("nan")], {}), ([1.0, -2.0], {}),          # the likelihood is divided by T (R:204)
                      ([1.0, 2.0], {"l_prob": 1.5}), ([1.0, 2.0], {"learn_rate": float("inf")}),
                      ([1.0, 2.0], {"window_plan": 3}), ([1.0, 2.0], {"swap_kind": 7}), ([1.0, 2.0], {"barrier_timeout_ms": -1})):
        with pytest.raises(capi.PtfnnError) as e:
            Sampler(capi.TASK_REGRESSION, (4, 5, 1), temps, 10, 5, **kw)
        assert e.value.code == capi.E_INVALID


def test_topology_is_compiled_on_demand(lib):
    """A topology that is not built into libptfnn.so is compiled from the same sources into its own shared
    library and registered (capi.ensure_topology); nvcc cross-compiles sm_100a without a GPU."""
    topo = (3, 7, 1)
    assert not capi.has_topology(capi.TASK_REGRESSION, topo) or "reg[3,7,1]" in capi.build_info()
    capi.ensure_topology(capi.TASK_REGRESSION, topo)
    assert capi.has_topology(capi.TASK_REGRESSION, topo)
    assert "reg[3,7,1]" in capi.build_info()
    capi.ensure_topology(capi.TASK_REGRESSION, topo)                          # idempotent
