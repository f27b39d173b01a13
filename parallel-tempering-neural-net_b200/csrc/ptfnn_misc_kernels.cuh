// Topology-independent kernels (swap sweep, prior, Philox dump, multi-GPU row install).
// Included by ptfnn_api.cu only (non-template __global__ definitions).
#pragma once
#include "ptfnn_device.cuh"

namespace ptfnn {

// ---- topology-independent kernels ----
__global__ void op_sweep_kernel(int n, const double *lhood, const float *u, int *src, uint8_t *swapped, int *ns_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *l = reinterpret_cast<double *>(smem_raw);
    int *s = reinterpret_cast<int *>(smem_raw + (size_t)n * 8);
    for (int k = threadIdx.x; k < n; k += blockDim.x) { l[k] = lhood[k]; s[k] = k; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ns = swap_sweep_serial(n, l, s, swapped, [&](int k) { return u[k]; });
        if (ns_out) *ns_out = ns;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) src[k] = s[k];
}

__global__ void op_prior_kernel(int task, int I, int H, int O, const float *w, double sigma_sq, double nu1,
                                double nu2, double tausq, double *out) {
    __shared__ double red[32];
    const int P = I * H + H * O + H + O;
    double s = 0.0;
    for (int j = threadIdx.x; j < P; j += blockDim.x) s += (double)w[j] * (double)w[j];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
        if (task == kTaskReg)
            *out = -((I * H + H + 2) / 2.0) * log(sigma_sq) - t / (2.0 * sigma_sq) - (1.0 + nu1) * log(tausq) - nu2 / tausq;
        else
            *out = -((I * H + H + O + H * O) / 2.0) * log(sigma_sq) - t / (2.0 * sigma_sq);
    }
}

// Dump the Philox draws of free-running mode (verification only).
__global__ void draws_kernel(uint64_t seed, int crn, int replica_offset, int R, int P, int i0, int n, float *lx,
                             float *z, float *z_eta, float *u) {
    const int r = blockIdx.y, k = blockIdx.x;   // replica, step offset
    const uint32_t gr = (uint32_t)(replica_offset + r);
    const uint32_t stream = crn ? kStreamCommon : gr;
    const size_t q = (size_t)r * n + k;
    if (threadIdx.x == 0) {
        const StepDraws sd = philox_step_scalars(seed, (uint32_t)(i0 + k), stream, gr);
        lx[q] = sd.lx; z_eta[q] = sd.z_eta; u[q] = sd.u;
    }
    for (int b = threadIdx.x; b < (P + 3) / 4; b += blockDim.x) {
        float z4[4];
        philox_step_normals4(seed, (uint32_t)(i0 + k), stream, (uint32_t)b, z4);
        for (int c = 0; c < 4; ++c)
            if (4 * b + c < P) z[q * P + 4 * b + c] = z4[c];
    }
}

// Network.ForwardPass on ONE row with runtime dimensions (R:51-55): hid = sigmoid(x.W1 - B1),
// out = sigmoid(hid.W2 - B2).  One block; hidden units strided over threads.
__global__ void op_forward_row_kernel(int I, int H, int O, const float *x, const float *w, float *hid, float *out) {
    extern __shared__ float s_hid[];
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float z = -w[I * H + H * O + h];
        for (int i = 0; i < I; ++i) z = fmaf(x[i], w[i * H + h], z);
        s_hid[h] = sigmoid_precise(z);
        hid[h] = s_hid[h];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        float z = -w[I * H + H * O + H + o];
        for (int h = 0; h < H; ++h) z = fmaf(s_hid[h], w[I * H + h * O + o], z);
        out[o] = sigmoid_precise(z);
    }
}

// multi-GPU: install the rows selected by the sweep (R:435-437)
__global__ void swap_apply_kernel(int R, int P, int replica_offset, const int *src, const float *rows_local,
                                  const float *rows_in, float *w, double *eta, int *gd_valid) {
    const int r = blockIdx.x;
    const int s = src[replica_offset + r];
    if (s == replica_offset + r) return;
    const bool local = s >= replica_offset && s < replica_offset + R;
    const float *row = local ? rows_local + (size_t)(s - replica_offset) * (P + 1) : rows_in + (size_t)r * (P + 1);
    for (int j = threadIdx.x; j < P; j += blockDim.x) w[(size_t)r * P + j] = row[j];
    if (threadIdx.x == 0) { eta[r] = (double)row[P]; gd_valid[r] = 0; }
}

}  // namespace ptfnn
