// ICP refinement (--iter=0 path: reference model/icp_model.py:26-108 ICP.forward, driven by vcrnetIcpNet,
// model/vcrnet_model.py:46-62).
//
//   vcr_icp_nearest   nearest_neighbor (:52-75): for every source point the destination point maximising
//                     pd = (-xx - (-2 s.d)) - yy (reference op order; dot / norms are fma chains, ties -> lower index),
//                     gathers it into corr [B,3,Ns] and accumulates sum(pd_best) for the mean error.
//   vcr_icp_advance   the rest of one loop iteration (:36-40): src <- R src + t, then the convergence test
//                     |prev_error - mean_error| < tolerance -- evaluated ON THE DEVICE: a `done` flag in a small state
//                     record freezes src for all later iterations, so the loop needs no host synchronisation (the
//                     reference syncs every iteration through the Python `if`) and still stops where the reference does.
// The rigid fit between them is vcr_svd_head (best_fit_transform == SVDHead arithmetic, :77-108).
// Roofline: nearest is FP32-ALU bound (3 fma + 3 sub + compare per pair, 4*3*(Ns+Nt) bytes per cloud pair).
#include "common.cuh"

namespace {

struct IcpState {
    int done;
    int iters;
    double prev_error;
};

constexpr int NN_T = 128;      // threads = source points per CTA
constexpr int NN_TILE = 1024;  // destination points per smem tile

__global__ void __launch_bounds__(NN_T)
icp_nearest_kernel(const float* __restrict__ src, const float* __restrict__ dst, int Ns, int Nt,
                   float* __restrict__ corr, int* __restrict__ nn_idx, double* __restrict__ err_sum) {
    __shared__ float4 tile[NN_TILE];
    __shared__ double red[NN_T / 32];
    const int b = blockIdx.y;
    const int i = blockIdx.x * NN_T + threadIdx.x;
    const float* sb = src + (size_t)b * 3 * Ns;
    const float* db = dst + (size_t)b * 3 * Nt;
    const bool valid = i < Ns;
    const float sx = valid ? sb[i] : 0.f, sy = valid ? sb[Ns + i] : 0.f, sz = valid ? sb[2 * Ns + i] : 0.f;
    const float nxx = -fmaf(sz, sz, fmaf(sy, sy, fmaf(sx, sx, 0.f)));
    float best = -INFINITY;
    int bj = 0;
    for (int j0 = 0; j0 < Nt; j0 += NN_TILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < NN_TILE; j += NN_T) {
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + j < Nt) {
                d.x = db[j0 + j]; d.y = db[Nt + j0 + j]; d.z = db[2 * Nt + j0 + j];
                d.w = fmaf(d.z, d.z, fmaf(d.y, d.y, fmaf(d.x, d.x, 0.f)));
            }
            tile[j] = d;
        }
        __syncthreads();
        const int n = min(NN_TILE, Nt - j0);
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float4 d = tile[j];
            const float dot = fmaf(sz, d.z, fmaf(sy, d.y, fmaf(sx, d.x, 0.f)));
            const float pd = __fsub_rn(__fsub_rn(nxx, -2.f * dot), d.w);
            if (pd > best) { best = pd; bj = j0 + j; }
        }
    }
    if (valid) {
        float* cb = corr + (size_t)b * 3 * Ns;
        cb[i] = db[bj]; cb[Ns + i] = db[Nt + bj]; cb[2 * Ns + i] = db[2 * Nt + bj];
        if (nn_idx) nn_idx[(size_t)b * Ns + i] = bj;
    }
    double v = valid ? (double)best : 0.0;
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < NN_T / 32; ++w) s += red[w];
        atomicAdd(err_sum, s);
    }
}

__global__ void icp_apply_kernel(float* __restrict__ src, const float* __restrict__ R, const float* __restrict__ t,
                                 int Ns, const IcpState* __restrict__ state) {
    if (state->done) return;
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ns) return;
    float* sb = src + (size_t)b * 3 * Ns;
    const float* r = R + b * 9;
    const float x = sb[i], y = sb[Ns + i], z = sb[2 * Ns + i];
    // same accumulation order as vcr_rigid_apply (matmul row . column, then + t)
    sb[i] = fmaf(r[2], z, fmaf(r[1], y, r[0] * x)) + t[b * 3 + 0];
    sb[Ns + i] = fmaf(r[5], z, fmaf(r[4], y, r[3] * x)) + t[b * 3 + 1];
    sb[2 * Ns + i] = fmaf(r[8], z, fmaf(r[7], y, r[6] * x)) + t[b * 3 + 2];
}

__global__ void icp_check_kernel(IcpState* state, const double* err_sum, double count, float tolerance) {
    if (state->done) return;
    const double mean = *err_sum / count;
    state->iters += 1;
    if (fabs(state->prev_error - mean) < (double)tolerance) state->done = 1;
    state->prev_error = mean;
}

}  // namespace

VCR_API size_t vcr_icp_state_bytes(void) { return sizeof(IcpState); }

// src [B,3,Ns], dst [B,3,Nt] -> corr [B,3,Ns] (nearest destination point), optional nn_idx [B,Ns];
// *err_sum (double, caller-zeroed) += sum over all source points of the best pd.
VCR_API int vcr_icp_nearest(const float* src, const float* dst, int B, int Ns, int Nt, float* corr, int* nn_idx,
                            double* err_sum, cudaStream_t stream) {
    VCR_REQUIRE(src && dst && corr && err_sum && B > 0 && Ns > 0 && Nt > 0 && B <= 65535);
    dim3 grid(vcr_cdiv(Ns, NN_T), B);
    icp_nearest_kernel<<<grid, NN_T, 0, stream>>>(src, dst, Ns, Nt, corr, nn_idx, err_sum);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// One loop tail: if (!state.done) { src <- R src + t; mean = *err_sum / (B*Ns); done = |prev - mean| < tol; prev = mean }.
// state: vcr_icp_state_bytes() bytes, zero-initialised before the first iteration.
VCR_API int vcr_icp_advance(float* src, const float* R, const float* t, int B, int Ns, const double* err_sum,
                            float tolerance, void* state, cudaStream_t stream) {
    VCR_REQUIRE(src && R && t && err_sum && state && B > 0 && Ns > 0 && B <= 65535);
    dim3 grid(vcr_cdiv(Ns, 256), B);
    icp_apply_kernel<<<grid, 256, 0, stream>>>(src, R, t, Ns, reinterpret_cast<const IcpState*>(state));
    VCR_CHECK_LAUNCH();
    icp_check_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<IcpState*>(state), err_sum, (double)B * Ns, tolerance);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}
