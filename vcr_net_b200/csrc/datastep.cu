// On-device data step and evaluation metrics either side of the registration path.
//
//   vcr_make_pairs     util/data.py:247-309 (ModelNet40.__getitem__) minus the random draws: the host draws the per-item
//                      pose and permutations exactly as the reference does (they are data, drawn from numpy's
//                      RandomState(item)), the device gathers, applies the float64 rigid transform
//                      (Rotation.from_euler('zyx').apply + t, :289-291) and casts to fp32 -- the clouds never exist
//                      on the host.
//   vcr_crop_nearest   util/data.py:320-329 nearest_neighbor(): keep the int(N*reserve) points nearest to the LAST point,
//                      nearest first (float64 squared distances, ties -> lower index), by rank counting in shared memory.
//   vcr_eval_metrics   model/vcrnet_model.py:583-630: pose loss, cycle loss, mse/mae of the a->b and b->a transformed
//                      clouds, accumulated into a device-resident double[8] (the reference does 6+ .item() host
//                      syncs per batch).
// Roofline: HBM / latency bound helpers (a few bytes per point); they exist to keep the 8 GPUs fed without host work.
#include "common.cuh"

namespace {

// base [P, Nb, 3] fp32; idx_src / idx_tgt [P, N] int32 (indices into the item's base cloud, permutations composed on
// the host); pose [P, 12] fp64 = row-major R (9) then t (3).  src[p,:,n] = base[idx_src[n]], tgt[p,:,n] = R base[idx_tgt[n]] + t.
__global__ void make_pairs_kernel(const float* __restrict__ base, int Nb, const int* __restrict__ idx_src,
                                  const int* __restrict__ idx_tgt, const double* __restrict__ pose, int N,
                                  double* __restrict__ src64, double* __restrict__ tgt64,
                                  float* __restrict__ src, float* __restrict__ tgt) {
    const int p = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* bp = base + (size_t)p * Nb * 3;
    const double* ps = pose + (size_t)p * 12;
    const int is = idx_src[(size_t)p * N + n], it = idx_tgt[(size_t)p * N + n];
    const double sx = bp[is * 3 + 0], sy = bp[is * 3 + 1], sz = bp[is * 3 + 2];
    const double x = bp[it * 3 + 0], y = bp[it * 3 + 1], z = bp[it * 3 + 2];
    double o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        o[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(ps[i * 3 + 0], x), __dmul_rn(ps[i * 3 + 1], y)),
                                   __dmul_rn(ps[i * 3 + 2], z)), ps[9 + i]);
    const size_t o0 = (size_t)p * 3 * N + n;
    if (src64) { src64[o0] = sx; src64[o0 + N] = sy; src64[o0 + 2 * N] = sz; }
    if (tgt64) { tgt64[o0] = o[0]; tgt64[o0 + N] = o[1]; tgt64[o0 + 2 * N] = o[2]; }
    if (src) { src[o0] = (float)sx; src[o0 + N] = (float)sy; src[o0 + 2 * N] = (float)sz; }
    if (tgt) { tgt[o0] = (float)o[0]; tgt[o0 + N] = (float)o[1]; tgt[o0 + 2 * N] = (float)o[2]; }
}

// pc64 [P,3,N] fp64 -> out [P,3,keep] fp32: the keep points nearest to point N-1, nearest first
__global__ void crop_nearest_kernel(const double* __restrict__ pc64, int N, int keep, float* __restrict__ out) {
    extern __shared__ double d2[];
    const int p = blockIdx.x;
    const double* pc = pc64 + (size_t)p * 3 * N;
    const double lx = pc[N - 1], ly = pc[2 * N - 1], lz = pc[3 * N - 1];
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double dx = pc[i] - lx, dy = pc[N + i] - ly, dz = pc[2 * N + i] - lz;
        d2[i] = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));   // ((x-l)**2).sum(axis=1)
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double di = d2[i];
        int rank = 0;
        for (int j = 0; j < N; ++j) {
            const double dj = d2[j];
            rank += (dj < di) || (dj == di && j < i);
        }
        if (rank < keep) {
            float* o = out + (size_t)p * 3 * keep;
            o[rank] = (float)pc[i]; o[keep + rank] = (float)pc[N + i]; o[2 * keep + rank] = (float)pc[2 * N + i];
        }
    }
}

__device__ __forceinline__ void rigid(const float* R, const float* t, float x, float y, float z, float* o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float acc = R[i * 3 + 0] * x;
        acc = fmaf(R[i * 3 + 1], y, acc);
        acc = fmaf(R[i * 3 + 2], z, acc);
        o[i] = acc + t[i];
    }
}

// acc[0] pose loss, [1] cycle loss, [2] mse_ab, [3] mae_ab, [4] mse_ba, [5] mae_ba, [6] number of examples (each already
// multiplied by the batch size as the reference accumulates `.item() * batch_size`)
__global__ void eval_metrics_kernel(const float* __restrict__ src, const float* __restrict__ tgt, int N,
                                    const float* __restrict__ srcK, const float* __restrict__ corrK, int M,
                                    const float* __restrict__ R_gt, const float* __restrict__ t_gt,
                                    const float* __restrict__ R_ab, const float* __restrict__ t_ab,
                                    const float* __restrict__ R_ba, const float* __restrict__ t_ba, int B,
                                    double* __restrict__ acc) {
    __shared__ double red[4][8];
    double s_ab2 = 0.0, s_ab1 = 0.0, s_ba2 = 0.0, s_ba1 = 0.0;
    const int stride = gridDim.x * blockDim.x, g = blockIdx.x * blockDim.x + threadIdx.x;
    for (long long e = g; e < (long long)B * M; e += stride) {             // transformed_srcK vs src_corrK (:591,623-624)
        const int b = (int)(e / M), m = (int)(e - (long long)b * M);
        const float* s = srcK + (size_t)b * 3 * M;
        const float* c = corrK + (size_t)b * 3 * M;
        float o[3];
        rigid(R_gt + b * 9, t_gt + b * 3, s[m], s[M + m], s[2 * M + m], o);
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double d = (double)(o[i] - c[i * M + m]); s_ab2 += d * d; s_ab1 += fabs(d); }
    }
    for (long long e = g; e < (long long)B * N; e += stride) {             // transformed_target vs src (:589,626-627)
        const int b = (int)(e / N), n = (int)(e - (long long)b * N);
        const float* t = tgt + (size_t)b * 3 * N;
        const float* s = src + (size_t)b * 3 * N;
        float o[3];
        rigid(R_ba + b * 9, t_ba + b * 3, t[n], t[N + n], t[2 * N + n], o);
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double d = (double)(o[i] - s[i * N + n]); s_ba2 += d * d; s_ba1 += fabs(d); }
    }
    double v[4] = {s_ab2, s_ab1, s_ba2, s_ba1};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[q] = warp_sum(v[q]);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        const double denom = threadIdx.x < 2 ? 3.0 * M : 3.0 * N;          // mean over [B,3,*] times batch size
        atomicAdd(acc + 2 + threadIdx.x, s / denom);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {                              // pose / cycle losses (:598-617), tiny
        double lr = 0.0, lt = 0.0, cr = 0.0, ct = 0.0;
        for (int b = 0; b < B; ++b) {
            const float* Rp = R_ab + b * 9; const float* Rg = R_gt + b * 9; const float* Rb = R_ba + b * 9;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    float m1 = 0.f, m2 = 0.f;                                 // (R_pred^T R_gt)_ij and (R_ba R_ab)_ij
                    for (int k = 0; k < 3; ++k) { m1 = fmaf(Rp[k * 3 + i], Rg[k * 3 + j], m1); m2 = fmaf(Rb[i * 3 + k], Rp[k * 3 + j], m2); }
                    const double d1 = (double)m1 - (i == j), d2 = (double)m2 - (i == j);
                    lr += d1 * d1; cr += d2 * d2;
                }
            for (int i = 0; i < 3; ++i) {
                const double d = (double)t_ab[b * 3 + i] - (double)t_gt[b * 3 + i];
                lt += d * d;
                float m = 0.f;                                               // (R_ba^T t_ab + t_ba)_i
                for (int k = 0; k < 3; ++k) m = fmaf(Rb[k * 3 + i], t_ab[b * 3 + k], m);
                const double c = (double)m + (double)t_ba[b * 3 + i];
                ct += c * c;
            }
        }
        const double pose = lr / (9.0 * B) + lt / (3.0 * B);
        const double cyc = cr / (9.0 * B) + ct / (3.0 * B);
        atomicAdd(acc + 0, pose * B);
        atomicAdd(acc + 1, cyc * B);
        atomicAdd(acc + 6, (double)B);
    }
}

}  // namespace

VCR_API int vcr_make_pairs(const float* base, int P, int Nb, const int* idx_src, const int* idx_tgt, const double* pose,
                           int N, double* src64, double* tgt64, float* src, float* tgt, cudaStream_t stream) {
    VCR_REQUIRE(base && idx_src && idx_tgt && pose && P > 0 && Nb > 0 && N > 0 && P <= 65535 && (src64 || src) && (tgt64 || tgt));
    dim3 g(vcr_cdiv(N, 256), P);
    make_pairs_kernel<<<g, 256, 0, stream>>>(base, Nb, idx_src, idx_tgt, pose, N, src64, tgt64, src, tgt);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_crop_nearest(const double* pc64, int P, int N, int keep, float* out, cudaStream_t stream) {
    VCR_REQUIRE(pc64 && out && P > 0 && N > 0 && keep > 0 && keep <= N);
    if ((size_t)N * sizeof(double) > 200 * 1024) return VCR_ERR_UNSUPPORTED;
    const size_t smem = (size_t)N * sizeof(double);
    // set on every launch: the attribute is per device, and one process may drive several (nn.DataParallel)
    if (cudaFuncSetAttribute(crop_nearest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return VCR_ERR_LAUNCH;
    crop_nearest_kernel<<<P, 256, smem, stream>>>(pc64, N, keep, out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// acc: device double[8], zero-initialised once per evaluation run; src/tgt [B,3,N], srcK/corrK [B,3,M].
VCR_API int vcr_eval_metrics(const float* src, const float* tgt, int N, const float* srcK, const float* corrK, int M,
                             const float* R_gt, const float* t_gt, const float* R_ab, const float* t_ab,
                             const float* R_ba, const float* t_ba, int B, double* acc, cudaStream_t stream) {
    VCR_REQUIRE(src && tgt && srcK && corrK && R_gt && t_gt && R_ab && t_ab && R_ba && t_ba && acc && B > 0 && N > 0 && M > 0);
    const long long work = (long long)B * (N > M ? N : M);
    int blocks = vcr_cdiv(work, 256);
    if (blocks > 592) blocks = 592;
    eval_metrics_kernel<<<blocks, 256, 0, stream>>>(src, tgt, N, srcK, corrK, M, R_gt, t_gt, R_ab, t_ab, R_ba, t_ba, B, acc);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}
