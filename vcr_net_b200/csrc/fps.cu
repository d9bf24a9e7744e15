// Farthest-point sampling (reference util/util.py:107-140 == util/fps.py:10-49).
// One CTA per cloud; coordinates and running min-distances live in shared memory; each of the
// npoint sequential steps is one fused pass (distance, running min, arg-max) + a block arg-max.
//
// Canonical arithmetic, identical to oracle/canon.c so indices are bit-exact:
//   barycentre = float(double-sum) / float(N)           (IEEE fp32 division)
//   d = (dx*dx + dy*dy) + dz*dz with separately rounded multiplies and adds (no FMA contraction)
//   arg-max returns the FIRST maximal index.
// Roofline: 12*N bytes read per cloud once; the loop is latency/ALU bound (npoint*N*~9 flops).
#include "common.cuh"

namespace {

constexpr int FPS_THREADS = 512;

__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// (value, index) arg-max with ties -> lower index
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

__device__ __forceinline__ int block_argmax(float v, int i, float* sv, int* si) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        argmax_combine(v, i, ov, oi);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                       // protect sv/si from the previous round's readers
    if (lane == 0) { sv[warp] = v; si[warp] = i; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        v = lane < nw ? sv[lane] : -INFINITY;
        i = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            argmax_combine(v, i, ov, oi);
        }
        if (lane == 0) si[32] = i;
    }
    __syncthreads();
    return si[32];
}

__global__ void __launch_bounds__(FPS_THREADS)
fps_kernel(const float* __restrict__ xyz, int N, int npoint, int32_t* __restrict__ out32,
           int64_t* __restrict__ out64) {
    extern __shared__ __align__(16) float sm[];
    float* X = sm; float* Y = X + N; float* Z = Y + N; float* Dm = Z + N;
    __shared__ float sv[32];
    __shared__ int si[33];
    __shared__ double sd[3][FPS_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* p = xyz + (size_t)b * 3 * N;
    double sx = 0, sy = 0, sz = 0;
    for (int n = tid; n < N; n += FPS_THREADS) {
        const float x = p[n], y = p[N + n], z = p[2 * N + n];
        X[n] = x; Y[n] = y; Z[n] = z; Dm[n] = 1e10f;
        sx += x; sy += y; sz += z;
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if ((tid & 31) == 0) { sd[0][tid >> 5] = sx; sd[1][tid >> 5] = sy; sd[2][tid >> 5] = sz; }
    __syncthreads();
    sx = sy = sz = 0;
    for (int w = 0; w < FPS_THREADS / 32; ++w) { sx += sd[0][w]; sy += sd[1][w]; sz += sd[2][w]; }
    const float bx = __fdiv_rn((float)sx, (float)N), by = __fdiv_rn((float)sy, (float)N),
                bz = __fdiv_rn((float)sz, (float)N);
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int n = tid; n < N; n += FPS_THREADS) {
        const float d = sqdist3(X[n], Y[n], Z[n], bx, by, bz);
        if (d > bv) { bv = d; bi = n; }
    }
    int far = block_argmax(bv, bi, sv, si);
    for (int s = 0; s < npoint; ++s) {
        if (tid == 0) {
            if (out32) out32[(size_t)b * npoint + s] = far;
            if (out64) out64[(size_t)b * npoint + s] = far;
        }
        const float cx = X[far], cy = Y[far], cz = Z[far];
        bv = -INFINITY; bi = 0x7fffffff;
        for (int n = tid; n < N; n += FPS_THREADS) {
            const float d = sqdist3(X[n], Y[n], Z[n], cx, cy, cz);
            float cur = Dm[n];
            if (d < cur) { cur = d; Dm[n] = d; }
            if (cur > bv) { bv = cur; bi = n; }
        }
        far = block_argmax(bv, bi, sv, si);
    }
}

}  // namespace

// xyz [B,3,N] -> idx [B,npoint] (int32 and/or int64).  N <= 14000 (16*N bytes of shared memory).
VCR_API int vcr_fps(const float* xyz, int B, int N, int npoint, int32_t* idx32, int64_t* idx64, cudaStream_t stream) {
    VCR_REQUIRE(xyz && (idx32 || idx64) && B > 0 && N > 0 && npoint > 0);
    const size_t smem = (size_t)4 * N * sizeof(float);
    if (smem > 224 * 1024) return VCR_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return VCR_ERR_LAUNCH;
    fps_kernel<<<B, FPS_THREADS, smem, stream>>>(xyz, N, npoint, idx32, idx64);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}
