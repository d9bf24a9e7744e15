// k-nearest-neighbour graph kernel (reference util/util.py:143-160).
//
// Canonical arithmetic (identical, op for op, to oracle/canon.c so indices are bit-exact):
//   xx_i   = fma chain over d = 0..D-1 of x_i[d]*x_i[d], start 0
//   dot_ij = fma chain over d = 0..D-1 of x_i[d]*x_j[d], start 0
//   pd_ij  = (-xx_j - (-2*dot_ij)) - xx_i           (reference op order, util/util.py:157-158)
//   neighbours(i) = ranks 1..k of pd_i* sorted descending, ties -> lower j
// The N x N matrix is never materialised: a CTA owns 32 query rows (one per lane, the query
// vector lives in registers) and streams candidate tiles through shared memory; its 8 warps
// each scan a different slice of every tile with a register-resident sorted top-(k+1) list
// (branch-free insertion), then the 8 lists are merged lexicographically by warp 0.
//
// Roofline: bytes = 4*D*N in + 4*k*N out per cloud (nothing else touches HBM; candidate tiles
// are re-read from L2 by the N/32 CTAs of a cloud), flops = 2*D*N^2 + ~3N^2 select ops:
// at N = 1024 this kernel is FP32-ALU / issue bound, not HBM bound.
#include "common.cuh"

namespace {

constexpr int QPB = 32;       // queries per CTA (one per lane)
constexpr int NWARP = 2;      // candidate slices (the final lexicographic merge is serial per query: keep it short)
constexpr int TJ = 128;       // candidates per smem tile
constexpr int CPW = TJ / NWARP;

template <int KS>
struct TopList {
    float v[KS];
    int i[KS];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int p = 0; p < KS; ++p) { v[p] = -INFINITY; i[p] = 0x7fffffff; }
    }
    // candidates arrive in increasing index order: equal values stay behind earlier ones
    __device__ __forceinline__ void push(float pd, int j) {
        if (pd > v[KS - 1]) {
#pragma unroll
            for (int p = KS - 1; p >= 1; --p) {
                if (v[p - 1] < pd) { v[p] = v[p - 1]; i[p] = i[p - 1]; }
                else if (v[p] < pd) { v[p] = pd; i[p] = j; }
            }
            if (v[0] < pd) { v[0] = pd; i[0] = j; }
        }
    }
    // arbitrary arrival order: order by (value desc, index asc)
    static __device__ __forceinline__ bool before(float av, int ai, float bv, int bi) {
        return av > bv || (av == bv && ai < bi);
    }
    __device__ __forceinline__ void push_lex(float pd, int j) {
        if (before(pd, j, v[KS - 1], i[KS - 1])) {
#pragma unroll
            for (int p = KS - 1; p >= 1; --p) {
                if (before(pd, j, v[p - 1], i[p - 1])) { v[p] = v[p - 1]; i[p] = i[p - 1]; }
                else if (before(pd, j, v[p], i[p])) { v[p] = pd; i[p] = j; }
            }
            if (before(pd, j, v[0], i[0])) { v[0] = pd; i[0] = j; }
        }
    }
};

// xx[b*N + i] = fma chain.  One thread per point.
template <bool TOKEN_MAJOR>
__global__ void knn_sqnorm_kernel(const float* __restrict__ x, int D, int N, float* __restrict__ xx) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* xb = x + (size_t)b * D * N;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) {
        const float v = TOKEN_MAJOR ? xb[(size_t)i * D + d] : xb[(size_t)d * N + i];
        acc = fmaf(v, v, acc);
    }
    xx[(size_t)b * N + i] = acc;
}

template <int D, int KS, bool TOKEN_MAJOR>
__global__ void __launch_bounds__(QPB * NWARP)
knn_topk_kernel(const float* __restrict__ x, const float* __restrict__ xx, int N, int k,
                int32_t* __restrict__ idx32, int64_t* __restrict__ idx64) {
    constexpr int DP = (D % 4 == 0) ? D + 4 : ((D + 3) / 4) * 4;   // padded row, 16-byte aligned
    constexpr int D4 = (D + 3) / 4;
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;                      // [TJ][DP]
    float* txx = smem + TJ * DP;             // [TJ]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int qi = blockIdx.x * QPB + lane;
    const bool qvalid = qi < N;
    const float* xb = x + (size_t)b * D * N;
    const float* xxb = xx + (size_t)b * N;

    float q[D4 * 4];
#pragma unroll
    for (int d = 0; d < D4 * 4; ++d) {
        float v = 0.f;
        if (qvalid && d < D) v = TOKEN_MAJOR ? xb[(size_t)qi * D + d] : xb[(size_t)d * N + qi];
        q[d] = v;
    }
    const float xxq = qvalid ? xxb[qi] : 0.f;

    TopList<KS> top;
    top.init();

    for (int j0 = 0; j0 < N; j0 += TJ) {
        __syncthreads();
        // cooperative tile load -> tile[j][d]
        if (TOKEN_MAJOR) {
            for (int e = threadIdx.x; e < TJ * D; e += blockDim.x) {
                const int j = e / D, d = e - j * D;
                tile[j * DP + d] = (j0 + j < N) ? xb[(size_t)(j0 + j) * D + d] : 0.f;
            }
        } else {
            for (int e = threadIdx.x; e < TJ * D; e += blockDim.x) {
                const int d = e / TJ, j = e - d * TJ;
                tile[j * DP + d] = (j0 + j < N) ? xb[(size_t)d * N + j0 + j] : 0.f;
            }
        }
        if constexpr (D % 4 != 0) {
            constexpr int PADN = D4 * 4 - D;
            for (int e = threadIdx.x; e < TJ * PADN; e += blockDim.x) {
                const int j = e / PADN, d = D + e % PADN;
                tile[j * DP + d] = 0.f;
            }
        }
        for (int j = threadIdx.x; j < TJ; j += blockDim.x) txx[j] = (j0 + j < N) ? xxb[j0 + j] : 0.f;
        __syncthreads();

#pragma unroll 1
        for (int jj = 0; jj < CPW; jj += 4) {
            const int jl = warp * CPW + jj;
            const float* c = tile + jl * DP;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int d4 = 0; d4 < D4; ++d4) {
                const float4 v0 = *reinterpret_cast<const float4*>(c + 0 * DP + d4 * 4);
                const float4 v1 = *reinterpret_cast<const float4*>(c + 1 * DP + d4 * 4);
                const float4 v2 = *reinterpret_cast<const float4*>(c + 2 * DP + d4 * 4);
                const float4 v3 = *reinterpret_cast<const float4*>(c + 3 * DP + d4 * 4);
                // chain order d = 0,1,2,... exactly as oracle/canon.c; padded dims multiply 0*0
                a0 = fmaf(q[d4 * 4 + 0], v0.x, a0); a1 = fmaf(q[d4 * 4 + 0], v1.x, a1);
                a2 = fmaf(q[d4 * 4 + 0], v2.x, a2); a3 = fmaf(q[d4 * 4 + 0], v3.x, a3);
                if (d4 * 4 + 1 < D) {
                    a0 = fmaf(q[d4 * 4 + 1], v0.y, a0); a1 = fmaf(q[d4 * 4 + 1], v1.y, a1);
                    a2 = fmaf(q[d4 * 4 + 1], v2.y, a2); a3 = fmaf(q[d4 * 4 + 1], v3.y, a3);
                }
                if (d4 * 4 + 2 < D) {
                    a0 = fmaf(q[d4 * 4 + 2], v0.z, a0); a1 = fmaf(q[d4 * 4 + 2], v1.z, a1);
                    a2 = fmaf(q[d4 * 4 + 2], v2.z, a2); a3 = fmaf(q[d4 * 4 + 2], v3.z, a3);
                }
                if (d4 * 4 + 3 < D) {
                    a0 = fmaf(q[d4 * 4 + 3], v0.w, a0); a1 = fmaf(q[d4 * 4 + 3], v1.w, a1);
                    a2 = fmaf(q[d4 * 4 + 3], v2.w, a2); a3 = fmaf(q[d4 * 4 + 3], v3.w, a3);
                }
            }
            const int jg = j0 + jl;
            float p0 = __fsub_rn(__fsub_rn(-txx[jl + 0], -2.f * a0), xxq);
            float p1 = __fsub_rn(__fsub_rn(-txx[jl + 1], -2.f * a1), xxq);
            float p2 = __fsub_rn(__fsub_rn(-txx[jl + 2], -2.f * a2), xxq);
            float p3 = __fsub_rn(__fsub_rn(-txx[jl + 3], -2.f * a3), xxq);
            if (jg + 0 < N) top.push(p0, jg + 0);
            if (jg + 1 < N) top.push(p1, jg + 1);
            if (jg + 2 < N) top.push(p2, jg + 2);
            if (jg + 3 < N) top.push(p3, jg + 3);
        }
    }

    // merge the NWARP partial lists of each query (smem reuse: tile is dead now)
    __syncthreads();
    float* mv = smem;                                   // [NWARP][QPB][KS]
    int* mi = reinterpret_cast<int*>(smem + NWARP * QPB * KS);
    if (warp > 0) {
#pragma unroll
        for (int p = 0; p < KS; ++p) {
            mv[(warp * QPB + lane) * KS + p] = top.v[p];
            mi[(warp * QPB + lane) * KS + p] = top.i[p];
        }
    }
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < NWARP; ++w) {
#pragma unroll 1
            for (int p = 0; p < KS; ++p) {
                const float v = mv[(w * QPB + lane) * KS + p];
                const int i = mi[(w * QPB + lane) * KS + p];
                if (i == 0x7fffffff) break;
                top.push_lex(v, i);
            }
        }
        if (qvalid) {
            const size_t o = ((size_t)b * N + qi) * k;
#pragma unroll
            for (int p = 1; p < KS; ++p) {
                if (p <= k) {
                    if (idx32) idx32[o + p - 1] = top.i[p];
                    if (idx64) idx64[o + p - 1] = (int64_t)top.i[p];
                }
            }
        }
    }
}

template <int D, int KS, bool TM>
int launch_knn(const float* x, const float* xx, int B, int N, int k, int32_t* i32, int64_t* i64,
               cudaStream_t st) {
    constexpr int DP = (D % 4 == 0) ? D + 4 : ((D + 3) / 4) * 4;
    size_t tile_bytes = (size_t)(TJ * DP + TJ) * sizeof(float);
    size_t merge_bytes = (size_t)NWARP * QPB * KS * 8;
    size_t smem = tile_bytes > merge_bytes ? tile_bytes : merge_bytes;
    auto kern = knn_topk_kernel<D, KS, TM>;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return VCR_ERR_LAUNCH;
    }
    dim3 grid(vcr_cdiv(N, QPB), B);
    kern<<<grid, QPB * NWARP, smem, st>>>(x, xx, N, k, i32, i64);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

template <int D, bool TM>
int dispatch_ks(const float* x, const float* xx, int B, int N, int k, int32_t* i32, int64_t* i64,
                cudaStream_t st) {
    if (k == 20) return launch_knn<D, 21, TM>(x, xx, B, N, k, i32, i64, st);
    if (k <= 1) return launch_knn<D, 2, TM>(x, xx, B, N, k, i32, i64, st);
    if (k <= 8) return launch_knn<D, 9, TM>(x, xx, B, N, k, i32, i64, st);
    if (k <= 31) return launch_knn<D, 32, TM>(x, xx, B, N, k, i32, i64, st);
    return VCR_ERR_UNSUPPORTED;
}


// Generic feature width (any D >= 1): same canonical arithmetic, the d-chain is walked in chunks of
// GD_DC dims staged through shared memory (query chunk + candidate chunk); every thread keeps the
// CPW running dot products of its warp's candidate slice in registers across chunks.
constexpr int GD_DC = 32;

template <int KS, bool TOKEN_MAJOR>
__global__ void __launch_bounds__(QPB * NWARP)
knn_generic_kernel(const float* __restrict__ x, const float* __restrict__ xx, int D, int N, int k,
                   int32_t* __restrict__ idx32, int64_t* __restrict__ idx64) {
    constexpr int DP = GD_DC + 4;
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;                      // [TJ][DP]
    float* qs = tile + TJ * DP;              // [QPB][GD_DC + 1]
    float* txx = qs + QPB * (GD_DC + 1);     // [TJ]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * QPB;
    const int qi = q0 + lane;
    const bool qvalid = qi < N;
    const float* xb = x + (size_t)b * D * N;
    const float* xxb = xx + (size_t)b * N;
    const float xxq = qvalid ? xxb[qi] : 0.f;
    TopList<KS> top;
    top.init();
    for (int j0 = 0; j0 < N; j0 += TJ) {
        float acc[CPW];
#pragma unroll
        for (int c = 0; c < CPW; ++c) acc[c] = 0.f;
        for (int d0 = 0; d0 < D; d0 += GD_DC) {
            const int dc = min(GD_DC, D - d0);
            __syncthreads();
            for (int e = threadIdx.x; e < TJ * GD_DC; e += blockDim.x) {
                int j, d;
                if (TOKEN_MAJOR) { j = e / GD_DC; d = e - j * GD_DC; } else { d = e / TJ; j = e - d * TJ; }
                float v = 0.f;
                if (d < dc && j0 + j < N)
                    v = TOKEN_MAJOR ? xb[(size_t)(j0 + j) * D + d0 + d] : xb[(size_t)(d0 + d) * N + j0 + j];
                tile[j * DP + d] = v;
            }
            for (int e = threadIdx.x; e < QPB * GD_DC; e += blockDim.x) {
                int q, d;
                if (TOKEN_MAJOR) { q = e / GD_DC; d = e - q * GD_DC; } else { d = e / QPB; q = e - d * QPB; }
                float v = 0.f;
                if (d < dc && q0 + q < N)
                    v = TOKEN_MAJOR ? xb[(size_t)(q0 + q) * D + d0 + d] : xb[(size_t)(d0 + d) * N + q0 + q];
                qs[q * (GD_DC + 1) + d] = v;
            }
            if (d0 == 0)
                for (int j = threadIdx.x; j < TJ; j += blockDim.x) txx[j] = (j0 + j < N) ? xxb[j0 + j] : 0.f;
            __syncthreads();
            const float* qrow = qs + lane * (GD_DC + 1);
            for (int d = 0; d < dc; ++d) {
                const float qv = qrow[d];
#pragma unroll
                for (int c = 0; c < CPW; ++c) acc[c] = fmaf(qv, tile[(warp * CPW + c) * DP + d], acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const int jl = warp * CPW + c, jg = j0 + jl;
            const float pd = __fsub_rn(__fsub_rn(-txx[jl], -2.f * acc[c]), xxq);
            if (jg < N) top.push(pd, jg);
        }
    }
    __syncthreads();
    float* mv = smem;
    int* mi = reinterpret_cast<int*>(smem + NWARP * QPB * KS);
    if (warp > 0) {
#pragma unroll
        for (int p = 0; p < KS; ++p) {
            mv[(warp * QPB + lane) * KS + p] = top.v[p];
            mi[(warp * QPB + lane) * KS + p] = top.i[p];
        }
    }
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < NWARP; ++w) {
#pragma unroll 1
            for (int p = 0; p < KS; ++p) {
                const float v = mv[(w * QPB + lane) * KS + p];
                const int i = mi[(w * QPB + lane) * KS + p];
                if (i == 0x7fffffff) break;
                top.push_lex(v, i);
            }
        }
        if (qvalid) {
            const size_t o = ((size_t)b * N + qi) * k;
#pragma unroll
            for (int p = 1; p < KS; ++p) {
                if (p <= k) {
                    if (idx32) idx32[o + p - 1] = top.i[p];
                    if (idx64) idx64[o + p - 1] = (int64_t)top.i[p];
                }
            }
        }
    }
}

template <int KS, bool TM>
int launch_knn_generic(const float* x, const float* xx, int B, int D, int N, int k, int32_t* i32, int64_t* i64,
                       cudaStream_t st) {
    size_t tile_bytes = (size_t)(TJ * (GD_DC + 4) + QPB * (GD_DC + 1) + TJ) * sizeof(float);
    size_t merge_bytes = (size_t)NWARP * QPB * KS * 8;
    size_t smem = tile_bytes > merge_bytes ? tile_bytes : merge_bytes;
    auto kern = knn_generic_kernel<KS, TM>;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return VCR_ERR_LAUNCH;
    }
    dim3 grid(vcr_cdiv(N, QPB), B);
    kern<<<grid, QPB * NWARP, smem, st>>>(x, xx, D, N, k, i32, i64);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

template <bool TM>
int dispatch_generic(const float* x, const float* xx, int B, int D, int N, int k, int32_t* i32, int64_t* i64,
                     cudaStream_t st) {
    if (k == 20) return launch_knn_generic<21, TM>(x, xx, B, D, N, k, i32, i64, st);
    if (k <= 8) return launch_knn_generic<9, TM>(x, xx, B, D, N, k, i32, i64, st);
    if (k <= 31) return launch_knn_generic<32, TM>(x, xx, B, D, N, k, i32, i64, st);
    return VCR_ERR_UNSUPPORTED;
}

}  // namespace

VCR_API size_t vcr_knn_workspace_bytes(int B, int N) { return (size_t)B * N * sizeof(float); }

// x: [B,D,N] (token_major=0, the reference layout) or [B,N,D] (token_major=1); idx: [B,N,k].
// Either idx32 or idx64 (or both) may be given.  Requires 1 <= k <= 31, N >= k+1; D = 3 and D = 64 take the
// register-resident fast path, any other D the chunked generic kernel.
VCR_API int vcr_knn_topk(const float* x, int B, int D, int N, int k, int token_major, int32_t* idx32,
                         int64_t* idx64, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    VCR_REQUIRE(x && (idx32 || idx64) && B > 0 && N > 0 && k >= 1);
    if (N < k + 1) return VCR_ERR_INVALID;
    if (!workspace || workspace_bytes < vcr_knn_workspace_bytes(B, N)) return VCR_ERR_WORKSPACE;
    float* xx = reinterpret_cast<float*>(workspace);
    dim3 g(vcr_cdiv(N, 256), B);
    if (token_major) knn_sqnorm_kernel<true><<<g, 256, 0, stream>>>(x, D, N, xx);
    else knn_sqnorm_kernel<false><<<g, 256, 0, stream>>>(x, D, N, xx);
    VCR_CHECK_LAUNCH();
    if (D == 3) {
        return token_major ? dispatch_ks<3, true>(x, xx, B, N, k, idx32, idx64, stream)
                           : dispatch_ks<3, false>(x, xx, B, N, k, idx32, idx64, stream);
    } else if (D == 64) {
        return token_major ? dispatch_ks<64, true>(x, xx, B, N, k, idx32, idx64, stream)
                           : dispatch_ks<64, false>(x, xx, B, N, k, idx32, idx64, stream);
    }
    return token_major ? dispatch_generic<true>(x, xx, B, D, N, k, idx32, idx64, stream)
                       : dispatch_generic<false>(x, xx, B, D, N, k, idx32, idx64, stream);
}
