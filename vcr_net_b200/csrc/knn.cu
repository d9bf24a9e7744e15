// k-nearest-neighbour graph kernel (reference util/util.py:143-160).
//
// Canonical arithmetic (identical, op for op, to oracle/canon.c so indices are bit-exact):
//   xx_i   = fma chain over d = 0..D-1 of x_i[d]*x_i[d], start 0
//   dot_ij = fma chain over d = 0..D-1 of x_i[d]*x_j[d], start 0
//   pd_ij  = (-xx_j - (-2*dot_ij)) - xx_i           (reference op order, util/util.py:157-158)
//   neighbours(i) = ranks 1..k of pd_i* sorted descending, ties -> lower j
// The N x N matrix is never materialised.  A CTA owns 32 query rows and walks the cloud in chunks of
// 512 candidates:
//   compute  an SGEMM-style 32 x 512 tile, 8 x 8 register block per thread, the d-chain of every
//            (i, j) pair kept in ONE accumulator in d order (bit-exact by construction), operands
//            staged through shared memory 16 dims at a time (cp.async when rows are 16-byte aligned);
//   select   the 32 x 512 distance tile goes through shared memory to a warp-per-query selection:
//            a lane holds 16 distances in registers, a threshold tau (first chunk: the (k+1)-th largest
//            of the 32 lane maxima, found by a shuffle bitonic sort; later chunks: the running
//            (k+1)-th best) filters the chunk down to ~30 survivors with ballot/popc compaction, the
//            survivors are sorted 32 at a time as 64-bit (ordered-float, ~index) keys by shuffle
//            bitonic networks and bitonic-merged into the running top-32 list (one entry per lane).
//            Nothing is dropped unless k+1 seen elements are strictly better, so ranks 0..k are exact.
//
// Roofline: bytes = 4*D*N in + 4*k*N out per cloud (candidate chunks are re-read from L2 by the N/32
// CTAs of a cloud), flops = 2*D*N^2 (+ ~N^2 select steps): FP32-FMA bound at D = 64, selection
// (issue) bound at D = 3 -- DESIGN.md section 4.
#include "common.cuh"

namespace {

constexpr int TQ = 32;        // queries per CTA
constexpr int TC = 512;       // candidates per chunk
constexpr int NT = 256;       // threads per CTA: 8 warps = 4 query groups x 2 candidate groups, 8 x 8 per thread
constexpr int RPL = TC / 32;  // distances per lane in the selection phase

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

// order-preserving float -> uint32 (-0.0 canonicalised to +0.0 first: the two compare equal)
__device__ __forceinline__ uint32_t okey(float v) {
    const uint32_t b = __float_as_uint(v + 0.f);
    return b ^ (uint32_t)(((int32_t)b >> 31) | 0x80000000);
}
__device__ __forceinline__ float okey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
// 64-bit key (ordered value, ~index) kept as two 32-bit registers; keys of real entries are distinct
struct Key { uint32_t hi, lo; };
__device__ __forceinline__ Key key_xchg(Key v, int j, bool take_max) {
    Key o;
    o.hi = __shfl_xor_sync(0xffffffffu, v.hi, j);
    o.lo = __shfl_xor_sync(0xffffffffu, v.lo, j);
    const bool gt = v.hi > o.hi || (v.hi == o.hi && v.lo > o.lo);
    return gt == take_max ? v : o;
}
__device__ __forceinline__ Key key_max(Key a, Key b) {
    return (a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo)) ? a : b;
}

// W independent 32-lane bitonic sorts (descending), one element per lane each
template <int W>
__device__ __forceinline__ void sort32_desc(Key (&v)[W], int lane) {
#pragma unroll
    for (int ks = 2; ks <= 32; ks <<= 1) {
#pragma unroll
        for (int j = ks >> 1; j > 0; j >>= 1) {
            const bool take_max = ((lane & j) == 0) == ((lane & ks) == 0);
#pragma unroll
            for (int w = 0; w < W; ++w) v[w] = key_xchg(v[w], j, take_max);
        }
    }
}
template <int W>
__device__ __forceinline__ void sort32_desc_f(float (&v)[W], int lane) {
#pragma unroll
    for (int ks = 2; ks <= 32; ks <<= 1) {
#pragma unroll
        for (int j = ks >> 1; j > 0; j >>= 1) {
            const bool take_max = ((lane & j) == 0) == ((lane & ks) == 0);
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const float o = __shfl_xor_sync(0xffffffffu, v[w], j);
                v[w] = take_max ? fmaxf(v[w], o) : fminf(v[w], o);
            }
        }
    }
}
// run, add: descending-sorted 32-lists -> run = top 32 of the union, descending
template <int W>
__device__ __forceinline__ void merge32_desc(Key (&run)[W], const Key (&add)[W], int lane) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
        Key o;
        o.hi = __shfl_sync(0xffffffffu, add[w].hi, 31 - lane);
        o.lo = __shfl_sync(0xffffffffu, add[w].lo, 31 - lane);
        run[w] = key_max(run[w], o);
    }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const bool take_max = (lane & j) == 0;
#pragma unroll
        for (int w = 0; w < W; ++w) run[w] = key_xchg(run[w], j, take_max);
    }
}

// Filter W queries' chunk rows by their thresholds into buf (capacity TC keys), then sort the survivors 32 at a
// time and merge them into the running lists.  Returns false, leaving `run` untouched, if buf would overflow
// (impossible for W == 1).
template <int W>
__device__ __forceinline__ bool filter_merge(const float* drows, const float (&tau)[W], Key (&run)[W], uint2* buf,
                                             int j0, int lane) {
    const uint32_t lt = (1u << lane) - 1u;
    int cnt[W], off[W];
    int base = 0;
    __syncwarp();
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const float* dr = drows + w * TC + lane;
        float v[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) v[r] = dr[r * 32];
        off[w] = base;
        int n = base;
        // branch-free: the 16 ballots are independent, only the running offset chains
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const bool p = v[r] >= tau[w];
            const uint32_t bal = __ballot_sync(0xffffffffu, p);
            const int pos = n + __popc(bal & lt);
            if (p && pos < TC) buf[pos] = make_uint2(~(uint32_t)(j0 + r * 32 + lane), okey(v[r]));
            n += __popc(bal);
        }
        cnt[w] = n - base;
        base = n;
    }
    __syncwarp();
    if (base > TC) return false;
    int rounds = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) rounds = max(rounds, (cnt[w] + 31) >> 5);
    for (int r = 0; r < rounds; ++r) {
        Key e[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int p = r * 32 + lane;
            const uint2 t = p < cnt[w] ? buf[off[w] + p] : make_uint2(0u, 0u);
            e[w].lo = t.x; e[w].hi = t.y;
        }
        sort32_desc<W>(e, lane);
        merge32_desc<W>(run, e, lane);
    }
    return true;
}

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

constexpr int QW = TQ / (NT / 32);   // queries per warp in the selection phase (4)
constexpr size_t SURV_BYTES = (size_t)(NT / 32) * TC * sizeof(uint2);   // per-warp survivor buffers (alias the operand stage)

template <int DCH>
struct Smem {
    static constexpr int DCP = DCH == 4 ? 12 : DCH + 4;   // row stride (floats): conflict-free LDS.128 across 8 lanes
    static constexpr size_t stage_bytes = (size_t)(TC + TQ) * DCP * sizeof(float);
    static constexpr size_t region0 = stage_bytes > SURV_BYTES ? stage_bytes : SURV_BYTES;
    static constexpr size_t total = region0 + (size_t)TC * sizeof(float) + (size_t)TQ * TC * sizeof(float);
};

template <int DCH, bool TOKEN_MAJOR>
__global__ void __launch_bounds__(NT, 2)
knn_select_kernel(const float* __restrict__ x, const float* __restrict__ xx, int D, int N, int k,
                  int32_t* __restrict__ idx32, int64_t* __restrict__ idx64) {
    using S = Smem<DCH>;
    constexpr int DCP = S::DCP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Cs = reinterpret_cast<float*>(smem_raw);                       // [TC][DCP]
    float* Qs = Cs + TC * DCP;                                            // [TQ][DCP]
    uint2* surv = reinterpret_cast<uint2*>(smem_raw);                     // [8][TC], aliases Cs/Qs during selection
    float* xxc = reinterpret_cast<float*>(smem_raw + S::region0);         // [TC]
    float* dist = xxc + TC;                                               // [TQ][TC]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qg = warp & 3, cg = warp >> 2;
    const int b = blockIdx.y, q0 = blockIdx.x * TQ;
    const float* xb = x + (size_t)b * D * N;
    const float* xxb = xx + (size_t)b * N;
    const bool vec = TOKEN_MAJOR && (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);

    Key run[QW];
#pragma unroll
    for (int w = 0; w < QW; ++w) run[w].hi = run[w].lo = 0u;

    for (int j0 = 0; j0 < N; j0 += TC) {
        float acc[8][8];
#pragma unroll
        for (int qi = 0; qi < 8; ++qi)
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) acc[qi][ci] = 0.f;

        for (int d0 = 0; d0 < D; d0 += DCH) {
            __syncthreads();                       // previous compute / selection no longer reads region 0
            if (vec) {
                constexpr int PPR = DCH / 4;       // 16-byte pieces per row
                constexpr int RPP = NT / PPR;      // rows covered per pass
                const int part = tid % PPR, r0 = tid / PPR;
                const bool dok = d0 + 4 * part < D;
                const float* src = xb + (size_t)(j0 + r0) * D + d0 + 4 * part;
                float* dst = Cs + r0 * DCP + 4 * part;
#pragma unroll
                for (int n = 0; n < TC / RPP; ++n) {
                    const bool ok = dok && j0 + r0 + n * RPP < N;
                    cp_async16(dst + n * RPP * DCP, ok ? src + (size_t)n * RPP * D : xb, ok);
                }
                if (r0 < TQ) {
                    const bool ok = dok && q0 + r0 < N;
                    cp_async16(Qs + r0 * DCP + 4 * part, ok ? xb + (size_t)(q0 + r0) * D + d0 + 4 * part : xb, ok);
                }
            } else {
                for (int e = tid; e < (TC + TQ) * DCH; e += NT) {
                    int row, d;
                    if (TOKEN_MAJOR) { row = e / DCH; d = e - row * DCH; }
                    else if (e < TC * DCH) { d = e / TC; row = e - d * TC; }
                    else { const int e2 = e - TC * DCH; d = e2 / TQ; row = TC + e2 - d * TQ; }
                    const int g = row < TC ? j0 + row : q0 + (row - TC);
                    float v = 0.f;
                    if (g < N && d0 + d < D)
                        v = TOKEN_MAJOR ? xb[(size_t)g * D + d0 + d] : xb[(size_t)(d0 + d) * N + g];
                    Cs[row * DCP + d] = v;
                }
            }
            if (d0 == 0)
                for (int c = tid; c < TC; c += NT) xxc[c] = (j0 + c < N) ? xxb[j0 + c] : 0.f;
            if (vec) cp_async_wait_all();
            __syncthreads();

            const float* qrow = Qs + (qg * 8) * DCP;
            const float* crow = Cs + (cg * 256 + lane) * DCP;
            if (j0 + cg * 256 >= N) continue;      // this warp's 256 candidates are all past N (e.g. N = 768): warp-uniform
#pragma unroll 1   // (a fully unrolled body makes ptxas rotate the 64 accumulators: +20% MOVs and spills at 128 regs)
            for (int d4 = 0; d4 < DCH / 4; ++d4) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 cv[4];
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci)
                        cv[ci] = *reinterpret_cast<const float4*>(crow + ((h * 4 + ci) * 32) * DCP + d4 * 4);
#pragma unroll
                    for (int qi = 0; qi < 8; ++qi) {
                        const float4 qv = *reinterpret_cast<const float4*>(qrow + qi * DCP + d4 * 4);
                        // chain order d = 0,1,2,... per accumulator, exactly as oracle/canon.c; padded dims add 0*0
#pragma unroll
                        for (int ci = 0; ci < 4; ++ci) acc[qi][h * 4 + ci] = fmaf(qv.x, cv[ci].x, acc[qi][h * 4 + ci]);
#pragma unroll
                        for (int ci = 0; ci < 4; ++ci) acc[qi][h * 4 + ci] = fmaf(qv.y, cv[ci].y, acc[qi][h * 4 + ci]);
#pragma unroll
                        for (int ci = 0; ci < 4; ++ci) acc[qi][h * 4 + ci] = fmaf(qv.z, cv[ci].z, acc[qi][h * 4 + ci]);
#pragma unroll
                        for (int ci = 0; ci < 4; ++ci) acc[qi][h * 4 + ci] = fmaf(qv.w, cv[ci].w, acc[qi][h * 4 + ci]);
                    }
                }
            }
        }

        // distances of this chunk -> dist[q][c]; candidates past N become -inf
        float xxq[8];
#pragma unroll
        for (int qi = 0; qi < 8; ++qi) {
            const int q = q0 + qg * 8 + qi;
            xxq[qi] = q < N ? xxb[q] : 0.f;
        }
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            const int c = cg * 256 + ci * 32 + lane;
            const float nxc = -xxc[c];
            const bool cvalid = j0 + c < N;
#pragma unroll
            for (int qi = 0; qi < 8; ++qi) {
                const float pd = __fsub_rn(__fsub_rn(nxc, -2.f * acc[qi][ci]), xxq[qi]);
                dist[(qg * 8 + qi) * TC + c] = cvalid ? pd : -INFINITY;
            }
        }
        __syncthreads();

        // ---- selection: warp w owns queries QW*w .. QW*w+QW-1 ------------------------------------------
        {
            uint2* buf = surv + (size_t)warp * TC;
            const float* drows = dist + (size_t)(warp * QW) * TC;
            float tau[QW];
            if (j0 == 0) {
                float m[QW];
#pragma unroll
                for (int w = 0; w < QW; ++w) {
                    const float* dr = drows + w * TC + lane;
                    float mm = dr[0];
#pragma unroll
                    for (int r = 1; r < RPL; ++r) mm = fmaxf(mm, dr[r * 32]);
                    m[w] = mm;
                }
                sort32_desc_f<QW>(m, lane);
#pragma unroll
                for (int w = 0; w < QW; ++w) tau[w] = __shfl_sync(0xffffffffu, m[w], k);
            } else {
#pragma unroll
                for (int w = 0; w < QW; ++w)
                    tau[w] = okey_inv(__shfl_sync(0xffffffffu, run[w].hi, k));
            }
            // common case: the survivors of all QW queries fit the warp's buffer and are sorted / merged together
            if (!filter_merge<QW>(drows, tau, run, buf, j0, lane)) {
                // heavy ties (or an adversarial candidate order): one query at a time, any survivor count fits
#pragma unroll
                for (int w = 0; w < QW; ++w) {
                    float t1[1] = {tau[w]};
                    Key r1[1] = {run[w]};
                    filter_merge<1>(drows + w * TC, t1, r1, buf, j0, lane);
                    run[w] = r1[0];
                }
            }
        }
        // the loop-top __syncthreads orders these smem reads before the next chunk's stores
    }

#pragma unroll
    for (int w = 0; w < QW; ++w) {
        const int q = q0 + warp * QW + w;
        if (q < N && lane >= 1 && lane <= k) {
            const int j = (int)(~run[w].lo);
            const size_t o = ((size_t)b * N + q) * k + lane - 1;
            if (idx32) idx32[o] = j;
            if (idx64) idx64[o] = (int64_t)j;
        }
    }
}

template <int DCH, bool TM>
int launch_knn(const float* x, const float* xx, int B, int D, int N, int k, int32_t* i32, int64_t* i64,
               cudaStream_t st) {
    auto kern = knn_select_kernel<DCH, TM>;
    const size_t smem = Smem<DCH>::total;
    static bool configured = false;     // idempotent attribute, benign race
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return VCR_ERR_LAUNCH;
        configured = true;
    }
    dim3 grid(vcr_cdiv(N, TQ), B);
    kern<<<grid, NT, smem, st>>>(x, xx, D, N, k, i32, i64);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

}  // namespace

VCR_API size_t vcr_knn_workspace_bytes(int B, int N) { return (size_t)B * N * sizeof(float); }

// x: [B,D,N] (token_major=0, the reference layout) or [B,N,D] (token_major=1); idx: [B,N,k].
// Either idx32 or idx64 (or both) may be given.  Requires 1 <= k <= 31, N >= k+1; any D >= 1 (D <= 4 is staged
// 4 dims at a time, wider features 16 dims at a time).
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
    if (k > 31) return VCR_ERR_UNSUPPORTED;
    if (D <= 4)
        return token_major ? launch_knn<4, true>(x, xx, B, D, N, k, idx32, idx64, stream)
                           : launch_knn<4, false>(x, xx, B, D, N, k, idx32, idx64, stream);
    return token_major ? launch_knn<16, true>(x, xx, B, D, N, k, idx32, idx64, stream)
                       : launch_knn<16, false>(x, xx, B, D, N, k, idx32, idx64, stream);
}
