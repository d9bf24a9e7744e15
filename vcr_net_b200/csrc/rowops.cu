// Row-wise / elementwise kernels of the path (HBM-bound by construction): layout changes at the
// module boundary, the 3->64 pointwise conv, the reference's custom LayerNorm, row softmax (with the
// partial-overlap key mask), column statistics of attention probabilities, and the soft
// virtual-correspondence reduction.  One warp per row, float4 accesses, no shared memory.
#include "tc_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// [B,C,N] <-> [B,N,C] through a 32x32 smem tile (both sides coalesced)
// ---------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc,
                                 int ld_in, int ld_out, long long s_in, long long s_out) {
    // in: [R, Cc] (row stride ld_in)  ->  out: [Cc, R] (row stride ld_out), per batch blockIdx.z
    __shared__ float t[32][33];
    const float* ib = in + blockIdx.z * s_in;
    float* ob = out + blockIdx.z * s_out;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < Cc) t[i][threadIdx.x] = ib[(size_t)r * ld_in + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < Cc) ob[(size_t)c * ld_out + r] = t[threadIdx.x][i];
    }
}

// ---------------------------------------------------------------------------------------------
// conv1_lpd: [B,3,N] channel-major xyz -> token-major [B*N, Cout], LeakyReLU
// (model/lpdnet_model.py:111).  Accumulation order c = 0,1,2 then + bias like a GEMM with bias.
// ---------------------------------------------------------------------------------------------
__global__ void conv3_kernel(const float* __restrict__ xyz, const float* __restrict__ w /*[Cout,3]*/,
                             const float* __restrict__ bias, int N, int Cout, float slope,
                             float* __restrict__ out, int ldo) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.y + threadIdx.y;
    if (n >= N) return;
    const float* p = xyz + (size_t)b * 3 * N;
    const float x = p[n], y = p[N + n], z = p[2 * N + n];
    float* o = out + ((size_t)b * N + n) * ldo;
    for (int c = threadIdx.x; c < Cout; c += 32) {
        float acc = w[c * 3 + 0] * x;
        acc = fmaf(w[c * 3 + 1], y, acc);
        acc = fmaf(w[c * 3 + 2], z, acc);
        o[c] = leaky(acc + bias[c], slope);
    }
}

// ---------------------------------------------------------------------------------------------
// conv1_lpd + conv2_lpd fused (model/lpdnet_model.py:111-112): xyz [B,3,N] -> h1 = act(W1 x + b1) [B*N,64] (optional store),
// h2 = act(W2 h1 + b2) [B*N,64] fp32 and, optionally, h2 in "h3" operand format for the kNN prefilter / the DG1 GEMM.
// A warp owns 32 points (lane = point), W2 is read from shared memory as broadcast float4.  Arithmetic is the two separate
// kernels' (conv3_kernel: w0*x, fma, fma, + b; sgemm_kernel: one fma chain over k = 0..63 starting at 0, then + b), so h1 / h2
// are bit-identical to the unfused path.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
lpd_point_mlp_kernel(const float* __restrict__ xyz, const float* __restrict__ w1, const float* __restrict__ b1,
                     const float* __restrict__ w2, const float* __restrict__ b2, int N, float slope,
                     float* __restrict__ h1o, float* __restrict__ h2o, __half* __restrict__ op, int ldop, long long op_plane) {
    __shared__ __align__(16) float sw2[64 * 64];
    __shared__ float sw1[64 * 3], sb1[64], sb2[64];
    for (int e = threadIdx.x; e < 64 * 64; e += 128) sw2[e] = w2[e];
    for (int e = threadIdx.x; e < 64 * 3; e += 128) sw1[e] = w1[e];
    if (threadIdx.x < 64) { sb1[threadIdx.x] = b1[threadIdx.x]; sb2[threadIdx.x] = b2[threadIdx.x]; }
    __syncthreads();
    const int b = blockIdx.y;
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    const float* p = xyz + (size_t)b * 3 * N;
    const float x = p[n], y = p[N + n], z = p[2 * N + n];
    float h1[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) {
        float acc = sw1[c * 3 + 0] * x;
        acc = fmaf(sw1[c * 3 + 1], y, acc);
        acc = fmaf(sw1[c * 3 + 2], z, acc);
        h1[c] = leaky(acc + sb1[c], slope);
    }
    const size_t row = (size_t)b * N + n;
    if (h1o != nullptr) {
#pragma unroll
        for (int c = 0; c < 64; c += 4)
            *reinterpret_cast<float4*>(h1o + row * 64 + c) = make_float4(h1[c], h1[c + 1], h1[c + 2], h1[c + 3]);
    }
#pragma unroll 1
    for (int o0 = 0; o0 < 64; o0 += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k4 = 0; k4 < 64; k4 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 w = *reinterpret_cast<const float4*>(sw2 + (o0 + u) * 64 + k4);
                acc[u] = fmaf(h1[k4 + 0], w.x, acc[u]); acc[u] = fmaf(h1[k4 + 1], w.y, acc[u]);
                acc[u] = fmaf(h1[k4 + 2], w.z, acc[u]); acc[u] = fmaf(h1[k4 + 3], w.w, acc[u]);
            }
        }
        float4 r;
        r.x = leaky(acc[0] + sb2[o0 + 0], slope); r.y = leaky(acc[1] + sb2[o0 + 1], slope);
        r.z = leaky(acc[2] + sb2[o0 + 2], slope); r.w = leaky(acc[3] + sb2[o0 + 3], slope);
        *reinterpret_cast<float4*>(h2o + row * 64 + o0) = r;
        if (op != nullptr) {
            __half* orow = op + row * ldop + o0;
            *reinterpret_cast<uint2*>(orow) = make_uint2(tc::pack_h2(r.x, r.y, 0), tc::pack_h2(r.z, r.w, 0));
            *reinterpret_cast<uint2*>(orow + op_plane) =
                make_uint2(tc::pack_h2(tc::lo_part(r.x, 0), tc::lo_part(r.y, 0), 0),
                           tc::pack_h2(tc::lo_part(r.z, 0), tc::lo_part(r.w, 0), 0));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm of model/transformer.py:141-144: a*(x-mean)/(std_unbiased+eps)+b.  D % 128 == 0, D <= 1024.
// ---------------------------------------------------------------------------------------------
// EXTRA: also write the row in operand format (fp16 hi / lo * 2^11 planes) and its squared norm -- the two things the
// VCP head derives from the Transformer output (to_operand + sqnorm_rows), with the summation order of sqnorm_rows_kernel.
// n / den with the reciprocal taken once per row: q = n * r, one Newton step on the residual (the sequence the IEEE division
// routine runs after its range checks): correctly rounded for normal-range operands, 3 instructions instead of ~10 per element
__device__ __forceinline__ float div_by(float n, float den, float rcp) {
    const float q = n * rcp;
    return fmaf(fmaf(-q, den, n), rcp, q);
}

template <int VPL, bool EXTRA>  // float4 per lane
__global__ void layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ a,
                                 const float* __restrict__ b, float eps, int M, int D,
                                 const float* __restrict__ res, int ldr, float* __restrict__ out, int ldo,
                                 __half* __restrict__ op, int ldop, long long plane, float* __restrict__ sq) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        v[i] = xr[lane + i * 32];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float stdv = sqrtf(warp_sum(q) / (float)(D - 1));
    const float den = stdv + eps;
    const float rcp = 1.f / den;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float4* o = reinterpret_cast<float4*>(out + (size_t)row * ldo);
    const float4* rr = res ? reinterpret_cast<const float4*>(res + (size_t)row * ldr) : nullptr;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float4 aa = a4[lane + i * 32], bb = b4[lane + i * 32];
        float4 r;
        r.x = div_by(aa.x * v[i].x, den, rcp) + bb.x;
        r.y = div_by(aa.y * v[i].y, den, rcp) + bb.y;
        r.z = div_by(aa.z * v[i].z, den, rcp) + bb.z;
        r.w = div_by(aa.w * v[i].w, den, rcp) + bb.w;
        if (rr) {
            const float4 e = rr[lane + i * 32];
            r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
        }
        if (!EXTRA || out != nullptr) o[lane + i * 32] = r;      // EXTRA: the fp32 copy is optional (the head reads op / sq)
        if (EXTRA) v[i] = r;
    }
    if (EXTRA) {
        float n2 = 0.f;
        __half* oh = op + (size_t)row * ldop;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 r = v[i];
            n2 += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
            const int c = (lane + i * 32) * 4;
            *reinterpret_cast<uint2*>(oh + c) = make_uint2(tc::pack_h2(r.x, r.y, 0), tc::pack_h2(r.z, r.w, 0));
            *reinterpret_cast<uint2*>(oh + plane + c) =
                make_uint2(tc::pack_h2(tc::lo_part(r.x, 0), tc::lo_part(r.y, 0), 0),
                           tc::pack_h2(tc::lo_part(r.z, 0), tc::lo_part(r.w, 0), 0));
        }
        n2 = warp_sum(n2);
        if (lane == 0) sq[row] = n2;
    }
}

// ---------------------------------------------------------------------------------------------
// In-place row softmax over S[rows, n] (row stride ld).  keep (optional): uint8 [batch, n], the row's
// batch is row / rows_per_batch; keys with keep == 0 are filled with -1e9 before the softmax exactly
// as model/transformer.py:51-52 does.  One warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(float* __restrict__ S, int ld, long long rows, int n,
                                    const uint8_t* __restrict__ keep, long long rows_per_batch) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* r = S + row * ld;
    const uint8_t* kp = keep ? keep + (row / rows_per_batch) * n : nullptr;
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) {
        float v = r[j];
        if (kp && !kp[j]) v = -1e9f;
        m = fmaxf(m, v);
    }
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < n; j += 32) {
        float v = r[j];
        if (kp && !kp[j]) v = -1e9f;
        const float e = expf(v - m);
        r[j] = e;
        s += e;
    }
    s = warp_sum(s);
    for (int j = lane; j < n; j += 32) r[j] = r[j] / s;
}

// colsum[b, j] = sum over the rows of batch b of P[row, j]   (model/transformer.py:39 and
// model/vcrnet_model.py:222).  Deterministic: a thread owns a column and walks its rows in order,
// rows are split into `gridDim.y` slabs per batch whose partials are added in slab order by pass 2.
__global__ void colsum_partial_kernel(const float* __restrict__ P, int ld, long long rows_per_batch, int n,
                                      int slabs, float* __restrict__ part /*[B, slabs, n]*/) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int slab = blockIdx.y, b = blockIdx.z;
    if (j >= n) return;
    const long long per = (rows_per_batch + slabs - 1) / slabs;
    const long long r0 = slab * per, r1 = min(rows_per_batch, r0 + per);
    const float* p = P + ((long long)b * rows_per_batch) * ld + j;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    long long r = r0;
    for (; r + 3 < r1; r += 4) {
        a0 += p[(r + 0) * ld]; a1 += p[(r + 1) * ld]; a2 += p[(r + 2) * ld]; a3 += p[(r + 3) * ld];
    }
    for (; r < r1; ++r) a0 += p[r * ld];
    part[((size_t)b * slabs + slab) * n + j] = (a0 + a1) + (a2 + a3);
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int slabs, int n, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (j >= n) return;
    float s = 0.f;
    for (int t = 0; t < slabs; ++t) s += part[((size_t)b * slabs + t) * n + j];
    out[(size_t)b * n + j] = s;
}


// ---------------------------------------------------------------------------------------------
// Operand-format producers for the tensor-core GEMMs (gemm_tc.cu): LayerNorm and softmax write the
// 16-bit hi/lo planes directly, so no separate conversion pass sits between them and the next GEMM.
// ---------------------------------------------------------------------------------------------
template <int VPL, int PLANES, int BF16>
__global__ void layernorm_operand_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ a,
                                         const float* __restrict__ b, float eps, int M, int D,
                                         __half* __restrict__ out, int ldo, long long plane) {
    constexpr int planes = PLANES, bf16 = BF16;        // compile-time: no per-element format branches
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        v[i] = xr[lane + i * 32];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float den = sqrtf(warp_sum(q) / (float)(D - 1)) + eps;
    const float rcp = 1.f / den;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    __half* o = out + (size_t)row * ldo;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float4 aa = a4[lane + i * 32], bb = b4[lane + i * 32];
        float4 r;
        r.x = div_by(aa.x * v[i].x, den, rcp) + bb.x; r.y = div_by(aa.y * v[i].y, den, rcp) + bb.y;
        r.z = div_by(aa.z * v[i].z, den, rcp) + bb.z; r.w = div_by(aa.w * v[i].w, den, rcp) + bb.w;
        const int c = (lane + i * 32) * 4;
        *reinterpret_cast<uint2*>(o + c) = make_uint2(tc::pack_h2(r.x, r.y, bf16), tc::pack_h2(r.z, r.w, bf16));
        if (planes == 2)
            *reinterpret_cast<uint2*>(o + plane + c) =
                make_uint2(tc::pack_h2(tc::lo_part(r.x, bf16), tc::lo_part(r.y, bf16), bf16),
                           tc::pack_h2(tc::lo_part(r.z, bf16), tc::lo_part(r.w, bf16), bf16));
    }
}

// row statistics of the (optionally key-masked) softmax: m_i = max_j s_ij, z_i = sum_j exp(s_ij - m_i)
__global__ void row_lse_kernel(const float* __restrict__ S, int ld, long long rows, int n,
                               const uint8_t* __restrict__ keep, long long rows_per_batch,
                               float* __restrict__ rmax, float* __restrict__ rsum) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* r = S + row * ld;
    const uint8_t* kp = keep ? keep + (row / rows_per_batch) * n : nullptr;
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) {
        float v = r[j];
        if (kp && !kp[j]) v = -1e9f;
        m = fmaxf(m, v);
    }
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < n; j += 32) {
        float v = r[j];
        if (kp && !kp[j]) v = -1e9f;
        s += expf(v - m);
    }
    s = warp_sum(s);
    if (lane == 0) { rmax[row] = m; rsum[row] = s; }
}

// S fp32 -> softmax probabilities in operand format (S is left untouched)
__global__ void softmax_operand_kernel(const float* __restrict__ S, int ld, long long rows, int n,
                                       const uint8_t* __restrict__ keep, long long rows_per_batch,
                                       __half* __restrict__ out, int ldo, long long plane, int planes, int bf16) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* r = S + row * ld;
    const uint8_t* kp = keep ? keep + (row / rows_per_batch) * n : nullptr;
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) {
        float v = r[j];
        if (kp && !kp[j]) v = -1e9f;
        m = fmaxf(m, v);
    }
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < n; j += 32) {
        float v = r[j];
        if (kp && !kp[j]) v = -1e9f;
        s += expf(v - m);
    }
    s = warp_sum(s);
    unsigned short* o = reinterpret_cast<unsigned short*>(out + row * ldo);
    for (int j = lane; j < n; j += 32) {
        float v = r[j];
        if (kp && !kp[j]) v = -1e9f;
        const float pj = expf(v - m) / s;
        o[j] = (unsigned short)(tc::pack_h2(pj, 0.f, bf16) & 0xffff);
        if (planes == 2) o[plane + j] = (unsigned short)(tc::pack_h2(tc::lo_part(pj, bf16), 0.f, bf16) & 0xffff);
    }
}

// colsum[b, j] = sum over rows i of batch b of exp(S_ij - m_i) / z_i  (deterministic slab order)
__global__ void colsum_softmax_partial_kernel(const float* __restrict__ S, int ld, long long rows_per_batch, int n,
                                              const float* __restrict__ rmax, const float* __restrict__ rsum,
                                              int slabs, float* __restrict__ part) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int slab = blockIdx.y, b = blockIdx.z;
    if (j >= n) return;
    const long long per = (rows_per_batch + slabs - 1) / slabs;
    const long long r0 = slab * per, r1 = min(rows_per_batch, r0 + per);
    const long long base = (long long)b * rows_per_batch;
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r)
        acc += expf(S[(base + r) * ld + j] - rmax[base + r]) / rsum[base + r];
    part[((size_t)b * slabs + slab) * n + j] = acc;
}

// Fused partial-overlap key statistic (model/transformer.py:35-39): colsum[b, j] = sum over the rows i of batch b of
// softmax_j(S_i*)_j, with S read from HBM ONCE: a CTA stages RB whole rows in shared memory, takes their max / sum there,
// then one thread per column adds the normalised exponentials of its RB rows (fixed order) into a per-slab partial.
// Deterministic: partials are reduced in slab order by colsum_final_kernel.
__global__ void __launch_bounds__(256, 2)
softmax_colsum_fused_kernel(const float* __restrict__ S, int ld, long long rows_per_batch, int n, int RB,
                            float* __restrict__ part) {
    extern __shared__ __align__(16) float fsm[];
    float* tile = fsm;                       // [RB][ld]
    float* rinv = fsm + (size_t)RB * ld;     // [RB]  1 / row sum
    const int slab = blockIdx.x, b = blockIdx.y, slabs = gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r0 = (long long)slab * RB;
    const int nr = (int)min((long long)RB, rows_per_batch - r0);
    const float* base = S + ((long long)b * rows_per_batch + r0) * ld;
    // the whole slab is requested up front (cp.async, 16 B per request): ~100 KB in flight per CTA, two CTAs per SM
    const int ld4 = ld >> 2;
    for (int e = threadIdx.x; e < nr * ld4; e += 256) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(tile + (size_t)e * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(base + (size_t)e * 4) : "memory");
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
    for (int r = warp; r < nr; r += 8) {
        float* t = tile + (size_t)r * ld;
        float m = -INFINITY;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, t[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < n; j += 32) {
            const float e = expf(t[j] - m);
            t[j] = e;
            s += e;
        }
        s = warp_sum(s);
        if (lane == 0) rinv[r] = 1.f / s;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += 256) {
        float acc = 0.f;
        for (int r = 0; r < nr; ++r) acc = fmaf(tile[(size_t)r * ld + j], rinv[r], acc);
        part[((size_t)b * slabs + slab) * n + j] = acc;
    }
}

// rowsum[r] = sum_j P[r, j]   (model/vcrnet_model.py:244 after the dim=1 softmax)
__global__ void rowsum_kernel(const float* __restrict__ P, int ld, long long rows, int n, float* __restrict__ out) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* r = P + row * ld;
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += r[j];
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

// squared row norms, plain fp32 (used by the VCP head: model/vcrnet_model.py:338-339)
__global__ void sqnorm_rows_kernel(const float* __restrict__ X, int ld, long long rows, int D, float* __restrict__ out) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* r = reinterpret_cast<const float4*>(X + row * ld);
    float s = 0.f;
    for (int j = lane; j < D / 4; j += 32) {
        const float4 v = r[j];
        s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

// ---------------------------------------------------------------------------------------------
// VCP head row pass (model/vcrnet_model.py:341-345): given dot[i,j] = s_i . t_j,
//   pd_ij = (-xx_i - (-2 dot_ij)) - yy_j ; P = softmax_j(pd) ; corr[:, i] = sum_j P_ij tgt[:, j].
// mode 0: write corr (soft correspondence, whole).  mode 1: write P in place (partial selectCom needs
// the matrix for column/row statistics).  mode 2: write (argmax_j, max_j P_ij) (partial getCopair).
// ---------------------------------------------------------------------------------------------
__global__ void softcorr_rows_kernel(float* __restrict__ dot, int ld, int Ns, int Nt,
                                     const float* __restrict__ xx, const float* __restrict__ yy,
                                     const float* __restrict__ tgt /*[B,3,Nt]*/, int mode,
                                     float* __restrict__ corr /*[B,3,Ns]*/,
                                     int* __restrict__ best_idx, float* __restrict__ best_val) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= Ns) return;
    float* r = dot + ((size_t)b * Ns + i) * ld;
    const float* yb = yy + (size_t)b * Nt;
    const float nx = -xx[(size_t)b * Ns + i];
    float m = -INFINITY;
    int mj = 0x7fffffff;
    for (int j = lane; j < Nt; j += 32) {
        const float pd = __fsub_rn(__fsub_rn(nx, -2.f * r[j]), yb[j]);
        r[j] = pd;
        if (pd > m) { m = pd; mj = j; }
    }
    // arg-max with ties -> lower index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oj = __shfl_xor_sync(0xffffffffu, mj, o);
        if (om > m || (om == m && oj < mj)) { m = om; mj = oj; }
    }
    float s = 0.f;
    for (int j = lane; j < Nt; j += 32) {
        const float e = expf(r[j] - m);
        r[j] = e;
        s += e;
    }
    s = warp_sum(s);
    if (mode == 0) {
        const float* tb = tgt + (size_t)b * 3 * Nt;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        for (int j = lane; j < Nt; j += 32) {
            const float pj = r[j] / s;
            cx = fmaf(pj, tb[j], cx);
            cy = fmaf(pj, tb[Nt + j], cy);
            cz = fmaf(pj, tb[2 * Nt + j], cz);
        }
        cx = warp_sum(cx); cy = warp_sum(cy); cz = warp_sum(cz);
        if (lane == 0) {
            float* cb = corr + (size_t)b * 3 * Ns;
            cb[i] = cx; cb[Ns + i] = cy; cb[2 * Ns + i] = cz;
        }
    } else if (mode == 1) {
        for (int j = lane; j < Nt; j += 32) r[j] = r[j] / s;
    } else {
        if (lane == 0) {
            best_idx[(size_t)b * Ns + i] = mj;
            best_val[(size_t)b * Ns + i] = 1.0f / s;       // exp(0)/s, the max probability
        }
    }
}

// column softmax statistics for selectCom's second pass (model/vcrnet_model.py:243-244):
//   rowsum_i = sum_j softmax over i (dim=1) of pd_ij.  Needs column max and column sum-exp first.
// Thread-per-column kernels over the pd matrix (coalesced along j).
// Column statistics are split over row slabs so that the whole GPU works on a [Ns, Nt] matrix (a thread-per-column
// kernel walking all Ns rows leaves 97 % of the SMs idle): slab maxima, then slab sums of exp(v - column max), each
// reduced in slab order (deterministic).
constexpr int COL_SLABS = 24;
__global__ void col_max_partial_kernel(const float* __restrict__ pd, int ld, int Ns, int Nt, float* __restrict__ pmax) {
    const int b = blockIdx.z, slab = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Nt) return;
    const int per = (Ns + COL_SLABS - 1) / COL_SLABS;
    const int i0 = slab * per, i1 = min(Ns, i0 + per);
    const float* p = pd + (size_t)b * Ns * ld + j;
    float m = -INFINITY;
    for (int i = i0; i < i1; ++i) m = fmaxf(m, p[(size_t)i * ld]);
    pmax[((size_t)b * COL_SLABS + slab) * Nt + j] = m;
}
__global__ void col_sum_partial_kernel(const float* __restrict__ pd, int ld, int Ns, int Nt,
                                       const float* __restrict__ pmax, float* __restrict__ cmax,
                                       float* __restrict__ psum) {
    const int b = blockIdx.z, slab = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Nt) return;
    float m = -INFINITY;
    for (int t = 0; t < COL_SLABS; ++t) m = fmaxf(m, pmax[((size_t)b * COL_SLABS + t) * Nt + j]);
    if (slab == 0) cmax[(size_t)b * Nt + j] = m;
    const int per = (Ns + COL_SLABS - 1) / COL_SLABS;
    const int i0 = slab * per, i1 = min(Ns, i0 + per);
    const float* p = pd + (size_t)b * Ns * ld + j;
    float s = 0.f;
    for (int i = i0; i < i1; ++i) s += expf(p[(size_t)i * ld] - m);
    psum[((size_t)b * COL_SLABS + slab) * Nt + j] = s;
}
__global__ void rowsum_colsoftmax_kernel(const float* __restrict__ pd, int ld, int Ns, int Nt,
                                         const float* __restrict__ cmax, const float* __restrict__ csum,
                                         float* __restrict__ out) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= Ns) return;
    const float* r = pd + ((size_t)b * Ns + i) * ld;
    const float* cm = cmax + (size_t)b * Nt;
    const float* cs = csum + (size_t)b * Nt;
    float s = 0.f;
    for (int j = lane; j < Nt; j += 32) s += expf(r[j] - cm[j]) / cs[j];
    s = warp_sum(s);
    if (lane == 0) out[(size_t)b * Ns + i] = s;
}

// pd matrix only (no softmax): pd_ij = (-xx_i + 2 dot_ij) - yy_j, in place
__global__ void negdist_kernel(float* __restrict__ dot, int ld, int Ns, int Nt, const float* __restrict__ xx,
                               const float* __restrict__ yy) {
    const int b = blockIdx.z;
    const int i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Nt) return;
    float* r = dot + ((size_t)b * Ns + i) * ld;
    r[j] = __fsub_rn(__fsub_rn(-xx[(size_t)b * Ns + i], -2.f * r[j]), yy[(size_t)b * Nt + j]);
}

// ---------------------------------------------------------------------------------------------
// selectCom statistics in two reads of the score products (model/vcrnet_model.py:213-222, 243-244):
//   pd_ij    = (-xx_i - (-2 dot_ij)) - yy_j                              (reference op order, never written back)
//   col_stat = sum_i softmax_j(pd)_ij   (row softmax, summed over rows)    -> picks the target points (:222)
//   row_stat = sum_j softmax_i(pd)_ij   (column softmax, summed over cols) -> picks the source points (:244)
// A CTA owns a slab of RS rows of one pair, staged as pd in shared memory.  Pass A: exact row (max, sum exp) per row and
// per-slab column (max, sum exp) partials; a small kernel combines the slabs per column (8 interleaved slab groups, fixed
// order); pass B recomputes pd and accumulates both weighted sums (rows complete, columns as per-slab partials reduced the
// same way by select_stats_final_kernel).  Replaces negdist (R+W) + col_max / col_sum partials + rowsum_colsoftmax + softmax_rows (3R+2W) +
// colsum: ~9 reads and 3 writes of the [Ns, Nt] matrix become 2 reads.  Deterministic (fixed slab order, no atomics).
// ---------------------------------------------------------------------------------------------
constexpr int SEL_T = 256;
constexpr int SEL_RS = SEL_T / 32;          // rows per slab: one warp per row

// warp `warp` loads row i0 + warp of the slab: pd values into shared memory (for the column phase) and, lane-strided by
// float4, into v[] (for the row phase) -- no second pass over shared memory for the rows.  Columns past Nt hold -inf.
template <int NV>
__device__ __forceinline__ void sel_load_row(const float* __restrict__ drow, int Nt, float nx, const float* __restrict__ yy,
                                             float* srow, float4 (&v)[NV], int lane, int vec) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int c = (lane + q * 32) * 4;
        float4 o = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (vec) {
            if (c < Nt) {
                const float4 d = *reinterpret_cast<const float4*>(drow + c);
                const float4 y = *reinterpret_cast<const float4*>(yy + c);
                o.x = __fsub_rn(__fsub_rn(nx, -2.f * d.x), y.x); o.y = __fsub_rn(__fsub_rn(nx, -2.f * d.y), y.y);
                o.z = __fsub_rn(__fsub_rn(nx, -2.f * d.z), y.z); o.w = __fsub_rn(__fsub_rn(nx, -2.f * d.w), y.w);
            }
        } else {                                          // ragged / unaligned: element-wise
            if (c + 0 < Nt) o.x = __fsub_rn(__fsub_rn(nx, -2.f * drow[c + 0]), yy[c + 0]);
            if (c + 1 < Nt) o.y = __fsub_rn(__fsub_rn(nx, -2.f * drow[c + 1]), yy[c + 1]);
            if (c + 2 < Nt) o.z = __fsub_rn(__fsub_rn(nx, -2.f * drow[c + 2]), yy[c + 2]);
            if (c + 3 < Nt) o.w = __fsub_rn(__fsub_rn(nx, -2.f * drow[c + 3]), yy[c + 3]);
        }
        v[q] = o;
        *reinterpret_cast<float4*>(srow + c) = o;         // ldp >= 128 * NV: always in range
    }
}

template <int NV>
__global__ void __launch_bounds__(SEL_T)
select_stats_a_kernel(const float* __restrict__ dot, int ld, int Ns, int Nt, int vec, const float* __restrict__ xx,
                      const float* __restrict__ yy, float* __restrict__ rmax, float* __restrict__ rsum,
                      float* __restrict__ pmax, float* __restrict__ psum) {
    constexpr int LDP = NV * 128;
    __shared__ __align__(16) float sm[SEL_RS * LDP];
    const int b = blockIdx.y, slab = blockIdx.x, slabs = gridDim.x;
    const int i0 = slab * SEL_RS, nr = min(SEL_RS, Ns - i0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < nr) {
        const int i = i0 + warp;
        float4 v[NV];
        sel_load_row<NV>(dot + ((size_t)b * Ns + i) * ld, Nt, -xx[(size_t)b * Ns + i], yy + (size_t)b * Nt, sm + warp * LDP, v,
                         lane, vec);
        float m = -INFINITY;
#pragma unroll
        for (int q = 0; q < NV; ++q) m = fmaxf(fmaxf(m, fmaxf(v[q].x, v[q].y)), fmaxf(v[q].z, v[q].w));
        m = warp_max(m);
        float sacc = 0.f;
#pragma unroll
        for (int q = 0; q < NV; ++q) sacc += (expf(v[q].x - m) + expf(v[q].y - m)) + (expf(v[q].z - m) + expf(v[q].w - m));
        sacc = warp_sum(sacc);                            // padded columns are -inf: exp gives exactly 0
        if (lane == 0) { rmax[(size_t)b * Ns + i] = m; rsum[(size_t)b * Ns + i] = sacc; }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Nt; j += SEL_T) {
        float m = -INFINITY;
        for (int r = 0; r < nr; ++r) m = fmaxf(m, sm[r * LDP + j]);
        float sacc = 0.f;
        for (int r = 0; r < nr; ++r) sacc += expf(sm[r * LDP + j] - m);
        pmax[((size_t)b * slabs + slab) * Nt + j] = m;
        psum[((size_t)b * slabs + slab) * Nt + j] = sacc;
    }
}
// Per column: (max, sum exp) over the slabs' partials and the reciprocal of the sum.  A CTA owns 32 columns; its 8 warps each
// scan every 8th slab (coalesced 128-byte rows), and the 8 partial results are combined in warp order -- fixed order, no
// atomics.  (One thread per column walking all ~100 slabs was a 25 us chain of dependent L2 loads.)
constexpr int SEL_CG = 8;                   // slab groups (warps) per CTA of the combine / final kernels
__global__ void __launch_bounds__(32 * SEL_CG)
select_stats_combine_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, int slabs, int Nt,
                            float* __restrict__ cmax, float* __restrict__ csum, float* __restrict__ crcp) {
    __shared__ float sm_m[SEL_CG][32], sm_s[SEL_CG][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane, b = blockIdx.y;
    const bool ok = j < Nt;
    const float* pm = pmax + (size_t)b * slabs * Nt + j;
    const float* ps = psum + (size_t)b * slabs * Nt + j;
    float m = -INFINITY;
    if (ok)
        for (int t = grp; t < slabs; t += SEL_CG) m = fmaxf(m, pm[(size_t)t * Nt]);
    sm_m[grp][lane] = m;
    __syncthreads();
#pragma unroll
    for (int g = 0; g < SEL_CG; ++g) m = fmaxf(m, sm_m[g][lane]);
    float sacc = 0.f;
    if (ok)
        for (int t = grp; t < slabs; t += SEL_CG) sacc += ps[(size_t)t * Nt] * expf(pm[(size_t)t * Nt] - m);
    sm_s[grp][lane] = sacc;
    __syncthreads();
    if (grp == 0 && ok) {
        float tot = 0.f;
#pragma unroll
        for (int g = 0; g < SEL_CG; ++g) tot += sm_s[g][lane];
        cmax[(size_t)b * Nt + j] = m;
        csum[(size_t)b * Nt + j] = tot;
        crcp[(size_t)b * Nt + j] = 1.f / tot;
    }
}
// out[b, j] = sum over slabs of part[b, slab, j]: 8 warps take every 8th slab, partials added in warp order
__global__ void __launch_bounds__(32 * SEL_CG)
select_stats_final_kernel(const float* __restrict__ part, int slabs, int Nt, float* __restrict__ out) {
    __shared__ float sm_s[SEL_CG][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane, b = blockIdx.y;
    const bool ok = j < Nt;
    const float* pp = part + (size_t)b * slabs * Nt + j;
    float sacc = 0.f;
    if (ok)
        for (int t = grp; t < slabs; t += SEL_CG) sacc += pp[(size_t)t * Nt];
    sm_s[grp][lane] = sacc;
    __syncthreads();
    if (grp == 0 && ok) {
        float tot = 0.f;
#pragma unroll
        for (int g = 0; g < SEL_CG; ++g) tot += sm_s[g][lane];
        out[(size_t)b * Nt + j] = tot;
    }
}
template <int NV>
__global__ void __launch_bounds__(SEL_T)
select_stats_b_kernel(const float* __restrict__ dot, int ld, int Ns, int Nt, int vec, const float* __restrict__ xx,
                      const float* __restrict__ yy, const float* __restrict__ rmax, const float* __restrict__ rsum,
                      const float* __restrict__ cmax, const float* __restrict__ csum, const float* __restrict__ crcp,
                      float* __restrict__ row_stat, float* __restrict__ cpart) {
    constexpr int LDP = NV * 128;
    __shared__ __align__(16) float sm[SEL_RS * LDP];
    __shared__ float s_rm[SEL_RS], s_rs[SEL_RS], s_rr[SEL_RS];
    const int b = blockIdx.y, slab = blockIdx.x, slabs = gridDim.x;
    const int i0 = slab * SEL_RS, nr = min(SEL_RS, Ns - i0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* cm = cmax + (size_t)b * Nt;
    const float* cs = csum + (size_t)b * Nt;
    const float* cr = crcp + (size_t)b * Nt;      // 1 / csum: the quotients below are div_by's (correctly rounded, 3 instructions)
    if (warp < nr) {                                       // column softmax, summed along the row (:243-244)
        const int i = i0 + warp;
        float4 v[NV];
        sel_load_row<NV>(dot + ((size_t)b * Ns + i) * ld, Nt, -xx[(size_t)b * Ns + i], yy + (size_t)b * Nt, sm + warp * LDP, v,
                         lane, vec);
        float sacc = 0.f;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int c = (lane + q * 32) * 4;
            const float e[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c + u < Nt) sacc += div_by(expf(e[u] - cm[c + u]), cs[c + u], cr[c + u]);
        }
        sacc = warp_sum(sacc);
        if (lane == 0) {
            row_stat[(size_t)b * Ns + i] = sacc;
            s_rm[warp] = rmax[(size_t)b * Ns + i]; s_rs[warp] = rsum[(size_t)b * Ns + i];
            s_rr[warp] = 1.f / s_rs[warp];
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Nt; j += SEL_T) {        // row softmax, summed down the column (:221-222)
        float sacc = 0.f;
        for (int r = 0; r < nr; ++r) sacc += div_by(expf(sm[r * LDP + j] - s_rm[r]), s_rs[r], s_rr[r]);
        cpart[((size_t)b * slabs + slab) * Nt + j] = sacc;
    }
}

// operand-format row gather: out[pl][b*K + r][:] = in[pl][b*Nin + idx[b, r]][:] for both planes, sq_out[b, r] = sq_in[b, idx[b, r]]
// (the rows and squared norms getCopair needs of the points selectCom kept; C % 8 == 0).  One warp per output row.
__global__ void gather_operand_rows_kernel(const __half* __restrict__ in, int ld_in, long long plane_in, int Nin,
                                           const int* __restrict__ idx, int K, int C, __half* __restrict__ out, int ld_out,
                                           long long plane_out, const float* __restrict__ sq_in, float* __restrict__ sq_out) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= K) return;
    const int src = idx[(size_t)b * K + r];
    const __half* ip = in + ((size_t)b * Nin + src) * ld_in;
    __half* op = out + ((size_t)b * K + r) * ld_out;
    for (int c = lane * 8; c < C; c += 256) {
        *reinterpret_cast<uint4*>(op + c) = *reinterpret_cast<const uint4*>(ip + c);
        *reinterpret_cast<uint4*>(op + plane_out + c) = *reinterpret_cast<const uint4*>(ip + plane_in + c);
    }
    if (lane == 0 && sq_in) sq_out[(size_t)b * K + r] = sq_in[(size_t)b * Nin + src];
}

// gather rows: out[b, r, :] = in[b, idx[b, r], :]   (C % 4 == 0)
__global__ void gather_rows_kernel(const float* __restrict__ in, int ld_in, int Nin, const int* __restrict__ idx,
                                   int K, int C, float* __restrict__ out, int ld_out) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= K) return;
    const int src = idx[(size_t)b * K + r];
    const float4* s = reinterpret_cast<const float4*>(in + ((size_t)b * Nin + src) * ld_in);
    float4* d = reinterpret_cast<float4*>(out + ((size_t)b * K + r) * ld_out);
    for (int c = lane; c < C / 4; c += 32) d[c] = s[c];
}

// gather columns of a channel-major cloud: out[b, c, r] = in[b, c, idx[b, r]]
__global__ void gather_cols_kernel(const float* __restrict__ in, int Cc, int Nin, const int* __restrict__ idx,
                                   int K, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= K) return;
    const int src = idx[(size_t)b * K + r];
    for (int c = 0; c < Cc; ++c)
        out[((size_t)b * Cc + c) * K + r] = in[((size_t)b * Cc + c) * Nin + src];
}

// getCopair's final gathers (model/vcrnet_model.py:300-331, tgtK == 1 so the weight is exactly 1)
__global__ void copair_gather_kernel(const float* __restrict__ src, const float* __restrict__ tgt, int Ns, int Nt,
                                     const int* __restrict__ keep, const int* __restrict__ best, int K,
                                     float* __restrict__ so, float* __restrict__ co) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= K) return;
    const int si = keep[(size_t)b * K + r];
    const int ti = best[(size_t)b * Ns + si];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        so[((size_t)b * 3 + c) * K + r] = src[((size_t)b * 3 + c) * Ns + si];
        co[((size_t)b * 3 + c) * K + r] = tgt[((size_t)b * 3 + c) * Nt + ti];
    }
}

__global__ void add_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o, size_t n4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 x = a[i], y = b[i];
    o[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

}  // namespace

// in [nb, R, C] (ld_in) -> out [nb, C, R] (ld_out)
VCR_API int vcr_transpose(const float* in, float* out, int nb, int R, int C, int ld_in, int ld_out,
                          long long stride_in, long long stride_out, cudaStream_t stream) {
    VCR_REQUIRE(in && out && nb > 0 && R > 0 && C > 0 && nb <= 65535);
    dim3 g(vcr_cdiv(C, 32), vcr_cdiv(R, 32), nb), blk(32, 8);
    transpose_kernel<<<g, blk, 0, stream>>>(in, out, R, C, ld_in, ld_out, stride_in, stride_out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_conv3_act(const float* xyz, const float* w, const float* bias, int B, int N, int Cout, float slope,
                          float* out, int ldo, cudaStream_t stream) {
    VCR_REQUIRE(xyz && w && bias && out && B > 0 && N > 0 && Cout > 0 && B <= 65535);
    dim3 g(vcr_cdiv(N, 8), B), blk(32, 8);
    conv3_kernel<<<g, blk, 0, stream>>>(xyz, w, bias, N, Cout, slope, out, ldo);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// out = a*(x-mean)/(std_unbiased+eps)+b (+ residual when residual != NULL)
// conv1_lpd + conv2_lpd (+ operand copy of the result): xyz [B,3,N]; w1 [64,3], w2 [64,64] row-major; h1 (nullable) / h2 [B*N,64];
// op (nullable): "h3" operand planes [2][B*N][ldop] of h2.
VCR_API int vcr_lpd_point_mlp(const float* xyz, const float* w1, const float* b1, const float* w2, const float* b2, int B, int N,
                              float slope, float* h1, float* h2, void* op, int ldop, long long op_plane, cudaStream_t stream) {
    VCR_REQUIRE(xyz && w1 && b1 && w2 && b2 && h2 && B > 0 && N > 0 && B <= 65535);
    if (op && ((ldop & 3) || (op_plane & 3) || (reinterpret_cast<uintptr_t>(op) & 7))) return VCR_ERR_INVALID;
    dim3 g(vcr_cdiv(N, 128), B);
    lpd_point_mlp_kernel<<<g, 128, 0, stream>>>(xyz, w1, b1, w2, b2, N, slope, h1, h2, reinterpret_cast<__half*>(op), ldop, op_plane);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_layernorm(const float* x, int ldx, const float* a, const float* b, float eps, long long M, int D,
                          const float* residual, int ldr, float* out, int ldo, cudaStream_t stream) {
    VCR_REQUIRE(x && a && b && out && M > 0);
    if (D % 128 != 0 || D > 1024 || (ldx & 3) || (ldo & 3) || (residual && (ldr & 3))) return VCR_ERR_UNSUPPORTED;
    const int wpb = 8;
    dim3 g(vcr_cdiv(M, wpb));
    switch (D / 128) {
#define LN_CASE(V) case V: layernorm_kernel<V, false><<<g, wpb * 32, 0, stream>>>(x, ldx, a, b, eps, (int)M, D, residual, ldr, out, ldo, nullptr, 0, 0, nullptr); break;
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
#undef LN_CASE
    }
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// vcr_layernorm that also emits what the VCP head needs from its output: the operand-format ("h3": fp16 hi, lo * 2^11)
// copy [2][M][ldop] and the squared row norms sq[M] (bit-identical to vcr_to_operand / vcr_sqnorm_rows of `out`).
VCR_API int vcr_layernorm_head(const float* x, int ldx, const float* a, const float* b, float eps, long long M, int D,
                               const float* residual, int ldr, float* out, int ldo, void* op, int ldop,
                               long long plane_stride, float* sq, cudaStream_t stream) {
    VCR_REQUIRE(x && a && b && op && sq && M > 0);             // out may be NULL: operand copy + norms only
    if (D % 128 != 0 || D > 1024 || (ldx & 3) || (out && (ldo & 3)) || (residual && (ldr & 3)) || (ldop & 3) || (plane_stride & 3))
        return VCR_ERR_UNSUPPORTED;
    const int wpb = 8;
    dim3 g(vcr_cdiv(M, wpb));
    __half* oh = reinterpret_cast<__half*>(op);
    switch (D / 128) {
#define LN_CASE(V) case V: layernorm_kernel<V, true><<<g, wpb * 32, 0, stream>>>(x, ldx, a, b, eps, (int)M, D, residual, ldr, out, ldo, oh, ldop, plane_stride, sq); break;
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
#undef LN_CASE
    }
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_softmax_rows(float* S, int ld, long long rows, int n, const uint8_t* keep, long long rows_per_batch,
                             cudaStream_t stream) {
    VCR_REQUIRE(S && rows > 0 && n > 0 && (!keep || rows_per_batch > 0));
    softmax_rows_kernel<<<vcr_cdiv(rows, 8), 256, 0, stream>>>(S, ld, rows, n, keep, rows_per_batch);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API size_t vcr_colsum_workspace_bytes(int B, int n) { return (size_t)B * 32 * n * sizeof(float); }

VCR_API int vcr_colsum(const float* P, int ld, int B, long long rows_per_batch, int n, float* out, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
    VCR_REQUIRE(P && out && B > 0 && rows_per_batch > 0 && n > 0 && B <= 65535);
    const int slabs = 32;
    if (!workspace || workspace_bytes < vcr_colsum_workspace_bytes(B, n)) return VCR_ERR_WORKSPACE;
    float* part = reinterpret_cast<float*>(workspace);
    dim3 g(vcr_cdiv(n, 128), slabs, B);
    colsum_partial_kernel<<<g, 128, 0, stream>>>(P, ld, rows_per_batch, n, slabs, part);
    VCR_CHECK_LAUNCH();
    dim3 g2(vcr_cdiv(n, 128), B);
    colsum_final_kernel<<<g2, 128, 0, stream>>>(part, slabs, n, out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}


VCR_API int vcr_layernorm_operand(const float* x, int ldx, const float* a, const float* b, float eps, long long M, int D,
                                  void* out, int ldo, long long plane_stride, int planes, int bf16, cudaStream_t stream) {
    VCR_REQUIRE(x && a && b && out && M > 0 && (planes == 1 || planes == 2));
    if (D % 128 != 0 || D > 1024 || (ldx & 3) || (ldo & 3)) return VCR_ERR_UNSUPPORTED;
    const int wpb = 8;
    dim3 g(vcr_cdiv(M, wpb));
    __half* o = reinterpret_cast<__half*>(out);
    switch (D / 128) {
#define LN_CASE(V) case V: \
        if (planes == 2) layernorm_operand_kernel<V, 2, 0><<<g, wpb * 32, 0, stream>>>(x, ldx, a, b, eps, (int)M, D, o, ldo, plane_stride); \
        else if (bf16) layernorm_operand_kernel<V, 1, 1><<<g, wpb * 32, 0, stream>>>(x, ldx, a, b, eps, (int)M, D, o, ldo, plane_stride); \
        else layernorm_operand_kernel<V, 1, 0><<<g, wpb * 32, 0, stream>>>(x, ldx, a, b, eps, (int)M, D, o, ldo, plane_stride); \
        break;
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
#undef LN_CASE
    }
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_row_lse(const float* S, int ld, long long rows, int n, const uint8_t* keep, long long rows_per_batch,
                        float* rmax, float* rsum, cudaStream_t stream) {
    VCR_REQUIRE(S && rmax && rsum && rows > 0 && n > 0 && (!keep || rows_per_batch > 0));
    row_lse_kernel<<<vcr_cdiv(rows, 8), 256, 0, stream>>>(S, ld, rows, n, keep, rows_per_batch, rmax, rsum);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_softmax_operand(const float* S, int ld, long long rows, int n, const uint8_t* keep,
                                long long rows_per_batch, void* out, int ldo, long long plane_stride, int planes,
                                int bf16, cudaStream_t stream) {
    VCR_REQUIRE(S && out && rows > 0 && n > 0 && (!keep || rows_per_batch > 0) && (planes == 1 || planes == 2));
    softmax_operand_kernel<<<vcr_cdiv(rows, 8), 256, 0, stream>>>(S, ld, rows, n, keep, rows_per_batch,
                                                                 reinterpret_cast<__half*>(out), ldo, plane_stride,
                                                                 planes, bf16);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// column sums of softmax(S) per batch without materialising the probabilities (workspace as vcr_colsum)
VCR_API int vcr_colsum_softmax(const float* S, int ld, int B, long long rows_per_batch, int n, const float* rmax,
                               const float* rsum, float* out, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream) {
    VCR_REQUIRE(S && rmax && rsum && out && B > 0 && rows_per_batch > 0 && n > 0 && B <= 65535);
    const int slabs = 32;
    if (!workspace || workspace_bytes < vcr_colsum_workspace_bytes(B, n)) return VCR_ERR_WORKSPACE;
    float* part = reinterpret_cast<float*>(workspace);
    dim3 g(vcr_cdiv(n, 128), slabs, B);
    colsum_softmax_partial_kernel<<<g, 128, 0, stream>>>(S, ld, rows_per_batch, n, rmax, rsum, slabs, part);
    VCR_CHECK_LAUNCH();
    dim3 g2(vcr_cdiv(n, 128), B);
    colsum_final_kernel<<<g2, 128, 0, stream>>>(part, slabs, n, out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

static int fused_colsum_rows_per_cta(int ld) {
    int rb = (int)((110 * 1024) / ((size_t)ld * sizeof(float) + sizeof(float)));    // two CTAs per SM
    return rb > 64 ? 64 : rb;
}
VCR_API size_t vcr_softmax_colsum_workspace_bytes(int B, long long rows_per_batch, int ld, int n) {
    const int rb = fused_colsum_rows_per_cta(ld);
    if (rb < 1) return 0;
    return (size_t)B * ((rows_per_batch + rb - 1) / rb) * n * sizeof(float);
}
// out [B, n] = per-batch column sums of the row softmax of S [B*rows_per_batch, ld] (first n columns), S read once.
VCR_API int vcr_softmax_colsum(const float* S, int ld, int B, long long rows_per_batch, int n, float* out,
                               void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    VCR_REQUIRE(S && out && B > 0 && rows_per_batch > 0 && n > 0 && n <= ld && B <= 65535 && (ld & 3) == 0 &&
                (reinterpret_cast<uintptr_t>(S) & 15) == 0);
    const int rb = fused_colsum_rows_per_cta(ld);
    if (rb < 1) return VCR_ERR_UNSUPPORTED;
    const long long slabs = (rows_per_batch + rb - 1) / rb;
    if (slabs > 0x7fffffff) return VCR_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < vcr_softmax_colsum_workspace_bytes(B, rows_per_batch, ld, n)) return VCR_ERR_WORKSPACE;
    const size_t smem = ((size_t)rb * ld + rb) * sizeof(float);
    // set on every launch: the attribute is per device, and one process may drive several (nn.DataParallel)
    if (cudaFuncSetAttribute(softmax_colsum_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return VCR_ERR_LAUNCH;
    float* part = reinterpret_cast<float*>(workspace);
    dim3 g((unsigned)slabs, B);
    softmax_colsum_fused_kernel<<<g, 256, smem, stream>>>(S, ld, rows_per_batch, n, rb, part);
    VCR_CHECK_LAUNCH();
    dim3 g2(vcr_cdiv(n, 128), B);
    colsum_final_kernel<<<g2, 128, 0, stream>>>(part, (int)slabs, n, out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_rowsum(const float* P, int ld, long long rows, int n, float* out, cudaStream_t stream) {
    VCR_REQUIRE(P && out && rows > 0 && n > 0);
    rowsum_kernel<<<vcr_cdiv(rows, 8), 256, 0, stream>>>(P, ld, rows, n, out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_sqnorm_rows(const float* X, int ld, long long rows, int D, float* out, cudaStream_t stream) {
    VCR_REQUIRE(X && out && rows > 0 && D > 0);
    if ((D & 3) || (ld & 3)) return VCR_ERR_UNSUPPORTED;
    sqnorm_rows_kernel<<<vcr_cdiv(rows, 8), 256, 0, stream>>>(X, ld, rows, D, out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// dot: [B, Ns, ld] holding s_i . t_j on entry (overwritten).  mode 0: corr out; 1: P in place;
// 2: best_idx / best_val out.
VCR_API int vcr_softcorr_rows(float* dot, int ld, int B, int Ns, int Nt, const float* xx, const float* yy,
                              const float* tgt, int mode, float* corr, int* best_idx, float* best_val,
                              cudaStream_t stream) {
    VCR_REQUIRE(dot && xx && yy && B > 0 && Ns > 0 && Nt > 0 && B <= 65535);
    if (mode == 0) VCR_REQUIRE(tgt && corr);
    if (mode == 2) VCR_REQUIRE(best_idx && best_val);
    dim3 g(vcr_cdiv(Ns, 8), B);
    softcorr_rows_kernel<<<g, 256, 0, stream>>>(dot, ld, Ns, Nt, xx, yy, tgt, mode, corr, best_idx, best_val);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_negdist(float* dot, int ld, int B, int Ns, int Nt, const float* xx, const float* yy,
                        cudaStream_t stream) {
    VCR_REQUIRE(dot && xx && yy && B > 0 && Ns > 0 && Nt > 0 && Ns <= 65535 && B <= 65535);
    dim3 g(vcr_cdiv(Nt, 256), Ns, B);
    negdist_kernel<<<g, 256, 0, stream>>>(dot, ld, Ns, Nt, xx, yy);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API size_t vcr_rowsum_colsoftmax_workspace_bytes(int B, int Nt) {
    return (size_t)(2 + 2 * COL_SLABS) * B * Nt * sizeof(float);
}
// rowsum of the column-softmax of pd (workspace: vcr_rowsum_colsoftmax_workspace_bytes)
VCR_API int vcr_rowsum_colsoftmax(const float* pd, int ld, int B, int Ns, int Nt, float* out, void* workspace,
                                  size_t workspace_bytes, cudaStream_t stream) {
    VCR_REQUIRE(pd && out && B > 0 && Ns > 0 && Nt > 0 && B <= 65535);
    if (!workspace || workspace_bytes < vcr_rowsum_colsoftmax_workspace_bytes(B, Nt)) return VCR_ERR_WORKSPACE;
    float* cmax = reinterpret_cast<float*>(workspace);
    float* csum = cmax + (size_t)B * Nt;
    float* crcp = csum + (size_t)B * Nt;
    float* pmax = crcp + (size_t)B * Nt;
    float* psum = pmax + (size_t)COL_SLABS * B * Nt;
    dim3 g(vcr_cdiv(Nt, 128), COL_SLABS, B);
    col_max_partial_kernel<<<g, 128, 0, stream>>>(pd, ld, Ns, Nt, pmax);
    VCR_CHECK_LAUNCH();
    col_sum_partial_kernel<<<g, 128, 0, stream>>>(pd, ld, Ns, Nt, pmax, cmax, psum);
    VCR_CHECK_LAUNCH();
    dim3 g1(vcr_cdiv(Nt, 128), B);
    colsum_final_kernel<<<g1, 128, 0, stream>>>(psum, COL_SLABS, Nt, csum);
    VCR_CHECK_LAUNCH();
    dim3 g2(vcr_cdiv(Ns, 8), B);
    rowsum_colsoftmax_kernel<<<g2, 256, 0, stream>>>(pd, ld, Ns, Nt, cmax, csum, out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API size_t vcr_select_stats_workspace_bytes(int B, int Ns, int Nt) {
    const size_t slabs = (size_t)(Ns + SEL_RS - 1) / SEL_RS;
    return ((size_t)2 * B * Ns + (size_t)3 * B * Nt + (size_t)3 * B * slabs * Nt) * sizeof(float);
}
// selectCom's two selection statistics from the score products dot[B, Ns, ld] (left untouched), xx [B, Ns], yy [B, Nt]:
// row_stat [B, Ns] = row sums of the column softmax of pd, col_stat [B, Nt] = column sums of the row softmax of pd.
// Nt <= 1024 (a warp holds a row in registers).
VCR_API int vcr_select_stats(const float* dot, int ld, int B, int Ns, int Nt, const float* xx, const float* yy,
                             float* row_stat, float* col_stat, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream) {
    VCR_REQUIRE(dot && xx && yy && row_stat && col_stat && B > 0 && Ns > 0 && Nt > 0 && B <= 65535);
    if (Nt > 1024) return VCR_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < vcr_select_stats_workspace_bytes(B, Ns, Nt)) return VCR_ERR_WORKSPACE;
    // float4 loads of the dot rows and of yy[b, :] when everything is 16-byte aligned, element-wise otherwise
    const int vec = (Nt & 3) == 0 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(dot) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(yy) & 15) == 0;
    const int slabs = (Ns + SEL_RS - 1) / SEL_RS;
    float* rmax = reinterpret_cast<float*>(workspace);
    float* rsum = rmax + (size_t)B * Ns;
    float* cmax = rsum + (size_t)B * Ns;
    float* csum = cmax + (size_t)B * Nt;
    float* crcp = csum + (size_t)B * Nt;
    float* pmax = crcp + (size_t)B * Nt;
    float* psum = pmax + (size_t)B * slabs * Nt;
    float* cpart = psum + (size_t)B * slabs * Nt;
    dim3 g(slabs, B), gc(vcr_cdiv(Nt, 32), B);
    switch ((Nt + 127) / 128) {
#define SEL_CASE(V) case V: select_stats_a_kernel<V><<<g, SEL_T, 0, stream>>>(dot, ld, Ns, Nt, vec, xx, yy, rmax, rsum, pmax, psum); break;
        SEL_CASE(1) SEL_CASE(2) SEL_CASE(3) SEL_CASE(4) SEL_CASE(5) SEL_CASE(6) SEL_CASE(7) SEL_CASE(8)
#undef SEL_CASE
    }
    VCR_CHECK_LAUNCH();
    select_stats_combine_kernel<<<gc, 32 * SEL_CG, 0, stream>>>(pmax, psum, slabs, Nt, cmax, csum, crcp);
    VCR_CHECK_LAUNCH();
    switch ((Nt + 127) / 128) {
#define SEL_CASE(V) case V: select_stats_b_kernel<V><<<g, SEL_T, 0, stream>>>(dot, ld, Ns, Nt, vec, xx, yy, rmax, rsum, cmax, csum, crcp, row_stat, cpart); break;
        SEL_CASE(1) SEL_CASE(2) SEL_CASE(3) SEL_CASE(4) SEL_CASE(5) SEL_CASE(6) SEL_CASE(7) SEL_CASE(8)
#undef SEL_CASE
    }
    VCR_CHECK_LAUNCH();
    select_stats_final_kernel<<<gc, 32 * SEL_CG, 0, stream>>>(cpart, slabs, Nt, col_stat);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// Rows idx[b, :] of an operand-format buffer ([2 planes][B*Nin][ld_in] 16-bit) and of its squared norms, per batch item.
VCR_API int vcr_gather_operand_rows(const void* in, int ld_in, long long plane_in, int B, int Nin, const int* idx, int K,
                                    int C, void* out, int ld_out, long long plane_out, const float* sq_in, float* sq_out,
                                    cudaStream_t stream) {
    VCR_REQUIRE(in && idx && out && B > 0 && K > 0 && C > 0 && B <= 65535 && (!sq_in || sq_out));
    if ((C & 7) || (ld_in & 7) || (ld_out & 7) || (plane_in & 7) || (plane_out & 7)) return VCR_ERR_UNSUPPORTED;
    dim3 g(vcr_cdiv(K, 8), B);
    gather_operand_rows_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const __half*>(in), ld_in, plane_in, Nin, idx, K, C,
                                                      reinterpret_cast<__half*>(out), ld_out, plane_out, sq_in, sq_out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_gather_rows(const float* in, int ld_in, int B, int Nin, const int* idx, int K, int C, float* out,
                            int ld_out, cudaStream_t stream) {
    VCR_REQUIRE(in && idx && out && B > 0 && K > 0 && C > 0 && B <= 65535);
    if ((C & 3) || (ld_in & 3) || (ld_out & 3)) return VCR_ERR_UNSUPPORTED;
    dim3 g(vcr_cdiv(K, 8), B);
    gather_rows_kernel<<<g, 256, 0, stream>>>(in, ld_in, Nin, idx, K, C, out, ld_out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_gather_cols(const float* in, int B, int C, int Nin, const int* idx, int K, float* out,
                            cudaStream_t stream) {
    VCR_REQUIRE(in && idx && out && B > 0 && K > 0 && C > 0 && B <= 65535);
    dim3 g(vcr_cdiv(K, 128), B);
    gather_cols_kernel<<<g, 128, 0, stream>>>(in, C, Nin, idx, K, out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_copair_gather(const float* src, const float* tgt, int B, int Ns, int Nt, const int* keep,
                              const int* best_idx, int K, float* src_out, float* corr_out, cudaStream_t stream) {
    VCR_REQUIRE(src && tgt && keep && best_idx && src_out && corr_out && B > 0 && K > 0 && B <= 65535);
    dim3 g(vcr_cdiv(K, 128), B);
    copair_gather_kernel<<<g, 128, 0, stream>>>(src, tgt, Ns, Nt, keep, best_idx, K, src_out, corr_out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_add(const float* a, const float* b, float* out, long long n, cudaStream_t stream) {
    VCR_REQUIRE(a && b && out && n > 0 && (n & 3) == 0);
    add_kernel<<<vcr_cdiv(n / 4, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(a),
                                                        reinterpret_cast<const float4*>(b),
                                                        reinterpret_cast<float4*>(out), (size_t)(n / 4));
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}
