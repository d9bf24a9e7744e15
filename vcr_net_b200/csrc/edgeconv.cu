// EdgeConv kernels of the LPDNet embedding (reference model/lpdnet_model.py:122-132 on top of
// util/util.py:176-199), FP32 SIMT flavour.
//
// Algebra: the reference's edge feature is the raw concat [f_j ; x_i]
// (util/util.py:197), so conv(W,[f_j;x_i]) = W_a f_j + W_b x_i: two per-POINT GEMMs (done by the
// caller with vcr_gemm_*: PQ = [P | Q], P = W_a f, Q = W_b f + bias) plus a gather-add per EDGE.
// LeakyReLU with slope >= 0 is monotone, so max_k act(z_k) = act(max_k z_k).
//
//   vcr_edgeconv_dg : e1 = act(P[j]+Q[i]); x1 = max_k e1; e2 = W2 e1 + b2; x2 = act(max_k e2).
//                     The [N*k,128] edge tensor lives only in shared memory (reference: 335 MB in HBM
//                     per 32 clouds); DG2 is a real (N*k) x 128 x 128 GEMM done from smem.
//   vcr_gather_max  : x3 = act(max_k P3[j] + Q3[i])  (convSN1 + max collapses to a gather-max).
//
// Roofline: edgeconv_dg is FP32-FMA bound (2*128*128*k flops per point vs 4*(256+256)+4k bytes);
// gather_max is L2/HBM bound: 4*k*C bytes read per point (L2 hits after the first touch) + 4*C write.
#include "tc_common.cuh"     // pack_h2 / lo_part (operand-format outputs)

namespace {

constexpr int DG_C = 128;           // channels in / out of convDG2
constexpr int DG_K = 20;            // neighbours (LPDNet.k, model/lpdnet_model.py:81)
constexpr int DG_PT = 8;            // points per tile
constexpr int DG_ROWS = DG_PT * DG_K;   // 160 edge rows per tile
constexpr int DG_LDE = DG_C + 4;
constexpr int DG_LDW = DG_C + 4;
constexpr int DG_RT = 10;           // rows per thread (two threads-rows groups per point)

__global__ void __launch_bounds__(256, 1)
edgeconv_dg_kernel(const float* __restrict__ PQ, int ldpq, const int* __restrict__ idx, int N,
                   long long total_pts, const float* __restrict__ W2, const float* __restrict__ b2,
                   float slope, float* __restrict__ x1, int ld1, float* __restrict__ x2, int ld2) {
    extern __shared__ __align__(16) float sm[];
    float* Wt = sm;                              // [128 c][DG_LDW]  Wt[c][o] = W2[o][c]
    float* E = Wt + DG_C * DG_LDW;               // [160][DG_LDE]
    float* red = E + DG_ROWS * DG_LDE;           // [16][128]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    for (int e = tid; e < DG_C * DG_C; e += 256) {
        const int o = e >> 7, c = e & 127;
        Wt[c * DG_LDW + o] = W2[e];
    }
    const long long ntiles = (total_pts + DG_PT - 1) / DG_PT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long pt0 = tile * DG_PT;
        __syncthreads();
        // ---- build the edge tile e1[r][c] = act(P[nbr] + Q[centre]) -----------------------------
        for (int e = tid; e < DG_ROWS * (DG_C / 4); e += 256) {
            const int r = e >> 5, c4 = e & 31;
            const int p = r / DG_K;
            const long long pt = pt0 + p;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pt < total_pts) {
                const long long cloud0 = (pt / N) * N;
                const int j = idx[pt * DG_K + (r - p * DG_K)];
                const float4 a = *reinterpret_cast<const float4*>(PQ + (cloud0 + j) * ldpq + c4 * 4);
                const float4 q = *reinterpret_cast<const float4*>(PQ + pt * ldpq + DG_C + c4 * 4);
                v.x = leaky(a.x + q.x, slope); v.y = leaky(a.y + q.y, slope);
                v.z = leaky(a.z + q.z, slope); v.w = leaky(a.w + q.w, slope);
            }
            *reinterpret_cast<float4*>(E + r * DG_LDE + c4 * 4) = v;
        }
        __syncthreads();
        // ---- x1 = max over the 20 edges ---------------------------------------------------------
        for (int e = tid; e < DG_PT * DG_C; e += 256) {
            const int p = e >> 7, c = e & 127;
            if (pt0 + p < total_pts) {
                float m = E[(p * DG_K) * DG_LDE + c];
#pragma unroll
                for (int kk = 1; kk < DG_K; ++kk) m = fmaxf(m, E[(p * DG_K + kk) * DG_LDE + c]);
                x1[(pt0 + p) * ld1 + c] = m;
            }
        }
        // ---- e2 = e1 . W2^T : thread = 10 rows x 8 cols ------------------------------------------
        float acc[DG_RT][8];
#pragma unroll
        for (int i = 0; i < DG_RT; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        const float* Er = E + (ty * DG_RT) * DG_LDE;
#pragma unroll 2
        for (int c = 0; c < DG_C; c += 4) {
            float4 a[DG_RT];
#pragma unroll
            for (int i = 0; i < DG_RT; ++i) a[i] = *reinterpret_cast<const float4*>(Er + i * DG_LDE + c);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float4 b0 = *reinterpret_cast<const float4*>(Wt + (c + cc) * DG_LDW + tx * 4);
                const float4 b1 = *reinterpret_cast<const float4*>(Wt + (c + cc) * DG_LDW + 64 + tx * 4);
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < DG_RT; ++i) {
                    const float av = cc == 0 ? a[i].x : cc == 1 ? a[i].y : cc == 2 ? a[i].z : a[i].w;
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
                }
            }
        }
        // ---- max over each thread's 10 rows, then over the two row groups of a point ----------------
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float m = acc[0][j];
#pragma unroll
            for (int i = 1; i < DG_RT; ++i) m = fmaxf(m, acc[i][j]);
            red[ty * DG_C + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4))] = m;
        }
        __syncthreads();
        for (int e = tid; e < DG_PT * DG_C; e += 256) {
            const int p = e >> 7, c = e & 127;
            if (pt0 + p < total_pts) {
                const float m = fmaxf(red[(2 * p) * DG_C + c], red[(2 * p + 1) * DG_C + c]);
                x2[(pt0 + p) * ld2 + c] = leaky(m + b2[c], slope);
            }
        }
    }
}

__global__ void edge_max_kernel(const float* __restrict__ E, int k, long long total_pts, int C, float* __restrict__ out,
                                int ldo) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int C4 = C >> 2;
    if (e >= total_pts * C4) return;
    const long long pt = e / C4;
    const int c4 = (int)(e - pt * C4);
    const float4* r = reinterpret_cast<const float4*>(E + pt * k * C) + c4;
    float4 m = r[0];
    for (int kk = 1; kk < k; ++kk) {
        const float4 t = r[(size_t)kk * C4];
        m.x = fmaxf(m.x, t.x); m.y = fmaxf(m.y, t.y); m.z = fmaxf(m.z, t.z); m.w = fmaxf(m.w, t.w);
    }
    *reinterpret_cast<float4*>(out + pt * ldo + c4 * 4) = m;
}

// out[pt, c] = act( max_k P[cloud0 + idx[pt,k], c] + Q[pt, c] ),  C % 128 == 0: one warp per point
template <int VPL>
__global__ void gather_max_kernel(const float* __restrict__ P, int ldp, const float* __restrict__ Q, int ldq,
                                  const int* __restrict__ idx, int k, int N, long long total_pts, float slope,
                                  float* __restrict__ out, int ldo, __half* __restrict__ op, int ldop, long long op_plane) {
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pt >= total_pts) return;
    const long long cloud0 = (pt / N) * N;
    float4 m[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) m[v] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    const int* ip = idx + pt * k;
    for (int kk = 0; kk < k; ++kk) {
        const int j = ip[kk];
        const float4* r = reinterpret_cast<const float4*>(P + (cloud0 + j) * ldp);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const float4 t = r[lane + v * 32];
            m[v].x = fmaxf(m[v].x, t.x); m[v].y = fmaxf(m[v].y, t.y);
            m[v].z = fmaxf(m[v].z, t.z); m[v].w = fmaxf(m[v].w, t.w);
        }
    }
    const float4* q = reinterpret_cast<const float4*>(Q + pt * ldq);
    float4* o = reinterpret_cast<float4*>(out + pt * ldo);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const float4 qq = q[lane + v * 32];
        float4 r;
        r.x = leaky(m[v].x + qq.x, slope); r.y = leaky(m[v].y + qq.y, slope);
        r.z = leaky(m[v].z + qq.z, slope); r.w = leaky(m[v].w + qq.w, slope);
        o[lane + v * 32] = r;
        if (op != nullptr) {              // the same row in "h3" operand format for the next GEMM
            __half* orow = op + pt * ldop + (lane + v * 32) * 4;
            *reinterpret_cast<uint2*>(orow) = make_uint2(tc::pack_h2(r.x, r.y, 0), tc::pack_h2(r.z, r.w, 0));
            *reinterpret_cast<uint2*>(orow + op_plane) =
                make_uint2(tc::pack_h2(tc::lo_part(r.x, 0), tc::lo_part(r.y, 0), 0),
                           tc::pack_h2(tc::lo_part(r.z, 0), tc::lo_part(r.w, 0), 0));
        }
    }
}

// generic edge tensor builder for the public get_graph_feature(): out[b, c, n, kk] (reference layout
// [B, 2D, N, k], contiguous) = c < D ? x[b, idx[b,n,kk], c] : x[b, n, c-D];  x token-major [B,N,D]
__global__ void graph_feature_kernel(const float* __restrict__ xt, int D, int N, int k,
                                     const int* __restrict__ idx, float* __restrict__ out) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;     // n*k + kk
    if (e >= N * k) return;
    const int n = e / k;
    const float* xb = xt + (size_t)b * N * D;
    float v;
    if (c < D) v = xb[(size_t)idx[((size_t)b * N) * k + e] * D + c];
    else v = xb[(size_t)n * D + (c - D)];
    out[(((size_t)b * 2 * D + c) * N) * k + e] = v;
}

}  // namespace

VCR_API int vcr_edgeconv_dg(const float* PQ, int ldpq, const int* idx, int k, int N, long long total_pts,
                            const float* W2, const float* b2, float slope, float* x1, int ld1, float* x2, int ld2,
                            cudaStream_t stream) {
    VCR_REQUIRE(PQ && idx && W2 && b2 && x1 && x2 && N > 0 && total_pts > 0);
    if (k != DG_K) return VCR_ERR_UNSUPPORTED;
    if (slope < 0.f) return VCR_ERR_UNSUPPORTED;          // monotone-max shortcut needs slope >= 0
    if ((ldpq & 3) || (reinterpret_cast<uintptr_t>(PQ) & 15)) return VCR_ERR_INVALID;
    const size_t smem = (size_t)(DG_C * DG_LDW + DG_ROWS * DG_LDE + 16 * DG_C) * sizeof(float);
    if (cudaFuncSetAttribute(edgeconv_dg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return VCR_ERR_LAUNCH;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long ntiles = (total_pts + DG_PT - 1) / DG_PT;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    edgeconv_dg_kernel<<<grid, 256, smem, stream>>>(PQ, ldpq, idx, N, total_pts, W2, b2, slope, x1, ld1, x2, ld2);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

VCR_API int vcr_gather_max(const float* P, int ldp, const float* Q, int ldq, const int* idx, int k, int N,
                           long long total_pts, int C, float slope, float* out, int ldo, void* op, int ldop,
                           long long op_plane, cudaStream_t stream) {
    VCR_REQUIRE(P && Q && idx && out && k > 0 && N > 0 && total_pts > 0);
    if (op && ((ldop & 3) || (op_plane & 3) || (reinterpret_cast<uintptr_t>(op) & 7))) return VCR_ERR_INVALID;
    __half* oph = reinterpret_cast<__half*>(op);
    if (slope < 0.f || C % 128 != 0 || C > 512) return VCR_ERR_UNSUPPORTED;
    if ((ldp & 3) || (ldq & 3) || (ldo & 3)) return VCR_ERR_INVALID;
    const int wpb = 8;
    const int grid = vcr_cdiv(total_pts, wpb);
    switch (C / 128) {
        case 1: gather_max_kernel<1><<<grid, wpb * 32, 0, stream>>>(P, ldp, Q, ldq, idx, k, N, total_pts, slope, out, ldo, oph, ldop, op_plane); break;
        case 2: gather_max_kernel<2><<<grid, wpb * 32, 0, stream>>>(P, ldp, Q, ldq, idx, k, N, total_pts, slope, out, ldo, oph, ldop, op_plane); break;
        case 3: gather_max_kernel<3><<<grid, wpb * 32, 0, stream>>>(P, ldp, Q, ldq, idx, k, N, total_pts, slope, out, ldo, oph, ldop, op_plane); break;
        case 4: gather_max_kernel<4><<<grid, wpb * 32, 0, stream>>>(P, ldp, Q, ldq, idx, k, N, total_pts, slope, out, ldo, oph, ldop, op_plane); break;
    }
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// out[pt, c] = max_kk E[pt, kk, c]   (x.max(dim=-1) over a materialised per-edge tensor, DGCNN model/vcrnet_model.py:108-118)
VCR_API int vcr_edge_max(const float* E, int k, long long total_pts, int C, float* out, int ldo, cudaStream_t stream) {
    VCR_REQUIRE(E && out && k >= 1 && total_pts > 0 && C > 0 && C % 4 == 0 && ldo % 4 == 0);
    const long long n = total_pts * (C / 4);
    edge_max_kernel<<<vcr_cdiv(n, 256), 256, 0, stream>>>(E, k, total_pts, C, out, ldo);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// x token-major [B,N,D]; idx [B,N,k] int32; out [B,2D,N,k] contiguous (util/util.py:176-199 layout)
VCR_API int vcr_graph_feature(const float* xt, int B, int D, int N, int k, const int* idx, float* out,
                              cudaStream_t stream) {
    VCR_REQUIRE(xt && idx && out && B > 0 && D > 0 && N > 0 && k > 0 && B <= 65535 && 2 * D <= 65535);
    dim3 g(vcr_cdiv((long long)N * k, 256), 2 * D, B);
    graph_feature_kernel<<<g, 256, 0, stream>>>(xt, D, N, k, idx, out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}
