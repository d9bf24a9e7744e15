// FP32 SIMT GEMM with fused epilogue: the exact-precision ("fp32" mode) matrix engine of the
// path.  Every 1x1 conv / nn.Linear of the reference (model/lpdnet_model.py:111-135,
// model/transformer.py:210-224,238) and the score / PV products of attention
// (model/transformer.py:30,55) and of the VCP head (model/vcrnet_model.py:337) go through
// it in fp32 mode; the tcgen05 kernels (gemm_tc.cu) replace it in the tensor-core modes.
//
//   C[z] = epilogue( alpha * A[z] (M x K) * op(B[z]) ),   op(B) = B^T for B stored [N,K] ("NT":
//   weights, keys) or B for B stored [K,N] ("NN": values).
//   epilogue: + bias[n], LeakyReLU(slope) (act=1), + residual[m,n].
//   z = outer * nb_inner + inner selects pointers through two strides per operand, which is how a
//   head of a fused [B,N,3*512] QKV buffer is addressed without any transpose/copy
//   (the reference .contiguous()-copies every head split, model/transformer.py:210-212).
//
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile per thread, register-prefetched
// double buffering.  Roofline: FP32 FMA pipe (not HBM): 2*M*N*K flops, bytes 4(MK+NK+MN).
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int APAD = 4;

struct GemmParams {
    const float* A; const float* B; float* C;
    const float* bias; const float* residual;
    int M, N, K;
    int lda, ldb, ldc, ldr;
    long long sAo, sAi, sBo, sBi, sCo, sCi, sRo, sRi;
    int nb_inner;
    float alpha, slope;
    int act;
};

template <bool B_KN>
__global__ void __launch_bounds__(256)
sgemm_kernel(const GemmParams p) {
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN + APAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int zo = blockIdx.z / p.nb_inner, zi = blockIdx.z - zo * p.nb_inner;
    const float* __restrict__ A = p.A + zo * p.sAo + zi * p.sAi;
    const float* __restrict__ B = p.B + zo * p.sBo + zi * p.sBi;
    float* __restrict__ C = p.C + zo * p.sCo + zi * p.sCi;
    const float* __restrict__ R = p.residual ? p.residual + zo * p.sRo + zi * p.sRi : nullptr;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    // global -> register staging
    const int a_r = tid >> 2, a_k = (tid & 3) * 4;          // A rows a_r, a_r+64 ; k offset a_k
    const int b_k = tid >> 5, b_n = (tid & 31) * 4;          // NN: B rows b_k, b_k+8 ; n offset b_n
    float4 ra[2], rb[2];

    // 4 consecutive K elements of a row; the last group of a ragged K (K % 4 != 0) is read element-wise
    auto ld_k4 = [&](const float* row, int kk) -> float4 {
        if (kk + 3 < p.K) return *reinterpret_cast<const float4*>(row + kk);
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < 4; ++c) if (kk + c < p.K) t[c] = row[kk + c];
        return make_float4(t[0], t[1], t[2], t[3]);
    };
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + a_r + h * 64, kk = k0 + a_k;
            ra[h] = (m < p.M && kk < p.K) ? ld_k4(A + (size_t)m * p.lda, kk) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (!B_KN) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + a_r + h * 64, kk = k0 + a_k;
                rb[h] = (n < p.N && kk < p.K) ? ld_k4(B + (size_t)n * p.ldb, kk) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = k0 + b_k + h * 8, n = n0 + b_n;
                if (kk < p.K && n + 3 < p.N) {
                    rb[h] = *reinterpret_cast<const float4*>(B + (size_t)kk * p.ldb + n);
                } else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    if (kk < p.K)
                        for (int c = 0; c < 4; ++c) if (n + c < p.N) t[c] = B[(size_t)kk * p.ldb + n + c];
                    rb[h] = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
        }
    };
    auto store_tiles = [&](int s) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = a_r + h * 64;
            As[s][a_k + 0][r] = ra[h].x; As[s][a_k + 1][r] = ra[h].y;
            As[s][a_k + 2][r] = ra[h].z; As[s][a_k + 3][r] = ra[h].w;
        }
        if (!B_KN) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = a_r + h * 64;
                Bs[s][a_k + 0][r] = rb[h].x; Bs[s][a_k + 1][r] = rb[h].y;
                Bs[s][a_k + 2][r] = rb[h].z; Bs[s][a_k + 3][r] = rb[h].w;
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                *reinterpret_cast<float4*>(&Bs[s][b_k + h * 8][b_n]) = rb[h];
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (p.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[s][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[s][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[s][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[s][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(s ^ 1);
            __syncthreads();
        }
    }

    // epilogue
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                        (!R || (((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0)));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            if (n >= p.N) continue;
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float t = p.alpha * acc[i][jh * 4 + c];
                if (p.bias && n + c < p.N) t += p.bias[n + c];
                if (p.act == 1) t = leaky(t, p.slope);
                v[c] = t;
            }
            if (vec_ok && n + 3 < p.N) {
                if (R) {
                    const float4 r = *reinterpret_cast<const float4*>(R + (size_t)m * p.ldr + n);
                    v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
                }
                *reinterpret_cast<float4*>(C + (size_t)m * p.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
                for (int c = 0; c < 4; ++c) {
                    if (n + c < p.N) {
                        float t = v[c];
                        if (R) t += R[(size_t)m * p.ldr + n + c];
                        C[(size_t)m * p.ldc + n + c] = t;
                    }
                }
            }
        }
    }
}

}  // namespace

// Generic batched GEMM.  b_layout: 0 = B stored [N,K] (ldb >= K), 1 = B stored [K,N] (ldb >= N).
// Requirements: lda % 4 == 0, A/B 16-byte aligned and ldb % 4 == 0 (any K: a ragged tail is read element-wise);
// batch strides are in elements; nb_outer * nb_inner <= 65535.  act: 0 none, 1 LeakyReLU(slope).
VCR_API int vcr_gemm_f32(const float* A, int lda, long long sAo, long long sAi,
                         const float* B, int ldb, long long sBo, long long sBi, int b_layout,
                         float* C, int ldc, long long sCo, long long sCi,
                         const float* bias, const float* residual, int ldr, long long sRo, long long sRi,
                         int M, int N, int K, int nb_outer, int nb_inner,
                         float alpha, int act, float slope, cudaStream_t stream) {
    VCR_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && nb_outer > 0 && nb_inner > 0);
    if ((lda & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
        return VCR_ERR_INVALID;
    if ((sAo & 3) || (sAi & 3) || (sBo & 3) || (sBi & 3)) return VCR_ERR_INVALID;
    if (ldb & 3) return VCR_ERR_INVALID;
    if ((long long)nb_outer * nb_inner > 65535) return VCR_ERR_UNSUPPORTED;
    GemmParams p;
    p.A = A; p.B = B; p.C = C; p.bias = bias; p.residual = residual;
    p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldr = ldr;
    p.sAo = sAo; p.sAi = sAi; p.sBo = sBo; p.sBi = sBi; p.sCo = sCo; p.sCi = sCi; p.sRo = sRo; p.sRi = sRi;
    p.nb_inner = nb_inner; p.alpha = alpha; p.slope = slope; p.act = act;
    dim3 grid(vcr_cdiv(N, BN), vcr_cdiv(M, BM), nb_outer * nb_inner);
    if (b_layout == 0) sgemm_kernel<false><<<grid, 256, 0, stream>>>(p);
    else sgemm_kernel<true><<<grid, 256, 0, stream>>>(p);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}
