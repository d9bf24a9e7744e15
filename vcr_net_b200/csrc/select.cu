// Top-K selection with the canonical order (value descending, ties -> lower index) used by the
// partial-overlap path: key subset of the cross attention (reference model/transformer.py:39-47),
// selectCom (model/vcrnet_model.py:222-223, 244-245) and getCopair (:309).  torch.topk leaves tie
// order unspecified; the oracle (oracle/vcr_oracle.py topk_desc) fixes the same rule.
// One CTA per batch row: bitonic sort of 64-bit (order-preserving float key, index) in shared memory.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned int float_desc_key(float v) {
    if (v == 0.f) v = 0.f;                               // -0.0 and +0.0 compare equal: one key
    unsigned int u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // ascending order-preserving
    return ~u;                                           // descending
}

__global__ void topk_sort_kernel(const float* __restrict__ vals, int n, int npad, int K,
                                 int* __restrict__ idx_out, uint8_t* __restrict__ mask_out) {
    extern __shared__ unsigned long long keys[];
    const int b = blockIdx.x;
    const float* v = vals + (size_t)b * n;
    for (int i = threadIdx.x; i < npad; i += blockDim.x)
        keys[i] = i < n ? (((unsigned long long)float_desc_key(v[i]) << 32) | (unsigned int)i) : ~0ull;
    __syncthreads();
    for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < npad / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = keys[lo], c = keys[hi];
                if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    if (mask_out)
        for (int i = threadIdx.x; i < n; i += blockDim.x) mask_out[(size_t)b * n + i] = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < K; r += blockDim.x) {
        const int i = (int)(keys[r] & 0xffffffffu);
        if (idx_out) idx_out[(size_t)b * K + r] = i;
        if (mask_out) mask_out[(size_t)b * n + i] = 1;
    }
}

}  // namespace

// vals [B,n] -> idx_out [B,K] (sorted, optional) and/or mask_out [B,n] uint8 (1 = selected, optional)
VCR_API int vcr_topk_select(const float* vals, int B, int n, int K, int* idx_out, uint8_t* mask_out,
                            cudaStream_t stream) {
    VCR_REQUIRE(vals && (idx_out || mask_out) && B > 0 && n > 0 && K > 0 && K <= n);
    int npad = 2;
    while (npad < n) npad <<= 1;
    const size_t smem = (size_t)npad * sizeof(unsigned long long);
    if (smem > 224 * 1024) return VCR_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return VCR_ERR_LAUNCH;
    const int threads = npad / 2 < 1024 ? (npad / 2 < 32 ? 32 : npad / 2) : 1024;
    topk_sort_kernel<<<B, threads, smem, stream>>>(vals, n, npad, K, idx_out, mask_out);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}
