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

// npad <= 1024: one key per thread.  Compare-exchange partners closer than a warp are reached with shuffles, the 15 stages
// with stride >= 32 go through double-buffered shared memory (one barrier each): 15 barriers instead of 55.  Same network,
// same result as topk_sort_kernel.
__global__ void __launch_bounds__(1024)
topk_sort_warp_kernel(const float* __restrict__ vals, int n, int npad, int K, int* __restrict__ idx_out,
                      uint8_t* __restrict__ mask_out) {
    __shared__ unsigned long long xch[2][1024];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* v = vals + (size_t)b * n;
    unsigned long long key = tid < n ? (((unsigned long long)float_desc_key(v[tid]) << 32) | (unsigned int)tid) : ~0ull;
    int buf = 0;
    for (int size = 2; size <= npad; size <<= 1) {
        const bool up = (tid & size) == 0;
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            unsigned long long other;
            if (stride >= 32) {
                xch[buf][tid] = key;
                __syncthreads();
                other = xch[buf][tid ^ stride];
                buf ^= 1;                                  // the next smem stage writes the other buffer: no second barrier
            } else {
                other = __shfl_xor_sync(0xffffffffu, key, stride);
            }
            const bool lower = (tid & stride) == 0;
            const bool take_min = lower == up;
            key = take_min ? (key < other ? key : other) : (key > other ? key : other);
        }
    }
    if (mask_out)
        for (int i = tid; i < n; i += blockDim.x) mask_out[(size_t)b * n + i] = 0;
    __syncthreads();
    if (tid < K) {
        const int i = (int)(key & 0xffffffffu);
        if (idx_out) idx_out[(size_t)b * K + tid] = i;
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
    if (npad >= 32 && npad <= 1024) {
        topk_sort_warp_kernel<<<B, npad, 0, stream>>>(vals, n, npad, K, idx_out, mask_out);
        VCR_CHECK_LAUNCH();
        return VCR_OK;
    }
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
