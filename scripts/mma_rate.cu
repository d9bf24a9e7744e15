// GPU diagnostic (not part of the library): tcgen05.mma issue rates for the shapes of the flash attention kernel -- is the
// 3-term tile (24 x QK^T 128x64x16 + 12 x PV 128x128x16, every operand from shared memory) paced by the tensor pipe or by the
// shared-memory operand reads?  Also checks the layout of an A operand held in TMEM (tcgen05.mma [d], [a], b-desc) against a
// host product, and times the tile with P taken from TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I vcr_net_b200/csrc -o gpurun_out/mma_rate scripts/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

namespace {

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

constexpr int Q_TILE = 128 * 128, K_TILE = 64 * 128, V_TILE = 128 * 128, P_TILE = 128 * 128;
constexpr int OFF_Q = 0, OFF_K = 4 * Q_TILE, OFF_V = OFF_K + 4 * K_TILE * 2, OFF_P = OFF_V + 2 * V_TILE, SMEM = OFF_P + 2 * P_TILE + 1024;

// mode: 0 QK N=64 SS x24 | 1 PV N=128 SS x12 | 2 N=256 SS x12 | 3 PV TS x12 | 4 tile SS (24+12) | 5 tile, PV TS | 6 QK N=128 x24
//       7 / 8 per 128 keys: 24 x QK N=128 + 2 x 12 PV (SS / P in TMEM) | 9 QK N=64, Q in TMEM | 10 tile, Q and P in TMEM
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int reps, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (SMEM - 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 1 && tc::elect_one()) {
        const uint32_t q = tc::smem_u32(smem + OFF_Q), k = tc::smem_u32(smem + OFF_K), v = tc::smem_u32(smem + OFF_V),
                       pp = tc::smem_u32(smem + OFF_P);
        const uint32_t i64 = tc::umma_idesc(128, 64, 0), i128 = tc::umma_idesc(128, 128, 0), i256 = tc::umma_idesc(128, 256, 0);
        auto qk = [&](int st, uint32_t idesc, int nbytes) {          // 24 products into S buffer st
            const uint32_t d = tb + st * 64;
            for (int kb = 0; kb < 2; ++kb)
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t qh = tc::umma_desc_k_sw128(q + (kb * 2) * Q_TILE) + kk * 2, ql = tc::umma_desc_k_sw128(q + (kb * 2 + 1) * Q_TILE) + kk * 2;
                    const uint64_t kh = tc::umma_desc_k_sw128(k + st * 4 * nbytes + (kb * 2) * nbytes) + kk * 2,
                                   kl = tc::umma_desc_k_sw128(k + st * 4 * nbytes + (kb * 2 + 1) * nbytes) + kk * 2;
                    tc::umma_f16(d, qh, kl, idesc, (kb | kk) != 0);
                    tc::umma_f16(d, ql, kh, idesc, 1);
                    tc::umma_f16(d, qh, kh, idesc, 1);
                }
        };
        auto qk_ts = [&](int st) {                                     // Q (hi kb0 | hi kb1 | lo kb0 | lo kb1, 32 columns each) in TMEM
            const uint32_t d = tb + st * 64, qt = tb + 384;
            for (int kb = 0; kb < 2; ++kb)
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t kh = tc::umma_desc_k_sw128(k + st * 4 * K_TILE + (kb * 2) * K_TILE) + kk * 2,
                                   kl = tc::umma_desc_k_sw128(k + st * 4 * K_TILE + (kb * 2 + 1) * K_TILE) + kk * 2;
                    umma_f16_ts(d, qt + kb * 32 + kk * 8, kl, i64, (kb | kk) != 0);
                    umma_f16_ts(d, qt + 64 + kb * 32 + kk * 8, kh, i64, 1);
                }
            for (int kb = 0; kb < 2; ++kb)
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t kh = tc::umma_desc_k_sw128(k + st * 4 * K_TILE + (kb * 2) * K_TILE) + kk * 2;
                    umma_f16_ts(d, qt + kb * 32 + kk * 8, kh, i64, 1);
                }
        };
        auto pv_alias = [&](int st) {                                  // P (hi 32 | lo 32 columns) in the S buffer's columns
            const uint32_t o0 = tb + 128, o1 = tb + 256, pb = tb + st * 64;
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t vh = tc::umma_desc_k_sw128(v) + kk * 2, vl = tc::umma_desc_k_sw128(v + V_TILE) + kk * 2;
                umma_f16_ts(o0, pb + kk * 8, vh, i128, 1); umma_f16_ts(o1, pb + kk * 8, vl, i128, 1); umma_f16_ts(o1, pb + 32 + kk * 8, vh, i128, 1);
            }
        };
        auto pv = [&](bool ts, uint32_t idesc) {
            const uint32_t o0 = tb + 128, o1 = tb + 256;
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t vh = tc::umma_desc_k_sw128(v) + kk * 2, vl = tc::umma_desc_k_sw128(v + V_TILE) + kk * 2;
                if (ts) {
                    const uint32_t ph = tb + 384 + kk * 8, pl = tb + 384 + 32 + kk * 8;
                    umma_f16_ts(o0, ph, vh, idesc, 1); umma_f16_ts(o1, ph, vl, idesc, 1); umma_f16_ts(o1, pl, vh, idesc, 1);
                } else {
                    const uint64_t ph = tc::umma_desc_k_sw128(pp) + kk * 2, pl = tc::umma_desc_k_sw128(pp + P_TILE) + kk * 2;
                    tc::umma_f16(o0, ph, vh, idesc, 1); tc::umma_f16(o1, ph, vl, idesc, 1); tc::umma_f16(o1, pl, vh, idesc, 1);
                }
            }
        };
        long long t0 = 0;
        unsigned long long g0 = 0, g1;
        for (int r = -8; r < reps; ++r) {                             // 8 warm-up rounds
            if (r == 0) {
                tc::umma_commit(&bar); tc::mbar_wait(&bar, 0);
                t0 = clock64();
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
            }
            switch (mode) {
                case 0: qk(r & 1, i64, K_TILE); break;
                case 1: pv(false, i128); break;
                case 2: pv(false, i256); break;                       // (reads past the V tile: timing only)
                case 3: pv(true, i128); break;
                case 4: qk(r & 1, i64, K_TILE); pv(false, i128); break;
                case 5: qk(r & 1, i64, K_TILE); pv(true, i128); break;
                case 6: qk(0, i128, 2 * K_TILE); break;
                case 7: qk(0, i128, 2 * K_TILE); pv(false, i128); pv(false, i128); break;
                case 9: qk_ts(r & 1); break;
                case 10: qk_ts(r & 1); pv_alias((r & 1) ^ 1); break;
                default: qk(0, i128, 2 * K_TILE); pv(true, i128); pv(true, i128); break;
            }
        }
        tc::umma_commit(&bar); tc::mbar_wait(&bar, 1);
        cycles[blockIdx.x] = clock64() - t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        cycles[148 + blockIdx.x] = (long long)(g1 - g0);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 512);
}

// D[128 x 64] = A[128 x 16] (TMEM, fp16 pairs packed in 32-bit columns) * B[64 x 16]^T (smem, K-major SW128)
__global__ void __launch_bounds__(128, 1) ts_layout_kernel(float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * 16; e += 128) {
        const int n = e / 16, kq = e % 16;
        const float val = (float)((n + 2 * kq) % 5 - 2);
        *reinterpret_cast<__half*>(smem + n * 128 + (((kq / 8) ^ (n & 7)) * 16) + (kq % 8) * 2) = __float2half(val);
    }
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) { tc::tmem_alloc(&slot, 128); tc::tmem_relinquish(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tb = slot;
    const int m = threadIdx.x;
    uint32_t a[8];
    for (int c = 0; c < 8; ++c) a[c] = tc::pack_h2((float)((m * 3 + 2 * c) % 7 - 3), (float)((m * 3 + 2 * c + 1) % 7 - 3), 0);
    tmem_st_32x8(tb + 64 + ((uint32_t)(warp * 32) << 16), a);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1 && tc::elect_one()) {
        tc::tc_fence_after();
        umma_f16_ts(tb, tb + 64, tc::umma_desc_k_sw128(tc::smem_u32(smem)), tc::umma_idesc(128, 64, 0), 0);
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    uint32_t r[32];
    for (int h = 0; h < 2; ++h) {
        tc::tmem_ld_32x32(tb + ((uint32_t)(warp * 32) << 16) + h * 32, r);
        tc::tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[m * 64 + h * 32 + i] = __uint_as_float(r[i]);
    }
    (void)lane;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 128);
}

// TMEM port: tcgen05.ld streams from 4 or 8 warps (32x32b.x32, 64 KB of fp32 per round of 128 lanes x 128 columns) alone and
// while the tensor core runs 128x128x16 / 128x64x16 products -- do accumulator reads and MMAs overlap?
// what: 0 = MMA only, 1 = LDTM only, 2 = both.  ldw = number of loading warps (4 or 8).  ts = A operand from TMEM.
__global__ void __launch_bounds__(384, 1) port_kernel(int what, int ldw, int ts, int n64, int reps, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (SMEM - 1024) / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 1 && what != 1 && tc::elect_one()) {
        const uint32_t q = tc::smem_u32(smem + OFF_Q), k = tc::smem_u32(smem + OFF_K);
        const uint32_t idesc = n64 ? tc::umma_idesc(128, 64, 0) : tc::umma_idesc(128, 128, 0);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int i = 0; i < 24; ++i) {
                const uint64_t kd = tc::umma_desc_k_sw128(k + (i & 3) * 2 * K_TILE) + (i & 3) * 2;
                if (ts) umma_f16_ts(tb + (r & 1) * 128, tb + 384 + (i & 7) * 8, kd, idesc, i != 0);
                else tc::umma_f16(tb + (r & 1) * 128, tc::umma_desc_k_sw128(q + (i & 3) * Q_TILE) + (i & 3) * 2, kd, idesc, i != 0);
            }
        tc::umma_commit(&bar); tc::mbar_wait(&bar, 0);
        cycles[blockIdx.x * 2] = clock64() - t0;
    }
    if (warp >= 4 && warp < 4 + ldw && what != 0) {
        const uint32_t lane_adr = (uint32_t)((warp & 3) * 32) << 16;
        const int half = (warp - 4) >> 2;
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int c = 0; c < (ldw == 8 ? 2 : 4); ++c) {       // one round = 128 lanes x 128 columns across the loading warps
                uint32_t v[32];
                tc::tmem_ld_32x32(tb + 256 + lane_adr + (ldw == 8 ? half * 64 : 0) + c * 32, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc ^= v[i];
            }
        if (acc == 0x12345678u) cycles[1000] = acc;
        if ((threadIdx.x & 31) == 0 && warp == 4) cycles[blockIdx.x * 2 + 1] = clock64() - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 512);
}

}  // namespace


int main() {
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    // ---- A-in-TMEM layout ----
    float* d_out;
    cudaMalloc(&d_out, 128 * 64 * 4);
    ts_layout_kernel<<<1, 128, 64 * 128 + 1024>>>(d_out);
    std::vector<float> h(128 * 64);
    cudaError_t e = cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("ts_layout_kernel: %s\n", cudaGetErrorString(e)); return 1; }
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            float want = 0.f;
            for (int k = 0; k < 16; ++k) want += (float)((m * 3 + k) % 7 - 3) * (float)((n + 2 * k) % 5 - 2);
            if (h[m * 64 + n] != want && bad++ < 5) printf("  D[%d][%d] = %g, want %g\n", m, n, h[m * 64 + n], want);
        }
    printf("A operand in TMEM (lane = row, fp16 pair k = 2c, 2c+1 in column c): %s (%d mismatches)\n", bad ? "MISMATCH" : "exact", bad);
    // ---- rates ----
    long long* d_cyc;
    cudaMalloc(&d_cyc, 2 * 148 * 8);
    const char* names[] = {"24 x QK 128x64x16 SS", "12 x PV 128x128x16 SS", "12 x 128x256x16 SS", "12 x PV 128x128x16, A in TMEM",
                           "tile: 24 QK(N=64) + 12 PV, all SS", "tile: 24 QK(N=64) SS + 12 PV A-in-TMEM", "24 x QK 128x128x16 SS",
                           "128 keys: 24 QK(N=128) + 24 PV, all SS", "128 keys: 24 QK(N=128) SS + 24 PV A-in-TMEM",
                           "24 x QK 128x64x16, Q in TMEM", "tile: 24 QK(N=64) + 12 PV, Q and P in TMEM (P over S)"};
    const int floors[] = {24 * 32, 12 * 64, 12 * 128, 12 * 64, 24 * 32 + 12 * 64, 24 * 32 + 12 * 64, 24 * 64, 24 * 64 + 24 * 64, 24 * 64 + 24 * 64, 24 * 32, 24 * 32 + 12 * 64};
    for (int grid : {1, 148})
        for (int mode = 0; mode < 11; ++mode) {
            const int reps = grid == 1 ? 400 : 40000;                 // the full-chip runs are long enough (tens of ms) for the power cap to act
            rate_kernel<<<grid, 128, SMEM>>>(mode, reps, d_cyc);
            std::vector<long long> c(296);
            e = cudaMemcpy(c.data(), d_cyc, 296 * 8, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("rate_kernel mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            double s = 0, ns = 0;
            for (int b = 0; b < grid; ++b) { s += (double)c[b]; ns += (double)c[148 + b]; }
            printf("grid %3d  %-46s %8.1f cycles   floor %5d   x%.2f   SM clock %.2f GHz\n", grid, names[mode], s / (grid * (double)reps), floors[mode],
                   s / (grid * (double)reps) / floors[mode], s / ns);
        }
    cudaFuncSetAttribute(port_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    long long* d_c2;
    cudaMalloc(&d_c2, 2048 * 8);
    printf("TMEM port: 24 products per round (floor 1536 / 768 cycles) and one 64 KB accumulator read per round\n");
    for (int n64 = 0; n64 < 2; ++n64)
        for (int ts = 0; ts < 2; ++ts)
            for (int ldw : {4, 8})
                for (int what = 0; what < 3; ++what) {
                    const int reps = 200;
                    cudaMemset(d_c2, 0, 2048 * 8);
                    port_kernel<<<148, 384, SMEM>>>(what, ldw, ts, n64, reps, d_c2);
                    std::vector<long long> c(296);
                    e = cudaMemcpy(c.data(), d_c2, 296 * 8, cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) { printf("port_kernel: %s\n", cudaGetErrorString(e)); return 1; }
                    double m = 0, l = 0;
                    for (int b = 0; b < 148; ++b) { m += (double)c[2 * b]; l += (double)c[2 * b + 1]; }
                    printf("  N=%3d A in %s, %d loading warps, %-9s  MMA %8.1f cycles / round   LDTM %8.1f cycles / round\n", n64 ? 64 : 128,
                           ts ? "TMEM" : "smem", ldw, what == 0 ? "MMA only" : what == 1 ? "LDTM only" : "both", m / 148 / reps, l / 148 / reps);
                }
    return 0;
}
