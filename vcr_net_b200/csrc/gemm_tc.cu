// tcgen05 / TMEM / TMA GEMM: the tensor-core matrix engine of the path (sm_100a only).
//
//   C[z] = epilogue( alpha * A[z] (M x K) * B[z]^T (N x K) ),  both operands K-major 16-bit,
//   fed by TMA (128-byte swizzle) into a multi-stage shared-memory ring, multiplied by
//   tcgen05.mma (cta_group::1, 128 x 128 x 16 per instruction) into TMEM accumulators that are
//   double-buffered so the epilogue of tile i overlaps the main loop of tile i+1 (persistent CTAs,
//   one per SM, warp-specialised: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
//   warps 4-7 epilogue).
//
// Precision modes:
//   NTERMS = 3, fp16 "split" operands ("h3", the fp32-parity mode): every fp32 value x is carried as
//       hi = fp16(x), lo = fp16((x - hi) * 2^11)  (two planes of the operand buffer), and
//       x*y ~= hi*hi' + 2^-11 (hi*lo' + lo*hi') with the main term and the correction term in SEPARATE
//       TMEM accumulators (D0, D1), combined in fp32 in the epilogue: ~2^-22 relative error per product,
//       3 tensor-core passes at the fp16 rate (a tf32 x3 scheme would cost 3 passes at HALF that rate).
//   NTERMS = 1: plain fp16 or bf16 operands (throughput modes, reported separately).
//
// Epilogue (fused, straight from TMEM): alpha, + bias[n], LeakyReLU, + residual, then any of
//   - fp32 row-major C,
//   - operand-format ("H2": [plane][row][ld] fp16 hi/lo) row-major output for columns < h_split,
//   - operand-format TRANSPOSED output for columns >= h_split (V^T for attention: with one TMEM lane
//     = one row per thread the transposed store is the coalesced one).
// so consecutive GEMMs never round-trip through a separate conversion kernel.
#include <atomic>
#include "tc_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int TILE_BYTES = 128 * 128;       // 128 rows x 128 B
constexpr int EPI_WARPS = 8;                // two per TMEM lane quarter (column halves)
constexpr int NTHREADS = 128 + EPI_WARPS * 32;

struct TcGemmParams {
    int M, N, K;
    int nb_outer, nb_inner;
    long long a_row_o, a_row_i; int a_col_o, a_col_i;
    long long b_row_o, b_row_i; int b_col_o, b_col_i;
    float alpha; const float* bias; int act; float slope;
    float* C; int ldc; long long c_so, c_si;
    const float* R; int ldr; long long r_so, r_si;
    __half* H; int ldh; long long h_plane, h_so, h_si; int h_split;
    __half* HT; int ldt; long long t_plane, t_so, t_si;
    int out_planes;      // 2: write hi and lo planes, 1: hi only
    int out_bf16;        // operand-format outputs are bf16 instead of fp16 (bf16 throughput mode)
};

// CTAS = 2: CTA pairs (cta_group::2, 256 x 128 tile per pair): a CTA stages its 128 A rows and HALF of the B tile
// (64 rows) per k-block -- 48 KB instead of 64 KB in the 3-term mode, so 4 stages fit and the L2 -> SM operand traffic
// per MMA drops by a quarter.
template <int NTERMS, int CTAS = 1>
struct Cfg {
    static constexpr int PLANES = NTERMS == 3 ? 2 : 1;
    static constexpr int B_TILE_BYTES = TILE_BYTES / CTAS;
    static constexpr int STAGE_BYTES = PLANES * (TILE_BYTES + B_TILE_BYTES);
    static constexpr int NSTAGES = CTAS == 2 ? 4 : (NTERMS == 3 ? 3 : 6);
    static constexpr int ACC_COLS = NTERMS == 3 ? 2 * BN : BN;      // D0 | D1
    static constexpr int TMEM_COLS = 2 * ACC_COLS;                  // double-buffered
    static constexpr int SMEM_BYTES = NSTAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                                      EPI_WARPS * 32 * 32 * 4 /*epilogue transpose staging*/;
};

using tc::pack_h2;
using tc::lo_part;

// Row-major outputs of one 32 x 32 chunk whose rows all exist and whose accesses are 16-byte aligned.  Every run-time switch
// of the epilogue is a bit of V (1: fp32 C, 2: operand-format output, 4: residual, 8: LeakyReLU, 16: lo plane), so the
// 8-row loop is straight-line code: the generic loop spent ~900 issue slots per chunk on predicates and branches, which
// made the K = 512 GEMMs epilogue-bound (two epilogue warps per scheduler).  Same arithmetic, same order.
template <int V, int OBF>
__device__ __forceinline__ void store_rows_full(const float* stg, const float4 b4, const float alpha, const float slope,
                                                const float4 (&rq)[8], float* c0, const int ldc, __half* h0,
                                                const int ldh, const long long h_plane, const int rsub, const int c4) {
    constexpr bool HAS_C = (V & 1) != 0, HAS_H = (V & 2) != 0, HAS_R = (V & 4) != 0, ACT = (V & 8) != 0, PL2 = (V & 16) != 0;
#pragma unroll
    for (int itr = 0; itr < 8; ++itr) {
        const int rl = itr * 4 + rsub;
        float4 x = *reinterpret_cast<const float4*>(stg + rl * 32 + ((c4 ^ (rl & 7)) << 2));
        x.x = fmaf(alpha, x.x, b4.x); x.y = fmaf(alpha, x.y, b4.y);
        x.z = fmaf(alpha, x.z, b4.z); x.w = fmaf(alpha, x.w, b4.w);
        if (ACT) { x.x = leaky(x.x, slope); x.y = leaky(x.y, slope); x.z = leaky(x.z, slope); x.w = leaky(x.w, slope); }
        if (HAS_R) { x.x += rq[itr].x; x.y += rq[itr].y; x.z += rq[itr].z; x.w += rq[itr].w; }
        if (HAS_C) *reinterpret_cast<float4*>(c0 + (size_t)(itr * 4) * ldc) = x;
        if (HAS_H) {
            __half* hh = h0 + (size_t)(itr * 4) * ldh;
            *reinterpret_cast<uint2*>(hh) = make_uint2(pack_h2(x.x, x.y, OBF), pack_h2(x.z, x.w, OBF));
            if (PL2)
                *reinterpret_cast<uint2*>(hh + h_plane) =
                    make_uint2(pack_h2(lo_part(x.x, OBF), lo_part(x.y, OBF), OBF),
                               pack_h2(lo_part(x.z, OBF), lo_part(x.w, OBF), OBF));
        }
    }
}

// MC = 2 (with CTAS = 2): a cluster is TWO pairs working on the same 256 rows and adjacent 128-column tiles.  The A tile is
// the same for both pairs, so each of the four CTAs fetches one 64-row quarter of the 256 x 64 A block and MULTICASTS it to
// its counterpart in the other pair: 32 KB instead of 48 KB of L2 -> SM traffic per CTA and k-block.  The main loop is
// bound by the L2 slice throughput (~6300 B/clk chip-wide = 42 B/clk/SM; 64 KB per 768 MMA-clocks = 83 B/clk wanted by the
// single-CTA kernel, 62 by a pair, 42 by this one).  A stage is refilled only after BOTH pairs have consumed it.
template <int NTERMS, int FMT, int CTAS, int MC>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcGemmParams p) {
    using C_ = Cfg<NTERMS, CTAS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // stays in the shared window
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_::NSTAGES * C_::STAGE_BYTES);
    uint64_t* full = bars;                         // [NSTAGES]
    uint64_t* empty = bars + C_::NSTAGES;          // [NSTAGES]
    uint64_t* tfull = bars + 2 * C_::NSTAGES;      // [2]
    uint64_t* tempty = tfull + 2;                  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* stage_buf = reinterpret_cast<float*>(smem + C_::NSTAGES * C_::STAGE_BYTES + 256);   // [EPI_WARPS][32][32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a scheduling unit is a 128-row tile (CTAS = 1) or a 256-row tile pair (CTAS = 2: this CTA takes rows rank*128..)
    const int tiles_m = (p.M + CTAS * BM - 1) / (CTAS * BM), tiles_n = ((p.N + BN - 1) / BN + MC - 1) / MC;
    const int tiles_per_z = tiles_m * tiles_n;
    const long long total_tiles = (long long)tiles_per_z * p.nb_outer * p.nb_inner;
    const int nkb = (p.K + BK - 1) / BK;
    const int crank = CTAS == 2 ? (int)tc::cluster_ctarank() : 0;
    const int rank = crank & 1, pairq = crank >> 1;          // rank within the pair, pair within the cluster
    const int unit0 = blockIdx.x / (CTAS * MC), unit_step = gridDim.x / (CTAS * MC);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C_::NSTAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], MC); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], EPI_WARPS * 32 * CTAS); }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        if (CTAS == 2) { tc::tmem_alloc_2sm(tmem_slot, C_::TMEM_COLS); tc::tmem_relinquish_2sm(); }
        else { tc::tmem_alloc(tmem_slot, C_::TMEM_COLS); tc::tmem_relinquish(); }
    }
    tc::tc_fence_before();
    if (CTAS == 2) tc::cluster_sync_all();      // the peer's barriers exist before anything signals them
    else __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (long long t = unit0; t < total_tiles; t += unit_step) {
                const int z = (int)(t / tiles_per_z), r = (int)(t - (long long)z * tiles_per_z);
                const int m_blk = (r / tiles_n) * CTAS + rank, n_blk = (r % tiles_n) * MC + pairq;
                const int zo = z / p.nb_inner, zi = z - zo * p.nb_inner;
                const int a_row = (int)(zo * p.a_row_o + zi * p.a_row_i) + m_blk * BM;
                const int b_row = (int)(zo * p.b_row_o + zi * p.b_row_i) + n_blk * BN + rank * (BN / CTAS);
                const int a_col = zo * p.a_col_o + zi * p.a_col_i;
                const int b_col = zo * p.b_col_o + zi * p.b_col_i;
                for (int kb = 0; kb < nkb; ++kb) {
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * C_::STAGE_BYTES;
                    if (CTAS == 2) {
                        // the leader's barrier collects the bytes of both CTAs; only the leader arms it
                        if (rank == 0) tc::mbar_expect_tx(&full[s], 2 * C_::STAGE_BYTES);
#pragma unroll
                        for (int pl = 0; pl < C_::PLANES; ++pl) {
                            if (MC == 2)     // this CTA's 64-row half of the A tile, to itself and its counterpart in the other pair
                                tc::tma_load_3d_2sm_mc(st + pl * TILE_BYTES + pairq * (TILE_BYTES / 2), &tmA, &full[s],
                                                       a_col + kb * BK, a_row + pairq * (BM / 2), pl, (uint16_t)(5u << rank));
                            else
                                tc::tma_load_3d_2sm(st + pl * TILE_BYTES, &tmA, &full[s], a_col + kb * BK, a_row, pl);
                            tc::tma_load_3d_2sm(st + C_::PLANES * TILE_BYTES + pl * C_::B_TILE_BYTES, &tmB, &full[s],
                                                b_col + kb * BK, b_row, pl);
                        }
                    } else {
                        tc::mbar_expect_tx(&full[s], C_::STAGE_BYTES);
#pragma unroll
                        for (int pl = 0; pl < C_::PLANES; ++pl) {
                            tc::tma_load_3d(st + pl * TILE_BYTES, &tmA, &full[s], a_col + kb * BK, a_row, pl);
                            tc::tma_load_3d(st + (C_::PLANES + pl) * TILE_BYTES, &tmB, &full[s], b_col + kb * BK, b_row, pl);
                        }
                    }
                    if (++s == C_::NSTAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (the leader CTA of a pair) =====================
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::umma_idesc(CTAS * BM, BN, FMT);
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (long long t = unit0; t < total_tiles; t += unit_step, ++it) {
                const int a = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                tc::mbar_wait(&tempty[a], aph ^ 1);
                tc::tc_fence_after();
                const uint32_t d0 = tmem_base + a * C_::ACC_COLS;
                const uint32_t d1 = d0 + BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    tc::mbar_wait(&full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t st = tc::smem_u32(smem + s * C_::STAGE_BYTES);
                    const uint64_t a_hi = tc::umma_desc_k_sw128(st);
                    const uint64_t b_hi = tc::umma_desc_k_sw128(st + C_::PLANES * TILE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        const uint32_t acc = (kb | kk) != 0;
                        const uint64_t adv = (uint64_t)(kk * 32 >> 4);      // 16 elements = 32 B along K
                        if (CTAS == 2) {
                            tc::umma_f16_2sm(d0, a_hi + adv, b_hi + adv, idesc, acc);
                            if (NTERMS == 3) {
                                const uint64_t a_lo = tc::umma_desc_k_sw128(st + TILE_BYTES);
                                const uint64_t b_lo = tc::umma_desc_k_sw128(st + 2 * TILE_BYTES + C_::B_TILE_BYTES);
                                tc::umma_f16_2sm(d1, a_hi + adv, b_lo + adv, idesc, acc);
                                tc::umma_f16_2sm(d1, a_lo + adv, b_hi + adv, idesc, 1);
                            }
                        } else {
                            tc::umma_f16(d0, a_hi + adv, b_hi + adv, idesc, acc);
                            if (NTERMS == 3) {
                                const uint64_t a_lo = tc::umma_desc_k_sw128(st + TILE_BYTES);
                                const uint64_t b_lo = tc::umma_desc_k_sw128(st + 3 * TILE_BYTES);
                                tc::umma_f16(d1, a_hi + adv, b_lo + adv, idesc, acc);
                                tc::umma_f16(d1, a_lo + adv, b_hi + adv, idesc, 1);
                            }
                        }
                    }
                    if (CTAS == 2) tc::umma_commit_2sm(&empty[s], MC == 2 ? 15 : 3);   // frees the stage in every CTA that fills it
                    else tc::umma_commit(&empty[s]);               // smem slot free once these MMAs retire
                    if (++s == C_::NSTAGES) { s = 0; ph ^= 1; }
                }
                if (CTAS == 2) tc::umma_commit_2sm(&tfull[a], (uint16_t)(3u << (2 * pairq)));   // both CTAs' epilogues
                else tc::umma_commit(&tfull[a]);                    // accumulator ready for the epilogue
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> registers -> smem transpose -> global =====================
        // 8 warps: warp w reads TMEM lane quarter (w % 4) and column half ((w - 4) / 4) of the 128 x 128 tile.
        // tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns per load).  Row-major
        // outputs are transposed through a swizzled per-warp smem tile so that each global access of the
        // warp covers whole 128-byte row segments; alpha / bias / activation / residual are applied in that
        // coalesced layout, where a lane owns the same 4 columns for all rows.  The residual rows of a chunk
        // are requested (8 x 16 B per lane, unconditionally, rows clamped) BEFORE the TMEM load so that their
        // DRAM latency overlaps the accumulator read instead of serialising per row.
        // The transposed operand output (V^T) re-reads the same smem tile column-wise: a lane then owns one
        // output column = 32 consecutive V^T elements (64 contiguous bytes per plane).
        const int ew = warp & 3;                               // == warp % 4: TMEM lane quarter
        const int chalf = (warp - 4) >> 2;                     // column half of the tile
        const float alpha = p.alpha, slope = p.slope;
        const int act = p.act, obf = p.out_bf16, nplanes = p.out_planes;
        const int N = p.N, M = p.M, h_split = p.h_split;
        const int ldc = p.ldc, ldr = p.ldr, ldh = p.ldh;
        const bool vecC = (ldc & 3) == 0, vecR = (ldr & 3) == 0, vecH = (ldh & 3) == 0;
        float* stg = stage_buf + (warp - 4) * (32 * 32);
        const int c4 = lane & 7, rsub = lane >> 3;
        int it = 0;
        for (long long t = unit0; t < total_tiles; t += unit_step, ++it) {
            const int z = (int)(t / tiles_per_z), r = (int)(t - (long long)z * tiles_per_z);
            const int m_blk = (r / tiles_n) * CTAS + rank, n_blk = (r % tiles_n) * MC + pairq;
            const int zo = z / p.nb_inner, zi = z - zo * p.nb_inner;
            const int a = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int row0 = m_blk * BM + ew * 32;             // first row of this warp's lane quarter
            const int row = row0 + lane;
            const bool row_ok = row < M;
            const uint32_t tadr = tmem_base + a * C_::ACC_COLS + ((uint32_t)(ew * 32) << 16);
            float* Cz = p.C ? p.C + zo * p.c_so + zi * p.c_si : nullptr;
            const float* Rz = p.R ? p.R + zo * p.r_so + zi * p.r_si : nullptr;
            __half* Hz = p.H ? p.H + zo * p.h_so + zi * p.h_si : nullptr;
            __half* Tz = p.HT ? p.HT + zo * p.t_so + zi * p.t_si : nullptr;
            bool waited = false;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c0 = chalf * 64 + cc * 32;
                const int col0 = n_blk * BN + c0;
                if (col0 >= N) break;                          // warp-uniform
                const int col = col0 + c4 * 4;
                const bool want_h = Hz && col0 < h_split;
                const bool want_t = Tz && col0 + 32 > h_split;
                const bool rowmajor = Cz != nullptr || want_h;
                const bool full = col0 + 32 <= N;
                const bool fast = full && (!Cz || vecC) && (!Rz || vecR) && (!want_h || (vecH && col0 + 32 <= h_split));
                // ---- residual rows of this chunk: all 8 requests in flight before the accumulator is read ----
                float4 rq[8];
                if (Rz && rowmajor && fast) {
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr) {
                        const int gr = min(row0 + itr * 4 + rsub, M - 1);
                        rq[itr] = __ldg(reinterpret_cast<const float4*>(Rz + (size_t)gr * ldr + col));
                    }
                }
                if (!waited) {
                    tc::mbar_wait(&tfull[a], aph);
                    tc::tc_fence_after();
                    waited = true;
                }
                uint32_t r0[32];
                float v[32];
                tc::tmem_ld_32x32(tadr + c0, r0);
                if (NTERMS == 3) {
                    uint32_t r1[32];
                    tc::tmem_ld_32x32(tadr + BN + c0, r1);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r1[j]), 1.f / 2048.f, __uint_as_float(r0[j]));
                } else {
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
                }
                const bool t_fast = want_t && full && col0 >= h_split && row0 + 32 <= M && (p.ldt & 7) == 0;
                if (want_t && !t_fast) {                       // ragged tile: scalar transposed stores
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int cj = col0 + j;
                            if (cj < N && cj >= h_split) {
                                float x = alpha * v[j];
                                if (p.bias) x += __ldg(p.bias + cj);
                                if (act == 1) x = leaky(x, slope);
                                unsigned short* tt = reinterpret_cast<unsigned short*>(Tz) + (size_t)(cj - h_split) * p.ldt + row;
                                tt[0] = (unsigned short)(pack_h2(x, 0.f, obf) & 0xffff);
                                if (nplanes == 2) tt[p.t_plane] = (unsigned short)(pack_h2(lo_part(x, obf), 0.f, obf) & 0xffff);
                            }
                        }
                    }
                }
                if (!rowmajor && !t_fast) continue;            // warp-uniform
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) =
                        make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                if (t_fast) {
                    // Four lanes per output column, 8 consecutive rows (16 bytes per plane) each: a store instruction covers
                    // 8 columns x 64 contiguous bytes of V^T (one lane per column, 4 x 16 B each, cost 32 L1 wavefronts per
                    // instruction and made the epilogue of a V^T tile as long as its K = 512 main loop).  The rows are read
                    // in an order rotated by 2 * (lane & 3) so that the four lanes of a column, which share the swizzle
                    // pattern, hit different banks; the packed row pairs are rotated back with selects.
                    const int q = lane & 3, cs = lane >> 2;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int cl = cs + 8 * i;                                    // column of the chunk
                        const float bz = p.bias ? __ldg(p.bias + col0 + cl) : 0.f;
                        uint32_t wh[4], wl[4];
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const int rp = (m + q) & 3;                               // row pair held by word m
                            const int rr = q * 8 + 2 * rp;
                            float y0 = fmaf(alpha, stg[rr * 32 + (((cl >> 2) ^ (rr & 7)) << 2) + (cl & 3)], bz);
                            float y1 = fmaf(alpha, stg[(rr + 1) * 32 + (((cl >> 2) ^ ((rr + 1) & 7)) << 2) + (cl & 3)], bz);
                            if (act == 1) { y0 = leaky(y0, slope); y1 = leaky(y1, slope); }
                            wh[m] = pack_h2(y0, y1, obf);
                            wl[m] = pack_h2(lo_part(y0, obf), lo_part(y1, obf), obf);
                        }
                        // word j of the store = row pair j = wh[(j - q) & 3]
                        uint32_t th[4], tl[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            th[j] = (q & 1) ? wh[(j + 3) & 3] : wh[j];
                            tl[j] = (q & 1) ? wl[(j + 3) & 3] : wl[j];
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            wh[j] = (q & 2) ? th[(j + 2) & 3] : th[j];
                            wl[j] = (q & 2) ? tl[(j + 2) & 3] : tl[j];
                        }
                        __half* tt = Tz + (size_t)(col0 + cl - h_split) * p.ldt + row0 + q * 8;
                        *reinterpret_cast<uint4*>(tt) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                        if (nplanes == 2) *reinterpret_cast<uint4*>(tt + p.t_plane) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                    }
                }
                if (rowmajor && fast && row0 + 32 <= M) {
                    const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const int vsel = (Cz ? 1 : 0) | (want_h ? 2 : 0) | (Rz ? 4 : 0) | (act == 1 ? 8 : 0) |
                                     (want_h && nplanes == 2 ? 16 : 0);
                    float* c0p = Cz ? Cz + (size_t)(row0 + rsub) * ldc + col : nullptr;
                    __half* h0p = want_h ? Hz + (size_t)(row0 + rsub) * ldh + col : nullptr;
                    // warp-uniform switch; direct calls keep rq[] in registers
#define VCR_ROWS_CASE(v) \
    case v: store_rows_full<v, FMT == 1 ? 1 : 0>(stg, b4, alpha, slope, rq, c0p, ldc, h0p, ldh, p.h_plane, rsub, c4); break;
                    switch (vsel) {
                        VCR_ROWS_CASE(1) VCR_ROWS_CASE(5) VCR_ROWS_CASE(9) VCR_ROWS_CASE(13)
                        VCR_ROWS_CASE(2) VCR_ROWS_CASE(6) VCR_ROWS_CASE(10) VCR_ROWS_CASE(14)
                        VCR_ROWS_CASE(3) VCR_ROWS_CASE(7) VCR_ROWS_CASE(11) VCR_ROWS_CASE(15)
                        VCR_ROWS_CASE(18) VCR_ROWS_CASE(22) VCR_ROWS_CASE(26) VCR_ROWS_CASE(30)
                        VCR_ROWS_CASE(19) VCR_ROWS_CASE(23) VCR_ROWS_CASE(27) VCR_ROWS_CASE(31)
                        default: break;
                    }
#undef VCR_ROWS_CASE
                } else if (rowmajor && fast) {
                    const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr) {
                        const int rl = itr * 4 + rsub;
                        const int grow = row0 + rl;
                        float4 x = *reinterpret_cast<const float4*>(stg + rl * 32 + ((c4 ^ (rl & 7)) << 2));
                        x.x = fmaf(alpha, x.x, b4.x); x.y = fmaf(alpha, x.y, b4.y);
                        x.z = fmaf(alpha, x.z, b4.z); x.w = fmaf(alpha, x.w, b4.w);
                        if (act == 1) { x.x = leaky(x.x, slope); x.y = leaky(x.y, slope); x.z = leaky(x.z, slope); x.w = leaky(x.w, slope); }
                        if (Rz) { x.x += rq[itr].x; x.y += rq[itr].y; x.z += rq[itr].z; x.w += rq[itr].w; }
                        if (grow < M) {
                            if (Cz) *reinterpret_cast<float4*>(Cz + (size_t)grow * ldc + col) = x;
                            if (want_h) {
                                __half* hh = Hz + (size_t)grow * ldh + col;
                                *reinterpret_cast<uint2*>(hh) = make_uint2(pack_h2(x.x, x.y, obf), pack_h2(x.z, x.w, obf));
                                if (nplanes == 2)
                                    *reinterpret_cast<uint2*>(hh + p.h_plane) =
                                        make_uint2(pack_h2(lo_part(x.x, obf), lo_part(x.y, obf), obf),
                                                   pack_h2(lo_part(x.z, obf), lo_part(x.w, obf), obf));
                            }
                        }
                    }
                } else if (rowmajor) {
                    // ragged / unaligned chunk: element-wise guarded path
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias) {
                        if (col < N) b4.x = p.bias[col];
                        if (col + 1 < N) b4.y = p.bias[col + 1];
                        if (col + 2 < N) b4.z = p.bias[col + 2];
                        if (col + 3 < N) b4.w = p.bias[col + 3];
                    }
                    if (col < N) {
#pragma unroll 1
                        for (int itr = 0; itr < 8; ++itr) {
                            const int rl = itr * 4 + rsub;
                            const int grow = row0 + rl;
                            if (grow >= M) continue;
                            const float4 x4 = *reinterpret_cast<const float4*>(stg + rl * 32 + ((c4 ^ (rl & 7)) << 2));
                            float ys[4] = {fmaf(alpha, x4.x, b4.x), fmaf(alpha, x4.y, b4.y), fmaf(alpha, x4.z, b4.z), fmaf(alpha, x4.w, b4.w)};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (col + q >= N) continue;
                                float y = ys[q];
                                if (act == 1) y = leaky(y, slope);
                                if (Rz) y += Rz[(size_t)grow * ldr + col + q];
                                if (Cz) Cz[(size_t)grow * ldc + col + q] = y;
                                if (want_h && col + q < h_split) {
                                    unsigned short* hh = reinterpret_cast<unsigned short*>(Hz + (size_t)grow * ldh + col + q);
                                    hh[0] = (unsigned short)(pack_h2(y, 0.f, obf) & 0xffff);
                                    if (nplanes == 2) hh[p.h_plane] = (unsigned short)(pack_h2(lo_part(y, obf), 0.f, obf) & 0xffff);
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (!waited) { tc::mbar_wait(&tfull[a], aph); tc::tc_fence_after(); }
            tc::tc_fence_before();
            if (CTAS == 2) tc::mbar_arrive_cluster(&tempty[a], crank & ~1);   // the pair leader's MMA issuer waits for both CTAs
            else tc::mbar_arrive(&tempty[a]);
        }
    }
    tc::tc_fence_before();
    if (CTAS == 2) {
        tc::cluster_sync_all();                 // neither CTA leaves (or frees TMEM) while the pair is still working
        if (warp == 2) tc::tmem_dealloc_2sm(tmem_base, C_::TMEM_COLS);
    } else {
        __syncthreads();
        if (warp == 2) tc::tmem_dealloc(tmem_base, C_::TMEM_COLS);
    }
}

// fp32 [rows, cols] (ld) -> operand format planes (hi, and lo*2^11 when planes == 2)
__global__ void to_operand_kernel(const float* __restrict__ x, int ld, long long rows, int cols,
                                  __half* __restrict__ out, int ldo, long long plane, int planes, int bf16) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = cols >> 2;
    if (i >= rows * c4) return;
    const long long r = i / c4;
    const int c = (int)(i - r * c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
    uint2 w;
    w.x = pack_h2(v.x, v.y, bf16); w.y = pack_h2(v.z, v.w, bf16);
    *reinterpret_cast<uint2*>(out + r * ldo + c) = w;
    if (planes == 2) {
        w.x = pack_h2(lo_part(v.x, bf16), lo_part(v.y, bf16), bf16);
        w.y = pack_h2(lo_part(v.z, bf16), lo_part(v.w, bf16), bf16);
        *reinterpret_cast<uint2*>(out + plane + r * ldo + c) = w;
    }
}

// CTA-pair launch (cta_group::2): cluster of 2 * MC CTAs (MC = 2: two pairs sharing the multicast A tile).  Returns
// VCR_ERR_UNSUPPORTED when the device cannot co-schedule such a cluster (the caller then uses a smaller variant).
template <int MC>
int launch_tc_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcGemmParams& p, cudaStream_t stream) {
    using C_ = Cfg<3, 2>;
    constexpr int CL = 2 * MC;
    auto kern = gemm_tc_kernel<3, 0, 2, MC>;
    // per device (nn.DataParallel drives several devices from one process, one thread each): the function attribute and the
    // number of co-schedulable clusters.  Published with release/acquire: a thread either sees the final value or
    // recomputes the same one (the probe is idempotent), never a half-written slot.
    static std::atomic<int> max_clusters_dev[64];     // 0 = unknown, n + 1 = n clusters
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev & 63;
    if (max_clusters_dev[slot].load(std::memory_order_acquire) == 0) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES) != cudaSuccess)
            return VCR_ERR_LAUNCH;
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(CL, 1, 1); q.blockDim = dim3(NTHREADS, 1, 1); q.dynamicSmemBytes = C_::SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        q.attrs = at; q.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess) { cudaGetLastError(); n = 0; }
        max_clusters_dev[slot].store(n + 1, std::memory_order_release);
    }
    const int max_clusters = max_clusters_dev[slot].load(std::memory_order_acquire) - 1;
    if (max_clusters < 1) return VCR_ERR_UNSUPPORTED;
    const long long units = (long long)vcr_cdiv(p.M, 2 * BM) * vcr_cdiv(vcr_cdiv(p.N, BN), MC) * p.nb_outer * p.nb_inner;
    const int clusters = (int)(units < max_clusters ? units : max_clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * clusters, 1, 1); cfg.blockDim = dim3(NTHREADS, 1, 1);
    cfg.dynamicSmemBytes = C_::SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p) != cudaSuccess) { cudaGetLastError(); return VCR_ERR_LAUNCH; }
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

// clusters of 4 the device can co-schedule (0 when unsupported); diagnostic for the quad variant
int quad_clusters() {
    using C_ = Cfg<3, 2>;
    auto kern = gemm_tc_kernel<3, 0, 2, 2>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES) != cudaSuccess) return 0;
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(4, 1, 1); q.blockDim = dim3(NTHREADS, 1, 1); q.dynamicSmemBytes = C_::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    q.attrs = at; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess) { cudaGetLastError(); n = 0; }
    return n;
}

template <int NTERMS, int FMT>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcGemmParams& p, cudaStream_t stream) {
    using C_ = Cfg<NTERMS>;
    auto kern = gemm_tc_kernel<NTERMS, FMT, 1, 1>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES) != cudaSuccess)
        return VCR_ERR_LAUNCH;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (long long)vcr_cdiv(p.M, BM) * vcr_cdiv(p.N, BN) * p.nb_outer * p.nb_inner;
    const int grid = (int)(tiles < sms ? tiles : sms);
    kern<<<grid, NTHREADS, C_::SMEM_BYTES, stream>>>(tmA, tmB, p);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}

}  // namespace

static std::atomic<int> g_vcr_gemm_pair{2};      // tuning knob, 0: never, 1: always, 2: auto (see vcr_set_gemm_pair)

vcr_tmap_encode_fn vcr_get_tmap_encoder() {
    static std::atomic<vcr_tmap_encode_fn> cached{nullptr};      // idempotent lookup, published atomically
    vcr_tmap_encode_fn fn = cached.load(std::memory_order_acquire);
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<vcr_tmap_encode_fn>(ptr);
        cached.store(fn, std::memory_order_release);
    }
    return fn;
}

int vcr_make_operand_tmap(CUtensorMap* out, const void* base, int cols, long long rows, int ld_elems,
                          long long plane_stride_elems, int planes, int box_rows) {
    vcr_tmap_encode_fn enc = vcr_get_tmap_encoder();
    if (!enc) return VCR_ERR_LAUNCH;
    if ((ld_elems & 7) || (reinterpret_cast<uintptr_t>(base) & 15) || (plane_stride_elems & 7)) return VCR_ERR_INVALID;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(planes > 0 ? planes : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld_elems * 2, (cuuint64_t)(planes > 1 ? plane_stride_elems : (long long)ld_elems * rows) * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? VCR_OK : VCR_ERR_INVALID;
}

// Tensor-core GEMM on operand-format inputs.
//   A: [a_planes][a_rows_total][lda] 16-bit, B likewise (both K-major).  mode: 0 = fp16 split 3-term ("h3"),
//   1 = fp16 single, 2 = bf16 single.  Per batch z = (zo, zi): A rows start at zo*a_row_o + zi*a_row_i and its
//   K range at column zo*a_col_o + zi*a_col_i (same for B); M, N, K are per batch.
//   Outputs (each optional): C fp32 (+residual R), H operand-format row-major for columns < h_split,
//   HT operand-format transposed for columns >= h_split.  out_planes 2 writes hi+lo, 1 hi only.
VCR_API int vcr_gemm_tc(const void* A, int lda, long long a_rows_total, int a_cols_total, long long a_plane,
                        long long a_row_o, long long a_row_i, int a_col_o, int a_col_i,
                        const void* B, int ldb, long long b_rows_total, int b_cols_total, long long b_plane,
                        long long b_row_o, long long b_row_i, int b_col_o, int b_col_i,
                        int M, int N, int K, int nb_outer, int nb_inner, int mode,
                        float alpha, const float* bias, int act, float slope,
                        float* C, int ldc, long long c_so, long long c_si,
                        const float* R, int ldr, long long r_so, long long r_si,
                        void* H, int ldh, long long h_plane, long long h_so, long long h_si, int h_split,
                        void* HT, int ldt, long long t_plane, long long t_so, long long t_si,
                        int out_planes, cudaStream_t stream) {
    VCR_REQUIRE(A && B && M > 0 && N > 0 && K > 0 && nb_outer > 0 && nb_inner > 0);   // no output at all = timing probe
    if (mode < 0 || mode > 2) return VCR_ERR_INVALID;
    if (((a_col_o | a_col_i | b_col_o | b_col_i) != 0) && (K % BK) != 0) return VCR_ERR_UNSUPPORTED;
    const int planes = mode == 0 ? 2 : 1;
    CUtensorMap tmA, tmB;
    int rc = vcr_make_operand_tmap(&tmA, A, a_cols_total, a_rows_total, lda, a_plane, planes, BM);
    if (rc != VCR_OK) return rc;
    rc = vcr_make_operand_tmap(&tmB, B, b_cols_total, b_rows_total, ldb, b_plane, planes, BN);
    if (rc != VCR_OK) return rc;
    TcGemmParams p;
    p.M = M; p.N = N; p.K = K; p.nb_outer = nb_outer; p.nb_inner = nb_inner;
    p.a_row_o = a_row_o; p.a_row_i = a_row_i; p.a_col_o = a_col_o; p.a_col_i = a_col_i;
    p.b_row_o = b_row_o; p.b_row_i = b_row_i; p.b_col_o = b_col_o; p.b_col_i = b_col_i;
    p.alpha = alpha; p.bias = bias; p.act = act; p.slope = slope;
    p.C = C; p.ldc = ldc; p.c_so = c_so; p.c_si = c_si;
    p.R = R; p.ldr = ldr; p.r_so = r_so; p.r_si = r_si;
    p.H = reinterpret_cast<__half*>(H); p.ldh = ldh; p.h_plane = h_plane; p.h_so = h_so; p.h_si = h_si;
    p.h_split = H ? (HT ? h_split : N) : 0;
    p.HT = reinterpret_cast<__half*>(HT); p.ldt = ldt; p.t_plane = t_plane; p.t_so = t_so; p.t_si = t_si;
    p.out_planes = out_planes; p.out_bf16 = mode == 2;
    // auto: pairs everywhere except the residual epilogue at K < 1024, the one shape class where the pair measured slower
    // (0.94-0.97x; 1.03-1.12x elsewhere, scripts/pair_diag.py on B200)
    const int pol = g_vcr_gemm_pair.load(std::memory_order_relaxed);
    if (mode == 0 && pol == 3) {
        // two pairs per cluster, A multicast: both CTAs of a pair fetch 64-row boxes of A and of B
        CUtensorMap tmA2, tmB2;
        rc = vcr_make_operand_tmap(&tmA2, A, a_cols_total, a_rows_total, lda, a_plane, planes, BM / 2);
        if (rc != VCR_OK) return rc;
        rc = vcr_make_operand_tmap(&tmB2, B, b_cols_total, b_rows_total, ldb, b_plane, planes, BN / 2);
        if (rc != VCR_OK) return rc;
        rc = launch_tc_pair<2>(tmA2, tmB2, p, stream);
        if (rc != VCR_ERR_UNSUPPORTED) return rc;
    }
    if (mode == 0 && (pol == 1 || pol == 3 || (pol == 2 && (R == nullptr || K >= 1024)))) {
        // CTA pairs (cta_group::2): each CTA of a pair stages half of the B tile (box of 64 rows)
        CUtensorMap tmB2;
        rc = vcr_make_operand_tmap(&tmB2, B, b_cols_total, b_rows_total, ldb, b_plane, planes, BN / 2);
        if (rc != VCR_OK) return rc;
        rc = launch_tc_pair<1>(tmA, tmB2, p, stream);
        if (rc != VCR_ERR_UNSUPPORTED) return rc;
    }
    if (mode == 0) return launch_tc<3, 0>(tmA, tmB, p, stream);
    if (mode == 1) return launch_tc<1, 0>(tmA, tmB, p, stream);
    return launch_tc<1, 1>(tmA, tmB, p, stream);
}

// Process-wide policy for the 3-term ("h3") GEMMs on CTA pairs (cta_group::2): 0 never, 1 always, 2 auto (default),
// 3 clusters of two pairs with the A tile multicast (falls back to pairs when such clusters cannot be scheduled).
// Returns the previous setting.  Results are bit-identical under every setting.
VCR_API int vcr_gemm_quad_clusters(void) { return quad_clusters(); }

VCR_API int vcr_set_gemm_pair(int on) {
    return g_vcr_gemm_pair.exchange(on < 0 ? 0 : (on > 3 ? 2 : on));
}

// fp32 [rows, cols] (row stride ld) -> operand format [planes][rows][ldo] (fp16, or bf16 when bf16 != 0)
VCR_API int vcr_to_operand(const float* x, int ld, long long rows, int cols, void* out, int ldo, long long plane_stride,
                           int planes, int bf16, cudaStream_t stream) {
    VCR_REQUIRE(x && out && rows > 0 && cols > 0 && (planes == 1 || planes == 2));
    if ((cols & 3) || (ld & 3) || (ldo & 3)) return VCR_ERR_INVALID;
    const long long n = rows * (cols >> 2);
    to_operand_kernel<<<vcr_cdiv(n, 256), 256, 0, stream>>>(x, ld, rows, cols, reinterpret_cast<__half*>(out), ldo,
                                                            plane_stride, planes, bf16);
    VCR_CHECK_LAUNCH();
    return VCR_OK;
}
