/*
 * Canonical (bit-defined) CPU restatement of the two integer-valued ops on the VCR-Net
 * registration path.  TEST INFRASTRUCTURE ONLY: built by oracle/canon.py (gcc), loaded
 * only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 *   canon_knn  -- util/util.py:143-160 (knn): pd_ij = (-xx_j - (-2*dot_ij)) - xx_i,
 *                 top-(k+1) descending, drop rank 0.
 *   canon_fps  -- util/util.py:107-140 (farthest_point_sample).
 *
 * torch leaves two things unspecified that decide integer results: the accumulation
 * order of the matmul / sum, and the order of equal values in topk.  The canonical
 * definition fixes both so that a GPU kernel can be compared bit for bit:
 *   - dot_ij and xx_i are one fused-multiply-add chain over d = 0..D-1 starting from 0
 *     (fmaf is exactly rounded, so this is the same number on any IEEE machine);
 *   - equal distances are ordered by lower index;
 *   - FPS: barycentre = fp32( double-sum / 1 ) / fp32(N); squared distances are
 *     (dx*dx + dy*dy) + dz*dz with separately rounded products (no contraction);
 *     argmax returns the first maximal index.
 * On dyadic-grid inputs every operation above is exact, so the live torch reference
 * must agree wherever the k-th and (k+1)-th distances differ (tests/golden pins that).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC canon.c -o libvcr_canon.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* x: [B, D, N] fp32 (channel-major, as the reference passes it); idx: [B, N, k] int32 */
int canon_knn(const float* x, int B, int D, int N, int k, int32_t* idx)
{
    if (k + 1 > N) return -1;
    const int K1 = k + 1;
    float* xx = (float*)malloc(sizeof(float) * (size_t)N);
    float* bv = (float*)malloc(sizeof(float) * (size_t)K1);
    int32_t* bi = (int32_t*)malloc(sizeof(int32_t) * (size_t)K1);
    if (!xx || !bv || !bi) return -2;
    for (int b = 0; b < B; ++b) {
        const float* xb = x + (size_t)b * D * N;
        for (int i = 0; i < N; ++i) {
            float acc = 0.0f;
            for (int d = 0; d < D; ++d) acc = fmaf(xb[(size_t)d * N + i], xb[(size_t)d * N + i], acc);
            xx[i] = acc;
        }
        for (int i = 0; i < N; ++i) {
            int cnt = 0;
            for (int j = 0; j < N; ++j) {
                float dot = 0.0f;
                for (int d = 0; d < D; ++d) dot = fmaf(xb[(size_t)d * N + i], xb[(size_t)d * N + j], dot);
                const float inner = -2.0f * dot;
                float pd = -xx[j] - inner;
                pd = pd - xx[i];
                /* sorted insert, descending; equal values keep the earlier (lower) index first */
                if (cnt < K1) {
                    int p = cnt++;
                    while (p > 0 && bv[p - 1] < pd) { bv[p] = bv[p - 1]; bi[p] = bi[p - 1]; --p; }
                    bv[p] = pd; bi[p] = j;
                } else if (pd > bv[K1 - 1]) {
                    int p = K1 - 1;
                    while (p > 0 && bv[p - 1] < pd) { bv[p] = bv[p - 1]; bi[p] = bi[p - 1]; --p; }
                    bv[p] = pd; bi[p] = j;
                }
            }
            int32_t* out = idx + ((size_t)b * N + i) * k;
            for (int r = 0; r < k; ++r) out[r] = bi[r + 1];
        }
    }
    free(xx); free(bv); free(bi);
    return 0;
}

static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    const float x2 = dx * dx, y2 = dy * dy, z2 = dz * dz;   /* -ffp-contract=off: no FMA */
    const float s = x2 + y2;
    return s + z2;
}

/* xyz: [B, 3, N] fp32; out: [B, npoint] int32 */
int canon_fps(const float* xyz, int B, int N, int npoint, int32_t* out)
{
    float* dist = (float*)malloc(sizeof(float) * (size_t)N);
    if (!dist) return -2;
    for (int b = 0; b < B; ++b) {
        const float* X = xyz + (size_t)b * 3 * N;
        const float* Y = X + N;
        const float* Z = Y + N;
        double sx = 0, sy = 0, sz = 0;
        for (int n = 0; n < N; ++n) { sx += X[n]; sy += Y[n]; sz += Z[n]; }
        const float bx = (float)sx / (float)N, by = (float)sy / (float)N, bz = (float)sz / (float)N;
        int far = 0; float best = -1.0f;
        for (int n = 0; n < N; ++n) {
            const float d = sqdist3(X[n], Y[n], Z[n], bx, by, bz);
            if (d > best) { best = d; far = n; }
            dist[n] = 1e10f;
        }
        for (int s = 0; s < npoint; ++s) {
            out[(size_t)b * npoint + s] = far;
            const float cx = X[far], cy = Y[far], cz = Z[far];
            int nf = 0; float nb = -1.0f;
            for (int n = 0; n < N; ++n) {
                const float d = sqdist3(X[n], Y[n], Z[n], cx, cy, cz);
                if (d < dist[n]) dist[n] = d;
                if (dist[n] > nb) { nb = dist[n]; nf = n; }
            }
            far = nf;
        }
    }
    free(dist);
    return 0;
}
