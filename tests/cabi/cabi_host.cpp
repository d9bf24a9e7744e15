// Plain C++ host of the C ABI (no Python, no torch): what a compiled-language caller of include/vcr_b200.h looks like.
// Builds random clouds on a 2^-8 grid (exact fp32 arithmetic, real ties), runs vcr_knn_topk (util/util.py:143-160) and
// vcr_fps (util/util.py:107-140) on the device through caller-allocated buffers and the caller's stream, and compares the
// indices bit for bit with the CPU oracle (oracle/canon.c, test infrastructure linked only here).
//
//   g++ -std=c++17 -I include -I /usr/local/cuda/include tests/cabi/cabi_host.cpp -o /tmp/cabi_host \
//       -L vcr_net_b200 -lvcr_b200 -L oracle -lvcr_canon -L /usr/local/cuda/lib64 -lcudart
//   LD_LIBRARY_PATH=vcr_net_b200:oracle /tmp/cabi_host        (exit code 0 = identical)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "vcr_b200.h"

extern "C" int canon_knn(const float* x, int B, int D, int N, int k, int32_t* idx);
extern "C" int canon_fps(const float* xyz, int B, int N, int npoint, int32_t* out);

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 2; } } while (0)

int main() {
    const int B = 3, D = 3, N = 777, k = 20, npoint = 32;
    std::vector<float> x((size_t)B * D * N);
    uint32_t s = 12345u;
    for (auto& v : x) { s = s * 1664525u + 1013904223u; v = (float)((int)((s >> 16) & 255) - 128) / 256.0f; }

    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    float* dx; int32_t *dknn, *dfps; void* ws;
    const size_t wsb = vcr_knn_workspace_bytes(B, N);
    CK(cudaMalloc(&dx, x.size() * sizeof(float)));
    CK(cudaMalloc(&dknn, (size_t)B * N * k * sizeof(int32_t)));
    CK(cudaMalloc(&dfps, (size_t)B * npoint * sizeof(int32_t)));
    CK(cudaMalloc(&ws, wsb ? wsb : 16));
    CK(cudaMemcpyAsync(dx, x.data(), x.size() * sizeof(float), cudaMemcpyHostToDevice, st));

    int rc = vcr_knn_topk(dx, B, D, N, k, /*token_major=*/0, dknn, nullptr, ws, wsb, st);
    if (rc != 0) { std::fprintf(stderr, "vcr_knn_topk rc=%d\n", rc); return 3; }
    rc = vcr_fps(dx, B, N, npoint, dfps, nullptr, st);
    if (rc != 0) { std::fprintf(stderr, "vcr_fps rc=%d\n", rc); return 3; }
    // error behaviour of the boundary: invalid arguments come back as a negative code, nothing is launched
    if (vcr_knn_topk(nullptr, B, D, N, k, 0, dknn, nullptr, ws, wsb, st) >= 0) { std::fprintf(stderr, "null input accepted\n"); return 4; }
    if (vcr_knn_topk(dx, B, D, N, 40, 0, dknn, nullptr, ws, wsb, st) >= 0) { std::fprintf(stderr, "k=40 accepted\n"); return 4; }

    std::vector<int32_t> knn((size_t)B * N * k), fps((size_t)B * npoint);
    CK(cudaMemcpyAsync(knn.data(), dknn, knn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(fps.data(), dfps, fps.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));

    std::vector<int32_t> knn_ref(knn.size()), fps_ref(fps.size());
    canon_knn(x.data(), B, D, N, k, knn_ref.data());
    canon_fps(x.data(), B, N, npoint, fps_ref.data());
    size_t bad = 0;
    for (size_t i = 0; i < knn.size(); ++i) bad += knn[i] != knn_ref[i];
    for (size_t i = 0; i < fps.size(); ++i) bad += fps[i] != fps_ref[i];
    std::printf("abi %d: kNN %zu indices, FPS %zu indices, %zu mismatches\n", vcr_abi_version(), knn.size(), fps.size(), bad);
    cudaFree(dx); cudaFree(dknn); cudaFree(dfps); cudaFree(ws); cudaStreamDestroy(st);
    return bad ? 1 : 0;
}
