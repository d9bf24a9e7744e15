"""BASELINE config 5: kernel sweep 512-16384 pts -- kNN/top-k, attention, soft-correspondence and SVD kernels vs roofline.
CUDA-event time per call (after warm-up, L2 flushed between calls), against MEASURED_PEAKS.json.  GPU diagnostic; its
output is committed under profiles/."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcr_net_b200 import ops, config
from vcr_net_b200 import functional as Fn
dev = "cuda:0"
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
TC, HBM = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
FMA = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
config.set_precision("h3")


def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


print(f"# peaks: bf16 sustained {TC} TFLOP/s, HBM {HBM} GB/s, fp32 FMA {FMA:.1f} TFLOP/s (148 SM x 128 lanes x 2 x max clock)")
print("# kernel, N, batch, us, achieved, unit, frac_of_peak, note")
torch.manual_seed(0)
for N in (512, 1024, 2048, 4096, 8192, 16384):
    B = max(1, 65536 // N)                      # clouds per call (64 Ki points in flight)
    # --- kNN / top-k -------------------------------------------------------------------------------------
    f = torch.randn(B, N, 64, device=dev)
    x = torch.rand(B, 3, N, device=dev) - 0.5
    ms = timeit(lambda: ops.knn_topk(f, 20, token_major=True))
    fl = 2.0 * 64 * N * N * B
    print(f"knn_topk_64d, {N}, {B}, {ms*1e3:.1f}, {fl/ms/1e9:.2f}, TFLOP/s fp32, {fl/ms/1e9/FMA:.3f}, of fp32 FMA peak")
    fop = ops.to_operand(f.view(B * N, 64), "h3")
    ms_tc = timeit(lambda: ops.knn_topk_tc(f, fop, 20))
    print(f"knn_topk_64d_tc_prefilter, {N}, {B}, {ms_tc*1e3:.1f}, {ms/ms_tc:.2f}, x vs SIMT, , tcgen05 distance tiles + exact re-rank (bit-identical); selection-bound")
    ms = timeit(lambda: ops.knn_topk(x, 20, token_major=False))
    print(f"knn_topk_3d, {N}, {B}, {ms*1e3:.1f}, {B*N*N/ms/1e6:.1f}, Gpair/s, , selection-bound ({2.0*3*N*N*B/ms/1e9:.2f} TFLOP/s)")
    # --- flash attention (4 heads x 128) --------------------------------------------------------------------
    Ba = max(1, 32768 // N)
    q = ops.to_operand(torch.randn(Ba * N, 512, device=dev), "h3")
    k = ops.to_operand(torch.randn(Ba * N, 512, device=dev), "h3")
    vt = ops.to_operand(torch.randn(Ba * 512, N, device=dev), "h3")
    o = ops.Operand.empty(Ba * N, 512, "h3", dev)
    ms = timeit(lambda: ops.flash_attn_tc(q, k, vt, o, Ba, 4, N, N, 128, 1.0 / math.sqrt(128)))
    fl = 4.0 * N * N * 512 * Ba
    print(f"flash_attn_h3, {N}, {Ba}, {ms*1e3:.1f}, {fl/ms/1e9:.1f}, TFLOP/s algorithmic, {fl/ms/1e9/TC:.3f}, tensor work = 3x: {3*fl/ms/1e9/TC:.3f} of measured bf16 sustained")
    # --- soft correspondence (getCopairALL) ---------------------------------------------------------------------
    Bs = max(1, 16384 // N)
    s_tok = torch.randn(Bs, N, 512, device=dev) * 0.1
    t_tok = torch.randn(Bs, N, 512, device=dev) * 0.1
    txyz = torch.rand(Bs, 3, N, device=dev)
    ms = timeit(lambda: Fn.vcp_whole(s_tok, t_tok, txyz))
    fl = (2.0 * 512 + 6) * N * N * Bs
    print(f"softcorr_whole, {N}, {Bs}, {ms*1e3:.1f}, {fl/ms/1e9:.1f}, TFLOP/s algorithmic, {fl/ms/1e9/TC:.3f}, fused tcgen05 GEMM + online softmax + weighted target sum (no {4.0*N*N*Bs/1e6:.0f} MB score matrix in HBM); tensor work = 3x: {3*fl/ms/1e9/TC:.3f}")
    # --- SVD head -------------------------------------------------------------------------------------------
    for Bp in (1, 16, 256):
        a = torch.rand(Bp, 3, N, device=dev); b = torch.rand(Bp, 3, N, device=dev)
        ms = timeit(lambda: ops.svd_head(a, b))
        by = 2.0 * 4 * 3 * N * Bp
        print(f"svd_head, {N}, {Bp}, {ms*1e3:.1f}, {by/ms/1e6:.1f}, GB/s, {by/ms/1e6/HBM:.4f}, {ms*1e3/Bp:.2f} us/pair (latency-bound)")
