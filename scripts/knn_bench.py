"""kNN kernel micro-benchmark (GPU diagnostic): CUDA-event time per launch + FP32-FMA fraction, and a bit-exactness
spot check against oracle/canon.c."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vcr_net_b200 import ops
from oracle import canon
dev = "cuda:0"
torch.manual_seed(0)
FP32_PEAK = 148 * 128 * 2 * 1.965e9 / 1e12   # TFLOP/s at max clock
for (B, D, N, tm) in [(32, 64, 1024, True), (32, 3, 1024, False), (48, 64, 768, True), (8, 64, 4096, True), (8, 3, 4096, False),
                      (2, 64, 16384, True), (32, 128, 1024, True)]:
    x = torch.randn(B, N, D, device=dev) if tm else torch.randn(B, D, N, device=dev)
    for _ in range(3):
        idx = ops.knn_topk(x, 20, token_major=tm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        idx = ops.knn_topk(x, 20, token_major=tm)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * D * N * N * B
    nb = min(B, 2)
    xc = (x[:nb].transpose(1, 2) if tm else x[:nb]).contiguous().cpu().numpy()
    ok = np.array_equal(idx[:nb].cpu().numpy(), canon.knn(xc, 20)) if N <= 4096 else None
    print(f"B={B} D={D} N={N} tm={tm}: {ms*1e3:8.1f} us  {fl/ms/1e9:6.2f} TFLOP/s ({fl/ms/1e9/FP32_PEAK*100:4.1f}% of {FP32_PEAK:.1f} fp32 FMA peak)  exact={ok}")
    if tm and ops.knn_tc_supported(D, 20):
        xop = ops.to_operand(x.view(B * N, D), "h3")
        for _ in range(3):
            idx2, fl_q = ops.knn_topk_tc(x, xop, 20, want_flagged=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            idx2 = ops.knn_topk_tc(x, xop, 20)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 10
        print(f"    tcgen05 prefilter + exact re-rank: {ms2*1e3:8.1f} us ({ms/ms2:4.2f}x)  identical={bool(torch.equal(idx, idx2))}  "
              f"queries through the exact kernel: {int(fl_q.item())} of {B*N}")
