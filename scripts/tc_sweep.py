import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcr_net_b200 import ops
dev = "cuda:0"
def t(mode, M, N, K, out="c"):
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev)
    A, B = ops.to_operand(a, mode), ops.to_operand(w, mode)
    c = torch.empty(M, N, device=dev) if out == "c" else None
    h = ops.Operand.empty(M, N, mode, dev) if out == "h" else None
    kw = dict(c=c) if out == "c" else (dict(h=h, h_split=N) if out == "h" else {})
    for _ in range(3): ops.gemm_tc(A, B, M, N, K, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.gemm_tc(A, B, M, N, K, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    print(f"{mode:5s} out={out:4s} {M:6d}x{N:5d}x{K:5d} tiles/cta={tiles/148:5.2f}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
for mode in ("fp16", "h3"):
    for out in ("none", "c", "h"):
        t(mode, 128 * 37, 512, 512, out)     # 148 tiles: 1 per CTA
        t(mode, 128 * 37 * 4, 512, 512, out)  # 4 per CTA
        t(mode, 32768, 512, 512, out)
    t(mode, 32768, 512, 64, "none")
    t(mode, 32768, 512, 2048, "none")
    t(mode, 32768, 1024, 512, "c")
