import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcr_net_b200 import ops
dev = "cuda:0"
def t(mode, M, N, K, out="c", iters=20):
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev)
    A, B = ops.to_operand(a, mode), ops.to_operand(w, mode)
    c = torch.empty(M, N, device=dev) if out == "c" else None
    kw = dict(c=c) if out == "c" else {}
    for _ in range(3): ops.gemm_tc(A, B, M, N, K, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm_tc(A, B, M, N, K, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"dbg={os.environ.get('VCR_TC_DEBUG','0')} {mode:5s} out={out:4s} {M:6d}x{N:5d}x{K:5d}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
for mode in ("fp16", "h3"):
    t(mode, 32768, 512, 512, "c")
    t(mode, 131072, 512, 512, "c")
    t(mode, 131072, 512, 512, "none")
