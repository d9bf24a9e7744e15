"""GPU diagnostic: time vcr_gemm_tc on the shapes one registration step launches (whole, B=16, N=1024)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcr_net_b200 import ops
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def t(mode, M, N, K, out="c", res=False, nbo=1, iters=10, label=""):
    a = torch.randn(nbo * M, K, device=dev); w = torch.randn(N, K, device=dev)
    A, B = ops.to_operand(a, mode), ops.to_operand(w, mode)
    kw = {}
    if nbo > 1: kw.update(nbo=nbo, a_off=(M, 0, 0, 0))
    if out == "c":
        kw["c"] = torch.empty(nbo * M, N, device=dev)
        if nbo > 1: kw["c_strides"] = (M * N, 0)
        if res:
            kw["residual"] = torch.randn(nbo * M, N, device=dev)
            if nbo > 1: kw["r_strides"] = (M * N, 0)
    elif out == "h":
        h = ops.Operand.empty(nbo * M, N, mode, dev); kw.update(h=h, h_split=N, h_strides=(M * h.ld, 0))
    elif out == "qkv":      # Q,K row-major operand + V transposed (per batch), as mha_tc does
        D = N // 3
        h = ops.Operand.empty(nbo * M, 2 * D, mode, dev); vt = ops.Operand.empty(nbo * D, M, mode, dev)
        kw.update(h=h, h_strides=(M * h.ld, 0), h_split=2 * D, ht=vt, ht_strides=(D * vt.ld, 0))
    for _ in range(2): ops.gemm_tc(A, B, M, N, K, **kw)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm_tc(A, B, M, N, K, **kw); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    tiles = ((M + 127) // 128) * ((N + 127) // 128) * nbo
    print(f"{mode:5s} {label:10s} out={out:4s} res={int(res)} {nbo:3d}x{M:6d}x{N:5d}x{K:5d} tiles/cta={tiles/148:5.2f}: "
          f"{ms*1e3:8.1f} us  {2*nbo*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)

for mode in ("h3", "fp16"):
    t(mode, 32768, 512, 512, "none", label="mainloop")
    t(mode, 32768, 512, 512, "c", label="q/conv3")
    t(mode, 32768, 512, 512, "c", res=True, label="wo+res")
    t(mode, 32768, 512, 1024, "c", res=True, label="ffn2+res")
    t(mode, 32768, 1024, 512, "h", label="ffn1")
    t(mode, 1024, 1536, 512, "qkv", nbo=32, label="qkv")
    t(mode, 1024, 1024, 512, "c", nbo=16, label="vcp-dot")
