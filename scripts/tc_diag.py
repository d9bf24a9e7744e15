"""GPU diagnostic for the tcgen05 GEMM (run under `timeout`): error vs fp64 for each precision mode."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vcr_net_b200 import ops

dev = "cuda:0"
def run(M, N, K, mode, seed=0, bias=False, act=0, res=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) * 0.05
    ad, wd = a.to(dev), w.to(dev)
    A, B = ops.to_operand(ad, mode), ops.to_operand(wd, mode)
    c = torch.full((M, N), float("nan"), device=dev)
    bt = torch.randn(N, generator=g).to(dev) if bias else None
    rt = torch.randn(M, N, generator=g).to(dev) if res else None
    ops.gemm_tc(A, B, M, N, K, c=c, bias=bt, act=act, slope=0.2, residual=rt)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T
    if bias: ref = ref + bt.cpu().double()
    if act: ref = torch.where(ref >= 0, ref, 0.2 * ref)
    if res: ref = ref + rt.cpu().double()
    err = (c.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    simt = ops.gemm(ad, wd)
    e2 = (simt.cpu().double() - a.double() @ w.double().T).abs().max().item() / ref.abs().max().item()
    print(f"M={M} N={N} K={K} mode={mode:5s} bias={bias} act={act} res={res}: rel err {err:.3e}   (simt fp32 {e2:.3e})", flush=True)
    return err

print(torch.cuda.get_device_name(0))
for mode in ("fp16", "h3", "bf16"):
    run(128, 128, 64, mode)
    run(128, 128, 512, mode)
    run(256, 384, 128, mode)
run(1000, 520, 200, "h3", bias=True, act=1, res=True)
run(4096, 1536, 512, "h3", bias=True)
run(16384, 512, 1024, "h3", res=True)
# timing
for mode, M, N, K in (("h3", 32768, 512, 512), ("fp16", 32768, 512, 512), ("bf16", 32768, 1024, 512), ("h3", 32768, 1536, 512)):
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev)
    A, B = ops.to_operand(a, mode), ops.to_operand(w, mode)
    c = torch.empty(M, N, device=dev)
    for _ in range(3): ops.gemm_tc(A, B, M, N, K, c=c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm_tc(A, B, M, N, K, c=c)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{mode} {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s (fp32-equivalent)")
