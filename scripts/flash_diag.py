"""GPU diagnostic: flash attention softmax organisations (vcr_set_flash_warps 1 = tile ping-pong, 2 = 8 warps per tile, 3 = Q and P in TMEM, 4 = 16 warps): agreement and timing on
the step's shapes.  Run under `timeout`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math
import torch
from vcr_net_b200 import ops
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)


def case(B, H, Nq, Nk, masked, mode="h3", iters=10):
    dk = 128
    q = torch.randn(B * Nq, H * dk, device=dev); k = torch.randn(B * Nk, H * dk, device=dev)
    vt = torch.randn(B * H * dk, Nk, device=dev)
    Q, K, VT = ops.to_operand(q, mode), ops.to_operand(k, mode), ops.to_operand(vt, mode)
    keep = (torch.rand(B, Nk, device=dev) < 0.766).to(torch.uint8) if masked else None
    res = {}
    for nwq in (1, 2, 3, 4):
        ops.set_flash_warps(nwq)
        out = ops.Operand.empty(B * Nq, H * dk, mode, dev)
        ops.flash_attn_tc(Q, K, VT, out, B, H, Nq, Nk, dk, 1.0 / math.sqrt(dk), keep=keep)
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.flash_attn_tc(Q, K, VT, out, B, H, Nq, Nk, dk, 1.0 / math.sqrt(dk), keep=keep); e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        res[nwq] = (out.to_float().clone(), ms / iters)
    ops.set_flash_warps(3)
    a, b = res[2][0], res[3][0]
    diff = float((a - b).abs().max() / a.abs().max())
    fl = 4.0 * B * H * Nq * Nk * dk
    print(f"{mode:5s} B={B:3d} H={H} Nq={Nq:5d} Nk={Nk:5d} masked={int(masked)}  rel diff {diff:.2e}  finite={bool(torch.isfinite(b).all())}  "
          f"TMEM operands {res[3][1]*1e3:8.1f} us {fl/res[3][1]/1e9:6.1f} TF/s | smem operands: 8 warps {res[2][1]*1e3:8.1f} us | 16 warps {res[4][1]*1e3:8.1f} us | "
          f"ping-pong {res[1][1]*1e3:8.1f} us   x{res[2][1]/res[3][1]:4.2f} vs 8 warps", flush=True)


case(2, 4, 200, 332, False, iters=1)
case(2, 4, 200, 332, True, iters=1)
case(1, 1, 129, 64, True, iters=1)
case(24, 4, 768, 768, False)
case(48, 4, 768, 768, False)
case(48, 4, 768, 768, True)
case(32, 4, 1024, 1024, False)
case(8, 4, 4096, 4096, False, iters=3)
case(32, 4, 1024, 1024, False, mode="fp16")
