"""GPU diagnostic: flash attention time per work item as a function of the number of 64-key tiles, with a grid that has no
wave quantisation (37 x 4 x 6 = 888 items = 6 per SM): slope = time per key tile, intercept = per-item overhead."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math
import torch
from vcr_net_b200 import ops
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)
B, H, Nq, dk, mode = 37, 4, 768, 128, sys.argv[1] if len(sys.argv) > 1 else "h3"
q = torch.randn(B * Nq, H * dk, device=dev)
Q = ops.to_operand(q, mode)
out = ops.Operand.empty(B * Nq, H * dk, mode, dev)
pts = []
for Nk in (128, 256, 512, 768, 1024, 1536, 3072):
    k = torch.randn(B * Nk, H * dk, device=dev); vt = torch.randn(B * H * dk, Nk, device=dev)
    K, VT = ops.to_operand(k, mode), ops.to_operand(vt, mode)
    for org in (3, 2):
        ops.set_flash_warps(org)
        ops.flash_attn_tc(Q, K, VT, out, B, H, Nq, Nk, dk, 1.0 / math.sqrt(dk)); torch.cuda.synchronize()
        ms = []
        for _ in range(7):
            if not os.environ.get('NOFLUSH'): flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.flash_attn_tc(Q, K, VT, out, B, H, Nq, Nk, dk, 1.0 / math.sqrt(dk)); e1.record()
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        us_item = sorted(ms)[3] * 1e3 / 6
        pts.append((org, Nk // 64, us_item))
        print(f"{mode} org {org} Nk={Nk:5d} tiles={Nk // 64:3d}  {sorted(ms)[3] * 1e3:8.1f} us  = {us_item:7.2f} us per item", flush=True)
ops.set_flash_warps(3)
for org in (3, 2):
    xs = [(t, u) for o, t, u in pts if o == org]
    n = len(xs); sx = sum(t for t, _ in xs); sy = sum(u for _, u in xs); sxx = sum(t * t for t, _ in xs); sxy = sum(t * u for t, u in xs)
    slope = (n * sxy - sx * sy) / (n * sxx - sx * sx); icpt = (sy - slope * sx) / n
    print(f"org {org}: {slope:.3f} us per 64-key tile + {icpt:.2f} us per item")
