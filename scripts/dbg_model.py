import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vcr_net_b200 as V
from vcr_net_b200.synthetic import default_args, state_dict
from vcr_net_b200 import ops
dev = "cuda:0"
B, N = int(sys.argv[1]), int(sys.argv[2])
a = default_args()
net = V.VCRNet(a).to(dev).eval()
net.load_state_dict(state_dict(1234), strict=True)
torch.manual_seed(0)
src = torch.rand(B, 3, N, device=dev) - 0.5; tgt = torch.rand(B, 3, N, device=dev) - 0.5
if len(sys.argv) > 3: ops.set_flash_warps(int(sys.argv[3]))
_orig = ops.flash_attn_tc
variant = os.environ.get("DBG", "")
def _wrapped(q, k, vt, out, B_, H, Nq, Nk, dk, scale, keep=None, lse=None):
    if "poison" in variant: out.buf.view(torch.int16).fill_(0x7e00)
    if "zero" in variant: out.buf.zero_()
    if "pre" in variant: torch.cuda.synchronize()
    _orig(q, k, vt, out, B_, H, Nq, Nk, dk, scale, keep=keep, lse=lse)
    if "post" in variant: torch.cuda.synchronize()
    if "check" in variant:
        torch.cuda.synchronize()
        raw = out.buf.view(torch.int16)[: out.planes * out.plane_stride].view(out.planes, -1, out.ld)[:, : out.rows, : out.cols]
        f = raw.view(torch.float16).float()
        bad = ~torch.isfinite(f)
        print("flash out bad", int(bad.sum()), "per plane", [int(bad[p_].sum()) for p_ in range(out.planes)], flush=True)
        if bad.any():
            br = bad.any(dim=0).any(dim=1).nonzero().flatten().cpu().numpy(); bc = bad.any(dim=0).any(dim=0).nonzero().flatten().cpu().numpy()
            print("   rows", len(br), br[:8], br[-4:], "row tiles(128)", sorted(set((br // 128).tolist()))[:20], " rows mod 32:", sorted(set((br % 32).tolist()))[:40])
            print("   cols", len(bc), bc[:8], bc[-4:], "col blocks(64)", sorted(set((bc // 64).tolist())))
            big = (f.abs() > 1e4) & torch.isfinite(f); print("   huge finite", int(big.sum()))
    return out
ops.flash_attn_tc = _wrapped
out = V.vcrnetIter(net, src, tgt, iter=1)
torch.cuda.synchronize()
print("ok", [bool(torch.isfinite(o).all()) for o in out[:4]])
