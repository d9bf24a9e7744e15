import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, torch
from vcr_net_b200 import ops
from vcr_net_b200._lib import lib
dev = "cuda:0"
B, H, Nq, Nk, dk, mode = 37, 4, 768, 768, 128, "h3"
q = torch.randn(B * Nq, H * dk, device=dev); k = torch.randn(B * Nk, H * dk, device=dev); vt = torch.randn(B * H * dk, Nk, device=dev)
Q, K, VT = ops.to_operand(q, mode), ops.to_operand(k, mode), ops.to_operand(vt, mode)
out = ops.Operand.empty(B * Nq, H * dk, mode, dev)
for _ in range(3):
    ops.flash_attn_tc(Q, K, VT, out, B, H, Nq, Nk, dk, 1.0 / math.sqrt(dk))
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 256)()
assert lib().cdll.vcr_debug_flash_timeline(buf) == 0
t = [[buf[w * 16 + s] for s in range(16)] for w in range(8)]
t00 = t[0][0]
names = ["sm item start", "sm tiles done", "sm l exch done", "sm copy_q done", "sm o_full", "sm O in regs", "sm write-out done", "sm tile1 start",
         "mma qt_full", "mma QK0 issued", "mma QK1 issued", "mma p_full(0)", "mma before p_full(last)", "mma o_full commit"]
for w in range(6):
    print("item", w, " ".join(f"{names[s]}={t[w][s] - t00}" for s in (0, 7, 1, 2, 3, 4, 5, 6)))
    print("      ", " ".join(f"{names[s]}={t[w][s] - t00}" for s in (8, 9, 10, 11, 12, 13)))
