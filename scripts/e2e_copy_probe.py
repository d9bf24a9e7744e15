"""GPU diagnostic: what the e2e bracket adds to a step -- the pinned H2D copies of one batch and the four pose D2H copies."""
import torch
dev = "cuda:0"
B, N = 24, 768
hs = torch.rand(B, 3, N).pin_memory(); ht = torch.rand(B, 3, N).pin_memory()
sb = torch.empty(B, 3, N, device=dev); tb = torch.empty(B, 3, N, device=dev)
outs = [torch.rand(s, device=dev) for s in ((B, 3, 3), (B, 3), (B, 3, 3), (B, 3))]
hp = [torch.empty(o.shape).pin_memory() for o in outs]
spin = torch.empty(64 << 20, device=dev)
def timed(fn, n=50):
    ms = 0.0
    for _ in range(n):
        spin.fill_(1.0)                      # keep the stream busy so that the host runs ahead, as in the bench loop
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ms += e0.elapsed_time(e1)
    return ms / n * 1e3
def h2d():
    sb.copy_(hs, non_blocking=True); tb.copy_(ht, non_blocking=True)
def d2h():
    for h, o in zip(hp, outs): h.copy_(o, non_blocking=True)
def both():
    h2d(); d2h()
for name, fn in (("2 x H2D 221 KB", h2d), ("4 x D2H poses", d2h), ("both", both), ("empty bracket", lambda: None)):
    print(f"{name:18s} {timed(fn):8.1f} us")
