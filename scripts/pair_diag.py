"""GPU diagnostic: the CTA-pair (cta_group::2) variant of vcr_gemm_tc against the single-CTA kernel -- results must be
bit-identical -- and its timing on the shapes of one registration step.  Run under `timeout`: a pipeline bug hangs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcr_net_b200 import ops
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.manual_seed(0)


def case(M, N, K, out="c", res_flag=False, nbo=1, bias=False, act=0, iters=10, label=""):
    a = torch.randn(nbo * M, K, device=dev); w = torch.randn(N, K, device=dev)
    A, B = ops.to_operand(a, "h3"), ops.to_operand(w, "h3")
    bia = torch.randn(N, device=dev) if bias else None
    R = torch.randn(nbo * M, N, device=dev) if res_flag else None

    def run(pair, time_it):
        ops.set_gemm_pair(pair)
        kw = dict(bias=bia, act=act, slope=0.2)
        if nbo > 1: kw.update(nbo=nbo, a_off=(M, 0, 0, 0))
        bufs = []
        if out == "c":
            kw["c"] = torch.full((nbo * M, N), float("nan"), device=dev); bufs.append(kw["c"])
            if nbo > 1: kw["c_strides"] = (M * N, 0)
            if res_flag:
                kw["residual"] = R
                if nbo > 1: kw["r_strides"] = (M * N, 0)
        elif out == "h":
            h = ops.Operand.empty(nbo * M, N, "h3", dev); h.buf.zero_()
            kw.update(h=h, h_split=N, h_strides=(M * h.ld, 0)); bufs.append(h.buf)
        elif out == "qkv":
            D = N // 3
            h = ops.Operand.empty(nbo * M, 2 * D, "h3", dev); vt = ops.Operand.empty(nbo * D, M, "h3", dev)
            h.buf.zero_(); vt.buf.zero_()
            kw.update(h=h, h_strides=(M * h.ld, 0), h_split=2 * D, ht=vt, ht_strides=(D * vt.ld, 0))
            bufs += [h.buf, vt.buf]
        ops.gemm_tc(A, B, M, N, K, **kw)
        torch.cuda.synchronize()
        ms = 0.0
        if time_it:
            for _ in range(iters):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ops.gemm_tc(A, B, M, N, K, **kw); e1.record(); torch.cuda.synchronize()
                ms += e0.elapsed_time(e1)
            ms /= iters
        return [b.clone() for b in bufs], ms

    ref, ms0 = run(0, iters > 0)
    res, line = [], ""
    for pol in POLICIES:
        got, ms = run(pol, iters > 0)
        res.append(all(torch.equal(x.view(torch.uint8), y.view(torch.uint8)) for x, y in zip(ref, got)))
        if iters:
            line += f" | pol{pol} {ms*1e3:7.1f} us {2.0*nbo*M*N*K/ms/1e9:6.1f} TF/s x{ms0/ms:4.2f}"
    ops.set_gemm_pair("auto")
    fl = 2.0 * nbo * M * N * K
    t = f"{ms0*1e3:7.1f} us {fl/ms0/1e9:6.1f} TF/s" + line if iters else ""
    print(f"{label:10s} out={out:4s} res={int(res_flag)} {nbo:3d}x{M:6d}x{N:5d}x{K:5d}  identical={res}  {t}", flush=True)
    return all(res)


POLICIES = [int(x) for x in os.environ.get("PAIR_POLICIES", "1,3").split(",")]
from vcr_net_b200._lib import lib
print("quad clusters schedulable:", lib().vcr_gemm_quad_clusters(), flush=True)
ok = True
# correctness first, small and ragged (M not a multiple of 256 / 128, N tails, K tails, batches)
for (M, N, K, nbo) in [(256, 128, 64, 1), (128, 128, 64, 1), (300, 200, 72, 1), (494, 494, 512, 3), (768, 768, 512, 2),
                       (1000, 130, 520, 1), (64, 64, 64, 5), (257, 512, 1024, 1), (512, 128, 512, 2), (1024, 384, 128, 1)]:
    ok &= case(M, N, K, "c", res_flag=True, nbo=nbo, bias=True, act=1, iters=0, label="ragged")
    ok &= case(M, N, K, "h", nbo=nbo, iters=0, label="ragged")
ok &= case(768, 1536, 512, "qkv", nbo=4, iters=0, label="qkv")
print("ALL IDENTICAL" if ok else "MISMATCH", flush=True)
# timing on the step's shapes (partial: 48 clouds x 768 points; whole: 32 x 1024)
case(36864, 512, 512, "none", label="mainloop")
case(36864, 512, 512, "c", label="q/conv3")
case(36864, 512, 512, "c", res_flag=True, label="wo+res")
case(36864, 512, 1024, "c", res_flag=True, label="ffn2+res")
case(36864, 1024, 512, "h", label="ffn1")
case(768, 1536, 512, "qkv", nbo=48, label="qkv")
case(768, 1024, 512, "c", nbo=48, label="kv")
case(32768, 512, 512, "c", label="whole q")
case(1024, 1536, 512, "qkv", nbo=32, label="whole qkv")
case(1024, 1024, 512, "c", nbo=16, label="vcp-dot")
