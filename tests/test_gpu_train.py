"""GPU parity of the LPD pre-training path (BASELINE config 3: LPDNet forward + backward, reference
model/lpdnet_model.py:103-229 under torch autograd).  Gradients from the hand-written backward (csrc/train.cu +
vcr_gemm_f32) are compared with the live reference's autograd gradients stored in tests/golden/lpd_train.npz.
Tolerance 1e-4 relative (max-norm per tensor) with the reference's neighbour sets injected; the free-running LPD
loss test allows 3e-4 because a flipped near-tie neighbour changes the arg-max routing (SURVEY.md section 7;
measured 7e-6 .. 2.7e-5)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import vcr_net_b200 as V
    from vcr_net_b200 import ops
from oracle.ref_harness import default_args

DEV = "cuda:0"
TOL = 1e-4


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def nump(t):
    return t.detach().cpu().numpy()


@pytest.fixture(params=["fp32", "h3"])
def precision(request):
    from vcr_net_b200 import config
    old = config.precision
    config.set_precision(request.param)
    yield request.param
    config.set_precision(old)


def _lpd(num_points=256):
    lpd_w = load_golden("lpd_pretrained_weights")
    net = V.LPD(default_args(model="lpd", num_points=num_points)).to(DEV).train()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in lpd_w.items()}, strict=True)
    return net


def test_wgrad_and_data_grad_kernels():
    torch.manual_seed(0)
    for (M, N, K) in [(1000, 64, 3), (4096, 128, 128), (777, 512, 256), (70000, 256, 64)]:
        g = torch.randn(M, N, device=DEV)
        x = torch.randn(M, K, device=DEV)
        dW, db = ops.wgrad(g, x)
        want = (g.double().t() @ x.double())
        assert float((dW.double() - want).abs().max() / want.abs().max()) < 1e-5
        assert float((db.double() - g.double().sum(0)).abs().max() / g.double().sum(0).abs().max()) < 1e-5
    g = torch.randn(300, 96, device=DEV)
    w = torch.randn(96, 40, device=DEV)                  # [N_out, K_in]
    got = ops.gemm(g, w, b_layout=1)                     # g @ w
    assert float((got.double() - g.double() @ w.double()).abs().max()) < 1e-4
    y = torch.randn(50, 33, device=DEV)
    gy = torch.randn(50, 33, device=DEV)
    gz = ops.act_bwd(gy, y, 0.2)
    assert torch.equal(gz, gy * torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.2)))


def test_lpdnet_backward_vs_reference_autograd(precision):
    g = load_golden("lpd_train")
    net = _lpd()
    emb = net.emb_nn
    x = cu(np.concatenate([g["src"], g["tgt"]], axis=0))
    keep = np.unpackbits(g["Rw_keep_bits"])[:4 * 512 * 256].reshape(4, 512, 256).astype(np.float32)
    Rw = cu(np.random.RandomState(7).standard_normal((4, 512, 256)).astype(np.float32) * keep)   # see make_golden.py
    out = emb(x, idx_feat=cu(g["idx_feat"], torch.int32), idx_xyz=cu(g["idx_xyz"], torch.int32))
    assert out.requires_grad and tuple(out.shape) == (4, 512, 256)
    assert rel_err(nump(out)[:, :, ::16], g["emb_sub"]) < TOL
    (out * Rw).sum().backward()
    errs = {}
    for name, p in emb.named_parameters():
        want = g["ga." + name]
        assert p.grad is not None and tuple(p.grad.shape) == want.shape, name
        errs[name] = rel_err(nump(p.grad), want)
    print({k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs               # measured 2e-7 .. 1.5e-6 on B200


def test_lpd_loss_and_grads_free_running():
    g = load_golden("lpd_train")
    net = _lpd()
    se, te, loss, mse, mae = net(cu(g["src"]), cu(g["tgt"]))
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    assert abs(float(mse.detach()) - float(g["mse"])) < 1e-3 * abs(float(g["mse"]))
    loss.backward()
    errs = {name: rel_err(nump(p.grad), g["gb." + name]) for name, p in net.emb_nn.named_parameters()}
    print({k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < 3e-4, errs


def test_lpd_train_step_reduces_loss():
    """A few SGD steps on one synthetic batch must lower the pre-training loss (end-to-end sanity of the gradients)."""
    from oracle import synth
    net = _lpd(512)
    pa = synth.make_pairs(4, 512, aligned=True, first_item=300)
    src, tgt = cu(pa["src"]), cu(pa["tgt"])
    opt = torch.optim.SGD(net.parameters(), lr=1e-3)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        loss = net(src, tgt)[2]
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses


def test_inference_path_has_no_autograd_graph():
    from oracle import synth
    lpd_w = load_golden("lpd_pretrained_weights")
    ck = synth.make_checkpoint(1234, emb_weights=dict(lpd_w))
    net = V.VCRNet(default_args()).to(DEV).eval()
    net.load_state_dict(synth.checkpoint_to_torch(ck), strict=True)
    p = synth.make_pairs(1, 128, first_item=5)
    out = net(cu(p["src"]), cu(p["tgt"]))
    assert all(not o.requires_grad for o in out)
