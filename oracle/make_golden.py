"""Generate tests/golden/*.npz from the LIVE reference (CPU torch, fp32, 1 thread).

TEST INFRASTRUCTURE ONLY; runs only where /root/reference is mounted (the build container).
The reference holds no golden vectors of its own, so these files ARE the
parity pin: every tensor below is produced by calling the unmodified reference modules
through oracle/ref_harness.py.  Re-run with ``python -m oracle.make_golden``.

Weights: ``lpd_pretrained_weights.npz`` holds the 12 ``emb_nn.*`` tensors of the reference's
only shipped checkpoint (pretrained/lpd-pretrained.t7, data not source) so that the GPU
box -- which has no /root/reference -- can rebuild the same synthetic 59-key checkpoint:
``synth.make_checkpoint(1234, emb_weights=lpd)``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
OV2 = synth.OVERLAP2_0575


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def N_(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name:28s} {os.path.getsize(path) / 1e6:7.2f} MB  " +
          " ".join(f"{k}{tuple(v.shape)}" for k, v in arrs.items()))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)
    torch.manual_seed(0)
    ref = ref_harness.import_reference()
    U, LP, TR, VM = ref.util, ref.lpdnet_model, ref.transformer, ref.vcrnet_model

    # ---- weights -------------------------------------------------------------------------
    lpd_sd = torch.load(os.path.join(ref_harness.REF_ROOT, "pretrained", "lpd-pretrained.t7"),
                        map_location="cpu", weights_only=True)
    lpd = {k: N_(v).astype(np.float32) for k, v in lpd_sd.items()}
    save("lpd_pretrained_weights", **lpd)
    ckpt = synth.make_checkpoint(1234, emb_weights=lpd)
    sd_t = synth.checkpoint_to_torch(ckpt)

    def build(partial):
        args = ref_harness.default_args(partial=partial, overlap2=OV2 if partial else 0.75)
        net = VM.VCRNet(args).eval()
        ref_keys = list(net.state_dict().keys())
        assert ref_keys == list(ckpt.keys()), "synthetic checkpoint key order != reference"
        net.load_state_dict(sd_t, strict=True)
        return net

    net_w, net_p = build(False), build(True)
    save("state_dict_layout", keys=np.array(list(ckpt.keys())),
         shapes=np.array([str(tuple(v.shape)) for v in ckpt.values()]))

    rs = np.random.RandomState(7)
    with torch.no_grad():
        # ---- kNN ---------------------------------------------------------------------------
        x3g = synth.grid_cloud(rs, (2, 3, 512), 8, -0.5, 0.5)
        x64g = synth.grid_cloud(rs, (2, 64, 256), 6, -2.0, 2.0)
        pairs = synth.make_pairs(2, 1024)
        x3f = pairs["src"]
        f64 = N_(F.leaky_relu(net_w.emb_nn.conv2_lpd(F.leaky_relu(net_w.emb_nn.conv1_lpd(T(x3f)), 0.0)), 0.0))
        save("knn", x3g=x3g, idx3g=N_(U.knn(T(x3g), 20)), x64g=x64g, idx64g=N_(U.knn(T(x64g), 20)),
             x3f=x3f, idx3f=N_(U.knn(T(x3f), 20)), x64f=f64, idx64f=N_(U.knn(T(f64), 20)))
        gf = N_(U.get_graph_feature(T(x64g[:1, :8, :64]), k=5))
        save("graph_feature", x=x64g[:1, :8, :64], k=np.array(5), out=gf)

        # ---- FPS ---------------------------------------------------------------------------
        pg = synth.grid_cloud(rs, (4, 3, 1024), 8, -0.5, 0.5)
        pg2 = synth.grid_cloud(rs, (2, 3, 768), 8, -0.5, 0.5)
        pf = synth.make_pairs(3, 1024, first_item=5)["tgt"]
        save("fps", pg=pg, ig=N_(U.farthest_point_sample(T(pg), 32)),
             pg2=pg2, ig2=N_(U.farthest_point_sample(T(pg2), 32)),
             pf=pf, i_f=N_(U.farthest_point_sample(T(pf), 32)))

        # ---- LPDNet ------------------------------------------------------------------------
        xl = synth.make_pairs(1, 512, first_item=3)["src"]
        emb0 = net_w.emb_nn
        h = F.leaky_relu(emb0.conv2_lpd(F.leaky_relu(emb0.conv1_lpd(T(xl)), 0.0)), 0.0)
        args = ref_harness.default_args()
        emb02 = LP.LPDNet(args, negative_slope=0.2).eval()
        emb02.load_state_dict({k[len("emb_nn."):]: v for k, v in lpd_sd.items()})
        h02 = F.leaky_relu(emb02.conv2_lpd(F.leaky_relu(emb02.conv1_lpd(T(xl)), 0.2)), 0.2)
        save("lpdnet", x=xl, out_s0=N_(emb0(T(xl))), out_s02=N_(emb02(T(xl))),
             idx_feat_s0=N_(U.knn(h, 20)), idx_feat_s02=N_(U.knn(h02, 20)), idx_xyz=N_(U.knn(T(xl), 20)))

        # ---- Transformer (whole + partial) ----------------------------------------------------
        pr = synth.make_pairs(1, 256, first_item=11)
        se, te = emb0(T(pr["src"])), emb0(T(pr["tgt"]))
        sp, tp = net_w.pointer(se, te)
        spp, tpp = net_p.pointer(se, te)
        save("transformer", src_emb=N_(se), tgt_emb=N_(te), src_p=N_(sp), tgt_p=N_(tp),
             src_p_partial=N_(spp), tgt_p_partial=N_(tpp), overlap2=np.array(OV2))

        ln = TR.LayerNorm(512)
        ln.a_2.data = torch.from_numpy(rs.randn(512).astype(np.float32))
        ln.b_2.data = torch.from_numpy(rs.randn(512).astype(np.float32))
        xln = rs.randn(3, 17, 512).astype(np.float32) * 3 + 1
        save("layernorm", x=xln, a=N_(ln.a_2), b=N_(ln.b_2), out=N_(ln(T(xln))))

        q = rs.randn(2, 4, 96, 128).astype(np.float32)
        k = rs.randn(2, 4, 80, 128).astype(np.float32)
        v = rs.randn(2, 4, 80, 128).astype(np.float32)
        o0, p0 = TR.attention(T(q), T(k), T(v))
        o1, p1 = TR.attention(T(q), T(k), T(v), is_src=True, overlap2=OV2)
        save("attention", q=q, k=k, v=v, out=N_(o0), out_src=N_(o1), overlap2=np.array(OV2),
             colsum=N_(p0.sum(dim=(1, 2))), kept=N_((p1.sum(dim=(1, 2)) > 0)))

        # ---- VCP head ----------------------------------------------------------------------
        se2, te2 = se + sp, te + tp
        s_a, c_a = net_w.head(se2, te2, T(pr["src"]), T(pr["tgt"]))
        hp = net_p.head
        so, seo, to, teo, _, _ = hp.selectCom(T(pr["src"]), se2, T(pr["tgt"]), te2, overlap2=OV2)
        s_p, c_p = hp.getCopair(so, seo, to, teo, OV2)
        save("vcp_head", src=pr["src"], tgt=pr["tgt"], src_emb=N_(se2), tgt_emb=N_(te2),
             corr_all=N_(c_a), sel_src=N_(so), sel_tgt=N_(to), sel_src_emb=N_(seo), sel_tgt_emb=N_(teo),
             part_src=N_(s_p), part_corr=N_(c_p), overlap2=np.array(OV2))

        # ---- SVD head ----------------------------------------------------------------------
        ps = synth.make_pairs(8, 128, first_item=20)
        a = ps["src"]
        b = ps["tgt"] + rs.randn(*ps["tgt"].shape).astype(np.float32) * 0.01
        b[4:6] = b[4:6] * np.array([1, 1, -1], np.float32).reshape(1, 3, 1)   # mirrored => det<0 branch
        b[6] = rs.randn(3, 128).astype(np.float32)                            # unrelated clouds
        a[7, 2] = 0.0                                                         # planar source => rank-2 H
        R, t = net_w.svd(T(a), T(b))
        save("svd_head", src=a, corr=b, R=N_(R), t=N_(t))

        # ---- full network -------------------------------------------------------------------
        pw = synth.make_pairs(2, 512, first_item=30)
        outs = VM.vcrnetIter(net_w, T(pw["src"]), T(pw["tgt"]), iter=1)
        se_w = emb0(T(pw["src"]))
        save("vcrnet_whole", src=pw["src"], tgt=pw["tgt"], R_gt=pw["R_ab"], t_gt=pw["t_ab"],
             srcK=N_(outs[0]), corrK=N_(outs[1]), R_ab=N_(outs[2]), t_ab=N_(outs[3]),
             R_ba=N_(outs[4]), t_ba=N_(outs[5]), src_emb0=N_(se_w))
        outs2 = VM.vcrnetIter(net_w, T(pw["src"]), T(pw["tgt"]), iter=2)
        save("vcrnet_whole_iter2", R_ab=N_(outs2[2]), t_ab=N_(outs2[3]), corrK=N_(outs2[1]))

        pp = synth.make_pairs(2, 512, partial=True, first_item=40)
        outp1 = VM.vcrnetIter(net_p, T(pp["src"]), T(pp["tgt"]), iter=1)
        outp = VM.vcrnetIter(net_p, T(pp["src"]), T(pp["tgt"]), iter=3)
        save("vcrnet_partial", src=pp["src"], tgt=pp["tgt"], R_gt=pp["R_ab"], t_gt=pp["t_ab"],
             overlap2=np.array(OV2),
             srcK1=N_(outp1[0]), corrK1=N_(outp1[1]), R_ab1=N_(outp1[2]), t_ab1=N_(outp1[3]),
             srcK=N_(outp[0]), corrK=N_(outp[1]), R_ab=N_(outp[2]), t_ab=N_(outp[3]),
             R_ba=N_(outp[4]), t_ba=N_(outp[5]))

        # ---- LPD pre-train forward (loss) ------------------------------------------------------
        pa = synth.make_pairs(2, 512, aligned=True, first_item=50)
        largs = ref_harness.default_args(model="lpd", num_points=512)
        lnet = LP.LPD(largs).eval()
        lnet.load_state_dict(lpd_sd, strict=True)
        se_l, te_l, loss, mse, mae = lnet(T(pa["src"]), T(pa["tgt"]))
        save("lpd_loss", src=pa["src"], tgt=pa["tgt"], loss=N_(loss), mse=N_(mse), mae=N_(mae),
             src_emb=N_(se_l))

    make_lpd_train(ref, lpd_sd)
    make_variants(ref)
    make_ragged(ref)


def make_ragged(ref=None):
    """Live-reference outputs at the ragged sizes of tests/test_gpu_parity.py::test_vcrnet_ragged_sizes_vs_oracle (odd point
    counts, one pair; source and target clouds of different sizes) and an odd partial crop with iter=2: pins the oracle where
    the GPU test leans on it.  Standalone: python -c "from oracle import make_golden as m; m.make_ragged()"."""
    torch.set_num_threads(1)
    ref = ref or ref_harness.import_reference()
    VM = ref.vcrnet_model
    lpd = dict(np.load(os.path.join(GOLD, "lpd_pretrained_weights.npz")))
    sd_t = synth.checkpoint_to_torch(synth.make_checkpoint(1234, emb_weights=lpd))

    def build(partial):
        net = VM.VCRNet(ref_harness.default_args(partial=partial, overlap2=OV2 if partial else 0.75)).eval()
        net.load_state_dict(sd_t, strict=True)
        return net

    with torch.no_grad():
        net_w, net_p = build(False), build(True)
        p = synth.make_pairs(1, 333, first_item=70)
        o1 = VM.vcrnetIter(net_w, T(p["src"]), T(p["tgt"]), iter=1)
        q = synth.make_pairs(2, 300, first_item=80)
        tgt_short = np.ascontiguousarray(q["tgt"][:, :, :257])
        o2 = VM.vcrnetIter(net_w, T(q["src"]), T(tgt_short), iter=1)
        r = synth.make_pairs(1, 331, partial=True, first_item=90)
        o3 = VM.vcrnetIter(net_p, T(r["src"]), T(r["tgt"]), iter=2)
    save("vcrnet_ragged",
         corrK_333=N_(o1[1]), R_333=N_(o1[2]), t_333=N_(o1[3]),
         corrK_300_257=N_(o2[1]), R_300_257=N_(o2[2]), t_300_257=N_(o2[3]),
         srcK_p331=N_(o3[0]), corrK_p331=N_(o3[1]), R_p331=N_(o3[2]), t_p331=N_(o3[3]), overlap2=np.array(OV2))


def make_lpd_train(ref=None, lpd_sd=None):
    """BASELINE config 3 pin (LPD pre-training forward + backward, model/lpdnet_model.py:149-229 under autograd):
    parameter gradients of (a) a fixed linear functional of the LPDNet embedding and (b) the LPD loss, from the live
    reference, together with the neighbour sets the reference used (so a test can inject them)."""
    torch.set_num_threads(1)
    if ref is None:
        ref = ref_harness.import_reference()
        lpd_sd = torch.load(os.path.join(ref_harness.REF_ROOT, "pretrained", "lpd-pretrained.t7"),
                            map_location="cpu", weights_only=True)
    U, LP = ref.util, ref.lpdnet_model
    pa = synth.make_pairs(2, 256, aligned=True, first_item=60)
    largs = ref_harness.default_args(model="lpd", num_points=256)
    lnet = LP.LPD(largs).train()
    lnet.load_state_dict(lpd_sd, strict=True)
    emb = lnet.emb_nn
    both = T(np.concatenate([pa["src"], pa["tgt"]], axis=0))               # [4,3,256]
    with torch.no_grad():
        h = F.leaky_relu(emb.conv2_lpd(F.leaky_relu(emb.conv1_lpd(both), 0.2)), 0.2)
        idx_feat = U.knn(h, 20)
        idx_xyz = U.knn(both, 20)
    # (a) LPDNet level: L = sum(emb(x) * Rw)
    rng = np.random.RandomState(7)
    Rw = rng.standard_normal((4, 512, 256)).astype(np.float32)
    lnet.zero_grad()
    out = emb(both)
    # The pre-trained conv3_lpd has dead output channels (|weights| ~ 1e-11) whose pre-activations are ~1e-10: the
    # sign there, hence LeakyReLU', is summation-order noise.  The functional is supported away from that kink.
    keep = (out.detach().abs() > 1e-6)
    Rw = Rw * N_(keep).astype(np.float32)
    (out * T(Rw)).sum().backward()
    ga = {"ga." + k: N_(p.grad).copy() for k, p in emb.named_parameters()}
    # (b) LPD level: the pre-training loss
    lnet.zero_grad()
    se, te, loss, mse, mae = lnet(T(pa["src"]), T(pa["tgt"]))
    loss.backward()
    gb = {"gb." + k: N_(p.grad).copy() for k, p in emb.named_parameters()}
    # Rw is reproducible from RandomState(7); the embedding is stored subsampled (every 16th point)
    save("lpd_train", src=pa["src"], tgt=pa["tgt"], idx_feat=N_(idx_feat).astype(np.int32),
         idx_xyz=N_(idx_xyz).astype(np.int32), emb_sub=N_(out)[:, :, ::16].copy(),
         Rw_keep_bits=np.packbits(N_(keep).reshape(-1)), loss=N_(loss), mse=N_(mse), mae=N_(mae), **ga, **gb)


def make_variants(ref=None):
    """the flag-selectable embeddings (--emb_nn dgcnn | pointnet) and heads
    (--vcp_nn att | dist), eval mode, from the live reference.  BatchNorm running statistics and affine parameters are
    randomised so that folding them into the convs is actually exercised."""
    torch.set_num_threads(1)
    if ref is None:
        ref = ref_harness.import_reference()
    VM, U = ref.vcrnet_model, ref.util
    torch.manual_seed(11)
    rng = np.random.RandomState(11)
    out = {}

    def randomise_bn(net):
        for mod in net.modules():
            if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                n = mod.num_features
                mod.running_mean.copy_(T(rng.normal(0, 0.2, n).astype(np.float32)))
                mod.running_var.copy_(T(rng.uniform(0.5, 1.5, n).astype(np.float32)))
                mod.weight.data.copy_(T(rng.uniform(0.5, 1.5, n).astype(np.float32)))
                mod.bias.data.copy_(T(rng.normal(0, 0.1, n).astype(np.float32)))

    pa = synth.make_pairs(2, 256, first_item=90)
    x = pa["src"]
    with torch.no_grad():
        for name, cls in (("dgcnn", VM.DGCNN), ("pointnet", VM.PointNet)):
            net = cls(emb_dims=128).eval()
            randomise_bn(net)
            for k, v in net.state_dict().items():
                if "num_batches_tracked" not in k:
                    out[f"{name}.{k}"] = N_(v)
            out[f"{name}.out"] = N_(net(T(x)))
        out["x"] = x
        out["idx"] = N_(U.knn(T(x), 20)).astype(np.int32)
        # heads on 64-d embeddings of 128 points
        se = rng.standard_normal((2, 64, 128)).astype(np.float32)
        te = rng.standard_normal((2, 64, 128)).astype(np.float32)
        ph = synth.make_pairs(2, 128, first_item=95)
        args = ref_harness.default_args(emb_dims=64)
        att = VM.VcpAtt(args).eval()
        for k, v in att.state_dict().items():
            out[f"att.{k}"] = N_(v)
        out.update(h_src_emb=se, h_tgt_emb=te, h_src=ph["src"], h_tgt=ph["tgt"],
                   att_corr=N_(att(T(se), T(te), T(ph["src"]), T(ph["tgt"]))[1]),
                   dist_corr=N_(VM.VcpByDis(args)(T(se), T(te), T(ph["src"]), T(ph["tgt"]))[1]))
    # ICP (model/icp_model.py) on a mildly perturbed copy: converges in a few iterations, exercising the early break
    IM = ref.icp_model if hasattr(ref, "icp_model") else __import__("model.icp_model", fromlist=["ICP"])
    pi = synth.make_pairs(3, 256, first_item=120)
    from scipy.spatial.transform import Rotation
    Rs = Rotation.from_euler("zyx", [[0.08, -0.05, 0.06], [0.02, 0.1, -0.07], [0.0, 0.0, 0.0]]).as_matrix().astype(np.float32)
    ts = np.array([[0.03, -0.02, 0.01], [-0.04, 0.0, 0.02], [0.0, 0.0, 0.0]], dtype=np.float32)
    icp_src = pi["src"]
    perm = np.random.RandomState(5).permutation(256)
    icp_dst = (np.matmul(Rs, icp_src) + ts[:, :, None])[:, :, perm].astype(np.float32)
    with torch.no_grad():
        for mi in (10, 2):
            o = IM.ICP(max_iterations=mi)(T(icp_src), T(icp_dst))
            out.update({f"icp{mi}_src": N_(o[1]), f"icp{mi}_R": N_(o[2]), f"icp{mi}_t": N_(o[3]),
                        f"icp{mi}_R_ba": N_(o[4]), f"icp{mi}_t_ba": N_(o[5])})
    out.update(icp_in_src=icp_src, icp_in_dst=icp_dst)
    save("variants", **out)


def make_tnet(ref=None):
    """LPDNet with --t3d / --tfea TranformNets (model/lpdnet_model.py:19-70, 107-118), eval
    mode, from the live reference.  BatchNorm statistics are randomised so the folding is exercised; the neighbour sets
    the reference's forward used are recomputed with the reference's own knn on the reference's intermediates."""
    torch.set_num_threads(1)
    if ref is None:
        ref = ref_harness.import_reference()
    LP, U = ref.lpdnet_model, ref.util
    out = {}
    x = synth.make_pairs(2, 256, first_item=140)["src"]
    with torch.no_grad():
        for name, t3d, tfea in (("both", True, True), ("t3d", True, False)):
            net = LP.LPDNet(ref_harness.default_args(t3d=t3d, tfea=tfea, emb_dims=128), negative_slope=0.0).eval()
            sd = synth.make_tnet_lpdnet_weights(21, t3d, tfea, 128)        # weights are reproducible, not stored
            ref_keys = [k for k in net.state_dict().keys() if "num_batches_tracked" not in k]
            assert ref_keys == list(sd.keys()), "synthetic TranformNet key order != reference"
            net.load_state_dict(synth.checkpoint_to_torch(sd), strict=False)
            out[f"{name}.keys"] = np.array(list(net.state_dict().keys()))
            xt = T(x)
            trans = net.t_net3d(xt)
            h = torch.bmm(xt.transpose(2, 1), trans).transpose(2, 1)
            h = F.leaky_relu(net.conv2_lpd(F.leaky_relu(net.conv1_lpd(h), 0.0)), 0.0)
            out[f"{name}.trans"] = N_(trans)
            if tfea:
                tf = net.t_net_fea(h)
                h = torch.bmm(h.transpose(2, 1), tf).transpose(2, 1)
                out[f"{name}.trans_feat"] = N_(tf)
            out[f"{name}.idx_feat"] = N_(U.knn(h.contiguous(), 20)).astype(np.int32)
            out[f"{name}.out"] = N_(net(xt))
        out["x"] = x
        out["idx_xyz"] = N_(U.knn(T(x), 20)).astype(np.int32)
    save("tnet", **out)


def _loader_from_pairs(p, batch):
    """Batches shaped like the reference's DataLoader over ModelNet40.__getitem__ (util/data.py:296-318)."""
    n = p["src"].shape[0]
    R_ba = np.transpose(p["R_ab"], (0, 2, 1)).copy()
    t_ba = -np.einsum("pij,pj->pi", R_ba, p["t_ab"]).astype(np.float32)
    e_ba = -p["euler_ab"][:, ::-1].copy()
    out = []
    for b0 in range(0, n, batch):
        s = slice(b0, min(n, b0 + batch))
        out.append((T(p["src"][s]), T(p["tgt"][s]), T(p["R_ab"][s]), T(p["t_ab"][s]), T(R_ba[s]), T(t_ba[s]),
                    T(p["euler_ab"][s]), T(e_ba[s]), torch.zeros(s.stop - s.start, dtype=torch.int64)))
    return out


def aggregate_metrics(ref, res):
    """testVCRNet's printed aggregates (model/vcrnet_model.py:776-806) from test_one_epoch's return tuple, through the
    reference's own npmat2euler."""
    (loss_pose, cycle, mse_ab, mae_ab, mse_ba, mae_ba, R_ab, t_ab, R_ab_p, t_ab_p, R_ba, t_ba, R_ba_p, t_ba_p,
     e_ab, e_ba, loss_vcr) = res
    eul = ref.util.npmat2euler(R_ab_p)
    r_mse = np.mean((eul - np.degrees(e_ab)) ** 2)
    t_mse = np.mean((t_ab - t_ab_p) ** 2)
    eul_ba = ref.util.npmat2euler(R_ba_p, 'xyz')
    r_mse_ba = np.mean((eul_ba - np.degrees(e_ba)) ** 2)
    t_mse_ba = np.mean((t_ba - t_ba_p) ** 2)
    return dict(loss=loss_vcr, loss_pose=loss_pose, cycle_loss=cycle, mse_ab=mse_ab, rmse_ab=np.sqrt(mse_ab), mae_ab=mae_ab,
                mse_ba=mse_ba, mae_ba=mae_ba,
                r_mse_ab=r_mse, r_rmse_ab=np.sqrt(r_mse), r_mae_ab=np.mean(np.abs(eul - np.degrees(e_ab))),
                t_mse_ab=t_mse, t_rmse_ab=np.sqrt(t_mse), t_mae_ab=np.mean(np.abs(t_ab - t_ab_p)),
                r_mse_ba=r_mse_ba, r_mae_ba=np.mean(np.abs(eul_ba - np.degrees(e_ba))),
                t_mse_ba=t_mse_ba, t_mae_ba=np.mean(np.abs(t_ba - t_ba_p)))


HEADLINE = {   # BASELINE.json configs[0] / configs[1] at their own sizes; 48 items = the staged synthetic test partition
    "cfg1": dict(partial=False, batch=16, num_points=1024, iters=1, n_pairs=48),
    "cfg2": dict(partial=True, batch=24, num_points=1024, iters=3, n_pairs=48),
}


def make_headline(ref=None, which=("cfg1", "cfg2"), threads=1):
    """The benchmark's own configurations pinned from the LIVE reference: the reference's ``test_one_epoch``
    (model/vcrnet_model.py:521-649, run on the host through ref_harness.cpu_mode) over the 48-item synthetic test partition,
    in fp32 (the pin) and again with the network and inputs in fp64 (``net.double()``: the same reference code, a numerical
    yardstick for how far fp32 rounding alone moves each pair).  Stored per config: R/t of every pair (fp32 + fp64), the
    first batch's srcK / src_corrK, and the aggregate metrics testVCRNet prints."""
    torch.set_num_threads(threads)
    ref = ref or ref_harness.import_reference()
    VM = ref.vcrnet_model
    lpd = dict(np.load(os.path.join(GOLD, "lpd_pretrained_weights.npz")))
    sd_t = synth.checkpoint_to_torch(synth.make_checkpoint(1234, emb_weights=lpd))
    for name in which:
        c = HEADLINE[name]
        ov2 = OV2 if c["partial"] else 0.75
        args = ref_harness.default_args(partial=c["partial"], overlap2=ov2, iter=c["iters"], num_points=c["num_points"])
        p = synth.make_pairs(c["n_pairs"], c["num_points"], partial=c["partial"])
        out = {}
        for tag, dt in (("", torch.float32), ("64", torch.float64)):
            net = VM.VCRNet(args).eval()
            net.load_state_dict(sd_t, strict=True)
            loader = _loader_from_pairs(p, c["batch"])
            if dt == torch.float64:
                net = net.double()
                loader = [tuple(x.double() if x.is_floating_point() else x for x in b) for b in loader]
            with ref_harness.cpu_mode(), torch.no_grad():
                res = VM.test_one_epoch(args, net, loader)
                first = VM.vcrnetIter(net, loader[0][0], loader[0][1], iter=c["iters"])
                it1 = VM.vcrnetIter(net, loader[0][0], loader[0][1], iter=1) if c["iters"] > 1 else None
            m = aggregate_metrics(ref, res)
            out.update({f"R_ab{tag}": res[8], f"t_ab{tag}": res[9], f"R_ba{tag}": res[12], f"t_ba{tag}": res[13],
                        f"srcK{tag}": N_(first[0]), f"corrK{tag}": N_(first[1])})
            if it1 is not None:      # first refinement iteration of the first batch: srcK are ORIGINAL source points, so the
                out.update({f"srcK_it1{tag}": N_(it1[0]), f"corrK_it1{tag}": N_(it1[1]),       # selections compare as sets
                            f"R_ab_it1{tag}": N_(it1[2]), f"t_ab_it1{tag}": N_(it1[3])})
            out.update({f"m{tag}.{k}": np.float64(v) for k, v in m.items()})
            print(name, tag or "32", {k: float(f"{v:.6g}") for k, v in m.items()})
        d = np.abs(out["R_ab"] - out["R_ab64"]).reshape(c["n_pairs"], -1).max(1)
        print(name, "fp32 vs fp64 reference: max|dR| per pair: median %.3g max %.3g" % (np.median(d), d.max()))
        save("headline_" + name, R_gt=p["R_ab"], t_gt=p["t_ab"], euler_gt=p["euler_ab"], overlap2=np.array(ov2),
             batch=np.array(c["batch"]), iters=np.array(c["iters"]), num_points=np.array(c["num_points"]),
             **out)


def make_identity(ref=None):
    """--pointer identity (model/vcrnet_model.py:477-478, 502-505): Identity returns its inputs, so both embeddings are
    doubled before the head.  Whole-to-whole, 2 pairs x 256 points, from the live reference."""
    torch.set_num_threads(1)
    ref = ref or ref_harness.import_reference()
    VM = ref.vcrnet_model
    lpd = dict(np.load(os.path.join(GOLD, "lpd_pretrained_weights.npz")))
    sd_t = synth.checkpoint_to_torch(synth.make_checkpoint(1234, emb_weights=lpd))
    net = VM.VCRNet(ref_harness.default_args(pointer="identity")).eval()
    missing = net.load_state_dict(sd_t, strict=False)
    assert not missing.missing_keys, missing
    p = synth.make_pairs(2, 256, first_item=150)
    with torch.no_grad():
        o = VM.vcrnetIter(net, T(p["src"]), T(p["tgt"]), iter=1)
    save("vcrnet_identity_pointer", src=p["src"], tgt=p["tgt"], corrK=N_(o[1]), R_ab=N_(o[2]), t_ab=N_(o[3]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        make_variants()
    elif len(sys.argv) > 1 and sys.argv[1] == "tnet":
        make_tnet()
    elif len(sys.argv) > 1 and sys.argv[1] == "ragged":
        make_ragged()
    elif len(sys.argv) > 1 and sys.argv[1] == "identity":
        make_identity()
    elif len(sys.argv) > 1 and sys.argv[1] == "headline":
        make_headline(which=tuple(sys.argv[2:]) or ("cfg1", "cfg2"))
    elif len(sys.argv) > 1 and sys.argv[1] == "lpd_train":
        make_lpd_train()
    else:
        main()
