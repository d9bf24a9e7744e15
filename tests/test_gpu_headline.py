"""GPU parity on BASELINE.json's OWN configurations (VERDICT r1 "missing #1"):

  cfg 1  whole-to-whole, 1024 pts, batch 16, iter 1     (tests/golden/headline_cfg1.npz)
  cfg 2  partial-to-partial, 768 of 1024 pts, batch 24, iter 3   (tests/golden/headline_cfg2.npz)

The fixtures come from the LIVE reference's own ``test_one_epoch`` (model/vcrnet_model.py:521-649) over the 48-item synthetic
test partition (oracle/make_golden.make_headline), once in fp32 -- the pin -- and once with the same reference code in
fp64 -- a yardstick for how far fp32 rounding ALONE moves a pair through the network's hard selections (kNN sets, top-K
keys / points, arg-max correspondences).

What is asserted, and why in this form:
  * per pair, R / t against the fp32 reference, at the tightest bar that holds (ACHIEVED errors are printed and written
    to gpurun_out/parity_headline_*.json, copied to profiles/);
  * per pair, the distance to the fp64 reference is no larger than a small multiple of the fp32 reference's OWN distance
    to it -- the statement "as close to exact arithmetic as the reference is" that stays meaningful where a flipped
    near-tie makes per-pair fp32-vs-fp32 comparison chaotic (VERDICT r1 weak #1c);
  * the aggregate metrics testVCRNet prints (rot / trans MSE, RMSE, MAE, point MSE / MAE, :776-806) within a stated relative
    bound of the reference's;
  * partial path: Jaccard overlap of the selected source sets and agreement of the hard correspondences at iteration 1.
``rel_err`` is max-abs error over the tensor's max-abs value (conftest.py), not element-wise relative error.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import vcr_net_b200 as V
    from vcr_net_b200.data import EvalAccumulator
    from vcr_net_b200.util.util import npmat2euler
from oracle import synth
from oracle.ref_harness import default_args

DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def nump(t):
    return t.detach().cpu().numpy()


@pytest.fixture(params=["fp32", "h3"])
def precision(request):
    from vcr_net_b200 import config
    old = config.precision
    config.set_precision(request.param)
    yield request.param
    config.set_precision(old)


def _net(ckpt, partial):
    net = V.VCRNet(default_args(partial=partial, overlap2=synth.OVERLAP2_0575 if partial else 0.75)).to(DEV).eval()
    net.load_state_dict(synth.checkpoint_to_torch(ckpt), strict=True)
    return net


def _run_epoch(net, g, partial):
    """The reference's test loop shape (test_one_epoch): batches of g['batch'] over the 48 items, vcrnetIter per batch,
    running metrics; returns per-pair poses and the aggregates of testVCRNet."""
    B, iters, N = int(g["batch"]), int(g["iters"]), int(g["num_points"])
    n = g["R_gt"].shape[0]
    p = synth.make_pairs(n, N, partial=partial)
    assert np.array_equal(p["R_ab"], g["R_gt"]) and np.array_equal(p["t_ab"], g["t_gt"])
    acc = EvalAccumulator(DEV)
    Rs, ts, Rbs, tbs = [], [], [], []
    first = None
    for b0 in range(0, n, B):
        s = slice(b0, b0 + B)
        src, tgt = cu(p["src"][s]), cu(p["tgt"][s])
        out = V.vcrnetIter(net, src, tgt, iter=iters)
        if first is None:
            first = [nump(o) for o in out]
        acc.update(src, tgt, out[0], out[1], cu(p["R_ab"][s]), cu(p["t_ab"][s]), out[2], out[3], out[4], out[5])
        Rs.append(nump(out[2])); ts.append(nump(out[3])); Rbs.append(nump(out[4])); tbs.append(nump(out[5]))
    R, t, Rb, tb = (np.concatenate(x) for x in (Rs, ts, Rbs, tbs))
    res = acc.result()
    e_ab = p["euler_ab"]
    e_ba = -e_ab[:, ::-1]
    t_ba_gt = -np.einsum("pji,pj->pi", p["R_ab"], p["t_ab"])
    eul, eul_ba = npmat2euler(R), npmat2euler(Rb, "xyz")
    m = dict(loss=res["mse_ab"], loss_pose=res["loss"], mse_ab=res["mse_ab"], rmse_ab=np.sqrt(res["mse_ab"]),
             mae_ab=res["mae_ab"], mse_ba=res["mse_ba"], mae_ba=res["mae_ba"],
             r_mse_ab=np.mean((eul - np.degrees(e_ab)) ** 2), r_mae_ab=np.mean(np.abs(eul - np.degrees(e_ab))),
             t_mse_ab=np.mean((p["t_ab"] - t) ** 2), t_mae_ab=np.mean(np.abs(p["t_ab"] - t)),
             r_mse_ba=np.mean((eul_ba - np.degrees(e_ba)) ** 2), r_mae_ba=np.mean(np.abs(eul_ba - np.degrees(e_ba))),
             t_mse_ba=np.mean((t_ba_gt - tb) ** 2), t_mae_ba=np.mean(np.abs(t_ba_gt - tb)))
    m["r_rmse_ab"], m["t_rmse_ab"] = np.sqrt(m["r_mse_ab"]), np.sqrt(m["t_mse_ab"])
    return p, R, t, Rb, tb, first, {k: float(v) for k, v in m.items()}


def _pair_err(a, b):
    """max-abs difference per pair."""
    n = a.shape[0]
    return np.abs(a.astype(np.float64) - b.astype(np.float64)).reshape(n, -1).max(axis=1)


def _report(name, rec):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"parity_headline_{name}.json"), "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    print("\n[parity achieved] " + name + " " + json.dumps(rec, sort_keys=True))


def _metric_errs(m, g):
    errs = {}
    for k, v in m.items():
        ref = float(g["m." + k])
        errs[k] = abs(v - ref) / max(abs(ref), 1e-12)
    return errs


def _stats(x):
    return {"median": float(np.median(x)), "p90": float(np.quantile(x, 0.9)), "max": float(np.max(x)),
            "frac_pairs_below_1e-4": float(np.mean(np.asarray(x) < 1e-4))}


# bars: set from the achieved numbers of the committed run (profiles/r02_parity_headline_*.json), with ~2x slack
CFG1_R_MAX, CFG1_T_MAX = 1.5e-3, 1.5e-3            # worst pair (a flipped near-tie neighbour / a re-ordered key)
CFG1_R_MEDIAN = 1e-4                            # north_star's bar holds for the typical pair
CFG1_METRIC_REL = 5e-4                          # achieved <= 1.2e-4 (the fp32 and fp64 reference differ by 2.3e-4)
CFG2_METRIC_REL = 2e-2                          # achieved <= 7.4e-3 (the fp32 and fp64 reference differ by 7.3e-3)
YARDSTICK = 4.0                                 # ours-vs-fp64 <= YARDSTICK * (fp32 reference vs fp64), on median and p90


def test_cfg1_whole_batch16_n1024_vs_live_reference(ckpt, precision, capsys, product_defaults):
    g = load_golden("headline_cfg1")
    net = _net(ckpt, False)
    p, R, t, Rb, tb, first, m = _run_epoch(net, g, False)
    eR, et = _pair_err(R, g["R_ab"]), _pair_err(t, g["t_ab"])
    eR64, et64 = _pair_err(R, g["R_ab64"]), _pair_err(t, g["t_ab64"])
    rR64, rt64 = _pair_err(g["R_ab"], g["R_ab64"]), _pair_err(g["t_ab"], g["t_ab64"])
    corr = np.abs(first[1] - g["corrK"]).reshape(first[1].shape[0], -1).max(axis=1) / np.abs(g["corrK"]).max()
    merr = _metric_errs(m, g)
    rec = {"precision": precision, "pairs": int(R.shape[0]), "batch": int(g["batch"]),
           "R_vs_ref32": _stats(eR), "t_vs_ref32": _stats(et), "R_vs_ref64": _stats(eR64), "t_vs_ref64": _stats(et64),
           "ref32_vs_ref64_R": _stats(rR64), "ref32_vs_ref64_t": _stats(rt64), "corrK_rel_first_batch": _stats(corr),
           "R_ba_vs_ref32": _stats(_pair_err(Rb, g["R_ba"])), "t_ba_vs_ref32": _stats(_pair_err(tb, g["t_ba"])),
           "metrics": m, "metrics_rel_err": merr}
    with capsys.disabled():
        _report(f"cfg1_{precision}", rec)
    assert np.array_equal(first[0], p["src"][:int(g["batch"])])                 # whole-to-whole: srcK is the source cloud
    assert eR.max() < CFG1_R_MAX and et.max() < CFG1_T_MAX, (eR.max(), et.max())
    assert np.median(eR) < CFG1_R_MEDIAN and np.median(et) < CFG1_R_MEDIAN, (np.median(eR), np.median(et))
    for q in (0.5, 0.9):
        assert np.quantile(eR64, q) <= YARDSTICK * np.quantile(rR64, q) + 1e-6, (q, np.quantile(eR64, q), np.quantile(rR64, q))
        assert np.quantile(et64, q) <= YARDSTICK * np.quantile(rt64, q) + 1e-6, (q, np.quantile(et64, q), np.quantile(rt64, q))
    bad = {k: v for k, v in merr.items() if v > CFG1_METRIC_REL}
    assert not bad, bad


def _rows_to_index(pts, cloud):
    """pts [3,M] are exact copies of columns of cloud [3,N] -> their column indices."""
    key = {tuple(c): i for i, c in enumerate(cloud.T.tolist())}
    return np.array([key[tuple(c)] for c in pts.T.tolist()])


def test_cfg2_partial_batch24_iter3_vs_live_reference(ckpt, precision, capsys, product_defaults):
    g = load_golden("headline_cfg2")
    net = _net(ckpt, True)
    p, R, t, Rb, tb, first, m = _run_epoch(net, g, True)
    assert p["src"].shape[2] == 768
    eR, et = _pair_err(R, g["R_ab"]), _pair_err(t, g["t_ab"])
    eR64, et64 = _pair_err(R, g["R_ab64"]), _pair_err(t, g["t_ab64"])
    rR64, rt64 = _pair_err(g["R_ab"], g["R_ab64"]), _pair_err(g["t_ab"], g["t_ab64"])
    merr = _metric_errs(m, g)
    # iteration 1 of the first batch: srcK are original source points, src_corrK original target points
    B = int(g["batch"])
    o1 = V.vcrnetIter(net, cu(p["src"][:B]), cu(p["tgt"][:B]), iter=1)
    sK, cK = nump(o1[0]), nump(o1[1])
    jac, agree = [], []
    for b in range(B):
        ours_s, ours_c = _rows_to_index(sK[b], p["src"][b]), _rows_to_index(cK[b], p["tgt"][b])
        ref_s, ref_c = _rows_to_index(g["srcK_it1"][b], p["src"][b]), _rows_to_index(g["corrK_it1"][b], p["tgt"][b])
        a, r = dict(zip(ours_s.tolist(), ours_c.tolist())), dict(zip(ref_s.tolist(), ref_c.tolist()))
        common = set(a) & set(r)
        jac.append(len(common) / len(set(a) | set(r)))
        agree.append(np.mean([a[i] == r[i] for i in common]) if common else 0.0)
    e1R = _pair_err(nump(o1[2]), g["R_ab_it1"])
    rec = {"precision": precision, "pairs": int(R.shape[0]), "batch": B, "iter": int(g["iters"]),
           "R_vs_ref32": _stats(eR), "t_vs_ref32": _stats(et), "R_vs_ref64": _stats(eR64), "t_vs_ref64": _stats(et64),
           "ref32_vs_ref64_R": _stats(rR64), "ref32_vs_ref64_t": _stats(rt64),
           "iter1_R_vs_ref32": _stats(e1R), "iter1_selected_src_jaccard": {"min": float(min(jac)), "median": float(np.median(jac))},
           "iter1_correspondence_agreement_on_common": {"min": float(min(agree)), "median": float(np.median(agree))},
           "metrics": m, "metrics_rel_err": merr}
    with capsys.disabled():
        _report(f"cfg2_{precision}", rec)
    assert tuple(first[0].shape) == g["srcK"].shape
    # iteration 1 (no chaos amplified yet): same selected sets, same hard correspondences, same pose
    assert min(jac) >= 0.98 and np.median(jac) >= 0.99, (min(jac), np.median(jac))
    assert min(agree) >= 0.99, min(agree)
    assert np.mean(e1R < 1e-4) >= 0.9, np.mean(e1R < 1e-4)
    # after 3 iterations the per-pair error is BIMODAL (a pair either keeps every hard selection, error ~1e-6, or flips
    # one and lands ~1e-2 away): the fp32 reference itself agrees with its fp64 run on only 2/3 of these 48 pairs.  Bars:
    # the share of agreeing pairs is within 0.25 of the reference's own, the tail is within YARDSTICK of the reference's.
    ref_share = np.mean(rR64 < 1e-4)
    assert np.mean(eR64 < 1e-4) >= ref_share - 0.25 and np.mean(eR < 1e-4) >= ref_share - 0.25, (np.mean(eR64 < 1e-4), ref_share)
    assert np.quantile(eR64, 0.9) <= YARDSTICK * np.quantile(rR64, 0.9) + 1e-5, (np.quantile(eR64, 0.9), np.quantile(rR64, 0.9))
    assert np.quantile(et64, 0.9) <= YARDSTICK * np.quantile(rt64, 0.9) + 1e-5, (np.quantile(et64, 0.9), np.quantile(rt64, 0.9))
    assert eR64.max() <= YARDSTICK * rR64.max() and et64.max() <= YARDSTICK * rt64.max(), (eR64.max(), rR64.max())
    bad = {k: v for k, v in merr.items() if v > CFG2_METRIC_REL}
    assert not bad, bad
