"""Synthetic inputs and a synthetic 59-key checkpoint for the VCR-Net registration path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``vcr_net_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (as data generator / checker / cpu_baseline) may use it.

What it mirrors (reference file:line, relative to /root/reference):
  * pair generation            util/data.py:255-303 (euler U(0, pi/factor), R = Rx.Ry.Rz,
                               t ~ U(-0.5, 0.5)^3, independent permutations of both clouds)
  * nearest-to-anchor crop     util/data.py:320-329 (keep int(N*reserve) nearest to the LAST point)
  * base clouds                model/icp_model.py:124 (np.random.rand(n, 3) - 0.5)
  * reserve / overlap2         util/initPara.py:115-124 (cubic solved there with sympy; the
                               constants below are its root for --overlap=0.575)
  * checkpoint key layout      SURVEY.md section 8b (59 keys / 5 625 161 params)

Everything is numpy + RandomState so the same bytes come out on any box.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

# util/initPara.py:115-124 evaluated for --overlap=0.575 (values from SURVEY.md section 5)
RESERVE_0575 = 0.750681278
OVERLAP2_0575 = 0.765970881


def solve_reserve(overlap: float) -> tuple[float, float]:
    """Root of the reference's cubic (util/initPara.py:115-124) by bisection, no sympy.

    f(n) = ((a+b)*2 + (1-2n)^3)/(1-n)^2 - overlap on n in [0, 0.5]; reserve = 1-n,
    overlap2 = overlap/reserve.
    """

    def f(n: float) -> float:
        a = (n - 1.5 * n * n) * (1.0 - 2.0 * n)
        b = 0.5 * (n - 1.0) ** 2 * n - (1.0 - n) ** 3 / 6.0 + (1.0 - 2.0 * n) ** 3 / 6.0
        return ((a + b) * 2.0 + (1.0 - 2.0 * n) ** 3) / (1.0 - n) ** 2 - overlap

    lo, hi = 0.0, 0.5
    flo = f(lo)
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        fm = f(mid)
        if (fm > 0) == (flo > 0):
            lo, flo = mid, fm
        else:
            hi = mid
    n = 0.5 * (lo + hi)
    reserve = 1.0 - n
    return reserve, overlap / reserve


def euler_to_R(anglex: float, angley: float, anglez: float) -> np.ndarray:
    """R = Rx.Ry.Rz as in util/data.py:262-277 (float64)."""
    cx, cy, cz = math.cos(anglex), math.cos(angley), math.cos(anglez)
    sx, sy, sz = math.sin(anglex), math.sin(angley), math.sin(anglez)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float64)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float64)
    return Rx @ Ry @ Rz


def crop_nearest_to_last(pc: np.ndarray, reserve: float) -> np.ndarray:
    """pc [3, N] -> [3, int(N*reserve)] nearest neighbours of the last point, nearest first.

    util/data.py:320-329 uses sklearn NearestNeighbors; a stable argsort of the squared
    distance gives the same ordered set (ties broken by index) without the dependency.
    """
    pts = pc.T
    keep = int(max(pts.shape) * reserve)
    d2 = ((pts - pts[-1:]) ** 2).sum(axis=1)
    order = np.argsort(d2, kind="stable")[:keep]
    return pts[order].T


def make_pairs(n_pairs: int, num_points: int = 1024, *, partial: bool = False,
               reserve: float = RESERVE_0575, factor: float = 4.0, base_points: int = 2048,
               seed: int = 1234, first_item: int = 0, aligned: bool = False):
    """ModelNet40-shaped synthetic pairs.

    Returns dict of float32 arrays: src [P,3,M], tgt [P,3,M], R_ab [P,3,3], t_ab [P,3],
    euler_ab [P,3] (zyx order as util/data.py:293).  M = num_points, or
    int(num_points*reserve) when ``partial``.  ``aligned`` reproduces the --model=lpd
    branch (util/data.py:304-309): one shared permutation, src/tgt stay index-aligned.
    """
    base_points = max(base_points, num_points)
    base = np.random.RandomState(seed).rand(first_item + n_pairs, base_points, 3).astype(np.float32) - 0.5
    srcs, tgts, Rs, ts, eulers = [], [], [], [], []
    for item in range(first_item, first_item + n_pairs):
        rs = np.random.RandomState(item)           # util/data.py:255-256 np.random.seed(item)
        ax = rs.uniform() * math.pi / factor
        ay = rs.uniform() * math.pi / factor
        az = rs.uniform() * math.pi / factor
        R = euler_to_R(ax, ay, az)
        t = np.array([rs.uniform(-0.5, 0.5), rs.uniform(-0.5, 0.5), rs.uniform(-0.5, 0.5)])
        pc1 = rs.permutation(base[item])[:num_points].T.astype(np.float64)       # [3,N]
        # Rotation.from_euler('zyx',[az,ay,ax]).apply(p) == (Rx.Ry.Rz ... ) see note below
        pc2 = _apply_zyx(pc1, az, ay, ax) + t[:, None]
        if aligned:
            perm = rs.permutation(num_points)
            pc1, pc2 = pc1[:, perm], pc2[:, perm]
        else:
            pc1 = pc1[:, rs.permutation(num_points)]
            if partial:
                pc1 = crop_nearest_to_last(pc1, reserve)
            pc2 = pc2[:, rs.permutation(num_points)]
            if partial:
                pc2 = crop_nearest_to_last(pc2, reserve)
        srcs.append(pc1.astype(np.float32))
        tgts.append(pc2.astype(np.float32))
        Rs.append(R.astype(np.float32))
        ts.append(t.astype(np.float32))
        eulers.append(np.array([az, ay, ax], dtype=np.float32))
    return {
        "src": np.stack(srcs), "tgt": np.stack(tgts), "R_ab": np.stack(Rs),
        "t_ab": np.stack(ts), "euler_ab": np.stack(eulers),
    }


def _apply_zyx(pc: np.ndarray, az: float, ay: float, ax: float) -> np.ndarray:
    """scipy Rotation.from_euler('zyx', [az, ay, ax]).apply(pc.T).T without scipy.

    Lower-case 'zyx' is extrinsic: rotate about z, then y, then x, i.e. the matrix
    Rx @ Ry @ Rz -- the same R_ab the dataset returns (util/data.py:278,290-291).
    """
    return euler_to_R(ax, ay, az) @ pc


def grid_cloud(rs: np.random.RandomState, shape, frac_bits: int, lo: float, hi: float) -> np.ndarray:
    """Dyadic-rational values k / 2**frac_bits in [lo, hi): every fp32 product/sum of a few
    hundred of them is exact, so distance values are independent of accumulation order
    and ties are real (exact-arithmetic kNN/FPS inputs)."""
    scale = float(1 << frac_bits)
    return (rs.randint(int(lo * scale), int(hi * scale), size=shape).astype(np.float32) / scale)


# --------------------------------------------------------------------------------------
# synthetic checkpoint
# --------------------------------------------------------------------------------------

def _linear_init(rs, out_f, in_f):
    """nn.Linear / nn.Conv default: kaiming_uniform(a=sqrt(5)) => U(-1/sqrt(fan_in), +)."""
    bound = 1.0 / math.sqrt(in_f)
    w = rs.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    b = rs.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    return w, b


def _mha_keys(prefix):
    return [f"{prefix}.linears.{i}" for i in range(4)]


def make_checkpoint(seed: int = 1234, emb_dims: int = 512, ff_dims: int = 1024,
                    emb_weights: dict | None = None) -> "OrderedDict[str, np.ndarray]":
    """59-key VCRNet state_dict as numpy arrays.

    Transformer / emb_nn weights are drawn from the nn.Linear default distribution with a
    numpy RandomState (portable); LayerNorm a_2=1, b_2=0; svd.reflect = diag(1,1,-1).
    ``emb_weights`` (the 12 ``emb_nn.*`` tensors of lpd-pretrained.t7, shipped as
    tests/golden/lpd_pretrained_weights.npz) overrides the random embedding weights.
    """
    rs = np.random.RandomState(seed)
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def conv(name, co, ci, extra):
        w, b = _linear_init(rs, co, ci)
        sd[f"{name}.weight"] = w.reshape((co, ci) + extra)
        sd[f"{name}.bias"] = b

    conv("emb_nn.convDG1.0", 128, 128, (1, 1))
    conv("emb_nn.convDG2.0", 128, 128, (1, 1))
    conv("emb_nn.convSN1.0", 256, 256, (1, 1))
    conv("emb_nn.conv1_lpd", 64, 3, (1,))
    conv("emb_nn.conv2_lpd", 64, 64, (1,))
    conv("emb_nn.conv3_lpd", emb_dims, 512, (1,))

    def lin(name, o, i):
        w, b = _linear_init(rs, o, i)
        sd[f"{name}.weight"] = w
        sd[f"{name}.bias"] = b

    def ln(name):
        sd[f"{name}.a_2"] = np.ones(emb_dims, np.float32)
        sd[f"{name}.b_2"] = np.zeros(emb_dims, np.float32)

    enc = "pointer.model.encoder"
    for k in _mha_keys(f"{enc}.layers.0.self_attn"):
        lin(k, emb_dims, emb_dims)
    lin(f"{enc}.layers.0.feed_forward.w_1", ff_dims, emb_dims)
    lin(f"{enc}.layers.0.feed_forward.w_2", emb_dims, ff_dims)
    ln(f"{enc}.layers.0.sublayer.0.norm")
    ln(f"{enc}.layers.0.sublayer.1.norm")
    ln(f"{enc}.norm")
    dec = "pointer.model.decoder"
    for k in _mha_keys(f"{dec}.layers.0.self_attn"):
        lin(k, emb_dims, emb_dims)
    for k in _mha_keys(f"{dec}.layers.0.src_attn"):
        lin(k, emb_dims, emb_dims)
    lin(f"{dec}.layers.0.feed_forward.w_1", ff_dims, emb_dims)
    lin(f"{dec}.layers.0.feed_forward.w_2", emb_dims, ff_dims)
    for i in range(3):
        ln(f"{dec}.layers.0.sublayer.{i}.norm")
    ln(f"{dec}.norm")
    sd["svd.reflect"] = np.diag([1.0, 1.0, -1.0]).astype(np.float32)

    if emb_weights is not None:
        for k, v in emb_weights.items():
            assert k in sd and sd[k].shape == v.shape, (k, v.shape)
            sd[k] = np.asarray(v, dtype=np.float32)
    return sd


def checkpoint_to_torch(sd):
    import torch
    return OrderedDict((k, torch.from_numpy(np.ascontiguousarray(v)).clone()) for k, v in sd.items())


def make_tnet_lpdnet_weights(seed: int, t3d: bool = True, tfea: bool = True, emb_dims: int = 128):
    """LPDNet(t3d, tfea) state_dict (model/lpdnet_model.py:19-42, 78-99 key order, num_batches_tracked omitted) from a
    numpy RandomState: nn default-style uniform weights, randomised BatchNorm1d statistics / affine parameters."""
    rs = np.random.RandomState(seed)
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def conv(name, co, ci, extra):
        w, b = _linear_init(rs, co, ci)
        sd[f"{name}.weight"] = w.reshape((co, ci) + extra)
        sd[f"{name}.bias"] = b

    conv("convDG1.0", 128, 128, (1, 1))
    conv("convDG2.0", 128, 128, (1, 1))
    conv("convSN1.0", 256, 256, (1, 1))
    conv("conv1_lpd", 64, 3, (1,))
    conv("conv2_lpd", 64, 64, (1,))
    conv("conv3_lpd", emb_dims, 512, (1,))
    for on, name, k in ((t3d, "t_net3d", 3), (tfea, "t_net_fea", 64)):
        if not on:
            continue
        conv(f"{name}.conv1", 64, k, (1,))
        conv(f"{name}.conv2", 128, 64, (1,))
        conv(f"{name}.conv3", 1024, 128, (1,))
        conv(f"{name}.fc1", 512, 1024, ())
        conv(f"{name}.fc2", 256, 512, ())
        conv(f"{name}.fc3", k * k, 256, ())
        for i, n in enumerate((64, 128, 1024, 512, 256), start=1):
            sd[f"{name}.bn{i}.weight"] = rs.uniform(0.5, 1.5, n).astype(np.float32)
            sd[f"{name}.bn{i}.bias"] = rs.normal(0, 0.1, n).astype(np.float32)
            sd[f"{name}.bn{i}.running_mean"] = rs.normal(0, 0.2, n).astype(np.float32)
            sd[f"{name}.bn{i}.running_var"] = rs.uniform(0.5, 1.5, n).astype(np.float32)
    return sd
