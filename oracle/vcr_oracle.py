"""CPU oracle: a numpy (float32) restatement of VCR-Net's registration inference path.

TEST INFRASTRUCTURE ONLY -- the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; ``vcr_net_b200`` must never do so (tests/test_layout.py
greps for it).

Parity pin: the reference ships no tests or golden vectors, so this
restatement is pinned against outputs of the LIVE reference (CPU torch, imported from
/root/reference by ``oracle/ref_harness.py``) stored under ``tests/golden/`` by
``oracle/make_golden.py``; ``tests/test_oracle_vs_golden.py`` re-checks every function
here against those vectors on every run.

Each function cites the reference lines it restates (paths relative to /root/reference).
Weights are passed as a flat ``dict[str, np.ndarray]`` using the reference's
``state_dict`` key names.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------

def _f32(a):
    return np.ascontiguousarray(a, dtype=F32)


def leaky_relu(x, slope):
    return np.where(x >= 0, x, x * F32(slope)).astype(F32)


def softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m, dtype=F32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=F32)).astype(F32)


def topk_desc(v, k, axis=-1):
    """Indices of the k largest along ``axis``, descending, ties -> lower index.

    torch.topk leaves tie order unspecified (hard part 1); the
    canonical rule used throughout this repo is "lower index wins"."""
    order = np.argsort(-v, axis=axis, kind="stable")
    return np.take(order, np.arange(k), axis=axis)


def conv1x1(x, w, b):
    """nn.Conv1d / nn.Conv2d with kernel 1.  x [B,Ci,...], w [Co,Ci,1(,1)], b [Co]."""
    w2 = w.reshape(w.shape[0], w.shape[1])
    lead = x.shape[2:]
    y = np.matmul(w2, x.reshape(x.shape[0], x.shape[1], -1))
    if b is not None:
        y = y + b.reshape(1, -1, 1)
    return y.reshape((x.shape[0], w2.shape[0]) + lead).astype(F32)


def linear(x, w, b):
    """nn.Linear: x [..., in], w [out, in]."""
    return (np.matmul(x, w.T) + b).astype(F32)


# --------------------------------------------------------------------------------------
# L1 point-cloud ops (util/util.py)
# --------------------------------------------------------------------------------------

def neg_sqdist_self(x):
    """util/util.py:153-158.  x [B,D,N] -> pd [B,N,N], pd_ij = (-xx_j - (-2 x_i.x_j)) - xx_i."""
    x = _f32(x)
    inner = F32(-2.0) * np.matmul(x.transpose(0, 2, 1), x)
    xx = np.sum(x * x, axis=1, keepdims=True, dtype=F32)          # [B,1,N]
    pd = -xx - inner
    pd = pd - xx.transpose(0, 2, 1)
    return pd.astype(F32)


def knn(x, k):
    """util/util.py:143-160: top-(k+1) of the negative squared distance, drop rank 0.

    Returns int64 [B,N,k]."""
    pd = neg_sqdist_self(x)
    return topk_desc(pd, k + 1, axis=-1)[:, :, 1:].astype(np.int64)


def get_graph_feature(x, k=20, idx=None):
    """util/util.py:176-199: edge tensor [B,2D,N,k] = concat(neighbour f_j, centre x_i)."""
    x = _f32(x)
    B, D, N = x.shape
    if idx is None:
        idx = knn(x, k)
    xt = x.transpose(0, 2, 1)                                      # [B,N,D]
    nbr = np.stack([xt[b][idx[b]] for b in range(B)])              # [B,N,k,D]
    ctr = np.broadcast_to(xt[:, :, None, :], nbr.shape)
    return np.concatenate([nbr, ctr], axis=3).transpose(0, 3, 1, 2).astype(F32)


def farthest_point_sample(xyz, npoint):
    """util/util.py:107-140 (== util/fps.py:10-49).  xyz [B,3,N] -> int64 [B,npoint].

    Seed = farthest point from the barycentre; distances are (dx^2+dy^2)+dz^2 with separate
    multiply and add; argmax takes the first maximal index."""
    p = _f32(xyz).transpose(0, 2, 1)                               # [B,N,3]
    B, N, _ = p.shape
    out = np.zeros((B, npoint), dtype=np.int64)
    distance = np.full((B, N), 1e10, dtype=F32)
    bary = (np.sum(p, axis=1, dtype=F32) / F32(N)).reshape(B, 1, 3)

    def sq(d):
        d2 = (d * d).astype(F32)
        return ((d2[..., 0] + d2[..., 1]).astype(F32) + d2[..., 2]).astype(F32)

    farthest = np.argmax(sq(p - bary), axis=1)
    rows = np.arange(B)
    for i in range(npoint):
        out[:, i] = farthest
        c = p[rows, farthest].reshape(B, 1, 3)
        dist = sq(p - c)
        distance = np.where(dist < distance, dist, distance)
        farthest = np.argmax(distance, axis=1)
    return out


def transform_point_cloud(pc, R, t):
    """util/util.py:91-96 (rotation-matrix branch).  pc [B,3,N], R [B,3,3], t [B,3]."""
    return (np.matmul(_f32(R), _f32(pc)) + _f32(t)[:, :, None]).astype(F32)


def quat2mat(q):
    """util/util.py:76-88.  q [B,4] as (x,y,z,w)."""
    q = _f32(q)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    m = np.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                  2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                  2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], axis=1)
    return m.reshape(-1, 3, 3).astype(F32)


# --------------------------------------------------------------------------------------
# LPDNet embedding (model/lpdnet_model.py:103-137) and its optional TranformNets (:19-70)
# --------------------------------------------------------------------------------------

def transform_net_forward(p, x, prefix, k):
    """TranformNet.forward (model/lpdnet_model.py:44-70), eval mode.  x [B,k,N] -> [B,k,k].
    The activations are F.relu (the module's LeakyReLU member is never called)."""
    h = _f32(x)
    for i in (1, 2, 3):                                                                 # :54-56
        h = np.maximum(batch_norm_eval(conv1x1(h, p[prefix + f"conv{i}.weight"], p[prefix + f"conv{i}.bias"]),
                                       p, prefix + f"bn{i}"), 0)
    h = h.max(axis=2)                                                                   # :57-58 [B,1024]
    for i, bn in ((1, 4), (2, 5)):                                                      # :60-61
        h = np.maximum(batch_norm_eval(linear(h, p[prefix + f"fc{i}.weight"], p[prefix + f"fc{i}.bias"]),
                                       p, prefix + f"bn{bn}"), 0)
    h = linear(h, p[prefix + "fc3.weight"], p[prefix + "fc3.bias"])                     # :62
    h = h + np.eye(k, dtype=F32).reshape(1, k * k)                                      # :66-68
    return _f32(h.reshape(-1, k, k))


def lpdnet_forward(p, x, slope=0.0, prefix="emb_nn.", k=20, idx_feat=None, idx_xyz=None,
                   return_stages=False, t3d=False, tfea=False):
    """x [B,3,N] -> [B,emb_dims,N].  ``idx_feat`` / ``idx_xyz`` inject neighbour sets
    (the reference allows this through get_graph_feature(x, idx=...), util/util.py:176)."""
    x = _f32(x)
    B, _, N = x.shape
    x_init = x
    g = lambda name: p[prefix + name]
    trans = trans_feat = None
    if t3d:                                                                             # :107-109
        trans = transform_net_forward(p, x, prefix + "t_net3d.", 3)
        x = _f32(np.matmul(x.transpose(0, 2, 1), trans).transpose(0, 2, 1))
    h = leaky_relu(conv1x1(x, g("conv1_lpd.weight"), g("conv1_lpd.bias")), slope)       # :111
    h = leaky_relu(conv1x1(h, g("conv2_lpd.weight"), g("conv2_lpd.bias")), slope)       # :112
    if tfea:                                                                            # :114-118
        trans_feat = transform_net_forward(p, h, prefix + "t_net_fea.", 64)
        h = _f32(np.matmul(h.transpose(0, 2, 1), trans_feat).transpose(0, 2, 1))
    f64 = h
    if idx_feat is None:
        idx_feat = knn(h, k)                                                            # :122
    e = get_graph_feature(h, k, idx_feat)
    e = leaky_relu(conv1x1(e, g("convDG1.0.weight"), g("convDG1.0.bias")), slope)       # :123
    x1 = e.max(axis=-1)                                                                 # :124
    e = leaky_relu(conv1x1(e, g("convDG2.0.weight"), g("convDG2.0.bias")), slope)       # :125
    x2 = e.max(axis=-1)                                                                 # :126
    if idx_xyz is None:
        idx_xyz = knn(x_init, k)                                                        # :129
    e = get_graph_feature(x2, k, idx_xyz)                                               # :130
    e = leaky_relu(conv1x1(e, g("convSN1.0.weight"), g("convSN1.0.bias")), slope)       # :131
    x3 = e.max(axis=-1)                                                                 # :132
    cat = np.concatenate([x1, x2, x3], axis=1)                                          # :134
    out = leaky_relu(conv1x1(cat, g("conv3_lpd.weight"), g("conv3_lpd.bias")), slope)   # :135
    if return_stages:
        return out, {"f64": f64, "idx_feat": idx_feat, "idx_xyz": idx_xyz,
                     "x1": x1, "x2": x2, "x3": x3, "trans": trans, "trans_feat": trans_feat}
    return out


def batch_norm_eval(x, p, prefix, eps=1e-5):
    """nn.BatchNorm1d/2d in eval mode: channel axis 1, running statistics."""
    sh = (1, -1) + (1,) * (x.ndim - 2)
    mean, var = p[prefix + ".running_mean"].reshape(sh), p[prefix + ".running_var"].reshape(sh)
    w, b = p[prefix + ".weight"].reshape(sh), p[prefix + ".bias"].reshape(sh)
    return _f32((x - mean) / np.sqrt(var + np.float32(eps)) * w + b)


def dgcnn_forward(p, x, prefix="", k=20, idx=None):
    """DGCNN.forward (model/vcrnet_model.py:105-123), eval mode.  x [B,3,N] -> [B,emb_dims,N]."""
    x = _f32(x)
    g = lambda name: p[prefix + name]
    e = get_graph_feature(x, k, idx)                                                    # :107 [B,6,N,k]
    outs = []
    for i in (1, 2, 3, 4):                                                              # :109-119
        e = np.maximum(batch_norm_eval(conv1x1(e, g(f"conv{i}.weight"), None), p, prefix + f"bn{i}"), 0)
        outs.append(e.max(axis=-1))
    cat = np.concatenate(outs, axis=1)                                                  # :120
    return np.maximum(batch_norm_eval(conv1x1(cat, g("conv5.weight"), None), p, prefix + "bn5"), 0)   # :122


def pointnet_forward(p, x, prefix=""):
    """PointNet.forward (model/vcrnet_model.py:82-88), eval mode."""
    h = _f32(x)
    for i in range(1, 6):
        h = np.maximum(batch_norm_eval(conv1x1(h, p[prefix + f"conv{i}.weight"], None), p, prefix + f"bn{i}"), 0)
    return h


# --------------------------------------------------------------------------------------
# Transformer pointer (model/transformer.py)
# --------------------------------------------------------------------------------------

def layer_norm(x, a2, b2, eps=1e-6):
    """model/transformer.py:141-144: a*(x-mean)/(std_unbiased+eps)+b  (NOT nn.LayerNorm)."""
    mean = x.mean(axis=-1, keepdims=True, dtype=F32)
    std = x.std(axis=-1, keepdims=True, ddof=1, dtype=F32)
    return (a2 * (x - mean) / (std + F32(eps)) + b2).astype(F32)


def attention(q, k, v, is_src=False, overlap2=0.75):
    """model/transformer.py:13-55.  q [B,h,Nq,dk], k/v [B,h,Nk,dk] -> (out, p_attn).

    ``is_src`` (partial cross-attention): sum p_attn over heads and queries, keep the
    int(Nk*overlap2) keys with the largest column sum, mask the rest to -1e9, re-softmax."""
    dk = q.shape[-1]
    scores = (np.matmul(q, k.transpose(0, 1, 3, 2)) / F32(math.sqrt(dk))).astype(F32)
    p = softmax(scores, -1)
    if is_src:
        B, h, Nk, _ = k.shape
        colsum = p.sum(axis=(1, 2), dtype=F32)                     # [B,Nk]
        keep_n = int(Nk * overlap2)
        keep = topk_desc(colsum, keep_n, axis=-1)
        mask = np.zeros((B, Nk), dtype=bool)
        for b in range(B):
            mask[b, keep[b]] = True
        scores = np.where(mask[:, None, None, :], scores, F32(-1e9)).astype(F32)
        p = softmax(scores, -1)
    return np.matmul(p, v).astype(F32), p


def mha(p, prefix, query, key, value, h=4, is_src=False, overlap2=0.75):
    """model/transformer.py:202-224.  [B,N,D] x3 -> [B,N,D]."""
    B, Nq, D = query.shape
    dk = D // h

    def proj(i, x):
        y = linear(x, p[f"{prefix}.linears.{i}.weight"], p[f"{prefix}.linears.{i}.bias"])
        return y.reshape(B, -1, h, dk).transpose(0, 2, 1, 3)

    q, k, v = proj(0, query), proj(1, key), proj(2, value)
    x, _ = attention(q, k, v, is_src=is_src, overlap2=overlap2)
    x = x.transpose(0, 2, 1, 3).reshape(B, Nq, D)
    return linear(x, p[f"{prefix}.linears.3.weight"], p[f"{prefix}.linears.3.bias"])


def feed_forward(p, prefix, x):
    """model/transformer.py:237-238: w_2(relu(w_1 x))."""
    hdn = np.maximum(linear(x, p[f"{prefix}.w_1.weight"], p[f"{prefix}.w_1.bias"]), 0)
    return linear(hdn, p[f"{prefix}.w_2.weight"], p[f"{prefix}.w_2.bias"])


def _ln(p, name, x):
    return layer_norm(x, p[f"{name}.a_2"], p[f"{name}.b_2"])


def encoder_decoder(p, src, tgt, prefix="pointer.model", h=4, partial=False, overlap2=0.75,
                    n_blocks=1):
    """model/transformer.py:72-82 with Encoder :108-117, EncoderLayer :156-166, Decoder
    :120-131, DecoderLayer :169-185.  Returns decode(encode(src), tgt).  [B,N,D] tensors."""
    x = src
    for l in range(n_blocks):
        L = f"{prefix}.encoder.layers.{l}"
        n = _ln(p, f"{L}.sublayer.0.norm", x)
        x = x + mha(p, f"{L}.self_attn", n, n, n, h)
        x = x + feed_forward(p, f"{L}.feed_forward", _ln(p, f"{L}.sublayer.1.norm", x))
    mem = _ln(p, f"{prefix}.encoder.norm", x)
    y = tgt
    for l in range(n_blocks):
        L = f"{prefix}.decoder.layers.{l}"
        n = _ln(p, f"{L}.sublayer.0.norm", y)
        y = y + mha(p, f"{L}.self_attn", n, n, n, h)
        n = _ln(p, f"{L}.sublayer.1.norm", y)
        y = y + mha(p, f"{L}.src_attn", n, mem, mem, h, is_src=partial, overlap2=overlap2)
        y = y + feed_forward(p, f"{L}.feed_forward", _ln(p, f"{L}.sublayer.2.norm", y))
    return _ln(p, f"{prefix}.decoder.norm", y).astype(F32)


def transformer_forward(p, src_emb, tgt_emb, h=4, partial=False, overlap2=0.75, n_blocks=1):
    """model/transformer.py:264-272.  [B,D,N] x2 -> (src_p, tgt_p) both [B,D,N]."""
    s = _f32(src_emb).transpose(0, 2, 1)
    t = _f32(tgt_emb).transpose(0, 2, 1)
    tgt_p = encoder_decoder(p, s, t, h=h, partial=partial, overlap2=overlap2, n_blocks=n_blocks)
    src_p = encoder_decoder(p, t, s, h=h, partial=partial, overlap2=overlap2, n_blocks=n_blocks)
    return (np.ascontiguousarray(src_p.transpose(0, 2, 1)),
            np.ascontiguousarray(tgt_p.transpose(0, 2, 1)))


# --------------------------------------------------------------------------------------
# VcpTopK head (model/vcrnet_model.py:162-347)
# --------------------------------------------------------------------------------------

def neg_sqdist_cross(a, b):
    """model/vcrnet_model.py:337-342 (same lines at :210-215, :286-291).
    a [B,D,Na], b [B,D,Nb] -> [B,Na,Nb] = (-|a_i|^2 - (-2 a_i.b_j)) - |b_j|^2."""
    a, b = _f32(a), _f32(b)
    inner = F32(-2.0) * np.matmul(a.transpose(0, 2, 1), b)
    xx = np.sum(a * a, axis=1, keepdims=True, dtype=F32).transpose(0, 2, 1)
    yy = np.sum(b * b, axis=1, keepdims=True, dtype=F32)
    return ((-xx - inner) - yy).astype(F32)


def get_copair_all(src, src_emb, tgt, tgt_emb):
    """model/vcrnet_model.py:334-347 (whole): src_corr = tgt . softmax_j(pd)^T."""
    scores = softmax(neg_sqdist_cross(src_emb, tgt_emb), 2)
    src_corr = np.matmul(_f32(tgt), scores.transpose(0, 2, 1))
    return _f32(src), src_corr.astype(F32)


def vcp_by_dis(src_emb, tgt_emb, src, tgt):
    """VcpByDis.forward (model/vcrnet_model.py:407-421)."""
    d_k = src_emb.shape[1]
    scores = _f32(np.matmul(src_emb.transpose(0, 2, 1), tgt_emb) / np.float32(math.sqrt(d_k)))
    scores = softmax(scores, axis=2)
    return src, _f32(np.matmul(tgt, scores.transpose(0, 2, 1)))


def vcp_att(p, prefix, src_emb, tgt_emb, src, tgt):
    """VcpAtt.forward (model/vcrnet_model.py:434-460): Linear on both embeddings, then the getCopairALL arithmetic."""
    q = linear(src_emb.transpose(0, 2, 1), p[prefix + "linears_emb.0.weight"], p[prefix + "linears_emb.0.bias"])
    k = linear(tgt_emb.transpose(0, 2, 1), p[prefix + "linears_emb.1.weight"], p[prefix + "linears_emb.1.bias"])
    return get_copair_all(src, np.ascontiguousarray(q.transpose(0, 2, 1)), tgt, np.ascontiguousarray(k.transpose(0, 2, 1)))


def _gather_cols(x, idx):
    """x [B,C,N], idx [B,K] -> [B,C,K]."""
    return np.stack([x[b][:, idx[b]] for b in range(x.shape[0])]).astype(F32)


def select_com(src, src_emb, tgt, tgt_emb, overlap2):
    """model/vcrnet_model.py:190-262 minus the unused ``*_remain`` outputs (:228,249).

    Returns (src_o, src_emb_o, tgt_o, tgt_emb_o, idx_src, idx_tgt); order = top-k order."""
    Ns, Nt = src.shape[2], tgt.shape[2]
    srcK = int(Ns * 0.84 * overlap2)
    tgtK = int(Nt * 0.84 * overlap2)
    scores = neg_sqdist_cross(src_emb, tgt_emb)
    col = softmax(scores, 2).sum(axis=1, dtype=F32)                # [B,Nt]  (:221-222)
    idx_t = topk_desc(col, tgtK, axis=-1)
    row = softmax(scores, 1).sum(axis=2, dtype=F32)                # [B,Ns]  (:243-244)
    idx_s = topk_desc(row, srcK, axis=-1)
    return (_gather_cols(_f32(src), idx_s), _gather_cols(_f32(src_emb), idx_s),
            _gather_cols(_f32(tgt), idx_t), _gather_cols(_f32(tgt_emb), idx_t), idx_s, idx_t)


def get_copair(src, src_emb, tgt, tgt_emb, overlap2):
    """model/vcrnet_model.py:264-332 (partial).  tgtK = 1 => val/val_sum == 1, i.e. a HARD
    correspondence to the arg-max target for the int(Ns*0.52*overlap2) most confident sources."""
    Ns = src.shape[2]
    srcK = int(Ns * 0.52 * overlap2)
    P = softmax(neg_sqdist_cross(src_emb, tgt_emb), 2)
    best = topk_desc(P, 1, axis=-1)[..., 0]                         # [B,Ns]
    val = np.take_along_axis(P, best[..., None], axis=2)[..., 0]    # [B,Ns]
    keep = topk_desc(val, srcK, axis=-1)                            # [B,srcK]
    w = (val / val).astype(F32)                                     # :323-324
    B = src.shape[0]
    corr = np.stack([_f32(tgt)[b][:, best[b][keep[b]]] * w[b][keep[b]][None, :] for b in range(B)])
    return _gather_cols(_f32(src), keep), corr.astype(F32), keep, best


def vcp_topk_forward(src_emb, tgt_emb, src, tgt, partial=False, overlap2=0.75):
    """model/vcrnet_model.py:173-188."""
    if partial:
        s, se, t, te, _, _ = select_com(src, src_emb, tgt, tgt_emb, overlap2)
        s2, corr, _, _ = get_copair(s, se, t, te, overlap2)
        return s2, corr
    return get_copair_all(src, src_emb, tgt, tgt_emb)


# --------------------------------------------------------------------------------------
# SVD head (model/vcrnet_model.py:350-399)
# --------------------------------------------------------------------------------------

def svd_head(src, src_corr, reflect=None):
    """Centre, H = S_c C_c^T, per-item SVD, R = V U^T (V's last column flipped when
    det < 0), t = -R mean(src) + mean(corr).  [B,3,M] x2 -> R [B,3,3], t [B,3]."""
    src, src_corr = _f32(src), _f32(src_corr)
    if reflect is None:
        reflect = np.diag([1.0, 1.0, -1.0]).astype(F32)
    ms = src.mean(axis=2, keepdims=True, dtype=F32)
    mc = src_corr.mean(axis=2, keepdims=True, dtype=F32)
    H = np.matmul(src - ms, (src_corr - mc).transpose(0, 2, 1)).astype(F32)
    Rs = []
    for b in range(src.shape[0]):
        u, s, vt = np.linalg.svd(H[b].astype(np.float64))
        v = vt.T
        r = v @ u.T
        if np.linalg.det(r) < 0:
            v = v @ reflect.astype(np.float64)
            r = v @ u.T
        Rs.append(r.astype(F32))
    R = np.stack(Rs)
    t = np.matmul(-R, ms) + mc
    return R, t.reshape(-1, 3).astype(F32)


# --------------------------------------------------------------------------------------
# VCRNet assembly + refinement loop (model/vcrnet_model.py:21-43, 495-518)
# --------------------------------------------------------------------------------------

def vcrnet_forward(p, src, tgt, partial=False, overlap2=0.75, h=4, pointer="transformer",
                   slope=0.0, return_stages=False):
    src, tgt = _f32(src), _f32(tgt)
    se = lpdnet_forward(p, src, slope)                                                  # :499
    te = lpdnet_forward(p, tgt, slope)                                                  # :500
    stages = {"src_emb0": se, "tgt_emb0": te}
    if pointer == "transformer":
        sp, tp = transformer_forward(p, se, te, h=h, partial=partial, overlap2=overlap2)  # :503
        se, te = (se + sp).astype(F32), (te + tp).astype(F32)                           # :504-505
    elif pointer == "identity":                                                         # Identity returns its inputs (:65)
        se, te = (se + se).astype(F32), (te + te).astype(F32)                           # :504-505
    stages.update(src_emb=se, tgt_emb=te)
    sK, cK = vcp_topk_forward(se, te, src, tgt, partial=partial, overlap2=overlap2)     # :507
    R, t = svd_head(sK, cK, p.get("svd.reflect"))                                       # :509
    R_ba = np.ascontiguousarray(R.transpose(0, 2, 1))                                   # :515
    t_ba = -np.matmul(R_ba, t[:, :, None])[:, :, 0]                                     # :516
    out = (sK, cK, R, t, R_ba, t_ba.astype(F32))
    return (out, stages) if return_stages else out


def eval_metrics_batch(src, tgt, srcK, corrK, R_gt, t_gt, R_ab, t_ab, R_ba, t_ba):
    """One batch of test_one_epoch's running sums (model/vcrnet_model.py:589-627), each already times batch_size:
    (loss_pose, cycle_loss, mse_ab, mae_ab, mse_ba, mae_ba, batch_size)."""
    B = src.shape[0]
    eye = np.eye(3, dtype=np.float64)[None]
    t_target = transform_point_cloud(tgt, R_ba, t_ba).astype(np.float64)                  # :589
    t_srcK = transform_point_cloud(srcK, R_gt, t_gt).astype(np.float64)                   # :591
    loss_pose = ((np.matmul(R_ab.transpose(0, 2, 1), R_gt).astype(np.float64) - eye) ** 2).mean() \
        + ((t_ab.astype(np.float64) - t_gt) ** 2).mean()                                  # :614-615
    rot = ((np.matmul(R_ba, R_ab).astype(np.float64) - eye) ** 2).mean()                  # :620
    tr = ((np.matmul(R_ba.transpose(0, 2, 1), t_ab[:, :, None])[:, :, 0].astype(np.float64) + t_ba) ** 2).mean()
    return (loss_pose * B, (rot + tr) * B, ((t_srcK - corrK) ** 2).mean() * B, np.abs(t_srcK - corrK).mean() * B,
            ((t_target - src) ** 2).mean() * B, np.abs(t_target - src).mean() * B, B)


def icp_forward(src_init, dst, max_iterations=10, tolerance=0.001):
    """ICP.forward (model/icp_model.py:26-50) -> (srcInit, src, R_ab, t_ab, R_ba, t_ba, iterations)."""
    src_init, dst = _f32(src_init), _f32(dst)
    src = src_init
    prev = 0.0
    iters = 0
    for _ in range(max_iterations):
        pd = neg_sqdist_cross(src, dst)                                     # nearest_neighbor :57-62
        idx = pd.argmax(axis=2)
        mean_error = float(pd.max(axis=2).mean())
        corr = _gather_cols(dst, idx)
        R, t = svd_head(src, corr)                                          # best_fit_transform :77-108
        src = transform_point_cloud(src, R, t)
        iters += 1
        if abs(prev - mean_error) < tolerance:
            break
        prev = mean_error
    R, t = svd_head(src_init, src)
    R_ba = np.ascontiguousarray(R.transpose(0, 2, 1))
    t_ba = (-np.matmul(R_ba, t[:, :, None])[:, :, 0]).astype(F32)
    return src_init, src, R, t, R_ba, t_ba, iters


def vcrnet_iter(p, src, tgt, n_iter=1, **kw):
    """model/vcrnet_model.py:21-43: R_f <- R_i R_f, t_f <- R_i t_f + t_i, inverse at the end."""
    cur = _f32(src)
    R_f = t_f = None
    for i in range(n_iter):
        sK, cK, R, t, _, _ = vcrnet_forward(p, cur, tgt, **kw)
        cur = transform_point_cloud(cur, R, t)
        if R_f is None:
            R_f, t_f = R, t
        else:
            t_f = (np.matmul(R, t_f[:, :, None])[:, :, 0] + t).astype(F32)
            R_f = np.matmul(R, R_f).astype(F32)
    R_ba = np.ascontiguousarray(R_f.transpose(0, 2, 1))
    t_ba = (-np.matmul(R_ba, t_f[:, :, None])[:, :, 0]).astype(F32)
    return sK, cK, R_f, t_f, R_ba, t_ba


# --------------------------------------------------------------------------------------
# LPD pre-training loss (model/lpdnet_model.py:149-229), forward only
# --------------------------------------------------------------------------------------

def kfn(x, k):
    """model/lpdnet_model.py:163-171: k FARTHEST (top-k of +squared distance)."""
    x = _f32(x)
    inner = F32(-2.0) * np.matmul(x.transpose(0, 2, 1), x)
    xx = np.sum(x * x, axis=1, keepdims=True, dtype=F32)
    pd = (xx + inner) + xx.transpose(0, 2, 1)
    return topk_desc(pd.astype(F32), k, axis=-1)


def lpd_loss(src, src_emb, tgt_emb, k=32, neg_k=8):
    """model/lpdnet_model.py:191-229."""
    B, _, N = src.shape
    sidx = farthest_point_sample(src, k)                                                # :195
    src_k = _gather_cols(_f32(src), sidx)
    se_k = _gather_cols(_f32(src_emb), sidx)                                            # [B,D,k]
    te_k = _gather_cols(_f32(tgt_emb), sidx)
    far = kfn(src_k, neg_k)                                                             # [B,k,neg_k]
    neg = np.stack([te_k[b][:, far[b]] for b in range(B)])                              # [B,D,k,neg_k]
    dp = ((se_k - te_k) ** 2).mean(axis=1, dtype=F32)                                   # [B,k]
    dn = ((se_k[..., None] - neg) ** 2).mean(axis=(1, 3), dtype=F32)
    trip = np.maximum(F32(0.0), 1 - dn / (F32(1.0) + dp)).astype(F32)                    # :186
    sl = np.sqrt((_f32(src_emb) ** 2).sum(axis=1, dtype=F32))
    tl = np.sqrt((_f32(tgt_emb) ** 2).sum(axis=1, dtype=F32))
    n1 = np.sqrt(((sl - 1) ** 2).mean(dtype=F32))
    n2 = np.sqrt(((tl - 1) ** 2).mean(dtype=F32))
    return F32(trip.mean(dtype=F32) + (n1 + n2) / 2.0 * 0.03)
