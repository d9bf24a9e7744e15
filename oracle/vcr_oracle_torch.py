"""CPU timing port of the registration loop on torch CPU operators.

TEST / BENCH INFRASTRUCTURE ONLY (imported by tests/ and by bench.py's cpu_baseline / --impl reference legs, never by
the product).  The reference is Python/torch and does not exist on the GPU box, so the CPU arm of the benchmark is a
port.  oracle/vcr_oracle.py (numpy) is the parity oracle; its element-wise passes are single-threaded, which makes it
about 3x slower than the reference's own ATen kernels on a many-core host.  This file restates the SAME functions on
torch CPU ops -- the operators the reference itself runs (bmm, softmax, topk, svd, ...) with all intra-op threads --
so that the reported CPU number is a fair stand-in for "the reference's PyTorch CPU path".  It is checked against the
numpy oracle and the live-reference golden vectors in tests/test_oracle_vs_golden.py.

Each function cites the reference lines it follows (same as the numpy oracle):
  knn / get_graph_feature   util/util.py:143-199          lpdnet   model/lpdnet_model.py:103-137
  attention / mha / ffn     model/transformer.py:13-55, 202-238      encoder_decoder  :72-82, 108-185
  VcpTopK                   model/vcrnet_model.py:173-347  SVDHead  :356-399   vcrnetIter  :21-43
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(a)


def knn(x, k):
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    return pd.topk(k=k + 1, dim=-1)[1][:, :, 1:]


def graph_feature(x, k=20, idx=None):
    B, D, N = x.shape
    if idx is None:
        idx = knn(x, k)
    xt = x.transpose(2, 1).contiguous()
    flat = (idx + torch.arange(B).view(-1, 1, 1) * N).view(-1)
    nbr = xt.view(B * N, D)[flat].view(B, N, k, D)
    ctr = xt.view(B, N, 1, D).expand(-1, -1, k, -1)
    return torch.cat((nbr, ctr), dim=3).permute(0, 3, 1, 2)


def _conv(x, w, b):
    return F.conv1d(x, w, b) if w.dim() == 3 else F.conv2d(x, w, b)


def lpdnet(p, x, slope=0.0, prefix="emb_nn.", k=20):
    g = lambda n: p[prefix + n]
    h = F.leaky_relu(_conv(x, g("conv1_lpd.weight"), g("conv1_lpd.bias")), slope)
    h = F.leaky_relu(_conv(h, g("conv2_lpd.weight"), g("conv2_lpd.bias")), slope)
    e = F.leaky_relu(_conv(graph_feature(h, k), g("convDG1.0.weight"), g("convDG1.0.bias")), slope)
    x1 = e.max(dim=-1)[0]
    e = F.leaky_relu(_conv(e, g("convDG2.0.weight"), g("convDG2.0.bias")), slope)
    x2 = e.max(dim=-1)[0]
    e = F.leaky_relu(_conv(graph_feature(x2, k, knn(x, k)), g("convSN1.0.weight"), g("convSN1.0.bias")), slope)
    x3 = e.max(dim=-1)[0]
    return F.leaky_relu(_conv(torch.cat((x1, x2, x3), dim=1), g("conv3_lpd.weight"), g("conv3_lpd.bias")), slope)


def layer_norm(x, a2, b2, eps=1e-6):
    return a2 * (x - x.mean(-1, keepdim=True)) / (x.std(-1, keepdim=True) + eps) + b2


def attention(q, k, v, is_src=False, overlap2=0.75):
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(q.size(-1))
    p = F.softmax(scores, dim=-1)
    if is_src:
        Nk = k.size(2)
        keep = p.sum(dim=(1, 2)).topk(int(Nk * overlap2), dim=-1)[1]
        mask = torch.zeros(k.size(0), Nk, dtype=torch.bool).scatter_(1, keep, True)
        p = F.softmax(scores.masked_fill(~mask[:, None, None, :], -1e9), dim=-1)
    return torch.matmul(p, v)


def mha(p, prefix, query, key, value, h=4, is_src=False, overlap2=0.75):
    B, Nq, D = query.shape
    proj = lambda i, x: F.linear(x, p[f"{prefix}.linears.{i}.weight"], p[f"{prefix}.linears.{i}.bias"]) \
        .view(B, -1, h, D // h).transpose(1, 2)
    x = attention(proj(0, query), proj(1, key), proj(2, value), is_src, overlap2)
    x = x.transpose(1, 2).contiguous().view(B, Nq, D)
    return F.linear(x, p[f"{prefix}.linears.3.weight"], p[f"{prefix}.linears.3.bias"])


def ffn(p, prefix, x):
    return F.linear(F.relu(F.linear(x, p[f"{prefix}.w_1.weight"], p[f"{prefix}.w_1.bias"])),
                    p[f"{prefix}.w_2.weight"], p[f"{prefix}.w_2.bias"])


def encoder_decoder(p, src, tgt, prefix="pointer.model", h=4, partial=False, overlap2=0.75, n_blocks=1):
    ln = lambda name, x: layer_norm(x, p[f"{name}.a_2"], p[f"{name}.b_2"])
    x = src
    for l in range(n_blocks):
        L = f"{prefix}.encoder.layers.{l}"
        n = ln(f"{L}.sublayer.0.norm", x)
        x = x + mha(p, f"{L}.self_attn", n, n, n, h)
        x = x + ffn(p, f"{L}.feed_forward", ln(f"{L}.sublayer.1.norm", x))
    mem = ln(f"{prefix}.encoder.norm", x)
    y = tgt
    for l in range(n_blocks):
        L = f"{prefix}.decoder.layers.{l}"
        n = ln(f"{L}.sublayer.0.norm", y)
        y = y + mha(p, f"{L}.self_attn", n, n, n, h)
        n = ln(f"{L}.sublayer.1.norm", y)
        y = y + mha(p, f"{L}.src_attn", n, mem, mem, h, is_src=partial, overlap2=overlap2)
        y = y + ffn(p, f"{L}.feed_forward", ln(f"{L}.sublayer.2.norm", y))
    return ln(f"{prefix}.decoder.norm", y)


def transformer(p, src_emb, tgt_emb, **kw):
    s, t = src_emb.transpose(2, 1).contiguous(), tgt_emb.transpose(2, 1).contiguous()
    tgt_p = encoder_decoder(p, s, t, **kw).transpose(2, 1).contiguous()
    src_p = encoder_decoder(p, t, s, **kw).transpose(2, 1).contiguous()
    return src_p, tgt_p


def neg_sqdist(a, b):
    inner = -2 * torch.matmul(a.transpose(2, 1).contiguous(), b)
    xx = torch.sum(a ** 2, dim=1, keepdim=True).transpose(2, 1)
    yy = torch.sum(b ** 2, dim=1, keepdim=True)
    return -xx - inner - yy


def _gather(x, idx):
    return torch.gather(x, 2, idx.unsqueeze(1).expand(-1, x.size(1), -1))


def vcp_topk(src_emb, tgt_emb, src, tgt, partial=False, overlap2=0.75):
    if not partial:                                                       # getCopairALL :334-347
        return src, torch.matmul(tgt, F.softmax(neg_sqdist(src_emb, tgt_emb), dim=2).transpose(2, 1))
    Ns, Nt = src.size(2), tgt.size(2)                                     # selectCom :190-262
    scores = neg_sqdist(src_emb, tgt_emb)
    idx_t = F.softmax(scores, dim=2).sum(dim=1).topk(int(Nt * 0.84 * overlap2), dim=-1)[1]
    idx_s = F.softmax(scores, dim=1).sum(dim=2).topk(int(Ns * 0.84 * overlap2), dim=-1)[1]
    s, se, t, te = _gather(src, idx_s), _gather(src_emb, idx_s), _gather(tgt, idx_t), _gather(tgt_emb, idx_t)
    P = F.softmax(neg_sqdist(se, te), dim=2)                              # getCopair :264-332
    val, best = P.topk(1, dim=-1)
    val, best = val[..., 0], best[..., 0]
    keep = val.topk(int(s.size(2) * 0.52 * overlap2), dim=-1)[1]
    corr = _gather(t, torch.gather(best, 1, keep)) * (torch.gather(val, 1, keep) / torch.gather(val, 1, keep)).unsqueeze(1)
    return _gather(s, keep), corr


def svd_head(src, corr):
    ms, mc = src.mean(dim=2, keepdim=True), corr.mean(dim=2, keepdim=True)
    H = torch.matmul(src - ms, (corr - mc).transpose(2, 1))
    reflect = torch.diag(torch.tensor([1.0, 1.0, -1.0]))
    Rs = []
    for i in range(src.size(0)):                                          # per-item svd like :369-381
        u, _, vh = torch.linalg.svd(H[i])
        v = vh.transpose(0, 1)
        r = v @ u.t()
        if torch.det(r) < 0:
            r = (v @ reflect) @ u.t()
        Rs.append(r)
    R = torch.stack(Rs)
    return R, (torch.matmul(-R, ms) + mc).view(-1, 3)


def vcrnet_forward(p, src, tgt, partial=False, overlap2=0.75):
    se, te = lpdnet(p, src), lpdnet(p, tgt)
    sp, tp = transformer(p, se, te, partial=partial, overlap2=overlap2)
    sK, cK = vcp_topk(se + sp, te + tp, src, tgt, partial, overlap2)
    R, t = svd_head(sK, cK)
    return sK, cK, R, t


@torch.no_grad()
def vcrnet_iter(params, src, tgt, n_iter=1, partial=False, overlap2=0.75):
    """model/vcrnet_model.py:21-43 on torch CPU ops.  params: dict name -> numpy array / tensor."""
    p = {k: _t(v) for k, v in params.items()}
    cur, tgt = _t(src), _t(tgt)
    R_f = t_f = None
    for _ in range(n_iter):
        sK, cK, R, t = vcrnet_forward(p, cur, tgt, partial, overlap2)
        cur = torch.matmul(R, cur) + t.unsqueeze(2)
        if R_f is None:
            R_f, t_f = R, t
        else:
            t_f = torch.matmul(R, t_f.unsqueeze(2)).squeeze(2) + t
            R_f = torch.matmul(R, R_f)
    R_ba = R_f.transpose(2, 1).contiguous()
    t_ba = -torch.matmul(R_ba, t_f.unsqueeze(2)).squeeze(2)
    return tuple(x.numpy() for x in (sK, cK, R_f, t_f, R_ba, t_ba))
