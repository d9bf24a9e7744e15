"""GPU parity tests: CUDA path (through the C ABI) vs the oracle and the live-reference golden
vectors.  Integer outputs bit-exact; floating point within 1e-4 relative (BASELINE.json
north_star) unless a looser bound is justified inline.  Run on the B200 box: pytest -m gpu.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import vcr_net_b200 as V
    from vcr_net_b200 import ops
    from vcr_net_b200 import functional as Fn
from oracle import canon, synth
from oracle import vcr_oracle as O
from oracle.ref_harness import default_args

TOL = 1e-4
DEV = "cuda:0"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def nump(t):
    return t.detach().cpu().numpy()


@pytest.fixture(params=["fp32", "h3"])
def precision(request):
    """Both fp32-parity engines: FP32 SIMT GEMMs and the tcgen05 3-term fp16-split GEMMs."""
    from vcr_net_b200 import config
    old = config.precision
    config.set_precision(request.param)
    yield request.param
    config.set_precision(old)


@pytest.fixture(scope="module")
def net_whole(ckpt):
    net = V.VCRNet(default_args()).to(DEV).eval()
    net.load_state_dict(synth.checkpoint_to_torch(ckpt), strict=True)
    return net


@pytest.fixture(scope="module")
def net_partial(ckpt):
    net = V.VCRNet(default_args(partial=True, overlap2=synth.OVERLAP2_0575)).to(DEV).eval()
    net.load_state_dict(synth.checkpoint_to_torch(ckpt), strict=True)
    return net


# ---------------------------------------------------------------- kNN ---------------------------------
@pytest.mark.parametrize("key", ["x3g", "x64g", "x3f", "x64f"])
def test_knn_bit_exact_vs_canonical(key):
    x = load_golden("knn")[key]
    want = canon.knn(x, 20)
    got = nump(V.knn(cu(x), 20))
    assert got.dtype == np.int64 and np.array_equal(got, want), (got != want).any(-1).mean()


def test_knn_vs_live_reference_grid():
    g = load_golden("knn")
    for x, idx in ((g["x3g"], g["idx3g"]), (g["x64g"], g["idx64g"])):
        got = nump(V.knn(cu(x), 20))
        pd = O.neg_sqdist_self(x)
        assert np.array_equal(np.take_along_axis(pd, got, -1), np.take_along_axis(pd, idx, -1))


@pytest.mark.parametrize("D,N,k,tm", [(3, 777, 20, False), (3, 33, 20, False), (64, 130, 20, True),
                                      (3, 1024, 1, False), (64, 300, 8, False), (3, 4096, 20, False),
                                      (64, 2048, 20, True), (3, 21, 20, False), (8, 200, 5, False),
                                      (128, 515, 20, True), (6, 300, 20, False), (200, 64, 31, False)])
def test_knn_shapes(D, N, k, tm):
    rs = np.random.RandomState(N + D)
    x = rs.randn(2, D, N).astype(np.float32)
    want = canon.knn(x, k)
    xin = cu(x.transpose(0, 2, 1)) if tm else cu(x)
    got = nump(ops.knn_topk(xin, k, token_major=tm))
    assert np.array_equal(got, want)


@pytest.fixture(params=[0, 1], ids=["warp-per-query", "thread-per-query"])
def knn_tc_selection(request):
    """Both selections of the tensor-core kNN route (csrc/knn.cu knn_tc_kernel, csrc/knn_tpq.cuh knn_tc2_kernel); the
    default routes by N (thread-per-query from N >= 8192), so each is forced here."""
    from vcr_net_b200._lib import lib
    old = lib().vcr_set_knn_tc_tpq(request.param)
    yield request.param
    lib().vcr_set_knn_tc_tpq(old)


@pytest.mark.parametrize("D,N,k", [(64, 130, 20), (64, 1024, 20), (64, 768, 20), (64, 2048, 20), (128, 515, 20),
                                   (16, 300, 5), (96, 257, 30), (32, 21, 20), (64, 4096, 20), (128, 1000, 1)])
def test_knn_tc_prefilter_bit_exact(D, N, k, knn_tc_selection):
    """tcgen05 prefilter + exact re-rank == canonical C oracle, bit for bit; on generic data almost nothing needs the
    exact kernel."""
    rs = np.random.RandomState(N + D + k)
    x = rs.randn(3, D, N).astype(np.float32)
    want = canon.knn(x, k)
    xt = cu(np.ascontiguousarray(x.transpose(0, 2, 1)))
    got, flagged = ops.knn_topk_tc(xt, None, k, want_flagged=True)
    assert np.array_equal(nump(got), want)
    assert int(flagged.item()) <= 0.02 * 3 * N, int(flagged.item())
    got32, got64 = ops.knn_topk_tc(xt, ops.to_operand(xt.view(3 * N, D), "h3"), k, want64=True)
    assert np.array_equal(nump(got64), want) and got64.dtype == torch.int64 and torch.equal(got32.long(), got64)


def test_knn_tc_prefilter_adversarial_inputs(knn_tc_selection):
    """Inputs that defeat the certificate (exact ties, duplicates, cancellation, out-of-range magnitudes) fall back to the
    exact kernel inside the same call: still bit-exact."""
    rs = np.random.RandomState(17)
    g = load_golden("knn")
    cases = {
        "grid_ties": g["x64g"],                                                         # dyadic grid: real ties
        "real_features": g["x64f"],                                                     # LPDNet conv2 output
        "duplicates": np.repeat(rs.randn(2, 64, 100).astype(np.float32), 4, axis=2),    # every point 4 times
        "zeros": np.zeros((1, 64, 200), np.float32),
        "dead_half": np.concatenate([np.zeros((2, 64, 300), np.float32), rs.randn(2, 64, 300).astype(np.float32)], 2),
        "offset": (rs.randn(2, 64, 500) * 0.01 + 100.0).astype(np.float32),             # cancellation: eps >> gaps
        "huge": (rs.randn(2, 64, 300) * 1e6).astype(np.float32),                        # outside the fp16 split range
        "tiny": (rs.randn(2, 64, 300) * 1e-6).astype(np.float32),
        "mixed_scale": (rs.randn(2, 64, 400) * np.exp(rs.randn(2, 1, 400) * 2)).astype(np.float32),
    }
    n_flag = {}
    for name, x in cases.items():
        xt = cu(np.ascontiguousarray(x.transpose(0, 2, 1)))
        got, flagged = ops.knn_topk_tc(xt, None, 20, want_flagged=True)
        assert np.array_equal(nump(got), canon.knn(x, 20)), name
        n_flag[name] = int(flagged.item())
    assert n_flag["real_features"] <= 0.05 * g["x64f"].shape[0] * g["x64f"].shape[2], n_flag
    assert n_flag["zeros"] == 200 and n_flag["huge"] == 600 and n_flag["tiny"] == 600, n_flag
    assert n_flag["offset"] > 0, n_flag


def test_knn_tc_matches_simt_inside_lpdnet(net_whole):
    """The LPDNet feature-space neighbour sets are identical with the prefilter on and off (h3 mode)."""
    from vcr_net_b200 import config
    x = cu(synth.make_pairs(4, 1024, first_item=33)["src"])
    old_p, old_k = config.precision, config.knn_tc
    try:
        config.set_precision("h3")
        st_on, st_off = {}, {}
        config.knn_tc = "1"
        a = net_whole.emb_nn.forward_tokens(x, stages=st_on)
        config.knn_tc = "0"
        b = net_whole.emb_nn.forward_tokens(x, stages=st_off)
        assert torch.equal(st_on["idx_feat"], st_off["idx_feat"]) and torch.equal(a, b)
    finally:
        config.set_precision(old_p)
        config.knn_tc = old_k


@pytest.mark.parametrize("B,N,tm,grid", [(2, 21, False, False), (3, 33, True, False), (2, 777, False, True), (2, 1024, True, True),
                                         (1, 1500, False, True)])
def test_knn_3d_direct_and_tile_kernels_agree(B, N, tm, grid):
    """D == 3 has two routes (on-the-fly distances, default; the generic distance-tile kernel): both equal the canonical
    oracle bit for bit, on tie-heavy grids too."""
    rs = np.random.RandomState(N)
    x = rs.rand(B, 3, N).astype(np.float32) - 0.5
    if grid:
        x = np.round(x * 32) / 32
    xt = cu(x.transpose(0, 2, 1) if tm else x)
    want = canon.knn(x, 20)
    for direct in (True, False):
        old = ops.set_knn3_direct(direct)
        try:
            got = nump(ops.knn_topk(xt, 20, token_major=tm))
        finally:
            ops.set_knn3_direct(old)
        assert np.array_equal(got, want), direct


def test_knn_duplicates_ties():
    rs = np.random.RandomState(5)
    base = synth.grid_cloud(rs, (1, 3, 64), 3, -0.5, 0.5)      # coarse grid: many exact ties / duplicates
    x = np.concatenate([base, base, base, base], axis=2)
    assert np.array_equal(nump(V.knn(cu(x), 20)), canon.knn(x, 20))


def test_knn_errors():
    x = torch.zeros(1, 3, 16, device=DEV)
    with pytest.raises(RuntimeError):
        V.knn(x, 20)                       # N < k+1
    with pytest.raises(RuntimeError):
        V.knn(torch.zeros(1, 3, 64), 20)   # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        V.knn(torch.zeros(1, 3, 64, device=DEV), 40)   # k > 31


def test_graph_feature():
    g = load_golden("graph_feature")
    out = nump(V.get_graph_feature(cu(g["x"]), k=int(g["k"])))
    assert np.array_equal(out, g["out"])
    idx = O.knn(g["x"], int(g["k"]))
    out2 = nump(V.get_graph_feature(cu(g["x"]), k=int(g["k"]), idx=cu(idx)))
    assert np.array_equal(out2, g["out"])


# ---------------------------------------------------------------- FPS ---------------------------------
def test_fps_bit_exact():
    g = load_golden("fps")
    for p, idx in ((g["pg"], g["ig"]), (g["pg2"], g["ig2"]), (g["pf"], g["i_f"])):
        got = nump(V.farthest_point_sample(cu(p), 32))
        assert got.dtype == np.int64
        assert np.array_equal(got, canon.fps(p, 32))
        assert np.array_equal(got, idx)                     # live reference
    rs = np.random.RandomState(3)
    for N, npnt in ((100, 7), (5000, 64), (12000, 16)):
        p = rs.randn(3, 3, N).astype(np.float32)
        assert np.array_equal(nump(V.farthest_point_sample(cu(p), npnt)), canon.fps(p, npnt))


# ---------------------------------------------------------------- GEMM --------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 512, 512), (77, 130, 20), (2048, 1536, 512), (5, 3, 4)])
def test_gemm_epilogues(M, N, K):
    rs = np.random.RandomState(M + N + K)
    a = rs.randn(M, K).astype(np.float32)
    w = rs.randn(N, K).astype(np.float32)
    b = rs.randn(N).astype(np.float32)
    r = rs.randn(M, N).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    got = nump(ops.gemm(cu(a), cu(w)))
    assert rel_err(got, ref) < 1e-5
    got = nump(ops.gemm(cu(a), cu(w), cu(b), act=1, slope=0.2, residual=cu(r), alpha=0.5))
    z = 0.5 * ref + b
    want = np.where(z >= 0, z, 0.2 * z) + r
    assert rel_err(got, want) < 1e-5


def test_gemm_strided_views_and_batched():
    rs = np.random.RandomState(0)
    B, h, Nq, Nk, dk = 2, 4, 96, 80, 128
    qkv = rs.randn(B, Nq, 3 * h * dk).astype(np.float32)
    t = cu(qkv)
    S = torch.empty(B, h, Nq, Nq, device=DEV)
    D = h * dk
    ops.bgemm(t, 3 * D, Nq * 3 * D, dk, t[:, :, D:], 3 * D, Nq * 3 * D, dk, 0, S, Nq, h * Nq * Nq, Nq * Nq,
              Nq, Nq, dk, B, h, alpha=0.25)
    q = qkv[:, :, :D].reshape(B, Nq, h, dk).transpose(0, 2, 1, 3)
    k = qkv[:, :, D:2 * D].reshape(B, Nq, h, dk).transpose(0, 2, 1, 3)
    assert rel_err(nump(S), 0.25 * q @ k.transpose(0, 1, 3, 2)) < 1e-5
    # NN layout: out = S @ V written into a strided [B,Nq,D] buffer per head
    out = torch.zeros(B, Nq, D, device=DEV)
    ops.bgemm(S, Nq, h * Nq * Nq, Nq * Nq, t[:, :, 2 * D:], 3 * D, Nq * 3 * D, dk, 1, out, D, Nq * D, dk,
              Nq, dk, Nq, B, h)
    v = qkv[:, :, 2 * D:].reshape(B, Nq, h, dk).transpose(0, 2, 1, 3)
    want = (nump(S) @ v).transpose(0, 2, 1, 3).reshape(B, Nq, D)
    assert rel_err(nump(out), want) < 1e-5
    # column slice in, column slice out
    cat = torch.zeros(4, 50, 512, device=DEV)
    a = rs.randn(4, 50, 128).astype(np.float32)
    cat[:, :, 128:256] = cu(a)
    w = rs.randn(256, 128).astype(np.float32)
    ops.gemm(cat[:, :, 128:256], cu(w), out=cat[:, :, 256:512])
    assert rel_err(nump(cat[:, :, 256:512]), a @ w.T) < 1e-5
    assert float(cat[:, :, :128].abs().max()) == 0.0


# ---------------------------------------------------------------- row ops -----------------------------
def test_layernorm():
    g = load_golden("layernorm")
    out = nump(ops.layernorm(cu(g["x"]), cu(g["a"]), cu(g["b"])))
    assert rel_err(out, g["out"]) < 1e-5
    res = np.random.RandomState(1).randn(*g["x"].shape).astype(np.float32)
    out = nump(ops.layernorm(cu(g["x"]), cu(g["a"]), cu(g["b"]), residual=cu(res)))
    assert rel_err(out, g["out"] + res) < 1e-5


def test_softmax_colsum_topk():
    rs = np.random.RandomState(2)
    B, rows, n = 3, 40, 333
    s = (rs.randn(B * rows, n) * 3).astype(np.float32)
    keep = (rs.rand(B, n) > 0.3).astype(np.uint8)
    p = nump(ops.softmax_rows_(cu(s).clone()))
    assert rel_err(p, O.softmax(s, -1)) < 1e-5
    pm = nump(ops.softmax_rows_(cu(s).clone(), cu(keep), rows))
    sm = np.where(np.repeat(keep, rows, axis=0) > 0, s, np.float32(-1e9))
    assert rel_err(pm, O.softmax(sm, -1)) < 1e-5
    cs = nump(ops.colsum(cu(p), B))
    assert rel_err(cs, p.reshape(B, rows, n).sum(1)) < 1e-5
    vals = np.round(rs.randn(4, 768), 1).astype(np.float32)           # rounding => many ties
    idx, mask = ops.topk_select(cu(vals), 588, want_idx=True, want_mask=True)
    want = O.topk_desc(vals, 588)
    assert np.array_equal(nump(idx), want)
    wm = np.zeros_like(vals, dtype=np.uint8)
    np.put_along_axis(wm, want, 1, axis=1)
    assert np.array_equal(nump(mask), wm)
    idx, _ = ops.topk_select(cu(vals[:, :494]), 196)
    assert np.array_equal(nump(idx), O.topk_desc(vals[:, :494], 196))


@pytest.mark.parametrize("n,K", [(21, 20), (32, 7), (100, 64), (333, 200), (1024, 1024), (1500, 700), (4096, 2408)])
def test_topk_select_sizes_and_ties(n, K):
    """vcr_topk_select on both kernels (one key per thread with warp shuffles for n <= 1024, the shared-memory network above):
    sorted indices with ties broken by lower index, and the membership mask."""
    rs = np.random.RandomState(n + K)
    vals = np.round(rs.randn(5, n), 1).astype(np.float32)             # rounding => many ties
    vals[0, : n // 2] = 0.5                                           # one row dominated by a single tied value
    idx, mask = ops.topk_select(cu(vals), K, want_idx=True, want_mask=True)
    want = O.topk_desc(vals, K)
    assert np.array_equal(nump(idx), want)
    wm = np.zeros_like(vals, dtype=np.uint8)
    np.put_along_axis(wm, want, 1, axis=1)
    assert np.array_equal(nump(mask), wm)


def test_attention_vs_golden():
    g = load_golden("attention")
    q, k, v = g["q"], g["k"], g["v"]
    B, h, Nq, dk = q.shape
    Nk = k.shape[2]
    tok = lambda x: cu(x.transpose(0, 2, 1, 3).reshape(x.shape[0], x.shape[2], h * dk))
    qt, kt, vt = tok(q), tok(k), tok(v)
    D = h * dk
    out = torch.empty(B, Nq, D, device=DEV)
    args = ((qt, D, Nq * D, dk), (kt, D, Nk * D, dk), (vt, D, Nk * D, dk), B, h, Nq, Nk, dk, 1.0 / np.sqrt(dk),
            (out, D, Nq * D, dk))
    ops.attention_f32(*args)
    want = g["out"].transpose(0, 2, 1, 3).reshape(B, Nq, D)
    assert rel_err(nump(out), want) < 1e-5
    cs = ops.attention_f32(*args, colsum_out=True)
    assert rel_err(nump(cs), g["colsum"]) < 1e-5
    _, keep = ops.topk_select(cs, int(Nk * float(g["overlap2"])), want_idx=False, want_mask=True)
    assert np.array_equal(nump(keep).astype(bool), g["kept"])
    ops.attention_f32(*args, keep=keep)
    want = g["out_src"].transpose(0, 2, 1, 3).reshape(B, Nq, D)
    assert rel_err(nump(out), want) < 1e-5


# ---------------------------------------------------------------- LPDNet ------------------------------
def test_lpdnet_vs_golden(net_whole, precision):
    g = load_golden("lpdnet")
    emb = net_whole.emb_nn
    x = cu(g["x"])
    out = emb.forward_tokens(x, idx_feat=cu(g["idx_feat_s0"], torch.int32), idx_xyz=cu(g["idx_xyz"], torch.int32))
    assert rel_err(nump(out).transpose(0, 2, 1), g["out_s0"]) < TOL
    # free-running: the kNN on OUR 64-d features must equal the canonical kNN of those same features
    st = {}
    out_free = emb.forward_tokens(x, stages=st)
    f64 = nump(st["f64"]).transpose(0, 2, 1)
    assert np.array_equal(nump(st["idx_feat"]), canon.knn(f64, 20))
    assert np.array_equal(np.sort(nump(st["idx_xyz"]), -1), np.sort(g["idx_xyz"], -1))
    flips = (~(np.sort(nump(st["idx_feat"]), -1) == np.sort(g["idx_feat_s0"], -1)).all(-1)).mean()
    assert flips < 0.01, flips
    # module API: [B,3,N] -> [B,512,N]
    y = emb(x)
    assert tuple(y.shape) == (1, 512, 512) and y.is_contiguous()
    assert np.array_equal(nump(y), nump(out_free).transpose(0, 2, 1))


def test_lpdnet_slope02(precision):
    g = load_golden("lpdnet")
    lpd = load_golden("lpd_pretrained_weights")
    m = V.LPDNet(default_args(), negative_slope=0.2).to(DEV).eval()
    m.load_state_dict({k[len("emb_nn."):]: torch.from_numpy(v) for k, v in lpd.items()})
    out = m.forward_tokens(cu(g["x"]), idx_feat=cu(g["idx_feat_s02"], torch.int32),
                           idx_xyz=cu(g["idx_xyz"], torch.int32))
    assert rel_err(nump(out).transpose(0, 2, 1), g["out_s02"]) < TOL


@pytest.mark.parametrize("B,N,slope", [(3, 101, 0.0), (2, 1024, 0.2), (1, 21, 0.0)])
def test_edgeconv_dg_tensor_core_vs_simt(B, N, slope):
    """csrc/edgeconv_tc.cu (tcgen05, transposed DG2 GEMM) against the FP32 SIMT kernel and a float64 restatement
    of model/lpdnet_model.py:122-126 on ragged point counts (tiles of 8 points)."""
    rs = np.random.RandomState(5)
    pq = rs.randn(B, N, 256).astype(np.float32)
    idx = rs.randint(0, N, size=(B, N, 20)).astype(np.int32)
    w2 = (rs.randn(128, 128) / 11.0).astype(np.float32)
    b2 = rs.randn(128).astype(np.float32)
    e1 = pq[np.arange(B)[:, None, None], idx, :128] + pq[:, :, None, 128:]          # fp32, as the kernels do
    e1 = np.where(e1 >= 0, e1, e1 * np.float32(slope)).astype(np.float64)
    want1 = e1.max(2)
    e2 = e1 @ w2.T.astype(np.float64) + b2
    e2 = np.where(e2 >= 0, e2, e2 * slope)
    want2 = e2.max(2)
    for mode, tol in (("h3", 2e-6), ("fp16", 3e-3)):
        x1 = torch.full((B, N, 128), float("nan"), device=DEV)
        x2 = torch.full((B, N, 128), float("nan"), device=DEV)
        ops.edgeconv_dg_tc(cu(pq), cu(idx), cu(w2), cu(b2), slope, x1, x2, mode)
        assert np.array_equal(nump(x1), want1.astype(np.float32))
        assert rel_err(nump(x2), want2) < tol, mode
    y1, y2 = torch.empty((B, N, 128), device=DEV), torch.empty((B, N, 128), device=DEV)
    ops.edgeconv_dg(cu(pq), cu(idx), cu(w2), cu(b2), slope, y1, y2)
    assert np.array_equal(nump(y1), want1.astype(np.float32))
    assert rel_err(nump(y2), want2) < 2e-6


# ---------------------------------------------------------------- Transformer -------------------------
def test_transformer_vs_golden(net_whole, net_partial, precision):
    g = load_golden("transformer")
    sp, tp = net_whole.pointer(cu(g["src_emb"]), cu(g["tgt_emb"]))
    assert rel_err(nump(sp), g["src_p"]) < TOL and rel_err(nump(tp), g["tgt_p"]) < TOL
    sp, tp = net_partial.pointer(cu(g["src_emb"]), cu(g["tgt_emb"]))
    assert rel_err(nump(sp), g["src_p_partial"]) < TOL and rel_err(nump(tp), g["tgt_p_partial"]) < TOL


# ---------------------------------------------------------------- VCP head + SVD -----------------------
def _col_set_diff(a, b):
    n = 0
    for x, y in zip(a, b):
        n += len(set(map(tuple, x.T.tolist())) ^ set(map(tuple, y.T.tolist())))
    return n


def test_vcp_head_vs_golden(net_whole, net_partial, precision):
    g = load_golden("vcp_head")
    ov2 = float(g["overlap2"])
    s, c = net_whole.head(cu(g["src_emb"]), cu(g["tgt_emb"]), cu(g["src"]), cu(g["tgt"]))
    assert np.array_equal(nump(s), g["src"]) and rel_err(nump(c), g["corr_all"]) < TOL
    hp = net_partial.head
    so, seo, to, teo, _, _ = hp.selectCom(cu(g["src"]), cu(g["src_emb"]), cu(g["tgt"]), cu(g["tgt_emb"]), ov2)
    assert so.shape == g["sel_src"].shape and teo.shape == g["sel_tgt_emb"].shape
    assert _col_set_diff(nump(so), g["sel_src"]) <= 2 and _col_set_diff(nump(to), g["sel_tgt"]) <= 2
    s2, c2 = hp.getCopair(cu(g["sel_src"]), cu(g["sel_src_emb"]), cu(g["sel_tgt"]), cu(g["sel_tgt_emb"]), ov2)
    got = np.concatenate([nump(s2), nump(c2)], axis=1)
    want = np.concatenate([g["part_src"], g["part_corr"]], axis=1)
    assert got.shape == want.shape and _col_set_diff(got, want) <= 2


def _copair_all_f64(se, te, tgt):
    se, te, tgt = se.astype(np.float64), te.astype(np.float64), tgt.astype(np.float64)
    pd = 2.0 * np.matmul(se.transpose(0, 2, 1), te) - (se ** 2).sum(1)[:, :, None] - (te ** 2).sum(1)[:, None, :]
    pd -= pd.max(-1, keepdims=True)
    P = np.exp(pd)
    P /= P.sum(-1, keepdims=True)
    return np.matmul(tgt, P.transpose(0, 2, 1))


@pytest.mark.parametrize("B,Ns,Nt,D,common", [(2, 200, 333, 512, 0.0), (2, 200, 333, 512, 0.8), (3, 1024, 1024, 512, 0.8),
                                              (1, 129, 64, 128, 0.0), (2, 768, 500, 64, 0.3), (1, 2048, 4096, 512, 0.8)])
def test_fused_softcorr_vs_oracle_and_row_pass(B, Ns, Nt, D, common):
    """csrc/softcorr_tc.cu (GEMM + online softmax + weighted target sum in one kernel, no score matrix in HBM) against the
    numpy oracle's getCopairALL, a float64 evaluation of the same formula, and the materialised GEMM -> row-pass path,
    incl. ragged / unequal sizes.  common > 0: embeddings share a large component (|f|^2 >> gaps, the cancellation regime
    of the real network): there every fp32 evaluation -- the oracle's included -- carries ~|f|^2 * 2^-23
    of logit noise, so the bar is "as close to float64 as the fp32 oracle is", plus fused == row pass."""
    from vcr_net_b200 import config, functional as Fn
    rs = np.random.RandomState(B + Ns + Nt)
    base = rs.randn(1, D, 1).astype(np.float32) * common
    se = (base + 0.05 * rs.randn(B, D, Ns)).astype(np.float32)
    te = (base + 0.05 * rs.randn(B, D, Nt)).astype(np.float32)
    src = rs.rand(B, 3, Ns).astype(np.float32) - 0.5
    tgt = rs.rand(B, 3, Nt).astype(np.float32) - 0.5
    _, want = O.get_copair_all(src, se, tgt, te)
    truth = _copair_all_f64(se, te, tgt)
    s_tok, t_tok = cu(np.ascontiguousarray(se.transpose(0, 2, 1))), cu(np.ascontiguousarray(te.transpose(0, 2, 1)))
    old_p, old_f = config.precision, config.fused_softcorr
    try:
        config.set_precision("h3")
        config.fused_softcorr = True
        fused = nump(Fn.vcp_whole(s_tok, t_tok, cu(tgt)))
        config.fused_softcorr = False
        rows = nump(Fn.vcp_whole(s_tok, t_tok, cu(tgt)))
    finally:
        config.set_precision(old_p)
        config.fused_softcorr = old_f
    assert fused.shape == want.shape == (B, 3, Ns)
    assert rel_err(fused, rows) < 2e-5                       # same products, same pd op order; only the softmax schedule differs
    bar = max(TOL, 2.0 * rel_err(want, truth))               # fp32 oracle's own distance from float64
    assert rel_err(fused, truth) < bar and rel_err(rows, truth) < bar, (rel_err(fused, truth), rel_err(want, truth))
    if common == 0.0:
        assert rel_err(fused, want) < TOL


@pytest.mark.parametrize("B,Ns,Nt,D", [(2, 494, 494, 512), (3, 200, 333, 512), (1, 129, 64, 128)])
def test_fused_copair_argmax_matches_row_pass(B, Ns, Nt, D):
    """getCopair statistics (argmax_j, max_j P_ij) from the fused tensor-core kernel == the materialised GEMM -> row pass
    (same products and pd op order, so the integer outcome must be identical), incl. exact ties (duplicated targets)."""
    from vcr_net_b200 import functional as Fn
    rs = np.random.RandomState(B + Ns + Nt)
    base = rs.randn(1, 1, D).astype(np.float32) * 0.5
    s_tok = (base + 0.05 * rs.randn(B, Ns, D)).astype(np.float32)
    t_tok = (base + 0.05 * rs.randn(B, Nt, D)).astype(np.float32)
    t_tok[:, Nt // 2:Nt // 2 + 8] = t_tok[:, 3:11]                   # duplicated targets: exact ties -> lower index
    s_tok[:, :5] = t_tok[:, 3:8]                                      # sources that coincide with a duplicated target
    st, tt = cu(s_tok), cu(t_tok)
    xx, yy = ops.sqnorm_rows(st), ops.sqnorm_rows(tt)
    bi, bv = ops.softcorr_best_tc(ops.to_operand(st.reshape(B * Ns, D), "h3"), ops.to_operand(tt.reshape(B * Nt, D), "h3"),
                                  xx, yy, B, Ns, Nt, D)
    dot, ld = Fn.pair_dots(st, tt)
    _, ri, rv = ops.softcorr_rows(dot, ld, Ns, Nt, xx, yy, mode=2)
    assert torch.equal(bi, ri)
    assert rel_err(nump(bv), nump(rv)) < 2e-5
    assert (nump(bi)[:, :5] == np.arange(3, 8)[None]).all()           # the lower of the two equal targets


def test_svd_head_vs_golden(net_whole):
    g = load_golden("svd_head")
    R, t = net_whole.svd(cu(g["src"]), cu(g["corr"]))
    assert np.abs(nump(R) - g["R"]).max() < 1e-5 and np.abs(nump(t) - g["t"]).max() < 1e-5
    Ro, to = O.svd_head(g["src"], g["corr"])
    assert np.abs(nump(R) - Ro).max() < 1e-5 and np.abs(nump(t) - to).max() < 1e-5
    # exact recovery of a known rigid motion, incl. a big batch (size-independent property)
    p = synth.make_pairs(64, 300, first_item=100, aligned=True)   # index-aligned => exact rigid motion
    R2, t2, Rb, tb = ops.svd_head(cu(p["src"]), cu(p["tgt"]))
    assert np.abs(nump(R2) - p["R_ab"]).max() < 1e-5 and np.abs(nump(t2) - p["t_ab"]).max() < 1e-5
    assert np.abs(nump(Rb) - p["R_ab"].transpose(0, 2, 1)).max() < 1e-5
    assert np.abs(nump(tb) + np.einsum("bji,bj->bi", p["R_ab"], p["t_ab"])).max() < 1e-5


def test_pose_algebra():
    p = synth.make_pairs(5, 64, first_item=7)
    out = nump(V.transform_point_cloud(cu(p["src"]), cu(p["R_ab"]), cu(p["t_ab"])))
    assert rel_err(out, O.transform_point_cloud(p["src"], p["R_ab"], p["t_ab"])) < 1e-6
    q = np.random.RandomState(0).randn(5, 4).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    assert rel_err(nump(V.quat2mat(cu(q))), O.quat2mat(q)) < 1e-6
    out = nump(V.transform_point_cloud(cu(p["src"]), cu(q), cu(p["t_ab"])))
    assert rel_err(out, O.transform_point_cloud(p["src"], O.quat2mat(q), p["t_ab"])) < 1e-5


# ---------------------------------------------------------------- full network ------------------------
def test_vcrnet_whole_vs_golden(net_whole, precision):
    g = load_golden("vcrnet_whole")
    with torch.no_grad():
        out = V.vcrnetIter(net_whole, cu(g["src"]), cu(g["tgt"]), iter=1)
    for n, o in zip(("srcK", "corrK", "R_ab", "t_ab", "R_ba", "t_ba"), out):
        # free-running kNN: a flipped near-tie neighbour moves src_corr by more than fp32 noise
        # (SURVEY.md section 7 hard part 1), hence 5e-4 here; the staged tests above hold 1e-4.
        assert rel_err(nump(o), g[n]) < 5e-4, n
    g2 = load_golden("vcrnet_whole_iter2")
    out2 = V.vcrnetIter(net_whole, cu(g["src"]), cu(g["tgt"]), iter=2)
    assert rel_err(nump(out2[2]), g2["R_ab"]) < 5e-4 and rel_err(nump(out2[3]), g2["t_ab"]) < 5e-4


def test_vcrnet_whole_vs_oracle_seeded(net_whole, ckpt, precision):
    p = synth.make_pairs(2, 384, first_item=60)
    out = V.vcrnetIter(net_whole, cu(p["src"]), cu(p["tgt"]), iter=1)
    want = O.vcrnet_iter(ckpt, p["src"], p["tgt"], 1)
    for n, o, w in zip(("srcK", "corrK", "R_ab", "t_ab", "R_ba", "t_ba"), out, want):
        assert rel_err(nump(o), w) < 5e-4, n


def test_vcrnet_partial_vs_golden(net_partial, precision):
    g = load_golden("vcrnet_partial")
    out1 = V.vcrnetIter(net_partial, cu(g["src"]), cu(g["tgt"]), iter=1)
    assert tuple(out1[0].shape) == g["srcK1"].shape
    # hard correspondences: selection sets may differ at near-ties, poses must agree closely
    assert rel_err(nump(out1[2]), g["R_ab1"]) < 2e-3 and rel_err(nump(out1[3]), g["t_ab1"]) < 2e-3
    out3 = V.vcrnetIter(net_partial, cu(g["src"]), cu(g["tgt"]), iter=3)
    assert tuple(out3[0].shape) == g["srcK"].shape
    # the headline loop (iter=3) against the live reference's pose; hard selections => same 2e-3 bar as iter=1
    e3R, e3t = rel_err(nump(out3[2]), g["R_ab"]), rel_err(nump(out3[3]), g["t_ab"])
    print(f"[parity achieved] vcrnet_partial iter=3 ({precision}): R {e3R:.3g} t {e3t:.3g}; "
          f"iter=1: R {rel_err(nump(out1[2]), g['R_ab1']):.3g} t {rel_err(nump(out1[3]), g['t_ab1']):.3g}")
    assert e3R < 2e-3 and e3t < 2e-3, (e3R, e3t)
    assert rel_err(nump(out3[4]), g["R_ba"]) < 2e-3 and rel_err(nump(out3[5]), g["t_ba"]) < 2e-3
    R = nump(out3[2]).astype(np.float64)
    assert np.allclose(np.einsum("bij,bkj->bik", R, R), np.eye(3), atol=1e-5)
    assert np.allclose(np.linalg.det(R), 1.0, atol=1e-5)


def test_full_size_properties(net_whole, precision):
    """BASELINE cfg 1 size (B=16, N=1024): size-independent properties instead of an oracle run."""
    p = synth.make_pairs(16, 1024, first_item=200)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    out = V.vcrnetIter(net_whole, src, tgt, iter=1)
    R, t = nump(out[2]).astype(np.float64), nump(out[3]).astype(np.float64)
    assert np.isfinite(R).all() and np.isfinite(t).all()
    assert np.allclose(np.einsum("bij,bkj->bik", R, R), np.eye(3), atol=1e-5)
    assert np.allclose(np.linalg.det(R), 1.0, atol=1e-5)
    # inverse pose composes to identity
    Rb, tb = nump(out[4]).astype(np.float64), nump(out[5]).astype(np.float64)
    assert np.allclose(np.einsum("bij,bjk->bik", Rb, R), np.eye(3), atol=1e-5)
    assert np.allclose(np.einsum("bij,bj->bi", Rb, t) + tb, 0, atol=1e-5)
    # batch independence: pair 3 alone gives the same answer as pair 3 inside the batch
    one = V.vcrnetIter(net_whole, src[3:4], tgt[3:4], iter=1)
    assert np.abs(nump(one[2]) - nump(out[2])[3:4]).max() < 1e-5
    # permutation equivariance of the source cloud: R,t unchanged up to fp noise
    perm = torch.randperm(1024, device=DEV)
    outp = V.vcrnetIter(net_whole, src[:2][:, :, perm], tgt[:2], iter=1)
    assert np.abs(nump(outp[2]) - nump(out[2])[:2]).max() < 5e-4



def test_identity_pointer_vs_live_reference(ckpt, precision):
    """--pointer identity (model/vcrnet_model.py:477-478, 502-505): Identity returns its inputs, the residual
    add doubles both embeddings, the head logits scale by 4."""
    g = load_golden("vcrnet_identity_pointer")
    net = V.VCRNet(default_args(pointer="identity")).to(DEV).eval()
    res = net.load_state_dict(synth.checkpoint_to_torch(ckpt), strict=False)
    assert not res.missing_keys
    out = V.vcrnetIter(net, cu(g["src"]), cu(g["tgt"]), iter=1)
    assert rel_err(nump(out[1]), g["corrK"]) < 5e-4 and rel_err(nump(out[2]), g["R_ab"]) < 5e-4
    assert rel_err(nump(out[3]), g["t_ab"]) < 5e-4


def test_vcrnet_training_mode_raises_clearly(net_whole):
    """VCRNet is the inference path: train() + grad enabled must fail loudly up front, not late in backward."""
    p = synth.make_pairs(1, 128, first_item=3)
    net_whole.train()
    try:
        with torch.enable_grad(), pytest.raises(NotImplementedError, match="training is not implemented"):
            net_whole(cu(p["src"]), cu(p["tgt"]))
    finally:
        net_whole.eval()


@pytest.mark.parametrize("B,Ns,Nt", [(2, 768, 768), (3, 257, 331), (1, 40, 1024), (2, 100, 64)])
def test_select_stats_fused_two_read_pass(B, Ns, Nt):
    """vcr_select_stats: selectCom's two statistics (model/vcrnet_model.py:213-222, 243-244) straight from the score
    products, pd formed on the fly -- against fp64 numpy, incl. ragged sizes (element-wise load path) and the old
    negdist -> column-softmax / row-softmax kernels."""
    rs = np.random.RandomState(Ns + Nt)
    s = (rs.randn(B, Ns, 32) * 0.7).astype(np.float32)
    t = (rs.randn(B, Nt, 32) * 0.7).astype(np.float32)
    st, tt = cu(s), cu(t)
    dot, ld = ops.pair_dots(st, tt)
    xx, yy = ops.sqnorm_rows(st), ops.sqnorm_rows(tt)
    row_stat, col_stat = ops.select_stats(dot, ld, Ns, Nt, xx, yy)
    d64 = nump(dot)[:, :, :Nt].astype(np.float64)
    pd = (-nump(xx).astype(np.float64)[:, :, None] + 2.0 * d64) - nump(yy).astype(np.float64)[:, None, :]
    e_r = np.exp(pd - pd.max(axis=2, keepdims=True)); p_r = e_r / e_r.sum(axis=2, keepdims=True)
    e_c = np.exp(pd - pd.max(axis=1, keepdims=True)); p_c = e_c / e_c.sum(axis=1, keepdims=True)
    assert rel_err(nump(col_stat), p_r.sum(axis=1)) < 2e-5
    assert rel_err(nump(row_stat), p_c.sum(axis=2)) < 2e-5
    # the materialised path of round 1 (kept for the reference-named API) agrees
    pdm = ops.negdist_(dot.clone(), ld, Ns, Nt, xx, yy)
    old_row = ops.rowsum_colsoftmax(pdm, ld, Ns, Nt)
    P = ops.softmax_rows_(pdm.view(B * Ns, ld)[:, :Nt])
    old_col = ops.colsum(P, B)
    assert rel_err(nump(row_stat), nump(old_row)) < 2e-5 and rel_err(nump(col_stat), nump(old_col)) < 2e-5


def test_gather_operand_rows_matches_fp32_gather():
    """getCopair's inputs gathered in operand format: same bits as gather_rows + to_operand + sqnorm_rows of the fp32 rows."""
    rs = np.random.RandomState(5)
    B, N, D, K = 3, 200, 512, 77
    x = cu(rs.randn(B, N, D).astype(np.float32))
    idx = cu(np.stack([rs.permutation(N)[:K] for _ in range(B)]).astype(np.int32))
    op = ops.to_operand(x.view(B * N, D), "h3")
    sq = ops.sqnorm_rows(x)
    g_op, g_sq = ops.gather_operand_rows(op, sq, idx, B, N)
    rows = ops.gather_rows(x, idx)
    want = ops.to_operand(rows.view(B * K, D), "h3")
    assert torch.equal(g_op.buf, want.buf) and torch.equal(g_sq, ops.sqnorm_rows(rows))


def test_lpdnet_producers_write_operand_copies_bit_identically(net_whole):
    """Round 2: conv1+conv2 fused (vcr_lpd_point_mlp) and the operand-format outputs of edgeconv_dg_tc / gather_max carry
    exactly the bits of the unfused path (conv3_act + gemm, to_operand of the fp32 outputs)."""
    rs = np.random.RandomState(3)
    p = synth.make_pairs(2, 333, first_item=9)
    xyz = cu(p["src"])
    W = Fn.lpdnet_weights(net_whole.emb_nn)
    for slope in (0.0, 0.2):
        h1, h2, hop = ops.lpd_point_mlp(xyz, W["w1"], W["b1"], W["w2"], W["b2"], slope, want_h1=True, want_operand=True)
        r1 = ops.conv3_act(xyz, W["w1"], W["b1"], slope)
        r2 = ops.gemm(r1, W["w2"], W["b2"], act=1, slope=slope)
        assert torch.equal(h1, r1) and torch.equal(h2, r2)
        assert torch.equal(hop.buf, ops.to_operand(r2.view(-1, 64), "h3").buf)
    B, N = 2, 333
    pq = cu(rs.randn(B, N, 256).astype(np.float32))
    idx = cu(rs.randint(0, N, size=(B, N, 20)).astype(np.int32))
    x12 = torch.empty((B, N, 256), device=DEV)
    cat_op = ops.Operand.empty(B * N, 512, "h3", DEV)
    ops.edgeconv_dg_tc(pq, idx, W["dg2_w"], W["dg2_b"], 0.0, x12[:, :, 0:128], x12[:, :, 128:256], "h3",
                       op1=cat_op.cols_view(0, 128), op2=cat_op.cols_view(128, 128))
    want = ops.to_operand(x12.view(B * N, 256), "h3")
    assert torch.equal(cat_op.buf[:, :, :256], want.buf[:, :, :256])
    pq3 = cu(rs.randn(B, N, 512).astype(np.float32))
    x3 = torch.empty((B, N, 256), device=DEV)
    ops.gather_max(pq3[:, :, 0:256], pq3[:, :, 256:512], idx, 0.0, x3, op=cat_op.cols_view(256, 256))
    assert torch.equal(cat_op.buf[:, :, 256:512], ops.to_operand(x3.view(B * N, 256), "h3").buf[:, :, :256])

# ---------------------------------------------------------------- tensor-core GEMM --------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 520, 200), (4096, 1536, 512), (300, 64, 1024)])
def test_gemm_tc_h3_matches_fp64(M, N, K):
    rs = np.random.RandomState(M + N)
    a = rs.randn(M, K).astype(np.float32)
    w = (rs.randn(N, K) * 0.05).astype(np.float32)
    b = rs.randn(N).astype(np.float32)
    r = rs.randn(M, N).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    c = torch.empty(M, N, device=DEV)
    ops.gemm_tc(ops.to_operand(cu(a), "h3"), ops.to_operand(cu(w), "h3"), M, N, K, c=c)
    assert rel_err(nump(c), ref) < 3e-6                    # fp32-level (SIMT fp32 gives ~1e-6 here)
    h = ops.Operand.empty(M, N, "h3", DEV)
    ops.gemm_tc(ops.to_operand(cu(a), "h3"), ops.to_operand(cu(w), "h3"), M, N, K, bias=cu(b), act=1, slope=0.2,
                alpha=0.5, c=c, residual=cu(r), h=h, h_split=N)
    z = 0.5 * ref + b
    z = np.where(z >= 0, z, 0.2 * z)
    assert rel_err(nump(c), z + r) < 3e-6
    assert rel_err(nump(h.to_float()), z + r) < 3e-6       # operand-format output carries hi + lo/2^11
    for mode, tol in (("fp16", 2e-3), ("bf16", 2e-2)):
        ops.gemm_tc(ops.to_operand(cu(a), mode), ops.to_operand(cu(w), mode), M, N, K, c=c)
        assert rel_err(nump(c), ref) < tol


def test_gemm_tc_batched_heads_and_transposed_output():
    rs = np.random.RandomState(1)
    B, h, N, dk = 2, 4, 256, 128
    D = h * dk
    x = rs.randn(B * N, D).astype(np.float32)
    w = (rs.randn(3 * D, D) * 0.05).astype(np.float32)
    qkv_ref = x.astype(np.float64) @ w.astype(np.float64).T
    qk = ops.Operand.empty(B * N, 2 * D, "h3", DEV)
    vt = ops.Operand.empty(B * D, N, "h3", DEV)
    ops.gemm_tc(ops.to_operand(cu(x), "h3"), ops.to_operand(cu(w), "h3"), N, 3 * D, D, nbo=B, a_off=(N, 0, 0, 0),
                h=qk, h_strides=(N * qk.ld, 0), h_split=2 * D, ht=vt, ht_strides=(D * vt.ld, 0))
    assert rel_err(nump(qk.to_float()), qkv_ref[:, :2 * D]) < 3e-6
    v_ref = qkv_ref[:, 2 * D:].reshape(B, N, D).transpose(0, 2, 1).reshape(B * D, N)
    assert rel_err(nump(vt.to_float()), v_ref) < 3e-6
    S = torch.empty(B, h, N, N, device=DEV)
    ops.gemm_tc(qk.cols_view(0, D), qk.cols_view(D, D), N, N, dk, nbo=B, nbi=h, a_off=(N, 0, 0, dk),
                b_off=(N, 0, 0, dk), alpha=0.25, c=S, c_strides=(h * N * N, N * N))
    q = qkv_ref[:, :D].reshape(B, N, h, dk).transpose(0, 2, 1, 3)
    k = qkv_ref[:, D:2 * D].reshape(B, N, h, dk).transpose(0, 2, 1, 3)
    assert rel_err(nump(S), 0.25 * q @ k.transpose(0, 1, 3, 2)) < 3e-6


@pytest.mark.parametrize("mode,tol", [("fp16", 3e-3), ("bf16", 3e-2)])
def test_throughput_modes_reported_tolerance(net_whole, mode, tol):
    """Single-pass tensor-core modes: NOT parity modes; their measured tolerance is recorded here."""
    from vcr_net_b200 import config
    g = load_golden("transformer")
    old = config.precision
    config.set_precision(mode)
    try:
        sp, tp = net_whole.pointer(cu(g["src_emb"]), cu(g["tgt_emb"]))
    finally:
        config.set_precision(old)
    assert rel_err(nump(sp), g["src_p"]) < tol and rel_err(nump(tp), g["tgt_p"]) < tol


# ---------------------------------------------------------------- flash attention (tcgen05) -------------
def _attn_operands(q, k, v, mode):
    """[B,h,N,dk] numpy -> (Q, K row-major operands [B*N, h*dk], V^T operand [B*h*dk, Nk])."""
    B, h, Nq, dk = q.shape
    Nk = k.shape[2]
    tok = lambda x: cu(x.transpose(0, 2, 1, 3).reshape(x.shape[0] * x.shape[2], h * dk))
    vt = cu(v.transpose(0, 1, 3, 2).reshape(B * h * dk, Nk))
    return ops.to_operand(tok(q), mode), ops.to_operand(tok(k), mode), ops.to_operand(vt, mode)


@pytest.fixture(params=[1, 2, 3, 4], ids=["tile-ping-pong", "8-warps-per-tile", "operands-in-tmem", "16-warps-per-tile"])
def flash_warps(request):
    """The organisations of the flash attention kernel: two groups of 4 warps alternating key tiles, 8 warps on every tile with
    every operand in shared memory, 8 warps with Q and P in tensor memory, 16 warps on every tile."""
    old = ops.set_flash_warps(request.param)
    yield request.param
    ops.set_flash_warps(old)


@pytest.mark.parametrize("mode,tol", [("h3", 2e-5), ("fp16", 3e-3), ("bf16", 3e-2)])
def test_flash_attention_vs_golden(mode, tol, flash_warps):
    g = load_golden("attention")
    q, k, v = g["q"], g["k"], g["v"]
    B, h, Nq, dk = q.shape
    Nk = k.shape[2]
    Q, K, VT = _attn_operands(q, k, v, mode)
    out = ops.Operand.empty(B * Nq, h * dk, mode, DEV)
    ops.flash_attn_tc(Q, K, VT, out, B, h, Nq, Nk, dk, 1.0 / np.sqrt(dk))
    want = g["out"].transpose(0, 2, 1, 3).reshape(B * Nq, h * dk)
    assert rel_err(nump(out.to_float()), want) < tol
    keep = cu(g["kept"].astype(np.uint8))
    ops.flash_attn_tc(Q, K, VT, out, B, h, Nq, Nk, dk, 1.0 / np.sqrt(dk), keep=keep)
    want = g["out_src"].transpose(0, 2, 1, 3).reshape(B * Nq, h * dk)
    assert rel_err(nump(out.to_float()), want) < tol


@pytest.mark.parametrize("Nq,Nk,spread", [(1024, 1024, 1.0), (768, 768, 1.0), (300, 1000, 1.0), (256, 2048, 12.0)])
def test_flash_attention_shapes_and_rescale(Nq, Nk, spread, flash_warps):
    """Ragged tiles, many persistent work items per CTA and (spread = 12) logits whose running maximum
    keeps growing, which exercises the lazy O-rescale path; oracle = numpy attention."""
    rs = np.random.RandomState(Nq + Nk)
    B, h, dk = 3, 4, 128
    q = (rs.randn(B, h, Nq, dk) * spread).astype(np.float32)
    k = rs.randn(B, h, Nk, dk).astype(np.float32)
    k *= np.linspace(0.2, 1.5, Nk, dtype=np.float32)[None, None, :, None]      # later keys score higher
    v = rs.randn(B, h, Nk, dk).astype(np.float32)
    want, _ = O.attention(q, k, v)
    Q, K, VT = _attn_operands(q, k, v, "h3")
    out = ops.Operand.empty(B * Nq, h * dk, "h3", DEV)
    lse = torch.empty(B, h, Nq, device=DEV)
    ops.flash_attn_tc(Q, K, VT, out, B, h, Nq, Nk, dk, 1.0 / np.sqrt(dk), lse=lse)
    got = nump(out.to_float()).reshape(B, Nq, h, dk).transpose(0, 2, 1, 3)
    assert rel_err(got, want) < 3e-5
    sc = (q.astype(np.float64) @ k.astype(np.float64).transpose(0, 1, 3, 2)) / np.sqrt(dk)
    ref_lse = (np.log(np.exp(sc - sc.max(-1, keepdims=True)).sum(-1)) + sc.max(-1)) / np.log(2.0)
    assert np.abs(nump(lse) - ref_lse).max() < 1e-3 * max(1.0, np.abs(ref_lse).max())


@pytest.mark.parametrize("masked", [False, True])
def test_flash_attention_several_items_per_cta_into_a_dirty_buffer(masked, flash_warps):
    """More work items (10 x 4 x 6 = 240) than SMs, so every persistent CTA walks the item boundary code (next Q, O write-out,
    barrier phases), and an output buffer pre-filled with NaN, so a row that is not written cannot hide behind stale data."""
    rs = np.random.RandomState(11)
    B, h, Nq, Nk, dk = 10, 4, 768, 320, 128
    q = rs.randn(B, h, Nq, dk).astype(np.float32)
    k = rs.randn(B, h, Nk, dk).astype(np.float32)
    v = rs.randn(B, h, Nk, dk).astype(np.float32)
    kept = (rs.rand(B, Nk) < 0.7).astype(np.uint8) if masked else None
    sc = (q.astype(np.float64) @ k.astype(np.float64).transpose(0, 1, 3, 2)) / np.sqrt(dk)
    if masked:
        sc = np.where(kept[:, None, None, :] != 0, sc, -1e9)
    p = np.exp(sc - sc.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    want = (p @ v.astype(np.float64)).transpose(0, 2, 1, 3).reshape(B * Nq, h * dk)
    Q, K, VT = _attn_operands(q, k, v, "h3")
    out = ops.Operand.empty(B * Nq, h * dk, "h3", DEV)
    out.buf.view(torch.int16).fill_(0x7E00)                       # fp16 NaN in both planes
    ops.flash_attn_tc(Q, K, VT, out, B, h, Nq, Nk, dk, 1.0 / np.sqrt(dk), keep=cu(kept) if masked else None)
    got = nump(out.to_float())
    assert np.isfinite(got).all()
    assert rel_err(got, want) < 3e-5


# ---------------------------------------------------------------- embeddings / heads variants ----------
def _load_prefixed(module, g, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items()
          if k.startswith(prefix) and not k.endswith(".out")}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.endswith("num_batches_tracked") for m in missing), (missing, unexpected)


def test_dgcnn_pointnet_vs_golden(precision):
    g = load_golden("variants")
    x = cu(g["x"])
    net = V.DGCNN(emb_dims=128).to(DEV).eval()
    _load_prefixed(net, g, "dgcnn.")
    out = net(x, idx=cu(g["idx"], torch.int32))
    assert tuple(out.shape) == (2, 128, 256)
    assert rel_err(nump(out), g["dgcnn.out"]) < TOL
    assert rel_err(nump(net(x)), g["dgcnn.out"]) < 5e-4            # free-running kNN (near-tie flips, cf. LPDNet)
    pn = V.PointNet(emb_dims=128).to(DEV).eval()
    _load_prefixed(pn, g, "pointnet.")
    assert rel_err(nump(pn(x)), g["pointnet.out"]) < TOL
    with pytest.raises(RuntimeError):
        net.train()(x)                                             # batch-statistics BatchNorm is not implemented


def test_lpdnet_transform_nets_vs_golden(precision):
    """--t3d / --tfea TranformNets (model/lpdnet_model.py:19-70, 107-118), eval mode."""
    from vcr_net_b200.model.lpdnet_model import LPDNet
    g = load_golden("tnet")
    x = cu(g["x"])
    for name, t3d, tfea in (("both", True, True), ("t3d", True, False)):
        net = LPDNet(default_args(t3d=t3d, tfea=tfea, emb_dims=128)).to(DEV).eval()
        net.load_state_dict(synth.checkpoint_to_torch(synth.make_tnet_lpdnet_weights(21, t3d, tfea, 128)), strict=False)
        with torch.no_grad():
            assert rel_err(nump(net.t_net3d(x)), g[f"{name}.trans"]) < TOL
            st = {}
            tok = net.forward_tokens(x, idx_feat=cu(g[f"{name}.idx_feat"], torch.int32),
                                     idx_xyz=cu(g["idx_xyz"], torch.int32), stages=st)
            if tfea:
                assert rel_err(nump(st["trans_feat"]), g[f"{name}.trans_feat"]) < TOL
            assert rel_err(nump(tok).transpose(0, 2, 1), g[f"{name}.out"]) < TOL
            out = net(x)                                                # free-running kNN
            assert tuple(out.shape) == (2, 128, 256)
            assert rel_err(nump(out), g[f"{name}.out"]) < 5e-4
            assert torch.equal(st["idx_xyz"], cu(g["idx_xyz"], torch.int32))
            with pytest.raises(RuntimeError):
                net.train()(x)                                         # batch-statistics BatchNorm1d is not implemented
            net.eval()
    # a 1024-point cloud exercises the two-level cloud max and the oracle on a second size
    p = synth.make_tnet_lpdnet_weights(21, True, True, 128)
    x2 = synth.make_pairs(3, 1000, first_item=150)["tgt"]
    net = LPDNet(default_args(t3d=True, tfea=True, emb_dims=128)).to(DEV).eval()
    net.load_state_dict(synth.checkpoint_to_torch(p), strict=False)
    want, wst = O.lpdnet_forward(p, x2, prefix="", t3d=True, tfea=True, return_stages=True)
    with torch.no_grad():
        tok = net.forward_tokens(cu(x2), idx_feat=cu(wst["idx_feat"], torch.int32), idx_xyz=cu(wst["idx_xyz"], torch.int32))
    assert rel_err(nump(tok).transpose(0, 2, 1), want) < TOL


def test_vcp_att_dist_heads_vs_golden(precision):
    g = load_golden("variants")
    args = default_args(emb_dims=64, vcp_nn="att")
    att = V.VcpAtt(args).to(DEV).eval()
    att.load_state_dict({k[len("att."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("att.")}, strict=True)
    se, te, s, t = cu(g["h_src_emb"]), cu(g["h_tgt_emb"]), cu(g["h_src"]), cu(g["h_tgt"])
    s_out, c = att(se, te, s, t)
    assert torch.equal(s_out, s) and rel_err(nump(c), g["att_corr"]) < TOL
    _, c = V.VcpByDis(args)(se, te, s, t)
    assert rel_err(nump(c), g["dist_corr"]) < TOL


def test_vcrnet_with_dgcnn_and_att_head_runs(ckpt):
    """--emb_nn dgcnn --vcp_nn att assembles and registers a pair (random init; checks wiring + a proper rotation)."""
    torch.manual_seed(3)
    net = V.VCRNet(default_args(emb_nn="dgcnn", vcp_nn="att")).to(DEV).eval()
    p = synth.make_pairs(2, 256, first_item=9)
    out = V.vcrnetIter(net, cu(p["src"]), cu(p["tgt"]), iter=2)
    R = nump(out[2])
    assert np.allclose(np.matmul(R, R.transpose(0, 2, 1)), np.eye(3)[None], atol=1e-5)
    assert np.allclose(np.linalg.det(R), 1.0, atol=1e-5)


def test_icp_vs_golden_device_side_convergence():
    """ICP (model/icp_model.py) with the convergence test on the device: same result as the reference's early break."""
    g = load_golden("variants")
    src, dst = cu(g["icp_in_src"]), cu(g["icp_in_dst"])
    for mi, want_iters in ((10, 3), (2, 2)):
        icp = V.ICP(max_iterations=mi).to(DEV)
        s0, s1, R, t, R_ba, t_ba = icp(src, dst)
        assert icp.iterations_run() == want_iters
        assert torch.equal(s0, src)
        for got, key in zip((s1, R, t, R_ba, t_ba), ("src", "R", "t", "R_ba", "t_ba")):
            assert np.abs(nump(got) - g[f"icp{mi}_{key}"]).max() < 5e-6, (mi, key)
    # nearest-neighbour kernel vs the oracle's arg-max (ties -> lower index), ragged sizes
    rng = np.random.RandomState(3)
    a = rng.rand(2, 3, 301).astype(np.float32) - 0.5
    b = rng.rand(2, 3, 1111).astype(np.float32) - 0.5
    err = torch.zeros(1, dtype=torch.float64, device=DEV)
    corr, idx = ops.icp_nearest(cu(a), cu(b), err, want_idx=True)
    pd = O.neg_sqdist_cross(a, b)
    want = pd.argmax(axis=2)
    same = nump(idx) == want
    assert same.mean() > 0.995                       # fp32 near-ties between matmul orders only
    assert abs(float(err) - float(pd.max(axis=2).astype(np.float64).sum())) < 1e-3
    assert np.array_equal(nump(corr)[0][:, same[0]], b[0][:, want[0][same[0]]])


def test_vcrnet_icp_net_refines_pose(net_whole):
    """--iter=0 path (vcrnetIcpNet, model/vcrnet_model.py:46-62): composes the network pose with the ICP refinement."""
    from types import SimpleNamespace
    p = synth.make_pairs(2, 256, first_item=40)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    out = V.vcrnetIcpNet(SimpleNamespace(max_iterations=5), net_whole, src, tgt)
    want = O.vcrnet_forward(synth_ckpt_cache(), p["src"], p["tgt"])
    ts = O.transform_point_cloud(p["src"], want[2], want[3])
    oi = O.icp_forward(ts, p["tgt"], max_iterations=5)
    R_f = np.matmul(oi[2], want[2])
    t_f = np.matmul(oi[2], want[3][:, :, None])[:, :, 0] + oi[3]
    assert rel_err(nump(out[2]), R_f) < 2e-3 and rel_err(nump(out[3]), t_f) < 2e-3
    R = nump(out[2])
    assert np.allclose(np.matmul(R, R.transpose(0, 2, 1)), np.eye(3)[None], atol=1e-5)


_CK = {}


def synth_ckpt_cache():
    if "ck" not in _CK:
        _CK["ck"] = synth.make_checkpoint(1234, emb_weights=dict(load_golden("lpd_pretrained_weights")))
    return _CK["ck"]


@pytest.mark.parametrize("B,rpb,n", [(3, 40, 333), (2, 4 * 768, 768), (2, 130, 1024), (1, 70, 4096)])
def test_fused_softmax_colsum(B, rpb, n):
    """Single-read fused statistic (partial-overlap key selection, model/transformer.py:35-39) vs float64, and vs the
    two-kernel path it replaces."""
    rs = np.random.RandomState(B + n)
    ld = (n + 3) // 4 * 4
    s = np.zeros((B * rpb, ld), dtype=np.float32)
    s[:, :n] = (rs.randn(B * rpb, n) * 3).astype(np.float32)
    S = cu(s)
    want = np.exp(s[:, :n].astype(np.float64) - s[:, :n].max(-1, keepdims=True))
    want = (want / want.sum(-1, keepdims=True)).reshape(B, rpb, n).sum(1)
    got = nump(ops.colsum_softmax(S, n, B, fused=True))
    old = nump(ops.colsum_softmax(S, n, B, fused=False))
    assert rel_err(got, want) < 1e-5 and rel_err(old, want) < 1e-5
    assert np.array_equal(nump(ops.colsum_softmax(S, n, B, fused=True)), got)      # deterministic


@pytest.mark.parametrize("B,H,Nq,Nk", [(2, 4, 200, 333), (3, 4, 768, 768), (1, 2, 129, 64), (1, 4, 1024, 2048), (2, 1, 50, 130)])
def test_attn_colsum_tc_two_sweep_statistic(B, H, Nq, Nk):
    """csrc/attn_colsum_tc.cu (partial-overlap key statistic, model/transformer.py:33-39, no score matrix in HBM) vs a
    float64 evaluation and vs the materialised score-GEMM -> column-sum path; ragged / unequal sizes; deterministic."""
    import math
    dk = 128
    rs = np.random.RandomState(B * 7 + Nq + Nk)
    q = rs.randn(B * Nq, H * dk).astype(np.float32)
    k = rs.randn(B * Nk, H * dk).astype(np.float32)
    scale = 1.0 / math.sqrt(dk)
    q4 = q.astype(np.float64).reshape(B, Nq, H, dk).transpose(0, 2, 1, 3)
    k4 = k.astype(np.float64).reshape(B, Nk, H, dk).transpose(0, 2, 1, 3)
    sc = np.matmul(q4, k4.transpose(0, 1, 3, 2)) * scale
    pr = np.exp(sc - sc.max(-1, keepdims=True))
    want = (pr / pr.sum(-1, keepdims=True)).sum(axis=(1, 2))                       # [B, Nk]
    qo, ko = ops.to_operand(cu(q), "h3"), ops.to_operand(cu(k), "h3")
    got = ops.attn_colsum_tc(qo, ko, B, H, Nq, Nk, dk, scale)
    assert tuple(got.shape) == (B, Nk)
    assert rel_err(nump(got), want) < 1e-5
    assert torch.equal(ops.attn_colsum_tc(qo, ko, B, H, Nq, Nk, dk, scale), got)   # deterministic
    # the path it replaces: scaled scores through the GEMM kernel, then the single-read column-sum kernel
    ldS = (Nk + 3) // 4 * 4
    S = torch.zeros((B, H, Nq, ldS), dtype=torch.float32, device=DEV)
    ops.gemm_tc(qo, ko, Nq, Nk, dk, nbo=B, nbi=H, a_off=(Nq, 0, 0, dk), b_off=(Nk, 0, 0, dk), alpha=scale, c=S,
                c_strides=(H * Nq * ldS, Nq * ldS))
    old = ops.colsum_softmax(S.view(B * H * Nq, ldS), Nk, B)
    assert rel_err(nump(got), nump(old)) < 1e-5


@pytest.mark.parametrize("which", ["partial", "whole"])
def test_vcrnet_iter_hoisting_is_bit_identical(net_partial, net_whole, which):
    """config.hoist lifts everything that depends on the target cloud alone out of the --iter loop (emb_nn(tgt),
    encoder(tgt_emb) + its K/V projections, the decoder's first self-attention sublayer on tgt): same bits out at every
    level as recomputing it per iteration like the reference (model/vcrnet_model.py:24-28)."""
    from vcr_net_b200 import config
    net = net_partial if which == "partial" else net_whole
    p = synth.make_pairs(3, 512, partial=which == "partial", first_item=21)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    old = config.hoist
    outs = {}
    try:
        for level in ("none", "emb", "all"):
            config.hoist = level
            outs[level] = V.vcrnetIter(net, src, tgt, iter=3)
    finally:
        config.hoist = old
    for level in ("emb", "all"):
        for a, b in zip(outs["none"], outs[level]):
            assert torch.equal(a, b), level


def test_vcrnet_iter_hoisting_two_blocks_is_bit_identical():
    """--n_blocks 2 (util/initPara.py:176): only decoder layer 0's self-attention on tgt is loop-invariant, the K / V
    projections of encoder(tgt) are hoisted for every decoder layer; same bits as the plain loop."""
    from vcr_net_b200 import config
    torch.manual_seed(5)
    net = V.VCRNet(default_args(partial=True, overlap2=synth.OVERLAP2_0575, n_blocks=2)).to(DEV).eval()
    p = synth.make_pairs(2, 384, partial=True, first_item=44)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    old = config.hoist
    try:
        config.hoist = "none"
        base = V.vcrnetIter(net, src, tgt, iter=2)
        config.hoist = "all"
        hoisted = V.vcrnetIter(net, src, tgt, iter=2)
    finally:
        config.hoist = old
    for a, b in zip(base, hoisted):
        assert torch.equal(a, b)


def test_attention_probabilities_recorded_on_request(net_whole):
    """MultiHeadedAttention.attn (model/transformer.py:216-219, plot-only) is opt-in: module.record_attn = True fills it
    with the head-summed probabilities [B, Nq, Nk]; rows sum to the number of heads."""
    p = synth.make_pairs(2, 256, first_item=31)
    mha = net_whole.pointer.model.decoder.layers[0].src_attn
    assert mha.attn is None
    mha.record_attn = True
    try:
        V.vcrnetIter(net_whole, cu(p["src"]), cu(p["tgt"]), iter=1)
        a = mha.attn
        assert tuple(a.shape) == (4, 256, 256)                      # both directions run as one batch of 2B
        assert torch.allclose(a.sum(dim=-1), torch.full((4, 256), 4.0, device=a.device), atol=1e-4)
    finally:
        mha.record_attn = False
        mha.attn = None


# ---------------------------------------------------------------- data step + metrics -------------
@pytest.mark.parametrize("partial,aligned", [(False, False), (True, False), (False, True)])
def test_device_data_step_matches_host_generator(partial, aligned):
    """PairGenerator (device gather + fp64 transform + nearest-to-last crop) vs oracle/synth.make_pairs, the host
    restatement of util/data.py:247-329: identical fp32 clouds, same order."""
    from vcr_net_b200.data import PairGenerator
    n_items, first, N = 6, 3, 512
    base = np.random.RandomState(1234).rand(first + n_items, 2048, 3).astype(np.float32) - 0.5   # synth.make_pairs' base
    gen = PairGenerator(cu(base), num_points=N, partial=partial, reserve=synth.RESERVE_0575, aligned=aligned)
    got = gen.batch(range(first, first + n_items))
    want = synth.make_pairs(n_items, N, partial=partial, reserve=synth.RESERVE_0575, first_item=first, aligned=aligned)
    assert tuple(got["src"].shape) == want["src"].shape
    for k in ("src", "tgt"):
        d = np.abs(nump(got[k]) - want[k])
        assert d.max() <= 6e-8, (k, d.max())                       # fp64 transform, <= 1 fp32 ulp after the cast
        assert (d > 0).mean() < 1e-3
    assert np.allclose(nump(got["R_ab"]), want["R_ab"], atol=1e-7) and np.allclose(nump(got["t_ab"]), want["t_ab"], atol=1e-7)
    assert np.allclose(nump(got["euler_ab"]), want["euler_ab"], atol=1e-7)


def test_eval_accumulator_vs_oracle(net_whole):
    from vcr_net_b200.data import EvalAccumulator
    acc = EvalAccumulator(DEV)
    want = np.zeros(7)
    for first in (0, 4):
        p = synth.make_pairs(4, 256, first_item=first)
        src, tgt = cu(p["src"]), cu(p["tgt"])
        out = V.vcrnetIter(net_whole, src, tgt, iter=1)
        acc.update(src, tgt, out[0], out[1], cu(p["R_ab"]), cu(p["t_ab"]), out[2], out[3], out[4], out[5])
        want += np.array(O.eval_metrics_batch(p["src"], p["tgt"], nump(out[0]), nump(out[1]), p["R_ab"], p["t_ab"],
                                              nump(out[2]), nump(out[3]), nump(out[4]), nump(out[5])), dtype=np.float64)
    res = acc.result()
    assert res["num_examples"] == 8
    for i, k in enumerate(acc.KEYS):
        assert abs(res[k] - want[i] / 8) <= 1e-5 * max(abs(want[i] / 8), 1e-6), (k, res[k], want[i] / 8)


# ---------------------------------------------------------------- ragged / full-size cases ------------------------
def test_vcrnet_ragged_sizes_vs_oracle(net_whole, ckpt, precision):
    """Odd point counts (not multiples of any tile), a single pair, and clouds of different sizes."""
    p = synth.make_pairs(1, 333, first_item=70)
    out = V.vcrnetIter(net_whole, cu(p["src"]), cu(p["tgt"]), iter=1)
    want = O.vcrnet_iter(ckpt, p["src"], p["tgt"], 1)
    for n, o, w in zip(("srcK", "corrK", "R_ab", "t_ab"), out, want):
        assert rel_err(nump(o), w) < 5e-4, n
    q = synth.make_pairs(2, 300, first_item=80)
    tgt_short = np.ascontiguousarray(q["tgt"][:, :, :257])                 # Ns = 300, Nt = 257
    out = V.vcrnetIter(net_whole, cu(q["src"]), cu(tgt_short), iter=1)
    want = O.vcrnet_iter(ckpt, q["src"], tgt_short, 1)
    assert tuple(out[1].shape) == (2, 3, 300)
    for n, o, w in zip(("srcK", "corrK", "R_ab", "t_ab"), out, want):
        assert rel_err(nump(o), w) < 5e-4, n


def test_partial_full_size_properties(net_partial):
    """BASELINE cfg 2 size (B=24, 768 of 1024 points, iter=3): shapes of the hard-correspondence sets, proper rotations,
    inverse pose, batch independence."""
    p = synth.make_pairs(24, 1024, partial=True, first_item=400)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    assert src.shape[2] == 768
    out = V.vcrnetIter(net_partial, src, tgt, iter=3)
    M = int(int(768 * 0.84 * synth.OVERLAP2_0575) * 0.52 * synth.OVERLAP2_0575)
    assert tuple(out[0].shape) == (24, 3, M) and tuple(out[1].shape) == (24, 3, M)
    R, t = nump(out[2]).astype(np.float64), nump(out[3]).astype(np.float64)
    assert np.isfinite(R).all() and np.isfinite(t).all()
    assert np.allclose(np.einsum("bij,bkj->bik", R, R), np.eye(3), atol=1e-5)
    assert np.allclose(np.linalg.det(R), 1.0, atol=1e-5)
    Rb, tb = nump(out[4]).astype(np.float64), nump(out[5]).astype(np.float64)
    assert np.allclose(np.einsum("bij,bjk->bik", Rb, R), np.eye(3), atol=1e-5)
    assert np.allclose(np.einsum("bij,bj->bi", Rb, t) + tb, 0, atol=1e-5)
    one = V.vcrnetIter(net_partial, src[5:6], tgt[5:6], iter=3)
    assert np.abs(nump(one[2]) - nump(out[2])[5:6]).max() < 1e-5
    # iteration 1: every selected source point is one of the input points, every correspondence one of the target points
    o1 = V.vcrnetIter(net_partial, src[:2], tgt[:2], iter=1)
    for b in range(2):
        src_set = {tuple(c) for c in p["src"][b].T.tolist()}
        tgt_set = {tuple(c) for c in p["tgt"][b].T.tolist()}
        sel = nump(o1[0])[b].T.tolist()
        assert all(tuple(c) in src_set for c in sel) and len({tuple(c) for c in sel}) == M
        assert all(tuple(c) in tgt_set for c in nump(o1[1])[b].T.tolist())


def test_cfg4_size_4096_points(net_whole):
    """BASELINE cfg 4 point count (4096 per cloud; 4 pairs here): the path runs at that size, kNN stays bit-exact."""
    p = synth.make_pairs(4, 4096, first_item=500, base_points=4096)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    out = V.vcrnetIter(net_whole, src, tgt, iter=1)
    R = nump(out[2]).astype(np.float64)
    assert np.isfinite(R).all() and np.allclose(np.einsum("bij,bkj->bik", R, R), np.eye(3), atol=1e-5)
    assert np.array_equal(nump(V.knn(src[:1], 20)), canon.knn(p["src"][:1], 20))
    one = V.vcrnetIter(net_whole, src[2:3], tgt[2:3], iter=1)
    assert np.abs(nump(one[2]) - nump(out[2])[2:3]).max() < 1e-5


def test_gemm_f32_ragged_k():
    """K not a multiple of 4 (odd key counts in fp32-mode attention): rows padded to 16 bytes, tail read element-wise."""
    rs = np.random.RandomState(9)
    M, N, K = 70, 48, 333
    a = np.zeros((M, 336), dtype=np.float32); a[:, :K] = rs.randn(M, K); a[:, K:] = np.nan     # padding must never be read
    w = np.zeros((N, 336), dtype=np.float32); w[:, :K] = rs.randn(N, K); w[:, K:] = np.nan
    got = nump(ops.gemm(cu(a)[:, :K], cu(w)[:, :K]))
    want = a[:, :K].astype(np.float64) @ w[:, :K].astype(np.float64).T
    assert np.isfinite(got).all() and rel_err(got, want) < 1e-5
    v = rs.randn(K, 64).astype(np.float32)                                                      # [K, N] layout
    got = nump(ops.gemm(cu(a)[:, :K], cu(v), b_layout=1))
    assert rel_err(got, a[:, :K].astype(np.float64) @ v.astype(np.float64)) < 1e-5


# ---------------------------------------------------------------- CUDA-graph replay, CTA-pair GEMM ---------------------
@pytest.mark.parametrize("partial,B,N,it", [(False, 2, 512, 1), (True, 1, 1024, 3)])
def test_graphed_registration_is_bit_identical(net_whole, net_partial, partial, B, N, it):
    """vcr_net_b200.graph.GraphedRegistration: the whole --iter loop as ONE CUDA graph (the path has no host sync, no
    data-dependent shape, no hidden allocation); replay on new inputs == eager path, bit for bit."""
    from vcr_net_b200.graph import GraphedRegistration
    net = net_partial if partial else net_whole
    p = synth.make_pairs(2 * B, N, partial=partial, first_item=61)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    reg = GraphedRegistration(net, batch=B, num_points=src.shape[2], iter=it)
    for lo in (0, B):                                             # two different batches through the same graph
        with torch.no_grad():
            eager = V.vcrnetIter(net, src[lo:lo + B], tgt[lo:lo + B], iter=it)
        got = reg(src[lo:lo + B], tgt[lo:lo + B])
        for a, b in zip(eager, got):
            assert torch.equal(a, b)
    with pytest.raises(ValueError):
        reg(src[:B, :, :100], tgt[:B])


def test_vcrnet_iter_transparent_graph_cache(net_partial):
    """config.cuda_graph (VCR_CUDA_GRAPH=1): the reference-facing vcrnetIter serves repeated shapes from a captured graph --
    first call eager, second captures, later ones replay; fresh result tensors, same bits; an in-place weight update
    drops the cache."""
    from vcr_net_b200 import config
    p = synth.make_pairs(8, 512, partial=True, first_item=33)
    src, tgt = cu(p["src"]), cu(p["tgt"])
    was = config.cuda_graph
    config.cuda_graph = False
    eager = [V.vcrnetIter(net_partial, src[i:i + 2], tgt[i:i + 2], iter=3) for i in (0, 2, 4, 6)]
    config.cuda_graph = True
    try:
        got = [V.vcrnetIter(net_partial, src[i:i + 2], tgt[i:i + 2], iter=3) for i in (0, 2, 4, 6)]
        cache = net_partial.__dict__["_vcr_graph_cache"]
        assert len(cache["entries"]) == 1 and next(iter(cache["entries"].values())) != "seen"
        for e, g in zip(eager, got):
            for a, b in zip(e, g):
                assert torch.equal(a, b)
        assert got[2][2].data_ptr() != got[3][2].data_ptr()            # results are not the graph's static buffers
        with torch.no_grad():
            next(net_partial.parameters()).mul_(1.0)                    # bumps the parameter version
        again = V.vcrnetIter(net_partial, src[:2], tgt[:2], iter=3)
        assert next(iter(net_partial.__dict__["_vcr_graph_cache"]["entries"].values())) == "seen"
        for a, b in zip(eager[0], again):
            assert torch.equal(a, b)
        # a REPLACED parameter tensor (same version counter, new storage) must drop the cache too
        V.vcrnetIter(net_partial, src[:2], tgt[:2], iter=3)
        assert next(iter(net_partial.__dict__["_vcr_graph_cache"]["entries"].values())) != "seen"
        prm = net_partial.pointer.model.decoder.norm.b_2
        prm.data = prm.data.clone()
        V.vcrnetIter(net_partial, src[:2], tgt[:2], iter=3)
        assert next(iter(net_partial.__dict__["_vcr_graph_cache"]["entries"].values())) == "seen"
    finally:
        config.cuda_graph = was
        net_partial.__dict__.pop("_vcr_graph_cache", None)


@pytest.mark.parametrize("M,N,K,nbo", [(256, 128, 64, 1), (300, 200, 72, 1), (494, 494, 512, 3), (1000, 130, 520, 1),
                                       (64, 64, 64, 5), (2048, 1536, 512, 2)])
def test_gemm_cta_pair_is_bit_identical(M, N, K, nbo):
    """cta_group::2 variant of the h3 GEMM (256 x 128 tile per SM pair, B tile split across the pair): same bits as the
    single-CTA kernel for fp32 / operand-format / residual / activation epilogues, ragged M, N, K and batches."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(nbo * M, K, generator=g).to(DEV)
    w = torch.randn(N, K, generator=g).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    R = torch.randn(nbo * M, N, generator=g).to(DEV)
    A, Bm = ops.to_operand(a, "h3"), ops.to_operand(w, "h3")

    def run(pair):
        old = ops.set_gemm_pair(pair)
        try:
            kw = dict(bias=bias, act=1, slope=0.2)
            if nbo > 1:
                kw.update(nbo=nbo, a_off=(M, 0, 0, 0))
            c = torch.full((nbo * M, N), float("nan"), device=DEV)
            ops.gemm_tc(A, Bm, M, N, K, c=c, residual=R, c_strides=(M * N, 0) if nbo > 1 else (0, 0),
                        r_strides=(M * N, 0) if nbo > 1 else (0, 0), **kw)
            h = ops.Operand.empty(nbo * M, N, "h3", DEV)
            h.buf.zero_()
            ops.gemm_tc(A, Bm, M, N, K, h=h, h_split=N, h_strides=(M * h.ld, 0), **kw)
            torch.cuda.synchronize()
            return c, h.buf.clone()
        finally:
            ops.set_gemm_pair(old)

    c0, h0 = run(False)
    c1, h1 = run(True)
    assert torch.isfinite(c0).all()
    assert torch.equal(c0, c1) and torch.equal(h0.view(torch.uint8), h1.view(torch.uint8))
    ref = (a.double().view(nbo, M, K) @ w.double().T + bias.double()).view(nbo * M, N)
    ref = torch.where(ref >= 0, ref, 0.2 * ref) + R.double()
    assert rel_err(nump(c1), nump(ref)) < 1e-5
