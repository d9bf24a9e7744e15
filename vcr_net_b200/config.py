"""Run-time precision policy of the matrix engine.

  "fp32"  hand-written FP32 SIMT GEMMs everywhere (reference-exact arithmetic type)
  "h3"    tcgen05 tensor cores with 3-term fp16 split operands: fp32-level accuracy (the parity mode
          on tensor cores)
  "fp16" / "bf16"  single-pass tensor-core throughput modes (reported separately with their tolerance)
kNN / FPS / top-K selections and the SVD head always return the exact fp32 / fp64 result (the tensor-core kNN
prefilter only narrows the candidates; the kept neighbours are re-ranked with the canonical fp32 chain).
"""
import os

precision = os.environ.get("VCR_PRECISION", "h3")
VALID = ("fp32", "h3", "fp16", "bf16")
# flash attention kernel (attn_tc.cu) vs materialised scores through the GEMM kernel (tensor-core modes only)
flash_attention = os.environ.get("VCR_FLASH", "1") != "0"
# vcrnetIter (--iter > 1): what is hoisted out of the refinement loop.  The target cloud never changes inside the loop
# (model/vcrnet_model.py:24-28), so emb_nn(tgt), encoder(tgt_emb) with its K / V projections in the decoder's src_attn, and
# the decoder's first self-attention sublayer on tgt are loop-invariant (functional.TargetInvariants).  Outputs are
# bit-identical at every level (tests/test_gpu_parity.py::test_vcrnet_iter_hoisting_is_bit_identical):
#   "all" (default)  everything above, computed once per call
#   "emb"            only emb_nn(tgt)
#   "none"           recompute everything every iteration, exactly the work the reference does (bench.py reports this as
#                    the labelled variant `variant_no_hoisting`)
hoist = os.environ.get("VCR_HOIST", "all")
assert hoist in ("all", "emb", "none"), hoist


# feature-space kNN (16 <= D <= 128): tcgen05 prefilter + exact re-rank (csrc/knn.cu, bit-identical indices).
# Measured on B200 (scripts/knn_bench.py, D = 64, profiles/r01_knn_tc_prefilter_v6.txt): 0.94x of the FP32 SIMT kernel at
# N = 1024 (two 512-candidate chunks per CTA do not amortise the TMA -> MMA -> TMEM latency chain and the re-rank),
# 1.47x at N = 4096, 1.85x at N = 16384.  "auto" (default) therefore uses it from N >= 2048 in the tensor-core precision
# modes; "1" always, "0" never.
knn_tc = os.environ.get("VCR_KNN_TC", "auto")
KNN_TC_MIN_N = 2048


def use_knn_tc(N: int) -> bool:
    if precision == "fp32" or knn_tc in ("0", False):
        return False
    return True if knn_tc in ("1", True) else N >= KNN_TC_MIN_N


# partial-overlap key statistic (model/transformer.py:35-39): the materialised score chunk [cb, h, Nq, Nk] fp32 that
# feeds the softmax column sums.  Sized to stay L2-resident between the score GEMM and the column-sum kernel.
stat_chunk_bytes = int(float(os.environ.get("VCR_STAT_CHUNK_MB", "3072")) * (1 << 20))


# whole-to-whole VCP head (getCopairALL): fused tcgen05 GEMM + online softmax + weighted target sum (csrc/softcorr_tc.cu)
# instead of GEMM -> HBM score matrix -> row pass (tensor-core precision modes)
fused_softcorr = os.environ.get("VCR_FUSED_SOFTCORR", "1") != "0"


# partial-overlap key statistic: two-sweep tcgen05 kernel (csrc/attn_colsum_tc.cu) instead of score GEMM -> HBM -> column sums
fused_key_stat = os.environ.get("VCR_FUSED_KEY_STAT", "1") != "0"


def set_precision(p: str):
    global precision
    if p not in VALID:
        raise ValueError(f"precision must be one of {VALID}")
    precision = p


# 3-term ("h3") GEMMs on CTA pairs (tcgen05 cta_group::2, 256 x 128 tile per pair of SMs; csrc/gemm_tc.cu): "auto"
# (default: pairs except for the residual epilogue at K < 1024, measured with scripts/pair_diag.py), "1" always, "0" never.
# Same results bit for bit; applied to the library when it is loaded (_lib.lib()) and switchable with ops.set_gemm_pair().
gemm_pair = os.environ.get("VCR_GEMM_PAIR", "auto")
GEMM_PAIR_CODES = {"0": 0, "1": 1, "auto": 2}


# flash attention organisation: 3 (default) = Q and P held in tensor memory (only K and V tiles are read from shared memory
# by the tensor core), 8 softmax warps; 2 = every operand in shared memory, 8 warps; 4 = 16 warps; 1 = two groups of 4 warps
# alternating key tiles.  2 and 3 give the same bits.  Applied to the library when it is loaded, switchable with
# ops.set_flash_warps()
flash_warps = int(os.environ.get("VCR_FLASH_WARPS", "3"))


# vcrnetIter: serve repeated calls of one (network, shapes, iter) from a captured CUDA graph (vcr_net_b200/graph.py;
# bit-identical outputs, fresh result tensors).  ON by default for eval-mode VCRNet: the first call of a
# shape runs eagerly, the second captures, later ones replay -- one graph launch instead of ~150 Python -> ctypes -> launch
# round trips per call.  VCR_CUDA_GRAPH=0 keeps every call on the eager path (bench.py reports that as `variant_eager`).
cuda_graph = os.environ.get("VCR_CUDA_GRAPH", "1") == "1"

def graph_key():
    """Every switch that changes which kernels a captured vcrnetIter graph contains."""
    return (precision, hoist, flash_attention, fused_softcorr, fused_key_stat, str(knn_tc),
            str(gemm_pair), int(flash_warps))
