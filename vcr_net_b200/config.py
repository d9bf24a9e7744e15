"""Run-time precision policy of the matrix engine.

  "fp32"  hand-written FP32 SIMT GEMMs everywhere (reference-exact arithmetic type)
  "h3"    tcgen05 tensor cores with 3-term fp16 split operands: fp32-level accuracy (the parity mode
          on tensor cores)
  "fp16" / "bf16"  single-pass tensor-core throughput modes (reported separately with their tolerance)
kNN / FPS / top-K selections and the SVD head always run in exact fp32 / fp64 arithmetic.
"""
import os

precision = os.environ.get("VCR_PRECISION", "h3")
VALID = ("fp32", "h3", "fp16", "bf16")
# flash attention kernel (attn_tc.cu) vs materialised scores through the GEMM kernel (tensor-core modes only)
flash_attention = os.environ.get("VCR_FLASH", "1") != "0"
# vcrnetIter: compute the loop-invariant target embedding emb_nn(tgt) once per call instead of once per --iter iteration
# (bit-identical outputs).  Off by default so the default path does exactly the work the reference does per iteration;
# bench.py reports the throughput with the switch on as a separate, labelled field.
reuse_target_embedding = os.environ.get("VCR_REUSE_TGT_EMB", "0") == "1"


def set_precision(p: str):
    global precision
    if p not in VALID:
        raise ValueError(f"precision must be one of {VALID}")
    precision = p
