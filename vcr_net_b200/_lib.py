"""ctypes binding of libvcr_b200.so.  Signatures are parsed from include/vcr_b200.h so the header
is the single source of truth for the C ABI.  There is NO fallback: if the library is missing or
a symbol cannot be bound this raises, and every op raises RuntimeError on a non-zero return code.
"""
from __future__ import annotations

import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libvcr_b200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "vcr_b200.h")

_ERR = {-1: "invalid argument or alignment", -2: "unsupported shape", -3: "CUDA launch failure",
        -4: "workspace too small"}

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
    "size_t": ctypes.c_size_t, "cudaStream_t": ctypes.c_void_p,
}


def parse_header(path: str = HEADER_PATH):
    """-> {name: (restype, [(ctype, argname), ...])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"\b(int|size_t|long long)\s+(vcr_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        parsed = []
        for a in [x.strip() for x in args.split(",") if x.strip()]:
            if a == "void":
                continue
            if "*" in a:
                parsed.append((ctypes.c_void_p, a.split("*")[-1].strip()))
            else:
                toks = a.split()
                ty = " ".join(toks[:-1]).replace("const ", "").strip()
                parsed.append((_CTYPES[ty], toks[-1]))
        protos[name] = (_CTYPES[ret], parsed)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m vcr_net_b200.build` "
                "(or __graft_entry__.build()).  vcr_net_b200 has no CPU / PyTorch fallback.")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self._prof = None
        for name, (ret, args) in self.protos.items():
            fn = getattr(self.cdll, name)          # AttributeError if the .so lacks the symbol
            fn.restype = ret
            fn.argtypes = [t for t, _ in args]
            takes_stream = bool(args) and args[-1][1] == "stream"
            setattr(self, name, self._wrap(name, fn) if takes_stream else fn)

    def _wrap(self, name, fn):
        """Optional per-call CUDA-event timing on the call's own stream (bench.py roofline leg)."""
        def call(*a):
            if self._prof is None:
                return fn(*a)
            import torch
            st = torch.cuda.ExternalStream(a[-1]) if a[-1] else torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = fn(*a)
            e1.record(st)
            self._prof.append((name, e0, e1, a))
            return rc
        return call

    def profile_begin(self):
        self._prof = []

    def profile_end(self):
        """-> list of (name, ms, args) for every stream-taking C-ABI call since profile_begin()."""
        import torch
        torch.cuda.synchronize()
        out = [(n, e0.elapsed_time(e1), a) for n, e0, e1, a in self._prof]
        self._prof = None
        return out

    def check(self, rc: int, name: str):
        if rc != 0:
            raise RuntimeError(f"{name} failed: {rc} ({_ERR.get(rc, 'unknown')})")


_instance = None


def lib() -> _Lib:
    global _instance
    if _instance is None:
        _instance = _Lib()
        from . import config
        _instance.vcr_set_gemm_pair(config.GEMM_PAIR_CODES[str(config.gemm_pair)])
        _instance.vcr_set_flash_warps(int(config.flash_warps))
    return _instance
