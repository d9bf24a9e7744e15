"""ICP refinement with the reference's module API (reference model/icp_model.py:15-108) on the CUDA path.

The loop runs without a single host synchronisation: the convergence test of icp_model.py:41-43 is evaluated on the
device (csrc/icp.cu) and freezes the source cloud for the remaining iterations, so the result equals the reference's
early ``break`` while every launch is enqueued up front."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class ICP(nn.Module):
    """model/icp_model.py:16-50.  forward(srcInit, dst [B,3,N]) -> (srcInit, src, R_ab, t_ab, R_ba, t_ba)."""

    def __init__(self, max_iterations=10, tolerance=0.001):
        super().__init__()
        self.max_iterations = max_iterations
        self.tolerance = tolerance
        self.reflect = nn.Parameter(torch.eye(3), requires_grad=False)      # state_dict parity (:22-23)
        self.reflect[2, 2] = -1
        self.last_state = None

    @torch.no_grad()
    def forward(self, srcInit, dst):
        srcInit = srcInit.contiguous()
        dst = dst.contiguous()
        src = srcInit.clone()
        state = ops.icp_state(src.device)
        errs = torch.zeros(max(1, self.max_iterations), dtype=torch.float64, device=src.device)
        for i in range(self.max_iterations):
            corr = ops.icp_nearest(src, dst, errs[i:i + 1])                 # nearest_neighbor (:33)
            R, t, _, _ = ops.svd_head(src, corr)                            # best_fit_transform (:35)
            ops.icp_advance_(src, R, t, errs[i:i + 1], self.tolerance, state)   # :36-43, break -> device flag
        R_ab, t_ab, R_ba, t_ba = ops.svd_head(srcInit, src)                 # :46-49
        self.last_state = state      # [done, iters | prev_error bits]: diagnostics only, reading it synchronises
        return srcInit, src, R_ab, t_ab, R_ba, t_ba

    def iterations_run(self):
        """Number of loop iterations the reference would have executed (reads the device state: host sync)."""
        return int((self.last_state[0].item() >> 32) & 0xffffffff) if self.last_state is not None else 0
