"""Point-cloud ops with the reference's names and signatures (reference util/util.py).

knn / get_graph_feature / farthest_point_sample / transform_point_cloud run as CUDA kernels
(vcr_net_b200/csrc); quat2mat and npmat2euler are a handful of scalar ops per pair on the host
side of the boundary and stay torch / scipy.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import config, ops


def knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """util/util.py:143-160.  x [B,D,N] -> int64 [B,N,k]; ties -> lower index (canonical order)."""
    B, D, N = x.shape
    if config.use_knn_tc(N) and ops.knn_tc_supported(D, k) and N >= k + 1:
        # feature-space kNN: tcgen05 prefilter + exact re-rank (bit-identical to the FP32 SIMT kernel)
        _, idx64 = ops.knn_topk_tc(ops.transpose_batched(x.contiguous()), None, k, want64=True)
        return idx64
    _, idx64 = ops.knn_topk(x, k, token_major=False, want64=True)
    return idx64


def get_graph_feature(x: torch.Tensor, k: int = 20, idx: torch.Tensor | None = None) -> torch.Tensor:
    """util/util.py:176-199.  x [B,D,N] -> [B,2D,N,k] = concat(neighbour features, centre)."""
    B, D, N = x.shape
    x = x.reshape(B, -1, N)
    if idx is None:
        idx32 = ops.knn_topk(x, k, token_major=False)
    else:
        idx32 = idx.to(torch.int32).contiguous()
    xt = ops.transpose_batched(x)                     # [B,N,D]
    return ops.graph_feature(xt, idx32)


def farthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """util/util.py:107-140.  xyz [B,3,N] -> int64 [B,npoint]."""
    return ops.fps(xyz, npoint, want64=True)


def quat2mat(quat: torch.Tensor) -> torch.Tensor:
    """util/util.py:76-88.  quat [B,4] as (x,y,z,w) -> [B,3,3]."""
    x, y, z, w = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    B = quat.size(0)
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).reshape(B, 3, 3)


def transform_point_cloud(point_cloud: torch.Tensor, rotation: torch.Tensor, translation: torch.Tensor):
    """util/util.py:91-96.  point_cloud [B,3,N]; rotation [B,3,3] or quaternion [B,4]; translation [B,3]."""
    rot = quat2mat(rotation) if rotation.dim() == 2 else rotation
    return ops.rigid_apply(point_cloud, rot, translation)


def npmat2euler(mats, seq="zyx"):
    """util/util.py:99-104 (scipy renamed from_dcm to from_matrix)."""
    from scipy.spatial.transform import Rotation
    return np.asarray([Rotation.from_matrix(m).as_euler(seq, degrees=True) for m in mats], dtype="float32")
