"""LPDNet embedding with the reference's module API (reference model/lpdnet_model.py:73-229).

Parameter containers are ordinary nn.Conv1d/Conv2d so the ``state_dict`` keys and shapes match the
reference's ``.t7`` layout (emb_nn.convDG1.0.weight [128,128,1,1] ...); the forward pass is the
CUDA pipeline of vcr_net_b200/functional.py.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import ops


class TranformNet(nn.Module):
    """model/lpdnet_model.py:19-70 (the reference's spelling).  forward(x [B,k,N]) -> [B,k,k]; eval mode only
    (BatchNorm1d running statistics folded into the convs / linears)."""

    def __init__(self, k=3, negative_slope=1e-2):
        super().__init__()
        self.negative_slope = negative_slope
        self.conv1 = nn.Conv1d(k, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k * k)
        self.relu = nn.LeakyReLU(negative_slope=self.negative_slope)      # never called by the reference's forward
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.bn4 = nn.BatchNorm1d(512)
        self.bn5 = nn.BatchNorm1d(256)
        self.k = k

    @torch.no_grad()
    def forward(self, x):
        if self.k == 3:
            return Fn.transform_net_matrix(self, x)
        return Fn.transform_net_matrix(self, ops.transpose_batched(x.contiguous()))


class LPDNet(nn.Module):
    """model/lpdnet_model.py:73-137.  forward(x [B,3,N]) -> [B,emb_dims,N]."""

    def __init__(self, args, negative_slope=0.0):
        super().__init__()
        self.negative_slope = negative_slope
        self.k = 20
        self.t3d = args.t3d
        self.tfea = args.tfea
        self.emb_dims = args.emb_dims
        act = lambda: nn.LeakyReLU(negative_slope=self.negative_slope)
        self.convDG1 = nn.Sequential(nn.Conv2d(64 * 2, 128, kernel_size=1, bias=True), act())
        self.convDG2 = nn.Sequential(nn.Conv2d(128, 128, kernel_size=1, bias=True), act())
        self.convSN1 = nn.Sequential(nn.Conv2d(128 * 2, 256, kernel_size=1, bias=True), act())
        self.conv1_lpd = nn.Conv1d(3, 64, kernel_size=1, bias=True)
        self.conv2_lpd = nn.Conv1d(64, 64, kernel_size=1, bias=True)
        self.conv3_lpd = nn.Conv1d(512, self.emb_dims, kernel_size=1, bias=True)
        if self.t3d:
            self.t_net3d = TranformNet(3)
        if self.tfea:
            self.t_net_fea = TranformNet(64)

    def forward_tokens(self, x, idx_feat=None, idx_xyz=None, stages=None):
        """x [B,3,N] -> tokens [B,N,emb_dims] (internal fast path, no output transpose)."""
        # train-mode test on the conv attributes, not parameters(): replicas of nn.DataParallel have no parameters
        if stages is None and torch.is_grad_enabled() and any(t.requires_grad for t in Fn.lpdnet_param_tensors(self)):
            if self.t3d or self.tfea:
                raise Exception("Not implemented: backward through TranformNet (t3d/tfea); run under torch.no_grad()")
            return Fn.lpdnet_tokens_train(self, x, idx_feat=idx_feat, idx_xyz=idx_xyz)     # training: custom backward
        return Fn.lpdnet_tokens(self, x, idx_feat=idx_feat, idx_xyz=idx_xyz, stages=stages)

    def forward(self, x, idx_feat=None, idx_xyz=None):
        """idx_feat / idx_xyz (int32 [B,N,20], optional) inject neighbour sets like get_graph_feature(x, idx=...)."""
        tok = self.forward_tokens(x, idx_feat=idx_feat, idx_xyz=idx_xyz)
        if tok.requires_grad:
            return tok.transpose(1, 2).contiguous()   # autograd-visible; values identical to the kernel transpose
        return ops.transpose_batched(tok)


class LPD(nn.Module):
    """model/lpdnet_model.py:140-229 (pre-training wrapper), forward only: embeddings by the CUDA
    path, FPS anchors by the CUDA FPS kernel, the O(B*32*512) triplet arithmetic in torch."""

    def __init__(self, args):
        super().__init__()
        self.emb_dims = args.emb_dims
        self.num_points = args.num_points
        self.negative_slope = 0.2
        self.emb_nn = LPDNet(args, negative_slope=self.negative_slope)
        self.cycle = args.cycle

    def forward(self, *input):
        src, tgt = input[0], input[1]
        batch_size = src.size(0)
        both = self.emb_nn(torch.cat([src, tgt], dim=0))
        src_embedding, tgt_embedding = both[:batch_size], both[batch_size:]
        loss = self.getLoss(src, src_embedding, tgt_embedding)
        mse_ab_ = torch.mean((src_embedding - tgt_embedding) ** 2, dim=[0, 1, 2]) * batch_size
        mae_ab_ = torch.mean(torch.abs(src_embedding - tgt_embedding), dim=[0, 1, 2]) * batch_size
        return src_embedding, tgt_embedding, loss, mse_ab_, mae_ab_

    def kfn(self, x, k=20):
        """:163-171: k FARTHEST among the anchors (top-k of +squared distance), [B,3,32] sized."""
        inner = -2 * torch.matmul(x.transpose(2, 1).contiguous(), x)
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        pd = xx + inner
        pd = pd + xx.transpose(2, 1).contiguous()
        return pd.topk(k=k, dim=-1)[1]

    def triplet_loss(self, src_embedding_k, tgt_embedding_k, topFarTgt):
        margin = 1.0
        s = src_embedding_k.unsqueeze(3)
        dp_loss = torch.mean((s - tgt_embedding_k) ** 2, dim=[1, 3])
        dn_loss = torch.mean((s - topFarTgt) ** 2, dim=[1, 3])
        return torch.clamp_min(1 - dn_loss / (margin + dp_loss), 0.0)

    def getLoss(self, src, src_embedding, tgt_embedding, k=32, neg_k=8):
        B, pt_dims, N = src.size()
        emb_dims = src_embedding.size(1)
        from ..util.util import farthest_point_sample
        sampleIdx = farthest_point_sample(src, npoint=k)                        # CUDA FPS kernel
        src_k = torch.gather(src, 2, sampleIdx.unsqueeze(1).expand(-1, pt_dims, -1))
        eidx = sampleIdx.unsqueeze(1).expand(-1, emb_dims, -1)
        se_k = torch.gather(src_embedding, 2, eidx)
        te_k = torch.gather(tgt_embedding, 2, eidx)
        far = self.kfn(src_k, k=neg_k)                                          # [B,k,neg_k]
        tek_t = te_k.transpose(2, 1)                                            # [B,k,D]
        neg = torch.gather(tek_t.unsqueeze(1).expand(-1, k, -1, -1), 2,
                           far.unsqueeze(-1).expand(-1, -1, -1, emb_dims))      # [B,k,neg_k,D]
        topFarTgt = neg.permute(0, 3, 1, 2)
        loss_triplet = self.triplet_loss(se_k, te_k.unsqueeze(3), topFarTgt)
        src_length = torch.norm(src_embedding.transpose(2, 1), dim=-1)
        tgt_length = torch.norm(tgt_embedding.transpose(2, 1), dim=-1)
        one = torch.ones_like(src_length)
        loss_norm1 = torch.sqrt(torch.nn.functional.mse_loss(src_length, one))
        loss_norm2 = torch.sqrt(torch.nn.functional.mse_loss(tgt_length, one))
        return loss_triplet.mean() + (loss_norm1 + loss_norm2) / 2.0 * 0.03
