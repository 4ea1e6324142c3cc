"""VertexPredictor -- drop-in for the reference's models/VertexPredictor.py.

`point_pool_proj` is created lazily inside the first forward, exactly like the reference
(models/VertexPredictor.py:94-97): an optimizer built before the first forward does not
see it, and a fresh model loaded with strict=False drops the checkpoint's copy.  Set
WF_B200_EAGER_POOL_PROJ=1 to register it in __init__ instead (checkpoint value honoured)."""
import os

import torch
import torch.nn as nn

from wf_b200 import ops
from wf_b200._lib import ACT_RELU


class VertexPredictor(nn.Module):
    def __init__(self, global_feature_dim=512, max_vertices=64, vertex_dim=4):
        super(VertexPredictor, self).__init__()
        self.max_vertices = max_vertices
        self.vertex_dim = vertex_dim
        blk = lambda i, o: nn.Sequential(nn.Linear(i, o), nn.LayerNorm(o), nn.ReLU(inplace=True), nn.Dropout(0.0))
        self.vertex_mlp1 = blk(global_feature_dim, 4096)
        self.vertex_mlp2 = blk(4096, 2048)
        self.vertex_mlp3 = blk(2048, 2048)
        self.vertex_mlp4 = blk(2048, 1024)
        self.final_layer = nn.Linear(1024, max_vertices * vertex_dim)
        self.residual_proj1 = nn.Linear(global_feature_dim, 2048)
        self.residual_proj2 = nn.Linear(global_feature_dim, 1024)
        if os.environ.get("WF_B200_EAGER_POOL_PROJ", "0") == "1":
            self.point_pool_proj = nn.Linear(2 * global_feature_dim, global_feature_dim)

    def _proj(self, pooled):
        if not hasattr(self, "point_pool_proj"):
            self.point_pool_proj = nn.Linear(pooled.shape[1], self.residual_proj1.in_features)
            self.point_pool_proj = self.point_pool_proj.to(pooled.device)
        return self.point_pool_proj

    def forward_pooled(self, global_features, pooled_mean, pooled_max):
        """Same computation as forward(), fed with the already-reduced point features."""
        batch_size = global_features.shape[0]
        if pooled_mean is not None:
            pooled = torch.cat([pooled_mean, pooled_max], dim=1)
            pp = self._proj(pooled)
            eg = ops.linear_ln_act(pooled, pp.weight, pp.bias, residual=global_features)
        else:
            eg = global_features
        m1, m2, m3, m4 = self.vertex_mlp1, self.vertex_mlp2, self.vertex_mlp3, self.vertex_mlp4
        x = ops.linear_ln_act(eg, m1[0].weight, m1[0].bias, m1[1].weight, m1[1].bias, ACT_RELU)
        x = ops.linear_ln_act(x, m2[0].weight, m2[0].bias, m2[1].weight, m2[1].bias, ACT_RELU)
        r1 = ops.linear_ln_act(eg, self.residual_proj1.weight, self.residual_proj1.bias)
        x = ops.linear_ln_act(x, m3[0].weight, m3[0].bias, m3[1].weight, m3[1].bias, ACT_RELU, residual=r1)
        r2 = ops.linear_ln_act(eg, self.residual_proj2.weight, self.residual_proj2.bias)
        x = ops.linear_ln_act(x, m4[0].weight, m4[0].bias, m4[1].weight, m4[1].bias, ACT_RELU, residual=r2)
        vf = ops.linear_ln_act(x, self.final_layer.weight, self.final_layer.bias)
        if self.vertex_dim != 4:
            # the reference views (B, V, vertex_dim) and reads columns 0:3 and 3 (models/VertexPredictor.py:118-122): wider
            # rows carry unused columns, narrower ones fail on the existence column with torch's IndexError
            if self.vertex_dim < 4:
                raise IndexError(f"index 3 is out of bounds for dimension 2 with size {self.vertex_dim}")
            vf = vf.view(batch_size, self.max_vertices, self.vertex_dim)[:, :, :4].reshape(batch_size, self.max_vertices * 4)
        coords, prob, count = ops.VertexSplit.apply(vf, self.max_vertices)
        return {'vertices': coords, 'existence_probabilities': prob, 'actual_vertex_counts': count}

    def forward(self, global_features, point_features, target_vertex_counts=None):
        """Reference signature (models/VertexPredictor.py:63)."""
        if point_features is None:
            return self.forward_pooled(global_features, None, None)
        ops._need_cuda(point_features)
        B, N, _ = point_features.shape
        mask = torch.ones(B, N, device=point_features.device, dtype=torch.uint8)
        valid = torch.full((B,), float(N), device=point_features.device)
        _, _, max_u, mean_u, _, _ = ops.PoolPoints.apply(point_features, mask, valid)
        return self.forward_pooled(global_features, mean_u, max_u)
