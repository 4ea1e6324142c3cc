"""PointNetEncoder -- drop-in for the reference's models/PointNetEncoder.py (same ctor, forward
signature, parameter names: encoder.mlp.{0,1,4,5,8,9,12,13,16}.*, encoder.feature_fusion.{0,1,3,4,6}.*).

The nn.Sequential containers below exist to hold and initialise the parameters exactly as the
reference does (same registration order, same default init, so the same torch seed gives the same
weights); they are never called.  All compute goes through wf_b200.ops -> libwf_b200.so."""
import torch
import torch.nn as nn

from wf_b200 import ops
from wf_b200._lib import ACT_RELU


class PointNetEncoder(nn.Module):
    def __init__(self, input_dim=8, hidden_dims=[512, 1024, 2048, 1024], output_dim=512):
        super(PointNetEncoder, self).__init__()
        layers = []
        prev_dim = input_dim
        for hidden_dim in hidden_dims:
            layers.extend([nn.Linear(prev_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(inplace=True), nn.Dropout(0.0)])
            prev_dim = hidden_dim
        layers.append(nn.Linear(prev_dim, output_dim))
        self.mlp = nn.Sequential(*layers)
        self.global_max_pool = nn.AdaptiveMaxPool1d(1)      # registered but unused, as in the reference (:52-53)
        self.global_avg_pool = nn.AdaptiveAvgPool1d(1)
        self.feature_fusion = nn.Sequential(
            nn.Linear(output_dim * 2, output_dim * 4), nn.LayerNorm(output_dim * 4), nn.ReLU(inplace=True),
            nn.Linear(output_dim * 4, output_dim * 2), nn.LayerNorm(output_dim * 2), nn.ReLU(inplace=True),
            nn.Linear(output_dim * 2, output_dim))
        self._n_hidden = len(hidden_dims)
        # the tensor-core path is built for the shipped widths; fewer input features (the dataset's use_color / use_intensity
        # switches give 3, 4 or 7, datasets/building3d.py:103-111) ride on it zero-padded to 8
        self._tc_shape = (1 <= input_dim <= 8 and list(hidden_dims) == [512, 1024, 2048, 1024] and output_dim == 512)
        self._input_dim = input_dim

    def tc_inputs(self, x):
        """(x, parameter list) as the tensor-core encoder kernels take them: 8 input features."""
        p = []
        for li in range(self._n_hidden):
            lin, ln = self.mlp[4 * li], self.mlp[4 * li + 1]
            p += [lin.weight, lin.bias, ln.weight, ln.bias]
        last = self.mlp[4 * self._n_hidden]
        p += [last.weight, last.bias]
        if x.shape[2] < 8:
            # zero features and zero weight columns leave every activation unchanged (the centred layer's factorisation
            # drops the rank-deficient directions); the pad's autograd slices the weight gradient back
            x = torch.nn.functional.pad(x, (0, 8 - x.shape[2]))
            p[0] = torch.nn.functional.pad(p[0], (0, 8 - p[0].shape[1]))
        return x, p

    # ---- per-point MLP + the four pooled reductions (reference :85-111 and VertexPredictor :86-87)
    def pooled(self, x, want_point_features=False):
        """Returns (max_masked, avg_masked, max_unmasked, mean_unmasked, argmax_masked, argmax_unmasked,
        point_features or None)."""
        if x.dim() != 3:
            raise ValueError("expected point cloud of shape (batch, num_points, input_dim)")
        if ops.get_precision() == "bf16" and self._tc_shape:
            x, p = self.tc_inputs(x)
            if (not torch.is_grad_enabled() and not want_point_features and ops.FUSED_POOL
                    and x.shape[1] >= ops._FUSED_MIN_POINTS):
                # inference: chunked over points, nothing saved (evaluate.py:71 runs under torch.no_grad())
                return (*ops.encoder_pooled_infer(x, p), None)
            r = ops.EncoderPointMLP_TC.apply(x, bool(want_point_features), *p)
            return (*r[:6], r[6] if want_point_features else None)
        B, N, D = x.shape
        ops._need_cuda(x)
        xf = ops._f32c(x)
        mask, valid = ops.point_mask(xf.detach())
        h = xf.reshape(B * N, D)
        for li in range(self._n_hidden):
            lin, ln = self.mlp[4 * li], self.mlp[4 * li + 1]
            h = ops.linear_ln_act(h, lin.weight, lin.bias, ln.weight, ln.bias, ACT_RELU)
        last = self.mlp[4 * self._n_hidden]
        pf = ops.linear_ln_act(h, last.weight, last.bias).reshape(B, N, -1)
        r = ops.PoolPoints.apply(pf, mask, valid)
        return (*r, pf if want_point_features else None)

    def fuse(self, max_features, avg_features):
        """reference :115-116"""
        ff = self.feature_fusion
        g = torch.cat([max_features, avg_features], dim=1)
        g = ops.linear_ln_act(g, ff[0].weight, ff[0].bias, ff[1].weight, ff[1].bias, ACT_RELU)
        g = ops.linear_ln_act(g, ff[3].weight, ff[3].bias, ff[4].weight, ff[4].bias, ACT_RELU)
        return ops.linear_ln_act(g, ff[6].weight, ff[6].bias)

    def forward(self, x):
        """Reference signature: returns (global_features (B,512), point_features (B,N,512))."""
        max_m, avg_m, _, _, _, _, pf = self.pooled(x, want_point_features=True)
        return self.fuse(max_m, avg_m), pf
