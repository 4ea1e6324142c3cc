"""EdgePredictor -- drop-in for the reference's models/EdgePredictor.py (same parameters, incl. the
unused `spatial_proj`).  All samples of a batch run as ONE ragged launch sequence; the
(E,1031) pair-feature concat is never materialised (first edge layer is applied in decomposed form)."""
import torch
import torch.nn as nn

from wf_b200 import ops
from wf_b200._lib import ACT_GELU, ACT_NONE

_PAIR_LISTS = {}


def pair_list(c):
    """[[i, j] for i < j] as the reference's `edge_indices.tolist()` (models/EdgePredictor.py:84-89,140);
    cached per vertex count (the reference rebuilds it with an O(c^2) Python loop every call)."""
    lst = _PAIR_LISTS.get(c)
    if lst is None:
        lst = [[i, j] for i in range(c) for j in range(i + 1, c)]
        _PAIR_LISTS[c] = lst
    return lst


class EdgePredictor(nn.Module):
    def __init__(self, vertex_dim=3, hidden_dim=512, num_heads=8):
        super(EdgePredictor, self).__init__()
        self.vertex_proj = nn.Sequential(
            nn.Linear(vertex_dim, hidden_dim // 2), nn.LayerNorm(hidden_dim // 2), nn.GELU(),
            nn.Linear(hidden_dim // 2, hidden_dim), nn.LayerNorm(hidden_dim), nn.Dropout(0.1))
        self.attention = nn.MultiheadAttention(embed_dim=hidden_dim, num_heads=num_heads, dropout=0.1, batch_first=True)
        self.spatial_proj = nn.Sequential(nn.Linear(vertex_dim, hidden_dim // 4), nn.GELU(),
                                          nn.Linear(hidden_dim // 4, hidden_dim // 4))
        self.edge_mlp = nn.Sequential(
            nn.Linear(hidden_dim * 2 + vertex_dim * 2 + 1, hidden_dim), nn.LayerNorm(hidden_dim), nn.GELU(), nn.Dropout(0.1),
            nn.Linear(hidden_dim, hidden_dim // 2), nn.LayerNorm(hidden_dim // 2), nn.GELU(), nn.Dropout(0.1),
            nn.Linear(hidden_dim // 2, hidden_dim // 4), nn.GELU(),
            nn.Linear(hidden_dim // 4, 1))
        self._dims = (vertex_dim, hidden_dim, num_heads)
        # nn.MultiheadAttention above has already raised the reference's AssertionError when num_heads does not divide
        # hidden_dim.  The kernels are instantiated for hidden_dim in {128, 256, 512, 1024} (the pair layer keeps hidden_dim / 32
        # channels per lane, the output layer hidden_dim / 128) and head widths 16 / 32 / 64 / 128; vertices are 3-D
        # (PointCloudToWireframe.py:41 is the only constructor call in the reference: vertex_dim=3, defaults otherwise).
        if vertex_dim != 3 or hidden_dim not in (128, 256, 512, 1024) or hidden_dim // num_heads not in (16, 32, 64, 128):
            raise ops._lib.WfError(f"EdgePredictor(vertex_dim={vertex_dim}, hidden_dim={hidden_dim}, num_heads={num_heads}): kernels "
                                   "exist for vertex_dim=3, hidden_dim in {128,256,512,1024}, hidden_dim/num_heads in {16,32,64,128}")

    def _get_edge_indices(self, num_vertices):
        return torch.tensor(pair_list(num_vertices), dtype=torch.long, device=next(self.parameters()).device)

    def _keep(self, shape, drop: nn.Dropout, device):
        return ops.dropout_keep(shape, drop.p, drop.training, device)

    def forward_ragged(self, verts, rg):
        """verts: packed (T,3) vertices of all samples; rg: ops.Ragged.  Returns probs (B, max_e), zero padded."""
        H = self._dims[1]
        dev = verts.device
        vp, att, em = self.vertex_proj, self.attention, self.edge_mlp
        f = ops.linear_ln_act(verts, vp[0].weight, vp[0].bias, vp[1].weight, vp[1].bias, ACT_GELU)
        k, ks = self._keep((rg.T, H), vp[5], dev)
        f = ops.linear_ln_act(f, vp[3].weight, vp[3].bias, vp[4].weight, vp[4].bias, ACT_NONE, None, k, ks)
        qkv = ops.linear_ln_act(f, att.in_proj_weight, att.in_proj_bias)
        p_att = att.dropout if self.training else 0.0
        ka, kas = ops.dropout_keep((rg.Ptot,), p_att, p_att > 0.0, dev)
        o = ops.AttentionCore.apply(qkv, rg, ka, kas, self._dims[2])
        f = ops.linear_ln_act(o, att.out_proj.weight, att.out_proj.bias, residual=f)       # reference :114
        # first edge layer on [f_i | f_j | v_i | v_j | dist] = P[i] + Q[j] + w*dist: both feature blocks in ONE product against the
        # stacked, contiguous (2H, H) weight (the 2H + 7-wide rows of edge_mlp.0.weight are not 16-byte aligned for TMA); the
        # pair kernels read P and Q as the two halves of that product
        Wf, Wv, wd = ops.SplitPairWeight.apply(em[0].weight)                                # (H, 2H + 7): (512, 1031)
        PQ = ops.linear_ln_act(verts, Wv, residual=ops.linear_ln_act(f, Wf))               # (T, 2H)
        z1 = ops.EdgePairLayer.apply(PQ, verts, wd, em[0].bias, rg)
        k, ks = self._keep((rg.E, H), em[3], dev)
        e = ops.LNAct.apply(z1, em[1].weight, em[1].bias, ACT_GELU, k, ks)
        k, ks = self._keep((rg.E, H // 2), em[7], dev)
        e = ops.linear_ln_act(e, em[4].weight, em[4].bias, em[5].weight, em[5].bias, ACT_GELU, None, k, ks)
        e = ops.linear_ln_act(e, em[8].weight, em[8].bias, None, None, ACT_GELU)
        return ops.EdgeOut.apply(e, em[10].weight, em[10].bias, rg)

    def forward(self, vertices):
        """Reference signature: vertices (batch, num_vertices, 3) -> (edge_probs (batch, E), [[i,j],...])."""
        batch_size, num_vertices, vertex_dim = vertices.shape
        if num_vertices <= 1:
            # the reference indexes a 1-D empty index tensor here (models/EdgePredictor.py:117-119)
            raise IndexError("too many indices for tensor of dimension 1")
        ops._need_cuda(vertices)
        rg = ops.Ragged([num_vertices] * batch_size, vertices.device)
        probs = self.forward_ragged(vertices.reshape(batch_size * num_vertices, vertex_dim), rg)
        return probs, pair_list(num_vertices)
