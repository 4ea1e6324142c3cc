"""CPU-side checks of the boundary: the C-ABI library loads without a GPU and exports every symbol
include/wf_b200.h declares, with the argument counts the ctypes binding assumes; the drop-in modules keep
the reference's state_dict keys, constructor defaults and loud no-GPU failure."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_protos():
    txt = open(os.path.join(ROOT, "include", "wf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(wf_\w+)\s*\(([^;]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(1)] = n
    return protos


def test_library_exports_header_symbols():
    from wf_b200 import _lib
    lib = _lib.load()
    protos = _header_protos()
    assert len(protos) >= 30
    for name, nargs in protos.items():
        assert hasattr(lib, name), f"{name} declared in include/wf_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib.SIGNATURES[name]) == nargs, f"{name}: header has {nargs} args, binding {len(_lib.SIGNATURES[name])}"
    for name in _lib.SIGNATURES:
        assert name in protos, f"binding declares {name} which the header does not"
    assert lib.wf_version() >= 100


def test_every_entry_point_the_package_calls_is_bound():
    """Every `call("wf_...")` in the product package names an entry point of the binding (a typo would otherwise only surface
    on a GPU box), with as many arguments as its ctypes signature for the calls whose argument list is on one level."""
    import glob
    from wf_b200 import _lib
    used = {}
    for f in glob.glob(os.path.join(ROOT, "wireframe-3d-prediction_b200", "**", "*.py"), recursive=True):
        for m in re.finditer(r'call\(\s*"(wf_[a-z0-9_]+)"', open(f).read()):
            used.setdefault(m.group(1), []).append(os.path.relpath(f, ROOT))
    assert len(used) >= 40
    for name, where in used.items():
        assert name in _lib.SIGNATURES, f"{where[0]} calls {name}, which the binding does not declare"


def test_state_dict_keys_match_reference_inventory():
    from oracle import wireframe_oracle as wo
    from models.PointCloudToWireframe import PointCloudToWireframe
    m = PointCloudToWireframe(input_dim=8, max_vertices=64)
    keys = list(m.state_dict().keys())
    expect = [k for k, _, _ in wo.param_shapes(64) if "point_pool_proj" not in k]
    assert keys == expect                                   # names AND registration order
    assert sum(p.numel() for p in m.parameters()) == 30528897
    shapes = {k: tuple(s) for k, s, _ in wo.param_shapes(64)}
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == shapes[k], k
    assert m.max_vertices == 64
    # strict=False load of a checkpoint that carries the lazy layer drops it, like the reference (Q2)
    sd = wo.make_state_dict(0, 64)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert unexpected == ["vertex_predictor.point_pool_proj.weight", "vertex_predictor.point_pool_proj.bias"]


def test_no_cpu_fallback():
    from wf_b200._lib import WfError
    from models.PointCloudToWireframe import PointCloudToWireframe
    from losses.WireframeLoss import WireframeLoss
    m = PointCloudToWireframe(max_vertices=4)
    with pytest.raises(WfError):
        m(torch.rand(1, 8, 8))
    crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
    assert (crit.vertex_weight, crit.edge_weight, crit.existence_weight) == (3.0, 1.0, 1.5)
    with pytest.raises(WfError):
        crit({"vertices": torch.zeros(1, 4, 3), "existence_probabilities": torch.full((1, 4), 0.5), "edge_probs": torch.zeros(1, 0)},
             {"vertices": torch.zeros(1, 4, 3), "vertex_existence": torch.zeros(1, 4), "edge_labels": torch.zeros(1, 0),
              "vertex_counts": torch.tensor([2])})


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "wireframe-3d-prediction_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} imports the oracle"


def test_helpers_importable_like_train_py():
    from models.utils import create_edge_labels_from_edge_set
    lab = create_edge_labels_from_edge_set({(0, 1), (1, 2)}, [(0, 1), (0, 2), (1, 2)])
    assert lab.tolist() == [[1.0, 0.0, 1.0]]
    from models.WireframeHungarianMatcher import build_wireframe_matcher
    from models.HungarianMatcher import build_matcher, box_iou, generalized_box_iou, box_cxcywh_to_xyxy  # noqa: F401
    assert build_wireframe_matcher(2.0, 0.5).cost_vertex == 2.0
