"""wf_b200 -- host-side binding of libwf_b200.so, the B200-native hot path of
cansdev/wireframe-3d-prediction (PointCloudToWireframe forward/backward, WireframeLoss, matchers).

`import wf_b200` works without a GPU (symbol binding only); any compute call needs a CUDA device
and raises WfError otherwise.  There is no CPU or PyTorch-eager fallback."""
from ._lib import LIB_PATH, WfError, load  # noqa: F401
from .ops import get_precision, set_precision  # noqa: F401

__all__ = ["LIB_PATH", "WfError", "load", "get_precision", "set_precision"]
