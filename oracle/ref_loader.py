"""Loads the LIVE reference hot path (vector_quantizer/vq_img.py) by file path -- TEST INFRASTRUCTURE.

Only usable where /root/reference exists (the authoring container).  It is used to pin the
restated oracle (tests/test_oracle.py) and to generate tests/golden/*.pt
(tests/golden/make_golden.py).  Nothing that runs on the GPU box may depend on it.
`import vector_quantizer` (the package) fails here because easydict is not installed, but
vq_img.py itself needs only torch + einops.
"""
import importlib.util
import os

_CANDIDATES = [os.environ.get("VQSEG_REF", ""), "/root/reference"]


def reference_root():
    for root in _CANDIDATES:
        if root and os.path.isfile(os.path.join(root, "vector_quantizer", "vq_img.py")):
            return root
    return None


def load_reference_vq_img():
    root = reference_root()
    if root is None:
        return None
    path = os.path.join(root, "vector_quantizer", "vq_img.py")
    spec = importlib.util.spec_from_file_location("ref_vq_img", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_seghead():
    """models/modules/vq_segmentation_head.py of the live reference (needs only torch + einops)."""
    root = reference_root()
    if root is None:
        return None
    path = os.path.join(root, "models", "modules", "vq_segmentation_head.py")
    if not os.path.isfile(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_vq_segmentation_head", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
