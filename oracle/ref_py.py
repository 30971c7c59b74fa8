"""Loads the reference's UNMODIFIED Python (staged by build_ref.stage_python into oracle/_ref/py/) as a package tree.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and by bench.py's `reference_cuda` extra), never by de6d_b200.

    tree = load_tree("pcdet", extensions=None)        # `from . import pointnet2_batch_cuda` resolves through
                                                      # sys.modules, i.e. to whatever de6d_b200.compat.install() put there
    tree = load_tree("pcdet_ref", extensions=build_ref.load())   # the same files over the reference's own kernels

`tree.pointnet2_utils`, `.pointnet2_modules`, `.iou3d_nms_utils`, `.roiaware_pool3d_utils`, `.model_nms_utils`,
`.box_utils`, `.pointnet2_backbone` are the reference modules.  The package objects are created by hand (types.ModuleType with __path__) so no
reference __init__.py runs: pcdet/__init__.py needs a generated version.py the mount does not have.
Third-party imports the mount lacks are stubbed: SharedArray (pcdet/utils/common_utils.py:7) and the out-of-scope
pointnet2_stack_cuda extension (pointnet2_stack/pointnet2_utils.py:8, imported by pointnet2_modules.py:7 only for the
experimental samplers).
"""
import importlib
import os
import sys
import tempfile
import types
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
PY_ROOT = os.path.join(HERE, "_ref", "py", "pcdet")

_PACKAGES = ["", "ops", "ops.pointnet2", "ops.pointnet2.pointnet2_batch", "ops.pointnet2.pointnet2_stack",
             "ops.iou3d_nms", "ops.roiaware_pool3d", "models", "models.model_utils", "models.backbones_3d", "utils"]
_EXT_PATHS = {
    "pointnet2_batch_cuda": "ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda",
    "iou3d_nms_cuda": "ops.iou3d_nms.iou3d_nms_cuda",
    "roiaware_pool3d_cuda": "ops.roiaware_pool3d.roiaware_pool3d_cuda",
}


def available():
    return os.path.exists(os.path.join(PY_ROOT, "ops", "pointnet2", "pointnet2_batch", "pointnet2_modules.py"))


def load_tree(root="pcdet", extensions=None):
    """Build the package tree `root` over oracle/_ref/py/pcdet and import the op-level reference modules.
    extensions: {"pointnet2_batch_cuda": module, ...} registered under the tree; None = keep what sys.modules
    already holds under `root` (de6d_b200.compat.install() for root == "pcdet")."""
    if not available():
        raise FileNotFoundError("oracle/_ref/py not staged: run `python oracle/build_ref.py` where /root/reference is mounted")
    if "SharedArray" not in sys.modules:
        try:
            importlib.import_module("SharedArray")
        except ImportError:
            sys.modules["SharedArray"] = types.ModuleType("SharedArray")
    # pointnet2_stack/pointnet2_utils.py jits helpers with numba cache=True; the cache index is keyed by file path but the
    # pickles name the importing module, so a tree loaded under a second root must not see the first root's cache
    cache_dir = os.path.join(tempfile.gettempdir(), "de6d_numba_cache_%s_%d" % (root, os.getuid()))
    os.environ["NUMBA_CACHE_DIR"] = cache_dir
    try:
        import numba
        numba.config.CACHE_DIR = cache_dir
    except Exception:
        pass
    for sub in _PACKAGES:
        name = root + ("." + sub if sub else "")
        if name not in sys.modules or not hasattr(sys.modules[name], "__path__"):
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(PY_ROOT, *sub.split("."))] if sub else [PY_ROOT]
            pkg.__package__ = name
            sys.modules[name] = pkg
            if sub:
                parent, leaf = name.rsplit(".", 1)
                setattr(sys.modules[parent], leaf, pkg)
    stack_ext = root + ".ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda"
    sys.modules.setdefault(stack_ext, types.ModuleType(stack_ext))
    for short, sub in _EXT_PATHS.items():
        name = root + "." + sub
        if extensions is not None:
            sys.modules[name] = extensions[short]
        elif name not in sys.modules:
            raise RuntimeError("%s is not registered: call de6d_b200.compat.install() first or pass extensions=" % name)
        setattr(sys.modules[name.rsplit(".", 1)[0]], short, sys.modules[name])
    imp = importlib.import_module
    return SimpleNamespace(
        root=root,
        pointnet2_utils=imp(root + ".ops.pointnet2.pointnet2_batch.pointnet2_utils"),
        pointnet2_modules=imp(root + ".ops.pointnet2.pointnet2_batch.pointnet2_modules"),
        iou3d_nms_utils=imp(root + ".ops.iou3d_nms.iou3d_nms_utils"),
        roiaware_pool3d_utils=imp(root + ".ops.roiaware_pool3d.roiaware_pool3d_utils"),
        model_nms_utils=imp(root + ".models.model_utils.model_nms_utils"),
        box_utils=imp(root + ".utils.box_utils"),
        pointnet2_backbone=imp(root + ".models.backbones_3d.pointnet2_backbone"),
    )


def load_pair():
    """(reference python over de6d_b200.compat, the same files over the reference's own extension modules)."""
    from de6d_b200 import compat
    from . import build_ref
    compat.install()
    ours = load_tree("pcdet", None)
    theirs = load_tree("pcdet_ref", build_ref.load())
    return ours, theirs
