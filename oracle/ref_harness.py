"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Imports and runs the UNMODIFIED reference model graph (/root/reference/models/HRegNet/*.py,
models/model_v2/*.py, models/utils.py) on the CPU, with ONLY the native ops underneath replaced by the
oracle's CPU restatements (oracle/native.py):

  * `point_utils_cuda`  (reference pybind module, point_utils_api.cpp:6-13)  -> oracle C FPS / gather
  * `pytorch3d.ops.knn_points / knn_gather` (third-party, not vendored)       -> oracle C kNN / index gather

The reference hard-codes `.cuda()` (models.py:100,104,120; layers.py:305,311,354,360,488-490,497) and the
legacy `torch.cuda.IntTensor/FloatTensor` constructors (models/utils.py:24-25); those four entry points are
redirected to the CPU while the harness is active.  Every layer above the native ops is the reference's own
Python, which is why its outputs are usable as golden vectors (tests/golden/make_golden.py).

/root/reference exists only in the build container: nothing that runs on the GPU box imports this module.
"""
import contextlib
import os
import sys
import types

import torch

from . import native

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference():
    """/root/reference in the build container; on the GPU box the staged copy baseline/_ref (oracle/build_ref.py:
    stage_reference, git-ignored, unmodified files)."""
    for cand in (os.environ.get("HREGNET_REFERENCE"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "models", "HRegNet", "layers.py")):
            return cand
    return "/root/reference"


REF = _find_reference()


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "HRegNet", "layers.py"))


def _fake_point_utils():
    m = types.ModuleType("point_utils_cuda")

    def furthest_point_sampling_wrapper(b, n, npoint, xyz, temp, out):
        out.copy_(native.fps(xyz, npoint, None, temp))
        return 1

    def weighted_furthest_point_sampling_wrapper(b, n, npoint, xyz, w, temp, out):
        out.copy_(native.fps(xyz, npoint, w, temp))
        return 1

    def gather_points_wrapper(b, c, n, npoint, pts, idx, out):
        out.copy_(native.gather_points(pts, idx))
        return 1

    def gather_points_grad_wrapper(b, c, n, npoint, grad_out, idx, grad_points):
        grad_points.add_(native.gather_points_grad(grad_out, idx, n))
        return 1

    m.furthest_point_sampling_wrapper = furthest_point_sampling_wrapper
    m.weighted_furthest_point_sampling_wrapper = weighted_furthest_point_sampling_wrapper
    m.gather_points_wrapper = gather_points_wrapper
    m.gather_points_grad_wrapper = gather_points_grad_wrapper
    return m


def _fake_pytorch3d():
    def _missing(name):
        def f(*a, **k):
            raise NotImplementedError(f"pytorch3d.{name} is not part of the forward hot path")
        return f

    p3d = types.ModuleType("pytorch3d")
    ops = types.ModuleType("pytorch3d.ops")
    ops.knn_points = native.knn_points
    ops.knn_gather = native.knn_gather
    loss = types.ModuleType("pytorch3d.loss")
    loss.chamfer_distance = _missing("loss.chamfer_distance")
    tr = types.ModuleType("pytorch3d.transforms")
    for n in ("matrix_to_euler_angles", "axis_angle_to_matrix", "rotation_6d_to_matrix",
              "euler_angles_to_matrix", "matrix_to_quaternion", "quaternion_to_matrix"):
        setattr(tr, n, _missing("transforms." + n))
    p3d.ops, p3d.loss, p3d.transforms = ops, loss, tr
    return {"pytorch3d": p3d, "pytorch3d.ops": ops, "pytorch3d.loss": loss, "pytorch3d.transforms": tr}


_loaded = None
_loaded_device = None


def _gpu_pytorch3d():
    """pytorch3d stand-in for the reference graph ON THE GPU (BASELINE.md plan B3): pytorch3d cannot be installed offline,
    so `knn_points` is the torch-library route to the same result, torch.cdist + topk (squared distances, ascending),
    and `knn_gather` the index gather it is.  Timing stand-in only -- never a parity oracle."""
    mods = _fake_pytorch3d()

    def knn_points(p1, p2, K=1, return_nn=False, **kw):
        d = torch.cdist(p1, p2) ** 2
        dists, idx = d.topk(K, dim=2, largest=False, sorted=True)
        nn = knn_gather(p2, idx) if return_nn else None
        return dists, idx, nn

    def knn_gather(x, idx):
        B, M, K = idx.shape
        return x[torch.arange(B, device=x.device)[:, None, None], idx]

    mods["pytorch3d.ops"].knn_points = knn_points
    mods["pytorch3d.ops"].knn_gather = knn_gather
    return mods


def _real_point_utils():
    """The reference's own extension compiled for sm_100a from its unmodified sources (oracle/build_ref.py)."""
    import importlib.util
    so = os.path.join(_ROOT, "oracle", "_ref", "point_utils_cuda.so")
    if not os.path.exists(so):
        raise FileNotFoundError(f"{so} not built (oracle/build_ref.py, build container only)")
    spec = importlib.util.spec_from_file_location("point_utils_cuda", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference(device="cpu"):
    """Returns a namespace with the reference's HRegNet / Model_V2 classes and its layers module.
    device="cpu": the four CPU redirections + oracle native ops (parity oracle, CPU reference arm).
    device="cuda": nothing redirected -- the reference's own CUDA extension (oracle/_ref) and a torch cdist+topk
    stand-in for pytorch3d: the reference graph as it runs on the GPU box (timing only).  One mode per process."""
    global _loaded, _loaded_device
    if _loaded is not None:
        if _loaded_device != device:
            raise RuntimeError(f"reference already loaded for {_loaded_device}")
        return _loaded
    if not available():
        raise FileNotFoundError(f"reference not found under {REF}")
    if device == "cuda":
        sys.modules["point_utils_cuda"] = _real_point_utils()
        sys.modules.update(_gpu_pytorch3d())
    else:
        # CPU redirection of the hard-coded device moves
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.IntTensor = lambda *s: torch.empty(*s, dtype=torch.int32)
        torch.cuda.FloatTensor = lambda *s: torch.empty(*s, dtype=torch.float32)
        sys.modules["point_utils_cuda"] = _fake_point_utils()
        sys.modules.update(_fake_pytorch3d())
    _loaded_device = device
    # models/__init__.py imports the spconv-based model_v6 (absent here): register a bare package instead
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = pkg
    import models.utils as U  # noqa: E402  (the reference's own wrappers, unmodified)
    for n in ("furthest_point_sample", "weighted_furthest_point_sample", "gather_operation", "set_seed"):
        setattr(pkg, n, getattr(U, n))
    from models.HRegNet import layers as L
    from models.HRegNet.models import HRegNet, HierFeatureExtraction
    ns = types.SimpleNamespace(utils=U, layers=L, HRegNet=HRegNet, HierFeatureExtraction=HierFeatureExtraction)
    try:
        from models.model_v2 import layers as L2
        from models.model_v2.models import Model_V2
        ns.layers_v2, ns.Model_V2 = L2, Model_V2
    except Exception as e:  # pragma: no cover
        ns.layers_v2, ns.Model_V2, ns.v2_error = None, None, e
    try:
        from models.model_v4 import layers as L4
        from models.model_v4.models import Model_V4
        ns.layers_v4, ns.Model_V4 = L4, Model_V4
    except Exception as e:  # pragma: no cover
        ns.layers_v4, ns.Model_V4, ns.v4_error = None, None, e
    _loaded = ns
    return ns


_losses = None


def load_reference_losses():
    """The reference's losses/losses.py, unmodified, loaded as a stand-alone module (its package __init__ pulls in
    unrelated losses with absent dependencies).  pytorch3d.transforms.matrix_to_euler_angles is absent: the restated
    conversion of oracle/ref_metrics.py is injected for it, so only the non-Euler outputs are pinned by the reference."""
    global _losses
    if _losses is not None:
        return _losses
    import importlib.util
    import numpy as np
    from . import ref_metrics
    load_reference()                                            # CPU redirections + pytorch3d stubs
    spec = importlib.util.spec_from_file_location("_ref_losses", os.path.join(REF, "losses", "losses.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.matrix_to_euler_angles = lambda M, convention="XYZ": torch.from_numpy(
        ref_metrics.matrix_to_euler_angles_xyz(M.detach().cpu().numpy())).to(M.dtype)
    _losses = mod
    return mod


def load_reference_calibeval():
    """The reference's metrics/calibeval.py, unmodified, as a stand-alone module.  Its imports `config.Config` and
    `transform.SO3` (a type annotation and an unused member) are satisfied by empty stand-ins; pytorch3d's
    matrix_to_euler_angles is absent, so the restated conversion of oracle/ref_metrics.py is injected for it (as for the
    losses) -- the evaluator's own arithmetic, bookkeeping and dictionary layout are the reference's."""
    import importlib.util
    from . import ref_metrics
    load_reference()
    cfg = types.ModuleType("config"); cfg.Config = object
    tr = types.ModuleType("transform"); tr.SO3 = lambda *a, **k: None
    saved = {k: sys.modules.get(k) for k in ("config", "transform")}
    sys.modules["config"], sys.modules["transform"] = cfg, tr
    try:
        spec = importlib.util.spec_from_file_location("_ref_calibeval", os.path.join(REF, "metrics", "calibeval.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod.matrix_to_euler_angles = lambda M, convention="XYZ": torch.from_numpy(
        ref_metrics.matrix_to_euler_angles_xyz(M.detach().cpu().numpy())).to(M.dtype)
    return mod


def load_reference_file(rel_path, name, stubs=()):
    """An unmodified reference source file loaded as a stand-alone module; `stubs` = names of absent third-party modules
    (e.g. open3d) registered as empty modules first -- the functions exercised by the tests do not touch them."""
    import importlib.util
    for st in stubs:
        sys.modules.setdefault(st, types.ModuleType(st))
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel_path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Args:
    use_fps = True
    use_weights = True
    freeze_detector = False
    freeze_feats = False


def pretrained_feats_path():
    return os.path.join(REF, "ckpt", "pretrained", "nusc_feats.pth")


def randomize_bn_(module, gen):
    """Give every BatchNorm non-trivial running stats / affine params so BN folding is actually tested."""
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            n = m.num_features
            m.running_mean.copy_(torch.randn(n, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=gen) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(n, generator=gen) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(n, generator=gen) * 0.1)


def build_reference_hregnet(seed=7, pretrained=True, randomize_bn=True, device="cpu"):
    """HRegNet(args).eval(): feature extractor from ckpt/pretrained/nusc_feats.pth, registration heads seeded."""
    ns = load_reference(device)
    torch.manual_seed(seed)
    net = ns.HRegNet(Args())
    if pretrained and os.path.isfile(pretrained_feats_path()):
        net.feature_extraction.load_state_dict(torch.load(pretrained_feats_path(), map_location="cpu"))
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for name in ("coarse_corres", "fine_corres_2", "fine_corres_1"):
            randomize_bn_(getattr(net, name), g)
    return net.eval().to(device) if device != "cpu" else net.eval()
