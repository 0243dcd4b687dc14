"""CPU: host-side logic of the product (no kernels run): C-ABI surface, BN folding, column permutations,
error behaviour, the closed-form pose from a covariance (library host function)."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

from oracle import ref_layers as RL
from pcd_reg_hregnet_b200 import _lib, fold, layers, ops
from common import Args, build_product_hregnet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hregnet_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(hrn_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 18
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert declared - {"hrn_version", "hrn_level_pack_bytes", "hrn_level_bias_count", "hrn_level_ws_pack_bytes",
                       "hrn_level_ws_bias_count"} == set(_lib.SIGNATURES), "python binding table out of sync with the header"
    assert b"sm_100a" in _lib.lib().hrn_version()


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.Seg) == 32 and ctypes.sizeof(_lib.Rows) == 4 * 32 + 8 + 16


def test_cpu_tensors_are_rejected_not_silently_computed():
    with pytest.raises(_lib.HrnError):
        ops.furthest_point_sample(torch.rand(1, 64, 3), 8)
    with pytest.raises(_lib.HrnError):
        ops.knn_points(torch.rand(1, 8, 3), torch.rand(1, 64, 3), K=4)


def test_training_mode_is_rejected():
    net = build_product_hregnet()
    net.train()
    with pytest.raises(RuntimeError):
        net.feature_extraction.detector_1.folded()


def test_fold_matches_torch_modules():
    torch.manual_seed(0)
    seq = nn.Sequential(nn.Conv2d(12, 20, 1, bias=False), nn.BatchNorm2d(20), nn.ReLU(),
                        nn.Conv2d(20, 7, 1, bias=False), nn.BatchNorm2d(7), nn.ReLU()).eval()
    for m in seq:
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.normal_(); m.running_var.uniform_(0.5, 2); m.weight.data.uniform_(0.5, 2); m.bias.data.normal_()
    x = torch.randn(3, 12, 5, 4)
    want = seq(x)
    rows = x.permute(0, 2, 3, 1).reshape(-1, 12)
    for W, b, act in fold.fold_sequential(seq):
        rows = rows @ W.t() + b
        if act == _lib.ACT_RELU:
            rows = torch.relu(rows)
    got = rows.view(3, 5, 4, 7).permute(0, 3, 1, 2)
    assert torch.allclose(got, want, atol=1e-5)
    h1 = nn.Sequential(nn.Conv1d(7, 7, 1), nn.BatchNorm1d(7), nn.ReLU()).eval()
    h1[1].running_mean.normal_(); h1[1].running_var.uniform_(0.5, 2)
    (W, b, act), = fold.fold_sequential(h1)
    y = torch.randn(2, 7, 9)
    assert torch.allclose(torch.relu(torch.einsum("oc,bcn->bon", W, y) + b[None, :, None]), h1(y), atol=1e-5)


def test_pair_perm_is_a_permutation_and_orders_blocks():
    for C, ns in ((64, 0), (256, 4)):
        p = fold.pair_perm(C, ns, "cpu")
        K = 2 * C + 12 + ns
        assert sorted(p.tolist()) == list(range(K))
        assert p[:10].tolist() == list(range(10)) and p[10].item() == 10 + 2 * C and p[12 + ns].item() == 10


def test_pose_from_covariance_host_matches_svd():
    g = torch.Generator().manual_seed(0)
    L = _lib.lib()
    for trial in range(50):
        H = torch.randn(3, 3, generator=g, dtype=torch.float64)
        if trial % 5 == 0:
            H[:, 2] = -H[:, 2]                                # negative determinant -> reflection branch
        xb, yb = torch.randn(3, generator=g, dtype=torch.float64), torch.randn(3, generator=g, dtype=torch.float64)
        R9, t3 = torch.empty(9, dtype=torch.float64), torch.empty(3, dtype=torch.float64)
        assert L.hrn_pose_from_covariance_host(H.contiguous().data_ptr(), xb.data_ptr(), yb.data_ptr(), R9.data_ptr(),
                                               t3.data_ptr()) == 0
        U, _, Vh = torch.linalg.svd(H)
        V = Vh.t()
        D = torch.diag(torch.tensor([1.0, 1.0, float(torch.det(V @ U.t()))], dtype=torch.float64))
        R = V @ D @ U.t()                                      # layers.py:495-499
        assert torch.allclose(R9.view(3, 3), R, atol=1e-10), trial
        assert torch.allclose(t3, yb - R @ xb, atol=1e-10)
        assert abs(float(torch.det(R9.view(3, 3))) - 1.0) < 1e-10
    # rank-1 covariance -> the reference's SVD-failure fallback (identity, zero)
    H = torch.zeros(3, 3, dtype=torch.float64); H[0, 0] = 1.0
    R9, t3 = torch.empty(9, dtype=torch.float64), torch.empty(3, dtype=torch.float64)
    z = torch.zeros(3, dtype=torch.float64)
    L.hrn_pose_from_covariance_host(H.data_ptr(), z.data_ptr(), z.data_ptr(), R9.data_ptr(), t3.data_ptr())
    assert torch.equal(R9.view(3, 3), torch.eye(3, dtype=torch.float64))


def test_module_signatures_match_reference_api():
    import inspect
    assert list(inspect.signature(layers.KeypointDetector.__init__).parameters)[1:] == ["nsample", "k", "in_channels", "out_channels", "fps"]
    assert list(inspect.signature(layers.DescExtractor.__init__).parameters)[1:] == ["in_channels", "out_channels", "C_detector", "desc_dim"]
    assert list(inspect.signature(layers.CoarseReg.__init__).parameters)[1:] == ["k", "in_channels", "use_sim", "use_neighbor"]
    assert list(inspect.signature(layers.FineReg.__init__).parameters)[1:] == ["k", "in_channels"]
    assert list(inspect.signature(layers.CoarseReg.forward).parameters)[1:] == ["src_xyz", "src_desc", "dst_xyz", "dst_desc", "src_weights", "dst_weights"]
    assert list(inspect.signature(layers.WeightedSVDHead.forward).parameters)[1:] == ["src", "src_corres", "weights"]
    for n in ("furthest_point_sampling_wrapper", "weighted_furthest_point_sampling_wrapper", "gather_points_wrapper",
              "gather_points_grad_wrapper"):
        assert hasattr(ops.point_utils_cuda, n)


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly ONE JSON line on stdout with the
    contract's keys, exit 0, no GPU needed.  Tiny sample so that it runs in seconds."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--points", "1024",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    # "reference" = the unmodified reference graph (/root/reference here, the staged baseline/_ref on the GPU box);
    # "port" only where neither exists
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert set(d["config"]) == {"workload", "pairs_per_gpu", "points", "parallelism"}       # the same dict as our arm's
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_refuses_to_run_without_a_gpu():
    """Our own arm has no CPU fallback: without a CUDA device it fails loudly instead of printing a number."""
    import subprocess
    import sys
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode != 0 and p.stdout.strip() == ""
    assert "needs a CUDA device" in p.stderr


def test_stack_clouds_is_a_view_when_the_clouds_are_adjacent():
    from pcd_reg_hregnet_b200 import engine
    both = torch.arange(2 * 3 * 5 * 3, dtype=torch.float32).reshape(6, 5, 3)
    src, dst = both[:3], both[3:]
    st = engine.stack_clouds(src, dst)
    assert st.data_ptr() == both.data_ptr() and st.shape == both.shape and torch.equal(st, both)
    # not adjacent / different allocations / swapped order: a copy with the same contents
    for a, b in ((src.clone(), dst.clone()), (dst, src), (both[:2], both[3:5])):
        st = engine.stack_clouds(a, b)
        assert st.data_ptr() not in (a.data_ptr(), b.data_ptr()) and torch.equal(st, torch.cat([a, b]))
    # a view in the middle of an allocation
    big = torch.randn(10, 5, 3)
    st = engine.stack_clouds(big[2:4], big[4:6])
    assert st.data_ptr() == big[2:].data_ptr() and torch.equal(st, big[2:6])


def test_registrar_refuses_a_cpu_net():
    from pcd_reg_hregnet_b200.runner import Registrar
    with pytest.raises(RuntimeError, match="CUDA device"):
        Registrar(build_product_hregnet(seed=7), batch=2, n_points=1024)
