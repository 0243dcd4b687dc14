"""CPU (any box): the oracle restatement reproduces the committed reference outputs (tests/golden/*.npz)."""
import pytest
import torch

from oracle import ref_layers as RL
from common import build_product_hregnet, load_golden, unflatten


@pytest.mark.parametrize("name", ["hregnet_b2_n2048", "hregnet_uniform_b1_n1500"])
def test_oracle_reproduces_golden(name):
    torch.set_num_threads(8)
    gd = load_golden(name)
    sd = build_product_hregnet(seed=7).state_dict()     # same keys / values as the reference net (see common.py)
    trace = {}
    with torch.no_grad():
        out = RL.hregnet_forward(sd, gd["src"], gd["dst"], trace=trace)
    for side in ("src", "dst"):
        for lv in (1, 2, 3):
            assert torch.equal(trace[f"{side}_trace"][f"fps_idx_{lv}"], gd[f"{side}_fps_idx_{lv}"])
        for k, v in unflatten(gd, f"{side}_feats.").items():
            assert torch.equal(out[f"{side}_feats"][k], v), k            # bit-identical feature extraction
    assert torch.equal(trace["coarse_idx"], gd["coarse_idx"])
    for i in range(3):
        assert float(RL.rotation_angle_deg(out["rotation"][i], gd[f"rotation.{i}"]).max()) < 1e-4
        assert float((out["translation"][i] - gd[f"translation.{i}"]).abs().max()) < 1e-5
    for lv in (3, 2, 1):
        assert float((out[f"src_xyz_corres_{lv}"] - gd[f"src_xyz_corres_{lv}"]).abs().max()) < 1e-4
        assert float((out[f"src_dst_weights_{lv}"] - gd[f"src_dst_weights_{lv}"]).abs().max()) < 1e-6
