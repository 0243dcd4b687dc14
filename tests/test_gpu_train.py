"""GPU: the differentiable forward of the drop-in modules (pcd_reg_hregnet_b200/train_path.py).
Stage-wise and teacher-forced like the inference gates: with gradients recorded, every module must (i) return what its
fused inference path returns and (ii) give its parameters / inputs the gradients the oracle gives under CPU autograd
(the oracle's layers are plain torch functions on a state_dict: its tensors are made leaves here)."""
import pytest
import torch

from oracle import ref_layers as RL
from pcd_reg_hregnet_b200 import synth, train_path
from common import build_product_hregnet, load_golden, rel_err, unflatten

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def full_fp32_library_kernels():
    """The differentiable path runs its convolutions through ATen like the reference does, and ATen's cuDNN convolutions
    default to TF32 (10-bit significand) on this GPU.  The comparisons below are at fp32 accuracy: TF32 off for them."""
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved


def _leaf_sd(net):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone())
            for k, v in net.state_dict().items()}


def _grad_close(g_gpu, g_cpu, tol=2e-3):
    scale = float(g_cpu.abs().max())
    assert scale > 0
    err = float((g_gpu.cpu() - g_cpu).abs().max()) / scale
    assert err < tol, err
    return err


def test_level1_detector_descriptor_outputs_and_gradients():
    cpu, gpu = build_product_hregnet(seed=7), build_product_hregnet(seed=7, device=DEV)
    src = synth.make_batch([91, 92], 2048)[0]
    det, desc = gpu.feature_extraction.detector_1, gpu.feature_extraction.desc_extractor_1
    with torch.no_grad():                                            # fused inference path
        kp0, sig0, af0, G0, afm0 = det(src.to(DEV), None, None)
        d0 = desc(G0, afm0)
    gpu.zero_grad()
    kp, sig, af, G, afm = det(src.to(DEV), None, None)               # gradients recorded -> train_path
    d = desc(G, afm)
    assert kp.requires_grad and d.requires_grad
    assert rel_err(kp.detach().cpu(), kp0.cpu()) < 1e-4 and rel_err(sig.detach().cpu(), sig0.cpu()) < 1e-3
    assert rel_err(d.detach().cpu(), d0.cpu()) < 1e-3 and rel_err(af.detach().cpu(), af0.cpu()) < 1e-3
    (d.square().mean() + sig.mean() + kp.square().mean() * 1e-3).backward()
    sd = _leaf_sd(cpu)
    kp_o, sig_o, af_o, G_o, afm_o, _ = RL.keypoint_detector(sd, "feature_extraction.detector_1.", src, None, None, 1024, 64)
    d_o = RL.desc_extractor(sd, "feature_extraction.desc_extractor_1.", G_o, afm_o)
    (d_o.square().mean() + sig_o.mean() + kp_o.square().mean() * 1e-3).backward()
    for name in ("detector_1.convs.0.weight", "detector_1.convs.4.weight", "detector_1.mlp2.0.bias", "desc_extractor_1.mlp1.0.weight",
                 "desc_extractor_1.mlp2.1.bias"):
        g_gpu = dict(gpu.feature_extraction.named_parameters())[name].grad
        e = _grad_close(g_gpu, sd["feature_extraction." + name].grad)
        print(f"grad {name}: rel err {e:.1e}")


@pytest.mark.parametrize("stage", ["coarse", "fine2"])
def test_correspondence_stage_gradients(stage):
    cpu, gpu = build_product_hregnet(seed=7), build_product_hregnet(seed=7, device=DEV)
    gd = load_golden("hregnet_b2_n2048")
    S, D = unflatten(gd, "src_feats."), unflatten(gd, "dst_feats.")
    lv = 3 if stage == "coarse" else 2
    mod = gpu.coarse_corres if stage == "coarse" else gpu.fine_corres_2
    pfx = "coarse_corres." if stage == "coarse" else "fine_corres_2."
    sx = S[f"xyz_{lv}"] if stage == "coarse" else RL._apply(gd["rotation.0"], gd["translation.0"], S["xyz_2"])
    args = (sx, S[f"desc_{lv}"], D[f"xyz_{lv}"], D[f"desc_{lv}"], S[f"sigmas_{lv}"], D[f"sigmas_{lv}"])
    g = lambda t: t.to(DEV).contiguous()
    with torch.no_grad():
        cor0, w0 = mod(*[g(a) for a in args])
    gpu.zero_grad()
    sdesc = g(args[1]).requires_grad_(True)
    cor, w = mod(g(args[0]), sdesc, g(args[2]), g(args[3]), g(args[4]), g(args[5]))
    assert float((cor.detach() - cor0).abs().max()) < 1e-3 * float(cor0.abs().max()) and float((w.detach() - w0).abs().max()) < 1e-3
    (cor.square().mean() + w.mean()).backward()
    sd = _leaf_sd(cpu)
    sdesc_o = args[1].clone().requires_grad_(True)
    fn = RL.coarse_reg if stage == "coarse" else RL.fine_reg
    cor_o, w_o = fn(sd, pfx, args[0], sdesc_o, args[2], args[3], args[4], args[5])
    (cor_o.square().mean() + w_o.mean()).backward()
    names = ["convs_1.0.weight", "convs_1.6.weight", "mlp1.0.weight", "mlp3.0.bias"] + (["convs_2.0.weight"] if stage == "coarse" else [])
    for name in names:
        # convs_2 reaches the loss only through the max-normalised neighbour-aware cosine features (two divisions by row /
        # column maxima of a 256 x 256 matrix): its gradient is a small difference of large terms, fp32 summation order shows
        e = _grad_close(dict(mod.named_parameters())[name].grad, sd[pfx + name].grad, tol=1e-2 if name.startswith("convs_2") else 2e-3)
        print(f"{stage} grad {name}: rel err {e:.1e}")
    _grad_close(sdesc.grad, sdesc_o.grad)


def test_kabsch_backward_matches_autograd_of_the_formula():
    g = torch.Generator().manual_seed(3)
    B, N = 4, 256
    src = (torch.rand(B, N, 3, generator=g) * 2 - 1) * torch.tensor([40.0, 40.0, 3.0])
    R_gt = torch.stack([synth._rodrigues((torch.rand(3, generator=g, dtype=torch.float64) - 0.5) * 0.6) for _ in range(B)]).float()
    cor = torch.einsum("bij,bnj->bni", R_gt, src) + 0.05 * torch.randn(B, N, 3, generator=g)
    w = torch.rand(B, N, generator=g)
    GR, Gt = torch.randn(B, 3, 3, generator=g), torch.randn(B, 3, generator=g)
    leaves = [t.clone().to(DEV).requires_grad_(True) for t in (src, cor, w)]
    R, t = train_path.svd_head(*leaves)
    ((R * GR.to(DEV)).sum() + (t * Gt.to(DEV)).sum()).backward()
    ref = [t.clone().double().requires_grad_(True) for t in (src, cor, w)]
    R_o, t_o = train_path.kabsch_formula(*ref)
    ((R_o * GR.double()).sum() + (t_o * Gt.double()).sum()).backward()
    assert float(RL.rotation_angle_deg(R.detach().cpu(), R_o.detach().float()).max()) < 1e-4
    for a, b in zip(leaves, ref):
        _grad_close(a.grad, b.grad.float(), tol=1e-4)


def test_training_mode_steps_reduce_the_loss_and_update_batchnorm():
    """net.train(): BatchNorm on batch statistics, running statistics updated, every parameter gets a gradient, a few Adam
    steps on one batch reduce a pose loss; back in eval mode the fused inference path runs on the updated weights."""
    net = build_product_hregnet(seed=7, device=DEV)
    src, dst, R_gt, t_gt = synth.make_batch([95, 96], 4096)
    src, dst, R_gt, t_gt = src.to(DEV), dst.to(DEV), R_gt.to(DEV), t_gt.to(DEV)
    net.train()
    rm0 = net.coarse_corres.convs_1[1].running_mean.clone()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        out = net(src, dst)
        loss = sum(((R - R_gt) ** 2).mean() + ((t - t_gt) ** 2).mean() for R, t in zip(out["rotation"], out["translation"]))
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0], losses
    assert not torch.equal(rm0, net.coarse_corres.convs_1[1].running_mean)
    missing = [n for n, p in net.named_parameters() if p.grad is None]
    assert not missing, missing[:5]
    net.eval()
    with torch.no_grad():
        out = net(src, dst)                                          # fused path, re-folded weights
    assert all(torch.isfinite(v).all() for v in out["rotation"] + out["translation"])
    print("training losses", [round(l, 5) for l in losses])
