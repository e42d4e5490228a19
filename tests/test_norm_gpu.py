"""Fused BatchNorm(+residual)(+ReLU) kernels (csrc/cuda/bn_fused.cu) against an fp32 PyTorch reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, res, gamma, beta, relu, eps=1e-5):
    """fp32 reference on the bf16-rounded inputs."""
    x32 = x.float().requires_grad_(True)
    r32 = res.float().requires_grad_(True) if res is not None else None
    g32, b32 = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    y = F.batch_norm(x32, None, None, g32, b32, True, 0.1, eps)
    if r32 is not None:
        y = y + r32
    if relu:
        y = F.relu(y)
    return x32, r32, g32, b32, y


@pytest.mark.parametrize("shape", [(128, 64, 32, 32), (32, 128, 16, 16), (16, 256, 8, 8), (128, 512, 4, 4), (3, 8, 5, 7), (2, 2048, 4, 4)])
@pytest.mark.parametrize("relu,with_res", [(True, True), (True, False), (False, False), (False, True)])
def test_fused_bn_matches_fp32_reference(shape, relu, with_res):
    """Statistics + apply (forward) and reduce + apply (backward) streaming kernels vs an fp32 reference."""
    from draco_b200.ops.norm import FusedBatchNorm2d, backend_counters
    dev = torch.device("cuda", 0)
    torch.manual_seed(sum(shape))
    n, c, h, w = shape
    x = (torch.randn(shape, device=dev) * 1.7 + 0.3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    res = torch.randn(shape, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if with_res else None
    bn = FusedBatchNorm2d(c).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    xg = x.clone().requires_grad_(True)
    rg = res.clone().requires_grad_(True) if with_res else None
    before = backend_counters["fused"]
    y = bn(xg, residual=rg, relu=relu)
    assert backend_counters["fused"] == before + 1, "fused kernel path did not run"
    gy = torch.randn_like(y)
    y.backward(gy)
    torch.cuda.synchronize()

    x32, r32, g32, b32, yr = _ref(x, res, bn.weight, bn.bias, relu)
    yr.backward(gy.float())
    tol = 2e-2
    assert torch.allclose(y.float(), yr, atol=tol, rtol=tol), float((y.float() - yr).abs().max())
    # where the bf16 output rounds across the ReLU kink the masks may differ; compare gradients on agreeing elements
    agree = ((y.float() > 0) == (yr > 0)) if relu else torch.ones_like(yr, dtype=torch.bool)
    assert agree.float().mean() > 0.995
    dx_err = ((xg.grad.float() - x32.grad).abs() * agree).max() / x32.grad.abs().max().clamp_min(1e-6)
    assert dx_err < 3e-2, float(dx_err)
    if with_res:
        dr_err = ((rg.grad.float() - r32.grad).abs() * agree).max() / r32.grad.abs().max().clamp_min(1e-6)
        assert dr_err < 2e-2, float(dr_err)
    scale = max(1.0, float(g32.grad.abs().max()))
    assert float((bn.weight.grad - g32.grad).abs().max()) < 3e-2 * scale
    assert float((bn.bias.grad - b32.grad).abs().max()) < 3e-2 * max(1.0, float(b32.grad.abs().max()))
    # running statistics follow nn.BatchNorm2d
    ref_bn = torch.nn.BatchNorm2d(c).to(dev)
    ref_bn(x.float())
    assert torch.allclose(bn.running_mean, ref_bn.running_mean, atol=2e-3)
    assert torch.allclose(bn.running_var, ref_bn.running_var, atol=5e-3, rtol=5e-3)
    assert int(bn.num_batches_tracked) == 1


def test_fused_bn_is_deterministic_and_eval_falls_back():
    from draco_b200.ops.norm import FusedBatchNorm2d, backend_counters
    dev = torch.device("cuda", 0)
    x = torch.randn(64, 128, 16, 16, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    bn = FusedBatchNorm2d(128).to(dev)
    outs, grads = [], []
    for _ in range(3):
        xg = x.clone().requires_grad_(True)
        y = bn(xg, relu=True)
        y.backward(torch.ones_like(y))
        outs.append(y.detach().clone())
        grads.append((xg.grad.clone(), bn.weight.grad.clone()))
        bn.weight.grad = None
        bn.bias.grad = None
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    assert all(torch.equal(grads[0][0], g[0]) and torch.equal(grads[0][1], g[1]) for g in grads[1:])
    bn.eval()
    before = backend_counters["aten"]
    bn(x)
    assert backend_counters["aten"] == before + 1


def test_resnet18_fused_vs_aten_training_step():
    """Whole-model check: one bf16 ResNet-18 fwd/bwd with the fused BN kernels is as close to the fp32 model as the ATen
    bf16 path is (bf16 noise through 20 layers makes the two bf16 runs differ from each other by design)."""
    import os
    from draco_b200.models import ResNet18
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    x32 = torch.randn(64, 3, 32, 32, device=dev)
    y = torch.randint(0, 10, (64,), device=dev)

    def run(mode):
        torch.manual_seed(1)
        m = ResNet18().to(dev)
        xin = x32
        if mode != "fp32":
            os.environ["DRACO_BN"] = mode
            for mod in m.modules():
                if not isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                    for p in mod.parameters(recurse=False):
                        p.data = p.data.to(torch.bfloat16)
            xin = x32.to(torch.bfloat16)
        m = m.to(memory_format=torch.channels_last)
        loss = F.cross_entropy(m(xin.contiguous(memory_format=torch.channels_last)).float(), y)
        loss.backward()
        return float(loss), [p.grad.detach().float().clone() for p in m.parameters()]

    from draco_b200.ops.norm import backend_counters
    ref_loss, ref_g = run("fp32")
    before = backend_counters["fused"]
    f_loss, f_g = run("fused")
    assert backend_counters["fused"] - before == 20, "all 20 BatchNorm layers of ResNet-18 should take the fused path"
    a_loss, a_g = run("aten")
    os.environ["DRACO_BN"] = "fused"

    def rel(gs):
        num = sum(float((g - r).norm() ** 2) for g, r in zip(gs, ref_g)) ** 0.5
        den = sum(float(r.norm() ** 2) for r in ref_g) ** 0.5
        return num / den

    ef, ea = rel(f_g), rel(a_g)
    assert abs(f_loss - ref_loss) < 0.05 and abs(a_loss - ref_loss) < 0.05
    assert ef < 1.5 * ea + 0.02, (ef, ea)


@pytest.mark.parametrize("shape", [(128, 512, 4, 4), (128, 64, 32, 32), (5, 64, 8, 8), (7, 1024, 2, 2)])
def test_fused_bn_is_deterministic(shape):
    """Fixed summation orders everywhere: two runs agree bit for bit (forward output, running statistics, every gradient)."""
    from draco_b200.ops import norm
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    x = torch.randn(shape, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(shape, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    outs = []
    for _ in range(2):
        bn = norm.FusedBatchNorm2d(shape[1]).to(dev)
        xi = x.clone().requires_grad_(True)
        y = bn(xi, relu=True)
        y.backward(gy)
        outs.append((y.detach(), xi.grad, bn.weight.grad, bn.bias.grad, bn.running_var.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("shape", [(128, 64, 32, 32), (16, 256, 8, 8), (2, 2048, 4, 4), (3, 8, 5, 7)])
def test_fused_bn_relu_mask_recomputed_from_x_is_bit_identical(shape, monkeypatch):
    """DRACO_BN_MASK=x: backward recomputes the ReLU mask from x with the forward's own arithmetic instead of reading the saved
    output (one saved activation less).  Same mask, so the gradients are bit-identical to the default path."""
    from draco_b200.ops.norm import FusedBatchNorm2d
    dev = torch.device("cuda", 0)
    torch.manual_seed(sum(shape))
    x = (torch.randn(shape, device=dev) * 1.7 + 0.3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(shape, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    grads = {}
    for mode in ("y", "x"):
        monkeypatch.setenv("DRACO_BN_MASK", mode)
        bn = FusedBatchNorm2d(shape[1]).to(dev)
        with torch.no_grad():
            bn.weight.copy_(torch.linspace(0.5, 1.5, shape[1]))
            bn.bias.copy_(torch.linspace(-0.5, 0.5, shape[1]))
        xg = x.clone().requires_grad_(True)
        y = bn(xg, relu=True)
        y.backward(gy)
        grads[mode] = (y.detach().clone(), xg.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone())
    for a, b in zip(grads["y"], grads["x"]):
        assert torch.equal(a, b)
