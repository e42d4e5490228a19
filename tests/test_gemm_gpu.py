"""tcgen05 GEMM (csrc/cuda/gemm_tcgen05.cu) and the Linear layer built on it, against fp32 PyTorch references."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from draco_b200.ops import kernels
    return kernels


def _rel_err(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6))


SHAPES = [(128, 128, 64), (256, 192, 512), (300, 504, 784), (128, 800, 784), (77, 136, 72), (4096, 512, 4608),
          (1000, 16, 512), (64, 64, 8), (130, 264, 1032)]


@pytest.mark.parametrize("M,N,Kd", SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_all_operand_orders(K, M, N, Kd, a_mn, b_mn):
    dev = torch.device("cuda", 0)
    torch.manual_seed(M * 7 + N * 3 + Kd)
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operand needs a 16-byte row pitch")
    A = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
    B = torch.randn(N, Kd, device=dev).to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    A_in = A.t().contiguous() if a_mn else A
    B_in = B.t().contiguous() if b_mn else B
    out = K.gemm_bf16(A_in, B_in, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert out.shape == (M, N)
    assert _rel_err(out, ref) < 2e-3, _rel_err(out, ref)
    out16 = K.gemm_bf16(A_in, B_in, a_mn=a_mn, b_mn=b_mn)
    assert _rel_err(out16, ref) < 1.5e-2


@pytest.mark.parametrize("block_n", [32, 64, 128, 256])
def test_gemm_block_n_variants_bias_relu_accumulate(K, block_n):
    dev = torch.device("cuda", 0)
    torch.manual_seed(block_n)
    M, N, Kd = 515, 328, 1000
    A = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
    B = torch.randn(N, Kd, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    ref = torch.relu(A.float() @ B.float().t() + bias)
    out = K.gemm_bf16(A, B, out_dtype=torch.float32, bias=bias, relu=True, block_n=block_n)
    assert _rel_err(out, ref) < 2e-3
    out_b = K.gemm_bf16(A, B, bias=bias.to(torch.bfloat16), relu=True, block_n=block_n)
    assert _rel_err(out_b, ref) < 2e-2
    acc = torch.ones(M, N, device=dev)
    K.gemm_bf16(A, B, out=acc, accumulate=True, block_n=block_n)
    assert _rel_err(acc, A.float() @ B.float().t() + 1.0) < 2e-3


def test_gemm_odd_ldc_scalar_store_path(K):
    dev = torch.device("cuda", 0)
    A = torch.randn(200, 512, device=dev).to(torch.bfloat16)
    B = torch.randn(10, 512, device=dev).to(torch.bfloat16)       # the CIFAR classifier: N = 10
    ref = A.float() @ B.float().t()
    assert _rel_err(K.gemm_bf16(A, B, out_dtype=torch.float32), ref) < 2e-3
    assert _rel_err(K.gemm_bf16(A, B), ref) < 1.5e-2


def test_gemm_is_deterministic(K):
    dev = torch.device("cuda", 0)
    A = torch.randn(1024, 2048, device=dev).to(torch.bfloat16)
    B = torch.randn(768, 2048, device=dev).to(torch.bfloat16)
    o1 = K.gemm_bf16(A, B)
    o2 = K.gemm_bf16(A, B)
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("batch,fin,fout", [(128, 512, 512), (128, 784, 800), (128, 512, 10), (96, 800, 500)])
def test_linear_layer_forward_backward(batch, fin, fout):
    from draco_b200.ops.linear import Linear, backend_counters
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    lin = Linear(fin, fout).to(dev).to(torch.bfloat16)
    x = torch.randn(batch, fin, device=dev).to(torch.bfloat16).requires_grad_(True)
    before = backend_counters["tcgen05"]
    y = lin(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    assert backend_counters["tcgen05"] > before, "the tensor-core path did not run"
    w32, b32 = lin.weight.detach().float().requires_grad_(True), lin.bias.detach().float().requires_grad_(True)
    x32 = x.detach().float().requires_grad_(True)
    y32 = F.linear(x32, w32, b32)
    y32.backward(gy.float())
    assert _rel_err(y, y32) < 1.5e-2
    assert _rel_err(x.grad, x32.grad) < 2e-2
    assert _rel_err(lin.weight.grad, w32.grad) < 2e-2
    assert _rel_err(lin.bias.grad, b32.grad) < 2e-2


@pytest.mark.parametrize("n,cin,cout,hw", [(32, 64, 256, 16), (16, 256, 64, 8), (8, 512, 2048, 4), (128, 64, 64, 32)])
def test_pointwise_conv_on_tcgen05(n, cin, cout, hw):
    from draco_b200.ops.conv import Conv2d, backend_counters
    dev = torch.device("cuda", 0)
    torch.manual_seed(cin + cout)
    conv = Conv2d(cin, cout, kernel_size=1, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.randn(n, cin, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    before = backend_counters["tcgen05"]
    y = conv(x)
    assert backend_counters["tcgen05"] == before + 1
    assert y.shape == (n, cout, hw, hw) and y.is_contiguous(memory_format=torch.channels_last)
    gy = torch.randn_like(y)
    y.backward(gy)
    x32 = x.detach().float().requires_grad_(True)
    w32 = conv.weight.detach().float().requires_grad_(True)
    y32 = F.conv2d(x32, w32)
    y32.backward(gy.float())
    assert _rel_err(y, y32) < 1.5e-2
    assert _rel_err(x.grad, x32.grad) < 2e-2
    assert _rel_err(conv.weight.grad, w32.grad) < 2e-2
    # strided / 3x3 convolutions take the cuDNN path
    c3 = Conv2d(cin, cout, kernel_size=3, padding=1, bias=False).to(dev).to(torch.bfloat16)
    b2 = backend_counters["cudnn"]
    c3(x.detach())
    assert backend_counters["cudnn"] == b2 + 1


@pytest.mark.parametrize("n,cin,cout,hw", [(4, 64, 64, 32), (8, 128, 128, 16), (16, 256, 256, 8), (32, 512, 512, 4), (5, 64, 128, 16),
                                           (128, 64, 64, 32), (2, 512, 64, 2)])
def test_conv3x3_tcgen05_fprop_and_dgrad(n, cin, cout, hw):
    """TMA-patch implicit-GEMM 3x3 convolution (csrc/cuda/conv_tcgen05.cu) vs F.conv2d in fp32."""
    from draco_b200.ops.conv import conv3x3_tcgen05
    dev = torch.device("cuda", 0)
    torch.manual_seed(n + cin + cout + hw)
    x = torch.randn(n, cin, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn(cout, device=dev)
    y = conv3x3_tcgen05(x, w, False, b)
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert _rel_err(y, ref) < 1.5e-2, _rel_err(y, ref)
    dy = torch.randn(n, cout, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dx = conv3x3_tcgen05(dy, w, True)
    dref = torch.nn.grad.conv2d_input(x.shape, w.float(), dy.float(), padding=1)
    assert _rel_err(dx, dref) < 1.5e-2, _rel_err(dx, dref)
    assert torch.equal(conv3x3_tcgen05(x, w, False, b), y)            # deterministic


@pytest.mark.parametrize("n,cin,cout,hw", [(4, 64, 64, 32), (128, 64, 64, 32), (128, 128, 128, 16), (64, 256, 256, 8), (128, 512, 512, 4),
                                           (5, 64, 128, 16), (7, 512, 64, 2), (3, 128, 192, 4)])
def test_conv3x3_tcgen05_wgrad(n, cin, cout, hw):
    """split-K weight gradient (MN-major dy and x patches through 4-D TMA) vs conv2d_weight in fp32."""
    from draco_b200.ops.conv import conv3x3_wgrad_tcgen05
    dev = torch.device("cuda", 0)
    torch.manual_seed(n * 3 + cin + cout + hw)
    x = torch.randn(n, cin, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(n, cout, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dw = conv3x3_wgrad_tcgen05(dy, x)
    ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, 3, 3), dy.float(), padding=1)
    assert dw.shape == ref.shape and dw.is_contiguous(memory_format=torch.channels_last)
    assert _rel_err(dw, ref) < 1.5e-2, _rel_err(dw, ref)
    assert torch.equal(conv3x3_wgrad_tcgen05(dy, x), dw)              # deterministic split-K


def test_conv3x3_layer_autograd_path(monkeypatch):
    from draco_b200.ops.conv import Conv2d, backend_counters
    monkeypatch.setenv("DRACO_CONV3X3", "tcgen05")
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    conv = Conv2d(64, 128, 3, padding=1, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.randn(16, 64, 16, 16, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    before = backend_counters["tcgen05"]
    y = conv(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    assert backend_counters["tcgen05"] >= before + 3                  # fprop + dgrad + wgrad
    x32, w32 = x.detach().float().requires_grad_(True), conv.weight.detach().float().requires_grad_(True)
    y32 = F.conv2d(x32, w32, padding=1)
    y32.backward(gy.float())
    assert _rel_err(y, y32) < 1.5e-2 and _rel_err(x.grad, x32.grad) < 2e-2 and _rel_err(conv.weight.grad, w32.grad) < 2e-2


# ---------------------------------------------------------------------------------------------------------------------
# tap-table kernels (strided / 1x1 / 3x3) and the native stem: numerics validated on a B200 at the end of round 1 (14/14);
# still opt-in in the model path (DRACO_CONV_STRIDED / DRACO_CONV_STEM) until they have been timed against cuDNN.
# ---------------------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("n,cin,cout,hw,ks,stride", [(128, 64, 128, 32, 3, 2), (128, 128, 256, 16, 3, 2), (128, 256, 512, 8, 3, 2),
                                                     (128, 64, 128, 32, 1, 2), (64, 128, 256, 16, 1, 2), (5, 256, 512, 8, 1, 2),
                                                     (8, 64, 64, 32, 3, 1), (16, 128, 128, 16, 3, 1), (6, 64, 128, 16, 1, 1)])
def test_convg_tap_table_kernels(n, cin, cout, hw, ks, stride):
    from draco_b200.ops.conv import convg_tcgen05, convg_wgrad_tcgen05
    dev = torch.device("cuda", 0)
    torch.manual_seed(n + cin + cout + hw + ks + stride)
    pad = ks // 2
    x = torch.randn(n, cin, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(cout, cin, ks, ks, device=dev) * 0.05).to(torch.bfloat16)
    w = w.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)                 # [Cout, ks, ks, Cin] storage
    b = torch.randn(cout, device=dev)
    y = convg_tcgen05(x, w, (hw, hw), stride, False, b)
    ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=pad)
    assert y.shape == ref.shape and _rel_err(y, ref) < 1.5e-2, _rel_err(y, ref)
    dy = torch.randn_like(ref).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dx = convg_tcgen05(dy, w, (hw, hw), stride, True)
    dref = torch.nn.grad.conv2d_input(x.shape, w.float(), dy.float(), stride=stride, padding=pad)
    assert dx.shape == dref.shape and _rel_err(dx, dref) < 1.5e-2, _rel_err(dx, dref)
    dw = convg_wgrad_tcgen05(dy, x, ks, stride)
    wref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, ks, ks), dy.float(), stride=stride, padding=pad)
    assert dw.shape == wref.shape and _rel_err(dw, wref) < 1.5e-2, _rel_err(dw, wref)
    assert torch.equal(convg_tcgen05(x, w, (hw, hw), stride, False, b), y)


def test_convg_layer_autograd_path(monkeypatch):
    from draco_b200.ops.conv import Conv2d, backend_counters
    monkeypatch.setenv("DRACO_CONV_STRIDED", "tcgen05")
    dev = torch.device("cuda", 0)
    torch.manual_seed(9)
    for ks in (3, 1):
        conv = Conv2d(64, 128, ks, stride=2, padding=ks // 2, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        x = torch.randn(16, 64, 32, 32, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        before = backend_counters["tcgen05"]
        y = conv(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        assert backend_counters["tcgen05"] >= before + 2
        x32, w32 = x.detach().float().requires_grad_(True), conv.weight.detach().float().requires_grad_(True)
        y32 = F.conv2d(x32, w32, stride=2, padding=ks // 2)
        y32.backward(gy.float())
        assert _rel_err(y, y32) < 1.5e-2 and _rel_err(x.grad, x32.grad) < 2e-2 and _rel_err(conv.weight.grad, w32.grad) < 2e-2


@pytest.mark.parametrize("n,hw", [(128, 32), (5, 32), (16, 16), (3, 64)])
def test_conv_stem_native_kernels(n, hw):
    from draco_b200.ops.conv import conv_stem_fprop, conv_stem_wgrad
    dev = torch.device("cuda", 0)
    torch.manual_seed(n + hw)
    x = torch.randn(n, 3, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 3, 3, 3, device=dev) * 0.2).to(torch.bfloat16).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    b = torch.randn(64, device=dev)
    y = conv_stem_fprop(x, w, b)
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last) and _rel_err(y, ref) < 1e-2
    dy = torch.randn(n, 64, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dw = conv_stem_wgrad(dy, x)
    wref = torch.nn.grad.conv2d_weight(x.float(), (64, 3, 3, 3), dy.float(), padding=1)
    assert dw.shape == wref.shape and _rel_err(dw, wref) < 1e-2, _rel_err(dw, wref)
    assert torch.equal(conv_stem_wgrad(dy, x), dw)


@pytest.mark.timeout(120)       # a never-run kernel that hangs must not eat the GPU budget
@pytest.mark.skipif(__import__("os").environ.get("DRACO_EXPERIMENTAL", "0") != "1",
                    reason="conv_halo_tcgen05.cu has not run on hardware yet; set DRACO_EXPERIMENTAL=1 (and try DRACO_HALO_DESC=0 / 1)")
@pytest.mark.parametrize("n,hw", [(4, 32), (128, 32), (3, 16), (16, 64)])
@pytest.mark.parametrize("pw,desc", [(10, 0), (10, 1), (16, 0), (16, 1)])
def test_conv3x3_halo_reuse_kernels(n, hw, pw, desc, monkeypatch):
    """Halo patch loaded once per tile, nine taps through row-shifted UMMA descriptors (64 -> 64 channels).  The four
    (patch pitch, base-offset convention) combinations are all run: at least one must be numerically right -- see the header of
    csrc/cuda/conv_halo_tcgen05.cu; keep the cheapest passing one as the default."""
    monkeypatch.setenv("DRACO_HALO_PW", str(pw))
    monkeypatch.setenv("DRACO_HALO_DESC", str(desc))
    from draco_b200.ops.conv import conv3x3_halo
    dev = torch.device("cuda", 0)
    torch.manual_seed(n + hw)
    x = torch.randn(n, 64, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 64, 3, 3, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn(64, device=dev)
    y = conv3x3_halo(x, w, False, b)
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    assert _rel_err(y, ref) < 1.5e-2, _rel_err(y, ref)
    dy = torch.randn(n, 64, hw, hw, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dx = conv3x3_halo(dy, w, True)
    dref = torch.nn.grad.conv2d_input(x.shape, w.float(), dy.float(), padding=1)
    assert _rel_err(dx, dref) < 1.5e-2, _rel_err(dx, dref)
    assert torch.equal(conv3x3_halo(x, w, False, b), y)
