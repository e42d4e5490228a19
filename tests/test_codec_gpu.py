"""Device codec (csrc/cuda/codec.cu): round trip on the GPU and stream compatibility with the host codec."""
import numpy as np
import pytest
import torch

from draco_b200.utils import codec

pytestmark = pytest.mark.gpu


def _cases():
    g = torch.Generator().manual_seed(0)
    yield "gauss", torch.randn(1_000_003, generator=g) * 1e-3
    yield "zeros", torch.zeros(50_000)
    yield "sparse", torch.randn(300_000, generator=g) * (torch.rand(300_000, generator=g) > 0.9)
    yield "tiny", torch.randn(5, generator=g)
    yield "one_block_exact", torch.randn(4096, generator=g)
    yield "int", torch.randint(-5, 5, (70_001,), generator=g, dtype=torch.int32)
    yield "bytes", torch.randint(0, 7, (9_999,), generator=g, dtype=torch.uint8)
    yield "complex", torch.view_as_complex(torch.randn(40_000, 2, generator=g) * 1e-2)


@pytest.mark.parametrize("name,t", list(_cases()))
def test_device_roundtrip_and_host_compat(name, t):
    dev = torch.device("cuda", 0)
    x = t.to(dev).contiguous()
    s = codec.compress_tensor(x)
    y = codec.decompress_tensor(s, x.dtype, x.shape)
    assert torch.equal(torch.view_as_real(y) if y.is_complex() else y, torch.view_as_real(x) if x.is_complex() else x)
    # the device stream is the host stream, byte for byte
    arr = (torch.view_as_real(t).numpy().view(np.complex64).reshape(t.shape) if t.is_complex() else t.numpy())
    host_msg = codec.compress(arr)
    hdr_len = len(host_msg) - (len(bytes(s.cpu().numpy())))
    assert host_msg[hdr_len:] == bytes(s.cpu().numpy())
    # and the host decoder reads it
    back = codec.decompress(host_msg[:hdr_len] + bytes(s.cpu().numpy()))
    assert np.array_equal(back.view(np.uint8), arr.view(np.uint8))
    if name in ("zeros", "gauss", "bytes"):       # randomly scattered sparsity does not pack plane-wise (no LZ stage)
        assert s.numel() < x.numel() * x.element_size()


def test_device_decoder_rejects_garbage():
    dev = torch.device("cuda", 0)
    with pytest.raises(ValueError):
        codec.decompress_tensor(torch.zeros(64, dtype=torch.uint8, device=dev), torch.float32, (4,))
