"""Host C++ natives: locator (N1), codec (N4), ABI of the CUDA library."""
import ctypes as C
import itertools

import numpy as np
import pytest

from draco_b200 import _native as N
from draco_b200.codes import cyclic
from draco_b200.utils import codec


def _locate(E, n, s, tol=1e-6):
    Eh = np.ascontiguousarray(np.stack([E.real, E.imag], -1)).astype(np.float64)
    T = Eh.shape[0]
    v = np.zeros((T, n, 2))
    mask = np.zeros(T, dtype=np.uint32)
    fl = np.zeros(T, dtype=np.int32)
    assert N.host().drc_host_locate(Eh.ctypes.data, T, n, s, tol, v.ctypes.data, mask.ctypes.data, fl.ctypes.data) == 0
    return v[..., 0] + 1j * v[..., 1], mask, fl


@pytest.mark.parametrize("n,s", [(5, 1), (7, 2), (7, 3), (9, 2)])
def test_cpp_locator_matches_numpy_oracle(n, s):
    rng = np.random.RandomState(0)
    c = cyclic.search_w(n, s)
    G = rng.randn(n, 64)
    f = rng.randn(64) + 1.0
    R0 = np.stack([cyclic.encode(c, i, G) for i in range(n)])
    for k in range(s + 1):
        for liars in itertools.islice(itertools.combinations(range(n), k), 12):
            R = R0.copy()
            for l in liars:
                R[l] += -100.0 * (1 + rng.rand(64))
            v, mask, fl = _locate((R @ f)[None], n, s)
            assert fl[0] >= len(liars) and fl[0] <= s
            for l in liars:
                assert not (mask[0] >> l) & 1, "a liar was used for recombination"
            dec = np.real(v[0] @ R)
            assert np.abs(dec - G.sum(0)).max() < 1e-6 * max(1.0, np.abs(G).sum())


def test_cpp_solve_poly_a_is_the_locator_polynomial():
    n, s = 7, 2
    rng = np.random.RandomState(3)
    c = cyclic.search_w(n, s)
    G = rng.randn(n, 32)
    R = np.stack([cyclic.encode(c, i, G) for i in range(n)])
    R[1] += 100.0
    R[4] -= 70.0
    E = R @ (rng.randn(32) + 1.0)
    Eh = np.ascontiguousarray(np.stack([E.real, E.imag], -1))
    alpha = np.zeros((s, 2))
    assert N.host().drc_host_solve_poly_a(Eh.ctypes.data, n, s, alpha.ctypes.data) == 0
    p = cyclic.locator_values(alpha[:, 0] + 1j * alpha[:, 1], n)
    mag = np.abs(p)
    assert set(np.argsort(mag)[:2]) == {1, 4} and mag[[1, 4]].max() < 1e-6 * mag.max()


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.int32, np.uint8])
@pytest.mark.parametrize("shape", [(0,), (1,), (4097,), (64, 3, 3, 3), (10000,)])
def test_codec_roundtrip(dtype, shape):
    rng = np.random.RandomState(0)
    a = (rng.randn(*shape) * 1e-3).astype(dtype)
    b = codec.decompress(codec.compress(a))
    assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_codec_compresses_structured_gradients():
    z = np.zeros(100000, dtype=np.float32)
    assert codec.ratio(z) < 0.01
    g = (np.random.RandomState(0).randn(100000) * 1e-3).astype(np.float32)
    assert codec.ratio(g) < 0.97              # sign/exponent plane packs, mantissa planes stay raw
    sparse = g.copy(); sparse[np.random.RandomState(1).rand(100000) < 0.9] = 0
    assert codec.ratio(sparse) <= 1.01
    with pytest.raises(ValueError):
        codec.decompress(b"garbage-not-a-stream-----")


def test_cuda_library_abi_matches_ctypes():
    lib = N.cuda()            # loads without a GPU (static cudart, libcuda resolved lazily)
    for name in ("PushArgs", "OmniArgs", "VoteArgs", "ResolveArgs", "UpdateArgs", "CastArgs", "WaitArgs", "SetFlagArgs",
                 "TensorMeta", "HyperParams", "TileView", "FlagList", "ProjectArgs", "LocateArgs", "GeoMedArgs",
                 "GeoMedPrepArgs", "PairDistArgs", "KrumSelectArgs"):
        assert getattr(lib, "drc_sizeof_" + name)() == C.sizeof(getattr(N, name)), name
