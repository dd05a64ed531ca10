"""CPU checks of the lazy-residue arithmetic the CUDA kernels are built on (csrc/glx.cuh, csrc/poseidon2_core.cuh): the
host halves of those headers, compiled with g++, must agree with the oracle on canonical results -- including for
non-canonical input representatives (x and x + p), which is the whole point of the lazy representation."""
import ctypes

import numpy as np
import pytest

from tests import oracle_lib

P = oracle_lib.P
vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def hc():
    return oracle_lib.build_hostcheck()


@pytest.fixture(scope="module")
def oracle():
    return oracle_lib.load()


def _edge_mix(rng, n):
    a = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)  # ANY 64-bit representative
    edges = np.array([0, 1, P - 1, P, P + 1, (1 << 64) - 1, (1 << 32) - 1, 1 << 32, (1 << 64) - (1 << 32)], dtype=np.uint64)
    a[: len(edges)] = edges
    return a


def test_lazy_mul_matches_oracle(hc, oracle):
    rng = np.random.default_rng(11)
    a, b = _edge_mix(rng, 4096), _edge_mix(rng, 4096)[::-1].copy()
    out = np.empty_like(a)
    hc.hc_glx_mul(a.ctypes.data_as(vp), b.ctypes.data_as(vp), out.ctypes.data_as(vp), ctypes.c_size_t(a.size))
    exp = np.array([(int(x) * int(y)) % P for x, y in zip(a, b)], dtype=np.uint64)
    assert (out == exp).all()


def test_lazy_sub_and_reduce96(hc):
    rng = np.random.default_rng(12)
    a, b = _edge_mix(rng, 4096), _edge_mix(rng, 4096)[::-1].copy()
    out = np.empty_like(a)
    hc.hc_glx_sub(a.ctypes.data_as(vp), b.ctypes.data_as(vp), out.ctypes.data_as(vp), ctypes.c_size_t(a.size))
    assert (out == np.array([(int(x) - int(y)) % P for x, y in zip(a, b)], dtype=np.uint64)).all()
    hi = rng.integers(0, 1 << 32, size=a.size, dtype=np.uint32)
    hi[:4] = [0, 1, (1 << 32) - 1, 1 << 31]
    hc.hc_glx_reduce96(a.ctypes.data_as(vp), hi.ctypes.data_as(vp), out.ctypes.data_as(vp), ctypes.c_size_t(a.size))
    assert (out == np.array([(int(x) + (int(h) << 64)) % P for x, h in zip(a, hi)], dtype=np.uint64)).all()


def test_lazy_poseidon2_matches_oracle_permutation(hc, oracle):
    rng = np.random.default_rng(13)
    st = rng.integers(0, 1 << 64, size=(512, 12), dtype=np.uint64)
    st[0, :] = 0
    st[1, :] = P - 1
    st[2, :] = (1 << 64) - 1
    st[3, :4] = P
    ref = (st % np.uint64(P)).copy()
    for row in ref:
        oracle.lib.orc_poseidon2_permute(row.ctypes.data_as(vp))
    mine = st.copy()
    hc.hc_p2x_permute(mine.ctypes.data_as(vp), ctypes.c_size_t(len(mine)))
    assert (mine == ref).all()


def test_mul_2exp_all_exponents(hc):
    rng = np.random.default_rng(14)
    a = rng.integers(0, P, size=2048, dtype=np.uint64)
    a[:6] = [0, 1, P - 1, (1 << 32) - 1, 1 << 32, P - (1 << 32)]
    out = np.empty_like(a)
    for s in range(96):
        assert hc.hc_glx_mul_2exp(a.ctypes.data_as(vp), out.ctypes.data_as(vp), ctypes.c_size_t(a.size), s) == 0
        exp = np.array([(int(x) << s) % P for x in a], dtype=np.uint64)
        assert (out == exp).all(), s


def test_reduce128_any_hi(hc):
    rng = np.random.default_rng(15)
    lo, hi = _edge_mix(rng, 4096), _edge_mix(rng, 4096)[::-1].copy()
    out = np.empty_like(lo)
    hc.hc_glx_reduce128(lo.ctypes.data_as(vp), hi.ctypes.data_as(vp), out.ctypes.data_as(vp), ctypes.c_size_t(lo.size))
    assert (out == np.array([(int(l) + (int(h) << 64)) % P for l, h in zip(lo, hi)], dtype=np.uint64)).all()
