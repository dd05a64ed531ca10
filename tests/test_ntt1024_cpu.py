"""The 32x32 four-step lane programs of csrc/ntt1024_core.cuh (power-of-two twiddles inside the 32-point transforms),
run on the CPU over all 32 lanes, against the oracle's radix-2 NTT."""
import ctypes

import numpy as np
import pytest

from tests import oracle_lib

P = oracle_lib.P
vp = ctypes.c_void_p


@pytest.mark.parametrize("inverse", [0, 1])
def test_ntt1024_lane_programs_match_oracle(inverse):
    hc = oracle_lib.build_hostcheck()
    oracle = oracle_lib.load()
    rng = np.random.default_rng(21 + inverse)
    for trial in range(3):
        x = oracle_lib.rand_field(rng, (1024,))
        if trial == 0:
            x[:4] = [0, 1, P - 1, P - 2]
        nat = np.empty_like(x)
        br = np.empty_like(x)
        hc.hc_ntt1024(x.ctypes.data_as(vp), nat.ctypes.data_as(vp), br.ctypes.data_as(vp), inverse)
        ref = oracle.ntt(x.copy(), inverse=bool(inverse))
        if inverse:  # the oracle scales by 1/n, the lane programs leave that to the post-twiddle table
            ref = np.array([(int(v) * 1024) % P for v in ref], dtype=np.uint64)
        assert (nat == ref).all()
        rev = np.array([int(format(i, "010b")[::-1], 2) for i in range(1024)])
        assert (br == ref[rev]).all()
