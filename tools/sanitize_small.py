#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): one 2^20 NTT round trip on 2 columns (the shared-memory
warp-exchange kernels), generic-size NTTs, a Merkle build, and one small proof checked against the oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from era_zkevm_test_harness_b200 import GpuContext, geometry as G, prover_utils as PU  # noqa: E402
from era_zkevm_test_harness_b200.context import to_device_u64, to_numpy_u64  # noqa: E402
from tests import oracle_lib  # noqa: E402

oracle = oracle_lib.load()
ctx = GpuContext(0)
rng = np.random.default_rng(3)
vals = oracle_lib.rand_field(rng, (2, 1 << 20))
d = to_device_u64(vals, ctx.device)
mono = ctx.ntt_inverse(d, 20)
ev = ctx.ntt_forward(mono, 20, 7)
assert (to_numpy_u64(ev)[0] == oracle.coset_evals_bitrev(to_numpy_u64(mono)[0], 7)).all()
for log_n in (5, 11, 13):
    a = oracle_lib.rand_field(rng, (3, 1 << log_n))
    assert (to_numpy_u64(ctx.ntt_forward(to_device_u64(a, ctx.device), log_n, 7)) == oracle.coset_evals_bitrev(a, 7)).all()
    assert (to_numpy_u64(ctx.ntt_inverse(to_device_u64(a, ctx.device), log_n)) == oracle.ntt(a, inverse=True)).all()
geo = G.small_test_geometry(8, 16, True)
cfg = G.make_proof_config(8, 2, 4, security_level=12)
wit, setup = PU.synth_trace(geo, seed=11)
sd = PU.create_setup_data(ctx, geo, cfg, setup)
proof = PU.prove_circuit(ctx, sd, wit)
assert (proof == oracle.prove(geo, cfg, wit, setup)).all()
g2 = G.mainvm_like_geometry(9)
c2 = G.make_proof_config(9, 2, 16, security_level=8)
w2, s2 = PU.synth_trace(g2, seed=12)
sd2 = PU.create_setup_data(ctx, g2, c2, s2)
assert (PU.prove_circuit(ctx, sd2, w2) == oracle.prove(g2, c2, w2, s2)).all()
# compression mode 2 (plain witness columns, LDE 512: all cosets in one launch per NTT pass) and mode 4 (cap 256, MatMul gates)
import json  # noqa: E402
fx = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "vk_shapes.json")))
for key, cgeo, ccfg, _ in G.compression_geometries_from_fixture(fx):
    if key not in ("compression_2", "compression_4"):
        continue
    g3 = cgeo.scaled(5)
    c3 = G.make_proof_config(5, 1 << ccfg.log_lde, ccfg.cap_size, security_level=2 * ccfg.log_lde)
    w3, s3 = PU.synth_trace(g3, seed=13)
    sd3 = PU.create_setup_data(ctx, g3, c3, s3)
    assert (PU.prove_circuit(ctx, sd3, w3) == oracle.prove(g3, c3, w3, s3)).all()
    sd3.close()
# staged upload: both slots, reuse
wa, _ = PU.synth_trace(g2, seed=12, pinned=True)
PU.stage_witness(ctx, sd2, wa, 0)
PU.stage_witness(ctx, sd2, wa, 1)
pa = PU.prove_staged(ctx, sd2, 0).copy()
PU.stage_witness(ctx, sd2, wa, 0)
assert (PU.prove_staged(ctx, sd2, 1) == pa).all() and (PU.prove_staged(ctx, sd2, 0) == pa).all()
torch.cuda.synchronize()
print("sanitize_small: ok")
