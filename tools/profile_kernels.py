#!/usr/bin/env python3
"""Small driver for ncu captures (see profiles/README.md): runs either the commitment primitives on W=156 columns x 2^20
(`primitives`: forward coset NTT, inverse NTT, Poseidon2 Merkle build) or one full MainVM-shaped proof (`prove`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from era_zkevm_test_harness_b200 import GpuContext, geometry as G, prover_utils as PU  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "primitives"
    log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    ctx = GpuContext(0)
    n = 1 << log_n
    if what == "primitives":
        W = 156
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randint(0, 2**62, (W, n), dtype=torch.int64, device="cuda", generator=g)
        out = torch.empty_like(x)
        tmp = torch.empty_like(x)
        ctx.ntt_forward(x, log_n, 7, out=out)      # table build + first launches (not the ones captured: use ncu -s)
        ctx.ntt_inverse(x, log_n, out=out, tmp=tmp)
        torch.cuda.synchronize()
        ctx.ntt_forward(x, log_n, 7, out=out)
        ctx.ntt_inverse(x, log_n, out=out, tmp=tmp)
        lde = torch.empty((W, 2 * n), dtype=torch.int64, device="cuda")
        lde[:, :n] = x
        lde[:, n:] = out
        ctx.merkle_build(lde, 2 * n, 1, 16)
        torch.cuda.synchronize()
    else:
        geo = G.mainvm_like_geometry(log_n)
        cfg = G.base_layer_proof_config(log_n)
        wit, setup = PU.synth_trace(geo, seed=7)
        sd = PU.create_setup_data(ctx, geo, cfg, setup)
        d_wit = torch.from_numpy(wit.view(np.int64)).cuda()
        proof = PU.prove_circuit(ctx, sd, d_wit)
        ok, msg = PU.verify_proof(geo, cfg, sd.vk_cap, proof)
        print("verified", ok, msg)
    ctx.close()


if __name__ == "__main__":
    main()
