"""Micro-benchmarks of the primitive kernels on cuda:0 (CUDA events, warm-up, inputs larger than L2)."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from era_zkevm_test_harness_b200 import GpuContext  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    ctx = GpuContext(0)
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n_cols = int(sys.argv[2]) if len(sys.argv) > 2 else 156
    n = 1 << log_n
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randint(0, 2**62, (n_cols, n), dtype=torch.int64, device="cuda", generator=g)
    out = torch.empty_like(x); tmp = torch.empty_like(x)
    res = {}
    t = timeit(lambda: ctx.ntt_forward(x, log_n, 7, out=out))
    res["ntt_forward_coset_ms"] = t; res["ntt_forward_GBps"] = 16 * n * n_cols / t / 1e6
    t = timeit(lambda: ctx.ntt_forward(x, log_n, 1, out=out))
    res["ntt_forward_ms"] = t; res["ntt_forward_nocoset_GBps"] = 16 * n * n_cols / t / 1e6
    t = timeit(lambda: ctx.ntt_inverse(x, log_n, out=out, tmp=tmp))
    res["ntt_inverse_ms"] = t; res["ntt_inverse_GBps"] = 16 * n * n_cols / t / 1e6
    del tmp
    lde = torch.randint(0, 2**62, (n_cols, 2 * n), dtype=torch.int64, device="cuda", generator=g)
    tree = torch.empty((4 * n - 16, 4), dtype=torch.int64, device="cuda")
    t = timeit(lambda: ctx.merkle_build(lde, 2 * n, 1, 16, tree=tree), iters=3, warm=1)
    perms = 2 * n * ((n_cols + 7) // 8) + 2 * n
    res["merkle_ms"] = t; res["poseidon2_perms_per_s"] = perms / t * 1e3; res["merkle_read_GBps"] = 8 * 2 * n * n_cols / t / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
