"""GpuContext: Python handle over zkgpu_ctx.  PyTorch is used only for device memory and streams (plumbing)."""
import ctypes

import numpy as np
import torch

from . import _lib

P = (1 << 64) - (1 << 32) + 1


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def to_device_u64(arr, device):
    """numpy uint64 array -> int64 CUDA tensor holding the same bits."""
    a = np.ascontiguousarray(arr, dtype=np.uint64)
    return torch.from_numpy(a.view(np.int64)).to(device)


def to_numpy_u64(t):
    return t.detach().cpu().numpy().view(np.uint64)


class GpuContext:
    """One per GPU; plays the role of the caller-owned `&Worker` of the reference (src/prover_utils.rs:50)."""

    def __init__(self, device=0, use_torch_stream=True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.ZkGpuError("CUDA device required: the prover has no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream if use_torch_stream else 0
        h = ctypes.c_void_p()
        _lib.check(self.lib.zkgpu_ctx_create(device, ctypes.c_void_p(stream), ctypes.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.zkgpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _lib.check(self.lib.zkgpu_ctx_synchronize(self.h))

    @property
    def kernel_launches(self):
        return int(self.lib.zkgpu_ctx_kernel_launches(self.h))

    # ---- primitives on torch int64 tensors of shape [n_cols, n] (bits are u64 field elements) ----
    def ntt_forward(self, x, log_n, coset_shift=1, out=None):
        n_cols = x.shape[0]
        out = torch.empty_like(x) if out is None else out
        _lib.check(self.lib.zkgpu_ntt_forward(self.h, _ptr(x), x.stride(0), _ptr(out), out.stride(0), log_n, n_cols, coset_shift))
        return out

    def ntt_inverse(self, x, log_n, out=None, tmp=None):
        n_cols = x.shape[0]
        out = torch.empty_like(x) if out is None else out
        tmp = torch.empty_like(x) if tmp is None else tmp
        _lib.check(self.lib.zkgpu_ntt_inverse(self.h, _ptr(x), x.stride(0), _ptr(out), out.stride(0), _ptr(tmp), tmp.stride(0), log_n, n_cols))
        return out

    def lde(self, vals, log_n, log_lde, mono=None, lde=None):
        n_cols, n = vals.shape
        mono = torch.empty_like(vals) if mono is None else mono
        lde = torch.empty((n_cols, n << log_lde), dtype=torch.int64, device=vals.device) if lde is None else lde
        _lib.check(self.lib.zkgpu_lde(self.h, _ptr(vals), vals.stride(0), _ptr(mono), mono.stride(0), _ptr(lde), lde.stride(0), log_n, log_lde, n_cols))
        return mono, lde

    def poseidon2_permute(self, states):
        _lib.check(self.lib.zkgpu_poseidon2_permute(self.h, _ptr(states), states.shape[0]))
        return states

    def merkle_build(self, cols, n_leaves, elems_per_leaf, cap_size, tree=None):
        n_cols = cols.shape[0]
        n_dig = 2 * n_leaves - cap_size
        tree = torch.empty((n_dig, 4), dtype=torch.int64, device=cols.device) if tree is None else tree
        _lib.check(self.lib.zkgpu_merkle_build(self.h, _ptr(cols), cols.stride(0), n_cols, n_leaves, elems_per_leaf, cap_size, _ptr(tree)))
        return tree

    def fri_fold(self, c0, c1, log_dom, shift, challenge):
        half = 1 << (log_dom - 1)
        o0 = torch.empty(half, dtype=torch.int64, device=c0.device)
        o1 = torch.empty(half, dtype=torch.int64, device=c0.device)
        ch = (ctypes.c_uint64 * 2)(int(challenge[0]), int(challenge[1]))
        _lib.check(self.lib.zkgpu_fri_fold(self.h, _ptr(c0), _ptr(c1), log_dom, shift, ctypes.byref(ch), _ptr(o0), _ptr(o1)))
        return o0, o1

    def commit_columns_host(self, h_cols, log_n, log_lde, cap_size, h_cap=None):
        """h_cols: pinned (or plain) CPU int64 tensor [n_cols, n]; returns CPU tensor [cap_size, 4]."""
        n_cols = h_cols.shape[0]
        h_cap = torch.empty((cap_size, 4), dtype=torch.int64).pin_memory() if h_cap is None else h_cap
        _lib.check(self.lib.zkgpu_commit_columns_host(self.h, ctypes.c_void_p(h_cols.data_ptr()), n_cols, log_n, log_lde, cap_size,
                                                      ctypes.c_void_p(h_cap.data_ptr())))
        return h_cap
