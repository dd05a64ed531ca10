"""Host-side mirror of the reference's prover boundary `src/prover_utils.rs` over the C ABI of libzkgpu.so.

  create_setup_data   <-> create_base_layer_setup_data / create_recursive_layer_setup_data  (prover_utils.rs:48-197, :383-464)
  prove_circuit       <-> prove_base_layer_circuit / prove_recursion_layer_circuit          (prover_utils.rs:205-349, :466-544)
  verify_proof        <-> verify_base_layer_proof / verify_recursion_layer_proof            (prover_utils.rs:351-372, :546-564)

The reference functions take a circuit instance and run Rust synthesis first; the Rust toolchain is absent from this
image, so the "circuit" here is the materialised trace (witness columns + setup columns) for a `Geometry`
(geometry.py).  Errors raise (the reference panics).
"""
import ctypes

import numpy as np

from . import _lib
from .geometry import Geometry, ProofConfig


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def num_columns(geo):
    lib = _lib.load()
    return dict(witness=lib.zkgpu_num_witness_cols(ctypes.byref(geo)), permuted=lib.zkgpu_num_permuted_cols(ctypes.byref(geo)),
                setup=lib.zkgpu_num_setup_cols(ctypes.byref(geo)), stage2=lib.zkgpu_num_stage2_cols(ctypes.byref(geo)),
                quotient=lib.zkgpu_num_quotient_cols(ctypes.byref(geo)))


def proof_size_u64(geo, cfg):
    n = _lib.load().zkgpu_proof_size_u64(ctypes.byref(geo), ctypes.byref(cfg))
    if n == 0:
        raise _lib.ZkGpuError(_lib.load().zkgpu_last_error().decode())
    return int(n)


def synth_trace(geo, seed=0, pinned=False, witness_seed=None):
    """Synthetic satisfying trace (stands in for Rust synthesis): returns (witness_cols [W,n], setup_cols [S,n]) uint64.
    With `witness_seed`, `seed` fixes the circuit type (setup columns) and `witness_seed` the instance (witness columns)."""
    lib = _lib.load()
    n = 1 << geo.log_n
    if pinned:
        import torch
        wit_t = torch.empty((geo.n_witness, n), dtype=torch.int64).pin_memory()
        set_t = torch.empty((geo.n_setup, n), dtype=torch.int64).pin_memory()
        wit, setup = wit_t.numpy().view(np.uint64), set_t.numpy().view(np.uint64)
    else:
        wit = np.empty((geo.n_witness, n), dtype=np.uint64)
        setup = np.empty((geo.n_setup, n), dtype=np.uint64)
    if witness_seed is None:
        _lib.check(lib.zkgpu_synth_trace(ctypes.byref(geo), seed, _p(wit), _p(setup)))
    else:
        _lib.check(lib.zkgpu_synth_trace_instance(ctypes.byref(geo), seed, witness_seed, _p(wit), _p(setup)))
    return wit, setup


class SetupData:
    """The 7-tuple the reference returns from create_*_setup_data, collapsed to what the GPU prover needs: a device-resident
    setup (monomials, LDE, Merkle tree) + the verification key cap."""

    def __init__(self, ctx, geo, cfg, handle, vk_cap):
        self.ctx, self.geo, self.cfg, self.handle, self.vk_cap = ctx, geo, cfg, handle, vk_cap

    def close(self):
        if self.handle:
            self.ctx.lib.zkgpu_setup_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def create_setup_data(ctx, geo: Geometry, cfg: ProofConfig, setup_cols):
    setup_cols = np.ascontiguousarray(setup_cols, dtype=np.uint64)
    assert setup_cols.shape == (geo.n_setup, 1 << geo.log_n), setup_cols.shape
    h = ctypes.c_void_p()
    vk_cap = np.empty((cfg.cap_size, 4), dtype=np.uint64)
    _lib.check(ctx.lib.zkgpu_setup_create(ctx.h, ctypes.byref(geo), ctypes.byref(cfg), _p(setup_cols), ctypes.byref(h), _p(vk_cap)))
    return SetupData(ctx, geo, cfg, h, vk_cap)


def prove_circuit(ctx, setup: SetupData, witness_cols, proof_out=None):
    """witness_cols: numpy uint64 [W, n] (host) or a torch CUDA int64 tensor [W, n] (already resident)."""
    n_u64 = proof_size_u64(setup.geo, setup.cfg)
    proof = np.empty(n_u64, dtype=np.uint64) if proof_out is None else proof_out
    if isinstance(witness_cols, np.ndarray):
        w = np.ascontiguousarray(witness_cols, dtype=np.uint64)
        assert w.shape == (setup.geo.n_witness, 1 << setup.geo.log_n), w.shape
        _lib.check(ctx.lib.zkgpu_prove(ctx.h, setup.handle, _p(w), _p(proof), n_u64))
    else:
        assert witness_cols.is_cuda and witness_cols.is_contiguous()
        _lib.check(ctx.lib.zkgpu_prove_device(ctx.h, setup.handle, ctypes.c_void_p(witness_cols.data_ptr()), _p(proof), n_u64))
    return proof


def stage_witness(ctx, setup: SetupData, witness_cols, slot):
    """Starts the upload of a (pinned) host witness into staging slot 0/1 and returns immediately; prove_staged(slot) proves it.
    Keep `witness_cols` alive until then."""
    w = witness_cols
    assert isinstance(w, np.ndarray) and w.dtype == np.uint64 and w.flags.c_contiguous
    assert w.shape == (setup.geo.n_witness, 1 << setup.geo.log_n), w.shape
    _lib.check(ctx.lib.zkgpu_witness_stage(ctx.h, setup.handle, _p(w), slot))


def prove_staged(ctx, setup: SetupData, slot, proof_out=None):
    n_u64 = proof_size_u64(setup.geo, setup.cfg)
    proof = np.empty(n_u64, dtype=np.uint64) if proof_out is None else proof_out
    _lib.check(ctx.lib.zkgpu_prove_staged(ctx.h, setup.handle, slot, _p(proof), n_u64))
    return proof


def set_variable_maps(ctx, setup: SetupData, var_maps):
    """var_maps: uint32 [n_perm, n] -- `DenseVariablesCopyHint` of the circuit type (row -> variable index per copy column,
    0xFFFFFFFF = placeholder).  Uploaded once per setup; see prove_from_variables."""
    m = np.ascontiguousarray(var_maps, dtype=np.uint32)
    assert m.shape == (setup.geo.n_perm, 1 << setup.geo.log_n), m.shape
    _lib.check(ctx.lib.zkgpu_setup_set_variable_maps(ctx.h, setup.handle, _p(m)))


def prove_from_variables(ctx, setup: SetupData, variable_values, multiplicities=None, proof_out=None):
    """The reference's hand-off (prove_from_precomputations(.., vars_hint, wits_hint, ..), src/prover_utils.rs:338-348): ship the
    assembly's variable values (+ lookup multiplicities), gather the trace columns on the GPU, prove."""
    n_u64 = proof_size_u64(setup.geo, setup.cfg)
    proof = np.empty(n_u64, dtype=np.uint64) if proof_out is None else proof_out
    v = np.ascontiguousarray(variable_values, dtype=np.uint64)
    mp = None if multiplicities is None else np.ascontiguousarray(multiplicities, dtype=np.uint64)
    _lib.check(ctx.lib.zkgpu_prove_from_variables(ctx.h, setup.handle, _p(v), v.size, _p(mp) if mp is not None else ctypes.c_void_p(0),
                                                  _p(proof), n_u64))
    return proof


def verify_proof(geo: Geometry, cfg: ProofConfig, vk_cap, proof):
    """-> (bool, message).  CPU only, like the reference verifier."""
    lib = _lib.load()
    vk_cap = np.ascontiguousarray(vk_cap, dtype=np.uint64)
    proof = np.ascontiguousarray(proof, dtype=np.uint64)
    rc = lib.zkgpu_verify(ctypes.byref(geo), ctypes.byref(cfg), _p(vk_cap), _p(proof), proof.size)
    if rc == 0:
        return True, ""
    msg = lib.zkgpu_last_error().decode()
    if rc == 1:
        return False, msg
    raise _lib.ZkGpuError(msg)
