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


def synth_trace(geo, seed=0, pinned=False, witness_seed=None, out_witness=None):
    """Synthetic satisfying trace (stands in for Rust synthesis): returns (witness_cols [W,n], setup_cols [S,n]) uint64.
    With `witness_seed`, `seed` fixes the circuit type (setup columns) and `witness_seed` the instance (witness columns).
    `out_witness`: a caller-owned (e.g. pinned, PinnedPool) uint64 buffer of at least W*n words to generate the witness into."""
    lib = _lib.load()
    n = 1 << geo.log_n
    if out_witness is not None:
        assert out_witness.dtype == np.uint64 and out_witness.flags.c_contiguous and out_witness.size >= geo.n_witness * n
        wit = out_witness.reshape(-1)[: geo.n_witness * n].reshape(geo.n_witness, n)
        setup = np.empty((geo.n_setup, n), dtype=np.uint64)
    elif pinned:
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


class PinnedPool:
    """A few page-locked host buffers (zkgpu_host_alloc = cudaHostAlloc) of one size, handed out and taken back: witness
    columns generated into them upload at full PCIe speed and asynchronously (zkgpu_witness_stage)."""

    def __init__(self, n_buffers, n_u64):
        import queue
        self.lib = _lib.load()
        self.n_u64 = int(n_u64)
        self.ptrs, self.free = [], queue.Queue()
        for _ in range(n_buffers):
            p = self.lib.zkgpu_host_alloc(self.n_u64 * 8)
            if not p:
                raise _lib.ZkGpuError("zkgpu_host_alloc failed: " + self.lib.zkgpu_last_error().decode())
            self.ptrs.append(p)
            self.free.put(np.ctypeslib.as_array((ctypes.c_uint64 * self.n_u64).from_address(p)))

    def take(self):
        return self.free.get()      # blocks until a buffer is handed back

    def give(self, buf):
        self.free.put(buf)

    def close(self):
        for p in self.ptrs:
            self.lib.zkgpu_host_free(p)
        self.ptrs = []


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


def _proof_buffer(setup, proof_out):
    """-> (buffer, capacity in u64).  A caller-supplied buffer must be a C-contiguous uint64 array able to hold the proof: the
    library writes `zkgpu_proof_size_u64` words into it and is told the buffer's real size."""
    n_u64 = proof_size_u64(setup.geo, setup.cfg)
    if proof_out is None:
        return np.empty(n_u64, dtype=np.uint64), n_u64
    if not (isinstance(proof_out, np.ndarray) and proof_out.dtype == np.uint64 and proof_out.flags.c_contiguous and proof_out.ndim == 1):
        raise ValueError("proof_out must be a one-dimensional C-contiguous numpy uint64 array")
    if proof_out.size < n_u64:
        raise ValueError(f"proof_out holds {proof_out.size} u64, the proof needs {n_u64}")
    return proof_out, int(proof_out.size)


def prove_circuit(ctx, setup: SetupData, witness_cols, proof_out=None):
    """witness_cols: numpy uint64 [W, n] (host) or a torch CUDA int64 tensor [W, n] (already resident)."""
    proof, n_u64 = _proof_buffer(setup, proof_out)
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
    proof, n_u64 = _proof_buffer(setup, proof_out)
    _lib.check(ctx.lib.zkgpu_prove_staged(ctx.h, setup.handle, slot, _p(proof), n_u64))
    return proof


def set_variable_maps(ctx, setup: SetupData, var_maps):
    """var_maps: uint32 [n_perm, n] -- `DenseVariablesCopyHint` of the circuit type (row -> variable index per copy column,
    0xFFFFFFFF = placeholder).  Uploaded once per setup; see prove_from_variables."""
    m = np.ascontiguousarray(var_maps, dtype=np.uint32)
    assert m.shape == (setup.geo.n_perm, 1 << setup.geo.log_n), m.shape
    _lib.check(ctx.lib.zkgpu_setup_set_variable_maps(ctx.h, setup.handle, _p(m)))


def set_witness_maps(ctx, setup: SetupData, wit_maps):
    """wit_maps: uint32 [n_witness_plain, n] -- `DenseWitnessCopyHint` of the circuit type (row -> index into the assembly's
    witness values per plain witness column; compression modes 1-3).  Uploaded once per setup; see prove_from_hints."""
    m = np.ascontiguousarray(wit_maps, dtype=np.uint32)
    assert m.shape == (setup.geo.n_witness_plain, 1 << setup.geo.log_n), m.shape
    _lib.check(ctx.lib.zkgpu_setup_set_witness_maps(ctx.h, setup.handle, _p(m)))


def prove_from_hints(ctx, setup: SetupData, variable_values, witness_values=None, multiplicities=None, proof_out=None):
    """The reference's hand-off (prove_from_precomputations(.., vars_hint, wits_hint, ..), src/prover_utils.rs:338-348): ship the
    assembly's variable values, witness values (circuits with plain witness columns) and lookup multiplicities; the trace
    columns are gathered on the GPU through the resident maps, then proven."""
    proof, n_u64 = _proof_buffer(setup, proof_out)
    v = np.ascontiguousarray(variable_values, dtype=np.uint64)
    wv = None if witness_values is None else np.ascontiguousarray(witness_values, dtype=np.uint64)
    mp = None if multiplicities is None else np.ascontiguousarray(multiplicities, dtype=np.uint64)
    null = ctypes.c_void_p(0)
    _lib.check(ctx.lib.zkgpu_prove_from_hints(ctx.h, setup.handle, _p(v), v.size, _p(wv) if wv is not None else null,
                                              0 if wv is None else wv.size, _p(mp) if mp is not None else null, _p(proof), n_u64))
    return proof


def prove_from_variables(ctx, setup: SetupData, variable_values, multiplicities=None, proof_out=None):
    """prove_from_hints for circuits without plain witness columns (every base- and recursion-layer circuit)."""
    proof, n_u64 = _proof_buffer(setup, proof_out)
    v = np.ascontiguousarray(variable_values, dtype=np.uint64)
    mp = None if multiplicities is None else np.ascontiguousarray(multiplicities, dtype=np.uint64)
    _lib.check(ctx.lib.zkgpu_prove_from_variables(ctx.h, setup.handle, _p(v), v.size, _p(mp) if mp is not None else ctypes.c_void_p(0),
                                                  _p(proof), n_u64))
    return proof


def verify_proof(geo: Geometry, cfg: ProofConfig, vk_cap, proof, skip_quotient_identity=False):
    """-> (bool, message).  CPU only, like the reference verifier.  skip_quotient_identity: diagnostic (zkgpu_verify_ex)."""
    lib = _lib.load()
    vk_cap = np.ascontiguousarray(vk_cap, dtype=np.uint64)
    proof = np.ascontiguousarray(proof, dtype=np.uint64)
    rc = lib.zkgpu_verify_ex(ctypes.byref(geo), ctypes.byref(cfg), _p(vk_cap), _p(proof), proof.size, 1 if skip_quotient_identity else 0)
    if rc == 0:
        return True, ""
    msg = lib.zkgpu_last_error().decode()
    if rc == 1:
        return False, msg
    raise _lib.ZkGpuError(msg)
