"""ctypes binding of libzkgpu.so (include/zkgpu.h).  Fails loudly: there is no CPU fallback on the product path."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKGPU_LIB") or os.path.join(_HERE, "libzkgpu.so")   # ZKGPU_LIB: a build variant for A/B measurements

u64 = ctypes.c_uint64
u64p = ctypes.c_void_p  # device or host pointer passed as an integer address
sz = ctypes.c_size_t
ci = ctypes.c_int

# every symbol include/zkgpu.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "zkgpu_abi_version": (ci, []),
    "zkgpu_last_error": (ctypes.c_char_p, []),
    "zkgpu_ctx_create": (ci, [ci, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "zkgpu_ctx_destroy": (None, [ctypes.c_void_p]),
    "zkgpu_ctx_synchronize": (ci, [ctypes.c_void_p]),
    "zkgpu_ctx_kernel_launches": (u64, [ctypes.c_void_p]),
    "zkgpu_ntt_forward": (ci, [ctypes.c_void_p, u64p, sz, u64p, sz, ci, ci, u64]),
    "zkgpu_ntt_inverse": (ci, [ctypes.c_void_p, u64p, sz, u64p, sz, u64p, sz, ci, ci]),
    "zkgpu_lde": (ci, [ctypes.c_void_p, u64p, sz, u64p, sz, u64p, sz, ci, ci, ci]),
    "zkgpu_poseidon2_permute": (ci, [ctypes.c_void_p, u64p, sz]),
    "zkgpu_merkle_build": (ci, [ctypes.c_void_p, u64p, sz, sz, sz, sz, sz, u64p]),
    "zkgpu_fri_fold": (ci, [ctypes.c_void_p, u64p, u64p, ci, u64, ctypes.POINTER(u64 * 2), u64p, u64p]),
    "zkgpu_commit_columns_host": (ci, [ctypes.c_void_p, u64p, sz, ci, ci, sz, u64p]),
    "zkgpu_num_witness_cols": (ctypes.c_uint32, [ctypes.c_void_p]),
    "zkgpu_num_permuted_cols": (ctypes.c_uint32, [ctypes.c_void_p]),
    "zkgpu_num_setup_cols": (ctypes.c_uint32, [ctypes.c_void_p]),
    "zkgpu_num_stage2_cols": (ctypes.c_uint32, [ctypes.c_void_p]),
    "zkgpu_num_quotient_cols": (ctypes.c_uint32, [ctypes.c_void_p]),
    "zkgpu_proof_size_u64": (sz, [ctypes.c_void_p, ctypes.c_void_p]),
    "zkgpu_setup_create": (ci, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, u64p, ctypes.POINTER(ctypes.c_void_p), u64p]),
    "zkgpu_setup_destroy": (None, [ctypes.c_void_p]),
    "zkgpu_prove": (ci, [ctypes.c_void_p, ctypes.c_void_p, u64p, u64p, sz]),
    "zkgpu_prove_device": (ci, [ctypes.c_void_p, ctypes.c_void_p, u64p, u64p, sz]),
    "zkgpu_host_alloc": (ctypes.c_void_p, [sz]),
    "zkgpu_host_free": (None, [ctypes.c_void_p]),
    "zkgpu_witness_stage": (ci, [ctypes.c_void_p, ctypes.c_void_p, u64p, ci]),
    "zkgpu_prove_staged": (ci, [ctypes.c_void_p, ctypes.c_void_p, ci, u64p, sz]),
    "zkgpu_setup_set_variable_maps": (ci, [ctypes.c_void_p, ctypes.c_void_p, u64p]),
    "zkgpu_prove_from_variables": (ci, [ctypes.c_void_p, ctypes.c_void_p, u64p, sz, u64p, u64p, sz]),
    "zkgpu_setup_set_witness_maps": (ci, [ctypes.c_void_p, ctypes.c_void_p, u64p]),
    "zkgpu_prove_from_hints": (ci, [ctypes.c_void_p, ctypes.c_void_p, u64p, sz, u64p, sz, u64p, u64p, sz]),
    "zkgpu_verify": (ci, [ctypes.c_void_p, ctypes.c_void_p, u64p, u64p, sz]),
    "zkgpu_verify_ex": (ci, [ctypes.c_void_p, ctypes.c_void_p, u64p, u64p, sz, ctypes.c_uint32]),
    "zkgpu_synth_trace": (ci, [ctypes.c_void_p, u64, u64p, u64p]),
    "zkgpu_synth_trace_instance": (ci, [ctypes.c_void_p, u64, u64, u64p, u64p]),
}

_lib = None


class ZkGpuError(RuntimeError):
    pass


def load():
    """Loads libzkgpu.so; raises if it has not been built (python -m era_zkevm_test_harness_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZkGpuError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                         "(the GPU prover has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise ZkGpuError(f"libzkgpu error {rc}: {load().zkgpu_last_error().decode()}")
