/* zkgpu.h -- C ABI of libzkgpu.so: the B200 (sm_100a) replacement for the boojum prover calls made by
 * matter-labs/era-zkevm_test_harness.
 *
 * The reference has no FFI today; its prover boundary is two generic Rust calls into the `boojum` crate:
 *     cs.get_full_setup(worker, fri_lde_factor, merkle_tree_cap_size)        src/prover_utils.rs:185-186 (:452 :619 :749 :886)
 *     cs.prove_from_precomputations::<EXT,TR,H,POW>(proof_config, setup_base, setup, setup_tree, vk, vars_hint,
 *                                                   wits_hint, (), worker)    src/prover_utils.rs:338-348 (:533 :689 :797 :956)
 * and  verifier.verify::<H,TR,POW>((), vk, proof)                             src/prover_utils.rs:351-372 (:546 :702 :810)
 * Each entry point below names the reference interface it replaces.  A Rust maintainer binds them with the
 * `extern "C"` block shown in INTEGRATION.md; this repo's own host side (Python/ctypes, the reference toolchain being
 * absent from the build image) lives in era_zkevm_test_harness_b200/.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero code on failure, with
 * a thread-local message from zkgpu_last_error() (the reference's convention is panic-on-error -- the Rust shim turns
 * non-zero into panic!).  "d_" parameters are device pointers on the context's GPU, "h_" parameters are host pointers.
 * Field elements are canonical Goldilocks u64 (< 2^64 - 2^32 + 1).  Columns are column-major: column c of a batch
 * starts at base + c*stride (stride in elements).
 */
#ifndef ZKGPU_H
#define ZKGPU_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ZKGPU_API __attribute__((visibility("default")))
#else
#define ZKGPU_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkgpu_ctx zkgpu_ctx;     /* one per GPU: device ordinal, stream, twiddle tables, scratch arenas        */
typedef struct zkgpu_setup zkgpu_setup; /* device-resident per-circuit-type setup (replaces SetupStorage + setup tree) */

/* ---- context (replaces `&Worker`, src/prover_utils.rs:50,207: the caller-owned execution resource) ---- */
ZKGPU_API int zkgpu_ctx_create(int device, void* cuda_stream /* cudaStream_t; NULL = the legacy default stream */, zkgpu_ctx** out);
ZKGPU_API void zkgpu_ctx_destroy(zkgpu_ctx* ctx);
ZKGPU_API int zkgpu_ctx_synchronize(zkgpu_ctx* ctx);
ZKGPU_API uint64_t zkgpu_ctx_kernel_launches(const zkgpu_ctx* ctx); /* kernels launched through this context so far */
ZKGPU_API const char* zkgpu_last_error(void);
ZKGPU_API int zkgpu_abi_version(void);

/* ---- primitives on device memory (exported for the parity tests and the NTT roofline metric) ---- */
/* natural-order monomials -> evaluations over coset_shift*<omega_n>, BIT-REVERSED order. d_in == d_out allowed.
 * (boojum: the per-coset FFT inside the LDE of every committed oracle) */
ZKGPU_API int zkgpu_ntt_forward(zkgpu_ctx* ctx, const uint64_t* d_in, size_t in_stride, uint64_t* d_out, size_t out_stride, int log_n,
                      int n_cols, uint64_t coset_shift);
/* natural-order evaluations over <omega_n> -> natural-order monomials. d_tmp: scratch, same shape as d_out
 * (needed when log_n > 11). (boojum: the iFFT turning trace columns into monomial form) */
ZKGPU_API int zkgpu_ntt_inverse(zkgpu_ctx* ctx, const uint64_t* d_in, size_t in_stride, uint64_t* d_out, size_t out_stride, uint64_t* d_tmp,
                      size_t tmp_stride, int log_n, int n_cols);
/* values on the trace domain -> monomials (d_mono, stride mono_stride) and LDE by 2^log_lde (d_lde, column stride
 * lde_stride >= n<<log_lde; coset-major, coset c = 7*omega_{lde*n}^bitrev(c), each coset bit-reversed). */
ZKGPU_API int zkgpu_lde(zkgpu_ctx* ctx, const uint64_t* d_values, size_t val_stride, uint64_t* d_mono, size_t mono_stride, uint64_t* d_lde,
              size_t lde_stride, int log_n, int log_lde, int n_cols);
/* in-place Poseidon2 permutation of n states of 12 elements */
ZKGPU_API int zkgpu_poseidon2_permute(zkgpu_ctx* ctx, uint64_t* d_states, size_t n_states);
/* Merkle tree with cap over column-major data (replaces MerkleTreeWithCap::construct for H = Poseidon2 sponge).
 * leaf i = for each column c: elems_per_leaf consecutive values starting at i*elems_per_leaf.
 * d_tree receives (2*n_leaves - cap_size) digests of 4 u64: levels concatenated from leaf hashes down to the cap. */
ZKGPU_API int zkgpu_merkle_build(zkgpu_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t n_cols, size_t n_leaves,
                       size_t elems_per_leaf, size_t cap_size, uint64_t* d_tree);
/* one un-normalised FRI fold step over the domain shift*<omega_{2^log_dom}> (bit-reversed): Ext2 split storage */
ZKGPU_API int zkgpu_fri_fold(zkgpu_ctx* ctx, const uint64_t* d_in_c0, const uint64_t* d_in_c1, int log_dom, uint64_t shift,
                   const uint64_t challenge[2], uint64_t* d_out_c0, uint64_t* d_out_c1);

/* ---- host-buffer entry point used for the end-to-end numbers: H2D of n_cols trace columns (natural order),
 * iNTT + LDE + Merkle commit on the GPU, D2H of the cap (cap_size*4 u64).  This is round 1 of
 * prove_from_precomputations (witness commitment) and the whole of get_full_setup's commitment work. ---- */
ZKGPU_API int zkgpu_commit_columns_host(zkgpu_ctx* ctx, const uint64_t* h_cols, size_t n_cols, int log_n, int log_lde, size_t cap_size,
                              uint64_t* h_cap_out);

#ifdef __cplusplus
}
#endif
#endif /* ZKGPU_H */
