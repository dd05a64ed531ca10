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


/* ==================================================================================================================
 * Circuit description, setup, prove, verify  (the drop-in for get_full_setup / prove_from_precomputations / verify)
 * ================================================================================================================== */

/* Gate kinds.  The VK (setup/base_layer/vk_N.json: fixed_parameters.selectors_placement) stores per gate only
 * (gate_idx, num_constants, degree) and the selector path; the gate polynomials are boojum source.  The kinds below
 * restate the gate set named by the reference's configure_builder calls (circuit_definitions/src/circuit_definitions/
 * base_layer/vm_main.rs:55-117 and siblings); SURVEY.md section 8a maps gate_idx -> kind for every VK. */
enum {
    ZKGPU_GATE_NOP = 0,               /* NopGate / PublicInputGate: no relation                                   */
    ZKGPU_GATE_CONSTANTS_ALLOCATOR,   /* v_i - k_i, one instance per gate constant                       (deg 1) */
    ZKGPU_GATE_FMA,                   /* k0*a*b + k1*c - d                     FmaGateInBaseFieldWithoutConstant  */
    ZKGPU_GATE_REDUCTION4,            /* sum_i k_i*a_i - out                                   ReductionGate<_,4> */
    ZKGPU_GATE_SELECTION,             /* s*a + (1-s)*b - out                                        SelectionGate */
    ZKGPU_GATE_PARALLEL_SELECTION4,   /* 4x selection with a shared selector                ParallelSelectionGate */
    ZKGPU_GATE_ZERO_CHECK,            /* x*inv - (1-z), x*z                                         ZeroCheckGate */
    ZKGPU_GATE_UINTX_ADD,             /* a + b + cin - c - k*cout, cout^2 - cout                     UIntXAddGate */
    ZKGPU_GATE_DOT_PRODUCT4,          /* sum_i a_i*b_i - out                                   DotProductGate<4> */
    ZKGPU_GATE_U8X4_FMA,              /* byte-decomposed a*b + c + carry = low + 2^32*high               U8x4FMAGate */
    ZKGPU_GATE_POSEIDON2_FLATTENED,   /* one whole Poseidon2 permutation per row, 118 relations of degree 7       */
    ZKGPU_GATE_FMA_EXT,               /* FMA over Ext2 (4 constants)                FmaGateInExtensionWithoutConstant */
    ZKGPU_GATE_U32_TRI_ADD_CARRY,     /* a + b + c - out - 2^32*carry (carry range-checked as a lookup chunk)  U32TriAddCarryAsChunkGate */
    ZKGPU_GATE_BOUNDED_BOOLEAN,       /* x^2 - x on the first min(10, n_copy) copy columns (deg 2)         BoundedBooleanConstraintGate */
    ZKGPU_GATE_MATMUL12_EXTERNAL,     /* out_i - sum_j M_E[i][j] in_j, 12 relations (deg 1)  MatrixMultiplicationGate<12, Poseidon2 external> */
    ZKGPU_GATE_MATMUL12_INNER,        /* out_i - sum_j M_I[i][j] in_j, 12 relations (deg 1)  MatrixMultiplicationGate<12, Poseidon2 inner>    */
    ZKGPU_GATE_NONLINEARITY7,         /* y - (x + k)^7, one gate constant (deg 7)                         SimpleNonlinearityGate<7> */
    ZKGPU_GATE_CONDITIONAL_SWAP4,     /* s*(b_i-a_i)+a_i-ra_i, s*(a_i-b_i)+b_i-rb_i, i<4 (deg 2)             ConditionalSwapGate<4> */
    ZKGPU_GATE_ZERO_CHECK_WITNESS,    /* ZeroCheckGate with the inverse in a plain witness column (use_witness = true)        */
    ZKGPU_GATE_BOOLEAN_ALL,           /* x^2 - x on EVERY copy column (deg 2): BooleanConstraintGate on general-purpose columns (eip4844/mod.rs:95-98) */
    ZKGPU_GATE_KINDS
};

#define ZKGPU_MAX_GATES 24
#define ZKGPU_MAX_PUBLIC_INPUTS 8
#define ZKGPU_MAX_FRI_ORACLES 16

typedef struct {
    uint32_t kind;       /* ZKGPU_GATE_* */
    uint32_t n_consts;   /* gate constants, read from constant columns [path_len, path_len + n_consts) */
    uint32_t path_len;   /* selector = prod over i < path_len of (c_i if bit i of path_bits else 1 - c_i) */
    uint32_t path_bits;
} zkgpu_gate;

/* Mirrors VerificationKey.fixed_parameters (boojum VerificationKeyCircuitGeometry) as read from setup/ vk_N.json. */
typedef struct {
    uint32_t log_n;             /* domain_size = 2^log_n                                                           */
    uint32_t n_copy;            /* parameters.num_columns_under_copy_permutation                                    */
    uint32_t n_witness_plain;   /* parameters.num_witness_columns: witness columns NOT under the copy permutation
                                 * (compression modes 1-3: 78 / 74 / 62).  Gate cells [n_copy, n_copy + n_witness_plain)
                                 * live there: the flattened Poseidon2 gate's 130 cells span both kinds of column.     */
    uint32_t n_const_cols;      /* num_constant_columns + extra_constant_polys_for_selectors (+1 table-id column)   */
    uint32_t lookup_width;      /* lookup_parameters width (0 = no lookup)                                          */
    uint32_t lookup_reps;       /* num_repetitions                                                                  */
    uint32_t table_id_col;      /* table_ids_column_idxes[0]: constant column carrying the row's table id           */
    uint32_t has_boolean_col;   /* one specialised boolean-constrained column (the Boolean gate's own column)       */
    uint32_t quotient_degree;   /* quotient_degree (8)                                                              */
    uint32_t table_len;         /* total_tables_len                                                                 */
    uint32_t n_public_inputs;
    uint32_t pi_col[ZKGPU_MAX_PUBLIC_INPUTS], pi_row[ZKGPU_MAX_PUBLIC_INPUTS]; /* public_inputs_locations */
    uint32_t n_gates;
    zkgpu_gate gates[ZKGPU_MAX_GATES];
} zkgpu_geometry;

/* Mirrors ProofConfig (circuit_definitions/src/lib.rs:29-57) with the derived query count and folding schedule. */
typedef struct {
    uint32_t log_lde;           /* log2(fri_lde_factor) */
    uint32_t cap_size;          /* merkle_tree_cap_size */
    uint32_t n_queries;         /* ceil(security_level / log2(fri_lde_factor)) */
    uint32_t pow_bits;          /* 0 everywhere in the reference (NoPow) */
    uint32_t n_fri_oracles;     /* base oracle + intermediates */
    uint32_t fri_schedule[ZKGPU_MAX_FRI_ORACLES]; /* log2 fold factor per oracle, e.g. 3,3,3,3,3,2 */
} zkgpu_proof_config;

/* derived column counts (identical formulas in prover, verifier and oracle) */
ZKGPU_API uint32_t zkgpu_num_witness_cols(const zkgpu_geometry* g);  /* n_copy + boolean + width*reps + plain + multiplicity */
ZKGPU_API uint32_t zkgpu_num_permuted_cols(const zkgpu_geometry* g); /* witness columns under copy permutation     */
ZKGPU_API uint32_t zkgpu_num_setup_cols(const zkgpu_geometry* g);    /* sigmas + constants + (width+1) table cols   */
ZKGPU_API uint32_t zkgpu_num_stage2_cols(const zkgpu_geometry* g);   /* 2 * (ceil(perm/qdeg) + reps + (reps?1:0))   */
ZKGPU_API uint32_t zkgpu_num_quotient_cols(const zkgpu_geometry* g); /* 2 * quotient_degree                         */
ZKGPU_API size_t zkgpu_proof_size_u64(const zkgpu_geometry* g, const zkgpu_proof_config* cfg); /* proof buffer length */

/* get_full_setup (src/prover_utils.rs:185-186): uploads the setup columns (column-major, natural row order:
 * sigma columns, constant columns, table columns -- zkgpu_num_setup_cols() x 2^log_n), keeps monomials + LDE +
 * Merkle tree resident on the GPU, returns the VK's setup_merkle_tree_cap (cap_size*4 u64). */
ZKGPU_API int zkgpu_setup_create(zkgpu_ctx* ctx, const zkgpu_geometry* g, const zkgpu_proof_config* cfg, const uint64_t* h_setup_cols,
                                 zkgpu_setup** out, uint64_t* h_vk_cap_out);
ZKGPU_API void zkgpu_setup_destroy(zkgpu_setup* s);

/* prove_from_precomputations (src/prover_utils.rs:338-348): h_witness_cols = zkgpu_num_witness_cols() x 2^log_n
 * (column-major, natural row order, the materialised trace: copy columns, boolean column, lookup columns,
 * multiplicities).  Writes the proof (layout: DESIGN.md "Proof buffer") into h_proof_out. */
ZKGPU_API int zkgpu_prove(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* h_witness_cols, uint64_t* h_proof_out,
                          size_t proof_capacity_u64);
/* same, witness already resident on the device (d_witness_cols, column stride = 2^log_n) */
ZKGPU_API int zkgpu_prove_device(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* d_witness_cols, uint64_t* h_proof_out,
                                 size_t proof_capacity_u64);

/* Page-locked host memory for witness / setup columns (cudaHostAlloc): uploads from pinned buffers run at PCIe speed and
 * asynchronously; from a pageable buffer (a plain Vec<u64>) the copy is staged by the driver and a base-layer proof takes
 * 240-270 ms instead of 152 ms.  Lets the Rust side allocate its column buffers without linking the CUDA runtime itself. */
ZKGPU_API void* zkgpu_host_alloc(size_t bytes);
ZKGPU_API void zkgpu_host_free(void* p);

/* Witness upload ahead of the proof: `basic_test` proves its circuits in a loop (src/tests/complex_tests/mod.rs:316-410), so the
 * witness of circuit k+1 is known while circuit k is being proven.  zkgpu_witness_stage starts the host->device copy of a
 * witness into one of two staging slots on the context's copy stream and returns at once (h_witness_cols should be pinned
 * and must stay valid until the matching zkgpu_prove_staged returns); zkgpu_prove_staged proves the witness of a slot.
 * Staging slot k+1 before proving slot k hides the PCIe transfer behind the previous proof. */
ZKGPU_API int zkgpu_witness_stage(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* h_witness_cols, int slot);
ZKGPU_API int zkgpu_prove_staged(zkgpu_ctx* ctx, const zkgpu_setup* s, int slot, uint64_t* h_proof_out, size_t proof_capacity_u64);

/* Witness hand-off as the reference does it: boojum's prove_from_precomputations takes `vars_hint: &DenseVariablesCopyHint`
 * (src/prover_utils.rs:346) -- per copy-permutation column a dense map row -> variable index -- and materialises the trace
 * columns itself from the assembly's variable values.  The maps are per circuit TYPE (part of the setup 7-tuple,
 * src/prover_utils.rs:186-196), so they are uploaded once and stay resident; a proof then ships only the variable values.
 *   h_var_maps: zkgpu_num_permuted_cols() x 2^log_n u32, column-major; entry = index into the variable-value array,
 *               ZKGPU_VAR_PLACEHOLDER = unassigned cell (reads as 0, boojum's placeholder variable). */
#define ZKGPU_VAR_PLACEHOLDER 0xFFFFFFFFu
ZKGPU_API int zkgpu_setup_set_variable_maps(zkgpu_ctx* ctx, zkgpu_setup* s, const uint32_t* h_var_maps);
/* h_variable_values: n_vars canonical field elements; h_multiplicities: 2^log_n lookup multiplicities (NULL when the
 * circuit has no lookup) -- boojum keeps them in the assembly next to the variable values.  GPU: gather -> columns -> prove. */
ZKGPU_API int zkgpu_prove_from_variables(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* h_variable_values, size_t n_vars,
                                         const uint64_t* h_multiplicities, uint64_t* h_proof_out, size_t proof_capacity_u64);
/* The second hint of the same call: `wits_hint: &DenseWitnessCopyHint` (src/prover_utils.rs:347) -- per plain witness column (the
 * columns NOT under the copy permutation: compression modes 1-3 have 78 / 74 / 62 of them) a dense map row -> index into the
 * assembly's witness-value array.  h_wit_maps: n_witness_plain x 2^log_n u32, column-major, ZKGPU_VAR_PLACEHOLDER = 0.
 * Both setters reject nothing by themselves; the largest index of each map is remembered and checked against n_vars / n_wits
 * by the prove calls (an out-of-range index is an error, not a silent zero). */
ZKGPU_API int zkgpu_setup_set_witness_maps(zkgpu_ctx* ctx, zkgpu_setup* s, const uint32_t* h_wit_maps);
/* zkgpu_prove_from_variables plus the witness values behind the plain witness columns (h_witness_values may be NULL iff the
 * circuit has no plain witness columns). */
ZKGPU_API int zkgpu_prove_from_hints(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* h_variable_values, size_t n_vars,
                                     const uint64_t* h_witness_values, size_t n_wits, const uint64_t* h_multiplicities,
                                     uint64_t* h_proof_out, size_t proof_capacity_u64);

/* Verifier::verify (src/prover_utils.rs:351-372): CPU only, as in the reference. Returns 0 iff the proof is valid,
 * 1 if invalid (message says which check failed), >1 on malformed input. */
ZKGPU_API int zkgpu_verify(const zkgpu_geometry* g, const zkgpu_proof_config* cfg, const uint64_t* vk_cap, const uint64_t* proof,
                           size_t proof_len_u64);
/* Diagnostic form of zkgpu_verify: `flags` switches single checks off so the remaining ones can be run on proofs made by the
 * reference itself (tools/golden_verify.py, tests/test_golden_verify_cpu.py).  flags = 0 is zkgpu_verify. */
#define ZKGPU_VERIFY_SKIP_QUOTIENT_IDENTITY 1u
ZKGPU_API int zkgpu_verify_ex(const zkgpu_geometry* g, const zkgpu_proof_config* cfg, const uint64_t* vk_cap, const uint64_t* proof,
                              size_t proof_len_u64, uint32_t flags);

/* Synthetic satisfying trace for a geometry (stands in for the reference's Rust synthesis, which cannot run in this
 * image): fills witness (W x n) and setup (S x n) columns deterministically from `seed`.  Host only. */
ZKGPU_API int zkgpu_synth_trace(const zkgpu_geometry* g, uint64_t seed, uint64_t* h_witness_cols, uint64_t* h_setup_cols);
/* Same, with the two roles of the seed separated: `setup_seed` fixes what belongs to the circuit TYPE (gate constants; the
 * setup columns depend on it alone), `witness_seed` the free witness values of one circuit INSTANCE -- many instances of
 * one type share a setup / verification key, as in the reference (src/tests/complex_tests/mod.rs:316-410). */
ZKGPU_API int zkgpu_synth_trace_instance(const zkgpu_geometry* g, uint64_t setup_seed, uint64_t witness_seed, uint64_t* h_witness_cols,
                                         uint64_t* h_setup_cols);

#ifdef __cplusplus
}
#endif
#endif /* ZKGPU_H */
