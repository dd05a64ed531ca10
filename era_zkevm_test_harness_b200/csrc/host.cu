// host.cu -- host-only parts of libzkgpu.so: shapes, host Poseidon2, synthetic trace generator, CPU verifier.
//
// zkgpu_verify mirrors `verifier.verify::<H, TR, POW>((), vk, proof)` (/root/reference/src/prover_utils.rs:351-372):
// CPU-only there, CPU-only here.  zkgpu_synth_trace stands in for the reference's Rust circuit synthesis
// (`synthesis_inner`, circuit_definitions/src/circuit_definitions/base_layer/mod.rs:286-313), which cannot run in this
// image (no Rust toolchain): it produces a satisfying trace of the requested geometry and gate set.
#include "host_common.cuh"

namespace zk {

extern thread_local std::string g_last_error;

const uint64_t H_P2_RC[360] = {ZK_P2_RC_INIT};

void host_poseidon2_permute(uint64_t (&s)[12]) {
    p2g_external(s);
    int r = 0;
    for (int k = 0; k < 4; k++, r++) {
        for (int i = 0; i < 12; i++) s[i] = f_pow7(gl::add(s[i], H_P2_RC[12 * r + i]));
        p2g_external(s);
    }
    for (int k = 0; k < 22; k++, r++) {
        s[0] = f_pow7(gl::add(s[0], H_P2_RC[12 * r]));
        p2g_internal(s);
    }
    for (int k = 0; k < 4; k++, r++) {
        for (int i = 0; i < 12; i++) s[i] = f_pow7(gl::add(s[i], H_P2_RC[12 * r + i]));
        p2g_external(s);
    }
}
void host_hash_leaf(const uint64_t* els, size_t n, uint64_t out[4]) {
    uint64_t st[12] = {0};
    for (size_t i = 0; i < n; i += 8) {
        for (size_t k = 0; k < 8; k++) st[k] = i + k < n ? els[i + k] : 0;
        host_poseidon2_permute(st);
    }
    memcpy(out, st, 32);
}
void host_hash_node(const uint64_t* l, const uint64_t* r, uint64_t out[4]) {
    uint64_t st[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    host_poseidon2_permute(st);
    memcpy(out, st, 32);
}

void validate(const zkgpu_geometry& g, const zkgpu_proof_config& cfg) {
    ZK_REQUIRE(g.log_n >= 4 && g.log_n <= 24, "geometry: log_n out of range [4,24]");
    ZK_REQUIRE(g.n_copy >= 1 && g.n_copy < 260, "geometry: n_copy out of range [1,260)");
    ZK_REQUIRE(g.n_witness_plain < 260, "geometry: n_witness_plain out of range [0,260)");
    ZK_REQUIRE(g.quotient_degree >= 2 && g.quotient_degree <= 16 && (g.quotient_degree & (g.quotient_degree - 1)) == 0,
               "geometry: quotient_degree must be a power of two in [2,16]");
    ZK_REQUIRE(g.n_gates <= ZKGPU_MAX_GATES, "geometry: too many gates");
    ZK_REQUIRE(g.n_public_inputs <= ZKGPU_MAX_PUBLIC_INPUTS, "geometry: too many public inputs");
    ZK_REQUIRE((g.lookup_reps == 0) == (g.lookup_width == 0), "geometry: lookup width/reps inconsistent");
    ZK_REQUIRE(g.lookup_reps <= 48, "geometry: more than 48 lookup repetitions (stage-2 per-row scratch)");
    ZK_REQUIRE(g.lookup_width <= 8, "geometry: lookup width > 8");
    ZK_REQUIRE(g.lookup_reps == 0 || g.table_id_col < g.n_const_cols, "geometry: table id column outside the constant columns");
    ZK_REQUIRE(g.table_len <= ((uint32_t)1 << g.log_n), "geometry: table longer than the trace");
    for (uint32_t i = 0; i < g.n_gates; i++) {
        const zkgpu_gate& gt = g.gates[i];
        ZK_REQUIRE(gt.kind < ZKGPU_GATE_KINDS, "geometry: unknown gate kind");
        ZK_REQUIRE(gt.path_len + gt.n_consts <= g.n_const_cols, "geometry: gate selector path + constants exceed constant columns");
        ZK_REQUIRE(gate_width(gt.kind) <= g.n_copy + (gt.kind == ZKGPU_GATE_POSEIDON2_FLATTENED ? g.n_witness_plain : 0),
                   "geometry: gate wider than the columns it may use");
        if (gt.kind == ZKGPU_GATE_NONLINEARITY7) ZK_REQUIRE(gt.n_consts >= 1, "geometry: nonlinearity gate needs 1 constant");
        if (gt.kind == ZKGPU_GATE_FMA) ZK_REQUIRE(gt.n_consts >= 2, "geometry: FMA gate needs 2 constants");
        if (gt.kind == ZKGPU_GATE_REDUCTION4 || gt.kind == ZKGPU_GATE_FMA_EXT) ZK_REQUIRE(gt.n_consts >= 4, "geometry: gate needs 4 constants");
        if (gt.kind == ZKGPU_GATE_UINTX_ADD) ZK_REQUIRE(gt.n_consts >= 1, "geometry: UIntXAdd gate needs 1 constant");
    }
    for (uint32_t i = 0; i < g.n_public_inputs; i++)
        ZK_REQUIRE(g.pi_col[i] < g.n_copy && g.pi_row[i] < ((uint32_t)1 << g.log_n), "geometry: public input location out of range");
    ZK_REQUIRE(cfg.log_lde >= 1 && cfg.log_lde <= 12, "config: log_lde out of range [1,12]");
    ZK_REQUIRE(g.log_n + cfg.log_lde <= 32, "config: LDE domain larger than 2^32 (query indexes and bit reversals are 32-bit)");
    ZK_REQUIRE(cfg.cap_size >= 1 && (cfg.cap_size & (cfg.cap_size - 1)) == 0, "config: cap_size must be a power of two");
    ZK_REQUIRE(cfg.cap_size <= ((size_t)1 << (g.log_n + cfg.log_lde)), "config: cap larger than the LDE domain");
    ZK_REQUIRE(cfg.n_queries >= 1 && cfg.n_queries <= 1024, "config: n_queries out of range");
    ZK_REQUIRE(cfg.n_fri_oracles >= 1 && cfg.n_fri_oracles <= ZKGPU_MAX_FRI_ORACLES, "config: n_fri_oracles out of range");
    ZK_REQUIRE(cfg.pow_bits == 0, "config: proof of work is not used by the reference (NoPow) and not implemented");
    uint32_t total = 0;
    for (uint32_t k = 0; k < cfg.n_fri_oracles; k++) {
        ZK_REQUIRE(cfg.fri_schedule[k] >= 1 && cfg.fri_schedule[k] <= 5, "config: fold factor log2 must be in [1,5]");
        total += cfg.fri_schedule[k];
    }
    ZK_REQUIRE(total <= g.log_n, "config: folding schedule folds below degree 1");
}

Shape make_shape(const zkgpu_geometry& g, const zkgpu_proof_config& cfg) {
    Shape s{};
    s.log_n = g.log_n; s.log_lde = cfg.log_lde;
    s.N = (size_t)1 << g.log_n;
    s.LN = s.N << cfg.log_lde;
    s.depth = ilog2(s.LN / cfg.cap_size);
    s.QD = g.quotient_degree;
    s.NP = g.n_copy + (g.has_boolean_col ? 1 : 0) + g.lookup_width * g.lookup_reps;
    s.lookup_col0 = g.n_copy + (g.has_boolean_col ? 1 : 0);
    s.plain_col0 = s.NP;
    s.W = s.NP + g.n_witness_plain + (g.lookup_reps ? 1 : 0);
    s.S = s.NP + g.n_const_cols + (g.lookup_reps ? g.lookup_width + 1 : 0);
    s.C = (s.NP + s.QD - 1) / s.QD;
    s.E2 = s.C + g.lookup_reps + (g.lookup_reps ? 1 : 0);
    s.S2 = 2 * s.E2;
    s.Q = 2 * s.QD;
    s.n_at_z = s.W + s.S + s.E2 + s.QD;
    s.n_at_zw = 1;
    s.n_at_0 = g.lookup_reps ? g.lookup_reps + 1 : 0;
    // public inputs are NOT quotient terms: the reference opens them through the DEEP polynomial (pinned on golden proofs)
    s.n_terms = total_gate_terms(g) + (g.has_boolean_col ? 1 : 0) + (g.lookup_reps ? g.lookup_reps + 1 : 0) + 1 + s.C;
    s.NF = cfg.n_fri_oracles;
    size_t ld = g.log_n + cfg.log_lde;
    for (uint32_t k = 0; k < cfg.n_fri_oracles; k++) {
        s.fri_dom_log[k] = ld;
        s.fri_leaves[k] = ((size_t)1 << ld) >> cfg.fri_schedule[k];
        s.fri_cap[k] = cfg.cap_size < s.fri_leaves[k] ? cfg.cap_size : s.fri_leaves[k];
        s.fri_depth[k] = ilog2(s.fri_leaves[k] / s.fri_cap[k]);
        ld -= cfg.fri_schedule[k];
    }
    s.fri_dom_log[cfg.n_fri_oracles] = ld;
    s.n_final = (uint32_t)(((size_t)1 << ld) >> cfg.log_lde);
    size_t n = 32 + g.n_public_inputs + 3 * (size_t)cfg.cap_size * 4 + 2 * s.n_final + 2 * (s.n_at_z + s.n_at_zw + s.n_at_0);
    for (uint32_t k = 0; k < cfg.n_fri_oracles; k++) n += s.fri_cap[k] * 4;
    size_t per_q = s.W + s.S2 + s.Q + s.S + 4 * s.depth * 4;
    for (uint32_t k = 0; k < cfg.n_fri_oracles; k++) per_q += 2 * ((size_t)1 << cfg.fri_schedule[k]) + s.fri_depth[k] * 4;
    s.proof_len = n + per_q * cfg.n_queries + 1;
    return s;
}

// ---------------------------------------------------------------------------------------------- synthetic trace
struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    uint64_t field() { return next() % GL_P; }
    uint64_t below(uint64_t n) { return next() % n; }
};

static uint64_t table_entry(uint32_t t, uint32_t j, uint32_t width) {
    if (j == width) return 1;  // table id
    if (j == 0) return t;
    SplitMix m{0x7AB1E000ULL + (uint64_t)t * 16 + j};
    return m.next() & 0xFFFFFFFFULL;
}

// setup_seed drives what a circuit TYPE fixes (gate constants); witness_seed drives the free witness values of one INSTANCE.
// split = false keeps the original single-stream generator (zkgpu_synth_trace).
static void synth_trace(const zkgpu_geometry& g, uint64_t setup_seed, uint64_t seed, bool split, uint64_t* wit, uint64_t* setup) {
    zkgpu_proof_config dummy{};
    dummy.log_lde = 1; dummy.cap_size = 1; dummy.n_queries = 1; dummy.n_fri_oracles = 1; dummy.fri_schedule[0] = 1;
    validate(g, dummy);
    Shape sh = make_shape(g, dummy);
    const size_t N = sh.N;
    const uint32_t NP = sh.NP;
    uint64_t* sigma = setup;
    uint64_t* consts = setup + (size_t)NP * N;
    uint64_t* tables = consts + (size_t)g.n_const_cols * N;
    memset(consts, 0, (size_t)g.n_const_cols * N * 8);
    const uint64_t omega = gl::omega(g.log_n);

    // identity permutation: sigma_i(w^r) = k_i * w^r
    const std::vector<uint64_t> knr = copy_permutation_non_residues(NP, (int)g.log_n);
    {
        std::vector<uint64_t> wp(N);
        uint64_t x = 1;
        for (size_t r = 0; r < N; r++) { wp[r] = x; x = gl::mul(x, omega); }
        for (uint32_t i = 0; i < NP; i++)
            for (size_t r = 0; r < N; r++) sigma[(size_t)i * N + r] = gl::mul(knr[i], wp[r]);
    }
    // tables and multiplicities
    std::vector<uint64_t> mult;
    if (g.lookup_reps) {
        mult.assign(N, 0);
        for (uint32_t j = 0; j <= g.lookup_width; j++)
            for (size_t r = 0; r < N; r++) tables[(size_t)j * N + r] = r < g.table_len ? table_entry((uint32_t)r, j, g.lookup_width) : 0;
    }
    std::vector<long> last_fma_row(1, -1);
    SplitMix rng{seed ^ 0xB2000000ULL};
    SplitMix rng_setup{setup_seed ^ 0x5E7000000ULL};
    SplitMix& rk = split ? rng_setup : rng;
    const uint32_t n_cells = g.n_copy + g.n_witness_plain;   // gate cells: copy columns, then plain witness columns
    std::vector<uint64_t> v(n_cells), kc(g.n_const_cols);
    for (size_t r = 0; r < N; r++) {
        for (uint32_t c = 0; c < n_cells; c++) v[c] = rng.field();
        for (auto& x : kc) x = 0;
        const zkgpu_gate* gt = g.n_gates ? &g.gates[r % g.n_gates] : nullptr;
        if (gt) {
            for (uint32_t b = 0; b < gt->path_len; b++) kc[b] = (gt->path_bits >> b) & 1;
            uint64_t* k = kc.data() + gt->path_len;
            const uint32_t inst = gate_instances(*gt, g);
            switch (gt->kind) {
                case ZKGPU_GATE_CONSTANTS_ALLOCATOR:
                    for (uint32_t t = 0; t < inst; t++) { k[t] = rk.field(); v[t] = k[t]; }
                    break;
                case ZKGPU_GATE_FMA: {
                    k[0] = rk.field(); k[1] = rk.field();
                    long prev = last_fma_row[0];
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 4 * t;
                        if (prev >= 0) {
                            // wire: input a of this row = output d of the previous FMA row (a 2-cycle in sigma)
                            x[0] = wit[(size_t)(4 * t + 3) * N + prev];
                            uint64_t ka = knr[4 * t], kd = knr[4 * t + 3];
                            sigma[(size_t)(4 * t) * N + r] = gl::mul(kd, gl::pow(omega, (uint64_t)prev));
                            sigma[(size_t)(4 * t + 3) * N + prev] = gl::mul(ka, gl::pow(omega, r));
                        }
                        x[3] = gl::add(gl::mul(k[0], gl::mul(x[0], x[1])), gl::mul(k[1], x[2]));
                    }
                    last_fma_row[0] = (long)r;
                } break;
                case ZKGPU_GATE_REDUCTION4:
                    for (int i = 0; i < 4; i++) k[i] = rk.field();
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 5 * t;
                        uint64_t s = 0;
                        for (int i = 0; i < 4; i++) s = gl::add(s, gl::mul(k[i], x[i]));
                        x[4] = s;
                    }
                    break;
                case ZKGPU_GATE_SELECTION:
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 4 * t;
                        x[0] = rng.next() & 1;
                        x[3] = x[0] ? x[1] : x[2];
                    }
                    break;
                case ZKGPU_GATE_PARALLEL_SELECTION4:
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 13 * t;
                        x[0] = rng.next() & 1;
                        for (int i = 0; i < 4; i++) x[3 + 3 * i] = x[0] ? x[1 + 3 * i] : x[2 + 3 * i];
                    }
                    break;
                case ZKGPU_GATE_ZERO_CHECK:
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 3 * t;
                        if (rng.next() & 3) { x[1] = gl::inv(x[0] ? x[0] : (x[0] = 5)); x[2] = 0; }
                        else { x[0] = 0; x[2] = 1; }
                    }
                    break;
                case ZKGPU_GATE_UINTX_ADD:
                    k[0] = (uint64_t)1 << 32;
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 5 * t;
                        x[0] = rng.next() & 0xFFFFFFFFULL; x[1] = rng.next() & 0xFFFFFFFFULL; x[2] = rng.next() & 1;
                        uint64_t s = x[0] + x[1] + x[2];
                        x[3] = s & 0xFFFFFFFFULL; x[4] = s >> 32;
                    }
                    break;
                case ZKGPU_GATE_U32_TRI_ADD_CARRY:
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 5 * t;
                        for (int i = 0; i < 3; i++) x[i] = rng.next() & 0xFFFFFFFFULL;
                        uint64_t s = x[0] + x[1] + x[2];
                        x[3] = s & 0xFFFFFFFFULL; x[4] = s >> 32;  // carry in {0, 1, 2}
                    }
                    break;
                case ZKGPU_GATE_BOUNDED_BOOLEAN:
                case ZKGPU_GATE_BOOLEAN_ALL:
                    for (uint32_t t = 0; t < inst; t++) v[t] = rng.next() & 1;
                    break;
                case ZKGPU_GATE_MATMUL12_EXTERNAL:
                case ZKGPU_GATE_MATMUL12_INNER:
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 24 * t;
                        uint64_t s[12];
                        for (int i = 0; i < 12; i++) s[i] = x[i];
                        if (gt->kind == ZKGPU_GATE_MATMUL12_EXTERNAL) p2g_external(s);
                        else p2g_internal(s);
                        for (int i = 0; i < 12; i++) x[12 + i] = s[i];
                    }
                    break;
                case ZKGPU_GATE_NONLINEARITY7:
                    k[0] = rk.field();
                    for (uint32_t t = 0; t < inst; t++) v[2 * t + 1] = f_pow7(gl::add(v[2 * t], k[0]));
                    break;
                case ZKGPU_GATE_CONDITIONAL_SWAP4:
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 17 * t;
                        x[8] = rng.next() & 1;
                        for (int i = 0; i < 4; i++) { x[9 + i] = x[8] ? x[4 + i] : x[i]; x[13 + i] = x[8] ? x[i] : x[4 + i]; }
                    }
                    break;
                case ZKGPU_GATE_ZERO_CHECK_WITNESS:
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 2 * t;
                        uint64_t& inv = v[g.n_copy + t];
                        if (rng.next() & 3) { inv = gl::inv(x[0] ? x[0] : (x[0] = 5)); x[1] = 0; }
                        else { x[0] = 0; x[1] = 1; }
                    }
                    break;
                case ZKGPU_GATE_DOT_PRODUCT4:
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 9 * t;
                        uint64_t s = 0;
                        for (int i = 0; i < 4; i++) s = gl::add(s, gl::mul(x[2 * i], x[2 * i + 1]));
                        x[8] = s;
                    }
                    break;
                case ZKGPU_GATE_U8X4_FMA:
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 26 * t;
                        uint64_t a = rng.next() & 0x7FFFFFFFULL, b = rng.next() & 0x7FFFFFFFULL, c = rng.next() & 0xFFFFFFFFULL,
                                 ci = rng.next() & 0xFFFFFFFFULL;
                        uint64_t tot = a * b + c + ci;
                        for (int i = 0; i < 4; i++) {
                            x[i] = (a >> (8 * i)) & 0xFF; x[4 + i] = (b >> (8 * i)) & 0xFF; x[8 + i] = (c >> (8 * i)) & 0xFF;
                            x[12 + i] = (ci >> (8 * i)) & 0xFF; x[16 + i] = (tot >> (8 * i)) & 0xFF; x[20 + i] = (tot >> (32 + 8 * i)) & 0xFF;
                        }
                    }
                    break;
                case ZKGPU_GATE_FMA_EXT:
                    for (int i = 0; i < 4; i++) k[i] = rk.field();
                    for (uint32_t t = 0; t < inst; t++) {
                        uint64_t* x = v.data() + 8 * t;
                        gl::e2 d = gl::add(gl::mul(gl::make2(k[0], k[1]), gl::mul(gl::make2(x[0], x[1]), gl::make2(x[2], x[3]))),
                                           gl::mul(gl::make2(k[2], k[3]), gl::make2(x[4], x[5])));
                        x[6] = d.c0; x[7] = d.c1;
                    }
                    break;
                case ZKGPU_GATE_POSEIDON2_FLATTENED:
                    if (inst) {
                        uint64_t s[12];
                        for (int i = 0; i < 12; i++) s[i] = v[i];
                        p2g_external(s);
                        uint32_t col = 12;
                        int rr = 0;
                        for (int q = 0; q < 4; q++, rr++) {
                            for (int i = 0; i < 12; i++) { s[i] = f_pow7(gl::add(s[i], H_P2_RC[12 * rr + i])); v[col + i] = s[i]; }
                            col += 12;
                            p2g_external(s);
                        }
                        for (int q = 0; q < 22; q++, rr++) {
                            s[0] = f_pow7(gl::add(s[0], H_P2_RC[12 * rr]));
                            v[col++] = s[0];
                            p2g_internal(s);
                        }
                        for (int q = 0; q < 4; q++, rr++) {
                            for (int i = 0; i < 12; i++) { s[i] = f_pow7(gl::add(s[i], H_P2_RC[12 * rr + i])); v[col + i] = s[i]; }
                            col += 12;
                            p2g_external(s);
                        }
                    }
                    break;
                default: break;
            }
        }
        for (uint32_t c = 0; c < g.n_copy; c++) wit[(size_t)c * N + r] = v[c];
        for (uint32_t c = 0; c < g.n_witness_plain; c++) wit[(size_t)(sh.plain_col0 + c) * N + r] = v[g.n_copy + c];
        if (g.has_boolean_col) wit[(size_t)g.n_copy * N + r] = rng.next() & 1;
        if (g.lookup_reps) {
            kc[g.table_id_col] = 1;
            for (uint32_t i = 0; i < g.lookup_reps; i++) {
                uint32_t t = (uint32_t)rng.below(g.table_len ? g.table_len : 1);
                mult[t]++;
                for (uint32_t j = 0; j < g.lookup_width; j++)
                    wit[(size_t)(sh.lookup_col0 + i * g.lookup_width + j) * N + r] = table_entry(t, j, g.lookup_width);
            }
        }
        for (uint32_t c = 0; c < g.n_const_cols; c++) consts[(size_t)c * N + r] = kc[c];
    }
    if (g.lookup_reps)
        for (size_t r = 0; r < N; r++) wit[(size_t)(sh.W - 1) * N + r] = mult[r];
}

// ---------------------------------------------------------------------------------------------- verifier
static bool merkle_verify(const uint64_t* leaf, size_t leaf_len, const uint64_t* path, size_t depth, const uint64_t* cap, size_t idx) {
    uint64_t cur[4];
    host_hash_leaf(leaf, leaf_len, cur);
    for (size_t k = 0; k < depth; k++) {
        if (idx & 1) host_hash_node(path + 4 * k, cur, cur);
        else host_hash_node(cur, path + 4 * k, cur);
        idx >>= 1;
    }
    return memcmp(cur, cap + 4 * idx, 32) == 0;
}

struct Fail : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define V_CHECK(cond, msg) do { if (!(cond)) throw Fail(msg); } while (0)

static void verify(const zkgpu_geometry& g, const zkgpu_proof_config& cfg, const uint64_t* vk_cap, const uint64_t* proof, size_t len, uint32_t flags) {
    validate(g, cfg);
    const Shape sh = make_shape(g, cfg);
    V_CHECK(len == sh.proof_len, "proof length does not match geometry/config");
    const uint64_t* p = proof;
    V_CHECK(p[0] == PROOF_MAGIC, "bad magic");
    V_CHECK(p[1] == g.log_n && p[2] == cfg.log_lde && p[3] == cfg.cap_size && p[4] == cfg.n_queries && p[5] == cfg.n_fri_oracles,
            "proof header does not match the proof config");
    V_CHECK(p[6] == sh.W && p[7] == sh.S2 && p[8] == sh.Q && p[9] == sh.S && p[10] == sh.n_at_z && p[11] == sh.n_at_zw && p[12] == sh.n_at_0 &&
                p[13] == g.n_public_inputs && p[14] == sh.n_final && p[15] == cfg.pow_bits,
            "proof header does not match the geometry");
    for (uint32_t k = 0; k < sh.NF; k++) V_CHECK(p[16 + k] == cfg.fri_schedule[k], "folding schedule mismatch");
    // boojum serialises Goldilocks elements as raw u64 and accepts representatives >= p (golden base-layer proofs 4 and 8 contain
    // some); they denote the same field elements, so they are reduced here instead of being rejected
    std::vector<uint64_t> canon_copy;
    for (size_t i = 32; i < len; i++)
        if (proof[i] >= GL_P) {
            if (canon_copy.empty()) canon_copy.assign(proof, proof + len);
            canon_copy[i] = proof[i] - GL_P;
        }
    if (!canon_copy.empty()) { proof = canon_copy.data(); p = proof; }
    p += 32;
    const size_t cap = cfg.cap_size;
    const uint64_t* pi = p; p += g.n_public_inputs;
    const uint64_t* cap_w = p; p += cap * 4;
    const uint64_t* cap_2 = p; p += cap * 4;
    const uint64_t* cap_q = p; p += cap * 4;
    const uint64_t* fin0 = p; p += sh.n_final;
    const uint64_t* fin1 = p; p += sh.n_final;
    const gl::e2* at_z_proof = reinterpret_cast<const gl::e2*>(p); p += 2 * sh.n_at_z;   // reference order (opening_positions)
    const std::vector<uint32_t> open_pos = opening_positions(g, sh);
    std::vector<gl::e2> at_z_v(sh.n_at_z);   // oracle order: witness, setup, stage 2, quotient
    for (uint32_t i = 0; i < sh.n_at_z; i++) at_z_v[i] = at_z_proof[open_pos[i]];
    const gl::e2* at_z = at_z_v.data();
    const gl::e2 at_zw = *reinterpret_cast<const gl::e2*>(p); p += 2;
    const gl::e2* at_0 = reinterpret_cast<const gl::e2*>(p); p += 2 * sh.n_at_0;
    const uint64_t* fri_caps[ZKGPU_MAX_FRI_ORACLES];
    for (uint32_t k = 0; k < sh.NF; k++) { fri_caps[k] = p; p += sh.fri_cap[k] * 4; }
    const uint64_t* queries = p;

    // ---- transcript
    Transcript tr;
    tr.absorb(vk_cap, cap * 4);
    tr.absorb(pi, g.n_public_inputs);
    tr.absorb(cap_w, cap * 4);
    gl::e2 beta = tr.challenge_ext(), gamma = tr.challenge_ext(), lbeta = gl::make2(0, 0), lgamma = gl::make2(0, 0);
    if (g.lookup_reps) { lbeta = tr.challenge_ext(); lgamma = tr.challenge_ext(); }
    tr.absorb(cap_2, cap * 4);
    gl::e2 alpha = tr.challenge_ext();
    tr.absorb(cap_q, cap * 4);
    gl::e2 z = tr.challenge_ext();
    tr.absorb(reinterpret_cast<const uint64_t*>(at_z_proof), 2 * sh.n_at_z);
    tr.absorb(at_zw);
    tr.absorb(reinterpret_cast<const uint64_t*>(at_0), 2 * sh.n_at_0);
    gl::e2 phi = tr.challenge_ext();
    gl::e2 fri_ch[ZKGPU_MAX_FRI_ORACLES];
    for (uint32_t k = 0; k < sh.NF; k++) { tr.absorb(fri_caps[k], sh.fri_cap[k] * 4); fri_ch[k] = tr.challenge_ext(); }
    tr.absorb(fin0, sh.n_final);
    tr.absorb(fin1, sh.n_final);

    // ---- quotient identity at z
    {
        const gl::e2* w = at_z;
        const gl::e2* s = at_z + sh.W;
        const gl::e2* e2v = at_z + sh.W + sh.S;
        const gl::e2* qv = e2v + sh.E2;
        const gl::e2* sigma = s;
        const gl::e2* consts = s + sh.NP;
        const gl::e2* tables = consts + g.n_const_cols;
        const gl::e2 one = gl::make2(1, 0);
        gl::e2 zn = gl::pow(z, (uint64_t)sh.N);
        gl::e2 zh = gl::sub(zn, one);
        gl::e2 acc = gl::make2(0, 0), ap = one;
        for (uint32_t gi = 0; gi < g.n_gates; gi++) {
            const zkgpu_gate& gt = g.gates[gi];
            if (!gate_relations(gt.kind) || !gate_instances(gt, g)) continue;
            gl::e2 sel = one;
            for (uint32_t b = 0; b < gt.path_len; b++) sel = gl::mul(sel, ((gt.path_bits >> b) & 1) ? consts[b] : gl::sub(one, consts[b]));
            gl::e2 ga = gl::make2(0, 0);
            eval_gate<gl::e2>(gt, g, H_P2_RC, [&](uint32_t c) { return w[c < g.n_copy ? c : sh.plain_col0 + (c - g.n_copy)]; }, [&](uint32_t i) { return consts[gt.path_len + i]; },
                              [&](gl::e2 r) { ga = gl::add(ga, gl::mul(ap, r)); ap = gl::mul(ap, alpha); });
            acc = gl::add(acc, gl::mul(ga, sel));
        }
        if (g.has_boolean_col) {
            gl::e2 b = w[g.n_copy];
            acc = gl::add(acc, gl::mul(ap, gl::sub(gl::mul(b, b), b)));
            ap = gl::mul(ap, alpha);
        }
        const uint64_t n_field = (uint64_t)sh.N % GL_P;
        // (public inputs are not quotient terms: they are opened through the DEEP polynomial below)
        if (g.lookup_reps) {
            const uint32_t LW = g.lookup_width;
            gl::e2 gp[16];
            gp[0] = one;
            for (uint32_t j = 1; j <= LW; j++) gp[j] = gl::mul(gp[j - 1], lgamma);
            gl::e2 tid = gl::mul(gp[LW], consts[g.table_id_col]);
            for (uint32_t i = 0; i < g.lookup_reps; i++) {
                gl::e2 den = gl::add(lbeta, tid);
                for (uint32_t j = 0; j < LW; j++) den = gl::add(den, gl::mul(gp[j], w[sh.lookup_col0 + i * LW + j]));
                acc = gl::add(acc, gl::mul(ap, gl::sub(gl::mul(e2v[sh.C + i], den), one)));
                ap = gl::mul(ap, alpha);
            }
            gl::e2 den = lbeta;
            for (uint32_t j = 0; j <= LW; j++) den = gl::add(den, gl::mul(gp[j], tables[j]));
            acc = gl::add(acc, gl::mul(ap, gl::sub(gl::mul(e2v[sh.C + g.lookup_reps], den), w[sh.W - 1])));
            ap = gl::mul(ap, alpha);
            // logUp sum check: sum_i A_i(0) == B(0)  (sum over H of f equals N*f(0) for deg f < N)
            gl::e2 sa = gl::make2(0, 0);
            for (uint32_t i = 0; i < g.lookup_reps; i++) sa = gl::add(sa, at_0[i]);
            V_CHECK(gl::eq(sa, at_0[g.lookup_reps]), "lookup sum check failed (sum A_i(0) != B(0))");
        }
        {
            gl::e2 l0 = gl::mul(zh, gl::inv(gl::mul_base(gl::sub(z, one), n_field)));
            acc = gl::add(acc, gl::mul(ap, gl::mul(l0, gl::sub(e2v[0], one))));
            ap = gl::mul(ap, alpha);
            const std::vector<uint64_t> knr = copy_permutation_non_residues(sh.NP, (int)g.log_n);
            for (uint32_t j = 0; j < sh.C; j++) {
                gl::e2 num = one, den = one;
                for (uint32_t i = j * sh.QD; i < (j + 1) * sh.QD && i < sh.NP; i++) {
                    const gl::e2 kx = gl::mul_base(z, knr[i]);
                    num = gl::mul(num, gl::add(gl::add(w[i], gl::mul(beta, kx)), gamma));
                    den = gl::mul(den, gl::add(gl::add(w[i], gl::mul(beta, sigma[i])), gamma));
                }
                gl::e2 cur = (j + 1 < sh.C) ? e2v[j + 1] : at_zw;
                acc = gl::add(acc, gl::mul(ap, gl::sub(gl::mul(cur, den), gl::mul(e2v[j], num))));
                ap = gl::mul(ap, alpha);
            }
        }
        gl::e2 t = gl::make2(0, 0), zp = one;
        for (uint32_t c = 0; c < sh.QD; c++) { t = gl::add(t, gl::mul(zp, qv[c])); zp = gl::mul(zp, zn); }
        V_CHECK((flags & ZKGPU_VERIFY_SKIP_QUOTIENT_IDENTITY) || gl::eq(acc, gl::mul(t, zh)), "quotient identity does not hold at z");
    }

    // ---- queries
    // phi powers by position in values_at_z, then z*omega, the openings at 0, the public inputs; phip[] is indexed in oracle order
    std::vector<gl::e2> phi_pow(sh.n_at_z + 1 + sh.n_at_0 + g.n_public_inputs), phip(phi_pow.size());
    phi_pow[0] = gl::make2(1, 0);
    for (size_t i = 1; i < phi_pow.size(); i++) phi_pow[i] = gl::mul(phi_pow[i - 1], phi);
    for (size_t i = 0; i < phip.size(); i++) phip[i] = i < sh.n_at_z ? phi_pow[open_pos[i]] : phi_pow[i];
    gl::e2 sum_at_z = gl::make2(0, 0);
    for (uint32_t i = 0; i < sh.n_at_z; i++) sum_at_z = gl::add(sum_at_z, gl::mul(phip[i], at_z[i]));
    const uint64_t omega = gl::omega(g.log_n);
    const gl::e2 zw = gl::mul_base(z, omega);
    const int log_ln = g.log_n + cfg.log_lde;
    const uint64_t omega_ln = gl::omega(log_ln);
    const uint64_t* q = queries;
    for (uint32_t qi = 0; qi < cfg.n_queries; qi++) {
        size_t idx = tr.query_index(ilog2(sh.LN));
        const uint64_t* leaf_w = q; q += sh.W; const uint64_t* path_w = q; q += sh.depth * 4;
        const uint64_t* leaf_2 = q; q += sh.S2; const uint64_t* path_2 = q; q += sh.depth * 4;
        const uint64_t* leaf_q = q; q += sh.Q; const uint64_t* path_q = q; q += sh.depth * 4;
        const uint64_t* leaf_s = q; q += sh.S; const uint64_t* path_s = q; q += sh.depth * 4;
        V_CHECK(merkle_verify(leaf_w, sh.W, path_w, sh.depth, cap_w, idx), "witness oracle Merkle path invalid");
        V_CHECK(merkle_verify(leaf_2, sh.S2, path_2, sh.depth, cap_2, idx), "stage-2 oracle Merkle path invalid");
        V_CHECK(merkle_verify(leaf_q, sh.Q, path_q, sh.depth, cap_q, idx), "quotient oracle Merkle path invalid");
        V_CHECK(merkle_verify(leaf_s, sh.S, path_s, sh.depth, vk_cap, idx), "setup oracle Merkle path invalid");
        uint64_t x = gl::mul(GL_GEN, gl::pow(omega_ln, gl::bitrev((uint32_t)idx, log_ln)));
        gl::e2 s = gl::make2(0, 0);
        uint32_t k = 0;
        for (uint32_t i = 0; i < sh.W; i++) s = gl::add(s, gl::mul_base(phip[k++], leaf_w[i]));
        for (uint32_t i = 0; i < sh.S; i++) s = gl::add(s, gl::mul_base(phip[k++], leaf_s[i]));
        for (uint32_t i = 0; i < sh.E2; i++) s = gl::add(s, gl::mul(phip[k++], gl::make2(leaf_2[2 * i], leaf_2[2 * i + 1])));
        for (uint32_t i = 0; i < sh.QD; i++) s = gl::add(s, gl::mul(phip[k++], gl::make2(leaf_q[2 * i], leaf_q[2 * i + 1])));
        gl::e2 xe = gl::make2(x, 0);
        gl::e2 h = gl::mul(gl::sub(s, sum_at_z), gl::inv(gl::sub(xe, z)));
        h = gl::add(h, gl::mul(gl::mul(phip[k++], gl::sub(gl::make2(leaf_2[0], leaf_2[1]), at_zw)), gl::inv(gl::sub(xe, zw))));
        uint64_t xinv = gl::inv(x);
        for (uint32_t i = 0; i < sh.n_at_0; i++) {
            gl::e2 a = gl::make2(leaf_2[2 * (sh.C + i)], leaf_2[2 * (sh.C + i) + 1]);
            h = gl::add(h, gl::mul(phip[k++], gl::mul_base(gl::sub(a, at_0[i]), xinv)));
        }
        for (uint32_t i = 0; i < g.n_public_inputs; i++) {   // (w_col(x) - value) / (x - omega^row)
            const uint64_t den = gl::sub(x, gl::pow(omega, g.pi_row[i]));
            h = gl::add(h, gl::mul_base(phip[k++], gl::mul(gl::sub(leaf_w[g.pi_col[i]], pi[i]), gl::inv(den))));
        }
        // FRI chain
        size_t di = idx;
        gl::e2 expected = h;
        uint64_t shift = GL_GEN;
        for (uint32_t o = 0; o < sh.NF; o++) {
            const uint32_t sbits = cfg.fri_schedule[o];
            const size_t epl = (size_t)1 << sbits;
            const uint64_t* leaf = q; q += 2 * epl;
            const uint64_t* path = q; q += sh.fri_depth[o] * 4;
            size_t leaf_idx = di >> sbits, pos = di & (epl - 1);
            V_CHECK(merkle_verify(leaf, 2 * epl, path, sh.fri_depth[o], fri_caps[o], leaf_idx), "FRI oracle Merkle path invalid");
            V_CHECK(leaf[pos] == expected.c0 && leaf[epl + pos] == expected.c1,
                    o == 0 ? "DEEP combination does not match the FRI base oracle" : "FRI folding inconsistent between oracles");
            gl::e2 cur[32];
            for (size_t e = 0; e < epl; e++) cur[e] = gl::make2(leaf[e], leaf[epl + e]);
            gl::e2 c = fri_ch[o];
            int ld = (int)sh.fri_dom_log[o];
            size_t base = leaf_idx << sbits, n = epl;
            while (n > 1) {
                uint64_t w_inv = gl::inv(gl::omega(ld)), s_inv = gl::inv(shift);
                for (size_t e = 0; e < n / 2; e++) {
                    uint64_t xi = gl::mul(s_inv, gl::pow(w_inv, gl::bitrev((uint32_t)(base + 2 * e), ld)));
                    gl::e2 a = cur[2 * e], b = cur[2 * e + 1];
                    cur[e] = gl::add(gl::add(a, b), gl::mul(gl::mul_base(gl::sub(a, b), xi), c));
                }
                n >>= 1; ld--; shift = gl::sqr(shift); base >>= 1; c = gl::sqr(c);
            }
            expected = cur[0];
            di = leaf_idx;
        }
        // final polynomial at the point of index di in the last domain
        const int ldf = (int)sh.fri_dom_log[sh.NF];
        uint64_t xf = gl::mul(shift, gl::pow(gl::omega(ldf), gl::bitrev((uint32_t)di, ldf)));
        gl::e2 fv = gl::make2(0, 0);
        for (size_t i = sh.n_final; i-- > 0;) fv = gl::add(gl::mul_base(fv, xf), gl::make2(fin0[i], fin1[i]));
        V_CHECK(gl::eq(fv, expected), "final FRI polynomial does not match the last folded value");
    }
    V_CHECK(*q == 0, "pow_challenge must be 0 (NoPow)");
}

}  // namespace zk

extern "C" {
uint32_t zkgpu_num_witness_cols(const zkgpu_geometry* g) { zkgpu_proof_config c{}; c.log_lde = 1; c.cap_size = 1; return zk::make_shape(*g, c).W; }
uint32_t zkgpu_num_permuted_cols(const zkgpu_geometry* g) { zkgpu_proof_config c{}; c.log_lde = 1; c.cap_size = 1; return zk::make_shape(*g, c).NP; }
uint32_t zkgpu_num_setup_cols(const zkgpu_geometry* g) { zkgpu_proof_config c{}; c.log_lde = 1; c.cap_size = 1; return zk::make_shape(*g, c).S; }
uint32_t zkgpu_num_stage2_cols(const zkgpu_geometry* g) { zkgpu_proof_config c{}; c.log_lde = 1; c.cap_size = 1; return zk::make_shape(*g, c).S2; }
uint32_t zkgpu_num_quotient_cols(const zkgpu_geometry* g) { return 2 * g->quotient_degree; }
size_t zkgpu_proof_size_u64(const zkgpu_geometry* g, const zkgpu_proof_config* cfg) {
    try {
        zk::validate(*g, *cfg);
        return zk::make_shape(*g, *cfg).proof_len;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 0;
    }
}
int zkgpu_synth_trace(const zkgpu_geometry* g, uint64_t seed, uint64_t* h_witness_cols, uint64_t* h_setup_cols) {
    try {
        zk::synth_trace(*g, seed, seed, false, h_witness_cols, h_setup_cols);
        return 0;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 1;
    }
}
int zkgpu_synth_trace_instance(const zkgpu_geometry* g, uint64_t setup_seed, uint64_t witness_seed, uint64_t* h_witness_cols,
                               uint64_t* h_setup_cols) {
    try {
        zk::synth_trace(*g, setup_seed, witness_seed, true, h_witness_cols, h_setup_cols);
        return 0;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 1;
    }
}
int zkgpu_verify(const zkgpu_geometry* g, const zkgpu_proof_config* cfg, const uint64_t* vk_cap, const uint64_t* proof, size_t proof_len_u64) {
    return zkgpu_verify_ex(g, cfg, vk_cap, proof, proof_len_u64, 0);
}
int zkgpu_verify_ex(const zkgpu_geometry* g, const zkgpu_proof_config* cfg, const uint64_t* vk_cap, const uint64_t* proof, size_t proof_len_u64,
                    uint32_t flags) {
    try {
        zk::verify(*g, *cfg, vk_cap, proof, proof_len_u64, flags);
        return 0;
    } catch (const zk::Fail& f) {
        zk::g_last_error = std::string("proof rejected: ") + f.what();
        return 1;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 2;
    }
}
}
