// host_common.cuh -- host-side pieces shared by the prover driver and the CPU verifier of libzkgpu.so:
// geometry-derived shapes, the Poseidon2 transcript, proof buffer layout.
//
// Reference: these are the parts of boojum's `prove_from_precomputations` / `Verifier::verify`
// (/root/reference/src/prover_utils.rs:338-372) that stay on the host in any implementation -- Fiat-Shamir transcript
// (`GoldilocksPoisedon2Transcript`, prover_utils.rs:38), proof assembly (`Proof<F,H,EXT>`, field order as in the golden
// JSON files under /root/reference/test_proofs/).
#pragma once
#include <cstring>
#include <vector>
#include "../../include/zkgpu.h"
#include "gates.cuh"
#include "zk_internal.cuh"

namespace zk {

// ---- host Poseidon2 (same parameters as poseidon2.cu) ----
extern const uint64_t H_P2_RC[360];
void host_poseidon2_permute(uint64_t (&s)[12]);
void host_hash_leaf(const uint64_t* els, size_t n, uint64_t out[4]);
void host_hash_node(const uint64_t* l, const uint64_t* r, uint64_t out[4]);

inline uint32_t ilog2(size_t x) { uint32_t r = 0; while (((size_t)1 << r) < x) r++; return r; }

struct Shape {
    size_t N, LN, depth;
    uint32_t log_n, log_lde, W, S, S2, Q, NP, C, E2, QD, n_at_z, n_at_zw, n_at_0, n_final, n_terms, NF;
    uint32_t lookup_col0;  // first lookup column inside the witness
    uint32_t plain_col0;   // first plain (not copy-permuted) witness column = NP; gate cell c >= n_copy is column plain_col0 + c - n_copy
    size_t fri_dom_log[ZKGPU_MAX_FRI_ORACLES + 1], fri_leaves[ZKGPU_MAX_FRI_ORACLES], fri_cap[ZKGPU_MAX_FRI_ORACLES],
        fri_depth[ZKGPU_MAX_FRI_ORACLES];
    size_t proof_len;
};
Shape make_shape(const zkgpu_geometry& g, const zkgpu_proof_config& cfg);

// Position of every opened polynomial inside the proof's `values_at_z` (and so its power of the DEEP challenge).  The code keeps
// openings internally in ORACLE order -- index i: witness columns [0,W), setup columns [W,W+S), stage-2 Ext2 polys, quotient
// Ext2 polys -- and permutes at the boundary.  Reference order (boojum's verifier walks values_at_z as: variables, witness
// columns, constants, copy-permutation sigmas, grand product z, partial products, lookup multiplicities, lookup A polys, lookup B,
// lookup table columns, quotient chunks).  This order -- with the setup leaf itself stored sigmas, constants, table columns and
// the multiplicity column last in the witness leaf -- is pinned hash-free on the reference's golden proofs (tools/golden_deep.py,
// tests/golden/deep_*.json): node_layer_proof_3_0_0, compression modes 1 and 2, and five base-layer circuit types with lookups
// of width 1, 3 and 4 (basic_circuit_proof_{3,4,8,10,13}_0).
inline std::vector<uint32_t> opening_positions(const zkgpu_geometry& g, const Shape& sh) {
    std::vector<uint32_t> pos(sh.n_at_z);
    uint32_t p = 0;
    const uint32_t n_wit_first = sh.W - (g.lookup_reps ? 1 : 0);
    const uint32_t o_s = sh.W, o_2 = sh.W + sh.S, o_q = sh.W + sh.S + sh.E2;
    for (uint32_t i = 0; i < n_wit_first; i++) pos[i] = p++;                                   // variables, plain witness columns
    for (uint32_t i = 0; i < g.n_const_cols; i++) pos[o_s + sh.NP + i] = p++;                  // constants
    for (uint32_t i = 0; i < sh.NP; i++) pos[o_s + i] = p++;                                   // sigmas
    for (uint32_t i = 0; i < sh.C; i++) pos[o_2 + i] = p++;                                    // z, partial products
    if (g.lookup_reps) pos[sh.W - 1] = p++;                                                    // multiplicities
    for (uint32_t i = sh.C; i < sh.E2; i++) pos[o_2 + i] = p++;                                // lookup A polys, B
    for (uint32_t i = sh.NP + g.n_const_cols; i < sh.S; i++) pos[o_s + i] = p++;               // lookup table columns
    for (uint32_t i = 0; i < sh.QD; i++) pos[o_q + i] = p++;                                   // quotient chunks
    return pos;
}
void validate(const zkgpu_geometry& g, const zkgpu_proof_config& cfg);

// Copy-permutation non-residues k_0 .. k_{n-1} (column i of the permutation argument lives on the coset k_i * H).  boojum's
// `non_residues_for_copy_permutation` -> `make_non_residues` (the bellman routine): k_0 = 1, then the successive smallest quadratic
// non-residues whose cosets are new: 7, 11, 13, 14, 19, 21, 22 ... [recalled; the sigma columns the Rust side hands over are built
// with boojum's values, so this table has to equal them -- not observable on the golden proofs until the quotient identity is
// pinned, DESIGN.md section 5].
inline std::vector<uint64_t> copy_permutation_non_residues(uint32_t n, int log_n) {
    std::vector<uint64_t> k, seen;
    if (n == 0) return k;
    k.push_back(1); seen.push_back(1);
    for (uint64_t cur = 2; k.size() < n; cur++) {
        if (gl::pow(cur, (GL_P - 1) / 2) != GL_P - 1) continue;   // a square
        uint64_t t = cur;
        for (int i = 0; i < log_n; i++) t = gl::sqr(t);            // cur^(domain size) decides the coset
        bool dup = false;
        for (uint64_t s : seen) dup |= s == t;
        if (dup) continue;
        k.push_back(cur); seen.push_back(t);
    }
    return k;
}

constexpr uint64_t PROOF_MAGIC = 0x5A4B50524F4F4631ULL;

// boojum's AlgebraicSpongeBasedTranscript<F, 8, 12, 4, Poseidon2Goldilocks>, PINNED on the reference's golden proofs
// (tools/golden_transcript.py, tests/test_golden_verify_cpu.py: z, the DEEP challenge and every FRI challenge equal the values
// recovered hash-free, and every transcript-derived query index opens all Merkle paths of compression_1..4 and proof.json):
// witnessed elements are buffered; a challenge request absorbs the buffer followed by a ONE (rate 8, overwrite, then zero fill,
// one permutation per block; the state is never reset) and the 8 RATE lanes of the state become the list of available
// challenges; when that list runs out the state is permuted once more.  Query indexes come from a bit buffer (`BoolsBuffer`):
// every challenge contributes its 64 - log2(LDE domain) low bits, LSB first, and an index takes log2(LDE domain) of them.
struct Transcript {
    static constexpr int CHALLENGES_PER_PERMUTATION = 8;   // the rate lanes
    uint64_t st[12] = {0};
    std::vector<uint64_t> buf;
    int pos = CHALLENGES_PER_PERMUTATION;
    uint64_t bitbuf = 0;
    int nbits = 0;
    void absorb(const uint64_t* v, size_t n) { buf.insert(buf.end(), v, v + n); }
    void absorb(const gl::e2& e) { buf.push_back(e.c0); buf.push_back(e.c1); }
    uint64_t challenge() {
        if (!buf.empty()) {
            buf.push_back(1);
            for (size_t i = 0; i < buf.size(); i += 8) {
                for (size_t k = 0; k < 8; k++) st[k] = i + k < buf.size() ? buf[i + k] : 0;
                host_poseidon2_permute(st);
            }
            buf.clear();
            pos = 0;
        } else if (pos == CHALLENGES_PER_PERMUTATION) {
            host_poseidon2_permute(st);
            pos = 0;
        }
        return st[pos++];
    }
    size_t query_index(uint32_t bits) {   // bits = log2(LDE domain) <= 32
        const int take = 64 - (int)bits;
        while (nbits < (int)bits) {
            const uint64_t c = challenge();
            bitbuf |= (take == 64 ? c : (c & (((uint64_t)1 << take) - 1))) << nbits;   // nbits < bits, so nbits + take <= 63
            nbits += take;
        }
        const size_t idx = (size_t)(bitbuf & (((uint64_t)1 << bits) - 1));
        bitbuf >>= bits;
        nbits -= (int)bits;
        return idx;
    }
    gl::e2 challenge_ext() { uint64_t a = challenge(); uint64_t b = challenge(); return gl::make2(a, b); }
};

}  // namespace zk
