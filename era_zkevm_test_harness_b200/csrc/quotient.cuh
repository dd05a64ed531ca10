// quotient.cuh -- parameters of the per-coset quotient kernels (quotient.cu), filled by the prover driver (prover.cu).
#pragma once
#include "host_common.cuh"

namespace zk {

#ifndef ZKGPU_QG_WARPS
#define ZKGPU_QG_WARPS 2   // warps per 32-point tile of quotient_gates_kernel (measured per proof: 2 warps 14.1 ms, 3: 14.8, 4: 14.9)
#endif

struct QuotParams {
    zkgpu_geometry g;
    const uint64_t* wit;    // coset evaluations, column stride cs_w, already offset to the coset
    const uint64_t* setup;  // sigma columns, then constant columns, then table columns
    const uint64_t* s2;     // stage-2 columns (c0, c1 interleaved per Ext2 polynomial)
    size_t cs_w, cs_s, cs_2;
    const uint64_t* omega_br;  // w^bitrev(j)
    const uint64_t* l0_inv;    // 1 / (n (x_j - 1)) on this coset (prover.cu get_l0_inv_table)
    const uint64_t* apow;      // alpha^k, interleaved (c0, c1)
    const uint64_t* beta_k;    // beta * k_i for every copy-permuted column, interleaved (c0, c1); k_i = copy-permutation non-residues
    const uint64_t* rc;        // Poseidon2 round constants (device copy)
    uint64_t* t0;
    uint64_t* t1;              // output (split Ext2), offset to the coset
    uint32_t NP, C, E2, W, lookup_col0;
    uint32_t gate_term0[ZKGPU_MAX_GATES];  // index of the first alpha power of each gate
    uint16_t gate_t0[ZKGPU_QG_WARPS][ZKGPU_MAX_GATES];   // quotient_gates_kernel: warp q evaluates instances [gate_t0[q][g], gate_t1[q][g])
    uint16_t gate_t1[ZKGPU_QG_WARPS][ZKGPU_MAX_GATES];   // of gate g (contiguous cost-balanced segments of the work list)
    uint32_t p2_gate;                      // index of the flattened Poseidon2 gate in g.gates, 0xFFFFFFFF if absent
    uint32_t tail_term0;                   // first alpha power after the gate terms (boolean column, PI, lookup, copy permutation)
    uint64_t shift, xn_minus_1, zh_inv, n_field;
    gl::e2 beta, gamma, lbeta, lgamma;
    gl::e2 lgamma_pow[9];                  // lgamma^q, q <= lookup_width
};

// one coset of the quotient domain: gates -> (Poseidon2 gate) -> boolean/PI/lookup/copy-permutation and division by Z_H
void launch_quotient_coset(Ctx* ctx, const QuotParams& p);

}  // namespace zk
