// poseidon2.cu -- Poseidon2-Goldilocks (width 12, rate 8, capacity 4, x^7, 4+22+4 rounds) sponge and Merkle tree with cap.
//
// Replaces `GoldilocksPoseidon2Sponge<AbsorptionModeOverwrite>` as tree hasher `H`
// (/root/reference/src/prover_utils.rs:43) inside boojum's oracle commitment.  Parameters: see oracle/primitives.c
// -- PARITY UNPINNED against the reference digests; bit-exact against the CPU oracle.
#include "zk_internal.cuh"
#include "poseidon2_consts.cuh"
#include "poseidon2_core.cuh"

namespace zk {

__constant__ uint64_t P2_RC[360] = {ZK_P2_RC_INIT};

// The permutation itself lives in poseidon2_core.cuh (lazy residues, 96-bit linear layers); lanes are canonicalised only
// where they leave a kernel.
__device__ __forceinline__ void p2_permute(uint64_t (&s)[12]) { p2x_permute(s, P2_RC); }

__global__ void p2_permute_kernel(uint64_t* states, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = states[12 * i + k];
    p2_permute(s);
#pragma unroll
    for (int k = 0; k < 12; k++) states[12 * i + k] = glx::canon(s[k]);
}
void poseidon2_permute_batch(Ctx* ctx, uint64_t* d_states, size_t n_states) {
    if (!n_states) return;
    p2_permute_kernel<<<(unsigned)((n_states + 127) / 128), 128, 0, ctx->stream>>>(d_states, n_states);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
}

// one thread per leaf; leaf i = for each column c: elems_per_leaf consecutive values at cols[c*stride + i*epl ..]
__global__ void __launch_bounds__(128) leaf_hash_kernel(const uint64_t* __restrict__ cols, size_t col_stride, int n_cols, size_t n_leaves,
                                                        int epl, uint64_t* __restrict__ digests) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_leaves) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    int fill = 0;
    for (int c = 0; c < n_cols; c++) {
        const uint64_t* src = cols + (size_t)c * col_stride + i * (size_t)epl;
        for (int e = 0; e < epl; e++) {
            uint64_t v = src[e];
            // s[fill] = v with a static-index switch so the state stays in registers
            switch (fill) {
                case 0: s[0] = v; break; case 1: s[1] = v; break; case 2: s[2] = v; break; case 3: s[3] = v; break;
                case 4: s[4] = v; break; case 5: s[5] = v; break; case 6: s[6] = v; break; default: s[7] = v; break;
            }
            if (++fill == 8) { p2_permute(s); fill = 0; }
        }
    }
    if (fill) {
#pragma unroll
        for (int k = 0; k < 8; k++) if (k >= fill) s[k] = 0;
        p2_permute(s);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) digests[4 * i + k] = glx::canon(s[k]);
}

__global__ void __launch_bounds__(128) node_hash_kernel(const uint64_t* __restrict__ prev, uint64_t* __restrict__ next, size_t n_next) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_next) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = prev[8 * i + k];
#pragma unroll
    for (int k = 8; k < 12; k++) s[k] = 0;
    p2_permute(s);
#pragma unroll
    for (int k = 0; k < 4; k++) next[4 * i + k] = glx::canon(s[k]);
}

void merkle_build(Ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t n_cols, size_t n_leaves, size_t elems_per_leaf,
                  size_t cap_size, uint64_t* d_tree) {
    ZK_REQUIRE(n_leaves && (n_leaves & (n_leaves - 1)) == 0, "merkle: n_leaves must be a power of two");
    ZK_REQUIRE(cap_size && (cap_size & (cap_size - 1)) == 0 && cap_size <= n_leaves, "merkle: bad cap size");
    leaf_hash_kernel<<<(unsigned)((n_leaves + 127) / 128), 128, 0, ctx->stream>>>(d_cols, col_stride, (int)n_cols, n_leaves,
                                                                                  (int)elems_per_leaf, d_tree);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
    uint64_t* prev = d_tree;
    size_t width = n_leaves;
    while (width > cap_size) {
        uint64_t* next = prev + 4 * width;
        size_t nn = width / 2;
        node_hash_kernel<<<(unsigned)((nn + 127) / 128), 128, 0, ctx->stream>>>(prev, next, nn);
        CUDA_CHECK(cudaGetLastError());
        ctx->kernel_launches++;
        prev = next;
        width = nn;
    }
}

}  // namespace zk
