// poseidon2.cu -- Poseidon2-Goldilocks (width 12, rate 8, capacity 4, x^7, 4+22+4 rounds) sponge and Merkle tree with cap.
//
// Replaces `GoldilocksPoseidon2Sponge<AbsorptionModeOverwrite>` as tree hasher `H`
// (/root/reference/src/prover_utils.rs:43) inside boojum's oracle commitment.  Parameters: see oracle/primitives.c and
// tools/gen_poseidon_constants.py -- PINNED: the oracle's leaf and node hashes reproduce the digests of the reference's golden
// proofs (tests/test_hash_pin_cpu.py, tests/test_golden_verify_cpu.py), and these kernels are bit-exact against the oracle.
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

// one thread per leaf; leaf i = for each column c: elems_per_leaf consecutive values at cols[c*stride + i*epl ..].
// The sponge absorbs 8 elements per permutation (overwrite mode, zero padding of the last chunk); the loop is arranged so
// the permutation has a single call site (code size, see poseidon2_core.cuh).
__global__ void __launch_bounds__(128, 10) leaf_hash_kernel(const uint64_t* __restrict__ cols, size_t col_stride, int n_cols, size_t n_leaves,
                                                        int log_epl, uint64_t* __restrict__ digests) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_leaves) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    const int leaf_len = n_cols << log_epl;
    const int epl_mask = (1 << log_epl) - 1;
    const uint64_t* base = cols + (i << log_epl);
#pragma unroll 1
    for (int e0 = 0; e0 < leaf_len; e0 += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int e = e0 + k;
            s[k] = e < leaf_len ? base[(size_t)(e >> log_epl) * col_stride + (e & epl_mask)] : 0;
        }
        p2_permute(s);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) digests[4 * i + k] = glx::canon(s[k]);
}

// one thread per node of the next level (wide levels: every lane busy)
__global__ void __launch_bounds__(128, 10) node_hash_kernel(const uint64_t* __restrict__ prev, uint64_t* __restrict__ next, size_t n_next) {
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

// The narrow top of the tree, several levels per launch: a CTA takes 2*NT adjacent digests of one level, hashes them pairwise
// (NT threads), keeps the results in shared memory, halves again ... and finishes the last five levels inside one warp with
// shuffles.  Every level is also written to the tree in global memory (the query phase reads Merkle paths from it).
// Wide levels use node_hash_kernel instead: hashing is ALU-bound, and a CTA that idles most of its threads while it walks up
// the tree costs more than the launches it saves (measured: +2 ms per proof when used from the leaf level).
template <int NT>
__global__ void __launch_bounds__(NT) node_levels_kernel(uint64_t* __restrict__ tree, size_t in_off, size_t in_width, int in_per_block, int levels) {
    __shared__ uint64_t sm[NT][4];
    const int t = threadIdx.x;
    size_t off = in_off, width = in_width;   // current input level: digest offset in the tree, number of digests
    int active = in_per_block >> 1;           // nodes this CTA produces at the next level
    uint64_t d[4] = {0, 0, 0, 0};
    for (int lvl = 0; lvl < levels; lvl++) {
        const size_t out_off = off + width;
        const bool mine = t < active;
        uint64_t s[12];
        if (lvl == 0) {
            if (mine) {
                const uint64_t* src = tree + 4 * (off + (size_t)blockIdx.x * in_per_block + 2 * (size_t)t);
#pragma unroll
                for (int k = 0; k < 8; k++) s[k] = src[k];
            }
        } else if (2 * active <= 32) {
            // inputs live in lanes 0 .. 2*active-1 of warp 0 (registers d[]): fetch children with shuffles
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint64_t l = __shfl_sync(0xffffffffu, d[k], (2 * t) & 31), r = __shfl_sync(0xffffffffu, d[k], (2 * t + 1) & 31);
                s[k] = l; s[4 + k] = r;
            }
        } else {
            __syncthreads();
            if (t < 2 * active) {
#pragma unroll
                for (int k = 0; k < 4; k++) sm[t][k] = d[k];
            }
            __syncthreads();
            if (mine) {
#pragma unroll
                for (int k = 0; k < 4; k++) { s[k] = sm[2 * t][k]; s[4 + k] = sm[2 * t + 1][k]; }
            }
        }
        if (mine) {
#pragma unroll
            for (int k = 8; k < 12; k++) s[k] = 0;
            p2_permute(s);
            uint64_t* dst = tree + 4 * (out_off + (size_t)blockIdx.x * active + t);
#pragma unroll
            for (int k = 0; k < 4; k++) { d[k] = glx::canon(s[k]); dst[k] = d[k]; }
        }
        off = out_off;
        width >>= 1;
        active >>= 1;
        if (t >= 32 && 2 * active <= 16) return;   // only warp 0 is needed from here on (no further __syncthreads)
    }
}

void merkle_build(Ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t n_cols, size_t n_leaves, size_t elems_per_leaf,
                  size_t cap_size, uint64_t* d_tree) {
    ZK_REQUIRE(n_leaves && (n_leaves & (n_leaves - 1)) == 0, "merkle: n_leaves must be a power of two");
    ZK_REQUIRE(cap_size && (cap_size & (cap_size - 1)) == 0 && cap_size <= n_leaves, "merkle: bad cap size");
    ZK_REQUIRE(elems_per_leaf && (elems_per_leaf & (elems_per_leaf - 1)) == 0, "merkle: elems_per_leaf must be a power of two");
    int log_epl = 0;
    while (((size_t)1 << log_epl) < elems_per_leaf) log_epl++;
    leaf_hash_kernel<<<(unsigned)((n_leaves + 127) / 128), 128, 0, ctx->stream>>>(d_cols, col_stride, (int)n_cols, n_leaves, log_epl, d_tree);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
    constexpr int NT = 256;
    constexpr size_t FUSE_BELOW = 8192;   // levels with at most this many digests are finished by the fused kernel
    size_t off = 0, width = n_leaves;
    while (width > cap_size) {
        if (width > FUSE_BELOW) {
            const size_t nn = width / 2;
            node_hash_kernel<<<(unsigned)((nn + 127) / 128), 128, 0, ctx->stream>>>(d_tree + 4 * off, d_tree + 4 * (off + width), nn);
            CUDA_CHECK(cudaGetLastError());
            ctx->kernel_launches++;
            off += width;
            width = nn;
            continue;
        }
        // blocks of up to 2*NT inputs; as many levels per launch as the block (and the cap) allow
        const size_t per_block = width < 2 * NT ? width : 2 * NT;
        const size_t n_blocks = width / per_block;
        int levels = 0;
        while (((size_t)1 << (levels + 1)) <= per_block && (width >> (levels + 1)) >= cap_size) levels++;
        node_levels_kernel<NT><<<(unsigned)n_blocks, NT, 0, ctx->stream>>>(d_tree, off, width, (int)per_block, levels);
        CUDA_CHECK(cudaGetLastError());
        ctx->kernel_launches++;
        for (int l = 0; l < levels; l++) { off += width; width >>= 1; }
    }
}

}  // namespace zk
