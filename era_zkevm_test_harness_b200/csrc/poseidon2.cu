// poseidon2.cu -- Poseidon2-Goldilocks (width 12, rate 8, capacity 4, x^7, 4+22+4 rounds) sponge and Merkle tree with cap.
//
// Replaces `GoldilocksPoseidon2Sponge<AbsorptionModeOverwrite>` as tree hasher `H`
// (/root/reference/src/prover_utils.rs:43) inside boojum's oracle commitment.  Parameters: see oracle/primitives.c
// -- PARITY UNPINNED against the reference digests; bit-exact against the CPU oracle.
#include "zk_internal.cuh"
#include "poseidon2_consts.cuh"

namespace zk {

__constant__ uint64_t P2_RC[360] = {ZK_P2_RC_INIT};

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
    uint64_t x2 = gl::sqr(x), x4 = gl::sqr(x2);
    return gl::mul(gl::mul(x4, x2), x);
}

// circ(2*M4, M4, M4) with M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]
__device__ __forceinline__ void p2_external(uint64_t (&s)[12]) {
#pragma unroll
    for (int b = 0; b < 3; b++) {
        uint64_t x0 = s[4 * b], x1 = s[4 * b + 1], x2 = s[4 * b + 2], x3 = s[4 * b + 3];
        uint64_t t0 = gl::add(x0, x1), t1 = gl::add(x2, x3);
        uint64_t t2 = gl::add(gl::dbl(x1), t1), t3 = gl::add(gl::dbl(x3), t0);
        uint64_t t4 = gl::add(gl::dbl(gl::dbl(t1)), t3), t5 = gl::add(gl::dbl(gl::dbl(t0)), t2);
        s[4 * b] = gl::add(t3, t5);
        s[4 * b + 1] = t5;
        s[4 * b + 2] = gl::add(t2, t4);
        s[4 * b + 3] = t4;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint64_t t = gl::add(gl::add(s[i], s[4 + i]), s[8 + i]);
        s[i] = gl::add(s[i], t);
        s[4 + i] = gl::add(s[4 + i], t);
        s[8 + i] = gl::add(s[8 + i], t);
    }
}

// J + diag(2^s), s = [4,14,11,8,0,5,2,9,13,6,3,12]
__device__ __forceinline__ void p2_internal(uint64_t (&s)[12]) {
    constexpr unsigned SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
    uint64_t sum = s[0];
#pragma unroll
    for (int i = 1; i < 12; i++) sum = gl::add(sum, s[i]);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::add(gl::mul_pow2(s[i], SH[i]), sum);
}

__device__ __forceinline__ void p2_permute(uint64_t (&s)[12]) {
    p2_external(s);
    int r = 0;
#pragma unroll 1
    for (int k = 0; k < 4; k++, r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox7(gl::add(s[i], P2_RC[12 * r + i]));
        p2_external(s);
    }
#pragma unroll 1
    for (int k = 0; k < 22; k++, r++) {
        s[0] = sbox7(gl::add(s[0], P2_RC[12 * r]));
        p2_internal(s);
    }
#pragma unroll 1
    for (int k = 0; k < 4; k++, r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox7(gl::add(s[i], P2_RC[12 * r + i]));
        p2_external(s);
    }
}

__global__ void p2_permute_kernel(uint64_t* states, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = states[12 * i + k];
    p2_permute(s);
#pragma unroll
    for (int k = 0; k < 12; k++) states[12 * i + k] = s[k];
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
    for (int k = 0; k < 4; k++) digests[4 * i + k] = s[k];
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
    for (int k = 0; k < 4; k++) next[4 * i + k] = s[k];
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
