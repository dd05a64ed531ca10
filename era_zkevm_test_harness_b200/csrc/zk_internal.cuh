// zk_internal.cuh -- internal declarations shared by the CUDA translation units of libzkgpu.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "gl.cuh"

namespace zk {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                                         \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess)                                                                                   \
            throw zk::Error(2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                                   std::to_string(__LINE__));                                                    \
    } while (0)

#define ZK_REQUIRE(cond, msg)                                  \
    do {                                                       \
        if (!(cond)) throw zk::Error(1, std::string(msg));     \
    } while (0)

struct NttPlan {
    int log_n;
    bool inverse;
    int L1, L2;          // n = 2^L1 (rows) * 2^L2 (cols); L2 == 0 -> single pass
    uint64_t* tw1;       // stage tables
    uint64_t* tw2;
    uint64_t* twA;       // inter-pass twiddle w^x = twA[x & mask] * twB[x >> LA]
    uint64_t* twB;
    int LA;
    uint64_t n_inv;
};
struct CosetTables {
    uint64_t* pre_e;
    uint64_t* pre_t;
};
// powers of EVERY coset shift of a (size, LDE factor) pair: row c holds shift_c^(n2*i1), i1 < 2^L1 (pre_e) and shift_c^i0,
// i0 < 2^L2 (pre_t), shift_c = lde_coset_shift(log_n, log_e, c).  (2^L1 + 2^L2) * E words: 6 MB for 2^15 x 2048.
struct CosetBatchTables {
    uint64_t* pre_e;
    uint64_t* pre_t;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::vector<void*> persistent;  // freed at destroy
    std::map<int, NttPlan> ntt_plans;
    std::map<std::pair<int, uint64_t>, CosetTables> coset_tables;
    std::map<std::pair<int, int>, CosetBatchTables> coset_batch_tables;   // key (log_n, log_e)
    uint64_t kernel_launches = 0;
    // per-context scratch arena for one proof at a time (prover.cu): a bump allocator over one cudaMalloc'd block, so a proof
    // performs no driver allocations at all once the arena has reached its size
    char* arena_base = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    // two staging slots for witnesses uploaded ahead of their proof (zkgpu_witness_stage / zkgpu_prove_staged)
    cudaStream_t copy_stream = nullptr;
    uint64_t* staged[2] = {nullptr, nullptr};
    size_t staged_words[2] = {0, 0};
    bool staged_valid[2] = {false, false};
    cudaEvent_t staged_ready[2] = {nullptr, nullptr}, staged_free[2] = {nullptr, nullptr};
    // the staged upload goes in up to 8 column chunks, each with its own event: a proof whose witness is still crossing PCIe starts on
    // the chunks that have landed (first proof of a run); a fully landed witness is proven in one batch
    cudaEvent_t staged_chunk[2][8] = {{nullptr}, {nullptr}};
    uint32_t staged_chunk_cols[2] = {0, 0}, staged_n_chunks[2] = {0, 0};

    void* alloc_persistent(size_t bytes) {
        void* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 8));
        persistent.push_back(p);
        return p;
    }
};

// ntt.cu
const NttPlan& get_ntt_plan(Ctx* ctx, int log_n, bool inverse);
void ntt_forward_coset(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, int log_n, int n_polys,
                       uint64_t shift);
// the same for cosets c0 .. c0+nc-1 of the 2^log_e-coset LDE in ONE launch per pass: coset c is written at out + (c - c0) * 2^log_n
// (+ poly * out_stride).  The high-LDE compression proofs (512-2048 cosets of 2^12..2^15 points) are launch- and
// latency-bound coset by coset; batched, every pass fills the GPU.
void ntt_forward_cosets(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, int log_n, int n_polys, int log_e,
                        uint32_t c0, uint32_t nc);
void ntt_inverse(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, uint64_t* tmp, size_t tmp_stride,
                 int log_n, int n_polys);
void bitrev_copy(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, int log_n, int n_polys);
inline uint64_t lde_coset_shift(int log_n, int log_lde, uint32_t c) {
    return gl::mul(GL_GEN, gl::pow(gl::omega(log_n + log_lde), gl::bitrev(c, log_lde)));
}

// ntt1024.cu -- the 2^20 fast path (two passes of 1024-point warp transforms)
void ntt1024_forward_coset(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, int n_polys, uint64_t shift);
void ntt1024_inverse(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, uint64_t* tmp, size_t tmp_stride,
                     int n_polys);
void ntt1024_forget(Ctx* ctx);

// poseidon2.cu
void poseidon2_permute_batch(Ctx* ctx, uint64_t* d_states, size_t n_states);
void merkle_build(Ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t n_cols, size_t n_leaves, size_t elems_per_leaf,
                  size_t cap_size, uint64_t* d_tree);
inline size_t merkle_tree_digests(size_t n_leaves, size_t cap_size) { return 2 * n_leaves - cap_size; }
inline size_t merkle_cap_offset(size_t n_leaves, size_t cap_size) { return 2 * n_leaves - 2 * cap_size; }

}  // namespace zk
