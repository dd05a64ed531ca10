// capi.cu -- extern "C" boundary of libzkgpu.so (include/zkgpu.h).  No torch types; exceptions stop here.
#include "../../include/zkgpu.h"
#include "zk_internal.cuh"

namespace zk {
thread_local std::string g_last_error;

template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const Error& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 99;
    }
}

// ---------------------------------------------------------------- FRI fold
__global__ void fri_fold_kernel(const uint64_t* __restrict__ in0, const uint64_t* __restrict__ in1, int log_dom, uint64_t shift_inv,
                                uint64_t w_inv, gl::e2 ch, uint64_t* __restrict__ out0, uint64_t* __restrict__ out1) {
    size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t half = (size_t)1 << (log_dom - 1);
    if (m >= half) return;
    uint64_t xinv = gl::mul(shift_inv, gl::pow(w_inv, gl::bitrev((uint32_t)m, log_dom - 1)));
    ulonglong2 a0 = reinterpret_cast<const ulonglong2*>(in0)[m], a1 = reinterpret_cast<const ulonglong2*>(in1)[m];
    gl::e2 a = gl::make2(a0.x, a1.x), b = gl::make2(a0.y, a1.y);
    gl::e2 sum = gl::add(a, b), dif = gl::mul_base(gl::sub(a, b), xinv);
    gl::e2 r = gl::add(sum, gl::mul(dif, ch));
    out0[m] = r.c0;
    out1[m] = r.c1;
}
void fri_fold(Ctx* ctx, const uint64_t* in0, const uint64_t* in1, int log_dom, uint64_t shift, gl::e2 ch, uint64_t* out0, uint64_t* out1) {
    ZK_REQUIRE(log_dom >= 1 && log_dom <= 32, "fri_fold: bad domain");
    size_t half = (size_t)1 << (log_dom - 1);
    fri_fold_kernel<<<(unsigned)((half + 255) / 256), 256, 0, ctx->stream>>>(in0, in1, log_dom, gl::inv(shift), gl::inv(gl::omega(log_dom)),
                                                                             ch, out0, out1);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
}

void lde_batch(Ctx* ctx, const uint64_t* d_values, size_t val_stride, uint64_t* d_mono, size_t mono_stride, uint64_t* d_lde,
               size_t lde_stride, int log_n, int log_lde, int n_cols) {
    size_t n = (size_t)1 << log_n;
    ZK_REQUIRE(lde_stride >= (n << log_lde), "lde: lde_stride too small");
    // scratch for the two-pass inverse: first coset region of the LDE output (overwritten afterwards)
    ntt_inverse(ctx, d_values, val_stride, d_mono, mono_stride, d_lde, lde_stride, log_n, n_cols);
    ntt_forward_cosets(ctx, d_mono, mono_stride, d_lde, lde_stride, log_n, n_cols, log_lde, 0, 1u << log_lde);
}
}  // namespace zk

using zk::Ctx;
struct zkgpu_ctx {
    Ctx c;
};

extern "C" {

int zkgpu_abi_version(void) { return 1; }
const char* zkgpu_last_error(void) { return zk::g_last_error.c_str(); }

int zkgpu_ctx_create(int device, void* cuda_stream, zkgpu_ctx** out) {
    return zk::guarded([&] {
        ZK_REQUIRE(out != nullptr, "ctx_create: out is NULL");
        int n_dev = 0;
        cudaError_t e = cudaGetDeviceCount(&n_dev);
        if (e != cudaSuccess || n_dev == 0)
            throw zk::Error(3, std::string("no CUDA device available (libzkgpu has no CPU fallback): ") + cudaGetErrorString(e));
        ZK_REQUIRE(device >= 0 && device < n_dev, "ctx_create: bad device ordinal");
        CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) throw zk::Error(3, "libzkgpu is built for sm_100a (Blackwell B200) only");
        zkgpu_ctx* ctx = new zkgpu_ctx();
        ctx->c.device = device;
        // NULL is CUDA's own handle for the legacy default stream: work is then ordered with every other
        // default-stream user of the process (e.g. torch's current stream when none was set).
        ctx->c.stream = (cudaStream_t)cuda_stream;
        // keep freed stream-ordered allocations cached in the pool: a proof allocates and frees tens of GB of scratch
        cudaMemPool_t pool;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ULL;
        CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        *out = ctx;
    });
}

void zkgpu_ctx_destroy(zkgpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    zk::ntt1024_forget(&ctx->c);
    for (void* p : ctx->c.persistent) cudaFree(p);
    if (ctx->c.arena_base) cudaFree(ctx->c.arena_base);
    if (ctx->c.copy_stream) { cudaStreamSynchronize(ctx->c.copy_stream); cudaStreamDestroy(ctx->c.copy_stream); }
    for (int k = 0; k < 2; k++) {
        if (ctx->c.staged[k]) cudaFree(ctx->c.staged[k]);
        if (ctx->c.staged_ready[k]) cudaEventDestroy(ctx->c.staged_ready[k]);
        if (ctx->c.staged_free[k]) cudaEventDestroy(ctx->c.staged_free[k]);
        for (int q = 0; q < 8; q++) if (ctx->c.staged_chunk[k][q]) cudaEventDestroy(ctx->c.staged_chunk[k][q]);
    }
    if (ctx->c.own_stream) cudaStreamDestroy(ctx->c.stream);
    delete ctx;
}

void* zkgpu_host_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 8, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        zk::g_last_error = std::string("host_alloc: ") + cudaGetErrorString(e);
        return nullptr;
    }
    return p;
}
void zkgpu_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int zkgpu_ctx_synchronize(zkgpu_ctx* ctx) {
    return zk::guarded([&] { CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream)); });
}
uint64_t zkgpu_ctx_kernel_launches(const zkgpu_ctx* ctx) { return ctx->c.kernel_launches; }

int zkgpu_ntt_forward(zkgpu_ctx* ctx, const uint64_t* d_in, size_t in_stride, uint64_t* d_out, size_t out_stride, int log_n, int n_cols,
                      uint64_t coset_shift) {
    return zk::guarded([&] {
        ZK_REQUIRE(log_n >= 0 && log_n <= 24, "ntt_forward: log_n out of range [0,24]");
        ZK_REQUIRE(coset_shift != 0 && coset_shift < GL_P, "ntt_forward: bad coset shift");
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::ntt_forward_coset(&ctx->c, d_in, in_stride, d_out, out_stride, log_n, n_cols, coset_shift);
    });
}
int zkgpu_ntt_inverse(zkgpu_ctx* ctx, const uint64_t* d_in, size_t in_stride, uint64_t* d_out, size_t out_stride, uint64_t* d_tmp,
                      size_t tmp_stride, int log_n, int n_cols) {
    return zk::guarded([&] {
        ZK_REQUIRE(log_n >= 0 && log_n <= 24, "ntt_inverse: log_n out of range [0,24]");
        ZK_REQUIRE(log_n <= 11 || d_tmp != nullptr, "ntt_inverse: scratch buffer required for log_n > 11");
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::ntt_inverse(&ctx->c, d_in, in_stride, d_out, out_stride, d_tmp, tmp_stride, log_n, n_cols);
    });
}
int zkgpu_lde(zkgpu_ctx* ctx, const uint64_t* d_values, size_t val_stride, uint64_t* d_mono, size_t mono_stride, uint64_t* d_lde,
              size_t lde_stride, int log_n, int log_lde, int n_cols) {
    return zk::guarded([&] {
        ZK_REQUIRE(log_n >= 0 && log_n <= 24 && log_lde >= 0 && log_lde <= 12, "lde: size out of range");
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::lde_batch(&ctx->c, d_values, val_stride, d_mono, mono_stride, d_lde, lde_stride, log_n, log_lde, n_cols);
    });
}
int zkgpu_poseidon2_permute(zkgpu_ctx* ctx, uint64_t* d_states, size_t n_states) {
    return zk::guarded([&] {
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::poseidon2_permute_batch(&ctx->c, d_states, n_states);
    });
}
int zkgpu_merkle_build(zkgpu_ctx* ctx, const uint64_t* d_cols, size_t col_stride, size_t n_cols, size_t n_leaves, size_t elems_per_leaf,
                       size_t cap_size, uint64_t* d_tree) {
    return zk::guarded([&] {
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::merkle_build(&ctx->c, d_cols, col_stride, n_cols, n_leaves, elems_per_leaf, cap_size, d_tree);
    });
}
int zkgpu_fri_fold(zkgpu_ctx* ctx, const uint64_t* d_in_c0, const uint64_t* d_in_c1, int log_dom, uint64_t shift, const uint64_t challenge[2],
                   uint64_t* d_out_c0, uint64_t* d_out_c1) {
    return zk::guarded([&] {
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::fri_fold(&ctx->c, d_in_c0, d_in_c1, log_dom, shift, gl::make2(challenge[0], challenge[1]), d_out_c0, d_out_c1);
    });
}

int zkgpu_commit_columns_host(zkgpu_ctx* ctx, const uint64_t* h_cols, size_t n_cols, int log_n, int log_lde, size_t cap_size,
                              uint64_t* h_cap_out) {
    return zk::guarded([&] {
        ZK_REQUIRE(log_n >= 0 && log_n <= 24 && log_lde >= 0 && log_lde <= 12, "commit: size out of range");
        Ctx* c = &ctx->c;
        CUDA_CHECK(cudaSetDevice(c->device));
        size_t n = (size_t)1 << log_n, ln = n << log_lde;
        ZK_REQUIRE(cap_size <= ln, "commit: cap larger than the LDE domain");
        uint64_t *d_vals = nullptr, *d_mono = nullptr, *d_lde = nullptr, *d_tree = nullptr;
        size_t tree_digests = zk::merkle_tree_digests(ln, cap_size);
        CUDA_CHECK(cudaMallocAsync((void**)&d_vals, n_cols * n * 8, c->stream));
        CUDA_CHECK(cudaMallocAsync((void**)&d_mono, n_cols * n * 8, c->stream));
        CUDA_CHECK(cudaMallocAsync((void**)&d_lde, n_cols * ln * 8, c->stream));
        CUDA_CHECK(cudaMallocAsync((void**)&d_tree, tree_digests * 32, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_vals, h_cols, n_cols * n * 8, cudaMemcpyHostToDevice, c->stream));
        zk::lde_batch(c, d_vals, n, d_mono, n, d_lde, ln, log_n, log_lde, (int)n_cols);
        zk::merkle_build(c, d_lde, ln, n_cols, ln, 1, cap_size, d_tree);
        CUDA_CHECK(cudaMemcpyAsync(h_cap_out, d_tree + 4 * zk::merkle_cap_offset(ln, cap_size), cap_size * 32, cudaMemcpyDeviceToHost,
                                   c->stream));
        CUDA_CHECK(cudaFreeAsync(d_vals, c->stream));
        CUDA_CHECK(cudaFreeAsync(d_mono, c->stream));
        CUDA_CHECK(cudaFreeAsync(d_lde, c->stream));
        CUDA_CHECK(cudaFreeAsync(d_tree, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

}  // extern "C"
