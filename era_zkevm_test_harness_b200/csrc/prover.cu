// prover.cu -- the GPU proving pipeline behind zkgpu_setup_create / zkgpu_prove (include/zkgpu.h).
//
// Replaces boojum's `get_full_setup` and `prove_from_precomputations` as called by the reference at
// /root/reference/src/prover_utils.rs:185-186 and :338-348 (recursion :452, :533).  Stage order and oracle shapes follow
// the reference's golden proofs; the bit-level contract is oracle/prover.c.
//
// Data layout in HBM (all column-major, u64 Goldilocks):
//   values on H      [cols][N]          natural row order (as uploaded / as computed by stage 2)
//   monomials        [cols][N]          natural coefficient order
//   coset evals      [cols][E*N]        E = max(lde, quotient_degree) cosets, coset c = 7*w_{E*N}^bitrev(c), each coset
//                                       bit-reversed; the committed LDE is the prefix [0, lde*N) of every column, the
//                                       quotient kernel reads cosets [0, quotient_degree)
//   Merkle trees     [(2*leaves - cap)][4]
#include <memory>
#include "host_common.cuh"
#include "quotient.cuh"
#include "dot.cuh"
#include "glx.cuh"

namespace zk {

extern thread_local std::string g_last_error;
void fri_fold(Ctx* ctx, const uint64_t* in0, const uint64_t* in1, int log_dom, uint64_t shift, gl::e2 ch, uint64_t* out0, uint64_t* out1);

// ------------------------------------------------------------------------------------------------ device buffers
// Scratch of one proof comes from the context's arena while an ArenaScope is open on this thread (bump allocation, freed
// wholesale when the scope closes; a buffer released while it is the top of the stack gives its space back at once).
// Outside a scope (setup creation) DevBuf falls back to the stream-ordered allocator.
struct ArenaScope;
static thread_local Ctx* g_arena_ctx = nullptr;

struct DevBuf {
    uint64_t* p = nullptr;
    size_t n = 0;
    cudaStream_t stream = nullptr;
    Ctx* arena = nullptr;
    size_t arena_mark = 0;
    void alloc(size_t n_u64, cudaStream_t s) {
        release();
        stream = s;
        n = n_u64;
        if (g_arena_ctx) {
            Ctx* c = g_arena_ctx;
            const size_t bytes = ((n_u64 ? n_u64 : 1) * 8 + 255) & ~(size_t)255;
            if (c->arena_off + bytes > c->arena_cap)
                throw Error(4, "prove: scratch arena exhausted (" + std::to_string(c->arena_cap >> 20) + " MiB)");
            arena = c;
            arena_mark = c->arena_off;
            p = reinterpret_cast<uint64_t*>(c->arena_base + c->arena_off);
            c->arena_off += bytes;
            return;
        }
        CUDA_CHECK(cudaMallocAsync((void**)&p, (n_u64 ? n_u64 : 1) * 8, s));
    }
    void release() {
        if (p) {
            if (arena) {
                const size_t bytes = ((n ? n : 1) * 8 + 255) & ~(size_t)255;
                if (arena->arena_off == arena_mark + bytes) arena->arena_off = arena_mark;   // top of the stack
                arena = nullptr;
            } else {
                cudaFreeAsync(p, stream);
            }
        }
        p = nullptr;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

struct ArenaScope {
    Ctx* ctx;
    bool owner;
    ArenaScope(Ctx* c, size_t bytes) : ctx(c), owner(g_arena_ctx == nullptr) {
        if (!owner) return;
        if (c->arena_cap < bytes) {
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            if (c->arena_base) CUDA_CHECK(cudaFree(c->arena_base));
            c->arena_base = nullptr;
            c->arena_cap = 0;
            CUDA_CHECK(cudaMalloc((void**)&c->arena_base, bytes));
            c->arena_cap = bytes;
        }
        c->arena_off = 0;
        g_arena_ctx = c;
    }
    ~ArenaScope() {
        if (owner) g_arena_ctx = nullptr;
    }
};

struct Setup {
    zkgpu_geometry g;
    zkgpu_proof_config cfg;
    Shape sh;
    uint32_t E;            // cosets kept per column
    DevBuf vals, mono, cosets, tree;
    DevBuf var_maps;       // optional: [NP][N] u32 variable index per copy-permutation cell (zkgpu_setup_set_variable_maps)
    DevBuf wit_maps;       // optional: [n_witness_plain][N] u32 witness index per plain witness cell (zkgpu_setup_set_witness_maps)
    int64_t max_var_index = -1, max_wit_index = -1;   // largest non-placeholder entry of each map
    std::vector<uint64_t> vk_cap;
    int device;
};

// ------------------------------------------------------------------------------------------------ witness materialisation
// cols[c][r] = values[maps[c][r]] (0 for the placeholder): the step boojum performs from `vars_hint` at the top of
// prove_from_precomputations.  One thread per cell; map reads and column writes are coalesced, the value reads are a gather
// (the variable array of a 2^20 circuit is <= a few hundred MB, mostly L2 hits for the hot small-index variables).
__global__ void materialize_columns_kernel(const uint32_t* __restrict__ maps, const uint64_t* __restrict__ values, size_t n_vars,
                                           uint64_t* __restrict__ cols, size_t n_cells) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const uint32_t v = maps[i];
    cols[i] = v == ZKGPU_VAR_PLACEHOLDER ? 0 : values[v];   // v < n_vars is checked on the host against the map's largest index
}

// ------------------------------------------------------------------------------------------------ small kernels
__global__ void omega_br_kernel(uint64_t* out, uint64_t omega, int log_n) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >> log_n) return;
    out[j] = gl::pow(omega, gl::bitrev((uint32_t)j, log_n));
}

// 1 / (n * (x - 1)) on every point of the quotient cosets, coset-major: the denominator of the Lagrange polynomial L_0 used by the
// z(1) = 1 term.  Depends on the domain only, so it is computed once per context (one field inversion per point, 64 MB at 2^20 x 8)
// instead of once per point and proof inside quotient_perm_kernel.
__global__ void l0_inv_table_kernel(uint64_t* out, const uint64_t* omega_br, uint64_t shift, uint64_t n_field, size_t n) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    out[j] = gl::inv(gl::mul(n_field, gl::sub(gl::mul(shift, omega_br[j]), 1)));
}
static const uint64_t* get_omega_br(Ctx* ctx, int log_n);
static void h2d(Ctx* ctx, void* dst, const void* src, size_t bytes);
// copy-permutation non-residues on the device (512 entries per trace length, uploaded once per context)
static const uint64_t* get_non_residues(Ctx* ctx, uint32_t n, int log_n) {
    ZK_REQUIRE(n <= 512, "prove: more than 512 copy-permuted columns");
    auto key = std::make_pair(-2000 - log_n, (uint64_t)0);
    auto it = ctx->coset_tables.find(key);
    if (it != ctx->coset_tables.end()) return it->second.pre_e;
    const std::vector<uint64_t> k = copy_permutation_non_residues(512, log_n);
    uint64_t* d = (uint64_t*)ctx->alloc_persistent(512 * 8);
    h2d(ctx, d, k.data(), 512 * 8);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // k is a stack-lifetime host vector
    CosetTables ct{};
    ct.pre_e = d;
    ctx->coset_tables.emplace(key, ct);
    return d;
}
static const uint64_t* get_l0_inv_table(Ctx* ctx, int log_n, int log_qd) {
    auto key = std::make_pair(-1000 - log_n, (uint64_t)log_qd);
    auto it = ctx->coset_tables.find(key);
    if (it != ctx->coset_tables.end()) return it->second.pre_e;
    const size_t n = (size_t)1 << log_n, qd = (size_t)1 << log_qd;
    const uint64_t* obr = get_omega_br(ctx, log_n);
    uint64_t* d = (uint64_t*)ctx->alloc_persistent(n * qd * 8);
    for (uint32_t c = 0; c < qd; c++) {
        l0_inv_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d + (size_t)c * n, obr, lde_coset_shift(log_n, log_qd, c), (uint64_t)n % GL_P, n);
        CUDA_CHECK(cudaGetLastError());
        ctx->kernel_launches++;
    }
    CosetTables ct{};
    ct.pre_e = d;
    ctx->coset_tables.emplace(key, ct);
    return d;
}

static const uint64_t* get_omega_br(Ctx* ctx, int log_n) {
    auto key = std::make_pair(-log_n - 1, (uint64_t)0);
    auto it = ctx->coset_tables.find(key);
    if (it != ctx->coset_tables.end()) return it->second.pre_e;
    size_t n = (size_t)1 << log_n;
    uint64_t* d = (uint64_t*)ctx->alloc_persistent(n * 8);
    omega_br_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d, gl::omega(log_n), log_n);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
    CosetTables ct{};
    ct.pre_e = d;
    ctx->coset_tables.emplace(key, ct);
    return d;
}

// values on H -> monomials -> E cosets (bit-reversed), all columns
static void extend_columns(Ctx* ctx, const uint64_t* vals, uint64_t* mono, uint64_t* cosets, int log_n, uint32_t E, uint32_t n_cols) {
    size_t N = (size_t)1 << log_n;
    int log_e = (int)ilog2(E);
    // two-pass inverse needs scratch: use the first coset region of the output
    ntt_inverse(ctx, vals, N, mono, N, cosets, (size_t)E * N, log_n, (int)n_cols);
    ntt_forward_cosets(ctx, mono, N, cosets, (size_t)E * N, log_n, (int)n_cols, log_e, 0, E);
}

// ------------------------------------------------------------------------------------------------ stage 2
#define ZKGPU_MAX_LOOKUP_REPS 48   // per-row scratch of stage2_rows_kernel (validate(): lookup_reps <= 48; the reference's largest is 26)
struct Stage2Params {
    const uint64_t* wit;     // [W][N]
    const uint64_t* setup;   // [S][N]
    uint64_t* s2;            // [S2][N]
    uint64_t* rowprod;       // [2][N]  (c0 | c1)
    const uint64_t* knr;     // [NP] copy-permutation non-residues k_i (host_common.cuh)
    uint32_t log_n, NP, C, QD, W, n_const_cols, lookup_width, lookup_reps, table_id_col, lookup_col0;
    gl::e2 beta, gamma, lbeta, lgamma;
    uint64_t omega;
};

// per row: q_j = prod_{l<=j} N_l/D_l for every chunk (q_j for j < C-1 parked in the p_j slots), row total -> rowprod;
// lookup polys A_i = 1/den_i, B = m/den_table.
__global__ void __launch_bounds__(128) stage2_rows_kernel(Stage2Params p) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t N = (size_t)1 << p.log_n;
    if (r >= N) return;
    const uint64_t* sigma = p.setup;
    const uint64_t* consts = p.setup + (size_t)p.NP * N;
    const uint64_t* tables = consts + (size_t)p.n_const_cols * N;
    uint64_t x = gl::pow(p.omega, r);
    // q_j = (prod_{l<=j} N_l) * (prod_{l>j} D_l) / (prod_l D_l): one Ext2 inversion per row
    // suffix products of chunk denominators need a second sweep from the end; C <= 33 so keep them in local memory
    gl::e2 dsuf[40];
    gl::e2 inv_all;
    {
        gl::e2 suf = gl::make2(1, 0);
#pragma unroll 1
        for (int j = (int)p.C - 1; j >= 0; j--) {
            dsuf[j] = suf;  // product of D_l for l > j
            gl::e2 d = gl::make2(1, 0);
#pragma unroll 1
            for (uint32_t i = j * p.QD; i < (j + 1) * p.QD && i < p.NP; i++) {
                // w + gamma + beta*sigma as lazy residues (fused multiply-adds; the Ext2 multiply reduces once)
                const uint64_t wg0 = gl::add(p.wit[(size_t)i * N + r], p.gamma.c0), sv = sigma[(size_t)i * N + r];
                const gl::e2 f = gl::make2(glx::fma(p.beta.c0, sv, wg0), glx::fma(p.beta.c1, sv, p.gamma.c1));
                if (i == j * p.QD) d = f; else d = gl::mul(d, f);   // no multiplication by one for the first column of a chunk
            }
            suf = gl::mul(suf, d);
        }
        inv_all = gl::inv(suf);
    }
    gl::e2 npre = gl::make2(1, 0);
#pragma unroll 1
    for (uint32_t j = 0; j < p.C; j++) {
#pragma unroll 1
        for (uint32_t i = j * p.QD; i < (j + 1) * p.QD && i < p.NP; i++) {
            const uint64_t wg0 = gl::add(p.wit[(size_t)i * N + r], p.gamma.c0);
            const uint64_t kx = gl::mul(p.knr[i], x);   // k_i * x
            npre = gl::mul(npre, gl::make2(glx::fma(p.beta.c0, kx, wg0), glx::fma(p.beta.c1, kx, p.gamma.c1)));
        }
        gl::e2 q = gl::mul(gl::mul(npre, dsuf[j]), inv_all);
        if (j + 1 < p.C) {
            p.s2[(size_t)(2 * (j + 1)) * N + r] = q.c0;
            p.s2[(size_t)(2 * (j + 1) + 1) * N + r] = q.c1;
        } else {
            p.rowprod[r] = q.c0;
            p.rowprod[N + r] = q.c1;
        }
    }
    if (p.lookup_reps) {
        const uint32_t LW = p.lookup_width;
        gl::e2 gp[9];
        gp[0] = gl::make2(1, 0);
        for (uint32_t j = 1; j <= LW; j++) gp[j] = gl::mul(gp[j - 1], p.lgamma);
        gl::e2 tid = gl::mul_base(gp[LW], consts[(size_t)p.table_id_col * N + r]);
        // all lookup_reps + 1 denominators are inverted together (Montgomery's trick: one field inversion and three Ext2
        // multiplications per denominator instead of an inversion each)
        gl::e2 den[ZKGPU_MAX_LOOKUP_REPS + 1], pre[ZKGPU_MAX_LOOKUP_REPS + 1];
        const uint32_t nd = p.lookup_reps + 1;
#pragma unroll 1
        for (uint32_t i = 0; i < p.lookup_reps; i++) {
            gl::e2 d = gl::add(p.lbeta, tid);
            for (uint32_t j = 0; j < LW; j++) d = gl::add(d, gl::mul_base(gp[j], p.wit[(size_t)(p.lookup_col0 + i * LW + j) * N + r]));
            den[i] = d;
        }
        {
            gl::e2 d = p.lbeta;
            for (uint32_t j = 0; j <= LW; j++) d = gl::add(d, gl::mul_base(gp[j], tables[(size_t)j * N + r]));
            den[p.lookup_reps] = d;
        }
        gl::e2 run = gl::make2(1, 0);
#pragma unroll 1
        for (uint32_t i = 0; i < nd; i++) { pre[i] = run; run = gl::mul(run, den[i]); }
        gl::e2 inv = gl::inv(run);   // a zero denominator (probability ~2^-120 per row) makes every inverse of the row zero, as inv(0) = 0 does
#pragma unroll 1
        for (uint32_t i = nd; i-- > 0;) {
            const gl::e2 a = gl::mul(inv, pre[i]);   // 1 / den[i]
            inv = gl::mul(inv, den[i]);
            if (i < p.lookup_reps) {
                p.s2[(size_t)(2 * (p.C + i)) * N + r] = a.c0;
                p.s2[(size_t)(2 * (p.C + i) + 1) * N + r] = a.c1;
            } else {
                const gl::e2 b = gl::mul_base(a, p.wit[(size_t)(p.W - 1) * N + r]);
                p.s2[(size_t)(2 * (p.C + p.lookup_reps)) * N + r] = b.c0;
                p.s2[(size_t)(2 * (p.C + p.lookup_reps) + 1) * N + r] = b.c1;
            }
        }
    }
}

// exclusive prefix product over Ext2 (split storage in0/in1 -> out0/out1), three kernels
__global__ void scan_chunk_prod_kernel(const uint64_t* in0, const uint64_t* in1, size_t n, uint32_t ch, uint64_t* cp0, uint64_t* cp1) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t * ch >= n) return;
    gl::e2 a = gl::make2(1, 0);
    for (uint32_t e = 0; e < ch; e++) a = gl::mul(a, gl::make2(in0[t * ch + e], in1[t * ch + e]));
    cp0[t] = a.c0;
    cp1[t] = a.c1;
}
__global__ void scan_chunks_kernel(uint64_t* cp0, uint64_t* cp1, size_t m) {  // in-place exclusive scan of m chunk products, one CTA
    __shared__ uint64_t s0[1024], s1[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    size_t per = (m + nt - 1) / nt;
    size_t b = (size_t)t * per, e = b + per < m ? b + per : m;
    gl::e2 a = gl::make2(1, 0);
    for (size_t i = b; i < e; i++) a = gl::mul(a, gl::make2(cp0[i], cp1[i]));
    s0[t] = a.c0; s1[t] = a.c1;
    __syncthreads();
    if (t == 0) {
        gl::e2 run = gl::make2(1, 0);
        for (int i = 0; i < nt; i++) {
            gl::e2 v = gl::make2(s0[i], s1[i]);
            s0[i] = run.c0; s1[i] = run.c1;
            run = gl::mul(run, v);
        }
    }
    __syncthreads();
    gl::e2 run = gl::make2(s0[t], s1[t]);
    for (size_t i = b; i < e; i++) {
        gl::e2 v = gl::make2(cp0[i], cp1[i]);
        cp0[i] = run.c0; cp1[i] = run.c1;
        run = gl::mul(run, v);
    }
}
// z[r] = exclusive prefix product; p_j[r] = z[r] * q_j[r]
__global__ void scan_apply_kernel(const uint64_t* in0, const uint64_t* in1, size_t n, uint32_t ch, const uint64_t* cp0, const uint64_t* cp1,
                                  uint64_t* s2, uint32_t C) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t * ch >= n) return;
    gl::e2 run = gl::make2(cp0[t], cp1[t]);
    for (uint32_t e = 0; e < ch; e++) {
        size_t r = t * ch + e;
        gl::e2 v = gl::make2(in0[r], in1[r]);
        s2[r] = run.c0;
        s2[n + r] = run.c1;
        for (uint32_t j = 1; j < C; j++) {
            gl::e2 q = gl::make2(s2[(size_t)(2 * j) * n + r], s2[(size_t)(2 * j + 1) * n + r]);
            q = gl::mul(q, run);
            s2[(size_t)(2 * j) * n + r] = q.c0;
            s2[(size_t)(2 * j + 1) * n + r] = q.c1;
        }
        run = gl::mul(run, v);
    }
}

// ------------------------------------------------------------------------------------------------ quotient (kernels: quotient.cu)
// coefficient i of the big-coset interpolation -> chunk monomials, undoing the shift 7^i
__global__ void quotient_split_kernel(const uint64_t* t0, const uint64_t* t1, uint64_t* qmono, int log_n, size_t qn, uint64_t ginv) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= qn) return;
    size_t N = (size_t)1 << log_n, c = i >> log_n, r = i & (N - 1);
    uint64_t s = gl::pow(ginv, i);
    qmono[(2 * c) * N + r] = gl::mul(t0[i], s);
    qmono[(2 * c + 1) * N + r] = gl::mul(t1[i], s);
}

// ------------------------------------------------------------------------------------------------ evaluation at a point
// powers table pw[i] = z^i (split c0 | c1)
__global__ void ext_pow_table_kernel(uint64_t* pw, size_t n, gl::e2 z) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ch = 16;
    if (t * ch >= n) return;
    gl::e2 v = gl::pow(z, t * ch);
    for (uint32_t e = 0; e < ch && t * ch + e < n; e++) {
        pw[t * ch + e] = v.c0;
        pw[n + t * ch + e] = v.c1;
        v = gl::mul(v, z);
    }
}
// partial[col][block] = sum over the block's slice of mono[col][i] * pw[i]
__global__ void __launch_bounds__(256) eval_partial_kernel(const uint64_t* mono, size_t stride, size_t n, const uint64_t* pw, uint64_t* partial) {
    __shared__ uint64_t s0[256], s1[256];
    const uint64_t* m = mono + (size_t)blockIdx.y * stride;
    DotE2 d;   // sum_i c_i * z^i accumulated unreduced (dot.cuh): n / (gridDim.x * 256) <= 2^15 terms per thread
    dote_zero(d);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dote_add(d, m[i], make_ulonglong2(pw[i], pw[n + i]));
    const gl::e2 acc = dote_reduce(d);
    s0[threadIdx.x] = acc.c0; s1[threadIdx.x] = acc.c1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s0[threadIdx.x] = gl::add(s0[threadIdx.x], s0[threadIdx.x + o]);
            s1[threadIdx.x] = gl::add(s1[threadIdx.x], s1[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x)] = s0[0];
        partial[2 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) + 1] = s1[0];
    }
}
__global__ void eval_final_kernel(const uint64_t* partial, uint32_t nblk, uint64_t* out) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= gridDim.x * blockDim.x) return;
    gl::e2 acc = gl::make2(0, 0);
    for (uint32_t b = 0; b < nblk; b++) acc = gl::add(acc, gl::make2(partial[2 * ((size_t)c * nblk + b)], partial[2 * ((size_t)c * nblk + b) + 1]));
    out[2 * c] = acc.c0; out[2 * c + 1] = acc.c1;
}

// ------------------------------------------------------------------------------------------------ DEEP
struct DeepParams {
    const uint64_t *wit, *setup, *s2, *q;
    size_t cs_w, cs_s, cs_2, cs_q;
    uint32_t W, S, E2, QD, C, n_at_0, log_ln;
    const uint64_t* phip;   // phi^k interleaved
    const uint64_t* at_0;   // interleaved
    gl::e2 sum_at_z, at_zw, z, zw;
    uint64_t omega_ln;
    uint32_t n_pi, pi_col[ZKGPU_MAX_PUBLIC_INPUTS];        // public inputs: (w_col(x) - value) / (x - omega^row)
    uint64_t pi_val[ZKGPU_MAX_PUBLIC_INPUTS], pi_root[ZKGPU_MAX_PUBLIC_INPUTS];
    uint64_t *f0, *f1;
};
__global__ void __launch_bounds__(128) deep_kernel(const __grid_constant__ DeepParams p) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >> p.log_ln) return;
    const ulonglong2* phip = reinterpret_cast<const ulonglong2*>(p.phip);
    uint64_t x = gl::mul(GL_GEN, gl::pow(p.omega_ln, gl::bitrev((uint32_t)idx, p.log_ln)));
    uint32_t k = 0;
    DotE2 d;   // sum_k phi^k * f_k(x) over the base-field columns, unreduced (dot.cuh)
    dote_zero(d);
#pragma unroll 4
    for (uint32_t i = 0; i < p.W; i++) dote_add(d, p.wit[(size_t)i * p.cs_w + idx], phip[k++]);
#pragma unroll 4
    for (uint32_t i = 0; i < p.S; i++) dote_add(d, p.setup[(size_t)i * p.cs_s + idx], phip[k++]);
    gl::e2 s = dote_reduce(d);
    for (uint32_t i = 0; i < p.E2; i++) {
        ulonglong2 a = phip[k++];
        s = gl::add(s, gl::mul(gl::make2(a.x, a.y), gl::make2(p.s2[(size_t)(2 * i) * p.cs_2 + idx], p.s2[(size_t)(2 * i + 1) * p.cs_2 + idx])));
    }
    for (uint32_t i = 0; i < p.QD; i++) {
        ulonglong2 a = phip[k++];
        s = gl::add(s, gl::mul(gl::make2(a.x, a.y), gl::make2(p.q[(size_t)(2 * i) * p.cs_q + idx], p.q[(size_t)(2 * i + 1) * p.cs_q + idx])));
    }
    // 1/(x - z), 1/(x - zw): (a - b u)^-1 = (a + b u)/(a^2 - 7 b^2); 1/x -- one shared base-field inversion
    uint64_t a1 = gl::sub(x, p.z.c0), b1 = p.z.c1, a2 = gl::sub(x, p.zw.c0), b2 = p.zw.c1;
    uint64_t n1 = gl::sub(gl::sqr(a1), gl::mul(7, gl::sqr(b1))), n2 = gl::sub(gl::sqr(a2), gl::mul(7, gl::sqr(b2)));
    uint64_t n12 = gl::mul(n1, n2);
    uint64_t inv = gl::inv(gl::mul(n12, x));
    uint64_t xinv = gl::mul(inv, n12);
    uint64_t i12 = gl::mul(inv, x);
    uint64_t i1 = gl::mul(i12, n2), i2 = gl::mul(i12, n1);
    gl::e2 inv_xz = gl::make2(gl::mul(a1, i1), gl::mul(b1, i1));
    gl::e2 inv_xzw = gl::make2(gl::mul(a2, i2), gl::mul(b2, i2));
    gl::e2 h = gl::mul(gl::sub(s, p.sum_at_z), inv_xz);
    {
        ulonglong2 a = phip[k++];
        gl::e2 zp = gl::make2(p.s2[idx], p.s2[p.cs_2 + idx]);
        h = gl::add(h, gl::mul(gl::mul(gl::make2(a.x, a.y), gl::sub(zp, p.at_zw)), inv_xzw));
    }
    for (uint32_t i = 0; i < p.n_at_0; i++) {
        ulonglong2 a = phip[k++];
        gl::e2 v = gl::make2(p.s2[(size_t)(2 * (p.C + i)) * p.cs_2 + idx], p.s2[(size_t)(2 * (p.C + i) + 1) * p.cs_2 + idx]);
        gl::e2 a0 = gl::make2(p.at_0[2 * i], p.at_0[2 * i + 1]);
        h = gl::add(h, gl::mul(gl::make2(a.x, a.y), gl::mul_base(gl::sub(v, a0), xinv)));
    }
    // public inputs (reference: opened through the DEEP polynomial, not constrained in the quotient).  The reference circuits put
    // all of theirs in ONE row, so consecutive equal roots share the inversion.
    uint64_t last_root = ~0ULL, last_inv = 0;
    for (uint32_t i = 0; i < p.n_pi; i++) {
        ulonglong2 a = phip[k++];
        if (p.pi_root[i] != last_root) { last_root = p.pi_root[i]; last_inv = gl::inv(gl::sub(x, last_root)); }
        h = gl::add(h, gl::mul_base(gl::make2(a.x, a.y), gl::mul(gl::sub(p.wit[(size_t)p.pi_col[i] * p.cs_w + idx], p.pi_val[i]), last_inv)));
    }
    p.f0[idx] = h.c0;
    p.f1[idx] = h.c1;
}

// ------------------------------------------------------------------------------------------------ query gather
// one CTA per query: leaf elements then Merkle path (leaf -> cap), written at out + q*q_stride + offset
__global__ void gather_kernel(const uint64_t* cols, size_t col_stride, uint32_t n_cols, uint32_t epl, const uint64_t* tree, size_t n_leaves,
                              uint32_t depth, const uint32_t* leaf_idx, uint64_t* out, size_t q_stride, size_t offset) {
    const uint32_t q = blockIdx.x;
    const size_t leaf = leaf_idx[q];
    uint64_t* o = out + (size_t)q * q_stride + offset;
    const uint32_t leaf_len = n_cols * epl;
    for (uint32_t i = threadIdx.x; i < leaf_len; i += blockDim.x) {
        uint32_t c = i / epl, e = i % epl;
        o[i] = cols[(size_t)c * col_stride + leaf * epl + e];
    }
    o += leaf_len;
    for (uint32_t i = threadIdx.x; i < depth * 4; i += blockDim.x) {
        uint32_t lvl = i >> 2;
        // offset of level lvl: sum_{l<lvl} n_leaves >> l = 2*n_leaves - (n_leaves >> (lvl-1)) for lvl>0
        size_t off = lvl == 0 ? 0 : 2 * n_leaves - (n_leaves >> (lvl - 1));
        size_t node = (leaf >> lvl) ^ 1;
        o[i] = tree[4 * (off + node) + (i & 3)];
    }
}

// ------------------------------------------------------------------------------------------------ host driver
static void d2h(Ctx* ctx, void* dst, const void* src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
static void h2d(Ctx* ctx, void* dst, const void* src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
}
#define LAUNCH_CHECK(ctx)            \
    do {                             \
        CUDA_CHECK(cudaGetLastError()); \
        (ctx)->kernel_launches++;    \
    } while (0)

static Setup* setup_create(Ctx* ctx, const zkgpu_geometry& g, const zkgpu_proof_config& cfg, const uint64_t* h_setup_cols, uint64_t* h_vk_cap) {
    validate(g, cfg);
    std::unique_ptr<Setup> s(new Setup());
    s->g = g; s->cfg = cfg; s->sh = make_shape(g, cfg); s->device = ctx->device;
    const Shape& sh = s->sh;
    const uint32_t L = 1u << cfg.log_lde;
    s->E = L > sh.QD ? L : sh.QD;
    const size_t N = sh.N;
    s->vals.alloc((size_t)sh.S * N, ctx->stream);
    s->mono.alloc((size_t)sh.S * N, ctx->stream);
    s->cosets.alloc((size_t)sh.S * s->E * N, ctx->stream);
    s->tree.alloc(merkle_tree_digests(sh.LN, cfg.cap_size) * 4, ctx->stream);
    h2d(ctx, s->vals.p, h_setup_cols, (size_t)sh.S * N * 8);
    extend_columns(ctx, s->vals.p, s->mono.p, s->cosets.p, g.log_n, s->E, sh.S);
    merkle_build(ctx, s->cosets.p, (size_t)s->E * N, sh.S, sh.LN, 1, cfg.cap_size, s->tree.p);
    s->vk_cap.resize((size_t)cfg.cap_size * 4);
    d2h(ctx, s->vk_cap.data(), s->tree.p + 4 * merkle_cap_offset(sh.LN, cfg.cap_size), (size_t)cfg.cap_size * 32);
    if (h_vk_cap) memcpy(h_vk_cap, s->vk_cap.data(), (size_t)cfg.cap_size * 32);
    return s.release();
}

// Host witness upload in column chunks on a second stream: chunk k's iNTT + coset NTTs run while chunk k+1 is still
// crossing PCIe (zkgpu_prove).  ready[k] is recorded after columns [k*chunk, (k+1)*chunk) have landed.
struct UploadPlan {
    uint32_t chunk_cols = 0;
    std::vector<cudaEvent_t> ready;
};

static void commit(Ctx* ctx, const Setup& st, const uint64_t* vals, uint32_t n_cols, DevBuf& mono, DevBuf& cosets, uint32_t n_cosets, DevBuf& tree,
                   uint64_t* h_cap, const UploadPlan* plan = nullptr) {
    const Shape& sh = st.sh;
    mono.alloc((size_t)n_cols * sh.N, ctx->stream);
    cosets.alloc((size_t)n_cols * n_cosets * sh.N, ctx->stream);
    tree.alloc(merkle_tree_digests(sh.LN, st.cfg.cap_size) * 4, ctx->stream);
    if (plan && plan->chunk_cols) {
        for (size_t k = 0; k < plan->ready.size(); k++) {
            const uint32_t c0 = (uint32_t)k * plan->chunk_cols, c1 = c0 + plan->chunk_cols < n_cols ? c0 + plan->chunk_cols : n_cols;
            CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, plan->ready[k], 0));
            extend_columns(ctx, vals + (size_t)c0 * sh.N, mono.p + (size_t)c0 * sh.N, cosets.p + (size_t)c0 * n_cosets * sh.N, sh.log_n, n_cosets,
                           c1 - c0);
        }
    } else
    extend_columns(ctx, vals, mono.p, cosets.p, sh.log_n, n_cosets, n_cols);
    merkle_build(ctx, cosets.p, (size_t)n_cosets * sh.N, n_cols, sh.LN, 1, st.cfg.cap_size, tree.p);
    d2h(ctx, h_cap, tree.p + 4 * merkle_cap_offset(sh.LN, st.cfg.cap_size), (size_t)st.cfg.cap_size * 32);
}

static void scan_stage2(Ctx* ctx, const uint64_t* rowprod, uint64_t* s2, size_t N, uint32_t C) {
    uint32_t ch = N >= 4096 ? 64 : 4;
    size_t m = N / ch;
    DevBuf cp;
    cp.alloc(2 * m, ctx->stream);
    scan_chunk_prod_kernel<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(rowprod, rowprod + N, N, ch, cp.p, cp.p + m);
    LAUNCH_CHECK(ctx);
    int nt = m >= 1024 ? 1024 : (int)m;
    scan_chunks_kernel<<<1, nt, 0, ctx->stream>>>(cp.p, cp.p + m, m);
    LAUNCH_CHECK(ctx);
    scan_apply_kernel<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(rowprod, rowprod + N, N, ch, cp.p, cp.p + m, s2, C);
    LAUNCH_CHECK(ctx);
}

// evaluate n_cols base-coefficient columns at an Ext2 point given its power table; results (interleaved c0,c1) -> h_out
static void eval_columns(Ctx* ctx, const uint64_t* mono, size_t stride, uint32_t n_cols, size_t n, const uint64_t* pw, uint64_t* d_partial,
                         uint64_t* d_out, gl::e2* h_out) {
    const uint32_t nblk = 32;
    dim3 grid(nblk, n_cols);
    eval_partial_kernel<<<grid, 256, 0, ctx->stream>>>(mono, stride, n, pw, d_partial);
    LAUNCH_CHECK(ctx);
    eval_final_kernel<<<1, n_cols, 0, ctx->stream>>>(d_partial, nblk, d_out);
    LAUNCH_CHECK(ctx);
    d2h(ctx, h_out, d_out, (size_t)n_cols * 16);
}

// upper bound of the scratch one proof takes from the arena (every buffer of prove(), no reuse assumed)
static size_t prove_scratch_bytes(const Setup& st, bool with_witness_upload) {
    const Shape& sh = st.sh;
    const size_t N = sh.N, L = (size_t)1 << st.cfg.log_lde, E = st.E;
    size_t words = (size_t)sh.W * (E + 1) * N + (size_t)sh.S2 * (E + 2) * N + 6 * (size_t)sh.QD * N + (size_t)sh.Q * (L + 1) * N + 8 * L * N;
    words += 4 * merkle_tree_digests(sh.LN, st.cfg.cap_size) * 4 + 2 * L * N;   // oracle trees, FRI trees
    if (with_witness_upload) words += (size_t)sh.W * N;
    return words * 8 + ((size_t)64 << 20);
}

static void prove(Ctx* ctx, const Setup& st, const uint64_t* d_wit, uint64_t* proof, size_t capacity, const UploadPlan* plan = nullptr) {
    const zkgpu_geometry& g = st.g;
    const zkgpu_proof_config& cfg = st.cfg;
    const Shape& sh = st.sh;
    ZK_REQUIRE(capacity >= sh.proof_len, "prove: proof buffer too small");
    ZK_REQUIRE(st.device == ctx->device, "prove: setup lives on another device");
    ArenaScope arena_scope(ctx, prove_scratch_bytes(st, false));
    const size_t N = sh.N, LN = sh.LN, cap = cfg.cap_size;
    const uint32_t W = sh.W, S = sh.S, S2 = sh.S2, Q = sh.Q, QD = sh.QD, E = st.E, L = 1u << cfg.log_lde;
    const int log_n = (int)g.log_n;
    cudaStream_t stream = ctx->stream;
    std::vector<uint64_t> cap_w(cap * 4), cap_2(cap * 4), cap_q(cap * 4);

    // ---- round 1: witness commitment
    DevBuf mono_w, cos_w, tree_w;
    commit(ctx, st, d_wit, W, mono_w, cos_w, E, tree_w, cap_w.data(), plan);
    std::vector<uint64_t> pi(g.n_public_inputs ? g.n_public_inputs : 1);
    for (uint32_t i = 0; i < g.n_public_inputs; i++)
        CUDA_CHECK(cudaMemcpyAsync(&pi[i], d_wit + (size_t)g.pi_col[i] * N + g.pi_row[i], 8, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));

    Transcript tr;
    tr.absorb(st.vk_cap.data(), cap * 4);
    tr.absorb(pi.data(), g.n_public_inputs);
    tr.absorb(cap_w.data(), cap * 4);
    gl::e2 beta = tr.challenge_ext(), gamma = tr.challenge_ext(), lbeta = gl::make2(0, 0), lgamma = gl::make2(0, 0);
    if (g.lookup_reps) { lbeta = tr.challenge_ext(); lgamma = tr.challenge_ext(); }

    // ---- stage 2
    DevBuf s2v, rowprod, mono_2, cos_2, tree_2;
    s2v.alloc((size_t)S2 * N, stream);
    rowprod.alloc(2 * N, stream);
    {
        Stage2Params p{};
        p.wit = d_wit; p.setup = st.vals.p; p.s2 = s2v.p; p.rowprod = rowprod.p;
        p.log_n = g.log_n; p.NP = sh.NP; p.C = sh.C; p.QD = QD; p.W = W; p.n_const_cols = g.n_const_cols;
        p.lookup_width = g.lookup_width; p.lookup_reps = g.lookup_reps; p.table_id_col = g.table_id_col; p.lookup_col0 = sh.lookup_col0;
        p.beta = beta; p.gamma = gamma; p.lbeta = lbeta; p.lgamma = lgamma; p.omega = gl::omega(log_n);
        p.knr = get_non_residues(ctx, sh.NP, log_n);
        ZK_REQUIRE(sh.C <= 40, "prove: too many copy-permutation chunks");
        stage2_rows_kernel<<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(p);
        LAUNCH_CHECK(ctx);
        scan_stage2(ctx, rowprod.p, s2v.p, N, sh.C);
    }
    commit(ctx, st, s2v.p, S2, mono_2, cos_2, E, tree_2, cap_2.data());
    s2v.release(); rowprod.release();
    tr.absorb(cap_2.data(), cap * 4);
    gl::e2 alpha = tr.challenge_ext();

    // ---- quotient
    DevBuf tq, d_apow, d_rc, qmono, cos_q, tree_q;
    const size_t QN = N * QD;
    const int log_qd = (int)ilog2(QD);
    tq.alloc(4 * QN, stream);  // t0 | t1 | bit-reversal scratch for both
    {
        std::vector<uint64_t> apow(2 * (size_t)sh.n_terms);
        gl::e2 a = gl::make2(1, 0);
        for (uint32_t i = 0; i < sh.n_terms; i++) { apow[2 * i] = a.c0; apow[2 * i + 1] = a.c1; a = gl::mul(a, alpha); }
        // beta * k_i per copy-permuted column (Ext2, interleaved), behind the alpha powers in the same upload
        const size_t bk_off = apow.size();
        {
            const std::vector<uint64_t> knr = copy_permutation_non_residues(sh.NP, log_n);
            for (uint32_t i = 0; i < sh.NP; i++) { const gl::e2 bk = gl::mul_base(beta, knr[i]); apow.push_back(bk.c0); apow.push_back(bk.c1); }
        }
        d_apow.alloc(apow.size(), stream);
        h2d(ctx, d_apow.p, apow.data(), apow.size() * 8);
        d_rc.alloc(360, stream);
        h2d(ctx, d_rc.p, H_P2_RC, 360 * 8);
        CUDA_CHECK(cudaStreamSynchronize(stream));  // apow is a stack-lifetime host vector
        QuotParams p{};
        p.g = g;
        p.cs_w = (size_t)E * N; p.cs_s = (size_t)st.E * N; p.cs_2 = (size_t)E * N;
        p.omega_br = get_omega_br(ctx, log_n);
        const uint64_t* l0_tab = get_l0_inv_table(ctx, log_n, log_qd);
        p.apow = d_apow.p; p.rc = d_rc.p; p.beta_k = d_apow.p + bk_off;
        p.NP = sh.NP; p.C = sh.C; p.E2 = sh.E2; p.W = W; p.lookup_col0 = sh.lookup_col0;
        p.n_field = (uint64_t)N % GL_P;
        p.beta = beta; p.gamma = gamma; p.lbeta = lbeta; p.lgamma = lgamma;
        {
            uint32_t k = 0;
            p.p2_gate = 0xFFFFFFFFu;
            for (uint32_t gi = 0; gi < g.n_gates; gi++) {
                p.gate_term0[gi] = k;
                const uint32_t inst = gate_instances(g.gates[gi], g);
                if (inst && g.gates[gi].kind == ZKGPU_GATE_POSEIDON2_FLATTENED) {
                    ZK_REQUIRE(p.p2_gate == 0xFFFFFFFFu, "prove: more than one flattened Poseidon2 gate");
                    p.p2_gate = gi;
                }
                k += gate_relations(g.gates[gi].kind) * inst;
            }
            p.tail_term0 = k;
            {   // work list of the gates kernel: cut into ZKGPU_QG_WARPS contiguous segments of equal estimated cost
                // rough instruction counts per instance (ncu per-line counts of the MainVM / recursion / compression circuits)
                static const uint32_t COST[ZKGPU_GATE_KINDS] = {0, 25, 60, 110, 45, 200, 80, 80, 110, 600, 0, 330, 60, 40, 1500, 1500, 120, 360, 80, 40};
                const uint32_t PER_GATE = 250;   // selector product + alpha-dot reduction, paid by every warp that touches the gate
                uint64_t total = 0;
                for (uint32_t gi = 0; gi < g.n_gates; gi++) {
                    const uint32_t inst = g.gates[gi].kind == ZKGPU_GATE_POSEIDON2_FLATTENED ? 0 : gate_instances(g.gates[gi], g);
                    if (inst) total += PER_GATE + (uint64_t)inst * COST[g.gates[gi].kind];
                }
                uint64_t base = 0;
                for (uint32_t gi = 0; gi < g.n_gates; gi++) {
                    const uint32_t inst = g.gates[gi].kind == ZKGPU_GATE_POSEIDON2_FLATTENED ? 0 : gate_instances(g.gates[gi], g);
                    const uint64_t c = COST[g.gates[gi].kind] ? COST[g.gates[gi].kind] : 1;
                    uint32_t prev = 0;
                    for (uint32_t q = 0; q < ZKGPU_QG_WARPS; q++) {
                        // instances whose start cost lies below the upper cut of warp q
                        const uint64_t hi = total * (q + 1) / ZKGPU_QG_WARPS;
                        uint32_t end = inst;
                        if (q + 1 < ZKGPU_QG_WARPS) {
                            const uint64_t b0 = base + PER_GATE;
                            end = hi <= b0 ? 0 : (uint32_t)std::min<uint64_t>(inst, (hi - b0 + c - 1) / c);
                        }
                        if (end < prev) end = prev;
                        p.gate_t0[q][gi] = (uint16_t)prev; p.gate_t1[q][gi] = (uint16_t)end;
                        prev = end;
                    }
                    if (inst) base += PER_GATE + (uint64_t)inst * c;
                }
            }
            ZK_REQUIRE(g.lookup_width < 9, "prove: lookup width too large");
            p.lgamma_pow[0] = gl::make2(1, 0);
            for (uint32_t q = 1; q < 9; q++) p.lgamma_pow[q] = gl::mul(p.lgamma_pow[q - 1], lgamma);
        }
        for (uint32_t c = 0; c < QD; c++) {
            p.wit = cos_w.p + (size_t)c * N; p.setup = st.cosets.p + (size_t)c * N; p.s2 = cos_2.p + (size_t)c * N;
            p.shift = lde_coset_shift(log_n, log_qd, c);
            p.xn_minus_1 = gl::sub(gl::pow(p.shift, N), 1);
            p.zh_inv = gl::inv(p.xn_minus_1);
            p.l0_inv = l0_tab + (size_t)c * N;
            p.t0 = tq.p + (size_t)c * N; p.t1 = tq.p + QN + (size_t)c * N;
            launch_quotient_coset(ctx, p);
        }
        // interpolate over the big coset: bit-reversed -> natural, inverse NTT (two columns), undo shift, split
        uint64_t* nat = tq.p + 2 * QN;
        bitrev_copy(ctx, tq.p, QN, nat, QN, log_n + log_qd, 2);
        ntt_inverse(ctx, nat, QN, nat, QN, tq.p, QN, log_n + log_qd, 2);
        qmono.alloc((size_t)Q * N, stream);
        quotient_split_kernel<<<(unsigned)((QN + 255) / 256), 256, 0, stream>>>(nat, nat + QN, qmono.p, log_n, QN, gl::inv(GL_GEN));
        LAUNCH_CHECK(ctx);
    }
    tq.release();
    cos_q.alloc((size_t)Q * LN, stream);
    tree_q.alloc(merkle_tree_digests(LN, cap) * 4, stream);
    ntt_forward_cosets(ctx, qmono.p, N, cos_q.p, LN, log_n, (int)Q, (int)cfg.log_lde, 0, L);
    merkle_build(ctx, cos_q.p, LN, Q, LN, 1, cap, tree_q.p);
    d2h(ctx, cap_q.data(), tree_q.p + 4 * merkle_cap_offset(LN, cap), cap * 32);
    tr.absorb(cap_q.data(), cap * 4);
    gl::e2 z = tr.challenge_ext();

    // ---- openings
    std::vector<gl::e2> at_z(sh.n_at_z), at_0(sh.n_at_0 ? sh.n_at_0 : 1);
    gl::e2 at_zw;
    {
        DevBuf pw, partial, dout;
        pw.alloc(2 * N, stream);
        const uint32_t max_cols = W > S ? W : S;
        partial.alloc(2 * 32 * (size_t)(max_cols > S2 ? max_cols : S2), stream);
        dout.alloc(2 * (size_t)(max_cols > S2 ? max_cols : S2), stream);
        ext_pow_table_kernel<<<(unsigned)((N / 16 + 127) / 128 + 1), 128, 0, stream>>>(pw.p, N, z);
        LAUNCH_CHECK(ctx);
        std::vector<gl::e2> tmp(S2 > Q ? S2 : Q);
        eval_columns(ctx, mono_w.p, N, W, N, pw.p, partial.p, dout.p, at_z.data());
        eval_columns(ctx, st.mono.p, N, S, N, pw.p, partial.p, dout.p, at_z.data() + W);
        eval_columns(ctx, mono_2.p, N, S2, N, pw.p, partial.p, dout.p, tmp.data());
        for (uint32_t e = 0; e < sh.E2; e++) {  // f0(z) + u*f1(z)
            gl::e2 a = tmp[2 * e], b = tmp[2 * e + 1];
            at_z[W + S + e] = gl::make2(gl::add(a.c0, gl::mul(7, b.c1)), gl::add(a.c1, b.c0));
        }
        eval_columns(ctx, qmono.p, N, Q, N, pw.p, partial.p, dout.p, tmp.data());
        for (uint32_t e = 0; e < QD; e++) {
            gl::e2 a = tmp[2 * e], b = tmp[2 * e + 1];
            at_z[W + S + sh.E2 + e] = gl::make2(gl::add(a.c0, gl::mul(7, b.c1)), gl::add(a.c1, b.c0));
        }
        gl::e2 zw = gl::mul_base(z, gl::omega(log_n));
        ext_pow_table_kernel<<<(unsigned)((N / 16 + 127) / 128 + 1), 128, 0, stream>>>(pw.p, N, zw);
        LAUNCH_CHECK(ctx);
        eval_columns(ctx, mono_2.p, N, 2, N, pw.p, partial.p, dout.p, tmp.data());
        at_zw = gl::make2(gl::add(tmp[0].c0, gl::mul(7, tmp[1].c1)), gl::add(tmp[0].c1, tmp[1].c0));
        for (uint32_t i = 0; i < sh.n_at_0; i++) {
            CUDA_CHECK(cudaMemcpyAsync(&at_0[i].c0, mono_2.p + (size_t)(2 * (sh.C + i)) * N, 8, cudaMemcpyDeviceToHost, stream));
            CUDA_CHECK(cudaMemcpyAsync(&at_0[i].c1, mono_2.p + (size_t)(2 * (sh.C + i) + 1) * N, 8, cudaMemcpyDeviceToHost, stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    // at_z is kept in oracle order (witness, setup, stage 2, quotient); the proof and the transcript carry the reference's order
    const std::vector<uint32_t> open_pos = opening_positions(g, sh);
    std::vector<gl::e2> at_z_out(sh.n_at_z);
    for (uint32_t i = 0; i < sh.n_at_z; i++) at_z_out[open_pos[i]] = at_z[i];
    tr.absorb(reinterpret_cast<const uint64_t*>(at_z_out.data()), 2 * sh.n_at_z);
    tr.absorb(at_zw);
    tr.absorb(reinterpret_cast<const uint64_t*>(at_0.data()), 2 * sh.n_at_0);
    gl::e2 phi = tr.challenge_ext();

    // ---- DEEP
    const uint32_t NF = sh.NF;
    std::vector<DevBuf> fri(NF + 1);   // fri[k]: pair buffer (c0 | c1) of oracle k's domain; fri[NF] = final values
    std::vector<DevBuf> fri_tree(NF);
    fri[0].alloc(2 * LN, stream);
    {
        // weights in the kernel's (oracle) order: phi^(position in values_at_z), then z*omega, the openings at 0, the public inputs
        const uint32_t n_deep = sh.n_at_z + 1 + sh.n_at_0 + g.n_public_inputs;
        std::vector<gl::e2> phi_pow(n_deep);
        phi_pow[0] = gl::make2(1, 0);
        for (uint32_t i = 1; i < n_deep; i++) phi_pow[i] = gl::mul(phi_pow[i - 1], phi);
        std::vector<uint64_t> phip(2 * (size_t)n_deep);
        gl::e2 sum_at_z = gl::make2(0, 0);
        for (uint32_t i = 0; i < n_deep; i++) {
            const gl::e2 a = i < sh.n_at_z ? phi_pow[open_pos[i]] : phi_pow[i];
            phip[2 * i] = a.c0; phip[2 * i + 1] = a.c1;
            if (i < sh.n_at_z) sum_at_z = gl::add(sum_at_z, gl::mul(a, at_z[i]));
        }
        DevBuf d_phip, d_at0;
        d_phip.alloc(phip.size(), stream);
        h2d(ctx, d_phip.p, phip.data(), phip.size() * 8);
        d_at0.alloc(2 * (size_t)(sh.n_at_0 ? sh.n_at_0 : 1), stream);
        if (sh.n_at_0) h2d(ctx, d_at0.p, at_0.data(), (size_t)sh.n_at_0 * 16);
        DeepParams p{};
        p.wit = cos_w.p; p.setup = st.cosets.p; p.s2 = cos_2.p; p.q = cos_q.p;
        p.cs_w = (size_t)E * N; p.cs_s = (size_t)st.E * N; p.cs_2 = (size_t)E * N; p.cs_q = LN;
        p.W = W; p.S = S; p.E2 = sh.E2; p.QD = QD; p.C = sh.C; p.n_at_0 = sh.n_at_0; p.log_ln = g.log_n + cfg.log_lde;
        p.phip = d_phip.p; p.at_0 = d_at0.p; p.sum_at_z = sum_at_z; p.at_zw = at_zw; p.z = z; p.zw = gl::mul_base(z, gl::omega(log_n));
        p.omega_ln = gl::omega((int)p.log_ln);
        p.n_pi = g.n_public_inputs;
        for (uint32_t i = 0; i < g.n_public_inputs; i++) {
            p.pi_col[i] = g.pi_col[i]; p.pi_val[i] = pi[i]; p.pi_root[i] = gl::pow(gl::omega(log_n), g.pi_row[i]);
        }
        p.f0 = fri[0].p; p.f1 = fri[0].p + LN;
        deep_kernel<<<(unsigned)((LN + 127) / 128), 128, 0, stream>>>(p);
        LAUNCH_CHECK(ctx);
        CUDA_CHECK(cudaStreamSynchronize(stream));  // phip / at_0 host vectors go out of scope
    }

    // ---- FRI commit phase
    std::vector<std::vector<uint64_t>> fri_caps(NF);
    {
        uint64_t shift = GL_GEN;
        for (uint32_t k = 0; k < NF; k++) {
            int ld = (int)sh.fri_dom_log[k];
            size_t D = (size_t)1 << ld;
            fri_tree[k].alloc(merkle_tree_digests(sh.fri_leaves[k], sh.fri_cap[k]) * 4, stream);
            merkle_build(ctx, fri[k].p, D, 2, sh.fri_leaves[k], (size_t)1 << cfg.fri_schedule[k], sh.fri_cap[k], fri_tree[k].p);
            fri_caps[k].resize(sh.fri_cap[k] * 4);
            d2h(ctx, fri_caps[k].data(), fri_tree[k].p + 4 * merkle_cap_offset(sh.fri_leaves[k], sh.fri_cap[k]), sh.fri_cap[k] * 32);
            tr.absorb(fri_caps[k].data(), sh.fri_cap[k] * 4);
            gl::e2 c = tr.challenge_ext();
            const uint64_t* in = fri[k].p;
            size_t in_d = D;
            DevBuf tmp_a, tmp_b;
            for (uint32_t stp = 0; stp < cfg.fri_schedule[k]; stp++) {
                size_t out_d = in_d >> 1;
                uint64_t* out;
                if (stp + 1 == cfg.fri_schedule[k]) { fri[k + 1].alloc(2 * out_d, stream); out = fri[k + 1].p; }
                else { DevBuf& t = (stp & 1) ? tmp_b : tmp_a; t.alloc(2 * out_d, stream); out = t.p; }
                fri_fold(ctx, in, in + in_d, ld, shift, c, out, out + out_d);
                in = out; in_d = out_d;
                c = gl::sqr(c); shift = gl::sqr(shift); ld--;
            }
        }
        // final polynomial (tiny): values -> monomials on the host
        const int ldf = (int)sh.fri_dom_log[NF];
        const size_t DF = (size_t)1 << ldf;
        std::vector<uint64_t> fv(2 * DF);
        d2h(ctx, fv.data(), fri[NF].p, 2 * DF * 8);
        std::vector<uint64_t> fin0(DF), fin1(DF);
        uint64_t w_inv = gl::inv(gl::omega(ldf)), n_inv = gl::inv((uint64_t)DF % GL_P), s_inv = gl::inv(shift);
        // inverse NTT on the host (DF <= final degree x LDE factor: 16 for the base layer, 2048 for compression mode 4):
        // bit-reversed values -> natural order, radix-2 decimation in time with w^-1, then n^-1 and the coset shift undone
        {
            std::vector<uint64_t> tw(DF / 2 ? DF / 2 : 1);
            uint64_t x = 1;
            for (size_t j = 0; j < DF / 2; j++) { tw[j] = x; x = gl::mul(x, w_inv); }
            for (int half = 0; half < 2; half++) {
                std::vector<uint64_t>& a = half ? fin1 : fin0;
                // DIT consumes its input in bit-reversed order: the stored order is already that
                for (size_t pos = 0; pos < DF; pos++) a[pos] = fv[(size_t)half * DF + pos];
                for (size_t len = 2; len <= DF; len <<= 1) {
                    const size_t step = DF / len;
                    for (size_t blk = 0; blk < DF; blk += len)
                        for (size_t j = 0; j < len / 2; j++) {
                            uint64_t u = a[blk + j], v = gl::mul(a[blk + j + len / 2], tw[j * step]);
                            a[blk + j] = gl::add(u, v);
                            a[blk + j + len / 2] = gl::sub(u, v);
                        }
                }
                uint64_t sc = n_inv;
                for (size_t i = 0; i < DF; i++) { a[i] = gl::mul(a[i], sc); sc = gl::mul(sc, s_inv); }
            }
        }
        for (size_t i = sh.n_final; i < DF; i++)
            ZK_REQUIRE(fin0[i] == 0 && fin1[i] == 0, "prove: final FRI polynomial exceeds its degree bound -- the trace does not satisfy the circuit");

        // ---- assemble the proof (DESIGN.md "Proof buffer")
        uint64_t* p = proof;
        memset(p, 0, 32 * 8);
        p[0] = PROOF_MAGIC; p[1] = g.log_n; p[2] = cfg.log_lde; p[3] = cap; p[4] = cfg.n_queries; p[5] = NF; p[6] = W; p[7] = S2; p[8] = Q; p[9] = S;
        p[10] = sh.n_at_z; p[11] = sh.n_at_zw; p[12] = sh.n_at_0; p[13] = g.n_public_inputs; p[14] = sh.n_final; p[15] = cfg.pow_bits;
        for (uint32_t k = 0; k < NF; k++) p[16 + k] = cfg.fri_schedule[k];
        p += 32;
        memcpy(p, pi.data(), g.n_public_inputs * 8); p += g.n_public_inputs;
        memcpy(p, cap_w.data(), cap * 32); p += cap * 4;
        memcpy(p, cap_2.data(), cap * 32); p += cap * 4;
        memcpy(p, cap_q.data(), cap * 32); p += cap * 4;
        memcpy(p, fin0.data(), sh.n_final * 8); p += sh.n_final;
        memcpy(p, fin1.data(), sh.n_final * 8); p += sh.n_final;
        memcpy(p, at_z_out.data(), (size_t)sh.n_at_z * 16); p += 2 * sh.n_at_z;
        memcpy(p, &at_zw, 16); p += 2;
        memcpy(p, at_0.data(), (size_t)sh.n_at_0 * 16); p += 2 * sh.n_at_0;
        for (uint32_t k = 0; k < NF; k++) { memcpy(p, fri_caps[k].data(), sh.fri_cap[k] * 32); p += sh.fri_cap[k] * 4; }
        tr.absorb(fin0.data(), sh.n_final);
        tr.absorb(fin1.data(), sh.n_final);

        // ---- queries
        const uint32_t NQ = cfg.n_queries;
        size_t per_q = W + S2 + Q + S + 4 * sh.depth * 4;
        for (uint32_t k = 0; k < NF; k++) per_q += 2 * ((size_t)1 << cfg.fri_schedule[k]) + sh.fri_depth[k] * 4;
        std::vector<uint32_t> idx((size_t)NQ * (NF + 1));
        for (uint32_t q = 0; q < NQ; q++) {
            size_t di = tr.query_index(ilog2(LN));
            idx[q] = (uint32_t)di;
            for (uint32_t k = 0; k < NF; k++) { di >>= cfg.fri_schedule[k]; idx[(size_t)(k + 1) * NQ + q] = (uint32_t)di; }
        }
        DevBuf d_idx, d_q;
        d_idx.alloc((idx.size() + 1) / 2, stream);
        h2d(ctx, d_idx.p, idx.data(), idx.size() * 4);
        d_q.alloc(per_q * NQ, stream);
        const uint32_t* di32 = reinterpret_cast<const uint32_t*>(d_idx.p);
        size_t off = 0;
        struct Src { const uint64_t* cols; size_t stride; uint32_t n; const uint64_t* tree; };
        Src srcs[4] = {{cos_w.p, (size_t)E * N, W, tree_w.p}, {cos_2.p, (size_t)E * N, S2, tree_2.p}, {cos_q.p, LN, Q, tree_q.p},
                       {st.cosets.p, (size_t)st.E * N, S, st.tree.p}};
        for (int o = 0; o < 4; o++) {
            gather_kernel<<<NQ, 128, 0, stream>>>(srcs[o].cols, srcs[o].stride, srcs[o].n, 1, srcs[o].tree, LN, (uint32_t)sh.depth, di32, d_q.p, per_q, off);
            LAUNCH_CHECK(ctx);
            off += srcs[o].n + sh.depth * 4;
        }
        for (uint32_t k = 0; k < NF; k++) {
            size_t D = (size_t)1 << sh.fri_dom_log[k];
            uint32_t epl = 1u << cfg.fri_schedule[k];
            gather_kernel<<<NQ, 64, 0, stream>>>(fri[k].p, D, 2, epl, fri_tree[k].p, sh.fri_leaves[k], (uint32_t)sh.fri_depth[k],
                                                 di32 + (size_t)(k + 1) * NQ, d_q.p, per_q, off);
            LAUNCH_CHECK(ctx);
            off += 2 * epl + sh.fri_depth[k] * 4;
        }
        d2h(ctx, p, d_q.p, per_q * NQ * 8);
        p += per_q * NQ;
        *p++ = 0;  // pow_challenge (NoPow)
        ZK_REQUIRE((size_t)(p - proof) == sh.proof_len, "prove: internal proof length mismatch");
    }
}

}  // namespace zk

struct zkgpu_ctx {
    zk::Ctx c;
};
struct zkgpu_setup {
    zk::Setup* s;
};

extern "C" {

int zkgpu_setup_create(zkgpu_ctx* ctx, const zkgpu_geometry* g, const zkgpu_proof_config* cfg, const uint64_t* h_setup_cols, zkgpu_setup** out,
                       uint64_t* h_vk_cap_out) {
    try {
        ZK_REQUIRE(ctx && g && cfg && h_setup_cols && out, "setup_create: NULL argument");
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::Setup* s = zk::setup_create(&ctx->c, *g, *cfg, h_setup_cols, h_vk_cap_out);
        *out = new zkgpu_setup{s};
        return 0;
    } catch (const zk::Error& e) {
        zk::g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 99;
    }
}
void zkgpu_setup_destroy(zkgpu_setup* s) {
    if (!s) return;
    if (s->s) {
        cudaSetDevice(s->s->device);
        delete s->s;
    }
    delete s;
}
int zkgpu_prove_device(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* d_witness_cols, uint64_t* h_proof_out, size_t proof_capacity_u64) {
    try {
        ZK_REQUIRE(ctx && s && s->s && d_witness_cols && h_proof_out, "prove: NULL argument");
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::prove(&ctx->c, *s->s, d_witness_cols, h_proof_out, proof_capacity_u64);
        return 0;
    } catch (const zk::Error& e) {
        zk::g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 99;
    }
}
int zkgpu_prove(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* h_witness_cols, uint64_t* h_proof_out, size_t proof_capacity_u64) {
    cudaStream_t up = nullptr;
    zk::UploadPlan plan;
    cudaEvent_t allocated = nullptr;
    int rc = 0;
    try {
        ZK_REQUIRE(ctx && s && s->s && h_witness_cols && h_proof_out, "prove: NULL argument");
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::ArenaScope arena_scope(&ctx->c, zk::prove_scratch_bytes(*s->s, true));
        zk::DevBuf wit;
        const size_t N = s->s->sh.N;
        const uint32_t W = s->s->sh.W;
        wit.alloc((size_t)W * N, ctx->c.stream);
        // upload in (up to) 8 column chunks on a side stream: the NTTs of chunk k overlap the PCIe transfer of chunk k+1
        plan.chunk_cols = W >= 16 ? (W + 7) / 8 : W;
        const uint32_t n_chunks = (W + plan.chunk_cols - 1) / plan.chunk_cols;
        CUDA_CHECK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&allocated, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventRecord(allocated, ctx->c.stream));
        CUDA_CHECK(cudaStreamWaitEvent(up, allocated, 0));   // the stream-ordered allocation of `wit` comes first
        for (uint32_t k = 0; k < n_chunks; k++) {
            const uint32_t c0 = k * plan.chunk_cols, c1 = c0 + plan.chunk_cols < W ? c0 + plan.chunk_cols : W;
            CUDA_CHECK(cudaMemcpyAsync(wit.p + (size_t)c0 * N, h_witness_cols + (size_t)c0 * N, (size_t)(c1 - c0) * N * 8, cudaMemcpyHostToDevice, up));
            cudaEvent_t ev;
            CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            plan.ready.push_back(ev);
            CUDA_CHECK(cudaEventRecord(ev, up));
        }
        zk::prove(&ctx->c, *s->s, wit.p, h_proof_out, proof_capacity_u64, &plan);
    } catch (const zk::Error& e) {
        zk::g_last_error = e.what();
        rc = e.code;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        rc = 99;
    }
    if (up) { cudaStreamSynchronize(up); cudaStreamDestroy(up); }
    for (cudaEvent_t e : plan.ready) cudaEventDestroy(e);
    if (allocated) cudaEventDestroy(allocated);
    return rc;
}

int zkgpu_witness_stage(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* h_witness_cols, int slot) {
    try {
        ZK_REQUIRE(ctx && s && s->s && h_witness_cols, "witness_stage: NULL argument");
        ZK_REQUIRE(slot == 0 || slot == 1, "witness_stage: slot must be 0 or 1");
        zk::Ctx& c = ctx->c;
        CUDA_CHECK(cudaSetDevice(c.device));
        const size_t words = (size_t)s->s->sh.W * s->s->sh.N;
        if (!c.copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++)
            if (!c.staged_ready[k]) {
                CUDA_CHECK(cudaEventCreateWithFlags(&c.staged_ready[k], cudaEventDisableTiming));
                CUDA_CHECK(cudaEventCreateWithFlags(&c.staged_free[k], cudaEventDisableTiming));
                CUDA_CHECK(cudaEventRecord(c.staged_free[k], c.stream));
            }
        if (c.staged_words[slot] < words) {   // both slots grow together to the widest circuit seen; steady state allocates nothing
            CUDA_CHECK(cudaStreamSynchronize(c.stream));
            CUDA_CHECK(cudaStreamSynchronize(c.copy_stream));
            for (int k = 0; k < 2; k++) {
                if (c.staged_words[k] >= words) continue;
                ZK_REQUIRE(k == slot || c.staged_words[k] == 0 || cudaEventQuery(c.staged_ready[k]) == cudaSuccess,
                           "witness_stage: cannot grow a slot that holds a pending witness");
                uint64_t* fresh = nullptr;
                CUDA_CHECK(cudaMalloc((void**)&fresh, words * 8));
                if (c.staged[k]) {   // keep a witness already staged in the other slot
                    CUDA_CHECK(cudaMemcpy(fresh, c.staged[k], c.staged_words[k] * 8, cudaMemcpyDeviceToDevice));
                    CUDA_CHECK(cudaFree(c.staged[k]));
                }
                c.staged[k] = fresh;
                c.staged_words[k] = words;
            }
        }
        CUDA_CHECK(cudaStreamWaitEvent(c.copy_stream, c.staged_free[slot], 0));   // the proof that last read this slot is done
        {
            const uint32_t W = s->s->sh.W;
            const size_t N = s->s->sh.N;
            const uint32_t chunk_cols = W >= 16 ? (W + 7) / 8 : W, n_chunks = (W + chunk_cols - 1) / chunk_cols;
            for (uint32_t k = 0; k < n_chunks; k++) {
                const uint32_t c0 = k * chunk_cols, c1 = c0 + chunk_cols < W ? c0 + chunk_cols : W;
                CUDA_CHECK(cudaMemcpyAsync(c.staged[slot] + (size_t)c0 * N, h_witness_cols + (size_t)c0 * N, (size_t)(c1 - c0) * N * 8,
                                           cudaMemcpyHostToDevice, c.copy_stream));
                if (!c.staged_chunk[slot][k]) CUDA_CHECK(cudaEventCreateWithFlags(&c.staged_chunk[slot][k], cudaEventDisableTiming));
                CUDA_CHECK(cudaEventRecord(c.staged_chunk[slot][k], c.copy_stream));
            }
            c.staged_chunk_cols[slot] = chunk_cols; c.staged_n_chunks[slot] = n_chunks;
        }
        CUDA_CHECK(cudaEventRecord(c.staged_ready[slot], c.copy_stream));
        c.staged_valid[slot] = true;
        return 0;
    } catch (const zk::Error& e) {
        zk::g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 99;
    }
}

int zkgpu_prove_staged(zkgpu_ctx* ctx, const zkgpu_setup* s, int slot, uint64_t* h_proof_out, size_t proof_capacity_u64) {
    try {
        ZK_REQUIRE(ctx && s && s->s && h_proof_out, "prove_staged: NULL argument");
        ZK_REQUIRE(slot == 0 || slot == 1, "prove_staged: slot must be 0 or 1");
        zk::Ctx& c = ctx->c;
        ZK_REQUIRE(c.staged_valid[slot] && c.staged_words[slot] >= (size_t)s->s->sh.W * s->s->sh.N, "prove_staged: no witness staged in this slot");
        c.staged_valid[slot] = false;
        CUDA_CHECK(cudaSetDevice(c.device));
        if (cudaEventQuery(c.staged_ready[slot]) == cudaSuccess) {   // the witness has landed: one batched pass over all columns
            CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.staged_ready[slot], 0));
            zk::prove(&c, *s->s, c.staged[slot], h_proof_out, proof_capacity_u64);
        } else {   // still crossing PCIe (first proof of a run): start on the column chunks as they land, like zkgpu_prove
            (void)cudaGetLastError();   // cudaErrorNotReady from the query is not an error
            zk::UploadPlan plan;
            plan.chunk_cols = c.staged_chunk_cols[slot];
            for (uint32_t k = 0; k < c.staged_n_chunks[slot]; k++) plan.ready.push_back(c.staged_chunk[slot][k]);
            zk::prove(&c, *s->s, c.staged[slot], h_proof_out, proof_capacity_u64, &plan);
        }
        CUDA_CHECK(cudaEventRecord(c.staged_free[slot], c.stream));
        CUDA_CHECK(cudaStreamSynchronize(c.stream));
        return 0;
    } catch (const zk::Error& e) {
        zk::g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 99;
    }
}

static int64_t upload_index_map(zk::Ctx& c, zk::DevBuf& dst, const uint32_t* h_maps, size_t cells) {
    dst.alloc((cells + 1) / 2, c.stream);   // u32 cells in a u64 buffer
    CUDA_CHECK(cudaMemcpyAsync(dst.p, h_maps, cells * 4, cudaMemcpyHostToDevice, c.stream));
    int64_t mx = -1;
    for (size_t i = 0; i < cells; i++)
        if (h_maps[i] != ZKGPU_VAR_PLACEHOLDER && (int64_t)h_maps[i] > mx) mx = h_maps[i];
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    return mx;
}

int zkgpu_setup_set_variable_maps(zkgpu_ctx* ctx, zkgpu_setup* s, const uint32_t* h_var_maps) {
    try {
        ZK_REQUIRE(ctx && s && s->s && h_var_maps, "set_variable_maps: NULL argument");
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::Setup& st = *s->s;
        st.max_var_index = upload_index_map(ctx->c, st.var_maps, h_var_maps, (size_t)st.sh.NP * st.sh.N);
        return 0;
    } catch (const zk::Error& e) {
        zk::g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 99;
    }
}

int zkgpu_setup_set_witness_maps(zkgpu_ctx* ctx, zkgpu_setup* s, const uint32_t* h_wit_maps) {
    try {
        ZK_REQUIRE(ctx && s && s->s && h_wit_maps, "set_witness_maps: NULL argument");
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        zk::Setup& st = *s->s;
        ZK_REQUIRE(st.g.n_witness_plain > 0, "set_witness_maps: the circuit has no plain witness columns");
        st.max_wit_index = upload_index_map(ctx->c, st.wit_maps, h_wit_maps, (size_t)st.g.n_witness_plain * st.sh.N);
        return 0;
    } catch (const zk::Error& e) {
        zk::g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 99;
    }
}

int zkgpu_prove_from_hints(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* h_variable_values, size_t n_vars,
                           const uint64_t* h_witness_values, size_t n_wits, const uint64_t* h_multiplicities, uint64_t* h_proof_out,
                           size_t proof_capacity_u64) {
    try {
        ZK_REQUIRE(ctx && s && s->s && h_variable_values && h_proof_out, "prove_from_hints: NULL argument");
        CUDA_CHECK(cudaSetDevice(ctx->c.device));
        const zk::Setup& st = *s->s;
        ZK_REQUIRE(st.var_maps.p != nullptr, "prove_from_hints: call zkgpu_setup_set_variable_maps first");
        ZK_REQUIRE(n_vars < ZKGPU_VAR_PLACEHOLDER && n_wits < ZKGPU_VAR_PLACEHOLDER, "prove_from_hints: too many values");
        ZK_REQUIRE(st.max_var_index < (int64_t)n_vars, "prove_from_hints: the variable maps reference an index >= n_vars");
        ZK_REQUIRE(st.g.lookup_reps == 0 || h_multiplicities != nullptr, "prove_from_hints: lookup circuit needs multiplicities");
        const uint32_t n_plain = st.g.n_witness_plain;
        if (n_plain) {
            ZK_REQUIRE(st.wit_maps.p != nullptr && h_witness_values != nullptr,
                       "prove_from_hints: circuits with plain witness columns (compression modes 1-3) need zkgpu_setup_set_witness_maps "
                       "and the witness values");
            ZK_REQUIRE(st.max_wit_index < (int64_t)n_wits, "prove_from_hints: the witness maps reference an index >= n_wits");
        }
        const size_t N = st.sh.N, cells = (size_t)st.sh.NP * N, wcells = (size_t)n_plain * N;
        zk::ArenaScope arena_scope(&ctx->c, zk::prove_scratch_bytes(st, true) + (n_vars + n_wits) * 8 + 64);
        zk::DevBuf wit, vals, wvals;
        wit.alloc((size_t)st.sh.W * N, ctx->c.stream);
        vals.alloc(n_vars ? n_vars : 1, ctx->c.stream);
        CUDA_CHECK(cudaMemcpyAsync(vals.p, h_variable_values, n_vars * 8, cudaMemcpyHostToDevice, ctx->c.stream));
        if (st.g.lookup_reps)   // the multiplicity column is the last witness column
            CUDA_CHECK(cudaMemcpyAsync(wit.p + (size_t)(st.sh.W - 1) * N, h_multiplicities, N * 8, cudaMemcpyHostToDevice, ctx->c.stream));
        zk::materialize_columns_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->c.stream>>>(
            reinterpret_cast<const uint32_t*>(st.var_maps.p), vals.p, n_vars, wit.p, cells);
        CUDA_CHECK(cudaGetLastError());
        ctx->c.kernel_launches++;
        if (n_plain) {   // plain witness columns follow the copy-permuted ones (zk_internal Shape::plain_col0 == NP)
            wvals.alloc(n_wits ? n_wits : 1, ctx->c.stream);
            CUDA_CHECK(cudaMemcpyAsync(wvals.p, h_witness_values, n_wits * 8, cudaMemcpyHostToDevice, ctx->c.stream));
            zk::materialize_columns_kernel<<<(unsigned)((wcells + 255) / 256), 256, 0, ctx->c.stream>>>(
                reinterpret_cast<const uint32_t*>(st.wit_maps.p), wvals.p, n_wits, wit.p + cells, wcells);
            CUDA_CHECK(cudaGetLastError());
            ctx->c.kernel_launches++;
        }
        zk::prove(&ctx->c, st, wit.p, h_proof_out, proof_capacity_u64);
        return 0;
    } catch (const zk::Error& e) {
        zk::g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        zk::g_last_error = e.what();
        return 99;
    }
}

int zkgpu_prove_from_variables(zkgpu_ctx* ctx, const zkgpu_setup* s, const uint64_t* h_variable_values, size_t n_vars,
                               const uint64_t* h_multiplicities, uint64_t* h_proof_out, size_t proof_capacity_u64) {
    if (s && s->s && s->s->g.n_witness_plain) {
        zk::g_last_error = "prove_from_variables: circuits with plain witness columns (compression modes 1-3) also need the witness "
                           "hint; use zkgpu_prove_from_hints";
        return 2;
    }
    return zkgpu_prove_from_hints(ctx, s, h_variable_values, n_vars, nullptr, 0, h_multiplicities, h_proof_out, proof_capacity_u64);
}
}
