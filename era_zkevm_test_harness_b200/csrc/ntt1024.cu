// ntt1024.cu -- the 2^20-point batched Goldilocks NTT / iNTT / coset NTT as two passes of 1024-point warp transforms.
//
// Replaces the FFT/LDE work of boojum's `prove_from_precomputations` / `get_full_setup`
// (/root/reference/src/prover_utils.rs:338, :186) at the size every base- and recursion-layer circuit uses (domain 2^20).
//
//   n = 1024 x 1024, index i = i1*1024 + i0, k = k1 + 1024*k0
//   pass A  (tile = 8 adjacent i0, staged through shared memory so global accesses are 64-byte segments):
//           one WARP per i0: 1024-point transform over i1 (ntt1024_core.cuh: 32 x 32 in registers, power-of-two twiddles),
//           times the inter-pass factor omega_n^(i0*k1) (* shift^i0 on a coset, * 1/n backwards) from an 8 MB L2-resident table
//   pass B  one warp per row k1 (8 KB contiguous): 1024-point transform over i0
// Forward: natural monomials -> BIT-REVERSED evaluations on shift*<omega_n> (in place capable);
// inverse: natural evaluations -> natural monomials (second pass writes transposed).
// General multiplications per element: 3 in pass A (coset pre-scale by a per-b constant, 32x32 twiddle, inter-pass factor),
// 1 in pass B; everything else is add/sub and shifts.
#define ZK_CANON_SEL 0   // see gl.cuh canon(): the select form costs pass A more in spills than it saves
#include "ntt1024_core.cuh"
#include <mutex>
#include "zk_internal.cuh"

namespace zk {

static constexpr int NT_T = 8;        // members (warps) per CTA
static constexpr int NT_SP = 1060;    // member pitch in shared memory (>= 33*32, = 4 mod 16 so staging writes spread over banks)

struct Ntt1024Params {
    const uint64_t* in;
    uint64_t* out;
    size_t in_stride, out_stride;   // distance between polynomials of the batch
    const uint64_t* twid;           // [32*32]  rho^(a*kb) (* coset factor of lane a), index kb*32 + a
    const uint64_t* pre;            // [32] or null: u[a + 32*b] *= pre[b] before step 1
    const uint64_t* post;           // [1024*1024] or null: X[k] *= post[member*1024 + k]
};

template <bool IN_STRIDED, bool OUT_STRIDED, bool OUT_NATURAL, bool INVERSE>
__global__ void __launch_bounds__(32 * NT_T, 2) ntt1024_kernel(const __grid_constant__ Ntt1024Params p) {
    constexpr int E32 = INVERSE ? NTT32_E_INV : NTT32_E_FWD;
    extern __shared__ __align__(128) uint64_t smem[];
    __shared__ __align__(8) uint64_t bars[NT_T];   // one mbarrier per warp (pass B bulk loads)
    uint64_t* twid_s = smem;                 // 1024
    uint64_t* pre_s = smem + 1024;           // 32
    uint64_t* data_all = smem + 1024 + 32;   // NT_T * NT_SP (member pitch 8480 B, 16-byte aligned)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t t0 = (size_t)blockIdx.x * NT_T;
    const size_t member = t0 + warp;
    const uint64_t* __restrict__ in = p.in + (size_t)blockIdx.y * p.in_stride;
    uint64_t* __restrict__ out = p.out + (size_t)blockIdx.y * p.out_stride;
    uint64_t* data = data_all + warp * NT_SP;

#pragma unroll
    for (int i = 0; i < 4; i++) twid_s[tid + 256 * i] = p.twid[tid + 256 * i];
    if (p.pre != nullptr && tid < 32) pre_s[tid] = p.pre[tid];

    uint64_t v[32];
    if constexpr (IN_STRIDED) {
        // element e of member m lives at in[e*1024 + t0 + m]: a warp instruction reads 4 rows x 64 bytes
#pragma unroll 8
        for (int kk = 0; kk < 32; kk++) {
            const int idx = tid + 256 * kk, m = idx & (NT_T - 1), e = idx >> 3;
            data_all[m * NT_SP + e + (e >> 5)] = in[((size_t)e << 10) + t0 + m];
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < 32; b++) v[b] = data[33 * b + lane];
        __syncwarp();
    } else {
        // the member's row is 8 KB contiguous: one TMA bulk copy (cp.async.bulk, SASS UBLKCP) per warp straight into the
        // warp's exchange buffer, completion on a per-warp mbarrier; the 32 lanes then pick their stride-32 elements from
        // shared memory (conflict-free).  Unaligned callers (row not 16-byte aligned) take plain coalesced loads.
        const uint64_t* row = in + (member << 10);
        if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp]);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(data);
            if (lane == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 8192;" ::"r"(bar) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 8192, [%2];"
                             ::"r"(dst), "l"(row), "r"(bar) : "memory");
            }
            __syncwarp();
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                             : "=r"(done) : "r"(bar) : "memory");
            }
#pragma unroll
            for (int b = 0; b < 32; b++) v[b] = data[32 * b + lane];
            __syncwarp();
        } else {
#pragma unroll
            for (int b = 0; b < 32; b++) v[b] = row[32 * b + lane];
        }
        __syncthreads();   // twiddle table visible
    }
    if (p.pre != nullptr) {
#pragma unroll
        for (int b = 1; b < 32; b++) v[b] = gl::mul(v[b], pre_s[b]);
    }
    ntt1024_step1<E32>(v, lane, twid_s, data);
    __syncwarp();
    ntt1024_step2<E32>(v, lane, data);
    if (p.post != nullptr) {
        const uint64_t* __restrict__ post = p.post + (member << 10);
#pragma unroll
        for (int r = 0; r < 32; r++) v[r] = gl::mul(v[r], post[ntt1024_k(lane, r)]);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; r++) {
        const int o = OUT_NATURAL ? ntt1024_k(lane, r) : ntt1024_pos(lane, r);
        data[o + (o >> 5)] = v[r];
    }
    if constexpr (OUT_STRIDED) {
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; kk++) {
            const int idx = tid + 256 * kk, m = idx & (NT_T - 1), o = idx >> 3;
            out[((size_t)o << 10) + t0 + m] = data_all[m * NT_SP + o + (o >> 5)];
        }
    } else {
        __syncwarp();
        uint64_t* row = out + (member << 10);
#pragma unroll
        for (int c = 0; c < 32; c++) row[32 * c + lane] = data[33 * c + lane];
    }
}

// ---------------------------------------------------------------- tables
// post[i0*1024 + k1] = scale * shift^i0 * w^(i0*k1)
__global__ void ntt1024_post_table_kernel(uint64_t* out, uint64_t w, uint64_t shift, uint64_t scale) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;   // 2^20 entries
    const uint64_t i0 = idx >> 10, k1 = idx & 1023;
    out[idx] = gl::mul(gl::mul(scale, gl::pow(shift, i0)), gl::pow(w, i0 * k1));
}

struct Ntt1024Tables {
    uint64_t* twid_std[2] = {nullptr, nullptr};   // forward / inverse, no coset
    uint64_t* post_inv = nullptr;                 // w^-(i0*k1) / n
    struct Coset { uint64_t* twid; uint64_t* pre; uint64_t* post; };
    std::map<uint64_t, Coset> cosets;             // by shift (shift = 1: plain forward transform)
};
static std::map<Ctx*, Ntt1024Tables> g_tables;    // keyed by context; the device memory belongs to the context's persistent list
static std::mutex g_tables_mu;

static uint64_t* upload(Ctx* ctx, const std::vector<uint64_t>& h) {
    uint64_t* d = (uint64_t*)ctx->alloc_persistent(h.size() * 8);
    CUDA_CHECK(cudaMemcpyAsync(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return d;
}
static uint64_t* make_twid(Ctx* ctx, uint64_t rho, uint64_t lane_factor_base) {  // rho^(a*kb) * base^a
    std::vector<uint64_t> h(1024);
    for (int kb = 0; kb < 32; kb++)
        for (int a = 0; a < 32; a++) h[kb * 32 + a] = gl::mul(gl::pow(rho, (uint64_t)a * kb), gl::pow(lane_factor_base, a));
    return upload(ctx, h);
}
static uint64_t* make_post(Ctx* ctx, uint64_t w, uint64_t shift, uint64_t scale) {
    uint64_t* d = (uint64_t*)ctx->alloc_persistent((size_t)8 << 20);
    ntt1024_post_table_kernel<<<4096, 256, 0, ctx->stream>>>(d, w, shift, scale);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
    return d;
}

void ntt1024_forget(Ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_tables_mu);
    g_tables.erase(ctx);
}

static Ntt1024Tables& tables(Ctx* ctx) {
    Ntt1024Tables* tp;
    {
        std::lock_guard<std::mutex> lk(g_tables_mu);
        tp = &g_tables[ctx];   // std::map nodes are stable; a context is used by one thread at a time
    }
    Ntt1024Tables& t = *tp;
    if (!t.twid_std[0]) {
        ZK_REQUIRE(gl::pow(2, NTT32_E_FWD) == gl::omega(5) && gl::mul(gl::pow(2, NTT32_E_INV), gl::omega(5)) == 1,
                   "ntt1024: omega_32 is not the expected power of two");
        const uint64_t rho = gl::omega(10);
        t.twid_std[0] = make_twid(ctx, rho, 1);
        t.twid_std[1] = make_twid(ctx, gl::inv(rho), 1);
    }
    return t;
}
static const Ntt1024Tables::Coset& coset_tables(Ctx* ctx, uint64_t shift) {
    Ntt1024Tables& t = tables(ctx);
    auto it = t.cosets.find(shift);
    if (it != t.cosets.end()) return it->second;
    Ntt1024Tables::Coset c{};
    const uint64_t rho = gl::omega(10), w = gl::omega(20);
    if (shift != 1) {
        c.twid = make_twid(ctx, rho, gl::pow(shift, 1024));                 // * shift^(1024*a)
        std::vector<uint64_t> pre(32);
        const uint64_t sb = gl::pow(shift, 32 * 1024);
        for (int b = 0; b < 32; b++) pre[b] = gl::pow(sb, b);                // * shift^(1024*32*b)
        c.pre = upload(ctx, pre);
    } else {
        c.twid = t.twid_std[0];
        c.pre = nullptr;
    }
    c.post = make_post(ctx, w, shift, 1);                                    // * shift^i0 * w^(i0*k1)
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return t.cosets.emplace(shift, c).first->second;
}

template <bool A, bool B, bool C, bool D>
static void launch(Ctx* ctx, const Ntt1024Params& p, int n_polys) {
#ifndef ZK_NTT_SMEM_PAD
#define ZK_NTT_SMEM_PAD 0   // occupancy experiments: extra dynamic shared memory per CTA
#endif
    constexpr size_t smem = (size_t)(1024 + 32 + NT_T * NT_SP) * sizeof(uint64_t) + ZK_NTT_SMEM_PAD;
    static bool attr_set[64] = {};   // the attribute is per device: one process may hold contexts on several GPUs
    if (!attr_set[ctx->device & 63]) {
        CUDA_CHECK(cudaFuncSetAttribute(ntt1024_kernel<A, B, C, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ctx->device & 63] = true;
    }
    // blockIdx.y is limited to 65535 polynomials per launch, far above any batch here
    dim3 grid(1024 / NT_T, (unsigned)n_polys);
    ntt1024_kernel<A, B, C, D><<<grid, 32 * NT_T, smem, ctx->stream>>>(p);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
}

// natural monomials -> bit-reversed evaluations over shift*<omega_2^20>; in == out allowed
void ntt1024_forward_coset(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, int n_polys, uint64_t shift) {
    const Ntt1024Tables::Coset& c = coset_tables(ctx, shift);
    const Ntt1024Tables& t = tables(ctx);
    Ntt1024Params a{in, out, in_stride, out_stride, c.twid, c.pre, c.post};
    launch<true, true, false, false>(ctx, a, n_polys);
    Ntt1024Params b{out, out, out_stride, out_stride, t.twid_std[0], nullptr, nullptr};
    launch<false, false, false, false>(ctx, b, n_polys);
}

// natural evaluations over <omega_2^20> -> natural monomials; tmp: scratch of the shape of out (may alias neither)
void ntt1024_inverse(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, uint64_t* tmp, size_t tmp_stride,
                     int n_polys) {
    Ntt1024Tables& t = tables(ctx);
    if (!t.post_inv) {
        t.post_inv = make_post(ctx, gl::inv(gl::omega(20)), 1, gl::inv((uint64_t)1 << 20));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    Ntt1024Params a{in, tmp, in_stride, tmp_stride, t.twid_std[1], nullptr, t.post_inv};
    launch<true, true, true, true>(ctx, a, n_polys);
    Ntt1024Params b{tmp, out, tmp_stride, out_stride, t.twid_std[1], nullptr, nullptr};
    launch<false, true, true, true>(ctx, b, n_polys);
}

}  // namespace zk
