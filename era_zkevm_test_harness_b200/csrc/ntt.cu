// ntt.cu -- batched radix-2 Goldilocks NTT / iNTT / coset-LDE for sm_100a.
//
// Replaces the forward/inverse FFT and LDE that boojum runs on CPU threads inside
// `prove_from_precomputations` / `get_full_setup` (/root/reference/src/prover_utils.rs:338,186).
//
// Decomposition (n = n1*n2, n1 = 2^L1 rows, n2 = 2^L2 columns of the row-major matrix view x[i1*n2 + i0]):
//   X[k1 + n1*k0] = sum_i0 w_n^(i0 k1) w_n2^(i0 k0) [ sum_i1 x[n2 i1 + i0] w_n1^(i1 k1) ]
//   pass 1 ("strided" tile): T adjacent columns i0..i0+T-1, an n1-point transform over i1 per column, then the
//           twiddle w_n^(i0 k1); global accesses are T*8-byte segments.
//   pass 2 ("contiguous" tile): T rows, an n2-point transform over i0 per row.
// Each sub-transform is a decimation-in-frequency NTT in shared memory, done in register groups of up to 5 stages
// (32 values per thread), so a 2^20 NTT touches HBM twice per element-pass and the 2^10 sub-transforms need only
// two shared-memory round trips.  Forward: natural-order coefficients -> BIT-REVERSED evaluations (the storage
// order of every committed oracle), in place capable.  Inverse: natural-order evaluations -> natural-order
// coefficients (second pass writes transposed).
#include "zk_internal.cuh"

namespace zk {

static constexpr int RMAX = 5;  // stages per register group

__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 5); }

// One register group of R DIF stages starting at stage s0 of a 2^LB transform held in shared memory (padded).
template <int R>
__device__ __forceinline__ void dif_group(uint64_t* __restrict__ s, const uint64_t* __restrict__ tw, int LB, int s0, int item) {
    const int b_lo = LB - s0 - R;
    const int lo = item & ((1 << b_lo) - 1);
    const int hi = item >> b_lo;
    const int base = (hi << (b_lo + R)) | lo;
    uint64_t v[1 << R];
#pragma unroll
    for (int e = 0; e < (1 << R); e++) v[e] = s[pad_idx(base + (e << b_lo))];
#pragma unroll
    for (int q = 0; q < R; q++) {
        const int hr = 1 << (R - 1 - q);
        const int st = s0 + q;
        // per-stage table: tw_st[j] = w_{2^(LB-st)}^j, j < 2^(LB-1-st), stored at offset 2^LB - 2^(LB-st)
        const uint64_t* twst = tw + ((1 << LB) - (1 << (LB - st)));
#pragma unroll
        for (int e = 0; e < (1 << R); e++) {
            if (e & hr) continue;
            const int j = ((e & (hr - 1)) << b_lo) | lo;
            uint64_t a = v[e], b = v[e + hr];
            v[e] = gl::add(a, b);
            uint64_t d = gl::sub(a, b);
            v[e + hr] = (LB - 1 - st == 0) ? d : gl::mul(d, twst[j]);  // last stage: twiddle is 1
        }
    }
#pragma unroll
    for (int e = 0; e < (1 << R); e++) s[pad_idx(base + (e << b_lo))] = v[e];
}

struct NttPass {
    const uint64_t* in;
    uint64_t* out;
    size_t in_col_stride, out_col_stride;  // distance between polynomials of the batch
    int LB;            // log2 sub-transform size
    int LM;            // log2 of the other dimension (n = 2^(LB+LM))
    int T;             // tile members per CTA
    int in_strided;    // 1: element e of member t at e*2^LM + t ; 0: at t*2^LB + e
    int out_strided;   // same for the output
    int out_natural;   // 1: output index = natural k (read shared position bitrev(k)) ; 0: output index = position
    const uint64_t* tw;      // per-stage twiddle table for size 2^LB (2^LB - 1 entries)
    const uint64_t* pre_e;   // optional: multiply input element e by pre_e[e]
    const uint64_t* pre_t;   // optional: multiply input of member t (global id) by pre_t[t]
    const uint64_t* twA;     // optional post twiddle: out_k *= twA[x & maskA] * twB[x >> LA], x = t * k_natural
    const uint64_t* twB;
    int LA;
    uint64_t post_scale;     // optional (non-zero, != 1): multiply every output (used for n^-1 when no post twiddle)
    // batch over blockIdx.z (cosets): element offsets added per z
    size_t in_z_stride, out_z_stride, pre_e_z_stride, pre_t_z_stride;
};

__global__ void __launch_bounds__(512) ntt_pass_kernel(NttPass p) {
    extern __shared__ uint64_t smem[];
    const int n_sub = 1 << p.LB;
    const int SP = n_sub + (n_sub >> 5) + 1;
    uint64_t* tw_s = smem;                  // n_sub entries (n_sub - 1 used)
    uint64_t* data = smem + n_sub;          // T * SP
    const int tid = threadIdx.x, nthr = blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * p.T;  // first member (column or row id) of this tile
    const uint64_t* in = p.in + (size_t)blockIdx.y * p.in_col_stride + (size_t)blockIdx.z * p.in_z_stride;
    uint64_t* out = p.out + (size_t)blockIdx.y * p.out_col_stride + (size_t)blockIdx.z * p.out_z_stride;
    const uint64_t* __restrict__ pre_e = p.pre_e ? p.pre_e + (size_t)blockIdx.z * p.pre_e_z_stride : nullptr;
    const uint64_t* __restrict__ pre_t = p.pre_t ? p.pre_t + (size_t)blockIdx.z * p.pre_t_z_stride : nullptr;

    for (int i = tid; i < n_sub - 1; i += nthr) tw_s[i] = p.tw[i];

    const int total = p.T << p.LB;
    if (p.in_strided) {
        for (int idx = tid; idx < total; idx += nthr) {
            int t = idx % p.T, e = idx / p.T;
            uint64_t v = in[((size_t)e << p.LM) + t0 + t];
            if (pre_e) v = gl::mul(v, pre_e[e]);
            if (pre_t) v = gl::mul(v, pre_t[t0 + t]);
            data[t * SP + pad_idx(e)] = v;
        }
    } else {
        for (int idx = tid; idx < total; idx += nthr) {
            int e = idx & (n_sub - 1), t = idx >> p.LB;
            uint64_t v = in[((t0 + t) << p.LB) + e];
            if (pre_e) v = gl::mul(v, pre_e[e]);
            if (pre_t) v = gl::mul(v, pre_t[t0 + t]);
            data[t * SP + pad_idx(e)] = v;
        }
    }
    __syncthreads();

    // stage groups
    for (int s0 = 0; s0 < p.LB;) {
        int R = p.LB - s0 < RMAX ? p.LB - s0 : RMAX;
        const int items = p.T << (p.LB - R);
        for (int w = tid; w < items; w += nthr) {
            int t = w >> (p.LB - R), item = w & ((1 << (p.LB - R)) - 1);
            uint64_t* s = data + t * SP;
            switch (R) {
                case 5: dif_group<5>(s, tw_s, p.LB, s0, item); break;
                case 4: dif_group<4>(s, tw_s, p.LB, s0, item); break;
                case 3: dif_group<3>(s, tw_s, p.LB, s0, item); break;
                case 2: dif_group<2>(s, tw_s, p.LB, s0, item); break;
                default: dif_group<1>(s, tw_s, p.LB, s0, item); break;
            }
        }
        s0 += R;
        __syncthreads();
    }

    const bool post_tw = p.twA != nullptr;
    const uint32_t maskA = (1u << p.LA) - 1;
    if (p.out_strided) {
        for (int idx = tid; idx < total; idx += nthr) {
            int t = idx % p.T, o = idx / p.T;  // o = output index
            int pos = p.out_natural ? (int)gl::bitrev((uint32_t)o, p.LB) : o;
            uint64_t v = data[t * SP + pad_idx(pos)];
            if (post_tw) {
                uint32_t k = p.out_natural ? (uint32_t)o : gl::bitrev((uint32_t)o, p.LB);
                uint64_t x = (uint64_t)(t0 + t) * k;
                v = gl::mul(v, gl::mul(p.twA[x & maskA], p.twB[x >> p.LA]));
            } else if (p.post_scale) {
                v = gl::mul(v, p.post_scale);
            }
            out[((size_t)o << p.LM) + t0 + t] = v;
        }
    } else {
        for (int idx = tid; idx < total; idx += nthr) {
            int o = idx & (n_sub - 1), t = idx >> p.LB;
            int pos = p.out_natural ? (int)gl::bitrev((uint32_t)o, p.LB) : o;
            uint64_t v = data[t * SP + pad_idx(pos)];
            if (post_tw) {
                uint32_t k = p.out_natural ? (uint32_t)o : gl::bitrev((uint32_t)o, p.LB);
                uint64_t x = (uint64_t)(t0 + t) * k;
                v = gl::mul(v, gl::mul(p.twA[x & maskA], p.twB[x >> p.LA]));
            } else if (p.post_scale) {
                v = gl::mul(v, p.post_scale);
            }
            out[((t0 + t) << p.LB) + o] = v;
        }
    }
}

// ---------------------------------------------------------------- host side: tables and launch plans
__global__ void pow_table_kernel(uint64_t* out, uint64_t base, uint64_t scale, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gl::mul(scale, gl::pow(base, i));
}

static uint64_t* make_pow_table(Ctx* ctx, uint64_t base, uint64_t scale, size_t n) {
    uint64_t* d = (uint64_t*)ctx->alloc_persistent(n * sizeof(uint64_t));
    pow_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d, base, scale, n);
    ctx->kernel_launches++;
    return d;
}

// per-stage twiddle table for a 2^L DIF transform with root w (order 2^L): stage st holds w^(2^st * j), j < 2^(L-1-st)
static uint64_t* make_stage_table(Ctx* ctx, int L, uint64_t w) {
    size_t n = (size_t)1 << L;
    std::vector<uint64_t> h(n ? n : 1, 0);
    uint64_t ws = w;
    for (int st = 0; st < L; st++) {
        size_t off = n - (n >> st), cnt = n >> (st + 1);
        uint64_t x = 1;
        for (size_t j = 0; j < cnt; j++) { h[off + j] = x; x = gl::mul(x, ws); }
        ws = gl::sqr(ws);
    }
    uint64_t* d = (uint64_t*)ctx->alloc_persistent(h.size() * sizeof(uint64_t));
    CUDA_CHECK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return d;
}

const NttPlan& get_ntt_plan(Ctx* ctx, int log_n, bool inverse) {
    int key = log_n * 2 + (inverse ? 1 : 0);
    auto it = ctx->ntt_plans.find(key);
    if (it != ctx->ntt_plans.end()) return it->second;
    NttPlan pl{};
    pl.log_n = log_n;
    pl.inverse = inverse;
    // single pass up to 2^11, else split (L1 rows >= L2 cols)
    if (log_n <= 11) { pl.L1 = log_n; pl.L2 = 0; }
    else { pl.L1 = (log_n + 1) / 2; pl.L2 = log_n - pl.L1; }
    uint64_t w = gl::omega(log_n);
    if (inverse) w = gl::inv(w);
    uint64_t w1 = w, w2 = w;
    for (int i = 0; i < pl.L2; i++) w1 = gl::sqr(w1);  // order 2^L1
    for (int i = 0; i < pl.L1; i++) w2 = gl::sqr(w2);  // order 2^L2
    pl.tw1 = make_stage_table(ctx, pl.L1, w1);
    pl.tw2 = pl.L2 ? make_stage_table(ctx, pl.L2, w2) : nullptr;
    uint64_t n_inv = gl::inv(((uint64_t)1 << log_n) % GL_P);
    pl.n_inv = n_inv;
    if (pl.L2) {
        pl.LA = (log_n + 1) / 2;
        pl.twA = make_pow_table(ctx, w, inverse ? n_inv : 1, (size_t)1 << pl.LA);
        pl.twB = make_pow_table(ctx, gl::pow(w, (uint64_t)1 << pl.LA), 1, (size_t)1 << (log_n - pl.LA));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return ctx->ntt_plans.emplace(key, pl).first->second;
}

// Tables of shift^i for a coset.  The first MAX_CACHED_COSETS shifts of a size are cached (the LDE/quotient cosets of the
// base and recursion layers: 8); beyond that (compression layers: LDE factor up to 2048) one transient table per size is
// refilled on the stream before each transform -- no allocation and no synchronisation per coset.
static constexpr size_t MAX_CACHED_COSETS = 16;
static void fill_pow_table(Ctx* ctx, uint64_t* d, uint64_t base, size_t n) {
    pow_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d, base, 1, n);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
}
const CosetTables& get_coset_tables(Ctx* ctx, int log_n, uint64_t shift) {
    auto key = std::make_pair(log_n, shift);
    auto it = ctx->coset_tables.find(key);
    if (it != ctx->coset_tables.end()) return it->second;
    const NttPlan& pl = get_ntt_plan(ctx, log_n, false);
    size_t cached = 0;
    for (auto& kv : ctx->coset_tables) cached += kv.first.first == log_n;
    const uint64_t base_e = pl.L2 ? gl::pow(shift, (uint64_t)1 << pl.L2) : shift;   // shift^(n2*i1), or shift^i for one pass
    if (cached >= MAX_CACHED_COSETS) {
        auto tkey = std::make_pair(log_n + 1000, (uint64_t)0);   // the transient slot of this size
        auto tit = ctx->coset_tables.find(tkey);
        if (tit == ctx->coset_tables.end()) {
            CosetTables ct{};
            ct.pre_e = (uint64_t*)ctx->alloc_persistent(((size_t)8) << pl.L1);
            ct.pre_t = pl.L2 ? (uint64_t*)ctx->alloc_persistent(((size_t)8) << pl.L2) : nullptr;
            tit = ctx->coset_tables.emplace(tkey, ct).first;
        }
        // stream order protects the previous transform that read the slot
        fill_pow_table(ctx, tit->second.pre_e, base_e, (size_t)1 << pl.L1);
        if (pl.L2) fill_pow_table(ctx, tit->second.pre_t, shift, (size_t)1 << pl.L2);
        return tit->second;
    }
    CosetTables ct{};
    ct.pre_e = make_pow_table(ctx, base_e, 1, (size_t)1 << pl.L1);
    ct.pre_t = pl.L2 ? make_pow_table(ctx, shift, 1, (size_t)1 << pl.L2) : nullptr;   // shift^i0
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return ctx->coset_tables.emplace(key, ct).first->second;
}

static int pick_tile(int LB, int LM) {
    // members per CTA: 2^LB * T * 8 bytes of shared memory, aim for <= 64 KB and >= 64-byte global segments
    int T = 8;
    while (T > 1 && ((size_t)T << LB) * 8 > 72 * 1024) T >>= 1;
    while (T > 1 && (1 << LM) < T) T >>= 1;
    if (LB <= 7) { while (((size_t)(2 * T) << LB) * 8 <= 32 * 1024 && (1 << LM) >= 2 * T) T <<= 1; }
    return T;
}

static void launch_pass(Ctx* ctx, NttPass p, int n_polys, unsigned n_z = 1) {
    p.T = pick_tile(p.LB, p.LM);
    int n_sub = 1 << p.LB;
    int SP = n_sub + (n_sub >> 5) + 1;
    size_t smem = ((size_t)n_sub + (size_t)p.T * SP) * sizeof(uint64_t);
    int thr_per = p.LB > RMAX ? (1 << (p.LB - RMAX)) : 1;
    int threads = thr_per * p.T;
    if (threads < 64) threads = 64;
    if (threads > 512) threads = 512;
    static bool attr_set[64] = {};   // per device
    if (!attr_set[ctx->device & 63]) {
        CUDA_CHECK(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[ctx->device & 63] = true;
    }
    dim3 grid((unsigned)((1u << p.LM) / p.T), (unsigned)n_polys, n_z);
    ntt_pass_kernel<<<grid, threads, smem, ctx->stream>>>(p);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
}

// natural-order monomials -> evaluations on shift*<w_n> in bit-reversed order.  in == out allowed.
void ntt_forward_coset(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, int log_n, int n_polys,
                       uint64_t shift) {
    if (n_polys == 0) return;
    if (log_n == 20) { ntt1024_forward_coset(ctx, in, in_stride, out, out_stride, n_polys, shift); return; }
    const NttPlan& pl = get_ntt_plan(ctx, log_n, false);
    const CosetTables* ct = shift != 1 ? &get_coset_tables(ctx, log_n, shift) : nullptr;
    NttPass p{};
    p.in = in; p.out = out; p.in_col_stride = in_stride; p.out_col_stride = out_stride;
    if (pl.L2 == 0) {
        p.LB = pl.L1; p.LM = 0; p.in_strided = 0; p.out_strided = 0; p.out_natural = 0; p.tw = pl.tw1;
        p.pre_e = ct ? ct->pre_e : nullptr;
        launch_pass(ctx, p, n_polys);
        return;
    }
    // pass 1: columns
    p.LB = pl.L1; p.LM = pl.L2; p.in_strided = 1; p.out_strided = 1; p.out_natural = 0; p.tw = pl.tw1;
    p.pre_e = ct ? ct->pre_e : nullptr; p.pre_t = ct ? ct->pre_t : nullptr;
    p.twA = pl.twA; p.twB = pl.twB; p.LA = pl.LA;
    launch_pass(ctx, p, n_polys);
    // pass 2: rows, in place on out
    NttPass q{};
    q.in = out; q.out = out; q.in_col_stride = out_stride; q.out_col_stride = out_stride;
    q.LB = pl.L2; q.LM = pl.L1; q.in_strided = 0; q.out_strided = 0; q.out_natural = 0; q.tw = pl.tw2;
    launch_pass(ctx, q, n_polys);
}

// ---- all cosets of an LDE at once
__global__ void coset_pow_table_kernel(uint64_t* out, const uint64_t* __restrict__ bases, size_t n) {  // out[c][i] = bases[c]^i
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[(size_t)blockIdx.y * n + i] = gl::pow(bases[blockIdx.y], i);
}
static const CosetBatchTables& get_coset_batch_tables(Ctx* ctx, int log_n, int log_e) {
    auto key = std::make_pair(log_n, log_e);
    auto it = ctx->coset_batch_tables.find(key);
    if (it != ctx->coset_batch_tables.end()) return it->second;
    const NttPlan& pl = get_ntt_plan(ctx, log_n, false);
    const size_t E = (size_t)1 << log_e, ne = (size_t)1 << pl.L1, nt = pl.L2 ? (size_t)1 << pl.L2 : 0;
    std::vector<uint64_t> h(2 * E);
    for (size_t c = 0; c < E; c++) {
        const uint64_t shift = lde_coset_shift(log_n, log_e, (uint32_t)c);
        h[c] = pl.L2 ? gl::pow(shift, (uint64_t)1 << pl.L2) : shift;   // base of pre_e: shift^(n2*i1), or shift^i for one pass
        h[E + c] = shift;                                               // base of pre_t
    }
    uint64_t* d_bases = (uint64_t*)ctx->alloc_persistent(2 * E * 8);
    CUDA_CHECK(cudaMemcpyAsync(d_bases, h.data(), 2 * E * 8, cudaMemcpyHostToDevice, ctx->stream));
    CosetBatchTables t{};
    t.pre_e = (uint64_t*)ctx->alloc_persistent(E * ne * 8);
    coset_pow_table_kernel<<<dim3((unsigned)((ne + 127) / 128), (unsigned)E), 128, 0, ctx->stream>>>(t.pre_e, d_bases, ne);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
    if (nt) {
        t.pre_t = (uint64_t*)ctx->alloc_persistent(E * nt * 8);
        coset_pow_table_kernel<<<dim3((unsigned)((nt + 127) / 128), (unsigned)E), 128, 0, ctx->stream>>>(t.pre_t, d_bases + E, nt);
        CUDA_CHECK(cudaGetLastError());
        ctx->kernel_launches++;
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));   // h is a stack-lifetime host vector
    return ctx->coset_batch_tables.emplace(key, t).first->second;
}

void ntt_forward_cosets(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, int log_n, int n_polys, int log_e,
                        uint32_t c0, uint32_t nc) {
    if (n_polys == 0 || nc == 0) return;
    if (log_n == 20) {   // the 2^20 fast path keeps per-coset inter-pass tables (base / recursion layers: 8 cosets)
        for (uint32_t c = c0; c < c0 + nc; c++)
            ntt1024_forward_coset(ctx, in, in_stride, out + ((size_t)(c - c0) << 20), out_stride, n_polys, lde_coset_shift(log_n, log_e, c));
        return;
    }
    const NttPlan& pl = get_ntt_plan(ctx, log_n, false);
    const CosetBatchTables& bt = get_coset_batch_tables(ctx, log_n, log_e);
    const size_t N = (size_t)1 << log_n, ne = (size_t)1 << pl.L1, nt = pl.L2 ? (size_t)1 << pl.L2 : 0;
    // grid.z is limited to 65535 and grid.y * grid.z CTAs should stay reasonable: at most 2048 cosets per launch
    for (uint32_t b0 = 0; b0 < nc; b0 += 2048) {
        const uint32_t nb = nc - b0 < 2048 ? nc - b0 : 2048;
        uint64_t* o = out + (size_t)b0 * N;
        NttPass p{};
        p.in = in; p.out = o; p.in_col_stride = in_stride; p.out_col_stride = out_stride;
        p.in_z_stride = 0; p.out_z_stride = N;
        p.pre_e = bt.pre_e + (size_t)(c0 + b0) * ne; p.pre_e_z_stride = ne;
        if (pl.L2 == 0) {
            p.LB = pl.L1; p.LM = 0; p.in_strided = 0; p.out_strided = 0; p.out_natural = 0; p.tw = pl.tw1;
            launch_pass(ctx, p, n_polys, nb);
            continue;
        }
        p.LB = pl.L1; p.LM = pl.L2; p.in_strided = 1; p.out_strided = 1; p.out_natural = 0; p.tw = pl.tw1;
        p.pre_t = bt.pre_t + (size_t)(c0 + b0) * nt; p.pre_t_z_stride = nt;
        p.twA = pl.twA; p.twB = pl.twB; p.LA = pl.LA;
        launch_pass(ctx, p, n_polys, nb);
        NttPass q{};
        q.in = o; q.out = o; q.in_col_stride = out_stride; q.out_col_stride = out_stride;
        q.in_z_stride = N; q.out_z_stride = N;
        q.LB = pl.L2; q.LM = pl.L1; q.in_strided = 0; q.out_strided = 0; q.out_natural = 0; q.tw = pl.tw2;
        launch_pass(ctx, q, n_polys, nb);
    }
}

// natural-order evaluations on <w_n> -> natural-order monomials.  `tmp` (same shape as out) is scratch when the
// transform needs two passes; in == out is allowed only for the single-pass sizes.
void ntt_inverse(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, uint64_t* tmp, size_t tmp_stride,
                 int log_n, int n_polys) {
    if (n_polys == 0) return;
    if (log_n == 20) { ntt1024_inverse(ctx, in, in_stride, out, out_stride, tmp, tmp_stride, n_polys); return; }
    const NttPlan& pl = get_ntt_plan(ctx, log_n, true);
    NttPass p{};
    p.in = in; p.in_col_stride = in_stride;
    if (pl.L2 == 0) {
        p.out = out; p.out_col_stride = out_stride;
        p.LB = pl.L1; p.LM = 0; p.in_strided = 0; p.out_strided = 0; p.out_natural = 1; p.tw = pl.tw1;
        p.post_scale = pl.n_inv;
        launch_pass(ctx, p, n_polys);
        return;
    }
    // pass 1: columns -> tmp, rows k1 in natural order, twiddle (with n^-1 folded in)
    p.out = tmp; p.out_col_stride = tmp_stride;
    p.LB = pl.L1; p.LM = pl.L2; p.in_strided = 1; p.out_strided = 1; p.out_natural = 1; p.tw = pl.tw1;
    p.twA = pl.twA; p.twB = pl.twB; p.LA = pl.LA;
    launch_pass(ctx, p, n_polys);
    // pass 2: row k1 over i0 -> k0, written transposed: out[k1 + n1*k0]
    NttPass q{};
    q.in = tmp; q.in_col_stride = tmp_stride; q.out = out; q.out_col_stride = out_stride;
    q.LB = pl.L2; q.LM = pl.L1; q.in_strided = 0; q.out_strided = 1; q.out_natural = 1; q.tw = pl.tw2;
    launch_pass(ctx, q, n_polys);
}

__global__ void bitrev_copy_kernel(const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, int log_n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    out[(size_t)blockIdx.y * out_stride + gl::bitrev((uint32_t)i, log_n)] = in[(size_t)blockIdx.y * in_stride + i];
}
void bitrev_copy(Ctx* ctx, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, int log_n, int n_polys) {
    size_t n = (size_t)1 << log_n;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_polys);
    bitrev_copy_kernel<<<grid, 256, 0, ctx->stream>>>(in, in_stride, out, out_stride, log_n);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
}

}  // namespace zk
