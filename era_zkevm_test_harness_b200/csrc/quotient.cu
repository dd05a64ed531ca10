// quotient.cu -- the quotient numerator over one coset of the 8N domain, as three compact kernels.
//
// Replaces the "compute quotient" stage of boojum's `prove_from_precomputations`
// (/root/reference/src/prover_utils.rs:338-348); term order and relations are those of oracle/prover.c
// `quotient_numerator` (gates in gate-list order, boolean column, public inputs, lookup, copy permutation).
//
// Why three kernels and why the loops are NOT unrolled: the first version (one kernel, everything inlined) was 286 KB of
// SASS and ran at 1.4 IPC/SM with `stall_no_instruction` as the top stall (profiles/r01_b_prove_raw.csv) -- it was bound by
// instruction fetch, not by arithmetic or HBM.  Here every hot loop body is a few KB:
//   quotient_gates_kernel  general-purpose gates; instance loops stay rolled
//   quotient_p2_kernel     the flattened Poseidon2 gate (118 relations of degree 7), 4 lanes per loop iteration
//   quotient_perm_kernel   boolean column, public inputs, logUp lookup, copy-permutation chunks, division by Z_H
// and the alpha-weighted sums  sum_k alpha^k * r_k  are accumulated UNREDUCED: a 64x64 product is four IMAD.WIDE.U32 into
// three 96-bit column accumulators (1 fma + 1 alu instruction per partial product) and reduced mod p once per gate,
// instead of two modular multiplies and two modular adds per relation.
#include "host_common.cuh"
#include "poseidon2_core.cuh"
#include "quotient.cuh"
#include "dot.cuh"

#ifndef ZK_PERM_UNROLL
#define ZK_PERM_UNROLL 1
#endif
#define ZK_PRAGMA_(x) _Pragma(#x)
#define ZK_UNROLL(n) ZK_PRAGMA_(unroll n)

namespace zk {

__device__ __forceinline__ uint64_t selector(const uint64_t* __restrict__ kc, size_t cs, uint32_t path_len, uint32_t path_bits) {
    uint64_t sel = 1;
    for (uint32_t b = 0; b < path_len; b++) {
        uint64_t c = kc[(size_t)b * cs];
        sel = gl::mul(sel, ((path_bits >> b) & 1) ? c : gl::sub(1, c));
    }
    return sel;
}

// ------------------------------------------------------------------------------------------------ general-purpose gates
// One CTA = QG_P adjacent points of the coset.  The gate cells of those points (copy columns, then plain witness columns) and
// the constant columns are staged ONCE in shared memory -- one 256-byte TMA bulk copy (cp.async.bulk, 16-byte units) per column
// onto an mbarrier -- and every gate kind reads them from there: the first version walked the columns in global memory once per
// gate kind and pulled 6.2 GB per coset through HBM for 1.1 GB of columns (profiles/r01_m_summary.txt).  The QG_SPLIT warps of
// the CTA share the (gate, instance) work list of the same 32 points: the host cuts the list into QG_SPLIT contiguous segments of
// equal estimated cost (QuotParams::gate_t0 / gate_t1), so a gate is set up -- selector product, alpha-dot reduction -- by ONE warp
// unless it straddles a cut; the partial alpha-weighted sums are added at the end (exact field additions: the split does not
// change a bit of the result).  The first version strided every gate's instances over the warps, so each warp repeated the per-gate
// set-up (14.95 ms per proof with 2 warps; the contiguous split: 14.1 ms with 2 warps, 14.8 with 3, 14.9 with 4).
constexpr int QG_P = 32;
#ifndef ZK_QG_SPLIT
#define ZK_QG_SPLIT ZKGPU_QG_WARPS
#endif
constexpr int QG_SPLIT = ZK_QG_SPLIT;
static size_t quotient_gates_smem(const zkgpu_geometry& g) {
    return ((size_t)(g.n_copy + g.n_witness_plain + g.n_const_cols) * QG_P + 2 * QG_SPLIT * QG_P) * sizeof(uint64_t);
}
__global__ void __launch_bounds__(32 * QG_SPLIT, 16 / QG_SPLIT) quotient_gates_kernel(const __grid_constant__ QuotParams p) {
    extern __shared__ __align__(128) uint64_t qg_smem[];
    __shared__ __align__(8) uint64_t qg_bar;
    const size_t N = (size_t)1 << p.g.log_n;
    const uint32_t n_copy = p.g.n_copy, n_cells = p.g.n_copy + p.g.n_witness_plain, n_rows = n_cells + p.g.n_const_cols;
    uint64_t* tile = qg_smem;                               // [n_cells][QG_P]
    uint64_t* ktile = qg_smem + (size_t)n_cells * QG_P;      // [n_const_cols][QG_P]
    uint64_t* red = qg_smem + (size_t)n_rows * QG_P;         // [2][QG_SPLIT][QG_P]
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t j0 = (size_t)blockIdx.x * QG_P;
    const uint32_t npts = (uint32_t)(N - j0 < QG_P ? N - j0 : QG_P);
    auto row_src = [&](uint32_t r) -> const uint64_t* {
        if (r < n_copy) return p.wit + (size_t)r * p.cs_w + j0;
        if (r < n_cells) return p.wit + (size_t)(p.NP + r - n_copy) * p.cs_w + j0;          // plain witness columns start at NP
        return p.setup + (size_t)(p.NP + r - n_cells) * p.cs_s + j0;                          // constant columns
    };
    const bool bulk = npts == QG_P && ((reinterpret_cast<uintptr_t>(p.wit) | reinterpret_cast<uintptr_t>(p.setup)) & 15) == 0 &&
                      ((p.cs_w | p.cs_s) & 1) == 0;
    if (bulk) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&qg_bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n_rows * QG_P * 8u) : "memory");
        for (uint32_t r = tid; r < n_rows; r += 32 * QG_SPLIT) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(qg_smem + (size_t)r * QG_P);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(row_src(r)), "r"(QG_P * 8u), "r"(bar) : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
        }
    } else {   // short or unaligned tiles (parity-test sizes): plain loads
        for (uint32_t i = tid; i < n_rows * QG_P; i += 32 * QG_SPLIT) {
            const uint32_t r = i / QG_P, c = i % QG_P;
            qg_smem[i] = c < npts ? row_src(r)[c] : 0;
        }
        __syncthreads();
    }
    // Relation values are handed to the alpha dot product as LAZY residues (dot.cuh multiplies exact integers, any u64
    // representative is fine): products are glx::mul / glx::fma (17 / 20 instructions, no canonicalisation), a canonical
    // subtrahend is taken off with gl::sub, whose single borrow fix-up is exact for ANY minuend when the subtrahend is < p.
    const uint64_t* __restrict__ w = tile + lane;
    const uint64_t* __restrict__ kc = ktile + lane;
    const ulonglong2* __restrict__ apow = reinterpret_cast<const ulonglong2*>(p.apow);
    constexpr size_t cw = QG_P, cs = QG_P;
    gl::e2 acc = gl::make2(0, 0);

#pragma unroll 1
    for (uint32_t gi = 0; gi < p.g.n_gates; gi++) {
        const zkgpu_gate gt = p.g.gates[gi];
        const uint32_t tb = p.gate_t0[warp][gi], te = p.gate_t1[warp][gi];   // this warp's instances of the gate
        if (tb >= te || gt.kind == ZKGPU_GATE_POSEIDON2_FLATTENED) continue;
        const uint64_t* __restrict__ gk = kc + (size_t)gt.path_len * cs;
        const ulonglong2* __restrict__ ap = apow + p.gate_term0[gi];
        DotE2 d;
        dote_zero(d);
        switch (gt.kind) {
            case ZKGPU_GATE_CONSTANTS_ALLOCATOR:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) dote_add(d, gl::sub(w[(size_t)t * cw], gk[(size_t)t * cs]), ap[t]);
                break;
            case ZKGPU_GATE_FMA: {
                const uint64_t k0 = gk[0], k1 = gk[cs];
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(4 * t) * cw;
                    uint64_t r = gl::sub(glx::fma(glx::mul(k0, x[0]), x[cw], glx::mul(k1, x[2 * cw])), x[3 * cw]);
                    dote_add(d, r, ap[t]);
                }
            } break;
            case ZKGPU_GATE_REDUCTION4: {
                const uint64_t k0 = gk[0], k1 = gk[cs], k2 = gk[2 * cs], k3 = gk[3 * cs];
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(5 * t) * cw;
                    uint64_t s = glx::fma(k3, x[3 * cw], glx::fma(k2, x[2 * cw], glx::fma(k1, x[cw], glx::mul(k0, x[0]))));
                    dote_add(d, gl::sub(s, x[4 * cw]), ap[t]);
                }
            } break;
            case ZKGPU_GATE_SELECTION:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(4 * t) * cw;
                    uint64_t b = x[2 * cw];
                    // s*a + (1-s)*b - out = s*(a - b) + b - out
                    dote_add(d, gl::sub(glx::fma(x[0], gl::sub(x[cw], b), b), x[3 * cw]), ap[t]);
                }
                break;
            case ZKGPU_GATE_PARALLEL_SELECTION4:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(13 * t) * cw;
                    const uint64_t s = x[0];
#pragma unroll 1
                    for (uint32_t i = 0; i < 4; i++) {
                        const uint64_t* y = x + (size_t)(1 + 3 * i) * cw;
                        uint64_t b = y[cw];
                        dote_add(d, gl::sub(glx::fma(s, gl::sub(y[0], b), b), y[2 * cw]), ap[4 * t + i]);
                    }
                }
                break;
            case ZKGPU_GATE_ZERO_CHECK:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(3 * t) * cw;
                    uint64_t xv = x[0], zf = x[2 * cw];
                    dote_add(d, gl::sub(glx::mul(xv, x[cw]), gl::sub(1, zf)), ap[2 * t]);
                    dote_add(d, glx::mul(xv, zf), ap[2 * t + 1]);
                }
                break;
            case ZKGPU_GATE_UINTX_ADD: {
                const uint64_t k0 = gk[0];
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(5 * t) * cw;
                    uint64_t co = x[4 * cw];
                    // a + b + cin - c - k*cout, as (a + b + cin) + (p - c) + k*(p - cout): every term canonical, sums lazy
                    uint64_t lhs = glx::add_canon(glx::add_canon(x[0], x[cw]), x[2 * cw]);
                    lhs = glx::add_canon(lhs, gl::neg(x[3 * cw]));
                    dote_add(d, glx::fma(k0, gl::neg(co), lhs), ap[2 * t]);
                    dote_add(d, gl::sub(glx::mul(co, co), co), ap[2 * t + 1]);
                }
            } break;
            case ZKGPU_GATE_U32_TRI_ADD_CARRY:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(5 * t) * cw;
                    uint64_t lhs = gl::add(gl::add(x[0], x[cw]), x[2 * cw]);
                    dote_add(d, gl::sub(lhs, gl::add(x[3 * cw], gl::mul_pow2(x[4 * cw], 32))), ap[t]);
                }
                break;
            case ZKGPU_GATE_BOUNDED_BOOLEAN:
            case ZKGPU_GATE_BOOLEAN_ALL:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t x = w[(size_t)t * cw];
                    dote_add(d, gl::sub(gl::sqr(x), x), ap[t]);
                }
                break;
            case ZKGPU_GATE_MATMUL12_EXTERNAL:
            case ZKGPU_GATE_MATMUL12_INNER:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(24 * t) * cw;
                    uint64_t s[12];
#pragma unroll
                    for (int i = 0; i < 12; i++) s[i] = x[(size_t)i * cw];
                    if (gt.kind == ZKGPU_GATE_MATMUL12_EXTERNAL) p2x_external(s);
                    else p2x_internal(s);
#pragma unroll
                    for (int i = 0; i < 12; i++) dote_add(d, gl::sub(x[(size_t)(12 + i) * cw], glx::canon(s[i])), ap[12 * t + i]);
                }
                break;
            case ZKGPU_GATE_NONLINEARITY7: {
                const uint64_t k0 = gk[0];
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(2 * t) * cw;
                    dote_add(d, gl::sub(x[cw], glx::canon(glx::pow7(glx::add_canon(x[0], k0)))), ap[t]);
                }
            } break;
            case ZKGPU_GATE_CONDITIONAL_SWAP4:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(17 * t) * cw;
                    const uint64_t sw = x[8 * cw];
#pragma unroll 1
                    for (uint32_t i = 0; i < 4; i++) {
                        const uint64_t a = x[(size_t)i * cw], b = x[(size_t)(4 + i) * cw];
                        const uint64_t dl = gl::mul(sw, gl::sub(b, a));
                        dote_add(d, gl::sub(gl::add(dl, a), x[(size_t)(9 + i) * cw]), ap[8 * t + 2 * i]);
                        dote_add(d, gl::sub(gl::sub(b, dl), x[(size_t)(13 + i) * cw]), ap[8 * t + 2 * i + 1]);
                    }
                }
                break;
            case ZKGPU_GATE_ZERO_CHECK_WITNESS: {
                const uint64_t* pw = w + (size_t)n_copy * cw;   // plain witness cells follow the copy columns in the tile
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(2 * t) * cw;
                    uint64_t xv = x[0], zf = x[cw];
                    dote_add(d, gl::sub(gl::mul(xv, pw[(size_t)t * cw]), gl::sub(1, zf)), ap[2 * t]);
                    dote_add(d, gl::mul(xv, zf), ap[2 * t + 1]);
                }
            } break;
            case ZKGPU_GATE_DOT_PRODUCT4:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(9 * t) * cw;
                    uint64_t s = glx::fma(x[6 * cw], x[7 * cw], glx::fma(x[4 * cw], x[5 * cw], glx::fma(x[2 * cw], x[3 * cw], glx::mul(x[0], x[cw]))));
                    dote_add(d, gl::sub(s, x[8 * cw]), ap[t]);
                }
                break;
            case ZKGPU_GATE_U8X4_FMA:
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(26 * t) * cw;
                    // sum_{i,j} a_i b_j 2^(8(i+j)), grouped by i+j; then the linear part, byte position by byte position
                    uint64_t a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) { a[i] = x[(size_t)i * cw]; b[i] = x[(size_t)(4 + i) * cw]; }
                    uint64_t r = gl::mul(a[0], b[0]);
#pragma unroll
                    for (int s = 1; s < 7; s++) {
                        uint64_t g = 0;
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            if (s - i >= 0 && s - i < 4) g = gl::add(g, gl::mul(a[i], b[s - i]));
                        r = gl::add(r, gl::mul_pow2(g, 8 * s));
                    }
#pragma unroll 1
                    for (uint32_t i = 0; i < 4; i++) {
                        const uint64_t* y = x + (size_t)(8 + i) * cw;
                        uint64_t lin = gl::sub(gl::add(y[0], y[4 * cw]), gl::add(y[8 * cw], gl::mul_pow2(y[12 * cw], 32)));
                        r = gl::add(r, gl::mul_pow2(lin, 8 * i));
                    }
                    dote_add(d, r, ap[t]);
                }
                break;
            case ZKGPU_GATE_FMA_EXT: {
                const gl::e2 k0 = gl::make2(gk[0], gk[cs]), k1 = gl::make2(gk[2 * cs], gk[3 * cs]);
#pragma unroll 1
                for (uint32_t t = tb; t < te; t++) {
                    const uint64_t* x = w + (size_t)(8 * t) * cw;
                    gl::e2 ab = gl::mul(gl::make2(x[0], x[cw]), gl::make2(x[2 * cw], x[3 * cw]));
                    gl::e2 r = gl::sub(gl::add(gl::mul(k0, ab), gl::mul(k1, gl::make2(x[4 * cw], x[5 * cw]))), gl::make2(x[6 * cw], x[7 * cw]));
                    dote_add(d, r.c0, ap[2 * t]);
                    dote_add(d, r.c1, ap[2 * t + 1]);
                }
            } break;
            default: break;
        }
        const uint64_t sel = selector(kc, cs, gt.path_len, gt.path_bits);
        acc = gl::add(acc, gl::mul_base(dote_reduce(d), sel));
    }
    red[(0 * QG_SPLIT + warp) * QG_P + lane] = acc.c0;
    red[(1 * QG_SPLIT + warp) * QG_P + lane] = acc.c1;
    __syncthreads();
    if (warp == 0 && lane < npts) {
        uint64_t r0 = red[lane], r1 = red[QG_SPLIT * QG_P + lane];
#pragma unroll
        for (int q = 1; q < QG_SPLIT; q++) {
            r0 = gl::add(r0, red[q * QG_P + lane]);
            r1 = gl::add(r1, red[(QG_SPLIT + q) * QG_P + lane]);
        }
        p.t0[j0 + lane] = r0;
        p.t1[j0 + lane] = r1;
    }
}

// ------------------------------------------------------------------------------------------------ flattened Poseidon2 gate
// columns: [0,12) input state, then one variable per S-box output in round order (48 + 22 + 48); relation k:
//   v_k - (lin_k + rc_k)^7   where lin_k is the running linear-layer image of the previous variables.
struct P2GateRC {
    const uint64_t* rc;
    __device__ __forceinline__ uint64_t operator[](int i) const { return rc[i]; }
};
__global__ void __launch_bounds__(128) quotient_p2_kernel(const __grid_constant__ QuotParams p) {
    const size_t N = (size_t)1 << p.g.log_n;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const uint64_t* __restrict__ w = p.wit + j;
    const uint64_t* __restrict__ kc = p.setup + (size_t)p.NP * p.cs_s + j;
    const size_t cw = p.cs_w;
    const zkgpu_gate gt = p.g.gates[p.p2_gate];
    const ulonglong2* __restrict__ ap = reinterpret_cast<const ulonglong2*>(p.apow) + p.gate_term0[p.p2_gate];
    const uint64_t* __restrict__ rc = p.rc;

    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = w[(size_t)i * cw];
    p2x_external(s);
    DotE2 d;
    dote_zero(d);
    // gate cell c: copy column c below n_copy, plain witness column NP + (c - n_copy) above (compression modes 1-3)
    const uint32_t n_copy = p.g.n_copy, gap = p.NP - p.g.n_copy;
    auto cell = [&](uint32_t c) -> uint64_t { return w[(size_t)(c < n_copy ? c : c + gap) * cw]; };
    uint32_t v = 12;   // next variable cell
    int r = 0;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
        for (int q = 0; q < 4; q++, r++) {
#pragma unroll 1
            for (int it = 0; it < 3; it++) {
                uint64_t nv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    nv[u] = cell(v + u);
                    uint64_t rel = glx::sub(nv[u], glx::pow7(glx::add_canon(s[u], rc[12 * r + 4 * it + u])));
                    dote_add(d, rel, ap[u]);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) s[i] = s[i + 4];
#pragma unroll
                for (int u = 0; u < 4; u++) s[8 + u] = nv[u];
                v += 4;
                ap += 4;
            }
            p2x_external(s);
        }
        if (half == 0) {
#pragma unroll 1
            for (int q = 0; q < 22; q++, r++) {
                uint64_t nv = cell(v);
                uint64_t rel = glx::sub(nv, glx::pow7(glx::add_canon(s[0], rc[12 * r])));
                dote_add(d, rel, ap[0]);
                s[0] = nv;
                v += 1;
                ap += 1;
                p2x_internal(s);
            }
        }
    }
    const uint64_t sel = selector(kc, p.cs_s, gt.path_len, gt.path_bits);
    gl::e2 acc = gl::mul_base(dote_reduce(d), sel);
    p.t0[j] = gl::add(p.t0[j], acc.c0);
    p.t1[j] = gl::add(p.t1[j], acc.c1);
}

// ------------------------------------------------------------------------------------------------ boolean, PI, lookup, copy permutation
__global__ void __launch_bounds__(128) quotient_perm_kernel(const __grid_constant__ QuotParams p) {
    const uint32_t log_n = p.g.log_n;
    const size_t N = (size_t)1 << log_n;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const zkgpu_geometry& g = p.g;
    const uint64_t* __restrict__ w = p.wit + j;
    const uint64_t* __restrict__ sg = p.setup + j;
    const uint64_t* __restrict__ kc = p.setup + (size_t)p.NP * p.cs_s + j;
    const uint64_t* __restrict__ tb = kc + (size_t)g.n_const_cols * p.cs_s;
    const uint64_t* __restrict__ e2 = p.s2 + j;
    const ulonglong2* __restrict__ apow = reinterpret_cast<const ulonglong2*>(p.apow);
    const size_t cw = p.cs_w, cs = p.cs_s, c2 = p.cs_2;
    const uint64_t x = gl::mul(p.shift, p.omega_br[j]);
    gl::e2 acc = gl::make2(p.t0[j], p.t1[j]);
    uint32_t k = p.tail_term0;

    // 2. boolean column
    if (g.has_boolean_col) {
        uint64_t b = w[(size_t)g.n_copy * cw];
        ulonglong2 a = apow[k++];
        acc = gl::add(acc, gl::mul_base(gl::make2(a.x, a.y), gl::sub(gl::sqr(b), b)));
    }
    // 5a. Lagrange denominator N(x - 1) of L_0 (public inputs are not quotient terms: the reference opens them through the DEEP
    // polynomial, prover.cu deep_kernel)
    const uint64_t l0_inv = p.l0_inv[j];   // = 1 / (n (x - 1)), tabulated per context
    // 4. lookup
    if (g.lookup_reps) {
        const uint32_t LW = g.lookup_width;
        const uint64_t* __restrict__ lw = w + (size_t)p.lookup_col0 * cw;
        const gl::e2 tid = gl::mul_base(p.lgamma_pow[LW], kc[(size_t)g.table_id_col * cs]);
#pragma unroll 1
        for (uint32_t i = 0; i < g.lookup_reps; i++) {
            gl::e2 den2 = gl::add(p.lbeta, tid);
#pragma unroll 1
            for (uint32_t q = 0; q < LW; q++) den2 = gl::add(den2, gl::mul_base(p.lgamma_pow[q], lw[(size_t)(i * LW + q) * cw]));
            gl::e2 A = gl::make2(e2[(size_t)(2 * (p.C + i)) * c2], e2[(size_t)(2 * (p.C + i) + 1) * c2]);
            gl::e2 t = gl::mul(A, den2);
            t.c0 = gl::sub(t.c0, 1);
            ulonglong2 a = apow[k++];
            acc = gl::add(acc, gl::mul(gl::make2(a.x, a.y), t));
        }
        gl::e2 den2 = p.lbeta;
#pragma unroll 1
        for (uint32_t q = 0; q <= LW; q++) den2 = gl::add(den2, gl::mul_base(p.lgamma_pow[q], tb[(size_t)q * cs]));
        gl::e2 B = gl::make2(e2[(size_t)(2 * (p.C + g.lookup_reps)) * c2], e2[(size_t)(2 * (p.C + g.lookup_reps) + 1) * c2]);
        gl::e2 t = gl::mul(B, den2);
        t.c0 = gl::sub(t.c0, w[(size_t)(p.W - 1) * cw]);
        ulonglong2 a = apow[k++];
        acc = gl::add(acc, gl::mul(gl::make2(a.x, a.y), t));
    }
    // 5. copy permutation
    gl::e2 zv = gl::make2(e2[0], e2[c2]);
    {
        uint64_t l0 = gl::mul(p.xn_minus_1, l0_inv);
        gl::e2 t = gl::mul_base(gl::make2(gl::sub(zv.c0, 1), zv.c1), l0);
        ulonglong2 a = apow[k++];
        acc = gl::add(acc, gl::mul(gl::make2(a.x, a.y), t));
    }
    // z(w*x): position of the next natural index inside the bit-reversed coset
    const uint32_t nat = gl::bitrev((uint32_t)j, log_n);
    const size_t jn = gl::bitrev((nat + 1) & (uint32_t)(N - 1), log_n);
    const gl::e2 zs = gl::make2(p.s2[jn], p.s2[c2 + jn]);
    // a_i = w_i + gamma + (beta*k_i)*x with beta*k_i from the per-proof table (k_i = copy-permutation non-residues)
    const ulonglong2* __restrict__ bkt = reinterpret_cast<const ulonglong2*>(p.beta_k);
    gl::e2 prev = zv;
    const uint32_t QD = g.quotient_degree;
    // one flat loop over the copy-permuted columns with the loads of column i+1 issued before the arithmetic of column i (ncu: the
    // two-level loop waited on its own loads, stall_long_scoreboard 4.4 of 17 cycles per instruction at 14 % DRAM throughput)
    gl::e2 num = gl::make2(1, 0), dn = gl::make2(1, 0);
    uint64_t wv = w[0], sgv = sg[0];
    uint32_t c = 0, left = min(QD, p.NP);
    bool first = true;
#pragma unroll 1
    for (uint32_t i = 0; i < p.NP; i++) {
        const uint32_t in = min(i + 1, p.NP - 1);
        const uint64_t wn = w[(size_t)in * cw], sgn = sg[(size_t)in * cs];
        // the two linear forms as LAZY residues (the Ext2 multiply takes any u64 representative and reduces once):
        // a = w + gamma + beta*k_i*x, b = w + gamma + beta*sigma_i with the additions folded into fused multiply-adds
        const uint64_t wg0 = gl::add(wv, p.gamma.c0);
        const ulonglong2 bk = bkt[i];
        gl::e2 a = gl::make2(glx::fma(bk.x, x, wg0), glx::fma(bk.y, x, p.gamma.c1));
        gl::e2 b = gl::make2(glx::fma(p.beta.c0, sgv, wg0), glx::fma(p.beta.c1, sgv, p.gamma.c1));
        if (first) {   // first column of a chunk: no multiplication by one (uniform branch: every thread is at the same column)
            num = a; dn = b; first = false;   // lazy residues are fine: the Ext2 multiply takes any u64 representative
        } else {
            num = gl::mul(num, a);
            dn = gl::mul(dn, b);
        }
        wv = wn; sgv = sgn;
        if (--left == 0) {   // end of chunk c
            gl::e2 cur = (c + 1 < p.C) ? gl::make2(e2[(size_t)(2 * (c + 1)) * c2], e2[(size_t)(2 * (c + 1) + 1) * c2]) : zs;
            gl::e2 t = gl::sub(gl::mul(cur, dn), gl::mul(prev, num));
            ulonglong2 ak = apow[k++];
            acc = gl::add(acc, gl::mul(gl::make2(ak.x, ak.y), t));
            prev = cur;
            first = true;
            c++;
            left = min(QD, p.NP - (i + 1));
        }
    }
    acc = gl::mul_base(acc, p.zh_inv);
    p.t0[j] = acc.c0;
    p.t1[j] = acc.c1;
}

void launch_quotient_coset(Ctx* ctx, const QuotParams& p) {
    const size_t N = (size_t)1 << p.g.log_n;
    const unsigned grid = (unsigned)((N + 127) / 128);
    {
        const size_t smem = quotient_gates_smem(p.g);
        static size_t attr_smem[64] = {};   // per device: the opt-in limit only ever grows
        if (smem > attr_smem[ctx->device & 63]) {
            CUDA_CHECK(cudaFuncSetAttribute(quotient_gates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_smem[ctx->device & 63] = smem;
        }
        quotient_gates_kernel<<<(unsigned)((N + QG_P - 1) / QG_P), 32 * QG_SPLIT, smem, ctx->stream>>>(p);
    }
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
    if (p.p2_gate != 0xFFFFFFFFu) {
        quotient_p2_kernel<<<grid, 128, 0, ctx->stream>>>(p);
        CUDA_CHECK(cudaGetLastError());
        ctx->kernel_launches++;
    }
    quotient_perm_kernel<<<grid, 128, 0, ctx->stream>>>(p);
    CUDA_CHECK(cudaGetLastError());
    ctx->kernel_launches++;
}

}  // namespace zk
