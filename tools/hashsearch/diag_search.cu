// tools/hashsearch/diag_search.cu -- brute-force search for the Poseidon2 internal-matrix diagonal (test tooling, not product code).
//
// Hypothesis family: boojum's Poseidon2-Goldilocks is the structure restated in oracle/primitives.c (plonky2 round-constant table,
// circ(2*M4, M4, M4) external matrix, M_I = J + diag(2^s_i)) with a diagonal (s_0..s_11) of twelve DISTINCT exponents in [0, 16)
// that differs from the recollected (4,14,11,8,0,5,2,9,13,6,3,12).  16!/4! = 8.7e11 ordered selections; each candidate costs the 22
// partial and the last 4 full rounds on a state precomputed up to the end of the first 4 full rounds.  A hit = some output lane
// equals an element of the golden cap (single-permutation KAT, tests/golden/poseidon2_kat.json: one 8-element FRI leaf -> cap entry).
//   usage: diag_search <kat.bin> [first_fraction last_fraction]      (kat.bin written by tools/hashsearch/run_diag_search.py)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "../../era_zkevm_test_harness_b200/csrc/glx.cuh"
#include "../../era_zkevm_test_harness_b200/csrc/poseidon2_consts.cuh"

__constant__ uint64_t RC[360] = {ZK_P2_RC_INIT};
__constant__ uint64_t MID[8][12];      // mid-states: variant v
__constant__ uint64_t TARGETS[64];
__constant__ int N_TARGETS;

__device__ __forceinline__ void ext_layer(uint64_t (&s)[12]) {
    glx::w96 y[12];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        glx::w96 x0 = glx::widen(s[4 * b]), x1 = glx::widen(s[4 * b + 1]), x2 = glx::widen(s[4 * b + 2]), x3 = glx::widen(s[4 * b + 3]);
        glx::w96 t0 = glx::add(x0, x1), t1 = glx::add(x2, x3);
        glx::w96 t2 = glx::add(glx::shl(x1, 1), t1), t3 = glx::add(glx::shl(x3, 1), t0);
        glx::w96 t4 = glx::add(glx::shl(t1, 2), t3), t5 = glx::add(glx::shl(t0, 2), t2);
        y[4 * b] = glx::add(t3, t5); y[4 * b + 1] = t5; y[4 * b + 2] = glx::add(t2, t4); y[4 * b + 3] = t4;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        glx::w96 t = glx::add(glx::add(y[i], y[4 + i]), y[8 + i]);
        s[i] = glx::reduce(glx::add(y[i], t)); s[4 + i] = glx::reduce(glx::add(y[4 + i], t)); s[8 + i] = glx::reduce(glx::add(y[8 + i], t));
    }
}

// mode 0: y_i = 2^s x_i + sum; mode 1: (2^s - 1) x_i + sum; mode 2: (2^s + 1) x_i + sum
template <int MODE>
__global__ void __launch_bounds__(128) search_kernel(unsigned long long first, unsigned long long count, int variant, unsigned long long* hits, int* n_hits) {
    __shared__ uint32_t bloom[2048];   // 65536-bit prefilter on the low 16 bits of a lane
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) bloom[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) for (int j = 0; j < N_TARGETS; j++) { uint32_t h = (uint32_t)TARGETS[j] & 0xFFFF; bloom[h >> 5] |= 1u << (h & 31); }
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = first + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < first + count; idx += stride) {
        // decode the ordered selection
        unsigned sh[12];
        unsigned long long r = idx;
        uint32_t avail = 0xFFFF;
#pragma unroll
        for (int i = 11; i >= 0; i--) {   // position i chooses among 16 - (11 - i) ... fixed radices: pos 0 has radix 16
        }
        unsigned digits[12];
#pragma unroll
        for (int i = 11; i >= 0; i--) { unsigned radix = 16 - i; digits[i] = (unsigned)(r % radix); r /= radix; }
#pragma unroll
        for (int i = 0; i < 12; i++) {
            uint32_t m = avail; unsigned k = digits[i];
            for (unsigned t = 0; t < k; t++) m &= m - 1;     // drop the k lowest set bits
            unsigned bit = __ffs(m) - 1;
            sh[i] = bit; avail &= ~(1u << bit);
        }
        uint64_t s[12];
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = MID[variant][i];
#pragma unroll 1
        for (int rd = 4; rd < 26; rd++) {
            s[0] = glx::pow7(glx::add_canon(s[0], RC[12 * rd]));
            glx::w96 sum = glx::widen(s[0]);
#pragma unroll
            for (int i = 1; i < 12; i++) sum = glx::add(sum, s[i]);
#pragma unroll
            for (int i = 0; i < 12; i++) {
                glx::w96 x = glx::widen(s[i]);
                glx::w96 y; y.lo = x.lo << sh[i]; y.hi = sh[i] ? (uint32_t)(x.lo >> (64 - sh[i])) : 0;
                if (MODE == 1) { // (2^s - 1) x = 2^s x - x : add p*2^.. to stay non-negative: use 2^s x + (p - canon(x))
                    y = glx::add(y, GL_P - glx::canon(s[i]));
                } else if (MODE == 2) y = glx::add(y, x);
                s[i] = glx::reduce(glx::add(y, sum));
            }
        }
#pragma unroll 1
        for (int rd = 26; rd < 30; rd++) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = glx::pow7(glx::add_canon(s[i], RC[12 * rd + i]));
            ext_layer(s);
        }
        bool maybe = false;
#pragma unroll
        for (int i = 0; i < 12; i++) { s[i] = glx::canon(s[i]); uint32_t h = (uint32_t)s[i] & 0xFFFF; maybe |= (bloom[h >> 5] >> (h & 31)) & 1; }
        if (maybe) {
            for (int i = 0; i < 12; i++) for (int j = 0; j < N_TARGETS; j++) if (s[i] == TARGETS[j]) { int k = atomicAdd(n_hits, 1); if (k < 64) hits[k] = idx; }
        }
    }
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint64_t hdr[2];
    if (fread(hdr, 8, 2, f) != 2) return 2;
    int n_var = (int)hdr[0], n_t = (int)hdr[1];
    std::vector<uint64_t> mid(n_var * 12), tg(n_t);
    if (fread(mid.data(), 8, mid.size(), f) != mid.size() || fread(tg.data(), 8, tg.size(), f) != tg.size()) return 2;
    fclose(f);
    double f0 = argc > 2 ? atof(argv[2]) : 0.0, f1 = argc > 3 ? atof(argv[3]) : 1.0;
    int modes = argc > 4 ? atoi(argv[4]) : 1;
    int vmask = argc > 5 ? atoi(argv[5]) : 0xFF;
    cudaMemcpyToSymbol(MID, mid.data(), mid.size() * 8);
    cudaMemcpyToSymbol(TARGETS, tg.data(), tg.size() * 8);
    cudaMemcpyToSymbol(N_TARGETS, &n_t, 4);
    unsigned long long total = 1; for (int i = 0; i < 12; i++) total *= 16 - i;
    unsigned long long first = (unsigned long long)(f0 * total), last = (unsigned long long)(f1 * total);
    unsigned long long* d_hits; int* d_n;
    cudaMalloc(&d_hits, 64 * 8); cudaMalloc(&d_n, 4);
    for (int v = 0; v < n_var; v++)
        for (int mode = 0; mode < modes; mode++) {
            if (!((vmask >> v) & 1)) continue;
            cudaMemset(d_n, 0, 4);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0);
            const unsigned long long chunk = 1ull << 34;
            for (unsigned long long a = first; a < last; a += chunk) {
                unsigned long long c = last - a < chunk ? last - a : chunk;
                if (mode == 0) search_kernel<0><<<148 * 16, 128>>>(a, c, v, d_hits, d_n);
                else if (mode == 1) search_kernel<1><<<148 * 16, 128>>>(a, c, v, d_hits, d_n);
                else search_kernel<2><<<148 * 16, 128>>>(a, c, v, d_hits, d_n);
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            int n; unsigned long long h[64];
            cudaMemcpy(&n, d_n, 4, cudaMemcpyDeviceToHost); cudaMemcpy(h, d_hits, 64 * 8, cudaMemcpyDeviceToHost);
            printf("variant %d mode %d: %llu candidates in %.1f s (%.2f G/s), hits %d, err %s\n", v, mode, last - first, ms / 1e3, (last - first) / ms / 1e6, n,
                   cudaGetErrorString(cudaGetLastError()));
            for (int i = 0; i < n && i < 64; i++) {
                unsigned long long r = h[i]; unsigned digits[12]; for (int k = 11; k >= 0; k--) { digits[k] = r % (16 - k); r /= (16 - k); }
                unsigned avail = 0xFFFF; printf("  HIT idx %llu exps:", h[i]);
                for (int k = 0; k < 12; k++) { unsigned m = avail; for (unsigned t = 0; t < digits[k]; t++) m &= m - 1; unsigned bit = __builtin_ffs(m) - 1; printf(" %u", bit); avail &= ~(1u << bit); }
                printf("\n");
            }
            fflush(stdout);
        }
    return 0;
}
