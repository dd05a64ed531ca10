// hostcheck.cpp -- compiles the host halves of the product's device headers (glx.cuh, poseidon2_core.cuh) with g++ so the
// CPU test suite can check their arithmetic against the oracle without a GPU.  Test infrastructure only.
#include <cstddef>
#include <cstdint>
#include "../../era_zkevm_test_harness_b200/csrc/poseidon2_core.cuh"
#include "../../era_zkevm_test_harness_b200/csrc/poseidon2_consts.cuh"

static const uint64_t RC[360] = {ZK_P2_RC_INIT};
#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT void hc_glx_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = glx::canon(glx::mul(a[i], b[i])); }
EXPORT void hc_glx_sub(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = glx::canon(glx::sub(a[i], b[i])); }
EXPORT void hc_glx_reduce96(const uint64_t* lo, const uint32_t* hi, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = glx::canon(glx::reduce96(lo[i], hi[i])); }
EXPORT void hc_p2x_permute(uint64_t* states, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint64_t s[12];
        for (int k = 0; k < 12; k++) s[k] = states[12 * i + k];
        zk::p2x_permute(s, RC);
        for (int k = 0; k < 12; k++) states[12 * i + k] = glx::canon(s[k]);
    }
}

template <int S> static void shl_case(const uint64_t* a, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = glx::mul_2exp<S>(a[i]); }
template <int S> struct ShlTable { static void fill(void (**t)(const uint64_t*, uint64_t*, size_t)) { t[S] = &shl_case<S>; ShlTable<S - 1>::fill(t); } };
template <> struct ShlTable<-1> { static void fill(void (**)(const uint64_t*, uint64_t*, size_t)) {} };
// o[i] = a[i] * 2^s mod p (canonical a), 0 <= s < 96
EXPORT int hc_glx_mul_2exp(const uint64_t* a, uint64_t* o, size_t n, int s) {
    static void (*table[96])(const uint64_t*, uint64_t*, size_t);
    static bool init = false;
    if (!init) { ShlTable<95>::fill(table); init = true; }
    if (s < 0 || s >= 96) return 1;
    table[s](a, o, n);
    return 0;
}
EXPORT void hc_glx_reduce128(const uint64_t* lo, const uint64_t* hi, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = glx::canon(glx::reduce128(lo[i], hi[i])); }

// ---- csrc/ntt1024_core.cuh: run the two lane programs of the 32x32 four-step NTT sequentially over the 32 lanes ----
#include <vector>
#include "../../era_zkevm_test_harness_b200/csrc/ntt1024_core.cuh"
template <int E32>
static void ntt1024_emulate(const uint64_t* in, uint64_t* out_natural, uint64_t* out_bitrev, uint64_t rho) {
    std::vector<uint64_t> twid(1024), buf(33 * 32 + 32);
    for (int kb = 0; kb < 32; kb++)
        for (int a = 0; a < 32; a++) twid[kb * 32 + a] = gl::pow(rho, (uint64_t)a * kb);
    for (int a = 0; a < 32; a++) {
        uint64_t v[32];
        for (int b = 0; b < 32; b++) v[b] = in[a + 32 * b];
        zk::ntt1024_step1<E32>(v, a, twid.data(), buf.data());
    }
    for (int kb = 0; kb < 32; kb++) {
        uint64_t v[32];
        zk::ntt1024_step2<E32>(v, kb, buf.data());
        for (int r = 0; r < 32; r++) {
            out_natural[zk::ntt1024_k(kb, r)] = v[r];
            out_bitrev[zk::ntt1024_pos(kb, r)] = v[r];
        }
    }
}
// forward: X[k] = sum x[i] w^(ik), w = omega_1024; inverse: the same with w^-1 (NO 1/n scaling)
EXPORT void hc_ntt1024(const uint64_t* in, uint64_t* out_natural, uint64_t* out_bitrev, int inverse) {
    uint64_t w = gl::omega(10);
    if (inverse) ntt1024_emulate<zk::NTT32_E_INV>(in, out_natural, out_bitrev, gl::inv(w));
    else ntt1024_emulate<zk::NTT32_E_FWD>(in, out_natural, out_bitrev, w);
}
