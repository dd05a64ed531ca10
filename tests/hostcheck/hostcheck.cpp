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
