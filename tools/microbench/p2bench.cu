// p2bench.cu -- Poseidon2 permutation throughput (compile with -DTPB=threads per CTA)
#include <cstdio>
#include "../../era_zkevm_test_harness_b200/csrc/poseidon2_core.cuh"
#include "../../era_zkevm_test_harness_b200/csrc/poseidon2_consts.cuh"
#ifndef TPB
#define TPB 128
#endif
__constant__ uint64_t RC[360] = {ZK_P2_RC_INIT};
#ifndef TWO
#define TWO 0
#endif
#if TWO
// two independent permutations per thread, interleaved round by round (more ILP, twice the registers)
__global__ void __launch_bounds__(TPB) k(uint64_t* st, int reps) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    uint64_t s[12], t[12];
    for (int j = 0; j < 12; j++) { s[j] = st[i * 12 + j]; t[j] = st[(i + 1) * 12 + j]; }
    for (int r = 0; r < reps; r++) {
        zk::p2x_external(s); zk::p2x_external(t);
        int rr = 0;
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
#pragma unroll 1
            for (int q = 0; q < 4; q++, rr++) { zk::p2x_full_round(s, RC, rr); zk::p2x_full_round(t, RC, rr); }
            if (half == 0) {
#pragma unroll 1
                for (int q = 0; q < 22; q++, rr++) {
                    s[0] = glx::pow7(glx::add_canon(s[0], RC[12 * rr]));
                    t[0] = glx::pow7(glx::add_canon(t[0], RC[12 * rr]));
                    zk::p2x_internal(s); zk::p2x_internal(t);
                }
            }
        }
    }
    for (int j = 0; j < 12; j++) { st[i * 12 + j] = glx::canon(s[j]); st[(i + 1) * 12 + j] = glx::canon(t[j]); }
}
#else
__global__ void __launch_bounds__(TPB) k(uint64_t* st, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s[12];
    for (int j = 0; j < 12; j++) s[j] = st[i * 12 + j];
    for (int r = 0; r < reps; r++) zk::p2x_permute(s, RC);
    for (int j = 0; j < 12; j++) st[i * 12 + j] = glx::canon(s[j]);
}
#endif
int main() {
    const size_t n = (size_t)148 * 2048 * 4;
    uint64_t* d;
    cudaMalloc(&d, n * 96);
    cudaMemset(d, 1, n * 96);
    const int reps = 8;
    k<<<n / TPB / (TWO ? 2 : 1), TPB>>>(d, reps);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<n / TPB / (TWO ? 2 : 1), TPB>>>(d, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("two=%d tpb %d: %.3f ms, %.3f G perm/s  (check %016llx)\n", TWO, TPB, ms, n * reps / (ms * 1e-3) / 1e9, (unsigned long long)h[0]);
    return 0;
}
