// p2bench.cu -- Poseidon2 permutation throughput (compile with -DTPB=threads per CTA)
#include <cstdio>
#include "../../era_zkevm_test_harness_b200/csrc/poseidon2_core.cuh"
#include "../../era_zkevm_test_harness_b200/csrc/poseidon2_consts.cuh"
#ifndef TPB
#define TPB 128
#endif
__constant__ uint64_t RC[360] = {ZK_P2_RC_INIT};
__global__ void __launch_bounds__(TPB) k(uint64_t* st, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s[12];
    for (int j = 0; j < 12; j++) s[j] = st[i * 12 + j];
    for (int r = 0; r < reps; r++) zk::p2x_permute(s, RC);
    for (int j = 0; j < 12; j++) st[i * 12 + j] = glx::canon(s[j]);
}
int main() {
    const size_t n = (size_t)148 * 2048 * 4;
    uint64_t* d;
    cudaMalloc(&d, n * 96);
    cudaMemset(d, 1, n * 96);
    const int reps = 8;
    k<<<n / TPB, TPB>>>(d, reps);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<n / TPB, TPB>>>(d, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    uint64_t h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("tpb %d: %.3f ms, %.3f G perm/s  (check %016llx)\n", TPB, ms, n * reps / (ms * 1e-3) / 1e9, (unsigned long long)h[0]);
    return 0;
}
