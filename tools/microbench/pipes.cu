// pipes.cu -- integer pipe throughput on sm_100a: how many IMAD / IMAD.WIDE / IMAD.HI / IADD3 / mixed sequences an SM
// retires per cycle.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed) {
    uint32_t a[CHAINS], b[CHAINS];
    uint64_t w[CHAINS];
    for (int i = 0; i < CHAINS; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + blockIdx.x; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(b[i]));                 // IMAD
            if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));     // IMAD.WIDE.U32
            if (OP == 2) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(b[i]));                 // IMAD.HI.U32
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                        // IADD3
            if (OP == 4) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(it)); // IADD3 + IADD3.X
            if (OP == 5) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));   // 1 wide : 2 adds
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                           asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
            if (OP == 6) { asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(b[i]));               // 1 imad : 1 add
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
            if (OP == 7) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(a[i]) : "r"(b[i]));              // SHF
            if (OP == 9) asm volatile("addc.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                       // IADD3.X alone
            if (OP == 10) asm volatile("{.reg .pred q; setp.lt.u32 q, %0, %1; selp.u32 %0, %1, %0, q;}" : "+r"(a[i]) : "r"(b[i]));  // ISETP + SEL
            if (OP == 11) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));   // LOP3
            if (OP == 12) asm volatile("prmt.b32 %0, %0, %1, 0x1032;" : "+r"(a[i]) : "r"(b[i]));              // PRMT
            if (OP == 13) asm volatile("{.reg .pred q; setp.lt.u32 q, %0, %1; @q add.u32 %0, %0, %1;}" : "+r"(a[i]) : "r"(b[i]));   // ISETP + predicated IADD3
            if (OP == 14) asm volatile("selp.u32 %0, %1, %0, %2;" : "+r"(a[i]) : "r"(b[i]), "r"((int)(it & 1)));  // hmm: setp+sel via int pred
            if (OP == 15) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.u32 %0, %0, 0;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(it)); // 3-chain
            if (OP == 16) asm volatile("sub.cc.u32 %0, %0, %2;\n\tsubc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(it)); // 64-bit sub
            if (OP == 17) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(((uint64_t)b[i] << 32) | a[i]));   // 64-bit add (compiler form)
            if (OP == 18) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i] ^ (uint32_t)w[i]), "r"(b[i]));  // IMAD.WIDE no addend (+xor)
            if (OP == 19) asm volatile("{.reg .pred q; setp.lt.u32 q, %0, %1; selp.u32 %0, %2, %0, q;}" : "+r"(a[i]) : "r"(b[i]), "r"(seed));  // ISETP + SEL (3 operands)
            if (OP == 20) asm volatile("{.reg .pred q; setp.lt.u32 q, %1, %2; selp.u32 %0, %2, %0, q; selp.u32 %1, %0, %1, q;}" : "+r"(a[i]), "+r"(b[i]) : "r"(seed));  // ISETP + 2 SEL
            if (OP == 21) asm volatile("{.reg .pred q; setp.lt.u32 q, %1, %2; setp.lt.and.u32 q, %0, %2, q; selp.u32 %0, %2, %0, q;}" : "+r"(a[i]), "+r"(b[i]) : "r"(seed));  // 2 ISETP + SEL
            if (OP == 22) asm volatile("add.cc.u32 %0, %0, %2;\n\tmadc.lo.u32 %1, %1, 1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(it)); // IADD3 + IMAD.X?
            if (OP == 23) asm volatile("mov.b32 %0, %1;" : "=r"(a[i]) : "r"(b[i] + it));   // MOV-ish
            if (OP == 8) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(it)); // fused wide w/ carry
        }
    }
    uint32_t acc = 0;
    for (int i = 0; i < CHAINS; i++) acc ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char* name, int ops_per_item) {
    int dev = 0, sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    uint32_t* out;
    const int threads = 256, blocks = sms * 8;
    cudaMalloc(&out, (size_t)threads * blocks * 4);
    k<OP><<<blocks, threads>>>(out, 12345);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)threads * blocks * ITERS * CHAINS * ops_per_item;
    double per_clk_sm = ops / (ms * 1e-3) / (1.965e9) / sms;   // assumes 1965 MHz
    printf("%-28s %8.3f ms  %7.1f thread-ops/clk/SM (of 128 issue slots)\n", name, ms, per_clk_sm);
    cudaFree(out);
}

int main() {
    run<0>("IMAD (mad.lo.u32)", 1);
    run<1>("IMAD.WIDE.U32", 1);
    run<2>("IMAD.HI.U32", 1);
    run<3>("IADD3 (add.u32)", 1);
    run<4>("IADD3+IADD3.X (64-bit add)", 2);
    run<5>("1 WIDE + add + xor", 3);
    run<6>("1 IMAD + 1 add", 2);
    run<7>("SHF", 1);
    run<8>("mad.lo.cc+madc.hi (fused?)", 2);
    run<19>("ISETP + SEL (3 operands)", 2);
    run<20>("ISETP + 2 SEL", 3);
    run<21>("2 ISETP + SEL", 3);
    run<22>("add.cc + madc.lo (IADD3+IMAD.X?)", 2);
    run<9>("IADD3.X alone (addc)", 1);
    run<10>("ISETP+SEL", 2);
    run<11>("LOP3", 1);
    run<12>("PRMT", 1);
    run<13>("ISETP + @p IADD3", 2);
    run<15>("add.cc+addc.cc+addc (3 chain)", 3);
    run<16>("sub.cc+subc (64-bit sub)", 2);
    run<17>("add.u64", 1);
    run<18>("mul.wide + xor", 2);
    return 0;
}
