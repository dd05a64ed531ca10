// ntt1024_core.cuh -- one 1024-point Goldilocks NTT as 32 x 32 (four-step) with both 32-point transforms held in registers.
//
//   X[kb + 32*ka] = sum_a  lambda^(a*ka) * [ rho^(a*kb) * sum_b u[a + 32*b] * lambda^(b*kb) ],   lambda = rho^32
//
// rho is a primitive 1024-th root of unity (omega_1024 forward, its inverse backward), so lambda is a primitive 32nd root
// and in Goldilocks that is a power of two: omega_32 = 2^78, omega_32^-1 = 2^114 (2 has order 192; checked at context
// creation).  Every butterfly twiddle of the two 32-point transforms is therefore +-2^s -- shifts and one reduction
// (glx::mul_2exp) instead of a 64x64 multiply -- and the only general multiplications are the 32x32 table rho^(a*kb)
// between the two steps (with the coset factor of lane a folded in).  The B200 has no 64-bit integer multiplier, so this
// is what decides the NTT's speed: the kernel is integer-issue-bound, not HBM-bound.
//
// The functions are host/device so the CPU suite can run the exact lane program against the oracle NTT
// (tests/hostcheck, tests/test_ntt1024_cpu.py).
#pragma once
#include "glx.cuh"

namespace zk {

constexpr int NTT32_E_FWD = 78;    // omega_32      = 2^78
constexpr int NTT32_E_INV = 114;   // omega_32^(-1) = 2^(192-78)

GL_HD constexpr int brev5(int x) { return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4); }

template <int E>
GL_HD uint64_t tw_butterfly(uint64_t x, uint64_t y) {  // (x - y) * 2^E, E in [0, 192)
    if constexpr (E == 0) return gl::sub(x, y);
    else if constexpr (E < 96) return glx::mul_2exp<E>(gl::sub(x, y));
    else if constexpr (E == 96) return gl::sub(y, x);
    else return glx::mul_2exp<E - 96>(gl::sub(y, x));
}

template <int E32, int Q, int BLK, int J>
GL_HD void ntt32_bfly(uint64_t (&v)[32]) {
    constexpr int H = 16 >> Q;
    constexpr int I = BLK * 2 * H + J;
    constexpr int E = (E32 * (1 << Q) * J) % 192;
    uint64_t x = v[I], y = v[I + H];
    v[I] = gl::add(x, y);
    v[I + H] = tw_butterfly<E>(x, y);
}
template <int E32, int Q, int T>
GL_HD void ntt32_stage_items(uint64_t (&v)[32]) {  // butterfly T of stage Q (16 per stage), then the next one
    constexpr int H = 16 >> Q;
    ntt32_bfly<E32, Q, T / H, T % H>(v);
    if constexpr (T + 1 < 16) ntt32_stage_items<E32, Q, T + 1>(v);
}
// radix-2 decimation in frequency: natural-order input, BIT-REVERSED output: v[r] = V[brev5(r)],
// V[k] = sum_b v_in[b] * (2^E32)^(b*k)
template <int E32>
GL_HD void ntt32(uint64_t (&v)[32]) {
    ntt32_stage_items<E32, 0, 0>(v);
    ntt32_stage_items<E32, 1, 0>(v);
    ntt32_stage_items<E32, 2, 0>(v);
    ntt32_stage_items<E32, 3, 0>(v);
    ntt32_stage_items<E32, 4, 0>(v);
}

// ---- lane programs (lane = a in step 1, lane = kb in step 2); buf is the member's 33-padded exchange buffer ----
// step 1 + twiddle: v[b] = u[a + 32*b] on entry (pre-scaling already applied); writes Y'[a][kb] to buf[33*kb + a]
template <int E32>
GL_HD void ntt1024_step1(uint64_t (&v)[32], int a, const uint64_t* twid /* [kb*32 + a] */, uint64_t* buf) {
    ntt32<E32>(v);
#pragma unroll
    for (int r = 0; r < 32; r++) {
        const int kb = brev5(r);
        buf[33 * kb + a] = gl::mul(v[r], twid[kb * 32 + a]);
    }
}
// step 2: reads Y'[a][kb] for a = 0..31, leaves X[kb + 32*brev5(r)] in v[r]
template <int E32>
GL_HD void ntt1024_step2(uint64_t (&v)[32], int kb, const uint64_t* buf) {
#pragma unroll
    for (int a = 0; a < 32; a++) v[a] = buf[33 * kb + a];
    ntt32<E32>(v);
}
// index of v[r] of lane kb after step 2: natural index k, or its position in the bit-reversed enumeration
GL_HD constexpr int ntt1024_k(int kb, int r) { return kb + 32 * brev5(r); }
GL_HD constexpr int ntt1024_pos(int kb, int r) { return 32 * brev5(kb) + r; }  // = bitrev10(k)

}  // namespace zk
