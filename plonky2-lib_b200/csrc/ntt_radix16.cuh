// Radix-16 decimation-in-frequency block in registers (shared by the NTT kernels).
//
// 2^192 = 1 (mod p): every root of unity of order <= 64 is a power of two (plonky2's w_16 is 2^156), so the 17
// twiddles inside a radix-16 block are compile-time constants with one or two set bits per 32-bit half and their
// products compile to shifts.  plonky2_field::fft (fft_classic) reaches the same values through its root table.
#pragma once
#include "gl_field.cuh"

__host__ __device__ constexpr u64 pow2_mod_p(int e) {  // 2^e for 0 <= e < 96
    return e < 64 ? ((u64)1 << e) : (((u64)1 << (e - 32)) - ((u64)1 << (e - 64)));
}

// (u - v) * w_16^(J * STEP), w_16 = 2^156 (forward) or 2^36 (inverse); 2^96 = -1 flips the subtraction
template <bool INV, int J>
GL_D u64 diff_times_w16(u64 u, u64 v) {
    constexpr int E = ((INV ? 36 : 156) * J) % 192;
    if (E == 0) return gl_sub(u, v);
    if (E >= 96) return gl_mul(gl_sub(v, u), pow2_mod_p(E - 96));
    return gl_mul(gl_sub(u, v), pow2_mod_p(E));
}

template <bool INV>
GL_D void radix16_dif(u64 x[16]) {
#define BF(i, j, J)                               \
    {                                             \
        u64 u_ = x[i], v_ = x[j];                 \
        x[i] = gl_add(u_, v_);                    \
        x[j] = diff_times_w16<INV, J>(u_, v_);    \
    }
    BF(0, 8, 0) BF(1, 9, 1) BF(2, 10, 2) BF(3, 11, 3) BF(4, 12, 4) BF(5, 13, 5) BF(6, 14, 6) BF(7, 15, 7)
    BF(0, 4, 0) BF(1, 5, 2) BF(2, 6, 4) BF(3, 7, 6) BF(8, 12, 0) BF(9, 13, 2) BF(10, 14, 4) BF(11, 15, 6)
    BF(0, 2, 0) BF(1, 3, 4) BF(4, 6, 0) BF(5, 7, 4) BF(8, 10, 0) BF(9, 11, 4) BF(12, 14, 0) BF(13, 15, 4)
    BF(0, 1, 0) BF(2, 3, 0) BF(4, 5, 0) BF(6, 7, 0) BF(8, 9, 0) BF(10, 11, 0) BF(12, 13, 0) BF(14, 15, 0)
#undef BF
}

GL_D unsigned brev4(unsigned x) { return __brev(x) >> 28; }

