// Poseidon-Goldilocks permutation (width 12, rate 8, x^7, 4 + 22 + 4 rounds), one state per thread.
//
// plonky2::hash::poseidon::{Poseidon::poseidon, PoseidonHash} -- SURVEY 8a P5.  Reached in the
// reference from PoseidonHash::two_to_one (src/smt/goldilocks_poseidon/mod.rs:165,
// src/zkdsa/account.rs:165, src/zkdsa/circuits/mod.rs:66-67), PoseidonHash::hash_pad
// (src/smt/goldilocks_poseidon/mod.rs:170) and, through data.prove(pw), from MerkleTree::new.
//
// B200 mapping (measured pipe rates: profiles/r1_pipe_peaks.json).  IMAD.WIDE.U32 issues at ~51
// lanes/clk/SM and does not overlap with ALU work, so the two halves of a round go to different
// pipes:
//   * S-box x^7: four 64x64 products = IMAD.WIDE.U32 on the FMA-heavy pipe + shift-reduction on ALU.
//   * MDS layer: on the FP64 pipe (B200 keeps full-rate DFMA).  A 32-bit half x of a lane is
//     reinterpreted as the double with bit pattern (hi = 0, lo = x), i.e. the denormal x * 2^-1074.
//     The circulant entries are <= 41 and each row sums to 264, so sum_i c_i * x_i < 2^42 stays below
//     2^52: every DFMA is exact and the accumulator's BIT PATTERN is the integer sum.  No int<->fp
//     conversion instruction is ever issued, in either direction.
//   * The next round's constants are the addend of the first DFMA of each row (a constant-bank
//     operand), so "add round constants" costs nothing.
//   * The 22 partial rounds run as 11 fused pairs (poseidon_partial_pair): two linear layers minus the
//     lane-0 path are one matrix N = M M' with entries < 2^14, still exact in FP64.
// The 12-lane state lives in 24 registers of one thread.  Measured: 1.39 G permutations/s on one B200.
#pragma once
#include "gl_field.cuh"

#define POSEIDON_WIDTH 12
#define POSEIDON_RATE 8
#define POSEIDON_FULL_HALF 4
#define POSEIDON_PARTIAL 22
#define POSEIDON_ROUNDS 30

// ALL_ROUND_CONSTANTS[round * 12 + lane]; filled by gl_poseidon_upload_constants() at ctx creation.
// Defined here (not extern): include this header from exactly one translation unit (hash_kernels.cu).
__constant__ u64 c_poseidon_rc[POSEIDON_ROUNDS * POSEIDON_WIDTH];
// The same constants split for the FP64 MDS: [round][lane] -> (double with bits rc & 0xffffffff,
// double with bits rc >> 32); one extra all-zero round so the last MDS adds nothing.
__constant__ double2 c_poseidon_rc_split[(POSEIDON_ROUNDS + 1) * POSEIDON_WIDTH];

#ifdef __CUDACC__
GL_D u64 poseidon_sbox(u64 x) {
    u64 x2 = gl_sqr(x);
    u64 x4 = gl_sqr(x2);
    u64 x3 = gl_mul(x2, x);
    return gl_mul(x3, x4);
}

GL_D double u32_as_denormal(u32 x) { return __hiloint2double(0, (int)x); }
GL_D u64 double_bits(double d) { return (u64)__double_as_longlong(d); }

#ifndef POSEIDON_NO_SPLIT_SBOX
#define POSEIDON_SPLIT_SBOX 1
#endif
#ifdef POSEIDON_SPLIT_SBOX
// x^7 handed to the FP64 linear layer without the last modular reduction: the 128-bit product x^3 * x^4 =
// r0 + r1 phi + r2 phi^2 + r3 phi^3 (phi = 2^32, phi^2 = phi - 1, phi^3 = -1) is (r0 - r2 - r3) + (r1 + r2) phi, and
// the two coefficients are formed by three exact FP64 additions on the limbs' denormal images: dl in (-2^33, 2^32),
// dh in [0, 2^33).  The linear layers are exact on signed integers below 2^52; the constants they add carry an offset
// that is a multiple of p and makes every sum positive again (POSEIDON_OFF_*), so the fold is unchanged.
GL_D void poseidon_sbox_split(u64 x, double& dl, double& dh) {
    const u64 x2 = gl_sqr(x), x4 = gl_sqr(x2), x3 = gl_mul(x2, x);
    const unsigned __int128 p = (unsigned __int128)x3 * x4;
    const u64 lo = (u64)p, hi = (u64)(p >> 64);
    const double r0 = u32_as_denormal((u32)lo), r1 = u32_as_denormal((u32)(lo >> 32));
    const double r2 = u32_as_denormal((u32)hi), r3 = u32_as_denormal((u32)(hi >> 32));
    dl = (r0 - r2) - r3;
    dh = r1 + r2;
}
#endif
// (the offsets live in the constant tables: poseidon_constants.h, linear_layer_tables(signed_sbox = true).  Exactness: every
// operand is an integer multiple of 2^-1074, so FP64 sums and products are exact while they stay below 2^53: a full
// round's layer sees |inputs| < 2^33 and peaks below 2^44; the fused pair's CIRC^2 form sees lane 0 in (-2^33, 2^32) and the
// other lanes in [0, 2^32) and peaks below 2^49, its outputs landing in [0, 2^50) after the 2^48 offset.)

// MDS entries: M[r][j] = CIRC[(j - r) mod 12] + (r == j ? DIAG[r] : 0)
__host__ __device__ constexpr int poseidon_mds_entry(int r, int j) {
    constexpr int C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    return C[(j - r + 12) % 12] + ((r == 0 && j == 0) ? 8 : 0);
}
// N = M * M' with row 0 of M' zeroed: the part of two consecutive linear layers that does not pass through
// lane 0 (the only lane the S-box touches in a partial round).  Entries < 2^14.
__host__ __device__ constexpr int poseidon_pair_entry(int r, int j) {
    int acc = 0;
    for (int i = 1; i < 12; i++) acc += poseidon_mds_entry(r, i) * poseidon_mds_entry(i, j);
    return acc;
}

// Constants of the 11 fused pairs of partial rounds: K[pair][lane] = sum_{i>=1} M[lane][i] * rc[a+1][i] + rc[a+2][lane]
// (a = 4 + 2 * pair), split in 32-bit-half sums like c_poseidon_rc_split.
__constant__ double2 c_poseidon_pair_k[(POSEIDON_PARTIAL / 2) * POSEIDON_WIDTH];

GL_D u64 poseidon_fold(double al, double ah);
GL_D u64 poseidon_fold(double al, double ah) {
    // value = A + 2^32 * B with A, B < 2^50 (integer bit patterns).  2^64 = 2^32 - 1 (mod p):
    //   t = A + (B >> 32) * (2^32 - 1)  (< 2^51, one IMAD.WIDE);  y = t + (B mod 2^32) * 2^32 wraps at most once
    u64 A = double_bits(al), B = double_bits(ah);
    u64 t = A + (u64)(u32)(B >> 32) * GL_EPS;
    u32 t0 = (u32)t, t1 = (u32)(t >> 32), b0 = (u32)B, y0, y1;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "add.cc.u32 %1, %3, %4;\n\t"
        "addc.u32 m, 0, 0;\n\t"          // carry (add family only)
        "neg.s32 m, m;\n\t"              // carry * (2^32 - 1)
        "add.cc.u32 %0, %2, m;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(y0), "=&r"(y1)
        : "r"(t0), "r"(t1), "r"(b0));
    return gl_pack(y0, y1);
}

// Two consecutive partial rounds a, a + 1 in one pass over the FP64 pipe.  On entry the state holds the input
// of round a's S-box (constants already added).  With x~ = state after that S-box (lane 0 only):
//   y0 = M[0] . x~ + rc[a+1][0]                     -> S-box of round a + 1 -> sigma
//   z  = N x~ + K + M[.][0] * sigma                 (N = M M' without the lane-0 path)
// 336 DFMA instead of 576, 13 folds instead of 24; every partial sum stays below 2^50 (exact).
// ------------------------------------------------------------------------------------------------------------
// Circulant products through a 4 x 3 Good-Thomas FFT (exact, FP64 pipe).
//
// y_r = sum_i CIRC[i] x_{(i+r) mod 12} is a cyclic convolution over Z_12 = Z_4 x Z_3 (index n <-> (n mod 4, n mod 3),
// no twiddles).  A real FFT of length 4 along the Z_4 axis leaves, for the frequencies 0, 1 (complex) and 2, one
// length-3 cyclic convolution each; the inverse FFT's 1/4 is absorbed into the kernel, which stays integral
// because plonky2's MDS row was chosen that way: FFT4(row)/4 = {16,32,16}, {(2,1),(-1,4),(-16,1)}/.., {-1,8,2}.
// 90 FP64 operations per 12-vector instead of 144 multiply-adds; every intermediate is an integer below 2^50 in
// units of 2^-1074 (signed: differences can be negative denormals), so the arithmetic is exact and the final,
// non-negative results are again integer bit patterns.  SQ = 1 applies the circulant twice (kernel convolved with
// itself), which is what the fused partial-round pair needs.
template <int SQ>
GL_D void poseidon_circ12(const double x[12], double y[12]) {
    constexpr double K0[2][3] = {{16., 32., 16.}, {5120., 5120., 6144.}};
    constexpr double K2[2][3] = {{-1., 8., 2.}, {132., -48., 240.}};
    constexpr double KR[2][3] = {{2., -1., -16.}, {54., 486., -162.}};
    constexpr double KI[2][3] = {{1., 4., 1.}, {-252., -36., -72.}};
    constexpr int IDX[4][3] = {{0, 4, 8}, {9, 1, 5}, {6, 10, 2}, {3, 7, 11}};   // lane with (n mod 4, n mod 3) = (a, b)
    double U0[3], U2[3], UR[3], UI[3];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const double p0 = x[IDX[0][b]], p1 = x[IDX[1][b]], p2 = x[IDX[2][b]], p3 = x[IDX[3][b]];
        const double t0 = p0 + p2, t1 = p1 + p3;
        U0[b] = t0 + t1;
        U2[b] = t0 - t1;
        UR[b] = p0 - p2;      // F1 = (p0 - p2) + i (p3 - p1)
        UI[b] = p3 - p1;
    }
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const int b1 = (b + 2) % 3, b2 = (b + 1) % 3;        // (b - 1) mod 3, (b - 2) mod 3
        const double v0 = __fma_rn(K0[SQ][2], U0[b2], __fma_rn(K0[SQ][1], U0[b1], K0[SQ][0] * U0[b]));
        const double v2 = __fma_rn(K2[SQ][2], U2[b2], __fma_rn(K2[SQ][1], U2[b1], K2[SQ][0] * U2[b]));
        double vr = KR[SQ][0] * UR[b], vi = KR[SQ][0] * UI[b];
        vr = __fma_rn(-KI[SQ][0], UI[b], vr);
        vi = __fma_rn(KI[SQ][0], UR[b], vi);
        vr = __fma_rn(KR[SQ][1], UR[b1], vr);
        vi = __fma_rn(KR[SQ][1], UI[b1], vi);
        vr = __fma_rn(-KI[SQ][1], UI[b1], vr);
        vi = __fma_rn(KI[SQ][1], UR[b1], vi);
        vr = __fma_rn(KR[SQ][2], UR[b2], vr);
        vi = __fma_rn(KR[SQ][2], UI[b2], vi);
        vr = __fma_rn(-KI[SQ][2], UI[b2], vr);
        vi = __fma_rn(KI[SQ][2], UR[b2], vi);
        const double sm = v0 + v2, df = v0 - v2;
        y[IDX[0][b]] = sm + vr;
        y[IDX[2][b]] = sm - vr;
        y[IDX[1][b]] = df - vi;
        y[IDX[3][b]] = df + vi;
    }
}

// One full round: S-box on every lane, then out = CIRC x + 8 x_0 e_0 + rc through the FFT form (206 FP64
// operations instead of 290).
GL_D void poseidon_full_round(u64 s[12], const double2* __restrict__ rc) {
    double dl[12], dh[12], yl[12], yh[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
#ifdef POSEIDON_SPLIT_SBOX
        poseidon_sbox_split(s[j], dl[j], dh[j]);
#else
        const u64 x = poseidon_sbox(s[j]);
        dl[j] = u32_as_denormal((u32)x);
        dh[j] = u32_as_denormal((u32)(x >> 32));
#endif
    }
    poseidon_circ12<0>(dl, yl);
    poseidon_circ12<0>(dh, yh);
#pragma unroll
    for (int r = 0; r < 12; r++) {
        const double2 k = rc[r];
        double al = yl[r] + k.x, ah = yh[r] + k.y;
        if (r == 0) {
            al = __fma_rn(8., dl[0], al);
            ah = __fma_rn(8., dh[0], ah);
        }
        s[r] = poseidon_fold(al, ah);
    }
}

// Two consecutive partial rounds in FFT form.  With M = CIRC + 8 e0 e0^T and M' = M without its row 0,
//   N x~ = M M' x~ = CIRC^2 x~ + col0(CIRC) * (8 x~_0 - Yraw),   Yraw = M[0] . x~  (the unfolded y0 without its constant)
// so z = CIRC^2 x~ (90 operations through poseidon_circ12<1>) + col0(CIRC) * w + M[.][0] * sigma + K: 280 FP64
// operations per pair instead of 336, 13 folds.
GL_D void poseidon_partial_pair(u64 s[12], int pair) {
    double dl[12], dh[12];
#pragma unroll
    for (int i = 1; i < 12; i++) {
        dl[i] = u32_as_denormal((u32)s[i]);
        dh[i] = u32_as_denormal((u32)(s[i] >> 32));
    }
    // everything that does not depend on lane 0 first: the FP64 pipe works while lane 0 goes through its S-box
    double yl = (double)poseidon_mds_entry(0, 1) * dl[1], yh = (double)poseidon_mds_entry(0, 1) * dh[1];
#pragma unroll
    for (int j = 2; j < 12; j++) {
        yl = __fma_rn((double)poseidon_mds_entry(0, j), dl[j], yl);
        yh = __fma_rn((double)poseidon_mds_entry(0, j), dh[j], yh);
    }
#ifdef POSEIDON_SPLIT_SBOX
    poseidon_sbox_split(s[0], dl[0], dh[0]);
#else
    const u64 x0 = poseidon_sbox(s[0]);
    dl[0] = u32_as_denormal((u32)x0);
    dh[0] = u32_as_denormal((u32)(x0 >> 32));
#endif
    yl = __fma_rn((double)poseidon_mds_entry(0, 0), dl[0], yl);      // Yraw
    yh = __fma_rn((double)poseidon_mds_entry(0, 0), dh[0], yh);
    const double2 ky = c_poseidon_rc_split[(POSEIDON_FULL_HALF + 2 * pair + 1) * 12];
#ifdef POSEIDON_SPLIT_SBOX
    double gl, gh;
    poseidon_sbox_split(poseidon_fold(yl + ky.x, yh + ky.y), gl, gh);
#else
    const u64 sigma = poseidon_sbox(poseidon_fold(yl + ky.x, yh + ky.y));
    const double gl = u32_as_denormal((u32)sigma), gh = u32_as_denormal((u32)(sigma >> 32));
#endif
    double zl[12], zh[12];
    poseidon_circ12<1>(dl, zl);
    poseidon_circ12<1>(dh, zh);
    const double wl = __fma_rn(8., dl[0], -yl), wh = __fma_rn(8., dh[0], -yh);
    const double2* kk = c_poseidon_pair_k + pair * 12;
#pragma unroll
    for (int r = 0; r < 12; r++) {
        const double2 k = kk[r];
        const double c0 = (double)(poseidon_mds_entry(r, 0) - (r == 0 ? 8 : 0));   // col0(CIRC)
        double tl = __fma_rn((double)poseidon_mds_entry(r, 0), gl, k.x);
        double th = __fma_rn((double)poseidon_mds_entry(r, 0), gh, k.y);
        tl = __fma_rn(c0, wl, tl);
        th = __fma_rn(c0, wh, th);
        s[r] = poseidon_fold(zl[r] + tl, zh[r] + th);
    }
}

GL_D void poseidon_permute(u64 s[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], c_poseidon_rc[i]);
    // One copy of each block (full round, fused partial pair) keeps the kernel inside the instruction cache.
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        // constants of round r + 1 go into the MDS of round r; round 30 is the all-zero row
        const double2* rc = c_poseidon_rc_split + 12 * (phase ? POSEIDON_FULL_HALF + POSEIDON_PARTIAL + 1 : 1);
#pragma unroll 1
        for (int r = 0; r < POSEIDON_FULL_HALF; r++, rc += 12) poseidon_full_round(s, rc);
        if (phase == 0) {
#pragma unroll 1
            for (int pair = 0; pair < POSEIDON_PARTIAL / 2; pair++) poseidon_partial_pair(s, pair);
        }
    }
}

// hashing::compress: state = l || r || 0000, permute, first four lanes
GL_D void poseidon_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        s[i] = l[i];
        s[4 + i] = r[i];
        s[8 + i] = 0;
    }
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl_canon(s[i]);
}
#endif
