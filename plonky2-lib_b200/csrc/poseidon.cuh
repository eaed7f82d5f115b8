// Poseidon-Goldilocks permutation (width 12, rate 8, x^7, 4 + 22 + 4 rounds), one state per thread.
//
// plonky2::hash::poseidon::{Poseidon::poseidon, PoseidonHash} -- SURVEY 8a P5.  Reached in the
// reference from PoseidonHash::two_to_one (src/smt/goldilocks_poseidon/mod.rs:165,
// src/zkdsa/account.rs:165, src/zkdsa/circuits/mod.rs:66-67), PoseidonHash::hash_pad
// (src/smt/goldilocks_poseidon/mod.rs:170) and, through data.prove(pw), from MerkleTree::new.
//
// B200 mapping: the 12-lane state lives in 24 registers of one thread; round constants sit in
// __constant__ memory (every lane of a warp reads the same word -> one broadcast).  The MDS layer
// uses that the circulant row sums to 256 (+8 on the diagonal): each 64-bit lane is split into two
// 32-bit halves, the two half-dot-products are 12 carry-free IMAD.WIDE.U32 each (< 2^42), and one
// more IMAD.WIDE folds them back into [0, 2^64).  The S-box is four 64x64 products (4 IMAD.WIDE
// each) with the Goldilocks shift-reduction on the ALU pipe.
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

#ifdef __CUDACC__
GL_D u64 poseidon_sbox(u64 x) {
    u64 x2 = gl_sqr(x);
    u64 x4 = gl_sqr(x2);
    u64 x3 = gl_mul(x2, x);
    return gl_mul(x3, x4);
}

// out[r] = sum_i s[(i + r) % 12] * CIRC[i] + s[r] * DIAG[r],  CIRC = 17 15 41 16 2 28 13 13 39 18 34 20, DIAG = 8 0 ...
GL_D void poseidon_mds(u64 s[12]) {
    constexpr u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u32 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = (u32)s[i];
        hi[i] = (u32)(s[i] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (u64)lo[(i + r) % 12] * C[i];
            ah += (u64)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) {
            al += (u64)lo[0] * 8u;
            ah += (u64)hi[0] * 8u;
        }
        // value = al + 2^32 * ah, al, ah < 2^42.  2^64 = 2^32 - 1:
        u64 x = al + (u64)(u32)(ah >> 32) * GL_EPS;  // < 2^43
        u64 y = x + (ah << 32);
        s[r] = y + ((y < x) ? GL_EPS : 0ULL);        // true value < 2^64 + 2^43: one wrap at most
    }
}

GL_D void poseidon_permute(u64 s[12]) {
    const u64* rc = c_poseidon_rc;
#pragma unroll 1
    for (int r = 0; r < POSEIDON_FULL_HALF; r++, rc += 12) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add_c(s[i], rc[i]));
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int r = 0; r < POSEIDON_PARTIAL; r++, rc += 12) {
#pragma unroll
        for (int i = 1; i < 12; i++) s[i] = gl_add_c(s[i], rc[i]);
        s[0] = poseidon_sbox(gl_add_c(s[0], rc[0]));
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int r = 0; r < POSEIDON_FULL_HALF; r++, rc += 12) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add_c(s[i], rc[i]));
        poseidon_mds(s);
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
