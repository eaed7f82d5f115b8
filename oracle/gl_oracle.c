/*
 * gl_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See gl_oracle.h for the
 * parity status ("pinned" for Poseidon, "parity unpinned" for NTT/LDE/Merkle/FRI layouts).
 *
 * Every function states which reference call site reaches it (paths relative to /root/reference)
 * and which upstream (plonky2 v0.1.4, source absent) routine it restates.  Values handed out are
 * always canonical (< p).  OpenMP is used only where upstream uses rayon.
 */
#include "gl_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
static double glo_now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;

#define P GLO_P
#define EPS 0xFFFFFFFFULL /* 2^64 mod p */

/* ------------------------------------------------------------------------------------------------
 * P0  GoldilocksField (plonky2_field::goldilocks_field; type used at
 *     src/smt/goldilocks_poseidon/mod.rs:9,179 and src/zkdsa/circuits/mod.rs:81-100)
 * ---------------------------------------------------------------------------------------------- */
static inline u64 canon(u64 x) { return x >= P ? x - P : x; }

static inline u64 reduce128(u128 x) {
    /* 2^64 = 2^32 - 1, 2^96 = -1 (mod p); written without data-dependent branches */
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hi_hi = hi >> 32, hi_lo = hi & EPS;
    u64 t0 = lo - hi_hi;
    t0 -= EPS & (0 - (u64)(lo < hi_hi)); /* borrow: -2^64 = -(2^32-1) */
    u64 t1 = hi_lo * EPS;
    u64 r = t0 + t1;
    r += EPS & (0 - (u64)(r < t1)); /* carry: +2^64 = +(2^32-1) */
    return canon(r);
}

static inline u64 f_add(u64 a, u64 b) {
    u128 s = (u128)canon(a) + canon(b);
    return s >= P ? (u64)(s - P) : (u64)s;
}
static inline u64 f_sub(u64 a, u64 b) {
    a = canon(a); b = canon(b);
    return a >= b ? a - b : a + (P - b);
}
static inline u64 f_mul(u64 a, u64 b) { return reduce128((u128)a * b); }
u64 glo_add(u64 a, u64 b) { return f_add(a, b); }
u64 glo_sub(u64 a, u64 b) { return f_sub(a, b); }
u64 glo_mul(u64 a, u64 b) { return f_mul(a, b); }
u64 glo_pow(u64 a, u64 e) {
    u64 r = 1;
    a = canon(a);
    while (e) {
        if (e & 1) r = f_mul(r, a);
        a = f_mul(a, a);
        e >>= 1;
    }
    return r;
}
u64 glo_inv(u64 a) { return glo_pow(a, P - 2); }

/* Field::primitive_root_of_unity: POWER_OF_TWO_GENERATOR^(2^(32-k)), POWER_OF_TWO_GENERATOR = 7^((p-1)/2^32) */
u64 glo_primitive_root_of_unity(unsigned lg_n) {
    u64 g = glo_pow(7, (P - 1) >> 32); /* = 1753635133440165772 */
    for (unsigned i = lg_n; i < 32; i++) g = f_mul(g, g);
    return g;
}

/* QuadraticExtension<GoldilocksField>: F[X]/(X^2 - 7), element [a0, a1] */
void glo_ext_mul(const u64 a[2], const u64 b[2], u64 out[2]) {
    u64 c0 = f_add(f_mul(a[0], b[0]), f_mul(7, f_mul(a[1], b[1])));
    u64 c1 = f_add(f_mul(a[0], b[1]), f_mul(a[1], b[0]));
    out[0] = c0;
    out[1] = c1;
}
static void ext_add(const u64 a[2], const u64 b[2], u64 out[2]) {
    out[0] = f_add(a[0], b[0]);
    out[1] = f_add(a[1], b[1]);
}

/* ------------------------------------------------------------------------------------------------
 * P5  Poseidon constants.  plonky2's ALL_ROUND_CONSTANTS table is not on this disk; SURVEY.md 8c:
 *     it is ChaCha8Rng::seed_from_u64(0) sampled with rand 0.8.5 gen_range(0..p)
 *     (rand / rand_chacha versions locked at Cargo.lock:1135-1149).  The result is pinned by the
 *     reference KAT src/zkdsa/circuits/mod.rs:85-105.
 * ---------------------------------------------------------------------------------------------- */
static inline u32 rotl32(u32 x, int k) { return (x << k) | (x >> (32 - k)); }
static inline u32 rotr32(u32 x, unsigned k) { k &= 31; return k ? (x >> k) | (x << (32 - k)) : x; }

static void chacha8_block(const u32 key[8], u64 counter, u32 out[16]) {
    u32 s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = (u32)counter;
    s[13] = (u32)(counter >> 32);
    s[14] = s[15] = 0;
    u32 x[16];
    memcpy(x, s, sizeof x);
#define QR(a, b, c, d)                                                                              \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);     \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    for (int r = 0; r < 4; r++) { /* 8 rounds = 4 double rounds */
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

void glo_poseidon_round_constants(u64 out[360]) {
    /* rand_core::SeedableRng::seed_from_u64(0): PCG32 expands the u64 into the 32-byte key */
    u32 key[8];
    u64 st = 0;
    for (int i = 0; i < 8; i++) {
        st = st * 6364136223846793005ULL + 11634580027462260723ULL;
        u32 xs = (u32)(((st >> 18) ^ st) >> 27);
        key[i] = rotr32(xs, (unsigned)(st >> 59));
    }
    u32 blk[16];
    u64 counter = 0;
    int pos = 16, n = 0;
    while (n < 360) {
        u32 w[2];
        for (int k = 0; k < 2; k++) {
            if (pos == 16) { chacha8_block(key, counter++, blk); pos = 0; }
            w[k] = blk[pos++];
        }
        u64 v = (u64)w[0] | ((u64)w[1] << 32);
        /* UniformInt<u64>::sample_single(0, p): widening multiply, zone = (p << clz(p)) - 1 = p - 1 */
        u128 m = (u128)v * P;
        if ((u64)m <= P - 1) out[n++] = (u64)(m >> 64);
    }
}

static u64 RC[360];
static int rc_ready = 0;
static const u64 MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const u64 MDS_DIAG0 = 8;

static void ensure_rc(void) {
    if (!rc_ready) {
#pragma omp critical(glo_rc)
        {
            if (!rc_ready) {
                glo_poseidon_round_constants(RC);
                rc_ready = 1;
            }
        }
    }
}

static inline u64 sbox7(u64 x) {
    u64 x2 = f_mul(x, x), x4 = f_mul(x2, x2), x3 = f_mul(x, x2);
    return f_mul(x3, x4);
}

/* Poseidon::mds_layer: out[r] = sum_i state[(i+r)%12]*CIRC[i] + state[r]*DIAG[r]  (literal form) */
static inline void mds_layer_naive(u64 s[12]) {
    u64 o[12];
    for (int r = 0; r < 12; r++) {
        u128 acc = 0;
        for (int i = 0; i < 12; i++) acc += (u128)s[(i + r) % 12] * MDS_CIRC[i];
        if (r == 0) acc += (u128)s[0] * MDS_DIAG0;
        o[r] = reduce128(acc);
    }
    memcpy(s, o, sizeof o);
}

/* Same map, arranged like upstream's mds_row_shf on 32-bit halves: the circulant entries sum to
 * 256, so the low-half and high-half dot products stay below 2^41 and need no carries. */
static inline void mds_layer(u64 s[12]) {
    u64 lo[24], hi[24], ol[12], oh[12];
    for (int i = 0; i < 12; i++) {
        lo[i] = lo[i + 12] = s[i] & EPS;
        hi[i] = hi[i + 12] = s[i] >> 32;
    }
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
        for (int i = 0; i < 12; i++) {
            al += lo[r + i] * MDS_CIRC[i];
            ah += hi[r + i] * MDS_CIRC[i];
        }
        ol[r] = al;
        oh[r] = ah;
    }
    ol[0] += lo[0] * MDS_DIAG0;
    oh[0] += hi[0] * MDS_DIAG0;
    for (int r = 0; r < 12; r++) s[r] = reduce128((u128)ol[r] + ((u128)oh[r] << 32));
}

static void permute_with(u64 s[12], void (*mds)(u64 *)) {
    ensure_rc();
    int rc = 0;
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
    for (int r = 0; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = f_add(s[i], RC[rc++]);
        if (r < 4 || r >= 26) {
            for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        } else {
            s[0] = sbox7(s[0]);
        }
        mds(s);
    }
}

/* Poseidon::poseidon (naive schedule: 4 full, 22 partial, 4 full; x^7).  Reached from
 * PoseidonHash::two_to_one (src/smt/goldilocks_poseidon/mod.rs:165, src/zkdsa/account.rs:165,
 * src/zkdsa/circuits/mod.rs:66-67) and PoseidonHash::hash_pad (src/smt/goldilocks_poseidon/mod.rs:170). */
void glo_poseidon_permute_naive(u64 s[12]) { permute_with(s, mds_layer_naive); }

/* The pre-round-2 schedule (literal rounds, mds_row_shf-style MDS): kept as a second cross-check. */
void glo_poseidon_permute_slow(u64 s[12]) {
    ensure_rc();
    const u64 *rc = RC;
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
    for (int r = 0; r < 30; r++, rc += 12) {
        for (int i = 0; i < 12; i++) {
            u64 t = s[i] + rc[i]; /* s canonical, rc canonical: at most one wrap of 2^64 */
            if (t < rc[i]) t += EPS;
            s[i] = t;
        }
        if (r < 4 || r >= 26) {
            for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        } else {
            s[0] = sbox7(s[0]);
        }
        mds_layer(s);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Upstream's CPU schedule of the SAME permutation (plonky2::hash::poseidon::Poseidon::poseidon:
 * full_rounds, then partial_rounds = partial_first_constant_layer + mds_partial_layer_init + 22 x
 * (sbox_monomial on lane 0, one lane-0 constant, mds_partial_layer_fast), then full_rounds).
 * Upstream ships the "fast partial round" tables as literals; that source is not on this disk, so the
 * tables are DERIVED here from the MDS matrix and the round constants (exact arithmetic mod p) and
 * the result is cross-checked against the literal schedule on random states (tests/test_oracle_cpu.py)
 * and against the reference's KAT.  Only outputs are comparable with upstream, not the tables.
 *
 * Derivation (column vectors; M = [[m00, v], [w, Mh]], S = x^7 on lane 0 only, R = 22):
 *   y_{r+1} = M S(y_r) + c_{r+1}.  With A_r = diag(1, Mh^-(R-r)) and u_r = A_r^-1 y_r, A_r commutes
 *   with S and u_{r+1} = Sp_r S(u_r) + A_{r+1}^-1 c_{r+1},  Sp_r = [[m00, v Mh^-(R-r)], [Mh^(R-r-1) w, I]]
 *   (A_R = I, so u_R is the true state).  A constant K = (K0, Kh) pending after Sp_r equals
 *   Sp_r (0, Kh) + (K0 - v_r . Kh) e0, and (0, Kh) passes backwards through S: every round keeps ONE
 *   lane-0 constant, the rest accumulates into the constants added before the initial dense matrix.
 * Cost per partial round: 1 S-box, a 12-term dot product (one 160-bit accumulation, one reduction)
 * and 11 multiply-accumulates -- upstream's mds_partial_layer_fast has exactly this shape.
 * ---------------------------------------------------------------------------------------------- */
#define NP 22
static u64 FP_FIRST[12];      /* added to the state that enters the partial rounds */
static u64 FP_INIT[11][11];   /* Mh^R: mds_partial_layer_init on lanes 1..11 */
static u64 FP_POST[11];       /* constants on lanes 1..11 after the initial matrix */
static u64 FP_ALPHA[NP];      /* lane-0 constant after the sparse matrix of round r */
static u64 FP_V[NP][11], FP_W[NP][11];
static int fp_ready = 0;

static void mat11_mul(const u64 a[11][11], const u64 b[11][11], u64 o[11][11]) {
    u64 t[11][11];
    for (int i = 0; i < 11; i++)
        for (int j = 0; j < 11; j++) {
            u64 acc = 0;
            for (int k = 0; k < 11; k++) acc = f_add(acc, f_mul(a[i][k], b[k][j]));
            t[i][j] = acc;
        }
    memcpy(o, t, sizeof t);
}
static void mat11_inv(const u64 a[11][11], u64 inv[11][11]) {
    u64 m[11][22];
    for (int i = 0; i < 11; i++)
        for (int j = 0; j < 11; j++) { m[i][j] = canon(a[i][j]); m[i][11 + j] = (i == j); }
    for (int c = 0; c < 11; c++) {
        int piv = c;
        while (piv < 11 && m[piv][c] == 0) piv++;
        if (piv == 11) abort(); /* the MDS minor is invertible */
        if (piv != c) for (int j = 0; j < 22; j++) { u64 t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; }
        u64 iv = glo_inv(m[c][c]);
        for (int j = 0; j < 22; j++) m[c][j] = f_mul(m[c][j], iv);
        for (int i = 0; i < 11; i++) {
            if (i == c || m[i][c] == 0) continue;
            u64 f = m[i][c];
            for (int j = 0; j < 22; j++) m[i][j] = f_sub(m[i][j], f_mul(f, m[c][j]));
        }
    }
    for (int i = 0; i < 11; i++) for (int j = 0; j < 11; j++) inv[i][j] = m[i][11 + j];
}

static void derive_fast_partial(void) {
    u64 Mh[11][11], Mhi[11][11], v[11], w[11];
    for (int i = 0; i < 11; i++) {
        v[i] = MDS_CIRC[(i + 1) % 12];            /* M[0][j], j = i + 1 */
        w[i] = MDS_CIRC[(12 - (i + 1)) % 12];     /* M[i + 1][0] */
        for (int j = 0; j < 11; j++) Mh[i][j] = MDS_CIRC[((j - i) % 12 + 12) % 12]; /* M[i+1][j+1]; the diagonal term is only at [0][0] */
    }
    mat11_inv(Mh, Mhi);
    /* Ainv[r] = Mh^(R-r) (r = 0..R), A[r] = Mh^-(R-r) */
    static u64 Apow[NP + 1][11][11], Aneg[NP + 1][11][11];
    for (int i = 0; i < 11; i++) for (int j = 0; j < 11; j++) Apow[NP][i][j] = Aneg[NP][i][j] = (i == j);
    for (int r = NP - 1; r >= 0; r--) {
        mat11_mul(Apow[r + 1], Mh, Apow[r]);
        mat11_mul(Aneg[r + 1], Mhi, Aneg[r]);
    }
    for (int r = 0; r < NP; r++) {
        for (int j = 0; j < 11; j++) {
            u64 a = 0, b = 0;
            for (int k = 0; k < 11; k++) {
                a = f_add(a, f_mul(v[k], Aneg[r][k][j]));        /* v_r = v Mh^-(R-r) */
                b = f_add(b, f_mul(Apow[r + 1][j][k], w[k]));    /* w_r = Mh^(R-r-1) w */
            }
            FP_V[r][j] = a;
            FP_W[r][j] = b;
        }
    }
    /* constants, from the last partial round backwards */
    u64 pend[11] = {0};
    FP_ALPHA[NP - 1] = 0;
    for (int j = NP - 1; j >= 1; j--) {
        const u64 *c = RC + 12 * (4 + j);
        u64 K0 = c[0], Kh[11];
        for (int i = 0; i < 11; i++) {
            u64 acc = pend[i];
            for (int k = 0; k < 11; k++) acc = f_add(acc, f_mul(Apow[j][i][k], c[1 + k]));
            Kh[i] = acc;
        }
        u64 dot = 0;
        for (int i = 0; i < 11; i++) dot = f_add(dot, f_mul(FP_V[j - 1][i], Kh[i]));
        FP_ALPHA[j - 1] = f_sub(K0, dot);
        memcpy(pend, Kh, sizeof pend);
    }
    memcpy(FP_POST, pend, sizeof pend);
    memcpy(FP_FIRST, RC + 12 * 4, sizeof FP_FIRST);
    memcpy(FP_INIT, Apow[0], sizeof FP_INIT);
}

static void ensure_fp(void) {
    ensure_rc();
    if (!fp_ready) {
#pragma omp critical(glo_fp)
        {
            if (!fp_ready) {
                derive_fast_partial();
                fp_ready = 1;
            }
        }
    }
}
void glo_poseidon_fast_tables(u64 *first12, u64 *init121, u64 *post11, u64 *alpha22, u64 *v242, u64 *w242) {
    ensure_fp();
    memcpy(first12, FP_FIRST, sizeof FP_FIRST);
    memcpy(init121, FP_INIT, sizeof FP_INIT);
    memcpy(post11, FP_POST, sizeof FP_POST);
    memcpy(alpha22, FP_ALPHA, sizeof FP_ALPHA);
    memcpy(v242, FP_V, sizeof FP_V);
    memcpy(w242, FP_W, sizeof FP_W);
}

/* x in [0, 2^64) -> x mod-p representative in [0, 2^64) (not canonical), from a 128-bit value */
static inline u64 red128_nc(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hi_hi = hi >> 32, hi_lo = hi & EPS;
    u64 t0 = lo - hi_hi;
    t0 -= EPS & (0 - (u64)(lo < hi_hi));
    u64 t1 = hi_lo * EPS;
    u64 r = t0 + t1;
    r += EPS & (0 - (u64)(r < t1));
    return r;
}
static inline u64 mul_nc(u64 a, u64 b) { return red128_nc((u128)a * b); }
static inline u64 sbox7_nc(u64 x) {
    u64 x2 = mul_nc(x, x), x4 = mul_nc(x2, x2), x3 = mul_nc(x, x2);
    return mul_nc(x3, x4);
}
static inline u64 add_nc(u64 a, u64 b_canonical) { /* a any u64, b < p: at most one wrap */
    u64 t = a + b_canonical;
    return t + (EPS & (0 - (u64)(t < b_canonical)));
}
/* One full round.  The S-box runs stage by stage over the twelve lanes (twelve independent multiply chains in flight) and the
 * MDS layer works on 32-bit halves with the shift as the outer loop, which is how upstream's mds_row_shf schedule reads
 * and what lets the compiler keep twelve accumulators in vector registers (32x32->64 products). */
static const u32 MDS_CIRC32[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static inline void full_round_nc(u64 s[12], const u64 *rc) {
    u64 x[12], x2[12], x3[12], x4[12];
    for (int i = 0; i < 12; i++) x[i] = add_nc(s[i], rc[i]);
    for (int i = 0; i < 12; i++) x2[i] = mul_nc(x[i], x[i]);
    for (int i = 0; i < 12; i++) x4[i] = mul_nc(x2[i], x2[i]);
    for (int i = 0; i < 12; i++) x3[i] = mul_nc(x2[i], x[i]);
    for (int i = 0; i < 12; i++) s[i] = mul_nc(x3[i], x4[i]);
    u32 lo[24] __attribute__((aligned(32))), hi[24] __attribute__((aligned(32)));
    for (int i = 0; i < 12; i++) {
        lo[i] = lo[i + 12] = (u32)s[i];
        hi[i] = hi[i + 12] = (u32)(s[i] >> 32);
    }
    u64 al[12] __attribute__((aligned(32))) = {0}, ah[12] __attribute__((aligned(32))) = {0};
    for (int i = 0; i < 12; i++) {
        const u64 c = MDS_CIRC32[i];
        for (int r = 0; r < 12; r++) {
            al[r] += (u64)lo[r + i] * c;
            ah[r] += (u64)hi[r + i] * c;
        }
    }
    al[0] += (u64)lo[0] * MDS_DIAG0;
    ah[0] += (u64)hi[0] * MDS_DIAG0;
    for (int r = 0; r < 12; r++) s[r] = red128_nc((u128)al[r] + ((u128)ah[r] << 32));
}
/* sum of up to 2^4 128-bit products kept as two running sums of their 64-bit halves (no overflow tests): lo + 2^64 hi with
 * both below 2^68; 2^64 = 2^32 - 1 (mod p) */
static inline u64 red_split(u128 lo_sum, u128 hi_sum) {
    const u64 rh = red128_nc(hi_sum);
    return red128_nc((u128)rh * EPS + lo_sum); /* < 2^96 + 2^68 */
}

/* The permutation every other oracle routine uses: upstream's schedule, canonical output. */
void glo_poseidon_permute(u64 s[12]) {
    ensure_fp();
    const u64 *rc = RC;
    for (int r = 0; r < 4; r++, rc += 12) full_round_nc(s, rc);
    /* partial_first_constant_layer + mds_partial_layer_init */
    u64 t[12];
    for (int i = 0; i < 12; i++) t[i] = add_nc(s[i], FP_FIRST[i]);
    s[0] = t[0];
    for (int i = 0; i < 11; i++) {
        u128 lo_sum = 0, hi_sum = 0;
        for (int k = 0; k < 11; k++) {
            const u128 pr = (u128)t[1 + k] * FP_INIT[i][k];
            lo_sum += (u64)pr;
            hi_sum += (u64)(pr >> 64);
        }
        s[1 + i] = add_nc(red_split(lo_sum, hi_sum), FP_POST[i]);
    }
    for (int r = 0; r < NP; r++) {
        /* mds_partial_layer_fast: the part of the first row that does not wait for the S-box goes first */
        u128 lo_sum = 0, hi_sum = 0;
        for (int k = 0; k < 11; k++) {
            const u128 pr = (u128)s[1 + k] * FP_V[r][k];
            lo_sum += (u64)pr;
            hi_sum += (u64)(pr >> 64);
        }
        const u64 s0 = sbox7_nc(s[0]);
        const u128 p0 = (u128)s0 * (MDS_CIRC[0] + MDS_DIAG0);
        lo_sum += (u64)p0;
        hi_sum += (u64)(p0 >> 64);
        const u64 d = red_split(lo_sum, hi_sum);
        for (int k = 0; k < 11; k++) {
            /* multiply_accumulate: s[k] + s0 * w, one 128-bit sum (s0 * w <= (2^64-1)^2, + s[k] cannot overflow) */
            s[1 + k] = red128_nc((u128)s0 * FP_W[r][k] + s[1 + k]);
        }
        s[0] = add_nc(d, FP_ALPHA[r]);
    }
    rc = RC + 12 * 26;
    for (int r = 0; r < 4; r++, rc += 12) full_round_nc(s, rc);
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
}

/* hashing::hash_n_to_m_no_pad with overwrite-mode absorption, RATE = 8, 4 outputs */
void glo_hash_no_pad(const u64 *in, size_t len, u64 out[4]) {
    u64 s[12] = {0};
    for (size_t off = 0; off < len; off += 8) {
        size_t k = len - off < 8 ? len - off : 8;
        for (size_t i = 0; i < k; i++) s[i] = canon(in[off + i]);
        glo_poseidon_permute(s);
    }
    memcpy(out, s, 4 * sizeof(u64));
}

/* Hasher::hash_pad: push 1, zeros until (len+1) % WIDTH == 0, push 1.  WIDTH = 12 is forced by
 * src/smt/gadgets/common.rs:87-101 == src/smt/goldilocks_poseidon/mod.rs:167-181. */
void glo_hash_pad(const u64 *in, size_t len, u64 out[4]) {
    size_t n = len + 1;
    while ((n + 1) % 12 != 0) n++;
    n++;
    u64 *buf = (u64 *)calloc(n, sizeof(u64));
    memcpy(buf, in, len * sizeof(u64));
    buf[len] = 1;
    buf[n - 1] = 1;
    glo_hash_no_pad(buf, n, out);
    free(buf);
}

/* Hasher::hash_or_noop: <= 4 elements are copied (zero padded), longer inputs hashed */
void glo_hash_or_noop(const u64 *in, size_t len, u64 out[4]) {
    if (len <= 4) {
        for (size_t i = 0; i < 4; i++) out[i] = i < len ? canon(in[i]) : 0;
    } else {
        glo_hash_no_pad(in, len, out);
    }
}

/* hashing::compress = PoseidonHash::two_to_one */
void glo_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
    u64 s[12] = {0};
    for (int i = 0; i < 4; i++) { s[i] = l[i]; s[4 + i] = r[i]; }
    glo_poseidon_permute(s);
    memcpy(out, s, 4 * sizeof(u64));
}

void glo_permute_batch(u64 *states, size_t m) {
    ensure_rc();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < m; i++) glo_poseidon_permute(states + 12 * i);
}
void glo_two_to_one_batch(const u64 *l, const u64 *r, u64 *out, size_t m) {
    ensure_rc();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < m; i++) glo_two_to_one(l + 4 * i, r + 4 * i, out + 4 * i);
}
void glo_hash_no_pad_batch(const u64 *in, size_t len_each, size_t m, u64 *out) {
    ensure_rc();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < m; i++) glo_hash_no_pad(in + len_each * i, len_each, out + 4 * i);
}

/* ------------------------------------------------------------------------------------------------
 * P1/P2/P9  plonky2_field::fft  (fft_classic: bit-reverse, then decimation-in-time butterflies
 *           with a table of roots; natural order in and out).
 * ---------------------------------------------------------------------------------------------- */
size_t glo_reverse_bits(size_t x, unsigned bits) {
    if (bits == 0) return 0;
    uint64_t v = (uint64_t)x;
    v = ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
    v = ((v >> 2) & 0x3333333333333333ULL) | ((v & 0x3333333333333333ULL) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((v & 0x0F0F0F0F0F0F0F0FULL) << 4);
    v = __builtin_bswap64(v);
    return (size_t)(v >> (64 - bits));
}

static void reverse_index_bits_u64(u64 *a, unsigned lg_n) {
    size_t n = (size_t)1 << lg_n;
    for (size_t i = 0; i < n; i++) {
        size_t j = glo_reverse_bits(i, lg_n);
        if (i < j) { u64 t = a[i]; a[i] = a[j]; a[j] = t; }
    }
}

/* roots[j] = w_n^j for j < n/2; level m uses stride n/m (same numbers as FftRootTable row lg_m-1) */
static u64 *root_cache[33];
static const u64 *roots_for(unsigned lg_n) {
    if (lg_n == 0) return NULL;
    if (!root_cache[lg_n]) {
#pragma omp critical(glo_roots)
        {
            if (!root_cache[lg_n]) {
                size_t half = (size_t)1 << (lg_n - 1);
                u64 *t = (u64 *)malloc(half * sizeof(u64));
                u64 w = glo_primitive_root_of_unity(lg_n), cur = 1;
                for (size_t j = 0; j < half; j++) { t[j] = cur; cur = f_mul(cur, w); }
                root_cache[lg_n] = t;
            }
        }
    }
    return root_cache[lg_n];
}

/* fft_with_options(poly, zero_factor, root_table): out[j] = sum_i a[i] * w_n^(i*j).
 * zero_factor = r promises the last n - n/2^r inputs are zero (PolynomialCoeffs::lde); the first r
 * butterfly layers then degenerate to copies, which is what upstream does; results are identical. */
void glo_fft(u64 *a, unsigned lg_n, unsigned zero_factor) {
    size_t n = (size_t)1 << lg_n;
    if (lg_n == 0) { a[0] = canon(a[0]); return; }
    const u64 *roots = roots_for(lg_n);
    for (size_t i = 0; i < n; i++) a[i] = canon(a[i]);
    reverse_index_bits_u64(a, lg_n);
    unsigned r = zero_factor > lg_n ? lg_n : zero_factor;
    if (r > 0) {
        size_t mask = ~(((size_t)1 << r) - 1);
        for (size_t i = 0; i < n; i++) a[i] = a[i & mask];
    }
    for (unsigned lg_half = r; lg_half < lg_n; lg_half++) {
        size_t half = (size_t)1 << lg_half, m = half << 1, stride = n / m;
        for (size_t k = 0; k < n; k += m) {
            for (size_t j = 0; j < half; j++) {
                u64 t = f_mul(roots[j * stride], a[k + half + j]);
                u64 u = a[k + j];
                a[k + j] = f_add(u, t);
                a[k + half + j] = f_sub(u, t);
            }
        }
    }
}

/* ifft_with_options: forward FFT, then out[0],out[n/2] *= 1/n and out[i] <-> out[n-i] scaled by 1/n */
void glo_ifft(u64 *a, unsigned lg_n) {
    size_t n = (size_t)1 << lg_n;
    glo_fft(a, lg_n, 0);
    u64 n_inv = glo_inv((u64)n % P);
    a[0] = f_mul(a[0], n_inv);
    if (n > 1) a[n / 2] = f_mul(a[n / 2], n_inv);
    for (size_t i = 1; i < n / 2; i++) {
        size_t j = n - i;
        u64 ci = f_mul(a[j], n_inv), cj = f_mul(a[i], n_inv);
        a[i] = ci;
        a[j] = cj;
    }
}

/* PolynomialCoeffs::coset_fft_with_options: coeffs[i] *= shift^i, then fft */
void glo_coset_fft(u64 *a, unsigned lg_n, u64 shift, unsigned zero_factor) {
    size_t n = (size_t)1 << lg_n;
    u64 cur = 1;
    for (size_t i = 0; i < n; i++) { a[i] = f_mul(a[i], cur); cur = f_mul(cur, shift); }
    glo_fft(a, lg_n, zero_factor);
}

/* PolynomialValues::coset_ifft: ifft, then coeffs[i] *= shift^-i */
void glo_coset_ifft(u64 *a, unsigned lg_n, u64 shift) {
    size_t n = (size_t)1 << lg_n;
    glo_ifft(a, lg_n);
    u64 sinv = glo_inv(shift), cur = 1;
    for (size_t i = 0; i < n; i++) { a[i] = f_mul(a[i], cur); cur = f_mul(cur, sinv); }
}

/* ------------------------------------------------------------------------------------------------
 * P4/P11  plonky2::hash::merkle_tree::MerkleTree::new / prove, merkle_proofs::verify_merkle_proof_to_cap
 * ---------------------------------------------------------------------------------------------- */
static unsigned log2_strict(size_t n) {
    unsigned k = 0;
    while (((size_t)1 << k) < n) k++;
    return k;
}

/* fill_subtree: digests_buf has 2*(num_leaves-1) entries, recursive in-order layout
 * [left subtree digests | left child | right child | right subtree digests]; returns the root. */
static void fill_subtree(u64 *buf, size_t buf_len, const u64 *leaves, size_t num_leaves, size_t leaf_len,
                         u64 root[4]) {
    if (buf_len == 0) {
        glo_hash_or_noop(leaves, leaf_len, root);
        return;
    }
    size_t half_buf = buf_len / 2, half_leaves = num_leaves / 2;
    u64 *left_slot = buf + 4 * (half_buf - 1);
    u64 *right_slot = buf + 4 * half_buf;
    u64 ld[4], rd[4];
    int par = num_leaves >= 1024; /* rayon::join upstream */
#pragma omp task shared(ld) if (par)
    fill_subtree(buf, half_buf - 1, leaves, half_leaves, leaf_len, ld);
#pragma omp task shared(rd) if (par)
    fill_subtree(buf + 4 * (half_buf + 1), half_buf - 1, leaves + half_leaves * leaf_len, half_leaves,
                 leaf_len, rd);
#pragma omp taskwait
    memcpy(left_slot, ld, sizeof ld);
    memcpy(right_slot, rd, sizeof rd);
    glo_two_to_one(ld, rd, root);
}

int glo_merkle_tree(const u64 *leaves, size_t num_leaves, size_t leaf_len, unsigned cap_height,
                    u64 *digests, u64 *cap) {
    if (num_leaves == 0 || (num_leaves & (num_leaves - 1))) return 1; /* log2_strict panics */
    unsigned lg = log2_strict(num_leaves);
    if (cap_height > lg) return 2; /* assert!(cap_height <= log2_leaves_len) */
    ensure_rc();
    size_t num_caps = (size_t)1 << cap_height;
    size_t num_digests = 2 * (num_leaves - num_caps);
    size_t sub_leaves = num_leaves / num_caps, sub_digests = num_digests / num_caps;
#pragma omp parallel
#pragma omp single
    {
        for (size_t s = 0; s < num_caps; s++) {
#pragma omp task firstprivate(s)
            fill_subtree(digests ? digests + 4 * s * sub_digests : NULL, sub_digests,
                         leaves + s * sub_leaves * leaf_len, sub_leaves, leaf_len, cap + 4 * s);
        }
#pragma omp taskwait
    }
    return 0;
}

void glo_merkle_prove(const u64 *digests, size_t num_leaves, unsigned cap_height, size_t leaf_index,
                      u64 *siblings) {
    unsigned lg = log2_strict(num_leaves);
    unsigned L = lg - cap_height;
    size_t num_caps = (size_t)1 << cap_height;
    size_t sub_digests = 2 * (num_leaves - num_caps) / num_caps;
    size_t subtree = leaf_index >> L;
    const u64 *buf = digests + 4 * subtree * sub_digests;
    size_t pair = leaf_index & (((size_t)1 << L) - 1);
    for (unsigned i = 0; i < L; i++) {
        size_t parity = pair & 1;
        pair >>= 1;
        size_t sib = 2 * ((pair << (i + 1)) + ((size_t)1 << i) - 1) + (1 - parity);
        memcpy(siblings + 4 * i, buf + 4 * sib, 4 * sizeof(u64));
    }
}

int glo_merkle_verify(const u64 *leaf, size_t leaf_len, size_t leaf_index, const u64 *siblings,
                      unsigned num_siblings, const u64 *cap, unsigned cap_height) {
    u64 d[4];
    (void)cap_height;
    glo_hash_or_noop(leaf, leaf_len, d);
    size_t idx = leaf_index;
    for (unsigned i = 0; i < num_siblings; i++) {
        u64 nd[4];
        if (idx & 1) glo_two_to_one(siblings + 4 * i, d, nd);
        else glo_two_to_one(d, siblings + 4 * i, nd);
        memcpy(d, nd, sizeof d);
        idx >>= 1;
    }
    return memcmp(d, cap + 4 * idx, sizeof d) == 0;
}

/* ------------------------------------------------------------------------------------------------
 * P*  plonky2::fri::oracle::PolynomialBatch::from_values / from_coeffs
 *     reached from data.prove(pw) (44 sites, e.g. src/ecdsa/gadgets/ecdsa.rs:349) and
 *     builder.build::<C>() (41 sites, e.g. src/ecdsa/gadgets/ecdsa.rs:298)
 * ---------------------------------------------------------------------------------------------- */
int glo_commit_from_coeffs(const u64 *coeffs, unsigned lg_n, unsigned c, unsigned rate_bits,
                           unsigned cap_height, u64 *leaves_out, u64 *digests_out, u64 *cap_out) {
    size_t n = (size_t)1 << lg_n, N = n << rate_bits;
    unsigned lg_N = lg_n + rate_bits;
    if (cap_height > lg_N) return 2;
    roots_for(lg_N);
    /* "FFT + blinding": lde_values[col] = coeffs[col].lde(rate_bits).coset_fft(7); blinding = false */
    double t_0 = glo_now();
    u64 *lde = (u64 *)malloc((size_t)c * N * sizeof(u64));
    if (!lde) return 3;
#pragma omp parallel for schedule(dynamic)
    for (unsigned col = 0; col < c; col++) {
        u64 *v = lde + (size_t)col * N;
        memcpy(v, coeffs + (size_t)col * n, n * sizeof(u64));
        memset(v + n, 0, (N - n) * sizeof(u64));
        glo_coset_fft(v, lg_N, 7, rate_bits);
    }
    /* "transpose LDEs" + reverse_index_bits_in_place(leaves) */
    double t_1 = glo_now();
    u64 *leaves = leaves_out ? leaves_out : (u64 *)malloc((size_t)c * N * sizeof(u64));
    if (!leaves) { free(lde); return 3; }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) {
        size_t src = glo_reverse_bits(i, lg_N);
        for (unsigned col = 0; col < c; col++) leaves[i * c + col] = lde[(size_t)col * N + src];
    }
    free(lde);
    double t_2 = glo_now();
    /* "build Merkle tree" */
    size_t num_digests = 2 * (N - ((size_t)1 << cap_height));
    u64 *digests = digests_out ? digests_out : (u64 *)malloc((num_digests ? num_digests : 1) * 4 * sizeof(u64));
    u64 cap_local[4 * 64];
    u64 *cap = cap_out ? cap_out : (cap_height <= 6 ? cap_local : (u64 *)malloc(((size_t)4 << cap_height) * sizeof(u64)));
    int rc = glo_merkle_tree(leaves, N, c, cap_height, digests, cap);
    if (getenv("GLO_TIMING"))
        fprintf(stderr, "glo_commit_from_coeffs 2^%u x %u: FFT %.2f s, transpose %.2f s, Merkle %.2f s\n", lg_n, c, t_1 - t_0, t_2 - t_1,
                glo_now() - t_2);
    if (!leaves_out) free(leaves);
    if (!digests_out) free(digests);
    if (!cap_out && cap_height > 6) free(cap);
    return rc;
}

int glo_commit_from_values(const u64 *values, unsigned lg_n, unsigned c, unsigned rate_bits,
                           unsigned cap_height, u64 *coeffs_out, u64 *leaves_out, u64 *digests_out,
                           u64 *cap_out) {
    size_t n = (size_t)1 << lg_n;
    u64 *coeffs = coeffs_out ? coeffs_out : (u64 *)malloc((size_t)c * n * sizeof(u64));
    if (!coeffs) return 3;
    roots_for(lg_n);
    /* "IFFT": values.into_par_iter().map(|v| v.ifft()) */
#pragma omp parallel for schedule(dynamic)
    for (unsigned col = 0; col < c; col++) {
        u64 *v = coeffs + (size_t)col * n;
        if (v != values + (size_t)col * n) memcpy(v, values + (size_t)col * n, n * sizeof(u64));
        glo_ifft(v, lg_n);
    }
    int rc = glo_commit_from_coeffs(coeffs, lg_n, c, rate_bits, cap_height, leaves_out, digests_out, cap_out);
    if (!coeffs_out) free(coeffs);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * P6  PoseidonNodeHash::calc_node_hash  (src/smt/goldilocks_poseidon/mod.rs:158-184)
 * ---------------------------------------------------------------------------------------------- */
void glo_smt_leaf_hash(const u64 key[4], const u64 value[4], u64 out[4]) {
    u64 in[9];
    memcpy(in, key, 32);
    memcpy(in + 4, value, 32);
    in[8] = 1;
    glo_hash_pad(in, 9, out); /* = hash_no_pad([k, v, 1, 1, 0, 1]), src/smt/gadgets/common.rs:87-101 */
}
void glo_smt_internal_hash(const u64 l[4], const u64 r[4], u64 out[4]) { glo_two_to_one(l, r, out); }

/* ------------------------------------------------------------------------------------------------
 * P7  verify_smt_process_proof  (src/smt/proof/process.rs:153-257), calc_old_new_root (:260-337),
 *     smt_processor_sm (:340-370), smt_lev_ins (src/smt/proof/common.rs:8-44).
 *     Key bits: HashOut::to_bytes (4 x u64 LE) -> LSB-first bits
 *     (src/smt/goldilocks_poseidon/mod.rs:27-48, src/smt/proof/common.rs:47-58).
 *     Returns 0 when no assert fires, else the ordinal of the first assert that would panic.
 * ---------------------------------------------------------------------------------------------- */
enum { ST_TOP = 0, ST_BOT = 1, ST_OLD0 = 2, ST_NEW1 = 3, ST_UPD = 4, ST_NA = 5 };

static inline int key_bit(const u64 k[4], unsigned i) { return (int)((k[i >> 6] >> (i & 63)) & 1); }
static inline int is_zero4(const u64 h[4]) { return (h[0] | h[1] | h[2] | h[3]) == 0; }
static inline int eq4(const u64 a[4], const u64 b[4]) { return memcmp(a, b, 32) == 0; }

static int processor_sm(int prev, int diff_bit, int is_old0, int lev_ins, int ins_or_rem) {
    if (prev == ST_TOP) {
        if (!lev_ins) return ST_TOP;
        if (!ins_or_rem) return ST_UPD;
        if (is_old0) return ST_OLD0;
        return diff_bit ? ST_NEW1 : ST_BOT;
    }
    if (prev == ST_BOT) return diff_bit ? ST_NEW1 : ST_BOT;
    return ST_NA;
}

int glo_smt_verify_process_proof(const glo_smt_process_proof *pf) {
    enum { LEVELS = 256 };
    int enabled = pf->fnc != 0;
    u32 fnc = pf->fnc;
    /* Every word is a field element: the reference takes key bits from HashOut::to_bytes (canonical u64s,
     * src/smt/proof/process.rs:193-203) and compares hashes as field elements -> reduce mod p on entry. */
    u64 hk[6][4];
    for (int j = 0; j < 4; j++) {
        hk[0][j] = canon(pf->old_key[j]); hk[1][j] = canon(pf->old_value[j]); hk[2][j] = canon(pf->old_root[j]);
        hk[3][j] = canon(pf->new_key[j]); hk[4][j] = canon(pf->new_value[j]); hk[5][j] = canon(pf->new_root[j]);
    }
    const u64 *old_key = hk[0], *old_value = hk[1], *old_root = hk[2];
    const u64 *new_key = hk[3], *new_value = hk[4], *new_root = hk[5];
    if (fnc == 3) { /* a remove proof is an insert proof with old and new flipped */
        fnc = 2;
        old_key = hk[3]; old_value = hk[4]; old_root = hk[5];
        new_key = hk[0]; new_value = hk[1]; new_root = hk[2];
    }
    if (pf->num_siblings >= LEVELS) return 1; /* assert!(siblings.len() < n2b_new.len()) */
    /* siblings.resize(256, default) : the struct is already zero padded */
    u64 sib[LEVELS][4];
    for (unsigned i = 0; i < LEVELS; i++) {
        for (int j = 0; j < 4; j++) sib[i][j] = i < pf->num_siblings ? canon(pf->siblings[i][j]) : 0;
    }
    /* smt_lev_ins */
    if (enabled && !is_zero4(sib[LEVELS - 1])) return 2; /* assert!(is_zeros.last()) */
    unsigned char lev_ins[LEVELS];
    {
        /* is_zeros reversed with a trailing false; scan from the bottom of the path upwards */
        int last_done = 0;
        for (unsigned i = 0; i < LEVELS; i++) {
            /* reversed index i+1 is original index LEVELS-2-i; beyond the top -> pushed `false` */
            int zero_next = (i + 1 < LEVELS) ? is_zero4(sib[LEVELS - 2 - i]) : 0;
            int v = !zero_next && !last_done;
            last_done = last_done || !zero_next;
            lev_ins[LEVELS - 1 - i] = (unsigned char)v;
        }
    }
    int sm[LEVELS];
    int prev = enabled ? ST_TOP : ST_NA;
    int ins_or_rem = fnc == 2;
    for (unsigned i = 0; i < LEVELS; i++) {
        int st = processor_sm(prev, key_bit(old_key, i) ^ key_bit(new_key, i), (int)pf->is_old0, lev_ins[i],
                              ins_or_rem);
        sm[i] = st;
        prev = st;
    }
    if (sm[LEVELS - 1] == ST_TOP || sm[LEVELS - 1] == ST_BOT) return 3;
    /* calc_old_new_root */
    u64 zero[4] = {0, 0, 0, 0};
    u64 old1_leaf[4], new1_leaf[4];
    glo_smt_leaf_hash(old_key, old_value, old1_leaf);
    glo_smt_leaf_hash(new_key, new_value, new1_leaf);
    u64 prev_old[4] = {0, 0, 0, 0}, prev_new[4] = {0, 0, 0, 0};
    for (int i = LEVELS - 1; i >= 0; i--) {
        int pos = key_bit(new_key, (unsigned)i);
        u64 old_hash[4], new_hash[4];
        if (pos) glo_two_to_one(sib[i], prev_old, old_hash);
        else glo_two_to_one(prev_old, sib[i], old_hash);
        const u64 *o_root, *n_left, *n_right, *n_root;
        switch (sm[i]) {
            case ST_TOP:  o_root = old_hash;  n_left = prev_new;  n_right = sib[i];    break;
            case ST_BOT:  o_root = old1_leaf; n_left = prev_new;  n_right = zero;      break;
            case ST_NEW1: o_root = old1_leaf; n_left = new1_leaf; n_right = old1_leaf; break;
            case ST_UPD:  o_root = old1_leaf; n_left = zero;      n_right = zero;      break;
            default:      o_root = zero;      n_left = zero;      n_right = zero;      break;
        }
        if (pos) glo_two_to_one(n_right, n_left, new_hash);
        else glo_two_to_one(n_left, n_right, new_hash);
        switch (sm[i]) {
            case ST_TOP: case ST_BOT: case ST_NEW1: n_root = new_hash; break;
            case ST_OLD0: case ST_UPD: n_root = new1_leaf; break;
            default: n_root = zero; break;
        }
        u64 t_old[4], t_new[4];
        memcpy(t_old, o_root, 32);
        memcpy(t_new, n_root, 32);
        memcpy(prev_old, t_old, 32);
        memcpy(prev_new, t_new, 32);
    }
    if (enabled) {
        if (!eq4(prev_old, old_root)) return 4;
        if (!eq4(prev_new, new_root)) return 5;
    } else {
        if (!eq4(old_root, new_root)) return 6;
        if (!eq4(old_value, new_value)) return 7;
    }
    if (fnc == 1 || !enabled) {
        if (!eq4(old_key, new_key)) return 8;
    }
    return 0;
}

void glo_smt_verify_process_batch(const glo_smt_process_proof *proofs, size_t m, int32_t *status) {
    ensure_rc();
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < m; i++) status[i] = glo_smt_verify_process_proof(proofs + i);
}

/* ------------------------------------------------------------------------------------------------
 * In-memory sparse Merkle tree: src/smt/tree.rs (find :588-676, update :174-253, insert :255-387,
 * remove :390-533, noop :536-559, calc_process_proof :561-586) over NodeDataMemory
 * (src/smt/goldilocks_poseidon/mod.rs:58-94: nodes are never deleted).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { u64 hash[4]; u64 a[4]; u64 b[4]; unsigned char used, is_leaf; } smt_node;
struct glo_smt { smt_node *tab; size_t cap, count; u64 root[4]; };

static size_t h4(const u64 h[4]) { return (size_t)(h[0] ^ (h[1] * 0x9E3779B97F4A7C15ULL) ^ (h[2] << 7) ^ (h[3] >> 3)); }

static smt_node *smt_lookup(const glo_smt *t, const u64 h[4]) {
    size_t i = h4(h) & (t->cap - 1);
    while (t->tab[i].used) {
        if (eq4(t->tab[i].hash, h)) return &t->tab[i];
        i = (i + 1) & (t->cap - 1);
    }
    return NULL;
}
static void smt_put(glo_smt *t, const u64 h[4], int is_leaf, const u64 a[4], const u64 b[4]);
static void smt_grow(glo_smt *t) {
    smt_node *old = t->tab;
    size_t oc = t->cap;
    t->cap *= 2;
    t->tab = (smt_node *)calloc(t->cap, sizeof(smt_node));
    t->count = 0;
    for (size_t i = 0; i < oc; i++)
        if (old[i].used) smt_put(t, old[i].hash, old[i].is_leaf, old[i].a, old[i].b);
    free(old);
}
static void smt_put(glo_smt *t, const u64 h[4], int is_leaf, const u64 a[4], const u64 b[4]) {
    if ((t->count + 1) * 2 > t->cap) smt_grow(t);
    size_t i = h4(h) & (t->cap - 1);
    while (t->tab[i].used && !eq4(t->tab[i].hash, h)) i = (i + 1) & (t->cap - 1);
    if (!t->tab[i].used) t->count++;
    t->tab[i].used = 1;
    t->tab[i].is_leaf = (unsigned char)is_leaf;
    memcpy(t->tab[i].hash, h, 32);
    memcpy(t->tab[i].a, a, 32);
    memcpy(t->tab[i].b, b, 32);
}

glo_smt *glo_smt_new(void) {
    glo_smt *t = (glo_smt *)calloc(1, sizeof(glo_smt));
    t->cap = 1024;
    t->tab = (smt_node *)calloc(t->cap, sizeof(smt_node));
    return t;
}
void glo_smt_free(glo_smt *t) { if (t) { free(t->tab); free(t); } }
void glo_smt_root(const glo_smt *t, u64 out[4]) { memcpy(out, t->root, 32); }

typedef struct {
    int found, is_old0;
    u32 ns;
    u64 sib[256][4];
    u64 value[4], nf_key[4], nf_value[4];
} find_res;

/* find_rec, iteratively: siblings[level] is the other child at that level, top first */
static int smt_find_i(const glo_smt *t, const u64 key[4], find_res *r) {
    memset(r, 0, sizeof *r);
    u64 cur[4];
    memcpy(cur, t->root, 32);
    for (unsigned level = 0;; level++) {
        if (is_zero4(cur)) { r->found = 0; r->is_old0 = 1; return 0; }
        const smt_node *nd = smt_lookup(t, cur);
        if (!nd) return -1; /* "searching node is not found" */
        if (nd->is_leaf) {
            if (eq4(nd->a, key)) { r->found = 1; memcpy(r->value, nd->b, 32); r->is_old0 = 0; }
            else { r->found = 0; memcpy(r->nf_key, nd->a, 32); memcpy(r->nf_value, nd->b, 32); r->is_old0 = 0; }
            return 0;
        }
        if (level >= 256) return -1;
        if (key_bit(key, level)) { memcpy(r->sib[r->ns++], nd->a, 32); memcpy(cur, nd->b, 32); }
        else { memcpy(r->sib[r->ns++], nd->b, 32); memcpy(cur, nd->a, 32); }
    }
}

int glo_smt_find(const glo_smt *t, const u64 key[4], u64 *siblings, u32 *num_siblings, u64 not_found_key[4],
                 u64 value[4], u32 *is_old0) {
    find_res r;
    if (smt_find_i(t, key, &r) < 0) return -1;
    if (siblings) memcpy(siblings, r.sib, (size_t)r.ns * 32);
    if (num_siblings) *num_siblings = r.ns;
    if (not_found_key) memcpy(not_found_key, r.nf_key, 32);
    if (value) memcpy(value, r.found ? r.value : r.nf_value, 32);
    if (is_old0) *is_old0 = (u32)r.is_old0;
    return r.found;
}

static void proof_set_siblings(glo_smt_process_proof *pf, u64 sib[][4], u32 ns) {
    memset(pf->siblings, 0, sizeof pf->siblings);
    memcpy(pf->siblings, sib, (size_t)ns * 32);
    pf->num_siblings = ns;
}

/* calc_process_proof: value == 0 ? (found ? remove : noop) : (found ? update : insert) */
int glo_smt_set(glo_smt *t, const u64 key[4], const u64 value[4], glo_smt_process_proof *pf) {
    find_res r;
    u64 zero[4] = {0, 0, 0, 0};
    if (smt_find_i(t, key, &r) < 0) return -1;
    memset(pf, 0, sizeof *pf);
    if (is_zero4(value) && !r.found) { /* noop */
        memcpy(pf->old_root, t->root, 32); memcpy(pf->new_root, t->root, 32);
        memcpy(pf->old_key, key, 32); memcpy(pf->new_key, key, 32);
        pf->is_old0 = 1; pf->fnc = 0;
        return 0;
    }
    if (!is_zero4(value) && r.found) { /* update */
        u64 rt_old[4], rt_new[4];
        glo_smt_leaf_hash(key, r.value, rt_old);
        glo_smt_leaf_hash(key, value, rt_new);
        smt_put(t, rt_new, 1, key, value);
        for (int lvl = (int)r.ns - 1; lvl >= 0; lvl--) {
            u64 no[4], nn[4];
            if (key_bit(key, (unsigned)lvl)) {
                glo_two_to_one(r.sib[lvl], rt_old, no); glo_two_to_one(r.sib[lvl], rt_new, nn);
                smt_put(t, nn, 0, r.sib[lvl], rt_new);
            } else {
                glo_two_to_one(rt_old, r.sib[lvl], no); glo_two_to_one(rt_new, r.sib[lvl], nn);
                smt_put(t, nn, 0, rt_new, r.sib[lvl]);
            }
            memcpy(rt_old, no, 32); memcpy(rt_new, nn, 32);
        }
        memcpy(pf->old_root, t->root, 32); memcpy(pf->old_key, key, 32); memcpy(pf->old_value, r.value, 32);
        memcpy(pf->new_root, rt_new, 32); memcpy(pf->new_key, key, 32); memcpy(pf->new_value, value, 32);
        proof_set_siblings(pf, r.sib, r.ns);
        pf->is_old0 = 0; pf->fnc = 1;
        memcpy(t->root, rt_new, 32);
        return 0;
    }
    if (!is_zero4(value)) { /* insert */
        u32 ns = r.ns;
        int mixed, added_one;
        u64 rt_old[4];
        if (!r.is_old0) {
            for (unsigned i = ns; i < 256; i++) {
                if (key_bit(r.nf_key, i) != key_bit(key, i)) break;
                memset(r.sib[ns++], 0, 32);
            }
            glo_smt_leaf_hash(r.nf_key, r.nf_value, rt_old);
            memcpy(r.sib[ns++], rt_old, 32);
            added_one = 1; mixed = 0;
        } else {
            mixed = r.ns != 0; added_one = 0;
            memset(rt_old, 0, 32);
        }
        u64 rt[4];
        glo_smt_leaf_hash(key, value, rt);
        smt_put(t, rt, 1, key, value);
        for (int lvl = (int)ns - 1, level = 0; lvl >= 0; lvl--, level++) {
            if (level != 0 && !is_zero4(r.sib[lvl])) mixed = 1;
            int bit = key_bit(key, (unsigned)lvl);
            if (mixed) {
                u64 no[4];
                if (bit) glo_two_to_one(r.sib[lvl], rt_old, no); else glo_two_to_one(rt_old, r.sib[lvl], no);
                memcpy(rt_old, no, 32);
            }
            u64 nn[4];
            if (bit) { glo_two_to_one(r.sib[lvl], rt, nn); smt_put(t, nn, 0, r.sib[lvl], rt); }
            else { glo_two_to_one(rt, r.sib[lvl], nn); smt_put(t, nn, 0, rt, r.sib[lvl]); }
            memcpy(rt, nn, 32);
        }
        if (added_one) ns--;
        while (ns > 0 && is_zero4(r.sib[ns - 1])) ns--;
        memcpy(pf->old_root, t->root, 32); memcpy(pf->old_key, r.nf_key, 32); memcpy(pf->old_value, r.nf_value, 32);
        memcpy(pf->new_root, rt, 32); memcpy(pf->new_key, key, 32); memcpy(pf->new_value, value, 32);
        proof_set_siblings(pf, r.sib, ns);
        pf->is_old0 = (u32)r.is_old0; pf->fnc = 2;
        memcpy(t->root, rt, 32);
        return 0;
    }
    /* remove */
    {
        u64 rt_old[4], rt_new[4], res_old_key[4], res_old_value[4];
        int mixed, res_is_old0;
        glo_smt_leaf_hash(key, r.value, rt_old);
        if (r.ns > 0) {
            const smt_node *nx = smt_lookup(t, r.sib[r.ns - 1]);
            if (nx && nx->is_leaf) {
                mixed = 0; memcpy(res_old_key, nx->a, 32); memcpy(res_old_value, nx->b, 32);
                res_is_old0 = 0; memcpy(rt_new, r.sib[r.ns - 1], 32);
            } else if (nx) {
                mixed = 1; memcpy(res_old_key, key, 32); memset(res_old_value, 0, 32);
                res_is_old0 = 1; memset(rt_new, 0, 32);
            } else if (is_zero4(r.sib[r.ns - 1])) {
                /* upstream: nodes_db.get(zero) -> None -> unreachable!(); cannot occur in a tree built by set() */
                return -2;
            } else return -2;
        } else {
            mixed = 0; memcpy(res_old_key, key, 32); memset(res_old_value, 0, 32);
            res_is_old0 = 1; memset(rt_new, 0, 32);
        }
        u64 res_sib[256][4];
        u32 rns = 0;
        for (int lvl = (int)r.ns - 1, level = 0; lvl >= 0; lvl--, level++) {
            const u64 *new_sibling = (level == 0 && !res_is_old0) ? zero : r.sib[lvl];
            int bit = key_bit(key, (unsigned)lvl);
            u64 no[4];
            if (bit) glo_two_to_one(r.sib[lvl], rt_old, no); else glo_two_to_one(rt_old, r.sib[lvl], no);
            memcpy(rt_old, no, 32);
            if (!is_zero4(new_sibling)) mixed = 1;
            if (mixed) {
                memmove(res_sib[1], res_sib[0], (size_t)rns * 32); /* push front */
                memcpy(res_sib[0], r.sib[lvl], 32);
                rns++;
                u64 nn[4];
                if (bit) { glo_two_to_one(new_sibling, rt_new, nn); smt_put(t, nn, 0, new_sibling, rt_new); }
                else { glo_two_to_one(rt_new, new_sibling, nn); smt_put(t, nn, 0, rt_new, new_sibling); }
                memcpy(rt_new, nn, 32);
            }
        }
        memcpy(pf->old_root, rt_old, 32); memcpy(pf->old_key, key, 32); memcpy(pf->old_value, r.value, 32);
        memcpy(pf->new_root, rt_new, 32); memcpy(pf->new_key, res_old_key, 32); memcpy(pf->new_value, res_old_value, 32);
        proof_set_siblings(pf, res_sib, rns);
        pf->is_old0 = (u32)res_is_old0; pf->fnc = 3;
        memcpy(t->root, rt_new, 32);
        return 0;
    }
}

/* ------------------------------------------------------------------------------------------------
 * P8/P10  plonky2::fri::prover::fri_committed_trees (one layer) and fri_proof_of_work
 * ---------------------------------------------------------------------------------------------- */
/* reverse_index_bits_in_place(values); leaves = values.chunks(arity).map(flatten); MerkleTree::new */
int glo_fri_layer_tree(const u64 *values_ext, size_t len, unsigned arity_bits, unsigned cap_height,
                       u64 *leaves_out, u64 *digests_out, u64 *cap_out) {
    unsigned lg = log2_strict(len);
    size_t arity = (size_t)1 << arity_bits;
    size_t num_leaves = len >> arity_bits, leaf_len = 2 * arity;
    u64 *leaves = leaves_out ? leaves_out : (u64 *)malloc(len * 2 * sizeof(u64));
    for (size_t i = 0; i < len; i++) {
        size_t src = glo_reverse_bits(i, lg);
        leaves[2 * i] = canon(values_ext[2 * src]);
        leaves[2 * i + 1] = canon(values_ext[2 * src + 1]);
    }
    int rc = glo_merkle_tree(leaves, num_leaves, leaf_len, cap_height, digests_out, cap_out);
    if (!leaves_out) free(leaves);
    return rc;
}

/* coeffs.chunks_exact(arity).map(|chunk| reduce_with_powers(chunk, beta)) */
void glo_fri_fold(const u64 *coeffs_ext, size_t len, unsigned arity_bits, const u64 beta[2], u64 *folded) {
    size_t arity = (size_t)1 << arity_bits;
    for (size_t k = 0; k < len >> arity_bits; k++) {
        u64 acc[2] = {0, 0};
        for (size_t j = arity; j-- > 0;) {
            u64 t[2];
            glo_ext_mul(acc, beta, t);
            ext_add(t, coeffs_ext + 2 * (k * arity + j), acc);
        }
        folded[2 * k] = acc[0];
        folded[2 * k + 1] = acc[1];
    }
}

/* PolynomialCoeffs<F::Extension>::coset_fft(shift.into()): roots and shift are base-field, so the two
 * coordinates transform independently */
void glo_ext_coset_fft(u64 *a_ext, unsigned lg_n, u64 shift) {
    size_t n = (size_t)1 << lg_n;
    u64 *t = (u64 *)malloc(n * sizeof(u64));
    for (int k = 0; k < 2; k++) {
        for (size_t i = 0; i < n; i++) t[i] = a_ext[2 * i + k];
        glo_coset_fft(t, lg_n, shift, 0);
        for (size_t i = 0; i < n; i++) a_ext[2 * i + k] = t[i];
    }
    free(t);
}

/* fri_proof_of_work: upstream uses rayon find_any (nondeterministic); the deterministic restatement
 * returns the SMALLEST satisfying candidate in [start, start+count), or UINT64_MAX if none. */
u64 glo_pow_grind(const u64 state[12], unsigned pos, unsigned out_pos, unsigned min_lz, u64 start, u64 count) {
    ensure_rc();
    u64 best = UINT64_MAX;
#pragma omp parallel for schedule(static) reduction(min : best)
    for (u64 k = 0; k < count; k++) {
        u64 cand = start + k;
        if (cand >= best) continue;
        u64 s[12];
        memcpy(s, state, sizeof s);
        s[pos] = cand;
        glo_poseidon_permute(s);
        unsigned lz = s[out_pos] ? (unsigned)__builtin_clzll(s[out_pos]) : 64;
        if (lz >= min_lz && cand < best) best = cand;
    }
    return best;
}

/* ------------------------------------------------------------------------------------------------
 * N1  plonky2::fri::oracle::PolynomialBatch::prove_openings, the part before fri_proof
 *     (ReducingFactor::reduce_polys_base, PolynomialCoeffs::divide_by_linear, shift_poly) and the
 *     OpeningSet evaluations the verifier's fri_combine_initial needs.
 * ---------------------------------------------------------------------------------------------- */
/* out = sum_j alpha^j * polys[j]   (polys: k base-field coefficient vectors of length n, pointers) */
void glo_reduce_polys_base(const u64 *const *polys, size_t k, size_t n, const u64 alpha[2], u64 *out_ext) {
    u64 *pw = (u64 *)malloc(2 * k * sizeof(u64));
    u64 cur[2] = {1, 0};
    for (size_t j = 0; j < k; j++) {
        pw[2 * j] = cur[0];
        pw[2 * j + 1] = cur[1];
        u64 t[2];
        glo_ext_mul(cur, alpha, t);
        cur[0] = t[0];
        cur[1] = t[1];
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        u64 a0 = 0, a1 = 0;
        for (size_t j = 0; j < k; j++) {
            u64 c = canon(polys[j][i]);
            a0 = f_add(a0, f_mul(c, pw[2 * j]));
            a1 = f_add(a1, f_mul(c, pw[2 * j + 1]));
        }
        out_ext[2 * i] = canon(a0);
        out_ext[2 * i + 1] = canon(a1);
    }
    free(pw);
}
/* PolynomialCoeffs::divide_by_linear(z) followed by the push(ZERO) of prove_openings: quot has n entries */
void glo_divide_by_linear(const u64 *poly_ext, size_t n, const u64 z[2], u64 *quot_ext) {
    u64 acc[2] = {0, 0};
    for (size_t i = n; i-- > 0;) {
        u64 t[2];
        glo_ext_mul(acc, z, t);
        ext_add(t, poly_ext + 2 * i, acc);      /* acc = b_i = b_{i+1} z + c_i */
        if (i > 0) {
            quot_ext[2 * (i - 1)] = canon(acc[0]);
            quot_ext[2 * (i - 1) + 1] = canon(acc[1]);
        }
    }
    quot_ext[2 * (n - 1)] = 0;
    quot_ext[2 * (n - 1) + 1] = 0;
}
/* acc = acc * scalar + add  (ReducingFactor::shift_poly then +=) */
void glo_ext_poly_scale_add(u64 *acc_ext, size_t n, const u64 scalar[2], const u64 *add_ext) {
    for (size_t i = 0; i < n; i++) {
        u64 t[2];
        glo_ext_mul(acc_ext + 2 * i, scalar, t);
        ext_add(t, add_ext + 2 * i, acc_ext + 2 * i);
        acc_ext[2 * i] = canon(acc_ext[2 * i]);
        acc_ext[2 * i + 1] = canon(acc_ext[2 * i + 1]);
    }
}
/* PolynomialCoeffs<F>::to_extension().eval(point) */
void glo_eval_base_poly_at_ext(const u64 *coeffs, size_t n, const u64 point[2], u64 out[2]) {
    u64 acc[2] = {0, 0};
    for (size_t i = n; i-- > 0;) {
        u64 t[2];
        glo_ext_mul(acc, point, t);
        acc[0] = f_add(t[0], canon(coeffs[i]));
        acc[1] = t[1];
    }
    out[0] = canon(acc[0]);
    out[1] = canon(acc[1]);
}

/* ------------------------------------------------------------------------------------------------
 * N3  plonky2::plonk::prover: the permutation argument's Z / partial products (prover step 5) and
 *     compute_quotient_polys (step 8), for circuits whose gates are NoopGate, ConstantGate,
 *     PublicInputGate (upstream) and the reference's three custom gates:
 *       U32InterleaveGate      /root/reference/src/u32/gates/interleave_u32.rs:89-126
 *       UninterleaveToU32Gate  /root/reference/src/u32/gates/uninterleave_to_u32.rs:98-145
 *       UninterleaveToB32Gate  /root/reference/src/u32/gates/uninterleave_to_b32.rs:101-149
 *     Everything else (vanishing_poly.rs, ZeroPolyOnCoset, selectors) restates upstream v0.1.4 from
 *     memory: "parity unpinned"; the reference-endorsed acceptance property is the low-degree test
 *     (interleave_u32.rs:341-352): a satisfying witness gives a quotient of degree < 7n.
 * ---------------------------------------------------------------------------------------------- */
#define UNUSED_SELECTOR 0xFFFFFFFFULL /* u32::MAX */

static unsigned gate_num_constraints(const glo_gate *g) {
    switch (g->kind) {
        case GLO_GATE_CONSTANT: return g->num_ops;
        case GLO_GATE_PUBLIC_INPUT: return 4;
        case GLO_GATE_U32_INTERLEAVE: return g->num_ops * 34;
        case GLO_GATE_UNINTERLEAVE_TO_U32:
        case GLO_GATE_UNINTERLEAVE_TO_B32: return g->num_ops * 67;
        default: return 0;
    }
}
unsigned glo_num_gate_constraints(const glo_gate *gates, unsigned num_gates) {
    unsigned m = 0;
    for (unsigned i = 0; i < num_gates; i++) {
        unsigned k = gate_num_constraints(gates + i);
        if (k > m) m = k;
    }
    return m;
}

/* plonk_common::reduce_with_powers(terms, alpha) = sum_j terms[j] alpha^j */
static u64 reduce_with_powers(const u64 *terms, unsigned n, long stride, u64 alpha) {
    u64 sum = 0;
    for (unsigned j = n; j-- > 0;) sum = f_add(f_mul(sum, alpha), terms[(long)j * stride]);
    return sum;
}

/* Gate::eval_unfiltered for one row; `consts` = local_constants after remove_prefix(num_selectors) */
static void gate_eval_unfiltered(const glo_gate *g, const u64 *wires, const u64 *consts, const u64 pih[4], u64 *out) {
    unsigned k = 0;
    switch (g->kind) {
        case GLO_GATE_CONSTANT: /* ConstantGate: local_constants[i] - local_wires[i] */
            for (unsigned i = 0; i < g->num_ops; i++) out[k++] = f_sub(consts[i], wires[i]);
            break;
        case GLO_GATE_PUBLIC_INPUT: /* PublicInputGate: local_wires[i] - public_inputs_hash[i] */
            for (unsigned i = 0; i < 4; i++) out[k++] = f_sub(wires[i], pih[i]);
            break;
        case GLO_GATE_U32_INTERLEAVE: { /* interleave_u32.rs:89-126 */
            for (unsigned i = 0; i < g->num_ops; i++) {
                u64 x = wires[2 * i], x_interleaved = wires[2 * i + 1];
                const u64 *bits = wires + g->num_ops * 2 + 32 * i; /* wires_ith_bit_decomposition, big-endian */
                /* reduce_with_powers(bits.iter().rev(), B): term_j = bits[31 - j] */
                out[k++] = f_sub(reduce_with_powers(bits + 31, 32, -1, 2), x);
                out[k++] = f_sub(reduce_with_powers(bits + 31, 32, -1, 4), x_interleaved);
                for (unsigned b = 0; b < 32; b++) out[k++] = f_mul(bits[b], f_sub(bits[b], 1)); /* (bit - 0)(bit - 1) */
            }
            break;
        }
        case GLO_GATE_UNINTERLEAVE_TO_U32:   /* uninterleave_to_u32.rs:98-145 */
        case GLO_GATE_UNINTERLEAVE_TO_B32: { /* uninterleave_to_b32.rs:101-149 */
            for (unsigned i = 0; i < g->num_ops; i++) {
                u64 x_interleaved = wires[3 * i], x_evens = wires[3 * i + 1], x_odds = wires[3 * i + 2];
                const u64 *bits = wires + g->num_ops * 3 + 64 * i;
                out[k++] = f_sub(reduce_with_powers(bits + 63, 64, -1, 2), x_interleaved);
                u64 ev = 0, od = 0;
                for (unsigned j = 0; j < 32; j++) {
                    u64 coeff = g->kind == GLO_GATE_UNINTERLEAVE_TO_U32 ? ((u64)1 << (32 - j - 1)) : ((u64)1 << (2 * (32 - j - 1)));
                    ev = f_add(ev, f_mul(coeff, bits[2 * j]));
                    od = f_add(od, f_mul(coeff, bits[2 * j + 1]));
                }
                out[k++] = f_sub(ev, x_evens);
                out[k++] = f_sub(od, x_odds);
                for (unsigned b = 0; b < 64; b++) out[k++] = f_mul(bits[b], f_sub(bits[b], 1));
            }
            break;
        }
        default: break; /* NoopGate */
    }
}

/* gate.rs compute_filter: prod_{i in group, i != row} (i - s) * (many_selectors ? (UNUSED_SELECTOR - s) : 1) */
static u64 compute_filter(unsigned row, unsigned g0, unsigned g1, u64 s, int many) {
    u64 f = 1;
    for (unsigned i = g0; i < g1; i++)
        if (i != row) f = f_mul(f, f_sub(i, s));
    if (many) f = f_mul(f, f_sub(UNUSED_SELECTOR, s));
    return f;
}

/* prover.rs wires_permutation_partial_products_and_zs / all_wires_permutation_partial_products:
 * wires [num_wires][n], sigmas [num_routed][n] (values k_j' * w^i' of the permutation), out [nch * (1 + num_prods)][n]
 * laid out as the prover commits them: Z_0 .. Z_{nch-1}, then the partial products of challenge 0, 1, ... */
int glo_permutation_zs(const glo_circuit *cd, const u64 *k_is, const u64 *wires, const u64 *sigmas, const u64 *betas,
                       const u64 *gammas, u64 *out) {
    const size_t n = (size_t)1 << cd->degree_bits;
    const unsigned R = cd->num_routed_wires, deg = cd->quotient_degree_factor;
    const unsigned chunks = (R + deg - 1) / deg, num_prods = chunks - 1, nch = cd->num_challenges;
    const u64 w = glo_primitive_root_of_unity(cd->degree_bits);
    u64 *subgroup = (u64 *)malloc(n * sizeof(u64));
    u64 cur = 1;
    for (size_t i = 0; i < n; i++) { subgroup[i] = cur; cur = f_mul(cur, w); }
    for (unsigned c = 0; c < nch; c++) {
        const u64 beta = betas[c], gamma = gammas[c];
        u64 *Z = out + (size_t)c * n;
        u64 *PP = out + ((size_t)nch + (size_t)c * num_prods) * n;
        /* the per-row chunk quotients are independent; the running product over rows is serial */
        u64 *chunkq = (u64 *)malloc(n * chunks * sizeof(u64));
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) {
            const u64 x = subgroup[i];
            for (unsigned q = 0; q < chunks; q++) {
                u64 num = 1, den = 1;
                for (unsigned j = q * deg; j < R && j < (q + 1) * deg; j++) {
                    const u64 wv = wires[(size_t)j * n + i];
                    u64 s_id = f_mul(k_is[j], x);
                    num = f_mul(num, f_add(f_add(wv, f_mul(beta, s_id)), gamma));
                    den = f_mul(den, f_add(f_add(wv, f_mul(beta, sigmas[(size_t)j * n + i])), gamma));
                }
                chunkq[i * chunks + q] = f_mul(num, glo_inv(den));
            }
        }
        u64 z = 1;
        for (size_t i = 0; i < n; i++) {
            Z[i] = z; /* "the last term is Z(gx), but we replace it with Z(x)" */
            u64 acc = z;
            for (unsigned q = 0; q < chunks; q++) {
                acc = f_mul(acc, chunkq[i * chunks + q]);
                if (q < num_prods) PP[(size_t)q * n + i] = acc;
            }
            z = acc;
        }
        free(chunkq);
        if (z != 1) { free(subgroup); return 1; } /* the grand product must close: sigma is not a permutation of equal values */
    }
    free(subgroup);
    return 0;
}

/* vanishing_poly.rs eval_vanishing_poly_base_batch for ONE point x: fills terms[] = vanishing_z_1_terms ++
 * vanishing_partial_products_terms ++ gate constraint slots (the list reduce_with_powers_multi takes).
 * lcs / lw / lz / nz: the rows of constants_sigmas, wires, zs_partial_products at x and of zs_partial_products at g x. */
static unsigned vanishing_terms(const glo_circuit *cd, const glo_gate *gates, const u64 *k_is, u64 x, u64 l0, const u64 *lcs,
                                const u64 *lw, const u64 *lz, const u64 *nz, const u64 pih[4], const u64 *betas,
                                const u64 *gammas, u64 *terms, u64 *gc) {
    const unsigned nch = cd->num_challenges, R = cd->num_routed_wires, deg = cd->quotient_degree_factor;
    const unsigned chunks = (R + deg - 1) / deg, num_prods = chunks - 1;
    const unsigned ngc = glo_num_gate_constraints(gates, cd->num_gates);
    const u64 *s_sigmas = lcs + cd->num_constants;
    const int many = cd->num_selectors > 1;
    unsigned t = 0;
    for (unsigned c = 0; c < nch; c++) terms[t++] = f_mul(l0, f_sub(lz[c], 1)); /* vanishing_z_1_terms */
    for (unsigned c = 0; c < nch; c++) { /* check_partial_products */
        const u64 *pp = lz + nch + c * num_prods;
        for (unsigned q = 0; q < chunks; q++) {
            u64 num = 1, den = 1;
            for (unsigned j = q * deg; j < R && j < (q + 1) * deg; j++) {
                num = f_mul(num, f_add(f_add(lw[j], f_mul(betas[c], f_mul(k_is[j], x))), gammas[c]));
                den = f_mul(den, f_add(f_add(lw[j], f_mul(betas[c], s_sigmas[j])), gammas[c]));
            }
            const u64 prev = q == 0 ? lz[c] : pp[q - 1], next = q == chunks - 1 ? nz[c] : pp[q];
            terms[t++] = f_sub(f_mul(prev, num), f_mul(next, den));
        }
    }
    /* evaluate_gate_constraints_base_batch: every gate, filtered, added into the shared slots */
    u64 *slots = terms + t;
    for (unsigned k = 0; k < ngc; k++) slots[k] = 0;
    for (unsigned gi = 0; gi < cd->num_gates; gi++) {
        const glo_gate *g = gates + gi;
        const unsigned k = gate_num_constraints(g);
        if (!k) continue;
        const u64 filter = compute_filter(gi, g->group_start, g->group_end, lcs[g->selector_index], many);
        gate_eval_unfiltered(g, lw, lcs + cd->num_selectors, pih, gc);
        for (unsigned j = 0; j < k; j++) slots[j] = f_add(slots[j], f_mul(filter, gc[j]));
    }
    return t + ngc;
}
/* The verifier's side of the same identity (plonk/verifier.rs via eval_vanishing_poly): the alpha-reduced vanishing
 * values at an arbitrary base-field point x, given the openings there.  out[nch]. */
void glo_vanishing_at_point(const glo_circuit *cd, const glo_gate *gates, const u64 *k_is, u64 x, const u64 *lcs, const u64 *lw,
                            const u64 *lz, const u64 *nz, const u64 pih[4], const u64 *betas, const u64 *gammas,
                            const u64 *alphas, u64 *out) {
    const size_t n = (size_t)1 << cd->degree_bits;
    const unsigned chunks = (cd->num_routed_wires + cd->quotient_degree_factor - 1) / cd->quotient_degree_factor;
    const unsigned ngc = glo_num_gate_constraints(gates, cd->num_gates);
    const unsigned nterms = cd->num_challenges * (1 + chunks) + ngc;
    u64 *terms = (u64 *)malloc(nterms * sizeof(u64)), *gc = (u64 *)malloc((ngc ? ngc : 1) * sizeof(u64));
    /* eval_l_0(x) = (x^n - 1) / (n (x - 1)) */
    const u64 l0 = f_mul(f_sub(glo_pow(x, (u64)n), 1), glo_inv(f_mul((u64)n % P, f_sub(x, 1))));
    vanishing_terms(cd, gates, k_is, x, l0, lcs, lw, lz, nz, pih, betas, gammas, terms, gc);
    for (unsigned c = 0; c < cd->num_challenges; c++) out[c] = reduce_with_powers(terms, nterms, 1, alphas[c]);
    free(terms);
    free(gc);
}

/* prover.rs compute_quotient_polys + vanishing_poly.rs eval_vanishing_poly_base_batch.
 * *_lde: row-major leaves of the three commits as PolynomialBatch holds them after transpose + reverse_index_bits
 * ([N][c], N = n << rate_bits); get_lde_values(i, step) = leaves[reverse_bits(i * step, lg N)].
 * out: [nch * quotient_degree_factor][n] = quotient_polys.flat_map(|p| p.chunks(n)). */
int glo_quotient_polys(const glo_circuit *cd, const glo_gate *gates, const u64 *k_is, const u64 *cs_lde, const u64 *wires_lde,
                       const u64 *zs_lde, unsigned rate_bits, const u64 pih[4], const u64 *betas, const u64 *gammas,
                       const u64 *alphas, u64 *out) {
    const unsigned lg_n = cd->degree_bits, nch = cd->num_challenges, R = cd->num_routed_wires, deg = cd->quotient_degree_factor;
    unsigned qdb = 0;
    while ((1u << qdb) < deg) qdb++; /* log2_ceil(quotient_degree_factor) */
    if (qdb > rate_bits) return 2;   /* "Having constraints of degree higher than the rate is not supported yet." */
    const size_t n = (size_t)1 << lg_n, lde_size = n << qdb, step = (size_t)1 << (rate_bits - qdb), next_step = (size_t)1 << qdb;
    const unsigned lg_N = lg_n + rate_bits;
    const unsigned chunks = (R + deg - 1) / deg, num_prods = chunks - 1;
    const unsigned c_cs = cd->num_constants + R, c_w = cd->num_wires, c_z = nch * (1 + num_prods);
    const unsigned ngc = glo_num_gate_constraints(gates, cd->num_gates);
    const unsigned nterms = nch + nch * chunks + ngc;
    const u64 w_lde = glo_primitive_root_of_unity(lg_n + qdb);
    /* ZeroPolyOnCoset::new(n_log, rate_bits = qdb): evals[i] = g^n * w_rate^i - 1, inverses */
    u64 zh[64], zh_inv[64];
    {
        const u64 g_pow_n = glo_pow(7, (u64)n), w_rate = glo_primitive_root_of_unity(qdb);
        u64 cur = 1;
        for (size_t i = 0; i < next_step; i++) { zh[i] = f_sub(f_mul(g_pow_n, cur), 1); zh_inv[i] = glo_inv(zh[i]); cur = f_mul(cur, w_rate); }
    }
    u64 *qv = (u64 *)malloc((size_t)nch * lde_size * sizeof(u64)); /* quotient values, transposed: [nch][lde_size] */
    u64 *points = (u64 *)malloc(lde_size * sizeof(u64));            /* F::two_adic_subgroup */
    { u64 cur = 1; for (size_t i = 0; i < lde_size; i++) { points[i] = cur; cur = f_mul(cur, w_lde); } }
#pragma omp parallel
    {
        u64 *terms = (u64 *)malloc(nterms * sizeof(u64));
        u64 *gc = (u64 *)malloc((ngc ? ngc : 1) * sizeof(u64));
#pragma omp for schedule(static)
        for (size_t i = 0; i < lde_size; i++) {
            const u64 x = f_mul(7, points[i]); /* shifted_x = F::coset_shift() * x */
            const size_t i_next = (i + next_step) % lde_size;
            const u64 *lcs = cs_lde + glo_reverse_bits(i * step, lg_N) * c_cs;
            const u64 *lw = wires_lde + glo_reverse_bits(i * step, lg_N) * c_w;
            const u64 *lz = zs_lde + glo_reverse_bits(i * step, lg_N) * c_z;
            const u64 *nz = zs_lde + glo_reverse_bits(i_next * step, lg_N) * c_z;
            /* z_h_on_coset.eval_l_0(index, x) = eval(index) * (n * (x - 1))^-1 */
            const u64 l0 = f_mul(zh[i % next_step], glo_inv(f_mul((u64)n % P, f_sub(x, 1))));
            vanishing_terms(cd, gates, k_is, x, l0, lcs, lw, lz, nz, pih, betas, gammas, terms, gc);
            /* reduce_with_powers_multi(vanishing_terms, alphas), then * z_h_on_coset.eval_inverse(i) */
            for (unsigned c = 0; c < nch; c++)
                qv[(size_t)c * lde_size + i] = f_mul(reduce_with_powers(terms, nterms, 1, alphas[c]), zh_inv[i % next_step]);
        }
        free(terms);
        free(gc);
    }
    free(points);
    /* values.coset_ifft(F::coset_shift()), then chunks(degree) */
    for (unsigned c = 0; c < nch; c++) glo_coset_ifft(qv + (size_t)c * lde_size, lg_n + qdb, 7);
    memcpy(out, qv, (size_t)nch * lde_size * sizeof(u64)); /* deg == 2^qdb chunks of n per challenge, in order */
    free(qv);
    return (((size_t)deg << lg_n) == lde_size) ? 0 : 3;
}

int glo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void glo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
