// The constant tables of the FP64 linear layers (csrc/poseidon_constants.h: linear_layer_tables with the signed S-box
// offsets) checked on the CPU with exact integers: the permutation is replayed the way csrc/poseidon.cuh schedules it --
// x^7 handed over as (r0 - r2 - r3, r1 + r2), full rounds through CIRC + 8 e0 e0^T with the next round's constants,
// the 22 partial rounds as 11 fused pairs (CIRC^2 form) -- and must (a) keep every accumulator the fold sees inside
// [0, 2^50) and (b) give the state of the plain definition (add constants, x^7, MDS mod p) on random and extreme inputs.
#include <cstdio>
#include <cstdlib>

#include "../../plonky2-lib_b200/csrc/poseidon_constants.h"

typedef unsigned __int128 u128;
typedef __int128 i128;
typedef uint64_t u64;
static const u64 P = 0xFFFFFFFF00000001ULL;
static const int CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static int mds(int r, int j) { return CIRC[(j - r + 12) % 12] + ((r == 0 && j == 0) ? 8 : 0); }
static u64 mulp(u64 a, u64 b) { return (u64)((u128)a * b % P); }
static u64 pow7(u64 x) {
    x %= P;
    u64 x2 = mulp(x, x), x4 = mulp(x2, x2), x3 = mulp(x2, x);
    return mulp(x3, x4);
}
static int bad = 0;
static void check(bool ok, const char* what) {
    if (!ok && bad++ < 10) std::printf("FAIL: %s\n", what);
}
// the device's hand-over: limbs of the last 128-bit product, not reduced
static void sbox_split(u64 x, i128& dl, i128& dh) {
    x %= P;
    const u64 x2 = mulp(x, x), x4 = mulp(x2, x2), x3 = mulp(x2, x);
    const u128 p = (u128)x3 * x4;
    const u64 lo = (u64)p, hi = (u64)(p >> 64);
    dl = (i128)(lo & 0xFFFFFFFFULL) - (i128)(hi & 0xFFFFFFFFULL) - (i128)(hi >> 32);
    dh = (i128)(lo >> 32) + (i128)(hi & 0xFFFFFFFFULL);
    check((u64)(((dl % (i128)P) + (i128)P + (dh % (i128)P) * (((i128)1 << 32) % (i128)P)) % (i128)P) == pow7(x), "split value");
}
static u64 fold(i128 A, i128 B, const char* where) {
    check(A >= 0 && A < ((i128)1 << 50) && B >= 0 && B < ((i128)1 << 50), where);
    return (u64)(((u128)A + ((u128)B << 32)) % P);
}

int main() {
    static u64 rc[360], split[31 * 12 * 2], pk[11 * 12 * 2];
    if (!poseidon_constants::generate(rc)) return 2;
    poseidon_constants::linear_layer_tables(rc, CIRC, 8, true, split, pk);
    u64 seed = 0x9E3779B97F4A7C15ULL;
    auto rnd = [&]() {
        seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17;
        return seed;
    };
    for (int t = 0; t < 400; t++) {
        u64 s[12], ref[12];
        for (int i = 0; i < 12; i++) {
            u64 v = rnd();
            if (t == 0) v = 0;
            if (t == 1) v = P - 1;
            if (t == 2) v = 0xFFFFFFFFFFFFFFFFULL;
            if (t == 3) v = (i & 1) ? 0xFFFFFFFF00000000ULL : 0xFFFFFFFFULL;
            s[i] = ref[i] = v;
        }
        // ---- the definition
        for (int r = 0; r < 30; r++) {
            u64 x[12];
            for (int i = 0; i < 12; i++) x[i] = (u64)(((u128)(ref[i] % P) + rc[r * 12 + i]) % P);
            if (r < 4 || r >= 26) for (int i = 0; i < 12; i++) x[i] = pow7(x[i]);
            else x[0] = pow7(x[0]);
            for (int i = 0; i < 12; i++) {
                u128 acc = 0;
                for (int j = 0; j < 12; j++) acc += (u128)mds(i, j) * x[j];
                ref[i] = (u64)(acc % P);
            }
        }
        // ---- the device's schedule
        for (int i = 0; i < 12; i++) s[i] = (u64)(((u128)(s[i] % P) + rc[i]) % P);
        for (int phase = 0; phase < 2; phase++) {
            for (int r = 0; r < 4; r++) {
                const u64* k = split + 2 * 12 * ((phase ? 27 : 1) + r);
                i128 dl[12], dh[12];
                for (int j = 0; j < 12; j++) sbox_split(s[j], dl[j], dh[j]);
                for (int i = 0; i < 12; i++) {
                    i128 A = (i128)k[2 * i], B = (i128)k[2 * i + 1];
                    for (int j = 0; j < 12; j++) { A += (i128)mds(i, j) * dl[j]; B += (i128)mds(i, j) * dh[j]; }
                    s[i] = fold(A, B, "full round accumulator");
                }
            }
            if (phase) break;
            for (int pair = 0; pair < 11; pair++) {
                i128 dl[12], dh[12];
                for (int i = 1; i < 12; i++) { dl[i] = (i128)(s[i] & 0xFFFFFFFFULL); dh[i] = (i128)(s[i] >> 32); }
                sbox_split(s[0], dl[0], dh[0]);
                i128 yl = 0, yh = 0;   // Yraw = M[0] . x~
                for (int j = 0; j < 12; j++) { yl += (i128)mds(0, j) * dl[j]; yh += (i128)mds(0, j) * dh[j]; }
                const u64* ky = split + 2 * 12 * (4 + 2 * pair + 1);
                i128 gl, gh;
                sbox_split(fold(yl + (i128)ky[0], yh + (i128)ky[1], "pair y0 accumulator"), gl, gh);
                const i128 wl = 8 * dl[0] - yl, wh = 8 * dh[0] - yh;
                u64 out[12];
                for (int r = 0; r < 12; r++) {
                    i128 A = (i128)pk[2 * (pair * 12 + r)], B = (i128)pk[2 * (pair * 12 + r) + 1];
                    for (int j = 0; j < 12; j++) {          // CIRC^2 x~
                        i128 c2 = 0;
                        for (int m = 0; m < 12; m++) c2 += (i128)CIRC[(m - r + 12) % 12] * CIRC[(j - m + 12) % 12];
                        A += c2 * dl[j];
                        B += c2 * dh[j];
                    }
                    const int c0 = mds(r, 0) - (r == 0 ? 8 : 0);   // col0(CIRC)
                    A += (i128)c0 * wl + (i128)mds(r, 0) * gl;
                    B += (i128)c0 * wh + (i128)mds(r, 0) * gh;
                    out[r] = fold(A, B, "pair accumulator");
                }
                for (int r = 0; r < 12; r++) s[r] = out[r];
            }
        }
        for (int i = 0; i < 12; i++) check(s[i] % P == ref[i] % P, "permutation differs from the definition");
    }
    if (bad) { std::printf("%d failures\n", bad); return 1; }
    std::puts("poseidon_tables_test ok");
    return 0;
}
