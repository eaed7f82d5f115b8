// Two-pass Goldilocks NTT for 2^16 .. 2^20 points on the Blackwell async-copy path.
//
// plonky2::fri::oracle::PolynomialBatch::from_coeffs runs, per polynomial, lde(rate_bits) + coset_fft(7) and then
// transposes and bit-reverses the values into leaves; here the same numbers come out of 2^rate_bits size-n coset
// transforms (decimation in frequency, natural order in, bit-reversed = leaf order out), each as two passes:
//
//   pass 1 (k_ntt_tma_strided): 256-point DFTs over the top 8 index bits.  The column is a [256][2^s] matrix; one work
//       item is the [256 rows][16 columns] tile (128-byte row segments), fetched by ONE cp.async.bulk.tensor (TMA) box
//       load into shared memory and written back by one TMA box store, so HBM sees whole 128-byte lines both ways.
//       Coset shift and four-step twiddle are one coalesced table read: T[coset][row][j] = (shift * w^brev(row))^j.
//   pass 2 (k_ntt_tma_contig): 2^s-point DFTs (s = 8..12) over contiguous 4096-element (32 KB) blocks, fetched by
//       one cp.async.bulk each; the finished tile is laid out in the TMA engine's 128-byte swizzle (every thread owns
//       whole 16-byte chunks of 128-byte rows, conflict-free) and leaves through one swizzled box store.
//
// Both kernels are persistent (two CTAs of 256 threads per SM), double-buffered: while a CTA works on tile t the TMA
// engine fills the other buffer with tile t + 1, so no warp ever waits on a strided global load.
//
// Arithmetic: 16 elements per thread, radix-16 rounds in registers.  Inside a round the values are kept in
// carry-save form (lo, hi, c) = lo + 2^32 hi + 2^64 c with a small signed c: a butterfly output is a three-instruction
// carry chain (IADD3, IADD3.X, IMAD.X) instead of an 8..12-instruction modular add / subtract, the power-of-two
// twiddles inside the block (w_16 = 2^156) are limb shifts with 2^64 = 2^32 - 1, 2^96 = -1, and each value is folded
// back to 64 bits once per round, before the table twiddle multiplies it.
#include <cuda.h>
#include <cuda_runtime.h>

#include "gl_field.cuh"
#include "ntt_kernels.h"
#include "ntt_radix16.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// carry-save field values
// ------------------------------------------------------------------------------------------------
struct cs {
    u32 lo, hi;
    int c;   // value = lo + 2^32 hi + 2^64 c  (mod p), |c| small
};
GL_D cs cs_from(u64 v) { return {(u32)v, (u32)(v >> 32), 0}; }
GL_D cs cs_add(cs a, cs b) {
    cs r;
    asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
        : "=r"(r.lo), "=r"(r.hi), "=r"(r.c)
        : "r"(a.lo), "r"(a.hi), "r"(a.c), "r"(b.lo), "r"(b.hi), "r"(b.c));
    return r;
}
GL_D cs cs_sub(cs a, cs b) {
    cs r;
    asm("sub.cc.u32 %0, %3, %6;\n\tsubc.cc.u32 %1, %4, %7;\n\tsubc.u32 %2, %5, %8;"
        : "=r"(r.lo), "=r"(r.hi), "=r"(r.c)
        : "r"(a.lo), "r"(a.hi), "r"(a.c), "r"(b.lo), "r"(b.hi), "r"(b.c));
    return r;
}
// x * 2^E, 0 <= E < 96.  With phi = 2^32: phi^2 = phi - 1, phi^3 = -1, phi^4 = -phi, phi^5 = -phi^2.  The shifted value
// is t0 + t1 phi + t2 phi^2 + t3 phi^3 with t3 a small signed number; times phi^a it becomes A + B phi with A, B sums of
// at most three limbs, accumulated with carries into (lo, hi, c).
template <int E>
GL_D cs cs_shl(cs x) {
    constexpr int a = E / 32, b = E % 32;
    u32 t0, t1, t2;
    int t3;
    if constexpr (b == 0) {
        t0 = x.lo; t1 = x.hi; t2 = (u32)x.c; t3 = x.c >> 31;
    } else {
        t0 = x.lo << b;
        t1 = __funnelshift_l(x.lo, x.hi, b);
        t2 = __funnelshift_l(x.hi, (u32)x.c, b);
        t3 = x.c >> (32 - b);
    }
    const int s3 = t3 >> 31;
    cs r;
    if constexpr (a == 0) {
        // (t0 - t2 - t3) + (t1 + t2) phi
        asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, 0;\n\tsubc.u32 %2, 0, 0;\n\t"
            "add.cc.u32 %1, %1, %5;\n\taddc.u32 %2, %2, 0;\n\t"
            "sub.cc.u32 %0, %0, %6;\n\tsubc.cc.u32 %1, %1, %7;\n\tsubc.u32 %2, %2, %7;"
            : "=&r"(r.lo), "=&r"(r.hi), "=&r"(r.c)
            : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(s3));
    } else if constexpr (a == 1) {
        // (-t1 - t2) + (t0 + t1 - t3) phi
        asm("sub.cc.u32 %0, 0, %4;\n\tsubc.cc.u32 %1, %3, 0;\n\tsubc.u32 %2, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, %5;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 %2, %2, 0;\n\t"
            "add.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;\n\t"
            "sub.cc.u32 %1, %1, %6;\n\tsubc.u32 %2, %2, %7;"
            : "=&r"(r.lo), "=&r"(r.hi), "=&r"(r.c)
            : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(s3));
    } else {
        // (-t0 - t1 + t3) + (t0 - t2 - t3) phi
        asm("sub.cc.u32 %0, 0, %3;\n\tsubc.cc.u32 %1, %3, 0;\n\tsubc.u32 %2, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, %4;\n\tsubc.cc.u32 %1, %1, %5;\n\tsubc.u32 %2, %2, 0;\n\t"
            "add.cc.u32 %0, %0, %6;\n\taddc.cc.u32 %1, %1, %7;\n\taddc.u32 %2, %2, %7;\n\t"
            "sub.cc.u32 %1, %1, %6;\n\tsubc.u32 %2, %2, %7;"
            : "=&r"(r.lo), "=&r"(r.hi), "=&r"(r.c)
            : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(s3));
    }
    return r;
}
// -> a u64 congruent to the value (not canonical).  2^64 c = c (2^32 - 1) = S, a signed 64-bit number; lo:hi + S wraps
// by k = carry - [c < 0] in {-1, 0, 1} times 2^64, folded back as k (2^32 - 1); for |c| < 2^20 that last step cannot
// wrap again (k = 1 leaves a sum below |S|, k = -1 one above 2^64 - |S|).
GL_D u64 cs_norm(cs a) {
    u32 r0, r1;
    asm("{\n\t.reg .u32 sx, s0, s1, k, nk, sg;\n\t"
        "shr.s32 sx, %4, 31;\n\t"
        "sub.cc.u32 s0, 0, %4;\n\t"
        "subc.u32 s1, %4, sx;\n\t"
        "add.cc.u32 %0, %2, s0;\n\t"
        "addc.cc.u32 %1, %3, s1;\n\t"
        "addc.u32 k, sx, 0;\n\t"
        "sub.u32 nk, 0, k;\n\t"
        "shr.s32 sg, k, 31;\n\t"
        "add.cc.u32 %0, %0, nk;\n\t"
        "addc.u32 %1, %1, sg;\n\t}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a.lo), "r"(a.hi), "r"(a.c));
    return gl_pack(r0, r1);
}

// (u - v) * w_16^J, w_16 = 2^156 (forward) or 2^36 (inverse); 2^96 = -1 flips the subtraction
template <bool INV, int J>
GL_D cs cs_diff_times_w16(cs u, cs v) {
    constexpr int E = ((INV ? 36 : 156) * J) % 192;
    if constexpr (E == 0) return cs_sub(u, v);
    else if constexpr (E >= 96) return cs_shl<E - 96>(cs_sub(v, u));
    else return cs_shl<E>(cs_sub(u, v));
}

// The last STAGES stages of the radix-16 DIF block: STAGES = 4 is one 16-point DFT, 3 two 8-point DFTs on x[0..7] and
// x[8..15], 2 four 4-point DFTs, 1 eight butterflies (results in place, bit-reversed inside each block).
template <bool INV, int STAGES>
GL_D void cs_radix16_dif(cs x[16]) {
#define BF(i, j, J)                                  \
    {                                                \
        cs u_ = x[i], v_ = x[j];                     \
        x[i] = cs_add(u_, v_);                       \
        x[j] = cs_diff_times_w16<INV, J>(u_, v_);    \
    }
    if (STAGES >= 4) {
        BF(0, 8, 0) BF(1, 9, 1) BF(2, 10, 2) BF(3, 11, 3) BF(4, 12, 4) BF(5, 13, 5) BF(6, 14, 6) BF(7, 15, 7)
    }
    if (STAGES >= 3) {
        BF(0, 4, 0) BF(1, 5, 2) BF(2, 6, 4) BF(3, 7, 6) BF(8, 12, 0) BF(9, 13, 2) BF(10, 14, 4) BF(11, 15, 6)
    }
    if (STAGES >= 2) {
        BF(0, 2, 0) BF(1, 3, 4) BF(4, 6, 0) BF(5, 7, 4) BF(8, 10, 0) BF(9, 11, 4) BF(12, 14, 0) BF(13, 15, 4)
    }
    if (STAGES >= 1) {
        BF(0, 1, 0) BF(2, 3, 0) BF(4, 5, 0) BF(6, 7, 0) BF(8, 9, 0) BF(10, 11, 0) BF(12, 13, 0) BF(14, 15, 0)
    }
#undef BF
}
template <bool INV, int STAGES>
GL_D void radix16_round(u64 x[16]) {
    cs y[16];
#pragma unroll
    for (int i = 0; i < 16; i++) y[i] = cs_from(x[i]);
    cs_radix16_dif<INV, STAGES>(y);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = cs_norm(y[i]);
}

// ------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX ISA: cp.async.bulk, cp.async.bulk.tensor, mbarrier)
// ------------------------------------------------------------------------------------------------
GL_D u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
GL_D void mbar_init(u64* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
GL_D void mbar_expect_tx(u64* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
GL_D void mbar_wait(u64* bar, unsigned parity) {
    u32 done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
GL_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
GL_D void tma_load_3d(void* dst, const CUtensorMap* tm, u64* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_addr(dst)), "l"(tm), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
GL_D void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tm), "r"(smem_addr(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
GL_D void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
GL_D void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
GL_D void bulk_load_1d(void* dst, const void* src, unsigned bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

constexpr unsigned TILE = 4096;            // elements per work item (32 KB)
constexpr unsigned TILE_BYTES = TILE * 8;

}  // namespace

// ------------------------------------------------------------------------------------------------
// pass 1: 256-point DFTs down the rows of the [256][2^s] view of every column, 16 adjacent columns per item
// ------------------------------------------------------------------------------------------------
template <bool INV>
__global__ void __launch_bounds__(256, 2)
k_ntt_tma_strided(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, ntt_tma_args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    u64* buf = reinterpret_cast<u64*>(smem_raw);   // [2][TILE], element (row, d) at row * 16 + d
    u64* wt = buf + 2 * TILE;                      // w_256^e
    u64* bars = wt + 256;
    const unsigned tid = threadIdx.x, d = tid & 15, q = tid >> 4;
    const u32 tiles = (1u << a.s) >> 4;
    const u64 total = (u64)a.cosets * a.columns * tiles;
    const u64 first = blockIdx.x, step = gridDim.x;
    const u64 count = first < total ? (total - first + step - 1) / step : 0;
    // item = (coset * columns + col) * tiles + jt, advanced by `step` per iteration without dividing again
    const u32 step_jt = (u32)(step % tiles), step_col = (u32)((step / tiles) % a.columns), step_coset = (u32)(step / tiles / a.columns);
    auto advance = [&](u32& jt, u32& col, u32& coset) {
        jt += step_jt;
        col += step_col;
        coset += step_coset;
        if (jt >= tiles) { jt -= tiles; col++; }
        if (col >= a.columns) { col -= a.columns; coset++; }
    };

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    wt[tid] = __ldg(a.wt1 + brev4(tid >> 4) * (tid & 15));   // wt[i * 16 + q] = w_256^(brev4(i) * q)
    __syncthreads();

    // (jt, col, coset) of the tile being computed, and (thread 0 only) of the next tile to fetch
    u32 jt = (u32)(first % tiles), col = (u32)((first / tiles) % a.columns), coset = (u32)(first / tiles / a.columns);
    u32 ljt = jt, lcol = col, lcoset = coset;
    auto issue_load = [&](u64 k) {   // loads are issued in order k = 0, 1, 2, ...
        u64* bar = &bars[k & 1];
        mbar_expect_tx(bar, TILE_BYTES);
        tma_load_3d(buf + (k & 1) * TILE, &tm_in, bar, (int)(ljt << 4), 0, (int)lcol);   // every coset reads the same input tile
        advance(ljt, lcol, lcoset);
    };
    if (tid == 0) {
        if (count > 0) issue_load(0);
        if (count > 1) issue_load(1);
    }

    for (u64 k = 0; k < count; k++, advance(jt, col, coset)) {
        u64* b = buf + (k & 1) * TILE;
        const u64 jlo = ((u64)jt << 4) + d;
        u64 x[16];
        // ---- round 1: rows i * 16 + q --------------------------------------------------------------------------
        u64 rf[16];
        if (a.rowfac) {
#pragma unroll
            for (int i = 0; i < 16; i++) rf[i] = __ldg(a.rowfac + coset * 256 + i * 16 + q);
        }
        mbar_wait(&bars[k & 1], (unsigned)(k >> 1) & 1);
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = b[((i * 16 + q) << 4) + d];
        if (a.rowfac) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = gl_mul(x[i], rf[i]);
        }
        radix16_round<INV, 4>(x);
#pragma unroll
        for (int i = 1; i < 16; i++) x[i] = gl_mul(x[i], wt[i * 16 + q]);
#pragma unroll
        for (int i = 0; i < 16; i++) b[((i * 16 + q) << 4) + d] = x[i];
        __syncthreads();
        // the other buffer's tile (k - 1) was handed to the TMA store at the end of the previous iteration: once
        // the engine has read it out, refill it with tile k + 1
        if (tid == 0 && k >= 1 && k + 1 < count) {
            tma_store_wait_read();
            issue_load(k + 1);
        }
        // ---- round 2: rows q * 16 + i; T[coset][row][jlo] carries coset shift and four-step twiddle ------------
        const u64* tw = a.post3 + ((((u64)coset << 8) + (q << 4)) << a.s) + jlo;
        u64 t[16];
#pragma unroll
        for (int i = 0; i < 16; i++) t[i] = __ldg(tw + ((u64)i << a.s));
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = b[((q * 16 + i) << 4) + d];
        radix16_round<INV, 4>(x);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            u64 v = gl_mul(x[i], t[i]);
            if (a.canonical_out) v = gl_canon(v);
            b[((q * 16 + i) << 4) + d] = v;
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) tma_store_3d(&tm_out, b, (int)(jt << 4), (int)(coset << 8), (int)col);
    }
    if (tid == 0) tma_store_wait_all();
}

// ------------------------------------------------------------------------------------------------
// pass 2: 2^M-point DFTs over contiguous blocks, in place; 4096 elements per work item
// ------------------------------------------------------------------------------------------------
// Layout of the second exchange and of the finished tile: the 128-byte swizzle of the TMA engine (16-byte chunk index
// xor the low three bits of the 128-byte row index), so the tile leaves through one swizzled box store.
GL_D unsigned swz128(unsigned e) { return e ^ (((e >> 4) & 7u) << 1); }

template <int M, bool INV>
__global__ void __launch_bounds__(256, 2) k_ntt_tma_contig(const __grid_constant__ CUtensorMap tm_out, ntt_tma_args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    u64* buf = reinterpret_cast<u64*>(smem_raw);   // [2][TILE]
    constexpr unsigned Q = 1u << (M - 4);          // threads per DFT
    constexpr unsigned LB = M - 8, GB = 4 - LB;    // bits of the third round / of its independent blocks
    // At M = 12 the second round reads one 128-byte run per half-warp from the natural layout; smaller M need the
    // first exchange swizzled (and a barrier between the round-1 reads and writes, which then touch different words).
    constexpr bool SWZ1 = M < 12;
    // twiddles by [i][thread] so that a warp reads consecutive words: wa[i * Q + q] = w^(brev4(i) * q) after round 1,
    // wb[i * 2^LB + plow] = w^(16 * brev4(i) * plow) after round 2 (w = w_{2^M})
    u64* wa = buf + 2 * TILE;
    u64* wb = wa + 16 * Q;
    u64* bars = wb + (16u << LB);
    const unsigned tid = threadIdx.x, q = tid & (Q - 1), dbase = (tid >> (M - 4)) << M;
    const u64 tiles = ((u64)1 << (a.s + 8)) / TILE;
    const u64 total = (u64)a.cosets * a.columns * tiles;
    const u64 first = blockIdx.x, step = gridDim.x;
    const u64 count = first < total ? (total - first + step - 1) / step : 0;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (unsigned e = tid; e < 16 * Q; e += 256) wa[e] = __ldg(a.wt2 + brev4(e / Q) * (e % Q));
    for (unsigned e = tid; e < (16u << LB); e += 256) wb[e] = __ldg(a.wt2 + ((brev4(e >> LB) * (e & ((1u << LB) - 1))) << 4));
    __syncthreads();

    // item -> (row of the [columns][rows][16] view, column)
    auto locate = [&](u64 k, u64& row, u32& col) {
        const u64 item = first + k * step;
        const u64 tile = item % tiles, rest = item / tiles;
        col = (u32)(rest % a.columns);
        row = ((rest / a.columns) * a.out_coset_stride + tile * TILE) >> 4;
    };
    auto issue_load = [&](u64 k) {
        u64 row; u32 col;
        locate(k, row, col);
        u64* bar = &bars[k & 1];
        mbar_expect_tx(bar, TILE_BYTES);
        bulk_load_1d(buf + (k & 1) * TILE, a.out + (u64)col * a.out_ld + (row << 4), TILE_BYTES, bar);
    };
    if (tid == 0) {
        if (count > 0) issue_load(0);
        if (count > 1) issue_load(1);
    }

    for (u64 k = 0; k < count; k++) {
        u64* tile = buf + (k & 1) * TILE;
        u64* b = tile + dbase;
        u64 x[16];
        // ---- round 1: positions i * Q + q ----------------------------------------------------------------------
        mbar_wait(&bars[k & 1], (unsigned)(k >> 1) & 1);
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = b[i * Q + q];
        if constexpr (SWZ1) __syncthreads();
        radix16_round<INV, 4>(x);
#pragma unroll
        for (int i = 1; i < 16; i++) x[i] = gl_mul(x[i], wa[i * Q + q]);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const unsigned p = i * Q + q;
            b[SWZ1 ? (p ^ ((unsigned)i << (M - 8))) : p] = x[i];
        }
        __syncthreads();
        // the other buffer's tile (k - 1) went to the TMA store at the end of the previous iteration: once the
        // engine has read it out, refill it with tile k + 1
        if (tid == 0 && k >= 1 && k + 1 < count) {
            tma_store_wait_read();
            issue_load(k + 1);
        }
        // ---- round 2: positions phigh * Q + i * 2^LB + plow ------------------------------------------------------
        const unsigned plow = q & ((1u << LB) - 1), phigh = q >> LB;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const unsigned p = phigh * Q + ((unsigned)i << LB) + plow;
            x[i] = b[SWZ1 ? (p ^ (phigh << (M - 8))) : p];
        }
        radix16_round<INV, 4>(x);
        if constexpr (LB == 0) {
            // in-place position p = phigh * 16 + i holds output index brev_M(p): the thread owns one 128-byte row
            // (read above, permuted inside the row), written back in the store layout
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                u64 v0 = x[i], v1 = x[i + 1];
                if (a.canonical_out) { v0 = gl_canon(v0); v1 = gl_canon(v1); }
                *reinterpret_cast<ulonglong2*>(tile + swz128(dbase + phigh * 16 + i)) = make_ulonglong2(v0, v1);
            }
        } else {
#pragma unroll
            for (int i = 1; i < 16; i++) x[i] = gl_mul(x[i], wb[(i << LB) + plow]);
            // ---- round 3: the last LB bits.  Second exchange; a thread then holds, for each of the 2^GB values g of
            // the TOP bits of p, the 2^LB positions f of the low bits: 2^GB independent 2^LB-point DFTs.
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const unsigned p = phigh * Q + ((unsigned)i << LB) + plow;
                tile[swz128(dbase + p)] = x[i];
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const unsigned gg = (unsigned)i >> LB, f = (unsigned)i & ((1u << LB) - 1);
                const unsigned p = (gg << (M - GB)) | (q << LB) | f;
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(tile + swz128(dbase + p));
                x[i] = v.x;
                x[i + 1] = v.y;
            }
            radix16_round<INV, (int)LB>(x);
            // results go back to the words this thread has just read
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const unsigned gg = (unsigned)i >> LB, f = (unsigned)i & ((1u << LB) - 1);
                const unsigned p = (gg << (M - GB)) | (q << LB) | f;
                u64 v0 = x[i], v1 = x[i + 1];
                if (a.canonical_out) { v0 = gl_canon(v0); v1 = gl_canon(v1); }
                *reinterpret_cast<ulonglong2*>(tile + swz128(dbase + p)) = make_ulonglong2(v0, v1);
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            u64 row; u32 col;
            locate(k, row, col);
            tma_store_3d(&tm_out, tile, 0, (int)row, (int)col);
        }
    }
    if (tid == 0) tma_store_wait_all();
}

// ------------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------------
// base^e from a 3 x 1024 table (e < 2^30)
static __device__ __forceinline__ u64 powtab3(const u64* __restrict__ tab, u64 e) {
    u64 r = __ldg(tab + (e & 1023));
    if (e >> 10) {
        r = gl_mul(r, __ldg(tab + 1024 + ((e >> 10) & 1023)));
        if (e >> 20) r = gl_mul(r, __ldg(tab + 2048 + ((e >> 20) & 1023)));
    }
    return r;
}
// T[coset][row][j] = w^(brev8(row) * j) * shift_coset^j; rowfac[coset][row] = shift_coset^(row << s)
__global__ void __launch_bounds__(256)
k_ntt_tma_tables(u64* __restrict__ T, u64* __restrict__ rowfac, const u64* __restrict__ post_tab,
                 const u64* __restrict__ pre_tabs, unsigned s, u32 cosets) {
    const u64 total = ((u64)cosets << 8) << s;
    for (u64 idx = blockIdx.x * (u64)256 + threadIdx.x; idx < total; idx += (u64)gridDim.x * 256) {
        const u64 j = idx & (((u64)1 << s) - 1);
        const unsigned row = (unsigned)(idx >> s) & 255u;
        const u64 coset = idx >> (s + 8);
        u64 v = powtab3(post_tab, (u64)(__brev(row) >> 24) * j);
        if (pre_tabs) v = gl_mul(v, powtab3(pre_tabs + coset * 3072, j));
        T[idx] = gl_canon(v);
        if (rowfac && j == 0) rowfac[coset * 256 + row] = gl_canon(powtab3(pre_tabs + coset * 3072, (u64)row << s));
    }
}
void launch_ntt_tma_tables(u64* T, u64* rowfac, const u64* post_tab, const u64* pre_tabs, unsigned s, u32 cosets,
                           cudaStream_t st) {
    u64 total = ((u64)cosets << 8) << s;
    u64 blocks = (total + 255) / 256;
    if (blocks > 8192) blocks = 8192;
    k_ntt_tma_tables<<<(unsigned)blocks, 256, 0, st>>>(T, pre_tabs ? rowfac : nullptr, post_tab, pre_tabs, s, cosets);
    ++g_gl_launches;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace {
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (encode_tiled_fn)p;
    }();
    return fn;
}
// [columns][rows][width] u64 view, box = [1][256][16] (32 KB); swizzled = 128-byte swizzle (width 16: rows are 128 bytes)
bool make_map(CUtensorMap* tm, const u64* base, u64 width, u64 rows, u64 ld, u32 columns, bool swizzled) {
    encode_tiled_fn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {width, rows, columns};
    cuuint64_t strides[2] = {width * 8, ld * 8};
    cuuint32_t box[3] = {16, 256, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<u64*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzled ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
int g_sms = 0;
template <int M>
void launch_contig(const CUtensorMap& tm, const ntt_tma_args& a, bool inverse, unsigned grid, cudaStream_t st) {
    const size_t smem = 2 * TILE_BYTES + ((size_t)8 << M) + ((size_t)128 << (M - 8)) + 64;
    if (inverse) {
        cudaFuncSetAttribute(k_ntt_tma_contig<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_ntt_tma_contig<M, true><<<grid, 256, smem, st>>>(tm, a);
    } else {
        cudaFuncSetAttribute(k_ntt_tma_contig<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_ntt_tma_contig<M, false><<<grid, 256, smem, st>>>(tm, a);
    }
    ++g_gl_launches;
}
}  // namespace

bool ntt_tma_supported(unsigned L) { return L >= 16 && L <= 20 && get_encode() != nullptr; }

bool launch_ntt_tma(const ntt_tma_job& j, cudaStream_t st) {
    if (!ntt_tma_supported(j.L)) return false;
    const unsigned s = j.L - 8;
    const u64 n = (u64)1 << j.L;
    if (j.cosets > 1 && j.in_coset_stride != 0) return false;   // every coset reads the same coefficients (the LDE)
    if (j.cosets > 1 && j.out_coset_stride != n) return false;
    if (g_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    CUtensorMap tm_in, tm_out;
    if (!make_map(&tm_in, j.in, (u64)1 << s, 256ull, j.in_ld, j.columns, false)) return false;
    if (!make_map(&tm_out, j.out, (u64)1 << s, 256ull * j.cosets, j.out_ld, j.columns, false)) return false;
    CUtensorMap tm_rows;   // the output as 128-byte rows, for the swizzled store of pass 2
    if (!make_map(&tm_rows, j.out, 16, (n * j.cosets) >> 4, j.out_ld, j.columns, true)) return false;
    ntt_tma_args a;
    a.wt1 = j.wt1; a.rowfac = j.rowfac; a.post3 = j.post3;
    a.out = j.out; a.out_ld = j.out_ld; a.out_coset_stride = j.cosets > 1 ? j.out_coset_stride : 0;
    a.wt2 = j.wt2; a.s = s; a.columns = j.columns; a.cosets = j.cosets;
    a.canonical_out = 0;
    const u64 items1 = (u64)j.cosets * j.columns * (((u64)1 << s) >> 4);
    const u64 items2 = (u64)j.cosets * j.columns * (n / TILE);
    const unsigned cap = 2u * (unsigned)g_sms;
    const size_t smem1 = 2 * TILE_BYTES + 256 * 8 + 64;
    const unsigned grid1 = (unsigned)(items1 < cap ? items1 : cap);
    if (j.inverse) {
        cudaFuncSetAttribute(k_ntt_tma_strided<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        k_ntt_tma_strided<true><<<grid1, 256, smem1, st>>>(tm_in, tm_out, a);
    } else {
        cudaFuncSetAttribute(k_ntt_tma_strided<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        k_ntt_tma_strided<false><<<grid1, 256, smem1, st>>>(tm_in, tm_out, a);
    }
    ++g_gl_launches;
    a.canonical_out = j.canonical_out;
    const unsigned grid2 = (unsigned)(items2 < cap ? items2 : cap);
    switch (s) {
        case 8: launch_contig<8>(tm_rows, a, j.inverse, grid2, st); break;
        case 9: launch_contig<9>(tm_rows, a, j.inverse, grid2, st); break;
        case 10: launch_contig<10>(tm_rows, a, j.inverse, grid2, st); break;
        case 11: launch_contig<11>(tm_rows, a, j.inverse, grid2, st); break;
        default: launch_contig<12>(tm_rows, a, j.inverse, grid2, st); break;
    }
    return cudaGetLastError() == cudaSuccess;
}
