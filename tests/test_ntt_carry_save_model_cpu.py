"""The carry-save arithmetic of csrc/ntt_tma.cu and the 160-bit accumulators of csrc/quotient_kernels.cu, restated
instruction by instruction with 32-bit wrap-around integers, against plain arithmetic mod p = 2^64 - 2^32 + 1:

  * a value is (lo, hi, c) = lo + 2^32 hi + 2^64 c with c a signed 32-bit word; add / sub are three-word carry chains;
  * x * 2^E (0 <= E < 96) through limb shifts and phi = 2^32 identities (phi^2 = phi - 1, phi^3 = -1), every one of the
    radix-16 block's twiddles w_16^J (forward w_16 = 2^156, inverse 2^36);
  * the fold back to 64 bits (cs_norm) with its claim that the second wrap cannot happen while |c| < 2^20;
  * a whole radix-16 DIF block on extreme inputs: the results equal the DFT over the field and |c| stays tiny.

These are the exactness arguments the device code relies on; the device itself is checked bit for bit against the
oracle in tests/test_gpu_parity.py (test_fft_tma_path_extreme_inputs, test_commit_tma_path_matches_oracle)."""
import random

P = 0xFFFFFFFF00000001
M32 = 0xFFFFFFFF


def s32(x):  # the signed reading of a 32-bit word
    x &= M32
    return x - (1 << 32) if x >> 31 else x


def val(v):
    lo, hi, c = v
    return (lo + (hi << 32) + (s32(c) << 64)) % P


def cs_from(x):
    return (x & M32, (x >> 32) & M32, 0)


def chain_add(a, b):  # add.cc / addc.cc / addc on three words
    out, carry = [], 0
    for x, y in zip(a, b):
        t = (x & M32) + (y & M32) + carry
        out.append(t & M32)
        carry = t >> 32
    return tuple(out)


def chain_sub(a, b):  # sub.cc / subc.cc / subc
    out, borrow = [], 0
    for x, y in zip(a, b):
        t = (x & M32) - (y & M32) - borrow
        out.append(t & M32)
        borrow = 1 if t < 0 else 0
    return tuple(out)


def cs_shl(v, E):
    lo, hi, c = v
    a, b = divmod(E, 32)
    if b == 0:
        t0, t1, t2, t3 = lo, hi, c & M32, (s32(c) >> 31) & M32
    else:
        t0 = (lo << b) & M32
        t1 = (((hi << 32) | lo) << b >> 32) & M32                  # shf.l.wrap(lo, hi, b)
        t2 = ((((c & M32) << 32) | hi) << b >> 32) & M32           # shf.l.wrap(hi, c, b)
        t3 = (s32(c) >> (32 - b)) & M32                            # arithmetic shift
    s3 = (s32(t3) >> 31) & M32
    if a == 0:      # (t0 - t2 - t3) + (t1 + t2) phi
        r = chain_sub((t0, t1, 0), (t2, 0, 0))
        h = chain_add((r[1], r[2]), (t2, 0))
        r = chain_sub((r[0], h[0], h[1]), (t3, s3, s3))
    elif a == 1:    # (-t1 - t2) + (t0 + t1 - t3) phi
        r = chain_sub((0, t0, 0), (t1, 0, 0))
        r = chain_sub(r, (t2, 0, 0))
        h = chain_add((r[1], r[2]), (t1, 0))
        h = chain_sub(h, (t3, s3))
        r = (r[0], h[0], h[1])
    else:           # (-t0 - t1 + t3) + (t0 - t2 - t3) phi
        r = chain_sub((0, t0, 0), (t0, 0, 0))
        r = chain_sub(r, (t1, t2, 0))
        r = chain_add(r, (t3, s3, s3))
        h = chain_sub((r[1], r[2]), (t3, s3))
        r = (r[0], h[0], h[1])
    return r


def cs_norm(v):
    lo, hi, c = v
    sx = (s32(c) >> 31) & M32
    s = chain_sub((0, c & M32), (c & M32, sx))                     # S = (c << 32) - sext(c), 64 bits
    t = (lo + s[0]) & M32
    c1 = (lo + s[0]) >> 32
    u = hi + s[1] + c1
    k = s32((sx + (u >> 32)) & M32)                                # carry - [c < 0]
    u &= M32
    assert k in (-1, 0, 1)
    nk, sg = (-k) & M32, (k >> 31) & M32
    w0 = t + nk
    w1 = u + sg + (w0 >> 32)
    assert (w1 >> 32) == (1 if k == -1 else 0), "second wrap"      # adding p = 2^64 - (2^32 - 1) wraps by construction
    return (w0 & M32) | ((w1 & M32) << 32)


def twiddle_exponent(inv, J):
    return ((36 if inv else 156) * J) % 192


def diff_times_w16(u, v, inv, J):
    E = twiddle_exponent(inv, J)
    if E == 0:
        return chain_sub(u, v)
    if E >= 96:
        return cs_shl(chain_sub(v, u), E - 96)
    return cs_shl(chain_sub(u, v), E)


BF = [[(i, i + 8, i) for i in range(8)],
      [(0, 4, 0), (1, 5, 2), (2, 6, 4), (3, 7, 6), (8, 12, 0), (9, 13, 2), (10, 14, 4), (11, 15, 6)],
      [(0, 2, 0), (1, 3, 4), (4, 6, 0), (5, 7, 4), (8, 10, 0), (9, 11, 4), (12, 14, 0), (13, 15, 4)],
      [(i, i + 1, 0) for i in range(0, 16, 2)]]


def radix16(xs, inv):
    y = [cs_from(x) for x in xs]
    worst = 0
    for stage in BF:
        for i, j, J in stage:
            u, v = y[i], y[j]
            y[i] = chain_add(u, v)
            y[j] = diff_times_w16(u, v, inv, J)
            worst = max(worst, abs(s32(y[i][2])), abs(s32(y[j][2])))
    return [cs_norm(v) for v in y], worst


EXTREMES = [0, 1, P - 1, P, P + 1, (1 << 64) - 1, 0xFFFFFFFF, 1 << 32, 0xFFFFFFFF00000000, 1 << 63, 0x7FFFFFFF80000001]


def rnd(rng):
    return rng.choice(EXTREMES) if rng.random() < 0.3 else rng.getrandbits(64)


def test_add_sub_shift_and_fold_back():
    rng = random.Random(1)
    for _ in range(3000):
        a, b = rnd(rng), rnd(rng)
        u, v = cs_from(a), cs_from(b)
        # push c away from zero the way a butterfly network does
        for _ in range(rng.randrange(4)):
            u, v = chain_add(u, v), chain_sub(v, u)
        assert val(chain_add(u, v)) == (val(u) + val(v)) % P
        assert val(chain_sub(u, v)) == (val(u) - val(v)) % P
        E = rng.randrange(96)
        assert val(cs_shl(u, E)) == val(u) * pow(2, E, P) % P
        assert abs(s32(cs_shl(u, E)[2])) <= 4
        assert cs_norm(u) % P == val(u)
    for c in (-(1 << 20) + 1, -17, -1, 0, 1, 17, (1 << 20) - 1):          # the bound cs_norm documents
        for x in EXTREMES:
            assert cs_norm((x & M32, x >> 32, c & M32)) % P == (x + (c << 64)) % P


def test_every_twiddle_of_the_radix16_block():
    for inv in (False, True):
        w16 = pow(2, 36 if inv else 156, P)
        assert pow(w16, 16, P) == 1 and pow(w16, 8, P) == P - 1
        for J in range(8):
            for a, b in [(0, 1), (P - 1, 1), ((1 << 64) - 1, 0), (0, (1 << 64) - 1), (12345, 0xFFFFFFFF00000000)]:
                got = val(diff_times_w16(cs_from(a), cs_from(b), inv, J))
                assert got == (a - b) * pow(w16, J, P) % P


def test_radix16_block_is_the_dft_and_c_stays_small():
    rng = random.Random(2)
    brev4 = lambda i: int("{:04b}".format(i)[::-1], 2)  # noqa: E731
    for inv in (False, True):
        w16 = pow(2, 36 if inv else 156, P)
        cases = [[P - 1] * 16, [(1 << 64) - 1] * 16, [(1 << 64) - 1 if i % 2 else 0 for i in range(16)]]
        cases += [[rnd(rng) for _ in range(16)] for _ in range(60)]
        for xs in cases:
            out, worst = radix16(xs, inv)
            assert worst < 64                                      # far below the 2^20 the fold-back allows
            for slot in range(16):                                 # in-place slot i holds output index brev4(i)
                k = brev4(slot)
                want = sum(x * pow(w16, (j * k) % 16, P) for j, x in enumerate(xs)) % P
                assert out[slot] % P == want


def test_quotient_accumulators():
    """acc160 of quotient_kernels.cu: sums of 128-bit products and Horner recompositions as plain integers in five words,
    reduced once: w0 + w1 phi + w2 (phi - 1) - w3 - w4 phi."""
    rng = random.Random(3)

    def reduce5(w):
        return (w[0] + (w[1] << 32) + w[2] * ((1 << 32) - 1) - w[3] - (w[4] << 32)) % P

    for _ in range(200):
        acc, true = 0, 0
        for _ in range(rng.randrange(1, 400)):
            a, b = rnd(rng), rnd(rng)
            acc += a * b
            true = (true + a * b) % P
        assert acc < 1 << 160
        assert reduce5([(acc >> (32 * i)) & M32 for i in range(5)]) == true
        h, th = 0, 0
        for _ in range(32):                                        # base-4 recomposition of 32 arbitrary field elements
            bit = rnd(rng)
            h = (h << 2) + bit
            th = (4 * th + bit) % P
        assert h < 1 << 160
        assert reduce5([(h >> (32 * i)) & M32 for i in range(5)]) == th
