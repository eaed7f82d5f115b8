"""GPU parity tests: every result of libgl_b200.so (through the C ABI) against the CPU oracle on the
same seeded inputs, plus the reference's own known-answer vectors.  Integer work: bit-exact."""
import numpy as np
import pytest

from conftest import P, adversarial_columns, rand_field

pytestmark = pytest.mark.gpu

# src/zkdsa/circuits/mod.rs:85-105 (test_default_simple_signature): PoseidonHash::two_to_one(0, 0)
KAT_TWO_TO_ONE_ZERO = [4330397376401421145, 14124799381142128323, 8742572140681234676, 14345658006221440202]


# ---- P5 Poseidon ---------------------------------------------------------------------------------
def test_reference_kat_two_to_one_zero(glb, ctx):
    out = glb.PoseidonHash.two_to_one(np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64))
    assert out.tolist() == KAT_TWO_TO_ONE_ZERO


def test_permutation_vectors(glb, ctx):
    v = glb.PoseidonHash.permute(np.arange(12, dtype=np.uint64))
    assert [hex(int(x)) for x in v[:4]] == ["0xd64e1e3efc5b8e9e", "0x53666633020aaa47", "0xd40285597c6a8825", "0x613a4f81e81231d2"]
    v = glb.PoseidonHash.permute(np.full(12, P - 1, dtype=np.uint64))
    assert [hex(int(x)) for x in v[:2]] == ["0xbe0085cfc57a8357", "0xd95af71847d05c09"]


def test_permute_batch_matches_oracle(glb, ctx, oracle, rng):
    states = rand_field(rng, (4099, 12))
    states[0] = 0
    states[1] = P - 1
    states[2] = np.uint64(0xFFFFFFFFFFFFFFFF)  # non-canonical input: taken mod p
    states[3] = np.uint64(0xFFFFFFFF)
    got = glb.PoseidonHash.permute(states)
    want = oracle.permute_batch(states)
    assert np.array_equal(got, want)
    assert (got < np.uint64(P)).all()


@pytest.mark.parametrize("length", [0, 1, 4, 5, 7, 8, 9, 12, 16, 20, 32, 85, 135, 136, 234])
def test_hash_no_pad_lengths(glb, ctx, oracle, rng, length):
    x = rand_field(rng, (257, length))
    got = glb.PoseidonHash.hash_no_pad(x)
    want = oracle.hash_no_pad_batch(x) if length else np.tile(oracle.hash_no_pad(np.zeros(0, dtype=np.uint64)), (257, 1))
    assert np.array_equal(got, want)


def test_two_to_one_batch(glb, ctx, oracle, rng):
    l, r = rand_field(rng, (1000, 4)), rand_field(rng, (1000, 4))
    assert np.array_equal(glb.PoseidonHash.two_to_one(l, r), oracle.two_to_one_batch(l, r))


def test_hash_pad_and_smt_leaf_consistency(glb, ctx, oracle):
    """src/smt/goldilocks_poseidon/mod.rs:167-181 (hash_pad([k, v, 1])) == src/smt/gadgets/common.rs:87-101
    (hash_no_pad([k, v, 1, 1, 0, 1])), inputs of test_calc_node_hash (common.rs:67-71)."""
    k, v = oracle.from_u128(1), oracle.from_u128(2)
    a = glb.PoseidonHash.hash_pad(np.concatenate([k, v, [np.uint64(1)]]))
    b = glb.PoseidonHash.hash_no_pad(np.concatenate([k, v, np.array([1, 1, 0, 1], dtype=np.uint64)]))
    c = glb.PoseidonNodeHash.calc_leaf_hash(k, v)
    want = [9613647271972624781, 17898244898336278454, 17153022918269186278, 8190762674233093240]
    assert a.tolist() == want and b.tolist() == want and c.tolist() == want
    assert oracle.smt_leaf_hash(k, v).tolist() == want
    assert glb.PoseidonNodeHash.calc_internal_hash(k, v).tolist() == [
        17484264305055072364, 557184275190569619, 7427655570849746255, 2765125522432587977]


def test_device_resident_buffers(glb, ctx, oracle, rng):
    import torch

    x = rand_field(rng, (513, 12))
    t = torch.from_numpy(x.view(np.int64)).cuda()
    glb.PoseidonHash.permute(t)
    assert np.array_equal(t.cpu().numpy().view(np.uint64), oracle.permute_batch(x))


# ---- P4 MerkleTree::new --------------------------------------------------------------------------
@pytest.mark.parametrize("lg,leaf_len,cap_height", [(0, 7, 0), (1, 3, 0), (1, 3, 1), (3, 4, 1), (5, 5, 0), (5, 9, 5),
                                                    (6, 135, 4), (10, 32, 4), (9, 20, 0), (7, 1, 2)])
def test_merkle_tree_new(glb, ctx, oracle, rng, lg, leaf_len, cap_height):
    leaves = rand_field(rng, (1 << lg, leaf_len))
    t = glb.MerkleTree.new(leaves, cap_height)
    digests, cap = oracle.merkle_tree(leaves, cap_height)
    assert np.array_equal(t.cap, cap)
    assert np.array_equal(t.digests, digests)
    for i in {0, (1 << lg) - 1, (1 << lg) // 3}:
        sib = t.prove(i)
        assert np.array_equal(sib, oracle.merkle_prove(digests, 1 << lg, cap_height, i))
        assert oracle.merkle_verify(leaves[i], i, sib, cap, cap_height)


def test_merkle_tree_panics_like_upstream(glb, ctx, rng):
    with pytest.raises(glb.GlPanic):
        glb.MerkleTree.new(rand_field(rng, (8, 5)), 4)  # cap_height > log2(len)
    with pytest.raises(glb.GlPanic):
        glb.MerkleTree.new(rand_field(rng, (6, 5)), 1)  # log2_strict


# ---- P1/P2/P9 FFT family -------------------------------------------------------------------------
@pytest.mark.parametrize("lg", [0, 1, 2, 3, 5, 8, 10, 11, 13, 16])
def test_fft_family(glb, ctx, oracle, rng, lg):
    n = 1 << lg
    a = rand_field(rng, (3, n))
    if n >= 8:
        a[1] = adversarial_columns(n)[4]
    assert np.array_equal(glb.fft(a), np.stack([oracle.fft(r) for r in a]))
    assert np.array_equal(glb.ifft(a), np.stack([oracle.ifft(r) for r in a]))
    assert np.array_equal(glb.coset_fft(a, 7), np.stack([oracle.coset_fft(r, 7) for r in a]))
    assert np.array_equal(glb.coset_ifft(a, 7), np.stack([oracle.coset_ifft(r, 7) for r in a]))
    assert np.array_equal(glb.ifft(glb.fft(a)), a)


@pytest.mark.parametrize("lg", [9, 12, 14, 15, 17, 18, 19, 20])
def test_fft_fast_pass_sizes(glb, ctx, oracle, rng, lg):
    """Every split of the radix-16 register kernel (m = 8, 9, 10; strided and contiguous) and its mix with the
    generic pass: forward, inverse and coset variants against the oracle."""
    a = rand_field(rng, (2, 1 << lg))
    assert np.array_equal(glb.fft(a), np.stack([oracle.fft(r) for r in a]))
    assert np.array_equal(glb.coset_ifft(a, 7), np.stack([oracle.coset_ifft(r, 7) for r in a]))
    assert np.array_equal(glb.coset_fft(a, 7)[1], oracle.coset_fft(a[1], 7))


@pytest.mark.parametrize("lg", [16, 17, 18, 19, 20])
def test_fft_tma_path_extreme_inputs(glb, ctx, oracle, rng, lg):
    """The TMA two-pass kernels (csrc/ntt_tma.cu: carry-save butterflies, 2^16 .. 2^20 points): forward and inverse
    transforms of the adversarial columns (all p - 1, p - 1 - i, 32-bit limbs, bits) and of non-canonical inputs
    (2^64 - 1, p, p + i: taken mod p like every u64 the ABI accepts) against the oracle."""
    n = 1 << lg
    adv = adversarial_columns(n)
    r = np.arange(n, dtype=np.uint64)
    noncanon = np.stack([np.full(n, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64), np.uint64(P) + (r & np.uint64(0xFFFFFFFE)),
                         np.where(r % np.uint64(3) == 0, np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(P - 1))])
    a = np.concatenate([adv[1:], noncanon, rand_field(rng, (1, n))])
    red = (a.astype(object) % P).astype(np.uint64)
    assert np.array_equal(glb.fft(a), np.stack([oracle.fft(r) for r in red]))
    assert np.array_equal(glb.ifft(a), np.stack([oracle.ifft(r) for r in red]))


@pytest.mark.parametrize("lg_n,c,rate_bits,cap_height", [(16, 5, 3, 4), (17, 3, 2, 4), (18, 2, 1, 3), (16, 9, 0, 0)])
def test_commit_tma_path_matches_oracle(glb, ctx, oracle, lg_n, c, rate_bits, cap_height):
    """from_values at sizes that take the TMA LDE (several cosets per launch): coefficients, every leaf, every digest
    and the cap equal the oracle's."""
    n = 1 << lg_n
    values = oracle.synthetic_values(c, n, seed=lg_n)
    values[0] = adversarial_columns(n)[1]
    if c > 1:
        values[1] = adversarial_columns(n)[4]
    want = oracle.commit_from_values(values, rate_bits, cap_height)
    b = glb.PolynomialBatch.from_values(values, rate_bits, False, cap_height)
    assert np.array_equal(b.polynomials, want["coeffs"])
    assert np.array_equal(b.merkle_tree.cap, want["cap"])
    assert np.array_equal(b.merkle_tree.leaves, want["leaves"])
    assert np.array_equal(b.merkle_tree.digests, want["digests"])
    b.free()


@pytest.mark.parametrize("lg", [21, 22, 24])
def test_fft_three_pass_sizes(glb, ctx, oracle, rng, lg):
    a = rand_field(rng, (1, 1 << lg))
    assert np.array_equal(glb.coset_fft(a, 7)[0], oracle.coset_fft(a[0], 7))
    assert np.array_equal(glb.coset_ifft(glb.coset_fft(a, 7), 7), a)


# ---- P* PolynomialBatch --------------------------------------------------------------------------
COMMIT_CASES = [
    # lg_n, c, rate_bits, cap_height
    (0, 1, 0, 0), (0, 3, 3, 0), (1, 2, 1, 1), (2, 4, 3, 4), (3, 5, 3, 4), (4, 9, 2, 0), (5, 16, 3, 4), (6, 20, 3, 4),
    (8, 135, 3, 4), (10, 136, 3, 4), (11, 7, 3, 4), (12, 85, 3, 4), (7, 234, 3, 4), (9, 2, 1, 10), (12, 33, 0, 3),
]


@pytest.mark.parametrize("lg_n,c,rate_bits,cap_height", COMMIT_CASES)
def test_commit_from_values_matches_oracle(glb, ctx, oracle, lg_n, c, rate_bits, cap_height):
    n = 1 << lg_n
    values = oracle.synthetic_values(c, n)
    if n >= 8 and c >= 7:
        values[:5] = adversarial_columns(n)
    want = oracle.commit_from_values(values, rate_bits, cap_height)
    b = glb.PolynomialBatch.from_values(values, rate_bits, False, cap_height)
    assert np.array_equal(b.merkle_tree.cap, want["cap"])
    assert np.array_equal(b.polynomials, want["coeffs"])
    assert np.array_equal(b.merkle_tree.leaves, want["leaves"])
    assert np.array_equal(b.merkle_tree.digests, want["digests"])
    N = n << rate_bits
    idx = sorted({0, N - 1, N // 2, (N * 5) // 7})
    rows, paths = b.open(idx)
    for q, i in enumerate(idx):
        assert np.array_equal(rows[q], want["leaves"][i])
        assert np.array_equal(paths[q], oracle.merkle_prove(want["digests"], N, cap_height, i))
        assert oracle.merkle_verify(rows[q], i, paths[q], want["cap"], cap_height)
    # get_lde_values(i, step) = leaves[reverse_bits(i * step, lg N)]
    step = 1 << rate_bits
    for i in {0, n - 1, n // 2}:
        r = oracle.lib().glo_reverse_bits(i * step, lg_n + rate_bits)
        assert np.array_equal(b.get_lde_values(i, step), want["leaves"][r])
    b.free()


def test_commit_from_coeffs_matches_oracle(glb, ctx, oracle, rng):
    coeffs = rand_field(rng, (21, 512))
    want = oracle.commit_from_coeffs(coeffs, 3, 4)
    b = glb.PolynomialBatch.from_coeffs(coeffs, 3, False, 4)
    assert np.array_equal(b.merkle_tree.cap, want["cap"])
    assert np.array_equal(b.merkle_tree.leaves, want["leaves"])
    assert np.array_equal(b.merkle_tree.digests, want["digests"])


def test_commit_panics_like_upstream(glb, ctx, rng):
    with pytest.raises(glb.GlPanic):
        glb.PolynomialBatch.from_values(rand_field(rng, (3, 12)), 3, False, 4)  # log2_strict
    with pytest.raises(glb.GlPanic):
        glb.PolynomialBatch.from_values(rand_field(rng, (3, 2)), 1, False, 4)  # cap_height > log2(N)


@pytest.mark.parametrize("count", [2, 4, 8])
def test_sharded_commit_equals_whole(glb, oracle, count):
    """shard s computes leaf block s = whole top-level subtrees; caps concatenate."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    values = oracle.synthetic_values(19, 256)
    want = oracle.commit_from_values(values, 3, 4)
    N = 256 << 3
    cap = np.zeros((16, 4), dtype=np.uint64)
    for s in range(count):
        c = glb.Context(0)
        c.set_shard(s, count)
        b = glb.PolynomialBatch.from_values(values, 3, False, 4, ctx=c)
        lo, hi = b.leaf_begin, b.leaf_end
        assert (lo, hi) == (s * N // count, (s + 1) * N // count)
        assert np.array_equal(b.merkle_tree.leaves, want["leaves"][lo:hi])
        per = 16 // count
        cap[s * per:(s + 1) * per] = b.merkle_tree.cap[s * per:(s + 1) * per]
        rows, paths = b.open([lo, hi - 1])
        assert oracle.merkle_verify(rows[1], hi - 1, paths[1], want["cap"], 4)
        with pytest.raises(glb.GlPanic):
            b.open([hi % N if count > 1 else N])
        b.free()
        c.close()
    assert np.array_equal(cap, want["cap"])


def test_lde_linearity_and_low_degree_at_scale(glb, ctx, oracle):
    """Size-independent properties at a size the oracle does not run in seconds: commit(a)+commit(b) rows
    = commit(a+b) rows, and every opened path verifies against the cap."""
    n, c = 1 << 16, 24
    a = oracle.synthetic_values(c, n, seed=1)
    b = oracle.synthetic_values(c, n, seed=2)
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    ba = glb.PolynomialBatch.from_values(a, 3, False, 4)
    bb = glb.PolynomialBatch.from_values(b, 3, False, 4)
    bs = glb.PolynomialBatch.from_values(s, 3, False, 4)
    idx = [0, 1, 12345, (n << 3) - 1, 77777]
    ra, pa = ba.open(idx)
    rb, _ = bb.open(idx)
    rs, _ = bs.open(idx)
    assert np.array_equal(((ra.astype(object) + rb.astype(object)) % P).astype(np.uint64), rs)
    for q, i in enumerate(idx):
        assert oracle.merkle_verify(ra[q], i, pa[q], ba.merkle_tree.cap, 4)
    # the n-coset restriction: leaves of block 0 are the values on 7<w_n> in bit-reversed order
    coef = ba.polynomials
    lde0 = ba.get_lde_values(np.arange(4), 8)
    col = oracle.coset_fft(coef[3], 7)
    assert [int(lde0[i][3]) for i in range(4)] == [int(col[i]) for i in range(4)]


# ---- P8/P10 FRI pieces ---------------------------------------------------------------------------
@pytest.mark.parametrize("lg,arity_bits,cap_height", [(4, 4, 0), (8, 4, 4), (12, 4, 4), (7, 3, 2), (10, 1, 0)])
def test_fri_layer_tree(glb, ctx, oracle, rng, lg, arity_bits, cap_height):
    v = rand_field(rng, (1 << lg, 2))
    _, digests, cap = oracle.fri_layer_tree(v, arity_bits, cap_height)
    d, c = glb.fri_layer_tree(v, arity_bits, cap_height)
    assert np.array_equal(c, cap) and np.array_equal(d, digests)


@pytest.mark.parametrize("lg,arity_bits", [(4, 4), (8, 4), (13, 4), (9, 3)])
def test_fri_fold(glb, ctx, oracle, rng, lg, arity_bits):
    cf = rand_field(rng, (1 << lg, 2))
    beta = rand_field(rng, (2,))
    shift = pow(7, 1 << arity_bits, P)
    folded, nxt = glb.fri_fold(cf, arity_bits, beta, shift)
    want = oracle.fri_fold(cf, arity_bits, beta)
    assert np.array_equal(folded, want)
    assert np.array_equal(nxt, oracle.ext_coset_fft(want, shift))


def test_pow_grind_smallest_witness(glb, ctx, oracle, rng):
    state = rand_field(rng, (12,))
    for bits in (4, 10, 16):
        w = glb.fri_proof_of_work(state, 0, bits)
        assert w == oracle.pow_grind(state, 0, bits, 0, 1 << 20)
        s = state.copy()
        s[0] = w
        assert int(oracle.permute(s)[7]) >> (64 - bits) == 0


# ---- P7 SparseMerkleProcessProof::check over batches ---------------------------------------------
def _smt_proofs(oracle, rng, n_ops=60):
    t = oracle.Smt()
    recs = []
    keys = [rand_field(rng, (4,)) for _ in range(n_ops // 2)]
    # inserts, updates, deletes (value 0 removes: src/smt/tree.rs:143-155), and noops
    for k in keys:
        recs.append(t.set(k, rand_field(rng, (4,))))
    for k in keys[::3]:
        recs.append(t.set(k, rand_field(rng, (4,))))
    for k in keys[1::4]:
        recs.append(t.set(k, np.zeros(4, dtype=np.uint64)))
    recs.append(t.set(rand_field(rng, (4,)), np.zeros(4, dtype=np.uint64)))
    return np.array(recs, dtype=oracle.SMT_PROOF_DTYPE)


def _pack(glb, recs):
    hd = np.zeros(recs.shape[0], dtype=glb.host.SMT_HDR_DTYPE)
    for f in ("old_root", "old_key", "old_value", "new_root", "new_key", "new_value", "is_old0", "fnc"):
        hd[f] = recs[f]
    off = np.zeros(recs.shape[0] + 1, dtype=np.uint64)
    off[1:] = np.cumsum(recs["num_siblings"])
    pool = np.concatenate([r["siblings"][: r["num_siblings"]] for r in recs] + [np.zeros((0, 4), dtype=np.uint64)])
    return hd, pool, off


def test_smt_process_proofs(glb, ctx, oracle, rng):
    recs = _smt_proofs(oracle, rng)
    assert set(recs["fnc"].tolist()) == {0, 1, 2, 3}
    want = oracle.smt_verify_process_batch(recs)
    assert (want == 0).all()
    hd, pool, off = _pack(glb, recs)
    assert np.array_equal(glb.smt_check_process_proofs(hd, pool, off), want)
    # corrupt: wrong new_root, wrong sibling, wrong key on update -> same assert ordinal as the oracle
    bad = recs.copy()
    bad["new_root"][0][0] ^= np.uint64(1)
    bad["old_value"][3][1] ^= np.uint64(5)
    for i in range(5, len(bad), 7):
        if bad["num_siblings"][i]:
            bad["siblings"][i][0][2] ^= np.uint64(9)
    want = oracle.smt_verify_process_batch(bad)
    assert (want != 0).any()
    hd, pool, off = _pack(glb, bad)
    assert np.array_equal(glb.smt_check_process_proofs(hd, pool, off), want)


def test_smt_process_proofs_take_header_words_mod_p(glb, ctx, oracle, rng):
    """include/gl_b200.h: inputs may be any u64 (taken mod p).  The reference reads key bits through HashOut::to_bytes
    (canonical, src/smt/proof/process.rs:193-203) and compares roots as field elements, so x and x + p are the same
    key / root / value / sibling: every word that fits gets p added and the statuses must not change."""
    recs = _smt_proofs(oracle, rng)
    want = oracle.smt_verify_process_batch(recs)
    nc = recs.copy()
    small = np.uint64(0xFFFFFFFF)          # x + p < 2^64  <=>  x < 2^32 - 1
    bumped = 0
    for name in ("old_root", "old_key", "old_value", "new_root", "new_key", "new_value", "siblings"):
        a = nc[name]
        m = a < small
        a[m] += np.uint64(P)
        bumped += int(m.sum())
    # from_u128-style keys (u32 limbs) guarantee non-canonical words even when random 64-bit words rarely qualify
    assert np.array_equal(oracle.smt_verify_process_batch(nc), want)
    hd, pool, off = _pack(glb, nc)
    assert np.array_equal(glb.smt_check_process_proofs(hd, pool, off), want)
    # small keys: every limb < 2^32, so every key / value word is bumped
    t = oracle.Smt()
    rs = [t.set(oracle.from_u128(k), oracle.from_u128(v)) for k, v in [(1, 2), (12, 1), (5, 51), (5, 7), (12, 0), (99, 0)]]
    recs2 = np.array(rs, dtype=oracle.SMT_PROOF_DTYPE)
    want2 = oracle.smt_verify_process_batch(recs2)
    assert (want2 == 0).all()
    nc2 = recs2.copy()
    for name in ("old_key", "old_value", "new_key", "new_value"):
        a = nc2[name]
        m = a < small
        a[m] += np.uint64(P)
        bumped += int(m.sum())
    assert bumped > 50
    hd, pool, off = _pack(glb, nc2)
    assert np.array_equal(glb.smt_check_process_proofs(hd, pool, off), want2)
    assert np.array_equal(oracle.smt_verify_process_batch(nc2), want2)


def test_smt_process_proofs_with_long_common_prefixes(glb, ctx, oracle, rng):
    """Keys that agree on 3 .. 255 leading path bits: the inserts push the old leaf down a chain of Bottom levels
    before NewOne.  The kernel computes only the hashes the verifier reads (state Top / Bottom / NewOne levels);
    statuses must equal the oracle's full 2 x 256-level walk, for valid proofs and for corrupted ones."""
    t = oracle.Smt()
    recs = []
    base = [rand_field(rng, (4,)) for _ in range(12)]
    for k in base:
        recs.append(t.set(k, rand_field(rng, (4,))))
    for j, k in enumerate(base):
        bit = (3, 17, 63, 64, 70, 127, 128, 200, 254, 255, 31, 191)[j]
        k2 = k.copy()
        k2[bit >> 6] ^= np.uint64(1) << np.uint64(bit & 63)
        k2 %= np.uint64(P)
        recs.append(t.set(k2, rand_field(rng, (4,))))          # insert next to its twin
    for k in base[::2]:
        recs.append(t.set(k, np.zeros(4, dtype=np.uint64)))    # delete: the twin climbs back up
    for k in base[1::2]:
        recs.append(t.set(k, rand_field(rng, (4,))))           # update deep in the tree
    recs = np.array(recs, dtype=oracle.SMT_PROOF_DTYPE)
    assert recs["num_siblings"].max() == 256       # the twin at bit 255: upstream's assert!(siblings.len() < 256) fires
    want = oracle.smt_verify_process_batch(recs)
    assert (want == 0).sum() == len(recs) - 1 and (want == 1).sum() == 1
    assert np.array_equal(glb.smt_check_process_proofs(*_pack(glb, recs)), want)
    bad = recs.copy()
    for i in range(len(bad)):
        which = i % 5
        if which == 0:
            bad["new_root"][i][i % 4] ^= np.uint64(1)
        elif which == 1:
            bad["old_value"][i][0] ^= np.uint64(2)
        elif which == 2 and bad["num_siblings"][i]:
            bad["siblings"][i][bad["num_siblings"][i] - 1][1] ^= np.uint64(4)
        elif which == 3:
            bad["new_key"][i][3] ^= np.uint64(1) << np.uint64(62)
        else:
            bad["is_old0"][i] ^= 1
    want = oracle.smt_verify_process_batch(bad)
    assert (want != 0).sum() > len(bad) // 2
    assert np.array_equal(glb.smt_check_process_proofs(*_pack(glb, bad)), want)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_smt_process_proof_batch_shares_equal_the_whole_batch(glb, ctx, oracle, rng, world):
    """a batch of process proofs is split evenly over the ranks with no exchange; every share, verified on
    its own (sibling offsets rebased by parallel.slice_proof_batch), gives the statuses of its slice of the batch."""
    import importlib

    par = importlib.import_module("plonky2-lib_b200.parallel")
    recs = _smt_proofs(oracle, rng)
    recs["new_root"][7][0] ^= np.uint64(1)
    recs["old_root"][20][1] ^= np.uint64(4)
    hd, pool, off = _pack(glb, recs)
    whole = glb.smt_check_process_proofs(hd, pool, off)
    assert np.array_equal(whole, oracle.smt_verify_process_batch(recs)) and (whole != 0).sum() >= 2
    got = []
    for rank in range(world):
        lo, hi = par.batch_range(rank, world, len(recs))
        got.append(glb.smt_check_process_proofs(*par.slice_proof_batch(hd, pool, off, lo, hi)))
    assert np.array_equal(np.concatenate(got), whole)


def test_smt_fixture_root_three_inserts(glb, ctx, oracle):
    """(1->2), (12->1), (5->51) as in src/smt/gadgets/verify/mod.rs:24-34."""
    t = oracle.Smt()
    recs = [t.set(oracle.from_u128(k), oracle.from_u128(v)) for k, v in [(1, 2), (12, 1), (5, 51)]]
    assert t.root().tolist() == [16994558480514381166, 8559105504417206749, 13458782878755336329, 17099432696459526118]
    recs = np.array(recs, dtype=oracle.SMT_PROOF_DTYPE)
    hd, pool, off = _pack(glb, recs)
    assert (glb.smt_check_process_proofs(hd, pool, off) == 0).all()


def test_golden_commit_caps(glb, ctx, oracle):
    """Committed fixtures (tests/golden/commit_caps.json, made by tests/golden/make_golden.py)."""
    import json
    import os

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "commit_caps.json")))
    for case in g["cases"]:
        v = oracle.synthetic_values(case["c"], 1 << case["lg_n"])
        b = glb.PolynomialBatch.from_values(v, case["rate_bits"], False, case["cap_height"])
        assert [f"{int(x):016x}" for x in b.merkle_tree.cap.reshape(-1)] == case["cap"]
        b.free()


def _horner(coeffs, x):
    acc = 0
    for c in coeffs[::-1].tolist():
        acc = (acc * x + c) % P
    return acc


def test_config2_full_size_properties(glb, ctx, oracle):
    """BASELINE.json configs[1] at full size (2^20 rows x 135 columns, rate_bits 3, cap_height 4), checked through
    size-independent properties: coefficients interpolate the values, opened LDE rows are the polynomial
    evaluated at 7 * w_N^bitrev(i), every opened Merkle path verifies against the cap, and the sub-sampled
    coset (get_lde_values) agrees with the opened rows."""
    lg_n, c, r, h = 20, 135, 3, 4
    n, N = 1 << lg_n, 1 << (lg_n + r)
    values = oracle.synthetic_values(c, n)
    b = glb.PolynomialBatch.from_values(values, r, False, h)
    coeffs = b.polynomials
    w_n = oracle.lib().glo_primitive_root_of_unity(lg_n)
    w_N = oracle.lib().glo_primitive_root_of_unity(lg_n + r)
    rng = np.random.default_rng(7)
    for col in (0, 134):
        j = int(rng.integers(0, n))
        assert _horner(coeffs[col], pow(w_n, j, P)) == int(values[col][j])
    idx = sorted(set(int(x) for x in rng.integers(0, N, size=6)) | {0, N - 1})
    rows, paths = b.open(idx)
    cap = b.merkle_tree.cap
    for q, i in enumerate(idx):
        assert oracle.merkle_verify(rows[q], i, paths[q], cap, h)
    for q in (0, len(idx) // 2):
        i = idx[q]
        x = 7 * pow(w_N, oracle.lib().glo_reverse_bits(i, lg_n + r), P) % P
        for col in (3, 77):
            assert _horner(coeffs[col], x) == int(rows[q][col])
    i = 123457
    leaf = oracle.lib().glo_reverse_bits(i * 8, lg_n + r)
    assert np.array_equal(b.get_lde_values(i, 8), b.open([leaf])[0][0])
    b.free()
    ctx.trim()


def test_streaming_commit_equals_from_coeffs(glb, ctx, oracle, rng):
    """gl_commit_begin / add_coeffs / finish (from_coeffs fed column block by column block, in any order)."""
    import ctypes as C

    lib, N = ctx._lib, glb._native
    c, lg = 21, 9
    coeffs = rand_field(rng, (c, 1 << lg))
    want = oracle.commit_from_coeffs(coeffs, 3, 4)
    h = C.c_void_p()
    ctx.check(lib.gl_commit_begin(ctx._h, lg, c, 3, 4, C.byref(h)))
    cap = np.zeros((16, 4), dtype=np.uint64)
    assert lib.gl_commit_finish(h, cap.ctypes.data, N.GL_HOST) == N.GL_E_STATE          # columns missing
    rows = np.zeros((1, c), dtype=np.uint64)
    idx = np.zeros(1, dtype=np.uint64)
    assert lib.gl_commit_open(h, idx.ctypes.data, 1, rows.ctypes.data, None, N.GL_HOST) == N.GL_E_STATE
    for col0, nc in [(16, 5), (0, 8), (8, 8)]:
        blk = np.ascontiguousarray(coeffs[col0:col0 + nc])
        ctx.check(lib.gl_commit_add_coeffs(h, col0, nc, blk.ctypes.data, N.GL_HOST))
    assert lib.gl_commit_add_coeffs(h, 20, 2, coeffs.ctypes.data, N.GL_HOST) == N.GL_E_ARG
    ctx.check(lib.gl_commit_finish(h, cap.ctypes.data, N.GL_HOST))
    assert np.array_equal(cap, want["cap"])
    leaves = np.zeros(((1 << lg) << 3, c), dtype=np.uint64)
    ctx.check(lib.gl_commit_download(h, leaves.ctypes.data, None, N.GL_HOST))
    assert np.array_equal(leaves, want["leaves"])
    lib.gl_commit_free(h)


STREAM_HASH_CASES = [
    # c, lg, block sizes fed in order (None = out of order: falls back to hashing at finish)
    (21, 9, [8, 8, 5]), (21, 9, [5, 5, 5, 6]), (21, 9, [21]), (21, 9, [1] * 21), (135, 8, [28, 28, 28, 28, 23]),
    (135, 8, [32, 32, 32, 32, 7]), (16, 7, [3, 13]), (9, 6, [8, 1]), (8, 6, [4, 4]), (5, 6, [2, 3]), (4, 6, [2, 2]), (1, 5, [1]),
    (21, 9, None),
]


@pytest.mark.parametrize("c,lg,blocks", STREAM_HASH_CASES)
def test_streamed_leaf_hashing_gives_the_same_commit(glb, ctx, oracle, rng, c, lg, blocks):
    """GL_COMMIT_STREAM_HASH: the sponge absorbs every complete group of 8 polynomials as soon as its block has been
    extended (partial groups wait for the next block; the last block closes the ragged group).  Digests, cap and
    openings must equal the one-shot commit for every way of cutting the batch into blocks."""
    import ctypes as C

    lib, N = ctx._lib, glb._native
    coeffs = rand_field(rng, (c, 1 << lg))
    want = oracle.commit_from_coeffs(coeffs, 3, 4)
    h = C.c_void_p()
    ctx.check(lib.gl_commit_begin_ex(ctx._h, lg, c, 3, 4, N.GL_COMMIT_STREAM_HASH, C.byref(h)))
    if blocks is None:
        order = [(16, 5), (0, 8), (8, 8)]
    else:
        order, at = [], 0
        for nc in blocks:
            order.append((at, nc))
            at += nc
        assert at == c
    for col0, nc in order:
        blk = np.ascontiguousarray(coeffs[col0:col0 + nc])
        ctx.check(lib.gl_commit_add_coeffs(h, col0, nc, blk.ctypes.data, N.GL_HOST))
    cap = np.zeros((16, 4), dtype=np.uint64)
    ctx.check(lib.gl_commit_finish(h, cap.ctypes.data, N.GL_HOST))
    assert np.array_equal(cap, want["cap"])
    NL = (1 << lg) << 3
    digests = np.zeros((2 * (NL - 16), 4), dtype=np.uint64)
    ctx.check(lib.gl_commit_download(h, None, digests.ctypes.data, N.GL_HOST))
    assert np.array_equal(digests, want["digests"])
    lib.gl_commit_free(h)
    assert lib.gl_commit_begin_ex(ctx._h, lg, c, 3, 4, 0x80, C.byref(h)) == N.GL_E_ARG   # unknown flag


def test_streamed_leaf_hashing_sharded(glb, oracle, rng):
    """The same on a shard (one coset block = top-level subtrees of one rank)."""
    import ctypes as C

    N = glb._native
    c, lg = 19, 8
    coeffs = rand_field(rng, (c, 1 << lg))
    want = oracle.commit_from_coeffs(coeffs, 3, 4)
    cap = np.zeros((16, 4), dtype=np.uint64)
    for rank in range(4):
        cx = glb.Context(0)
        cx.set_shard(rank, 4)
        h = C.c_void_p()
        cx.check(cx._lib.gl_commit_begin_ex(cx._h, lg, c, 3, 4, N.GL_COMMIT_STREAM_HASH, C.byref(h)))
        for col0, nc in [(0, 7), (7, 7), (14, 5)]:
            blk = np.ascontiguousarray(coeffs[col0:col0 + nc])
            cx.check(cx._lib.gl_commit_add_coeffs(h, col0, nc, blk.ctypes.data, N.GL_HOST))
        cx.check(cx._lib.gl_commit_finish(h, cap.ctypes.data, N.GL_HOST))    # writes its 4 entries at their global index
        cx._lib.gl_commit_free(h)
        cx.close()
    assert np.array_equal(cap, want["cap"])


def test_zkdsa_native_batch_and_public_input_json(glb, ctx, oracle, rng):
    """src/zkdsa: public keys, addresses and signatures over a batch, and the JSON the reference pins for the default
    (all-zero) signature (src/zkdsa/circuits/mod.rs:77-106, 136-153)."""
    host = glb.host
    z = np.zeros((1, 4), dtype=np.uint64)
    pk, sig = host.zkdsa_public_keys(z), host.zkdsa_sign(z, z)
    assert pk[0].tolist() == KAT_TWO_TO_ONE_ZERO and sig[0].tolist() == KAT_TWO_TO_ONE_ZERO
    want = ('{"message":"0x0000000000000000000000000000000000000000000000000000000000000000",'
            '"public_key":"0xc71603f33a1144ca7953db0ab48808f4c4055e3364a246c33c18a9786cb0b359",'
            '"signature":"0xc71603f33a1144ca7953db0ab48808f4c4055e3364a246c33c18a9786cb0b359"}')
    assert host.zkdsa_public_inputs_json(z[0], pk[0], sig[0]) == want
    sk, msg = rand_field(rng, (300, 4)), rand_field(rng, (300, 4))
    pks, sigs = host.zkdsa_public_keys(sk), host.zkdsa_sign(sk, msg)
    for i in (0, 7, 299):
        assert pks[i].tolist() == oracle.two_to_one(sk[i], sk[i]).tolist()
        assert sigs[i].tolist() == oracle.two_to_one(sk[i], msg[i]).tolist()
    assert host.zkdsa_addresses(pks).tolist() == pks[:, 0].tolist()


@pytest.mark.parametrize("lg_n,c,rate_bits,cap_height", [(6, 20, 3, 4), (9, 135, 3, 4), (5, 3, 1, 0), (4, 9, 2, 6), (7, 4, 3, 2)])
def test_merkle_verify_batch_accepts_openings_and_rejects_tampering(glb, ctx, oracle, rng, lg_n, c, rate_bits, cap_height):
    """verify_merkle_proof_to_cap over a batch: everything gl_commit_open returns verifies against the cap; a flipped bit in
    a leaf, a sibling, the index or the cap entry is caught -- proof by proof as the oracle's verifier decides."""
    n = 1 << lg_n
    N = n << rate_bits
    values = oracle.synthetic_values(c, n)
    b = glb.PolynomialBatch.from_values(values, rate_bits, False, cap_height)
    cap = b.merkle_tree.cap.copy()
    idx = np.unique(rng.integers(0, N, 64)).astype(np.uint64)
    rows, paths = b.open(idx)
    b.free()
    assert glb.host.merkle_verify_batch(rows, idx, paths, cap).all()
    bad_rows, bad_idx, bad_paths = rows.copy(), idx.copy(), paths.copy()
    k = len(idx)
    bad_rows[0, c - 1] ^= np.uint64(1)
    if paths.shape[1]:
        bad_paths[1, paths.shape[1] - 1, 2] ^= np.uint64(1 << 40)
        bad_paths[2, 0, 0] ^= np.uint64(1)
    bad_idx[3] ^= np.uint64(1)
    got = glb.host.merkle_verify_batch(bad_rows, bad_idx, bad_paths, cap)
    want = np.array([oracle.merkle_verify(bad_rows[i], int(bad_idx[i]), bad_paths[i], cap, cap_height) for i in range(k)])
    assert np.array_equal(got, want) and not got[0] and got[4:].all()
    cap2 = cap.copy()
    cap2[int(idx[5]) >> paths.shape[1], 3] ^= np.uint64(1)
    got = glb.host.merkle_verify_batch(rows, idx, paths, cap2)
    assert not got[5] and np.array_equal(got, np.array([oracle.merkle_verify(rows[i], int(idx[i]), paths[i], cap2, cap_height) for i in range(k)]))


# ---- blinding = true (upstream signature; never set by the reference's configs) -------------------
@pytest.mark.parametrize("lg_n,c,rate_bits,cap_height", [(6, 5, 3, 4), (10, 20, 3, 4), (3, 2, 1, 0), (16, 9, 3, 4)])
def test_commit_with_blinding(glb, ctx, oracle, lg_n, c, rate_bits, cap_height):
    """PolynomialBatch::from_values(.., blinding = true, ..): every leaf = the c LDE values + SALT_SIZE = 4 uniform field
    elements (lde_values' `.chain(F::rand_vec)`); the tree is the oracle's MerkleTree::new over those leaves; openings carry
    the salt and verify; get_lde_values strips it; the salt is reproducible under a seed and fresh otherwise."""
    n, N = 1 << lg_n, (1 << lg_n) << rate_bits
    values = oracle.synthetic_values(c, n, seed=7)
    plain = glb.PolynomialBatch.from_values(values, rate_bits, False, cap_height)
    ctx.set_salt_seed(1234)
    b = glb.PolynomialBatch.from_values(values, rate_bits, True, cap_height)
    assert b.blinding and b.leaf_len == c + 4 and b.num_columns == c
    assert np.array_equal(b.polynomials, plain.polynomials)
    leaves = b.merkle_tree.leaves
    assert leaves.shape == (N, c + 4)
    assert np.array_equal(leaves[:, :c], plain.merkle_tree.leaves)
    salt = leaves[:, c:]
    assert (salt < np.uint64(P)).all()
    assert len(np.unique(salt)) > 0.99 * salt.size or salt.size < 64          # uniform 64-bit draws do not repeat
    assert 0.3 < (salt >> np.uint64(63)).mean() < 0.7
    want_digests, want_cap = oracle.merkle_tree(leaves, cap_height)
    assert np.array_equal(b.merkle_tree.digests, want_digests)
    assert np.array_equal(b.merkle_tree.cap, want_cap)
    assert not np.array_equal(b.merkle_tree.cap, plain.merkle_tree.cap)
    idx = sorted({0, N - 1, N // 3})
    rows, paths = b.open(idx)
    for q, i in enumerate(idx):
        assert np.array_equal(rows[q], leaves[i])
        assert oracle.merkle_verify(rows[q], i, paths[q], b.merkle_tree.cap, cap_height)
    step = 1 << rate_bits
    got = b.get_lde_values(n // 2, step)
    assert got.shape == (c,) and np.array_equal(got, plain.get_lde_values(n // 2, step))
    # same seed -> same salt; the default seed (OS entropy, then a counter) -> different salt on every commit
    ctx.set_salt_seed(1234)
    b2 = glb.PolynomialBatch.from_coeffs(np.ascontiguousarray(b.polynomials), rate_bits, True, cap_height)
    assert np.array_equal(b2.merkle_tree.cap, b.merkle_tree.cap)
    b3 = glb.PolynomialBatch.from_values([v.copy() for v in values], rate_bits, True, cap_height)   # one array per polynomial
    assert not np.array_equal(b3.merkle_tree.cap, b.merkle_tree.cap) and b3.leaf_len == c + 4
    for x in (plain, b, b2, b3):
        x.free()
