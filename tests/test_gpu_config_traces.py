"""BASELINE.json configs 1, 3, 4, 5 as parity cases: there is no Rust toolchain, so each circuit is replayed as
its *commit trace* (SURVEY.md 8d): the PolynomialBatch commits and FRI layer trees a `data.prove(pw)` of that
circuit issues, on seeded synthetic polynomials of the same shapes.  Sizes the oracle finishes in seconds are
compared bit for bit; the larger ones go through size-independent properties."""
import numpy as np
import pytest

from conftest import P, rand_field

pytestmark = pytest.mark.gpu

RATE_BITS, CAP_HEIGHT = 3, 4   # CircuitConfig::standard_recursion_config().fri_config


def _commit_and_check(glb, oracle, c, lg_n, seed, full=True):
    v = oracle.synthetic_values(c, 1 << lg_n, seed=seed)
    b = glb.PolynomialBatch.from_values(v, RATE_BITS, False, CAP_HEIGHT)
    N = 1 << (lg_n + RATE_BITS)
    idx = [0, N // 3, N - 1]
    rows, paths = b.open(idx)
    for q, i in enumerate(idx):
        assert oracle.merkle_verify(rows[q], i, paths[q], b.merkle_tree.cap, CAP_HEIGHT)
    if full:
        want = oracle.commit_from_values(v, RATE_BITS, CAP_HEIGHT, want_leaves=False)
        assert np.array_equal(b.merkle_tree.cap, want["cap"])
        assert np.array_equal(b.polynomials, want["coeffs"])
    b.free()
    return b


def _fri_layers(glb, oracle, rng, degree_bits):
    """fri_committed_trees for ConstantArityBits(4, 5): every layer tree, fold and re-evaluation vs the oracle."""
    cfg = glb.FriConfig()
    arities = cfg.reduction_strategy.reduction_arity_bits(degree_bits, cfg.rate_bits, cfg.cap_height)
    lg = degree_bits + cfg.rate_bits
    coeffs = np.zeros((1 << lg, 2), dtype=np.uint64)
    coeffs[: 1 << degree_bits] = rand_field(rng, (1 << degree_bits, 2))   # final_poly.lde(rate_bits)
    shift = 7
    values = oracle.ext_coset_fft(coeffs, shift)
    for ab in arities:
        _, digests, cap = oracle.fri_layer_tree(values, ab, cfg.cap_height)
        d, c = glb.fri_layer_tree(values, ab, cfg.cap_height)
        assert np.array_equal(c, cap) and np.array_equal(d, digests)
        beta = rand_field(rng, (2,))   # stands in for challenger.get_extension_challenge()
        shift = pow(shift, 1 << ab, P)
        folded, nxt = glb.fri_fold(coeffs, ab, beta, shift)
        want = oracle.fri_fold(coeffs, ab, beta)
        assert np.array_equal(folded, want)
        coeffs = want
        values = oracle.ext_coset_fft(coeffs, shift)
        assert np.array_equal(nxt, values)
    return len(arities)


def test_config1_single_ecdsa_trace(glb, ctx, oracle, rng):
    """configs[0]: one secp256k1 ECDSA verify, ~2^16 rows: constants+sigmas (build), wires, Z/partial products,
    quotient chunks, then the FRI layers (degree_bits 16 -> arities [4, 4, 4])."""
    for c, seed in [(84, 11), (135, 12), (20, 13), (16, 14)]:
        _commit_and_check(glb, oracle, c, 16, seed)
    assert _fri_layers(glb, oracle, rng, 16) == 3


def test_config3_keccak_trace_properties(glb, ctx, oracle):
    """configs[2]: Keccak-256 over 64 blocks, ~2^18 rows: the wires commit at full width, checked by properties
    (the oracle needs ~20 s for it)."""
    _commit_and_check(glb, oracle, 135, 18, 31, full=False)
    _commit_and_check(glb, oracle, 20, 18, 32, full=False)
    ctx.trim()


def test_config4_smt_batch(glb, ctx, oracle, rng):
    """configs[3]: batches of native SMT updates verified with SparseMerkleProcessProof::check
    (src/smt/proof/process.rs:47-51): 2048 proofs against the oracle, status by status."""
    t = oracle.Smt()
    recs = []
    keys = [rand_field(rng, (4,)) for _ in range(1024)]
    for k in keys:
        recs.append(t.set(k, rand_field(rng, (4,))))
    for k in keys[:512]:
        recs.append(t.set(k, rand_field(rng, (4,))))
    for k in keys[512:1024]:
        recs.append(t.set(k, np.zeros(4, dtype=np.uint64)))
    recs = np.array(recs, dtype=oracle.SMT_PROOF_DTYPE)
    recs["new_root"][17][2] ^= np.uint64(1)
    recs["siblings"][900][0][0] ^= np.uint64(1)
    want = oracle.smt_verify_process_batch(recs)
    hd = np.zeros(recs.shape[0], dtype=glb.host.SMT_HDR_DTYPE)
    for f in ("old_root", "old_key", "old_value", "new_root", "new_key", "new_value", "is_old0", "fnc"):
        hd[f] = recs[f]
    off = np.zeros(recs.shape[0] + 1, dtype=np.uint64)
    off[1:] = np.cumsum(recs["num_siblings"])
    pool = np.concatenate([r["siblings"][: r["num_siblings"]] for r in recs])
    got = glb.smt_check_process_proofs(hd, pool, off)
    assert np.array_equal(got, want)
    assert (got != 0).sum() >= 2 and (got == 0).sum() >= 2000
    # the 256-proof membership circuit of the same config: ~2^15 rows
    _commit_and_check(glb, oracle, 135, 15, 41)


def test_config5_aggregation_trace(glb, ctx, oracle, rng):
    """configs[4]: 8 inner proofs are 8 independent provers (one per GPU: replicas, no exchange) and one outer
    recursion circuit (~2^13 rows) sharded by coset.  Here: distinct seeds give distinct caps, and the outer
    commit sharded over 8 contexts reassembles the single-context cap."""
    caps = set()
    for inner in range(3):
        v = oracle.synthetic_values(135, 1 << 12, seed=100 + inner)
        b = glb.PolynomialBatch.from_values(v, RATE_BITS, False, CAP_HEIGHT)
        caps.add(b.merkle_tree.cap.tobytes())
        b.free()
    assert len(caps) == 3
    v = oracle.synthetic_values(135, 1 << 13, seed=200)
    want = oracle.commit_from_values(v, RATE_BITS, CAP_HEIGHT, want_leaves=False)
    cap = np.zeros((16, 4), dtype=np.uint64)
    for s in range(8):
        c = glb.Context(0)
        c.set_shard(s, 8)
        b = glb.PolynomialBatch.from_values(v, RATE_BITS, False, CAP_HEIGHT, ctx=c)
        cap[2 * s: 2 * s + 2] = b.merkle_tree.cap[2 * s: 2 * s + 2]
        b.free()
        c.close()
    assert np.array_equal(cap, want["cap"])
    assert _fri_layers(glb, oracle, rng, 13) == 2


def test_non_canonical_inputs_are_taken_mod_p(glb, ctx, oracle, rng):
    """include/gl_b200.h: inputs may be any u64 (upstream keeps non-canonical representatives internally)."""
    v = oracle.synthetic_values(9, 1 << 8)
    v[:, ::3] %= np.uint64(1 << 31)                       # small values have a second representative below 2^64
    nc = v.copy()
    nc[:, ::3] += np.uint64(P)
    a = glb.PolynomialBatch.from_values(v, 3, False, 4)
    b = glb.PolynomialBatch.from_values(nc, 3, False, 4)
    assert np.array_equal(a.merkle_tree.cap, b.merkle_tree.cap)
    assert np.array_equal(a.polynomials, b.polynomials)
    x = rand_field(rng, (33, 12)) % np.uint64(1 << 30)
    assert np.array_equal(glb.PoseidonHash.permute(x), glb.PoseidonHash.permute(x + np.uint64(P)))
    assert np.array_equal(glb.PoseidonHash.hash_no_pad(x[:, :3]), glb.PoseidonHash.hash_no_pad(x[:, :3]))
    l = x[:, :4] + np.uint64(P)
    assert np.array_equal(glb.MerkleTree.new(l[:32], 2).cap, oracle.merkle_tree(x[:32, :4], 2)[1])
