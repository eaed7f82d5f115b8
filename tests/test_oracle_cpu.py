"""Pins the CPU oracle (oracle/gl_oracle.c) to the reference's own known-answer vectors and to the
self-consistency properties SURVEY.md 8c lists for the rows the reference does not pin."""
import json
import os

import numpy as np
import pytest

from conftest import P, rand_field

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reference_poseidon_kat(oracle):
    # src/zkdsa/circuits/mod.rs:85-105 test_default_simple_signature
    z = np.zeros(4, dtype=np.uint64)
    assert oracle.two_to_one(z, z).tolist() == [4330397376401421145, 14124799381142128323, 8742572140681234676, 14345658006221440202]
    # src/zkdsa/circuits/mod.rs:136-153: the same digest as the byte-reversed hex of HashOut::to_bytes
    hex_be = "".join(f"{int(x):016x}" for x in oracle.two_to_one(z, z)[::-1])
    assert hex_be == "c71603f33a1144ca7953db0ab48808f4c4055e3364a246c33c18a9786cb0b359"


def test_round_constants_fingerprints(oracle):
    rc = oracle.round_constants()
    assert hex(int(rc[0])) == "0xb585f766f2144405" and hex(int(rc[359])) == "0xbc8dfb627fe558fc"
    assert hex(int(np.bitwise_xor.reduce(rc))) == "0xd95d3c3bb2fe42e3"
    assert (rc < np.uint64(P)).all()


def test_permutation_vectors(oracle):
    g = json.load(open(os.path.join(GOLDEN, "poseidon_kat.json")))
    for case in g["permutation"]:
        got = oracle.permute(np.array([int(x, 16) for x in case["in"]], dtype=np.uint64))
        assert [f"{int(x):016x}" for x in got] == case["out"]


def test_fast_and_naive_mds_agree(oracle, rng):
    """glo_poseidon_permute runs upstream's CPU schedule (fast partial rounds with tables DERIVED from the MDS matrix,
    mds_row_shf-style full rounds); it must be the same map as the literal 30-round form on every kind of input."""
    P = 0xFFFFFFFF00000001
    cases = [rand_field(rng, (12,)) for _ in range(200)]
    cases += [np.zeros(12, dtype=np.uint64), np.full(12, P - 1, dtype=np.uint64),
              np.full(12, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64),          # non-canonical inputs are taken mod p
              np.full(12, 0xFFFFFFFF, dtype=np.uint64), np.full(12, 0xFFFFFFFF00000000, dtype=np.uint64),
              np.arange(12, dtype=np.uint64)]
    for e in range(12):
        v = np.zeros(12, dtype=np.uint64)
        v[e] = P - 1
        cases.append(v)
    for s in cases:
        a, b = s.copy(), s.copy()
        oracle.lib().glo_poseidon_permute_naive(a.ctypes.data_as(oracle.u64p))
        oracle.lib().glo_poseidon_permute_slow(b.ctypes.data_as(oracle.u64p))
        got = oracle.permute(s)
        assert np.array_equal(a, got) and np.array_equal(b, got)
        assert (got < np.uint64(P)).all()


def test_smt_leaf_hash_pad_consistency(oracle):
    # src/smt/goldilocks_poseidon/mod.rs:167-181 vs src/smt/gadgets/common.rs:87-101
    k, v = oracle.from_u128(1), oracle.from_u128(2)
    a = oracle.hash_pad(np.concatenate([k, v, [np.uint64(1)]]))
    b = oracle.hash_no_pad(np.concatenate([k, v, np.array([1, 1, 0, 1], dtype=np.uint64)]))
    assert a.tolist() == b.tolist() == oracle.smt_leaf_hash(k, v).tolist()
    assert a.tolist() == [9613647271972624781, 17898244898336278454, 17153022918269186278, 8190762674233093240]


def test_smt_fixtures(oracle):
    t = oracle.Smt()
    for k, v in [(1, 2), (12, 1), (5, 51)]:  # src/smt/gadgets/verify/mod.rs:24-34
        t.set(oracle.from_u128(k), oracle.from_u128(v))
    assert t.root().tolist() == [16994558480514381166, 8559105504417206749, 13458782878755336329, 17099432696459526118]
    f = t.find(oracle.from_u128(5))
    assert f["found"] and f["siblings"].shape[0] == 3
    assert f["siblings"][0].tolist() == [9925912915451152018, 8957749234548519353, 212602008497491439, 10987709841487322779]
    assert not f["siblings"][1].any()


def test_smt_process_proofs_verify_and_order_independence(oracle, rng):
    kv = [(rand_field(rng, (4,)), rand_field(rng, (4,))) for _ in range(12)]
    roots = []
    for order in (range(12), reversed(range(12))):
        t = oracle.Smt()
        recs = [t.set(*kv[i]) for i in order]
        roots.append(t.root().tolist())
        st = oracle.smt_verify_process_batch(np.array(recs, dtype=oracle.SMT_PROOF_DTYPE))
        assert (st == 0).all()
    assert roots[0] == roots[1]


def test_field_basics(oracle):
    assert oracle.lib().glo_pow(7, (P - 1) >> 32) == 1753635133440165772
    assert oracle.lib().glo_primitive_root_of_unity(32) == 1753635133440165772
    assert oracle.lib().glo_mul(P - 1, P - 1) == 1
    assert oracle.lib().glo_add(P - 1, 5) == 4


@pytest.mark.parametrize("lg", [0, 1, 3, 6, 10])
def test_fft_roundtrip_and_definition(oracle, rng, lg):
    n = 1 << lg
    a = rand_field(rng, (n,))
    assert np.array_equal(oracle.ifft(oracle.fft(a)), a)
    assert np.array_equal(oracle.coset_ifft(oracle.coset_fft(a, 7), 7), a)
    if n <= 64:  # O(n^2) definition
        w = oracle.lib().glo_primitive_root_of_unity(lg)
        want = [sum(int(a[i]) * pow(w, i * j, P) for i in range(n)) % P for j in range(n)]
        assert oracle.fft(a).tolist() == want
        want = [sum(int(a[i]) * pow(7 * pow(w, j, P), i, P) for i in range(n)) % P for j in range(n)]
        assert oracle.coset_fft(a, 7).tolist() == want


def test_commit_structure(oracle):
    """from_values: leaves[i] = lde[bitrev(i)], coset k of the LDE lands in leaf block bitrev_r(k)."""
    c, lg_n, r, h = 5, 4, 3, 2
    n, N = 1 << lg_n, 1 << (lg_n + r)
    values = oracle.synthetic_values(c, n)
    res = oracle.commit_from_values(values, r, h)
    for col in range(c):
        assert np.array_equal(oracle.ifft(values[col]), res["coeffs"][col])
        padded = np.zeros(N, dtype=np.uint64)
        padded[:n] = res["coeffs"][col]
        lde = oracle.coset_fft(padded, 7)
        for i in range(N):
            assert res["leaves"][i][col] == lde[oracle.lib().glo_reverse_bits(i, lg_n + r)]
    wN = oracle.lib().glo_primitive_root_of_unity(lg_n + r)
    for k in range(1 << r):
        shift = 7 * pow(wN, k, P) % P
        sub = oracle.coset_fft(res["coeffs"][0], shift)
        blk = oracle.lib().glo_reverse_bits(k, r)
        for i in range(n):
            assert res["leaves"][blk * n + oracle.lib().glo_reverse_bits(i, lg_n)][0] == sub[i]
    for i in range(N):
        sib = oracle.merkle_prove(res["digests"], N, h, i)
        assert oracle.merkle_verify(res["leaves"][i], i, sib, res["cap"], h)
    assert not oracle.merkle_verify(res["leaves"][1], 0, oracle.merkle_prove(res["digests"], N, h, 0), res["cap"], h)


def test_merkle_edge_cases(oracle, rng):
    leaves = rand_field(rng, (8, 3))  # hash_or_noop: short leaves are copied
    digests, cap = oracle.merkle_tree(leaves, 3)
    assert digests.shape[0] == 0
    assert np.array_equal(cap[:, :3], leaves) and not cap[:, 3].any()
    with pytest.raises(ValueError):
        oracle.merkle_tree(leaves, 4)


def test_golden_commit_caps(oracle):
    """tests/golden/commit_caps.json was generated by tests/golden/make_golden.py from this oracle; this
    keeps the oracle stable across edits (it does not add parity evidence)."""
    g = json.load(open(os.path.join(GOLDEN, "commit_caps.json")))
    for case in g["cases"]:
        v = oracle.synthetic_values(case["c"], 1 << case["lg_n"])
        res = oracle.commit_from_values(v, case["rate_bits"], case["cap_height"], want_leaves=False)
        assert [f"{int(x):016x}" for x in res["cap"].reshape(-1)] == case["cap"]


def test_golden_smt_sets(oracle):
    """tests/golden/smt_sets.json (make_golden.py): the oracle's tree.set / tree.find replayed call by call."""
    import json

    g = json.load(open(os.path.join(GOLDEN, "smt_sets.json")))
    un = lambda xs: np.array([int(x, 16) for x in xs], dtype=np.uint64)  # noqa: E731
    t = oracle.Smt()
    for c in g["calls"]:
        r = t.set(un(c["key"]), un(c["value"]))
        assert int(r["fnc"]) == c["fnc"] and int(r["is_old0"]) == c["is_old0"]
        for f in ("old_root", "new_root", "old_key", "old_value", "new_key", "new_value"):
            assert np.array_equal(r[f], un(c[f])), f
        ns = int(r["num_siblings"])
        assert np.array_equal(r["siblings"][:ns].reshape(-1), un(c["siblings"]))
    assert np.array_equal(t.root(), un(g["root"]))
    assert g["calls"][2]["new_root"][0] == f"{16994558480514381166:016x}"      # the three-insert fixture of SURVEY Appendix B
    for q in g["finds"]:
        f = t.find(un(q["key"]))
        assert bool(f["found"]) == q["found"] and bool(f["is_old0"]) == q["is_old0"]
        assert np.array_equal(f["siblings"].reshape(-1), un(q["siblings"]))
