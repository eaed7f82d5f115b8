"""The FRI half of the oracle (oracle/fri_oracle.py) checks itself the way the reference's circuit tests do:
prove, then verify (every reference test is `data.prove(pw)` followed by `data.verify(proof)`,
e.g. src/ecdsa/gadgets/ecdsa.rs:349-352)."""
import numpy as np
import pytest

from conftest import P, rand_field


def make_instance(oracle, rng, degree_bits, rate_bits=3, cap_height=4):
    from oracle import fri_oracle as fo

    n = 1 << degree_bits
    lde = n << rate_bits
    coeffs = np.zeros((lde, 2), dtype=np.uint64)
    coeffs[:n] = rand_field(rng, (n, 2))
    values = oracle.ext_coset_fft(coeffs, 7)
    base = oracle.commit_from_values(oracle.synthetic_values(6, n), rate_bits, cap_height)
    tree = fo.MerkleTree(base["leaves"], cap_height)
    return coeffs, values, tree


def direct_eval(coeffs, degree_bits):
    from oracle import fri_oracle as fo

    def f(x_index, rows, subgroup_x):
        return fo.eval_ext_poly(coeffs[: 1 << degree_bits], (subgroup_x, 0))

    return f


@pytest.mark.parametrize("degree_bits,pow_bits", [(6, 4), (10, 8), (9, 16)])
def test_fri_prove_then_verify(oracle, rng, degree_bits, pow_bits):
    from oracle import fri_oracle as fo

    coeffs, values, tree = make_instance(oracle, rng, degree_bits)
    ch = fo.Challenger()
    ch.observe_cap(tree.cap)
    proof = fo.fri_proof([tree], coeffs, values, ch, degree_bits, pow_bits=pow_bits, num_query_rounds=6)
    vch = fo.Challenger()
    vch.observe_cap(tree.cap)
    assert fo.verify_fri_proof(proof, [tree.cap], 4, vch, degree_bits, direct_eval(coeffs, degree_bits), pow_bits=pow_bits,
                               num_query_rounds=6)
    # the verifier rejects a tampered proof
    bad = dict(proof)
    bad["final_poly"] = proof["final_poly"].copy()
    bad["final_poly"][0][0] ^= np.uint64(1)
    vch = fo.Challenger()
    vch.observe_cap(tree.cap)
    with pytest.raises(AssertionError):
        fo.verify_fri_proof(bad, [tree.cap], 4, vch, degree_bits, direct_eval(coeffs, degree_bits), pow_bits=pow_bits,
                            num_query_rounds=6)


def test_challenger_duplex_semantics(oracle):
    """Outputs are popped from the back; observing clears pending outputs; 8 inputs trigger a duplex."""
    from oracle import fri_oracle as fo

    ch = fo.Challenger()
    for i in range(8):
        ch.observe_element(i)
    st = oracle.permute(np.array(list(range(8)) + [0, 0, 0, 0], dtype=np.uint64))
    assert ch.get_challenge() == int(st[7]) and ch.get_challenge() == int(st[6])
    ch.observe_element(5)
    s2 = st.copy()
    s2[0] = 5
    assert ch.get_challenge() == int(oracle.permute(s2)[7])


def test_reduction_strategy(oracle):
    from oracle import fri_oracle as fo

    assert fo.reduction_arity_bits(20, 3, 4) == [4, 4, 4, 4]
    assert fo.reduction_arity_bits(16, 3, 4) == [4, 4, 4]
    assert fo.reduction_arity_bits(5, 3, 4) == []


def test_prove_openings_then_verify(oracle, rng):
    """PolynomialBatch::prove_openings -> verify_fri_proof with fri_combine_initial: the full opening argument of a
    proof (two batches: every polynomial at zeta, a few at g * zeta), prove then verify, then a wrong opening."""
    from oracle import fri_oracle as fo

    degree_bits, rate_bits, cap_height = 8, 3, 4
    n = 1 << degree_bits
    cols = (4, 9, 3)
    polys, trees = [], []
    for k, c in enumerate(cols):
        res = oracle.commit_from_values(oracle.synthetic_values(c, n, seed=50 + k), rate_bits, cap_height)
        polys.append(res["coeffs"])
        trees.append(fo.MerkleTree(res["leaves"], cap_height))
    zeta = tuple(int(x) for x in rand_field(rng, (2,)))
    g = oracle.lib().glo_primitive_root_of_unity(degree_bits)
    batches = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]),
               (fo.ext_scalar(zeta, g), [(1, pi) for pi in range(4)])]
    openings = fo.opening_set(polys, batches)
    ch = fo.Challenger()
    for t in trees:
        ch.observe_cap(t.cap)
    proof = fo.prove_openings(polys, trees, batches, ch, degree_bits, rate_bits, cap_height, pow_bits=6, num_query_rounds=5)
    vch = fo.Challenger()
    for t in trees:
        vch.observe_cap(t.cap)
    assert fo.verify_openings(proof, openings, [t.cap for t in trees], batches, vch, degree_bits, rate_bits, cap_height, 6, 5)
    openings[0][2] = fo.ext_add(openings[0][2], (1, 0))
    vch = fo.Challenger()
    for t in trees:
        vch.observe_cap(t.cap)
    with pytest.raises(AssertionError):
        fo.verify_openings(proof, openings, [t.cap for t in trees], batches, vch, degree_bits, rate_bits, cap_height, 6, 5)


@pytest.mark.parametrize("times_x", [False, True])
def test_prove_openings_then_verify_in_both_fork_forms(oracle, rng, times_x):
    """The one fork-version switch of the FRI layer (GL_COMPAT_FRI_FINAL_POLY_TIMES_X): prover and verifier agree in either
    form, and a proof of one form is rejected by a verifier of the other."""
    from oracle import fri_oracle as fo

    degree_bits, cols = 6, (4, 3)
    n = 1 << degree_bits
    polys, trees = [], []
    for k, c in enumerate(cols):
        res = oracle.commit_from_values(oracle.synthetic_values(c, n, seed=50 + k), 3, 4)
        polys.append(res["coeffs"])
        trees.append(fo.MerkleTree(res["leaves"], 4))
    zeta = tuple(int(x) for x in rand_field(rng, (2,)))
    instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)])]
    ch, vch, xch = fo.Challenger(), fo.Challenger(), fo.Challenger()
    for t in trees:
        for c in (ch, vch, xch):
            c.observe_cap(t.cap)
    proof = fo.prove_openings(polys, trees, instance, ch, degree_bits, pow_bits=6, num_query_rounds=10, times_x=times_x)
    openings = fo.opening_set(polys, instance)
    caps = [t.cap for t in trees]
    assert fo.verify_openings(proof, openings, caps, instance, vch, degree_bits, pow_bits=6, num_query_rounds=10, times_x=times_x)
    with pytest.raises(AssertionError):
        fo.verify_openings(proof, openings, caps, instance, xch, degree_bits, pow_bits=6, num_query_rounds=10, times_x=not times_x)
