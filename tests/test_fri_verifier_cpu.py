"""Host arithmetic of the product-side FRI verifier (plonky2-lib_b200/fri_verifier.py) against the oracle's independent
restatement (oracle/fri_oracle.py): extension field, compute_evaluation (barycentric here, Lagrange double loop there),
fri_combine_initial with PrecomputedReducedOpenings, bit reversal.  No device needed."""
import importlib

import numpy as np
import pytest

from conftest import P


@pytest.fixture(scope="module")
def fv():
    return importlib.import_module("plonky2-lib_b200.fri_verifier")


def _r(rng):
    return int(rng.integers(0, P, dtype=np.uint64))


def test_extension_field_helpers(fv):
    from oracle import fri_oracle as fo

    rng = np.random.default_rng(11)
    for _ in range(50):
        a, b = (_r(rng), _r(rng)), (_r(rng), _r(rng))
        assert fv._emul(a, b) == tuple(fo.ext_mul(a, b)) and fv._eadd(a, b) == tuple(fo.ext_add(a, b))
        assert fv._esub(a, b) == tuple(fo.ext_sub(a, b))
        if a != (0, 0):
            assert fv._emul(a, fv._einv(a)) == (1, 0)
        e = int(rng.integers(0, 1000))
        assert fv._epow(a, e) == tuple(fo.ext_pow(a, e))
    assert fv._emul((0, 1), (0, 1)) == (7, 0)                      # X^2 = 7
    assert [fv._bitrev(x, 4) for x in (0, 1, 2, 8, 15)] == [0, 8, 4, 1, 15] and fv._bitrev(0, 0) == 0
    assert pow(fv._root_of_unity(5), 32, P) == 1 and pow(fv._root_of_unity(5), 16, P) != 1


@pytest.mark.parametrize("arity_bits", [1, 2, 3, 4])
def test_compute_evaluation_matches_the_oracle(fv, arity_bits):
    from oracle import fri_oracle as fo

    rng = np.random.default_rng(100 + arity_bits)
    for _ in range(5):
        evals = rng.integers(0, P, (1 << arity_bits, 2), dtype=np.uint64)
        x, within, beta = _r(rng) or 1, int(rng.integers(0, 1 << arity_bits)), (_r(rng), _r(rng))
        assert fv.compute_evaluation(x, within, arity_bits, evals, beta) == tuple(fo.compute_evaluation(x, within, arity_bits, evals, beta))
    # beta on the coset itself: the interpolant takes the tabulated value
    evals = rng.integers(0, P, (1 << arity_bits, 2), dtype=np.uint64)
    x = 12345
    g = fv._root_of_unity(arity_bits)
    pt = x * pow(g, 1, P) % P                      # within = 0 -> the coset starts at x; second point
    got = fv.compute_evaluation(x, 0, arity_bits, evals, (pt, 0))
    assert got == (int(evals[fv._bitrev(1, arity_bits)][0]), int(evals[fv._bitrev(1, arity_bits)][1]))


def test_fri_combine_initial_matches_the_oracle(fv):
    from oracle import fri_oracle as fo

    rng = np.random.default_rng(7)
    instance = [((_r(rng), _r(rng)), [(0, 0), (0, 1), (1, 0), (2, 3)]), ((_r(rng), _r(rng)), [(1, 1), (2, 0)])]
    rows = [rng.integers(0, P, 2, dtype=np.uint64), rng.integers(0, P, 2, dtype=np.uint64), rng.integers(0, P, 4, dtype=np.uint64)]
    openings = [[(_r(rng), _r(rng)) for _ in polys] for _, polys in instance]
    alpha, x = (_r(rng), _r(rng)), _r(rng)
    got = fv.fri_combine_initial(instance, rows, alpha, x, fv.precomputed_reduced_openings(openings, alpha))
    assert got == tuple(fo.fri_combine_initial(instance, openings, rows, alpha, x))


def test_final_polynomial_evaluation(fv):
    from oracle import fri_oracle as fo

    rng = np.random.default_rng(3)
    coeffs = rng.integers(0, P, (32, 2), dtype=np.uint64)
    x = _r(rng)
    assert fv._eval_final_poly(coeffs, x) == tuple(fo.eval_ext_poly(coeffs, (x, 0)))


def test_reduction_strategy_matches_the_oracle():
    """FriReductionStrategy::ConstantArityBits(4, 5) -> reduction_arity_bits for every degree the configs can see."""
    from oracle import fri_oracle as fo

    glb = importlib.import_module("plonky2-lib_b200")
    fri = importlib.import_module("plonky2-lib_b200.fri")
    for rate_bits, cap_height in ((3, 4), (1, 0), (2, 3), (4, 4)):
        cfg = glb.FriConfig(rate_bits=rate_bits, cap_height=cap_height)
        for degree_bits in range(0, 25):
            got = list(fri.FriParams.for_degree(cfg, degree_bits).reduction_arity_bits)
            assert got == fo.reduction_arity_bits(degree_bits, rate_bits, cap_height), (degree_bits, rate_bits, cap_height)
            assert sum(got) <= degree_bits or not got
