"""N3: gl_quotient_polys (compute_quotient_polys on the resident oracles) against the oracle's literal restatement, bit for
bit, on circuits built from the reference's three custom gates (tests/quotient_circuit.py), plus the acceptance
properties the reference / upstream verifier use: degree bound and the quotient identity at a random point."""
import numpy as np
import pytest

from conftest import P
from quotient_circuit import build

pytestmark = pytest.mark.gpu


def _setup(glb, ctx, oracle, lg_n, corrupt=None, seed=None):
    c = build(lg_n, seed=seed if seed is not None else lg_n, corrupt=corrupt)
    rng = np.random.default_rng(1234 + lg_n)
    betas, gammas, alphas = (rng.integers(0, P, 2, dtype=np.uint64) for _ in range(3))
    zs = oracle.permutation_zs(c["circuit"], c["k_is"], c["wires"], c["sigmas"], betas, gammas)
    vals = [np.concatenate([c["constants"], c["sigmas"]]), c["wires"], zs]
    batches = [glb.PolynomialBatch.from_values(v, 3, False, 4, ctx=ctx) for v in vals]
    return c, vals, batches, betas, gammas, alphas


@pytest.mark.parametrize("lg_n", [5, 8, 12, 14, 16])
def test_quotient_matches_oracle(glb, ctx, oracle, lg_n):
    c, vals, batches, betas, gammas, alphas = _setup(glb, ctx, oracle, lg_n)
    got = glb.host.compute_quotient_polys(c["circuit"], c["gates"], c["k_is"], *batches, c["pih"], betas, gammas, alphas, ctx=ctx)
    leaves = [oracle.commit_from_values(v, 3, 4)["leaves"] for v in vals]
    want = oracle.quotient_polys(c["circuit"], c["gates"], c["k_is"], *leaves, 3, c["pih"], betas, gammas, alphas)
    assert np.array_equal(got, want)
    n = 1 << lg_n
    for ch in range(2):                                   # degree <= 8 n - 9 for a satisfying witness
        assert not got[ch * 8 + 7][n - 8:].any() and got[ch * 8 + 7][: n - 8].any()
    # the chunks are what the prover commits next: from_coeffs takes them as they are
    qb = glb.PolynomialBatch.from_coeffs(got, 3, False, 4, ctx=ctx)
    assert np.array_equal(qb.merkle_tree.cap, oracle.commit_from_coeffs(want, 3, 4, want_leaves=False)["cap"])
    for b in batches + [qb]:
        b.free()


@pytest.mark.parametrize("corrupt", ["bit", "copy"])
def test_corrupted_witness_breaks_the_degree_bound(glb, ctx, oracle, corrupt):
    """interleave_u32.rs:341-352 (test_low_degree) turned around: a witness that violates a gate (or a copy constraint)
    leaves a 'quotient' whose top coefficients do not vanish -- and the device still equals the oracle on it."""
    lg_n = 7
    c = build(lg_n, seed=3, corrupt=corrupt)
    rng = np.random.default_rng(5)
    betas, gammas, alphas = (rng.integers(0, P, 2, dtype=np.uint64) for _ in range(3))
    good = build(lg_n, seed=3)
    zs = oracle.permutation_zs(good["circuit"], good["k_is"], good["wires"], good["sigmas"], betas, gammas)   # Z of the honest witness
    vals = [np.concatenate([c["constants"], c["sigmas"]]), c["wires"], zs]
    batches = [glb.PolynomialBatch.from_values(v, 3, False, 4, ctx=ctx) for v in vals]
    got = glb.host.compute_quotient_polys(c["circuit"], c["gates"], c["k_is"], *batches, c["pih"], betas, gammas, alphas, ctx=ctx)
    leaves = [oracle.commit_from_values(v, 3, 4)["leaves"] for v in vals]
    want = oracle.quotient_polys(c["circuit"], c["gates"], c["k_is"], *leaves, 3, c["pih"], betas, gammas, alphas)
    assert np.array_equal(got, want)
    n = 1 << lg_n
    assert any(got[ch * 8 + 7][n - 8:].any() for ch in range(2))
    for b in batches:
        b.free()


def test_quotient_rejects_mismatched_geometry(glb, ctx, oracle):
    c, vals, batches, betas, gammas, alphas = _setup(glb, ctx, oracle, 5)
    bad = list(c["circuit"])
    bad[2] = 72                                                      # num_routed_wires: sigma column count no longer matches
    with pytest.raises(glb.GlPanic):
        glb.host.compute_quotient_polys(bad, c["gates"], c["k_is"], *batches, c["pih"], betas, gammas, alphas, ctx=ctx)
    bad = list(c["circuit"])
    bad[6] = 16                                                      # quotient_degree_factor 16 > 2^rate_bits
    with pytest.raises(glb.GlPanic):
        glb.host.compute_quotient_polys(bad, c["gates"], c["k_is"], *batches, c["pih"], betas, gammas, alphas, ctx=ctx)
    gates = [list(g) for g in c["gates"]]
    gates[3][0] = 9                                                  # unknown gate kind
    with pytest.raises(glb.GlPanic):
        glb.host.compute_quotient_polys(c["circuit"], gates, c["k_is"], *batches, c["pih"], betas, gammas, alphas, ctx=ctx)
    for b in batches:
        b.free()
