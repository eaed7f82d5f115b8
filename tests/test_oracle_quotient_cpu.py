"""N3 oracle (oracle/gl_oracle.c glo_permutation_zs / glo_quotient_polys) checked the way the reference checks its gates and
the upstream verifier checks a proof: the quotient identity vanishing(zeta) = Z_H(zeta) * t(zeta) at a random point, and
the degree bound (src/u32/gates/interleave_u32.rs:341-352 test_low_degree), for a satisfying witness; both must FAIL for a
witness with a flipped decomposition bit or a broken copy constraint."""
import numpy as np
import pytest

from conftest import P
from quotient_circuit import build


def horner(coeffs, x):
    acc = 0
    for c in coeffs[::-1].tolist():
        acc = (acc * x + int(c)) % P
    return acc


def run(oracle, lg_n, corrupt):
    c = build(lg_n, seed=lg_n, corrupt=corrupt)
    rng = np.random.default_rng(99)
    betas, gammas, alphas = (rng.integers(0, P, 2, dtype=np.uint64) for _ in range(3))
    try:
        zs = oracle.permutation_zs(c["circuit"], c["k_is"], c["wires"], c["sigmas"], betas, gammas)
    except ValueError:
        return None, c                      # grand product does not close
    commits = [oracle.commit_from_values(v, 3, 4) for v in (np.concatenate([c["constants"], c["sigmas"]]), c["wires"], zs)]
    q = oracle.quotient_polys(c["circuit"], c["gates"], c["k_is"], commits[0]["leaves"], commits[1]["leaves"], commits[2]["leaves"], 3,
                              c["pih"], betas, gammas, alphas)
    return (q, commits, betas, gammas, alphas), c


def identity_holds(oracle, c, q, commits, betas, gammas, alphas):
    n = 1 << c["circuit"][0]
    zeta = 0x1234567890ABCDEF % P
    g = oracle.lib().glo_primitive_root_of_unity(c["circuit"][0])
    at = lambda co, x: np.array([horner(p, x) for p in co], dtype=np.uint64)   # noqa: E731
    lcs, lw, lz = (at(cm["coeffs"], zeta) for cm in commits)
    nz = at(commits[2]["coeffs"], zeta * g % P)
    van = oracle.vanishing_at_point(c["circuit"], c["gates"], c["k_is"], zeta, lcs, lw, lz, nz, c["pih"], betas, gammas, alphas)
    zh = (pow(zeta, n, P) - 1) % P
    ok = True
    for ch in range(2):
        t = 0
        for k in range(8)[::-1]:
            t = (t * pow(zeta, n, P) + horner(q[ch * 8 + k], zeta)) % P
        ok &= int(van[ch]) == zh * t % P
    return ok


@pytest.mark.parametrize("lg_n", [5, 8])
def test_quotient_identity_and_degree(oracle, lg_n):
    res, c = run(oracle, lg_n, None)
    assert c["num_copies"] > 0
    q = res[0]
    n = 1 << lg_n
    # degree: the permutation terms have degree 9 (n - 1), so t has degree <= 8 n - 9: the top 8 coefficients vanish
    for ch in range(2):
        assert not q[ch * 8 + 7][n - 8:].any()
        assert q[ch * 8 + 7][: n - 8].any()
    assert identity_holds(oracle, c, *res)


@pytest.mark.parametrize("corrupt", ["bit", "copy"])
def test_corrupted_witness_fails(oracle, corrupt):
    res, c = run(oracle, 6, corrupt)
    if res is None:
        assert corrupt == "copy"            # the grand product itself refuses
        return
    q = res[0]
    n = 1 << 6
    assert any(q[ch * 8 + 7][n - 8:].any() for ch in range(2)) or not identity_holds(oracle, c, *res)
    assert not identity_holds(oracle, c, *res)
