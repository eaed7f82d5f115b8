"""Bit-for-bit parity at the BENCHMARKED shapes (BASELINE.json configs[1] = 2^20 x 135 and config 3's trace shape
2^18 x 135; rate_bits 3, cap_height 4) against tests/golden/commit_fullsize.json, which the CPU oracle produced once
(tests/golden/make_golden_fullsize.py).  Every output of PolynomialBatch::from_values is compared in full:
coefficients, leaves (mirror mode, row-major leaf order), digests (plonky2's in-order layout) and the cap, plus the
sampled rows / Merkle paths through gl_commit_open.  The oracle is not in the loop here (the fixture is)."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "commit_fullsize.json")


def sha(a: np.ndarray) -> str:
    flat = np.ascontiguousarray(a, dtype="<u8").reshape(-1)
    h = hashlib.sha256()
    step = 1 << 24
    for i in range(0, flat.size, step):
        h.update(flat[i:i + step].tobytes())
    return h.hexdigest()


def hx(a):
    return [f"{int(x):016x}" for x in np.asarray(a).reshape(-1)]


def _cases():
    return json.load(open(GOLD))["cases"]


@pytest.mark.parametrize("lg_n", [18, 20])
def test_commit_bit_exact_at_benchmark_size(glb, ctx, oracle, lg_n):
    case = next(c for c in _cases() if c["lg_n"] == lg_n)
    c, r, h = case["c"], case["rate_bits"], case["cap_height"]
    values = oracle.synthetic_values(c, 1 << lg_n)          # input generator only (splitmix64), not the oracle's arithmetic
    b = glb.PolynomialBatch.from_values(values, r, False, h, ctx=ctx)
    assert hx(b.merkle_tree.cap) == case["cap"]
    coeffs = b.polynomials
    for j, want in case["coeff_cols"].items():
        assert sha(coeffs[int(j)]) == want, f"coefficient column {j}"
    assert sha(coeffs) == case["coeffs_sha256"]
    rows, paths = b.open(case["leaf_indices"])
    assert [hx(x) for x in rows] == case["leaf_rows"]
    assert [hx(x) for x in paths] == case["leaf_paths"]
    assert sha(b.merkle_tree.digests) == case["digests_sha256"]
    b.merkle_tree._digests = None
    assert sha(b.merkle_tree.leaves) == case["leaves_sha256"]      # 9 GB at 2^20: every LDE value, in leaf order
    b.merkle_tree._leaves = None
    b.free()
    ctx.trim()


def test_commit_from_device_resident_values_matches_golden(glb, ctx, oracle):
    """The path bench.py's `value` leg times (GL_DEVICE in, outputs left on the device) against the same fixture."""
    import torch

    case = next(c for c in _cases() if c["lg_n"] == 20)
    c, r, h = case["c"], case["rate_bits"], case["cap_height"]
    values = torch.from_numpy(oracle.synthetic_values(c, 1 << 20).view(np.int64)).cuda()
    b = glb.PolynomialBatch.from_values(values, r, False, h, ctx=ctx)
    assert hx(b.merkle_tree.cap) == case["cap"]
    for j, want in case["coeff_cols"].items():
        assert sha(b.polynomials[int(j)].cpu().numpy().view(np.uint64)) == want
    rows, paths = b.open(case["leaf_indices"])
    assert [hx(x) for x in rows] == case["leaf_rows"] and [hx(x) for x in paths] == case["leaf_paths"]
    b.free()
    del values
    ctx.trim()
