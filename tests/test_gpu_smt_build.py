"""N2: bulk build of the sparse Merkle tree on the device against the oracle's restatement of
src/smt/tree.rs driven one `set` at a time."""
import numpy as np
import pytest

from conftest import P, rand_field

pytestmark = pytest.mark.gpu


def _oracle_root(oracle, keys, values):
    t = oracle.Smt()
    for k, v in zip(keys, values):
        t.set(k, v)
    return t, t.root()


@pytest.mark.parametrize("m", [0, 1, 2, 3, 17, 500])
def test_bulk_root_matches_sequential_sets(glb, ctx, oracle, rng, m):
    keys, values = rand_field(rng, (m, 4)), rand_field(rng, (m, 4)) | np.uint64(1)
    _, want = _oracle_root(oracle, keys, values)
    assert np.array_equal(glb.host.smt_build_tree(keys, values), want)


def test_reference_fixture_three_inserts(glb, ctx, oracle):
    """(1 -> 2), (12 -> 1), (5 -> 51): src/smt/gadgets/verify/mod.rs:24-34, SURVEY Appendix B."""
    keys = np.stack([oracle.from_u128(k) for k in (1, 12, 5)])
    values = np.stack([oracle.from_u128(v) for v in (2, 1, 51)])
    root, nodes, leaf_hashes = glb.host.smt_build_tree(keys, values, want_nodes=True)
    assert root.tolist() == [16994558480514381166, 8559105504417206749, 13458782878755336329, 17099432696459526118]
    assert leaf_hashes[0].tolist() == oracle.smt_leaf_hash(keys[0], values[0]).tolist()
    # internal nodes: the split of {1, 5} at depth 2, its one-child parent at depth 1, the root
    assert nodes.shape[0] == 3
    by_hash = {tuple(n[:4].tolist()): n for n in nodes}
    top = by_hash[tuple(root.tolist())]
    assert top[4:8].tolist() == oracle.smt_leaf_hash(keys[1], values[1]).tolist()       # key 12: path bit 0 = 0 -> left
    mid = by_hash[tuple(top[8:12].tolist())]
    assert not mid[8:12].any()                                                           # chain node H(x, 0)
    low = by_hash[tuple(mid[4:8].tolist())]
    assert low[4:8].tolist() == oracle.smt_leaf_hash(keys[0], values[0]).tolist()
    assert low[8:12].tolist() == oracle.smt_leaf_hash(keys[2], values[2]).tolist()


def test_long_shared_prefixes_and_order_independence(glb, ctx, oracle, rng):
    """Keys that agree on their first 100..250 path bits (deep chains of one-child nodes), small limbs, and a
    shuffled copy of the batch."""
    base = rand_field(rng, (4,))
    keys = [base.copy()]
    for bit in (250, 200, 129, 128, 100, 64, 63, 5, 0):
        k = base.copy()
        k[bit // 64] ^= np.uint64(1) << np.uint64(bit % 64)
        keys.append(k)
    keys += [oracle.from_u128(x) for x in (1, 2, 3, 4, 12, 5, 1 << 100)]
    keys = np.stack(keys)
    values = rand_field(rng, keys.shape) | np.uint64(1)
    _, want = _oracle_root(oracle, keys, values)
    got, nodes, _ = glb.host.smt_build_tree(keys, values, want_nodes=True)
    assert np.array_equal(got, want)
    perm = rng.permutation(keys.shape[0])
    assert np.array_equal(glb.host.smt_build_tree(keys[perm], values[perm]), want)
    # the sibling path the oracle's tree reports for a key is made of the device's nodes / leaf hashes / zeros
    t, _ = _oracle_root(oracle, keys, values)
    known = {tuple(n[:4].tolist()) for n in nodes} | {tuple(oracle.smt_leaf_hash(k, v).tolist()) for k, v in zip(keys, values)}
    f = t.find(keys[3])
    assert f["found"] and f["siblings"].shape[0] >= 100
    assert all((not s.any()) or tuple(s.tolist()) in known for s in f["siblings"])


def test_zero_values_are_dropped_and_duplicates_panic(glb, ctx, oracle, rng):
    keys, values = rand_field(rng, (20, 4)), rand_field(rng, (20, 4)) | np.uint64(1)
    values[3] = 0
    values[11] = 0
    keep = values.any(axis=1)
    _, want = _oracle_root(oracle, keys[keep], values[keep])
    assert np.array_equal(glb.host.smt_build_tree(keys, values), want)
    keys[7] = keys[2]
    with pytest.raises(glb.GlPanic):
        glb.host.smt_build_tree(keys, values)


def test_bulk_build_scale(glb, ctx, oracle, rng):
    """2^16 random entries: the root against 2^16 sequential oracle inserts."""
    m = 1 << 16
    keys, values = rand_field(rng, (m, 4)), rand_field(rng, (m, 4)) | np.uint64(1)
    _, want = _oracle_root(oracle, keys, values)
    assert np.array_equal(glb.host.smt_build_tree(keys, values), want)
