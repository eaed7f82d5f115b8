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


def test_build_tree_is_what_the_sequence_of_sets_leaves(glb, ctx, oracle, rng):
    """`set` semantics for batches that are not plain inserts: a zero value removes its key (also one inserted earlier in
    the same batch), a repeated key is an update (the last value wins), a zero value for an absent key is a no-op."""
    keys, values = rand_field(rng, (20, 4)), rand_field(rng, (20, 4)) | np.uint64(1)
    values[3] = 0                      # no-op: key 3 was never inserted
    values[11] = 0
    _, want = _oracle_root(oracle, keys, values)
    assert np.array_equal(glb.host.smt_build_tree(keys, values), want)
    # set(k, v); set(k, 0): k is gone (the (k, 0) entry used to be discarded and k stayed in the tree)
    k2 = np.concatenate([keys, keys[[5]], keys[[6]], keys[[3]]])
    v2 = np.concatenate([values, np.zeros((1, 4), dtype=np.uint64), rand_field(rng, (1, 4)) | np.uint64(1),
                         rand_field(rng, (1, 4)) | np.uint64(1)])
    _, want2 = _oracle_root(oracle, k2, v2)
    assert not np.array_equal(want2, want)
    assert np.array_equal(glb.host.smt_build_tree(k2, v2), want2)
    # set(k, 0); set(k, v): k is there
    k3 = np.stack([keys[0], keys[1], keys[0]])
    v3 = np.stack([np.zeros(4, dtype=np.uint64), values[1], values[0]])
    assert np.array_equal(glb.host.smt_build_tree(k3, v3), _oracle_root(oracle, k3, v3)[1])
    # repeated keys without removals: updates
    keys[7] = keys[2]
    _, want4 = _oracle_root(oracle, keys, values)
    assert np.array_equal(glb.host.smt_build_tree(keys, values), want4)
    # the last root of the process proofs of the same sequence is the same tree
    hdr, _, _ = glb.host.smt_set_proofs(k2, v2)
    assert np.array_equal(hdr["new_root"][-1], want2)


def test_c_entry_point_rejects_zero_values_and_duplicates(glb, ctx, rng):
    """gl_smt_build is SparseMerkleTree::insert over a batch: "value must be non-zero", "given key already exists"."""
    import ctypes as C

    lib, N = ctx._lib, glb._native
    keys, values = rand_field(rng, (9, 4)), rand_field(rng, (9, 4)) | np.uint64(1)
    root, cnt = np.zeros(4, dtype=np.uint64), C.c_uint64(0)

    def build(k, v):
        return lib.gl_smt_build(ctx._h, k.ctypes.data, v.ctypes.data, k.shape[0], root.ctypes.data, None, 0, C.byref(cnt), None, N.GL_HOST)

    assert build(keys, values) == N.GL_OK
    v0 = values.copy()
    v0[4] = 0
    assert build(keys, v0) == N.GL_E_ARG and b"non-zero" in lib.gl_last_error(ctx._h)
    v0[4] = np.uint64(P)               # p = 0 (mod p)
    assert build(keys, v0) == N.GL_E_ARG
    k0 = keys.copy()
    k0[8] = k0[1]
    assert build(k0, values) == N.GL_E_ARG and b"already exists" in lib.gl_last_error(ctx._h)


def test_bulk_build_scale(glb, ctx, oracle, rng):
    """2^16 random entries: the root against 2^16 sequential oracle inserts."""
    m = 1 << 16
    keys, values = rand_field(rng, (m, 4)), rand_field(rng, (m, 4)) | np.uint64(1)
    _, want = _oracle_root(oracle, keys, values)
    assert np.array_equal(glb.host.smt_build_tree(keys, values), want)


# ---- process proofs of a batch of inserts and updates (gl_smt_set_proofs) ------------------------------------------------
def _sequential_proofs(oracle, keys, values):
    t = oracle.Smt()
    return np.array([t.set(k, v) for k, v in zip(keys, values)], dtype=oracle.SMT_PROOF_DTYPE), t.root()


def _check_proofs(glb, oracle, keys, values):
    want, root = _sequential_proofs(oracle, keys, values)
    hdr, pool, off = glb.host.smt_set_proofs(keys, values)
    m = len(keys)
    assert off.shape == (m + 1,) and int(off[0]) == 0 and int(off[-1]) == pool.shape[0]
    for f in ("old_root", "old_key", "old_value", "new_root", "new_key", "new_value", "is_old0", "fnc"):
        assert np.array_equal(hdr[f], want[f]), f
    assert np.array_equal(np.diff(off).astype(np.uint32), want["num_siblings"])
    for t in range(m):
        ns = int(want["num_siblings"][t])
        assert np.array_equal(pool[int(off[t]):int(off[t + 1])], want["siblings"][t][:ns]), t
    if m:
        assert np.array_equal(hdr["new_root"][-1], root)
        assert np.array_equal(glb.host.smt_build_tree(keys, values), root)
    # and they are what the batch verifier accepts
    assert not glb.smt_check_process_proofs(hdr, pool, off).any()
    return hdr, pool, off


@pytest.mark.parametrize("m", [1, 2, 3, 4, 33, 700])
def test_set_proofs_match_sequential_sets(glb, ctx, oracle, rng, m):
    keys, values = rand_field(rng, (m, 4)), rand_field(rng, (m, 4)) | np.uint64(1)
    _check_proofs(glb, oracle, keys, values)


def test_set_proofs_reference_fixture(glb, ctx, oracle):
    """(1 -> 2), (12 -> 1), (5 -> 51) in this order: src/smt/gadgets/verify/mod.rs:24-34."""
    keys = np.stack([oracle.from_u128(k) for k in (1, 12, 5)])
    values = np.stack([oracle.from_u128(v) for v in (2, 1, 51)])
    hdr, pool, off = _check_proofs(glb, oracle, keys, values)
    assert hdr["is_old0"].tolist() == [1, 0, 0]           # the first insert finds an empty tree, the others a leaf
    assert hdr["new_root"][2].tolist() == [16994558480514381166, 8559105504417206749, 13458782878755336329, 17099432696459526118]


def test_set_proofs_long_common_prefixes_and_orders(glb, ctx, oracle, rng):
    """Twins that agree on up to 255 leading path bits (deep one-child chains, zero siblings in the middle of a proof,
    trailing zeros trimmed), small keys that share long runs of zero bits, and the same set in three insertion orders."""
    base = rand_field(rng, (10, 4))
    twins = []
    for j, k in enumerate(base):
        for bit in ((0, 1, 5, 63, 64, 65, 128, 191, 254, 255)[j], (7, 200, 33, 100, 250, 2, 70, 129, 9, 130)[j]):
            k2 = k.copy()
            k2[bit >> 6] ^= np.uint64(1) << np.uint64(bit & 63)
            if int(k2[bit >> 6]) < P:
                twins.append(k2)
    small = np.stack([oracle.from_u128(x) for x in (0, 1, 2, 3, 4, 8, 12, 5, 1 << 40, (1 << 40) + 1, 1 << 100)])
    keys = np.unique(np.concatenate([base, np.array(twins), small]), axis=0)
    values = rand_field(rng, keys.shape) | np.uint64(1)
    for order in (np.arange(len(keys)), np.arange(len(keys))[::-1], rng.permutation(len(keys))):
        hdr, pool, off = _check_proofs(glb, oracle, keys[order].copy(), values[order].copy())
    assert int(np.diff(off).max()) > 100        # proofs that walk more than a hundred levels down


def _check_set_proofs(glb, oracle, keys, values):
    """like _check_proofs, for batches in which keys repeat (later occurrences are updates)"""
    want, root = _sequential_proofs(oracle, keys, values)
    hdr, pool, off = glb.host.smt_set_proofs(keys, values)
    for f in ("old_root", "old_key", "old_value", "new_root", "new_key", "new_value", "is_old0", "fnc"):
        assert np.array_equal(hdr[f], want[f]), f
    assert np.array_equal(np.diff(off).astype(np.uint32), want["num_siblings"])
    for t in range(len(keys)):
        ns = int(want["num_siblings"][t])
        assert np.array_equal(pool[int(off[t]):int(off[t + 1])], want["siblings"][t][:ns]), t
    assert np.array_equal(hdr["new_root"][-1], root)
    # the batch verifier says what the reference's says (a proof with 256 siblings -- twins that differ in their last
    # path bit only -- trips its assert!(siblings.len() < 256): status 1)
    assert np.array_equal(glb.smt_check_process_proofs(hdr, pool, off), oracle.smt_verify_process_batch(want))
    return hdr


@pytest.mark.parametrize("m,distinct", [(2, 1), (5, 1), (40, 7), (600, 150), (900, 900)])
def test_set_proofs_with_updates_match_sequential_sets(glb, ctx, oracle, rng, m, distinct):
    """A key may occur several times in the batch: its first occurrence is a ProcessInsert, the later ones are
    ProcessUpdates of the value the previous occurrence left (src/smt/tree.rs:174-253)."""
    pool_keys = rand_field(rng, (distinct, 4))
    keys = pool_keys[rng.integers(0, distinct, m)]
    values = rand_field(rng, (m, 4)) | np.uint64(1)
    hdr = _check_set_proofs(glb, oracle, keys, values)
    n_ins = len({tuple(k) for k in keys.tolist()})
    assert int((hdr["fnc"] == 2).sum()) == n_ins and int((hdr["fnc"] == 1).sum()) == m - n_ins


def test_set_proofs_updates_deep_in_the_tree(glb, ctx, oracle, rng):
    """Twins sharing up to 255 path bits, inserted, then updated several times in interleaved order: update proofs walk
    the whole chain down to the leaf (siblings are not trimmed for updates)."""
    base = rand_field(rng, (6, 4))
    keys = []
    for j, k in enumerate(base):
        k2 = k.copy()
        bit = (3, 64, 130, 200, 254, 255)[j]
        k2[bit >> 6] ^= np.uint64(1) << np.uint64(bit & 63)
        if int(k2[bit >> 6]) < P:
            keys += [k, k2]
        else:
            keys += [k]
    keys = np.array(keys)
    order = np.concatenate([np.arange(len(keys)), rng.permutation(len(keys)), rng.permutation(len(keys))[:7], np.arange(len(keys))[::-1]])
    ks = keys[order].copy()
    vs = rand_field(rng, ks.shape) | np.uint64(1)
    hdr = _check_set_proofs(glb, oracle, ks, vs)
    assert (hdr["fnc"][len(keys):] == 1).all()


def test_set_proofs_at_scale_verify(glb, ctx, rng):
    """2^17 inserts: every emitted proof passes the batch verifier and the chain of roots is consistent."""
    m = 1 << 17
    keys, values = rand_field(rng, (m, 4)), rand_field(rng, (m, 4)) | np.uint64(1)
    hdr, pool, off = glb.host.smt_set_proofs(keys, values)
    assert not glb.smt_check_process_proofs(hdr, pool, off).any()
    assert np.array_equal(hdr["old_root"][1:], hdr["new_root"][:-1]) and not hdr["old_root"][0].any()
    assert np.array_equal(hdr["new_root"][-1], glb.host.smt_build_tree(keys, values))
    ns = np.diff(off)
    assert 14 < ns[m // 2:].mean() < 20          # ~ log2 of the tree size at insertion time


def test_set_proofs_mixed_at_scale_verify(glb, ctx, rng):
    """2^16 inserts followed by 2^16 updates of random existing keys: the verifier accepts every proof, the roots chain,
    and the final root is the bulk-built tree over the last value of every key."""
    m = 1 << 16
    keys0, values0 = rand_field(rng, (m, 4)), rand_field(rng, (m, 4)) | np.uint64(1)
    pick = rng.integers(0, m, m)
    keys = np.concatenate([keys0, keys0[pick]])
    values = np.concatenate([values0, rand_field(rng, (m, 4)) | np.uint64(1)])
    hdr, pool, off = glb.host.smt_set_proofs(keys, values)
    assert not glb.smt_check_process_proofs(hdr, pool, off).any()
    assert np.array_equal(hdr["old_root"][1:], hdr["new_root"][:-1])
    assert (hdr["fnc"][:m] == 2).all() and (hdr["fnc"][m:] == 1).all()
    final = values0.copy()
    for i, p_ in enumerate(pick):
        final[p_] = values[m + i]
    assert np.array_equal(hdr["new_root"][-1], glb.host.smt_build_tree(keys0, final))


@pytest.mark.parametrize("m,distinct,p_zero,seed", [(6, 2, 0.5, 1), (60, 5, 0.4, 2), (400, 40, 0.3, 3), (1500, 300, 0.25, 4),
                                                      (300, 300, 0.5, 5), (50, 1, 0.5, 6)])
def test_set_proofs_with_removals_match_sequential_sets(glb, ctx, oracle, m, distinct, p_zero, seed):
    """Arbitrary `set` sequences: a zero value removes (ProcessDelete when the key is there, ProcessNoOp when it is not),
    keys come back after having been removed, subtrees collapse to a hoisted leaf and grow again
    (src/smt/tree.rs:143-155, remove :389-586)."""
    rng = np.random.default_rng(seed)
    pool_keys = rand_field(rng, (distinct, 4))
    keys = pool_keys[rng.integers(0, distinct, m)]
    values = rand_field(rng, (m, 4)) | np.uint64(1)
    values[rng.random(m) < p_zero] = 0
    hdr = _check_set_proofs(glb, oracle, keys, values)
    assert set(hdr["fnc"].tolist()) >= ({0, 1, 2, 3} if m >= 60 and distinct > 1 else set())


def test_set_proofs_removals_between_twins(glb, ctx, oracle, rng):
    """Twins sharing long prefixes: removing one hoists the other up the whole chain; removing a key whose neighbour is
    an internal node leaves a one-child node behind."""
    base = rand_field(rng, (5, 4))
    ks = []
    for j, k in enumerate(base):
        for bit in ((2, 70, 128, 250, 33)[j], (9, 71, 190, 254, 40)[j]):
            k2 = k.copy()
            k2[bit >> 6] ^= np.uint64(1) << np.uint64(bit & 63)
            if int(k2[bit >> 6]) < P:
                ks.append(k2)
        ks.append(k)
    ks = np.array(ks)
    n = len(ks)
    v = rand_field(rng, (n, 4)) | np.uint64(1)
    z = np.zeros((n, 4), dtype=np.uint64)
    perm = rng.permutation(n)
    keys = np.concatenate([ks, ks[perm], ks[perm][::-1], ks[::2], ks[::2]])
    values = np.concatenate([v, z, v, z[::2], z[::2]])           # insert all, remove all, insert again, remove half, no-ops
    hdr = _check_set_proofs(glb, oracle, keys, values)
    assert (hdr["fnc"][n:2 * n] == 3).all() and (hdr["fnc"][-len(ks[::2]):] == 0).all()
    assert not hdr["new_root"][2 * n - 1].any()                  # the tree was empty in between


# ---- tree.find over a batch of queries (gl_smt_find_batch) -----------------------------------------------------------
def _check_find(glb, oracle, keys, values, queries):
    t = oracle.Smt()
    for k, v in zip(keys, values):
        t.set(k, v)
    hdr, pool, off = glb.host.smt_find_batch(keys, values, queries)
    assert hdr.itemsize == 168 and off.shape == (len(queries) + 1,) and int(off[0]) == 0 and int(off[-1]) == pool.shape[0]
    root = t.root()
    for i, q in enumerate(queries):
        want = t.find(q)
        h = hdr[i]
        assert np.array_equal(h["root"], root) and np.array_equal(h["key"], q % np.uint64(P))
        assert bool(h["found"]) == want["found"] and bool(h["is_old0"]) == want["is_old0"], i
        assert np.array_equal(pool[int(off[i]):int(off[i + 1])], want["siblings"]), i
        if want["found"]:
            assert np.array_equal(h["value"], want["value"]) and not h["not_found_key"].any() and not h["not_found_value"].any()
        else:
            assert not h["value"].any()
            assert np.array_equal(h["not_found_key"], want["not_found_key"]) and np.array_equal(h["not_found_value"], want["value"])
    return hdr


@pytest.mark.parametrize("m,nq", [(0, 3), (1, 4), (2, 5), (40, 60), (700, 500)])
def test_find_batch_matches_sequential_find(glb, ctx, oracle, rng, m, nq):
    keys, values = rand_field(rng, (m, 4)), rand_field(rng, (m, 4)) | np.uint64(1)
    present = keys[rng.integers(0, m, nq // 2)] if m else np.zeros((0, 4), dtype=np.uint64)
    queries = np.concatenate([present, rand_field(rng, (nq - len(present), 4))])
    hdr = _check_find(glb, oracle, keys, values, queries)
    assert int(hdr["found"].sum()) == len(present)


def test_find_batch_near_misses_removed_keys_and_repeats(glb, ctx, oracle, rng):
    """Queries that share long prefixes with stored keys (they end at that key's leaf or in an empty slot deep down),
    keys that were removed again or overwritten by the sets, the same query several times, and the empty tree."""
    base = rand_field(rng, (8, 4))
    near = []
    for j, k in enumerate(base):
        for bit in ((0, 5, 64, 100, 191, 254, 255, 30)[j], (1, 63, 65, 128, 200, 250, 17, 90)[j]):
            k2 = k.copy()
            k2[bit >> 6] ^= np.uint64(1) << np.uint64(bit & 63)
            if int(k2[bit >> 6]) < P:
                near.append(k2)
    near = np.array(near)
    keys = np.concatenate([base, near[::2], base[:3], base[3:5]])
    values = np.concatenate([rand_field(rng, (8 + len(near[::2]) + 3, 4)) | np.uint64(1), np.zeros((2, 4), dtype=np.uint64)])
    queries = np.concatenate([base, near, base[:2], rand_field(rng, (9, 4))])
    hdr = _check_find(glb, oracle, keys, values, queries)
    assert not hdr["found"][3] and not hdr["found"][4] and hdr["found"][0]      # keys 3, 4 were removed by the last two sets
    z = np.zeros((0, 4), dtype=np.uint64)
    h0 = _check_find(glb, oracle, z, z, base[:2])
    assert h0["is_old0"].all() and not h0["root"].any()


def test_golden_smt_sets(glb, ctx):
    """The committed fixture tests/golden/smt_sets.json (40 `set` calls on 10 keys: inserts, updates, removals, no-ops, a
    twin sharing 137 path bits; then 14 `find`s) against the device, without the oracle in the loop."""
    import json
    import os

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smt_sets.json")))
    un = lambda xs: np.array([int(x, 16) for x in xs], dtype=np.uint64)  # noqa: E731
    keys = np.stack([un(c["key"]) for c in g["calls"]])
    values = np.stack([un(c["value"]) for c in g["calls"]])
    hdr, pool, off = glb.host.smt_set_proofs(keys, values)
    for t, c in enumerate(g["calls"]):
        assert int(hdr["fnc"][t]) == c["fnc"] and int(hdr["is_old0"][t]) == c["is_old0"], t
        for f in ("old_root", "new_root", "old_key", "old_value", "new_key", "new_value"):
            assert np.array_equal(hdr[f][t], un(c[f])), (t, f)
        assert np.array_equal(pool[int(off[t]):int(off[t + 1])].reshape(-1), un(c["siblings"])), t
    assert np.array_equal(hdr["new_root"][-1], un(g["root"]))
    queries = np.stack([un(q["key"]) for q in g["finds"]])
    inc, qpool, qoff = glb.host.smt_find_batch(keys, values, queries)
    for i, q in enumerate(g["finds"]):
        assert bool(inc["found"][i]) == q["found"] and bool(inc["is_old0"][i]) == q["is_old0"], i
        assert np.array_equal(inc["root"][i], un(g["root"]))
        assert np.array_equal(qpool[int(qoff[i]):int(qoff[i + 1])].reshape(-1), un(q["siblings"])), i
        if q["found"]:
            assert np.array_equal(inc["value"][i], un(q["value"]))
        else:
            assert np.array_equal(inc["not_found_key"][i], un(q["not_found_key"])) and np.array_equal(inc["not_found_value"][i], un(q["value"]))
