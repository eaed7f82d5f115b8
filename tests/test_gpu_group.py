"""gl_group: one commit sharded over ranks behind the C ABI -- column-sharded IFFT, NCCL all-gather of the
coefficients, coset-sharded LDE + subtrees, NCCL all-gather of the cap, NCCL exchange of query openings -- against the
oracle's unsharded commit.  The 2-rank cases need two GPUs (run with `gpurun --gpus 2`); the 1-rank group runs the same
code path without collectives on any box."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


def _check_against_oracle(glb, oracle, group, lg_n, c, r, h, device_inputs=False, stream_hash=False):
    values = oracle.synthetic_values(c, 1 << lg_n)
    want = oracle.commit_from_values(values, r, h)
    if device_inputs:
        import torch

        inp = [torch.from_numpy(values.view(np.int64)).to(f"cuda:{cx.device}") for cx in group.ctxs]
    else:
        inp = values
    b = group.commit(inp, r, h, stream_hash=stream_hash)
    for cap in b.caps:                      # the WHOLE cap on every local rank
        assert np.array_equal(cap, want["cap"])
    if device_inputs:
        for co in b.coeffs:                 # every rank holds all coefficients
            assert np.array_equal(co.cpu().numpy().view(np.uint64), want["coeffs"])
    elif group.nlocal == group.nranks:      # one process holds every rank: the shared host array is complete
        assert np.array_equal(b.coeffs, want["coeffs"])
    N = (1 << lg_n) << r
    idx = sorted({0, 1, N // 2 - 1, N // 2, N - 1, N // 8, 3 * N // 8 + 5, (0x9E3779B97F4A7C15 % N)})
    rows, paths = b.open(idx)               # global indices, owned by different ranks
    assert np.array_equal(rows, want["leaves"][idx])
    for q, i in enumerate(idx):
        assert np.array_equal(paths[q], oracle.merkle_prove(want["digests"], N, h, i))
        assert oracle.merkle_verify(rows[q], i, paths[q], want["cap"], h)
    for r_, p_ in zip(b.all_rows, b.all_paths):
        assert np.array_equal(r_, rows) and np.array_equal(p_, paths)
    b.free()


@pytest.mark.parametrize("lg_n,c", [(10, 135), (12, 20), (8, 3), (13, 136)])
def test_one_rank_group_equals_oracle(glb, oracle, lg_n, c):
    g = glb.Group.local([0])
    _check_against_oracle(glb, oracle, g, lg_n, c, 3, 4)
    _check_against_oracle(glb, oracle, g, lg_n, c, 3, 4, device_inputs=True)
    g.close()


@pytest.mark.parametrize("lg_n,c,r,h", [(10, 135, 3, 4), (12, 20, 3, 4), (9, 7, 1, 1), (14, 135, 3, 4), (16, 16, 3, 4)])
def test_two_rank_group_in_one_process(glb, oracle, lg_n, c, r, h):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    g = glb.Group.local([0, 1])
    assert g.nccl_version > 0
    _check_against_oracle(glb, oracle, g, lg_n, c, r, h)
    _check_against_oracle(glb, oracle, g, lg_n, c, r, h, device_inputs=True)
    _check_against_oracle(glb, oracle, g, lg_n, c, r, h, stream_hash=True)
    g.close()


@pytest.mark.parametrize("nranks", [1, 2])
def test_group_commit_with_blinding(glb, oracle, nranks):
    """GL_COMMIT_BLINDING over a group: every rank salts its own leaves; the gathered cap is the oracle's MerkleTree::new over
    the salted leaves as opened through gl_group_commit_open (every leaf, so the exchange is exercised for every owner)."""
    if _ngpu() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    lg_n, c, r, h = 8, 11, 2, 3
    g = glb.Group.local(list(range(nranks)))
    values = oracle.synthetic_values(c, 1 << lg_n, seed=5)
    plain = oracle.commit_from_values(values, r, h)
    b = g.commit(values, r, h, blinding=True)
    N = (1 << lg_n) << r
    rows, paths = b.open(list(range(N)))
    assert rows.shape == (N, c + 4) and b.leaf_len == c + 4
    assert np.array_equal(rows[:, :c], plain["leaves"])
    assert (rows[:, c:] < np.uint64(0xFFFFFFFF00000001)).all() and len(np.unique(rows[:, c:])) > 0.99 * 4 * N
    digests, cap = oracle.merkle_tree(rows, h)
    for cp in b.caps:
        assert np.array_equal(cp, cap)
    for i in (0, N // 2 - 1, N // 2, N - 1):
        assert np.array_equal(paths[i], oracle.merkle_prove(digests, N, h, i))
    b.free()
    g.close()


def test_group_rejects_bad_geometry(glb):
    g = glb.Group.local([0])
    with pytest.raises(glb.GlPanic):
        g.commit(np.zeros((3, 12), dtype=np.uint64), 3, 4)          # log2_strict
    with pytest.raises(glb.GlPanic):
        g.commit(np.zeros((3, 16), dtype=np.uint64), 3, 9)          # cap_height > log2(leaves)
    g.close()
    c0 = glb.Context(0)
    with pytest.raises(glb.GlPanic):
        glb.Group([c0, glb.Context(0)], 0, 2, None)                 # two ranks on one GPU
    with pytest.raises(glb.GlPanic):
        glb.Group([c0], 0, 2, None)                                 # spans processes: needs the token


@pytest.mark.parametrize("world", [2, 4])
def test_ranks_in_separate_processes(world):
    """One rank per process (torchrun, as bench.py --gpus N runs): token over torch.distributed, commit + openings checked
    against the oracle inside every rank (tests/mp_group_worker.py)."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mp_group_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("group worker ok") == world


def test_full_size_commit_over_two_ranks_matches_golden(glb, oracle):
    """BASELINE config 2 (2^20 x 135) sharded over 2 GPUs: the gathered cap and sampled openings against the full-size
    fixture the oracle produced (tests/golden/commit_fullsize.json)."""
    import json

    import torch

    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    case = next(c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "commit_fullsize.json")))["cases"] if c["lg_n"] == 20)
    hx = lambda a: [f"{int(x):016x}" for x in np.asarray(a).reshape(-1)]  # noqa: E731
    values = oracle.synthetic_values(case["c"], 1 << 20)
    g = glb.Group.local([0, 1])
    inp = [torch.from_numpy(values.view(np.int64)).to(f"cuda:{d}") for d in (0, 1)]
    b = g.commit(inp, 3, 4, want_coeffs=False)
    assert hx(b.cap) == case["cap"]
    rows, paths = b.open(case["leaf_indices"])
    assert [hx(x) for x in rows] == case["leaf_rows"] and [hx(x) for x in paths] == case["leaf_paths"]
    b.free()
    g.close()


def test_sharded_prove_openings_equals_single_gpu_and_verifies(glb, oracle, rng_seed=11):
    """E2: prove_openings over oracles sharded on 2 GPUs -- the query openings come from the owning rank over
    NCCL -- gives the same proof as the oracle's prover and passes the oracle's restatement of the upstream verifier."""
    import importlib

    from oracle import fri_oracle as fo

    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    fri = importlib.import_module("plonky2-lib_b200.fri")
    rng = np.random.default_rng(rng_seed)
    degree_bits, cols = 12, (84, 135, 20, 16)
    rate_bits, cap_height, pow_bits, rounds = 3, 4, 10, 28
    n = 1 << degree_bits
    g = glb.Group.local([0, 1])
    sharded, polys, trees = [], [], []
    for k, c in enumerate(cols):
        v = oracle.synthetic_values(c, n, seed=700 + k)
        sharded.append(g.commit(v, rate_bits, cap_height))
        res = oracle.commit_from_values(v, rate_bits, cap_height)
        polys.append(res["coeffs"])
        trees.append(fo.MerkleTree(res["leaves"], cap_height))
        assert np.array_equal(sharded[-1].cap, trees[-1].cap)
    zeta = tuple(int(x) for x in rng.integers(0, 0xFFFFFFFF00000001, 2, dtype=np.uint64))
    gen = oracle.lib().glo_primitive_root_of_unity(degree_bits)
    instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]), (fo.ext_scalar(zeta, gen), [(2, 0), (2, 1)])]
    cfg = glb.FriConfig(rate_bits=rate_bits, cap_height=cap_height, proof_of_work_bits=pow_bits, num_query_rounds=rounds)
    params = fri.FriParams.for_degree(cfg, degree_bits)
    ch, och, vch = fri.Challenger(g.ctxs[0]), fo.Challenger(), fo.Challenger()
    for t in trees:
        for c_ in (ch, och, vch):
            c_.observe_cap(t.cap)
    got = fri.prove_openings(sharded, instance, ch, params)
    want = fo.prove_openings(polys, trees, instance, och, degree_bits, rate_bits, cap_height, pow_bits, rounds)
    assert got["pow_witness"] == want["pow_witness"] and np.array_equal(got["final_poly"], want["final_poly"])
    owners = set()
    for ra, rb in zip(got["query_round_proofs"], want["query_round_proofs"]):
        assert ra["x_index"] == rb["x_index"]
        owners.add(ra["x_index"] >= (n << rate_bits) // 2)
        for (rowa, patha), (rowb, pathb) in zip(ra["initial_trees_proof"], rb["initial_trees_proof"]):
            assert np.array_equal(rowa, rowb) and np.array_equal(patha, pathb)
        for sa, sb in zip(ra["steps"], rb["steps"]):
            assert np.array_equal(sa["evals"], sb["evals"]) and np.array_equal(sa["merkle_proof"], sb["merkle_proof"])
    assert owners == {False, True}            # the 28 queries hit leaves of both ranks
    openings = fo.opening_set(polys, instance)
    assert [[tuple(int(x) for x in v) for v in b] for b in fri.opening_set(sharded, instance)] == \
           [[tuple(int(x) for x in v) for v in b] for b in openings]
    assert fo.verify_openings(got, openings, [t.cap for t in trees], instance, vch, degree_bits, rate_bits, cap_height, pow_bits, rounds)
    for b in sharded:
        b.free()
    g.close()
