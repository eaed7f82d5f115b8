"""The N > 1 host logic on CPU: world_size 2 over gloo.  Each rank runs the plan of
plonky2-lib_b200/parallel.py with the ORACLE standing in for the device (the oracle is test infrastructure;
the real device path of the same plan is bench.py --gpus N and tests/test_gpu_parity.py::test_sharded_*)."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, cols, lg_n, rate_bits, cap_height, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import pyoracle as o

    par = importlib.import_module("plonky2-lib_b200.parallel")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1 << lg_n
        N = n << rate_bits
        par.check_shardable(world, rate_bits, cap_height)
        values = o.synthetic_values(cols, n)
        c0, c1, per = par.column_slice(rank, world, cols)
        sl = np.zeros((per, n), dtype=np.uint64)
        for j in range(c0, c1):
            sl[j - c0] = o.ifft(values[j])
        slice_buf = torch.from_numpy(sl.view(np.int64))
        gathered = torch.empty((world * per, n), dtype=torch.int64)
        coeffs = par.all_gather_coefficients(dist, slice_buf, gathered, cols).numpy().view(np.uint64)
        whole = o.commit_from_coeffs(np.ascontiguousarray(coeffs), rate_bits, cap_height)
        lo, hi = par.leaf_range(rank, world, N)
        k0, k1 = par.cap_range(rank, world, cap_height)
        # this rank's subtrees, rebuilt from its own leaf block only
        _, local_cap = o.merkle_tree(whole["leaves"][lo:hi], cap_height - (world.bit_length() - 1))
        assert np.array_equal(local_cap, whole["cap"][k0:k1])
        cap_all = torch.empty((1 << cap_height, 4), dtype=torch.int64)
        par.all_gather_cap(dist, torch.from_numpy(local_cap.view(np.int64).copy()), cap_all)
        ref = o.commit_from_values(values, rate_bits, cap_height, want_leaves=False)
        ok = np.array_equal(cap_all.numpy().view(np.uint64), ref["cap"]) and np.array_equal(coeffs, ref["coeffs"])
        owners = [par.owner_of_leaf(i, world, N) for i in (0, lo, hi - 1, N - 1)]
        q.put((rank, bool(ok), owners, (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cols", [5, 8])
def test_sharded_commit_plan_world2(cols):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200 + cols
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cols, 6, 3, 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    N = 64 << 3
    assert all(r[1] for r in res)
    assert res[0][3] == (0, N // 2) and res[1][3] == (N // 2, N)
    assert res[0][2] == [0, 0, 0, 1] and res[1][2] == [0, 1, 1, 1]


def _worker_interleaved(rank, world, port, cols, lg_n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import pyoracle as o

    par = importlib.import_module("plonky2-lib_b200.parallel")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1 << lg_n
        values = o.synthetic_values(cols, n)
        G, rounds, mine = par.round_blocks(rank, world, cols, target_rounds=2)
        sl = np.zeros((rounds, G, n), dtype=np.uint64)
        for j, (c0, c1) in enumerate(mine):
            for col in range(c0, c1):
                sl[j, col - c0] = o.ifft(values[col])
        slice_buf = torch.from_numpy(sl.view(np.int64))
        coeffs = np.zeros((cols, n), dtype=np.uint64)
        for j in range(rounds):            # round j: G columns from every rank = world * G consecutive global columns
            stage = torch.empty((world * G, n), dtype=torch.int64)
            dist.all_gather_into_tensor(stage, slice_buf[j])
            c0 = j * world * G
            nc = min(world * G, cols - c0)
            coeffs[c0: c0 + nc] = stage.numpy().view(np.uint64)[:nc]
        ref = o.commit_from_values(values, 3, 4, want_leaves=False)
        q.put((rank, bool(np.array_equal(coeffs, ref["coeffs"]))))
    finally:
        dist.destroy_process_group()


def test_interleaved_column_plan_world2():
    """The pipelined plan of bench.py --gpus N: round j gathers G columns per rank into consecutive global columns."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_worker_interleaved, args=(r, 2, port, 7, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_plan_helpers():
    par = importlib.import_module("plonky2-lib_b200.parallel")
    G, rounds, mine = par.round_blocks(1, 8, 135)
    assert (G, rounds) == (4, 5) and mine[0] == (4, 8) and mine[4] == (132, 135)
    assert par.round_blocks(7, 8, 135)[2][4] == (135, 135)            # nothing left for the last rank in the last round
    assert sum(c1 - c0 for r in range(8) for c0, c1 in par.round_blocks(r, 8, 135)[2]) == 135
    par = importlib.import_module("plonky2-lib_b200.parallel")
    assert par.column_slice(0, 8, 135) == (0, 17, 17)
    assert par.column_slice(7, 8, 135) == (119, 135, 17)
    assert sum(par.column_slice(r, 8, 135)[1] - par.column_slice(r, 8, 135)[0] for r in range(8)) == 135
    assert par.column_slice(3, 4, 2) == (2, 2, 1)  # more ranks than columns: empty slice
    assert par.cap_range(3, 8, 4) == (6, 8)
    with pytest.raises(ValueError):
        par.check_shardable(16, 3, 4)
    with pytest.raises(ValueError):
        par.check_shardable(3, 3, 4)
    assert [par.batch_range(r, 4, 10) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [par.batch_range(r, 8, 3) for r in range(8)][2:5] == [(2, 3), (3, 3), (3, 3)]


def _worker_smt(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import pyoracle as o

    par = importlib.import_module("plonky2-lib_b200.parallel")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)            # same batch on every rank
        t = o.Smt()
        recs = []
        keys = [rng.integers(0, 2**63, size=4, dtype=np.uint64) for _ in range(11)]
        for k in keys:
            recs.append(t.set(k, rng.integers(1, 2**63, size=4, dtype=np.uint64)))
        for k in keys[::2]:
            recs.append(t.set(k, np.zeros(4, dtype=np.uint64)))
        recs = np.array(recs, dtype=o.SMT_PROOF_DTYPE)
        recs["new_root"][5][0] ^= np.uint64(1)    # one bad proof, owned by exactly one rank
        m = recs.shape[0]
        off = np.zeros(m + 1, dtype=np.uint64)
        off[1:] = np.cumsum(recs["num_siblings"])
        pool = np.concatenate([r["siblings"][: r["num_siblings"]] for r in recs])
        lo, hi = par.batch_range(rank, world, m)
        hd, pl, of = par.slice_proof_batch(recs, pool, off, lo, hi)
        # the share is self-contained: rebuilding the proofs of the share from the sliced arrays gives them back
        ok = int(of[0]) == 0 and int(of[-1]) == pl.shape[0]
        for i in range(hi - lo):
            sib = pl[int(of[i]):int(of[i + 1])]
            ok = ok and np.array_equal(sib, recs[lo + i]["siblings"][: recs[lo + i]["num_siblings"]])
        status = o.smt_verify_process_batch(np.ascontiguousarray(hd))       # the oracle stands in for the device
        # no data-path collective; the verdict is one scalar
        bad = torch.tensor([int((status != 0).sum())])
        dist.all_reduce(bad)
        q.put((rank, bool(ok), (lo, hi), int(bad.item()), [int(x) for x in np.nonzero(status)[0] + lo]))
    finally:
        dist.destroy_process_group()


def test_smt_proof_batch_split_world2():
    """process proofs are independent, so a batch is split evenly with no exchange of proof data."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 100
    procs = [ctx.Process(target=_worker_smt, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2][1] == res[1][2][0] and res[0][2][0] == 0      # contiguous shares covering the batch
    assert res[0][3] == res[1][3] == 1                             # both ranks learn there is one bad proof
    assert res[0][4] + res[1][4] == [5]                            # and its owner knows which
