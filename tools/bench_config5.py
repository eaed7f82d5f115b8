#!/usr/bin/env python
"""BASELINE config 5 as far as this path goes: 8 inner ECDSA proofs = 8 independent provers, one per
GPU ("replicas": no exchange), then the outer recursion proof.  Per proof the trace is the one bench.py's proof_trace
times: 3 commits (135 + 20 + 16 polynomials, page-able host arrays) + prove_openings over 4 oracles.  Launch with
torchrun, one rank per GPU; rank 0 prints one JSON object.  Inner proofs: 2^16 rows; outer: 2^13 rows (SURVEY's
estimates).  Witness generation and compute_quotient_polys are host stages of the reference and are not included."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
glb = importlib.import_module("plonky2-lib_b200")
fri = importlib.import_module("plonky2-lib_b200.fri")
ctx = glb.Context(local)
P = glb.host.P
COLS = (84, 135, 20, 16)


def make_trace(lg, seed):
    n = 1 << lg
    rng = np.random.default_rng(seed)
    vals = [rng.integers(0, P, size=(c, n), dtype=np.uint64) for c in COLS]
    cs = glb.PolynomialBatch.from_values(vals[0], 3, False, 4, want_coeffs=False, ctx=ctx)
    zeta = (0x123456789ABCDEF % P, 0x0FEDCBA987654321 % P)
    g = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - lg), P)
    inst = [(zeta, [(oi, pi) for oi, c in enumerate(COLS) for pi in range(c)]), ((zeta[0] * g % P, zeta[1] * g % P), [(2, pi) for pi in range(20)])]
    prm = fri.FriParams.for_degree(glb.FriConfig(), lg)

    def once():
        bs = [cs] + [glb.PolynomialBatch.from_values(v, 3, False, 4, want_coeffs=True, ctx=ctx) for v in vals[1:]]
        ch = fri.Challenger(ctx)
        for b in bs:
            ch.observe_cap(b.merkle_tree.cap)
        proof = fri.prove_openings(bs, inst, ch, prm, ctx)
        for b in bs[1:]:
            b.free()
        return proof

    return once


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


inner = make_trace(16, 100 + rank)
inner()
barrier()
best = None
for _ in range(3):
    barrier()
    t = time.perf_counter()
    inner()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    best = float(tt.item()) if best is None else min(best, float(tt.item()))
outer_ms = None
if rank == 0:
    outer = make_trace(13, 999)
    outer()
    b2 = None
    for _ in range(3):
        t = time.perf_counter()
        outer()
        dt = time.perf_counter() - t
        b2 = dt if b2 is None else min(b2, dt)
    outer_ms = b2 * 1e3
barrier()
if rank == 0:
    print(json.dumps({"inner_proofs": world, "inner_rows_log2": 16, "inner_ms_max_over_gpus": best * 1e3, "outer_rows_log2": 13, "outer_ms": outer_ms,
                      "total_ms": best * 1e3 + outer_ms, "note": "commit + opening traces only; one inner prover per GPU, no exchange"}))
if world > 1:
    dist.destroy_process_group()
