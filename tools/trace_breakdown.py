#!/usr/bin/env python
"""Where the wall clock of one small proof trace goes (3 commits + prove_openings at 2^13 .. 2^16 rows):
per-stage blocking wall-clock times.  Prints one JSON object."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

glb = importlib.import_module("plonky2-lib_b200")
fri = importlib.import_module("plonky2-lib_b200.fri")
ctx = glb.Context.default()
P = glb.host.P
cols = (84, 135, 20, 16)
zeta = (0x123456789ABCDEF % P, 0x0FEDCBA987654321 % P)


class Clock:
    def __init__(self):
        self.t = {}

    def lap(self, name, fn):
        t0 = time.perf_counter()
        r = fn()
        self.t[name] = self.t.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r


def trace(lg, reps=4):
    n = 1 << lg
    rng = np.random.default_rng(lg)
    vals = [rng.integers(0, P, size=(c, n), dtype=np.uint64) for c in cols]
    cs = glb.PolynomialBatch.from_values(vals[0], 3, False, 4, want_coeffs=False)
    g = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - lg), P)
    inst = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]),
            ((zeta[0] * g % P, zeta[1] * g % P), [(2, pi) for pi in range(20)])]
    prm = fri.FriParams.for_degree(glb.FriConfig(), lg)
    best = None
    for _ in range(reps + 1):
        ck = Clock()
        t0 = time.perf_counter()
        bs = [cs]
        for i, v in enumerate(vals[1:]):
            bs.append(ck.lap("commit_%d_cols" % cols[i + 1], lambda: glb.PolynomialBatch.from_values(v, 3, False, 4, want_coeffs=True)))
        ch = fri.Challenger()
        ck.lap("observe_caps", lambda: [ch.observe_cap(b.merkle_tree.cap) for b in bs])
        alpha = ck.lap("alpha", ch.get_extension_challenge)
        lc, lv = ck.lap("fri_final_poly", lambda: (fri.fri_final_poly(bs, inst, alpha, 3, ctx, resident=True), ctx.sync())[0])
        trees, final = ck.lap("fri_committed_trees", lambda: fri.fri_committed_trees(lc, lv, ch, prm, ctx))
        ck.lap("pow", lambda: fri.fri_proof_of_work(ch, prm.config, ctx))
        ck.lap("query_rounds", lambda: fri.fri_prover_query_rounds(bs, trees, ch, n << 3, prm))
        for t in trees:
            t.free()
        lc.free(); lv.free()
        for b in bs[1:]:
            b.free()
        ck.t["total"] = (time.perf_counter() - t0) * 1e3
        if best is None or ck.t["total"] < best["total"]:
            best = ck.t
    # the same trace with the one-call prover (gl_fri_prove)
    one = None
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        bs = [cs] + [glb.PolynomialBatch.from_values(v, 3, False, 4, want_coeffs=True) for v in vals[1:]]
        t1 = time.perf_counter()
        ch = fri.Challenger()
        for b in bs:
            ch.observe_cap(b.merkle_tree.cap)
        t2 = time.perf_counter()
        fri.prove_openings_device(bs, inst, ch, prm, ctx, flat=True)
        t3 = time.perf_counter()
        for b in bs[1:]:
            b.free()
        cur = {"one_call_commits": (t1 - t0) * 1e3, "one_call_observe_caps": (t2 - t1) * 1e3, "one_call_gl_fri_prove": (t3 - t2) * 1e3,
               "one_call_total": (time.perf_counter() - t0) * 1e3}
        if one is None or cur["one_call_total"] < one["one_call_total"]:
            one = cur
    best.update(one)
    cs.free()
    # the Challenger's own cost: single-state permutations through the C ABI
    st = np.zeros(12, dtype=np.uint64)
    t0 = time.perf_counter()
    for _ in range(200):
        st = glb.PoseidonHash.permute(st, ctx=ctx)
    best["one_challenger_permutation_us"] = (time.perf_counter() - t0) / 200 * 1e6
    return {k: round(v, 3) for k, v in best.items()}


print(json.dumps({"2^%d" % lg: trace(lg) for lg in [int(x) for x in (sys.argv[1:] or ["13", "15", "16"])]}))
