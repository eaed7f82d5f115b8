#!/usr/bin/env python
"""Mirror mode: gl_commit_download of the row-major leaves + digests into host memory."""
import ctypes as C
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
ctx = glb.Context.default()
lib, N = ctx._lib, glb._native
out = {}
for lg in (16, 18, 20):
    c, n = 135, 1 << lg
    v = torch.randint(0, 2**62, (c, n), dtype=torch.int64, device="cuda")
    b = glb.PolynomialBatch.from_values(v, 3, False, 4, want_coeffs=False)
    NL = n << 3
    nbytes = NL * c * 8 + 2 * (NL - 16) * 32
    r = {"gbytes": round(nbytes / 1e9, 2)}
    for kind in ("pageable", "pinned"):
        if kind == "pinned":
            leaves, dig = glb.pinned_empty((NL, c)), glb.pinned_empty((2 * (NL - 16), 4))
        else:
            leaves, dig = np.empty((NL, c), dtype=np.uint64), np.empty((2 * (NL - 16), 4), dtype=np.uint64)
        for rep in ("first", "again"):
            t = time.perf_counter()
            ctx.check(lib.gl_commit_download(b._h, leaves.ctypes.data, dig.ctypes.data, N.GL_HOST))
            dt = time.perf_counter() - t
            r["%s_%s_ms" % (kind, rep)] = round(dt * 1e3, 1)
            r["%s_%s_GBps" % (kind, rep)] = round(nbytes / dt / 1e9, 1)
        del leaves, dig
    b.free()
    del v
    out["2^%d x %d" % (lg, c)] = r
print(json.dumps(out))
