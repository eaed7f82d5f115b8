#!/usr/bin/env python
"""Wall clock of one gl_commit_from_values by where the buffers live (device / page-locked / page-able host,
flat or one array per polynomial) next to the device-side phase times: what the host path costs at small sizes."""
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
P = glb.host.P
rng = np.random.default_rng(1)


def best(fn, reps=6):
    fn()
    b = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        b = min(b, time.perf_counter() - t)
    return round(b * 1e3, 3)


out = {}
for lg, c in ((13, 16), (13, 135), (16, 20), (16, 135), (18, 135)):
    n = 1 << lg
    v = rng.integers(0, P, size=(c, n), dtype=np.uint64)
    co = np.empty_like(v)
    co[:] = 0
    cap = np.zeros((16, 4), dtype=np.uint64)
    pv, pc = glb.pinned_empty((c, n)), glb.pinned_empty((c, n))
    pv[:] = v
    dv = torch.from_numpy(v.view(np.int64)).cuda()
    dc = torch.empty_like(dv)
    dcap = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    cols = [v[j].copy() for j in range(c)]
    ocols = [np.zeros(n, dtype=np.uint64) for _ in range(c)]
    ip = (C.c_void_p * c)(*[a.ctypes.data for a in cols])
    op = (C.c_void_p * c)(*[a.ctypes.data for a in ocols])

    def call(inp, outp, capp, space):
        h = C.c_void_p()
        ctx.check(lib.gl_commit_from_values(ctx._h, inp, lg, c, 3, 4, outp, capp, C.byref(h), space))
        lib.gl_commit_free(h)

    def call_cols():
        h = C.c_void_p()
        ctx.check(lib.gl_commit_from_values_cols(ctx._h, ip, lg, c, 3, 4, op, cap.ctypes.data, C.byref(h)))
        lib.gl_commit_free(h)

    r = {}
    r["device_ms"] = best(lambda: call(dv.data_ptr(), dc.data_ptr(), dcap.data_ptr(), N.GL_DEVICE))
    ph = (C.c_float * 6)()
    lib.gl_ctx_commit_phase_ms(ctx._h, ph)
    r["device_phases_ms"] = [round(x, 3) for x in ph]
    r["pinned_ms"] = best(lambda: call(pv.ctypes.data, pc.ctypes.data, cap.ctypes.data, N.GL_HOST))
    r["pinned_no_coeffs_ms"] = best(lambda: call(pv.ctypes.data, None, cap.ctypes.data, N.GL_HOST))
    r["pageable_ms"] = best(lambda: call(v.ctypes.data, co.ctypes.data, cap.ctypes.data, N.GL_HOST))
    r["pageable_no_coeffs_ms"] = best(lambda: call(v.ctypes.data, None, cap.ctypes.data, N.GL_HOST))
    r["pageable_cols_ms"] = best(call_cols)
    l0 = ctx.kernel_launches
    call(dv.data_ptr(), dc.data_ptr(), dcap.data_ptr(), N.GL_DEVICE)
    r["launches"] = ctx.kernel_launches - l0
    r["mbytes_each_way"] = round(v.nbytes / 1e6, 1)
    out["2^%d x %d" % (lg, c)] = r
print(json.dumps(out))
