#!/usr/bin/env python
"""gl_copy between page-able host memory and HBM: staged (csrc/host_staging.cu) or handed to the driver
(GL_B200_STAGING=0).  Fresh destination pages (first touch) and warm ones are timed separately."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
glb = importlib.import_module("plonky2-lib_b200")
ctx = glb.Context.default()
lib, N = ctx._lib, glb._native
out = {"GL_B200_STAGING": os.environ.get("GL_B200_STAGING", "1"), "GL_B200_HOST_THREADS": os.environ.get("GL_B200_HOST_THREADS", "default"),
       "cpus": os.cpu_count()}
for mb in (8, 64, 256, 1024):
    n = mb * (1 << 20) // 8
    src = np.arange(n, dtype=np.uint64)
    d = glb.DeviceBuffer((n,), ctx)
    d.from_host(src)
    t = time.perf_counter(); d.from_host(src); h2d = time.perf_counter() - t
    t = time.perf_counter()
    fresh = np.empty(n, dtype=np.uint64)
    ctx.check(lib.gl_copy(ctx._h, fresh.ctypes.data, N.GL_HOST, d.ptr, N.GL_DEVICE, n * 8))
    d2h_fresh = time.perf_counter() - t
    t = time.perf_counter()
    ctx.check(lib.gl_copy(ctx._h, fresh.ctypes.data, N.GL_HOST, d.ptr, N.GL_DEVICE, n * 8))
    d2h_warm = time.perf_counter() - t
    assert np.array_equal(fresh, src)
    d.free()
    out["%d MB" % mb] = {"h2d_GBps": round(n * 8 / h2d / 1e9, 1), "d2h_fresh_pages_GBps": round(n * 8 / d2h_fresh / 1e9, 1),
                         "d2h_warm_pages_GBps": round(n * 8 / d2h_warm / 1e9, 1)}
print(json.dumps(out))
