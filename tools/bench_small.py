import importlib, sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
glb = importlib.import_module("plonky2-lib_b200")
import ctypes as C
ctx = glb.Context(0); lib, N = ctx._lib, glb._native
dev = torch.device("cuda", 0)
res = {}
for lg in (10, 12, 13, 14, 16):
    c = 135
    v = torch.randint(0, 2**62, (c, 1 << lg), dtype=torch.int64, device=dev)
    co = torch.empty_like(v); cap = torch.zeros((16, 4), dtype=torch.int64, device=dev)
    def step():
        h = C.c_void_p()
        ctx.check(lib.gl_commit_from_values(ctx._h, v.data_ptr(), lg, c, 3, 4, co.data_ptr(), cap.data_ptr(), C.byref(h), N.GL_DEVICE))
        lib.gl_commit_free(h)
    for _ in range(5): step()
    torch.cuda.synchronize()
    l0 = ctx.kernel_launches
    t = time.perf_counter()
    for _ in range(20): step()
    dt = (time.perf_counter() - t) / 20
    res[lg] = {"ms": dt * 1e3, "launches": (ctx.kernel_launches - l0) // 20, "cells_per_s": c * (1 << lg) / dt}
print(json.dumps(res))
