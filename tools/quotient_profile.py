"""gl_quotient_polys at standard_recursion_config geometry (2^lg rows, default 18): wall clock of three calls; run it
under `ncu --metrics gpu__time_duration.sum -k regex:k_quotient` for the kernel's share."""
import importlib, os, sys, time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
glb = importlib.import_module("plonky2-lib_b200")
ctx = glb.Context.default()
P = glb.host.P
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n_ = 1 << lg
def rand_dev(shape):
    return (torch.randint(0, 2**62, shape, dtype=torch.int64, device='cuda') % (P - 2**63) )
circuit = (lg, 135, 80, 4, 2, 2, 8, 6)
gates = [(0, 0, 0, 0, 3), (1, 2, 0, 0, 3), (2, 0, 0, 0, 3), (3, 3, 1, 3, 6), (4, 2, 1, 3, 6), (5, 2, 1, 3, 6)]
k_is = np.array([pow(7, j, P) for j in range(80)], dtype=np.uint64)
bs = [glb.PolynomialBatch.from_values(rand_dev((c, n_)), 3, False, 4, want_coeffs=False) for c in (84, 135, 20)]
ch = np.array([3, 5], dtype=np.uint64)
pih = np.arange(4, dtype=np.uint64)
for _ in range(3):
    t = time.perf_counter()
    glb.host.compute_quotient_polys(circuit, gates, k_is, *bs, pih, ch, ch + np.uint64(9), ch + np.uint64(77))
    print("call ms", (time.perf_counter() - t) * 1e3)
