"""Host-buffer paths of the commit (SURVEY.md 8b "Host buffers may be pageable; library pins/stages internally"):
page-able arrays go through the page-locked slot rings and helper threads of csrc/host_staging.cu, page-locked
ones go straight to the DMA engine, device-resident ones are not copied.  All three, and the one-array-per-
polynomial entry points (Vec<PolynomialValues<F>> as the reference holds it), must give identical bytes."""
import ctypes as C

import numpy as np
import pytest

from conftest import P, adversarial_columns, rand_field

pytestmark = pytest.mark.gpu


def _caps_coeffs(b):
    polys = b.polynomials
    if isinstance(polys, list):
        polys = np.stack(polys)
    return b.merkle_tree.cap.copy(), np.array(polys, copy=True)


# sizes around the ring geometry (4 slots of 4 MB): below the 1 MB staging threshold, one partial slot, several
# slots with a ragged tail, columns larger than a slot, many tiny columns
STAGING_CASES = [(10, 7), (12, 33), (14, 40), (15, 135), (16, 21), (20, 3), (8, 700), (19, 5)]


@pytest.mark.parametrize("lg_n,c", STAGING_CASES)
def test_pageable_pinned_device_and_per_column_inputs_agree(glb, ctx, oracle, lg_n, c):
    import torch

    n = 1 << lg_n
    values = oracle.synthetic_values(c, n)
    if c >= 7:
        values[:5] = adversarial_columns(n)
    # device-resident input is the reference point (no host copies at all)
    vd = torch.from_numpy(values.view(np.int64)).cuda()
    bd = glb.PolynomialBatch.from_values(vd, 3, False, 4)
    cap_d = bd.merkle_tree.cap.copy()
    coeffs_d = bd.polynomials.cpu().numpy().view(np.uint64)
    idx = [0, 1, (n << 3) - 1, (n << 3) // 3]
    rows_d, paths_d = bd.open(idx)
    bd.free()
    if lg_n <= 12:
        want = oracle.commit_from_values(values, 3, 4)
        assert np.array_equal(cap_d, want["cap"]) and np.array_equal(coeffs_d, want["coeffs"])

    # page-able flat array
    b = glb.PolynomialBatch.from_values(values, 3, False, 4)
    cap, coeffs = _caps_coeffs(b)
    rows, paths = b.open(idx)
    assert np.array_equal(cap, cap_d) and np.array_equal(coeffs, coeffs_d)
    assert np.array_equal(rows, rows_d) and np.array_equal(paths, paths_d)
    # gl_commit_coeffs into a page-able array (copy_out through the downloader)
    again = np.empty_like(values)
    ctx.check(ctx._lib.gl_commit_coeffs(b._h, again.ctypes.data, glb._native.GL_HOST))
    assert np.array_equal(again, coeffs_d)
    b.free()

    # page-locked flat array in and out
    pin = glb.pinned_empty((c, n))
    pin[:] = values
    pout = glb.pinned_empty((c, n))
    capb = np.zeros((16, 4), dtype=np.uint64)
    h = C.c_void_p()
    ctx.check(ctx._lib.gl_commit_from_values(ctx._h, pin.ctypes.data, lg_n, c, 3, 4, pout.ctypes.data, capb.ctypes.data,
                                             C.byref(h), glb._native.GL_HOST))
    ctx._lib.gl_commit_free(h)
    assert np.array_equal(capb, cap_d) and np.array_equal(pout, coeffs_d)

    # one array per polynomial (separately allocated, so not contiguous with each other)
    cols = [values[j].copy() for j in range(c)]
    b = glb.PolynomialBatch.from_values(cols, 3, False, 4)
    cap, coeffs = _caps_coeffs(b)
    assert isinstance(b.polynomials, list) and len(b.polynomials) == c
    assert np.array_equal(cap, cap_d) and np.array_equal(coeffs, coeffs_d)
    b.free()

    # from_coeffs, per polynomial
    ccols = [coeffs_d[j].copy() for j in range(c)]
    b = glb.PolynomialBatch.from_coeffs(ccols, 3, False, 4)
    assert np.array_equal(b.merkle_tree.cap, cap_d)
    rows, paths = b.open(idx)
    assert np.array_equal(rows, rows_d) and np.array_equal(paths, paths_d)
    b.free()


def test_per_column_entry_points_reject_null_pointers(glb, ctx):
    n, c = 1 << 8, 4
    cols = [np.zeros(n, dtype=np.uint64) for _ in range(c)]
    ptrs = (C.c_void_p * c)(*[a.ctypes.data for a in cols])
    ptrs[2] = None
    cap = np.zeros((16, 4), dtype=np.uint64)
    h = C.c_void_p()
    rc = ctx._lib.gl_commit_from_values_cols(ctx._h, ptrs, 8, c, 3, 4, None, cap.ctypes.data, C.byref(h))
    assert rc == glb._native.GL_E_ARG and not h.value
    assert b"NULL polynomial pointer" in ctx._lib.gl_last_error(ctx._h)
    with pytest.raises(glb.GlPanic):
        glb.PolynomialBatch.from_values([np.zeros(256, dtype=np.uint64), np.zeros(128, dtype=np.uint64)], 3, False, 4)


def test_mirror_mode_download_into_pageable_memory(glb, ctx, oracle):
    """gl_commit_download (leaves + digests into ordinary arrays: the 'mirror mode' of SURVEY 8b) through the staged path."""
    n, c = 1 << 13, 24     # leaves: 2^16 x 24 x 8 B = 12.6 MB, digests 4 MB
    values = oracle.synthetic_values(c, n)
    want = oracle.commit_from_values(values, 3, 4)
    b = glb.PolynomialBatch.from_values(values, 3, False, 4)
    assert np.array_equal(b.merkle_tree.leaves, want["leaves"])
    assert np.array_equal(b.merkle_tree.digests, want["digests"])
    b.free()


def test_gl_copy_large_pageable_round_trip(glb, ctx, rng):
    x = rand_field(rng, (3 * 1024 * 1024 + 77,))     # 24 MB + a ragged tail
    d = glb.DeviceBuffer(x.shape, ctx).from_host(x)
    y = d.to_host()
    d.free()
    assert np.array_equal(x, y)


def test_back_to_back_commits_reuse_the_rings(glb, ctx, oracle):
    n, c = 1 << 14, 50
    values = oracle.synthetic_values(c, n)
    first = None
    for it in range(4):
        v = values.copy()
        v[0, 0] = np.uint64(it)
        b = glb.PolynomialBatch.from_values(v, 3, False, 4)
        cap, coeffs = _caps_coeffs(b)
        b.free()
        vd = __import__("torch").from_numpy(v.view(np.int64)).cuda()
        bd = glb.PolynomialBatch.from_values(vd, 3, False, 4)
        assert np.array_equal(cap, bd.merkle_tree.cap)
        assert np.array_equal(coeffs, bd.polynomials.cpu().numpy().view(np.uint64))
        bd.free()


def test_fri_final_poly_host_outputs_above_the_staging_threshold(glb, ctx, oracle, rng):
    """gl_fri_final_poly writes the coefficients and then reuses the same device buffer for the values: with
    page-able outputs of 2 MB each both go through the downloader and must not overtake each other."""
    import importlib

    from oracle import fri_oracle as fo

    fri = importlib.import_module("plonky2-lib_b200.fri")
    degree_bits, cols = 14, (3, 2)
    n = 1 << degree_bits
    batches, polys = [], []
    for k, c in enumerate(cols):
        v = oracle.synthetic_values(c, n, seed=900 + k)
        batches.append(glb.PolynomialBatch.from_values(v, 3, False, 4))
        polys.append(np.array(batches[-1].polynomials))
    zeta = tuple(int(x) for x in rand_field(rng, (2,)))
    instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)])]
    alpha = tuple(int(x) for x in rand_field(rng, (2,)))
    want_final = fo.final_poly_of_openings(polys, instance, alpha)
    for _ in range(3):
        got_coeffs, got_values = fri.fri_final_poly(batches, instance, alpha, 3)
        assert np.array_equal(got_coeffs[:n], want_final) and not got_coeffs[n:].any()
        assert np.array_equal(got_values, oracle.ext_coset_fft(got_coeffs, 7))
    dc, dv = fri.fri_final_poly(batches, instance, alpha, 3, resident=True)
    assert np.array_equal(dc.to_host(), got_coeffs) and np.array_equal(dv.to_host(), got_values)
    dc.free(); dv.free()
    for b in batches:
        b.free()


def test_mixed_page_locked_and_pageable_polynomials(glb, ctx, oracle):
    """Per-polynomial arrays where some are page-locked (gl_host_alloc) and some are ordinary memory, inputs and outputs:
    large page-locked ones go straight to the DMA engine in the middle of a staged run, the rest through the rings."""
    lg_n, c = 19, 6                       # 4 MB per polynomial = one ring slot
    n = 1 << lg_n
    values = oracle.synthetic_values(c, n)
    cols, outs = [], []
    for j in range(c):
        if j % 2:
            a = glb.pinned_empty((n,))
            a[:] = values[j]
            cols.append(a)
            outs.append(np.zeros(n, dtype=np.uint64))
        else:
            cols.append(values[j].copy())
            outs.append(glb.pinned_empty((n,)))
    ip = (C.c_void_p * c)(*[a.ctypes.data for a in cols])
    op = (C.c_void_p * c)(*[a.ctypes.data for a in outs])
    cap = np.zeros((16, 4), dtype=np.uint64)
    h = C.c_void_p()
    ctx.check(ctx._lib.gl_commit_from_values_cols(ctx._h, ip, lg_n, c, 3, 4, op, cap.ctypes.data, C.byref(h)))
    ctx._lib.gl_commit_free(h)
    import torch

    bd = glb.PolynomialBatch.from_values(torch.from_numpy(values.view(np.int64)).cuda(), 3, False, 4)
    assert np.array_equal(cap, bd.merkle_tree.cap)
    want = bd.polynomials.cpu().numpy().view(np.uint64)
    for j in range(c):
        assert np.array_equal(outs[j], want[j]), j
    bd.free()
