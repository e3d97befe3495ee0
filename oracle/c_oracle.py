"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/nm_oracle.c (see its header).

PARITY UNPINNED (no runnable reference, no reference golden vectors).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        P = C.c_void_p
        L.nmo_candidates.restype = C.c_int64
        L.nmo_candidates.argtypes = [C.c_int, C.c_int64, P, P, C.c_int64, P, P, C.POINTER(P)]
        L.nmo_quad_split.argtypes = [C.c_int64, P, P]
        L.nmo_quad_split_code.restype = C.c_uint8
        L.nmo_quad_split_code.argtypes = [P]
        L.nmo_orient2d.argtypes = [P, P, P]
        L.nmo_incircle.argtypes = [P, P, P, P]
        L.nmo_pairs.argtypes = [C.c_int, C.c_int64, P, P, P, P, P, C.c_int, C.c_double, P, P, P, P]
        L.nmo_pairs_quadrature.argtypes = [C.c_int, C.c_int64, P, P, P, P, P, C.c_int, C.c_double, P, P, P]
        L.nmo_pairs_codim0.argtypes = [C.c_int64, P, P, P, P, C.c_int, C.c_double, P, P, P, P]
        L.nmo_pairs_quadrature_codim0.argtypes = [C.c_int64, P, P, P, P, C.c_int, C.c_double, P, P, P]
        L.nmo_build_csr.restype = C.c_int64
        L.nmo_build_csr.argtypes = [C.c_int, C.c_int64, P, P, P, P, C.c_int64, P, P, P, P,
                                    C.POINTER(P), C.POINTER(P), C.POINTER(P)]
        L.nmo_nitsche.restype = C.c_int64
        L.nmo_nitsche.argtypes = [C.c_int, C.c_int64, P, P, P, P, P, P, C.c_double, P, P, P, P, P, P, P,
                                  C.POINTER(P), C.POINTER(P), C.POINTER(P), P]
        L.nmo_spmv.argtypes = [C.c_int, C.c_int64, C.c_int64, P, P, P, P, P]
        L.nmo_free.argtypes = [P]
        L.nmo_num_threads.restype = C.c_int
        L.nmo_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _np(t, dtype):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def _take(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.empty(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr.value if hasattr(ptr, "value") else ptr)
    out = np.frombuffer(buf, dtype=dtype).copy()
    return out


def orient2d(a, b, c) -> int:
    a, b, c = (_np(v, np.float64) for v in (a, b, c))
    return lib().nmo_orient2d(_p(a), _p(b), _p(c))


def incircle(a, b, c, d) -> int:
    a, b, c, d = (_np(v, np.float64) for v in (a, b, c, d))
    return lib().nmo_incircle(_p(a), _p(b), _p(c), _p(d))


def quad_split(verts) -> np.ndarray:
    v = _np(verts, np.float64).reshape(-1, 12)
    codes = np.zeros(v.shape[0], dtype=np.uint8)
    lib().nmo_quad_split(v.shape[0], _p(v), _p(codes))
    return codes


def candidates(bg_lo, bg_hi, im_lo, im_hi) -> np.ndarray:
    """uint64 keys (im << 32 | bg), ascending."""
    bl, bh, il, ih = (_np(v, np.float64) for v in (bg_lo, bg_hi, im_lo, im_hi))
    d = bl.shape[1]
    out = C.c_void_p()
    n = lib().nmo_candidates(d, bl.shape[0], _p(bl), _p(bh), il.shape[0], _p(il), _p(ih), C.byref(out))
    keys = _take(out, n, np.uint64)
    lib().nmo_free(out)
    return keys


def pairs(d, cand, bg_lo, bg_hi, im_verts, codes, n1d=3, tol=1e-15):
    """Per-candidate (sum_w, local[2^d], n_points) and the number of dropped sub-simplices."""
    cand = _np(cand, np.uint64)
    bl, bh = _np(bg_lo, np.float64), _np(bg_hi, np.float64)
    iv = _np(im_verts, np.float64)
    n = cand.shape[0]
    sumw = np.zeros(n); local = np.zeros((n, 1 << d)); nq = np.zeros(n, dtype=np.int32)
    dropped = C.c_int64(0)
    cd = None if codes is None else _np(codes, np.uint8)
    rc = lib().nmo_pairs(d, n, _p(cand), _p(bl), _p(bh), _p(iv), _p(cd), n1d, tol, _p(sumw), _p(local), _p(nq),
                         C.cast(C.byref(dropped), C.c_void_p))
    if rc != 0:
        raise ValueError("unsupported quadrature rule n_points_1D=%d" % n1d)
    return sumw, local, nq, dropped.value


def pairs_codim0(cand, bg_lo, bg_hi, im_verts, n1d=3, tol=1e-15):
    """pairs() for <2,2,2>: immersed quads [n, 4, 2] in the plane of the background (nmo_pairs_codim0)."""
    cand = _np(cand, np.uint64)
    n = cand.shape[0]
    sumw = np.zeros(n); local = np.zeros((n, 4)); nq = np.zeros(n, dtype=np.int32)
    dropped = C.c_int64(0)
    rc = lib().nmo_pairs_codim0(n, _p(cand), _p(_np(bg_lo, np.float64)), _p(_np(bg_hi, np.float64)), _p(_np(im_verts, np.float64)), n1d, tol,
                                _p(sumw), _p(local), _p(nq), C.cast(C.byref(dropped), C.c_void_p))
    if rc != 0:
        raise ValueError("unsupported quadrature rule n_points_1D=%d" % n1d)
    return sumw, local, nq, dropped.value


def quadrature_codim0(pairs_keys, nq, bg_lo, bg_hi, im_verts, n1d=3, tol=1e-15):
    keys = _np(pairs_keys, np.uint64)
    q_off = np.concatenate([[0], np.cumsum(_np(nq, np.int64))]).astype(np.uint64)
    pts = np.zeros((int(q_off[-1]), 2)); wts = np.zeros(int(q_off[-1]))
    rc = lib().nmo_pairs_quadrature_codim0(keys.shape[0], _p(keys), _p(_np(bg_lo, np.float64)), _p(_np(bg_hi, np.float64)),
                                           _p(_np(im_verts, np.float64)), n1d, tol, _p(q_off), _p(pts), _p(wts))
    if rc != 0:
        raise ValueError("nmo_pairs_quadrature_codim0 failed (%d)" % rc)
    return q_off, pts, wts


def quadrature(d, pairs_keys, nq, bg_lo, bg_hi, im_verts, codes, n1d=3, tol=1e-15):
    """(q_offset, points, weights) of the given pairs in the reference's own subdivision (nmo_pairs_quadrature);
    `nq` = the per-pair point counts `pairs()` reported."""
    keys = _np(pairs_keys, np.uint64)
    q_off = np.concatenate([[0], np.cumsum(_np(nq, np.int64))]).astype(np.uint64)
    pts = np.zeros((int(q_off[-1]), d)); wts = np.zeros(int(q_off[-1]))
    cd = None if codes is None else _np(codes, np.uint8)
    rc = lib().nmo_pairs_quadrature(d, keys.shape[0], _p(keys), _p(_np(bg_lo, np.float64)), _p(_np(bg_hi, np.float64)),
                                    _p(_np(im_verts, np.float64)), _p(cd), n1d, tol, _p(q_off), _p(pts), _p(wts))
    if rc != 0:
        raise ValueError("nmo_pairs_quadrature failed (%d): unsupported rule or inconsistent point counts" % rc)
    return q_off, pts, wts


def line_table(bg):
    n = bg.n_dofs
    line_of = np.full(n, -1, dtype=np.int32)
    cd = _np(bg.c_dof, np.int64)
    line_of[cd] = np.arange(cd.shape[0], dtype=np.int32)
    return line_of


def build_csr(d, acc, local, bg, im):
    acc = _np(acc, np.uint64); local = _np(local, np.float64)
    bd = _np(bg.dofs, np.uint32); idf = _np(im.dofs, np.uint32)
    line_of = line_table(bg)
    cp = _np(bg.c_ptr, np.int64); cm = _np(bg.c_master, np.uint32); cw = _np(bg.c_weight, np.float64)
    rp, cl, vl = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nnz = lib().nmo_build_csr(d, acc.shape[0], _p(acc), _p(local), _p(bd), _p(idf), bg.n_dofs, _p(line_of), _p(cp),
                              _p(cm), _p(cw), C.byref(rp), C.byref(cl), C.byref(vl))
    rowptr = _take(rp, bg.n_dofs + 1, np.uint64)
    col = _take(cl, nnz, np.uint32)
    val = _take(vl, nnz, np.float64)
    for p in (rp, cl, vl):
        lib().nmo_free(p)
    return rowptr, col, val


def nitsche(bg, bg_ids, q_offset, points, weights, coefficient=None, rhs_values=None, penalty=1.0):
    """Nitsche matrix contribution (scipy CSR, duplicates summed, explicit zeros kept) and rhs (or None) on the given
    quadrature: reference code/include/coupling_utilities.h:306-481 restated in nmo_nitsche."""
    import scipy.sparse as sp
    d = bg.spacedim
    ids = _np(bg_ids, np.uint32); qo = _np(q_offset, np.uint64)
    pts = _np(points, np.float64); wts = _np(weights, np.float64)
    cf = None if coefficient is None else _np(coefficient, np.float64)
    fv = None if rhs_values is None else _np(rhs_values, np.float64)
    bl, bh = _np(bg.lo, np.float64), _np(bg.hi, np.float64)
    bd = _np(bg.dofs, np.uint32)
    has_c = _np(bg.c_dof, np.int64).shape[0] > 0
    line_of = line_table(bg) if has_c else None
    cp = _np(bg.c_ptr, np.int64); cm = _np(bg.c_master, np.uint32); cw = _np(bg.c_weight, np.float64)
    rhs = np.zeros(bg.n_dofs) if fv is not None else None
    r, c, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
    n = lib().nmo_nitsche(d, ids.shape[0], _p(ids), _p(qo), _p(pts), _p(wts), _p(cf), _p(fv), float(penalty), _p(bl), _p(bh),
                          _p(bd), _p(line_of), _p(cp), _p(cm), _p(cw), C.byref(r), C.byref(c), C.byref(v), _p(rhs))
    rows = _take(r, n, np.uint32).astype(np.int64); cols = _take(c, n, np.uint32).astype(np.int64); vals = _take(v, n, np.float64)
    for p in (r, c, v):
        lib().nmo_free(p)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(bg.n_dofs, bg.n_dofs)).tocsr()
    A.sum_duplicates(); A.sort_indices()
    return A, rhs


def spmv(transpose, rowptr, col, val, x, n_rows, n_cols):
    x = _np(x, np.float64)
    y = np.zeros(n_cols if transpose else n_rows)
    lib().nmo_spmv(int(transpose), n_rows, n_cols, _p(rowptr), _p(col), _p(val), _p(x), _p(y))
    return y


def assemble(bg, im, n1d: int = 3, tol: float = 1e-15, threads: Optional[int] = None):
    """Whole path on the CPU: candidates -> per-pair quadrature -> accepted pairs -> CSR."""
    if threads:
        lib().nmo_set_num_threads(threads)
    d = bg.spacedim
    ilo, ihi = im.boxes()
    cand = candidates(bg.lo, bg.hi, ilo, ihi)
    codim0 = im.dim == d
    codes = quad_split(im.verts) if d == 3 else None
    if codim0:
        sumw, local, nq, dropped = pairs_codim0(cand, bg.lo, bg.hi, im.verts, n1d, tol)
    else:
        sumw, local, nq, dropped = pairs(d, cand, bg.lo, bg.hi, im.verts, codes, n1d, tol)
    keep = sumw > tol
    acc = cand[keep]
    rowptr, col, val = build_csr(d, acc, local[keep], bg, im)
    return dict(candidates=cand, codes=codes, sum_w_all=sumw, accepted=acc, sum_w=sumw[keep], local=local[keep],
                nq=nq[keep], dropped=dropped, rowptr=rowptr, col=col, val=val)


def assemble_with_quadrature(bg, im, n1d: int = 3, tol: float = 1e-15):
    """assemble() plus the per-pair quadrature in the reference's subdivision: adds q_offset, points, weights."""
    r = assemble(bg, im, n1d, tol)
    if im.dim == bg.spacedim:
        r["q_offset"], r["points"], r["weights"] = quadrature_codim0(r["accepted"], r["nq"], bg.lo, bg.hi, im.verts, n1d, tol)
        return r
    r["q_offset"], r["points"], r["weights"] = quadrature(bg.spacedim, r["accepted"], r["nq"], bg.lo, bg.hi, im.verts, r["codes"], n1d, tol)
    return r
