"""ctypes binding of libnmgpu.so (the C ABI of include/nmgpu.h).

There is no fallback of any kind: if the shared library is missing or no CUDA device is
present, the calls raise.  torch is used only to own device memory handed to / received
from the library.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NMGPU_LIB") or os.path.join(_HERE, "libnmgpu.so")   # NMGPU_LIB: another build of the same library (kernel experiments)

NM_HOST, NM_DEVICE, NM_HOST_ASYNC, NM_DEVICE_INPLACE = 0, 1, 2, 3
NM_LOCAL_IMMERSED, NM_LOCAL_ROWS = 0, 1
INFO = {"merge_path": 0, "compact_rows": 1, "stored_rows": 2, "nnz": 3, "index_kind": 4, "host_syncs": 5, "merge_slots": 6}
STATUS = {0: "NM_OK", 1: "NM_ERR_INVALID", 2: "NM_ERR_CUDA", 3: "NM_ERR_UNSUPPORTED", 4: "NM_ERR_STATE", 5: "NM_ERR_NOMEM"}
PHASES = ["upload", "index", "candidates", "split", "clip", "merge", "transpose", "spmv", "spmv_t", "download",
          "k_clip", "k_cand_count", "k_cand_fill", "k_merge"]

# every symbol include/nmgpu.h declares
SYMBOLS = ["nm_version", "nm_create", "nm_destroy", "nm_last_error", "nm_set_background", "nm_set_background_boxes",
           "nm_set_background_dofs", "nm_set_immersed_dofs", "nm_set_immersed", "nm_flag_overlapped_background", "nm_intersect", "nm_get_candidates", "nm_get_pairs", "nm_get_quadrature",
           "nm_get_split_codes", "nm_build_sparsity", "nm_assemble", "nm_assemble_from_quadrature", "nm_assemble_host", "nm_assemble_nitsche", "nm_get_nitsche", "nm_get_csr",
           "nm_get_csr_compact", "nm_sync_downloads", "nm_spmv", "nm_spmv_local", "nm_spmv_local_pair", "nm_local_to_global", "nm_global_to_local",
           "nm_spmv_push", "nm_spmv_push_owned", "nm_spmv_rows", "nm_owner_offsets", "nm_pull_rows", "nm_set_finite_elements", "nm_assemble_fe", "nm_schur_solve", "nm_get_info", "nm_measure_fp64_peak", "nm_get_timings", "nm_get_launch_count", "nm_device_bytes", "nm_get_stream"]


class NmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s: %s" % (STATUS.get(code, str(code)), msg))
        self.code = code


class HostProblem(C.Structure):
    """nm_host_problem of include/nmgpu.h."""
    _fields_ = [("spacedim", C.c_int), ("n_bg", C.c_uint64), ("lo", C.c_void_p), ("hi", C.c_void_p), ("bg_dofs", C.c_void_p),
                ("n_space_dofs", C.c_uint64), ("n_constrained", C.c_uint64), ("c_dof", C.c_void_p), ("c_ptr", C.c_void_p),
                ("c_master", C.c_void_p), ("c_weight", C.c_void_p), ("dim", C.c_int), ("n_im", C.c_uint64), ("im_verts", C.c_void_p),
                ("im_dofs", C.c_void_p), ("n_im_dofs", C.c_uint64), ("degree", C.c_uint), ("tol", C.c_double), ("edge", C.c_void_p)]


class FeDesc(C.Structure):
    """nm_fe of include/nmgpu.h."""
    _fields_ = [("degree", C.c_uint), ("dofs_per_cell", C.c_uint), ("unit_support_points", C.c_void_p)]


class SchurProblem(C.Structure):
    """nm_schur_problem of include/nmgpu.h."""
    _fields_ = [("n_space", C.c_uint64), ("n_immersed", C.c_uint64), ("k_rowptr", C.c_void_p), ("k_col", C.c_void_p), ("k_val", C.c_void_p),
                ("m_rowptr", C.c_void_p), ("m_col", C.c_void_p), ("m_val", C.c_void_p), ("space_rhs", C.c_void_p), ("embedded_rhs", C.c_void_p),
                ("inhomogeneity", C.c_void_p), ("inner_max_it", C.c_uint), ("inner_tol", C.c_double), ("inner_reduce", C.c_double),
                ("outer_max_it", C.c_uint), ("outer_tol", C.c_double), ("outer_reduce", C.c_double), ("gmres_basis", C.c_uint)]


class SchurStats(C.Structure):
    _fields_ = [("outer_iterations", C.c_uint), ("inner_iterations", C.c_uint64), ("outer_residual", C.c_double),
                ("lambda_norm_sqr", C.c_double), ("ms", C.c_float)]


class Counts(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_candidates", "n_accepted", "n_qpoints", "n_dropped", "n_marginal", "n_split_ties")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def load():
    """Load libnmgpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libnmgpu.so not found at %s -- run `python __graft_entry__.py build` (nvcc, sm_100a)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    P, U64, I = C.c_void_p, C.c_uint64, C.c_int
    L.nm_version.restype = I
    L.nm_create.argtypes = [C.POINTER(P), I]
    L.nm_destroy.argtypes = [P]
    L.nm_last_error.restype = C.c_char_p
    L.nm_last_error.argtypes = [P]
    L.nm_set_background.argtypes = [P, I, U64, P, P, U64, U64, P, P, P, P, I]
    L.nm_set_background_boxes.argtypes = [P, I, U64, P, P, P, U64, U64, P, P, P, P, I]
    L.nm_set_immersed.argtypes = [P, I, I, U64, P, P, C.c_uint32, U64, I]
    L.nm_set_background_dofs.argtypes = [P, P, U64, U64, P, P, P, P, I]
    L.nm_set_immersed_dofs.argtypes = [P, P, C.c_uint32, U64, I]
    L.nm_flag_overlapped_background.argtypes = [P, P, I]
    L.nm_intersect.argtypes = [P, C.c_uint, C.c_double, C.POINTER(Counts)]
    L.nm_get_candidates.argtypes = [P, P, P, I]
    L.nm_get_pairs.argtypes = [P, P, P, P, I]
    L.nm_get_quadrature.argtypes = [P, P, P, P, I]
    L.nm_get_split_codes.argtypes = [P, P, I]
    L.nm_build_sparsity.argtypes = [P, C.POINTER(U64)]
    L.nm_assemble.argtypes = [P]
    L.nm_assemble_from_quadrature.argtypes = [P, U64, P, P, P, P, P, I]
    L.nm_assemble_host.argtypes = [P, C.POINTER(HostProblem), C.POINTER(Counts), C.POINTER(U64)]
    L.nm_assemble_nitsche.argtypes = [P, U64, P, P, P, P, P, P, C.c_double, I, I, C.POINTER(U64)]
    L.nm_get_nitsche.argtypes = [P, P, P, P, P, I]
    L.nm_get_csr.argtypes = [P, I, P, P, P, I]
    L.nm_spmv.argtypes = [P, I, P, P, I]
    L.nm_get_csr_compact.argtypes = [P, C.POINTER(U64), P, P, P, P, I]
    L.nm_sync_downloads.argtypes = [P]
    L.nm_spmv_local.argtypes = [P, I, P, P, I]
    L.nm_spmv_local_pair.argtypes = [P, P, P, P, P, I]
    L.nm_local_to_global.argtypes = [P, I, P, I, P]
    L.nm_global_to_local.argtypes = [P, I, P, P, I]
    L.nm_set_finite_elements.argtypes = [P, C.POINTER(FeDesc), P, U64, U64, P, P, P, P, C.POINTER(FeDesc), P, U64, I]
    L.nm_assemble_fe.argtypes = [P, U64, P, P, P, P, P, I, C.POINTER(U64)]
    L.nm_schur_solve.argtypes = [P, C.POINTER(SchurProblem), P, P, C.POINTER(SchurStats), I]
    L.nm_get_info.argtypes = [P, I, C.POINTER(U64)]
    L.nm_measure_fp64_peak.argtypes = [P, C.c_double, C.POINTER(C.c_double)]
    L.nm_spmv_push.argtypes = [P, P, I, C.POINTER(C.c_void_p), I]
    L.nm_spmv_push_owned.argtypes = [P, P, I, C.POINTER(C.c_void_p), U64]
    L.nm_spmv_rows.argtypes = [P, P, P, P, C.POINTER(U64)]
    L.nm_owner_offsets.argtypes = [P, I, U64, P]
    L.nm_pull_rows.argtypes = [P, P, I, P, P, U64, U64, P]
    L.nm_get_timings.argtypes = [P, P, I]
    L.nm_device_bytes.argtypes = [P, C.POINTER(U64)]
    L.nm_get_launch_count.argtypes = [P, C.POINTER(U64)]
    L.nm_get_stream.argtypes = [P, C.POINTER(P)]
    _lib = L
    return L


def _ptr(a):
    """(pointer, where) of a torch tensor (CPU or CUDA) or numpy array; None -> (None, host)."""
    if a is None:
        return None, NM_HOST
    if isinstance(a, torch.Tensor):
        assert a.is_contiguous(), "tensor must be contiguous"
        return C.c_void_p(a.data_ptr()), (NM_DEVICE if a.is_cuda else NM_HOST)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p), NM_HOST


def _where(*arrs) -> int:
    ws = {_ptr(a)[1] for a in arrs if a is not None}
    if len(ws) > 1:
        raise ValueError("all arrays of one call must live on the same side (host or device)")
    w = ws.pop() if ws else NM_HOST
    if w == NM_DEVICE:
        # the library works on its own stream: whatever produced these tensors on torch's
        # current stream must have finished before the pointers are handed over
        torch.cuda.current_stream().synchronize()
    return w


class Context:
    """One nm_ctx: one GPU, one background grid, one immersed shard."""

    def __init__(self, device: int = 0):
        self.L = load()
        self.h = C.c_void_p()
        rc = self.L.nm_create(C.byref(self.h), device)
        if rc != 0:
            raise NmError(rc, "nm_create(device=%d) failed -- libnmgpu.so needs a CUDA device; there is no CPU path" % device)
        self.device = device
        self.counts = None
        self.spacedim = None
        self.n_space_dofs = self.n_im_dofs = 0

    def close(self):
        if self.h:
            self.L.nm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int):
        if rc != 0:
            raise NmError(rc, self.L.nm_last_error(self.h).decode())

    # ---- inputs -------------------------------------------------------------------------
    def _inplace(self, w, inplace, *keep):
        """NM_DEVICE_INPLACE for device arrays when asked for: the library reads the DoF tables / immersed vertices where
        they are, so the tensors are kept alive here until the input is set again."""
        if inplace and w == NM_DEVICE:
            self._borrowed = getattr(self, "_borrowed", {})
            self._borrowed[keep[0]] = keep[1:]
            return NM_DEVICE_INPLACE
        return w

    def set_background(self, spacedim, verts=None, lo=None, hi=None, dofs=None, n_dofs=0, c_dof=None, c_ptr=None,
                       c_master=None, c_weight=None, inplace=False):
        """Either verts [n, 2^d, d] (deal.II order) or boxes lo/hi [n, d]; dofs [n, 2^d] uint32/int32;
        constraint lines as CSR over ascending c_dof (c_ptr uint64/int64)."""
        n_c = 0 if c_dof is None else int(c_dof.shape[0])
        w = self._inplace(_where(verts, lo, hi, dofs, c_dof, c_ptr, c_master, c_weight), inplace, "bg_dofs", dofs)
        n = int((verts if verts is not None else lo).shape[0])
        if verts is not None:
            rc = self.L.nm_set_background(self.h, spacedim, n, _ptr(verts)[0], _ptr(dofs)[0], n_dofs, n_c, _ptr(c_dof)[0],
                                          _ptr(c_ptr)[0], _ptr(c_master)[0], _ptr(c_weight)[0], w)
        else:
            rc = self.L.nm_set_background_boxes(self.h, spacedim, n, _ptr(lo)[0], _ptr(hi)[0], _ptr(dofs)[0], n_dofs, n_c,
                                                _ptr(c_dof)[0], _ptr(c_ptr)[0], _ptr(c_master)[0], _ptr(c_weight)[0], w)
        self._chk(rc)
        self.spacedim, self.n_space_dofs = spacedim, n_dofs

    def set_background_dofs(self, dofs, n_dofs, c_dof=None, c_ptr=None, c_master=None, c_weight=None, inplace=False):
        n_c = 0 if c_dof is None else int(c_dof.shape[0])
        w = self._inplace(_where(dofs, c_dof, c_ptr, c_master, c_weight), inplace, "bg_dofs", dofs)
        self._chk(self.L.nm_set_background_dofs(self.h, _ptr(dofs)[0], n_dofs, n_c, _ptr(c_dof)[0], _ptr(c_ptr)[0],
                                                _ptr(c_master)[0], _ptr(c_weight)[0], w))
        self.n_space_dofs = n_dofs

    def set_immersed_dofs(self, dofs, n_dofs, dofs_per_cell=1, inplace=False):
        self._chk(self.L.nm_set_immersed_dofs(self.h, _ptr(dofs)[0], dofs_per_cell, n_dofs, self._inplace(_where(dofs), inplace, "im_dofs", dofs)))
        self.n_im_dofs = n_dofs

    def set_immersed(self, dim, spacedim, verts, dofs, n_dofs, dofs_per_cell=1, inplace=False):
        w = self._inplace(_where(verts, dofs), inplace, "immersed", verts, dofs)
        self._chk(self.L.nm_set_immersed(self.h, dim, spacedim, int(verts.shape[0]), _ptr(verts)[0], _ptr(dofs)[0],
                                         dofs_per_cell, n_dofs, w))
        self.n_im_dofs = n_dofs
        self.n_im = int(verts.shape[0])

    def flag_overlapped_background(self, n_bg: int):
        """uint8 flag per background cell: closed box meets some immersed box (adjust_grids flagging)."""
        out = torch.empty(n_bg, dtype=torch.uint8)
        self._chk(self.L.nm_flag_overlapped_background(self.h, _ptr(out)[0], NM_HOST))
        return out

    # ---- intersection -------------------------------------------------------------------
    def intersect(self, degree: int = 3, tol: float = 1e-15) -> dict:
        c = Counts()
        self._chk(self.L.nm_intersect(self.h, degree, tol, C.byref(c)))
        self.counts = c.as_dict()
        return self.counts

    def _out(self, n, dtype, device):
        if device == "cpu":
            return torch.empty(n, dtype=dtype)
        return torch.empty(n, dtype=dtype, device="cuda:%d" % self.device)

    def candidates(self, device="cpu"):
        n = self.counts["n_candidates"]
        im, bg = self._out(n, torch.int32, device), self._out(n, torch.int32, device)
        self._chk(self.L.nm_get_candidates(self.h, _ptr(im)[0], _ptr(bg)[0], _where(im)))
        return im, bg

    def pairs(self, device="cpu"):
        n = self.counts["n_accepted"]
        im, bg, sw = self._out(n, torch.int32, device), self._out(n, torch.int32, device), self._out(n, torch.float64, device)
        self._chk(self.L.nm_get_pairs(self.h, _ptr(im)[0], _ptr(bg)[0], _ptr(sw)[0], _where(im)))
        return im, bg, sw

    def quadrature(self, device="cpu"):
        na, nq = self.counts["n_accepted"], self.counts["n_qpoints"]
        off = self._out(na + 1, torch.int64, device)
        pts = self._out(nq * self.spacedim, torch.float64, device).reshape(nq, self.spacedim)
        w = self._out(nq, torch.float64, device)
        self._chk(self.L.nm_get_quadrature(self.h, _ptr(off)[0], _ptr(pts)[0], _ptr(w)[0], _where(off)))
        return off, pts, w

    def split_codes(self):
        out = torch.empty(self.n_im, dtype=torch.uint8)
        self._chk(self.L.nm_get_split_codes(self.h, _ptr(out)[0], NM_HOST))
        return out

    # ---- matrix -------------------------------------------------------------------------
    def build_sparsity(self) -> int:
        nnz = C.c_uint64(0)
        self._chk(self.L.nm_build_sparsity(self.h, C.byref(nnz)))
        self.nnz = int(nnz.value)
        return self.nnz

    def assemble(self):
        self._chk(self.L.nm_assemble(self.h))

    def assemble_from_quadrature(self, im, bg, q_offset, points, weights):
        w = _where(im, bg, q_offset, points, weights)
        self._chk(self.L.nm_assemble_from_quadrature(self.h, int(im.shape[0]), _ptr(im)[0], _ptr(bg)[0], _ptr(q_offset)[0],
                                                     _ptr(points)[0], _ptr(weights)[0], w))

    def assemble_host(self, spacedim, lo, hi, dofs, n_dofs, c_dof, c_ptr, c_master, c_weight, im_dim, im_verts, im_dofs, n_im_dofs,
                      degree, tol, edge=None):
        """Whole assembly from HOST arrays (torch CPU tensors, ideally pinned) in one call with the uploads overlapped
        with the computation (nm_assemble_host); afterwards the context is in the same state as after
        set_background(lo, hi, ...) + set_immersed + intersect + build_sparsity + assemble."""
        arrs = (lo, hi, dofs, c_dof, c_ptr, c_master, c_weight, im_verts, im_dofs, edge)
        if _where(*arrs) != NM_HOST:
            raise ValueError("assemble_host takes host arrays")
        # edge [n_bg]: cubic cells as lower corner + edge length (hi = fl(lo + edge) on every axis; `hi` may then be None)
        hp = HostProblem(spacedim, int(lo.shape[0]), _ptr(lo)[0], _ptr(hi)[0], _ptr(dofs)[0], int(n_dofs), int(c_dof.shape[0]),
                         _ptr(c_dof)[0], _ptr(c_ptr)[0], _ptr(c_master)[0], _ptr(c_weight)[0], im_dim, int(im_verts.shape[0]),
                         _ptr(im_verts)[0], _ptr(im_dofs)[0], int(n_im_dofs), int(degree), float(tol), _ptr(edge)[0])
        cnt, nnz = Counts(), C.c_uint64(0)
        self._chk(self.L.nm_assemble_host(self.h, C.byref(hp), C.byref(cnt), C.byref(nnz)))
        self.spacedim, self.n_bg, self.n_space_dofs = spacedim, int(lo.shape[0]), int(n_dofs)
        self.n_im, self.n_im_dofs = int(im_verts.shape[0]), int(n_im_dofs)
        self.counts = cnt.as_dict()
        self.nnz = int(nnz.value)
        return self.counts

    def assemble_nitsche(self, bg, q_offset, points, weights, coefficient=None, rhs_values=None, penalty=1.0, matrix=True):
        """Nitsche terms on the given quadrature (reference coupling_utilities.h:306-481): returns
        (rowptr, col, val) of the n_space x n_space contribution (None when matrix=False) and the rhs contribution
        (None when rhs_values is None), on the device the inputs live on."""
        what = (1 if matrix else 0) | (2 if rhs_values is not None else 0)
        args = [a for a in (bg, q_offset, points, weights, coefficient, rhs_values) if a is not None]
        w = _where(*args)
        nnz = C.c_uint64(0)
        self._chk(self.L.nm_assemble_nitsche(self.h, int(bg.shape[0]), _ptr(bg)[0], _ptr(q_offset)[0], _ptr(points)[0], _ptr(weights)[0],
                                             _ptr(coefficient)[0], _ptr(rhs_values)[0], float(penalty), what, w,
                                             C.byref(nnz)))
        dev = "cuda" if w == NM_DEVICE else "cpu"
        n = self.n_space_dofs
        rp = col = val = rhs = None
        if matrix:
            rp, col, val = self._out(n + 1, torch.int64, dev), self._out(nnz.value, torch.int32, dev), self._out(nnz.value, torch.float64, dev)
        if rhs_values is not None:
            rhs = self._out(n, torch.float64, dev)
        self._chk(self.L.nm_get_nitsche(self.h, _ptr(rp)[0], _ptr(col)[0], _ptr(val)[0], _ptr(rhs)[0], w))
        return (rp, col, val), rhs

    # ---- general finite elements (FE_Q(p) x FE_DGQ(q)) ---------------------------------------
    def set_finite_elements(self, space_fe, space_dofs, n_space_dofs, immersed_fe, immersed_dofs, n_immersed_dofs,
                            c_dof=None, c_ptr=None, c_master=None, c_weight=None):
        """`space_fe` / `immersed_fe` = (degree, unit support points [dofs_per_cell, dim] in the element's dof order);
        dof tables [n_cells, dofs_per_cell]; constraint lines of the space as in set_background (nm_set_finite_elements)."""
        keep = []

        def desc(fe):
            deg, sp = fe
            sp = np.ascontiguousarray(np.asarray(sp, dtype=np.float64))
            keep.append(sp)
            return FeDesc(int(deg), int(sp.shape[0]), sp.ctypes.data_as(C.c_void_p) if sp.size else None)
        fs, fi = desc(space_fe), desc(immersed_fe)
        n_c = 0 if c_dof is None else int(c_dof.shape[0])
        w = _where(space_dofs, immersed_dofs, c_dof, c_ptr, c_master, c_weight)
        self._chk(self.L.nm_set_finite_elements(self.h, C.byref(fs), _ptr(space_dofs)[0], int(n_space_dofs), n_c, _ptr(c_dof)[0], _ptr(c_ptr)[0],
                                                _ptr(c_master)[0], _ptr(c_weight)[0], C.byref(fi), _ptr(immersed_dofs)[0], int(n_immersed_dofs), w))
        self.n_space_dofs, self.n_im_dofs = int(n_space_dofs), int(n_immersed_dofs)
        self.fe_dofs_per_cell = (fs.dofs_per_cell, fi.dofs_per_cell)

    def assemble_fe(self, im=None, bg=None, q_offset=None, points=None, weights=None) -> int:
        """Coupling matrix for the elements set with set_finite_elements: on the context's own intersection quadrature
        (no arguments) or on the given tuples (nm_assemble_fe).  Returns nnz."""
        nnz = C.c_uint64(0)
        if im is None:
            self._chk(self.L.nm_assemble_fe(self.h, 0, None, None, None, None, None, NM_DEVICE, C.byref(nnz)))
        else:
            w = _where(im, bg, q_offset, points, weights)
            self._chk(self.L.nm_assemble_fe(self.h, int(im.shape[0]), _ptr(im)[0], _ptr(bg)[0], _ptr(q_offset)[0], _ptr(points)[0],
                                            _ptr(weights)[0], w, C.byref(nnz)))
        self.nnz = int(nnz.value)
        return self.nnz

    def schur_solve(self, K, M, space_rhs, embedded_rhs, inner=(2000, 1e-14, 1e-12), outer=(0, 1e-11, 1e-2), inhomogeneity=None,
                    gmres_basis: int = 28):
        """The drivers' solve() around this context's coupling matrix (nm_schur_solve): K, M = (rowptr int64, col int32,
        val float64) CSR triples; inner / outer = ReductionControl (max iterations, tol, reduce) of the CG on K and of the
        GMRES on the Schur complement.  Returns (lambda, solution, stats dict), on the side the inputs live on."""
        arrs = list(K) + list(M) + [space_rhs, embedded_rhs] + ([inhomogeneity] if inhomogeneity is not None else [])
        w = _where(*arrs)
        pr = SchurProblem(int(space_rhs.shape[0]), int(embedded_rhs.shape[0]), _ptr(K[0])[0], _ptr(K[1])[0], _ptr(K[2])[0], _ptr(M[0])[0],
                          _ptr(M[1])[0], _ptr(M[2])[0], _ptr(space_rhs)[0], _ptr(embedded_rhs)[0], _ptr(inhomogeneity)[0], int(inner[0]),
                          float(inner[1]), float(inner[2]), int(outer[0]), float(outer[1]), float(outer[2]), int(gmres_basis))
        dev = "cuda" if w == NM_DEVICE else "cpu"
        lam = self._out(pr.n_immersed, torch.float64, dev)
        sol = self._out(pr.n_space, torch.float64, dev)
        st = SchurStats()
        self._chk(self.L.nm_schur_solve(self.h, C.byref(pr), _ptr(lam)[0], _ptr(sol)[0], C.byref(st), w))
        return lam, sol, {n: getattr(st, n) for n, _ in SchurStats._fields_}

    def csr(self, transpose: bool = False, values: bool = True, device="cpu", out=None):
        """CSR of the matrix; `out=(rowptr, col, val)` reuses caller buffers (e.g. pinned host memory)
        of at least the needed size and returns views of them."""
        nnz = self.build_sparsity()
        n_rows = self.n_im_dofs if transpose else self.n_space_dofs
        if out is not None:
            rp, col, val = out[0][:n_rows + 1], out[1][:nnz], (out[2][:nnz] if values and out[2] is not None else None)
        else:
            rp = self._out(n_rows + 1, torch.int64, device)
            col = self._out(nnz, torch.int32, device)
            val = self._out(nnz, torch.float64, device) if values else None
        self._chk(self.L.nm_get_csr(self.h, int(transpose), _ptr(rp)[0], _ptr(col)[0], _ptr(val)[0], _where(rp)))
        return rp, col, val

    def csr_compact(self, values: bool = True, device="cpu", out=None, asynchronous: bool = False):
        """The stored orientation without its empty rows: (row_ids, row_len, col, val) -- nm_get_csr_compact.
        `out=(row_ids, row_len, col, val)` reuses caller buffers (pinned host memory for `asynchronous=True`, in which
        case the arrays are complete only after `sync_downloads()`)."""
        nnz = self.build_sparsity()
        n = C.c_uint64(0)
        self._chk(self.L.nm_get_csr_compact(self.h, C.byref(n), None, None, None, None, NM_HOST))
        nr = int(n.value)
        if out is not None:
            ids, ln, col, val = out[0][:nr], out[1][:nr], out[2][:nnz], (out[3][:nnz] if values and out[3] is not None else None)
        else:
            ids, ln = self._out(nr, torch.int32, device), self._out(nr, torch.int32, device)
            col = self._out(nnz, torch.int32, device)
            val = self._out(nnz, torch.float64, device) if values else None
        w = _where(ids, ln, col, val)
        if asynchronous:
            if w != NM_HOST or not all(t.is_pinned() for t in (ids, ln, col) + ((val,) if val is not None else ())):
                raise ValueError("an asynchronous read-back needs pinned host buffers")
            w = NM_HOST_ASYNC
        self._chk(self.L.nm_get_csr_compact(self.h, C.byref(n), _ptr(ids)[0], _ptr(ln)[0], _ptr(col)[0], _ptr(val)[0], w))
        return ids, ln, col, val

    def sync_downloads(self):
        self._chk(self.L.nm_sync_downloads(self.h))

    def n_rows_local(self) -> int:
        """Number of non-empty rows of the stored orientation (length of the NM_LOCAL_ROWS numbering)."""
        n = C.c_uint64(0)
        self._chk(self.L.nm_get_csr_compact(self.h, C.byref(n), None, None, None, None, NM_HOST))
        return int(n.value)

    def spmv_local(self, x, transpose: bool = False, out=None):
        """Products in local numbering (nm_spmv_local): transpose=False: y[rows] = A x[cells]; True: y[cells] = A^T x[rows]."""
        ny = self.n_im * getattr(self, "fe_dofs_per_cell", (0, 1))[1] if transpose else self.n_rows_local()
        if out is None:
            out = torch.empty(ny, dtype=torch.float64, device=x.device)
        self._chk(self.L.nm_spmv_local(self.h, int(transpose), _ptr(x)[0], _ptr(out)[0], _where(x, out)))
        return out

    def spmv_local_pair(self, x_ct, x_c, out_ct=None, out_c=None):
        """Both products of one Schur iteration in local numbering (nm_spmv_local_pair): returns (Ct x_ct [rows], C x_c [cells]);
        with pinned host vectors the upload of x_c overlaps the download of the first result."""
        n_cells = self.n_im * getattr(self, "fe_dofs_per_cell", (0, 1))[1]
        if out_ct is None:
            out_ct = torch.empty(self.n_rows_local(), dtype=torch.float64, device=x_ct.device)
        if out_c is None:
            out_c = torch.empty(n_cells, dtype=torch.float64, device=x_ct.device)
        self._chk(self.L.nm_spmv_local_pair(self.h, _ptr(x_ct)[0], _ptr(out_ct)[0], _ptr(x_c)[0], _ptr(out_c)[0], _where(x_ct, out_ct, x_c, out_c)))
        return out_ct, out_c

    def spmv_rows(self, x: torch.Tensor, y_rows: torch.Tensor, row_ids: Optional[torch.Tensor] = None) -> int:
        """y_rows[r] = (A x)[dof of stored row r] for the rows with entries, optionally their dof ids (nm_spmv_rows);
        device tensors; returns the number of rows."""
        assert x.is_cuda and y_rows.is_cuda
        n = C.c_uint64(0)
        self._chk(self.L.nm_spmv_rows(self.h, C.c_void_p(x.data_ptr()), C.c_void_p(y_rows.data_ptr()),
                                      C.c_void_p(row_ids.data_ptr()) if row_ids is not None else None, C.byref(n)))
        return int(n.value)

    def owner_offsets(self, n_ranks: int, rows_per_rank: int, offsets: torch.Tensor):
        assert offsets.is_cuda and offsets.dtype == torch.int64 and offsets.numel() >= n_ranks + 1
        self._chk(self.L.nm_owner_offsets(self.h, int(n_ranks), int(rows_per_rank), C.c_void_p(offsets.data_ptr())))

    def pull_rows(self, src_offsets_ptr: int, me: int, src_ids_ptr: int, src_vals_ptr: int, first_row: int, y_owned: torch.Tensor):
        """y_owned[id - first_row] += val over this rank's segment of a (peer's) send buffer (nm_pull_rows); raw device pointers."""
        self._chk(self.L.nm_pull_rows(self.h, C.c_void_p(src_offsets_ptr), int(me), C.c_void_p(src_ids_ptr), C.c_void_p(src_vals_ptr),
                                      int(first_row), int(y_owned.numel()), C.c_void_p(y_owned.data_ptr())))

    def local_to_global(self, which: int, local, global_dev: torch.Tensor):
        assert global_dev.is_cuda
        w = _where(local)
        torch.cuda.current_stream().synchronize()
        self._chk(self.L.nm_local_to_global(self.h, which, _ptr(local)[0], w, C.c_void_p(global_dev.data_ptr())))
        return global_dev

    def global_to_local(self, which: int, global_dev: torch.Tensor, local):
        assert global_dev.is_cuda
        w = _where(local)
        torch.cuda.current_stream().synchronize()
        self._chk(self.L.nm_global_to_local(self.h, which, C.c_void_p(global_dev.data_ptr()), _ptr(local)[0], w))
        return local

    def info(self, key: str) -> int:
        v = C.c_uint64(0)
        self._chk(self.L.nm_get_info(self.h, INFO[key], C.byref(v)))
        r = int(v.value)
        return r - (1 << 64) if r >= (1 << 63) else r

    def measure_fp64_peak(self, ms_target: float = 20.0) -> float:
        v = C.c_double(0)
        self._chk(self.L.nm_measure_fp64_peak(self.h, float(ms_target), C.byref(v)))
        return float(v.value)

    def spmv(self, x, transpose: bool = False, out=None):
        """transpose=False: y[n_space] = A x[n_im]  (Ct.vmult);  True: y[n_im] = A^T x[n_space]  (C.Tvmult)."""
        ny = self.n_im_dofs if transpose else self.n_space_dofs
        if out is None:
            out = torch.empty(ny, dtype=torch.float64, device=x.device) if isinstance(x, torch.Tensor) else np.empty(ny)
        self._chk(self.L.nm_spmv(self.h, int(transpose), _ptr(x)[0], _ptr(out)[0], _where(x, out)))
        return out

    def spmv_push(self, x: torch.Tensor, target_ptrs, multicast: bool = False, sync_producer: bool = True):
        """Ct.vmult fused with its cross-GPU reduction: add every non-empty row result into each of the
        device pointers `target_ptrs` (own + peer-mapped result vectors, or one multicast pointer)."""
        assert x.is_cuda
        if sync_producer:
            torch.cuda.current_stream().synchronize()
        arr = (C.c_void_p * len(target_ptrs))(*[C.c_void_p(int(p)) for p in target_ptrs])
        self._chk(self.L.nm_spmv_push(self.h, C.c_void_p(x.data_ptr()), len(target_ptrs), arr, int(multicast)))

    def spmv_push_owned(self, x: torch.Tensor, rank_ptrs, rows_per_rank: int, sync_producer: bool = True):
        """Ct.vmult with reduce-scatter semantics: every non-empty row result is added into the vector of the rank
        that owns the dof (contiguous ranges of `rows_per_rank` dofs); `rank_ptrs[k]` = rank k's result vector."""
        assert x.is_cuda
        if sync_producer:
            torch.cuda.current_stream().synchronize()
        arr = (C.c_void_p * len(rank_ptrs))(*[C.c_void_p(int(p)) for p in rank_ptrs])
        self._chk(self.L.nm_spmv_push_owned(self.h, C.c_void_p(x.data_ptr()), len(rank_ptrs), arr, int(rows_per_rank)))

    # ---- introspection ------------------------------------------------------------------
    def timings(self) -> dict:
        buf = (C.c_float * 16)()
        self._chk(self.L.nm_get_timings(self.h, buf, 16))
        return {name: float(buf[i]) for i, name in enumerate(PHASES)}

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        self._chk(self.L.nm_get_launch_count(self.h, C.byref(n)))
        return int(n.value)

    def device_bytes(self) -> int:
        b = C.c_uint64(0)
        self._chk(self.L.nm_device_bytes(self.h, C.byref(b)))
        return int(b.value)

    def stream(self) -> int:
        s = C.c_void_p()
        self._chk(self.L.nm_get_stream(self.h, C.byref(s)))
        return s.value or 0
