"""QEqSolver: the host-side mirror of /root/reference/src/solver/implementation.rs:60-266.

Everything numerical happens behind the C ABI (include/cheq_b200.h) on the GPU.  This class does what the Rust
shim does: NoAtoms check, `fetch_element_data` (:348-362, first missing Z -> ParameterNotFound), the point-charge
element lookup (:378-387), the `ExternalPotential::is_empty` short-circuit (:245-247) and one FFI call.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _ffi
from .errors import CheqError, CudaError, LinalgError, NoAtoms, NotConverged, ParameterNotFound
from .options import BasisType, SolverOptions
from .params import Parameters
from .types import CalculationResult, ExternalPotential, atom_fields

_DAMPING_KIND = {"none": 0, "fixed": 1, "auto": 2}


class Context:
    """Owns one cheq_ctx (one CUDA stream + workspace).  Thread-safe: the C side serialises calls."""

    def __init__(self, device: int = 0):
        self._lib = _ffi.lib()
        handle = C.c_void_p()
        rc = self._lib.cheq_ctx_create(int(device), C.byref(handle))
        if rc != _ffi.OK:
            raise CudaError(f"cheq_ctx_create(device={device}) failed with status {rc}: "
                            "a CUDA device is required (there is no CPU fallback)")
        self.handle = handle
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cheq_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        return (self._lib.cheq_last_error(self.handle) or b"").decode()

    def stats(self) -> dict:
        s = _ffi.Stats()
        self._lib.cheq_get_stats(self.handle, C.byref(s))
        return s.as_dict()

    def scf_trace(self):
        """Per SCF iteration of the last single-frame solve: (delta, mode, gmres_steps); mode 0 = the hydrogen block
        was refactorised, 1 = frozen factor + GMRES."""
        cap = 4096
        d = (C.c_double * cap)()
        m = (C.c_int32 * cap)()
        k = (C.c_int32 * cap)()
        cnt = C.c_uint32()
        self._lib.cheq_get_scf_trace(self.handle, cap, d, m, k, C.byref(cnt))
        n = min(cap, cnt.value)
        return [(d[i], m[i], k[i]) for i in range(n)]

    def set_krylov(self, enabled: bool = True, min_hydrogen: int = 0, freeze_ev: float = 0.0,
                   force_iteration: int = 0):
        rc = self._lib.cheq_set_krylov(self.handle, 1 if enabled else 0, int(min_hydrogen), float(freeze_ev),
                                       int(force_iteration))
        if rc != _ffi.OK:
            raise CheqError(f"cheq_set_krylov failed with status {rc}")

    def set_ozaki(self, slices: int = -1, slices_r: int = -1, min_k: int = 0):
        """FP64-equivalent updates on the tcgen05 tensor cores (int8 slice products): slices = 0 switches the path off,
        4..8 slices per operand entry (7 = fp64 grade, the default); negative / zero arguments keep the setting."""
        rc = self._lib.cheq_set_ozaki(self.handle, int(slices), int(slices_r), int(min_k))
        if rc != _ffi.OK:
            raise CheqError(f"cheq_set_ozaki failed with status {rc}: {self.last_error()}")


_default_ctx = {}
_default_lock = threading.Lock()


def default_context(device: int = 0) -> Context:
    with _default_lock:
        if device not in _default_ctx:
            _default_ctx[device] = Context(device)
        return _default_ctx[device]


def make_options(o: SolverOptions) -> _ffi.Options:
    opt = _ffi.Options()
    opt.tolerance = float(o.tolerance)
    opt.lambda_scale = float(o.lambda_scale)
    opt.damping_value = float(o.damping.value)
    opt.max_iterations = int(o.max_iterations)
    opt.hydrogen_scf = 1 if o.hydrogen_scf else 0
    opt.basis = int(o.basis_type.value)
    opt.damping_kind = _DAMPING_KIND[o.damping.kind]
    return opt


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def raise_for_status(ctx: Context, rc: int, max_iterations: int = 0, delta: float = 0.0):
    if rc == _ffi.OK:
        return
    msg = ctx.last_error()
    if rc == _ffi.NO_ATOMS:
        raise NoAtoms()
    if rc == _ffi.NOT_CONVERGED:
        raise NotConverged(max_iterations, delta)
    if rc == _ffi.LINALG:
        raise LinalgError(msg)
    if rc in (_ffi.CUDA, _ffi.OUT_OF_MEMORY):
        raise CudaError(msg)
    raise CheqError(f"status {rc}: {msg}")


class QEqSolver:
    """implementation.rs:60-65: `parameters` + `options`.  `context` selects the GPU (default: device 0)."""

    def __init__(self, parameters: Parameters, options: SolverOptions | None = None, context: Context | None = None):
        self.parameters = parameters
        self.options = options or SolverOptions()
        self._context = context

    @staticmethod
    def new(parameters: Parameters) -> "QEqSolver":
        return QEqSolver(parameters)

    def with_options(self, options: SolverOptions) -> "QEqSolver":
        return QEqSolver(self.parameters, options, self._context)

    def with_context(self, context: Context) -> "QEqSolver":
        return QEqSolver(self.parameters, self.options, context)

    @property
    def context(self) -> Context:
        if self._context is None:
            self._context = default_context(0)
        return self._context

    # -- fetch_element_data (implementation.rs:348-362) -----------------------------------------------------
    def _gather(self, atomic_numbers):
        el = self.parameters.elements
        n = len(atomic_numbers)
        chi = np.empty(n)
        hard = np.empty(n)
        rad = np.empty(n)
        nq = np.empty(n, dtype=np.uint8)
        for i, z in enumerate(atomic_numbers):
            d = el.get(int(z))
            if d is None:
                raise ParameterNotFound(int(z))
            chi[i], hard[i], rad[i], nq[i] = d.electronegativity, d.hardness, d.radius, d.principal_quantum_number
        return chi, hard, rad, nq

    def solve(self, atoms, total_charge: float) -> CalculationResult:
        """implementation.rs:174-188."""
        return self._solve(atoms, total_charge, None)

    def solve_in_field(self, atoms, total_charge: float, external: ExternalPotential) -> CalculationResult:
        """implementation.rs:239-266."""
        if external.is_empty():
            return self._solve(atoms, total_charge, None)
        return self._solve(atoms, total_charge, external)

    def _solve(self, atoms, total_charge, external):
        n = len(atoms)
        if n == 0:
            raise NoAtoms()
        fields = [atom_fields(a) for a in atoms]
        zs = [f[0] for f in fields]
        xyz = np.ascontiguousarray([f[1] for f in fields], dtype=np.float64)
        chi, hard, rad, nq = self._gather(zs)
        ext_struct = None
        keep = []
        if external is not None:
            pcs = external.point_charges()
            _, _, prad, pnq = self._gather([pc.atomic_number for pc in pcs])
            pxyz = np.ascontiguousarray([[float(c) for c in pc.position] for pc in pcs], dtype=np.float64).reshape(-1, 3)
            pq = np.ascontiguousarray([float(pc.charge) for pc in pcs], dtype=np.float64)
            ext_struct = _ffi.External()
            ext_struct.num_point_charges = len(pcs)
            ext_struct.xyz = _ptr(pxyz)
            ext_struct.charge = _ptr(pq)
            ext_struct.radius = _ptr(prad)
            ext_struct.n = _ptr(pnq)
            f = external.uniform_field()
            ext_struct.field[0], ext_struct.field[1], ext_struct.field[2] = f[0], f[1], f[2]
            keep = [pxyz, pq, prad, pnq]
        res = self.solve_arrays(xyz, chi, hard, rad, nq, total_charge, ext_struct)
        del keep
        return res

    def solve_arrays(self, xyz, chi, hard, rad, nq, total_charge, ext_struct=None,
                     iterative: bool = False) -> CalculationResult:
        """One FFI call on already-resolved per-atom arrays (host numpy).  iterative=True forces the matrix-free
        path (cheq_solve_iterative); cheq_solve picks it by itself when the dense matrix cannot fit the device."""
        ctx = self.context
        n = len(chi)
        opt = make_options(self.options)
        charges = np.empty(n)
        pot = C.c_double()
        its = C.c_uint32()
        delta = C.c_double()
        fn = ctx._lib.cheq_solve_iterative if iterative else ctx._lib.cheq_solve
        rc = fn(ctx.handle, n, _ptr(xyz), _ptr(chi), _ptr(hard), _ptr(rad), _ptr(nq),
                                 float(total_charge), C.byref(opt),
                                 C.byref(ext_struct) if ext_struct is not None else None, _ptr(charges),
                                 C.byref(pot), C.byref(its), C.byref(delta))
        raise_for_status(ctx, rc, self.options.max_iterations, delta.value)
        return CalculationResult(charges.tolist(), pot.value, int(its.value))

    def solve_batch(self, atomic_numbers, frames_xyz, total_charge: float = 0.0):
        """Frames of one topology.  frames_xyz: (F, N, 3).

        Returns (charges (F, N), potentials (F,), iterations (F,), status (F,))."""
        ctx = self.context
        xyz = np.ascontiguousarray(frames_xyz, dtype=np.float64)
        if xyz.ndim != 3 or xyz.shape[2] != 3:
            raise ValueError("frames_xyz must have shape (frames, atoms, 3)")
        F, n = xyz.shape[0], xyz.shape[1]
        if n == 0:
            raise NoAtoms()
        chi, hard, rad, nq = self._gather(list(atomic_numbers))
        opt = make_options(self.options)
        charges = np.empty((F, n))
        pot = np.empty(F)
        its = np.empty(F, dtype=np.uint32)
        delta = np.empty(F)
        status = np.empty(F, dtype=np.int32)
        rc = ctx._lib.cheq_solve_batch(ctx.handle, F, n, _ptr(xyz), _ptr(chi), _ptr(hard), _ptr(rad), _ptr(nq),
                                       float(total_charge), C.byref(opt), _ptr(charges), _ptr(pot), _ptr(its),
                                       _ptr(delta), _ptr(status))
        raise_for_status(ctx, rc, self.options.max_iterations, 0.0)
        return charges, pot, its, status
