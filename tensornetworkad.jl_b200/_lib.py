"""ctypes binding of libtnad_b200.so -- the same C ABI a Julia `ccall` shim binds.

Fortran-ordered float64 NumPy arrays stand in for Julia arrays.  There is no fallback:
if the shared library is missing it is built with nvcc; if no sm_100 device is present
`Context()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libtnad_b200.so")

ERRORS = {1: "bad argument", 2: "CUDA error", 3: "out of device memory", 4: "SVD did not converge",
          5: "NCCL error", 6: "internal error"}

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int_p = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol declared in include/tnad.h
SIGNATURES = {
    "tnad_version": (C.c_int, []),
    "tnad_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "tnad_destroy": (C.c_int, [C.c_void_p]),
    "tnad_last_error": (C.c_char_p, [C.c_void_p]),
    "tnad_set_pointer_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "tnad_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "tnad_synchronize": (C.c_int, [C.c_void_p]),
    "tnad_launch_count": (C.c_int64, [C.c_void_p]),
    "tnad_reset_launch_count": (C.c_int, [C.c_void_p]),
    "tnad_dev_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "tnad_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tnad_dev_upload": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int64]),
    "tnad_dev_download": (C.c_int, [C.c_void_p, c_double_p, C.c_void_p, C.c_int64]),
    "tnad_contract": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, c_int64_p, C.c_int, C.c_void_p, c_int64_p,
                                C.c_int, C.c_double, C.c_double, C.c_void_p]),
    "tnad_contract_plan": (C.c_int, [C.c_char_p, c_int64_p, C.c_int, c_int64_p, C.c_int, c_int64_p]),
    "tnad_svd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, c_int_p]),
    "tnad_svd_sym": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, c_int_p]),
    "tnad_trg_svd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                               C.c_void_p, C.c_void_p, c_int_p]),
    "tnad_svd_back": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "tnad_trg_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                   c_double_p, C.POINTER(C.c_void_p)]),
    "tnad_trg_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "tnad_tape_free": (C.c_int, [C.c_void_p]),
    "tnad_ctmrg_init_random": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_ulonglong, C.c_void_p, C.c_void_p]),
    "tnad_ctmrg_init_raw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tnad_ctmrgstep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, c_double_p]),
    "tnad_ctmrg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                             c_int_p, c_double_p, C.POINTER(C.c_void_p)]),
    "tnad_ctmrg_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "tnad_expectationvalue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_int, c_double_p]),
    "tnad_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                              c_double_p, C.c_void_p, c_int_p]),
    "tnad_energy_fixedpoint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                         C.c_double, C.c_int, c_double_p, C.c_void_p, c_int_p, c_int_p]),
    "tnad_magnetisation_readout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_int, c_double_p]),
    "tnad_last_timing": (C.c_int, [C.c_void_p, c_double_p]),
    "tnad_ctmrgstep_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tnad_expectationvalue_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                 C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tnad_magnetisation_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tnad_trg_sweep": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                 c_int_p, c_double_p, C.c_void_p, C.c_char_p, C.c_int]),
    "tnad_svd_symmetrized": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, c_int_p]),
    "tnad_sytrd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tnad_symeig_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), c_int_p]),
    "tnad_symeig_backtransform": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "tnad_symeig_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tnad_symeig_free": (C.c_int, [C.c_void_p]),
    "tnad_sytrd2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tnad_stedc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "tnad_permute": (C.c_int, [C.c_void_p, C.c_void_p, c_int64_p, C.c_int, c_int_p, C.c_void_p]),
    "tnad_nccl_unique_id": (C.c_int, [C.c_char_p, C.c_void_p]),
    "tnad_comm_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
    "tnad_comm_destroy": (C.c_int, [C.c_void_p]),
    "tnad_ctmrg_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                                     C.POINTER(C.c_int), c_double_p]),
    "tnad_ctmrgstep_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         c_double_p, c_double_p]),
    "tnad_ctmrg_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tnad_host_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "tnad_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tnad_timer_start": (C.c_int, [C.c_void_p]),
    "tnad_timer_stop": (C.c_int, [C.c_void_p, c_double_p]),
    "tnad_set_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "tnad_kernel_timing": (C.c_int, [C.c_void_p, c_double_p, c_int64_p]),
    "tnad_dmma_peak": (C.c_int, [C.c_void_p, c_double_p]),
    "tnad_gemm_bench": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p]),
}

_lib = None


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libtnad_b200.so (building it in-tree with nvcc when absent) and declare all prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        # build_library() is incremental (mtime check per source): a no-op when the .so is current, a rebuild when
        # csrc/ was edited, an error when nvcc is missing and the .so is stale or absent
        try:
            from .build import build_library
            build_library()
        except Exception:
            if not os.path.exists(_LIBPATH):
                raise
    elif not os.path.exists(_LIBPATH):
        raise OSError(f"{_LIBPATH} is missing; run `python __graft_entry__.py build`")
    lib = C.CDLL(_LIBPATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class TnadError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tnad error {code} ({ERRORS.get(code, '?')}): {msg}")
        self.code = code


class DimensionMismatch(ValueError):
    """Mirror of Julia's DimensionMismatch (ipeps.jl:19-20)."""


def farray(x, shape=None):
    """float64, Fortran-ordered (Julia memory layout) copy/view of x."""
    a = np.asarray(x, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise DimensionMismatch(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return np.asfortranarray(a)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One device + one stream + workspace (tnad_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.tnad_create(int(device), C.byref(h))
        if rc != 0:
            msg = self.lib.tnad_last_error(None)
            raise TnadError(rc, msg.decode() if msg else "")
        self.h = h
        self.device = device
        self._tapes = weakref.WeakSet()

    def set_option(self, name: str, value=None):
        """A/B switch of this context (tnad_set_option); value None removes the entry."""
        self.check(self.lib.tnad_set_option(self.h, name.encode(), None if value is None else str(value).encode()))

    def close(self):
        if getattr(self, "h", None):
            for t in list(getattr(self, "_tapes", [])):     # tapes own device buffers of this context: free them first
                t.free()
            for p in getattr(self, "_pinned", []):
                self.lib.tnad_host_free(self.h, C.c_void_p(p))
            self._pinned = []
            self.lib.tnad_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            msg = self.lib.tnad_last_error(self.h)
            msg = msg.decode() if msg else ""
            if rc == 1 and ("size of tensor" in msg or "must be" in msg or "mismatch" in msg):
                raise DimensionMismatch(msg)
            raise TnadError(rc, msg)

    # ---- misc -----------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.tnad_launch_count(self.h))

    def reset_launch_count(self):
        self.lib.tnad_reset_launch_count(self.h)

    def last_timing(self):
        ms = (C.c_double * 8)()
        self.lib.tnad_last_timing(self.h, ms)
        return dict(total=ms[0], svd=ms[1], contractions=ms[2], backward=ms[3], svd_back=ms[4])

    def timer_start(self):
        self.check(self.lib.tnad_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double(0.0)
        self.check(self.lib.tnad_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def set_kernel_timing(self, enable: bool):
        self.check(self.lib.tnad_set_kernel_timing(self.h, 1 if enable else 0))

    def kernel_timing(self):
        ms = (C.c_double * 16)()
        cnt = (C.c_int64 * 16)()
        self.check(self.lib.tnad_kernel_timing(self.h, ms, cnt))
        names = {0: "m_update", 1: "eig_panel", 2: "q_update", 3: "gemm", 4: "other", 7: "chase", 8: "q2_stage",
                 9: "panel_qr", 10: "symm_y", 11: "rank64_update", 12: "stedc"}
        out = {n: dict(ms=ms[i], launches=int(cnt[i])) for i, n in names.items()}
        out["m_update"]["blocks"] = int(cnt[5])     # executed 64x64 two-sided block updates (2 x 2*64^3 flop each)
        out["q_update"]["slabs"] = int(cnt[6])      # executed 128x64 panel rotations (2*128*64*64 flop each)
        out["gemm"].update(flops=ms[13], flops_tma=ms[14], launches_tma=int(cnt[13]), launches_cp_async=int(cnt[14]))
        return out

    def dmma_peak(self) -> float:
        t = C.c_double(0.0)
        self.check(self.lib.tnad_dmma_peak(self.h, C.byref(t)))
        return t.value

    def gemm_bench(self, m, n, k, reps=5) -> float:
        t = C.c_double(0.0)
        self.check(self.lib.tnad_gemm_bench(self.h, int(m), int(n), int(k), int(reps), C.byref(t)))
        return t.value

    def host_alloc(self, shape) -> np.ndarray:
        """Pinned host array (Fortran order) backed by cudaMallocHost; freed with the context."""
        n = int(np.prod(shape))
        p = C.c_void_p()
        self.check(self.lib.tnad_host_alloc(self.h, n, C.byref(p)))
        buf = (C.c_double * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape, order="F")
        self._pinned = getattr(self, "_pinned", []) + [p.value]
        return arr

    def set_pointer_mode(self, mode: int):
        self.check(self.lib.tnad_set_pointer_mode(self.h, int(mode)))

    def dev_alloc(self, n: int) -> int:
        p = C.c_void_p()
        self.check(self.lib.tnad_dev_alloc(self.h, int(n), C.byref(p)))
        return p.value

    def dev_free(self, p: int):
        self.check(self.lib.tnad_dev_free(self.h, C.c_void_p(p)))

    def dev_upload(self, p: int, a: np.ndarray):
        a = farray(a)
        self.check(self.lib.tnad_dev_upload(self.h, C.c_void_p(p), a.ctypes.data_as(c_double_p), a.size))

    def dev_download(self, p: int, shape) -> np.ndarray:
        out = np.empty(shape, dtype=np.float64, order="F")
        self.check(self.lib.tnad_dev_download(self.h, out.ctypes.data_as(c_double_p), C.c_void_p(p), out.size))
        return out

    # ---- building blocks --------------------------------------------------------------------------
    def contract(self, spec: str, A, B, alpha=1.0, beta=0.0, Cin=None):
        A, B = farray(A), farray(B)
        lhs, out = spec.replace(" ", "").split("->")
        la, lb = lhs.split(",")
        ext = {}
        for l, n in zip(la, A.shape):
            ext[l] = n
        for l, n in zip(lb, B.shape):
            ext[l] = n
        cshape = tuple(ext[l] for l in out)
        Cout = farray(Cin, cshape).copy(order="F") if Cin is not None else np.zeros(cshape, order="F")
        da = (C.c_int64 * A.ndim)(*A.shape)
        db = (C.c_int64 * B.ndim)(*B.shape)
        self.check(self.lib.tnad_contract(self.h, spec.encode(), _p(A), da, A.ndim, _p(B), db, B.ndim,
                                          float(alpha), float(beta), _p(Cout)))
        return Cout

    def svd(self, A):
        A = farray(A)
        m, n = A.shape
        k = min(m, n)
        U = np.empty((m, k), order="F"); S = np.empty(k); V = np.empty((n, k), order="F")
        sw = C.c_int(0)
        self.check(self.lib.tnad_svd(self.h, _p(A), m, n, _p(U), _p(S), _p(V), C.byref(sw)))
        self.last_sweeps = sw.value
        return U, S, V

    def svd_sym(self, A):
        A = farray(A)
        n = A.shape[0]
        U = np.empty((n, n), order="F"); S = np.empty(n); V = np.empty((n, n), order="F")
        sw = C.c_int(0)
        self.check(self.lib.tnad_svd_sym(self.h, _p(A), n, _p(U), _p(S), _p(V), C.byref(sw)))
        self.last_sweeps = sw.value
        return U, S, V

    def trg_svd(self, t, dmax, tol):
        t = farray(t)
        d1, d2, d3, d4 = t.shape
        kmax = min(dmax, d1 * d2, d3 * d4)
        u = np.zeros((d1, d2, kmax), order="F"); v = np.zeros((kmax * d3 * d4,), order="F")
        k = C.c_int(0)
        self.check(self.lib.tnad_trg_svd(self.h, _p(t), d1, d2, d3, d4, int(dmax), float(tol), _p(u), _p(v),
                                         C.byref(k)))
        k = k.value
        return (np.asfortranarray(u[:, :, :k]),
                np.reshape(v[: k * d3 * d4], (k, d3, d4), order="F"))

    def svd_back(self, U, S, V, dU=None, dS=None, dV=None, eta=1e-40):
        U, S, V = farray(U), farray(S), farray(V)
        m, k = U.shape
        n = V.shape[0]
        dU = None if dU is None else farray(dU, (m, k))
        dS = None if dS is None else farray(dS, (k,))
        dV = None if dV is None else farray(dV, (n, k))
        if dU is None and dS is None and dV is None:
            return None                                     # trg.jl:73
        dA = np.empty((m, n), order="F")
        self.check(self.lib.tnad_svd_back(self.h, m, n, k, _p(U), _p(S), _p(V), _p(dU), _p(dS), _p(dV),
                                          float(eta), _p(dA)))
        return dA

    # ---- TRG ----------------------------------------------------------------------------------------
    def trg_forward(self, a, chi, niter, tol=1e-16, want_tape=False):
        a = farray(a)
        if a.ndim != 4 or a.shape[0] != a.shape[2] or a.shape[1] != a.shape[3]:
            raise DimensionMismatch(f"trg needs a (d0,d1,d0,d1) tensor, got {a.shape}")
        lnz = C.c_double(0.0)
        tape = C.c_void_p()
        self.check(self.lib.tnad_trg_forward(self.h, _p(a), a.shape[0], a.shape[1], int(chi), int(niter), float(tol),
                                             C.byref(lnz), C.byref(tape) if want_tape else None))
        return (lnz.value, Tape(self, tape, a.shape)) if want_tape else lnz.value

    def trg_backward(self, tape, dlnZ=1.0):
        da = np.empty(tape.shape, order="F")
        self.check(self.lib.tnad_trg_backward(self.h, tape.h, float(dlnZ), _p(da)))
        return da

    # ---- CTMRG --------------------------------------------------------------------------------------
    def ctmrg_init_raw(self, bulk, chi):
        bulk = farray(bulk)
        D = bulk.shape[0]
        corner = np.empty((chi, chi), order="F"); edge = np.empty((chi, D, chi), order="F")
        self.check(self.lib.tnad_ctmrg_init_raw(self.h, _p(bulk), D, int(chi), _p(corner), _p(edge)))
        return corner, edge

    def ctmrg_init_random(self, D, chi, seed):
        """:random environment generated on the device (reproducible from the seed)."""
        corner = np.empty((chi, chi), order="F"); edge = np.empty((chi, D, chi), order="F")
        self.check(self.lib.tnad_ctmrg_init_random(self.h, int(D), int(chi), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(corner), _p(edge)))
        return corner, edge

    def ctmrgstep(self, bulk, corner, edge):
        bulk = farray(bulk)
        D = bulk.shape[0]
        chi = np.shape(corner)[0]
        corner, edge = farray(corner, (chi, chi)), farray(edge, (chi, D, chi))
        co = np.empty_like(corner, order="F"); eo = np.empty_like(edge, order="F"); vals = np.empty(chi * D)
        self.check(self.lib.tnad_ctmrgstep(self.h, _p(bulk), D, chi, _p(corner), _p(edge), _p(co), _p(eo),
                                           vals.ctypes.data_as(c_double_p)))
        return co, eo, vals

    def ctmrg(self, bulk, corner, edge, tol, maxit, want_tape=False):
        bulk = farray(bulk)
        D = bulk.shape[0]
        chi = np.shape(corner)[0]
        if bulk.shape != (D, D, D, D):
            raise DimensionMismatch(f"bulk must be (D,D,D,D), got {bulk.shape}")
        co = farray(corner, (chi, chi)).copy(order="F"); ed = farray(edge, (chi, D, chi)).copy(order="F")
        vals = np.empty(chi * D)
        steps = C.c_int(0)
        tape = C.c_void_p()
        self.check(self.lib.tnad_ctmrg(self.h, _p(bulk), D, chi, _p(co), _p(ed), float(tol), int(maxit),
                                       C.byref(steps), vals.ctypes.data_as(c_double_p),
                                       C.byref(tape) if want_tape else None))
        out = (co, ed, vals, steps.value)
        return out + (Tape(self, tape, (D, chi)),) if want_tape else out

    def ctmrg_backward(self, tape, dcorner, dedge, want_init=False):
        D, chi = tape.shape
        dcorner, dedge = farray(dcorner, (chi, chi)), farray(dedge, (chi, D, chi))
        db = np.empty((D, D, D, D), order="F")
        dc0 = np.empty((chi, chi), order="F") if want_init else None
        de0 = np.empty((chi, D, chi), order="F") if want_init else None
        self.check(self.lib.tnad_ctmrg_backward(self.h, tape.h, _p(dcorner), _p(dedge), _p(db), _p(dc0), _p(de0)))
        return (db, dc0, de0) if want_init else db

    def ctmrgstep_backward(self, bulk, corner, edge, dcorner_out, dedge_out):
        """Pullback of one ctmrgstep: (dbulk, dcorner_in, dedge_in)."""
        bulk = farray(bulk)
        D, chi = bulk.shape[0], np.shape(corner)[0]
        corner, edge = farray(corner, (chi, chi)), farray(edge, (chi, D, chi))
        dco, deo = farray(dcorner_out, (chi, chi)), farray(dedge_out, (chi, D, chi))
        db = np.empty((D, D, D, D), order="F"); dc = np.empty((chi, chi), order="F"); de = np.empty((chi, D, chi), order="F")
        self.check(self.lib.tnad_ctmrgstep_backward(self.h, _p(bulk), D, chi, _p(corner), _p(edge), _p(dco), _p(deo),
                                                    _p(db), _p(dc), _p(de)))
        return db, dc, de

    # ---- energy --------------------------------------------------------------------------------------
    def expectationvalue_backward(self, h, ap, corner, edge, ybar=1.0):
        h, ap = farray(h), farray(ap)
        s, D, chi = h.shape[0], ap.shape[0], np.shape(corner)[0]
        corner, edge = farray(corner, (chi, chi)), farray(edge, (chi, D, chi))
        dap = np.empty(ap.shape, order="F"); dc = np.empty((chi, chi), order="F"); de = np.empty((chi, D, chi), order="F")
        self.check(self.lib.tnad_expectationvalue_backward(self.h, _p(h), _p(ap), D, s, _p(corner), _p(edge), chi,
                                                           float(ybar), _p(dap), _p(dc), _p(de)))
        return dap, dc, de

    def expectationvalue(self, h, ap, corner, edge):
        h, ap = farray(h), farray(ap)
        s, D, chi = h.shape[0], ap.shape[0], np.shape(corner)[0]
        corner, edge = farray(corner, (chi, chi)), farray(edge, (chi, D, chi))
        e = C.c_double(0.0)
        self.check(self.lib.tnad_expectationvalue(self.h, _p(h), _p(ap), D, s, _p(corner), _p(edge), chi, C.byref(e)))
        return e.value

    def energy(self, h, A, chi, tol, maxit, grad=False):
        h, A = farray(h), farray(A)
        if A.ndim != 5 or not (A.shape[0] == A.shape[1] == A.shape[2] == A.shape[3]):
            raise DimensionMismatch(f"size of tensor error, should be `(d, d, d, d, s)`, got {A.shape}.")
        d, s = A.shape[0], A.shape[4]
        if h.shape != (s, s, s, s):
            raise DimensionMismatch(f"h must be ({s},{s},{s},{s}), got {h.shape}")
        e = C.c_double(0.0)
        steps = C.c_int(0)
        g = np.empty(A.shape, order="F") if grad else None
        self.check(self.lib.tnad_energy(self.h, _p(h), _p(A), d, s, int(chi), float(tol), int(maxit), C.byref(e),
                                        _p(g), C.byref(steps)))
        self.last_steps = steps.value
        return (e.value, g) if grad else e.value

    def energy_fixedpoint(self, h, A, chi, tol, maxit, bwd_tol=1e-12, bwd_maxit=500):
        """Energy and its implicit (fixed-point) gradient: (e, grad); self.last_steps / self.last_bwd_iters are set."""
        h, A = farray(h), farray(A)
        d, s = A.shape[0], A.shape[4]
        e = C.c_double(0.0)
        steps, it = C.c_int(0), C.c_int(0)
        g = np.empty(A.shape, order="F")
        self.check(self.lib.tnad_energy_fixedpoint(self.h, _p(h), _p(A), d, s, int(chi), float(tol), int(maxit), float(bwd_tol),
                                                   int(bwd_maxit), C.byref(e), _p(g), C.byref(steps), C.byref(it)))
        self.last_steps, self.last_bwd_iters = steps.value, it.value
        return e.value, g

    def sytrd(self, a):
        """A = Q tridiag(d, e) Q' (stage of the direct eigensolver)."""
        a = farray(a)
        n = a.shape[0]
        d = np.empty(n); e = np.empty(max(n - 1, 1)); q = np.empty((n, n), order="F")
        self.check(self.lib.tnad_sytrd(self.h, _p(a), n, _p(d), _p(e), _p(q)))
        return d, e[:n - 1], q

    def sytrd2(self, a):
        """A = Q tridiag(d, e) Q' through the two-stage route; also returns the intermediate band (33 x n)."""
        a = farray(a)
        n = a.shape[0]
        d = np.empty(n); e = np.empty(max(n - 1, 1)); q = np.empty((n, n), order="F"); band = np.empty((33, n), order="F")
        self.check(self.lib.tnad_sytrd2(self.h, _p(a), n, _p(d), _p(e), _p(q), _p(band)))
        return d, e[:n - 1], q, band

    def stedc(self, d, e):
        d = np.ascontiguousarray(d, dtype=np.float64); e = np.ascontiguousarray(e, dtype=np.float64)
        n = d.size
        lam = np.empty(n); z = np.empty((n, n), order="F")
        self.check(self.lib.tnad_stedc(self.h, _p(d), _p(e) if n > 1 else None, n, _p(lam), _p(z)))
        return lam, z

    # ---- raw device-pointer calls (pointer mode DEVICE) used by sharded.py ---------------------------------
    def dev_contract(self, spec, pa, dims_a, pb, dims_b, pc, alpha=1.0, beta=0.0):
        da = (C.c_int64 * len(dims_a))(*dims_a)
        db = (C.c_int64 * len(dims_b))(*dims_b)
        self.check(self.lib.tnad_contract(self.h, spec.encode(), C.c_void_p(pa), da, len(dims_a), C.c_void_p(pb), db,
                                          len(dims_b), float(alpha), float(beta), C.c_void_p(pc)))

    def dev_permute(self, pin, dims, perm, pout):
        dd = (C.c_int64 * len(dims))(*dims)
        pp = (C.c_int * len(perm))(*perm)
        self.check(self.lib.tnad_permute(self.h, C.c_void_p(pin), dd, len(dims), pp, C.c_void_p(pout)))

    def dev_svd_symmetrized(self, pa, n, pu, ps, pv):
        sw = C.c_int(0)
        self.check(self.lib.tnad_svd_symmetrized(self.h, C.c_void_p(pa), int(n), C.c_void_p(pu), C.c_void_p(ps), C.c_void_p(pv),
                                         C.byref(sw)))
        return sw.value

    # three-phase symmetric eigensolver on device pointers (the back-transformation is shared between ranks)
    def dev_symeig_reduce(self, pa, n, add_transpose=True):
        h = C.c_void_p()
        N = C.c_int(0)
        self.check(self.lib.tnad_symeig_reduce(self.h, C.c_void_p(pa), int(n), 1 if add_transpose else 0, C.byref(h), C.byref(N)))
        return h, N.value

    def dev_symeig_backtransform(self, h, col0, ncols, pz):
        self.check(self.lib.tnad_symeig_backtransform(self.h, h, int(col0), int(ncols), C.c_void_p(pz)))

    def dev_symeig_finish(self, h, pzfull, pu, ps, pv):
        self.check(self.lib.tnad_symeig_finish(self.h, h, C.c_void_p(pzfull), C.c_void_p(pu), C.c_void_p(ps), C.c_void_p(pv)))

    def symeig_free(self, h):
        self.lib.tnad_symeig_free(h)

    # ---- chi-sharded ctmrgstep inside the library (NCCL loaded at run time) ----------------------------------
    @staticmethod
    def nccl_library_path():
        """libnccl.so.2 of the nvidia-nccl wheel torch uses (None: let the library find a mapped / system copy)."""
        try:
            import nvidia.nccl as _n
            p = os.path.join(os.path.dirname(_n.__file__), "lib", "libnccl.so.2")
            return p if os.path.exists(p) else None
        except Exception:   # noqa: BLE001
            return None

    def nccl_unique_id(self) -> bytes:
        buf = (C.c_ubyte * 128)()
        path = self.nccl_library_path()
        rc = self.lib.tnad_nccl_unique_id(path.encode() if path else None, buf)
        if rc != 0:
            raise TnadError(rc, "tnad_nccl_unique_id failed (is NCCL loadable?)")
        return bytes(buf)

    def comm_init(self, unique_id, rank: int, world: int):
        path = self.nccl_library_path()
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        self.check(self.lib.tnad_comm_init(self.h, path.encode() if path else None, buf, int(rank), int(world)))

    def comm_destroy(self):
        self.lib.tnad_comm_destroy(self.h)

    def ctmrgstep_sharded(self, bulk, corner, edge, timing=False):
        """Host arrays in, host arrays out (every rank passes the same environment); returns corner, edge, vals[, ms]."""
        D, chi = bulk.shape[0], corner.shape[0]
        b, co, ed = farray(bulk), farray(corner), farray(edge)
        cn, en = np.empty((chi, chi), order="F"), np.empty((chi, D, chi), order="F")
        vals = np.empty(chi * D)
        ms = (C.c_double * 3)()
        self.check(self.lib.tnad_ctmrgstep_sharded(self.h, _p(b), D, _p(co), _p(ed), chi, _p(cn), _p(en),
                                                   vals.ctypes.data_as(c_double_p), ms if timing else None))
        return (cn, en, vals, list(ms)) if timing else (cn, en, vals)

    def ctmrg_sharded(self, bulk, corner, edge, tol, maxit):
        """ctmrg over the communicator (forward): returns corner, edge, vals, steps."""
        D, chi = bulk.shape[0], corner.shape[0]
        b = farray(bulk)
        co, ed = np.array(corner, dtype=np.float64, order="F", copy=True), np.array(edge, dtype=np.float64, order="F", copy=True)
        vals = np.empty(chi * D)
        ns = C.c_int(0)
        self.check(self.lib.tnad_ctmrg_sharded(self.h, _p(b), D, chi, _p(co), _p(ed), float(tol), int(maxit), C.byref(ns),
                                               vals.ctypes.data_as(c_double_p)))
        return co, ed, vals, int(ns.value)

    def dev_ctmrgstep_sharded(self, pbulk, D, pcorner, pedge, chi, pcorner_out, pedge_out, timing=True):
        """Device pointers (pointer mode DEVICE must be set by the caller); returns vals (host) and the three device times."""
        vals = np.empty(chi * D)
        ms = (C.c_double * 3)()
        self.check(self.lib.tnad_ctmrgstep_sharded(self.h, C.c_void_p(pbulk), int(D), C.c_void_p(pcorner), C.c_void_p(pedge), int(chi),
                                                   C.c_void_p(pcorner_out), C.c_void_p(pedge_out), vals.ctypes.data_as(c_double_p),
                                                   ms if timing else None))
        return vals, list(ms)

    def dev_ctmrg_finish(self, pc1, pe1, D, chi, pco, peo):
        self.check(self.lib.tnad_ctmrg_finish(self.h, C.c_void_p(pc1), C.c_void_p(pe1), int(D), int(chi),
                                              C.c_void_p(pco), C.c_void_p(peo)))

    def energy_device(self, h_dptr: int, A_dptr: int, d: int, s: int, chi: int, tol: float, maxit: int,
                      grad_dptr: int = 0):
        """tnad_energy with inputs/outputs already resident in HBM (pointer mode DEVICE)."""
        e = C.c_double(0.0)
        steps = C.c_int(0)
        self.set_pointer_mode(1)
        try:
            self.check(self.lib.tnad_energy(self.h, C.c_void_p(h_dptr), C.c_void_p(A_dptr), d, s, int(chi), float(tol),
                                            int(maxit), C.byref(e), C.c_void_p(grad_dptr) if grad_dptr else None,
                                            C.byref(steps)))
        finally:
            self.set_pointer_mode(0)
        self.last_steps = steps.value
        return e.value

    def magnetisation_readout(self, a, m, corner, edge):
        a, m = farray(a), farray(m)
        D, chi = a.shape[0], np.shape(corner)[0]
        corner, edge = farray(corner, (chi, chi)), farray(edge, (chi, D, chi))
        mag = C.c_double(0.0)
        self.check(self.lib.tnad_magnetisation_readout(self.h, _p(a), _p(m), D, _p(corner), _p(edge), chi,
                                                       C.byref(mag)))
        return mag.value

    def magnetisation_backward(self, a, m, corner, edge, ybar=1.0):
        """Pullback of the read-out: (da, dm, dcorner, dedge)."""
        a, m = farray(a), farray(m)
        D, chi = a.shape[0], np.shape(corner)[0]
        corner, edge = farray(corner, (chi, chi)), farray(edge, (chi, D, chi))
        da = np.empty(a.shape, order="F"); dm = np.empty(a.shape, order="F")
        dc = np.empty((chi, chi), order="F"); de = np.empty((chi, D, chi), order="F")
        self.check(self.lib.tnad_magnetisation_backward(self.h, _p(a), _p(m), D, _p(corner), _p(edge), chi, float(ybar),
                                                        _p(da), _p(dm), _p(dc), _p(de)))
        return da, dm, dc, de


def trg_sweep(tensors, chi, niter, tol=1e-16, ngpu=1, devices=None, grad=False):
    """tnad_trg_sweep: independent TRG instances fanned out over `ngpu` devices inside the library (one context and
    one host thread per device).  tensors: (ninst, d0, d1, d0, d1).  Returns lnZ (ninst) [, grads like tensors]."""
    lib = load_library()
    ts = [farray(t) for t in tensors]
    ninst = len(ts)
    d0, d1 = ts[0].shape[0], ts[0].shape[1]
    flat = np.concatenate([t.ravel(order="F") for t in ts]) if ninst else np.zeros(0)
    lnz = np.zeros(ninst)
    g = np.zeros(flat.size) if grad else None
    dev = (C.c_int * ngpu)(*devices) if devices is not None else None
    err = C.create_string_buffer(512)
    rc = lib.tnad_trg_sweep(_p(flat), ninst, d0, d1, int(chi), int(niter), float(tol), int(ngpu), dev,
                            lnz.ctypes.data_as(c_double_p), _p(g), err, 512)
    if rc != 0:
        raise TnadError(rc, err.value.decode())
    if grad:
        return lnz, [np.reshape(g[i * ts[0].size:(i + 1) * ts[0].size], ts[0].shape, order="F") for i in range(ninst)]
    return lnz


class Tape:
    """Opaque device-resident record of a forward pass (tnad_tape)."""

    def __init__(self, ctx: Context, h, shape):
        self.ctx, self.h, self.shape = ctx, h, tuple(shape)
        ctx._tapes.add(self)

    def free(self):
        if self.h:
            self.ctx.lib.tnad_tape_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def contract_plan(spec: str, dimsA, dimsB):
    """Host-only: the GEMM descriptor for a pairwise einsum (used by the CPU tests)."""
    lib = load_library()
    da = (C.c_int64 * len(dimsA))(*dimsA)
    db = (C.c_int64 * len(dimsB))(*dimsB)
    plan = (C.c_int64 * 128)()
    rc = lib.tnad_contract_plan(spec.encode(), da, len(dimsA), db, len(dimsB), plan)
    if rc != 0:
        raise TnadError(rc, f"contract_plan({spec})")
    p = list(plan)
    names = ["am", "ak", "bk", "bn", "cm", "cn", "ab", "bb", "cb"]
    sets = {}
    for i, nm in enumerate(names):
        q = p[8 + 9 * i: 8 + 9 * (i + 1)]
        nl = q[0]
        sets[nm] = [(q[1 + l], q[5 + l]) for l in range(nl)]
    return dict(M=p[0], N=p[1], K=p[2], batch=p[3], a_kfast=p[4], b_kfast=p[5], a_vec=p[6], b_vec=p[7], **sets)
