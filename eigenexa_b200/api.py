"""Host-side mirror of EigenExa's public interface over the C ABI (include/eigenexa_b200.h).

Same names, argument meaning and error behaviour as the reference's Fortran module
``eigen_libs_mod`` (src/eigen_libs.F:70-216) and its C binding (C/EigenExa.h:12-46):
``eigen_init``, ``eigen_free``, ``eigen_get_matdims``, ``eigen_s``, ``eigen_sx``,
``eigen_get_procs``, ``eigen_get_id``, ``eigen_get_version`` and the index helpers.
Arrays are NumPy, column-major (``order='F'``), in the 2D cyclic local layout.

There is no CPU fallback: loading fails loudly when the CUDA library has not been built,
and ``eigen_init`` reports an error when no Blackwell GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libeigenexa_b200.so")
_lib = None


class _Comm(C.Structure):
    _fields_ = [("rank", C.c_int), ("nranks", C.c_int), ("device", C.c_int), ("reserved", C.c_int),
                ("unique_id", C.c_ubyte * 128)]


def lib() -> C.CDLL:
    """The C-ABI shared library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  eigenexa_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.eigen_init.argtypes = [C.POINTER(_Comm), C.c_char_p]
        L.eigen_init.restype = None
        L.eigen_free.restype = None
        for f in (L.eigen_s, L.eigen_sx):
            f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                          C.c_char_p]
            f.restype = None
        L.eigenexa_b200_eigen_s_dev.argtypes = L.eigen_s.argtypes
        L.eigenexa_b200_eigen_s_dev.restype = C.c_int
        L.eigenexa_b200_eigen_sx_dev.argtypes = L.eigen_s.argtypes
        L.eigenexa_b200_eigen_sx_dev.restype = C.c_int
        L.eigen_get_matdims.argtypes = [C.c_int, ip, ip, C.c_int, C.c_int, C.c_char_p]
        L.eigen_get_procs.argtypes = [ip, ip, ip]
        L.eigen_get_id.argtypes = [ip, ip, ip]
        L.eigen_get_version.argtypes = [ip, C.c_char_p, C.c_char_p]
        L.eigen_get_errinfo.argtypes = [ip]
        L.eigen_memory_internal.argtypes = [C.c_int] * 5
        L.eigen_memory_internal.restype = C.c_int64
        for name in ("eigen_loop_start", "eigen_loop_end", "eigen_translate_l2g", "eigen_translate_g2l",
                     "eigen_owner_node", "eigen_owner_index"):
            f = getattr(L, name)
            f.argtypes = [C.c_int] * 3
            f.restype = C.c_int
        L.eigenexa_b200_get_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
        L.eigenexa_b200_trd.argtypes = [C.c_int, dp, C.c_int, dp, dp, C.c_int]
        L.eigenexa_b200_trbakwy.argtypes = [C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int]
        L.eigenexa_b200_dc.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, C.c_int]
        L.eigenexa_b200_bisect.argtypes = [C.c_int, dp, dp, dp]
        L.eigenexa_b200_prd.argtypes = [C.c_int, dp, C.c_int, dp, dp, C.c_int, C.c_int]
        L.eigenexa_b200_dcx.argtypes = [C.c_int, C.c_int, dp, dp, C.c_int, dp, dp, C.c_int]
        L.eigenexa_b200_bisect2.argtypes = [C.c_int, dp, dp, C.c_int, dp]
        L.eigenexa_b200_trbakwy_nb.argtypes = [C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_int]
        L.eigenexa_b200_mat_set_dev.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64]
        L.eigenexa_b200_mat_set_host.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64]
        L.eigenexa_b200_ev_test_dev.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                dp]
        L.eigenexa_b200_dgemm_dev.argtypes = [C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                              C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
        L.eigenexa_b200_dgemm_tri_dev.argtypes = list(L.eigenexa_b200_dgemm_dev.argtypes) + [C.c_int] * 4
        L.eigenexa_b200_stream.restype = C.c_void_p
        L.eigenexa_b200_launch_count.argtypes = [C.c_int]
        L.eigenexa_b200_launch_count.restype = C.c_int64
        L.eigenexa_b200_last_timings.argtypes = [dp, C.c_int]
        L.eigenexa_b200_set_profiling.argtypes = [C.c_int]
        L.eigenexa_b200_symv_trace.argtypes = [C.POINTER(C.c_float), C.c_int]
        L.eigenexa_b200_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _dp(x: np.ndarray):
    if x.dtype != np.float64:
        raise TypeError("real(8) array expected")
    return x.ctypes.data_as(C.POINTER(C.c_double))


def _check_f(x: np.ndarray, name: str):
    if x.ndim == 2 and not x.flags.f_contiguous:
        raise ValueError(f"{name} must be column-major (order='F')")


# ---------------------------------------------------------------------------------------
# eigen_libs_mod
# ---------------------------------------------------------------------------------------
def get_unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    if lib().eigenexa_b200_get_unique_id(buf) != 0:
        raise RuntimeError(last_error())
    return bytes(buf)


def eigen_init(comm=None, order: str = "C") -> None:
    """``comm``: None (single rank) or (rank, nranks, unique_id_bytes[, device])."""
    c = _Comm()
    if comm is None:
        c.rank, c.nranks, c.device = 0, 1, -1
    else:
        c.rank, c.nranks = int(comm[0]), int(comm[1])
        uid = comm[2]
        c.device = int(comm[3]) if len(comm) > 3 else -1
        if uid is not None:
            C.memmove(c.unique_id, bytes(uid), 128)
    lib().eigen_init(C.byref(c), order.encode()[:1])


def eigen_init_torch(order: str = "C") -> None:
    """Bootstrap from a ``torch.distributed`` process group (one rank per GPU): rank 0 makes
    the NCCL id, the group broadcasts it (the job MPI_Bcast does in an MPI application)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return eigen_init(None, order)
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return eigen_init((0, 1, None, torch.cuda.current_device()), order)
    obj = [get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    eigen_init((rank, world, obj[0], torch.cuda.current_device()), order)


def eigen_free() -> None:
    lib().eigen_free()


def eigen_get_matdims(n: int, m_forward: int = 48, m_backward: int = 128, mode: str = "O"):
    nx, ny = C.c_int(), C.c_int()
    lib().eigen_get_matdims(n, C.byref(nx), C.byref(ny), m_forward, m_backward, mode.encode()[:1])
    return nx.value, ny.value


def eigen_get_procs():
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib().eigen_get_procs(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def eigen_get_id():
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib().eigen_get_id(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def eigen_get_version():
    v = C.c_int()
    date, code = C.create_string_buffer(64), C.create_string_buffer(64)
    lib().eigen_get_version(C.byref(v), date, code)
    return v.value, date.value.decode(), code.value.decode()


def eigen_get_errinfo() -> int:
    v = C.c_int()
    lib().eigen_get_errinfo(C.byref(v))
    return v.value


def eigen_loop_start(i, nnod, inod): return lib().eigen_loop_start(i, nnod, inod)
def eigen_loop_end(i, nnod, inod): return lib().eigen_loop_end(i, nnod, inod)
def eigen_translate_l2g(i, nnod, inod): return lib().eigen_translate_l2g(i, nnod, inod)
def eigen_translate_g2l(i, nnod, inod): return lib().eigen_translate_g2l(i, nnod, inod)
def eigen_owner_node(i, nnod, inod): return lib().eigen_owner_node(i, nnod, inod)
def eigen_owner_index(i, nnod, inod): return lib().eigen_owner_index(i, nnod, inod)


def _solve(fn, n, a, w, z, nvec, m_forward, m_backward, mode):
    _check_f(a, "a")
    lda = a.shape[0] if a.ndim == 2 else n
    if z is None:
        zp, ldz = None, lda
    else:
        _check_f(z, "z")
        zp, ldz = z.ctypes.data, (z.shape[0] if z.ndim == 2 else n)
    fn(n, n if nvec is None else nvec, a.ctypes.data, lda, w.ctypes.data, zp, ldz, m_forward, m_backward,
       mode.encode()[:1])


def eigen_s(n, a, w, z, nvec=None, m_forward=48, m_backward=128, mode="A") -> None:
    """eigen_s(n, nvec, a, lda, w, z, ldz, m_forward, m_backward, mode), src/eigen_libs.F:150-202."""
    _solve(lib().eigen_s, n, a, w, z, nvec, m_forward, m_backward, mode)


def eigen_sx(n, a, w, z, nvec=None, m_forward=48, m_backward=128, mode="A") -> None:
    _solve(lib().eigen_sx, n, a, w, z, nvec, m_forward, m_backward, mode)


def eigen_s_dev(n, a_ptr, lda, w_ptr, z_ptr, ldz, nvec=None, m_forward=48, m_backward=128, mode="A") -> None:
    """Device-pointer form (torch ``tensor.data_ptr()``): inputs already resident in HBM."""
    rc = lib().eigenexa_b200_eigen_s_dev(n, n if nvec is None else nvec, a_ptr, lda, w_ptr, z_ptr, ldz, m_forward,
                                         m_backward, mode.encode()[:1])
    if rc != 0:
        raise RuntimeError(last_error())


def eigen_sx_dev(n, a_ptr, lda, w_ptr, z_ptr, ldz, nvec=None, m_forward=48, m_backward=128, mode="A") -> None:
    """eigen_sx with device pointers (penta-diagonal path)."""
    rc = lib().eigenexa_b200_eigen_sx_dev(n, n if nvec is None else nvec, a_ptr, lda, w_ptr, z_ptr, ldz, m_forward,
                                          m_backward, mode.encode()[:1])
    if rc != 0:
        raise RuntimeError(last_error())


# ---------------------------------------------------------------------------------------
# stage-level procedures (eigen_trd_mod, trbakwy4_mod, dc2, bisect)
# ---------------------------------------------------------------------------------------
def eigen_trd(n, a, m_forward=48):
    """eigen_trd(n, a, lda, d, e, m) (src/eigen_trd.F:82): returns (d, e); a <- reflectors."""
    _check_f(a, "a")
    d, e = np.zeros(n), np.zeros(n)
    rc = lib().eigenexa_b200_trd(n, _dp(a), a.shape[0], _dp(d), _dp(e), m_forward)
    if rc != 0:
        raise RuntimeError(f"eigen_trd rc={rc}: {last_error()}")
    return d, e


def eigen_prd(n, a, m_forward=48):
    """eigen_prd(n, a, lda, d, e, ne, m) (src/eigen_prd.F:80): returns (d, e1, e2); a <- reflectors."""
    _check_f(a, "a")
    d, e = np.zeros(n), np.zeros((2, n))
    rc = lib().eigenexa_b200_prd(n, _dp(a), a.shape[0], _dp(d), _dp(e), n, m_forward)
    if rc != 0:
        raise RuntimeError(f"eigen_prd rc={rc}: {last_error()}")
    return d, e[0].copy(), e[1].copy()


def eigen_dcx(n, d, e1, e2, z, nvec=None):
    """Penta-diagonal eigen-decomposition (eigen_dcx, src/dcx.F:75).  Returns w; z filled in place."""
    _check_f(z, "z")
    w = np.zeros(n)
    e = np.ascontiguousarray(np.stack([e1, e2]))
    rc = lib().eigenexa_b200_dcx(n, n if nvec is None else nvec, _dp(np.ascontiguousarray(d)), _dp(e), n, _dp(w), _dp(z),
                                 z.shape[0])
    if rc != 0:
        raise RuntimeError(f"eigen_dcx rc={rc}: {last_error()}")
    return w


def eigen_bisect2(n, d, e1, e2):
    w = np.zeros(n)
    e = np.ascontiguousarray(np.stack([e1, e2]))
    rc = lib().eigenexa_b200_bisect2(n, _dp(np.ascontiguousarray(d)), _dp(e), n, _dp(w))
    if rc != 0:
        raise RuntimeError(f"eigen_bisect2 rc={rc}: {last_error()}")
    return w


def eigen_trbakwy(n, a, z, e, m_backward=128, nvec=None, nb=1):
    """eigen_common_trbakwy(n, nvec, a, lda, z, ldz, e, m, nb) (src/trbakwy4.F:77); z in place."""
    _check_f(a, "a")
    _check_f(z, "z")
    rc = lib().eigenexa_b200_trbakwy_nb(n, n if nvec is None else nvec, _dp(a), a.shape[0], _dp(z), z.shape[0],
                                        _dp(np.ascontiguousarray(e)), m_backward, nb)
    if rc != 0:
        raise RuntimeError(f"eigen_trbakwy rc={rc}: {last_error()}")
    return z


def eigen_dc(n, d, e, z, nvec=None):
    """Tridiagonal eigen-decomposition (eigen_dc2, src/dc2.F:78).  Returns w; z filled in place."""
    _check_f(z, "z")
    w = np.zeros(n)
    rc = lib().eigenexa_b200_dc(n, n if nvec is None else nvec, _dp(np.ascontiguousarray(d)),
                                _dp(np.ascontiguousarray(e)), _dp(w), _dp(z), z.shape[0])
    if rc != 0:
        raise RuntimeError(f"eigen_dc rc={rc}: {last_error()}")
    return w


def eigen_bisect(n, d, e):
    w = np.zeros(n)
    rc = lib().eigenexa_b200_bisect(n, _dp(np.ascontiguousarray(d)), _dp(np.ascontiguousarray(e)), _dp(w))
    if rc != 0:
        raise RuntimeError(f"eigen_bisect rc={rc}: {last_error()}")
    return w


def mat_set_host(n, a, mtype, seed=1):
    _check_f(a, "a")
    lib().eigenexa_b200_mat_set_host(n, a.ctypes.data, a.shape[0], mtype, seed)


def mat_set_dev(n, a_ptr, lda, mtype, seed=1):
    lib().eigenexa_b200_mat_set_dev(n, a_ptr, lda, mtype, seed)


def ev_test_dev(n, nvec, a_ptr, lda, w_ptr, z_ptr, ldz):
    out = np.zeros(4)
    lib().eigenexa_b200_ev_test_dev(n, nvec, a_ptr, lda, w_ptr, z_ptr, ldz, _dp(out))
    return float(out[0]), float(out[1])


def dgemm_dev(transa, transb, m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc):
    rc = lib().eigenexa_b200_dgemm_dev(transa.encode(), transb.encode(), m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta,
                                       c_ptr, ldc)
    if rc != 0:
        raise RuntimeError(last_error())


def dgemm_tri_dev(transa, transb, m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc, px=1, py=1, x=0, y=0):
    rc = lib().eigenexa_b200_dgemm_tri_dev(transa.encode(), transb.encode(), m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta,
                                           c_ptr, ldc, px, py, x, y)
    if rc != 0:
        raise RuntimeError(last_error())


def sync():
    lib().eigenexa_b200_sync()


def stream_ptr() -> int:
    return int(lib().eigenexa_b200_stream() or 0)


def launch_count(reset=False) -> int:
    return int(lib().eigenexa_b200_launch_count(int(reset)))


def last_timings():
    t = np.zeros(48)
    lib().eigenexa_b200_last_timings(_dp(t), 48)
    return t


def set_profiling(level: int):
    """0 off; 1 async CUDA events around every symv / syr2k launch; 2 sync per kernel class (debug)."""
    lib().eigenexa_b200_set_profiling(int(level))


def symv_trace():
    """Per-launch symv_kernel milliseconds of the last eigen_trd (column n first)."""
    cap = lib().eigenexa_b200_symv_trace(None, 0)
    out = np.zeros(max(cap, 1), dtype=np.float32)
    lib().eigenexa_b200_symv_trace(out.ctypes.data_as(C.POINTER(C.c_float)), cap)
    return out[:cap]


def set_debug_maxcols(ncols: int):
    lib().eigenexa_b200_set_debug_maxcols(int(ncols))


def last_error() -> str:
    return lib().eigenexa_b200_last_error().decode()
