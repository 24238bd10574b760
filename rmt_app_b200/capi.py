"""ctypes binding of librmtb200.so (include/rmt_b200.h).

This is the stub a maintainer of the reference would add to call the B200
engine from Python.  It contains no numerical code and no
fallback: when the shared library is missing it is built in-tree, and when no
CUDA driver/device is present every compute call raises `RmtError`.
"""
import ctypes as C
import hashlib
import os

import numpy as np

from . import build as _build

_PKG = os.path.dirname(os.path.abspath(__file__))
KERNELS_PATH = os.path.join(_PKG, "csrc", "rmt_kernels.cu")
CACHE_DIR = os.environ.get("RMT_B200_CACHE", os.path.join(_PKG, "_cubin_cache"))


class RmtError(RuntimeError):
    pass


class ModuleInfo(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "model", "n", "nc", "nr", "nin", "nconst", "nkp", "stages", "block", "iso",
        "flops_rhs_alg", "flops_rhs_wt", "flops_jac_alg", "flops_jac_wt", "m", "lanes")]


_lib = None

# every symbol include/rmt_b200.h declares: (restype, argtypes)
_vp, _i32, _i64, _u64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
_pi32, _pdbl = C.POINTER(C.c_int32), C.POINTER(C.c_double)
SIGNATURES = {
    "rmt_last_error": (C.c_char_p, []),
    "rmt_version": (C.c_char_p, []),
    "rmt_init": (C.c_int, [C.c_int]),
    "rmt_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rmt_shutdown": (C.c_int, []),
    "rmt_nvrtc_compile": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_int,
                                    C.POINTER(_u64)]),
    "rmt_blob_data": (C.c_int, [_u64, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "rmt_blob_log": (C.c_char_p, [_u64]),
    "rmt_blob_ptx": (C.c_int, [_u64, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]),
    "rmt_blob_free": (C.c_int, [_u64]),
    "rmt_module_load": (C.c_int, [_vp, C.c_size_t, C.POINTER(_u64)]),
    "rmt_module_get_info": (C.c_int, [_u64, C.POINTER(ModuleInfo)]),
    "rmt_module_free": (C.c_int, [_u64]),
    "rmt_setup": (C.c_int, [_u64, _i64, _vp, _i32, _pi32, _pdbl, _vp, _vp]),
    "rmt_n1_rhs": (C.c_int, [_u64, _i64, _vp, _vp, _vp, _vp]),
    "rmt_n1_jac": (C.c_int, [_u64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "rmt_n1_sys": (C.c_int, [_u64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "rmt_n1_solve": (C.c_int, [_u64, _i64, _vp, _i32, _pdbl, _dbl, _dbl, _i32, _i32, _i32, _vp, _vp, _vp, _pdbl,
                               _vp, _pdbl, _vp]),
    "rmt_n1_solve_population": (C.c_int, [_u64, _i64, _vp, _dbl, _dbl, _dbl, _i32, _vp, _vp, _vp, _pdbl, _vp, _vp, _i64, _pdbl,
                                          _vp]),
    "rmt_n1_solve_host": (C.c_int, [_u64, _i64, _vp, _i32, _pi32, _pdbl, _i32, _pdbl, _dbl, _dbl, _i32, _i32, _i32,
                                    _vp, _vp, _vp, _pdbl, _vp, _pdbl]),
    "rmt_n2_rhs": (C.c_int, [_u64, _i64, _i32, _vp, _vp, _vp, _vp]),
    "rmt_n2_work_doubles": (_i64, [_u64, _i64, _i32]),
    "rmt_n2_solve": (C.c_int, [_u64, _i64, _i32, _i32, _dbl, _vp, _dbl, _dbl, _i32, _i32, _vp, _vp, _vp, _vp, _pdbl,
                               _vp]),
    "rmt_reduce_objective": (C.c_int, [_u64, _i64, _vp, _i64, _pdbl, _pdbl, C.POINTER(_i64), _vp]),
    "rmt_comm_unique_id": (C.c_int, [_vp, C.c_size_t]),
    "rmt_comm_init": (C.c_int, [_i32, _i32, _vp, C.c_size_t, C.POINTER(_u64)]),
    "rmt_comm_info": (C.c_int, [_u64, _pi32, _pi32, _pi32]),
    "rmt_comm_allgather": (C.c_int, [_u64, _vp, _vp, _i64, _vp]),
    "rmt_comm_allreduce": (C.c_int, [_u64, _vp, _vp, _i64, _i32, _vp]),
    "rmt_comm_free": (C.c_int, [_u64]),
    "rmt_math_probe": (C.c_int, [_u64, _i32, _vp, _vp, _vp]),
    "rmt_debug_trace": (C.c_int, [_vp, _i32, _i64]),
    "rmt_fp64_peak": (C.c_int, [_u64, _i32, _i32, _pdbl]),
}


def lib():
    """Load (building if necessary) librmtb200.so.  Never falls back."""
    global _lib
    if _lib is None:
        path = _build.build_library()
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)       # AttributeError if the library lacks a declared symbol
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RmtError(lib().rmt_last_error().decode("utf-8", "replace"))


_initialised = None


def init(device=0):
    global _initialised
    if _initialised != device:
        _check(lib().rmt_init(int(device)))
        _initialised = device


def device_count():
    n = C.c_int(0)
    _check(lib().rmt_device_count(C.byref(n)))
    return n.value


def debug_trace(d_trace, cap=0, instance=-1):
    _check(lib().rmt_debug_trace(_ptr(d_trace), cap, instance))


def kernels_source():
    with open(KERNELS_PATH) as f:
        return f.read()


def nvrtc_compile(model_src, block=128, arch="sm_100a", extra_opts=(), want_ptx=False):
    """Generated header + hand-written kernels -> cubin bytes (no GPU needed)."""
    L = lib()
    blob = _u64(0)
    opts = (C.c_char_p*max(len(extra_opts), 1))(*[o.encode() for o in extra_opts])
    _check(L.rmt_nvrtc_compile(model_src.encode(), kernels_source().encode(), arch.encode(), int(block), opts,
                               len(extra_opts), C.byref(blob)))
    try:
        data, size = _vp(), C.c_size_t()
        _check(L.rmt_blob_data(blob, C.byref(data), C.byref(size)))
        cubin = C.string_at(data, size.value)
        log = L.rmt_blob_log(blob).decode("utf-8", "replace")
        ptx = None
        if want_ptx:
            p, n = C.c_char_p(), C.c_size_t()
            _check(L.rmt_blob_ptx(blob, C.byref(p), C.byref(n)))
            ptx = p.value.decode() if p.value else ""
    finally:
        L.rmt_blob_free(blob)
    return (cubin, log, ptx) if want_ptx else (cubin, log)


def cubin_key(model_src, block=128, arch="sm_100a", extra_opts=()):
    """Identity of a compiled module: hash of the exact translation unit (generated header + hand-written
    kernels) and its options.  Names the cache file; profiles/*calibration.json records it so that numbers
    read off an ncu capture can be recognised as stale when the kernel source has changed since."""
    h = hashlib.sha256()
    for part in (model_src, kernels_source(), arch, str(block), lib().rmt_version().decode()) + tuple(extra_opts):
        h.update(part.encode())
        h.update(b"\0")
    return h.hexdigest()[:24]


def cached_cubin(model_src, block=128, arch="sm_100a", extra_opts=()):
    """Disk cache keyed by the exact translation unit + options."""
    path = os.path.join(CACHE_DIR, cubin_key(model_src, block, arch, extra_opts) + ".cubin")
    if os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    cubin, _ = nvrtc_compile(model_src, block=block, arch=arch, extra_opts=extra_opts)
    try:
        os.makedirs(CACHE_DIR, exist_ok=True)
        tmp = path + ".%d.tmp" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(cubin)
        os.replace(tmp, path)
    except OSError:
        pass
    return cubin


def _ptr(t):
    """Device (or host) address of a torch tensor / numpy array / int / None."""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def _dptr(a):
    return a.ctypes.data_as(_pdbl)


class Module:
    """A compiled model loaded on the current device."""

    def __init__(self, cubin):
        self.handle = _u64(0)
        buf = C.create_string_buffer(cubin, len(cubin))
        _check(lib().rmt_module_load(C.cast(buf, _vp), len(cubin), C.byref(self.handle)))
        self.info = ModuleInfo()
        _check(lib().rmt_module_get_info(self.handle, C.byref(self.info)))

    def close(self):
        if self.handle is not None and self.handle.value:
            lib().rmt_module_free(self.handle)
            self.handle = None

    # thin 1:1 wrappers -----------------------------------------------------------
    def setup(self, B, d_rows, n_rows, row_map, uniform, d_consts, stream=None):
        row_map = np.ascontiguousarray(row_map, dtype=np.int32)
        uniform = np.ascontiguousarray(uniform, dtype=np.float64)
        assert row_map.size == self.info.nin and uniform.size == self.info.nin
        _check(lib().rmt_setup(self.handle, B, _ptr(d_rows), n_rows, row_map.ctypes.data_as(_pi32), _dptr(uniform),
                               _ptr(d_consts), stream))

    def n1_rhs(self, B, d_consts, d_y, d_f, stream=None):
        _check(lib().rmt_n1_rhs(self.handle, B, _ptr(d_consts), _ptr(d_y), _ptr(d_f), stream))

    def n1_jac(self, B, d_consts, d_y, d_f, d_J, stream=None):
        _check(lib().rmt_n1_jac(self.handle, B, _ptr(d_consts), _ptr(d_y), _ptr(d_f), _ptr(d_J), stream))

    def n1_sys(self, B, d_consts, d_y, d_g, d_A, stream=None):
        _check(lib().rmt_n1_sys(self.handle, B, _ptr(d_consts), _ptr(d_y), _ptr(d_g), _ptr(d_A), stream))

    def n1_solve(self, B, d_consts, z_eval, rtol, atol, d_out, d_status, d_stats, max_steps=100000, dense=True,
                 out_mode=1, obj_ref=None, d_obj=None, ctrl=None, stream=None):
        z = np.ascontiguousarray(z_eval, dtype=np.float64)
        ref = None if obj_ref is None else np.ascontiguousarray(obj_ref, dtype=np.float64)
        ctl = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float64)
        assert ctl is None or ctl.size == 6
        _check(lib().rmt_n1_solve(self.handle, B, _ptr(d_consts), z.size, _dptr(z), rtol, atol, max_steps,
                                  1 if dense else 0, out_mode, _ptr(d_out), _ptr(d_status), _ptr(d_stats),
                                  None if ref is None else _dptr(ref), _ptr(d_obj),
                                  None if ctl is None else _dptr(ctl), stream))

    def n1_solve_population(self, B, d_consts, z_end, rtol, atol, d_out, d_status, d_stats, obj_ref, d_obj, d_red,
                            index_offset=0, max_steps=100000, ctrl=None, stream=None):
        ref = np.ascontiguousarray(obj_ref, dtype=np.float64)
        ctl = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float64)
        _check(lib().rmt_n1_solve_population(self.handle, B, _ptr(d_consts), z_end, rtol, atol, max_steps, _ptr(d_out),
                                             _ptr(d_status), _ptr(d_stats), _dptr(ref), _ptr(d_obj), _ptr(d_red),
                                             index_offset, None if ctl is None else _dptr(ctl), stream))

    def n1_solve_host(self, B, h_rows, n_rows, row_map, uniform, z_eval, rtol, atol, h_out, h_status, h_stats=None,
                      max_steps=100000, dense=True, out_mode=1, obj_ref=None, h_obj=None, ctrl=None):
        row_map = np.ascontiguousarray(row_map, dtype=np.int32)
        uniform = np.ascontiguousarray(uniform, dtype=np.float64)
        z = np.ascontiguousarray(z_eval, dtype=np.float64)
        ref = None if obj_ref is None else np.ascontiguousarray(obj_ref, dtype=np.float64)
        ctl = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float64)
        _check(lib().rmt_n1_solve_host(self.handle, B, _ptr(h_rows), n_rows, row_map.ctypes.data_as(_pi32),
                                       _dptr(uniform), z.size, _dptr(z), rtol, atol, max_steps, 1 if dense else 0,
                                       out_mode, _ptr(h_out), _ptr(h_status), _ptr(h_stats),
                                       None if ref is None else _dptr(ref), _ptr(h_obj),
                                       None if ctl is None else _dptr(ctl)))

    def n2_rhs(self, B, zNo, d_consts, d_y, d_f, stream=None):
        _check(lib().rmt_n2_rhs(self.handle, B, zNo, _ptr(d_consts), _ptr(d_y), _ptr(d_f), stream))

    def n2_work_doubles(self, B, zNo):
        n = lib().rmt_n2_work_doubles(self.handle, B, zNo)
        if n < 0:
            _check(1)
        return n

    def n2_solve(self, B, zNo, tNo, period, d_consts, rtol, atol, d_out, d_status, d_stats, d_work,
                 max_steps=1000000, out_mode=1, ctrl=None, stream=None):
        ctl = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float64)
        _check(lib().rmt_n2_solve(self.handle, B, zNo, tNo, period, _ptr(d_consts), rtol, atol, max_steps, out_mode,
                                  _ptr(d_out), _ptr(d_status), _ptr(d_stats), _ptr(d_work),
                                  None if ctl is None else _dptr(ctl), stream))

    def reduce_objective(self, n, d_obj, index_offset=0, stream=None):
        s, mn, am = _dbl(0), _dbl(0), _i64(0)
        _check(lib().rmt_reduce_objective(self.handle, n, _ptr(d_obj), index_offset, C.byref(s), C.byref(mn),
                                          C.byref(am), stream))
        return s.value, mn.value, am.value

    def math_probe(self, n, d_x, d_out, stream=None):
        _check(lib().rmt_math_probe(self.handle, n, _ptr(d_x), _ptr(d_out), stream))

    def fp64_peak(self, iters=8192, repeats=5):
        v = _dbl(0)
        _check(lib().rmt_fp64_peak(self.handle, iters, repeats, C.byref(v)))
        return v.value


COMM_ID_BYTES = 128


def comm_unique_id():
    """Rank 0: a fresh communicator id (bytes) to ship to the other ranks."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(lib().rmt_comm_unique_id(C.cast(buf, _vp), COMM_ID_BYTES))
    return buf.raw


class Comm:
    """NCCL communicator behind the C ABI (rmt_comm_*): gather / reduce of sharded ensembles without
    torch.distributed.  `rmt_init(device)` must have been called in this process (capi.init)."""

    def __init__(self, nranks, rank, unique_id):
        self.handle = _u64(0)
        self.nranks, self.rank = int(nranks), int(rank)
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        _check(lib().rmt_comm_init(self.nranks, self.rank, C.cast(buf, _vp), COMM_ID_BYTES, C.byref(self.handle)))

    def nccl_version(self):
        v = C.c_int32(0)
        _check(lib().rmt_comm_info(self.handle, None, None, C.byref(v)))
        return v.value

    def allgather(self, d_send, d_recv, count, stream=None):
        _check(lib().rmt_comm_allgather(self.handle, _ptr(d_send), _ptr(d_recv), int(count), stream))

    def allreduce(self, d_send, d_recv, count, op="sum", stream=None):
        _check(lib().rmt_comm_allreduce(self.handle, _ptr(d_send), _ptr(d_recv), int(count),
                                        {"sum": 0, "min": 1, "max": 2}[op], stream))

    def close(self):
        if self.handle is not None and self.handle.value:
            lib().rmt_comm_free(self.handle)
            self.handle = None
