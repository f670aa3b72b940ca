"""ctypes binding of libsdepth.so (include/statdepth_b200.h) -- the only way the host reaches the GPU.

There is NO CPU fallback: if the CUDA library is missing or no sm_100 device is visible, every
compute entry point raises :class:`EngineUnavailable`.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STATDEPTH_B200_LIB", os.path.join(_HERE, "libsdepth.so"))  # override: tuning builds

SD_OK = 0
STATUS_NAMES = {1: "SD_ERR_INVALID", 2: "SD_ERR_CUDA", 3: "SD_ERR_NO_DEVICE", 4: "SD_ERR_NONFINITE",
                5: "SD_ERR_OVERFLOW", 6: "SD_ERR_UNSUPPORTED"}
LAYOUT_TN, LAYOUT_NT = 0, 1
BD_AUTO, BD_BITS, BD_GEMM, BD_MATCH = 0, 1, 2, 3
OPT_BD_IMPL, OPT_MBD_FORCE_FALLBACK, OPT_PROFILE, OPT_SIMPLICIAL_IMPL, OPT_ASYNC_DEVICE = 1, 2, 3, 4, 5
SIMPLICIAL_AUTO, SIMPLICIAL_ENUMERATE, SIMPLICIAL_COUNT = 0, 1, 2
PHASES = ("mbd_splitters", "mbd_partition", "mbd_rank", "mbd_generic", "bd_masks", "bd_pairs", "mbd_slab_hist",
          "mbd_slab_rank")


class EngineUnavailable(RuntimeError):
    """The CUDA extension or the GPU is missing; the B200 engine has no CPU path."""


class EngineError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status


class Timings(C.Structure):
    _fields_ = [("h2d_ns", C.c_int64), ("kernel_ns", C.c_int64), ("d2h_ns", C.c_int64),
                ("launches", C.c_int64), ("fallback_rows", C.c_int64), ("bd_impl_used", C.c_int64)]


class DevInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("sm_count", C.c_int), ("cc_major", C.c_int), ("cc_minor", C.c_int),
                ("total_mem", C.c_int64), ("name", C.c_char * 128)]


_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); must list every symbol include/statdepth_b200.h declares
SIGNATURES = {
    "sd_abi_version": (C.c_int, []),
    "sd_last_error": (C.c_char_p, []),
    "sd_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "sd_destroy": (C.c_int, [C.c_void_p]),
    "sd_device_info": (C.c_int, [C.c_void_p, C.POINTER(DevInfo)]),
    "sd_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "sd_get_timings": (C.c_int, [C.c_void_p, C.POINTER(Timings)]),
    "sd_get_phase_ns": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sd_probe_int8_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "sd_stream": (C.c_void_p, [C.c_void_p]),
    "sd_sync": (C.c_int, [C.c_void_p]),
    "sd_mbd_plan": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p]),
    "sd_band_depth_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                                    C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "sd_band_depth_f64_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "sd_band_ranks_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                                    C.c_void_p]),
    "sd_simplex_depth_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                                       C.c_int64, C.c_int, C.c_double, C.c_void_p]),
    "sd_pointcloud_simplicial_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                               C.c_double, C.c_void_p]),
    "sd_pointcloud_l1_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                       C.c_void_p]),
    "sd_pointcloud_oja_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_int64, C.c_double, C.c_void_p]),
    "sd_pointcloud_blocks_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "sd_band_depth_batched_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                            C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library(path: str = None):
    """dlopen libsdepth.so and attach prototypes.  Works without a GPU (symbol check only)."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise EngineUnavailable(
                "%s not found: build it with `python -m statdepth_b200.build` (nvcc, sm_100a). "
                "The B200 engine has no CPU fallback." % p)
        lib = C.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and this table ever drift apart
            fn.restype = res
            fn.argtypes = args
        if path is None:
            _lib = lib
        return lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def mbd_plan(n: int, ld: int = None) -> dict:
    """Rank pipeline relaxed depth takes for rows of n curves (host arithmetic of the library, no GPU needed)."""
    out = np.zeros(6, dtype=np.int64)
    lib = load_library()
    if lib.sd_mbd_plan(int(n), int(n if ld is None else ld), _ptr(out)) != 0:
        raise ValueError(lib.sd_last_error().decode())
    keys = ("slab", "ctas_per_row", "bins_per_cta", "entries_per_cta", "smem_rank", "smem_hist")
    return dict(zip(keys, (int(v) for v in out)))


class Engine:
    """One context (stream + workspace) on one GPU.  Not thread-safe; use one per thread."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self._ctx = C.c_void_p()
        st = self.lib.sd_init(int(device), C.byref(self._ctx))
        if st != SD_OK:
            msg = self.lib.sd_last_error().decode()
            self._ctx = C.c_void_p()
            if st == 3:
                raise EngineUnavailable("no usable B200: %s (the engine has no CPU fallback)" % msg)
            raise EngineError(st, msg)
        self.device = int(device)

    # -- plumbing ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.sd_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != SD_OK:
            raise EngineError(st, self.lib.sd_last_error().decode())

    def info(self) -> dict:
        di = DevInfo()
        self._check(self.lib.sd_device_info(self._ctx, C.byref(di)))
        return dict(device=di.device, sm_count=di.sm_count, cc=(di.cc_major, di.cc_minor),
                    total_mem=di.total_mem, name=di.name.decode())

    def timings(self) -> dict:
        t = Timings()
        self._check(self.lib.sd_get_timings(self._ctx, C.byref(t)))
        return dict(h2d_ns=t.h2d_ns, kernel_ns=t.kernel_ns, d2h_ns=t.d2h_ns, launches=t.launches,
                    fallback_rows=t.fallback_rows, bd_impl_used=t.bd_impl_used)

    def phase_ns(self) -> dict:
        """Per-phase device time of the last call (needs set_option(OPT_PROFILE, 1))."""
        out = np.zeros(len(PHASES), dtype=np.int64)
        self._check(self.lib.sd_get_phase_ns(self._ctx, _ptr(out)))
        return {k: int(v) for k, v in zip(PHASES, out) if v}

    def probe_int8_peak(self) -> float:
        """Measured tcgen05 kind::i8 rate of this GPU, int8 ops/s."""
        v = C.c_double()
        self._check(self.lib.sd_probe_int8_peak(self._ctx, C.byref(v)))
        return float(v.value)

    def set_option(self, option: int, value: int):
        self._check(self.lib.sd_set_option(self._ctx, int(option), int(value)))

    def stream(self) -> int:
        return int(self.lib.sd_stream(self._ctx) or 0)

    def sync(self):
        """Wait for the engine's stream; completes a call made under OPT_ASYNC_DEVICE (raises its error)."""
        self._check(self.lib.sd_sync(self._ctx))

    @staticmethod
    def _queries(queries, n):
        if queries is None:
            return None, n
        q = np.ascontiguousarray(queries, dtype=np.int64)
        return q, int(q.size)

    @staticmethod
    def _matrix(X):
        """float64 view + layout of a 2-D array: C-order -> [rows, cols]; F-order -> transposed."""
        X = np.asarray(X)
        if X.dtype != np.float64:
            X = X.astype(np.float64)
        if X.ndim != 2:
            raise ValueError("expected a 2-D array")
        if X.flags.c_contiguous:
            return X, False
        if X.flags.f_contiguous:
            return X, True
        return np.ascontiguousarray(X), False

    # -- band depth -------------------------------------------------------------------------------
    def band_depth_counts(self, X, queries=None, j=2, relax=False):
        """Integer numerators for subset size j.  X is [T, n] (rows = time points, columns = curves);
        an F-ordered X (what a column-built DataFrame gives) is passed as-is with SD_LAYOUT_NT."""
        X, is_f = self._matrix(X)
        T, n = X.shape
        q, nq = self._queries(queries, n)
        out = np.empty(nq, dtype=np.int64)
        if is_f:  # memory is [n, T] row-major
            st = self.lib.sd_band_depth_f64(self._ctx, C.c_void_p(X.ctypes.data), T, n, T, LAYOUT_NT, _ptr(q), nq,
                                            int(j), int(bool(relax)), _ptr(out))
        else:
            st = self.lib.sd_band_depth_f64(self._ctx, C.c_void_p(X.ctypes.data), T, n, n, LAYOUT_TN, _ptr(q), nq,
                                            int(j), int(bool(relax)), _ptr(out))
        self._check(st)
        return out

    def band_depth_counts_ptr(self, host_ptr, T, n, ld, queries=None, j=2, relax=False, out=None):
        """Same, from a raw host pointer (e.g. pinned memory) in SD_LAYOUT_TN."""
        q, nq = self._queries(queries, n)
        if out is None:
            out = np.empty(nq, dtype=np.int64)
        self._check(self.lib.sd_band_depth_f64(self._ctx, C.c_void_p(int(host_ptr)), int(T), int(n), int(ld),
                                               LAYOUT_TN, _ptr(q), nq, int(j), int(bool(relax)), _ptr(out)))
        return out

    def band_depth_counts_dev(self, dX_ptr, T, n, ld, d_out_ptr, d_query_ptr=None, nq=None, j=2, relax=False):
        """Device-pointer variant: no copies; result stays on the device."""
        nq = int(n if nq is None else nq)
        self._check(self.lib.sd_band_depth_f64_dev(self._ctx, C.c_void_p(int(dX_ptr)), int(T), int(n), int(ld),
                                                   C.c_void_p(int(d_query_ptr)) if d_query_ptr else None, nq,
                                                   int(j), int(bool(relax)), C.c_void_p(int(d_out_ptr))))

    def band_ranks(self, X):
        """(below, above) int32 [T, n]: strict ranks of every curve at every time point."""
        X, is_f = self._matrix(X)
        T, n = X.shape
        below = np.empty((T, n), dtype=np.int32)
        above = np.empty((T, n), dtype=np.int32)
        st = self.lib.sd_band_ranks_f64(self._ctx, C.c_void_p(X.ctypes.data), T, n, T if is_f else n,
                                        LAYOUT_NT if is_f else LAYOUT_TN, _ptr(below), _ptr(above))
        self._check(st)
        return below, above

    def band_depth_counts_batched(self, X, membership, queries, j=2, relax=False):
        """membership [B, n] (bool/uint8), queries [B, nqb] global curve ids -> counts [B, nqb]."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        T, n = X.shape
        mem = np.ascontiguousarray(membership, dtype=np.uint8)
        qs = np.ascontiguousarray(queries, dtype=np.int64)
        B, nqb = qs.shape
        assert mem.shape == (B, n)
        out = np.empty((B, nqb), dtype=np.int64)
        self._check(self.lib.sd_band_depth_batched_f64(self._ctx, _ptr(X), T, n, n, _ptr(mem), B, _ptr(qs), nqb,
                                                       int(j), int(bool(relax)), _ptr(out)))
        return out

    # -- multivariate / point clouds -----------------------------------------------------------------
    def simplex_depth_counts(self, F, queries=None, relax=False, tol=1e-7):
        F = np.ascontiguousarray(F, dtype=np.float64)
        N, T, d = F.shape
        q, nq = self._queries(queries, N)
        out = np.empty(nq, dtype=np.int64)
        self._check(self.lib.sd_simplex_depth_f64(self._ctx, _ptr(F), N, T, d, _ptr(q), nq, int(bool(relax)),
                                                  float(tol), _ptr(out)))
        return out

    def simplicial_counts(self, P, queries=None, tol=1e-7):
        P = np.ascontiguousarray(P, dtype=np.float64)
        n, d = P.shape
        q, nq = self._queries(queries, n)
        out = np.empty(nq, dtype=np.int64)
        self._check(self.lib.sd_pointcloud_simplicial_f64(self._ctx, _ptr(P), n, d, _ptr(q), nq, float(tol),
                                                          _ptr(out)))
        return out

    def l1_depth(self, P, queries=None):
        P = np.ascontiguousarray(P, dtype=np.float64)
        n, d = P.shape
        q, nq = self._queries(queries, n)
        out = np.empty(nq, dtype=np.float64)
        self._check(self.lib.sd_pointcloud_l1_f64(self._ctx, _ptr(P), n, d, _ptr(q), nq, _ptr(out)))
        return out

    def oja(self, P, hull_volume, queries=None, pool=None):
        P = np.ascontiguousarray(P, dtype=np.float64)
        n, d = P.shape
        q, nq = self._queries(queries, n)
        pl = None if pool is None else np.ascontiguousarray(pool, dtype=np.int64)
        npool = n if pl is None else int(pl.size)
        out = np.empty(nq, dtype=np.float64)
        self._check(self.lib.sd_pointcloud_oja_f64(self._ctx, _ptr(P), n, d, _ptr(q), nq, _ptr(pl), npool,
                                                   float(hull_volume), _ptr(out)))
        return out


    BLOCK_KINDS = {"simplex": 0, "l1": 1, "oja": 2}

    def cloud_blocks(self, P, members, offsets, query_pos, kind, tol=1e-7, hull_volumes=None):
        """One single-query depth per block of member points (K-sampled point-cloud depth), one launch for all.
        members: concatenated point ids; offsets [B+1]; query_pos [B] position of the query inside its block.
        Returns float64 [B]: simplicial COUNT (exact), L1 depth, or Oja sum / hull_volumes[b]."""
        P = np.ascontiguousarray(P, dtype=np.float64)
        n, d = P.shape
        mem = np.ascontiguousarray(members, dtype=np.int64)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        qp = np.ascontiguousarray(query_pos, dtype=np.int64)
        B = int(qp.size)
        hv = None if hull_volumes is None else np.ascontiguousarray(hull_volumes, dtype=np.float64)
        out = np.empty(B, dtype=np.float64)
        self._check(self.lib.sd_pointcloud_blocks_f64(self._ctx, _ptr(P), n, d, _ptr(mem), _ptr(off), _ptr(qp), B,
                                                      int(self.BLOCK_KINDS[kind]), float(tol), _ptr(hv), _ptr(out)))
        return out


_engines = {}


def get_engine(device: int = None) -> Engine:
    """Process-wide engine for `device` (default: $LOCAL_RANK or $STATDEPTH_DEVICE or 0)."""
    if device is None:
        device = int(os.environ.get("STATDEPTH_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        if "STATDEPTH_DEVICE" not in os.environ:
            # LOCAL_RANK is out of range when CUDA_VISIBLE_DEVICES is narrowed to one device per rank
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis is not None and device >= len([v for v in vis.split(",") if v.strip()]):
                device = 0
    eng = _engines.get(device)
    if eng is None:
        eng = Engine(device)
        _engines[device] = eng
    return eng
