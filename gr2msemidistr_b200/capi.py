"""ctypes binding of libgr2m_b200.so (include/gr2m_b200.h).

This is the only way the Python host code reaches the GPU; there is no fallback.  Loading fails
loudly if the shared library has not been built (``python -c "import __graft_entry__ as g; g.build()"``).
NumPy arrays cross as C-contiguous ``[ncell, ntime]`` float64 -- byte-identical to R's column-major
``[ntime x ncell]`` matrices that the C ABI specifies.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GR2M_B200_LIB") or os.path.join(_HERE, "libgr2m_b200.so")   # override: experiment builds only

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_NODEVICE = 0, -1, -2, -3, -4
PERC_F32_EXPONENT, SINK_UNROUNDED, MATH_LITERAL = 1, 8, 16
OF_KGE, OF_NSE, OF_RMSE = 0, 1, 2
WFA_F32, WFA_NO_CONTAMINATION, WFA_LEVEL_SYNC, WFA_DATAFLOW, WFA_NO_FORWARD, WFA_BY_BASIN, WFA_ROUTE_DENSE, WFA_NO_OVERLAP = 1, 2, 4, 8, 16, 32, 64, 128
WFA_NO_SEGMENTS = 256
OF_NAMES = {"KGE": OF_KGE, "NSE": OF_NSE, "RMSE": OF_RMSE}

# every symbol include/gr2m_b200.h declares (tests check the library exports exactly these)
SYMBOLS = (
    "gr2m_last_error", "gr2m_version", "gr2m_device_count", "gr2m_release_cached_memory", "gr2m_set_device", "gr2m_kernel_launches", "gr2m_guard_checked", "gr2m_guard_violations", "gr2m_fp64_peak",
    "gr2m_basin_create", "gr2m_basin_create_dev", "gr2m_basin_destroy", "gr2m_basin_set_stream", "gr2m_run", "gr2m_run_once",
    "gr2m_objective_batch", "gr2m_objective_batch_dev", "gr2m_objective_routed_batch", "gr2m_basin_last_kernel_ms",
    "wfa_plan_create", "wfa_plan_create_dev", "wfa_plan_destroy", "wfa_plan_set_stream", "wfa_plan_info",
    "wfa_plan_export", "wfa_plan_basin_info", "wfa_plan_reach_info", "wfa_plan_segment_info", "wfa_router_create", "wfa_router_destroy",
    "wfa_router_info", "wfa_router_gauge_upstream", "wfa_router_apply", "wfa_router_apply_dev", "wfa_route", "wfa_accumulate", "wfa_accumulate_dev", "wfa_mask_dev",
    "wfa_plan_last_kernel_ms", "wfa_plan_last_phase_ms", "wfa_flowdir_from_dem", "wfa_flowdir_from_dem_dev",
)


class GR2MError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgr2m_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libgr2m_b200.so once.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build the CUDA library first (__graft_entry__.build()); "
                "gr2msemidistr_b200 has no CPU implementation")
        L = C.CDLL(LIB_PATH)
        L.gr2m_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            getattr(L, name)          # AttributeError if the ABI and this binding drift apart
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise GR2MError(rc, lib().gr2m_last_error().decode(errors="replace"))


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def device_count() -> int:
    return lib().gr2m_device_count()


def guard_counts():
    """(buffers checked, buffers with a damaged guard band) in GR2M_GUARD=1 mode."""
    a, b = lib().gr2m_guard_checked, lib().gr2m_guard_violations
    a.restype = b.restype = C.c_longlong
    return int(a()), int(b())


def kernel_launches() -> int:
    """Kernels the library has launched since it was loaded (every launch site counts itself)."""
    f = lib().gr2m_kernel_launches
    f.restype = C.c_longlong
    return int(f())


def release_cached_memory():
    """Give the device blocks kept for the next handle back to the driver."""
    _check(lib().gr2m_release_cached_memory())


def set_device(device: int):
    _check(lib().gr2m_set_device(C.c_int(device)))


def fp64_peak_tflops(iters: int = 4096) -> float:
    """Measured DFMA throughput of the current device (TFLOP/s, 2 flop per DFMA)."""
    t = C.c_double()
    _check(lib().gr2m_fp64_peak(C.c_int(iters), C.byref(t), None))
    return t.value


class Basin:
    """gr2m_basin handle: the run-period forcing table and subbasin attributes resident in HBM."""

    def __init__(self, P, E, area, ndays, region_id, nreg: int, dev_ptrs=None, shape=None, stream: int = 0):
        """P, E: [ncell, ntime] host arrays -- or, with dev_ptrs=(P_ptr, E_ptr) and shape=(ncell, ntime), the
        same tables already in device memory, written on cudaStream_t `stream` (0 = complete)."""
        if dev_ptrs is None:
            P, E = _f64(P), _f64(E)
            if P.ndim != 2 or P.shape != E.shape:
                raise ValueError("P and E must be [ncell, ntime] arrays of the same shape")
            self.ncell, self.ntime = P.shape
        else:
            self.ncell, self.ntime = (int(x) for x in shape)
        self.nreg = int(nreg)
        area, ndays, region_id = _f64(area), _i32(ndays), _i32(region_id)
        if area.shape != (self.ncell,) or region_id.shape != (self.ncell,) or ndays.shape != (self.ntime,):
            raise ValueError("area/region_id must be [ncell], ndays [ntime]")
        self._h = C.c_void_p()
        if dev_ptrs is None:
            _check(lib().gr2m_basin_create(C.c_int(self.ntime), C.c_int(self.ncell), _dp(P), _dp(E), _dp(area),
                                           _ip(ndays), _ip(region_id), C.c_int(self.nreg), C.byref(self._h)))
        else:
            _check(lib().gr2m_basin_create_dev(C.c_int(self.ntime), C.c_int(self.ncell), C.c_void_p(dev_ptrs[0]),
                                               C.c_void_p(dev_ptrs[1]), _dp(area), _ip(ndays), _ip(region_id),
                                               C.c_int(self.nreg), C.c_void_p(stream), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h and _lib is not None:
            _lib.gr2m_basin_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int | None):
        _check(lib().gr2m_basin_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _check(lib().gr2m_basin_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def run(self, params, S0=None, R0=None, flags: int = 0, outputs=("PR", "AE", "SM", "RU", "QS")):
        """One simulation.  Returns dict of UNROUNDED [ncell, ntime] arrays + S_end, R_end [ncell]."""
        params = _f64(params)
        if params.shape != (4 * self.nreg,):
            raise ValueError(f"params must have 4*nreg = {4 * self.nreg} entries")
        S0, R0 = _f64(S0), _f64(R0)
        out = {k: np.empty((self.ncell, self.ntime)) for k in outputs}
        out["S_end"], out["R_end"] = np.empty(self.ncell), np.empty(self.ncell)
        _check(lib().gr2m_run(self._h, _dp(params), _dp(S0), _dp(R0), C.c_int(flags),
                              *[_dp(out.get(k)) for k in ("PR", "AE", "SM", "RU", "QS")],
                              _dp(out["S_end"]), _dp(out["R_end"])))
        return out

    def objective_batch(self, params, qobs, warmup: int = 0, which: int = OF_NSE, flags: int = 0,
                        want_crit: bool = True, want_qt: bool = False):
        """OFUN for many parameter sets.  params [nsets, 4*nreg].  Returns dict(of, crit, qt)."""
        params = _f64(params)
        if params.ndim != 2 or params.shape[1] != 4 * self.nreg:
            raise ValueError(f"params must be [nsets, {4 * self.nreg}]")
        nsets = params.shape[0]
        qobs = _f64(qobs)
        of = np.full(nsets, np.nan)
        crit = np.full((nsets, 3), np.nan) if want_crit else None
        qt = np.empty((nsets, self.ntime)) if want_qt else None
        _check(lib().gr2m_objective_batch(self._h, _dp(params), C.c_int(nsets), _dp(qobs), C.c_int(warmup or 0),
                                          C.c_int(which), C.c_int(flags), _dp(of), _dp(crit), _dp(qt)))
        return dict(of=of, crit=crit, qt=qt)

    def objective_routed_batch(self, router, params, qobs, gauge: int, warmup: int = 0, which: int = OF_NSE, flags: int = 0,
                               want_crit: bool = True, want_qt: bool = False):
        """Run -> Routing -> hydroGOF for many parameter sets: the objective on the ROUTED discharge of polygon `gauge`
        (0-based index into the router's subbasins).  Returns dict(of, crit, qt) like objective_batch."""
        params = _f64(params)
        if params.ndim != 2 or params.shape[1] != 4 * self.nreg:
            raise ValueError(f"params must be [nsets, {4 * self.nreg}]")
        nsets = params.shape[0]
        qobs = _f64(qobs)
        of = np.full(nsets, np.nan)
        crit = np.full((nsets, 3), np.nan) if want_crit else None
        qt = np.empty((nsets, self.ntime)) if want_qt else None
        _check(lib().gr2m_objective_routed_batch(self._h, router._h, _dp(params), C.c_int(nsets), _dp(qobs), C.c_int(gauge),
                                                 C.c_int(warmup or 0), C.c_int(which), C.c_int(flags), _dp(of), _dp(crit), _dp(qt)))
        return dict(of=of, crit=crit, qt=qt)

    def objective_batch_dev(self, params_ptr: int, nsets: int, qobs, warmup: int, which: int, flags: int,
                            of_ptr: int, crit_ptr: int = 0, qt_ptr: int = 0):
        """Device-pointer variant (asynchronous on the handle's stream)."""
        qobs = _f64(qobs)
        _check(lib().gr2m_objective_batch_dev(self._h, C.c_void_p(params_ptr), C.c_int(nsets), _dp(qobs),
                                              C.c_int(warmup or 0), C.c_int(which), C.c_int(flags),
                                              C.c_void_p(of_ptr), C.c_void_p(crit_ptr or 0), C.c_void_p(qt_ptr or 0)))


class Router:
    """wfa_router handle: the routing operator of one (D8 plan, source cells, polygons) geometry.

    apply(QS) gives what Plan.route(QS, ...) gives, for any number of columns (months, or months x sets)."""

    def __init__(self, plan: "Plan", src_cell, poly_off, poly_cells, flags: int = 0):
        self._h = C.c_void_p()
        self._plan = plan                      # the router borrows the plan's topology and stream
        src, poff, pc = _i32(src_cell), _i32(poly_off), _i32(poly_cells)
        self.nsub = int(src.shape[0])
        if poff.shape != (self.nsub + 1,):
            raise ValueError("src_cell must be [nsub], poly_off [nsub+1]")
        if pc.size == 0:
            pc = np.zeros(1, np.int32)
        _check(lib().wfa_router_create(plan._h, C.c_int(self.nsub), _ip(src), _ip(poff), _ip(pc), C.c_int(flags),
                                       C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h and _lib is not None:
            _lib.wfa_router_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        _check(lib().wfa_router_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(nnodes=a.value, nlevels=b.value, ncandidates=c.value, nactive=d.value)

    def gauge_upstream(self, gauge: int):
        """(nmax, owner[nsub]): the maximal candidate nodes of polygon `gauge` and, per subbasin, which of them it
        drains through (-1: it does not reach the polygon)."""
        n = C.c_int()
        owner = np.empty(self.nsub, np.int32)
        _check(lib().wfa_router_gauge_upstream(self._h, C.c_int(gauge), C.byref(n), _ip(owner)))
        return n.value, owner

    def apply(self, QS):
        """QS [nsub, ncols] -> QR [nsub, ncols] (unrounded per-polygon maxima)."""
        QS = _f64(QS)
        if QS.ndim != 2 or QS.shape[0] != self.nsub:
            raise ValueError("QS must be [nsub, ncols]")
        QR = np.empty_like(QS)
        _check(lib().wfa_router_apply(self._h, C.c_int(QS.shape[1]), _dp(QS), _dp(QR)))
        return QR

    def apply_dev(self, qs_ptr: int, qr_ptr: int, ncols: int):
        """Device-resident, asynchronous on the plan's stream: QS, QR are [nsub][ncols] FP64 in HBM."""
        _check(lib().wfa_router_apply_dev(self._h, C.c_int(ncols), C.c_void_p(qs_ptr), C.c_void_p(qr_ptr)))


def run_once(P, E, area, ndays, region_id, nreg, params, S0=None, R0=None, flags=0):
    """gr2m_run_once: create + run + destroy through the single C entry point."""
    P, E = _f64(P), _f64(E)
    ncell, ntime = P.shape
    area, ndays, region_id, params = _f64(area), _i32(ndays), _i32(region_id), _f64(params)
    S0, R0 = _f64(S0), _f64(R0)
    out = {k: np.empty((ncell, ntime)) for k in ("PR", "AE", "SM", "RU", "QS")}
    out["S_end"], out["R_end"] = np.empty(ncell), np.empty(ncell)
    _check(lib().gr2m_run_once(C.c_int(ntime), C.c_int(ncell), _dp(P), _dp(E), _dp(area), _ip(ndays),
                               _ip(region_id), C.c_int(nreg), _dp(params), _dp(S0), _dp(R0), C.c_int(flags),
                               *[_dp(out[k]) for k in ("PR", "AE", "SM", "RU", "QS", "S_end", "R_end")]))
    return out


class Plan:
    """wfa_plan handle: D8 topology (receiver, in-degree, Kahn levels, order, contamination) in HBM."""

    def __init__(self, flowdir=None, flags: int = 0, dev_ptr: int | None = None, shape=None):
        self._h = C.c_void_p()
        if dev_ptr is not None:
            self.nrow, self.ncol = shape
            _check(lib().wfa_plan_create_dev(C.c_void_p(dev_ptr), C.c_int(self.nrow), C.c_int(self.ncol),
                                             C.c_int(flags), C.byref(self._h)))
        else:
            d = np.ascontiguousarray(flowdir, dtype=np.int16)
            if d.ndim != 2:
                raise ValueError("flowdir must be a 2-D [nrow, ncol] array")
            self.nrow, self.ncol = d.shape
            _check(lib().wfa_plan_create(d.ctypes.data_as(C.POINTER(C.c_int16)), C.c_int(self.nrow),
                                         C.c_int(self.ncol), C.c_int(flags), C.byref(self._h)))
        nl, no, nc = C.c_int(), C.c_int(), C.c_int64()
        _check(lib().wfa_plan_info(self._h, C.byref(nl), C.byref(no), C.byref(nc)))
        self.nlevels, self.norder, self.ncell = nl.value, no.value, nc.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h and _lib is not None:
            _lib.wfa_plan_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int | None):
        _check(lib().wfa_plan_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _check(lib().wfa_plan_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def basin_info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().wfa_plan_basin_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(cut_level=a.value, nbasins=b.value, nupper=c.value, enabled=bool(d.value))

    def reach_info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().wfa_plan_reach_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(nreaches=a.value, nreach_levels=b.value, critical_cells=c.value, enabled=bool(d.value))

    def segment_info(self):
        a, b, c, d = C.c_longlong(), C.c_longlong(), C.c_int(), C.c_int()
        _check(lib().wfa_plan_segment_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(nsegments=a.value, ncells=b.value, levels=c.value, enabled=bool(d.value))

    def last_phase_ms(self):
        a, b = C.c_float(), C.c_float()
        _check(lib().wfa_plan_last_phase_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def export(self):
        n = self.ncell
        out = dict(receiver=np.empty(n, np.int32), inmask=np.empty(n, np.uint8), indeg=np.empty(n, np.int32),
                   level=np.empty(n, np.int32), order=np.empty(self.norder, np.int32),
                   level_off=np.empty(self.nlevels + 1, np.int32), contam=np.empty(n, np.uint8))
        u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
        _check(lib().wfa_plan_export(self._h, _ip(out["receiver"]), u8(out["inmask"]), _ip(out["indeg"]),
                                     _ip(out["level"]), _ip(out["order"]), _ip(out["level_off"]), u8(out["contam"])))
        out["nlevels"] = self.nlevels
        return out

    def route(self, QS, src_cell, poly_off, poly_cells, flags: int = 0):
        """Routing month loop for all months.  QS [nsub, ntime] -> QR [nsub, ntime] (unrounded)."""
        QS = _f64(QS)
        nsub, ntime = QS.shape
        src, poff, pc = _i32(src_cell), _i32(poly_off), _i32(poly_cells)
        if src.shape != (nsub,) or poff.shape != (nsub + 1,):
            raise ValueError("src_cell must be [nsub], poly_off [nsub+1]")
        if pc.size == 0:
            pc = np.zeros(1, np.int32)
        QR = np.empty((nsub, ntime))
        _check(lib().wfa_route(self._h, C.c_int(ntime), C.c_int(nsub), _dp(QS), _ip(src), _ip(poff), _ip(pc),
                               C.c_int(flags), _dp(QR)))
        return QR

    def router(self, src_cell, poly_off, poly_cells, flags: int = 0):
        """Month-invariant routing operator for these source cells and polygons (see Router)."""
        return Router(self, src_cell, poly_off, poly_cells, flags)

    def accumulate(self, W, flags: int = 0):
        """Dense AreaD8 -wg.  W [nmonth, nrow, ncol] (or [nrow, ncol]) -> A of the same shape, NaN = nodata."""
        W = _f64(W)
        single = W.ndim == 2
        W3 = W.reshape(1, *W.shape) if single else W
        if W3.shape[1:] != (self.nrow, self.ncol):
            raise ValueError("weight grid shape does not match the plan")
        A = np.empty_like(W3)
        _check(lib().wfa_accumulate(self._h, C.c_int(W3.shape[0]), _dp(W3), C.c_int(flags), _dp(A)))
        return A[0] if single else A

    def accumulate_dev(self, ptr: int, nmonth: int, ld: int, flags: int = 0):
        _check(lib().wfa_accumulate_dev(self._h, C.c_int(nmonth), C.c_int(ld), C.c_void_p(ptr), C.c_int(flags)))

    def mask_dev(self, ptr: int, nmonth: int, ld: int, flags: int = 0):
        _check(lib().wfa_mask_dev(self._h, C.c_int(nmonth), C.c_int(ld), C.c_void_p(ptr), C.c_int(flags)))


def flowdir_from_dem(dem, valid, want_filled: bool = False):
    """Depression fill + D8 directions on the device.  dem [nrow, ncol], valid bool mask.
    Returns flowdir int16 (TauDEM codes, -32768 nodata) and optionally the filled surface."""
    z = _f64(dem)
    v = np.ascontiguousarray(valid, dtype=np.uint8)
    if z.ndim != 2 or v.shape != z.shape:
        raise ValueError("dem and valid must be 2-D arrays of the same shape")
    fd = np.empty(z.shape, np.int16)
    filled = np.empty(z.shape) if want_filled else None
    _check(lib().wfa_flowdir_from_dem(_dp(z), v.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int(z.shape[0]),
                                      C.c_int(z.shape[1]), fd.ctypes.data_as(C.POINTER(C.c_int16)), _dp(filled)))
    return (fd, filled) if want_filled else fd


def flowdir_from_dem_dev(dem_ptr: int, valid_ptr: int, nrow: int, ncol: int, flowdir_ptr: int, filled_ptr: int = 0):
    """Device-pointer variant: dem float64, valid uint8, flowdir int16 (out), filled float64 (out, optional)."""
    _check(lib().wfa_flowdir_from_dem_dev(C.c_void_p(dem_ptr), C.c_void_p(valid_ptr), C.c_int(nrow), C.c_int(ncol),
                                          C.c_void_p(flowdir_ptr), C.c_void_p(filled_ptr or 0)))


def debug_math(op: int, x, y=None):
    """Test hook: element-wise device math (see k_debug_math in csrc/gr2m.cu)."""
    x = _f64(x).ravel()
    y = None if y is None else _f64(y).ravel()
    out = np.empty_like(x)
    _check(lib().gr2m_debug_math(C.c_int(op), C.c_int(x.size), _dp(x), _dp(y), _dp(out)))
    return out


def fp64_peak_rrr_tflops(iters: int = 4096) -> float:
    """DFMA throughput when all three operands are distinct per-thread registers (development probe)."""
    t = C.c_double()
    _check(lib().gr2m_fp64_peak_rrr(C.c_int(iters), C.byref(t)))
    return t.value
