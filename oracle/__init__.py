"""CPU ORACLE -- test infrastructure, not product code.

ctypes wrapper around oracle/liboracle.so (plain-C restatement of the PETLION.jl hot path,
see oracle/petlion_oracle.h).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

CATHODE = {"LCO": 0, "NMC": 1, "NMC_LGM50": 2, "LGM50": 2}
METHOD = {"I": 0, "V": 1, "P": 2, "dT": 3, "η_p": 4, "eta_p": 4, "dc": 6, "dc_alg": 7}
DC_KIND = {"dc_s_p_max": 0, "dc_s_p_min": 1, "dc_s_n_max": 2, "dc_s_n_min": 3, "dc_e_max": 4, "dc_e_min": 5}


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "residual_impl.inc", "petlion_oracle.h")]
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs)):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB


class Model(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("N_p", "N_s", "N_n", "N_a", "N_z", "N_r_p", "N_r_n", "temperature", "aging", "cathode",
                 "rxn_p", "rxn_n", "fickian_spectral")]


class Layout(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("Nx", "c_e", "c_s_p", "c_s_n", "T", "film", "SOH", "j", "phi_e", "phi_s", "j_s", "I",
                 "N_diff", "N_alg", "N_tot")]


class Run(C.Structure):
    _fields_ = [("method", C.c_int), ("value", C.c_double), ("tf", C.c_double),
                ("input_kind", C.c_int), ("new_run", C.c_int), ("t0", C.c_double),
                ("tab_n", C.c_int), ("tab_t", C.POINTER(C.c_double)), ("tab_v", C.POINTER(C.c_double)),
                ("scale", C.c_double), ("n_tdiscon", C.c_int), ("tdiscon", C.POINTER(C.c_double)),
                ("last_value", C.POINTER(C.c_double)), ("n_tstops", C.c_int), ("tstops", C.POINTER(C.c_double)),
                ("dense", C.c_void_p), ("dense_sys", C.c_int), ("dc_kind", C.c_int), ("dc_ind", C.c_int)]


class Dense(C.Structure):
    _fields_ = [("n", C.c_int), ("t", C.POINTER(C.c_double)), ("V", C.POINTER(C.c_double)),
                ("I", C.POINTER(C.c_double)), ("SOC", C.POINTER(C.c_double)), ("T", C.POINTER(C.c_double)),
                ("Y", C.POINTER(C.c_double)), ("done", C.POINTER(C.c_int))]


class Opts(C.Structure):
    _fields_ = [("abstol", C.c_double), ("reltol", C.c_double), ("abstol_init", C.c_double),
                ("reltol_init", C.c_double), ("maxiters", C.c_int), ("check_bounds", C.c_int),
                ("interp_final", C.c_int), ("ida_maxord", C.c_int), ("ida_maxcor", C.c_int),
                ("ida_maxnef", C.c_int), ("ida_maxncf", C.c_int), ("skip_alg_deriv", C.c_int)]


class Bounds(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("V_max", "V_min", "SOC_max", "SOC_min", "T_max", "c_s_n_max", "I_max", "I_min",
                 "eta_plating_min", "c_e_min", "dfilm_max")]


class Summary(C.Structure):
    _fields_ = [("t_end", C.c_double), ("V_end", C.c_double), ("I_end", C.c_double),
                ("SOC_end", C.c_double), ("T_end", C.c_double), ("flag", C.c_int), ("n_steps", C.c_int),
                ("n_res", C.c_int), ("n_jac", C.c_int), ("n_netf", C.c_int), ("n_ncfn", C.c_int),
                ("n_newton_init", C.c_int), ("n_reinit", C.c_int)]


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.orc_theta_name.restype = C.c_char_p
        L.orc_calc_I1C.restype = C.c_double
        L.orc_rng_u01.restype = C.c_double
        L.orc_rng_u01.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_uint]
        L.orc_jac_pattern.argtypes = [C.POINTER(Model), C.c_int, _ip, _ip]
        L.orc_residual.argtypes = [C.POINTER(Model), _dp, C.POINTER(Run), C.c_double, _dp, _dp, _dp]
        L.orc_jacobian.argtypes = [C.POINTER(Model), _dp, C.POINTER(Run), C.c_double, _dp, _dp,
                                   C.c_double, _dp]
        L.orc_initial_guess.argtypes = [C.POINTER(Model), _dp, C.c_double, _dp]
        L.orc_newton_init.argtypes = [C.POINTER(Model), _dp, C.POINTER(Run), C.POINTER(Opts), _dp, _dp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def theta_names():
    L = lib()
    return [L.orc_theta_name(i).decode() for i in range(L.orc_ntheta())]


def theta_defaults(cathode="LCO"):
    L = lib()
    th = np.zeros(L.orc_ntheta())
    L.orc_theta_defaults(CATHODE[cathode], _p(th))
    return th


def theta_dict(cathode="LCO"):
    return dict(zip(theta_names(), theta_defaults(cathode)))


RXN = {"BV": 0, "MHC": 1, "rxn_BV": 0, "rxn_MHC": 1}


def make_model(cathode="LCO", N_p=10, N_s=10, N_n=10, N_a=10, N_z=10, N_r_p=10, N_r_n=10,
               temperature=False, aging=False, rxn_p="BV", rxn_n="BV", Fickian_method="finite_difference"):
    return Model(N_p, N_s, N_n, N_a, N_z, N_r_p, N_r_n, int(bool(temperature)), int(bool(aging)),
                 CATHODE[cathode], RXN[rxn_p], RXN[rxn_n], {"finite_difference": 0, "spectral": 1}[Fickian_method])


def layout(m):
    Lo = Layout()
    lib().orc_layout_make(C.byref(m), C.byref(Lo))
    return Lo


def default_opts(**kw):
    o = Opts()
    lib().orc_opts_defaults(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def default_bounds(cathode="LCO", **kw):
    b = Bounds()
    lib().orc_bounds_defaults(CATHODE[cathode], C.byref(b))
    for k, v in kw.items():
        setattr(b, k, v)
    return b


INPUT = {"value": 0, "hold": 1, "rest": 2}


def make_run(method="I", value=-1.0, tf=1e6, input_kind="value", new_run=True, t0=0.0, table=None,
             tdiscon=(), scale=1.0, tstops=()):
    """table = (t_knots, v_knots): a run_function restricted to a piecewise-linear table (see orc_run)"""
    dc_ind = 0
    if isinstance(method, tuple):            # ("dc", state index): operator-level calls on a given row
        method, dc_ind = method
    dc_kind = DC_KIND.get(method, 0)
    r = Run(METHOD["dc" if method in DC_KIND else method], float(value), float(tf), INPUT[input_kind], int(new_run), float(t0))
    r.dc_kind = dc_kind; r.dc_ind = int(dc_ind)
    r.scale = float(scale)
    if len(tstops):
        ts = np.ascontiguousarray(tstops, dtype=np.float64)
        r._keep_ts = ts
        r.n_tstops = ts.size; r.tstops = _p(ts)
    if table is not None:
        tt = np.ascontiguousarray(table[0], dtype=np.float64)
        vv = np.ascontiguousarray(table[1], dtype=np.float64)
        td = np.ascontiguousarray(sorted(tdiscon), dtype=np.float64)
        assert tt.ndim == 1 and tt.shape == vv.shape and tt.size >= 1 and np.all(np.diff(tt) >= 0)
        r._keep = (tt, vv, td)            # the struct only holds pointers
        r.tab_n = tt.size; r.tab_t = _p(tt); r.tab_v = _p(vv)
        r.n_tdiscon = td.size; r.tdiscon = _p(td) if td.size else None
    return r


def table_eval(table, t):
    tt = np.ascontiguousarray(table[0], dtype=np.float64)
    vv = np.ascontiguousarray(table[1], dtype=np.float64)
    L = lib()
    L.orc_table_eval.restype = C.c_double
    L.orc_table_eval.argtypes = [C.c_int, _dp, _dp, C.c_double]
    return L.orc_table_eval(tt.size, _p(tt), _p(vv), float(t))


def calc_I1C(theta):
    return lib().orc_calc_I1C(_p(np.ascontiguousarray(theta, dtype=np.float64)))


def initial_guess(m, theta, SOC):
    Y0 = np.zeros(layout(m).N_tot)
    lib().orc_initial_guess(C.byref(m), _p(np.ascontiguousarray(theta)), float(SOC), _p(Y0))
    return Y0


def residual(m, theta, run, t, Y, YP):
    res = np.zeros_like(Y)
    lib().orc_residual(C.byref(m), _p(np.ascontiguousarray(theta)), C.byref(run), float(t),
                       _p(np.ascontiguousarray(Y)), _p(np.ascontiguousarray(YP)), _p(res))
    return res


def jac_pattern(m, method="I"):
    L = lib()
    nnz = L.orc_jac_pattern(C.byref(m), METHOD[method], None, None)
    N = layout(m).N_tot
    colptr = np.zeros(N + 1, dtype=np.int32)
    rowval = np.zeros(nnz, dtype=np.int32)
    L.orc_jac_pattern(C.byref(m), METHOD[method], colptr.ctypes.data_as(_ip), rowval.ctypes.data_as(_ip))
    return colptr, rowval


def jacobian(m, theta, run, t, Y, YP, gamma):
    L = lib()
    nnz = L.orc_jac_pattern(C.byref(m), run.method, None, None)
    nz = np.zeros(nnz)
    L.orc_jacobian(C.byref(m), _p(np.ascontiguousarray(theta)), C.byref(run), float(t),
                   _p(np.ascontiguousarray(Y)), _p(np.ascontiguousarray(YP)), float(gamma), _p(nz))
    return nz


def newton_init(m, theta, run, opts, Y):
    Y = np.array(Y, dtype=np.float64)
    YP = np.zeros_like(Y)
    it = lib().orc_newton_init(C.byref(m), _p(np.ascontiguousarray(theta)), C.byref(run),
                               C.byref(opts), _p(Y), _p(YP))
    return it, Y, YP


def simulate_batch(m, theta, run, opts, bounds, SOC0=1.0, values=None, state=None, n_save_max=0,
                   nthreads=1, dense_t=None, dense_Y=False):
    """theta: [B, ntheta] (oracle order).  Returns dict of numpy arrays.
    state: dict(Y, YP, SOC, t) from a previous call (simulate! continuation).
    dense_t: ascending global times -> res["dense"] = dict(t, V, I, SOC, T[, Y], n): the integrator's own
    interpolant at those times (rows past the end of a run stay NaN)."""
    L = lib()
    theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    B = theta.shape[0]
    N = layout(m).N_tot
    if state is None:
        sY = np.zeros((B, N)); sYP = np.zeros((B, N)); sSOC = np.zeros(B); st = np.zeros(B)
    else:
        sY = np.array(state["Y"], dtype=np.float64); sYP = np.array(state["YP"], dtype=np.float64)
        sSOC = np.array(state["SOC"], dtype=np.float64); st = np.array(state["t"], dtype=np.float64)
    soc0 = np.ascontiguousarray(np.broadcast_to(np.asarray(SOC0, dtype=np.float64), (B,)))
    out = (Summary * B)()
    ns = max(int(n_save_max), 1)
    tr = {k: np.full((B, ns), np.nan) for k in ("t", "V", "I", "SOC")}
    trn = np.zeros(B, dtype=np.int32)
    vals = None if values is None else np.ascontiguousarray(np.broadcast_to(np.asarray(values, dtype=np.float64), (B,)))
    dense = None
    if dense_t is not None:
        dt_ = np.ascontiguousarray(dense_t, dtype=np.float64)
        assert dt_.ndim == 1 and np.all(np.diff(dt_) >= 0)
        dense = dict(t=dt_, n=np.zeros(B, dtype=np.int32))
        for k in ("V", "I", "SOC", "T"):
            dense[k] = np.full((B, dt_.size), np.nan)
        if dense_Y:
            dense["Y"] = np.full((B, dt_.size, N), np.nan)
        dstruct = Dense(dt_.size, _p(dt_), _p(dense["V"]), _p(dense["I"]), _p(dense["SOC"]), _p(dense["T"]),
                        _p(dense["Y"]) if dense_Y else None, dense["n"].ctypes.data_as(_ip))
        run = _copy_run(run)
        run._keep_dense = (dstruct, dense)
        run.dense = C.cast(C.pointer(dstruct), C.c_void_p)
    L.orc_simulate_batch(C.byref(m), B, _p(theta), C.byref(run), None if vals is None else _p(vals),
                         C.byref(opts), C.byref(bounds), _p(soc0), _p(sY), _p(sYP), _p(sSOC), _p(st),
                         out, ns if n_save_max else 0,
                         _p(tr["t"]) if n_save_max else None, _p(tr["V"]) if n_save_max else None,
                         _p(tr["I"]) if n_save_max else None, _p(tr["SOC"]) if n_save_max else None,
                         trn.ctypes.data_as(_ip), int(nthreads))
    res = {f: np.array([getattr(o, f) for o in out]) for f, _ in Summary._fields_}
    res.update(state=dict(Y=sY, YP=sYP, SOC=sSOC, t=st), traj=tr, traj_n=trn, dense=dense)
    return res


def _copy_run(run):
    r = Run()
    C.memmove(C.byref(r), C.byref(run), C.sizeof(Run))
    for k in ("_keep", "_keep_ts"):
        if hasattr(run, k):
            setattr(r, k, getattr(run, k))
    return r


def rng_u01(seed, system_id, param_id):
    return lib().orc_rng_u01(int(seed), int(system_id), int(param_id))
