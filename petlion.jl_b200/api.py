"""Host-side mirror of the reference's public API for the hot path: petlion() / simulate() / simulate!().

Mirrors /root/reference/src/external.jl:2-36 (petlion), /root/reference/src/model_evaluation.jl:10-97
(simulate, simulate!) with the same keyword names, argument meaning and error behaviour -- but batched:
every parameter in `p.θ` and every input value may be a scalar or an array of length B, and one call
integrates B independent cells on the GPU through the C ABI (include/petlion_b200.h).

Julia is not available in the build image, so this Python module plays the role of the thin Julia
shim shown in INTEGRATION.md.  There is no CPU fallback.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _lib

CATHODES = {"LCO": 0, "NMC": 1, "NMC_LGM50": 2}
RXN = {"rxn_BV": 0, "rxn_MHC": 1}      # reaction rate laws (custom_functions.jl:212-298)
rxn_BV, rxn_MHC = "rxn_BV", "rxn_MHC"
METHODS = {"I": 0, "V": 1, "P": 2, "dT": 3, "η_p": 4, "eta_p": 4,
           # concentration-rate inputs (input_methods.jl:190-245): continuation runs of isothermal models without aging
           "dc_s_p_max": 6, "dc_s_p_min": 7, "dc_s_n_max": 8, "dc_s_n_min": 9, "dc_e_max": 10, "dc_e_min": 11}
_DC = {k for k in METHODS if k.startswith("dc_")}
EXIT_REASONS = {  # src/checks.jl
    -1: "running", 0: "Final time reached", 1: "Below min. voltage", 2: "Above max. voltage",
    3: "Below min. SOC", 4: "Above max. SOC", 5: "Above max. temperature", 6: "Above max. c_s_n",
    7: "Above max. C-rate", 8: "Below min. C-rate", 9: "Below min. c_e", 10: "Above max. film growth rate",
    11: "Below min. η_plating",
}
HARD_FAILURES = {
    -1: "Could not initialize DAE in 100 iterations.",          # model_evaluation.jl:456
    -2: "Model failed to converge",                              # checks.jl:233-236
    -3: "Model failed to converge (error test failures)",
    -4: "Reached max iterations",                                # checks.jl:239
    -5: "Non-finite state",
    -6: "The initial SOC is outside SOC_min/SOC_max for the requested (dis)charge",  # checks.jl:327-339
    -7: "An earlier segment of this simulation failed",
}


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def copy(self):
        return _NS(**self.__dict__)

    def __repr__(self):
        return "(" + ", ".join(f"{k}={v!r}" for k, v in self.__dict__.items()) + ")"


class Model:
    """`p = petlion(LCO)`: parameters p.θ, p.opts, p.bounds, p.N, p.numerics (src/external.jl:2-70)."""

    def __init__(self, cathode, N, numerics, device, devices=None):
        L = _lib.lib()
        self.cathode = cathode
        self.N = N
        self.numerics = numerics
        desc = _lib.ModelDesc(CATHODES[cathode], N.p, N.s, N.n, N.a, N.z, N.r_p, N.r_n,
                              int(bool(numerics.temperature)), int(bool(numerics.aging)), int(device),
                              RXN[numerics.rxn_p], RXN[numerics.rxn_n], int(numerics.Fickian_method == "spectral"))
        self._g = None
        self.devices = [int(device)] if devices is None else [int(d) for d in devices]
        if devices is not None and len(self.devices) > 1:
            # several GPUs behind one model: the library shards every batch contiguously over them
            g = C.c_void_p()
            devs = (C.c_int * len(self.devices))(*self.devices)
            _lib.check(L.plb_group_create(C.byref(desc), len(self.devices), devs, C.byref(g)))
            self._g = g
            h = C.c_void_p(L.plb_group_handle(g, 0))
        else:
            desc.device = self.devices[0]
            h = C.c_void_p()
            _lib.check(L.plb_create(C.byref(desc), C.byref(h)))
        self._h = h
        self.N.tot = L.plb_nstates(h)
        self.N.diff = L.plb_ndiff(h)
        self.N.alg = self.N.tot - self.N.diff
        self.ind = self._index_state()
        nth = L.plb_ntheta(h)
        keys = (C.c_char_p * nth)()
        L.plb_theta_keys(h, keys)
        self.θ_keys = [k.decode("utf-8") for k in keys]
        row = np.zeros(nth)
        L.plb_theta_defaults(h, row.ctypes.data)
        self.θ = OrderedDict(zip(self.θ_keys, row.tolist()))
        b = _lib.Bounds()
        L.plb_bounds_defaults(h, C.byref(b))
        self.bounds = _NS(**{n: getattr(b, n) for n, _ in _lib.Bounds._fields_})
        o = _lib.Opts()
        L.plb_opts_defaults(h, C.byref(o))
        self.opts = _NS(SOC=1.0, outputs=("t", "V"), abstol=o.abstol, reltol=o.reltol, maxiters=o.maxiters,
                        check_bounds=bool(o.check_bounds), interp_final=bool(o.interp_final), verbose=False,
                        n_save_max=512, tstops=[], tdiscon=[], initialize_algebraic_derivatives=True)

    theta = property(lambda self: self.θ)

    def _index_state(self):
        """state_indices: src/external.jl:275-365 (0-based slices into a state row; sections in their
        reference order, c_s_avg particle-major with the surface in the last slot of each particle)"""
        N, nm = self.N, self.numerics
        Nx = N.p + N.s + N.n
        sizes = [("c_e", Nx), ("c_s_avg", N.p * N.r_p + N.n * N.r_n)]
        if nm.temperature:
            sizes.append(("T", N.a + Nx + N.z))
        if nm.aging:
            sizes += [("film", N.n), ("SOH", 1)]
        sizes += [("j", N.p + N.n), ("Φ_e", Nx), ("Φ_s", N.p + N.n)]
        if nm.aging:
            sizes.append(("j_s", N.n))
        sizes.append(("I", 1))
        ind, k = {}, 0
        for name, n in sizes:
            ind[name] = slice(k, k + n)
            k += n
        assert k == N.tot, (k, N.tot)
        ind["Phi_e"], ind["Phi_s"] = ind["Φ_e"], ind["Φ_s"]
        return ind

    def __del__(self):
        try:
            if getattr(self, "_g", None):
                _lib.lib().plb_group_destroy(self._g)
                self._g = self._h = None
            elif getattr(self, "_h", None):
                _lib.lib().plb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def device_summaries(self, k=0):
        """(device pointer on self.devices[k], rows per device) of the all-gathered summaries of the last simulate()
        of a multi-GPU model: every device holds the whole batch's 80-byte records"""
        if self._g is None:
            raise ValueError("a single-device model has no gathered summaries")
        ptr, rows = C.c_void_p(), C.c_int()
        _lib.check(_lib.lib().plb_group_device_summaries(self._g, k, C.byref(ptr), C.byref(rows)))
        return ptr.value, rows.value

    # ---- parameter batch -----------------------------------------------------------------------
    def batch_size(self, *extra):
        B = 1
        for v in list(self.θ.values()) + list(extra):
            if v is None or isinstance(v, str):
                continue
            n = np.size(v)
            if n > 1:
                if B > 1 and n != B:
                    raise ValueError(f"inconsistent batch sizes {B} and {n}")
                B = n
        return B

    def theta_matrix(self, B=None):
        """update_θ! (generate_functions.jl:364-372): dict -> dense [B, nθ] in θ_keys order."""
        B = B or self.batch_size()
        th = np.empty((B, len(self.θ_keys)))
        for i, k in enumerate(self.θ_keys):
            th[:, i] = np.broadcast_to(np.asarray(self.θ[k], dtype=np.float64), (B,))
        return th

    def I1C(self, B=None):
        th = self.theta_matrix(B)
        out = np.zeros(th.shape[0])
        _lib.check(_lib.lib().plb_calc_I1C(self._h, th.shape[0], th.ctypes.data, out.ctypes.data))
        return out

    def jac_pattern(self, method="I", one_based=False):
        L = _lib.lib()
        nnz = L.plb_jac_nnz(self._h, METHODS[method])
        colptr = np.zeros(self.N.tot + 1, dtype=np.int32)
        rowval = np.zeros(nnz, dtype=np.int32)
        _lib.check(L.plb_jac_pattern(self._h, METHODS[method], colptr.ctypes.data_as(C.POINTER(C.c_int)),
                                     rowval.ctypes.data_as(C.POINTER(C.c_int)), int(one_based)))
        return colptr, rowval

    # ---- operator level (callback surface) -----------------------------------------------------
    def initial_guess(self, SOC, theta=None):
        th = self.theta_matrix() if theta is None else np.ascontiguousarray(theta)
        B = th.shape[0]
        soc = np.ascontiguousarray(np.broadcast_to(np.asarray(SOC, dtype=np.float64), (B,)))
        Y0 = np.zeros((B, self.N.tot))
        _lib.check(_lib.lib().plb_initial_guess(self._h, B, soc.ctypes.data, th.ctypes.data, Y0.ctypes.data, 0))
        return Y0

    def resjac(self, Y, YP, gamma, method="I", value=0.0, theta=None, want_res=True, want_jac=True):
        """R_full / J_full over a batch: returns (res [B,N], nzval [B,nnz])."""
        L = _lib.lib()
        Y = np.ascontiguousarray(np.atleast_2d(Y), dtype=np.float64)
        YP = np.ascontiguousarray(np.atleast_2d(YP), dtype=np.float64)
        B = Y.shape[0]
        th = self.theta_matrix(B) if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(gamma, dtype=np.float64), (B,)))
        vals = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (B,)))
        run = _lib.Run(METHODS[method], 0, 0.0, 1e6, 1, 0)
        nnz = L.plb_jac_nnz(self._h, METHODS[method])
        res = np.zeros((B, self.N.tot)) if want_res else None
        nz = np.zeros((B, nnz)) if want_jac else None
        _lib.check(L.plb_resjac(self._h, B, Y.ctypes.data, YP.ctypes.data, g.ctypes.data, th.ctypes.data,
                                C.byref(run), vals.ctypes.data, res.ctypes.data if want_res else None,
                                nz.ctypes.data if want_jac else None, 0))
        return res, nz

    def newton_init(self, Y, method="I", value=0.0, theta=None, reltol_init=None, abstol_init=None):
        L = _lib.lib()
        Y = np.array(np.atleast_2d(Y), dtype=np.float64, order="C")
        B = Y.shape[0]
        th = self.theta_matrix(B) if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        vals = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (B,)))
        YP = np.zeros_like(Y)
        st = np.zeros(B, dtype=np.int32)
        run = _lib.Run(METHODS[method], 0, 0.0, 1e6, 1, 0)
        o = _make_opts(self, {"reltol_init": reltol_init, "abstol_init": abstol_init})
        _lib.check(L.plb_newton_init(self._h, B, Y.ctypes.data, YP.ctypes.data, th.ctypes.data, C.byref(run),
                                     vals.ctypes.data, C.byref(o), st.ctypes.data, 0))
        return st, Y, YP

    def linear_solve(self, Y, YP, gamma, rhs, method="I", value=0.0, theta=None):
        """x = (dF/dY + gamma dF/dY')^{-1} rhs with the integrator's structured factorisation (KLU's role)."""
        L = _lib.lib()
        Y = np.ascontiguousarray(np.atleast_2d(Y), dtype=np.float64)
        YP = np.ascontiguousarray(np.atleast_2d(YP), dtype=np.float64)
        rhs = np.ascontiguousarray(np.atleast_2d(rhs), dtype=np.float64)
        B = Y.shape[0]
        th = self.theta_matrix(B) if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(gamma, dtype=np.float64), (B,)))
        vals = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (B,)))
        run = _lib.Run(METHODS[method], 0, 0.0, 1e6, 1, 0)
        x = np.zeros_like(Y)
        st = np.zeros(B, dtype=np.int32)
        _lib.check(L.plb_linear_solve(self._h, B, Y.ctypes.data, YP.ctypes.data, g.ctypes.data, th.ctypes.data,
                                      C.byref(run), vals.ctypes.data, rhs.ctypes.data, x.ctypes.data,
                                      st.ctypes.data, 0))
        return x, st


def model_key(p):
    """`<Cathode>_<Anode>/<sha1>` of a model's structural options -- the name of the reference's saved-model directory
    (strings_directory_func, src/external.jl:417-456), built from the same fields in the same order.  Models with the
    same key share the same compiled family and Jacobian pattern."""
    import hashlib
    nm, N = p.numerics, p.N
    anode, ocv_p, ocv_n, d_eff, k_eff = {
        "LCO": ("LiC6", "OCV_LCO", "OCV_LiC6", "D_eff_linear", "K_eff"),
        "NMC": ("LiC6_NMC", "OCV_NMC", "OCV_LiC6_NMC", "D_eff", "K_eff"),
        # (the OCVs of this set are closures named OCV_NMC / OCV_LiC6 inside NMC_LGM50 / LiC6_LGM50, params.jl:557, 632)
        "NMC_LGM50": ("LiC6_LGM50", "OCV_NMC", "OCV_LiC6", "D_eff_LGM50", "K_eff_LGM50")}[p.cathode]
    fields = [str(nm.temperature).lower(), nm.solid_diffusion, nm.Fickian_method, "SEI" if nm.aging else "false",
              nm.rxn_p, nm.rxn_n, ocv_p, ocv_n, "D_s_eff", "rxn_rate", d_eff, k_eff, "thermodynamic_factor_linear", nm.jacobian,
              f"Np{N.p}", f"Ns{N.s}", f"Nn{N.n}", f"Na{N.a}_Nz{N.z}" if nm.temperature else "",
              f"Nr_p{N.r_p}_Nr_n{N.r_n}" if nm.solid_diffusion == "Fickian" else ""]
    return f"{p.cathode}_{anode}/" + hashlib.sha1("_".join(fields).encode()).hexdigest()


class Solution:
    """`sol`: per-system trajectories t, V, I, SOC [B, n] (+ n_points), final state Y, results list."""

    def __init__(self):
        self.t = self.V = self.I = self.SOC = self.T = None
        self.states = None      # [B, n, N] when simulate(..., outputs=:all or state names)
        self.n_points = None
        self.truncated = None   # [B] bool: the run took more steps than n_save_max rows (the summary is complete)
        self.dense = None       # simulate(p, tf::Vector): dict(t, V, I, SOC, T[, Y], n) at the requested times
        self.Y = self.YP = None
        self._SOC_end = self._t_end = None
        self.results = []

    def __len__(self):
        return len(self.results)

    def isempty(self):
        return len(self.results) == 0

    def __call__(self, t, k=3, interp_bc="interpolate", system=None):
        """sol(t): the saved rows re-interpolated at times t with a spline of degree k per run, as the reference
        does with Dierckx.Spline1D (src/save_outputs.jl:74-133); scipy's splrep/splev are the same FITPACK
        routines.  interp_bc: "interpolate" (nearest value outside a run) or "extrapolate".
        Returns dict(t, V, I, SOC, T) with [B, len(t)] arrays (or [len(t)] for one `system`)."""
        from scipy.interpolate import splev, splrep
        if interp_bc not in ("interpolate", "extrapolate"):
            raise ValueError("Invalid interp_bc method.")
        t = np.atleast_1d(np.asarray(t, dtype=np.float64))
        systems = range(self.t.shape[0]) if system is None else [system]
        out = {key: np.full((len(systems), t.size), np.nan) for key in ("V", "I", "SOC", "T")}
        for row, s_ in enumerate(systems):
            # tspans of the runs of this system (results[i].tspan), rows of each run
            ends = np.cumsum([int(r.n_rows[s_]) for r in self.results])
            starts = np.concatenate([[0], ends[:-1]])
            spans = [(self.t[s_, a], self.t[s_, b - 1]) for a, b in zip(starts, ends) if b > a]
            rows = [(a, b) for a, b in zip(starts, ends) if b > a]
            which = np.full(t.size, len(spans) - 1)
            for i in range(len(spans) - 1, -1, -1):      # tspan_index: the first run whose span holds t
                which[(t >= spans[i][0]) & (t <= spans[i][1])] = i
            which[t < spans[0][0]] = 0
            for i, (a, b) in enumerate(rows):
                sel = which == i
                if not sel.any():
                    continue
                x = self.t[s_, a:b]
                keep = np.concatenate([[True], np.diff(x) > 0])      # FITPACK needs strictly increasing abscissae
                for key in out:
                    y = getattr(self, key)[s_, a:b]
                    if not np.all(np.isfinite(y[keep])):
                        continue
                    kk = min(k, int(keep.sum()) - 1)
                    if kk < 1:
                        out[key][row, sel] = y[0]
                        continue
                    tck = splrep(x[keep], y[keep], k=kk, s=0)
                    out[key][row, sel] = splev(t[sel], tck, ext=0 if interp_bc == "extrapolate" else 3)
        res = {"t": t}
        res.update({key: (v if system is None else v[0]) for key, v in out.items()})
        return res

    def state(self, p, name, system=0):
        """sol.c_e, sol.c_s_avg, sol.T, sol.j, sol.Φ_e, sol.Φ_s, sol.film, sol.j_s, sol.SOH of one system:
        rows = saved steps (needs simulate(..., outputs="all") or outputs=(name, ...))"""
        if self.states is None:
            raise ValueError("the states were not kept: pass outputs=\"all\" (or the state names) to simulate")
        return self.states[system, :self.n_points[system], p.ind[name]]


def petlion(cathode="LCO", *, N_p=10, N_s=10, N_n=10, N_a=10, N_z=10, N_r_p=10, N_r_n=10, temperature=None,
            solid_diffusion="Fickian", Fickian_method="finite_difference", aging=False, jacobian="symbolic",
            rxn_p="rxn_BV", rxn_n="rxn_BV", device=0, devices=None):
    """petlion(cathode; kwargs...) -- src/external.jl:2-18, src/params.jl:119-174.
    rxn_p / rxn_n: "rxn_BV" (default) or "rxn_MHC" (custom_functions.jl:212-298), per electrode."""
    if cathode not in CATHODES:
        raise ValueError(f"unknown cathode {cathode!r}; built: {list(CATHODES)}")
    if temperature is None:
        # system_LCO_LiC6 / system_NMC_LiC6 default to temperature = false, system_LGM50_NMC_LiC6 to true (params.jl:139, 686).
        # (That system's default aging = :stress has no implementation in the reference -- its first stop check throws a
        #  MethodError, checks.jl:204-206 -- so `aging` defaults to false here for every parameter set.)
        temperature = cathode == "NMC_LGM50"
    if solid_diffusion != "Fickian":
        # (:quadratic / :polynomial read states[:D_s_eff] before any method creates it and throw a KeyError at model
        #  construction in the reference itself: DESIGN.md section 7)
        raise NotImplementedError("only solid_diffusion=:Fickian is built")
    if Fickian_method not in ("finite_difference", "spectral"):
        raise ValueError("`Fickian_method` can either be :finite_difference or :spectral")     # params.jl:142
    if jacobian not in ("symbolic", "AD"):
        raise ValueError("`jacobian` can either be :symbolic or :AD")   # checks.jl:377-383
    if aging not in (False, True, "SEI"):       # params.jl:119-174: aging = false | :SEI
        raise ValueError(f"unknown aging model {aging!r}; built: False, \"SEI\"")
    for r in (rxn_p, rxn_n):
        if r not in RXN:
            raise ValueError(f"unknown reaction rate law {r!r}; built: {list(RXN)}")
    N = _NS(p=N_p, s=N_s, n=N_n, a=N_a, z=N_z, r_p=N_r_p, r_n=N_r_n)
    numerics = _NS(temperature=temperature, solid_diffusion=solid_diffusion, Fickian_method=Fickian_method,
                   aging=aging, jacobian=jacobian, cathode=cathode, rxn_p=rxn_p, rxn_n=rxn_n)
    return Model(cathode, N, numerics, device, devices)


def _make_opts(p, kw):
    o = _lib.Opts()
    o.abstol = kw.get("abstol") if kw.get("abstol") is not None else p.opts.abstol
    o.reltol = kw.get("reltol") if kw.get("reltol") is not None else p.opts.reltol
    o.abstol_init = kw.get("abstol_init") if kw.get("abstol_init") is not None else o.abstol
    o.reltol_init = kw.get("reltol_init") if kw.get("reltol_init") is not None else o.reltol
    o.maxiters = kw.get("maxiters") if kw.get("maxiters") is not None else p.opts.maxiters
    cb = kw.get("check_bounds")
    o.check_bounds = int(p.opts.check_bounds if cb is None else cb)
    itf = kw.get("interp_final")
    o.interp_final = int(p.opts.interp_final if itf is None else itf)
    iad = kw.get("initialize_algebraic_derivatives")
    o.skip_alg_deriv = int(not (p.opts.initialize_algebraic_derivatives if iad is None else iad))
    return o


_BOUND_NAMES = [n for n, _ in _lib.Bounds._fields_]
_BOUND_ALIASES = {"η_plating_min": "eta_plating_min"}


class Table:
    """A run_function input (structures.jl:55-63, examples/variable_input_functions.ipynb) as data: a
    piecewise-linear function of the run's local time.  A repeated knot time is a jump (right-continuous):
    `I_fun1(t) = t < 100 ? 1 : 0.5`  ==  Table([0, 100, 100], [1, 1, 0.5]).  `scale` (scalar or per-system
    array) multiplies the table per system.  Julia closures themselves cannot cross the C ABI."""

    def __init__(self, t, v, scale=None):
        self.t = np.ascontiguousarray(t, dtype=np.float64)
        self.v = np.ascontiguousarray(v, dtype=np.float64)
        if self.t.ndim != 1 or self.t.shape != self.v.shape or self.t.size < 1:
            raise ValueError("Table: t and v must be 1-d arrays of the same, non-zero length")
        if np.any(np.diff(self.t) < 0):
            raise ValueError("Table: knot times must be non-decreasing")
        self.scale = scale

    @classmethod
    def sample(cls, func, t_knots, scale=None):
        """tabulate a continuous func(t) on the given knots"""
        t = np.asarray(t_knots, dtype=np.float64)
        return cls(t, [func(float(x)) for x in t], scale)

    def jumps(self):
        """times of the jumps, the natural `tdiscon`"""
        return [float(a) for a, b in zip(self.t[:-1], self.t[1:]) if a == b]

    def __call__(self, t):
        k = int(np.searchsorted(self.t, t, side="right")) - 1
        if k < 0:
            return float(self.v[0])
        if k == self.t.size - 1:
            return float(self.v[-1])
        return float(self.v[k] + (self.v[k + 1] - self.v[k]) * ((t - self.t[k]) / (self.t[k + 1] - self.t[k])))


def simulate(p, tf=1e6, *, sol=None, SOC=None, abstol=None, reltol=None, abstol_init=None, reltol_init=None,
             maxiters=None, check_bounds=None, interp_final=None, n_save_max=None, tdiscon=None,
             initialize_algebraic_derivatives=None, outputs=None, tstops=None, dense_t=None, **inputs):
    """simulate(p, tf; I=..|V=..|P=.., SOC, V_max, V_min, SOC_max, ...) -- model_evaluation.jl:10-86.

    Inputs may be numbers (scalar or per-system arrays), "hold" or "rest" (Julia :hold / :rest), or a
    `Table` (a tabulated run_function; `tdiscon` = the known discontinuities, structures.jl:279)."""
    L = _lib.lib()
    bounds = _lib.Bounds(**{n: getattr(p.bounds, n) for n in _BOUND_NAMES})
    method_kw = {}
    for k, v in inputs.items():
        k2 = _BOUND_ALIASES.get(k, k)
        if k2 in _BOUND_NAMES:
            setattr(bounds, k2, float(v))
        elif k in METHODS:
            method_kw[k] = v
        else:
            # check_input_arguments, checks.jl:278-325
            raise TypeError(f"ERROR\n--------\n Invalid keyword argument: {k}")
    if len(method_kw) == 0:
        raise TypeError("ERROR\n--------\n No inputs are selected, choose one from: (I, V, P, dT, η_p)")
    if len(method_kw) > 1:
        raise TypeError("ERROR\n--------\n Cannot select more than one input from: (I, V, P, dT, η_p)")
    (name, inp), = method_kw.items()
    new_run = sol is None or sol.isempty()
    kind, value, vals, table = 0, 0.0, None, None
    if name == "dT" and not p.numerics.temperature:
        raise ValueError("Temperature must be enabled when using `dT`.")      # input_methods.jl:183
    if name in _DC:
        if new_run:
            raise ValueError(f"`{name}` needs a previous solution: use it with simulate!")     # @assert !isempty(sol.Y), input_methods.jl:196
        if isinstance(inp, Table) or (isinstance(inp, str) and inp != "hold"):
            raise ValueError(f"`{name}` takes a number or :hold")
    if isinstance(inp, str):
        if inp == "hold":
            if new_run:
                raise ValueError("Cannot use `:hold` without a previous simulation.")   # checks.jl:385
            kind = 1
        elif inp == "rest" and name in ("I", "P"):
            kind = 2
        else:
            raise ValueError("Unsupported input symbol.")                                # input_methods.jl:23
    elif isinstance(inp, Table):
        if name == "dT":
            raise ValueError("dT takes a number or :hold")
        table = inp
        if inp.scale is not None:
            vals = np.asarray(inp.scale, dtype=np.float64)
    elif callable(inp):
        raise NotImplementedError("a closure cannot cross the C ABI: tabulate it with Table.sample(func, knots)")
    else:
        vals = np.asarray(inp, dtype=np.float64)
    B = p.batch_size(vals, SOC) if new_run else sol.Y.shape[0]
    th = p.theta_matrix(B)
    if vals is not None:
        vals = np.ascontiguousarray(np.broadcast_to(vals, (B,)))
    # tf::AbstractVector (model_evaluation.jl:13, 80): run to tf[end], results also at every tf[i] (sol.dense)
    # dense_t: requested GLOBAL times without touching tf (what a continuation needs: its tf is local to the run)
    t_dense = None
    if np.ndim(tf) > 0 or dense_t is not None:
        t_dense = np.ascontiguousarray(np.ravel(tf if dense_t is None else dense_t), dtype=np.float64)
        if np.ndim(tf) > 0 and (np.size(tf) == 0 or np.any(np.diff(np.ravel(tf)) < 0)):
            raise ValueError("tf must be a number or an ascending vector of times")
        if t_dense.size == 0 or np.any(np.diff(t_dense) < 0):
            raise ValueError("the requested times must be ascending")
    # (like the reference, tf[end] is the run's local final time and the requested times are global ones)
    run = _lib.Run(METHODS[name], kind, value, float(np.ravel(tf)[-1]), int(new_run), 0)
    o = _make_opts(p, dict(abstol=abstol, reltol=reltol, abstol_init=abstol_init, reltol_init=reltol_init,
                           maxiters=maxiters, check_bounds=check_bounds, interp_final=interp_final,
                           initialize_algebraic_derivatives=initialize_algebraic_derivatives))
    N = p.N.tot
    if new_run:
        sol = Solution() if sol is None else sol
        sY = np.zeros((B, N)); sYP = np.zeros((B, N)); sSOC = np.zeros(B); st = np.zeros(B)
        soc0 = np.ascontiguousarray(np.broadcast_to(np.asarray(p.opts.SOC if SOC is None else SOC,
                                                               dtype=np.float64), (B,)))
    else:
        sY, sYP, sSOC, st = sol.Y, sol.YP, sol._SOC_end, sol._t_end
        soc0 = None
    ns = p.opts.n_save_max if n_save_max is None else n_save_max
    summ = np.zeros(B, dtype=_lib.SUMMARY_DTYPE)
    tr = {k: np.full((B, max(ns, 1)), np.nan) for k in ("t", "V", "I", "SOC", "T")}
    # outputs (params.jl:262, save_outputs.jl:11-40): t, V, I, SOC, T are always kept; :all or any state name
    # keeps the full state row of every saved step
    outs = p.opts.outputs if outputs is None else outputs
    outs = (outs,) if isinstance(outs, str) else tuple(outs)
    bad = [o for o in outs if o not in ("all", "t", "V", "I", "P", "SOC", "T", "Y", "YP") and o not in p.ind]
    if bad:
        raise ValueError(f"unknown output {bad[0]!r}")
    keep_states = ns > 0 and any(o in ("all", "Y") or (o in p.ind and o not in ("T", "I")) for o in outs)
    if keep_states and sol is not None and not sol.isempty() and sol.states is None:
        raise ValueError("cannot start keeping states in the middle of a solution")
    trY = np.full((B, ns, N), np.nan) if keep_states else None
    trn = np.zeros(B, dtype=np.int32)
    ts = np.ascontiguousarray(p.opts.tstops if tstops is None else tstops, dtype=np.float64)
    _lib.check(L.plb_set_tstops(p._h, ts.size, ts.ctypes.data if ts.size else None))
    tail = (C.byref(o), C.byref(bounds), None if soc0 is None else soc0.ctypes.data,
            sY.ctypes.data, sYP.ctypes.data, sSOC.ctypes.data, st.ctypes.data,
            summ.ctypes.data, ns, tr["t"].ctypes.data if ns else None,
            tr["V"].ctypes.data if ns else None, tr["I"].ctypes.data if ns else None,
            tr["SOC"].ctypes.data if ns else None, tr["T"].ctypes.data if ns else None,
            trY.ctypes.data if keep_states else None,
            trn.ctypes.data, 0)
    vptr = None if vals is None else vals.ctypes.data
    dense = None
    if t_dense is not None:
        nd = t_dense.size
        dense = dict(t=t_dense, n=np.zeros(B, dtype=np.int32))
        for key in ("V", "I", "SOC", "T"):
            dense[key] = np.full((B, nd), np.nan)
        dense["Y"] = np.full((B, nd, N), np.nan) if keep_states else None
        _lib.check(L.plb_set_dense_output(p._h, nd, t_dense.ctypes.data, dense["V"].ctypes.data, dense["I"].ctypes.data,
                                          dense["SOC"].ctypes.data, dense["T"].ctypes.data,
                                          dense["Y"].ctypes.data if keep_states else None, dense["n"].ctypes.data, 0))
    if p._g is not None:
        if table is not None or dense is not None or keep_states or ts.size:
            raise NotImplementedError("a multi-GPU model fans out plain simulate()/simulate!() calls; tabulated inputs, "
                                      "requested times, state rows and tstops are per-device features")
        _lib.check(L.plb_group_simulate(p._g, B, th.ctypes.data, C.byref(run), vptr, *tail[:-3], tail[-2]))
    elif table is not None:
        dpp = C.POINTER(C.c_double)
        td = np.ascontiguousarray(sorted(tdiscon if tdiscon is not None else p.opts.tdiscon), dtype=np.float64)
        it = _lib.InputTable(table.t.size, table.t.ctypes.data_as(dpp), table.v.ctypes.data_as(dpp), td.size,
                             td.ctypes.data_as(dpp) if td.size else None)
        _lib.check(L.plb_simulate_table(p._h, B, th.ctypes.data, C.byref(run), C.byref(it), vptr, *tail))
    else:
        _lib.check(L.plb_simulate(p._h, B, th.ctypes.data, C.byref(run), vptr, *tail))
    hard = summ["flag"] < 0
    if B == 1 and hard[0]:
        # the reference throws for a single simulation (model_evaluation.jl:456, checks.jl:233-239)
        raise RuntimeError(HARD_FAILURES.get(int(summ["flag"][0]), "simulation failed"))
    sol.Y, sol.YP, sol._SOC_end, sol._t_end = sY, sYP, sSOC, st
    # the reference keeps every step; here the rows live in caller-sized buffers: say so when a run outgrew them
    trunc = (summ["n_steps"] + 1 > ns) if ns > 0 else np.zeros(B, dtype=bool)
    if ns > 0 and trunc.any():
        import warnings
        warnings.warn(f"{int(trunc.sum())} of {B} runs took more steps (up to {int(summ['n_steps'].max())}) than "
                      f"n_save_max = {ns} rows: their trajectories are truncated (summaries and final states are "
                      f"complete); pass a larger n_save_max", RuntimeWarning, stacklevel=2)
    sol.truncated = trunc if sol.truncated is None or new_run else (sol.truncated | trunc)
    if dense is not None:
        sol.dense = dense
    if sol.t is None or new_run:
        sol.t, sol.V, sol.I, sol.SOC, sol.T, sol.n_points = tr["t"], tr["V"], tr["I"], tr["SOC"], tr["T"], trn.copy()
        sol.states = trY
    else:
        # append the new run's rows after the existing ones (per system)
        width = int((sol.n_points + trn).max())
        for k in ("t", "V", "I", "SOC", "T"):
            old = getattr(sol, k)
            new = np.full((B, width), np.nan)
            for s in range(B):
                new[s, :sol.n_points[s]] = old[s, :sol.n_points[s]]
                new[s, sol.n_points[s]:sol.n_points[s] + trn[s]] = tr[k][s, :trn[s]]
            setattr(sol, k, new)
        if sol.states is not None:
            new = np.full((B, width, N), np.nan)
            for s in range(B):
                new[s, :sol.n_points[s]] = sol.states[s, :sol.n_points[s]]
                if trY is not None:
                    new[s, sol.n_points[s]:sol.n_points[s] + trn[s]] = trY[s, :trn[s]]
            sol.states = new
        sol.n_points = sol.n_points + trn
    sol.results.append(_NS(run=_NS(method=name, input=inp, tf=run.tf), summary=summ, n_rows=trn.copy(), truncated=trunc,
                           exit_reason=[EXIT_REASONS.get(int(f), HARD_FAILURES.get(int(f), "?")) for f in summ["flag"]],
                           kernel_ms=L.plb_last_kernel_ms(p._h)))
    return sol


def simulate_(sol, p, tf=1e6, **kw):
    """simulate!(sol, p, tf; ...) -- model_evaluation.jl:87-97: continue `sol` with a new input."""
    return simulate(p, tf, sol=sol, **kw)
