"""Shared helpers for the parity tests: synthetic parameter batches (counter-based RNG),
theta mapping between the product's reference-ordered rows and the oracle's table."""
import ctypes as C

import numpy as np

import oracle as O

SEED = 20211


def splitmix_u01(seed, system_id, param_id):
    """Vectorised copy of oracle.orc_rng_u01 (splitmix64 counter RNG)."""
    sid = np.asarray(system_id, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) ^ (sid * np.uint64(1000003) + np.uint64(param_id))
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / 9007199254740992.0


# cfg2 randomisation: D_s, k log-uniform x/÷2 ; porosities uniform +-10 %
RANDOMISED = [("D_sp", "log"), ("D_sn", "log"), ("k_p", "log"), ("k_n", "log"),
              ("eps_p", "lin"), ("eps_n", "lin"), ("eps_s", "lin")]


def oracle_theta_batch(B, cathode="LCO", seed=SEED, first=0):
    names = O.theta_names()
    th = np.tile(O.theta_defaults(cathode), (B, 1))
    ids = np.arange(first, first + B)
    for pid, (name, kind) in enumerate(RANDOMISED):
        u = splitmix_u01(seed, ids, pid)
        col = names.index(name)
        if kind == "log":
            th[:, col] *= 10.0 ** (0.3 * (2 * u - 1))
        else:
            th[:, col] *= (0.9 + 0.2 * u)
    return th


def product_theta_from_oracle(p, th_oracle):
    """oracle-order [B, 62] -> product reference-order [B, ntheta] by key name"""
    from petlion_b200 import _lib
    L = _lib.lib()
    names = O.theta_names()
    out = np.zeros((th_oracle.shape[0], len(p.θ_keys)))
    filled = np.zeros(len(p.θ_keys), dtype=bool)
    for j, n in enumerate(names):
        i = L.plb_theta_index(p._h, n.encode())
        if i >= 0:
            out[:, i] = th_oracle[:, j]
            filled[i] = True
    assert filled.all()
    return out


def set_theta_batch(p, th_product):
    for i, k in enumerate(p.θ_keys):
        col = th_product[:, i]
        p.θ[k] = float(col[0]) if np.all(col == col[0]) else col.copy()


def random_states(m, th_oracle, seed=1, SOC_lo=0.1, SOC_hi=0.9, rel=0.02):
    """physically valid random states: initial guess at a random SOC, Newton-initialised by the oracle,
    then perturbed by a few percent (keeps concentrations positive)."""
    B = th_oracle.shape[0]
    L = O.layout(m)
    rng = np.random.default_rng(seed)
    Y = np.zeros((B, L.N_tot)); YP = np.zeros((B, L.N_tot))
    opts = O.default_opts()
    for s in range(B):
        soc = rng.uniform(SOC_lo, SOC_hi)
        cur = rng.choice([-2.0, -1.0, 0.5, 1.0, 2.0])
        run = O.make_run("I", cur)
        y0 = O.initial_guess(m, th_oracle[s], soc)
        y0[L.I] = cur
        it, y, yp = O.newton_init(m, th_oracle[s], run, opts, y0)
        assert it > 0
        scale = np.where(np.abs(y) > 1e-12, np.abs(y), 1e-3)
        Y[s] = y + rel * scale * rng.uniform(-1, 1, size=y.shape)
        YP[s] = yp * (1 + 0.1 * rng.uniform(-1, 1, size=y.shape))
    return Y, YP


# ---- protocols of BASELINE.json configs[1..4] (segments: method, input kind, value, local tf, bound overrides) --------
PROTOCOLS = {
    "cfg2": dict(cathode="LCO", soc0=1.0, segs=[("I", "value", -1.0, 1e6, {})]),
    "cfg3": dict(cathode="LCO", temperature=True, soc0=0.0,
                 segs=[("I", "value", 4.0, 1e6, {"V_max": 4.1}), ("V", "hold", 0.0, 1e6, {"V_max": 4.1})]),
    # the same with the reference's own end of a CV phase (examples/CC-CV.ipynb: I_min = 1/20) instead of running the
    # hold to SOC_max or 1e6 s
    "cfg3i": dict(cathode="LCO", temperature=True, soc0=0.0,
                  segs=[("I", "value", 4.0, 1e6, {"V_max": 4.1}), ("V", "hold", 0.0, 1e6, {"V_max": 4.1, "I_min": 0.05})]),
    "cfg4": dict(cathode="NMC", soc0=0.0,
                 segs=[s for _ in range(20) for s in (("I", "value", 1.0, 180.0, {}), ("I", "rest", 0.0, 7200.0, {}))]),
    "cfg5": dict(cathode="LCO", aging=True, grid=dict(N_p=20, N_s=20, N_n=20), soc0=0.0,
                 segs=[("I", "value", 1.0, 1e6, {"V_max": 4.2}), ("I", "value", -1.0, 1e6, {"V_max": 4.2})]),
    "cfg5n10": dict(cathode="LCO", aging=True, soc0=0.0,
                    segs=[("I", "value", 1.0, 1e6, {"V_max": 4.2}), ("I", "value", -1.0, 1e6, {"V_max": 4.2})]),
}


def oracle_protocol(W, tho, opts, dense_t=None, nthreads=8, n_segs=None):
    """the CPU oracle over a protocol; returns the list of per-segment results (dense rows per segment)"""
    m = O.make_model(W["cathode"], temperature=W.get("temperature", False), aging=W.get("aging", False), **W.get("grid", {}))
    out, state = [], None
    for k, (method, kind, value, tf, bo) in enumerate(W["segs"][:n_segs]):
        b = O.default_bounds(W["cathode"], **bo)
        run = O.make_run(method, value, tf=tf, input_kind=kind, new_run=(k == 0))
        r = O.simulate_batch(m, tho, run, opts, b, SOC0=W["soc0"], state=state, nthreads=nthreads, dense_t=dense_t)
        state = r["state"]
        out.append(r)
    return out


def gpu_protocol(P, p, W, dense_t=None, n_segs=None, **opts):
    """the product API over the same protocol: simulate() then simulate!(); returns (sol, [dense per segment])"""
    sol, dense = None, []
    for k, (method, kind, value, tf, bo) in enumerate(W["segs"][:n_segs]):
        inp = {method: value if kind == "value" else kind}
        if k == 0:
            sol = P.simulate(p, tf, SOC=W["soc0"], dense_t=dense_t, **inp, **bo, **opts)
        else:
            P.simulate_(sol, p, tf, dense_t=dense_t, **inp, **bo, **opts)
        dense.append(sol.dense)
    return sol, dense


def merge_dense(dense_list, keys=("V", "I", "SOC", "T")):
    """rows of a multi-segment protocol: every requested time is filled by the segment that covers it"""
    out = {k: np.full_like(dense_list[0][k], np.nan) for k in keys}
    for d in dense_list:
        for k in keys:
            fill = ~np.isnan(d[k])
            out[k][fill] = d[k][fill]
    return out
