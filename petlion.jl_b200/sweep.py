"""Synthetic parameter sweeps over a model's θ rows (the batch entry point is update_θ!,
/root/reference/src/generate_functions.jl:364-372: one dense θ row per simulation).

randomised_theta() is the sweep of BASELINE.json configs[1..4] / SURVEY.md 8(d): a counter-based RNG
(splitmix64 of seed, system id, parameter id -- the same stream on every rank and on the CPU side), D_s and k
log-uniform over x/÷2, porosities uniform over +-10 %."""
import numpy as np

SEED = 20211
RANDOMISED = [("D_sp", "log"), ("D_sn", "log"), ("k_p", "log"), ("k_n", "log"),
              ("ϵ_p", "lin"), ("ϵ_n", "lin"), ("ϵ_s", "lin")]


def splitmix_u01(seed, system_id, param_id):
    """u in [0, 1) from splitmix64(seed ^ (system_id * 1000003 + param_id))"""
    sid = np.asarray(system_id, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) ^ (sid * np.uint64(1000003) + np.uint64(param_id))
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / 9007199254740992.0


def randomised_theta(p, B, first=0, seed=SEED):
    """[B, nθ] rows in p.θ_keys order for systems first .. first+B-1 around the model's default parameters"""
    from . import _lib
    row = np.zeros(len(p.θ_keys))
    _lib.check(_lib.lib().plb_theta_defaults(p._h, row.ctypes.data))
    th = np.tile(row, (B, 1))
    ids = np.arange(first, first + B)
    for pid, (name, kind) in enumerate(RANDOMISED):
        u = splitmix_u01(seed, ids, pid)
        col = p.θ_keys.index(name)
        th[:, col] *= 10.0 ** (0.3 * (2 * u - 1)) if kind == "log" else (0.9 + 0.2 * u)
    return th
