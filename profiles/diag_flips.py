"""diagnostic: where do the GPU and the oracle first take a different step?  (time of the first differing step,
counters of the run up to there)"""
import numpy as np, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import petlion_b200 as P
import oracle as O
from tests import util
B = 4096
p = P.petlion("LCO")
tho = util.oracle_theta_batch(B)
util.set_theta_batch(p, util.product_theta_from_oracle(p, tho))
sol = P.simulate(p, I=-1, SOC=1, n_save_max=300)
ref = O.simulate_batch(O.make_model("LCO"), tho, O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"), SOC0=1.0, n_save_max=300, nthreads=16)
s = sol.results[-1].summary
first_t, first_k, frac = [], [], []
for i in range(B):
    n = min(sol.n_points[i], ref["traj_n"][i])
    a, b = sol.t[i, :n], ref["traj"]["t"][i, :n]
    d = np.abs(a - b) > 1e-9 * np.maximum(np.abs(b), 1e-3)
    if d.any() or sol.n_points[i] != ref["traj_n"][i]:
        k = int(np.argmax(d)) if d.any() else n
        first_k.append(k); first_t.append(b[min(k, n - 1)]); frac.append(b[min(k, n - 1)] / ref["t_end"][i])
first_t = np.array(first_t); frac = np.array(frac)
print("flipped", len(first_t), "of", B, "=", len(first_t) / B)
print("time of first differing step: quantiles", np.quantile(first_t, [0, 0.1, 0.25, 0.5, 0.75, 0.9, 1.0]).round(1))
print("as a fraction of the run:", np.quantile(frac, [0, 0.1, 0.25, 0.5, 0.75, 0.9, 1.0]).round(3))
print("step index:", np.quantile(first_k, [0, 0.25, 0.5, 0.75, 1.0]))
fl = np.array([sol.n_points[i] != ref["traj_n"][i] or np.any(np.abs(sol.t[i, :sol.n_points[i]] - ref["traj"]["t"][i, :sol.n_points[i]]) > 1e-9 * 3600) for i in range(B)])
print("failures among flipped: ncfn>0 gpu", np.mean(s["n_ncfn"][fl] > 0), "cpu", np.mean(ref["n_ncfn"][fl] > 0), " among identical:", np.mean(s["n_ncfn"][~fl] > 0))
print("netf>0 among flipped", np.mean(ref["n_netf"][fl] > 0), "identical", np.mean(ref["n_netf"][~fl] > 0), "mean netf flipped", ref["n_netf"][fl].mean(), "identical", ref["n_netf"][~fl].mean())
# V difference at the last common identical step
