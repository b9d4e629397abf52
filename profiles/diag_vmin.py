"""diagnostic: the systems of the full-size batch whose V_min exit is not exactly on the bound"""
import numpy as np, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import petlion_b200 as P
import oracle as O
from tests import util
B = 65536
p = P.petlion("LCO")
tho = util.oracle_theta_batch(B)
th = util.product_theta_from_oracle(p, tho)
util.set_theta_batch(p, th)
s = P.simulate(p, I=-1, SOC=1, n_save_max=0).results[-1].summary
bad = np.where((s["flag"] == 1) & (np.abs(s["V_end"] - 2.5) > 1e-9))[0]
print("offenders", bad, s["V_end"][bad] - 2.5)
for i in bad[:3]:
    p1 = P.petlion("LCO"); util.set_theta_batch(p1, th[i:i + 1])
    sol = P.simulate(p1, I=-1, SOC=1, n_save_max=400)
    q = sol.results[-1].summary
    n = sol.n_points[0]
    print(i, q, "\n t", sol.t[0, n - 5:n], "\n V", sol.V[0, n - 5:n])
    r = O.simulate_batch(O.make_model("LCO"), tho[i:i + 1], O.make_run("I", -1.0), O.default_opts(), O.default_bounds("LCO"), SOC0=1.0, n_save_max=400)
    m = r["traj_n"][0]
    print("oracle", r["flag"], r["n_steps"], r["V_end"] - 2.5, r["n_netf"], r["n_ncfn"], "\n t", r["traj"]["t"][0, m - 5:m], "\n V", r["traj"]["V"][0, m - 5:m])
