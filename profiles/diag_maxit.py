"""diagnostic: the thermal system that exhausts maxiters in the CV phase at tight tolerance"""
import numpy as np, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import petlion_b200 as P
import oracle as O
from tests import util
tol = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-7
W = util.PROTOCOLS["cfg3i"]
p = P.petlion("LCO", temperature=True)
B = 512
tho = util.oracle_theta_batch(B, first=80000)
th = util.product_theta_from_oracle(p, tho)
util.set_theta_batch(p, th)
sol, _ = util.gpu_protocol(P, p, W, reltol=tol, abstol=tol, n_save_max=0)
s = sol.results[1].summary
bad = np.where(s["flag"] < 0)[0]
print("bad", bad, s[bad])
for i in bad[:1]:
    p1 = P.petlion("LCO", temperature=True); util.set_theta_batch(p1, th[[i, i]])
    sol1, _ = util.gpu_protocol(P, p1, W, reltol=tol, abstol=tol, n_save_max=60000, maxiters=200000)
    q = sol1.results[1].summary[:1]
    n0 = sol1.results[0].n_rows[0]; n = sol1.n_points[0]
    t = sol1.t[0, n0:n]; I = sol1.I[0, n0:n]
    dt = np.diff(t)
    print("alone:", q, "rows", n - n0)
    print("t head", t[:12], "\ndt quantiles", np.quantile(dt, [0, 0.1, 0.5, 0.9, 1.0]), "\nt tail", t[-6:], "\nI tail", I[-6:], "T", sol1.T[0, n - 3:n])
    k = np.argmax(dt < 1e-6) if (dt < 1e-6).any() else -1
    print("first tiny step at row", k, t[max(k - 3, 0):k + 5] if k >= 0 else None)
    o = O.default_opts(reltol=tol, abstol=tol, reltol_init=tol, abstol_init=tol)
    rs = util.oracle_protocol(W, tho[i:i + 1], o, nthreads=1)
    print("oracle", rs[1]["flag"], rs[1]["n_steps"], rs[1]["n_netf"], rs[1]["n_ncfn"], rs[1]["t_end"])
