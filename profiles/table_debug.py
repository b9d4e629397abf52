import numpy as np, sys
sys.path.insert(0, ".")
import oracle as O, petlion_b200 as P
lco = P.petlion("LCO")
STEP = ([0.0, 100.0, 100.0], [1.0, 1.0, 0.5])
m = O.make_model("LCO")
for td in ([], [100.0]):
    sol = P.simulate(lco, 200, I=P.Table(*STEP), SOC=0, tdiscon=td)
    ref = O.simulate_batch(m, O.theta_defaults("LCO"), O.make_run("I", tf=200, table=STEP, tdiscon=td),
                           O.default_opts(), O.default_bounds("LCO"), SOC0=0.0, n_save_max=512)
    s = sol.results[-1].summary
    print("tdiscon", td, {k: int(s[k][0]) for k in ("n_steps", "n_res", "n_jac", "n_netf", "n_ncfn", "n_reinit")},
          {k: int(ref[k][0]) for k in ("n_steps", "n_res", "n_jac", "n_netf", "n_ncfn", "n_reinit")})
    n = min(sol.n_points[0], ref["traj_n"][0])
    tg, to = sol.t[0, :n], ref["traj"]["t"][0, :n]
    bad = np.where(np.abs(tg - to) > 1e-9 * np.maximum(1, to))[0]
    print("points", sol.n_points[0], ref["traj_n"][0], "first diff at", bad[:3])
    k = bad[0] if len(bad) else n - 3
    for i in range(max(0, k - 4), min(n, k + 6)):
        print(i, "%.12f %.12f" % (tg[i], to[i]), sol.I[0, i], ref["traj"]["I"][0, i], "%.9f %.9f" % (sol.V[0, i], ref["traj"]["V"][0, i]))
