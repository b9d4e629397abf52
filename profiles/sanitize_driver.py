"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck): plain and extended integrator,
K1, Newton init, linear solve.   compute-sanitizer --tool memcheck python profiles/sanitize_driver.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import petlion_b200 as P

fams = os.environ.get("SAN_FAMILIES", "iso,thermal,sei,wide,wsei,wth,thsei,wthsei,mhc,lgm,iso12,th14,sei14,isosp,thsp,seisp,widemhc,thseimhc,wthlgm").split(",")
for fam in fams:
    G = dict(N_p=20, N_s=20, N_n=20)
    kw = dict(iso={}, thermal=dict(temperature=True), sei=dict(aging="SEI"), wide=G, wsei=dict(aging="SEI", **G),
              wth=dict(temperature=True, **G), thsei=dict(temperature=True, aging="SEI"),
              wthsei=dict(temperature=True, aging="SEI", **G), mhc=dict(rxn_p="rxn_MHC", rxn_n="rxn_MHC"), lgm=dict(temperature=True),
              # N_r = 12 / 14 and Fickian_method = :spectral sibling builds
              iso12=dict(N_r_p=12, N_r_n=12), th14=dict(temperature=True, N_r_p=14, N_r_n=14), sei14=dict(aging="SEI", N_r_p=14, N_r_n=14),
              widemhc=dict(rxn_p="rxn_MHC", rxn_n="rxn_MHC", **G), thseimhc=dict(temperature=True, aging="SEI", rxn_p="rxn_MHC", rxn_n="rxn_MHC"),
              wthlgm=dict(temperature=True, **G),
              isosp=dict(Fickian_method="spectral"), thsp=dict(temperature=True, Fickian_method="spectral"),
              seisp=dict(aging="SEI", Fickian_method="spectral"))[fam]
    p = P.petlion("NMC_LGM50" if fam in ("lgm", "wthlgm") else "LCO", **kw)
    B = 5
    p.θ["D_sp"] = np.asarray(p.θ["D_sp"]) * np.linspace(0.8, 1.2, B)
    sol = P.simulate(p, 200, I=1, SOC=0.1)                                             # plain kernel
    P.simulate_(sol, p, 100, V="hold")
    tab = P.Table([0.0, 30.0, 30.0, 60.0], [1.0, 1.0, 2.0, 0.5])
    sol2 = P.simulate(p, 60, I=tab, SOC=0.2, outputs="all", tstops=[10.0])              # extended kernel
    sol3 = P.simulate(p, np.array([0.0, 7.0, 33.0, 150.0, 1e6]), I=1, SOC=0.1, V_max=3.9)  # dense output rows
    Y0 = p.initial_guess(np.full(B, 0.5))
    st, Y, YP = p.newton_init(Y0, method="I", value=1.0)
    res, nz = p.resjac(Y, YP, np.full(B, 0.1), method="I", value=1.0)                  # K1 (TMA-staged where built)
    os.environ["PLB_K1_NO_TMA"] = "1"
    p_old = P.petlion("NMC_LGM50" if fam in ("lgm", "wthlgm") else "LCO", **kw); p_old.θ["D_sp"] = p.θ["D_sp"]
    res_b, nz_b = p_old.resjac(Y, YP, np.full(B, 0.1), method="I", value=1.0)           # K1, per-lane loads
    del os.environ["PLB_K1_NO_TMA"]
    assert np.array_equal(res, res_b) and np.array_equal(nz, nz_b)
    x, ok = p.linear_solve(Y, YP, np.full(B, 0.1), res, method="I", value=1.0)
    if fam in ("iso", "wide"):            # the concentration-rate inputs run in a sibling build of these two families
        P.simulate_(sol, p, 50, dc_s_n_max="hold")
        P.simulate_(sol, p, 50, dc_e_min=-0.01)
        assert (sol.results[-1].summary["flag"] >= 0).all()
    print(fam, "ok", sol.results[-1].summary["flag"], sol2.results[-1].summary["n_steps"], int(np.isfinite(x).all()), flush=True)
