"""diagnostic for tests/test_gpu_tight.py: where do the two sides differ most, per segment?"""
import numpy as np, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import petlion_b200 as P
import oracle as O
from tests import util
from tests.test_gpu_tight import _run_both
name, B, tol, t1, dt = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5])
td = np.arange(0.0, t1, dt)
sol, dense, ref = _run_both(P, name, B, tol, td, first=int(os.environ.get("DIAG_FIRST", 80000)), maxiters=400000)
for k, r in enumerate(ref):
    s = sol.results[k].summary
    g, o = dense[k], r["dense"]
    mism = (np.isnan(g["V"]) != np.isnan(o["V"])).sum(axis=1)
    both = ~np.isnan(g["V"]) & ~np.isnan(o["V"])
    err = np.where(both, np.abs(g["V"] - o["V"]) / np.abs(o["V"]), 0.0)
    i = int(np.argmax(err.max(axis=1))); j = int(np.argmax(err[i]))
    dte = np.abs(s["t_end"] - r["t_end"])
    print(f"seg {k}: flags gpu {np.unique(s['flag'], return_counts=True)} cpu {np.unique(r['flag'], return_counts=True)} flagdiff {(s['flag'] != r['flag']).sum()}"
          f" max row mismatch {mism.max()} (sys {int(np.argmax(mism))}) max dt_end {dte.max():.3e} (sys {int(np.argmax(dte))})")
    print(f"   worst V: sys {i} t {td[j]} gpu {g['V'][i, j]:.9f} cpu {o['V'][i, j]:.9f} rel {err[i, j]:.2e}; t_end gpu {s['t_end'][i]:.6f} cpu {r['t_end'][i]:.6f}"
          f" flags {s['flag'][i]} {r['flag'][i]} steps {s['n_steps'][i]} {r['n_steps'][i]}")
    for q in (int(np.argmax(mism)),):
        print("   mismatch rows sys", q, "gpu filled", td[~np.isnan(g["V"][q])][[0, -1]] if (~np.isnan(g["V"][q])).any() else None,
              "cpu filled", td[~np.isnan(o["V"][q])][[0, -1]] if (~np.isnan(o["V"][q])).any() else None, "t_end", s["t_end"][q], r["t_end"][q], "flags", s["flag"][q], r["flag"][q])
    errI = np.where(both, np.abs(g["I"] - o["I"]) / np.maximum(np.abs(o["I"]), 1e-3), 0.0)
    order = np.argsort(-errI.max(axis=1))[:5]
    for q in order:
        jj = int(np.argmax(errI[q]))
        print(f"   worst I: sys {q} t {td[jj]} gpu {g['I'][q, jj]:.8f} cpu {o['I'][q, jj]:.8f} rel {errI[q, jj]:.2e} steps {s['n_steps'][q]} {r['n_steps'][q]} t_end {s['t_end'][q]:.4f} {r['t_end'][q]:.4f} flags {s['flag'][q]} {r['flag'][q]} t_start {sol.results[k-1].summary['t_end'][q] if k else 0:.4f}")
    for q in np.where(s["flag"] != r["flag"])[0][:4]:
        print(f"   FLAGDIFF seg {k} sys {q}: flags {s['flag'][q]} {r['flag'][q]} t_end {s['t_end'][q]:.6f} {r['t_end'][q]:.6f} V_end {s['V_end'][q]:.9f} {r['V_end'][q]:.9f} steps {s['n_steps'][q]} {r['n_steps'][q]}")
    if k > 0 and err.max() > 1e-5:
        # hold / later segments: first rows of the worst system
        fill = np.where(both[i])[0][:6]
        print("   first rows", td[fill], "\n   gpu", g["V"][i, fill], "\n   cpu", o["V"][i, fill], "\n   I gpu", g["I"][i, fill], "\n   I cpu", o["I"][i, fill])
