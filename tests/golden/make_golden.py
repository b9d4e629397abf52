#!/usr/bin/env python3
"""Extract the reference's own known-answer data into tests/golden/reference_goldens.json.

The reference (PETLION.jl) ships no test fixtures (test/runtests.jl is commented out);
the only known answers are the *executed outputs* stored inside examples/*.ipynb.
This script (run once, in the build container where /root/reference is mounted)
pulls them out so that the tests never need /root/reference at run time.

Sources (all under /root/reference/examples):
  updating_parameters.ipynb  cell 3  -> theta dict incl. I1C (full precision)
  updating_parameters.ipynb  cell 5  -> SVG polylines: every IDA step (t, V) of three
                                        1C discharges (eps_p = 0.385/0.485/0.585)
  model_inputs_and_outputs.ipynb     -> sol.V[1:13] and sol.c_e[1:5] (full precision)
                                        of simulate(p, I=2, SOC=0, V_max=4.1)
  CC-CV.ipynb cell 9                 -> SVG polyline of 2C CC + CV hold (older PETLION)
  getting_started.ipynb / CC-CV.ipynb-> printed run summaries (3-5 digits)
"""
import json, re, sys, os

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
EX = os.path.join(REF, "examples")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")


def cell_svg(nb, idx):
    for o in nb["cells"][idx].get("outputs", []):
        d = o.get("data", {})
        if "image/svg+xml" in d:
            return "".join(d["image/svg+xml"])
    raise KeyError(idx)


def polylines(svg):
    out = []
    for p in re.findall(r'<polyline[^>]*points="([^"]+)"', svg):
        pts = [tuple(float(v) for v in q.split(",")) for q in p.strip().split()]
        out.append(pts)
    return out


def decode_axes(pls):
    """x gridlines: vertical 2-point polylines spanning the plot; y gridlines: horizontal ones."""
    xs, ys = [], []
    spans = [p for p in pls if len(p) == 2]
    ymin = min(min(a[1], b[1]) for a, b in spans)
    ymax = max(max(a[1], b[1]) for a, b in spans)
    xmin = min(min(a[0], b[0]) for a, b in spans)
    xmax = max(max(a[0], b[0]) for a, b in spans)
    for a, b in spans:
        if a[0] == b[0] and abs(abs(a[1] - b[1]) - (ymax - ymin)) < 1e-6:
            xs.append(a[0])
        if a[1] == b[1] and abs(abs(a[0] - b[0]) - (xmax - xmin)) < 1e-6:
            ys.append(a[1])
    return sorted(set(xs)), sorted(set(ys)), (xmin, xmax, ymin, ymax)


gold = {"_source": "PETLION.jl v1.0.6 examples/*.ipynb executed outputs (Julia 1.7.1)"}

# ---- updating_parameters.ipynb ------------------------------------------------
nb = json.load(open(os.path.join(EX, "updating_parameters.ipynb")))
theta_txt = "".join(nb["cells"][3]["outputs"][0]["text"])
theta = {}
for line in theta_txt.strip().splitlines():
    k, v = line.split(": ")
    theta[k] = float(v)
gold["theta_LCO"] = theta

svg = cell_svg(nb, 5)
pls = polylines(svg)
xs, ys, box = decode_axes(pls)
# x gridlines: first is the axis edge (xmin); the remaining verticals are ticks at 0,1000,2000,3000 s
xticks = [x for x in xs if x != box[0]]
px_per_s = (xticks[1] - xticks[0]) / 1000.0
x0 = xticks[0]
# y gridlines at 3.0, 3.3, 3.6, 3.9 V (bottom -> top); the bottom axis edge is ymax
yticks = sorted([y for y in ys if y != box[3]], reverse=True)
v_per_px = 0.3 / (yticks[0] - yticks[1])
y_at_3V = yticks[0]
data = [p for p in pls if len(p) > 5]
ladders = {}
for name, p in zip(("0.385", "0.485", "0.585"), data):
    t = [(x - x0) / px_per_s for x, _ in p]
    V = [3.0 + (y_at_3V - y) * v_per_px for _, y in p]
    ladders[name] = {"t": t, "V": V, "n_points": len(p)}
gold["ladder_1C_discharge"] = {
    "note": "t decoded from SVG pixels (resolution ~1e-3 px = ~2 ms; ~1e-3 relative on big steps); "
            "V decoded assuming y ticks at 3.0/3.3/3.6/3.9 V (resolution ~1e-4 V, calibration ~3e-4 V)",
    "run": "p=petlion(LCO); p.opts.SOC=1; p.theta[eps_p]=X; simulate(p, I=-1)",
    "eps_p": ladders,
}

# ---- model_inputs_and_outputs.ipynb ------------------------------------------
nb = json.load(open(os.path.join(EX, "model_inputs_and_outputs.ipynb")))
for c in nb["cells"]:
    src = "".join(c.get("source", []))
    if src.strip() == "sol.V":
        txt = "".join(c["outputs"][0]["data"]["text/plain"])
        vals = [float(x) for x in re.findall(r"^\s+(\d\.\d+)\s*$", txt, flags=re.M)]
        gold["V_2C_charge"] = {"run": "simulate(p, I=2, SOC=0, V_max=4.1); simulate!(sol,p,V=:hold)",
                               "n_total": 121, "head13": vals[:13], "tail12": vals[13:]}
    if src.startswith("p.opts.outputs = (:t, :V, :c_e,)"):
        txt = "".join(c["outputs"][0]["data"]["text/plain"])
        rows = []
        for line in txt.splitlines():
            m = re.match(r"^\s*\[(.*)\]\s*$", line)
            if m:
                a, b = m.group(1).split("…")
                rows.append({"first10": [float(x) for x in a.split(",") if x.strip()],
                             "last10": [float(x) for x in b.split(",") if x.strip()]})
        gold["c_e_2C_charge_first5"] = rows

# ---- CC-CV.ipynb (older PETLION version: indicative only) --------------------
nb = json.load(open(os.path.join(EX, "CC-CV.ipynb")))
try:
    svg = cell_svg(nb, 9)
    pls = polylines(svg)
    xs, ys, box = decode_axes(pls)
    xticks = [x for x in xs if x != box[0]]
    px_per_s = (xticks[1] - xticks[0]) / 500.0
    # first tick is at t=0
    data = [p for p in pls if len(p) > 5]
    gold["ladder_CCCV_older_version"] = {
        "note": "x ticks every 500 s; executed with an older PETLION: indicative only",
        "t": [[(x - xticks[0]) / px_per_s for x, _ in p] for p in data],
    }
except Exception as e:  # pragma: no cover
    gold["ladder_CCCV_older_version"] = {"error": str(e)}

# ---- fast_charging_CC-CT-CV.ipynb (current PETLION wording; temperature=true) -------------
nb = json.load(open(os.path.join(EX, "fast_charging_CC-CT-CV.ipynb")))
svg = cell_svg(nb, 17)
pls = polylines(svg)
xs, ys, box = decode_axes(pls)
xticks = [x for x in xs if x != box[0]]          # ticks at 0, 500, 1000, 1500 s
px_per_s = (xticks[-1] - xticks[0]) / 1500.0
data = [p for p in pls if len(p) > 5][0]
tt = [(x - xticks[0]) / px_per_s for x, _ in data]
# segment 1 = simulate(p, I=4) until T_max: its last point is the interpolated end point, repeated
# as the first point of the next run
n1 = next(i for i in range(1, len(tt)) if tt[i] <= tt[i - 1])   # the duplicated hand-over point
gold["ladder_4C_thermal"] = {
    "run": "p=petlion(LCO; temperature=true); SOC=0; T_max=313.15; V_max=4.1; simulate(p, I=4)",
    "note": "t decoded from SVG pixels (~1 ms resolution); V as raw y pixels (linear in V)",
    "t": tt[:n1], "y_px": [y for _, y in data[:n1]], "n_points_all_runs": len(data),
}

# current trace of the same three runs (cell 18): calibrated with the two printed currents (4C on the first
# run, 0.1959C at the very end).  The dT=:hold run starts at the duplicated hand-over point.
svgI = cell_svg(nb, 18)
dataI = [p for p in polylines(svgI) if len(p) > 5][0]
yI = [y for _, y in dataI]
aI = (4.0 - 0.1959) / (yI[0] - yI[-1])
Iref = [4.0 + (y - yI[0]) * aI for y in yI]
n2 = next(i for i in range(n1 + 1, len(tt)) if tt[i] <= tt[i - 1] and tt[i] > 600.0)   # hand-over to V=:hold
gold["trace_CT_hold"] = {
    "run": "simulate!(sol, p, dT=:hold) after the 4C run above (fast_charging_CC-CT-CV.ipynb cell 11)",
    "note": "I in C-rate decoded from the SVG of cell 18 (resolution ~2e-4 C); t from cell 17 (~1 ms)",
    "t": tt[n1:n2], "I": Iref[n1:n2],
}

# ---- printed summaries --------------------------------------------------------
gold["summaries"] = {
    "1C_discharge": {"t_s": 3600.0, "V": 2.9357, "P": -85.8094, "SOC": -0.0, "exit": "SOC_min",
                     "src": "getting_started.ipynb:100-109 (older version)"},
    "2C_CC_to_4.1V": {"t_s": 1388.68, "V": 4.1, "P": 239.6861, "SOC": 0.7715, "exit": "V_max",
                      "src": "CC-CV.ipynb:66-75 (older version)"},
    "CV_hold": {"t_s": 2440.61, "I_C": 0.1955, "P": 23.432, "SOC": 1.0001, "exit": "SOC_max",
                "src": "CC-CV.ipynb:103-112 (older version)"},
    "thermal_4C_to_Tmax": {"t_s": 357.56, "V": 4.0312, "P": 471.33, "SOC": 0.3973, "T_C": 40.0,
                           "exit": "T_max", "src": "fast_charging_CC-CT-CV.ipynb cell 7 (current version)"},
    "thermal_CT_hold": {"t_s": 686.41, "I_C": 2.7892, "V": 4.1, "P": 334.26, "SOC": 0.6714, "T_C": 40.0,
                        "exit": "V_max", "src": "fast_charging_CC-CT-CV.ipynb cell 11 (current version)"},
    "thermal_CV_after_CT": {"t_s": 1865.61, "I_C": 0.1959, "P": 23.47, "SOC": 1.0, "T_C": 25.6963,
                            "exit": "SOC_max", "src": "fast_charging_CC-CT-CV.ipynb cell 13 (after a dT=:hold run)"},
    "benchmark_median_ms": 2.616,
}

# ---- function inputs: printed summaries of examples/variable_input_functions.ipynb ----------------------
import re
nbf = json.load(open(os.path.join(REF, "examples", "variable_input_functions.ipynb")))


def printed(cell):
    txt = ""
    for o in nbf["cells"][cell].get("outputs", []):
        t = o.get("text") or o.get("data", {}).get("text/plain") or []
        txt += "".join(t)
    f = lambda key: float(re.search(key + r":\s+([-0-9.]+)", txt).group(1))
    return {"t_s": f("Time"), "V": f("Voltage"), "P": f("Power"), "SOC": f("SOC")}


fi = {}
for name, cell, extra in (("step", 6, {"tdiscon": []}), ("step_tdiscon", 8, {"tdiscon": [100.0]}),
                          ("ramp_100", 13, {"ramp_val": 1 / 100}), ("ramp_10", 15, {"ramp_val": 1 / 10})):
    src = "".join(nbf["cells"][cell]["source"])
    assert ("I_fun1" in src) == name.startswith("step") and ("tdiscon" in src) == (name == "step_tdiscon"), (name, src)
    fi[name] = dict(printed(cell), **extra, src="variable_input_functions.ipynb cell %d: %s" % (cell, src.split("\n")[-1 if "ramp" not in name else 1].strip()))
for name, cell in (("ramp_100", 13), ("ramp_10", 15)):
    data = [q for q in polylines(cell_svg(nbf, cell)) if len(q) > 2][0]
    x0, x1 = data[0][0], data[-1][0]
    fi[name]["t_ladder"] = [(x - x0) / (x1 - x0) * 100.0 for x, _ in data]    # first vertex t = 0, last t = tf = 100
fi["_note"] = ("I_fun1(t) = t < 100 ? 1 : 0.5 ; I_ramp(t,p) = ramp_val*t ; SOC = 0, LCO defaults; the notebook was executed "
               "with an older PETLION/Sundials (power printed in W): 4-digit checks")
gold["function_inputs"] = fi

json.dump(gold, open(OUT, "w"), indent=1, ensure_ascii=False)
print("wrote", OUT)
for k, v in ladders.items():
    import numpy as np
    dt = np.diff(v["t"])
    print(k, v["n_points"], "t_end", v["t"][-1], "V0", v["V"][0], "V_end", v["V"][-1], "h0", dt[0])
