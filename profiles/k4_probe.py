"""Times the fused integrator (device-resident buffers, CUDA events) for the library named by PLB_LIB:
quick A/B of build variants.  usage: PLB_LIB=... python profiles/k4_probe.py [B] [iso|thermal|sei|wide|wsei|wth|thsei|wthsei|mhc|lgm|lgmth][_r12|_r14|_sp]
(one segment: a 1C discharge for iso / wide / mhc, a 4C charge to 4.1 V for the thermal families, a 1C charge to 4.2 V for
the SEI families)"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import petlion_b200 as P
from petlion_b200 import _lib
from petlion_b200 import sweep  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
fam = sys.argv[2] if len(sys.argv) > 2 else "iso"
sib = {}                       # sibling builds: N_r = 12 / 14, Fickian_method = :spectral
if fam.endswith("_r12") or fam.endswith("_r14"):
    sib = dict(N_r_p=int(fam[-2:]), N_r_n=int(fam[-2:])); fam = fam[:-4]
elif fam.endswith("_sp"):
    sib = dict(Fickian_method="spectral"); fam = fam[:-3]
L = _lib.lib()
grid = dict(N_p=20, N_s=20, N_n=20) if fam in ("wide", "wsei", "wth", "wthsei") else {}
thermal = fam in ("thermal", "wth", "thsei", "wthsei", "lgmth")
aging = fam in ("sei", "wsei", "thsei", "wthsei")
rx = dict(rxn_p="rxn_MHC", rxn_n="rxn_MHC") if fam == "mhc" else {}
p = P.petlion("NMC_LGM50" if fam.startswith("lgm") else "LCO", temperature=thermal, aging="SEI" if aging else False, **grid, **rx, **sib)

h = p._h; N = p.N.tot
dev = torch.device("cuda", 0); f64 = dict(dtype=torch.float64, device=dev)
th = sweep.randomised_theta(p, B)
if fam == "mhc":      # a physical reorganisation energy (in kT) instead of the reference's 6.26e-20
    th[:, p.θ_keys.index("λ_MHC_p")] = 15.0; th[:, p.θ_keys.index("λ_MHC_n")] = 12.0
d_theta = torch.from_numpy(th).to(dev)
cur, soc = (4.0, 0.0) if thermal else ((1.0, 0.0) if aging else (-1.0, 1.0))
d_soc0 = torch.full((B,), soc, **f64)
d_Y = torch.zeros(B, N, **f64); d_YP = torch.zeros(B, N, **f64); d_SOC = torch.zeros(B, **f64); d_t = torch.zeros(B, **f64)
d_sum = torch.zeros(B, 10, **f64); d_trn = torch.zeros(B, dtype=torch.int32, device=dev)
o = _lib.Opts(); L.plb_opts_defaults(h, C.byref(o)); b = _lib.Bounds(); L.plb_bounds_defaults(h, C.byref(b))
if thermal or aging:
    b.V_max = 4.1 if thermal else 4.2
L.plb_set_stream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
run = _lib.Run(0, 0, cur, 1e6, 1, 0)
ms = []
for k in range(4):
    _lib.check(L.plb_simulate(h, B, d_theta.data_ptr(), C.byref(run), None, C.byref(o), C.byref(b), d_soc0.data_ptr(),
                              d_Y.data_ptr(), d_YP.data_ptr(), d_SOC.data_ptr(), d_t.data_ptr(), d_sum.data_ptr(), 0,
                              None, None, None, None, None, None, d_trn.data_ptr(), 1))
    ms.append(L.plb_last_kernel_ms(h))
s = d_sum.cpu().numpy().view(_lib.SUMMARY_DTYPE).reshape(-1)
print(os.path.basename(os.environ.get("PLB_LIB", "default")), fam, sib, "B", B, "ms", [round(x, 1) for x in ms], "sims/s", round(B / (min(ms[1:]) * 1e-3)),
      "steps", float(np.mean(s["n_steps"])), "chk", float(np.sum(s["V_end"][s["flag"] >= 0])), flush=True)
