"""Small driver for ncu captures: K1 (k_resjac) over PROF_B1 systems and K4 (k_simulate) over PROF_B4.
PROF_THERMAL=1 profiles the temperature=true family (4C charge from SOC 0) instead of the 1C discharge."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import petlion_b200 as P  # noqa: E402
from petlion_b200 import _lib  # noqa: E402
from petlion_b200 import sweep  # noqa: E402

B1 = int(os.environ.get("PROF_B1", 65536))
B4 = int(os.environ.get("PROF_B4", 8192))
L = _lib.lib()
FAM = os.environ.get("PROF_FAMILY", "thermal" if int(os.environ.get("PROF_THERMAL", "0")) else "iso")   # iso|thermal|sei|wsei
TH = FAM == "thermal"
p = P.petlion("LCO", temperature=TH, aging="SEI" if FAM in ("sei", "wsei") else False,
              **(dict(N_p=20, N_s=20, N_n=20) if FAM == "wsei" else {}))
h = p._h
N = p.N.tot
dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)
th = sweep.randomised_theta(p, B1)
d_theta = torch.from_numpy(th).to(dev)
d_soc0 = torch.full((B1,), 1.0 if FAM == "iso" else 0.0, **f64)
CUR = 4.0 if TH else (1.0 if FAM in ("sei", "wsei") else -1.0)
T_MID = 150.0 if TH else 1800.0
d_Y = torch.zeros(B1, N, **f64); d_YP = torch.zeros(B1, N, **f64)
d_SOC = torch.zeros(B1, **f64); d_t = torch.zeros(B1, **f64); d_sum = torch.zeros(B1, 10, **f64)
d_trn = torch.zeros(B1, dtype=torch.int32, device=dev)
o = _lib.Opts(); L.plb_opts_defaults(h, C.byref(o))
b = _lib.Bounds(); L.plb_bounds_defaults(h, C.byref(b))
if FAM != "iso":
    b.V_max = 4.1 if TH else 4.2
L.plb_set_stream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream))


def sim(B, tf):
    run = _lib.Run(0, 0, CUR, tf, 1, 0)
    _lib.check(L.plb_simulate(h, B, d_theta.data_ptr(), C.byref(run), None, C.byref(o), C.byref(b), d_soc0.data_ptr(),
                              d_Y.data_ptr(), d_YP.data_ptr(), d_SOC.data_ptr(), d_t.data_ptr(), d_sum.data_ptr(), 0,
                              None, None, None, None, None, None, d_trn.data_ptr(), 1))


sim(B1, T_MID)                       # launch 1: mid-run states for K1
nnz = L.plb_jac_nnz(h, 0)
d_res = torch.empty(B1, N, **f64); d_nz = torch.empty(B1, nnz, **f64); d_gam = torch.full((B1,), 0.05, **f64)
runI = _lib.Run(0, 0, CUR, 1e6, 1, 0)
for _ in range(3):                   # launches 2-4: K1
    _lib.check(L.plb_resjac(h, B1, d_Y.data_ptr(), d_YP.data_ptr(), d_gam.data_ptr(), d_theta.data_ptr(), C.byref(runI),
                            None, d_res.data_ptr(), d_nz.data_ptr(), 1))
    print("k1 ms", L.plb_last_kernel_ms(h))
sim(B4, 1e6)                         # launch 5: full discharge over B4 systems
print("k4 ms", L.plb_last_kernel_ms(h), "B4", B4)
