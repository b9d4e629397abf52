"""Times K1 (k_resjac) standalone over a batch of mid-run states for the library named by PLB_LIB."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import petlion_b200 as P
from petlion_b200 import _lib
from petlion_b200 import sweep  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
fam = sys.argv[2] if len(sys.argv) > 2 else "iso"
L = _lib.lib()
grid = dict(N_p=20, N_s=20, N_n=20) if fam == "wsei" else {}
p = P.petlion("LCO", temperature=fam == "thermal", aging="SEI" if fam in ("sei", "wsei") else False, **grid)
h = p._h; N = p.N.tot; nth = len(p.θ_keys)
dev = torch.device("cuda", 0); f64 = dict(dtype=torch.float64, device=dev)
th = sweep.randomised_theta(p, B)
d_theta = torch.from_numpy(th).to(dev)
cur, soc, tmid = (4.0, 0.0, 150.0) if fam == "thermal" else ((1.0, 0.0, 1800.0) if fam in ("sei", "wsei") else (-1.0, 1.0, 1800.0))
d_soc0 = torch.full((B,), soc, **f64)
d_Y = torch.zeros(B, N, **f64); d_YP = torch.zeros(B, N, **f64); d_SOC = torch.zeros(B, **f64); d_t = torch.zeros(B, **f64)
d_sum = torch.zeros(B, 10, **f64); d_trn = torch.zeros(B, dtype=torch.int32, device=dev)
o = _lib.Opts(); L.plb_opts_defaults(h, C.byref(o)); b = _lib.Bounds(); L.plb_bounds_defaults(h, C.byref(b))
L.plb_set_stream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
run = _lib.Run(0, 0, cur, tmid, 1, 0)
_lib.check(L.plb_simulate(h, B, d_theta.data_ptr(), C.byref(run), None, C.byref(o), C.byref(b), d_soc0.data_ptr(), d_Y.data_ptr(),
                          d_YP.data_ptr(), d_SOC.data_ptr(), d_t.data_ptr(), d_sum.data_ptr(), 0, None, None, None, None, None, None, d_trn.data_ptr(), 1))
nnz = L.plb_jac_nnz(h, 0)
d_res = torch.empty(B, N, **f64); d_nz = torch.empty(B, nnz, **f64); d_gam = torch.full((B,), 0.05, **f64)
flush = torch.empty(256 * 1024 * 1024 // 8, **f64)
runI = _lib.Run(0, 0, cur, 1e6, 1, 0)
ms = []
for k in range(8):
    flush.zero_(); torch.cuda.synchronize()
    _lib.check(L.plb_resjac(h, B, d_Y.data_ptr(), d_YP.data_ptr(), d_gam.data_ptr(), d_theta.data_ptr(), C.byref(runI), None,
                            d_res.data_ptr(), d_nz.data_ptr(), 1))
    ms.append(L.plb_last_kernel_ms(h))
bytes_eval = 8 * (3 * N + nth + nnz) + 16
t = float(np.mean(ms[3:]))
print(os.path.basename(os.environ.get("PLB_LIB", "default")), fam, "B", B, "ms", round(t, 4), "GB/s", round(B * bytes_eval / t / 1e6), "frac", round(B * bytes_eval / t / 1e6 / 6551, 4),
      "chk", float(d_nz.sum().item()), float(d_res.abs().sum().item()), flush=True)
