"""Where does the end-to-end (host buffers) time of plb_simulate go?  Times the call with pinned and
pageable host buffers, with and without trajectories, next to the kernel time measured by CUDA events."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import petlion_b200 as P
from petlion_b200 import _lib
from petlion_b200 import sweep
L = _lib.lib()
p = P.petlion("LCO"); h = p._h
B = int(os.environ.get("B", 65536)); N = p.N.tot
th = sweep.randomised_theta(p, B)
o = _lib.Opts(); L.plb_opts_defaults(h, C.byref(o))
b = _lib.Bounds(); L.plb_bounds_defaults(h, C.byref(b))
run = _lib.Run(0, 0, -1.0, 1e6, 1, 0)
def bufs(pin, ns):
    mk = (lambda *s, dt=torch.float64: torch.zeros(*s, dtype=dt).pin_memory()) if pin else (lambda *s, dt=torch.float64: torch.zeros(*s, dtype=dt))
    d = dict(theta=torch.from_numpy(th).pin_memory() if pin else torch.from_numpy(th).clone(), soc=mk(B) + 1.0, Y=mk(B, N), SOC=mk(B), t=mk(B), sum=mk(B, 10),
             trt=mk(B, max(ns, 1)), trV=mk(B, max(ns, 1)), trn=mk(B, dt=torch.int32))
    if pin:
        d["soc"] = d["soc"].pin_memory()
    return d
for pin in (True, False):
    for ns in (0, 128):
        d = bufs(pin, ns)
        def call():
            _lib.check(L.plb_simulate(h, B, d["theta"].data_ptr(), C.byref(run), None, C.byref(o), C.byref(b), d["soc"].data_ptr(),
                                      d["Y"].data_ptr(), None, d["SOC"].data_ptr(), d["t"].data_ptr(), d["sum"].data_ptr(), ns,
                                      d["trt"].data_ptr() if ns else None, d["trV"].data_ptr() if ns else None, None, None, None, None,
                                      d["trn"].data_ptr(), 0))
        call(); torch.cuda.synchronize()
        t0 = time.perf_counter(); call(); call(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
        print(f"pinned={pin} n_save={ns}: call {dt*1e3:.1f} ms, kernel {L.plb_last_kernel_ms(h):.1f} ms, sims/s {B/dt:.0f}", flush=True)
