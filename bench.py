#!/usr/bin/env python3
"""bench.py -- headline benchmark of BASELINE.json: full-discharge sims/sec (LCO 301-DAE, FP64) and
the HBM roofline fraction of the residual+Jacobian kernel.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the CPU oracle port on all host cores)

One "step" = one pass of the hot path over one batch of synthetic input: B = 65 536 independent
1C CC discharges (LCO, N=(10,10,10), N_r=10, isothermal, SOC 1 -> SOC_min/V_min) with randomised
{D_s, k, eps} (BASELINE.json configs[1]), per GPU (weak scaling: simulations are independent,
ranks share nothing), followed by the one collective of this path: an NCCL all-gather of the
80-byte per-system summaries, INSIDE the timed region.

The other configs of BASELINE.json are measured by the same code (`--workload`), and one of them
rides along in the default run as `extra.configs` at the GPU count it is named for: configs[2]
(thermal CC-CV, 262 144 systems) at --gpus 8, configs[3] (NMC GITT, 131 072) at --gpus 4,
configs[4] (SEI, N=(20,20,20)) at --gpus 1.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 65536
N_SAVE_E2E = 128

# Workloads = BASELINE.json configs.  A protocol is a list of segments (method, input_kind, value, tf, bound
# overrides); segment 0 is simulate(), the rest simulate!().
WORKLOADS = {
    "cfg2": dict(name="configs[1]: batch=65536 LCO 1C CC discharges, randomised {D_s,k,eps}, N=(10,10,10), N_r=10, isothermal",
                 metric="full-discharge sims/sec (LCO 301-DAE, FP64)", cathode="LCO", temperature=False, batch=65536,
                 soc0=1.0, protocol=[("I", 0, -1.0, 1e6, {})], t_mid=1800.0, dense=(0.0, 3700.0, 30.0)),
    "cfg3": dict(name="configs[2]: LCO CC-CV (4C -> 4.1 V, then V=:hold to SOC_max) with temperature=true (351 DAEs), "
                      "randomised {D_s,k,eps}; 262144 systems over 8 GPUs = 32768 per GPU",
                 metric="CC-CV protocol sims/sec (LCO thermal 351-DAE, FP64)", cathode="LCO", temperature=True,
                 batch=32768, soc0=0.0, t_mid=150.0, dense=(0.0, 4000.0, 20.0),
                 protocol=[("I", 0, 4.0, 1e6, {"V_max": 4.1}), ("V", 1, 0.0, 1e6, {"V_max": 4.1})]),
    "cfg4": dict(name="configs[3]: NMC GITT, SOC0=0, 20 x {I=+1C for 180 s ; I=:rest for 7200 s} via simulate!, "
                      "randomised {D_s,k,eps}; 131072 systems over 4 GPUs = 32768 per GPU",
                 metric="GITT protocol sims/sec (NMC 301-DAE, FP64)", cathode="NMC", temperature=False,
                 batch=32768, soc0=0.0, t_mid=100.0, dense=(0.0, 147600.0, 180.0),
                 protocol=[seg for _ in range(20) for seg in (("I", 0, 1.0, 180.0, {}), ("I", 2, 0.0, 7200.0, {}))]),
    "cfg5": dict(name="configs[4]: batch=65536 LCO with aging=:SEI on the refined grid N=(20,20,20) (642 DAEs, two warps per "
                      "system), randomised {D_s,k,eps}; 1C charge to 4.2 V (side reaction active) then 1C discharge via simulate!",
                 metric="charge+discharge sims/sec (LCO SEI 642-DAE, FP64)", cathode="LCO", temperature=False, aging=True,
                 grid=dict(N_p=20, N_s=20, N_n=20), batch=65536, soc0=0.0, t_mid=1800.0, dense=(0.0, 7400.0, 60.0),
                 protocol=[("I", 0, 1.0, 1e6, {"V_max": 4.2}), ("I", 0, -1.0, 1e6, {"V_max": 4.2})]),
    "cfg5n10": dict(name="configs[4] physics on the N=(10,10,10) grid (322 DAEs): batch=65536 LCO with aging=:SEI, randomised "
                         "{D_s,k,eps}; 1C charge to 4.2 V (side reaction active) then 1C discharge via simulate!",
                    metric="charge+discharge sims/sec (LCO SEI 322-DAE, FP64)", cathode="LCO", temperature=False, aging=True,
                    batch=65536, soc0=0.0, t_mid=1800.0, dense=(0.0, 7400.0, 60.0),
                    protocol=[("I", 0, 1.0, 1e6, {"V_max": 4.2}), ("I", 0, -1.0, 1e6, {"V_max": 4.2})]),
}
# the config that rides along with the headline at a given GPU count (the count BASELINE.json names it for)
EXTRA_AT = {1: "cfg5", 4: "cfg4", 8: "cfg3"}
CPU_PER_CORE = {"cfg2": 256, "cfg3": 48, "cfg4": 12, "cfg5": 12, "cfg5n10": 96}     # reference-arm sample per host core
METH = {"I": 0, "V": 1, "P": 2}
KINDS = ("value", "hold", "rest")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def n_states(W):
    g = W.get("grid", {})
    Np, Ns, Nn = g.get("N_p", 10), g.get("N_s", 10), g.get("N_n", 10)
    Nx, Ne = Np + Ns + Nn, Np + Nn
    return 2 * Nx + 12 * Ne + 1 + (Nx + 20 if W["temperature"] else 0) + (2 * Nn + 1 if W.get("aging") else 0)


def config_block(W, world, reltol=1e-3, abstol=1e-6):
    """the same for the GPU arm and the reference arm: it names the workload, not what a leg happened to sample"""
    return {"workload": W["name"], "segments_per_step": len(W["protocol"]), "n_states": n_states(W),
            "batch_per_gpu": W["batch"], "l2": "flushed (256 MB write) between timed iterations",
            "reltol": reltol, "abstol": abstol,
            "parallelism": f"batch-sharded x{world}, NCCL all-gather of summaries inside the timed region"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# the CPU side (checker / baseline): everything that touches oracle/ lives in these three functions
# ---------------------------------------------------------------------------------------------------------------
def oracle_theta(p_keys, th_product, cathode="LCO"):
    """product-order theta rows -> the oracle's table, by key name (the GPU arm's parameters come from the
    product's own plb_theta_defaults + sweep.randomised_theta; the oracle only ever sees a copy)"""
    import oracle as O
    ascii_of = {"T₀": "T0", "c_e₀": "c_e0", "t₊": "t_plus"}
    tr = str.maketrans({"θ": "theta", "σ": "sigma", "ϵ": "eps", "λ": "lambda", "ρ": "rho"})
    names = O.theta_names()
    out = np.tile(O.theta_defaults(cathode), (th_product.shape[0], 1))
    for i, k in enumerate(p_keys):
        out[:, names.index(ascii_of.get(k, k.translate(tr)))] = th_product[:, i]
    return out


def oracle_protocol(W, tho, nthreads, dense_t=None, reltol=1e-3, abstol=1e-6):
    """the CPU oracle over a protocol: returns the per-segment results"""
    import oracle as O
    m = O.make_model(W["cathode"], temperature=W["temperature"], aging=W.get("aging", False), **W.get("grid", {}))
    opts = O.default_opts(reltol=reltol, abstol=abstol, reltol_init=reltol, abstol_init=abstol)
    out, state = [], None
    for k, (method, kind, value, tf, bo) in enumerate(W["protocol"]):
        b = O.default_bounds(W["cathode"], **bo)
        run = O.make_run(method, value, tf=tf, input_kind=KINDS[kind], new_run=(k == 0))
        r = O.simulate_batch(m, tho, run, opts, b, SOC0=W["soc0"], state=state, nthreads=nthreads, dense_t=dense_t)
        state = r["state"]
        out.append(r)
    return out


def oracle_default_theta(W, sample, first=0):
    """reference arm (no GPU, no product library): the same sweep generated on the oracle's own table"""
    from tests import util
    return util.oracle_theta_batch(sample, cathode=W["cathode"], first=first)


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's path (oracle port; the reference itself
    is Julia + SUNDIALS/KLU and cannot run here) on all host cores, same workload/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample = CPU_PER_CORE[args.workload] * cores if args.sample is None else args.sample
    tho = oracle_default_theta(W, sample)
    for _ in range(args.warmup):
        oracle_protocol(W, tho[:cores * 2], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rs = oracle_protocol(W, tho, cores)
    dt = (time.perf_counter() - t0) / args.steps
    done = int(np.sum(rs[-1]["flag"] >= 0))
    v = done / dt
    out = {
        "impl": "reference", "metric": W["metric"], "value": v, "unit": "sims/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(W, args.gpus),
        "stats": {"mean_steps": float(np.sum([np.mean(r["n_steps"]) for r in rs])), "failed_systems": sample - done,
                  "value_counting_failed_systems": sample / dt},
        "cpu_baseline": {"value": v, "unit": "sims/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} systems of the same batch per step (of {W['batch']} per GPU in the GPU arm), "
                                   f"one simulation per thread work item"},
        "e2e": {"value": v, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Runner:
    """one workload on this rank's GPU: device-resident buffers, the protocol as plb_simulate calls"""

    def __init__(self, W, B, rank, world, local_rank):
        import torch
        import petlion_b200 as P
        from petlion_b200 import _lib, sweep
        self.torch, self._lib, self.L = torch, _lib, _lib.lib()
        self.W, self.B, self.rank, self.world = W, B, rank, world
        self.dev = torch.device("cuda", local_rank)
        L = self.L
        self.p = P.petlion(W["cathode"], temperature=W["temperature"], aging=W.get("aging", False), device=local_rank,
                           **W.get("grid", {}))
        self.h = h = self.p._h
        self.N, self.nth = self.p.N.tot, len(self.p.θ_keys)
        self.th_host = sweep.randomised_theta(self.p, B, first=rank * B)      # every rank gets its own systems
        self.stream = torch.cuda.current_stream()
        L.plb_set_stream(h, C.c_void_p(self.stream.cuda_stream))
        f64 = self.f64 = dict(dtype=torch.float64, device=self.dev)
        N = self.N
        self.d_theta = torch.from_numpy(self.th_host).to(self.dev)
        self.d_soc0 = torch.full((B,), W["soc0"], **f64)
        self.d_Y = torch.zeros(B, N, **f64); self.d_YP = torch.zeros(B, N, **f64)
        self.d_SOC = torch.zeros(B, **f64); self.d_t = torch.zeros(B, **f64)
        self.d_trn = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.o = _lib.Opts(); L.plb_opts_defaults(h, C.byref(self.o))
        self.b = _lib.Bounds(); L.plb_bounds_defaults(h, C.byref(self.b))
        self.segs = []
        for k, (method, kind, value, tf, bo) in enumerate(W["protocol"]):
            bk = _lib.Bounds(); L.plb_bounds_defaults(h, C.byref(bk))
            for name, v in bo.items():
                setattr(bk, name, v)
            self.segs.append((_lib.Run(METH[method], kind, value, tf, 1 if k == 0 else 0, 0), bk))
        self.d_sums = [torch.zeros(B, 10, **f64) for _ in self.segs]          # 80-byte summary records
        self.gathered = torch.empty(world * B, 10, **f64) if world > 1 else self.d_sums[-1]

    def step_device(self, theta_ptr=None, soc_ptr=None, B=None, dense=None):
        # the whole protocol; the state (Y, Y', SOC, t) is handed from segment to segment on the device
        L, _lib = self.L, self._lib
        B = B or self.B
        for k, (rk, bk) in enumerate(self.segs):
            if dense is not None:
                td, dV, dS, dn = dense
                _lib.check(L.plb_set_dense_output(self.h, td.size, td.ctypes.data, dV[k].data_ptr(), None, dS[k].data_ptr(), None,
                                                  None, dn[k].data_ptr(), 1))
            _lib.check(L.plb_simulate(self.h, B, theta_ptr or self.d_theta.data_ptr(), C.byref(rk), None, C.byref(self.o),
                                      C.byref(bk), soc_ptr or self.d_soc0.data_ptr(), self.d_Y.data_ptr(), self.d_YP.data_ptr(),
                                      self.d_SOC.data_ptr(), self.d_t.data_ptr(), self.d_sums[k].data_ptr(), 0, None, None,
                                      None, None, None, None, self.d_trn.data_ptr(), 1))

    def gather(self):
        # the one collective of this path: fixed-size summaries of the last segment, every rank gets all of them
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.gathered, self.d_sums[-1])

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, steps, warmup, flush, sample_clocks=True):
        torch = self.torch
        for _ in range(warmup):
            self.step_device(); self.gather()
        self.barrier()
        sampler = ClockSampler(self.dev.index)
        if sample_clocks:
            sampler.start()
        launches0 = self.L.plb_launch_count(self.h)
        ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(steps)]
        self.barrier()
        for k in range(steps):
            flush.zero_()                                          # L2 flush between timed iterations
            ev[k][0].record(self.stream)
            self.step_device()
            ev[k][1].record(self.stream)
            self.gather()
            ev[k][2].record(self.stream)
        self.barrier()
        ms_local = sum(a.elapsed_time(c) for a, _, c in ev) / steps
        gather_ms = sum(b_.elapsed_time(c) for _, b_, c in ev) / steps
        launches = self.L.plb_launch_count(self.h) - launches0
        clocks = sampler.stop() if sample_clocks else None
        ms = ms_local
        if self.world > 1:
            import torch.distributed as dist
            tmax = torch.tensor([ms_local, gather_ms], **self.f64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms, gather_ms = float(tmax[0].item()), float(tmax[1].item())
        SD = self._lib.SUMMARY_DTYPE
        summ = self.gathered.cpu().numpy().view(SD).reshape(-1)
        seg_summ = [t.cpu().numpy().view(SD).reshape(-1) for t in self.d_sums]   # this rank's segments
        return dict(ms=ms, gather_ms=gather_ms, launches=int(launches), clocks=clocks, summ=summ, seg_summ=seg_summ)

    def e2e(self, reps):
        """the same metric through the C ABI with HOST buffers: parameters in from pinned memory, results out, every step"""
        torch, L, _lib = self.torch, self.L, self._lib
        B, N = self.B, self.N
        pin = lambda *shape, dtype=torch.float64: torch.zeros(*shape, dtype=dtype).pin_memory()   # noqa: E731
        h_theta = torch.from_numpy(self.th_host).pin_memory()
        h_soc0 = torch.full((B,), self.W["soc0"], dtype=torch.float64).pin_memory()
        h_Y, h_SOC, h_t, h_sum = pin(B, N), pin(B), pin(B), pin(B, 10)
        self.h_sum = h_sum
        if len(self.segs) == 1:
            h_trt, h_trV, h_trn = pin(B, N_SAVE_E2E), pin(B, N_SAVE_E2E), pin(B, dtype=torch.int32)
            self.h_tr = (h_trt, h_trV, h_trn)
            run, bk = self.segs[0]

            def step():
                _lib.check(L.plb_simulate(self.h, B, h_theta.data_ptr(), C.byref(run), None, C.byref(self.o), C.byref(bk),
                                          h_soc0.data_ptr(), h_Y.data_ptr(), None, h_SOC.data_ptr(), h_t.data_ptr(),
                                          h_sum.data_ptr(), N_SAVE_E2E, h_trt.data_ptr(), h_trV.data_ptr(), None, None, None,
                                          None, h_trn.data_ptr(), 0))
            d2h = (h_Y.numel() + h_SOC.numel() + h_t.numel() + h_sum.numel() + h_trt.numel() + h_trV.numel()) * 8 + h_trn.numel() * 4
            outputs = f"summary + final Y + (t,V) trajectories [{N_SAVE_E2E} rows]"
        else:
            # multi-segment protocol: parameters come from pinned host memory every step, the state stays on the
            # device between the simulate!/continuation calls, the final summaries and states go back to the host
            e_theta, e_soc = torch.empty_like(self.d_theta), torch.empty_like(self.d_soc0)

            def step():
                e_theta.copy_(h_theta, non_blocking=True); e_soc.copy_(h_soc0, non_blocking=True)
                self.step_device(e_theta.data_ptr(), e_soc.data_ptr())
                h_sum.copy_(self.d_sums[-1], non_blocking=True); h_Y.copy_(self.d_Y, non_blocking=True)
                torch.cuda.synchronize()
            d2h = (h_Y.numel() + h_sum.numel()) * 8
            outputs = "final summaries + final Y; state handed between segments on the device"
        h2d = h_theta.numel() * 8 + h_soc0.numel() * 8
        step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        self.barrier()
        s = (time.perf_counter() - t0) / reps
        if self.world > 1:
            import torch.distributed as dist
            tt = torch.tensor([s], **self.f64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            s = float(tt.item())
        done = int(np.sum(h_sum.numpy().view(_lib.SUMMARY_DTYPE).reshape(-1)["flag"] >= 0))
        if self.world > 1:
            dd = torch.tensor([float(done)], **self.f64)
            dist.all_reduce(dd, op=dist.ReduceOp.SUM)
            done = int(dd.item())
        return dict(seconds=s, h2d=int(h2d), d2h=int(d2h), outputs=outputs, completed=done)

    def k1_roofline(self, flush):
        """roofline of the residual+Jacobian kernel (K1) standalone over the batch, measured live with CUDA events"""
        torch, L, _lib = self.torch, self.L, self._lib
        B, N, W = self.B, self.N, self.W
        run0 = self.segs[0][0]
        # valid mid-run states: integrate the batch part of the way through segment 0, keep (Y, Y') on device
        run_mid = _lib.Run(run0.method, 0, run0.value, W["t_mid"], 1, 0)
        _lib.check(L.plb_simulate(self.h, B, self.d_theta.data_ptr(), C.byref(run_mid), None, C.byref(self.o), C.byref(self.b),
                                  self.d_soc0.data_ptr(), self.d_Y.data_ptr(), self.d_YP.data_ptr(), self.d_SOC.data_ptr(),
                                  self.d_t.data_ptr(), self.d_sums[0].data_ptr(), 0, None, None, None, None, None, None,
                                  self.d_trn.data_ptr(), 1))
        nnz = L.plb_jac_nnz(self.h, 0)
        d_res = torch.empty(B, N, **self.f64); d_nz = torch.empty(B, nnz, **self.f64)
        d_gam = torch.full((B,), 0.05, **self.f64)
        runI = _lib.Run(0, 0, run0.value, 1e6, 1, 0)

        def k1():
            _lib.check(L.plb_resjac(self.h, B, self.d_Y.data_ptr(), self.d_YP.data_ptr(), d_gam.data_ptr(),
                                    self.d_theta.data_ptr(), C.byref(runI), None, d_res.data_ptr(), d_nz.data_ptr(), 1))
        for _ in range(3):
            k1()
        kms = []
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            k1()
            kms.append(L.plb_last_kernel_ms(self.h))               # CUDA events on the launching stream
        k_ms = float(np.mean(kms))
        bytes_per_eval = 8 * (3 * N + self.nth + nnz) + 16          # Y, Y', res, theta, nzval, (t, gamma)
        achieved = B * bytes_per_eval / (k_ms * 1e-3) / 1e9
        peak, which = _peaks()
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get(f"{self._wname}_dram_bytes_per_launch")
            traffic_src = tj.get("source")
        except Exception:
            pass
        return {"kernel": "k_resjac (residual + CSC Jacobian, standalone over the batch)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "peak_source": which, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "bytes_per_eval": bytes_per_eval,
                "evals_per_launch": B, "kernel_ms": k_ms}

    def parity(self, sample, cores):
        """the first `sample` systems of this rank through the GPU (device buffers, dense output on the device) and
        through the CPU oracle: identical step sequences agree to 1e-6; the rest is bounded on a common grid"""
        torch, _lib = self.torch, self._lib
        W = self.W
        t0, t1, dt = W["dense"]
        td = np.arange(t0, t1, dt) + (0.0 if len(self.segs) > 1 else 7.0)
        nd = td.size
        dV = [torch.empty(sample, nd, **self.f64) for _ in self.segs]
        dS = [torch.empty(sample, nd, **self.f64) for _ in self.segs]
        dn = [torch.zeros(sample, dtype=torch.int32, device=self.dev) for _ in self.segs]
        self.step_device(B=sample, dense=(td, dV, dS, dn))
        torch.cuda.synchronize()
        SD = _lib.SUMMARY_DTYPE
        gs = [t[:sample].cpu().numpy().view(SD).reshape(-1) for t in self.d_sums]
        gV = [x.cpu().numpy() for x in dV]
        tho = oracle_theta(self.p.θ_keys, self.th_host[:sample], W["cathode"])
        tw = time.perf_counter()
        rs = oracle_protocol(W, tho, cores, dense_t=td)
        cpu_s = time.perf_counter() - tw
        same = np.ones(sample, dtype=bool)
        okb = np.ones(sample, dtype=bool)
        worst_same = worst_flip = worst_dt = 0.0
        for k, r in enumerate(rs):
            okb &= (gs[k]["flag"] >= 0) & (r["flag"] >= 0)
            same &= np.all([gs[k][f] == r[f] for f in ("n_steps", "flag", "n_res", "n_jac", "n_netf", "n_ncfn")], axis=0)
            a, b = gV[k], r["dense"]["V"]
            both = ~np.isnan(a) & ~np.isnan(b) & okb[:, None]
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
            m_same = both & same[:, None]
            if m_same.any():
                worst_same = max(worst_same, float(rel[m_same].max()))
            # the others: rows up to 30 s before the earlier of the two ends of this segment
            tend = np.minimum(gs[k]["t_end"], r["t_end"])
            m_flip = both & ~same[:, None] & (td[None, :] <= tend[:, None] - 30.0)
            if m_flip.any():
                worst_flip = max(worst_flip, float(rel[m_flip].max()))
            fl = okb & ~same
            if fl.any():
                worst_dt = max(worst_dt, float(np.max(np.abs(gs[k]["t_end"][fl] - r["t_end"][fl]) / np.maximum(np.abs(r["t_end"][fl]), 1.0))))
        fail_g = np.zeros(sample, dtype=bool); fail_c = np.zeros(sample, dtype=bool)
        for k, r in enumerate(rs):
            fail_g |= gs[k]["flag"] < 0; fail_c |= r["flag"] < 0
        return cpu_s, {
            "note": "GPU vs CPU oracle on the same systems, whole protocol; 'identical' = every counter of every segment agrees "
                    "(steps, residual and Jacobian evaluations, error-test and Newton failures, exit flag); V compared at fixed times through both sides' dense output",
            "sample": sample,
            "identical_trajectory_fraction": float(np.mean(same[~fail_c])) if (~fail_c).any() else None,
            "max_rel_dV_on_identical": worst_same,
            "max_rel_dV_nonidentical_common_grid": worst_flip,
            "max_rel_dt_end_nonidentical": worst_dt,
            "hard_failures_cpu": int(fail_c.sum()), "hard_failures_gpu": int(fail_g.sum()),
            "same_systems_fail": bool(np.array_equal(fail_c, fail_g))}


def measure(wname, B, rank, world, local_rank, steps, warmup, flush, cpu_sample, headline):
    """everything bench.py reports about one workload; rank 0 returns the record, the other ranks None"""
    W = WORKLOADS[wname]
    R = Runner(W, B, rank, world, local_rank)
    R._wname = wname
    t = R.timed(steps, warmup, flush)
    e = R.e2e(max(2, steps // 2) if headline else 1)
    ok = t["summ"]["flag"] >= 0
    done = int(ok.sum())
    total = world * B
    # completed simulations only: a hard failure (the reference would have thrown) is not a result
    value, value_all = done / (t["ms"] * 1e-3), total / (t["ms"] * 1e-3)
    rec = None
    roofline = cpu = None
    if rank == 0:
        roofline = R.k1_roofline(flush)
        if cpu_sample:
            cores = os.cpu_count() or 1
            sample = min(cpu_sample, B)
            cpu_s, par = R.parity(sample, cores)
            # the CPU baseline is timed without the dense rows (the parity run above asks for them)
            if headline:
                tho = oracle_theta(R.p.θ_keys, R.th_host[:sample], W["cathode"])
                tw = time.perf_counter()
                rs = oracle_protocol(W, tho, cores)
                cpu_s = time.perf_counter() - tw
                cpu_done = int(np.sum(rs[-1]["flag"] >= 0))
            else:
                cpu_done = sample - par["hard_failures_cpu"]
            cpu = {"value": cpu_done / cpu_s, "unit": "sims/s", "cores": cores, "kind": "port",
                   "sample": f"first {sample} systems of the same batch, whole protocol, {cores} threads, {cpu_s:.1f} s wall",
                   "parity": par}
        seg = t["seg_summ"]
        rec = {
            "metric": W["metric"], "value": value, "unit": "sims/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": t["ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(W, world, R.o.reltol, R.o.abstol),
            "stats": {"mean_steps": float(np.sum([np.mean(q["n_steps"]) for q in seg])),
                      "mean_res_evals": float(np.sum([np.mean(q["n_res"]) for q in seg])),
                      "mean_jac_evals": float(np.sum([np.mean(q["n_jac"]) for q in seg])),
                      "failed_systems": total - done,
                      "value_counting_failed_systems": value_all,
                      "batch_per_gpu_run": B,
                      "exit_flags": {str(int(k)): int(v) for k, v in zip(*np.unique(t["summ"]["flag"], return_counts=True))},
                      "mean_T_end_K": float(np.mean(t["summ"]["T_end"][ok])) if ok.any() else None,
                      "all_gather_ms": t["gather_ms"],
                      "integrator_steps_per_s": float(world * np.sum([np.sum(q["n_steps"]) for q in seg]) / (t["ms"] * 1e-3))},
            "e2e": {"value": e["completed"] / e["seconds"], "unit": "sims/s", "h2d_bytes_per_step": e["h2d"],
                    "d2h_bytes_per_step": e["d2h"], "outputs": e["outputs"]},
            "gpu_launches": t["launches"],
            "clocks": t["clocks"],
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
    del R
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=None, help="systems per GPU (default: the workload's)")
    ap.add_argument("--sample", type=int, default=None, help="CPU-baseline sample size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--extra", default="auto", help="auto | none | a workload name: the config that rides along in extra.configs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)        # > 126 MB L2
    W = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    # CPU sample (N = 1 only, rank 0): about 10-30 s of host work
    cpu_sample = 0
    if world == 1 and not args.no_cpu_baseline:
        cpu_sample = args.sample or (512 * cores if args.workload == "cfg2" else CPU_PER_CORE[args.workload] * cores)
    out = measure(args.workload, args.batch or W["batch"], rank, world, local_rank, args.steps, args.warmup, flush,
                  cpu_sample, headline=True)
    extra = args.extra if args.extra != "auto" else (EXTRA_AT.get(world) if args.workload == "cfg2" else None)
    if extra and extra != "none" and extra != args.workload:
        WE = WORKLOADS[extra]
        # 2 timed steps; parity of the first 256 systems against the oracle on rank 0
        rec = measure(extra, WE["batch"], rank, world, local_rank, 2, 1, flush, 0 if args.no_cpu_baseline else 256, headline=False)
        if rank == 0:
            out["extra"] = {"configs": {extra: rec}}
    if rank == 0:
        out["reference_published"] = {"value": 1e3 / 2.616, "unit": "sims/s",
                                      "note": "PETLION.jl 2.616 ms/sim median, 1 thread, unspecified laptop (examples/getting_started.ipynb)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
