#!/usr/bin/env python3
"""bench.py -- headline benchmark of BASELINE.json: full-discharge sims/sec (LCO 301-DAE, FP64) and
the HBM roofline fraction of the residual+Jacobian kernel.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the CPU oracle port on all host cores)

One "step" = one pass of the hot path over one batch of synthetic input: B = 65 536 independent
1C CC discharges (LCO, N=(10,10,10), N_r=10, isothermal, SOC 1 -> SOC_min/V_min) with randomised
{D_s, k, eps} (BASELINE.json configs[1]), per GPU (weak scaling: simulations are independent,
ranks share nothing; one NCCL all-gather collects the 80-byte per-system summaries).
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

# Workloads = BASELINE.json configs.  The default (and the line the driver records) is configs[1]; the
# others are measured with --workload and kept under profiles/.  A protocol is a list of segments
# (method, input_kind, value, tf, bound overrides); segment 0 is simulate(), the rest simulate!().
WORKLOADS = {
    "cfg2": dict(name="configs[1]: batch=65536 LCO 1C CC discharges, randomised {D_s,k,eps}, N=(10,10,10), N_r=10, isothermal",
                 metric="full-discharge sims/sec (LCO 301-DAE, FP64)", cathode="LCO", temperature=False, batch=65536,
                 soc0=1.0, protocol=[("I", 0, -1.0, 1e6, {})]),
    "cfg3": dict(name="configs[2]: LCO CC-CV (4C -> 4.1 V, then V=:hold to SOC_max) with temperature=true (351 DAEs), "
                      "randomised {D_s,k,eps}; 262144 systems over 8 GPUs = 32768 per GPU",
                 metric="CC-CV protocol sims/sec (LCO thermal 351-DAE, FP64)", cathode="LCO", temperature=True,
                 batch=32768, soc0=0.0,
                 protocol=[("I", 0, 4.0, 1e6, {"V_max": 4.1}), ("V", 1, 0.0, 1e6, {"V_max": 4.1})]),
    "cfg4": dict(name="configs[3]: NMC GITT, SOC0=0, 20 x {I=+1C for 180 s ; I=:rest for 7200 s} via simulate!, "
                      "randomised {D_s,k,eps}; 131072 systems over 4 GPUs = 32768 per GPU",
                 metric="GITT protocol sims/sec (NMC 301-DAE, FP64)", cathode="NMC", temperature=False,
                 batch=32768, soc0=0.0,
                 protocol=[seg for _ in range(20) for seg in (("I", 0, 1.0, 180.0, {}), ("I", 2, 0.0, 7200.0, {}))]),
    "cfg5": dict(name="configs[4]: batch=65536 LCO with aging=:SEI on the refined grid N=(20,20,20) (642 DAEs, two warps per "
                      "system), randomised {D_s,k,eps}; 1C charge to 4.2 V (side reaction active) then 1C discharge via simulate!",
                 metric="charge+discharge sims/sec (LCO SEI 642-DAE, FP64)", cathode="LCO", temperature=False, aging=True,
                 grid=dict(N_p=20, N_s=20, N_n=20), batch=65536, soc0=0.0,
                 protocol=[("I", 0, 1.0, 1e6, {"V_max": 4.2}), ("I", 0, -1.0, 1e6, {"V_max": 4.2})]),
    "cfg5n10": dict(name="configs[4] physics on the N=(10,10,10) grid (322 DAEs; the refined N=(20,20,20) grid is not built): "
                         "batch=65536 LCO with aging=:SEI, randomised {D_s,k,eps}; 1C charge to 4.2 V (side reaction active) "
                         "then 1C discharge via simulate!",
                    metric="charge+discharge sims/sec (LCO SEI 322-DAE, FP64)", cathode="LCO", temperature=False, aging=True,
                    batch=65536, soc0=0.0,
                    protocol=[("I", 0, 1.0, 1e6, {"V_max": 4.2}), ("I", 0, -1.0, 1e6, {"V_max": 4.2})]),
}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def synth_theta(p, B, first, cathode="LCO"):
    from tests import util
    tho = util.oracle_theta_batch(B, cathode=cathode, first=first)
    return util.product_theta_from_oracle(p, tho), tho


def oracle_protocol(W, tho, nthreads, n_save_max=0):
    """the CPU oracle over the same protocol (checker / CPU baseline): returns the per-segment results"""
    import oracle as O
    m = O.make_model(W["cathode"], temperature=W["temperature"], aging=W.get("aging", False), **W.get("grid", {}))
    opts = O.default_opts()
    out, state = [], None
    for k, (method, kind, value, tf, bo) in enumerate(W["protocol"]):
        b = O.default_bounds(W["cathode"], **bo)
        run = O.make_run(method, value, tf=tf, input_kind=("value", "hold", "rest")[kind], new_run=(k == 0))
        r = O.simulate_batch(m, tho, run, opts, b, SOC0=W["soc0"], state=state, nthreads=nthreads, n_save_max=n_save_max)
        state = r["state"]
        out.append(r)
    return out


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's path (oracle port; the reference itself
    is Julia + SUNDIALS/KLU and cannot run here) on all host cores, same workload/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import util
    W = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    per_core = {"cfg2": 256, "cfg3": 48, "cfg4": 12, "cfg5": 12, "cfg5n10": 96}[args.workload]
    sample = per_core * cores if args.sample is None else args.sample
    tho = util.oracle_theta_batch(sample, cathode=W["cathode"])
    for _ in range(args.warmup):
        oracle_protocol(W, tho[:cores * 2], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rs = oracle_protocol(W, tho, cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    out = {
        "impl": "reference", "metric": W["metric"], "value": v, "unit": "sims/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": W["name"], "sample_per_step": sample,
                   "steps_mean": float(np.sum([np.mean(r["n_steps"]) for r in rs]))},
        "cpu_baseline": {"value": v, "unit": "sims/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} systems of the same batch per step, one simulation per thread work item"},
        "e2e": {"value": v, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="systems per GPU")
    ap.add_argument("--sample", type=int, default=None, help="CPU-baseline sample size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    W = WORKLOADS[args.workload]
    if args.batch == B_PER_GPU:
        args.batch = W["batch"]
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
    import petlion_b200 as P
    from petlion_b200 import _lib
    L = _lib.lib()
    p = P.petlion(W["cathode"], temperature=W["temperature"], aging=W.get("aging", False), device=local_rank,
                  **W.get("grid", {}))
    h = p._h
    B = args.batch
    N, nth = p.N.tot, len(p.θ_keys)
    th_host, tho = synth_theta(p, B, first=rank * B, cathode=W["cathode"])   # every rank gets its own systems
    METH = {"I": 0, "V": 1, "P": 2}
    stream = torch.cuda.current_stream()
    L.plb_set_stream(h, C.c_void_p(stream.cuda_stream))

    # ---------------- device-resident buffers ------------------------------------------------------
    f64 = dict(dtype=torch.float64, device=dev)
    d_theta = torch.from_numpy(th_host).to(dev)
    d_soc0 = torch.full((B,), W["soc0"], **f64)
    d_Y = torch.zeros(B, N, **f64); d_YP = torch.zeros(B, N, **f64)
    d_SOC = torch.zeros(B, **f64); d_t = torch.zeros(B, **f64)
    d_sum = torch.zeros(B, 10, **f64)                          # 80-byte summary records
    d_trn = torch.zeros(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, **f64)        # > 126 MB L2
    o = _lib.Opts(); L.plb_opts_defaults(h, C.byref(o))
    b = _lib.Bounds(); L.plb_bounds_defaults(h, C.byref(b))
    segs = []
    for k, (method, kind, value, tf, bo) in enumerate(W["protocol"]):
        bk = _lib.Bounds(); L.plb_bounds_defaults(h, C.byref(bk))
        for name, v in bo.items():
            setattr(bk, name, v)
        segs.append((_lib.Run(METH[method], kind, value, tf, 1 if k == 0 else 0, 0), bk))
    run = segs[0][0]
    d_sums = [d_sum] + [torch.zeros(B, 10, **f64) for _ in segs[1:]]

    def step_device(theta_ptr=None, soc_ptr=None):
        # the whole protocol; the state (Y, Y', SOC, t) is handed from segment to segment on the device
        for k, (rk, bk) in enumerate(segs):
            _lib.check(L.plb_simulate(h, B, theta_ptr or d_theta.data_ptr(), C.byref(rk), None, C.byref(o), C.byref(bk),
                                      soc_ptr or d_soc0.data_ptr(), d_Y.data_ptr(), d_YP.data_ptr(), d_SOC.data_ptr(),
                                      d_t.data_ptr(), d_sums[k].data_ptr(), 0, None, None, None, None, None, None,
                                      d_trn.data_ptr(), 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.plb_launch_count(h)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.zero_()                                          # L2 flush between timed iterations
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    barrier()
    ms_local = sum(a.elapsed_time(c) for a, c in ev) / args.steps
    launches = L.plb_launch_count(h) - launches0
    clocks = sampler.stop()
    # one NCCL all-gather of the fixed-size summaries (the only collective on this path)
    d_last = d_sums[-1]
    if world > 1:
        gathered = torch.empty(world * B, 10, **f64)
        dist.all_gather_into_tensor(gathered, d_last)
        tmax = torch.tensor([ms_local], **f64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    else:
        gathered = d_last
        ms = ms_local
    summ = gathered.cpu().numpy().view(_lib.SUMMARY_DTYPE).reshape(-1)
    seg_summ = [t.cpu().numpy().view(_lib.SUMMARY_DTYPE).reshape(-1) for t in d_sums]   # this rank's segments
    value = world * B / (ms * 1e-3)

    # ---------------- e2e: public API, host buffers, H2D + D2H inside the timed region -------------
    h_theta = torch.from_numpy(th_host).pin_memory()
    h_soc0 = torch.full((B,), W["soc0"], dtype=torch.float64).pin_memory()
    h_Y = torch.zeros(B, N, dtype=torch.float64).pin_memory()
    h_SOC = torch.zeros(B, dtype=torch.float64).pin_memory(); h_t = torch.zeros(B, dtype=torch.float64).pin_memory()
    h_sum = torch.zeros(B, 10, dtype=torch.float64).pin_memory()
    h_trt = torch.zeros(B, N_SAVE_E2E, dtype=torch.float64).pin_memory()
    h_trV = torch.zeros(B, N_SAVE_E2E, dtype=torch.float64).pin_memory()
    h_trn = torch.zeros(B, dtype=torch.int32).pin_memory()

    if len(segs) == 1:
        def step_e2e():
            _lib.check(L.plb_simulate(h, B, h_theta.data_ptr(), C.byref(run), None, C.byref(o), C.byref(segs[0][1]),
                                      h_soc0.data_ptr(), h_Y.data_ptr(), None, h_SOC.data_ptr(), h_t.data_ptr(),
                                      h_sum.data_ptr(), N_SAVE_E2E, h_trt.data_ptr(), h_trV.data_ptr(), None, None, None, None,
                                      h_trn.data_ptr(), 0))
        h2d = h_theta.numel() * 8 + h_soc0.numel() * 8
        d2h = (h_Y.numel() + h_SOC.numel() + h_t.numel() + h_sum.numel() + h_trt.numel() + h_trV.numel()) * 8 + h_trn.numel() * 4
        e2e_outputs = f"summary + final Y + (t,V) trajectories [{N_SAVE_E2E} rows]"
    else:
        # multi-segment protocol: parameters come from pinned host memory every step, the state stays on the
        # device between the simulate!/continuation calls, the final summaries and states go back to the host
        e_theta = torch.empty_like(d_theta); e_soc = torch.empty_like(d_soc0)

        def step_e2e():
            e_theta.copy_(h_theta, non_blocking=True); e_soc.copy_(h_soc0, non_blocking=True)
            step_device(e_theta.data_ptr(), e_soc.data_ptr())
            h_sum.copy_(d_sums[-1], non_blocking=True); h_Y.copy_(d_Y, non_blocking=True)
            torch.cuda.synchronize()
        h2d = h_theta.numel() * 8 + h_soc0.numel() * 8
        d2h = (h_Y.numel() + h_sum.numel()) * 8
        e2e_outputs = "final summaries + final Y; state handed between segments on the device"
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(2, args.steps // 2)):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / max(2, args.steps // 2)
    if world > 1:
        tt = torch.tensor([e2e_s], **f64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * B / e2e_s

    # ---------------- roofline of the residual+Jacobian kernel (K1), measured live -----------------
    roofline = None
    cpu_baseline = None
    if rank == 0:
        # valid mid-run states: integrate the batch part of the way through segment 0, keep (Y, Y') on device
        t_mid = {"cfg2": 1800.0, "cfg3": 150.0, "cfg4": 100.0, "cfg5": 1800.0, "cfg5n10": 1800.0}[args.workload]
        run_mid = _lib.Run(run.method, 0, run.value, t_mid, 1, 0)
        _lib.check(L.plb_simulate(h, B, d_theta.data_ptr(), C.byref(run_mid), None, C.byref(o), C.byref(b),
                                  d_soc0.data_ptr(), d_Y.data_ptr(), d_YP.data_ptr(), d_SOC.data_ptr(),
                                  d_t.data_ptr(), d_sum.data_ptr(), 0, None, None, None, None, None, None, d_trn.data_ptr(), 1))
        nnz = L.plb_jac_nnz(h, 0)
        d_res = torch.empty(B, N, **f64); d_nz = torch.empty(B, nnz, **f64)
        d_gam = torch.full((B,), 0.05, **f64)
        runI = _lib.Run(0, 0, run.value, 1e6, 1, 0)

        def k1():
            _lib.check(L.plb_resjac(h, B, d_Y.data_ptr(), d_YP.data_ptr(), d_gam.data_ptr(), d_theta.data_ptr(),
                                    C.byref(runI), None, d_res.data_ptr(), d_nz.data_ptr(), 1))
        for _ in range(3):
            k1()
        kms = []
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            k1()
            kms.append(L.plb_last_kernel_ms(h))               # CUDA events on the launching stream
        k_ms = float(np.mean(kms))
        bytes_per_eval = 8 * (3 * N + nth + nnz) + 16          # Y, Y', res, theta, nzval, (t, gamma)
        achieved = B * bytes_per_eval / (k_ms * 1e-3) / 1e9
        peak, which = _peaks()
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                traffic = json.load(f).get({"cfg2": "dram_bytes_per_launch_at_65536"}.get(args.workload, args.workload + "_dram_bytes_per_launch"))
        except Exception:
            pass
        roofline = {"kernel": "k_resjac (residual + CSC Jacobian, standalone over the batch)", "bound": "hbm",
                    "achieved": achieved, "peak": peak, "peak_source": which, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "bytes_per_eval": bytes_per_eval, "evals_per_launch": B, "kernel_ms": k_ms}
        # ---------------- CPU baseline: oracle port on the host cores, bounded sample ---------------
        if world == 1 and not args.no_cpu_baseline and len(segs) > 1:
            cores = os.cpu_count() or 1
            per_core = {"cfg3": 48, "cfg4": 12, "cfg5": 12, "cfg5n10": 96}[args.workload]
            sample = min(per_core * cores if args.sample is None else args.sample, B)
            t0 = time.perf_counter()
            rs = oracle_protocol(W, tho[:sample], cores)
            dt = time.perf_counter() - t0
            gs = h_sum.numpy().view(_lib.SUMMARY_DTYPE).reshape(-1)[:sample]
            rl = rs[-1]
            same = (gs["n_steps"] == rl["n_steps"]) & (gs["flag"] == rl["flag"]) & (rl["flag"] >= 0)
            dv = np.abs(gs["V_end"] - rl["V_end"]) / np.abs(rl["V_end"])
            dt_end = np.abs(gs["t_end"] - rl["t_end"]) / np.abs(rl["t_end"])
            okb = (gs["flag"] >= 0) & (rl["flag"] >= 0)
            cpu_baseline = {"value": sample / dt, "unit": "sims/s", "cores": cores, "kind": "port",
                            "sample": f"first {sample} systems of the same batch, whole protocol, {cores} threads, {dt:.1f} s wall",
                            "parity": {"note": "GPU e2e run vs CPU oracle on the same systems, last segment of the protocol",
                                       "same_steps_and_flag_fraction": float(np.mean(same)),
                                       "max_rel_dV_end_on_same": float(dv[same].max()) if same.any() else None,
                                       "max_rel_dt_end_on_same": float(dt_end[same].max()) if same.any() else None,
                                       "max_rel_dV_end_all": float(dv[okb].max()) if okb.any() else None,
                                       "hard_failures_cpu": int(np.sum(rl["flag"] < 0)),
                                       "hard_failures_gpu": int(np.sum(gs["flag"] < 0))}}
        elif world == 1 and not args.no_cpu_baseline:
            import oracle as O
            cores = os.cpu_count() or 1
            sample = 512 * cores if args.sample is None else args.sample
            sample = min(sample, B)
            t0 = time.perf_counter()
            ref = O.simulate_batch(O.make_model("LCO"), tho[:sample], O.make_run("I", -1.0), O.default_opts(),
                                   O.default_bounds("LCO"), SOC0=1.0, nthreads=cores)
            dt = time.perf_counter() - t0
            # parity of the same systems (untimed): step-by-step trajectories of the e2e run vs the oracle
            refT = O.simulate_batch(O.make_model("LCO"), tho[:sample], O.make_run("I", -1.0), O.default_opts(),
                                    O.default_bounds("LCO"), SOC0=1.0, nthreads=cores, n_save_max=N_SAVE_E2E)
            gs = h_sum.numpy().view(_lib.SUMMARY_DTYPE).reshape(-1)[:sample]
            gt, gV, gn = h_trt.numpy()[:sample], h_trV.numpy()[:sample], h_trn.numpy()[:sample]
            ok_ref = refT["flag"] >= 0
            same_seq = np.zeros(sample, dtype=bool)
            worst_dv = 0.0
            for i in range(sample):
                n = int(refT["traj_n"][i])
                if not ok_ref[i] or gn[i] != n or gs["flag"][i] != refT["flag"][i]:
                    continue
                if np.allclose(gt[i, :n], refT["traj"]["t"][i, :n], rtol=1e-9, atol=1e-12):
                    dvi = float(np.max(np.abs(gV[i, :n] - refT["traj"]["V"][i, :n]) / np.abs(refT["traj"]["V"][i, :n])))
                    if dvi < 1e-6:
                        same_seq[i] = True
                        worst_dv = max(worst_dv, dvi)
            others = ok_ref & ~same_seq & (gs["flag"] >= 0)
            dv_others = float(np.max(np.abs(gs["V_end"][others] - refT["V_end"][others]) / np.abs(refT["V_end"][others]))) if others.any() else 0.0
            cpu_baseline = {"value": sample / dt, "unit": "sims/s", "cores": cores, "kind": "port",
                            "sample": f"first {sample} systems of the same batch, {cores} threads, {dt:.1f} s wall",
                            "parity": {"note": "GPU e2e run vs CPU oracle on the same systems; 'identical' = same step "
                                               "times (rtol 1e-9), same exit flag and V trace within rtol 1e-6",
                                       "identical_trajectory_fraction": float(np.mean(same_seq[ok_ref])),
                                       "max_rel_dV_on_identical": worst_dv,
                                       "max_rel_dV_end_on_round_off_decision_flips": dv_others,
                                       "hard_failures_cpu": int(np.sum(refT["flag"] < 0)),
                                       "hard_failures_gpu": int(np.sum(gs["flag"] < 0))}}

    if rank == 0:
        ok = summ["flag"] >= 0
        out = {
            "metric": W["metric"], "value": value, "unit": "sims/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": W["name"], "segments_per_step": len(segs), "n_states": N,
                       "batch_per_gpu": B, "l2": "flushed (256 MB write) between timed iterations",
                       "reltol": o.reltol, "abstol": o.abstol, "parallelism": f"batch-sharded x{world}, NCCL all-gather of summaries"},
            "stats": {"mean_steps": float(np.sum([np.mean(q["n_steps"]) for q in seg_summ])),
                      "mean_res_evals": float(np.sum([np.mean(q["n_res"]) for q in seg_summ])),
                      "mean_jac_evals": float(np.sum([np.mean(q["n_jac"]) for q in seg_summ])),
                      "failed_systems": int(np.sum(~ok)),
                      "exit_flags": {str(int(k)): int(v) for k, v in zip(*np.unique(summ["flag"], return_counts=True))},
                      "mean_T_end_K": float(np.mean(summ["T_end"][ok])) if ok.any() else None,
                      "integrator_steps_per_s": float(world * np.sum([np.sum(q["n_steps"]) for q in seg_summ]) / (ms * 1e-3))},
            "e2e": {"value": e2e_value, "unit": "sims/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "outputs": e2e_outputs},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "reference_published": {"value": 1e3 / 2.616, "unit": "sims/s", "note": "PETLION.jl 2.616 ms/sim median, 1 thread, unspecified laptop (examples/getting_started.ipynb)"},
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
