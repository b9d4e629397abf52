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


def synth_theta(p, B, first):
    from tests import util
    tho = util.oracle_theta_batch(B, first=first)
    return util.product_theta_from_oracle(p, tho), tho


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's path (oracle port; the reference itself
    is Julia + SUNDIALS/KLU and cannot run here) on all host cores, same workload/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    from tests import util
    cores = os.cpu_count() or 1
    m = O.make_model("LCO")
    sample = 256 * cores if args.sample is None else args.sample
    tho = util.oracle_theta_batch(sample)
    run = O.make_run("I", -1.0)
    opts, bounds = O.default_opts(), O.default_bounds("LCO")
    for _ in range(args.warmup):
        O.simulate_batch(m, tho[:cores * 8], run, opts, bounds, SOC0=1.0, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = O.simulate_batch(m, tho, run, opts, bounds, SOC0=1.0, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    out = {
        "impl": "reference", "metric": "full-discharge sims/sec (LCO 301-DAE, FP64)", "value": v, "unit": "sims/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: randomised LCO 1C CC discharge batch (N=10/10/10, N_r=10, isothermal)",
                   "sample_per_step": sample, "steps_mean": float(np.mean(r["n_steps"]))},
        "cpu_baseline": {"value": v, "unit": "sims/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} systems of the configs[1] batch per step, one simulation per thread work item"},
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
    import petlion_b200 as P
    from petlion_b200 import _lib
    L = _lib.lib()
    p = P.petlion("LCO", device=local_rank)
    h = p._h
    B = args.batch
    N, nth = p.N.tot, len(p.θ_keys)
    th_host, tho = synth_theta(p, B, first=rank * B)          # every rank gets its own systems
    stream = torch.cuda.current_stream()
    L.plb_set_stream(h, C.c_void_p(stream.cuda_stream))

    # ---------------- device-resident buffers ------------------------------------------------------
    f64 = dict(dtype=torch.float64, device=dev)
    d_theta = torch.from_numpy(th_host).to(dev)
    d_soc0 = torch.ones(B, **f64)
    d_Y = torch.zeros(B, N, **f64); d_YP = torch.zeros(B, N, **f64)
    d_SOC = torch.zeros(B, **f64); d_t = torch.zeros(B, **f64)
    d_sum = torch.zeros(B, 10, **f64)                          # 80-byte summary records
    d_trn = torch.zeros(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, **f64)        # > 126 MB L2
    run = _lib.Run(0, 0, -1.0, 1e6, 1, 0)
    o = _lib.Opts(); L.plb_opts_defaults(h, C.byref(o))
    b = _lib.Bounds(); L.plb_bounds_defaults(h, C.byref(b))

    def step_device():
        _lib.check(L.plb_simulate(h, B, d_theta.data_ptr(), C.byref(run), None, C.byref(o), C.byref(b),
                                  d_soc0.data_ptr(), d_Y.data_ptr(), d_YP.data_ptr(), d_SOC.data_ptr(),
                                  d_t.data_ptr(), d_sum.data_ptr(), 0, None, None, None, None, None,
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
    if world > 1:
        gathered = torch.empty(world * B, 10, **f64)
        dist.all_gather_into_tensor(gathered, d_sum)
        tmax = torch.tensor([ms_local], **f64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    else:
        gathered = d_sum
        ms = ms_local
    summ = gathered.cpu().numpy().view(_lib.SUMMARY_DTYPE).reshape(-1)
    value = world * B / (ms * 1e-3)

    # ---------------- e2e: public API, host buffers, H2D + D2H inside the timed region -------------
    h_theta = torch.from_numpy(th_host).pin_memory()
    h_soc0 = torch.ones(B, dtype=torch.float64).pin_memory()
    h_Y = torch.zeros(B, N, dtype=torch.float64).pin_memory()
    h_SOC = torch.zeros(B, dtype=torch.float64).pin_memory(); h_t = torch.zeros(B, dtype=torch.float64).pin_memory()
    h_sum = torch.zeros(B, 10, dtype=torch.float64).pin_memory()
    h_trt = torch.zeros(B, N_SAVE_E2E, dtype=torch.float64).pin_memory()
    h_trV = torch.zeros(B, N_SAVE_E2E, dtype=torch.float64).pin_memory()
    h_trn = torch.zeros(B, dtype=torch.int32).pin_memory()

    def step_e2e():
        _lib.check(L.plb_simulate(h, B, h_theta.data_ptr(), C.byref(run), None, C.byref(o), C.byref(b),
                                  h_soc0.data_ptr(), h_Y.data_ptr(), None, h_SOC.data_ptr(), h_t.data_ptr(),
                                  h_sum.data_ptr(), N_SAVE_E2E, h_trt.data_ptr(), h_trV.data_ptr(), None, None, None,
                                  h_trn.data_ptr(), 0))

    h2d = h_theta.numel() * 8 + h_soc0.numel() * 8
    d2h = (h_Y.numel() + h_SOC.numel() + h_t.numel() + h_sum.numel() + h_trt.numel() + h_trV.numel()) * 8 + h_trn.numel() * 4
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
        # valid mid-discharge states: integrate the batch to t = 1800 s, keep (Y, Y') on device
        run_mid = _lib.Run(0, 0, -1.0, 1800.0, 1, 0)
        _lib.check(L.plb_simulate(h, B, d_theta.data_ptr(), C.byref(run_mid), None, C.byref(o), C.byref(b),
                                  d_soc0.data_ptr(), d_Y.data_ptr(), d_YP.data_ptr(), d_SOC.data_ptr(),
                                  d_t.data_ptr(), d_sum.data_ptr(), 0, None, None, None, None, None, d_trn.data_ptr(), 1))
        nnz = L.plb_jac_nnz(h, 0)
        d_res = torch.empty(B, N, **f64); d_nz = torch.empty(B, nnz, **f64)
        d_gam = torch.full((B,), 0.05, **f64)
        runI = _lib.Run(0, 0, -1.0, 1e6, 1, 0)

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
                traffic = json.load(f).get("dram_bytes_per_launch_at_65536")
        except Exception:
            pass
        roofline = {"kernel": "k_resjac (residual + CSC Jacobian, standalone over the batch)", "bound": "hbm",
                    "achieved": achieved, "peak": peak, "peak_source": which, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "bytes_per_eval": bytes_per_eval, "evals_per_launch": B, "kernel_ms": k_ms}
        # ---------------- CPU baseline: oracle port on the host cores, bounded sample ---------------
        if world == 1 and not args.no_cpu_baseline:
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
            "metric": "full-discharge sims/sec (LCO 301-DAE, FP64)", "value": value, "unit": "sims/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: batch=65536 LCO 1C CC discharges, randomised {D_s,k,eps}, N=(10,10,10), N_r=10, isothermal",
                       "batch_per_gpu": B, "l2": "flushed (256 MB write) between timed iterations",
                       "reltol": o.reltol, "abstol": o.abstol, "parallelism": f"batch-sharded x{world}, NCCL all-gather of summaries"},
            "stats": {"mean_steps": float(np.mean(summ["n_steps"])), "mean_res_evals": float(np.mean(summ["n_res"])),
                      "mean_jac_evals": float(np.mean(summ["n_jac"])), "failed_systems": int(np.sum(~ok)),
                      "exit_flags": {str(int(k)): int(v) for k, v in zip(*np.unique(summ["flag"], return_counts=True))},
                      "integrator_steps_per_s": float(np.sum(summ["n_steps"]) / (ms * 1e-3))},
            "e2e": {"value": e2e_value, "unit": "sims/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "outputs": f"summary + final Y + (t,V) trajectories [{N_SAVE_E2E} rows]"},
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
