#!/usr/bin/env python3
"""bench.py -- chain-task resamples/sec of the hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--chains C] [--tasks T] [--sweeps S]
  python bench.py --impl reference ...        # the reference's own CPU path on the host cores

A "step" is S Gibbs sweeps (hidden-time slice resampling + posterior draw of the service parameters, the
work of one estimation.bayes iteration each) of C chains per GPU on BASELINE.json configs[1]: the paper's
3-tier web-service network, 5000 tasks (E = 20000 events), 20% of the tasks observed.  Inputs are synthetic:
simulated, sub-sampled and initialised by the package's own host code; every chain starts from the same state.

value  = resamples of all ranks / device time of the K timed steps, chain state resident in HBM.
e2e    = the same through the public host API with HOST buffers: initial state host->device, the sweeps,
         and parameters + per-sweep summaries device->host inside the timed region.
roofline.achieved = 16 B per resample (read + write of one fp64 departure time) x resamples of
         one launch / CUDA-event duration of that launch, against the measured HBM copy bandwidth.
Under torchrun every rank owns C chains on its GPU (weak scaling, no data-path collective); per-chain
summaries are gathered over NCCL once per step only to report R-hat / ESS.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QNET3 = """
states:
  - name: INITIAL
    queues: [ INITIAL ]
    successors: [ TIER1 ]
    initial: TRUE
  - name: TIER1
    queues: [ WEB1, WEB2, WEB3 ]
    successors: [ TIER2 ]
  - name: TIER2
    queues: [ APP1, APP2 ]
    successors: [ TIER3 ]
  - name: TIER3
    queues: [ DB ]
queues:
  - { name: INITIAL, service: [M, 1.0] }
  - { name: WEB1, service: [M, 0.2] }
  - { name: WEB2, service: [M, 0.2] }
  - { name: WEB3, service: [M, 0.2] }
  - { name: APP1, service: [M, 0.05] }
  - { name: APP2, service: [M, 0.05] }
  - { name: DB, service: [M, 0.3] }
"""
TANDEM3 = """
states:
  - {name: INITIAL, queues: [INITIAL], successors: [TIER1], initial: TRUE}
  - {name: TIER1, queues: [Q1], successors: [TIER2]}
  - {name: TIER2, queues: [Q2]}
queues:
  - { name: INITIAL, service: [M, 5.0] }
  - { name: Q1, type: GGk, processors: 3, service: [LN, 1.75, 1.0] }
  - { name: Q2, type: GGk, processors: 3, service: [LN, 0.1, 2.0] }
"""
PSDELAY = """
states:
  - {name: INITIAL, queues: [INITIAL], successors: [TIER1], initial: TRUE}
  - {name: TIER1, queues: [QPS], successors: [TIER2]}
  - {name: TIER2, queues: [QDELAY]}
queues:
  - { name: INITIAL, service: [M, 1.0] }
  - { name: QPS, type: PS, service: [M, 0.5] }
  - { name: QDELAY, type: DELAY, service: [M, 2.0] }
"""
# BASELINE.json configs[1..3] (the headline is configs[1]); algorithmic bytes per resample from SURVEY.md 8d
WORKLOADS = {
    "config2": dict(yaml=QNET3, init="DFN2", tasks=5000, chains=1024, sweeps=20, ref_tasks=5000, bytes=16.0,
                    kernel="sweep_static_kernel<true>", name="3-tier web-service network (qnet3)"),
    "config3": dict(yaml=TANDEM3, init="DFN2", tasks=20000, chains=4096, sweeps=2, ref_tasks=1000, bytes=32.0,
                    kernel="sweep_batch_kernel<4, 8, false, false>", name="tandem source -> G/G/3 -> G/G/3, LogNormal service"),
    "config4": dict(yaml=PSDELAY, init="PS", tasks=50000, chains=2048, sweeps=2, ref_tasks=2000, bytes=80.0 / 3.0,
                    kernel="sweep_batch_kernel<4, 4, false, true>", name="source -> processor sharing -> delay station (one GPU's share "
                                                         "of the 16384 chains)"),
}
BYTES_PER_RESAMPLE = 16.0  # source / GGk k=1 / DELAY visits
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the default configuration, from the ncu --set full
# capture summarised in profiles/r1_static_final_ncu_full.txt (166.0 MB + 108.0 MB)
NCU_TRAFFIC_DEFAULT = 273.98e6
# dram read + write of one sweep_batch_kernel launch of the default config-3 step (profiles/r1_config3_dram_traffic.csv)
NCU_TRAFFIC_CONFIG3 = 762250469376.0 + 251954478592.0
WORKLOAD = "%s, %d tasks, 20%% observed, %d chains/GPU x %d sweeps per step"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled while the timed region runs: through NVML (pynvml, one sample every
    ~2 ms) when available, else by spawning nvidia-smi every 200 ms."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag, self.active = index, [], False, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.bits = [pynvml.nvmlClocksThrottleReasonHwSlowdown, pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                         pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, pynvml.nvmlClocksThrottleReasonSwPowerCap]
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            nv = self.nvml
            while not self.stop_flag:
                if self.active:
                    try:
                        sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                        mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                        self.rows.append([sm, self.max_sm] + ["Active" if mask & b else "Not Active" for b in self.bits])
                    except Exception:
                        pass
                time.sleep(0.002)
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6 and self.active:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(str(r[2 + i]).lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_input(tasks, seed=13, yaml_text=QNET3, init="DFN2"):
    from bayes_qnet_b200 import qnet, qnetu
    net = qnetu.qnet_from_text(yaml_text)
    qnetu.set_seed(seed)
    smp = net.sample(tasks)
    sub = smp.subset_by_task(0.2)
    qnet.set_initialization_type(init)
    try:
        ini = net.gibbs_initialize(sub)
    finally:
        qnet.set_initialization_type("DFN2")
    return net, ini


def cpu_baseline_reference(tasks, sweeps, seed=13, yaml_text=QNET3, init="DFN2"):
    """The reference's own single-chain Cython path (oracle/_ref) on the same kind of input, one process per
    host core, each running `sweeps` iterations of estimation.bayes; returns resamples/s summed over cores."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, ".built")):
        return None
    ncores = os.cpu_count() or 1
    code = r"""
import sys, io, contextlib, time
sys.path.insert(0, %r)
with contextlib.redirect_stdout(io.StringIO()):
    import qnetu, qnet, sampling, estimation
    sampling.DEBUG = 0
    net = qnetu.qnet_from_text(%r)
    qnetu.set_seed(%d + int(sys.argv[1]))
    qnet.set_initialization_type(%r)
    smp = net.sample(%d); sub = smp.subset_by_task(0.2); ini = net.gibbs_initialize(sub)
    qnet.reset_stats()
    t0 = time.time()
    estimation.bayes(net, ini, %d)
    dt = time.time() - t0
    ne = qnet.stats()['nevents']
print("RESULT %%d %%.6f" %% (ne, dt))
""" % (ref, yaml_text, seed, init, tasks, sweeps)
    procs = []
    for c in range(ncores):
        cmd = [sys.executable, "-c", code, str(c)]
        if os.path.exists("/usr/bin/taskset"):
            cmd = ["taskset", "-c", str(c)] + cmd
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    total, per = 0.0, []
    for p in procs:
        out, err = p.communicate()
        for line in out.splitlines():
            if line.startswith("RESULT"):
                ne, dt = line.split()[1:]
                per.append(int(ne) / float(dt))
    if not per:
        return None
    return {"value": sum(per), "per_core": sum(per) / len(per), "cores": len(per)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    tasks = args.ref_tasks
    per_step = []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        r = cpu_baseline_reference(tasks, args.ref_sweeps, seed=13 + i, yaml_text=wl["yaml"], init=wl["init"])
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built on this box"}))
            return
        r["sec"] = time.time() - t0
        if i >= args.warmup:
            per_step.append(r)
    v = sum(r["value"] for r in per_step) / len(per_step)
    cores = per_step[0]["cores"]
    sample = ("reference Cython estimation.bayes (slice), one pinned single-chain process per core, %d-task %s, "
              "%d iterations per step" % (tasks, args.workload, args.ref_sweeps))
    line = {"impl": "reference", "metric": "chain-task resamples/sec", "value": v, "unit": "resamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(r["sec"] for r in per_step) / len(per_step), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (wl["name"], args.tasks, args.chains, args.sweeps), "reference_sample": sample},
            "cpu_baseline": {"value": v, "unit": "resamples/s", "cores": cores, "kind": "reference", "sample": sample,
                             "per_core": v / cores},
            "e2e": {"value": v, "unit": "resamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS),
                    help="BASELINE.json configs[1] (default, the headline), configs[2] or configs[3]")
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU")
    ap.add_argument("--tasks", type=int, default=None)
    ap.add_argument("--sweeps", type=int, default=None, help="Gibbs sweeps per step")
    ap.add_argument("--ref-tasks", type=int, default=None)
    ap.add_argument("--ref-sweeps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tight", action="store_true", help="BQ_FLAG_TIGHT bracket (not the default sampler)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    for k in ("chains", "tasks", "sweeps", "ref_tasks"):
        if getattr(args, k) is None:
            setattr(args, k, wl[k])

    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build(quiet=True)
    from bayes_qnet_b200 import GibbsEngine, capi, diagnostics, parallel
    if not torch.cuda.is_available() or capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: bayes_qnet_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()

    net, ini = make_input(args.tasks, yaml_text=wl["yaml"], init=wl["init"])  # same seed on every rank -> identical data
    C, S = args.chains, args.sweeps
    eng = GibbsEngine(net, ini, n_chains=C, seed=20261019, device=local, chain_offset=rank * C)
    n_elig = len(eng.schedule()[1])
    resamples_per_step = C * S * n_elig
    d0 = ini.soa()["d"].copy()
    hbm_peak, peak_kind = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        eng.sweep(S, "SLICE", "BAYES", sync=True)
    sampler = ClockSampler(local)
    sampler.start()
    launch_ms = []
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    barrier()
    sampler.active = True
    t_total = 0.0
    launches0 = eng.stats()["n_launches"]
    for _ in range(args.steps):
        l2_flush.zero_()
        barrier()
        eng.sweep(S, "SLICE", "BAYES", sync=True)
        launch_ms.append(eng.last_sweep_ms())  # CUDA events on the handle's own stream around the launch
    sampler.active = False
    launches = eng.stats()["n_launches"] - launches0
    t_local = sum(launch_ms) / 1e3
    barrier()
    # ---- ESS / R-hat: the warmed-up chains continue for 5 untimed steps with traces recorded ----------------------
    eng.clear_traces()
    eng.sweep(5 * S, "SLICE", "BAYES", trace=True, sync=True)
    th_ess, mw_ess, _ = eng.traces()
    ess_ms = eng.last_sweep_ms()

    # ---- end-to-end arm: host buffers in, summaries out, inside the timed region ----------------------------
    h2d = d0.nbytes * 1 + 8 * eng.P
    e2e_times = []
    d2h = 0
    for i in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        eng.broadcast_state(d0)                # host -> device: the initial state, replicated on the device
        eng.set_params(net.parameters)
        eng.clear_traces()
        eng.sweep(S, "SLICE", "BAYES", trace=True, sync=True)
        th, mw, lp = eng.traces()               # device -> host: per-sweep theta / mean wait / log-prob
        final = eng.params()
        dt = time.perf_counter() - t0
        d2h = th.nbytes + mw.nbytes + lp.nbytes + final.nbytes
        if i >= args.warmup:
            e2e_times.append(dt)
    h2d = d0.nbytes + C * 8 * eng.P
    t_e2e_local = sum(e2e_times)

    # ---- max over ranks, summaries over NCCL -------------------------------------------------------------------
    times = torch.tensor([t_local, t_e2e_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(times[0]), float(times[1])
    summ = numpy.concatenate((parallel.gather_traces(th_ess, world), parallel.gather_traces(mw_ess[:, :, 1:], world)),
                             axis=2)  # [n_rec, world*C, P + Q-1] on every rank
    barrier()
    sampler.stop_flag = True

    if rank == 0:
        value = world * resamples_per_step * args.steps / t_dev
        e2e = world * resamples_per_step * args.steps / t_e2e
        worst_ms = max(launch_ms)
        mean_ms = sum(launch_ms) / len(launch_ms)
        achieved = resamples_per_step * wl["bytes"] / (mean_ms / 1e3) / 1e9
        ess = diagnostics.ess(summ)
        rhat = diagnostics.rhat(summ)
        info = eng.kernel_info()
        line = {
            "metric": "chain-task resamples/sec", "value": value, "unit": "resamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (wl["name"], args.tasks, C, S), "baseline_config": args.workload,
                       "events_per_chain": eng.E,
                       "resampled_events_per_chain": n_elig, "chains_total": world * C, "sweeps_per_step": S,
                       "param_update": "BAYES (posterior draw every sweep)", "sampler": "slice" + (" tight" if args.tight else ""),
                       "l2": "256 MiB buffer written between timed steps (L2 flush); chain state 8*E*C = %.0f MB" % (8e-6 * eng.E * C),
                       "threads_per_cta": info["threads"], "state_in_smem": info["state_in_smem"],
                       "smem_bytes_per_cta": info["smem_bytes"], "n_colors": len(eng.schedule()[0]) - 1},
            "e2e": {"value": e2e, "unit": "resamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak,
                         "traffic": NCU_TRAFFIC_DEFAULT if (args.workload, args.tasks, C, S) == ("config2", 5000, 1024, 20) else
                                    (NCU_TRAFFIC_CONFIG3 if (args.workload, args.tasks, C, S) == ("config3", 20000, 4096, 2) else None),
                         "traffic_note": ("bytes per launch from profiles/r1_static_final_ncu_full.txt; far BELOW the "
                                          "algorithmic 16 B x resamples because the chain state stays in shared memory "
                                          "for all sweeps of a launch") if args.workload == "config2" else
                                         ("bytes per launch from profiles/r1_config3_dram_traffic.csv: 80x the algorithmic 32 B x resamples -- every "
                                          "evaluation re-reads the event's neighbourhood from HBM (L1 hit rate 54 %, "
                                          "profiles/r1_batch_tandem_4096_ncu_full.txt)") if args.workload == "config3" else
                                         "see profiles/r1_batch_tandem_4096_ncu_full.txt (256 GB per launch of 98 M resamples on the 2000-task tandem)",
                         "peak_kind": peak_kind,
                         "kernel": wl["kernel"], "bytes_per_resample": wl["bytes"],
                         "resamples_per_launch": resamples_per_step, "launch_ms_mean": mean_ms, "launch_ms_max": worst_ms},
            "clocks": sampler.summary(),
            "ess_per_s": float(numpy.min(ess)) / (ess_ms / 1e3) if numpy.all(numpy.isfinite(ess)) else None,
            "ess_note": "min over service parameters and per-queue mean waits, all chains pooled, %d traced sweeps "
                        "after %d warm-up sweeps" % (5 * S, (args.warmup + args.steps) * S),
            "rhat_max": float(numpy.max(rhat)) if numpy.all(numpy.isfinite(rhat)) else None,
            "slice_evals_per_resample": eng.stats()["slice_evals"] / max(eng.stats()["nevents"], 1),
        }
        if not args.no_cpu_baseline and world == 1:
            t0 = time.time()
            cb = cpu_baseline_reference(args.ref_tasks, args.ref_sweeps, yaml_text=wl["yaml"], init=wl["init"])
            if cb is not None:
                line["cpu_baseline"] = {
                    "value": cb["value"], "unit": "resamples/s", "cores": cb["cores"], "kind": "reference",
                    "per_core": cb["per_core"],
                    "sample": "reference Cython estimation.bayes (slice), one pinned single-chain process per core, "
                              "%d-task %s, %d iterations each (%.0f s wall)" % (args.ref_tasks, args.workload, args.ref_sweeps, time.time() - t0)}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
