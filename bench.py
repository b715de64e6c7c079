#!/usr/bin/env python3
"""bench.py -- chain-task resamples/sec of the hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config3|config4] [--chains C] ...
  python bench.py --impl reference ...        # the reference's own CPU path on the host cores
  python bench.py --workload config5 [--gpus N]   # chain-count sweep 1..65536 (strong and weak), one JSON line

A "step" is S Gibbs sweeps (hidden-time slice resampling + posterior draw of the service parameters, the work of one
estimation.bayes iteration each) of C chains per GPU.  Headline workload = BASELINE.json configs[1]: the paper's
3-tier web-service network, 5000 tasks (E = 20000 events), 20% of the tasks observed, 1024 chains per GPU.  The
initial state is the one the REFERENCE generated (sample -> subset_by_task -> gibbs_initialize, seed 13), committed as
tests/golden/big_*.npz by oracle/make_golden.py, so the GPU arm and the reference arm start from the identical
state; when a fixture is absent the package's own host code simulates and initialises one.

value  = resamples of all ranks / device time of the K timed steps, chain state resident in HBM.
e2e    = the same through the public host API with HOST buffers: initial state host->device, the sweeps, parameters
         + per-sweep summaries device->host, and (N > 1) the NCCL all-gather of the per-chain summaries plus the
         R-hat / ESS reductions on the device, all inside the timed region.
roofline.achieved = algorithmic bytes per resample (16 B source / single-server / delay visits, 32 B
         multi-server, 48 B PS) x resamples of one launch / CUDA-event duration of that launch, against the measured
         HBM copy bandwidth.
extra_workloads = short sub-runs of BASELINE configs[0], [2], [3] in the same process (same measurement rules), so that
         every config has a driver-timed figure; cpu_baseline beside each.
Under torchrun every rank owns C chains on its GPU (weak scaling, no data-path collective).
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
GOLDEN = os.path.join(ROOT, "tests", "golden")

MM1 = """
states:
  - {name: INITIAL, queues: [INITIAL], successors: [TIER1], initial: TRUE}
  - {name: TIER1, queues: [WEB1]}
queues:
  - { name: INITIAL, service: [M, 1.0] }
  - { name: WEB1, service: [M, 0.7] }
"""
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
# BASELINE.json configs[0..3] (the headline is configs[1]); algorithmic bytes per resample from SURVEY.md 8d.
# fixture / ref_fixture: reference-generated initial states (tests/golden/big_<name>.npz) for the GPU arm and for the
# bounded CPU sample; ref_tasks: size of the CPU sample when no fixture applies.
WORKLOADS = {
    "config1": dict(yaml=MM1, init="DFN2", tasks=200, pct=0.5, chains=1, sweeps=1000, ref_tasks=200, bytes=16.0,
                    fixture=None, ref_fixture=None, kernel="sweep_static_kernel<true, true>",
                    name="single M/M/1 queue (configs[0])"),
    "config2": dict(yaml=QNET3, init="DFN2", tasks=5000, pct=0.2, chains=1024, sweeps=20, ref_tasks=5000, bytes=16.0,
                    fixture="qnet3_5k", ref_fixture="qnet3_5k", kernel="sweep_static_kernel<true, true>",
                    name="3-tier web-service network (qnet3)"),
    "config3": dict(yaml=TANDEM3, init="DFN2", tasks=20000, pct=0.2, chains=4096, sweeps=2, ref_tasks=1000, bytes=32.0,
                    fixture="lnk_20k", ref_fixture=None, kernel="sweep_batch_kernel_nw<4, 4, false, 0, 2> (4 waves of 1024 chains)",
                    name="tandem source -> G/G/3 -> G/G/3, LogNormal service"),
    "config4": dict(yaml=PSDELAY, init="PS", tasks=50000, pct=0.2, chains=2048, sweeps=2, ref_tasks=2000, bytes=80.0 / 3.0,
                    fixture="psdelay_50k", ref_fixture=None, kernel="sweep_batch_kernel_nw<4, 4, false, 1, 2> (2 waves of 1024 chains)",
                    name="source -> processor sharing -> delay station (one GPU's share of the 16384 chains)"),
}
# dram__bytes_read.sum + dram__bytes_write.sum of one STEP (all its launches), from ncu captures under profiles/; the line
# reports it per launch
NCU_TRAFFIC = {
    ("config2", 5000, 1024, 20): (273.98e6, "profiles/r1_static_final_ncu_full.txt (166.0 MB read + 108.0 MB written): far BELOW "
                                            "the algorithmic 16 B x resamples because the chain state stays in shared memory "
                                            "for all sweeps of a launch"),
    ("config3", 20000, 4096, 2): (553471957504.0 + 163293786368.0, "profiles/r2_config3_dram_traffic.csv (launches 4-7 = the timed step, "
                                  "four waves of 1024 chains): 1.8 KB per resample, 57x the algorithmic 32 B -- witness scans and walks "
                                  "re-read each event's neighbourhood; round 1: 1014 GB per step (2.6 KB per resample) at 0.70x the rate, "
                                  "all 4096 chains resident at once: 1298 GB"),
    ("config4", 50000, 2048, 2): (405891099392.0 + 132882565888.0, "profiles/r2_config4_dram_traffic.csv (launches 2-3 = the timed step, two "
                                  "waves of 1024 chains): 1.1 KB per resample, 41x the algorithmic 26.7 B"),
}
WORKLOAD = "%s, %d tasks, %d%% observed, %d chains/GPU x %d sweeps per step"
# measured once, offline, at the REAL size on this repository's build container (tools/ref_fullsize.py; BASELINE.md
# section 3: "run a fixed small number of sweeps at the real size ... do not extrapolate"): profiles/r2_reference_fullsize.json
FULLSIZE = os.path.join(ROOT, "profiles", "r2_reference_fullsize.json")


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


# ---------------------------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------------------------
def fixture_path(name):
    p = os.path.join(GOLDEN, "big_%s.npz" % name) if name else None
    return p if p and os.path.exists(p) else None


def fixture_arrays(path, tag="s0"):
    """Flat arrays of a reference-generated state (layout of oracle/make_golden.py do_big): rows task-major in visit
    order, arrival = departure of the task's previous event (0 for its first)."""
    import numpy
    g = numpy.load(path, allow_pickle=False)
    d, tid = g[tag + "_d"], g["tid"]
    a = numpy.zeros_like(d)
    same = tid[1:] == tid[:-1]
    a[1:][same] = d[:-1][same]
    return dict(yaml=str(g["yaml"]), theta=g["theta"], tid=tid, qid=g["qid"], state=g["state"], a=a, d=d,
                obs_a=g["obs_a"], obs_d=g["obs_d"], eid=g["eid"])


def make_input(tasks, seed=13, yaml_text=QNET3, init="DFN2", fixture=None, pct=0.2):
    """(net, initialised Arrivals, description).  From the reference-generated fixture when there is one of this
    size, else simulated, sub-sampled and initialised by the package's own host code."""
    from bayes_qnet_b200 import qnet, qnetu
    from bayes_qnet_b200.arrivals import Arrivals
    path = fixture_path(fixture)
    if path is not None:
        f = fixture_arrays(path)
        if len(set(f["tid"].tolist())) == tasks:
            net = qnetu.qnet_from_text(f["yaml"])
            net.parameters = list(f["theta"])
            ini = Arrivals.from_arrays(net, f["tid"], f["qid"], f["state"], f["a"], f["d"], f["obs_a"], f["obs_d"], eid=f["eid"])
            return net, ini, "reference-generated state tests/golden/big_%s.npz (sample -> subset_by_task -> gibbs_initialize, seed 13)" % fixture
    net = qnetu.qnet_from_text(yaml_text)
    qnetu.set_seed(seed)
    smp = net.sample(tasks)
    sub = smp.subset_by_task(pct)
    qnet.set_initialization_type(init)
    try:
        ini = net.gibbs_initialize(sub)
    finally:
        qnet.set_initialization_type("DFN2")
    return net, ini, "simulated, sub-sampled and initialised by the package's host code (seed %d)" % seed


# ---------------------------------------------------------------------------------------------------------------
# the reference's own CPU path (oracle/_ref), in subprocesses
# ---------------------------------------------------------------------------------------------------------------
REF_CODE = r"""
import sys, io, os, contextlib, time
os.environ.setdefault("BQ_REF_BIG_N", "65536")
sys.path.insert(0, %(ref)r)
import numpy
with contextlib.redirect_stdout(io.StringIO()):
    import qnetu, qnet, sampling, estimation
    sampling.DEBUG = 0
    estimation.set_sampler(%(sampler)r)
    fixture = %(fixture)r
    if fixture:   # the committed reference-generated state, through the reference's own table reader (qnetu.py:474)
        g = numpy.load(fixture, allow_pickle=False)
        d, tid = g["s0_d"], g["tid"]
        a = numpy.zeros_like(d); same = tid[1:] == tid[:-1]; a[1:][same] = d[:-1][same]
        lines = ["eid tid qid a d s w state proc obs_a obs_d"]
        for i in range(len(d)):
            lines.append("%%d %%d %%d %%r %%r 0.0 0.0 %%d 0 %%d %%d" %% (g["eid"][i], tid[i], g["qid"][i], float(a[i]), float(d[i]),
                                                                   g["state"][i], g["obs_a"][i], g["obs_d"][i]))
        net = qnetu.qnet_from_text(str(g["yaml"]))
        net.parameters = list(g["theta"])
        ini = qnetu.read_csv_from_lines(net, lines)
        qnetu.set_seed(%(seed)d + int(sys.argv[1]))
    else:
        net = qnetu.qnet_from_text(%(yaml)r)
        qnetu.set_seed(%(seed)d + int(sys.argv[1]))
        qnet.set_initialization_type(%(init)r)
        smp = net.sample(%(tasks)d); sub = smp.subset_by_task(%(pct)r); ini = net.gibbs_initialize(sub)
    qnet.reset_stats()
    t0 = time.time()
    if %(sampler)r == "MH":   # estimation.bayes with sampling_type MH cannot report (estimation.py:141); its loop spelled out
        net.estimate(ini)
        arrv = ini
        for i in range(%(iters)d):
            arrv = net.gibbs_resample(arrv, burn=0, num_iter=1, return_arrv=False)[-1]
            net.sample_parameters(arrv)
    else:
        estimation.bayes(net, ini, %(iters)d)
    dt = time.time() - t0
    ne = qnet.stats()['nevents']
print("RESULT %%d %%.6f" %% (ne, dt))
"""


def reference_run(wl, iters, nproc, sampler="SLICE", seed=13, use_fixture=True, tasks=None):
    """`nproc` pinned single-chain processes of the reference, each running `iters` iterations of estimation.bayes
    (timed region: that call only); returns resamples/s summed over the processes."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, ".built")):
        return None
    fx = fixture_path(wl.get("ref_fixture")) if use_fixture else None
    code = REF_CODE % dict(ref=ref, sampler=sampler, fixture=fx, yaml=wl["yaml"], seed=seed, init=wl["init"],
                           tasks=tasks or wl["ref_tasks"], pct=wl["pct"], iters=iters)
    procs = []
    for c in range(nproc):
        cmd = [sys.executable, "-c", code, str(c)]
        if os.path.exists("/usr/bin/taskset"):
            cmd = ["taskset", "-c", str(c)] + cmd
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    per = []
    for p in procs:
        out, err = p.communicate()
        for line in out.splitlines():
            if line.startswith("RESULT"):
                ne, dt = line.split()[1:]
                per.append(int(ne) / float(dt))
    if not per:
        return None
    return {"value": sum(per), "per_core": sum(per) / len(per), "cores": len(per),
            "state": ("identical initial state (tests/golden/big_%s.npz through the reference's table reader)" % wl["ref_fixture"])
                     if fx else "own %d-task state (same generator and seed)" % (tasks or wl["ref_tasks"])}


def fullsize_record(workload):
    """The committed real-size measurement of the reference for this config: the estimation.bayes run (>= 2 iterations)
    when it exists, and the one-sweep figure taken while the fixture was generated."""
    try:
        with open(FULLSIZE) as f:
            data = json.load(f)
    except Exception:
        return None
    rec = data.get(workload)
    one = data.get(workload + "_one_sweep")
    if rec is None:
        return one
    if one is not None:
        rec = dict(rec, one_sweep=one)
    return rec


def cpu_baseline_block(wl, workload, iters, with_single=True, with_mh=True):
    """The reference on this box's host cores: all cores concurrently (value), one process alone (per_core_1proc),
    and the MH sampler, each a bounded sample; plus the committed real-size figure for configs 3 / 4."""
    ncores = os.cpu_count() or 1
    t0 = time.time()
    allc = reference_run(wl, iters, ncores)
    if allc is None:
        return None
    out = {"value": allc["value"], "unit": "resamples/s", "cores": allc["cores"], "kind": "reference",
           "per_core": allc["per_core"]}
    if with_single:
        one = reference_run(wl, iters, 1)
        if one:
            out["per_core_1proc"] = one["value"]
    if with_mh and "type: PS" not in wl["yaml"]:
        mh = reference_run(wl, iters, ncores, sampler="MH")
        if mh:
            out["mh_per_core"] = mh["per_core"]
            out["mh_unit"] = "MH proposals/s per core (Qnet.gibbs_resample + sample_parameters)"
    size = wl["tasks"] if fixture_path(wl.get("ref_fixture")) else wl["ref_tasks"]
    out["sample"] = ("reference Cython estimation.bayes (slice), one pinned single-chain process per core, %d-task %s, %s, "
                     "%d iterations each, timed region = the estimation.bayes call (%.0f s wall for all legs)"
                     % (size, workload, allc["state"], iters, time.time() - t0))
    fs = fullsize_record(workload)
    if fs:
        out["fullsize"] = fs
    return out


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    ncores = os.cpu_count() or 1
    per_step = []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        r = reference_run(wl, args.ref_sweeps if i >= args.warmup else 1, ncores, seed=13 + i)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built on this box"}))
            return
        r["sec"] = time.time() - t0
        if i >= args.warmup:
            per_step.append(r)
    v = sum(r["value"] for r in per_step) / len(per_step)
    cores = per_step[0]["cores"]
    size = wl["tasks"] if fixture_path(wl.get("ref_fixture")) else wl["ref_tasks"]
    sample = ("reference Cython estimation.bayes (slice), one pinned single-chain process per core, %d-task %s, %s, "
              "%d iterations per step (warm-up steps: 1), timed region = the estimation.bayes call"
              % (size, args.workload, per_step[0]["state"], args.ref_sweeps))
    one = reference_run(wl, args.ref_sweeps, 1)
    cb = {"value": v, "unit": "resamples/s", "cores": cores, "kind": "reference", "sample": sample, "per_core": v / cores}
    if one:
        cb["per_core_1proc"] = one["value"]
    if "type: PS" not in wl["yaml"]:
        mh = reference_run(wl, args.ref_sweeps, ncores, sampler="MH")
        if mh:
            cb["mh_per_core"] = mh["per_core"]
    fs = fullsize_record(args.workload)
    if fs:
        cb["fullsize"] = fs
    line = {"impl": "reference", "metric": "chain-task resamples/sec", "value": v, "unit": "resamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(r["sec"] for r in per_step) / len(per_step), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (wl["name"], args.tasks, int(100 * wl["pct"]), args.chains, args.sweeps),
                       "baseline_config": args.workload, "reference_sample": sample},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "resamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Ctx(object):
    """Process-wide handles shared by the sub-runs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            dist.barrier()
        self.l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        self.hbm_peak, self.peak_kind = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]


def measure(ctx, workload, tasks, C, S, steps, warmup, mode="SLICE", tight=False, e2e_steps=None, sampler=None):
    """One workload on this rank's GPU: device-resident arm, end-to-end arm, roofline.  Returns (dict, engine pieces)."""
    import numpy
    from bayes_qnet_b200 import GibbsEngine, parallel
    wl = WORKLOADS[workload]
    net, ini, how = make_input(tasks, yaml_text=wl["yaml"], init=wl["init"], fixture=wl["fixture"], pct=wl["pct"])
    eng = GibbsEngine(net, ini, n_chains=C, seed=20261019, device=ctx.local, chain_offset=ctx.rank * C)
    n_elig = len(eng.sequential_order())
    resamples_per_step = C * S * n_elig
    d0 = ini.soa()["d"].copy()
    theta0 = numpy.array(net.parameters, dtype=numpy.float64)

    # ---- device-resident arm ------------------------------------------------------------------------------------
    for _ in range(warmup):
        eng.sweep(S, mode, "BAYES", sync=True, tight=tight)
    launch_ms = []
    ctx.barrier()
    if sampler is not None:
        sampler.active = True
    launches0 = eng.stats()["n_launches"]
    for _ in range(steps):
        ctx.l2_flush.zero_()
        ctx.barrier()
        eng.sweep(S, mode, "BAYES", sync=True, tight=tight)
        launch_ms.append(eng.last_sweep_ms())  # CUDA events on the handle's own stream around the launch
    if sampler is not None:
        sampler.active = False
    launches = eng.stats()["n_launches"] - launches0
    t_local = sum(launch_ms) / 1e3
    ctx.barrier()

    # ---- end-to-end arm: host buffers in, summaries (+ gather, diagnostics) out, inside the timed region ----------
    e2e_times, gather_ms, d2h = [], [], 0
    n_e2e = steps if e2e_steps is None else e2e_steps
    for i in range(min(warmup, 1) + n_e2e):
        ctx.barrier()
        t0 = time.perf_counter()
        eng.broadcast_state(d0)                # host -> device: the initial state, replicated on the device
        eng.set_params(theta0)
        eng.clear_traces()
        eng.sweep(S, mode, "BAYES", trace=True, sync=True, tight=tight)
        tg = time.perf_counter()
        summ = parallel.gather_summaries_device(eng, ctx.world)     # NCCL all-gather on the device trace buffers
        diag = parallel.device_diagnostics(summ)                   # R-hat / ESS reductions on the device -> small host arrays
        gms = 1e3 * (time.perf_counter() - tg)
        th, mw, lp = eng.traces()               # device -> host: per-sweep theta / mean wait / log-prob of this rank
        final = eng.params()
        dt = time.perf_counter() - t0
        d2h = th.nbytes + mw.nbytes + lp.nbytes + final.nbytes + sum(v.nbytes for v in diag.values())
        if i >= min(warmup, 1):
            e2e_times.append(dt)
            gather_ms.append(gms)
    h2d = d0.nbytes + C * 8 * eng.P
    t_dev, t_e2e = ctx.max_over_ranks([t_local, sum(e2e_times)])
    world = ctx.world
    mean_ms = sum(launch_ms) / len(launch_ms)
    achieved = resamples_per_step * wl["bytes"] / (mean_ms / 1e3) / 1e9
    traffic = NCU_TRAFFIC.get((workload, tasks, C, S))
    info = eng.kernel_info()
    st = eng.stats()
    res = {
        "value": world * resamples_per_step * steps / t_dev,
        "ms_per_step": 1e3 * t_dev / steps,
        "e2e": {"value": world * resamples_per_step * n_e2e / t_e2e, "unit": "resamples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / n_e2e,
                "gather_ms": sum(gather_ms) / max(len(gather_ms), 1),
                "returns": "per-sweep theta / mean wait / log-prob of every chain and the final parameters; the per-chain "
                           "END STATE (%.0f MB here) stays on the device -- estimation.bayes copies back chain 0 only"
                           % (8e-6 * eng.E * C)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": ctx.hbm_peak, "unit": "GB/s",
                     "frac": achieved / ctx.hbm_peak, "traffic": traffic[0] / max(launches / steps, 1) if traffic else None,
                     "traffic_note": traffic[1] if traffic else "no ncu --set full capture of this exact launch shape",
                     "peak_kind": ctx.peak_kind, "kernel": wl["kernel"], "bytes_per_resample": wl["bytes"],
                     "resamples_per_launch": resamples_per_step / max(launches / steps, 1),
                     "launch_ms_mean": mean_ms / max(launches / steps, 1), "launch_ms_max": max(launch_ms) / max(launches / steps, 1),
                     "launches_per_step": launches / steps,
                     "launch_note": "a step of the dynamic-order kernels is one launch per wave of chains (two warps per chain, "
                                    "at most 8 x SMs chains resident); times are per launch = step time / launches per step"},
        "config": {"workload": WORKLOAD % (wl["name"], tasks, int(100 * wl["pct"]), C, S), "baseline_config": workload,
                   "initial_state": how, "events_per_chain": eng.E, "resampled_events_per_chain": n_elig,
                   "chains_total": world * C, "sweeps_per_step": S, "param_update": "BAYES (posterior draw every sweep)",
                   "sampler": ("slice" if mode == "SLICE" else "MH") + (" tight" if tight else ""),
                   "l2": "256 MiB buffer written between timed steps (L2 flush); chain state 8*E*C = %.0f MB" % (8e-6 * eng.E * C),
                   "threads_per_cta": info["threads"], "state_in_smem": info["state_in_smem"],
                   "smem_bytes_per_cta": info["smem_bytes"]},
        "slice_evals_per_resample": st["slice_evals"] / max(st["nevents"], 1),
        "redone_frac": st["n_redone"] / max(st["nevents"], 1),
    }
    return res, (net, ini, eng)


def convergence(ctx, net, ini, C, S):
    """R-hat and ESS that mean something.  (A) every chain gets a starting state of its own (bq_init_chains); the
    ensemble runs windows of 1000, 2000, 4000, 8000 sweeps (earlier windows are warm-up) until the largest split R-hat
    over the service parameters and per-queue mean waits of a window is below 1.05; ESS is then computed per chain on
    that window and SUMMED over the chains.  If no window passes, ess_valid is false and no number is given.
    (B) the same from the common reference-generated start, reported separately: chains that share a start also share
    whatever the sampler cannot move, so its R-hat flatters."""
    import numpy
    from bayes_qnet_b200 import GibbsEngine, parallel
    out = {}
    for label in ("per_chain_starts", "common_start"):
        eng = GibbsEngine(net, ini, n_chains=C, seed=777, device=ctx.local, chain_offset=ctx.rank * C)
        if label == "per_chain_starts":
            eng.init_chains(0.7)
            windows = [1000, 2000, 4000, 8000]
        else:
            eng.sweep(2000, "SLICE", "BAYES", sync=True)
            windows = [8000]
        swept, ok, diag, ms, history = (0 if label == "per_chain_starts" else 2000), False, None, 0.0, []
        for window in windows:
            eng.clear_traces()
            eng.sweep(window, "SLICE", "BAYES", trace=True, sync=True)
            ms = eng.last_sweep_ms()
            summ = parallel.gather_summaries_device(eng, ctx.world)
            diag = parallel.device_diagnostics(summ)
            rh = float(numpy.nanmax(diag["rhat"]))
            history.append([window, round(rh, 4)])
            if rh < 1.05:
                ok = True
                break
            swept += window
        ess_sum = diag["ess_per_chain_sum"]          # [K] sum over chains of the per-chain ESS of each summary
        window = history[-1][0]
        r = {"rhat_max": float(numpy.nanmax(diag["rhat"])), "rhat": [round(float(v), 3) for v in diag["rhat"]],
             "ess_valid": bool(ok), "warmup_sweeps_before_window": swept, "window_sweeps": window,
             "rhat_by_window": history, "chains": C * ctx.world}
        if ok:
            r["ess_per_s"] = float(numpy.min(ess_sum)) / (ms / 1e3)
            r["ess_per_sweep_per_chain"] = float(numpy.min(ess_sum)) / (window * C * ctx.world)
            r["slowest_summary"] = int(numpy.argmin(ess_sum))
        else:
            r["ess_per_s"] = None
        out[label] = r
        eng.close()
    out["summaries"] = "service parameters (Qnet.parameters order) then per-queue mean waits without the source queue"
    out["starts"] = ("per_chain_starts: one state per chain drawn on the device (bq_init_chains, factor exp(0.7 N(0,1))); "
                     "common_start: the reference-generated state replicated, 2000 warm-up sweeps")
    out["ess_note"] = ("min over the summaries of the SUM over chains of per-chain ESS (Geyer initial positive sequence) on the "
                       "first window whose split R-hat is below 1.05, divided by the device time of that window")
    ref = fullsize_record("config2_ess")
    if ref:
        out["reference"] = ref
    return out


def chain_sweep(ctx, counts=(1, 16, 148, 1024, 8192)):
    """BASELINE configs[4] in short: resamples/s of the 3-tier model at a few chain counts on this GPU (the full sweep
    1..65536 at 1/2/4/8 GPUs is `--workload config5`, profiles/r2_config5_*.json)."""
    out = []
    for c in counts:
        r, (net, ini, eng) = measure(ctx, "config2", 5000, c, 20, 2, 1, e2e_steps=1)
        eng.close()
        out.append({"chains_per_gpu": c, "value": r["value"], "e2e": r["e2e"]["value"], "ms_per_step": r["ms_per_step"]})
    return out


def run_config5(ctx, args):
    """Chain-count sweep: C total chains in {1, 2, 4, ..., 65536}; strong = split over the ranks, weak = per GPU."""
    rows = []
    c = 1
    while c <= 65536:
        for kind in ("strong", "weak"):
            per = c if kind == "weak" else (c + ctx.world - 1) // ctx.world
            if kind == "strong" and ctx.world == 1:
                continue
            if kind == "strong" and c < ctx.world:
                continue
            r, (net, ini, eng) = measure(ctx, "config2", 5000, per, 20, 2, 1, e2e_steps=1)
            eng.close()
            rows.append({"total_chains": per * ctx.world, "chains_per_gpu": per, "scaling": kind, "value": r["value"],
                         "e2e": r["e2e"]["value"], "ms_per_step": r["ms_per_step"]})
        c *= 2
    if ctx.rank == 0:
        print(json.dumps({"metric": "chain-task resamples/sec", "unit": "resamples/s", "n_gpus": ctx.world,
                          "workload": "config5: chain-count sweep on the 3-tier model, 20 sweeps per step, 2 timed steps",
                          "rows": rows}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS) + ["config5"],
                    help="BASELINE.json configs[1] (default, the headline), configs[0], [2], [3], or the chain sweep [4]")
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU")
    ap.add_argument("--tasks", type=int, default=None)
    ap.add_argument("--sweeps", type=int, default=None, help="Gibbs sweeps per step")
    ap.add_argument("--ref-tasks", type=int, default=None)
    ap.add_argument("--ref-sweeps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 1 / 3 / 4 sub-runs, convergence and chain sweep")
    ap.add_argument("--tight", action="store_true", help="BQ_FLAG_TIGHT bracket (not the default sampler)")
    args = ap.parse_args()
    wl = WORKLOADS["config2" if args.workload == "config5" else args.workload]
    for k in ("chains", "tasks", "sweeps", "ref_tasks"):
        if getattr(args, k) is None:
            setattr(args, k, wl[k])

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import __graft_entry__ as ge
    if int(os.environ.get("RANK", "0")) == 0:
        ge.build(quiet=True)
    from bayes_qnet_b200 import capi
    if not torch.cuda.is_available() or capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: bayes_qnet_b200 has no CPU fallback")
    ctx = Ctx(args)
    if args.workload == "config5":
        run_config5(ctx, args)
        if ctx.world > 1:
            ctx.dist.destroy_process_group()
        return

    sampler = ClockSampler(ctx.local)
    sampler.start()
    res, (net, ini, eng) = measure(ctx, args.workload, args.tasks, args.chains, args.sweeps, args.steps, args.warmup,
                                   tight=args.tight, sampler=sampler)
    eng.close()
    line = {"metric": "chain-task resamples/sec", "value": res["value"], "unit": "resamples/s", "n_gpus": ctx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": res["config"],
            "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "roofline": res["roofline"],
            "slice_evals_per_resample": res["slice_evals_per_resample"]}
    extras = not args.no_extras and args.workload == "config2"
    if extras:
        line["convergence"] = convergence(ctx, net, ini, min(args.chains, 256), args.sweeps)
        line["ess_per_s"] = line["convergence"]["per_chain_starts"]["ess_per_s"]
        line["ess_valid"] = line["convergence"]["per_chain_starts"]["ess_valid"]
        line["rhat_max"] = line["convergence"]["per_chain_starts"]["rhat_max"]
        line["ess_per_s_common_start"] = line["convergence"]["common_start"]["ess_per_s"]
    sampler.stop_flag = True
    line["clocks"] = sampler.summary()
    if extras:
        xs = []
        for name in ("config3", "config4", "config1"):
            w = WORKLOADS[name]
            r, (n2, i2, e2) = measure(ctx, name, w["tasks"], w["chains"], w["sweeps"], 2, 1, e2e_steps=1)
            e2.close()
            x = {"config": name, "workload": r["config"]["workload"], "initial_state": r["config"]["initial_state"],
                 "value": r["value"], "ms_per_step": r["ms_per_step"], "steps": 2, "warmup": 1, "e2e": r["e2e"],
                 "roofline": r["roofline"], "gpu_launches": r["gpu_launches"],
                 "slice_evals_per_resample": r["slice_evals_per_resample"], "redone_frac": r["redone_frac"]}
            if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
                cb = cpu_baseline_block(w, name, 2 if name != "config1" else 200, with_single=False, with_mh=(name != "config1"))
                if cb:
                    x["cpu_baseline"] = cb
            xs.append(x)
        line["extra_workloads"] = xs
        line["chain_sweep"] = chain_sweep(ctx)
    if ctx.rank == 0:
        if not args.no_cpu_baseline and ctx.world == 1:
            cb = cpu_baseline_block(wl, args.workload, args.ref_sweeps)
            if cb:
                line["cpu_baseline"] = cb
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
