#!/usr/bin/env python
"""bench.py -- noisy Surface-17 shots/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): PDD_model importance sampling, fixed-weight error subset
[0,0,2,1,1] placed uniformly over the two-round slot lists [24,36,48,156,48], s2 protocol
(calc_log_e_rate_arbitrary_error_model.py:355-457 of the reference): prep cycle, noisy cycle, second
noisy cycle for shots whose syndrome changed, lookup decode, logical read-out, failure tally.
A step = one batch of --shots shots per GPU.  Shots shard over ranks with no data-path collective; the
only exchange is one all-reduce of the counters at the end (inside the timed region).

`value`  : whole-job shots/s, device-timed (CUDA events on the library's launch stream), max over ranks.
`e2e`    : the same through the reference-shaped Python entry point
           surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model.s2_calculate_log_e_rate_error_subset
           with host inputs and host results, wall-clock around the calls.
`roofline`: the dominant kernel, k_cycle1 (one whole stabiliser cycle per shot and launch, the state in registers):
           algorithmic bytes = the 8 KiB block it reads (nothing in the preparation cycle, which starts from a basis
           state) + the 8 KiB block it writes per shot and cycle, divided by the summed CUDA-event time of its launches.  The kernel is bound by the FP64 pipe and shared memory, not by
           HBM; `roofline_dense` is the same gate stream as full-state sweeps (2 MiB read + 2 MiB written per shot
           and sweep, S17_DENSE=1), the HBM-bound configuration BASELINE.json's "% of HBM roofline per gate layer"
           refers to.
`value`, `e2e`, `roofline` are measured with S17_PREP_CACHE=0: every cycle of every shot -- the noiseless preparation
           cycle included -- is computed inside the timed region.
`with_memoisation`: the same steps in the library's default mode, where the preparation cycle (and the few event-free
           first cycles) of a shot is replayed from a cache of the kernel's own results for the 16 reachable outcome
           patterns; bit-identical records and counters.  Reported beside the headline, never as the headline.
`cpu_baseline` / --impl reference: the oracle port (one CPU sweep per gate, ProjectQ's execution model,
           python per-gate driver like the reference) on all host cores, bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

MODEL = "PDD_model"
ESET = [0, 0, 2, 1, 1]
LOCS2 = [24, 36, 48, 156, 48]
PROB_LISTS = [[1e-4], [1e-4], [5e-3], [1e-3], [5e-3]]
SWEEP_BYTES = 2 * (1 << 17) * 16          # read + write of one 2^17 complex128 state


def load_tables():
    from surface_17_iqt_b200 import api
    return (api.load_lookup_table("correction_table_depolarising.json"),
            api.load_lookup_table("correction_table_classical_bitflip.json"), api.load_error_models())


# ---------------------------------------------------------------------------------------------------
# CPU arm: oracle port, one process per host core
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    shots, seed = args
    import random
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import s17_oracle as so
    ct, bt, models = load_tables()
    orc = so.ShotOracle(models, ct, bt)
    t0 = time.perf_counter()
    failed, not0 = so.run_subset(orc, "s2", shots, MODEL, ESET, LOCS2, PROB_LISTS, rnd=random.Random(seed),
                                 meas_rng=random.Random(seed + 1))
    return shots, failed, not0, time.perf_counter() - t0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_arm(shots_per_core, cores=None, seed=0):
    import multiprocessing as mp
    from oracle import sv
    sv.lib()          # build liboracle_sv.so before forking if it is missing
    cores = cores or host_cores()
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(shots_per_core, seed + 17 * i) for i in range(cores)])
    wall = time.perf_counter() - t0
    shots = sum(r[0] for r in res)
    return {"shots": shots, "wall_s": wall, "shots_per_s": shots / wall, "cores": cores,
            "per_core_shots_per_s": sum(r[0] / r[3] for r in res) / cores,
            "failed": sum(r[1] for r in res), "not0": sum(r[2] for r in res)}


def cpu_arm_subprocess(shots_per_core, timeout=420):
    """The CPU arm in a fresh interpreter: forking a process that already holds a CUDA context can deadlock."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup",
                              "0", "--ref-shots-per-core", str(shots_per_core)], capture_output=True, text=True,
                             timeout=timeout, env=env).stdout
        line = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
        return line["cpu_baseline"]
    except Exception as exc:       # the baseline is a reported extra, never a reason to lose the GPU line
        return {"value": None, "unit": "shots/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: {}".format(exc)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_core = max(1, args.ref_shots_per_core)
    vals = []
    for _ in range(args.warmup):
        cpu_arm(max(1, per_core // 4), cores)
    t_all = time.perf_counter()
    for k in range(args.steps):
        vals.append(cpu_arm(per_core, cores, seed=1000 * k))
    wall = time.perf_counter() - t_all
    shots = sum(v["shots"] for v in vals)
    value = shots / sum(v["wall_s"] for v in vals)
    line = {
        "impl": "reference", "metric": "noisy Surface-17 shots/sec", "value": value, "unit": "shots/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": workload_config(per_core * cores, 1, None)["workload"],
                   "shots_per_step": per_core * cores,
                   "sample": "bounded sample of the same workload: {} shots per step through the CPU oracle port".format(per_core * cores)},
        "cpu_baseline": {"value": value, "unit": "shots/s", "cores": cores, "kind": "port",
                         "sample": "{} s2 shots per step on {} processes (oracle port: C sweep per gate, python driver)"
                         .format(per_core * cores, cores)},
        "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(shots, n_gpus, fusion):
    return {"workload": "Surface-17 {} s2 importance sampling, subset {} over {} two-round slots, uniform placement"
            .format(MODEL, ESET, LOCS2), "shots_per_step_per_gpu": shots, "fusion": fusion,
            "state": "complex128; between cycles the 2^9 amplitudes that can be non-zero after the ancilla resets "
                     "(8 KiB per shot) are resident in HBM; S17_DENSE=1 keeps and sweeps all 2^17 (roofline_dense)",
            "parallelism": "shots sharded x{}".format(n_gpus),
            "schedule": "the whole-cycle kernel (k_cycle1): a stabiliser cycle in the frame where its gates are "
                        "diagonal, X-type ancillas measured after timestep 4, one warp per shot; every cycle computed "
                        "per shot (S17_PREP_CACHE=0; the library's default memoised preparation is `with_memoisation`) -- "
                        "the preparation cycle and the first noisy cycle of a shot in one launch (block in registers in "
                        "between), the second noisy cycle in another; S17_DIAG=0 / S17_DENSE=1 select the general "
                        "amplitude paths",
            "l2": "per-step working set (shots x 8 KiB written by the first launch, read and written by the second "
                  "= {:.1f} GB) exceeds the 126 MB L2, no flush needed".format(shots * 8192 * 2.7 / 1e9)}


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region: NVML when importable (a sample per ~2 ms), nvidia-smi
    otherwise (a sample per ~0.2 s)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; NVML wants the physical one
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    n = self.nvml
                    sm = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                            n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
                    self.samples.append([str(sm), str(self.sm_max)] + ["Active" if r & b else "Not Active" for b in bits])
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                if self.nvml is not None:
                    self.nvml = None          # fall back to nvidia-smi
                    continue
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def note(msg):
    print("[bench] " + msg, file=sys.stderr, flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from surface_17_iqt_b200 import _lib as L
    from surface_17_iqt_b200 import api
    import surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model as drv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ct, bt, models = load_tables()
    B = args.shots
    # Headline arm: every cycle of every shot is computed inside the timed region (S17_PREP_CACHE=0).  The library's
    # default replays the noiseless preparation cycle (and event-free first cycles) from a cache of the kernel's own
    # results; that mode is measured too and reported as `with_memoisation` (--memo-headline swaps the two).
    headline_memo = bool(args.memo_headline)
    if headline_memo:
        os.environ.pop("S17_PREP_CACHE", None)
    else:
        os.environ["S17_PREP_CACHE"] = "0"
    sim = api.Simulation(MODEL, LOCS2, B, ct, bt, fusion=args.fusion, device=local, error_models=models)
    sim.backend.set_subset(ESET, LOCS2)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(k):
        first = (k * world + rank) * B
        c = sim.backend.run("s2", B, L.NOISE_SUBSET, seed=args.seed, first_shot_id=first)
        return c, sim.backend.timing()

    note("context ready, warming up")
    for k in range(args.warmup):
        step(k)
    torch.zeros(1, device="cuda")          # initialise torch's CUDA context outside the timed region
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    span_ms = gate_ms = 0.0
    sweeps = launches = gate_launches = failed = not0 = sweep_bytes = 0
    for k in range(args.steps):
        c, t = step(args.warmup + k)
        span_ms += t["span_ms"]
        gate_ms += t["gate_ms"]
        sweeps += c.sweeps
        sweep_bytes += c.sweep_bytes
        launches += t["launches"]
        gate_launches += t["gate_launches"]
        failed += c.failed
        not0 += c.not0
    note("timed steps done")
    tot = torch.tensor([failed, not0], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)           # the path's only collective: per-subset failure / s2 counts
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    sampler.stop_flag = True
    # max over ranks of the device-timed span; host wall clock reported beside it
    times = torch.tensor([span_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    span_ms_max, wall_ms_max = [float(v) for v in times.tolist()]
    total_shots = B * args.steps * world
    value = total_shots / (span_ms_max * 1e-3)

    # roofline of the gate sweep on this rank
    algo_bytes = sweep_bytes     # 8 KiB runs really read + written by the gate sweeps (known-zero regions are skipped)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo_bytes / (gate_ms * 1e-3) / 1e9 if gate_ms > 0 else 0.0

    # the same steps in the other mode (memoisation on unless --memo-headline): same shot ids, same counters
    sim.close()
    plain = None
    if rank == 0 and not args.no_plain:
        note("steps in the other memoisation mode")
        if headline_memo:
            os.environ["S17_PREP_CACHE"] = "0"
        else:
            os.environ.pop("S17_PREP_CACHE", None)
        try:
            psim = api.Simulation(MODEL, LOCS2, B, ct, bt, fusion=args.fusion, device=local, error_models=models)
            psim.backend.set_subset(ESET, LOCS2)
            for k in range(min(3, args.warmup)):
                psim.backend.run("s2", B, L.NOISE_SUBSET, seed=args.seed, first_shot_id=(k * world + rank) * B)
            p_span = p_gate = 0.0
            p_failed = p_not0 = 0
            p_steps = args.steps
            for k in range(p_steps):
                c = psim.backend.run("s2", B, L.NOISE_SUBSET, seed=args.seed,
                                     first_shot_id=((args.warmup + k) * world + rank) * B)
                t = psim.backend.timing()
                p_span, p_gate = p_span + t["span_ms"], p_gate + t["gate_ms"]
                p_failed, p_not0 = p_failed + c.failed, p_not0 + c.not0
            plain = {"value": B * p_steps / (p_span * 1e-3), "unit": "shots/s (this rank, device-timed)", "steps": p_steps,
                     "ms_per_step": p_span / p_steps, "gate_kernel_share_of_step": p_gate / p_span,
                     "counts_this_rank": {"failed": int(p_failed), "not0": int(p_not0), "shots": B * p_steps},
                     "mode": ("S17_PREP_CACHE=0: every cycle computed per shot" if headline_memo else
                              "library default: the noiseless preparation cycle and event-free first cycles are replayed "
                              "per shot from a cache of k_cycle1's own results (k_prep, k_quiet); bit-identical records")}
            psim.close()
            if world == 1:
                # the same through the drop-in entry point
                drv.VERBOSE = False
                drv.BATCH_SHOTS, drv.FUSION, drv.DEVICE = B, args.fusion, local
                drv.s2_calculate_log_e_rate_error_subset(B, ct, MODEL, ESET, LOCS2, bt, PROB_LISTS)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                for _ in range(args.steps):
                    drv.s2_calculate_log_e_rate_error_subset(B, ct, MODEL, ESET, LOCS2, bt, PROB_LISTS)
                torch.cuda.synchronize()
                plain["e2e"] = {"value": B * args.steps / (time.perf_counter() - t1), "unit": "shots/s"}
                drv.release()
        except Exception as exc:          # evidence only: never lose the main line over it
            plain = {"error": str(exc)[:200]}
        finally:
            if headline_memo:
                os.environ.pop("S17_PREP_CACHE", None)
            else:
                os.environ["S17_PREP_CACHE"] = "0"

    # the same gate stream as full-state sweeps (ProjectQ-equivalent traffic, one sweep per timestep): the HBM-bound
    # configuration the per-layer roofline fraction of BASELINE.json refers to
    dense = None
    note("dense-mode roofline run")
    if rank == 0 and not args.no_dense:
        os.environ["S17_DENSE"] = "1"
        try:
            D = args.dense_shots
            dsim = api.Simulation(MODEL, LOCS2, D, ct, bt, fusion=1, device=local, error_models=models,
                                  early_measure=False)
            dsim.backend.set_subset(ESET, LOCS2)
            dsim.backend.run("s2", D, L.NOISE_SUBSET, seed=args.seed)                 # warm-up
            g_ms = g_bytes = g_launch = 0
            c_ms = c_shots = c_launch = 0
            m_ms = m_launch = 0
            d_span = 0.0
            d_sweeps = 0
            for rep in range(args.dense_reps):
                dc = dsim.backend.run("s2", D, L.NOISE_SUBSET, seed=args.seed + 1 + rep)
                dt = dsim.backend.timing()
                g_ms += dt["gate_ms"]
                g_bytes += dc.sweep_bytes
                g_launch += dt["gate_launches"]
                d_sweeps += dc.sweeps
                d_span += dt["span_ms"]
                cl = dt["classes"]
                c_ms += cl["collapse"]["ms"]
                c_launch += cl["collapse"]["launches"]
                c_shots += cl["collapse"]["units"]
                m_ms += cl["sample"]["ms"]
                m_launch += cl["sample"]["launches"]
            d_ach = g_bytes / (g_ms * 1e-3) / 1e9
            # the measurement layer as this library runs it in full-state mode: the probability reduction is fused into
            # the gate sweep before the Measure (no extra pass over the state), k_measure samples from 256 weights
            # (2 KiB per shot), k_collapse reads the surviving 8 KiB block and writes the whole 2 MiB state (block at
            # ancilla value 0, zeros elsewhere).  SURVEY 8(d) counts reduce 2 MiB + collapse 4 MiB = 6 291 456 B for
            # ProjectQ's form of the same layer.
            COLLAPSE_MOVED = (1 << 17) * 16 + 8192
            c_ach = c_shots * COLLAPSE_MOVED / (c_ms * 1e-3) / 1e9 if c_ms > 0 else None
            dense = {"kernel": "k_layer2", "mode": "S17_DENSE=1, fusion=1 (8 full-state sweeps per cycle + explicit collapse)",
                     "achieved": d_ach, "peak": peak, "unit": "GB/s", "frac": d_ach / peak,
                     "shots_per_s": D * args.dense_reps / (d_span * 1e-3), "shots": D, "repetitions": args.dense_reps,
                     "bytes_per_shot": g_bytes / (D * args.dense_reps),
                     "gate_layer": {"kernel": "k_layer2", "bytes_per_shot_and_layer": SWEEP_BYTES, "layers": int(d_sweeps),
                                    "launches": int(g_launch), "ms": g_ms, "achieved": d_ach, "frac": d_ach / peak,
                                    "note": "layers = full-state sweeps counted by the kernel (one per shot and launch); "
                                            "bytes = layers x 4 194 304"},
                     "measure_layer": {"kernels": ["k_layer2 (fused probability reduction)", "k_measure", "k_collapse"],
                                       "collapse_launches": int(c_launch), "collapse_shots": int(c_shots), "collapse_ms": c_ms,
                                       "moved_bytes_per_shot": COLLAPSE_MOVED,
                                       "achieved": c_ach, "frac": (c_ach / peak) if c_ach else None,
                                       "sample_ms": m_ms, "sample_launches": int(m_launch),
                                       "survey_bytes_per_shot": 6291456,
                                       "survey_equivalent": (c_shots * 6291456 / ((c_ms + m_ms) * 1e-3) / 1e9 / peak) if c_ms > 0 else None,
                                       "note": "achieved/frac = bytes k_collapse really moves (8 KiB read + 2 MiB written per shot) "
                                               "over its CUDA-event time; survey_equivalent = SURVEY 8(d)'s 6 291 456 B per shot "
                                               "over k_measure + k_collapse time, as a fraction of peak -- above 1 means the fused "
                                               "reduction and the write-only collapse move fewer bytes than ProjectQ's form, not "
                                               "that the bus is faster"}}
            dsim.close()
        except Exception as exc:          # evidence only: never lose the main line over it
            dense = {"error": str(exc)[:200]}
        finally:
            os.environ.pop("S17_DENSE", None)

    # ProjectQ's own execution model on the GPU (the engine-level face, `s17_eng_*`): one full-state sweep per gate, and per
    # Measure a P(1) reduction over the state followed by collapse + renormalise -- exactly the layers SURVEY 8(d) counts
    # (gate layer 4 194 304 B per shot; measurement layer = reduce 2 097 152 B + collapse)
    engine = None
    note("engine-face roofline run")
    if rank == 0 and not args.no_dense:
        try:
            from surface_17_iqt_b200 import engine as eng_mod
            E = args.engine_shots
            eb = eng_mod.EngineBackend(17, E, local, seed=args.seed)
            ry, rxx = eng_mod.rotation_matrix("Ry", 0.5 * 3.141592653589793), eng_mod.rotation_matrix("Rxx", 0.5 * 3.141592653589793)

            def one_cycle():                          # a cycle-shaped gate stream: 54 sweeps, then All(Measure) | ancilla
                for k in range(30):
                    eb.gate1(k % 9, ry)
                for k in range(24):
                    eb.gate2(k % 9, 9 + (k % 8), rxx)
                for a in range(8):
                    eb.measure(9 + a)
            one_cycle()
            eb.timing(clear=True)
            for _ in range(args.dense_reps):
                eb.reset()
                one_cycle()
            et = eb.timing()
            eb.close()

            def gbs(nbytes, calls, ms):
                return nbytes * calls / (ms * 1e-3) / 1e9 if ms > 0 else None
            g = gbs(et["gate_bytes_per_call"], et["gate_calls"], et["gate_ms"])
            r = gbs(et["reduce_bytes_per_call"], et["reduce_calls"], et["reduce_ms"])
            c = gbs(et["collapse_bytes_per_call"], et["collapse_calls"], et["collapse_ms"])
            meas_ms = et["reduce_ms"] + et["collapse_ms"]
            engine = {"mode": "s17_eng_*: one sweep per gate over {} state vectors (ProjectQ's execution model, no fusion)".format(E),
                      "shots": E, "repetitions": args.dense_reps, "peak": peak, "unit": "GB/s",
                      "gate_layer": {"kernels": ["k_eng_gate1", "k_eng_gate2"], "bytes_per_shot": SWEEP_BYTES,
                                     "launches": int(et["gate_calls"]), "ms": et["gate_ms"], "achieved": g,
                                     "frac": g / peak if g else None},
                      "measure_layer": {"kernels": ["k_eng_prob", "k_eng_sample", "k_eng_collapse"],
                                        "reduce": {"bytes_per_shot": SWEEP_BYTES // 2, "launches": int(et["reduce_calls"]),
                                                   "ms": et["reduce_ms"], "achieved": r, "frac": r / peak if r else None},
                                        "collapse": {"bytes_per_shot": 3 * SWEEP_BYTES // 4, "launches": int(et["collapse_calls"]),
                                                     "ms": et["collapse_ms"], "achieved": c, "frac": c / peak if c else None,
                                                     "note": "reads the surviving half, writes both halves"},
                                        "survey_bytes_per_shot": 6291456,
                                        "survey_equivalent": (6291456 * E * et["reduce_calls"] / (meas_ms * 1e-3) / 1e9 / peak)
                                        if meas_ms > 0 else None}}
        except Exception as exc:          # evidence only: never lose the main line over it
            engine = {"error": str(exc)[:200]}

    # e2e through the reference-shaped entry point: host args in, host tuple out, wall clock
    note("e2e through the drop-in entry point")
    drv.VERBOSE = False
    drv.BATCH_SHOTS, drv.FUSION, drv.DEVICE = B, args.fusion, local
    drv.s2_calculate_log_e_rate_error_subset(B, ct, MODEL, ESET, LOCS2, bt, PROB_LISTS)      # warm: builds the context
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        r = drv.s2_calculate_log_e_rate_error_subset(B * world, ct, MODEL, ESET, LOCS2, bt, PROB_LISTS)
    barrier()
    e2e_s = time.perf_counter() - t1
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / float(e2e_t.item())
    drv.release()

    # the binding unit of k_cycle1 is the FP64 pipe, not HBM: instruction counts and DRAM traffic per shot-cycle come
    # from the committed ncu capture of this kernel (profiles/k_cycle1_counts.json, written by
    # profiles/tools/ncu_counts.py from an `ncu --set full` report), the time is this run's
    counts = {}
    try:
        counts = json.load(open(os.path.join(REPO, "profiles", "k_cycle1_counts.json")))
    except Exception:
        pass
    sm_hz = (sampler.summary().get("sm_mhz") or 1965) * 1e6
    n_sms = torch.cuda.get_device_properties(local).multi_processor_count
    fp64 = None
    if counts.get("fp64_warp_instr_per_shot_cycle") and gate_ms > 0:
        # 64 FP64 lanes per SM and clock: one warp instruction occupies an SM's FP64 pipes for half a clock
        floor_ms = 1e3 * counts["fp64_warp_instr_per_shot_cycle"] * sweeps * 0.5 / (n_sms * sm_hz)
        fp64 = {"warp_instr_per_shot_cycle": counts["fp64_warp_instr_per_shot_cycle"],
                "all_warp_instr_per_shot_cycle": counts.get("warp_instr_per_shot_cycle"),
                "floor_ms": floor_ms, "kernel_ms": gate_ms, "frac": floor_ms / gate_ms,
                "peak": "64 FP64 lanes per SM and clock x {} SMs x {:.0f} MHz".format(n_sms, sm_hz / 1e6),
                "source": counts.get("source")}
    issue = None
    if counts.get("warp_instr_per_shot_cycle") and gate_ms > 0:
        # four schedulers per SM, one warp instruction each per clock
        i_floor_ms = 1e3 * counts["warp_instr_per_shot_cycle"] * sweeps / (n_sms * 4 * sm_hz)
        issue = {"warp_instr_per_shot_cycle": counts["warp_instr_per_shot_cycle"], "floor_ms": i_floor_ms,
                 "kernel_ms": gate_ms, "frac": i_floor_ms / gate_ms,
                 "peak": "4 warp instructions per SM and clock x {} SMs x {:.0f} MHz".format(n_sms, sm_hz / 1e6)}
    traffic = None
    if counts.get("dram_bytes_per_shot_cycle") and gate_launches:
        traffic = counts["dram_bytes_per_shot_cycle"] * sweeps / gate_launches
    roofline = {"bound": "hbm", "kernel": "k_cycle1", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                "traffic": traffic,
                "traffic_source": counts.get("source"),
                "binding_unit": "fp64 pipe / instruction issue", "fp64": fp64, "issue": issue,
                "note": "`achieved` = bytes the cycle kernel really moves (8 KiB runs counted by the kernel) over its "
                        "time.  The launch of the first noisy cycle computes the shot's preparation cycle first and "
                        "keeps the block in registers (8 KiB written per shot for two cycles), the second noisy cycle "
                        "reads and writes 8 KiB per shot; everything else lives in registers / shared memory, so the "
                        "HBM fraction is low by construction -- it FELL from 0.46 to this value when the preparation "
                        "cycle stopped writing and re-reading its block, while the step got 8 % faster.  What binds "
                        "is the SM: `fp64` = counted FP64 warp instructions "
                        "x 0.5 clock over the kernel time (the floor of the arithmetic), `issue` = all counted warp "
                        "instructions over four issue slots per SM and clock (the floor of the instruction stream as "
                        "written: 1172 of its instructions are FP64, the rest is addressing, selects, shuffles and "
                        "control).  Round 1: 4280 instructions per shot-cycle (1810 FP64); counts of the committed ncu "
                        "capture in `profiles/k_cycle1_counts.json`.  SURVEY 8(d)'s per-layer bytes (4 194 304 per gate "
                        "layer, 6 291 456 per measurement layer) apply to the full-state configuration: `roofline_dense`",
                "launches": int(gate_launches), "sweeps": int(sweeps),
                "sweeps_per_shot": sweeps / (B * args.steps),
                "bytes_per_shot": sweep_bytes / (B * args.steps), "dense_bytes_per_sweep": SWEEP_BYTES,
                "gate_kernel_share_of_step": gate_ms / span_ms if span_ms else None}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            note("cpu baseline (subprocess)")
            cpu = cpu_arm_subprocess(args.cpu_shots_per_core)
        line = {
            "metric": "noisy Surface-17 shots/sec", "value": value, "unit": "shots/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": span_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)",
            "data": "synthetic", "config": workload_config(B, world, args.fusion),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "shots/s",
                    "h2d_bytes_per_step": 2048 * 8 + 3 * 4 * len(ESET) + 64, "d2h_bytes_per_step": 64 + 8 * 8,
                    "api": "s2_calculate_log_e_rate_error_subset(num_runs, ...) -> (rate, not0_rate, failed, not0)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "counts": {"failed": int(tot[0]), "not0": int(tot[1]), "shots": total_shots},
        }
        if plain is not None:
            line["without_memoisation" if headline_memo else "with_memoisation"] = plain
        line["memoisation_in_headline"] = headline_memo
        if dense is not None:
            line["roofline_dense"] = dense
        if engine is not None:
            line["roofline_engine"] = engine
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# run-til-fail workload (BASELINE config 5): --workload rtf
# ---------------------------------------------------------------------------------------------------
def run_gpu_rtf(args):
    """calculate_log_e_rate's loop (CL:33-148) at the size of BASELINE config 5: Dress_model (leakage) with true Bernoulli
    noise p_set = [12 x p/10, 18 x p/10, 24 x p, 78 x p_leak = 1e-4, 24 x 0], every GPU a fixed budget of rounds per step
    (10^8 rounds over 8 GPUs = 1.25e7 per GPU).  A step = one run-til-fail job of --shots slots per GPU that stops when
    its round budget is spent (a slot whose run fails takes the next run id).  value = rounds/s over all GPUs,
    device-timed span (max over ranks); e2e = the same through calculate_log_e_rate(..., round_budget=...) with host
    lists out.  The runs shard over the ranks; the only exchange is the gather of the per-run lists at the end."""
    import torch
    import torch.distributed as dist
    from surface_17_iqt_b200 import _lib as L
    from surface_17_iqt_b200 import api
    import surface_17_iqt_b200.calc_log_e_rate_arbitrary_error_model as drv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ct, bt, models = load_tables()
    p = args.p
    lens = [12, 18, 24, 78, 24]
    p_set = [[p / 10] * 12, [p / 10] * 18, [p] * 24, [1e-4] * 78, [0.0] * 24]
    B, budget = args.shots, args.rtf_rounds
    sim = api.Simulation("Dress_model", lens, B, ct, bt, fusion=args.fusion, device=local, error_models=models)
    sim.backend.set_slot_probs(p_set)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(k):
        c = sim.backend.run("rtf", B, L.NOISE_PROB, seed=args.seed + k, first_shot_id=rank * B, rtf_runs=B * 2,
                            leak_reset="run", rtf_round_budget=budget)
        return c, sim.backend.timing()

    for k in range(args.warmup):
        step(k)
    torch.zeros(1, device="cuda")
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    span_ms = gate_ms = 0.0
    rounds = launches = gate_launches = sweeps = failed_runs = 0
    for k in range(args.steps):
        c, t = step(args.warmup + k)
        span_ms += t["span_ms"]
        gate_ms += t["gate_ms"]
        rounds += c.rounds
        sweeps += c.sweeps
        failed_runs += c.failed
        launches += t["launches"]
        gate_launches += t["gate_launches"]
    tot = torch.tensor([rounds, failed_runs, sweeps], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    sampler.stop_flag = True
    times = torch.tensor([span_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    span_ms_max, wall_ms_max = [float(v) for v in times.tolist()]
    total_rounds = int(tot[0].item())
    value = total_rounds / (span_ms_max * 1e-3)
    sim.close()

    # e2e: the drop-in with host results (per-run lists gathered to every rank, no fit, no file)
    drv.VERBOSE = False
    drv.BATCH_SHOTS, drv.FUSION, drv.DEVICE = B, args.fusion, local
    e_model, e_probs = "Dress_model", dict(zip(models["Dress_model"]["probabilities"], p_set))
    drv.calculate_log_e_rate(B * 2 * world, "bench", ct, e_model, e_probs, bt, save=False, fit=False,
                             round_budget=budget)
    barrier()
    t1 = time.perf_counter()
    e_rounds = 0
    for _ in range(args.steps):
        drv.calculate_log_e_rate(B * 2 * world, "bench", ct, e_model, e_probs, bt, save=False, fit=False,
                                 round_budget=budget)
        e_rounds += drv.LAST_TIMING["rounds"]
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    drv.release()
    if rank == 0:
        line = {
            "metric": "noisy Surface-17 run-til-fail rounds/sec", "value": value, "unit": "rounds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": span_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)",
            "data": "synthetic",
            "config": {"workload": "Surface-17 Dress_model (leakage) run-til-fail, true Bernoulli noise p = {:g}, "
                                   "p_leak = 1e-4 (BASELINE config 5)".format(p),
                       "slots_per_gpu": B, "round_budget_per_gpu_and_step": budget, "leak_reset": "run",
                       "parallelism": "runs sharded x{}".format(world),
                       "l2": "per-round working set (slots x 16 KiB) exceeds the 126 MB L2, no flush needed"},
            "wall_ms_per_step": wall_ms_max / args.steps, "clocks": sampler.summary(),
            "e2e": {"value": e_rounds / float(e2e_t.item()), "unit": "rounds/s",
                    "h2d_bytes_per_step": 2048 * 8 + 64, "d2h_bytes_per_step": 2 * 4 * B * 2,
                    "api": "calculate_log_e_rate(num_runs, ..., round_budget=...) -> rounds_til_fail_list, syndrome_b_measured_list"},
            "gpu_launches": int(launches),
            "rounds": total_rounds, "runs_failed": int(tot[1].item()),
            "cycles_per_round": int(tot[2].item()) / max(1, total_rounds),
            "gate_kernel_share_of_step": gate_ms / span_ms if span_ms else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shots", type=int, default=int(os.environ.get("S17_BENCH_SHOTS", "524288")), help="shots per GPU per step")
    ap.add_argument("--fusion", type=int, default=2)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--workload", default="subset", choices=["subset", "rtf"],
                    help="subset: BASELINE config 2 (the headline); rtf: run-til-fail, BASELINE config 5")
    ap.add_argument("--p", type=float, default=1e-3, help="rtf: physical error rate")
    ap.add_argument("--rtf-rounds", type=int, default=12500000, help="rtf: round budget per GPU and step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dense", action="store_true")
    ap.add_argument("--no-plain", action="store_true")
    ap.add_argument("--memo-headline", action="store_true",
                    help="time the library default (memoised preparation cycle) as value / e2e and the full computation as the extra")
    ap.add_argument("--dense-shots", type=int, default=16384)
    ap.add_argument("--dense-reps", type=int, default=5)
    ap.add_argument("--engine-shots", type=int, default=8192)
    ap.add_argument("--cpu-shots-per-core", type=int, default=40)
    ap.add_argument("--ref-shots-per-core", type=int, default=40)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "rtf":
        run_gpu_rtf(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
