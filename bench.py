#!/usr/bin/env python3
"""Headline benchmark: SIS particles/sec on the README model (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (N>1 under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU SIS (restated oracle)

One "step" = one complete inference pass: every particle of the workload is generated, weighted and
folded into the posterior estimators (pilot + particle kernel + chunk merge [+ NCCL all-gather]).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "SIS particles/sec (device-timed)"
MODEL = "gaussian_unknown_mean"
OBS = [3.0, 4.0]
SEED = 0x5EED



def sass_budget():
    """Per-particle instruction budget of k_sis_fused<gaussian_unknown_mean_model, 1>, counted in the SASS of the particle
    loop when the library was built (cpprob_b200/build.py, tools/sass_mix.py).  `hot` excludes the call set-up that only
    the ziggurat's slow draws (0.06 %) execute; the slow path itself is not counted at all, so the figures are lower
    bounds of the work done.  "flop" counts DFMA as 2, DADD / DMUL as 1, DSETP as 0.  An FP64-pipe instruction holds the
    sub-partition's issue port for two cycles (measured), every other instruction for one:
    issue slots = 2 F + O."""
    with open(os.path.join(ROOT, "cpprob_b200", "lib", "sass_budget.json")) as f:
        b = json.load(f)
    per = b["particles_per_trip"]
    hot = b["total"] - b.get("cold", 0)
    return {"fp64_instr": b["fp64"] / per, "flop": (2 * b["dfma"] + b["dadd"] + b["dmul"]) / per, "loop_instr": hot / per,
            "issue_slots": (hot + b["fp64"]) / per}


def measured_peaks():
    """HBM copy bandwidth the rooflines are quoted against: MEASURED_PEAKS.json (driver-written) or the profiling
    recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def kernel_costs():
    """Per-kernel figures taken from the committed ncu captures (profiles/kernel_costs.json, written by
    tools/summarise_profiles.py): DRAM bytes and executed instructions per particle.  Absent entries read as None."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_costs.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML every few ms while the timed region runs
    (nvidia-smi -lms cannot start fast enough for a region of a few hundred ms)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index = index
        self.sm, self.power, self.mask = [], [], 0
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                self.mask |= nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=2)
        reasons = sorted(k for k, bit in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU SIS (oracle restatement, faithful flavour) on all host cores
# ---------------------------------------------------------------------------------------------------
def oracle_binary():
    exe = os.path.join(ROOT, "oracle", "oracle_sis")
    if not os.path.exists(exe):
        subprocess.run(["make"], cwd=os.path.join(ROOT, "oracle"), check=True, stdout=subprocess.DEVNULL)
    return exe


def scratch_dir():
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()


REF_STATS_PRINTER = os.path.join(ROOT, "oracle", "_ref", "ref_stats_printer")
REF_SIS = os.path.join(ROOT, "oracle", "_ref", "ref_sis")


def have_reference_loop():
    """oracle/_ref/ref_sis: the reference's OWN cpprob::inference(StateType::sis, ...) — its state.cpp, trace.cpp, utils.cpp,
    models ... compiled unmodified (oracle/Makefile, oracle/ref_sis.cpp); built where /root/reference exists, travels as a file."""
    return os.path.exists(REF_SIS) and os.path.exists(REF_STATS_PRINTER)


def cpu_kind():
    return "reference" if have_reference_loop() else "port"


def cpu_what():
    if have_reference_loop():
        return ("the reference's own cpprob::inference(StateType::sis) + StatsPrinter (oracle/_ref/ref_sis and ref_stats_printer: its "
                "state.cpp, trace.cpp, utils.cpp, models and post-processing headers compiled unmodified with -O2; Boost.Random's "
                "samplers stood in for by the standard library's, FlatBuffers / ZeroMQ by name-only stubs)")
    return "restated cpprob::inference (3 file appends per trace) + " + stats_printer_kind()


def stats_printer_kind():
    return ("reference's own StatsPrinter (oracle/_ref, compiled unmodified from the reference's headers)" if os.path.exists(REF_STATS_PRINTER)
            else "restated StatsPrinter")


def cpu_sis_step(n_procs, particles_each, flavour="faithful"):
    """One bounded sample: n_procs independent single-threaded runs (the reference is single-threaded,
    SURVEY.md §5), each writing its own posterior files and post-processing them — with the reference's own StatsPrinter
    where oracle/_ref was built.  Returns wall seconds."""
    exe = oracle_binary()
    env = dict(os.environ)
    if os.path.exists(REF_STATS_PRINTER):
        env["ORACLE_STATS_PRINTER"] = REF_STATS_PRINTER
    with tempfile.TemporaryDirectory(dir=scratch_dir()) as d:
        t0 = time.perf_counter()
        if flavour == "faithful" and have_reference_loop():
            # the README program of the reference: inference(...) then `std::cout << StatsPrinter{outfile}` (README.md:102-116)
            procs = [subprocess.Popen(["/bin/sh", "-c", f'"{REF_SIS}" {MODEL} {particles_each} "{d}/p{i}" - 3 4 && "{REF_STATS_PRINTER}" "{d}/p{i}"'],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(n_procs)]
        else:
            procs = [subprocess.Popen([exe, MODEL, str(particles_each), os.path.join(d, f"p{i}"), flavour, "3", "4"],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env) for i in range(n_procs)]
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("oracle_sis failed")
        return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    each = args.ref_particles
    for _ in range(args.warmup):
        cpu_sis_step(cores, each)
    times = [cpu_sis_step(cores, each) for _ in range(args.steps)]
    total = sum(times)
    value = cores * each * args.steps / total
    sample = f"{cores} processes x {each} particles per step: {cpu_what()}, files on {scratch_dir()}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particles/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "gaussian_unknown_mean x=(3,4), reference CPU SIS: " + cpu_what(), "particles_per_step": cores * each},
        "cpu_baseline": {"value": value, "unit": "particles/s", "cores": cores, "kind": cpu_kind(), "sample": sample},
        "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class _DeviceArray:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cpprob_b200 import Engine

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SIS engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    engine = Engine(device=local_rank, seed=SEED)
    per_gpu = args.particles
    total = per_gpu * world
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        # The library's own communicator (cpprob_sis_comm_init): rank 0 draws the NCCL id, torch.distributed is only the
        # out-of-band transport for its 128 bytes.  From here on the data path is one library call per inference
        # (cpprob_sis_run_dist): shard kernels, ONE ncclAllGather on the engine's stream, in-place merge, one host sync.
        from cpprob_b200 import capi
        id_t = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            id_t.copy_(torch.frombuffer(bytearray(capi.comm_get_id()), dtype=torch.uint8))
        dist.broadcast(id_t, 0)
        engine.comm_init(bytes(id_t.cpu().numpy().tobytes()), rank, world)

    from cpprob_b200.capi import stats_to_dict

    def step(n_total=None):
        """One inference pass of `n_total` particles over `world` GPUs.  Returns (stats, kernel_ms, launches)."""
        n_total = total if n_total is None else n_total
        st = engine.run(MODEL, OBS, n_total) if world == 1 else engine.run_dist(MODEL, OBS, n_total)
        return st, st["device_ms"], st["kernel_launches"]

    # The device-timed loops call the C ABI with everything marshalled beforehand (Engine.prepared): the bracket holds
    # cpprob_sis_run / cpprob_sis_run_dist and nothing of the Python wrapper (argument conversion, result dict), which is a
    # test harness, not the product.  The e2e figure below goes through the ordinary wrapper call.
    call_main = engine.prepared(MODEL, OBS, total, dist=world > 1)
    call_strong = engine.prepared(MODEL, OBS, args.strong_particles, dist=world > 1)

    for _ in range(args.warmup):
        step()
        flush.fill_(1)
    sampler = ClockSampler(local_rank)      # every rank watches its own GPU: the step is as slow as the slowest of them
    sampler.start()
    time.sleep(0.02)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kernel_ms_total, particle_ms_total, launches_total, last = 0.0, 0.0, 0, None
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                 # L2 flush between steps (outside the timed brackets)
        barrier()                             # every rank enters the step together (N > 1): the bracket times the step, not
        ev0[i].record()                       # the ranks' drift through the untimed flush
        raw = call_main()
        ev1[i].record()
        kernel_ms_total += raw.device_ms
        particle_ms_total += raw.particle_ms
        launches_total += raw.kernel_launches
    last = stats_to_dict(raw)
    barrier()
    clocks = sampler.stop()
    step_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    t = torch.tensor([step_ms, kernel_ms_total], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        # what every rank saw: its own device time per step (pilot + particle kernel + fold + merge, the merge including the
        # wait for the slowest peer's rows) and its own GPU's clocks — B200s of one box differ by a few per cent
        mine = {"rank": rank, "particle_pass_ms_per_step": particle_ms_total / args.steps, "kernel_ms_per_step": kernel_ms_total / args.steps,
                "step_ms": step_ms / args.steps,
                "sm_mhz": clocks["sm_mhz"], "reasons": clocks["reasons"], "power_w_max": clocks["power_w_max"]}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kernel_ms_total = t.tolist()

    # end to end through the public call with host buffers (cpprob_sis_run / run_shard+merge): wall clock,
    # includes the H2D of the observations, every launch and sync, and the D2H of the estimators
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        st_e2e, _, _ = step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = te.item()

    # Strong scaling (BASELINE.json configs[1] reads "1e9 particles on 1/2/4/8 B200"): the SAME 1e9 particles split over
    # the N GPUs, device-timed like the main figure; and a fingerprint of the merged sums of a fixed 2^30-particle run,
    # which must be the same string for every N (results are bit-identical for any GPU count).
    import hashlib
    strong_n = args.strong_particles
    for _ in range(3):
        step(strong_n)
    barrier()
    s_ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    s_ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        barrier()
        s_ev0[i].record()
        call_strong()
        s_ev1[i].record()
    barrier()
    ts = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(s_ev0, s_ev1))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    strong_ms = ts.item() / args.steps
    fp_st, _, _ = step(1 << 30)
    sums_sha = hashlib.sha256(fp_st["sums"].tobytes()).hexdigest()[:16]

    if rank == 0:
        value = total * args.steps / (step_ms * 1e-3)
        n_cols = last["n_cols"]
        line = {
            "metric": METRIC, "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"gaussian_unknown_mean x=(3,4), {per_gpu:.3g} particles per GPU (BASELINE.json configs[1])",
                       "particles_per_gpu": per_gpu, "total_particles": total, "seed": SEED, "parallelism": f"particle shards x{world}",
                       "l2": "256 MiB written between steps (flush); the path's only input is 16 B of observations"},
            "e2e": {"value": total * e2e_steps / e2e_s, "unit": "particles/s", "h2d_bytes_per_step": 8 * len(OBS),
                    "d2h_bytes_per_step": 24 + 8 * n_cols},
            "gpu_launches": int(launches_total),
            "clocks": clocks,
            "posterior": {"mean": float(last["real_mean"][0]), "variance": float(last["real_var"][0]), "log_evidence": last["log_evidence"],
                          "ess": last["ess"], "analytic_mean": 2.323529411764706, "analytic_variance": 1.0588235294117647,
                          "mean_err_in_mc_se": abs(float(last["real_mean"][0]) - 2.323529411764706) / (1.2973 / math.sqrt(total))},
            "kernel_ms_per_step": kernel_ms_total / args.steps,
            "strong_scaling": {"total_particles": strong_n, "ms_per_step": strong_ms, "value": strong_n / (strong_ms * 1e-3), "unit": "particles/s",
                               "note": "the same total split over the N GPUs (fixed total work); the main `value` is weak scaling"},
            "sums_sha": {"particles": 1 << 30, "sha256_16": sums_sha,
                         "note": "SHA-256 (first 16 hex digits) of the merged estimator sums of a 2^30-particle run: identical for every --gpus N"},
            "collective": "none (1 GPU)" if world == 1 else (
                "partial rows pushed into every peer's gather buffer over NVLink by the producing rank's own kernel (k_push_rows), epoch flags, "
                "merge kernel waits on the flags: no collective kernel, one host sync per inference (cpprob_sis_run_dist)"
                if engine.comm_exchange() == "peer" else "one ncclAllGather per inference inside libcpprob_sis.so (cpprob_sis_run_dist)"),
            "exchange": engine.comm_exchange(),
            "per_rank": per_rank,
        }
        if world == 1:
            peak_tflops, est_mhz = engine.dfma_peak()
            budget = sass_budget()
            k_s = kernel_ms_total * 1e-3 / args.steps
            achieved = budget["flop"] * per_gpu / k_s / 1e12
            slots = budget["issue_slots"] * per_gpu / k_s / 1e12        # thread-level issue slots per second, in T/s
            line["roofline"] = {
                "bound": "issue", "achieved": slots, "peak": peak_tflops, "unit": "Tslot/s", "frac": slots / peak_tflops,
                "traffic": 206592,
                "note": "dominant kernel k_sis_fused streams nothing from HBM and has no dense contraction: it is bound by the SM "
                        "sub-partitions' issue ports, where an FP64-pipe instruction holds the port for 2 cycles and any other for 1 "
                        "(cpprob_sis_probe_issue, DESIGN.md section 5).  unit = thread-level issue slots per second; peak = the DFMA "
                        "chain micro-benchmark of this run (one DFMA = 2 flop = 2 slots, so the figure equals its TFLOP/s; "
                        "MEASURED_PEAKS.json has no FP64 entry); achieved = (2 F + O) slots per particle from the SASS of the "
                        "particle loop x particles/s, slow ziggurat draws not counted.  traffic = dram bytes of one launch from "
                        "ncu --set full (profiles/): 207 KB read (the shared-memory tables of 148 CTAs), 0 written",
                "issue_slots_per_particle": budget["issue_slots"], "loop_instr_per_particle": budget["loop_instr"],
                "fp64_pipe_instr_per_particle": budget["fp64_instr"], "flop_per_particle": budget["flop"],
                "fp64_tflops": achieved, "fp64_frac_of_dfma_peak": achieved / peak_tflops,
                "fp64_pipe_util": budget["fp64_instr"] * per_gpu / k_s / (peak_tflops * 1e12 / 2.0),
                "dfma_peak_sm_mhz_equiv": est_mhz,
            }
            line["roofline"]["traffic"] = kernel_costs().get("C2", {}).get("dram_bytes_per_launch", line["roofline"]["traffic"])
            line["cpu_baseline"] = cpu_baseline(args)
            if not args.no_configs:
                line["configs"] = secondary_configs(engine, args)
                cf = line["configs"]["C2_files"]
                cb = line["cpu_baseline"]
                # the three ratios against the CPU arm of this run (1 thread; the driver computes the all-core ones itself)
                line["e2e_files"] = {"value": cf["value"], "unit": "particles/s", "particles": cf["particles"],
                                     "what": "cpprob_sis_infer_to_files, wall clock, files complete: like for like with the reference's inference()"}
                line["vs_cpu_1thread"] = {"files_vs_faithful": cf["value"] / cb["value"], "files_vs_buffered": cf["value"] / cb["fast_flavour_value"],
                                          "estimators_only_vs_faithful": line["e2e"]["value"] / cb["value"],
                                          "note": ("faithful = the reference's own SIS loop + StatsPrinter (oracle/_ref, one thread)" if cpu_kind() == "reference"
                                                   else "faithful = the restated loop, 3 file appends per trace as the reference") +
                                                  "; buffered = the restated loop with one ofstream per file kept open; "
                                                  "estimators only = cpprob_sis_run (CPPROB_SIS_EMIT=none), which writes no posterior file"}
        emit(line)
    engine.close()
    if world > 1:
        dist.destroy_process_group()


def timed_runs(fn, reps=3, warm=1):
    """best-of-`reps` of fn() -> (stats, wall seconds); the engine's own CUDA events give stats['device_ms']"""
    best = None
    for i in range(warm + reps):
        t0 = time.perf_counter()
        st = fn()
        wall = time.perf_counter() - t0
        if i >= warm and (best is None or st["device_ms"] < best[0]["device_ms"]):
            best = (st, wall)
    return best


def secondary_configs(engine, args):
    """BASELINE.json configs[2..4] and the file-emitting form of configs[1], on one GPU.  Each entry: particles/s from the
    engine's CUDA events around its particle + reduction kernels (best of 3 after a warm-up; every run regenerates
    all particles, and the row buffers of C3-C5 are far larger than L2), the same through the public call with host
    buffers (wall clock), and the roofline of the path that ran."""
    import analytic
    g = analytic.golden()
    hbm, hbm_src = measured_peaks()
    costs = kernel_costs()
    out = {}

    def entry(label, model, obs, n, note):
        st, wall = timed_runs(lambda: engine.run(model, obs, n))
        k_s = st["device_ms"] * 1e-3
        row_bytes = 8 * st["n_real"] + 4 * st["n_int"] + 16                  # SoA rows + log_w + w, written then read once
        e = {"workload": note, "particles": n, "value": n / k_s, "unit": "particles/s", "device_ms": st["device_ms"],
             "e2e_value": n / wall, "launches": int(st["kernel_launches"]), "passes": int(st["passes"]), "path": st.get("path", "rows"),
             "steps_per_s": n * max(st["n_real"], st["n_int"]) / k_s, "ess": st["ess"], "log_evidence": st["log_evidence"]}
        if e["path"] == "rows":
            algo = 2.0 * row_bytes * n                                       # each row element written once, read once
            e["roofline"] = {"bound": "hbm", "achieved": algo / k_s / 1e9, "peak": hbm, "unit": "GB/s", "frac": algo / k_s / 1e9 / hbm,
                             "peak_source": hbm_src, "algorithmic_bytes_per_particle": 2 * row_bytes,
                             "traffic": costs.get(label, {}).get("dram_bytes_per_particle")}
        else:
            c = costs.get(label, {})
            slots = c.get("issue_slots_per_particle")
            peak_t, _ = engine.dfma_peak()
            e["roofline"] = {"bound": "issue", "achieved": None if slots is None else slots * n / k_s / 1e12, "peak": peak_t, "unit": "Tslot/s",
                             "frac": None if slots is None else slots * n / k_s / 1e12 / peak_t,
                             "issue_slots_per_particle": slots, "traffic": c.get("dram_bytes_per_particle"),
                             "note": "no trace row touches HBM on this path; slots per particle from the committed ncu capture "
                                     "(executed warp instructions + FP64-pipe instructions, per particle), peak = DFMA probe of this run"}
        out[label] = e
        return e

    entry("C3", "linear_gaussian_1d", g["obs_linear_gaussian_32"], args.c3_particles,
          "linear_gaussian_1d, 32 synthetic observations (BASELINE.json configs[2]), per-(id,k) mean/variance")
    entry("C4", "hmm", g["obs_hmm_64"], args.c4_particles,
          "hmm, 3 states, 64-step synthetic observation sequence (configs[3]), per-address histograms / MAP")
    entry("C5_estimators", "hmm", g["obs_hmm_1000"], args.c5_particles,
          "hmm, 1000-step sequence (configs[4]), estimators only (no record leaves the GPU)")

    # configs[4] proper and the like-for-like form of configs[1]: every record written to the reference's posterior file
    def files(label, model, obs, n, note):
        d = scratch_dir()
        with tempfile.TemporaryDirectory(dir=d) as tmp:
            engine.infer_to_files(model, obs, n, os.path.join(tmp, "warm"))      # same size: pinned / text buffers reach their final size
            for f in os.listdir(tmp):
                os.remove(os.path.join(tmp, f))
            prefix = os.path.join(tmp, "post")
            t0 = time.perf_counter()
            st = engine.infer_to_files(model, obs, n, prefix)
            wall = time.perf_counter() - t0
            ts = engine.text_stage_stats()
            size = sum(os.path.getsize(prefix + ext) for ext in (".real", ".int") if os.path.exists(prefix + ext))
        row_bytes = 8 * st["n_real"] + 4 * st["n_int"] + 16
        k_s = st["device_ms"] * 1e-3
        out[label] = {
            "workload": note, "particles": n, "value": n / wall, "unit": "records/s (wall, files complete)", "wall_s": wall,
            "text_bytes": size, "text_GBps_wall": size / wall / 1e9, "particle_kernels_ms": st["device_ms"],
            "stages": {"text_kernels_ms": ts["kernel_ms"], "text_kernels_GBps": ts["bytes"] / max(ts["kernel_ms"], 1e-9) / 1e6,
                       "d2h_ms": ts["copy_ms"], "d2h_GBps": ts["bytes"] / max(ts["copy_ms"], 1e-9) / 1e6,
                       "file_write_s": ts["write_s"], "file_write_GBps": ts["bytes"] / max(ts["write_s"], 1e-9) / 1e9, "fixups": ts["fixups"]},
            "roofline": {"bound": "hbm", "achieved": row_bytes * n / k_s / 1e9, "peak": hbm, "unit": "GB/s", "frac": row_bytes * n / k_s / 1e9 / hbm,
                         "peak_source": hbm_src, "algorithmic_bytes_per_particle": row_bytes,
                         "note": "device stage only (rows written once by k_sis_rows + estimator kernels); the end-to-end rate is "
                                 "bounded by the file append on " + d, "traffic": costs.get(label, {}).get("dram_bytes_per_particle")},
            "files_on": d}

    files("C5", "hmm", g["obs_hmm_1000"], args.c5_file_particles,
          "hmm, 1000-step sequence, FULL trace emission to <prefix>.int in the reference's text format (configs[4])")
    files("C2_files", MODEL, OBS, args.file_particles,
          "gaussian_unknown_mean x=(3,4) through cpprob_sis_infer_to_files: every record appended to <prefix>.real, .ids written - "
          "the same work as the reference's cpprob::inference(..., outfile)")
    return out


def cpu_baseline(args):
    """The reference's CPU SIS (single thread, as the reference is) on a bounded sample of the same workload."""
    n = args.cpu_particles
    s = cpu_sis_step(1, n, "faithful")
    s_fast = cpu_sis_step(1, n, "fast")
    return {"value": n / s, "unit": "particles/s", "cores": 1, "kind": cpu_kind(),
            "sample": f"{n} particles of gaussian_unknown_mean x=(3,4): {cpu_what()}, files on {scratch_dir()}",
            "fast_flavour_value": n / s_fast,
            "fast_flavour": "the restated loop with one buffered ofstream per file kept open (oracle/, a port): what the reference would do "
                            "without its three open/append/close per trace"}


_RESULT_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one) is sent to stderr instead
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000_000, help="particles per GPU per step")
    ap.add_argument("--cpu-particles", type=int, default=1_000_000, help="particles of the cpu_baseline sample")
    ap.add_argument("--ref-particles", type=int, default=50_000, help="particles per process per step of --impl reference")
    ap.add_argument("--strong-particles", type=int, default=1_000_000_000, help="total particles of the strong-scaling extra")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3/C4/C5/file-emission block")
    ap.add_argument("--c3-particles", type=int, default=100_000_000)
    ap.add_argument("--c4-particles", type=int, default=100_000_000)
    ap.add_argument("--c5-particles", type=int, default=16_000_000)
    ap.add_argument("--c5-file-particles", type=int, default=1_000_000)
    ap.add_argument("--file-particles", type=int, default=20_000_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
