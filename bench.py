#!/usr/bin/env python
"""bench.py -- candidate palette evaluations per second (BASELINE.json metric).

Workload (BASELINE.json configs[4]): 64 synthetic 256x256 images, 8 subpalettes x 15
colours, RGB distance, no dither, 64 random candidates per image and step = 4096 candidate evaluations per step IN
TOTAL, whatever the number of GPUs (strong scaling).  One step = one `optimize_palette_entry_random` for every image:
error() + 64 x (optimize() + error()) + argmin + accept + optimize() (lib.rs:191-240).  At N > 1 the job is sharded as
driver.plan_shards says: with 64 images every rank owns 64/N images outright (nothing replicated); the exchange per
step is the all-gather of the 16-byte (error, index) records over NCCL.

  value     whole-job candidate evaluations / s, candidate lists resident in HBM (device-timed, max over ranks)
  e2e       the same step through the host-buffer entry points of the C ABI: candidate H2D + winning records D2H inside
            every call (N = 1: snes_batch_step_random; N > 1: snes_dist_step_random, the all-gather inside the library)
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (C restatement of the reference) on this box's host cores, bounded sample
  outside the headline's timed region: `weak` and `candidate_sharded` (N > 1), `modes` and `configs` (N = 1)

`--impl reference` times the CPU restatement of the reference's own loop on all host cores (the Rust
crate cannot be built in this image: no cargo/rustc, un-vendored crates -- see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate palette evals/sec (256x256, SSIMULACRA2)"
UNIT = "evals/s"
C, S = 8, 15
NIMG, NCAND = 64, 64
# algorithmic bytes per candidate evaluation (8x15 palettes)
B_ALG = 3_548_624
B_S2 = 3_219_560      # scoring share of B_ALG: 65,536 + 8,192 + 3,144,960 + 872
B_MIN = 403_664
WORKLOAD = ("cfg5: 64 synthetic 256x256 images (V family, seeds 0..63) x 64 random candidates per image = 4096 candidate "
            "evaluations per step in total, 8x15, RGB, no dither")


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def build_config(args, world: int) -> dict:
    """The `config` object of the JSON line -- one function for both arms, so they name the same workload key by key."""
    groups = max(d for d in range(1, world + 1) if world % d == 0 and d <= args.nimg)
    return {"workload": WORKLOAD, "subpalettes": C, "colours": S, "metric_mode": "rgb", "dither": False,
            "images": args.nimg, "candidates_per_image": args.ncand, "evals_per_step": args.nimg * args.ncand,
            "parallelism": f"{groups} image groups x {world // groups} candidate slices over {world} GPU(s), all-gather argmin (16 B/image)",
            "l2": "inputs larger than L2: per step and GPU 3.4 MB of source planes per image held + 326 KB of intermediates per "
                  "evaluation (>= 190 MB at 8 GPUs, 1.5 GB at one)",
            "bookkeeping_evals_per_step_not_counted": 2 * args.nimg, "chunk": args.chunk}


# ---- CPU side: the oracle as the reference's stand-in --------------------------------------------------
def _oracle_state(seed: int):
    """An oracle image in the state the timed workload starts from (after both k-means inits)."""
    from oracle import binding as ob
    from snesimage_b200 import synth
    o = ob.OracleImage(synth.image(seed, "V"), C, S)
    o.initialize_tiles()
    o.recalculate_palettes()
    return o


_WORKER_STATE = None


def _worker_init():
    global _WORKER_STATE
    from oracle import binding as ob
    ob.lib()
    _WORKER_STATE = _oracle_state((os.getpid() * 7) % NIMG)


def _worker_eval(n: int) -> int:
    """n candidate evaluations (optimize() + error() each, lib.rs:205-220) on this worker's image."""
    from snesimage_b200 import synth
    _WORKER_STATE.eval_candidates(0, 0, synth.candidates(os.getpid(), 0, n))
    return n


class CpuPool:
    """One process per host core (the reference itself is single-threaded; this is the most the box's cores can
    give it), each holding its own oracle image in the post-k-means state."""

    def __init__(self, procs: int):
        import multiprocessing as mp
        self.procs = procs
        self.pool = mp.get_context("fork").Pool(procs, initializer=_worker_init)
        self.pool.map(_worker_eval, [1] * procs)   # warm: page in, build states

    def run(self, evals_per_proc: int):
        t0 = time.perf_counter()
        done = sum(self.pool.map(_worker_eval, [evals_per_proc] * self.procs, chunksize=1))
        dt = time.perf_counter() - t0
        return done, dt

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_evals_per_sec(procs: int, evals_per_proc: int):
    pool = CpuPool(procs)
    try:
        done, dt = pool.run(evals_per_proc)
    finally:
        pool.close()
    return done / dt, dt


def run_reference(args):
    """The reference arm: CPU restatement on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_proc = 4
    t_all = time.perf_counter()
    pool = CpuPool(cores)
    try:
        for _ in range(args.warmup):
            pool.run(1)
        total, evals = 0.0, 0
        for _ in range(args.steps):
            done, dt = pool.run(per_proc)
            total += dt
            evals += done
    finally:
        pool.close()
    value = evals / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": build_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} processes x {per_proc} candidate evaluations per step (optimize()+error() each), "
                                   f"C restatement of lib.rs + crates (oracle/), gcc -O2; the Rust reference cannot be built here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    _emit(line)


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""
    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._th:
            self._th.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, sustained-in-step use)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- GPU arm -----------------------------------------------------------------------------------------
class Job:
    """One sharded optimiser job on this rank: its images (created and k-means-initialised), its BatchOptimizer and the
    candidate lists of every step, on the device and in pinned host memory."""

    def __init__(self, ctx, dev, rank, world, nimg, ncand_total, nsteps, mode="hybrid", cfg_kw=None, family="V"):
        import torch
        from snesimage_b200 import driver, engine, synth
        self.torch, self.dev, self.world = torch, dev, world
        self.plan = driver.plan_shards(nimg, rank, world, mode)
        cfg = engine.Config(**{"subpalette_count": C, "subpalette_size": S, **(cfg_kw or {})})
        self.images = [engine.OptimizedImage(ctx, synth.image(s, family), cfg) for s in range(self.plan.img_lo, self.plan.img_hi)]
        engine.batch_initialize_tiles(self.images)
        engine.batch_recalculate_palettes(self.images)
        self.opt = driver.BatchOptimizer(ctx, self.images, plan=self.plan, seed=0)
        self.ncand_total = ncand_total
        self.cand_host = []
        for it in range(nsteps):
            self.opt.iteration = it
            self.cand_host.append(self.opt.candidates_host(ncand_total))
        self.opt.iteration = 0
        self.d_cands = [torch.from_numpy(c).to(dev) for c in self.cand_host]
        self.pinned = [torch.from_numpy(c).pin_memory().numpy() for c in self.cand_host]
        self.evals_per_step = nimg * ncand_total
        torch.cuda.synchronize()

    def step_dev(self, it):
        self.opt.step_random_dev(self.d_cands[it], self.ncand_total)

    def step_host(self, it):
        self.opt.step_random_host(self.pinned[it])

    def checksum(self) -> int:
        h = 0xcbf29ce484222325
        for v in self.opt.state_checksums():
            h = ((h ^ v) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
        return h

    def close(self):
        for im in self.images:
            im.close()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from snesimage_b200 import _build, engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # CPU baseline first (forks a worker; done before this process touches CUDA)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, dt = cpu_evals_per_sec(1, args.cpu_evals)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"{args.cpu_evals} candidate evaluations of one image (optimize()+error() each) in one host process "
                                  f"(the reference's loop is single-threaded), {dt:.1f} s; box has {host_cores()} host cores"}
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps_range):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in steps_range:
            fn(it)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def gather_checksums(job):
        mine = (rank, job.plan.group, f"{job.checksum():016x}")
        if world == 1:
            return [mine]
        out = [None] * world
        dist.all_gather_object(out, mine)
        return out

    ctx = engine.Context(local, chunk=args.chunk)
    nsteps = args.warmup + args.steps

    # ---- headline: cfg5 as written, 4096 evaluations per step in total, sharded (strong scaling) ----
    job = Job(ctx, dev, rank, world, args.nimg, args.ncand, nsteps)
    for it in range(args.warmup):
        job.step_dev(it)
    clocks = ClockSampler(local)
    launches0 = ctx.kernel_launches
    ctx.profile_begin(only="k_score")          # event pairs around the dominant kernel only: the timed region stays undisturbed
    clocks.start()
    ms = timed(job.step_dev, range(args.warmup, nsteps))
    clock_info = clocks.stop()
    prof_top = ctx.profile_end()
    launches = ctx.kernel_launches - launches0
    evals_per_step = job.evals_per_step
    value = evals_per_step * args.steps / (ms * 1e-3)

    # ---- host-buffer arm (e2e): the same steps again from pinned host memory through the C ABI's host entry points ----
    if world > 1:
        from snesimage_b200 import driver
        driver.init_library_comm(ctx, rank, world)     # the library's own NCCL communicator: the sharded step is ONE C call
        job.opt.library_comm = True
    for it in range(min(3, args.warmup)):
        job.step_host(it)
    ms_e2e = timed(job.step_host, range(args.warmup, nsteps))
    e2e_value = evals_per_step * args.steps / (ms_e2e * 1e-3)
    # every kernel's CUDA-event time per step, from an extra (untimed) pass over the same steps
    ctx.profile_begin()
    for it in range(args.warmup, nsteps):
        job.step_dev(it)
    prof = ctx.profile_end()
    for k, v in prof_top.items():
        prof[k] = v                            # the dominant kernel's figures are the timed region's own
    sums = gather_checksums(job)
    h2d = int(job.cand_host[0].nbytes)
    d2h = int(job.plan.nloc * 16)
    nloc = job.plan.nloc
    plan_text = job.plan.describe()

    # ---- roofline of the dominant kernel (this rank's launches) ----
    # Algorithmic bytes per evaluation: S1 = assignment (source RGBA8 + tile/palette tables +
    # palette_map write), S2 = scoring (palette_map + alpha mask + 36 B of source planes per scale-pixel + partial sums).
    peak, peak_src = measured_peaks()
    total_ms = sum(v["ms"] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, {"ms": 0.0, "n": 0})
    roofline = None
    if top[0]:
        name, st = top
        # every launch of the scorer covers one chunk of evaluations; evaluations this rank's scorer processed in the
        # timed region = its candidates + the per-step error() of each of its images
        from snesimage_b200 import driver
        lo, hi = driver.shard_bounds(args.ncand, job.plan.slice, job.plan.cand_ranks)
        scored = (nloc * (hi - lo) + nloc) * args.steps
        alg_bytes = B_S2 if name.startswith("k_score") else B_ALG
        per_launch_bytes = alg_bytes * scored / st["n"]
        achieved = alg_bytes * scored / (st["ms"] * 1e-3) / 1e9
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("kernel", "").split("<")[0] != name.split("<")[0]:
                traffic_note = "profiles/roofline_traffic.json describes another kernel"
            elif tj.get("scorer_source_sha256") != _build.scorer_source_hash():
                traffic_note = "profiles/roofline_traffic.json was captured from other scorer sources than the tree's: not reported"
            else:
                traffic = tj["dram_bytes_per_eval"] * scored / st["n"]
                traffic_note = "ncu --set full capture of the same sources (profiles/roofline_traffic.json), scaled to this launch size"
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": traffic_note,
                    "kernel": name, "kernel_launches": st["n"], "kernel_ms_per_launch": st["ms"] / st["n"],
                    "kernel_share_of_gpu_time": st["ms"] / max(1e-9, total_ms),
                    "alg_bytes_per_eval": alg_bytes, "alg_bytes_per_launch": per_launch_bytes, "peak_source": peak_src,
                    "whole_path": {"alg_bytes_per_eval": B_ALG, "achieved": B_ALG * scored / (total_ms * 1e-3) / 1e9,
                                   "frac": B_ALG * scored / (total_ms * 1e-3) / 1e9 / peak, "gpu_ms_per_step": total_ms / args.steps},
                    "unique_bytes_per_eval": B_MIN,
                    "note": "per GPU (rank 0's launches); the scorer is FP32/FP64 issue-bound (recursive-Gaussian chains), not HBM-bound: see DESIGN.md",
                    "kernels": prof}
    job.close()

    # ---- outside the headline's timed region -------------------------------------------------------------------------
    extras = {}
    if not args.no_extras and world > 1:
        # the same 4096 evaluations per step with the round-1 layout (every rank holds all 64 images and evaluates a slice
        # of each image's candidates): what replicating the per-image work costs
        j2 = Job(ctx, dev, rank, world, args.nimg, args.ncand, nsteps, mode="candidates")
        for it in range(args.warmup):
            j2.step_dev(it)
        ms2 = timed(j2.step_dev, range(args.warmup, nsteps))
        sums2 = gather_checksums(j2)
        extras["candidate_sharded"] = {"value": j2.evals_per_step * args.steps / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / args.steps,
                                       "scaling": "strong", "parallelism": j2.plan.describe(),
                                       "replicas_identical": len({c for _, _, c in sums2}) == 1}
        j2.close()
        # weak scaling as in round 1: 64 candidates per image AND GPU (64 x N per image), every rank holds all images
        j3 = Job(ctx, dev, rank, world, args.nimg, args.ncand * world, nsteps, mode="candidates")
        for it in range(args.warmup):
            j3.step_dev(it)
        ms3 = timed(j3.step_dev, range(args.warmup, nsteps))
        extras["weak"] = {"value": j3.evals_per_step * args.steps / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3 / args.steps,
                          "scaling": "weak", "evals_per_step": j3.evals_per_step, "parallelism": j3.plan.describe()}
        j3.close()
    if not args.no_extras and world == 1:
        extras["modes"] = measure_modes(ctx, dev, args)
        extras["configs"] = measure_configs(ctx)

    if rank == 0:
        groups = {}
        for r, g, c in sums:
            groups.setdefault(g, set()).add(c)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": build_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "entry_points": "snes_batch_step_random" if world == 1 else "snes_dist_step_random (one call per rank and step: H2D, error + candidates, ncclAllGather inside the library, merge, accept, optimize, D2H)"},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "sharding": {"plan": plan_text, "images_per_rank": nloc,
                         "state_checksums": [{"rank": r, "image_group": g, "fnv1a64": c} for r, g, c in sums],
                         "replicas_identical": all(len(v) == 1 for v in groups.values())},
        }
        line.update(extras)
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_modes(ctx, dev, args) -> dict:
    """Device-timed step of the four reference modes on 64 images (one GPU): evaluations per second and the CUDA-event
    time of every kernel of a step.  rgb / lab / dither: optimize_palette_entry_random with 64 candidates; nes: the 56
    NES colours (lib.rs:242-284)."""
    import torch
    from snesimage_b200 import engine
    modes = {"rgb": {}, "lab": {"perceptual_palettes": True, "subpalette_count": 4, "subpalette_size": 7},
             "dither": {"dither": True}, "nes": {"nes": True, "dither": True, "subpalette_count": 4, "subpalette_size": 3}}
    out = {}
    warm, reps = 2, 4
    for name, kw in modes.items():
        job = Job(ctx, dev, 0, 1, args.nimg, args.ncand, warm + reps, cfg_kw=kw)
        nes = bool(kw.get("nes"))
        per_step = args.nimg * (56 if nes else args.ncand)

        def step(it):
            if nes:
                engine.batch_step_nes(job.images, it % kw["subpalette_count"], it % kw["subpalette_size"])
            else:
                job.step_dev(it)
        for it in range(warm):
            step(it)
        torch.cuda.synchronize()
        ctx.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(warm, warm + reps):
            step(it)
        e1.record()
        torch.cuda.synchronize()
        prof = ctx.profile_end()
        ms = e0.elapsed_time(e1) / reps
        cfg = job.images[0].config
        out[name] = {"value": per_step / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "evals_per_step": per_step,
                     "subpalettes": cfg.subpalette_count, "colours": cfg.subpalette_size,
                     "kernel_ms_per_step": {k: round(v["ms"] / reps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}
        job.close()
    return out


def measure_configs(ctx, iters: int = 100) -> dict:
    """BASELINE.json configs[0..3] the way the reference is used: ONE picture, k-means init + tile assignment, then 100
    iterations of the schedule of run() (lib.rs:889-933) through the headless driver; wall clock around synchronous calls."""
    from snesimage_b200 import driver, engine, synth
    cfgs = {"cfg1": dict(subpalette_count=8, subpalette_size=15),
            "cfg2": dict(subpalette_count=4, subpalette_size=7, perceptual_palettes=True),
            "cfg3": dict(subpalette_count=8, subpalette_size=15, dither=True),
            "cfg4": dict(subpalette_count=4, subpalette_size=3, nes=True, dither=True)}
    out = {}
    rgba = synth.image(0, "V")
    for name, kw in cfgs.items():
        cfg = engine.Config(**kw)
        r = driver.HeadlessRunner(ctx, rgba, cfg, seed=0, ncand=64)
        t0 = time.perf_counter()
        r.initialize()
        ctx.synchronize()
        t1 = time.perf_counter()
        e0 = r.image.error()
        r.iterate(3)
        ctx.synchronize()
        t2 = time.perf_counter()
        moves0 = len(r.log)
        r.iterate(iters)
        ctx.synchronize()
        t3 = time.perf_counter()
        moves1, e1 = len(r.log), r.image.error()
        later = 4 * iters
        r.iterate(later)                     # further down the same trajectory fewer iterations find a better colour
        ctx.synchronize()
        t4 = time.perf_counter()
        per_iter = 56 if cfg.nes else 64
        out[name] = {"init_ms": 1e3 * (t1 - t0), "ms_per_iteration": 1e3 * (t3 - t2) / iters, "iterations": iters,
                     "candidate_evals_per_s": iters * per_iter / (t3 - t2), "candidates_per_iteration": per_iter,
                     "iterations_that_moved": moves1 - moves0, "error_start": e0, "error_end": e1,
                     "next_iterations": later, "next_ms_per_iteration": 1e3 * (t4 - t3) / later,
                     "next_candidate_evals_per_s": later * per_iter / (t4 - t3), "next_iterations_that_moved": len(r.log) - moves1,
                     "error_after_all": r.image.error(), "lookahead_iterations_per_call": "adaptive (driver.HeadlessRunner.lookahead)",
                     "config": kw}
        r.image.close()
    return out


_RESULT_FD = None


def _claim_stdout():
    """Keep stdout for the one JSON line: everything else written to fd 1 from here on -- native libraries included (NCCL
    prints its version banner there when NCCL_DEBUG is set in the environment) -- goes to stderr."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def relaunch_command(args_gpus: int, argv) -> list:
    """`python bench.py --gpus N` outside torchrun: the command that runs the same arguments under torch.distributed.run."""
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args_gpus}", "--master-addr", "127.0.0.1",
            "--master-port", os.environ.get("BENCH_MASTER_PORT", "29517"), os.path.abspath(__file__)] + list(argv)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nimg", type=int, default=NIMG)
    ap.add_argument("--ncand", type=int, default=NCAND)
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--cpu-evals", type=int, default=128)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the measurements outside the headline (modes, configs, weak, candidate_sharded)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun.  Done BEFORE stdout is claimed, so the children inherit the real stdout
        # and rank 0's JSON line lands there.
        import subprocess
        launcher = os.environ.get("BENCH_LAUNCHER")   # tests substitute a stub for torch.distributed.run
        cmd = relaunch_command(args.gpus, sys.argv[1:])
        if launcher:
            cmd = [sys.executable, launcher] + cmd[3:]
        raise SystemExit(subprocess.call(cmd))
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
