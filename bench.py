#!/usr/bin/env python
"""bench.py -- candidate palette evaluations per second (BASELINE.json metric).

Workload (BASELINE.json configs[4]): 64 synthetic 256x256 images, 8 subpalettes x 15
colours, RGB distance, no dither.  One step = one `optimize_palette_entry_random` for every image:
error() + 64 candidate evaluations (optimize() + error() each) per image per GPU + argmin + accept +
optimize().  Weak scaling: every rank evaluates 64 candidates per image (64*N per image in total); the only
exchange is the all-gather argmin of 16 bytes per image.

  value     whole-job candidate evaluations / s, candidate lists resident in HBM (device-timed, max over ranks)
  e2e       the same step through the host-buffer entry point of the C ABI (snes_batch_step_random): candidate H2D +
            winning records D2H inside every call (N > 1: pinned host buffers around the sharded device-pointer step)
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (C restatement of the reference) on this box's host cores, bounded sample

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
WORKLOAD = "cfg5: 64 synthetic 256x256 images (V family, seeds 0..63) x 64 random candidates/image/GPU per step, 8x15, RGB, no dither"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "subpalettes": C, "colours": S, "metric_mode": "rgb", "dither": False},
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
            self._stop.wait(0.01)

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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from snesimage_b200 import driver, engine, synth

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

    ctx = engine.Context(local, chunk=args.chunk)
    cfg = engine.Config(subpalette_count=C, subpalette_size=S)
    images = [engine.OptimizedImage(ctx, synth.image(s, "V"), cfg) for s in range(args.nimg)]
    engine.batch_initialize_tiles(images)
    engine.batch_recalculate_palettes(images)
    opt = driver.BatchOptimizer(ctx, images, rank=rank, world=world, group=None, seed=0)

    ncand_total = args.ncand * world
    nsteps = args.warmup + args.steps
    # pre-generate every step's candidate list: (steps, nimg, ncand_total, 3)
    cand_host = []
    for it in range(nsteps):
        opt.iteration = it
        cand_host.append(opt.candidates_host(ncand_total))
    opt.iteration = 0
    d_cands = [torch.from_numpy(c).to(dev) for c in cand_host]
    torch.cuda.synchronize()

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

    # ---- device-resident arm (value) ----
    for it in range(args.warmup):
        opt.step_random_dev(d_cands[it], ncand_total)
    clocks = ClockSampler(local)
    launches0 = ctx.kernel_launches
    ctx.profile_begin()
    clocks.start()
    ms = timed(lambda it: opt.step_random_dev(d_cands[it], ncand_total), range(args.warmup, nsteps))
    clock_info = clocks.stop()
    prof = ctx.profile_end()
    launches = ctx.kernel_launches - launches0
    evals_per_step = args.nimg * ncand_total
    value = evals_per_step * args.steps / (ms * 1e-3)

    # ---- host-buffer arm (e2e): same steps again from pinned host memory ----
    pinned = [torch.from_numpy(c).pin_memory() for c in cand_host]
    d_stage = torch.empty_like(d_cands[0])
    best_pinned = torch.zeros(args.nimg * 2, dtype=torch.int64).pin_memory()
    if world == 1:
        # the host-buffer entry point of the C ABI itself: snes_batch_step_random(ctx, images, ..., cand /* host */, ...,
        # best /* host */): candidates are copied to the device, the step runs, the winning records come back, the call
        # returns when the stream has drained -- all inside the timed region
        pinned_np = [p.numpy() for p in pinned]
        best_np = best_pinned.numpy().view(engine.BEST_DTYPE)

        def host_step(it):
            p, i = opt.cursor.palette, opt.cursor.palette_index
            best, _ = engine.batch_step_random(images, p, i, pinned_np[it])
            best_np[:] = best
            opt.cursor.advance(opt.config)
            opt.iteration += 1
    else:
        def host_step(it):
            opt.step_random_host(pinned[it], d_stage, best_pinned)
    for it in range(min(3, args.warmup)):
        host_step(it)
    ms_e2e = timed(host_step, range(args.warmup, nsteps))
    e2e_value = evals_per_step * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel ----
    # Algorithmic bytes per evaluation: S1 = assignment (source RGBA8 + tile/palette tables +
    # palette_map write), S2 = scoring (palette_map + alpha mask + 36 B of source planes per scale-pixel + partial sums).
    peak, peak_src = measured_peaks()
    total_ms = sum(v["ms"] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, {"ms": 0.0, "n": 0})
    roofline = None
    if top[0]:
        name, st = top
        # every launch of the scorer covers one chunk of evaluations (bookkeeping error() launches are small ones);
        # evaluations it processed in the timed region = candidates + the per-step error() of each image
        scored = (args.nimg * args.ncand + args.nimg) * args.steps
        alg_bytes = B_S2 if name.startswith("k_score") else B_ALG
        per_launch_bytes = alg_bytes * scored / st["n"]
        achieved = alg_bytes * scored / (st["ms"] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("kernel", "").split("<")[0] == name.split("<")[0]:
                traffic = tj["dram_bytes_per_eval"] * scored / st["n"]
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": name, "kernel_launches": st["n"], "kernel_ms_per_launch": st["ms"] / st["n"],
                    "kernel_share_of_gpu_time": st["ms"] / max(1e-9, total_ms),
                    "alg_bytes_per_eval": alg_bytes, "alg_bytes_per_launch": per_launch_bytes, "peak_source": peak_src,
                    "whole_path": {"alg_bytes_per_eval": B_ALG, "achieved": B_ALG * scored / (total_ms * 1e-3) / 1e9,
                                   "frac": B_ALG * scored / (total_ms * 1e-3) / 1e9 / peak, "gpu_ms_per_step": total_ms / args.steps},
                    "unique_bytes_per_eval": B_MIN,
                    "note": "the scorer is FP32/FP64 issue-bound (recursive-Gaussian chains), not HBM-bound: see DESIGN.md for the instruction roofline",
                    "kernels": prof}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "subpalettes": C, "colours": S, "metric_mode": "rgb", "dither": False,
                       "images": args.nimg, "candidates_per_image_per_gpu": args.ncand, "evals_per_step": evals_per_step,
                       "parallelism": f"candidate-sharded x{world}, all-gather argmin (16 B/image)",
                       "l2": "inputs larger than L2: 64 images x 3.4 MB of source planes + per-evaluation intermediates (GBs per step)",
                       "bookkeeping_evals_per_step_not_counted": 2 * args.nimg, "chunk": args.chunk},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(cand_host[0].nbytes), "d2h_bytes_per_step": int(best_pinned.numel() * 8)},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
