"""Host-side logic without a GPU: the optimiser cursor (lib.rs:881-933), candidate sharding, and the
cross-rank argmin exercised over a real world_size-2 gloo process group."""
import os
import socket

import numpy as np
import pytest

from snesimage_b200 import driver, engine, synth


def _reference_schedule(C, S, nes, iterations):
    """Literal restatement of the loop body of run() (lib.rs:889-932): yields (method, palette, index, channel)."""
    step = palette = palette_index = channel = 0
    for _ in range(iterations):
        random = step % 5 < 4
        if nes:
            yield ("nes", palette, palette_index, channel)
        elif random:
            yield ("random", palette, palette_index, channel)
        else:
            yield ("channel", palette, palette_index, channel)
        channel += 1
        if channel == 3 or random:
            channel = 0
            palette_index += 1
            if palette_index == S:
                palette_index = 0
                palette += 1
                if palette == C:
                    palette = 0
                    step += 1


@pytest.mark.parametrize("C,S,nes", [(2, 3, False), (4, 7, False), (4, 3, True), (1, 2, False)])
def test_cursor_follows_reference_schedule(C, S, nes):
    cfg = engine.Config(subpalette_count=C, subpalette_size=S, nes=nes)
    cur = driver.Cursor()
    n = C * S * 4 + C * S * 3 + 5   # four random sweeps, one channel sweep, a bit more
    for want in _reference_schedule(C, S, nes, n):
        assert (cur.mode(cfg), cur.palette, cur.palette_index, cur.channel) == want
        cur.advance(cfg)
    assert cur.step >= 5


def test_cfg1_hundred_iterations_are_all_random():
    cfg = engine.Config(subpalette_count=8, subpalette_size=15)
    cur = driver.Cursor()
    for _ in range(100):
        assert cur.mode(cfg) == "random"
        cur.advance(cfg)
    assert (cur.palette, cur.palette_index, cur.step) == (6, 10, 0)


def test_lookahead_follows_the_accept_rate():
    """HeadlessRunner.lookahead: one iteration per call while nearly every iteration accepts a candidate, deeper look-ahead as
    accepts become rare, deeper still with dithering (a call's fixed cost is higher there), never beyond 4 x speculate, and 1 when
    look-ahead is switched off.  Pure host logic: no GPU."""
    class R(driver.HeadlessRunner):
        def __init__(self, rate, dither=False, speculate=4):
            self.speculate, self.accept_rate, self.config = speculate, rate, engine.Config(dither=dither)
    ks = [R(a).lookahead() for a in (0.95, 0.5, 0.3, 0.1, 0.03, 0.01)]
    assert ks[0] == 1 and ks == sorted(ks) and ks[-1] == 16
    assert all(R(a, dither=True).lookahead() >= R(a).lookahead() for a in (0.9, 0.5, 0.2, 0.05))
    assert R(0.01, speculate=1).lookahead() == 1 and R(0.01, speculate=2).lookahead() <= 8


def test_shard_bounds_partition():
    for n in (1, 7, 56, 64, 4096):
        for world in (1, 2, 3, 4, 8):
            spans = [driver.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_plan_shards_covers_every_image_and_candidate_once():
    """The (image group x candidate slice) grid: every (image, candidate) pair belongs to exactly one rank, images go to
    whole ranks while there are enough of them, and the all-gather slot count fits the largest group."""
    for nimg, world, ncand in [(64, 1, 64), (64, 2, 64), (64, 4, 64), (64, 8, 64), (1, 8, 64), (4, 8, 64), (5, 2, 7), (3, 8, 2), (6, 4, 1)]:
        plans = [driver.plan_shards(nimg, r, world) for r in range(world)]
        seen = np.zeros((nimg, ncand), np.int32)
        for pl in plans:
            assert pl.img_groups * pl.cand_ranks == world and pl.img_groups <= nimg
            assert pl.nloc <= pl.slots and pl.first_rank_of_group + pl.slice == pl.rank
            lo, hi = driver.shard_bounds(ncand, pl.slice, pl.cand_ranks)
            seen[pl.img_lo:pl.img_hi, lo:hi] += 1
        assert (seen == 1).all(), (nimg, world, ncand)
        if nimg >= world:
            assert plans[0].cand_ranks == 1          # enough images: nothing is replicated
        # the replicated layout of round 1 stays available
        pc = driver.plan_shards(nimg, world - 1, world, "candidates")
        assert (pc.img_groups, pc.cand_ranks, pc.img_lo, pc.img_hi, pc.slice) == (1, world, 0, nimg, world - 1)
    with pytest.raises(ValueError):
        driver.plan_shards(4, 2, 2)


def test_library_plan_equals_driver_plan():
    """snes_dist_plan (the C restatement a torch-free host uses) and driver.plan_shards name the same images, slices and slots."""
    for nimg in (1, 3, 4, 5, 7, 64):
        for world in (1, 2, 3, 4, 6, 8):
            for r in range(world):
                a, p = engine.dist_plan(nimg, r, world), driver.plan_shards(nimg, r, world)
                assert (a["img_lo"], a["img_hi"], a["cand_ranks"], a["slice"], a["slots"]) == (p.img_lo, p.img_hi, p.cand_ranks, p.slice, p.slots)
    with pytest.raises(engine.SnesGpuError):
        engine.dist_plan(4, 2, 2)


def test_bench_relaunch_keeps_the_json_line_on_stdout(tmp_path):
    """`python bench.py --gpus N` outside torchrun relaunches itself under torch.distributed.run; the child must inherit the
    real stdout (the parent used to point fd 1 at stderr first, so the line went to stderr).  A stub stands in
    for the launcher: it runs the same bench.py in the reference arm, whose line must arrive on the parent's stdout."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stub = tmp_path / "launcher_stub.py"
    stub.write_text(
        "import os, subprocess, sys\n"
        "args = sys.argv[1:]\n"
        "script = next(i for i, a in enumerate(args) if a.endswith('bench.py'))\n"
        "os.environ['WORLD_SIZE'] = '2'\n"           # what torchrun would set; the child then skips the relaunch
        "sys.exit(subprocess.call([sys.executable, args[script], '--impl', 'reference', '--steps', '1', '--warmup', '1']))\n")
    env = dict(os.environ, BENCH_LAUNCHER=str(stub))
    env.pop("WORLD_SIZE", None)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--gpus", "2"], capture_output=True, text=True, timeout=600,
                         cwd=root, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference", (out.stdout[:300], out.stderr[-300:])


def test_merge_best_is_first_minimum():
    rng = np.random.RandomState(0)
    nimg, ncand, world = 9, 24, 4
    scores = rng.randint(0, 6, (nimg, ncand)).astype(np.float64)   # many exact ties
    gathered = np.zeros((world, nimg), engine.BEST_DTYPE)
    for r in range(world):
        lo, hi = driver.shard_bounds(ncand, r, world)
        k = np.argmin(scores[:, lo:hi], axis=1)
        gathered[r]["idx"] = lo + k
        gathered[r]["err"] = scores[np.arange(nimg), lo + k]
    merged = driver.merge_best_host(gathered)
    assert np.array_equal(merged["idx"], np.argmin(scores, axis=1))   # np.argmin = first minimum = strict-< rule
    # rank order must not matter
    assert np.array_equal(driver.merge_best_host(gathered[::-1].copy())["idx"], merged["idx"])
    # a rank without candidates reports idx -1 and never wins
    gathered[2]["idx"] = -1
    gathered[2]["err"] = -1.0
    m2 = driver.merge_best_host(gathered)
    assert (m2["idx"] >= 0).all() and (m2["err"] >= 0).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, nimg, ncand, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank derives the same full score table, evaluates only its share of the (image group x candidate slice) grid,
    # then all-gathers 16-byte records, padded to the plan's slot count
    scores = (synth.hashn(3, np.arange(nimg)[:, None], np.arange(ncand)[None, :]) % np.uint64(5)).astype(np.float64)
    pl = driver.plan_shards(nimg, rank, world)
    lo, hi = driver.shard_bounds(ncand, pl.slice, pl.cand_ranks)
    local = np.zeros(pl.slots, engine.BEST_DTYPE)
    local["idx"] = -1
    if hi > lo:
        mine = scores[pl.img_lo:pl.img_hi, lo:hi]
        k = np.argmin(mine, axis=1)
        local["idx"][:pl.nloc] = lo + k
        local["err"][:pl.nloc] = mine[np.arange(pl.nloc), k]
    t_local = torch.from_numpy(local.view(np.int64).copy())
    t_all = torch.zeros(world * pl.slots * 2, dtype=torch.int64)
    dist.all_gather_into_tensor(t_all, t_local)
    gathered = t_all.numpy().view(engine.BEST_DTYPE).reshape(world, pl.slots)
    # the records of this rank's group decide its images; every rank can also rebuild the whole job's winners
    first = pl.first_rank_of_group
    merged = driver.merge_best_host(gathered[first:first + pl.cand_ranks])[:pl.nloc]
    ok = np.array_equal(merged["idx"], np.argmin(scores[pl.img_lo:pl.img_hi], axis=1))
    whole = []
    for g in range(pl.img_groups):
        glo, ghi = driver.shard_bounds(nimg, g, pl.img_groups)
        whole.append(driver.merge_best_host(gathered[g * pl.cand_ranks:(g + 1) * pl.cand_ranks])[:ghi - glo])
    whole = np.concatenate(whole)
    ok = ok and np.array_equal(whole["idx"], np.argmin(scores, axis=1))
    res = torch.tensor([int(ok), int(whole["idx"].sum())])
    got = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(got, res)
    if rank == 0:
        out.put([g.tolist() for g in got])
    dist.destroy_process_group()


@pytest.mark.parametrize("nimg,ncand", [(16, 64), (1, 64), (3, 1)])
def test_argmin_over_gloo_world_size_2(nimg, ncand):
    """world_size 2 over gloo: 16 images -> two image groups (no slicing); one image -> two candidate slices of it;
    three images -> groups of 1 and 2 images (padded records)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, nimg, ncand, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[0] == 1 for r in res)          # every rank reproduces the single-process argmin
    assert res[0][1] == res[1][1]               # and they agree with each other


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement on the host cores) must print one JSON line with the keys the
    driver reads, without touching CUDA."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    assert len(out.stdout.strip().splitlines()) == 1, out.stdout[:400]   # stdout carries the JSON line and nothing else
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["metric"].startswith("candidate palette evals/sec") and "workload" in line["config"]
    # both arms build `config` with one function, so the driver's same_config check compares like with like
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ns = argparse.Namespace(nimg=bench.NIMG, ncand=bench.NCAND, chunk=4096)
    assert line["config"] == bench.build_config(ns, 1)


def test_vertical_chain_thread_mapping_covers_every_plane_column_once():
    """Index logic of k_score_v3's vertical pass (score_v3.cuh): role threads 0..63 take word t of the interleaved
    (mu2, s22) row, i.e. (column t / 2, plane t % 2); threads 64..95 take column t - 64 of s12; lanes beyond a narrow
    scale's width sit out.  Every (plane, column) of the 32-column block must be owned by exactly one thread, and a
    warp's lanes must touch consecutive words."""
    BW = 32
    for D in (8, 16, 32, 64, 128, 256):
        owned = []
        for t in range(128):
            if t < 2 * BW:
                if (t >> 1) < D:
                    owned.append((t & 1, t >> 1, ("h01", t)))
            elif t < 3 * BW and t - 2 * BW < D:
                owned.append((2, t - 2 * BW, ("h2", t - 2 * BW)))
        want = {(pl, col) for pl in range(3) for col in range(min(D, BW))}
        assert sorted((pl, col) for pl, col, _ in owned) == sorted(want)
        for warp in range(3):
            words = [w for _, _, (buf, w) in owned if (buf == "h01" and w // 32 == warp) or (buf == "h2" and warp == 2)]
            assert not words or words == list(range(words[0], words[0] + len(words)))   # consecutive words: one wavefront per access
