"""The sharded optimiser over a real NCCL process group (one process per GPU), against the same job on one GPU.
Needs two GPUs; on a one-GPU box the test skips (the emulated-rank tests in test_gpu_parity.py cover the logic there)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

C, S, NCAND, STEPS = 3, 4, 64, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _job(rank, world, port, nimg, mode, out):
    import torch
    import torch.distributed as dist
    from snesimage_b200 import driver, engine, synth
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ctx = engine.Context(rank)
    cfg = engine.Config(subpalette_count=C, subpalette_size=S)
    plan = driver.plan_shards(nimg, rank, world, mode)
    ims = [engine.OptimizedImage(ctx, synth.image(400 + j, "V"), cfg) for j in range(plan.img_lo, plan.img_hi)]
    engine.batch_initialize_tiles(ims)
    engine.batch_recalculate_palettes(ims)
    opt = driver.BatchOptimizer(ctx, ims, plan=plan, seed=11)
    if world > 1 and mode == "hybrid":       # the host-buffer steps through the library's own communicator (one C call)
        driver.init_library_comm(ctx, rank, world)
        opt.library_comm = True
    records = []
    for it in range(STEPS):
        cand = opt.candidates_host(NCAND)
        if it % 2 == 0:
            opt.step_random_dev(torch.from_numpy(cand).to(dev), NCAND)
            torch.cuda.synchronize()
            records.append(opt.best_records().copy())
        else:
            records.append(np.array(opt.step_random_host(cand)))      # the host-buffer halves around the all-gather
    sums = opt.state_checksums()
    whole = opt.gathered_records().copy() if world > 1 else None
    out.put((rank, plan.img_lo, plan.img_hi, sums, [r.tobytes() for r in records], None if whole is None else whole.tobytes()))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def _run(world, nimg, mode):
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    out = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_job, args=(r, world, port, nimg, mode, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return sorted(res)


@pytest.mark.parametrize("nimg,mode", [(4, "hybrid"), (1, "hybrid"), (3, "candidates")])
def test_two_gpu_job_equals_one_gpu_job(nimg, mode):
    """Image groups (4 images on 2 ranks), candidate slices of one image, and the replicated layout: after three steps every
    rank's images carry the checksums of the single-GPU job, the per-step winners are the same records, and ranks that
    share images hold identical replicas."""
    import torch
    from snesimage_b200 import engine
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ref = _run(1, nimg, "hybrid")[0]
    want_sums = ref[3]
    want_rec = [np.frombuffer(b, engine.BEST_DTYPE) for b in ref[4]]
    got = _run(2, nimg, mode)
    for rank, lo, hi, sums, recs, whole in got:
        assert sums == want_sums[lo:hi], (rank, mode)
        for it, b in enumerate(recs):
            r = np.frombuffer(b, engine.BEST_DTYPE)
            assert np.array_equal(r["idx"], want_rec[it]["idx"][lo:hi]) and np.array_equal(r["err"], want_rec[it]["err"][lo:hi]), (rank, it)
        assert whole is not None
