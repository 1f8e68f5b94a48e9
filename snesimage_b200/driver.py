"""Headless optimiser driver: the schedule of `run()` (/root/reference/src/lib.rs:881-933) without the
SDL window, for one image or a batch of independent images, on one GPU or sharded over the GPUs of
one box (one process per GPU, `torch.distributed`).

Sharding: a step's evaluations are independent across images AND across the candidates
of one image, so the ranks form a grid of image groups x candidate slices (`plan_shards`):

  * while there are at least as many images as ranks, every rank owns a contiguous block of images outright
    -- it alone holds them, evaluates all their candidates and does their per-image bookkeeping (`error()`, the
    delta-assignment preparation, the final `optimize()`), so nothing is replicated;
  * ranks left over (fewer images than ranks, e.g. one picture on eight GPUs) split the candidate list of their
    group's images: rank slice s evaluates candidates [s*n/R, (s+1)*n/R) in place from the full list.

The one exchange per step is the argmin: each rank's per-image (error, global candidate index) record -- 16 bytes
per image -- is all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests of the host logic) and the ranks of a
group take the same lexicographic minimum, which reproduces the reference's strict-`<`, lowest-index-wins rule
(lib.rs:216) for any slice count.  Every rank of a group then applies the accept rule to its replica of the group's
images, so replicas stay bit-identical without further communication (checked by `state_checksums`).

torch is plumbing here (device buffers, streams, the process group); the arithmetic is libsnesgpu's.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import engine, synth


# ---- schedule (lib.rs:881-933) -------------------------------------------------------------------------
@dataclass
class Cursor:
    """The optimiser cursor of run(): (palette, palette_index, channel, step)."""
    palette: int = 0
    palette_index: int = 0
    channel: int = 0
    step: int = 0

    def mode(self, config: engine.Config) -> str:
        if config.nes:                      # lib.rs:892
            return "nes"
        return "random" if self.step % 5 < 4 else "channel"   # lib.rs:890

    def advance(self, config: engine.Config):
        """lib.rs:917-932"""
        random = self.step % 5 < 4
        self.channel += 1
        if self.channel == 3 or random:
            self.channel = 0
            self.palette_index += 1
            if self.palette_index == config.subpalette_size:
                self.palette_index = 0
                self.palette += 1
                if self.palette == config.subpalette_count:
                    self.palette = 0
                    self.step += 1


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of `rank` out of n items: [rank*n/world, (rank+1)*n/world) (may be empty when n < world)."""
    return (rank * n) // world, ((rank + 1) * n) // world


@dataclass(frozen=True)
class ShardPlan:
    """Where one rank sits in the (image group x candidate slice) grid of a job."""
    nimg: int          # images of the whole job
    rank: int
    world: int
    img_groups: int    # groups the images are split into
    cand_ranks: int    # ranks sharing one group's images, each with a slice of the candidates
    group: int         # this rank's image group
    slice: int         # this rank's candidate slice within the group
    img_lo: int        # this group's images: [img_lo, img_hi)
    img_hi: int
    slots: int         # records per rank in the all-gather (= the largest group; smaller groups pad)

    @property
    def nloc(self) -> int:
        return self.img_hi - self.img_lo

    @property
    def first_rank_of_group(self) -> int:
        return self.group * self.cand_ranks

    def describe(self) -> str:
        return f"{self.img_groups} image groups x {self.cand_ranks} candidate slices"


def plan_shards(nimg: int, rank: int, world: int, mode: str = "hybrid") -> ShardPlan:
    """mode "hybrid": as many image groups as divide the world and do not exceed nimg, the rest of the ranks slice the
    candidates; mode "candidates": one group, every rank holds every image (the replicated layout)."""
    if nimg < 1 or not (0 <= rank < world):
        raise ValueError("plan_shards: bad arguments")
    groups = 1 if mode == "candidates" else max(d for d in range(1, world + 1) if world % d == 0 and d <= nimg)
    cand_ranks = world // groups
    group, sl = rank // cand_ranks, rank % cand_ranks
    lo, hi = shard_bounds(nimg, group, groups)
    return ShardPlan(nimg, rank, world, groups, cand_ranks, group, sl, lo, hi, -(-nimg // groups))


def init_library_comm(ctx: engine.Context, rank: int, world: int, group=None):
    """Give the library its own NCCL communicator over the ranks of `group`: rank 0 makes the unique id (snes_comm_unique_id),
    the process group hands it round, every rank joins (snes_ctx_comm_init).  After this snes_dist_step_random runs a whole
    sharded step, all-gather included, in one C call -- what a host without torch (the reference is Rust) would use."""
    import torch.distributed as dist
    box = [engine.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(box[0], rank, world)


def merge_best_host(gathered: np.ndarray) -> np.ndarray:
    """Host restatement of k_merge_best for the gloo tests: gathered is (R, nimg) of BEST_DTYPE with
    global candidate indices; returns the lexicographic minimum of (err, idx) per image."""
    out = gathered[0].copy()
    for r in range(1, gathered.shape[0]):
        o = gathered[r]
        take = (o["idx"] >= 0) & ((out["idx"] < 0) | (o["err"] < out["err"]) | ((o["err"] == out["err"]) & (o["idx"] < out["idx"])))
        out[take] = o[take]
    return out


class BatchOptimizer:
    """A batch of independent OptimizedImages advancing through the reference's schedule together.  `images` are the
    images THIS rank holds (plan.img_lo .. plan.img_hi of the job); with the default single-rank plan, all of them."""

    def __init__(self, ctx: engine.Context, images: Sequence[engine.OptimizedImage], plan: Optional[ShardPlan] = None,
                 group=None, seed: int = 0):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.images = list(images)
        self.plan = plan or plan_shards(len(self.images), 0, 1)
        if len(self.images) != self.plan.nloc:
            raise ValueError(f"rank {self.plan.rank} holds {len(self.images)} images, its plan says {self.plan.nloc}")
        self.config = self.images[0].config
        self.group = group
        self.seed = seed
        self.cursor = Cursor()
        self.iteration = 0
        self.device = torch.device("cuda", ctx.device)
        slots, world = self.plan.slots, self.plan.world
        self._best_local = torch.zeros(slots * 2, dtype=torch.int64, device=self.device)   # slots x 16 bytes
        self._best_all = torch.zeros(world * slots * 2, dtype=torch.int64, device=self.device)
        self._best = torch.zeros(slots * 2, dtype=torch.int64, device=self.device)
        # a rank that owns its images outright (cand_ranks == 1) needs nobody's records to go on: its all-gather only
        # publishes them, so it runs asynchronously on the process group's stream from one of two send buffers
        self._send = [torch.zeros(slots * 2, dtype=torch.int64, device=self.device) for _ in range(2)]
        self._pending = [None, None]
        self.library_comm = False    # step_random_host through snes_dist_step_random (after init_library_comm)
        # enqueue library work on torch's current stream so it orders with the collectives and is seen by
        # torch.cuda.Event timing; torch reports the legacy default stream as handle 0, which the C ABI reads as
        # "own stream", so name it explicitly (cudaStreamLegacy == 0x1)
        ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream or 1)

    # ---- candidate lists ---------------------------------------------------------------------------
    def candidates_host(self, ncand_total: int) -> np.ndarray:
        """The explicit, seeded stand-in for the 64 rand::rng() draws of lib.rs:205-208, per local image (seeded by the
        image's number in the whole job, so the lists do not depend on how the job is sharded)."""
        return np.stack([synth.candidates(self.seed * 1000003 + self.plan.img_lo + j, self.iteration, ncand_total)
                         for j in range(len(self.images))])

    def _gathered_group_ptr(self) -> int:
        """Device address of the gathered records of this rank's group (cand_ranks blocks of `slots` records)."""
        return self._best_all.data_ptr() + self.plan.first_rank_of_group * self.plan.slots * 16

    def _exchange(self):
        """The step's one collective: all-gather of every rank's records (16 bytes per image slot)."""
        self.torch.distributed.all_gather_into_tensor(self._best_all, self._best_local, group=self.group)

    def _owns_images(self) -> bool:
        return self.plan.world > 1 and self.plan.cand_ranks == 1

    def _send_buffer(self):
        """The send buffer of this step; the collective that last read it (two steps ago) is ordered before its reuse."""
        k = self.iteration & 1
        if self._pending[k] is not None:
            self._pending[k].wait()          # a stream-side wait, not a host one
            self._pending[k] = None
        return self._send[k]

    def _publish(self, buf):
        self._pending[self.iteration & 1] = self.torch.distributed.all_gather_into_tensor(self._best_all, buf, group=self.group, async_op=True)

    def drain(self):
        """Order every outstanding all-gather before whatever the current stream does next (e.g. reading gathered_records)."""
        for k in (0, 1):
            if self._pending[k] is not None:
                self._pending[k].wait()
                self._pending[k] = None

    def gathered_records(self) -> np.ndarray:
        """The whole job's winning records of the last step as every rank holds them after the all-gather: (world, slots)."""
        self.drain()
        return self._best_all.cpu().numpy().view(engine.BEST_DTYPE).reshape(self.plan.world, self.plan.slots)

    # ---- one optimize_palette_entry_random over the batch, device-resident inputs ----------------------
    def begin_step_dev(self, d_cand, ncand_total: int):
        """First half: `best_error = self.error()` (lib.rs:199) + this rank's share of the candidate loop (lib.rs:205-220)
        in one pass; the slice is read in place from the full list and the records carry indices into the full list."""
        pl = self.plan
        p, i = self.cursor.palette, self.cursor.palette_index
        lo, hi = shard_bounds(ncand_total, pl.slice, pl.cand_ranks)
        out = self._best if pl.world == 1 else self._best_local
        engine.batch_error_eval_candidates_slice_dev(self.images, p, i, d_cand.data_ptr(), ncand_total, lo, hi - lo, None, out.data_ptr())

    def end_step_dev(self, d_cand, ncand_total: int):
        """Second half, after the exchange: merge the records of this group's ranks, accept if strictly better, optimize()
        with the winner (lib.rs:216-219, 236-237)."""
        pl = self.plan
        p, i = self.cursor.palette, self.cursor.palette_index
        if pl.world > 1:
            engine.merge_best_dev(self.ctx, self._gathered_group_ptr(), pl.cand_ranks, pl.nloc, self._best.data_ptr(), rank_stride=pl.slots)
        engine.batch_apply_best_dev(self.images, p, i, d_cand.data_ptr(), ncand_total, self._best.data_ptr())
        self.cursor.advance(self.config)
        self.iteration += 1

    def step_random_dev(self, d_cand, ncand_total: int):
        """d_cand: uint8 CUDA tensor (local images, ncand_total, 3): the full candidate list of this rank's images,
        identical on the ranks of its group."""
        if self._owns_images():
            pl = self.plan
            p, i = self.cursor.palette, self.cursor.palette_index
            buf = self._send_buffer()
            engine.batch_error_eval_candidates_slice_dev(self.images, p, i, d_cand.data_ptr(), ncand_total, 0, ncand_total, None, buf.data_ptr())
            self._publish(buf)               # the NCCL exchange of the step, off the critical path
            engine.batch_apply_best_dev(self.images, p, i, d_cand.data_ptr(), ncand_total, buf.data_ptr())
            self.cursor.advance(self.config)
            self.iteration += 1
            self._best = buf
            return
        self.begin_step_dev(d_cand, ncand_total)
        if self.plan.world > 1:
            self._exchange()
        self.end_step_dev(d_cand, ncand_total)

    # ---- the same through host buffers (the call a user of the library makes) --------------------------
    def step_random_host(self, cand_host: np.ndarray, best_host: Optional[np.ndarray] = None) -> np.ndarray:
        """cand_host: (local images, ncand_total, 3) uint8 host array (pinned or not); returns the winning (err, idx)
        records of the local images in host memory.  Copies in, both halves of the step, the all-gather between them and
        the copy out are all inside this call; it returns when the stream has drained."""
        pl = self.plan
        p, i = self.cursor.palette, self.cursor.palette_index
        ncand_total = cand_host.shape[1]
        if pl.world == 1:
            best, _ = engine.batch_step_random(self.images, p, i, cand_host)
        elif self.library_comm:
            # one C call: candidate H2D, error() + this rank's share, ncclAllGather inside the library, merge, accept, optimize()
            best, _ = engine.dist_step_random(self.images, pl.nimg, p, i, cand_host)
        elif self._owns_images():
            buf = self._send_buffer()
            engine.batch_step_random_shard_begin(self.images, p, i, cand_host, 0, ncand_total, buf.data_ptr())
            self._publish(buf)
            best, _ = engine.batch_step_random_shard_end(self.images, p, i, buf.data_ptr(), 1, pl.slots, best_host)
            self._best = buf
        else:
            lo, hi = shard_bounds(ncand_total, pl.slice, pl.cand_ranks)
            engine.batch_step_random_shard_begin(self.images, p, i, cand_host, lo, hi - lo, self._best_local.data_ptr())
            self._exchange()
            best, _ = engine.batch_step_random_shard_end(self.images, p, i, self._gathered_group_ptr(), pl.cand_ranks, pl.slots, best_host)
        self.cursor.advance(self.config)
        self.iteration += 1
        return best

    def best_records(self) -> np.ndarray:
        return self._best.cpu().numpy().view(engine.BEST_DTYPE)[:self.plan.nloc]

    def state_checksums(self) -> List[int]:
        """64-bit checksum of (palette, tile_palettes, palette_map) of every local image."""
        return [im.state_checksum() for im in self.images]


def sweep_random(images: Sequence[engine.OptimizedImage], seed: int, sweep: int, ncand: int = 64) -> np.ndarray:
    """Opt-in whole-sweep batching (TODO.md:33-35) -- NOT the reference's trajectory.

    The reference improves one palette entry at a time, each against the state the previous one left.  Here every
    entry's `ncand` random candidates are evaluated against ONE base state (C*S independent batches the GPU can run
    back to back without waiting for accept decisions), then per image the best strictly-improving candidate of each
    subpalette is tried in order of predicted gain and kept only if the image's real error() drops (subpalettes own
    disjoint tiles, so their winners rarely interfere; the re-check makes the step monotone regardless).
    Returns the error of every image after the sweep."""
    cfg = images[0].config
    C, S, nimg = cfg.subpalette_count, cfg.subpalette_size, len(images)
    base = engine.batch_error(images)
    winners = [[] for _ in range(nimg)]          # per image: (predicted error, subpalette, index, colour)
    # every entry's candidates in ONE launch sequence (snes_batch_eval_candidates_multi): the kernels read the replaced entry
    # per evaluation, so the C*S lists share one k_tables / prepare / assign / score pass
    steps = [(p, i) for p in range(C) for i in range(S)]
    cand = np.stack([np.stack([synth.candidates(seed * 1000003 + j, sweep * C * S + p * S + i, ncand) for p, i in steps]) for j in range(nimg)])
    r = engine.batch_eval_candidates_multi(images, steps, cand)
    for p in range(C):
        best_p = [None] * nimg
        for i in range(S):
            s = p * S + i
            for j in range(nimg):
                err, k = float(r["best"]["err"][j, s]), int(r["best"]["idx"][j, s])
                if k >= 0 and err < base[j] and (best_p[j] is None or err < best_p[j][0]):
                    best_p[j] = (err, p, i, cand[j, s, k].copy())
        for j in range(nimg):
            if best_p[j] is not None:
                winners[j].append(best_p[j])
    out = np.array(base, dtype=np.float64)
    for j, im in enumerate(images):
        cur = out[j]
        for err, p, i, colour in sorted(winners[j], key=lambda w: w[0]):
            pal = im.palette
            old = pal[p * S + i].copy()
            pal[p * S + i] = colour
            im.palette = pal
            im.optimize()
            e = im.error()
            if e < cur:
                cur = e
            else:                                 # interfered with an earlier winner: take it back
                pal[p * S + i] = old
                im.palette = pal
                im.optimize()
        out[j] = cur
    return out


class HeadlessRunner:
    """`run()` of the reference without the window: initialize_tiles -> recalculate_palettes -> N
    iterations of the optimiser schedule -> JSON (lib.rs:851-853, 987-989, 889-933, 999-1003)."""

    # Phase of run() (lib.rs:825-830): the green button (lib.rs:982-997) moves TileAssignment -> Clustering (which runs
    # recalculate_palettes) -> Optimization; the optimiser only iterates in the last phase (lib.rs:889).
    TILE_ASSIGNMENT, CLUSTERING, OPTIMIZATION = "TileAssignment", "Clustering", "Optimization"

    def __init__(self, ctx: engine.Context, rgba: np.ndarray, config: engine.Config, seed: int = 0, ncand: int = 64,
                 speculate: int = 4):
        """speculate: how many iterations ahead `iterate(n)` may evaluate in one call (1 = one at a time).  One picture's
        64 candidates fill a fraction of a B200; the candidates of the next few entries ride along, and are used up to the
        first iteration that accepts one (snes_image_iterate) -- the trajectory is the reference's either way."""
        self.image = engine.OptimizedImage(ctx, rgba, config)
        self.speculate = max(1, int(speculate))
        self.accept_rate = 0.5      # running estimate of the chance that an iteration accepts a candidate
        self.config = config
        self.cursor = Cursor()
        self.seed, self.ncand = seed, ncand
        self.iteration = 0
        self.last_error = float("inf")
        self.log: List[float] = []
        self.phase = self.TILE_ASSIGNMENT

    def initialize_tiles(self):
        self.image.initialize_tiles()        # lib.rs:851
        self.phase = self.TILE_ASSIGNMENT

    def green_button(self):
        """lib.rs:982-997"""
        if self.phase == self.TILE_ASSIGNMENT:
            self.phase = self.CLUSTERING
            self.image.recalculate_palettes()
        elif self.phase == self.CLUSTERING:
            self.phase = self.OPTIMIZATION

    def click_tile(self, tile_x: int, tile_y: int):
        """A left click on tile (tile_x, tile_y) of either picture (lib.rs:1005-1024): the tile moves to the next
        subpalette; outside the TileAssignment phase the palettes are re-clustered at once."""
        if not (0 <= tile_x < 32 and 0 <= tile_y < 32):
            raise ValueError("tile out of range")
        tp = self.image.tile_palettes
        index = tile_y * 32 + tile_x
        tp[index] = (int(tp[index]) + 1) % self.config.subpalette_count
        self.image.tile_palettes = tp
        if self.phase != self.TILE_ASSIGNMENT:
            self.image.recalculate_palettes()

    def optimize_tile(self, tile_x: int, tile_y: int) -> bool:
        """The automatic version of a tile click (TODO.md:36-37): every other subpalette is tried for the tile
        (optimize() + error() each, one batch on the GPU) and the best is kept if it is strictly better than the
        current error.  Returns whether the tile moved."""
        if not (0 <= tile_x < 32 and 0 <= tile_y < 32):
            raise ValueError("tile out of range")
        index = tile_y * 32 + tile_x
        cur = int(self.image.tile_palettes[index])
        moves = [(index, q) for q in range(self.config.subpalette_count) if q != cur]
        if not moves:
            return False
        r = engine.batch_step_tile_moves([self.image], np.asarray(moves, np.int32)[None])
        return bool(r["applied"][0])

    def initialize(self):
        """What a user does before the optimiser runs: start-up, then the green button twice."""
        self.initialize_tiles()
        self.green_button()                  # -> Clustering: recalculate_palettes, lib.rs:987-989
        self.green_button()                  # -> Optimization

    def resume(self, doc):
        """Continue from a JSON document written by `write_json` / the reference (lib.rs:579-625) instead of running the
        two k-means initialisations: palette and tile assignment are taken from the document, the per-pixel indices
        are re-derived by optimize() exactly as the reference does after every change (TODO.md:38-39)."""
        from . import ingest
        palette, tile_palettes, _palette_map, _transparent = ingest.state_from_json(doc, self.config.subpalette_count,
                                                                                     self.config.subpalette_size)
        self.image.tile_palettes = tile_palettes
        self.image.palette = palette
        self.image.optimize()
        self.phase = self.OPTIMIZATION

    def iterate(self, n: int = 1):
        if self.phase != self.OPTIMIZATION:      # lib.rs:889: the optimiser only runs in the last phase
            return
        im, c = self.image, self.cursor
        done = 0
        while done < n:
            mode = c.mode(self.config)
            # the iterations ahead of the cursor that share this one's mode
            ahead = min(self.lookahead(), n - done)
            steps, cands, look = [], [], Cursor(c.palette, c.palette_index, c.channel, c.step)
            for k in range(ahead):
                if look.mode(self.config) != mode:
                    break
                steps.append((look.palette, look.palette_index, look.channel))
                if mode == "random":
                    cands.append(synth.candidates(self.seed, self.iteration + k, self.ncand))
                look.advance(self.config)
            # optimize_palette_entry_* + optimize() + error() (lib.rs:892-910) of `used` iterations in one call
            used, before, error = im.iterate(mode, steps, np.stack(cands) if cands else None)
            # did the run end on an iteration that changed the palette?  (a NES run that used every step may have: counted as not)
            accepted = 1.0 if (used < len(steps) or (mode != "nes" and error != before)) else 0.0
            self.accept_rate += 0.25 * (accepted / used - self.accept_rate)
            # lib.rs:912-915 per iteration: the iterations before the last consumed one ended with the error they started from
            for e in ([before] if used > 1 and before == before else []) + [error]:
                if abs(e - self.last_error) > np.finfo(np.float64).eps:
                    self.last_error = e
                    self.log.append(e)
            for _ in range(used):
                c.advance(self.config)
            self.iteration += used
            done += used

    def lookahead(self) -> int:
        """How many iterations to evaluate in the next call.  A call costs about c0 + c1 * K (measured on a B200: c0 / c1 = 2.6
        without dithering, 5 with: the serial wavefront of one picture is long and K of them run side by side) and stands
        for (1 - (1 - a)^K) / a iterations when each accepts with probability a; take the K with the least cost per iteration."""
        if self.speculate <= 1:
            return 1
        a = min(0.95, max(0.01, self.accept_rate))
        ratio = 5.0 if self.config.dither else 2.6
        best_k, best_cost = 1, None
        for k in (1, 2, 3, 4, 6, 8, 12, 16):
            if k > 4 * self.speculate:
                break
            cost = (ratio + k) * a / (1.0 - (1.0 - a) ** k)
            if best_cost is None or cost < best_cost:
                best_k, best_cost = k, cost
        return best_k

    def write_json(self, path: str):
        with open(path, "w") as f:           # lib.rs:999-1003
            f.write(self.image.as_json_string())
