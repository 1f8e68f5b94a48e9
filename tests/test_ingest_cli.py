"""Data formats either side of the hot path: image ingest, the JSON document and its
inverse (resume), the reference's command line.  CPU tests use the oracle as the checker; the end-to-end CLI test
needs a GPU."""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob
from snesimage_b200 import ingest, synth
from snesimage_b200.__main__ import build_parser, config_from_args, main


def test_load_rgba_round_trip_and_size_check(tmp_path):
    from PIL import Image
    rgba = synth.image(3, "T")                      # has fully transparent tiles
    p = tmp_path / "img.png"
    Image.fromarray(rgba, "RGBA").save(p)
    got = ingest.load_rgba(str(p))
    assert got.dtype == np.uint8 and got.shape == (256, 256, 4) and got.flags["C_CONTIGUOUS"]
    assert np.array_equal(got, rgba)
    # an RGB file gains alpha = 255 (image::DynamicImage::into_rgba8)
    Image.fromarray(rgba[:, :, :3].copy(), "RGB").save(tmp_path / "rgb.png")
    got = ingest.load_rgba(str(tmp_path / "rgb.png"))
    assert np.array_equal(got[:, :, :3], rgba[:, :, :3]) and np.all(got[:, :, 3] == 255)
    # lib.rs:838-840: both sides wrong -> the reference's message; one side wrong -> refused as well (the reference would not)
    Image.new("RGBA", (128, 64)).save(tmp_path / "small.png")
    with pytest.raises(ValueError, match="^Image size must be 256x256$"):
        ingest.load_rgba(str(tmp_path / "small.png"))
    Image.new("RGBA", (256, 64)).save(tmp_path / "half.png")
    with pytest.raises(ValueError, match="256x64"):
        ingest.load_rgba(str(tmp_path / "half.png"))


@pytest.mark.parametrize("family,C,S,dither", [("V", 4, 7, False), ("T", 8, 15, False), ("B", 3, 3, True)])
def test_state_from_json_inverts_as_json(family, C, S, dither):
    """as_json (lib.rs:579-625, restated by the oracle) -> state_from_json gives back palette, tile_palettes and
    palette_map, with transparent pixels marked."""
    rgba = synth.image(11, family)
    o = ob.OracleImage(rgba, C, S, dither=dither)
    o.initialize_tiles()
    o.recalculate_palettes()
    doc = o.as_json()
    palette, tile_palettes, palette_map, transparent = ingest.state_from_json(json.dumps(doc), C, S)
    assert np.array_equal(tile_palettes, o.tile_palettes)
    want_pal = o.palette.reshape(-1, 3)
    keep = np.all(want_pal < 32, axis=1)            # a component of 32 does not survive as_u16 (lib.rs:679-681)
    assert np.array_equal(palette[keep], want_pal[keep])
    assert np.array_equal(transparent.reshape(256, 256), rgba[:, :, 3] == 0)
    assert np.array_equal(palette_map, o.palette_map.reshape(-1))   # transparent pixels hold 0 in both
    # shape / range errors
    bad = dict(doc)
    bad["palette"] = doc["palette"][:-1]
    with pytest.raises(ValueError):
        ingest.state_from_json(bad, C, S)
    bad = dict(doc)
    bad["tile_palettes"] = [C] * 1024
    with pytest.raises(ValueError):
        ingest.state_from_json(bad, C, S)


def test_cli_mirrors_config_rs():
    """config.rs:3-31: two positionals, -c/--subpalette-count (1), -s/--subpalette-size (7), -d/--dither,
    --perceptual-palettes, --nes."""
    a = build_parser().parse_args(["in.png", "out.json"])
    c = config_from_args(a)
    assert (c.source_filename, c.target_filename) == ("in.png", "out.json")
    assert (c.subpalette_count, c.subpalette_size, c.dither, c.perceptual_palettes, c.nes) == (1, 7, False, False, False)
    a = build_parser().parse_args(["-c", "8", "-s", "15", "-d", "--perceptual-palettes", "--nes", "a", "b"])
    c = config_from_args(a)
    assert (c.subpalette_count, c.subpalette_size, c.dither, c.perceptual_palettes, c.nes) == (8, 15, True, True, True)
    a = build_parser().parse_args(["--subpalette-count", "4", "--subpalette-size", "3", "--dither", "a", "b"])
    assert (a.subpalette_count, a.subpalette_size, a.dither) == (4, 3, True)


def test_cli_reports_errors_like_main_rs(tmp_path, capsys):
    """main.rs:16-19: the error is logged and the process exits with 1."""
    assert main([str(tmp_path / "missing.png"), str(tmp_path / "o.json"), "--quiet"]) == 1
    assert "[ERROR]" in capsys.readouterr().err


@pytest.mark.gpu
def test_cli_end_to_end_and_resume(tmp_path):
    """SOURCE -> JSON through the command line equals the oracle driven through the same schedule; a second run
    resumed from that JSON continues from the same state."""
    from PIL import Image
    from snesimage_b200 import driver
    rgba = synth.image(21, "V")
    src, out1, out2 = tmp_path / "src.png", tmp_path / "out1.json", tmp_path / "out2.json"
    Image.fromarray(rgba, "RGBA").save(src)
    C, S, n = 2, 3, 4
    assert main([str(src), str(out1), "-c", str(C), "-s", str(S), "--iterations", str(n), "--seed", "5", "--candidates", "8", "--quiet"]) == 0
    o = ob.OracleImage(rgba, C, S)
    o.initialize_tiles()
    o.recalculate_palettes()
    cur = driver.Cursor()
    cfg = driver.engine.Config(subpalette_count=C, subpalette_size=S)
    for it in range(n):
        assert cur.mode(cfg) == "random"
        o.optimize_palette_entry_random(cur.palette, cur.palette_index, synth.candidates(5, it, 8))
        o.optimize()
        cur.advance(cfg)
    with open(out1) as f:
        doc = json.load(f)
    assert doc == o.as_json()
    # resume: no k-means, the first iteration starts from the document's state
    assert main([str(src), str(out2), "-c", str(C), "-s", str(S), "--iterations", "0", "--resume", str(out1), "--quiet"]) == 0
    with open(out2) as f:
        assert json.load(f) == doc


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw,n", [
    ("cfg1 rgb", dict(subpalette_count=3, subpalette_size=4), 8),
    ("cfg3 dither", dict(subpalette_count=3, subpalette_size=4, dither=True), 6),
    ("cfg4 nes dither", dict(subpalette_count=2, subpalette_size=3, nes=True, dither=True), 3),
])
def test_headless_trajectory_matches_oracle(ctx, name, kw, n):
    """k-means init + tile assignment + n iterations of the schedule of run() (lib.rs:851-853, 889-933) on the GPU and
    in the oracle: identical palettes, tile assignment, palette_map and JSON after every iteration (integer outputs are
    bit-exact; the errors that steer the search agree to 1e-8, far below the gaps between candidates)."""
    from snesimage_b200 import driver, engine
    rgba = synth.image(33, "B")
    cfg = engine.Config(**kw)
    r = driver.HeadlessRunner(ctx, rgba, cfg, seed=9, ncand=6)
    o = ob.OracleImage(rgba, cfg.subpalette_count, cfg.subpalette_size, cfg.dither, cfg.perceptual_palettes, cfg.nes)
    r.initialize()
    o.initialize_tiles()
    o.recalculate_palettes()
    cur = driver.Cursor()
    for it in range(n):
        mode = cur.mode(cfg)
        if mode == "nes":
            o.optimize_palette_entry_nes(cur.palette, cur.palette_index)
        elif mode == "random":
            o.optimize_palette_entry_random(cur.palette, cur.palette_index, synth.candidates(9, it, 6))
        else:
            o.optimize_palette_entry_channel(cur.palette, cur.palette_index, cur.channel)
        o.optimize()
        cur.advance(cfg)
        r.iterate(1)
        assert np.array_equal(r.image.palette, o.palette), (name, it)
        assert np.array_equal(r.image.palette_map, o.palette_map), (name, it)
        assert abs(r.image.error() - o.error()) <= 1e-8, (name, it)
    assert np.array_equal(r.image.tile_palettes, o.tile_palettes)
    assert r.image.as_json() == o.as_json()
    assert (r.cursor.palette, r.cursor.palette_index, r.cursor.step) == (cur.palette, cur.palette_index, cur.step)
    r.image.close()


@pytest.mark.gpu
def test_phases_and_tile_clicks_match_oracle(ctx):
    """The interactive part of run() restated headless: tile clicks (lib.rs:1005-1024) before and after the first
    green button (lib.rs:982-997) give the oracle's state."""
    from snesimage_b200 import driver, engine
    rgba = synth.image(44, "B")
    C, S = 3, 4
    cfg = engine.Config(subpalette_count=C, subpalette_size=S)
    r = driver.HeadlessRunner(ctx, rgba, cfg)
    o = ob.OracleImage(rgba, C, S)
    r.initialize_tiles()
    o.initialize_tiles()
    assert r.phase == r.TILE_ASSIGNMENT

    def o_click(tx, ty, recluster):
        tp = o.tile_palettes
        tp[ty * 32 + tx] = (int(tp[ty * 32 + tx]) + 1) % C
        o.tile_palettes = tp
        if recluster:
            o.recalculate_palettes()

    r.click_tile(5, 7)          # TileAssignment: only the assignment changes
    o_click(5, 7, False)
    assert np.array_equal(r.image.tile_palettes, o.tile_palettes) and np.array_equal(r.image.palette, o.palette)
    r.green_button()            # -> Clustering
    o.recalculate_palettes()
    assert r.phase == r.CLUSTERING
    assert np.array_equal(r.image.palette, o.palette) and np.array_equal(r.image.palette_map, o.palette_map)
    r.click_tile(31, 0)         # later phases: re-cluster at once
    r.click_tile(31, 0)
    o_click(31, 0, True)
    o_click(31, 0, True)
    assert np.array_equal(r.image.tile_palettes, o.tile_palettes)
    assert np.array_equal(r.image.palette, o.palette) and np.array_equal(r.image.palette_map, o.palette_map)
    r.green_button()
    assert r.phase == r.OPTIMIZATION
    r.green_button()
    assert r.phase == r.OPTIMIZATION
    with pytest.raises(ValueError):
        r.click_tile(32, 0)
    r.image.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(dither=True), dict(perceptual_palettes=True)])
def test_tile_moves_match_oracle(ctx, kw):
    """Tile reassignment candidates: score of `tile_palettes[t] = q; optimize(); error()` for a batch of
    moves equals the oracle's, the accept rule takes the strict best, and the runner's optimize_tile applies it."""
    from snesimage_b200 import driver, engine
    rgba = synth.image(61, "B")
    C, S = 3, 4
    cfg = engine.Config(subpalette_count=C, subpalette_size=S, **kw)
    lab = bool(kw.get("perceptual_palettes"))
    r = driver.HeadlessRunner(ctx, rgba, cfg)
    o = ob.OracleImage(rgba, C, S, cfg.dither, cfg.perceptual_palettes, cfg.nes)
    r.initialize()
    o.initialize_tiles()
    o.recalculate_palettes()
    tp0 = o.tile_palettes.copy()
    tiles = [0, 37, 500, 1023]
    moves = [(t, (int(tp0[t]) + d) % C) for t in tiles for d in (1, 2)]
    want = []
    for t, q in moves:
        tp = tp0.copy()
        tp[t] = q
        o.tile_palettes = tp
        o.optimize()
        want.append((o.error(), o.palette_map.copy()))
    o.tile_palettes = tp0
    o.optimize()
    got = engine.batch_eval_tile_moves([r.image], np.asarray(moves, np.int32)[None], want_maps=not lab)
    tol = 1e-4 if lab else 1e-8
    assert np.max(np.abs(got["scores"][0] - np.array([w[0] for w in want]))) <= tol
    if not lab:
        for k, w in enumerate(want):
            assert np.array_equal(got["maps"][0, k], w[1]), moves[k]
    assert np.array_equal(r.image.tile_palettes, tp0)              # evaluation leaves the state alone
    k = int(np.argmin(got["scores"][0]))
    assert got["best"]["idx"][0] == k
    # accept: strictly better than the current error, else nothing changes
    cur = o.error()
    s = engine.batch_step_tile_moves([r.image], np.asarray(moves, np.int32)[None])
    take = want[k][0] < cur
    assert bool(s["applied"][0]) == take
    tp = tp0.copy()
    if take:
        tp[moves[k][0]] = moves[k][1]
    assert np.array_equal(r.image.tile_palettes, tp)
    o.tile_palettes = tp
    o.optimize()
    if not lab:
        assert np.array_equal(r.image.palette_map, o.palette_map)
    assert abs(r.image.error() - o.error()) <= tol
    # the runner's automatic tile click
    before = r.image.error()
    moved = r.optimize_tile(5, 9)
    assert r.image.error() <= before and (moved == (r.image.error() < before))
    with pytest.raises(engine.SnesGpuError):
        engine.batch_eval_tile_moves([r.image], np.asarray([[(2000, 0)]], np.int32))
    r.image.close()


@pytest.mark.gpu
def test_whole_sweep_batching_is_monotone_and_consistent(ctx):
    """Opt-in whole-sweep batching: not the reference's trajectory, so what is checked is what it
    promises -- the error never rises, and the state it leaves is a state of the reference's model: the oracle, given the
    same palette and tile assignment, derives the same palette_map and error."""
    from snesimage_b200 import driver, engine
    C, S = 3, 4
    cfg = engine.Config(subpalette_count=C, subpalette_size=S)
    rgbas = [synth.image(71 + j, "V") for j in range(2)]
    imgs = [engine.OptimizedImage(ctx, r, cfg) for r in rgbas]
    engine.batch_initialize_tiles(imgs)
    engine.batch_recalculate_palettes(imgs)
    before = engine.batch_error(imgs)
    after = driver.sweep_random(imgs, seed=3, sweep=0, ncand=8)
    assert np.all(after <= before)
    assert np.any(after < before)
    again = engine.batch_error(imgs)
    assert np.array_equal(again, after)
    for r, im in zip(rgbas, imgs):
        o = ob.OracleImage(r, C, S)
        o.tile_palettes = im.tile_palettes
        o.palette = im.palette
        o.optimize()
        assert np.array_equal(o.palette_map, im.palette_map)
        assert abs(o.error() - im.error()) <= 1e-8
        im.close()
