"""`python -m snesimage_b200 SOURCE TARGET [options]` -- the reference's command line (/root/reference/src/config.rs:3-31,
src/main.rs:9-20) driving the headless schedule of `run()` (lib.rs:833-1041 without the SDL window): the same two
positionals and five options, plus what a window-less run needs (how long to optimise, the seed of the explicit
candidate lists that stand in for the unseeded `rand::rng()` of lib.rs:201, resume).  Writes the reference's JSON
document (lib.rs:579-625, 999-1003).  Needs an sm_100 GPU: there is no CPU fallback.
"""
from __future__ import annotations

import argparse
import sys
import time

from . import driver, engine, ingest


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="snesimage_b200", description="B200-native palette optimiser for aexoden/snesimage")
    ap.add_argument("source_filename", help="Filename for the image to optimize.")
    ap.add_argument("target_filename", help="Filename for the generated JSON output.")
    ap.add_argument("-c", "--subpalette-count", type=int, default=1, help="Number of separate subpalettes to use.")
    ap.add_argument("-s", "--subpalette-size", type=int, default=7,
                    help="Number of colors within each subpalette (not including the transparent color).")
    ap.add_argument("-d", "--dither", action="store_true", help="Whether to dither the output.")
    ap.add_argument("--perceptual-palettes", action="store_true",
                    help="Whether to use more expensive CIELAB-based-computations for color comparisons.")
    ap.add_argument("--nes", action="store_true",
                    help="Enables a special mode that uses only colors similar to those available on the NES.")
    # headless-only knobs (the reference advances while its window is open and stops on a mouse click)
    ap.add_argument("--iterations", type=int, default=None, help="optimiser iterations to run (default: --sweeps 1)")
    ap.add_argument("--sweeps", type=float, default=1.0, help="full sweeps over all palette entries (one sweep = count x size iterations)")
    ap.add_argument("--seed", type=int, default=0, help="seed of the explicit candidate lists (the reference draws unseeded)")
    ap.add_argument("--candidates", type=int, default=64, help="random trial colours per iteration (lib.rs:205)")
    ap.add_argument("--resume", default=None, help="JSON written by an earlier run: continue from its palette / tile assignment")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--quiet", action="store_true")
    return ap


def config_from_args(args) -> engine.Config:
    return engine.Config(source_filename=args.source_filename, target_filename=args.target_filename,
                         subpalette_count=args.subpalette_count, subpalette_size=args.subpalette_size, dither=args.dither,
                         perceptual_palettes=args.perceptual_palettes, nes=args.nes)


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    config = config_from_args(args)
    say = (lambda *a: None) if args.quiet else (lambda *a: print(*a, flush=True))
    say(f"Using source image: {args.source_filename}")                     # lib.rs:834
    try:
        rgba = ingest.load_rgba(args.source_filename)
        ctx = engine.Context(args.device)
        runner = driver.HeadlessRunner(ctx, rgba, config, seed=args.seed, ncand=args.candidates)
        if args.resume:
            with open(args.resume) as f:
                runner.resume(f.read())
        else:
            runner.initialize()
        say("Finished assigning initial tiles")                            # lib.rs:131
        n = args.iterations if args.iterations is not None else int(round(args.sweeps * config.subpalette_count * config.subpalette_size))
        t0 = time.perf_counter()
        for done in range(n):
            before = len(runner.log)
            runner.iterate(1)
            if len(runner.log) > before:
                say(f"Current Error: {runner.last_error}")                 # lib.rs:912-915
        dt = time.perf_counter() - t0
        runner.write_json(args.target_filename)
        say(f"{n} iterations in {dt:.2f} s; error {runner.image.error():.9f}; wrote {args.target_filename}")
    except (engine.SnesGpuError, ValueError, OSError) as exc:
        print(f"[ERROR] {exc}", file=sys.stderr)                           # main.rs:16-19: log the error, exit(1)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
