// integration/ffi.rs -- what src/ffi.rs of aexoden/snesimage becomes.
// Written against include/snesgpu.h; NOT compiled in this repository's image (no cargo / rustc there).
// src/ffi.rs -- raw bindings to include/snesgpu.h
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct SnesCtx { _p: [u8; 0] }
#[repr(C)] pub struct SnesImage { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct SnesConfig {            // config::Config, src/config.rs:13-30
    pub subpalette_count: i32,
    pub subpalette_size: i32,
    pub dither: u8,
    pub perceptual_palettes: u8,
    pub nes: u8,
    pub reserved: u8,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct SnesBest { pub err: f64, pub idx: i32, pub pad: i32 }

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct SnesStep { pub palette: i32, pub index: i32, pub channel: i32, pub reserved: i32 }   // one iteration of lib.rs:889-933

#[link(name = "snesgpu")]
unsafe extern "C" {
    pub fn snes_last_error() -> *const c_char;
    pub fn snes_ctx_create(device: c_int, out: *mut *mut SnesCtx) -> c_int;
    pub fn snes_ctx_destroy(ctx: *mut SnesCtx);
    pub fn snes_image_new(ctx: *mut SnesCtx, rgba: *const u8, width: c_int, height: c_int,
                          cfg: *const SnesConfig, out: *mut *mut SnesImage) -> c_int;      // lib.rs:46-65
    pub fn snes_image_free(im: *mut SnesImage);
    pub fn snes_image_initialize_tiles(im: *mut SnesImage) -> c_int;                         // lib.rs:79-189
    pub fn snes_image_recalculate_palettes(im: *mut SnesImage) -> c_int;                     // lib.rs:407-415
    pub fn snes_image_optimize(im: *mut SnesImage) -> c_int;                                 // lib.rs:425-501
    pub fn snes_image_error(im: *mut SnesImage, err: *mut f64) -> c_int;                     // lib.rs:503-548
    pub fn snes_image_as_rgba(im: *mut SnesImage, out: *mut u8) -> c_int;                    // lib.rs:550-577
    pub fn snes_image_as_json(im: *mut SnesImage, buf: *mut c_char, cap: usize, len: *mut usize) -> c_int; // 579-625
    pub fn snes_image_optimize_palette_entry_random(im: *mut SnesImage, palette: c_int, index: c_int,
                                                    cand: *const u8, ncand: c_int) -> c_int; // lib.rs:191-240
    pub fn snes_image_optimize_palette_entry_nes(im: *mut SnesImage, palette: c_int, index: c_int) -> c_int;      // 242-284
    pub fn snes_image_optimize_palette_entry_channel(im: *mut SnesImage, palette: c_int, index: c_int,
                                                     channel: c_int) -> c_int;               // lib.rs:286-328
    pub fn snes_image_get_palette(im: *mut SnesImage, out: *mut u8) -> c_int;
    pub fn snes_image_get_tile_palettes(im: *mut SnesImage, out: *mut u8) -> c_int;
    pub fn snes_image_set_tile_palettes(im: *mut SnesImage, data: *const u8) -> c_int;       // mouse clicks, lib.rs:1005-1024
    pub fn snes_image_get_palette_map(im: *mut SnesImage, out: *mut u8) -> c_int;
    pub fn snes_image_state_checksum(im: *mut SnesImage, out: *mut u64) -> c_int;
    // nsteps consecutive iterations of the loop body of run() (lib.rs:889-910) in one call, evaluated ahead against the
    // current state and taken up to the first one that accepts a candidate: the reference's trajectory, fewer launches
    pub fn snes_image_iterate(im: *mut SnesImage, mode: c_int, steps: *const SnesStep, nsteps: c_int, cand: *const u8,
                              ncand: c_int, consumed: *mut c_int, error_before: *mut f64, error_after: *mut f64) -> c_int;
    // verified 256-entry sRGB -> linear tables (yuvxyb's, palette's) in place of the built-in ones; before any image exists
    pub fn snes_ctx_set_transfer_luts(ctx: *mut SnesCtx, yuvxyb_eotf: *const f32, palette_eotf: *const f32) -> c_int;
    // diagnostics: the kernels' cube root against the restated yuvxyb-math cbrtf on every f32 bit pattern in [lo_bits, hi_bits);
    // `mismatches` must come back 0 (a maintainer with the crate can also feed the same range through yuvxyb_math::cbrtf)
    pub fn snes_ctx_cbrt_selfcheck(ctx: *mut SnesCtx, lo_bits: u32, hi_bits: u32, mismatches: *mut u64, fallbacks: *mut u64) -> c_int;
}
