// integration/optimized_image.rs -- `struct OptimizedImage` of src/lib.rs:33-626 as a handle over libsnesgpu (INTEGRATION.md
// section 2): same method names, same anyhow::Result error behaviour, the bodies routed through the C ABI.  Everything else
// in src/lib.rs (run(), the SDL window, Palette::render) keeps calling these methods as it does today.
// NOT compiled in this repository's image (no cargo / rustc there).
use anyhow::Context as _;
use rand::distr::{Distribution, Uniform};

use crate::ffi;

struct PaletteView { sub_count: usize, sub_size: usize }

struct OptimizedImage {
    ctx: *mut ffi::SnesCtx,
    im: *mut ffi::SnesImage,
    width: usize, height: usize,
    palette: PaletteView,          // sub_count / sub_size only; colours are fetched for rendering
}

fn check(rc: i32) -> anyhow::Result<()> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(ffi::snes_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow::anyhow!(msg))
}

impl OptimizedImage {
    pub fn new(original: &image::RgbaImage, sub_count: usize, sub_size: usize,
               dither: bool, perceptual_palettes: bool, nes: bool) -> anyhow::Result<Self> {   // lib.rs:46-65
        let cfg = ffi::SnesConfig { subpalette_count: sub_count as i32, subpalette_size: sub_size as i32,
                                    dither: dither as u8, perceptual_palettes: perceptual_palettes as u8,
                                    nes: nes as u8, reserved: 0 };
        let (mut ctx, mut im) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(unsafe { ffi::snes_ctx_create(0, &mut ctx) })?;
        check(unsafe { ffi::snes_image_new(ctx, original.as_raw().as_ptr(), original.width() as i32,
                                           original.height() as i32, &cfg, &mut im) })?;
        Ok(Self { ctx, im, width: 256, height: 256, palette: PaletteView { sub_count, sub_size } })
    }
    pub fn initialize_tiles(&mut self) -> anyhow::Result<()> {
        check(unsafe { ffi::snes_image_initialize_tiles(self.im) }).context("Unable to optimize image")
    }
    pub fn recalculate_palettes(&mut self) -> anyhow::Result<()> {
        check(unsafe { ffi::snes_image_recalculate_palettes(self.im) }).context("Unable to optimize image")
    }
    pub fn optimize(&mut self) -> anyhow::Result<()> { check(unsafe { ffi::snes_image_optimize(self.im) }) }
    pub fn error(&self) -> anyhow::Result<f64> {
        let mut e = 0.0;
        check(unsafe { ffi::snes_image_error(self.im, &mut e) }).context("Failed to compute SSIMULACRA2")?;
        Ok(e)
    }
    pub fn optimize_palette_entry_random(&mut self, palette: usize, index: usize) -> anyhow::Result<()> {
        // the 64 draws of lib.rs:201-208 stay on this side of the boundary
        let mut rng = rand::rng();
        let distribution = Uniform::new(0u8, 32)?;
        let mut cand = [0u8; 64 * 3];
        for c in cand.iter_mut() { *c = distribution.sample(&mut rng); }
        check(unsafe { ffi::snes_image_optimize_palette_entry_random(self.im, palette as i32, index as i32,
                                                                     cand.as_ptr(), 64) })
    }
    pub fn optimize_palette_entry_nes(&mut self, palette: usize, index: usize) -> anyhow::Result<()> {
        check(unsafe { ffi::snes_image_optimize_palette_entry_nes(self.im, palette as i32, index as i32) })
    }
    pub fn optimize_palette_entry_channel(&mut self, palette: usize, index: usize, channel: usize) -> anyhow::Result<()> {
        check(unsafe { ffi::snes_image_optimize_palette_entry_channel(self.im, palette as i32, index as i32, channel as i32) })
    }
    pub fn as_rgba(&self) -> Vec<rgb::RGBA8> {
        let mut out = vec![rgb::RGBA8::default(); 65536];
        unsafe { ffi::snes_image_as_rgba(self.im, out.as_mut_ptr() as *mut u8) };
        out
    }
    pub fn as_json(&self) -> serde_json::Value {       // the library emits serde_json's own compact text
        let mut len = 0usize;
        unsafe { ffi::snes_image_as_json(self.im, std::ptr::null_mut(), 0, &mut len) };
        let mut buf = vec![0u8; len + 1];
        unsafe { ffi::snes_image_as_json(self.im, buf.as_mut_ptr() as *mut _, len + 1, &mut len) };
        serde_json::from_slice(&buf[..len]).expect("libsnesgpu emits valid JSON")
    }
}
impl Drop for OptimizedImage {
    fn drop(&mut self) { unsafe { ffi::snes_image_free(self.im); ffi::snes_ctx_destroy(self.ctx); } }
}

// The loop body of run() (lib.rs:889-933) for `steps.len()` iterations in one call.  `steps` are the cursor positions
// ahead of the current one; the 64 draws per iteration stay on this side of the boundary.  Returns how many iterations
// the call stands for: the caller advances its cursor by that many (lib.rs:917-932) and draws afresh for the rest.
impl OptimizedImage {
    pub fn iterate_random(&mut self, steps: &[(usize, usize)]) -> anyhow::Result<(usize, f64)> {
        let mut rng = rand::rng();
        let distribution = Uniform::new(0u8, 32)?;
        let ffi_steps: Vec<ffi::SnesStep> = steps.iter()
            .map(|&(p, i)| ffi::SnesStep { palette: p as i32, index: i as i32, channel: 0, reserved: 0 }).collect();
        let cand: Vec<u8> = (0..steps.len() * 64 * 3).map(|_| distribution.sample(&mut rng)).collect();
        let (mut used, mut before, mut after) = (0i32, 0f64, 0f64);
        check(unsafe { ffi::snes_image_iterate(self.im, 0, ffi_steps.as_ptr(), ffi_steps.len() as i32, cand.as_ptr(), 64,
                                               &mut used, &mut before, &mut after) })
            .context("Unable to optimize palette with the random method")?;
        Ok((used as usize, after))
    }
}
