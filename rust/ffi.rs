// src/ffi.rs for thomasantony/splat: the only `unsafe` in the crate.
// NOT compiled in this repository: the image has no Rust toolchain.  Binds include/splat.h via bindgen.
#![allow(non_camel_case_types, non_upper_case_globals, dead_code)]
include!(concat!(env!("OUT_DIR"), "/splat_sys.rs"));

use crate::{camera::Camera, gaussians::{Gaussian, GaussianList}};
use std::{cell::Cell, ffi::CStr, ptr};

pub struct B200 { ctx: *mut splat_ctx, uploaded: Cell<bool> }

impl B200 {
    /// lowpass: 0.01 for Pipeline01 (gaussians.rs:156-157), 0.3 for Pipeline02 (:517-518)
    pub fn new(device: i32, lowpass: f32) -> Self {
        let mut cfg = unsafe { std::mem::zeroed::<splat_config>() };
        unsafe { splat_config_default(&mut cfg) };
        cfg.device = device;
        cfg.lowpass = lowpass;
        let mut ctx = ptr::null_mut();
        let rc = unsafe { splat_create(&mut ctx, &cfg) };
        assert!(rc == 0, "splat_create failed: {rc}");
        Self { ctx, uploaded: Cell::new(false) }
    }
    fn check(&self, rc: i32) {
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(splat_last_error(self.ctx)) }.to_string_lossy().into_owned();
            panic!("libsplat_b200: {rc}: {msg}");      // the reference panics on every error as well
        }
    }
    /// GaussianList: nalgebra matrices are column-major, so each `as_ptr()` is already the
    /// contiguous k x N array splat_upload_soa wants (gaussians.rs:408-416).  cov3d is private
    /// (:415) and is recomputed on the device.
    pub fn upload_list(&self, g: &GaussianList) {
        self.check(unsafe { splat_upload_soa(self.ctx, g.positions.as_ptr(), g.scales.as_ptr(),
            g.opacities.as_ptr(), g.rotations.as_ptr(), g.sh.as_ptr(), g.num_gaussians as u64) });
        self.uploaded.set(true);
    }
    /// Vec<Gaussian> is not repr(C): copy field-wise into 59 floats per Gaussian
    /// (position 3, scale 3, opacity 1, rotation i j k w 4, sh 48; gaussians.rs:31-38).
    pub fn upload_vec(&self, gs: &[Gaussian]) {
        let mut buf = Vec::with_capacity(gs.len() * 59);
        for g in gs {
            buf.extend_from_slice(g.position.as_slice());
            buf.extend_from_slice(g.scale.as_slice());
            buf.push(g.opacity);
            buf.extend_from_slice(g.rotation.as_vector().as_slice());   // i, j, k, w
            buf.extend_from_slice(g.sh.as_slice());
        }
        self.check(unsafe { splat_upload_aos(self.ctx, buf.as_ptr(), gs.len() as u64) });
        self.uploaded.set(true);
    }
    pub fn is_uploaded(&self) -> bool { self.uploaded.get() }
    pub fn render(&self, cam: &Camera, color: &mut euc::Buffer<u32, 2>) {
        let [w, h] = color.size();
        let hf = cam.get_htanfovxy_focal();
        let mut c = unsafe { std::mem::zeroed::<splat_camera>() };
        c.view.copy_from_slice(cam.get_view_matrix().as_slice());       // column-major, camera.rs:70
        c.proj.copy_from_slice(cam.get_project_matrix().as_slice());    // camera.rs:80
        c.position.copy_from_slice(cam.position.as_slice());            // the pub field, camera.rs:10
        c.w = cam.w; c.h = cam.h;
        c.htanx = hf[0]; c.htany = hf[1]; c.focal = hf[2];
        self.check(unsafe { splat_render(self.ctx, &c, color.raw_mut().as_mut_ptr(), w as u32, h as u32) });
    }
}
impl Drop for B200 { fn drop(&mut self) { unsafe { splat_destroy(self.ctx) } } }
