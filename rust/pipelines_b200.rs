// The two method bodies of src/pipelines.rs that change.
// NOT compiled in this repository: the image has no Rust toolchain.
pub struct GaussianSplatPipeline02 {
    pub gaussians: GaussianList,
    pub camera: Camera,
    #[cfg(feature = "b200")] gpu: std::cell::OnceCell<crate::ffi::B200>,
}

impl GaussianSplatPipeline02 {
    #[cfg(feature = "b200")]
    pub fn render_to_buffer(&self, color: &mut euc::Buffer<u32, 2>) {     // pipelines.rs:260
        let gpu = self.gpu.get_or_init(|| crate::ffi::B200::new(0, 0.3));
        if !gpu.is_uploaded() { gpu.upload_list(&self.gaussians); }        // once per scene
        gpu.render(&self.camera, color);     // sort + project + rasterise + blend, H2D/D2H of `color`
    }
    #[cfg(not(feature = "b200"))]
    pub fn render_to_buffer(&self, color: &mut euc::Buffer<u32, 2>) { /* original body, :260-280 */ }
}
// GaussianSplatPipeline01 (pipelines.rs:54-86): identical with B200::new(0, 0.01) and upload_vec.
