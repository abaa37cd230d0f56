//! Replacement body of `Camera::render` (reference src/camera.rs:161-202).  Everything after
//! `self.initilize()` runs behind the C ABI; the progress bar goes away (a render is seconds).
use crate::ffi::*;
use crate::flatten::Flattener;

impl crate::camera::Camera {
    pub fn render(&mut self, world: &dyn crate::hit::Hittable, lights: Option<&dyn crate::hit::Hittable>) -> image::RgbImage {
        self.initilize();
        let mut f = Flattener::default();
        let world_root = world.flatten(&mut f).expect("scene contains a type the GPU core cannot express");
        let lights_root = lights.map(|l| l.flatten(&mut f).expect("unsupported light")).unwrap_or(RT_NONE);
        let background_tex = f.texture(&self.background.texture).expect("unsupported background texture");
        let cam = rt_camera {
            image_width: self.image_width, image_height: self.image_height, sqrt_spp: self.sqrt_spp, max_depth: self.max_depth,
            background_tex, toon_map: match self.toon_map { crate::utils::color::ToonMap::None => 0, _ => 1 },
            recip_sqrt_spp: self.recip_sqrt_spp, pixel_sample_scale: self.pixel_sample_scale,
            center: self.center.e(), pixel00_loc: self.pixel00_loc.e(),
            pixel_delta_u: self.pixel_delta_u.e(), pixel_delta_v: self.pixel_delta_v.e(),
            defocus_angle_in_degrees: self.defocus_angle_in_degrees,
            defocus_disk_u: self.defocus_disk_u.e(), defocus_disk_v: self.defocus_disk_v.e(),
        };
        let desc = f.desc(world_root, lights_root);
        let mut scene: *mut rt_scene = std::ptr::null_mut();
        check(unsafe { rt_scene_create(&desc, std::ptr::null(), &mut scene) });
        let n_px = (self.image_width * self.image_height) as usize;
        let mut accum = vec![0f64; n_px * 3];
        let opts = rt_render_opts { struct_size: std::mem::size_of::<rt_render_opts>() as u32, seed: 0x2025, accum_type: RT_ACCUM_F64, ..Default::default() };
        let mut stats = rt_stats::default();
        let rc = unsafe { rt_render(scene, &cam, &opts, accum.as_mut_ptr().cast(), &mut stats) };
        unsafe { rt_scene_destroy(scene) };
        check(rc);
        assert_eq!(stats.errors, 0, "a sample hit a state on which the reference panics (camera.rs:309,323 / pdf.rs:105-109)");
        let mut img: image::RgbImage = image::ImageBuffer::new(self.image_width, self.image_height);
        check(unsafe { rt_tonemap(accum.as_ptr().cast(), RT_ACCUM_F64, n_px as u64, cam.toon_map, img.as_mut_ptr()) });
        img
    }
}
