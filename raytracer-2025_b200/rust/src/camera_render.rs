//! `Camera::render` on the CUDA core (replaces the pixel loop of reference src/camera.rs:161-202; in the patched crate this file
//! is `src/camera/render_gpu.rs`, a child module of `camera`, so that it can read the private fields `initilize()` fills in).
//! Everything after `self.initilize()` runs behind the C ABI; the progress bar goes away (a render is seconds).
//! With RT2025_GPUS=n (n > 1) the frame is rendered by n GPUs of this process (rt_render_multi_rgb8).
use crate::ffi::*;
use crate::flatten::Flattener;

impl super::Camera {
    pub fn render(&mut self, world: &dyn crate::hit::Hittable, lights: Option<&dyn crate::hit::Hittable>) -> image::RgbImage {
        self.initilize();
        let mut f = Flattener::default();
        let world_root = world.flatten(&mut f).expect("scene contains a type the GPU core cannot express");
        let lights_root = lights.map(|l| l.flatten(&mut f).expect("unsupported light")).unwrap_or(RT_NONE);
        let background_tex = f.texture(&self.background.texture).expect("unsupported background texture");
        let cam = rt_camera {
            image_width: self.image_width, image_height: self.image_height, sqrt_spp: self.sqrt_spp as u32, max_depth: self.max_depth,
            background_tex, toon_map: match self.toon_map { crate::utils::color::ToonMap::None => 0, _ => 1 },
            recip_sqrt_spp: self.recip_sqrt_spp, pixel_sample_scale: self.pixel_sample_scale,
            center: self.center.e(), pixel00_loc: self.pixel00_loc.e(),
            pixel_delta_u: self.pixel_delta_u.e(), pixel_delta_v: self.pixel_delta_v.e(),
            defocus_angle_in_degrees: self.defocus_angle_in_degrees,
            defocus_disk_u: self.defocus_disk_u.e(), defocus_disk_v: self.defocus_disk_v.e(),
        };
        let desc = f.desc(world_root, lights_root);
        let n_gpus = std::env::var("RT2025_GPUS").ok().and_then(|s| s.parse::<usize>().ok()).unwrap_or(1).max(1);
        let mut scenes: Vec<*mut rt_scene> = Vec::new();
        for device in 0..n_gpus {
            let build = rt_build_opts { struct_size: std::mem::size_of::<rt_build_opts>() as u32, flags: 0, device: device as i32, reserved: 0 };
            let mut scene: *mut rt_scene = std::ptr::null_mut();
            check(unsafe { rt_scene_create(&desc, &build, &mut scene) });
            scenes.push(scene);
        }
        // the reference's RNG is unseeded (utils/random.rs:8-14); the core's draws are addressed Philox: any fixed seed is a valid run
        let opts = rt_render_opts { struct_size: std::mem::size_of::<rt_render_opts>() as u32, seed: 0x2025, ..Default::default() };
        let mut stats = rt_stats::default();
        let mut img: image::RgbImage = image::ImageBuffer::new(self.image_width, self.image_height);
        // render + Color::to_rgb on the device: only the RgbImage bytes cross PCIe (3 B per pixel instead of 24 + 24)
        let rc = unsafe {
            if n_gpus == 1 { rt_render_rgb8(scenes[0], &cam, &opts, img.as_mut_ptr(), &mut stats) }
            else { rt_render_multi_rgb8(scenes.as_ptr(), n_gpus as u32, &cam, &opts, img.as_mut_ptr(), &mut stats) }
        };
        for s in scenes { unsafe { rt_scene_destroy(s) }; }
        check(rc);
        assert_eq!(stats.errors, 0, "a sample hit a state on which the reference panics (camera.rs:309,323 / pdf.rs:105-109)");
        img
    }
}
