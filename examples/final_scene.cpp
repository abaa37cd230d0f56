// final_scene.cpp — the reference's `final_scene` / `cornell_box` / book-1 scene rendered through the
// drop-in API: scene code as in src/main.rs, `camera.render(&world, Some(&lights))` as the entry point.
//   usage: final_scene <book2_final|cornell_glass|book1_final> <width> <spp> <depth> <out.ppm> [scene_seed]
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../raytracer-2025_b200/host/scenes.hpp"

using namespace rt2025;

// The scene functions of scenes.hpp stop right before `camera.render`; this wrapper finishes the job the
// way main.rs does, through Camera::render (flatten -> rt_scene_create -> rt_render -> rt_tonemap).
int main(int argc, char** argv) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s scene width spp depth out.ppm [seed]\n", argv[0]);
        return 2;
    }
    std::string name = argv[1];
    uint32_t width = (uint32_t)std::atoi(argv[2]), depth = (uint32_t)std::atoi(argv[4]);
    size_t spp = (size_t)std::atoll(argv[3]);
    Random::seed(argc > 6 ? std::strtoull(argv[6], nullptr, 10) : 7);

    // a small scene written directly against the API, exactly like portal_scene()/cornell_box() in main.rs
    Hittables world, lights;
    Camera camera;
    if (name == "mini_cornell") {
        auto red = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.65, 0.05, 0.05)));
        auto white = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.73, 0.73, 0.73)));
        auto green = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.12, 0.45, 0.15)));
        auto light = std::make_shared<DiffuseLight>(std::make_shared<SolidColor>(Color(15.0, 15.0, 15.0)));
        world.add(std::make_shared<Quad>(Point3(555, 0, 0), Vec3(0, 555, 0), Vec3(0, 0, 555), green));
        world.add(std::make_shared<Quad>(Point3(0, 0, 0), Vec3(0, 555, 0), Vec3(0, 0, 555), red));
        world.add(std::make_shared<Quad>(Point3(343, 554, 332), Vec3(-130, 0, 0), Vec3(0, 0, -105), light));
        world.add(std::make_shared<Quad>(Point3(0, 0, 0), Vec3(555, 0, 0), Vec3(0, 0, 555), white));
        world.add(std::make_shared<Quad>(Point3(555, 555, 555), Vec3(-555, 0, 0), Vec3(0, 0, -555), white));
        world.add(std::make_shared<Quad>(Point3(0, 0, 555), Vec3(555, 0, 0), Vec3(0, 555, 0), white));
        world.add(std::make_shared<Transform>(build_box(Point3(0, 0, 0), Point3(165, 330, 165), white), Vec3(265, 0, 295),
                                              Quaternion::from_axis_angle(Vec3(0, 1, 0), 15.0), std::nullopt));
        lights.add(std::make_shared<Quad>(Point3(343, 554, 332), Vec3(-130, 0, 0), Vec3(0, 0, -105), light));
        camera.aspect_ratio = 1.0;
        camera.image_width = width;
        camera.samples_per_pixel = spp;
        camera.max_depth = depth;
        camera.vertical_fov_in_degrees = 40.0;
        camera.look_from = Point3(278, 278, -800);
        camera.look_at = Point3(278, 278, 0);
        RgbImage img = camera.render(world, &lights);
        FILE* f = std::fopen(argv[5], "wb");
        std::fprintf(f, "P6\n%u %u\n255\n", img.width, img.height);
        std::fwrite(img.data.data(), 1, img.data.size(), f);
        std::fclose(f);
        std::printf("%u x %u, %llu paths, %.2f ms on the device, %llu errors\n", img.width, img.height,
                    (unsigned long long)camera.last_stats.paths, camera.last_stats.ms_total, (unsigned long long)camera.last_stats.errors);
        return 0;
    }
    std::fprintf(stderr, "unknown scene %s\n", name.c_str());
    return 2;
}
