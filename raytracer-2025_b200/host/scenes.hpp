// scenes.hpp — the benchmark scenes of BASELINE.json, written against the host mirror the way
// src/main.rs writes them against the crate.  Scene randomness comes from the seeded Random.
#pragma once
#include "rt2025.hpp"

namespace rt2025 {

struct BuiltScene {
    Flattener flat;
    rt_camera camera{};
    rt_scene_desc desc{};
    void finish(const Camera& cam, const Hittable& world, const Hittable* lights) {
        camera = cam.initialize();
        cam.flatten(flat, camera, world, lights);
        desc = flat.desc();
    }
};

// Config 2 — final_scene(image_width, samples_per_pixel, max_depth), main.rs:384-539
inline void final_scene(BuiltScene& out, uint32_t image_width, size_t samples_per_pixel, uint32_t max_depth) {
    Hittables boxes1;
    auto ground_tex = std::make_shared<SolidColor>(Color(0.48, 0.83, 0.53));
    auto ground = std::make_shared<Lambertian>(ground_tex);

    const size_t BOXES_PER_SIDE = 20;
    for (size_t i = 0; i < BOXES_PER_SIDE; i++) {
        for (size_t j = 0; j < BOXES_PER_SIDE; j++) {
            double w = 100.0;
            double x0 = -1000.0 + (double)i * w;
            double z0 = -1000.0 + (double)j * w;
            double y0 = 0.0;
            double x1 = x0 + w;
            double y1 = Random::random_range(1.0, 101.0);
            double z1 = z0 + w;
            boxes1.add(build_box(Point3(x0, y0, z0), Point3(x1, y1, z1), ground));
        }
    }

    auto earth_tex = std::make_shared<ImageTexture>("earthmap.jpg");  // absent in the repo -> cyan
    auto earth_material = std::make_shared<Lambertian>(earth_tex);
    auto earth = std::make_shared<Sphere>(Point3(400.0, 200.0, 400.0), 100.0, earth_material);

    Hittables world;
    world.add(earth);
    world.add(std::make_shared<BVH>(std::move(boxes1)));

    auto light_tex = std::make_shared<SolidColor>(Color(7.0, 7.0, 7.0));
    auto light_material = std::make_shared<DiffuseLight>(light_tex);
    world.add(std::make_shared<Quad>(Point3(123.0, 554.0, 147.0), Vec3(300.0, 0.0, 0.0), Vec3(0.0, 0.0, 265.0),
                                     light_material));

    Point3 center1(400.0, 400.0, 200.0);
    Point3 center2 = center1 + Vec3(30.0, 0.0, 0.0);
    auto sphere_material = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.7, 0.3, 0.1)));
    world.add(Sphere::new_with_motion(center1, center2, 50.0, sphere_material));

    auto glass_material = std::make_shared<Dielectric>(std::make_shared<SolidColor>(Color(1, 1, 1)), 1.5);
    world.add(std::make_shared<Sphere>(Point3(260.0, 150.0, 45.0), 50.0, glass_material));

    auto metal_material = std::make_shared<Metal>(Color(0.8, 0.8, 0.9), 1.0);
    world.add(std::make_shared<Sphere>(Point3(0.0, 150.0, 145.0), 50.0, metal_material));

    world.add(std::make_shared<Sphere>(Point3(360.0, 150.0, 145.0), 70.0, glass_material));
    auto boundary = std::make_shared<Sphere>(Point3(360.0, 150.0, 145.0), 70.0, std::make_shared<EmptyMaterial>());
    world.add(ConstantMedium::new_with_tex(boundary, 0.2, std::make_shared<SolidColor>(Color(0.2, 0.4, 0.9))));
    auto boundary2 = std::make_shared<Sphere>(Point3(0.0, 0.0, 0.0), 5000.0, std::make_shared<EmptyMaterial>());
    world.add(ConstantMedium::new_with_tex(boundary2, 0.0001, std::make_shared<SolidColor>(Color(1, 1, 1))));

    auto pertext = std::make_shared<NoiseTexture>(0.2);
    world.add(std::make_shared<Sphere>(Point3(220.0, 280.0, 300.0), 80.0, std::make_shared<Lambertian>(pertext)));

    Hittables boxes2;
    auto white = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.73, 0.73, 0.73)));
    const size_t NS = 1000;
    for (size_t k = 0; k < NS; k++) boxes2.add(std::make_shared<Sphere>(random_vec3_range(0.0, 165.0), 10.0, white));

    world.add(std::make_shared<Transform>(std::make_shared<BVH>(std::move(boxes2)), Vec3(-100.0, 270.0, 395.0),
                                          Quaternion::from_axis_angle(Vec3(0.0, 1.0, 0.0), 15.0), std::nullopt));

    Hittables lights;
    lights.add(std::make_shared<Quad>(Point3(123.0, 554.0, 147.0), Vec3(300.0, 0.0, 0.0), Vec3(0.0, 0.0, 265.0),
                                      std::make_shared<EmptyMaterial>()));

    Camera camera;
    camera.aspect_ratio = 1.0;
    camera.image_width = image_width;
    camera.samples_per_pixel = samples_per_pixel;
    camera.max_depth = max_depth;
    camera.vertical_fov_in_degrees = 40.0;
    camera.look_from = Point3(478.0, 278.0, -600.0);
    camera.look_at = Point3(278.0, 278.0, 0.0);
    camera.vec_up = Vec3(0.0, 1.0, 0.0);
    camera.defocus_angle_in_degrees = 0.0;
    out.finish(camera, world, &lights);
}

// Config 3 — the Cornell walls of cornell_box() (main.rs:541-590) with the book-3 glass sphere
// that the reference left commented out (main.rs:606-611), lights = quad + sphere.
// `as_shipped` builds exactly what main.rs:541-639 builds (no sphere, 1080 px, 100 spp, depth 10).
inline void cornell_box(BuiltScene& out, bool as_shipped, uint32_t image_width, size_t spp, uint32_t max_depth) {
    Hittables world, lights;
    auto red = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.65, 0.05, 0.05)));
    auto white = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.73, 0.73, 0.73)));
    auto green = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.12, 0.45, 0.15)));
    auto light = std::make_shared<DiffuseLight>(std::make_shared<SolidColor>(Color(15.0, 15.0, 15.0)));

    world.add(std::make_shared<Quad>(Point3(555.0, 0.0, 0.0), Vec3(0.0, 555.0, 0.0), Vec3(0.0, 0.0, 555.0), green));
    world.add(std::make_shared<Quad>(Point3(0.0, 0.0, 0.0), Vec3(0.0, 555.0, 0.0), Vec3(0.0, 0.0, 555.0), red));
    world.add(std::make_shared<Quad>(Point3(343.0, 554.0, 332.0), Vec3(-130.0, 0.0, 0.0), Vec3(0.0, 0.0, -105.0), light));
    world.add(std::make_shared<Quad>(Point3(0.0, 0.0, 0.0), Vec3(555.0, 0.0, 0.0), Vec3(0.0, 0.0, 555.0), white));
    world.add(std::make_shared<Quad>(Point3(555.0, 555.0, 555.0), Vec3(-555.0, 0.0, 0.0), Vec3(0.0, 0.0, -555.0), white));
    world.add(std::make_shared<Quad>(Point3(0.0, 0.0, 555.0), Vec3(555.0, 0.0, 0.0), Vec3(0.0, 555.0, 0.0), white));

    auto box1 = build_box(Point3(0, 0, 0), Point3(165.0, 330.0, 165.0), white);
    world.add(std::make_shared<Transform>(box1, Vec3(265.0, 0.0, 295.0),
                                          Quaternion::from_axis_angle(Vec3(0.0, 1.0, 0.0), 15.0), std::nullopt));

    lights.add(std::make_shared<Quad>(Point3(343.0, 554.0, 332.0), Vec3(-130.0, 0.0, 0.0), Vec3(0.0, 0.0, -105.0), light));
    if (!as_shipped) {
        auto glass = std::make_shared<Dielectric>(std::make_shared<SolidColor>(Color(1, 1, 1)), 1.5);
        world.add(std::make_shared<Sphere>(Point3(190.0, 90.0, 190.0), 90.0, glass));
        lights.add(std::make_shared<Sphere>(Point3(190.0, 90.0, 190.0), 90.0, std::make_shared<EmptyMaterial>()));
    }

    Camera camera;
    camera.aspect_ratio = 1.0;
    camera.image_width = image_width;
    camera.samples_per_pixel = spp;
    camera.max_depth = max_depth;
    camera.vertical_fov_in_degrees = 40.0;
    camera.look_from = Point3(278.0, 278.0, -800.0);
    camera.look_at = Point3(278.0, 278.0, 0.0);
    camera.vec_up = Vec3(0.0, 1.0, 0.0);
    camera.defocus_angle_in_degrees = 0.0;
    out.finish(camera, world, &lights);
}

// Config 1 — "Ray Tracing in One Weekend" final scene.  The reference has no such scene
// function and no sky; it is composed from the crate's API plus GradientTexture.
inline void book1_final(BuiltScene& out, uint32_t image_width, size_t spp, uint32_t max_depth) {
    Hittables spheres;
    auto ground = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.5, 0.5, 0.5)));
    spheres.add(std::make_shared<Sphere>(Point3(0.0, -1000.0, 0.0), 1000.0, ground));
    auto white_tex = std::make_shared<SolidColor>(Color(1, 1, 1));
    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            double choose_mat = Random::f64();
            double cx = (double)a + 0.9 * Random::f64();
            double cz = (double)b + 0.9 * Random::f64();
            Point3 center(cx, 0.2, cz);
            if ((center - Point3(4.0, 0.2, 0.0)).length() > 0.9) {
                if (choose_mat < 0.8) {
                    Color r1(Random::f64(), Random::f64(), Random::f64());
                    Color r2(Random::f64(), Random::f64(), Random::f64());
                    auto m = std::make_shared<Lambertian>(std::make_shared<SolidColor>(r1 * r2));
                    spheres.add(std::make_shared<Sphere>(center, 0.2, m));
                } else if (choose_mat < 0.95) {
                    Color albedo = random_vec3_range(0.5, 1.0);
                    double fuzz = Random::random_range(0.0, 0.5);
                    spheres.add(std::make_shared<Sphere>(center, 0.2, std::make_shared<Metal>(albedo, fuzz)));
                } else {
                    spheres.add(std::make_shared<Sphere>(center, 0.2, std::make_shared<Dielectric>(white_tex, 1.5)));
                }
            }
        }
    }
    spheres.add(std::make_shared<Sphere>(Point3(0.0, 1.0, 0.0), 1.0, std::make_shared<Dielectric>(white_tex, 1.5)));
    spheres.add(std::make_shared<Sphere>(Point3(-4.0, 1.0, 0.0), 1.0,
                                         std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.4, 0.2, 0.1)))));
    spheres.add(std::make_shared<Sphere>(Point3(4.0, 1.0, 0.0), 1.0, std::make_shared<Metal>(Color(0.7, 0.6, 0.5), 0.0)));

    Hittables world;
    world.add(std::make_shared<BVH>(std::move(spheres)));

    Camera camera;
    camera.aspect_ratio = 16.0 / 9.0;
    camera.image_width = image_width;
    camera.samples_per_pixel = spp;
    camera.max_depth = max_depth;
    camera.background.texture = std::make_shared<GradientTexture>(Color(1.0, 1.0, 1.0), Color(0.5, 0.7, 1.0));
    camera.vertical_fov_in_degrees = 20.0;
    camera.look_from = Point3(13.0, 2.0, 3.0);
    camera.look_at = Point3(0.0, 0.0, 0.0);
    camera.vec_up = Vec3(0.0, 1.0, 0.0);
    camera.defocus_angle_in_degrees = 0.6;
    camera.focus_distance = 10.0;
    out.finish(camera, world, nullptr);
}

// Config 5 — synthetic closest-hit soup: N triangles (or spheres) with centres
// U(0,1)^3.  Triangle edge vectors U(-s,s)^3, s = 0.5*N^(-1/3); sphere radius 0.25*N^(-1/3).
// Written straight into the flat arrays (one heap object per primitive would not scale to
// 1e8), with the derived plane constants computed exactly like Triangle::new.  Each primitive
// draws from its own counter-based stream so the fill can run in parallel.
inline uint64_t soup_hash(uint64_t seed, uint64_t i, uint64_t k) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (i * 16 + k + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline double soup_u(uint64_t seed, uint64_t i, uint64_t k) { return (double)(soup_hash(seed, i, k) >> 11) * 0x1.0p-53; }

inline void soup(BuiltScene& out, uint64_t seed, uint64_t n, bool spheres) {
    Flattener& f = out.flat;
    auto mat = std::make_shared<Lambertian>(std::make_shared<SolidColor>(Color(0.73, 0.73, 0.73)));
    uint32_t mat_id = f.material(mat);
    double cbrt_n = std::cbrt((double)n);
    double s = 0.5 / cbrt_n, r = 0.25 / cbrt_n;
    f.objects.resize(n);
    if (spheres)
        f.spheres.resize(n);
    else
        f.planars.resize(n);
    std::vector<uint8_t> ok(n, 1);
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < (int64_t)n; ii++) {
        uint64_t i = (uint64_t)ii;
        Point3 c(soup_u(seed, i, 0), soup_u(seed, i, 1), soup_u(seed, i, 2));
        rt_object o{};
        o.material = mat_id;
        o.data = (uint32_t)i;
        AABB box;
        if (spheres) {
            rt_sphere sp{};
            for (int k = 0; k < 3; k++) sp.center[k] = c[k];
            sp.radius = r;
            f.spheres[i] = sp;
            Vec3 rvec(r, r, r);
            box = AABB::from_points(c - rvec, c + rvec);
            o.kind = RT_OBJ_SPHERE;
        } else {
            Vec3 u(-s + 2.0 * s * soup_u(seed, i, 3), -s + 2.0 * s * soup_u(seed, i, 4), -s + 2.0 * s * soup_u(seed, i, 5));
            Vec3 v(-s + 2.0 * s * soup_u(seed, i, 6), -s + 2.0 * s * soup_u(seed, i, 7), -s + 2.0 * s * soup_u(seed, i, 8));
            Point3 anchor = c - (u + v) / 3.0;  // centroid at c
            Vec3 nrm = u.cross(v);
            auto normal = unit_vector(nrm);
            if (!normal) {  // Triangle::new -> None; keep a tiny valid stand-in so indices stay dense
                u = Vec3(s, 0, 0), v = Vec3(0, s, 0);
                nrm = u.cross(v);
                normal = unit_vector(nrm);
            }
            rt_planar p{};
            Vec3 w = nrm / nrm.length_squared();
            for (int k = 0; k < 3; k++) p.anchor[k] = anchor[k], p.u[k] = u[k], p.v[k] = v[k], p.normal[k] = (*normal)[k], p.w[k] = w[k];
            p.parm_d = normal->dot(anchor);
            p.area = nrm.length() / 2.0;
            f.planars[i] = p;
            box = AABB::from_points(anchor, anchor + u).union_(AABB::from_points(anchor, anchor + v));
            o.kind = RT_OBJ_TRIANGLE;
        }
        o.bbox[0] = box.x.min, o.bbox[1] = box.x.max, o.bbox[2] = box.y.min;
        o.bbox[3] = box.y.max, o.bbox[4] = box.z.min, o.bbox[5] = box.z.max;
        f.objects[i] = o;
    }
    // world = Hittables[ BVH(all primitives) ]
    AABB all = AABB::EMPTY();
    for (uint64_t i = 0; i < n; i++) {
        const double* b = f.objects[i].bbox;
        all = all.union_(AABB{Interval::raw(b[0], b[1]), Interval::raw(b[2], b[3]), Interval::raw(b[4], b[5])});
    }
    std::vector<uint32_t> kids(n);
    for (uint64_t i = 0; i < n; i++) kids[i] = (uint32_t)i;
    uint32_t bvh = f.push_object(RT_OBJ_BVH, RT_NONE, RT_NONE, all, kids);
    f.world_root = f.push_object(RT_OBJ_LIST, RT_NONE, RT_NONE, AABB().union_(all), {bvh});
    f.lights_root = RT_NONE;
    Camera camera;  // pinhole at (0.5,0.5,-2) looking at the cube (SURVEY.md §8d ray set i)
    camera.aspect_ratio = 1.0;
    camera.image_width = 1024;
    camera.samples_per_pixel = 1;
    camera.max_depth = 1;
    camera.vertical_fov_in_degrees = 30.0;
    camera.look_from = Point3(0.5, 0.5, -2.0);
    camera.look_at = Point3(0.5, 0.5, 0.5);
    out.camera = camera.initialize();
    out.camera.background_tex = f.texture(camera.background.texture);
    out.desc = f.desc();
}

}  // namespace rt2025
