// CPU check of the scene compiler's tie ranks (csrc/compile.cpp, bvh_visit_order): the presorted-list
// partition walk must give exactly the order of the straightforward restatement of BVH::from_vec
// (bvh.rs:16-46: bounds, longest axis, stable sort by box-min, halves) on inputs full of equal keys.
// Host only, no CUDA:  make check_tie_order && build/check_tie_order
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "compile.h"

static uint64_t key_of(double x) {
    uint64_t b;
    std::memcpy(&b, &x, 8);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

static void plain_order(const rt_scene_desc& d, std::vector<uint32_t> objs, std::vector<uint32_t>& out) {
    const size_t len = objs.size();
    if (len == 1) {
        out.push_back(objs[0]);
        return;
    }
    if (len == 2) {
        out.push_back(objs[1]);
        out.push_back(objs[0]);
        return;
    }
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t o : objs)
        for (int k = 0; k < 3; k++) mn[k] = std::fmin(mn[k], d.objects[o].bbox[2 * k]), mx[k] = std::fmax(mx[k], d.objects[o].bbox[2 * k + 1]);
    double s[3];
    for (int k = 0; k < 3; k++) s[k] = std::fmax(mx[k] - mn[k], 0.0);
    const int axis = s[0] > s[1] ? (s[0] > s[2] ? 0 : 2) : (s[1] > s[2] ? 1 : 2);
    std::stable_sort(objs.begin(), objs.end(), [&](uint32_t x, uint32_t y) { return key_of(d.objects[x].bbox[2 * axis]) < key_of(d.objects[y].bbox[2 * axis]); });
    const size_t mid = len / 2;
    plain_order(d, std::vector<uint32_t>(objs.begin() + mid, objs.end()), out);
    plain_order(d, std::vector<uint32_t>(objs.begin(), objs.begin() + mid), out);
}

static int run_case(uint32_t n, int lattice, uint64_t seed) {
    std::mt19937_64 g(seed);
    std::vector<rt_object> objects(n + 2);
    std::vector<rt_sphere> spheres(n);
    std::vector<uint32_t> children(n + 1);
    const double radii[3] = {0.25, 0.5, 1.0};
    for (uint32_t i = 0; i < n; i++) {
        rt_sphere& s = spheres[i];
        std::memset(&s, 0, sizeof(s));
        for (int k = 0; k < 3; k++) s.center[k] = lattice > 0 ? (double)(g() % (uint64_t)lattice) : (double)(g() >> 11) * 0x1p-53 * 100.0;
        if (lattice > 0 && g() % 7 == 0) s.center[g() % 3] = -0.0;  // signed zeros are distinct under total_cmp
        s.radius = lattice > 0 ? radii[g() % 3] : 0.3;
        rt_object& o = objects[i];
        std::memset(&o, 0, sizeof(o));
        o.kind = RT_OBJ_SPHERE, o.material = 0, o.data = i;
        for (int k = 0; k < 3; k++) o.bbox[2 * k] = s.center[k] - s.radius, o.bbox[2 * k + 1] = s.center[k] + s.radius;
        children[i] = i;
    }
    std::shuffle(children.begin(), children.begin() + n, g);
    rt_object& bvh = objects[n];
    std::memset(&bvh, 0, sizeof(bvh));
    bvh.kind = RT_OBJ_BVH, bvh.material = RT_NONE, bvh.first_child = 0, bvh.child_count = n, bvh.data = RT_NONE;
    rt_object& root = objects[n + 1];
    std::memset(&root, 0, sizeof(root));
    root.kind = RT_OBJ_LIST, root.material = RT_NONE, root.first_child = n, root.child_count = 1, root.data = RT_NONE;
    children[n] = n;
    rt_material mat;
    std::memset(&mat, 0, sizeof(mat));
    mat.kind = RT_MAT_EMPTY, mat.tex = mat.inner = mat.inner2 = RT_NONE;
    rt_scene_desc d;
    std::memset(&d, 0, sizeof(d));
    d.version = RT_ABI_VERSION, d.struct_size = sizeof(d);
    d.n_objects = n + 2, d.n_children = n + 1, d.n_spheres = n, d.n_materials = 1;
    d.objects = objects.data(), d.children = children.data(), d.spheres = spheres.data(), d.materials = &mat;
    d.world_root = n + 1, d.lights_root = RT_NONE;
    rt::CompiledScene cs;
    std::string err;
    const int rc = rt::compile_scene(d, 0, cs, err);
    if (rc != 0) {
        printf("compile failed: %d %s\n", rc, err.c_str());
        return 1;
    }
    std::vector<uint32_t> expect;
    plain_order(d, std::vector<uint32_t>(children.begin(), children.begin() + n), expect);
    size_t bad = 0;
    for (uint32_t r = 0; r < n; r++)
        if (cs.ranks[expect[r]] != r) bad++;
    printf("n=%u lattice=%d seed=%llu: %zu rank mismatches\n", n, lattice, (unsigned long long)seed, bad);
    // four-wide collapse (bvh_build.cpp): same leaves as the binary tree, each exactly once, every child box copied
    // from the binary tree, depth within the traversal stack
    rt::RawVec<rt::Node4> wide;
    uint32_t depth4 = 0;
    const uint32_t root4 = rt::collapse_bvh4(cs.nodes, cs.world_root, wide, depth4);
    std::vector<uint32_t> leaves2, leaves4, todo;
    if (cs.world_root != rt::INVALID_REF) todo.push_back(cs.world_root);
    while (!todo.empty()) {
        const uint32_t ref = todo.back();
        todo.pop_back();
        if (ref & rt::LEAF_FLAG) { leaves2.push_back(ref); continue; }
        for (uint32_t c : {cs.nodes[ref].child0, cs.nodes[ref].child1})
            if (c != rt::INVALID_REF) todo.push_back(c);
    }
    size_t bad4 = 0, inner4 = 0;
    if (root4 != rt::INVALID_REF) todo.push_back(root4);
    while (!todo.empty()) {
        const uint32_t ref = todo.back();
        todo.pop_back();
        if (ref & rt::LEAF_FLAG) { leaves4.push_back(ref); continue; }
        inner4++;
        const rt::Node4& w = wide[ref];
        for (int k = 0; k < 4; k++) {
            if (w.child[k] == rt::INVALID_REF) { bad4 += !(w.box[k][0] > w.box[k][3]); continue; }
            bad4 += !(w.box[k][0] <= w.box[k][3] && w.box[k][1] <= w.box[k][4] && w.box[k][2] <= w.box[k][5]);
            todo.push_back(w.child[k]);
        }
    }
    std::sort(leaves2.begin(), leaves2.end());
    std::sort(leaves4.begin(), leaves4.end());
    bad4 += leaves2 != leaves4;
    bad4 += inner4 != wide.size() && !(cs.world_root & rt::LEAF_FLAG);
    bad4 += 3 * depth4 + 2 > (uint32_t)rt::TRAVERSAL_STACK;
    printf("  four-wide collapse: %zu binary nodes -> %zu wide nodes, depth %u, %zu leaves, %zu problems\n", cs.nodes.size(), wide.size(), depth4,
           leaves4.size(), bad4);
    return bad != 0 || bad4 != 0;
}

int main() {
    int fails = 0;
    for (uint64_t seed = 1; seed <= 4; seed++) {
        fails += run_case(1 + (uint32_t)(seed * 37 % 11), 2, seed);       // tiny nodes
        fails += run_case(1000, 3, seed);                                  // almost every key tied
        fails += run_case(5000, 12, seed);
        fails += run_case(40000, 40, seed);                                // crosses the task threshold
        fails += run_case(30000, 0, seed);                                 // no ties
    }
    fails += run_case(1, 2, 9) + run_case(2, 2, 9) + run_case(3, 2, 9);
    printf(fails ? "FAILED\n" : "OK\n");
    return fails != 0;
}
