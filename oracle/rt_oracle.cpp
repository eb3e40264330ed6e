// rt_oracle.cpp — CPU ORACLE for the per-pixel render path of ray-tracer-slave.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ray-tracer-s8_b200/, include/) may
// include, link, import or execute this file.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the shared object built from it.
//
// What it is: an operation-for-operation C++ restatement of the reference's Rust hot path
// (all citations relative to /root/reference/):
//   ray-tracer-slave/src/main.rs:32-146           worker() render loop + ray_color
//   ray-tracer-slave/src/camera.rs:19-47,109-129  Camera::new / get_ray
//   ray-tracer-slave/src/color.rs:13-19,29-35     as_slice / blend
//   ray-tracer-slave/src/shapes/mod.rs:12-13,106-129,158-191   t-range, root pick, nearest hit
//   ray-tracer-slave/src/shapes/sphere.rs:42-51,65-72          sphere roots / normal / aabb
//   ray-tracer-slave/src/shapes/mesh.rs:46-95,109-165          triangle aabb / roots / normal
//   ray-tracer-slave/local-dependencies/bvh/src/ray.rs:82-112,133-149,174-194
//   ray-tracer-slave/local-dependencies/bvh/src/aabb.rs:124,268,357,458,480,525,570-580
//   ray-tracer-slave/local-dependencies/bvh/src/bvh/bvh_impl.rs:229-364,373-398,421-442
//   ray-tracer-slave/local-dependencies/bvh/src/utils.rs:8-58
// and of the arithmetic living in un-vendored crates pinned by ray-tracer-slave/Cargo.lock
// (sources NOT under /root/reference; restated from the published crate sources):
//   glam 0.23.0 (Vec3A, SSE2 backend)  roots 0.0.8 (find_roots_quadratic)
//   rand 0.8.5 / rand_core 0.6.4 (SmallRng = xoshiro256++, seed_from_u64 = SplitMix64,
//   UniformFloat)  rand_distr 0.4.3 (UnitDisc, UnitSphere)
//
// Parity status: the reference cannot be compiled or run in this environment (no cargo/rustc,
// no crate sources, prebuilt binaries stripped).  PINNED by the reference's own tests: BVH build +
// traverse + slab test (21-unit-box KAT, bvh/src/testbase.rs:92-166; doc-test ray.rs:160-168) and
// the published xoshiro256++/SplitMix64 known-answer vectors.  PARITY UNPINNED (no reference test
// or fixture exists): camera, sphere/triangle roots, nearest-hit selection, ray_color, as_slice,
// and the glam/roots/rand_distr restatements.  See DESIGN.md.
//
// One required deviation: the reference seeds SmallRng::from_entropy() once per
// image row (main.rs:69) and is therefore not reproducible.  The oracle (and the GPU path) use one
// stream per pixel: SmallRng::seed_from_u64(seed + (y_global*width + x)); a pixel's samples and
// bounces draw from it in the reference's order.
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -std=c++17 -shared -fPIC -pthread
// (rustc never contracts a*b+c into an FMA; division and sqrt are IEEE correctly rounded).

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// glam 0.23.0 Vec3A, SSE2 backend
// ---------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
// dot3: mul_ps, then add_ss(x,y), then add_ss(.,z)
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline float length_recip(V3 a) { return 1.0f / length(a); }
// normalize(): _mm_div_ps(v, sqrt(dot)) — a division per lane
inline V3 normalize(V3 a) {
    float l = length(a);
    return v3(a.x / l, a.y / l, a.z / l);
}
inline bool try_normalize(V3 a, V3* out) {
    float rcp = length_recip(a);
    if (std::isfinite(rcp) && rcp > 0.0f) {
        *out = a * rcp;
        return true;
    }
    return false;
}
inline V3 normalize_or_zero(V3 a) {
    V3 r;
    if (try_normalize(a, &r)) return r;
    return v3(0.0f, 0.0f, 0.0f);
}
// cross: (a.zxy*b - a*b.zxy).zxy
inline V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float axis(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// ---------------------------------------------------------------------------------------------
// rand 0.8.5: SmallRng (64-bit) = xoshiro256++, SplitMix64 seeding
// ---------------------------------------------------------------------------------------------
inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

inline uint64_t splitmix64_next(uint64_t* state) {
    *state += 0x9e3779b97f4a7c15ull;
    uint64_t z = *state;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

struct Counters {
    uint64_t rays = 0;           // nearest-hit queries (ray_color calls with depth > 0)
    uint64_t primary = 0;        // get_ray calls
    uint64_t aabb_tests = 0;     // Ray::intersects_aabb calls
    uint64_t sphere_tests = 0;   // Sphere::get_roots calls
    uint64_t sphere_hits = 0;    // ... that produced an in-range point
    uint64_t tri_tests = 0;      // Triangle::get_roots calls
    uint64_t tri_exit[4] = {0, 0, 0, 0};  // exit stage: det / u / v / reached dist
    uint64_t tri_hits = 0;
    uint64_t shades_sphere = 0;  // non-terminal hits
    uint64_t shades_tri = 0;
    uint64_t emissive = 0;
    uint64_t sky = 0;
    uint64_t depth_exhausted = 0;
    uint64_t rng_draws = 0;      // next_u32 calls
    void add(const Counters& o) {
        rays += o.rays; primary += o.primary; aabb_tests += o.aabb_tests;
        sphere_tests += o.sphere_tests; sphere_hits += o.sphere_hits; tri_tests += o.tri_tests;
        for (int i = 0; i < 4; i++) tri_exit[i] += o.tri_exit[i];
        tri_hits += o.tri_hits; shades_sphere += o.shades_sphere; shades_tri += o.shades_tri;
        emissive += o.emissive; sky += o.sky; depth_exhausted += o.depth_exhausted;
        rng_draws += o.rng_draws;
    }
};

struct SmallRng {
    uint64_t s[4];
    Counters* ctr = nullptr;
    static SmallRng seed_from_u64(uint64_t state) {
        SmallRng r;
        for (int i = 0; i < 4; i++) r.s[i] = splitmix64_next(&state);
        // from_seed: an all-zero seed is replaced by seed_from_u64(0) (cannot occur from SplitMix64
        // in practice, kept for fidelity)
        if ((r.s[0] | r.s[1] | r.s[2] | r.s[3]) == 0) return seed_from_u64(0);
        return r;
    }
    uint64_t next_u64() {
        uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl64(s[3], 45);
        return result;
    }
    uint32_t next_u32() {
        if (ctr) ctr->rng_draws++;
        return (uint32_t)(next_u64() >> 32);
    }
    // UniformFloat<f32>: 23 random mantissa bits into [1,2), minus 1
    float value0_1() {
        uint32_t bits = next_u32() >> 9;
        uint32_t u = 0x3f800000u | bits;
        float f;
        std::memcpy(&f, &u, 4);
        return f - 1.0f;
    }
    // Rng::gen_range(0f32..1f32) → UniformFloat::sample_single: value0_1 * scale + low
    float gen_range_0_1() { return value0_1() * 1.0f + 0.0f; }
    // Uniform::new(-1f32, 1f32).sample: scale stays 2.0 (2*(1-2^-23) - 1 < 1)
    float uniform_m1_1() { return value0_1() * 2.0f + (-1.0f); }
};

// rand_distr 0.4.3 UnitDisc
inline void unit_disc(SmallRng& rng, float* a, float* b) {
    float x1, x2;
    for (;;) {
        x1 = rng.uniform_m1_1();
        x2 = rng.uniform_m1_1();
        if (x1 * x1 + x2 * x2 <= 1.0f) break;
    }
    *a = x1;
    *b = x2;
}
// rand_distr 0.4.3 UnitSphere (Marsaglia 1972)
inline V3 unit_sphere(SmallRng& rng) {
    for (;;) {
        float x1 = rng.uniform_m1_1();
        float x2 = rng.uniform_m1_1();
        float sum = x1 * x1 + x2 * x2;
        if (sum >= 1.0f) continue;
        float factor = 2.0f * std::sqrt(1.0f - sum);
        return v3(x1 * factor, x2 * factor, 1.0f - 2.0f * sum);
    }
}

// ---------------------------------------------------------------------------------------------
// bvh crate: Ray (ray.rs:133-149), custom min/max (ray.rs:82-112), AABB (aabb.rs)
// ---------------------------------------------------------------------------------------------
struct Ray {
    V3 origin, direction, inv_direction;
    int sign_x, sign_y, sign_z;
};
inline Ray ray_new(V3 origin, V3 direction) {
    Ray r;
    V3 d = normalize(direction);
    r.origin = origin;
    r.direction = d;
    r.inv_direction = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    r.sign_x = d.x < 0.0f;
    r.sign_y = d.y < 0.0f;
    r.sign_z = d.z < 0.0f;
    return r;
}
inline V3 ray_at(const Ray& r, float t) { return r.origin + t * r.direction; }
inline float rmin(float x, float y) { return x < y ? x : y; }
inline float rmax(float x, float y) { return x > y ? x : y; }

struct AABB {
    V3 min, max;
};
inline AABB aabb_empty() {
    const float inf = std::numeric_limits<float>::infinity();
    return AABB{v3(inf, inf, inf), v3(-inf, -inf, -inf)};
}
// f32::min / f32::max (IEEE minNum/maxNum) = fminf/fmaxf
inline AABB aabb_join(const AABB& a, const AABB& b) {
    return AABB{v3(fminf(a.min.x, b.min.x), fminf(a.min.y, b.min.y), fminf(a.min.z, b.min.z)),
                v3(fmaxf(a.max.x, b.max.x), fmaxf(a.max.y, b.max.y), fmaxf(a.max.z, b.max.z))};
}
inline AABB aabb_grow(const AABB& a, V3 p) {
    return AABB{v3(fminf(a.min.x, p.x), fminf(a.min.y, p.y), fminf(a.min.z, p.z)),
                v3(fmaxf(a.max.x, p.x), fmaxf(a.max.y, p.y), fmaxf(a.max.z, p.z))};
}
inline V3 aabb_size(const AABB& a) { return a.max - a.min; }
inline V3 aabb_center(const AABB& a) { return a.min + (aabb_size(a) / 2.0f); }
inline bool aabb_is_empty(const AABB& a) {
    return a.min.x > a.max.x || a.min.y > a.max.y || a.min.z > a.max.z;
}
inline float aabb_surface_area(const AABB& a) {
    V3 s = aabb_size(a);
    return 2.0f * (s.x * s.y + s.x * s.z + s.y * s.z);
}
inline int aabb_largest_axis(const AABB& a) {
    V3 s = aabb_size(a);
    if (s.x > s.y && s.x > s.z) return 0;
    if (s.y > s.z) return 1;
    return 2;
}
// Ray::intersects_aabb (ray.rs:174-194)
inline bool intersects_aabb(const Ray& r, const AABB& b) {
    const V3* bounds[2] = {&b.min, &b.max};
    float ray_min = (bounds[r.sign_x]->x - r.origin.x) * r.inv_direction.x;
    float ray_max = (bounds[1 - r.sign_x]->x - r.origin.x) * r.inv_direction.x;
    float y_min = (bounds[r.sign_y]->y - r.origin.y) * r.inv_direction.y;
    float y_max = (bounds[1 - r.sign_y]->y - r.origin.y) * r.inv_direction.y;
    ray_min = rmax(ray_min, y_min);
    ray_max = rmin(ray_max, y_max);
    float z_min = (bounds[r.sign_z]->z - r.origin.z) * r.inv_direction.z;
    float z_max = (bounds[1 - r.sign_z]->z - r.origin.z) * r.inv_direction.z;
    ray_min = rmax(ray_min, z_min);
    ray_max = rmin(ray_max, z_max);
    return rmax(ray_min, 0.0f) <= ray_max;
}

// ---------------------------------------------------------------------------------------------
// BVH build (bvh_impl.rs:229-364, utils.rs:19-58) and recursive traverse (bvh_impl.rs:373-398)
// ---------------------------------------------------------------------------------------------
struct BVHNode {
    bool leaf;
    uint32_t parent, depth;
    uint32_t shape;      // leaf
    uint32_t child_l, child_r;
    AABB aabb_l, aabb_r;  // inner
};
struct Bucket {
    size_t size;
    AABB aabb;
};

// float → usize `as` cast: truncates, saturates, NaN → 0
inline size_t f32_as_usize(float f) {
    if (!(f == f)) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 18446744073709551616.0f) return ~(size_t)0;
    return (size_t)f;
}

struct BuildError {};

size_t bvh_build_rec(const std::vector<AABB>& shape_aabb, const std::vector<size_t>& indices,
                     std::vector<BVHNode>& nodes, std::vector<uint32_t>& shape_node,
                     uint32_t parent, uint32_t depth) {
    if (indices.empty() || depth > 100000) throw BuildError{};  // reference recurses forever
    AABB aabb_bounds = aabb_empty(), centroid_bounds = aabb_empty();
    for (size_t idx : indices) {
        V3 c = aabb_center(shape_aabb[idx]);
        aabb_bounds = aabb_join(aabb_bounds, shape_aabb[idx]);
        centroid_bounds = aabb_grow(centroid_bounds, c);
    }
    if (indices.size() == 1) {
        size_t node_index = nodes.size();
        BVHNode n{};
        n.leaf = true;
        n.parent = parent;
        n.depth = depth;
        n.shape = (uint32_t)indices[0];
        nodes.push_back(n);
        shape_node[indices[0]] = (uint32_t)node_index;
        return node_index;
    }
    size_t node_index = nodes.size();
    nodes.push_back(BVHNode{});  // dummy, replaced below
    int split_axis = aabb_largest_axis(centroid_bounds);
    float split_axis_size = axis(centroid_bounds.max, split_axis) - axis(centroid_bounds.min, split_axis);

    size_t child_l, child_r;
    AABB aabb_l, aabb_r;
    if (split_axis_size < 0.00001f) {  // bvh::EPSILON, lib.rs:80
        size_t half = indices.size() / 2;
        std::vector<size_t> li(indices.begin(), indices.begin() + half);
        std::vector<size_t> ri(indices.begin() + half, indices.end());
        aabb_l = aabb_empty();
        for (size_t i : li) aabb_l = aabb_join(aabb_l, shape_aabb[i]);
        aabb_r = aabb_empty();
        for (size_t i : ri) aabb_r = aabb_join(aabb_r, shape_aabb[i]);
        child_l = bvh_build_rec(shape_aabb, li, nodes, shape_node, (uint32_t)node_index, depth + 1);
        child_r = bvh_build_rec(shape_aabb, ri, nodes, shape_node, (uint32_t)node_index, depth + 1);
    } else {
        const int NUM_BUCKETS = 6;
        Bucket buckets[NUM_BUCKETS];
        for (auto& b : buckets) b = Bucket{0, aabb_empty()};
        std::vector<size_t> assign[NUM_BUCKETS];
        for (size_t idx : indices) {
            const AABB& sa = shape_aabb[idx];
            V3 c = aabb_center(sa);
            float rel = (axis(c, split_axis) - axis(centroid_bounds.min, split_axis)) / split_axis_size;
            size_t bn = f32_as_usize(rel * ((float)NUM_BUCKETS - 0.01f));
            if (bn >= (size_t)NUM_BUCKETS) throw BuildError{};  // reference: index panic
            buckets[bn].size += 1;
            buckets[bn].aabb = aabb_join(buckets[bn].aabb, sa);
            assign[bn].push_back(idx);
        }
        size_t min_bucket = 0;
        float min_cost = std::numeric_limits<float>::infinity();
        aabb_l = aabb_empty();
        aabb_r = aabb_empty();
        for (int i = 0; i < NUM_BUCKETS - 1; i++) {
            Bucket l{0, aabb_empty()}, r{0, aabb_empty()};
            for (int k = 0; k <= i; k++) l = Bucket{l.size + buckets[k].size, aabb_join(l.aabb, buckets[k].aabb)};
            for (int k = i + 1; k < NUM_BUCKETS; k++) r = Bucket{r.size + buckets[k].size, aabb_join(r.aabb, buckets[k].aabb)};
            float cost = ((float)l.size * aabb_surface_area(l.aabb) + (float)r.size * aabb_surface_area(r.aabb)) /
                         aabb_surface_area(aabb_bounds);
            if (cost < min_cost) {
                min_bucket = i;
                min_cost = cost;
                aabb_l = l.aabb;
                aabb_r = r.aabb;
            }
        }
        std::vector<size_t> li, ri;
        for (size_t k = 0; k <= min_bucket; k++) li.insert(li.end(), assign[k].begin(), assign[k].end());
        for (size_t k = min_bucket + 1; k < (size_t)NUM_BUCKETS; k++) ri.insert(ri.end(), assign[k].begin(), assign[k].end());
        child_l = bvh_build_rec(shape_aabb, li, nodes, shape_node, (uint32_t)node_index, depth + 1);
        child_r = bvh_build_rec(shape_aabb, ri, nodes, shape_node, (uint32_t)node_index, depth + 1);
    }
    if (aabb_is_empty(aabb_l) || aabb_is_empty(aabb_r)) throw BuildError{};  // reference: assert!
    BVHNode n{};
    n.leaf = false;
    n.parent = parent;
    n.depth = depth;
    n.child_l = (uint32_t)child_l;
    n.child_r = (uint32_t)child_r;
    n.aabb_l = aabb_l;
    n.aabb_r = aabb_r;
    nodes[node_index] = n;
    return node_index;
}

struct BVH {
    std::vector<BVHNode> nodes;
    std::vector<uint32_t> shape_node;
};
bool bvh_build(const std::vector<AABB>& shape_aabb, BVH* out) {
    out->nodes.clear();
    out->nodes.reserve(shape_aabb.size() * 2);
    out->shape_node.assign(shape_aabb.size(), 0);
    std::vector<size_t> indices(shape_aabb.size());
    for (size_t i = 0; i < indices.size(); i++) indices[i] = i;
    try {
        bvh_build_rec(shape_aabb, indices, out->nodes, out->shape_node, 0, 0);
    } catch (const BuildError&) {
        return false;
    }
    return true;
}
void traverse_recursive(const std::vector<BVHNode>& nodes, size_t node_index, const Ray& ray,
                        std::vector<uint32_t>& indices, Counters* ctr) {
    const BVHNode& n = nodes[node_index];
    if (!n.leaf) {
        if (ctr) ctr->aabb_tests++;
        if (intersects_aabb(ray, n.aabb_l)) traverse_recursive(nodes, n.child_l, ray, indices, ctr);
        if (ctr) ctr->aabb_tests++;
        if (intersects_aabb(ray, n.aabb_r)) traverse_recursive(nodes, n.child_r, ray, indices, ctr);
    } else {
        indices.push_back(n.shape);
    }
}

// ---------------------------------------------------------------------------------------------
// roots 0.0.8 find_roots_quadratic
// ---------------------------------------------------------------------------------------------
struct Roots {
    int n;
    float r[2];
};
inline Roots find_roots_quadratic(float a2, float a1, float a0) {
    Roots out{0, {0.0f, 0.0f}};
    if (a2 == 0.0f) {  // linear (never taken: a2 = 1)
        if (a1 == 0.0f) {
            if (a0 == 0.0f) { out.n = 1; out.r[0] = 0.0f; }
            return out;
        }
        out.n = 1;
        out.r[0] = -a0 / a1;
        return out;
    }
    float discriminant = a1 * a1 - 4.0f * a2 * a0;
    if (discriminant < 0.0f) return out;
    float a2x2 = 2.0f * a2;
    if (discriminant == 0.0f) {
        out.n = 1;
        out.r[0] = -a1 / a2x2;
        return out;
    }
    float sq = std::sqrt(discriminant);
    float same_sign, diff_sign;
    if (a1 < 0.0f) {
        same_sign = -a1 + sq;
        diff_sign = -a1 - sq;
    } else {
        same_sign = -a1 - sq;
        diff_sign = -a1 + sq;
    }
    float x1, x2;
    if (std::fabs(same_sign) > std::fabs(a2x2)) {
        float a0x2 = 2.0f * a0;
        if (std::fabs(diff_sign) > std::fabs(a2x2)) {
            x1 = a0x2 / same_sign;
            x2 = a0x2 / diff_sign;
        } else {
            x1 = a0x2 / same_sign;
            x2 = same_sign / a2x2;
        }
    } else {
        x1 = diff_sign / a2x2;
        x2 = same_sign / a2x2;
    }
    out.n = 2;
    if (x1 < x2) {
        out.r[0] = x1;
        out.r[1] = x2;
    } else {
        out.r[0] = x2;
        out.r[1] = x1;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Scene primitives (shapes/*.rs)
// ---------------------------------------------------------------------------------------------
struct Color {
    float r, g, b;
};
struct Object {
    int kind;  // 0 sphere, 1 triangle
    float radius;
    V3 center;
    V3 a, b, c;
    Color albedo;
    float roughness, emission;
};

inline AABB object_aabb(const Object& o) {
    if (o.kind == 0) {  // sphere.rs:65-72
        V3 h = v3(o.radius, o.radius, o.radius);
        return AABB{o.center - h, o.center + h};
    }
    // mesh.rs:46-95: min_by(min_by(a,c),b) with partial_cmp().unwrap_or(Equal); min_by returns the
    // first argument unless compare(first, second) == Greater; max_by returns the second unless
    // compare(first, second) == Greater.
    auto mn = [](float x, float y) { return (x > y) ? y : x; };
    auto mx = [](float x, float y) { return (x > y) ? x : y; };
    V3 lo = v3(mn(mn(o.a.x, o.c.x), o.b.x), mn(mn(o.a.y, o.c.y), o.b.y), mn(mn(o.a.z, o.c.z), o.b.z));
    V3 hi = v3(mx(mx(o.a.x, o.c.x), o.b.x), mx(mx(o.a.y, o.c.y), o.b.y), mx(mx(o.a.z, o.c.z), o.b.z));
    return AABB{lo, hi};
}

inline Roots sphere_roots(const Object& s, const Ray& ray) {  // sphere.rs:42-47
    float a = 1.0f;
    float b = dot(2.0f * ray.direction, ray.origin - s.center);
    float l = length(ray.origin - s.center);
    float c = l * l - s.radius * s.radius;  // powi(2) lowers to x*x
    return find_roots_quadratic(a, b, c);
}
inline Roots triangle_roots(const Object& t, const Ray& ray, int* exit_stage) {  // mesh.rs:109-161
    const float EPSILON = 0.00001f;
    Roots no{0, {0.0f, 0.0f}};
    V3 a_to_b = t.b - t.a;
    V3 a_to_c = t.c - t.a;
    V3 u_vec = cross(ray.direction, a_to_c);
    float det = dot(a_to_b, u_vec);
    if (det < EPSILON && det > -EPSILON) { *exit_stage = 0; return no; }
    float inv_det = 1.0f / det;
    V3 a_to_origin = ray.origin - t.a;
    float u = dot(a_to_origin, u_vec) * inv_det;
    if (!(u >= 0.0f && u <= 1.0f)) { *exit_stage = 1; return no; }
    V3 v_vec = cross(a_to_origin, a_to_b);
    float v = dot(ray.direction, v_vec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) { *exit_stage = 2; return no; }
    float dist = dot(a_to_c, v_vec) * inv_det;
    *exit_stage = 3;
    if (dist > EPSILON) return Roots{1, {dist, 0.0f}};
    return no;
}
inline V3 object_normal(const Object& o, V3 point) {
    if (o.kind == 0) return normalize_or_zero(point - o.center);       // sphere.rs:49-51
    return normalize_or_zero(cross(o.a - o.b, o.a - o.c));             // mesh.rs:163-165
}

const float T_MIN = 0.001f, T_MAX = 1000.0f;  // shapes/mod.rs:12-13
inline bool in_range(float x) { return x >= T_MIN && x < T_MAX; }  // (T_MIN..T_MAX).contains

// Intersectable::get_intersection_point (shapes/mod.rs:106-129)
inline bool intersection_point(const Object& o, const Ray& ray, V3* point, Counters* ctr) {
    Roots rt;
    if (o.kind == 0) {
        if (ctr) ctr->sphere_tests++;
        rt = sphere_roots(o, ray);
    } else {
        int stage = 0;
        rt = triangle_roots(o, ray, &stage);
        if (ctr) { ctr->tri_tests++; ctr->tri_exit[stage]++; }
    }
    float t;
    if (rt.n == 0) return false;
    if (rt.n == 1) {
        if (!in_range(rt.r[0])) return false;
        t = rt.r[0];
    } else {
        bool xi = in_range(rt.r[0]), yi = in_range(rt.r[1]);
        if (xi && yi) t = rt.r[0] < rt.r[1] ? rt.r[0] : rt.r[1];
        else if (xi) t = rt.r[0];
        else if (yi) t = rt.r[1];
        else return false;
    }
    *point = ray_at(ray, t);
    if (ctr) { if (o.kind == 0) ctr->sphere_hits++; else ctr->tri_hits++; }
    return true;
}

struct IntersectionTable {
    V3 point, normal;
    Color albedo;
    float roughness, emission;
    int kind;
    uint32_t index;
};

// WorldRefList::intersect (shapes/mod.rs:158-191) over `cands` in the given order.
// Iterator::min_by keeps the incumbent unless compare(incumbent, new) == Greater;
// partial_cmp().unwrap_or(Less) makes NaN compare as Less (incumbent kept).
inline bool intersect_list(const std::vector<Object>& world, const uint32_t* cands, size_t n_cands,
                           const Ray& ray, IntersectionTable* out, Counters* ctr) {
    bool have = false;
    V3 best_p = v3(0, 0, 0);
    uint32_t best_i = 0;
    for (size_t k = 0; k < n_cands; k++) {
        uint32_t i = cands ? cands[k] : (uint32_t)k;
        V3 p;
        if (!intersection_point(world[i], ray, &p, ctr)) continue;
        if (!have) {
            have = true;
            best_p = p;
            best_i = i;
        } else {
            float la = length(best_p - ray.origin);
            float lb = length(p - ray.origin);
            if (la > lb) {  // Ordering::Greater → take the new one
                best_p = p;
                best_i = i;
            }
        }
    }
    if (!have) return false;
    const Object& o = world[best_i];
    out->emission = o.emission;
    out->point = best_p;
    out->normal = object_normal(o, best_p);
    out->albedo = o.albedo;
    out->roughness = o.roughness;
    out->kind = o.kind;
    out->index = best_i;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Camera (camera.rs:19-47, 109-129)
// ---------------------------------------------------------------------------------------------
struct Camera {
    V3 origin, lower_left_corner, horizontal, vertical;
    float aspect_ratio, image_height, aperture, focal_length, field_of_view, focus_distance;
};
inline Camera camera_new(V3 origin, float aspect_ratio, float aperture, float focus_distance,
                         float field_of_view, float focal_length, float image_height) {
    Camera c;
    float vh = 2.0f * std::tan(field_of_view / 2.0f);
    float vw = aspect_ratio * vh;
    c.horizontal = v3(vw, 0.0f, 0.0f);
    c.vertical = v3(0.0f, vh, 0.0f);
    c.origin = origin;
    c.focus_distance = focus_distance;
    c.image_height = image_height;
    c.field_of_view = field_of_view;
    c.focal_length = focal_length;
    c.aspect_ratio = aspect_ratio;
    c.aperture = aperture;
    c.lower_left_corner = origin - c.horizontal / 2.0f - c.vertical / 2.0f - v3(0.0f, 0.0f, focal_length);
    return c;
}
inline Ray camera_get_ray(const Camera& c, uint32_t x, uint32_t y, SmallRng& rng) {
    float lens_radius = c.aperture / 2.0f;
    float a, b;
    unit_disc(rng, &a, &b);
    V3 offset = v3(a * lens_radius, b * lens_radius, 0.0f);
    float u = ((float)x + rng.gen_range_0_1()) / (c.aspect_ratio * c.image_height - 1.0f);
    float v = ((float)y + rng.gen_range_0_1()) / (c.image_height - 1.0f);
    Ray fr = ray_new(c.origin,
                     normalize_or_zero(c.lower_left_corner + u * c.horizontal + v * c.vertical - c.origin));
    V3 focal_point = ray_at(fr, c.focus_distance);
    V3 final_origin = c.origin + offset;
    return ray_new(final_origin, normalize_or_zero(focal_point - final_origin));
}

// ---------------------------------------------------------------------------------------------
// ray_color (main.rs:108-146) — recursive, like the reference
// ---------------------------------------------------------------------------------------------
struct Scene {
    std::vector<Object> world;
    BVH bvh;
    int mode;  // 0 = reference BVH candidates, 1 = brute force over the whole world in order
};

// optional per-thread ray log (orc_trace_pixel): origin(3), direction(3), hit world position or -1
thread_local std::vector<float>* g_ray_log = nullptr;

Color ray_color(const Ray& ray, const Scene& sc, uint32_t depth, SmallRng& rng, Counters* ctr,
                std::vector<uint32_t>& scratch) {
    if (depth == 0) {
        if (ctr) ctr->depth_exhausted++;
        return Color{0.0f, 0.0f, 0.0f};
    }
    if (ctr) ctr->rays++;
    IntersectionTable tb;
    bool hit;
    if (sc.mode == 0) {
        scratch.clear();
        traverse_recursive(sc.bvh.nodes, 0, ray, scratch, ctr);
        hit = intersect_list(sc.world, scratch.data(), scratch.size(), ray, &tb, ctr);
    } else {
        hit = intersect_list(sc.world, nullptr, sc.world.size(), ray, &tb, ctr);
    }
    if (g_ray_log) {
        g_ray_log->insert(g_ray_log->end(), {ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x,
                                             ray.direction.y, ray.direction.z, hit ? (float)tb.index : -1.0f});
    }
    if (hit) {
        if (tb.emission > 0.0f) {
            if (ctr) ctr->emissive++;
            // f32 * Color → rhs * self → Color{r*e, g*e, b*e}
            return Color{tb.albedo.r * tb.emission, tb.albedo.g * tb.emission, tb.albedo.b * tb.emission};
        }
        if (ctr) { if (tb.kind == 0) ctr->shades_sphere++; else ctr->shades_tri++; }
        V3 diffuse_dir = unit_sphere(rng) + tb.normal;
        V3 glossy_dir = ray.direction - 2.0f * dot(ray.direction, tb.normal) * tb.normal;
        V3 scatter = diffuse_dir + tb.roughness * (glossy_dir - diffuse_dir);
        V3 nd;
        if (!try_normalize(scatter, &nd)) nd = tb.normal;
        Ray child = ray_new(tb.point, nd);
        Color c = ray_color(child, sc, depth - 1, rng, ctr, scratch);
        return Color{tb.albedo.r * c.r, tb.albedo.g * c.g, tb.albedo.b * c.b};
    }
    if (ctr) ctr->sky++;
    float t = normalize_or_zero(ray.direction).y * 0.5f + 1.0f;
    Color w = Color{1.0f * t, 1.0f * t, 1.0f * t};
    float k = 1.0f - t;
    Color s = Color{0.3f * k, 0.3f * k, 0.8f * k};
    return Color{w.r + s.r, w.g + s.g, w.b + s.b};
}

// `as u8`: truncate toward zero, saturate, NaN → 0 (color.rs:13-19)
inline uint8_t f32_as_u8(float f) {
    if (!(f == f)) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)f;
}

}  // namespace

// =============================================================================================
// C ABI (ctypes).  Struct layouts mirror include/rt_b200.h so the same numpy arrays feed both.
// =============================================================================================
extern "C" {

struct orc_sphere {
    float center[3];
    float radius;
    float albedo[3];
    float roughness;
    float emission;
};
struct orc_triangle {
    float a[3], b[3], c[3];
    float albedo[3];
    float roughness;
    float emission;
};
struct orc_params {
    uint32_t width, height, divisions, division_no;
    uint32_t spp, max_bounces;
    uint64_t seed;
    float cam_origin[3];
    float aperture, focus_distance, field_of_view, focal_length;
};
struct orc_stats {
    uint64_t rays, primary, aabb_tests, sphere_tests, sphere_hits, tri_tests, tri_exit[4], tri_hits,
        shades_sphere, shades_tri, emissive, sky, depth_exhausted, rng_draws;
    double build_ms, render_ms;
};

static bool make_world(const orc_sphere* sph, uint32_t ns, const orc_triangle* tri, uint32_t nt,
                       const uint32_t* world_index, std::vector<Object>* world) {
    size_t n = (size_t)ns + nt;
    world->assign(n, Object{});
    std::vector<char> seen(n, 0);
    for (size_t i = 0; i < n; i++) {
        size_t pos = world_index ? world_index[i] : i;
        if (pos >= n || seen[pos]) return false;
        seen[pos] = 1;
        Object o{};
        if (i < ns) {
            const orc_sphere& s = sph[i];
            o.kind = 0;
            o.radius = s.radius;
            o.center = v3(s.center[0], s.center[1], s.center[2]);
            o.albedo = Color{s.albedo[0], s.albedo[1], s.albedo[2]};
            o.roughness = s.roughness;
            o.emission = s.emission;
        } else {
            const orc_triangle& t = tri[i - ns];
            o.kind = 1;
            o.a = v3(t.a[0], t.a[1], t.a[2]);
            o.b = v3(t.b[0], t.b[1], t.b[2]);
            o.c = v3(t.c[0], t.c[1], t.c[2]);
            o.albedo = Color{t.albedo[0], t.albedo[1], t.albedo[2]};
            o.roughness = t.roughness;
            o.emission = t.emission;
        }
        (*world)[pos] = o;
    }
    return true;
}

static void fill_defaults(orc_params* p) {
    if (p->divisions == 0) p->divisions = 1;
    if (p->spp == 0) p->spp = 100;                      // main.rs:51
    if (p->max_bounces == 0) p->max_bounces = 10;       // main.rs:39
    if (p->aperture == 0.0f) p->aperture = 0.1f;        // main.rs:45
    if (p->focus_distance == 0.0f) p->focus_distance = 1.0f;
    if (p->field_of_view == 0.0f) p->field_of_view = 3.14159265358979323846f / 2.0f;  // PI / 2f32
    if (p->focal_length == 0.0f) p->focal_length = 1.0f;
}

static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// Renders rows [row0,row1) of the division band (row indices relative to the band; pass 0 and
// height/divisions for the whole band).  out has (row1-row0)*width*3 bytes.
// mode: 0 = reference BVH (bvh.traverse + intersect), 1 = brute force over the world in order.
// Returns 0, or -1 bad args, -2 BVH build failed (the reference would panic / overflow the stack).
int orc_render_rows(const orc_sphere* sph, uint32_t ns, const orc_triangle* tri, uint32_t nt,
                    const uint32_t* world_index, const orc_params* params_in, uint32_t row0, uint32_t row1,
                    uint8_t* out, size_t out_len, int mode, int threads, orc_stats* stats) {
    if (!params_in || !out) return -1;
    orc_params p = *params_in;
    fill_defaults(&p);
    if (p.width == 0 || p.height == 0 || (size_t)ns + nt == 0) return -1;
    if (p.height % p.divisions != 0 || p.division_no >= p.divisions) return -1;
    uint32_t band_h = p.height / p.divisions;
    if (row0 > row1 || row1 > band_h) return -1;
    if (out_len != (size_t)(row1 - row0) * p.width * 3) return -1;

    Scene sc;
    sc.mode = mode;
    if (!make_world(sph, ns, tri, nt, world_index, &sc.world)) return -1;
    double t0 = now_ms();
    if (mode == 0) {
        std::vector<AABB> boxes(sc.world.size());
        for (size_t i = 0; i < boxes.size(); i++) boxes[i] = object_aabb(sc.world[i]);
        if (!bvh_build(boxes, &sc.bvh)) return -2;
    }
    double t1 = now_ms();

    Camera cam = camera_new(v3(p.cam_origin[0], p.cam_origin[1], p.cam_origin[2]),
                            (float)p.width / (float)p.height, p.aperture, p.focus_distance,
                            p.field_of_view, p.focal_length, (float)p.height);

    int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    uint32_t nrows = row1 - row0;
    if ((uint32_t)nthreads > nrows && nrows > 0) nthreads = (int)nrows;
    std::atomic<uint32_t> next_row{0};
    std::vector<Counters> ctrs(nthreads);
    auto work = [&](int tid) {
        Counters* ctr = stats ? &ctrs[tid] : nullptr;
        std::vector<uint32_t> scratch;
        for (;;) {
            uint32_t r = next_row.fetch_add(1);
            if (r >= nrows) break;
            uint32_t y_band = row0 + r;
            // main.rs:66-68: y = (image_height / divisions) * division_no + y
            uint32_t y_global = band_h * p.division_no + y_band;
            uint8_t* row = out + (size_t)r * p.width * 3;
            for (uint32_t x = 0; x < p.width; x++) {
                // deviation: one xoshiro256++ stream per pixel
                SmallRng rng = SmallRng::seed_from_u64(p.seed + ((uint64_t)y_global * p.width + x));
                rng.ctr = ctr;
                uint32_t y_cam = p.height - y_global - 1;  // main.rs:71
                Color pix{0.0f, 0.0f, 0.0f};
                for (uint32_t s = 0; s < p.spp; s++) {
                    if (ctr) ctr->primary++;
                    Ray ray = camera_get_ray(cam, x, y_cam, rng);
                    Color c = ray_color(ray, sc, p.max_bounces + 1, rng, ctr, scratch);
                    pix = Color{pix.r + c.r, pix.g + c.g, pix.b + c.b};
                }
                float n = (float)p.spp;
                pix.r = std::sqrt(pix.r / n);
                pix.g = std::sqrt(pix.g / n);
                pix.b = std::sqrt(pix.b / n);
                row[x * 3 + 0] = f32_as_u8(pix.r * 255.999f);
                row[x * 3 + 1] = f32_as_u8(pix.g * 255.999f);
                row[x * 3 + 2] = f32_as_u8(pix.b * 255.999f);
            }
        }
    };
    if (nthreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    double t2 = now_ms();
    if (stats) {
        Counters tot;
        for (auto& c : ctrs) tot.add(c);
        stats->rays = tot.rays; stats->primary = tot.primary; stats->aabb_tests = tot.aabb_tests;
        stats->sphere_tests = tot.sphere_tests; stats->sphere_hits = tot.sphere_hits;
        stats->tri_tests = tot.tri_tests;
        for (int i = 0; i < 4; i++) stats->tri_exit[i] = tot.tri_exit[i];
        stats->tri_hits = tot.tri_hits; stats->shades_sphere = tot.shades_sphere;
        stats->shades_tri = tot.shades_tri; stats->emissive = tot.emissive; stats->sky = tot.sky;
        stats->depth_exhausted = tot.depth_exhausted; stats->rng_draws = tot.rng_draws;
        stats->build_ms = t1 - t0;
        stats->render_ms = t2 - t1;
    }
    return 0;
}

// Whole division band = worker() body main.rs:53-83.
int orc_render_division(const orc_sphere* sph, uint32_t ns, const orc_triangle* tri, uint32_t nt,
                        const uint32_t* world_index, const orc_params* params_in, uint8_t* out, size_t out_len,
                        int mode, int threads, orc_stats* stats) {
    if (!params_in) return -1;
    orc_params p = *params_in;
    fill_defaults(&p);
    if (p.height % p.divisions != 0) return -1;
    return orc_render_rows(sph, ns, tri, nt, world_index, params_in, 0, p.height / p.divisions, out, out_len,
                           mode, threads, stats);
}

// ------------------------------- KAT / unit-test entry points ---------------------------------
uint64_t orc_splitmix64(uint64_t* state) { return splitmix64_next(state); }
void orc_seed_from_u64(uint64_t seed, uint64_t state_out[4]) {
    SmallRng r = SmallRng::seed_from_u64(seed);
    for (int i = 0; i < 4; i++) state_out[i] = r.s[i];
}
void orc_xoshiro_next_u64(uint64_t state[4], uint32_t n, uint64_t* out) {
    SmallRng r;
    for (int i = 0; i < 4; i++) r.s[i] = state[i];
    for (uint32_t i = 0; i < n; i++) out[i] = r.next_u64();
    for (int i = 0; i < 4; i++) state[i] = r.s[i];
}
// kind: 0 value0_1, 1 gen_range(0..1), 2 Uniform(-1,1); writes n floats
void orc_rng_floats(uint64_t seed, int kind, uint32_t n, float* out) {
    SmallRng r = SmallRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < n; i++)
        out[i] = kind == 0 ? r.value0_1() : (kind == 1 ? r.gen_range_0_1() : r.uniform_m1_1());
}
void orc_unit_disc(uint64_t seed, uint32_t n, float* out2) {
    SmallRng r = SmallRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < n; i++) unit_disc(r, &out2[2 * i], &out2[2 * i + 1]);
}
void orc_unit_sphere(uint64_t seed, uint32_t n, float* out3) {
    SmallRng r = SmallRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < n; i++) {
        V3 v = unit_sphere(r);
        out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z;
    }
}
int orc_find_roots_quadratic(float a2, float a1, float a0, float out[2]) {
    Roots r = find_roots_quadratic(a2, a1, a0);
    out[0] = r.r[0]; out[1] = r.r[1];
    return r.n;
}
// Ray::new + intersects_aabb
int orc_ray_intersects_aabb(const float o[3], const float d[3], const float bmin[3], const float bmax[3]) {
    Ray r = ray_new(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]));
    AABB b{v3(bmin[0], bmin[1], bmin[2]), v3(bmax[0], bmax[1], bmax[2])};
    return intersects_aabb(r, b) ? 1 : 0;
}
// Generic BVH over AABBs (the 21-unit-box KAT, testbase.rs:92-166): builds, traverses with
// Ray::new(o,d); writes hit shape indices (DFS order) to out_idx, returns the count or -2.
int orc_bvh_traverse_boxes(const float* boxes6, uint32_t n, const float o[3], const float d[3],
                           uint32_t* out_idx, uint32_t* out_node_count, uint32_t* shape_node_out) {
    std::vector<AABB> bx(n);
    for (uint32_t i = 0; i < n; i++)
        bx[i] = AABB{v3(boxes6[6 * i], boxes6[6 * i + 1], boxes6[6 * i + 2]),
                     v3(boxes6[6 * i + 3], boxes6[6 * i + 4], boxes6[6 * i + 5])};
    BVH bvh;
    if (!bvh_build(bx, &bvh)) return -2;
    Ray r = ray_new(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]));
    std::vector<uint32_t> idx;
    traverse_recursive(bvh.nodes, 0, r, idx, nullptr);
    for (size_t i = 0; i < idx.size(); i++) out_idx[i] = idx[i];
    if (out_node_count) *out_node_count = (uint32_t)bvh.nodes.size();
    if (shape_node_out) for (uint32_t i = 0; i < n; i++) shape_node_out[i] = bvh.shape_node[i];
    return (int)idx.size();
}
// DFS (left-first) leaf order of the reference BVH over a world: rank_out[world position] = rank.
int orc_bvh_leaf_order(const orc_sphere* sph, uint32_t ns, const orc_triangle* tri, uint32_t nt,
                       const uint32_t* world_index, uint32_t* rank_out, uint32_t* depth_out) {
    std::vector<Object> world;
    if (!make_world(sph, ns, tri, nt, world_index, &world)) return -1;
    std::vector<AABB> boxes(world.size());
    for (size_t i = 0; i < boxes.size(); i++) boxes[i] = object_aabb(world[i]);
    BVH bvh;
    if (!bvh_build(boxes, &bvh)) return -2;
    uint32_t rank = 0, maxd = 0;
    for (const BVHNode& n : bvh.nodes) {  // nodes are stored in DFS pre-order
        if (n.leaf) rank_out[n.shape] = rank++;
        if (n.depth > maxd) maxd = n.depth;
    }
    if (depth_out) *depth_out = maxd;
    return 0;
}
// Primary ray for pixel (x, y_cam) from a fresh per-pixel stream: out = origin(3), direction(3)
void orc_get_ray(const orc_params* params_in, uint32_t x, uint32_t y_cam, uint64_t stream_seed, float out[6]) {
    orc_params p = *params_in;
    fill_defaults(&p);
    Camera cam = camera_new(v3(p.cam_origin[0], p.cam_origin[1], p.cam_origin[2]),
                            (float)p.width / (float)p.height, p.aperture, p.focus_distance,
                            p.field_of_view, p.focal_length, (float)p.height);
    SmallRng rng = SmallRng::seed_from_u64(stream_seed);
    Ray r = camera_get_ray(cam, x, y_cam, rng);
    out[0] = r.origin.x; out[1] = r.origin.y; out[2] = r.origin.z;
    out[3] = r.direction.x; out[4] = r.direction.y; out[5] = r.direction.z;
}
// Nearest hit of Ray::new(o,d) against a world (mode as in orc_render_rows).
// out = point(3), normal(3), distance-from-origin, world position of the winner. Returns 1 hit / 0 miss.
int orc_nearest_hit(const orc_sphere* sph, uint32_t ns, const orc_triangle* tri, uint32_t nt,
                    const uint32_t* world_index, const float o[3], const float d[3], int mode, float out[8]) {
    Scene sc;
    sc.mode = mode;
    if (!make_world(sph, ns, tri, nt, world_index, &sc.world)) return -1;
    if (mode == 0) {
        std::vector<AABB> boxes(sc.world.size());
        for (size_t i = 0; i < boxes.size(); i++) boxes[i] = object_aabb(sc.world[i]);
        if (!bvh_build(boxes, &sc.bvh)) return -2;
    }
    Ray r = ray_new(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]));
    IntersectionTable tb;
    bool hit;
    if (mode == 0) {
        std::vector<uint32_t> c;
        traverse_recursive(sc.bvh.nodes, 0, r, c, nullptr);
        hit = intersect_list(sc.world, c.data(), c.size(), r, &tb, nullptr);
    } else {
        hit = intersect_list(sc.world, nullptr, sc.world.size(), r, &tb, nullptr);
    }
    if (!hit) return 0;
    out[0] = tb.point.x; out[1] = tb.point.y; out[2] = tb.point.z;
    out[3] = tb.normal.x; out[4] = tb.normal.y; out[5] = tb.normal.z;
    out[6] = length(tb.point - r.origin);
    out[7] = (float)tb.index;
    return 1;
}
// Traces every sample of one pixel and logs each nearest-hit query: 7 floats per ray
// (origin, direction, winner's world position or -1). Returns the number of rays (<= max_rays logged).
int orc_trace_pixel(const orc_sphere* sph, uint32_t ns, const orc_triangle* tri, uint32_t nt,
                    const uint32_t* world_index, const orc_params* params_in, uint32_t x, uint32_t y_global, int mode,
                    float* log_out, uint32_t max_rays) {
    orc_params p = *params_in;
    fill_defaults(&p);
    Scene sc;
    sc.mode = mode;
    if (!make_world(sph, ns, tri, nt, world_index, &sc.world)) return -1;
    if (mode == 0) {
        std::vector<AABB> boxes(sc.world.size());
        for (size_t i = 0; i < boxes.size(); i++) boxes[i] = object_aabb(sc.world[i]);
        if (!bvh_build(boxes, &sc.bvh)) return -2;
    }
    Camera cam = camera_new(v3(p.cam_origin[0], p.cam_origin[1], p.cam_origin[2]),
                            (float)p.width / (float)p.height, p.aperture, p.focus_distance,
                            p.field_of_view, p.focal_length, (float)p.height);
    SmallRng rng = SmallRng::seed_from_u64(p.seed + ((uint64_t)y_global * p.width + x));
    std::vector<float> log;
    std::vector<uint32_t> scratch;
    g_ray_log = &log;
    for (uint32_t s = 0; s < p.spp; s++) {
        Ray ray = camera_get_ray(cam, x, p.height - y_global - 1, rng);
        ray_color(ray, sc, p.max_bounces + 1, rng, nullptr, scratch);
    }
    g_ray_log = nullptr;
    uint32_t n = (uint32_t)(log.size() / 7);
    for (uint32_t i = 0; i < n && i < max_rays; i++)
        for (int k = 0; k < 7; k++) log_out[7 * i + k] = log[7 * i + k];
    return (int)n;
}
uint8_t orc_f32_as_u8(float f) { return f32_as_u8(f); }
int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
