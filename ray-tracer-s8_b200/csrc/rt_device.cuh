// rt_device.cuh — device-side building blocks of the render path (sm_100a).
//
// Two arithmetic domains live side by side:
//   * EXACT  (x_* helpers): single IEEE-754 round-to-nearest f32 operations in the reference's
//     operation order (glam 0.23 Vec3A/SSE2, roots 0.0.8, rand 0.8.5,
//     rand_distr 0.4.3).  Built only from __fadd_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn, which nvcc never
//     contracts into FFMA, so the results are bit-identical to the CPU's SSE scalar ops.  Everything
//     that decides a path (roots, t-range, nearest hit, scatter direction, colour) is EXACT.
//   * FILTER (plain fmaf / fminf / fmaxf): conservative culling only (sphere discriminant pre-test,
//     BVH slab tests).  A filter may let a miss through; it must never reject an exact hit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtb {

constexpr int TILE_W = 8;        // a warp renders an 8x4-pixel tile
constexpr int TILE_H = 4;
constexpr int MAX_STACK = 64;    // BVH traversal stack entries (host rejects deeper trees)
constexpr int MAX_BIG = 8;      // primitives tested ahead of the traversal (split layout)
constexpr int MAX_PATH = 64;     // max_bounces + 1 <= MAX_PATH
constexpr float T_MIN = 0.001f;  // shapes/mod.rs:12
constexpr float T_MAX = 1000.0f; // shapes/mod.rs:13

// ---------------------------------------------------------------------------------------------
// Scene in HBM (structure of arrays).  Primitive id ("pid"): [0,ns) spheres, [ns,ns+nt) triangles,
// both stored in the DFS leaf order of the reference BVH.  BVH child code: >= 0 inner node index,
// < 0 leaf with pid = ~code.
// ---------------------------------------------------------------------------------------------
// 64 bytes: the same field of different nodes then falls on two of the eight 16-byte bank groups of shared memory (2.3 G
// bank conflicts on C3, shared wavefronts at 56 % of peak) — an 80-byte record spreads it over all eight and measures
// the same 36.33 ms (profiles/r2_notes.md), so the smaller record, which keeps larger scenes in shared memory, stays
constexpr int NODE_BYTES = 64;
constexpr int NODE_WORDS = NODE_BYTES / 4, NODE_F4 = NODE_BYTES / 16;
// Order of the twelve box floats inside a record.  RT_NODE_XY (default): x and y of a box's centre and of its half extent
// are adjacent 64-bit pairs, and so are the z of the two boxes, so the slab tests run on packed FMAs (fma.rn.f32x2, SASS
// FFMA2: nine instead of eighteen FFMA per visit):
//   l.c.x l.c.y l.h.x l.h.y | r.c.x r.c.y r.h.x r.h.y | l.c.z r.c.z l.h.z r.h.z
// otherwise (A/B build -DRT_NODE_XY=0, scalar FMAs): l.c.xyz l.h.x | l.h.yz r.c.xy | r.c.z r.h.xyz
#ifndef RT_NODE_XY
#define RT_NODE_XY 1
#endif
__host__ __device__ inline void node_box_words(const float lc[3], const float lh[3], const float rc[3], const float rh[3], float w[12]) {
#if RT_NODE_XY
    w[0] = lc[0]; w[1] = lc[1]; w[2] = lh[0]; w[3] = lh[1];
    w[4] = rc[0]; w[5] = rc[1]; w[6] = rh[0]; w[7] = rh[1];
    w[8] = lc[2]; w[9] = rc[2]; w[10] = lh[2]; w[11] = rh[2];
#else
    w[0] = lc[0]; w[1] = lc[1]; w[2] = lc[2]; w[3] = lh[0];
    w[4] = lh[1]; w[5] = lh[2]; w[6] = rc[0]; w[7] = rc[1];
    w[8] = rc[2]; w[9] = rh[0]; w[10] = rh[1]; w[11] = rh[2];
#endif
}
struct DevScene {
    // The first three arrays are adjacent in the scene blob, in this order, each a multiple of 16 bytes:
    //   sph | tri | lnode                      (the shared-memory image of the BVH kernel: ONE bulk copy per CTA)
    const float4* sph;     // [ns]    cx, cy, cz, r*r
    const float4* tri;     // [nt*4]  a | b-a | c-a | normalize_or_zero((a-b)x(a-c))
    // the traversal tree: one NODE_BYTES record per inner node — two child boxes in centre/half-extent form, padded for
    // FILTER rounding (a | b | c: l.c.xyz, l.h.x | l.h.yz, r.c.xy | r.c.z, r.h.xyz), then the two child codes as the
    // x, y of a fourth int4, then padding: >= 0 inner node, as the BYTE OFFSET of its record from lnode (a visit adds a base and
    // loads; the shared-memory kernel adds the base once, in its image, so its visits load from the code itself),
    // < 0 leaf ~(pid << 5)
    const float4* lnode;    // [lni*NODE_F4]
#ifdef RT_B200_EXPERIMENTS
    const float4* lnode_a;  // [lni*3] the same records as 48 bytes + codes with plain node indices, for the
    const int2* lnode_d;    // [lni]   experiment kernels
#endif
    const float4* sph2;    // [ceil8(ns)] sphere pairs for the packed f32x2 filter: -c.x pair, -c.y pair | -c.z pair, r*r pair
    uint32_t lni;          // inner nodes of the traversal tree
    int lroot;             // its root code
    // "split" traversal layout: the few primitives whose box is a large share of the scene's (a ground plane's two
    // triangles) are kept out of the tree and tested first, so they neither inflate the upper boxes nor cost node
    // visits; the tree then covers the remaining primitives only (ltree == 0: nothing remains, no traversal).
    uint32_t nbig;
    int ltree;
    const int* big_code;   // [nbig] their leaf codes ~(pid << 5), in HBM: the traversal stack starts with them
    const float4* mat;     // [ns+nt] albedo rgb, roughness
    const float* emis;     // [ns+nt]
    const uint32_t* rank;  // [ns+nt] DFS leaf rank in the REFERENCE tree (exact-distance tie-break, shapes/mod.rs:177-182)
    const float4* leaf_box;  // [(ns+nt)*2] the shape's own AABB exactly as the reference computes it (min | max)
    // the reference tree's ancestor chain, walked only by rays with a zero direction component (NaN / inf slabs,
    // ray.rs:82-112,174-194): up[pid] / up[n + node] = (parent inner node << 1) | side, UP_ROOT at the root;
    // ref_box[2*(2*node + side)] = that child's box (min | max), unpadded
    const uint32_t* ref_up;
    const float4* ref_box;
    int aux_ready;         // rank / ref_up / ref_box had landed when the kernel was launched (they are built beside the upload,
                           // rt_scene.cu).  While 0, a query that needs them polls *aux_flag (set once they land) and, if
                           // still 0, marks its pixel for a second pass
    const int* aux_flag;
    uint32_t ns, nt, ni;   // ni = inner nodes of the reference tree
#ifdef RT_B200_EXPERIMENTS
    uint32_t big_pid[MAX_BIG];  // the big primitives' ids, as the A/B kernels read them
    // 4-ary collapse of the host-built traversal tree (trace-bench experiment RT_B200_TB_ALT=2): 7 float4 per node:
    // [c0.xyz h0.x][h0.yz c1.xy][c1.z h1.xyz][c2.xyz h2.x][h2.yz c3.xy][c3.z h3.xyz][4 child codes]; an empty slot has
    // h = -1 (never hit)
    const float4* w4;
    uint32_t w4n;
    int w4root;
    // the reference-topology tree in the first kernels' formats (A/B kernels under csrc/experiments/ only)
    const float4* node_a;  // [ni]    l.min.xyz, l.max.x
    const float4* node_b;  // [ni]    l.max.yz,  r.min.xy
    const float4* node_c;  // [ni]    r.min.z,   r.max.xyz
    const int2* node_d;    // [ni]    left code, right code
    const float4* cnode_a; // [ni]    centre/half-extent form of the same boxes:
    const float4* cnode_b; //         l.c.xyz, l.h.x | l.h.yz, r.c.xy | r.c.z, r.h.xyz
    const float4* cnode_c;
    int root;              // child code of the root
#endif
};
constexpr uint32_t UP_ROOT = 0xffffffffu;
constexpr uint32_t REDO_VALID = 0x80000000u;

struct DevCamera {            // Camera::new (camera.rs:19-47), evaluated on the host in reference order
    float org[3], llc[3], hor[3], ver[3];
    float lens_radius;        // aperture / 2
    float u_den, v_den;       // aspect*image_height - 1, image_height - 1
    float focus;
};

struct DevParams {
    uint32_t width, height;
    uint32_t row0, row1;         // global image rows [row0,row1) rendered by this launch
    uint32_t spp, depth;         // depth = max_bounces + 1 nearest-hit queries per sample at most
    uint64_t seed;
    uint32_t tile_rank, tile_ranks;
    uint8_t* out;                // RGB8; buffer row 0 = global row out_row0 (may be a peer-mapped frame: stores travel over NVLink)
    uint32_t out_row0;
    uint32_t tiles_x, tiles_y;
    // work hand-out: tickets [0, tail_first) are whole 8x4 tiles pulled by warps; tickets [tail_first, my_tickets) are
    // handed out pixel by pixel to single lanes, so the last few per cent of the launch keep every lane busy
    unsigned int* tile_counter;  // [0] tile tickets, [1] pixel tickets of the tail; zeroed before launch
    uint32_t my_tickets, tail_first;
    int tile_order_reverse;      // 1: tickets walk the tile grid from the last tile to the first
    // completion counting (nullable): done[slab] += pixels written, slab = tile row / slab_tile_rows; released
    // with a system-scope fence so the frame owner may copy a slab out as soon as its count is complete
    unsigned long long* done;
    uint32_t slab_tile_rows;
    int stage_hint;              // the caller will stream slabs to the host: use the output stage even for a single rank
    unsigned long long* counters;  // NUM_COUNTERS
    // second pass: a pixel whose queries needed the tie-break tables before they had landed is appended to redo_list
    // (REDO_VALID | (y * width + x); the list is zeroed before the launch) and NOT added to `done`; redo_slab[slab] counts
    // such pixels.  When the launch runs out of tickets and the tables have landed meanwhile, idle lanes render listed
    // pixels again (ticket[2] = entries taken) — what is left (or everything, if *redo_count exceeds the capacity) is the
    // host's (rt_api.cu finish_redo), as a launch with pixel_list != nullptr: exactly those pixels, pixel tickets only
    unsigned int* redo_list;
    unsigned long long* redo_count;
    uint32_t redo_cap;
    unsigned long long* redo_slab;
    const unsigned int* pixel_list;
    uint32_t list_count;
#ifdef RT_B200_EXPERIMENTS
    // scheduled kernels: pool weights (NODE, LEAF, HIT, PRIM) and the NODE phase's stay-in-loop share num/den
    int sched_w[4];
    int sched_node_num, sched_node_den;
    // measurement aid (COUNT instantiation only): every query's ray is appended here (rt_trace_bench.cuh)
    float4* ray_dump;
    unsigned long long* ray_dump_n;
    unsigned long long ray_dump_cap;
#endif
};

enum CounterSlot {
    CTR_RAYS = 0, CTR_SLAB, CTR_SPH_TEST, CTR_SPH_EXACT, CTR_SPH_HIT, CTR_TRI_TEST, CTR_TRI_S1, CTR_TRI_S2, CTR_TRI_S3,
    CTR_TRI_HIT, CTR_SHADE_SPH, CTR_SHADE_TRI, CTR_EMISSIVE, CTR_SKY, CTR_ACTIVE_LANES, CTR_TOTAL_LANES, NUM_COUNTERS
};

// ---------------------------------------------------------------------------------------------
// EXACT domain
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float x_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float x_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float x_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float x_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float x_sqrt(float a) { return __fsqrt_rn(a); }

// Packed FP32 pairs (sm_100a FADD2 / FMUL2 / FFMA2): one issue slot carries two FP32 operations, each half rounded
// like the scalar instruction.  Used where the operands ARE pairs already (node records, sphere pairs); packing the
// EXACT domain's vector helpers (x and y of x_add / x_sub / x_dot) was measured and lost 2 % to the moves that build
// the pairs (profiles/r2_notes.md).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 x_add(V3 a, V3 b) { return mk(x_add(a.x, b.x), x_add(a.y, b.y), x_add(a.z, b.z)); }
__device__ __forceinline__ V3 x_sub(V3 a, V3 b) { return mk(x_sub(a.x, b.x), x_sub(a.y, b.y), x_sub(a.z, b.z)); }
__device__ __forceinline__ V3 x_scale(V3 a, float s) { return mk(x_mul(a.x, s), x_mul(a.y, s), x_mul(a.z, s)); }
// glam dot3: (x*x' + y*y') + z*z'
__device__ __forceinline__ float x_dot(V3 a, V3 b) {
    return x_add(x_add(x_mul(a.x, b.x), x_mul(a.y, b.y)), x_mul(a.z, b.z));
}
__device__ __forceinline__ float x_length(V3 a) { return x_sqrt(x_dot(a, a)); }
// glam cross: (a.zxy*b - a*b.zxy).zxy
__device__ __forceinline__ V3 x_cross(V3 a, V3 b) {
    return mk(x_sub(x_mul(a.y, b.z), x_mul(b.y, a.z)), x_sub(x_mul(a.z, b.x), x_mul(b.z, a.x)),
              x_sub(x_mul(a.x, b.y), x_mul(b.x, a.y)));
}
// The normalisations hold most of the kernel's divisions and square roots, each of which expands to a dozen
// instructions plus a slow-path call.  RT_OUTLINE_NORM keeps ONE copy of each out of line (arguments and results in
// registers): the render kernel's instruction working set sits at the edge of the instruction cache, and code bytes
// cost more there than a call does (profiles/r2_notes.md).
#ifdef RT_OUTLINE_NORM
#define RT_NORM_FN static __noinline__
#else
#define RT_NORM_FN __forceinline__
#endif
// Vec3A::normalize: v / sqrt(dot) per lane (used by Ray::new, ray.rs:134)
__device__ RT_NORM_FN V3 x_normalize_div(V3 a) {
    float l = x_length(a);
    return mk(x_div(a.x, l), x_div(a.y, l), x_div(a.z, l));
}
// Vec3A::try_normalize: rcp = 1/length; Some(v*rcp) iff rcp finite && rcp > 0.  `or` is returned otherwise.
__device__ RT_NORM_FN V3 x_normalize_or(V3 a, V3 orv) {
    float rcp = x_div(1.0f, x_length(a));
    if (isfinite(rcp) && rcp > 0.0f) return x_scale(a, rcp);
    return orv;
}
__device__ __forceinline__ bool x_try_normalize(V3 a, V3* out) {
    float rcp = x_div(1.0f, x_length(a));
    if (isfinite(rcp) && rcp > 0.0f) {
        *out = x_scale(a, rcp);
        return true;
    }
    return false;
}
__device__ __forceinline__ V3 x_normalize_or_zero(V3 a) { return x_normalize_or(a, mk(0.0f, 0.0f, 0.0f)); }

// rand 0.8.5 SmallRng on 64-bit targets = xoshiro256++; seed_from_u64 = 4 SplitMix64 outputs
struct Rng {
    uint64_t s0, s1, s2, s3;
    __device__ __forceinline__ static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    __device__ __forceinline__ static uint64_t splitmix(uint64_t& st) {
        st += 0x9e3779b97f4a7c15ull;
        uint64_t z = st;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    __device__ __forceinline__ void seed_from_u64(uint64_t st) {
        s0 = splitmix(st);
        s1 = splitmix(st);
        s2 = splitmix(st);
        s3 = splitmix(st);
        if ((s0 | s1 | s2 | s3) == 0) {  // from_seed: all-zero → seed_from_u64(0)
            uint64_t z = 0;
            s0 = splitmix(z);
            s1 = splitmix(z);
            s2 = splitmix(z);
            s3 = splitmix(z);
        }
    }
    __device__ __forceinline__ uint32_t next_u32() {  // high half of next_u64
        uint64_t result = rotl(s0 + s3, 23) + s0;
        uint64_t t = s1 << 17;
        s2 ^= s0;
        s3 ^= s1;
        s1 ^= s2;
        s0 ^= s3;
        s2 ^= t;
        s3 = rotl(s3, 45);
        return (uint32_t)(result >> 32);
    }
    // UniformFloat<f32>: 23 mantissa bits into [1,2), minus 1
    __device__ __forceinline__ float value0_1() {
        return x_sub(__uint_as_float(0x3f800000u | (next_u32() >> 9)), 1.0f);
    }
    // gen_range(0f32..1f32): value0_1 * 1 + 0
    __device__ __forceinline__ float gen_range_0_1() { return x_add(x_mul(value0_1(), 1.0f), 0.0f); }
    // Uniform::new(-1,1).sample: value0_1 * 2 + (-1)
    __device__ __forceinline__ float uniform_m1_1() { return x_add(x_mul(value0_1(), 2.0f), -1.0f); }
};

// rand_distr 0.4.3 UnitDisc: rejection, accept x1^2 + x2^2 <= 1
__device__ __forceinline__ void unit_disc(Rng& rng, float& a, float& b) {
    for (;;) {
        a = rng.uniform_m1_1();
        b = rng.uniform_m1_1();
        if (x_add(x_mul(a, a), x_mul(b, b)) <= 1.0f) break;
    }
}
// rand_distr 0.4.3 UnitSphere (Marsaglia 1972)
__device__ __forceinline__ V3 unit_sphere(Rng& rng) {
    for (;;) {
        float x1 = rng.uniform_m1_1();
        float x2 = rng.uniform_m1_1();
        float sum = x_add(x_mul(x1, x1), x_mul(x2, x2));
        if (sum >= 1.0f) continue;
        float factor = x_mul(2.0f, x_sqrt(x_sub(1.0f, sum)));
        return mk(x_mul(x1, factor), x_mul(x2, factor), x_sub(1.0f, x_mul(2.0f, sum)));
    }
}

__device__ __forceinline__ bool in_range(float t) { return t >= T_MIN && t < T_MAX; }

// Sphere::get_roots (sphere.rs:42-47) + find_roots_quadratic (roots 0.0.8) + root pick
// (shapes/mod.rs:106-129).  oc = origin - center (exact), r2 = radius*radius (exact).
// Returns true and the chosen in-range root.
__device__ __forceinline__ bool sphere_root_exact(V3 d, V3 oc, float r2, float* t_out) {
    float b = x_dot(x_scale(d, 2.0f), oc);  // (2*d).dot(o - c)
    float l = x_length(oc);
    float c = x_sub(x_mul(l, l), r2);       // length().powi(2) - radius.powi(2)
    float disc = x_sub(x_mul(b, b), x_mul(4.0f, c));  // a1*a1 - 4*a2*a0, a2 = 1
    if (disc < 0.0f) return false;
    if (disc == 0.0f) {
        float x = x_mul(-b, 0.5f);  // -a1 / (2*a2): a division by two is exact, and so is this product
        if (!in_range(x)) return false;
        *t_out = x;
        return true;
    }
    float sq = x_sqrt(disc);
    float same_sign, diff_sign;
    if (b < 0.0f) {
        same_sign = x_add(-b, sq);
        diff_sign = x_sub(-b, sq);
    } else {
        same_sign = x_sub(-b, sq);
        diff_sign = x_add(-b, sq);
    }
    float x1, x2;
    if (fabsf(same_sign) > 2.0f) {
        float a0x2 = x_mul(2.0f, c);
        x1 = x_div(a0x2, same_sign);
        x2 = (fabsf(diff_sign) > 2.0f) ? x_div(a0x2, diff_sign) : x_mul(same_sign, 0.5f);
    } else {
        x1 = x_mul(diff_sign, 0.5f);
        x2 = x_mul(same_sign, 0.5f);
    }
    float lo, hi;
    if (x1 < x2) {
        lo = x1;
        hi = x2;
    } else {
        lo = x2;
        hi = x1;
    }
    bool li = in_range(lo), hi_in = in_range(hi);
    if (li && hi_in) {
        *t_out = lo < hi ? lo : hi;
        return true;
    }
    if (li) {
        *t_out = lo;
        return true;
    }
    if (hi_in) {
        *t_out = hi;
        return true;
    }
    return false;
}

// Triangle::get_roots (mesh.rs:109-161), two-sided Moeller-Trumbore, + t-range (shapes/mod.rs:109-115).
// *stage = number of rejection tests passed (0 det, 1 u, 2 v, 3 reached dist) for the FLOP accounting.
__device__ __forceinline__ bool triangle_root_exact(V3 o, V3 d, V3 a, V3 ab, V3 ac, float* t_out, int* stage) {
    const float EPSILON = 0.00001f;
    *stage = 0;
    V3 u_vec = x_cross(d, ac);
    float det = x_dot(ab, u_vec);
    if (det < EPSILON && det > -EPSILON) return false;
    *stage = 1;
    float inv_det = x_div(1.0f, det);
    V3 ao = x_sub(o, a);
    float u = x_mul(x_dot(ao, u_vec), inv_det);
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    *stage = 2;
    V3 v_vec = x_cross(ao, ab);
    float v = x_mul(x_dot(d, v_vec), inv_det);
    if (v < 0.0f || x_add(u, v) > 1.0f) return false;
    *stage = 3;
    float dist = x_mul(x_dot(ac, v_vec), inv_det);
    if (!(dist > EPSILON)) return false;
    if (!in_range(dist)) return false;
    *t_out = dist;
    return true;
}

__device__ __forceinline__ V3 ld3(const float4& v) { return mk(v.x, v.y, v.z); }

// Ray::intersects_aabb (bvh/src/ray.rs:174-194) — EXACT, with Ray::new's cached 1/d and signs
// (ray.rs:133-143) and the crate's own min/max (ray.rs:82-112: `if x < y {x} else {y}`).
//
// Why the GPU path needs it: the reference's BVH is NOT a conservative cull.  A primitive whose exact
// hit lies a rounding error outside its own AABB (e.g. the border of an axis-aligned triangle) is
// dropped by bvh.traverse() before intersect() ever sees it (main.rs:113-114).  Float subtraction and
// multiplication are monotonic, and every ancestor's child box contains the shape's own box, so
// "the shape's own AABB passes this test" implies every ancestor passes: one test per exact hit
// reproduces the reference's candidate set (NaN slabs from 0*inf aside).
__device__ __forceinline__ bool ref_intersects_aabb(V3 o, V3 d, V3 lo, V3 hi) {
    const float ivx = x_div(1.0f, d.x), ivy = x_div(1.0f, d.y), ivz = x_div(1.0f, d.z);
    const bool sx = d.x < 0.0f, sy = d.y < 0.0f, sz = d.z < 0.0f;
    float ray_min = x_mul(x_sub(sx ? hi.x : lo.x, o.x), ivx);
    float ray_max = x_mul(x_sub(sx ? lo.x : hi.x, o.x), ivx);
    const float y_min = x_mul(x_sub(sy ? hi.y : lo.y, o.y), ivy);
    const float y_max = x_mul(x_sub(sy ? lo.y : hi.y, o.y), ivy);
    ray_min = (ray_min > y_min) ? ray_min : y_min;
    ray_max = (ray_max < y_max) ? ray_max : y_max;
    const float z_min = x_mul(x_sub(sz ? hi.z : lo.z, o.z), ivz);
    const float z_max = x_mul(x_sub(sz ? lo.z : hi.z, o.z), ivz);
    ray_min = (ray_min > z_min) ? ray_min : z_min;
    ray_max = (ray_max < z_max) ? ray_max : z_max;
    const float lo_t = (ray_min > 0.0f) ? ray_min : 0.0f;
    return lo_t <= ray_max;
}

}  // namespace rtb
