// rt_kernels.cu — the render megakernels (sm_100a) and their launcher.
//
// One persistent launch renders a set of 8x4-pixel tiles.  Warps pull tiles from a global ticket
// counter (dynamic load balance at warp granularity); a lane owns one pixel at a time and runs that pixel's
// whole sample/bounce chain from its own xoshiro256++ stream, because the reference draws all of a
// pixel's samples and bounces sequentially from one generator (ray-tracer-slave/src/main.rs:69-77).
// The reference's recursion (ray_color, main.rs:108-146) is flattened into ONE loop whose trip is a
// single nearest-hit query: a lane that finishes a path starts its next sample in the same trip
// structure, so lanes of a warp stay in the intersection code together regardless of bounce index.
//
// The product kernel is render_kernel_lanes (lanes are independent workers: a lane that finishes its pixel takes
// the next pixel of the warp's tile), launched as ONE 768-thread CTA per SM when the scene fits shared memory
// (geometry + traversal tree staged once per SM, the rest of the 228 KB left to L1 for the traversal stacks) and
// as 4 x 256 threads at 64 registers when the scene is read through L1/L2.  render_kernel (tile per warp) is the
// first form, kept for A/B runs together with rt_kernel_sched / _deferred / _wq.cuh and rt_wavefront.cuh.
//
//   K1 (ISECT_BRUTE): every primitive per query; the sphere FILTER runs on pairs of spheres in packed f32x2
//                     arithmetic (FADD2 / FMUL2 / FFMA2), exact reference arithmetic only where it passes.
//   K2 (ISECT_BVH):   ordered, distance-culled traversal of the traversal tree (big primitives
//                     first, then a SAH or LBVH tree over the rest) with conservative FMA slab tests in
//                     centre/half-extent form and a branch-free visit; exact arithmetic at the leaves.
#include "rt_device.cuh"
#include "rt_host.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace rtb {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

struct Hit {
    float dist;  // length(point - origin), exact domain
    int pid;     // -1 = miss
    V3 p;        // ray.at(t)
};

struct Ctr {
    unsigned long long v[NUM_COUNTERS];
};

// Shared-memory view of the geometry arrays (or the global pointers when SMEM == false)
struct SceneView {
    const float4* sph2;  // brute-force kernel only
    const float4* sph;
    const float4* tri;
    const float4* na;
    const float4* nb;
    const float4* nc;
    const int2* nd;
};

// ---------------------------------------------------------------------------------------------
// Leaf tests.  `best` is updated iff the candidate wins the reference's min_by: smaller
// length(p - o), ties to the smaller DFS leaf rank (shapes/mod.rs:174-182, bvh_impl.rs:373-398).
// ---------------------------------------------------------------------------------------------
// FILTER-domain shortcut for the reference's slab test on the shape's own box: when the hit point is inside
// the box by a margin that dominates every rounding error of ray.rs:174-194 (2^-23 relative on each slab
// product, plus the ~1e-6*t disagreement between a triangle's Moeller-Trumbore t and its flat box's slab t),
// the reference test passes for certain.  Axes on which the box is flat (lo == hi: an axis-aligned triangle)
// give the reference tmin == tmax bit for bit, so only the other axes need the margin.
__device__ __forceinline__ bool robustly_inside(V3 p, float t, V3 lo, V3 hi, float extra = 0.0f) {
    const float m0 = fmaf(4e-5f, fabsf(t), extra);
    bool ok = true;
    {
        const float m = fmaf(1e-6f, fabsf(p.x) + fabsf(lo.x) + fabsf(hi.x), m0);
        ok = ok && ((lo.x == hi.x) || ((p.x - lo.x >= m) && (hi.x - p.x >= m)));
    }
    {
        const float m = fmaf(1e-6f, fabsf(p.y) + fabsf(lo.y) + fabsf(hi.y), m0);
        ok = ok && ((lo.y == hi.y) || ((p.y - lo.y >= m) && (hi.y - p.y >= m)));
    }
    {
        const float m = fmaf(1e-6f, fabsf(p.z) + fabsf(lo.z) + fabsf(hi.z), m0);
        ok = ok && ((lo.z == hi.z) || ((p.z - lo.z >= m) && (hi.z - p.z >= m)));
    }
    return ok;
}

__device__ __forceinline__ void consider(const DevScene& sc, V3 o, V3 d, float t, int pid, Hit& best) {
    V3 p = x_add(o, x_scale(d, t));   // Ray::at: origin + t*direction
    // bvh.traverse() (main.rs:113): the shape is a candidate only if the reference's slab test lets it through
    if (sc.ns + sc.nt > 1) {
        const V3 blo = ld3(__ldg(&sc.leaf_box[2 * pid])), bhi = ld3(__ldg(&sc.leaf_box[2 * pid + 1]));
        if (!robustly_inside(p, t, blo, bhi) && !ref_intersects_aabb(o, d, blo, bhi)) return;
    }
    float dist = x_length(x_sub(p, o));
    bool take;
    if (best.pid < 0) {
        take = true;
    } else if (best.dist > dist) {
        take = true;
    } else if (best.dist == dist) {
        take = __ldg(&sc.rank[pid]) < __ldg(&sc.rank[best.pid]);
    } else {
        take = false;  // includes NaN: partial_cmp → None → Less → incumbent kept
    }
    if (take) {
        best.dist = dist;
        best.pid = pid;
        best.p = p;
    }
}

template <bool COUNT>
__device__ __forceinline__ void test_sphere(const DevScene& sc, const float4 s, int pid, V3 o, V3 d, Hit& best,
                                            Ctr& ctr) {
    // oc = origin - center is a single exact subtraction: shared by filter and exact path
    V3 oc = mk(x_sub(o.x, s.x), x_sub(o.y, s.y), x_sub(o.z, s.z));
    // FILTER: reference discriminant is 4*(bh*bh - (|oc|^2 - r^2)), bh = d.oc.  Evaluate it with FMAs
    // and reject only when it is negative by more than a generous rounding bound (~300 ulp of |oc|^2).
    float bh = fmaf(oc.z, d.z, fmaf(oc.y, d.y, oc.x * d.x));
    float oc2 = fmaf(oc.z, oc.z, fmaf(oc.y, oc.y, oc.x * oc.x));
    float cf = oc2 - s.w;
    float disc = fmaf(bh, bh, -cf);
    if (COUNT) ctr.v[CTR_SPH_TEST]++;
    if (fmaf(oc2, 2e-5f, disc) < 0.0f) return;
    // both roots behind the origin (ray points away, origin outside): cannot be in [T_MIN, T_MAX)
    if (bh > 0.0f && cf > 1e-4f * oc2) return;
    if (COUNT) ctr.v[CTR_SPH_EXACT]++;
    float t;
    if (!sphere_root_exact(d, oc, s.w, &t)) return;
    if (COUNT) ctr.v[CTR_SPH_HIT]++;
    consider(sc, o, d, t, pid, best);
}

template <bool COUNT>
__device__ __forceinline__ void test_triangle(const DevScene& sc, const float4* tri, int tidx, int pid, V3 o, V3 d,
                                              Hit& best, Ctr& ctr) {
    V3 a = ld3(tri[4 * tidx + 0]);
    V3 ab = ld3(tri[4 * tidx + 1]);
    V3 ac = ld3(tri[4 * tidx + 2]);
    if (COUNT) ctr.v[CTR_TRI_TEST]++;
    float t;
    int stage;
    bool hit = triangle_root_exact(o, d, a, ab, ac, &t, &stage);
    if (COUNT) {
        if (stage >= 1) ctr.v[CTR_TRI_S1]++;
        if (stage >= 2) ctr.v[CTR_TRI_S2]++;
        if (stage >= 3) ctr.v[CTR_TRI_S3]++;
        if (hit) ctr.v[CTR_TRI_HIT]++;
    }
    if (!hit) return;
    consider(sc, o, d, t, pid, best);
}

// ---- the same tests split into a cheap FILTER stage and an EXACT stage (scheduled kernel) ----
__device__ __forceinline__ bool sphere_filter(const float4 s, V3 o, V3 d) {
    const float ocx = x_sub(o.x, s.x), ocy = x_sub(o.y, s.y), ocz = x_sub(o.z, s.z);
    const float bh = fmaf(ocz, d.z, fmaf(ocy, d.y, ocx * d.x));
    const float oc2 = fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx));
    const float cf = oc2 - s.w;
    const float disc = fmaf(bh, bh, -cf);
    if (fmaf(oc2, 2e-5f, disc) < 0.0f) return false;
    if (bh > 0.0f && cf > 1e-4f * oc2) return false;
    return true;
}

template <bool COUNT>
__device__ __forceinline__ void sphere_exact(const DevScene& sc, const float4 s, int pid, V3 o, V3 d, Hit& best,
                                             Ctr& ctr) {
    const V3 oc = mk(x_sub(o.x, s.x), x_sub(o.y, s.y), x_sub(o.z, s.z));
    if (COUNT) ctr.v[CTR_SPH_EXACT]++;
    float t;
    if (!sphere_root_exact(d, oc, s.w, &t)) return;
    if (COUNT) ctr.v[CTR_SPH_HIT]++;
    consider(sc, o, d, t, pid, best);
}

// FMA Moeller-Trumbore with error-scaled margins: false only when the exact test (mesh.rs:109-161) must
// reject, or when the hit would be farther than the current best by more than the tie margin.
__device__ __forceinline__ bool triangle_filter(const float4* tri, int tidx, V3 o, V3 d, float cull) {
    const V3 a = ld3(tri[4 * tidx + 0]), ab = ld3(tri[4 * tidx + 1]), ac = ld3(tri[4 * tidx + 2]);
    const float ux = fmaf(d.y, ac.z, -ac.y * d.z), uy = fmaf(d.z, ac.x, -ac.z * d.x), uz = fmaf(d.x, ac.y, -ac.x * d.y);
    const float det = fmaf(ab.z, uz, fmaf(ab.y, uy, ab.x * ux));
    const float sdet = fabsf(ab.x * ux) + fabsf(ab.y * uy) + fabsf(ab.z * uz);
    if (fabsf(det) < 1e-5f + 1e-4f * sdet) return true;  // near-parallel: let the exact test decide
    const float inv = __frcp_rn(det), ainv = fabsf(inv);
    const float aox = o.x - a.x, aoy = o.y - a.y, aoz = o.z - a.z;
    const float mag = fabsf(aox) + fabsf(aoy) + fabsf(aoz);
    const float mab = fabsf(ab.x) + fabsf(ab.y) + fabsf(ab.z), mac = fabsf(ac.x) + fabsf(ac.y) + fabsf(ac.z);
    const float u = fmaf(aoz, uz, fmaf(aoy, uy, aox * ux)) * inv;
    // 1e-4 = ~800 ulp on the products actually summed; the second term covers cancellation inside d x ac
    const float eu = (1e-4f * (fabsf(aox * ux) + fabsf(aoy * uy) + fabsf(aoz * uz)) + 2e-6f * mag * mac) * ainv + 1e-5f;
    if (u < -eu || u > 1.0f + eu) return false;
    const float vx = fmaf(aoy, ab.z, -ab.y * aoz), vy = fmaf(aoz, ab.x, -ab.z * aox), vz = fmaf(aox, ab.y, -ab.x * aoy);
    const float v = fmaf(d.z, vz, fmaf(d.y, vy, d.x * vx)) * inv;
    const float ev = 1e-4f * mag * mab * ainv + 1e-5f;  // |d| = 1
    if (v < -ev || u + v > 1.0f + eu + ev) return false;
    const float t = fmaf(ac.z, vz, fmaf(ac.y, vy, ac.x * vx)) * inv;
    const float et = 1e-4f * mag * mab * mac * ainv + 1e-6f;
    if (t < 0.0009f - et || t > cull + et) return false;  // exact needs t in [T_MIN, T_MAX) and a chance to win
    return true;
}

template <bool COUNT>
__device__ __forceinline__ void triangle_exact(const DevScene& sc, const float4* tri, int tidx, int pid, V3 o, V3 d,
                                               Hit& best, Ctr& ctr) {
    const V3 a = ld3(tri[4 * tidx + 0]), ab = ld3(tri[4 * tidx + 1]), ac = ld3(tri[4 * tidx + 2]);
    float t;
    int stage;
    const bool hit = triangle_root_exact(o, d, a, ab, ac, &t, &stage);
    if (COUNT) {
        if (stage >= 1) ctr.v[CTR_TRI_S1]++;
        if (stage >= 2) ctr.v[CTR_TRI_S2]++;
        if (stage >= 3) ctr.v[CTR_TRI_S3]++;
        if (hit) ctr.v[CTR_TRI_HIT]++;
    }
    if (!hit) return;
    consider(sc, o, d, t, pid, best);
}

// ---- FILTER-domain distance bounds (deferred-exact traversal) ------------------------------------------
// Each returns CL_MISS when the reference's exact test must reject the primitive, else an interval [lo, hi]
// that contains the reference's length(point - origin) IF the exact test accepts it.  CL_SURE additionally
// guarantees that the exact test accepts (roots well conditioned, t-range and the own-box slab test passed
// by margins that dominate every rounding error), so `hi` may be used to cull farther candidates.
enum { CL_MISS = 0, CL_MAYBE = 1, CL_SURE = 2 };

__device__ __forceinline__ int sphere_bounds(const float4 s, V3 o, V3 d, float eo, bool check_box, float* lo,
                                             float* hi) {
    const float ocx = x_sub(o.x, s.x), ocy = x_sub(o.y, s.y), ocz = x_sub(o.z, s.z);
    const float bh = fmaf(ocz, d.z, fmaf(ocy, d.y, ocx * d.x));
    const float oc2 = fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx));
    const float cf = oc2 - s.w;
    const float disc = fmaf(bh, bh, -cf);
    const float e_d = fmaf(oc2 + s.w, 2e-5f, 1e-30f);  // >> every rounding of the reference's b*b - 4*c (~300 ulp)
    if (disc < -e_d) return CL_MISS;
    if (bh > 0.0f && cf > 1e-4f * oc2) return CL_MISS;   // both roots behind the origin
    const float m1 = fabsf(ocx) + fabsf(ocy) + fabsf(ocz);
    *hi = 0.0f;
    if (disc < 64.0f * e_d) {  // grazing: roots ill-conditioned, let the exact arithmetic decide
        const float sqm = sqrtf(fmaxf(disc, 0.0f) + e_d);
        const float eb = 4e-6f * m1 + eo;
        if (-bh + sqm + eb < T_MIN) return CL_MISS;
        *lo = -bh - sqm - eb;
        return CL_MAYBE;
    }
    const float sq = sqrtf(disc);
    const float e_t = __fdividef(0.51f * e_d, sq) + 2e-6f * (m1 + sq);
    const float t0 = -bh - sq, t1 = -bh + sq;
    float t;
    if (t0 > T_MIN + e_t) {
        t = t0;
    } else if (t0 < T_MIN - e_t) {
        if (t1 < T_MIN - e_t) return CL_MISS;
        if (t1 <= T_MIN + e_t) {
            *lo = t1 - e_t - eo;
            return CL_MAYBE;
        }
        t = t1;
    } else {
        *lo = t0 - e_t - eo;
        return CL_MAYBE;
    }
    const float e = e_t + eo + 1e-6f * t;
    *lo = t - e;
    *hi = t + e;
    if (t > 999.0f) return t > 1001.0f ? CL_MISS : CL_MAYBE;
    if (check_box) {  // the reference's slab test on the sphere's own box: certain when the point is well inside
        const float r = sqrtf(s.w);
        const V3 p = mk(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z));
        if (!robustly_inside(p, t, mk(s.x - r, s.y - r, s.z - r), mk(s.x + r, s.y + r, s.z + r), e)) return CL_MAYBE;
    }
    return CL_SURE;
}

__device__ __forceinline__ int triangle_bounds(const float4* tri, int tidx, V3 o, V3 d, float eo, bool check_box,
                                               float* lo, float* hi) {
    const V3 a = ld3(tri[4 * tidx + 0]), ab = ld3(tri[4 * tidx + 1]), ac = ld3(tri[4 * tidx + 2]);
    const float ux = fmaf(d.y, ac.z, -ac.y * d.z), uy = fmaf(d.z, ac.x, -ac.z * d.x), uz = fmaf(d.x, ac.y, -ac.x * d.y);
    const float det = fmaf(ab.z, uz, fmaf(ab.y, uy, ab.x * ux));
    const float sdet = fabsf(ab.x * ux) + fabsf(ab.y * uy) + fabsf(ab.z * uz);
    *hi = 0.0f;
    if (fabsf(det) < 1e-5f + 1e-4f * sdet) {  // near-parallel: only the exact test can tell
        *lo = 0.0f;
        return CL_MAYBE;
    }
    const float inv = __frcp_rn(det), ainv = fabsf(inv);
    const float aox = o.x - a.x, aoy = o.y - a.y, aoz = o.z - a.z;
    const float mag = fabsf(aox) + fabsf(aoy) + fabsf(aoz);
    const float mab = fabsf(ab.x) + fabsf(ab.y) + fabsf(ab.z), mac = fabsf(ac.x) + fabsf(ac.y) + fabsf(ac.z);
    const float u = fmaf(aoz, uz, fmaf(aoy, uy, aox * ux)) * inv;
    // 1e-4 = ~800 ulp on the products actually summed; the second term covers cancellation inside d x ac
    const float eu = (1e-4f * (fabsf(aox * ux) + fabsf(aoy * uy) + fabsf(aoz * uz)) + 2e-6f * mag * mac) * ainv + 1e-5f;
    if (u < -eu || u > 1.0f + eu) return CL_MISS;
    const float vx = fmaf(aoy, ab.z, -ab.y * aoz), vy = fmaf(aoz, ab.x, -ab.z * aox), vz = fmaf(aox, ab.y, -ab.x * aoy);
    const float v = fmaf(d.z, vz, fmaf(d.y, vy, d.x * vx)) * inv;
    const float ev = 1e-4f * mag * mab * ainv + 1e-5f;  // |d| = 1
    if (v < -ev || u + v > 1.0f + eu + ev) return CL_MISS;
    const float t = fmaf(ac.z, vz, fmaf(ac.y, vy, ac.x * vx)) * inv;
    const float et = 1e-4f * mag * mab * mac * ainv + 1e-6f;
    if (t < T_MIN - et || t > 1001.0f + et) return CL_MISS;  // EPSILON = 1e-5 < T_MIN
    const float e = et + eo + 1e-6f * fabsf(t);
    *lo = t - e;
    *hi = t + e;
    const bool inside = (u >= eu) && (u <= 1.0f - eu) && (v >= ev) && (u + v <= 1.0f - eu - ev);
    if (!inside || t <= T_MIN + et || t > 999.0f) return CL_MAYBE;
    if (check_box) {
        const V3 b = mk(a.x + ab.x, a.y + ab.y, a.z + ab.z), c = mk(a.x + ac.x, a.y + ac.y, a.z + ac.z);
        const V3 blo = mk(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
        const V3 bhi = mk(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
        const V3 p = mk(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z));
        if (!robustly_inside(p, t, blo, bhi, e)) return CL_MAYBE;
    }
    return CL_SURE;
}

// ---------------------------------------------------------------------------------------------
// K1: brute force
// ---------------------------------------------------------------------------------------------
// The hot loop is the FILTER alone, 11 FMA-pipe instructions + compare + branch per sphere:
//   oc = o - c (3 FADD, the reference's own first operation, so the exact path reuses it)
//   bh = d.oc, oc2 = oc.oc (2 FMUL + 4 FFMA);  m = bh*bh + (r^2 - 0.99998*oc2) (2 FFMA)
// m >= 0  <=>  the reference's discriminant 4*(bh^2 - (|oc|^2 - r^2)) is above -8e-5*|oc|^2 (~300 ulp of slack).
// Everything else (roots behind the origin, exact roots, slab check, min_by) runs only for the few spheres that pass.
template <bool COUNT>
__device__ __noinline__ void brute_sphere_slow(const DevScene& sc, const float4 s, int pid, V3 o, V3 d, Hit& best,
                                               Ctr& ctr) {
    const V3 oc = mk(x_sub(o.x, s.x), x_sub(o.y, s.y), x_sub(o.z, s.z));
    const float bh = fmaf(oc.z, d.z, fmaf(oc.y, d.y, oc.x * d.x));
    const float oc2 = fmaf(oc.z, oc.z, fmaf(oc.y, oc.y, oc.x * oc.x));
    if (bh > 0.0f && (oc2 - s.w) > 1e-4f * oc2) return;  // both roots behind the origin: not in [T_MIN, T_MAX)
    if (COUNT) ctr.v[CTR_SPH_EXACT]++;
    float t;
    if (!sphere_root_exact(d, oc, s.w, &t)) return;
    if (COUNT) ctr.v[CTR_SPH_HIT]++;
    consider(sc, o, d, t, pid, best);
}

__device__ __forceinline__ float brute_margin(const float4 s, V3 o, V3 d) {
    const float ocx = x_sub(o.x, s.x), ocy = x_sub(o.y, s.y), ocz = x_sub(o.z, s.z);
    const float bh = fmaf(ocz, d.z, fmaf(ocy, d.y, ocx * d.x));
    const float oc2 = fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx));
    return fmaf(bh, bh, fmaf(oc2, -0.99998f, s.w));
}

// a group of 8 spheres (first is a multiple of 8) in which at least one passed the filter: the four pairs once more
// in packed arithmetic, this time keeping WHICH spheres passed, then the slow path for exactly those
template <bool COUNT>
__device__ __noinline__ void brute_group_slow(const DevScene& sc, const float4* sph, const float4* sph2, int first, int count,
                                              V3 o, V3 d, bool all, Hit& best, Ctr& ctr) {
    const f32x2 ox2 = pk2(o.x, o.x), oy2 = pk2(o.y, o.y), oz2 = pk2(o.z, o.z);
    const f32x2 dx2 = pk2(d.x, d.x), dy2 = pk2(d.y, d.y), dz2 = pk2(d.z, d.z);
    const f32x2 kk2 = pk2(-0.99998f, -0.99998f);
    unsigned mask = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float4 A = sph2[2 * ((first >> 1) + k)], B = sph2[2 * ((first >> 1) + k) + 1];
        const f32x2 ocx = add2(ox2, pk2(A.x, A.y)), ocy = add2(oy2, pk2(A.z, A.w)), ocz = add2(oz2, pk2(B.x, B.y));
        const f32x2 bh = fma2(ocz, dz2, fma2(ocy, dy2, mul2(ocx, dx2)));
        const f32x2 oc2 = fma2(ocz, ocz, fma2(ocy, ocy, mul2(ocx, ocx)));
        float m_lo, m_hi;
        upk2(fma2(bh, bh, fma2(oc2, kk2, pk2(B.z, B.w))), m_lo, m_hi);
        mask |= (!(m_lo < 0.0f) ? 1u : 0u) << (2 * k);
        mask |= (!(m_hi < 0.0f) ? 1u : 0u) << (2 * k + 1);
    }
    if (all) mask = 0xffu;  // a non-finite ray: every sphere goes through the exact arithmetic
    mask &= (1u << count) - 1u;
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1u;
        brute_sphere_slow<COUNT>(sc, sph[first + k], first + k, o, d, best, ctr);
    }
}

template <bool COUNT>
__device__ __forceinline__ void trace_brute(const DevScene& sc, const SceneView& sv, V3 o, V3 d, Hit& best, Ctr& ctr) {
    best.pid = -1;
    best.dist = 0.0f;
    const int ns = (int)sc.ns;
    // a non-finite ray makes the margins NaN, which fmaxf would drop: send such a ray through the slow path whole
    const bool weird = !(isfinite(o.x) && isfinite(o.y) && isfinite(o.z) && isfinite(d.x) && isfinite(d.y) && isfinite(d.z));
    // Packed pairs: two spheres per instruction — 3 FADD2 + 2 FMUL2 + 6 FFMA2 per PAIR (5.5 FMA-pipe issue slots per
    // sphere instead of 11), two LDS.128 per pair, one FMNMX3 per pair, one branch per 16 spheres.  The FMA pipe still
    // does 11 lane-operations per sphere, so the loop is bound by the pipe, not by issue (profiles/r1_notes.md).
    const f32x2 ox2 = pk2(o.x, o.x), oy2 = pk2(o.y, o.y), oz2 = pk2(o.z, o.z);
    const f32x2 dx2 = pk2(d.x, d.x), dy2 = pk2(d.y, d.y), dz2 = pk2(d.z, d.z);
    const f32x2 kk2 = pk2(-0.99998f, -0.99998f);
    auto pair_margin = [&](int j, float m) {
        const float4 A = sv.sph2[2 * j], B = sv.sph2[2 * j + 1];
        const f32x2 ocx = add2(ox2, pk2(A.x, A.y)), ocy = add2(oy2, pk2(A.z, A.w)), ocz = add2(oz2, pk2(B.x, B.y));
        const f32x2 bh = fma2(ocz, dz2, fma2(ocy, dy2, mul2(ocx, dx2)));
        const f32x2 oc2 = fma2(ocz, ocz, fma2(ocy, ocy, mul2(ocx, ocx)));
        const f32x2 mm = fma2(bh, bh, fma2(oc2, kk2, pk2(B.z, B.w)));
        float m_lo, m_hi;
        upk2(mm, m_lo, m_hi);
        return fmaxf(fmaxf(m, m_lo), m_hi);
    };
    const int ns8 = (ns + 7) & ~7;
    const float NEG = -3.0e38f;
    int i = 0;
    for (; i + 16 <= ns8; i += 16) {
        float m0 = NEG, m1 = NEG;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            m0 = pair_margin((i >> 1) + k, m0);
            m1 = pair_margin((i >> 1) + 4 + k, m1);
        }
        if (COUNT) ctr.v[CTR_SPH_TEST] += min(16, ns - i);
        if (!(fmaxf(m0, m1) < 0.0f) || weird) {
            if (!(m0 < 0.0f) || weird) brute_group_slow<COUNT>(sc, sv.sph, sv.sph2, i, min(8, ns - i), o, d, weird, best, ctr);
            if ((!(m1 < 0.0f) || weird) && i + 8 < ns) brute_group_slow<COUNT>(sc, sv.sph, sv.sph2, i + 8, min(8, ns - i - 8), o, d, weird, best, ctr);
        }
    }
    if (i < ns8) {
        float m0 = NEG;
#pragma unroll
        for (int k = 0; k < 4; k++) m0 = pair_margin((i >> 1) + k, m0);
        if (COUNT) ctr.v[CTR_SPH_TEST] += ns - i;
        if (!(m0 < 0.0f) || weird) brute_group_slow<COUNT>(sc, sv.sph, sv.sph2, i, ns - i, o, d, weird, best, ctr);
    }
    const int nt = (int)sc.nt;
    for (int j = 0; j < nt; j++) test_triangle<COUNT>(sc, sv.tri, j, ns + j, o, d, best, ctr);
}

// ---------------------------------------------------------------------------------------------
// K2: BVH traversal.  FILTER-domain slab test; returns entry distance, hit flag.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool slab(float lx, float ly, float lz, float hx, float hy, float hz, float ix, float iy,
                                     float iz, float ox, float oy, float oz, float tmax, float* tnear) {
    float x0 = fmaf(lx, ix, ox), x1 = fmaf(hx, ix, ox);
    float y0 = fmaf(ly, iy, oy), y1 = fmaf(hy, iy, oy);
    float z0 = fmaf(lz, iz, oz), z1 = fmaf(hz, iz, oz);
    // fminf/fmaxf drop NaN operands (0*inf slabs): such an axis does not constrain → conservative
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    *tnear = tn;
    return tn <= tf * 1.000002f;
}

template <bool COUNT>
__device__ __forceinline__ void trace_bvh(const DevScene& sc, const SceneView& sv, V3 o, V3 d, Hit& best, Ctr& ctr) {
    best.pid = -1;
    best.dist = 0.0f;
    const float ix = __frcp_rn(d.x), iy = __frcp_rn(d.y), iz = __frcp_rn(d.z);
    const float ox = -o.x * ix, oy = -o.y * iy, oz = -o.z * iz;
    float cull = 1001.0f;  // a hit has t < T_MAX and length(p-o) ~ t
    int stack[MAX_STACK];
    int sp = 0;
    int cur = sc.root;
    const int ns = (int)sc.ns;
    for (;;) {
        while (cur >= 0) {
            const float4 a = sv.na[cur], b = sv.nb[cur], c = sv.nc[cur];
            const int2 ch = sv.nd[cur];
            float tl, tr;
            bool hl = slab(a.x, a.y, a.z, a.w, b.x, b.y, ix, iy, iz, ox, oy, oz, cull, &tl);
            bool hr = slab(b.z, b.w, c.x, c.y, c.z, c.w, ix, iy, iz, ox, oy, oz, cull, &tr);
            if (COUNT) ctr.v[CTR_SLAB] += 2;
            if (hl && hr) {
                int nearc = ch.x, farc = ch.y;
                if (tr < tl) {
                    nearc = ch.y;
                    farc = ch.x;
                }
                stack[sp++] = farc;
                cur = nearc;
            } else if (hl) {
                cur = ch.x;
            } else if (hr) {
                cur = ch.y;
            } else {
                if (sp == 0) return;
                cur = stack[--sp];
            }
        }
        // leaf
        int pid = ~cur;
        if (pid < ns) {
            test_sphere<COUNT>(sc, sv.sph[pid], pid, o, d, best, ctr);
        } else {
            test_triangle<COUNT>(sc, sv.tri, pid - ns, pid, o, d, best, ctr);
        }
        if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
        if (sp == 0) return;
        cur = stack[--sp];
    }
}

// ---------------------------------------------------------------------------------------------
// Camera::get_ray (camera.rs:109-129) — EXACT
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void primary_ray(const DevCamera& cam, uint32_t x, uint32_t y_cam, Rng& rng, V3* o_out,
                                            V3* d_out) {
    float a, b;
    unit_disc(rng, a, b);
    V3 offset = mk(x_mul(a, cam.lens_radius), x_mul(b, cam.lens_radius), 0.0f);
    float u = x_div(x_add((float)x, rng.gen_range_0_1()), cam.u_den);
    float v = x_div(x_add((float)y_cam, rng.gen_range_0_1()), cam.v_den);
    V3 org = mk(cam.org[0], cam.org[1], cam.org[2]);
    V3 llc = mk(cam.llc[0], cam.llc[1], cam.llc[2]);
    V3 hor = mk(cam.hor[0], cam.hor[1], cam.hor[2]);
    V3 ver = mk(cam.ver[0], cam.ver[1], cam.ver[2]);
    // lower_left_corner + u*horizontal + v*vertical - origin
    V3 target = x_sub(x_add(x_add(llc, x_scale(hor, u)), x_scale(ver, v)), org);
    // Ray::new(origin, normalize_or_zero(target)).at(focus_distance); Ray::new renormalises by division
    V3 d1 = x_normalize_div(x_normalize_or_zero(target));
    V3 focal_point = x_add(org, x_scale(d1, cam.focus));
    V3 fo = x_add(org, offset);
    *o_out = fo;
    *d_out = x_normalize_div(x_normalize_or_zero(x_sub(focal_point, fo)));
}

// `(c * 255.999) as u8`: truncation, saturation, NaN → 0 (color.rs:13-19)
__device__ __forceinline__ uint32_t quantise(float sum, float spp_f) {
    float c = x_sqrt(x_div(sum, spp_f));
    float s = x_mul(c, 255.999f);
    uint32_t q = __float2uint_rz(s);  // saturating, NaN → 0
    return q > 255u ? 255u : q;
}

// ---------------------------------------------------------------------------------------------
// The megakernel
// ---------------------------------------------------------------------------------------------
template <int ISECT, bool SMEM, bool COUNT>
__global__ void __launch_bounds__(THREADS) render_kernel(const DevScene sc, const DevCamera cam, const DevParams pr) {
    extern __shared__ float4 smem_dyn[];
    __shared__ __align__(16) uint8_t stage[WARPS][TILE_W * TILE_H * 3];

    SceneView sv;
    if (SMEM) {
        // layout: sph | tri | node_a | node_b | node_c | node_d
        float4* p = smem_dyn;
        float4* s_sph = p;  p += sc.ns;
        float4* s_tri = p;  p += 4 * sc.nt;
        float4* s_sph2 = p;  // brute force only: the pair-packed copy of the spheres
        if (ISECT == RT_INTERSECT_BRUTE) {
            const uint32_t ns8 = (sc.ns + 7u) & ~7u;
            for (uint32_t i = threadIdx.x; i < ns8; i += THREADS) s_sph2[i] = __ldg(&sc.sph2[i]);
        }
        sv.sph2 = s_sph2;
        float4* s_na = p;   p += sc.ni;
        float4* s_nb = p;   p += sc.ni;
        float4* s_nc = p;   p += sc.ni;
        int2* s_nd = reinterpret_cast<int2*>(p);
        for (uint32_t i = threadIdx.x; i < sc.ns; i += THREADS) s_sph[i] = __ldg(&sc.sph[i]);
        for (uint32_t i = threadIdx.x; i < 4 * sc.nt; i += THREADS) s_tri[i] = __ldg(&sc.tri[i]);
        if (ISECT == RT_INTERSECT_BVH) {
            for (uint32_t i = threadIdx.x; i < sc.ni; i += THREADS) {
                s_na[i] = __ldg(&sc.node_a[i]);
                s_nb[i] = __ldg(&sc.node_b[i]);
                s_nc[i] = __ldg(&sc.node_c[i]);
                s_nd[i] = __ldg(&sc.node_d[i]);
            }
        }
        __syncthreads();
        sv.sph = s_sph; sv.tri = s_tri; sv.na = s_na; sv.nb = s_nb; sv.nc = s_nc; sv.nd = s_nd;
    } else {
        sv.sph2 = sc.sph2;
        sv.sph = sc.sph; sv.tri = sc.tri; sv.na = sc.node_a; sv.nb = sc.node_b; sv.nc = sc.node_c; sv.nd = sc.node_d;
    }

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t lx = lane & (TILE_W - 1), ly = lane >> 3;
    const uint32_t total_tiles = pr.tiles_x * pr.tiles_y;
    const float spp_f = (float)pr.spp;

    Ctr ctr;
#pragma unroll
    for (int i = 0; i < NUM_COUNTERS; i++) ctr.v[i] = 0;
    unsigned long long rays = 0;

    for (;;) {
        // ---- fetch the next tile of this rank: ticket k → group k of `tile_ranks` tiles, rotated ----
        unsigned int k = 0;
        if (lane == 0) k = atomicAdd(pr.tile_counter, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        uint64_t g = (uint64_t)k * pr.tile_ranks + (pr.tile_rank + k) % pr.tile_ranks;
        if (g >= total_tiles) break;
        const uint32_t tx = (uint32_t)(g % pr.tiles_x), ty = (uint32_t)(g / pr.tiles_x);
        const uint32_t x = tx * TILE_W + lx;
        const uint32_t y = pr.row0 + ty * TILE_H + ly;  // global image row (0 = top)
        const bool valid = x < pr.width && y < pr.row1;

        float sr = 0.0f, sg = 0.0f, sb = 0.0f;
        if (valid) {
            Rng rng;
            rng.seed_from_u64(pr.seed + ((uint64_t)y * pr.width + x));
            const uint32_t y_cam = pr.height - y - 1;  // main.rs:71
            uint32_t path[MAX_PATH];                   // pids of the non-terminal hits of the current sample
            uint32_t s = 0, left = 0, np = 0;
            V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
            while (s < pr.spp) {
                if (left == 0) {  // start sample s
                    primary_ray(cam, x, y_cam, rng, &o, &d);
                    left = pr.depth;
                    np = 0;
                }
                // ---- one nearest-hit query (ray_color with depth > 0) ----
                rays++;
                if (COUNT) {
                    ctr.v[CTR_ACTIVE_LANES]++;
                    unsigned am = __activemask();
                    if (lane == (__ffs(am) - 1)) ctr.v[CTR_TOTAL_LANES] += 32;
                }
                Hit h;
                if (ISECT == RT_INTERSECT_BRUTE) trace_brute<COUNT>(sc, sv, o, d, h, ctr);
                else trace_bvh<COUNT>(sc, sv, o, d, h, ctr);

                bool done;
                float Lr, Lg, Lb;
                if (h.pid >= 0) {
                    const float e = __ldg(&sc.emis[h.pid]);
                    const float4 m = __ldg(&sc.mat[h.pid]);
                    if (e > 0.0f) {  // emission * albedo (main.rs:116-117)
                        Lr = x_mul(m.x, e); Lg = x_mul(m.y, e); Lb = x_mul(m.z, e);
                        done = true;
                        if (COUNT) ctr.v[CTR_EMISSIVE]++;
                    } else {
                        V3 n;
                        if (COUNT) ctr.v[h.pid < (int)sc.ns ? CTR_SHADE_SPH : CTR_SHADE_TRI]++;
                        if (h.pid < (int)sc.ns) {
                            const float4 sp4 = sv.sph[h.pid];
                            n = x_normalize_or_zero(x_sub(h.p, ld3(sp4)));  // sphere.rs:49-51
                        } else {
                            n = ld3(sv.tri[4 * (h.pid - (int)sc.ns) + 3]);  // mesh.rs:163-165 (host, same ops)
                        }
                        V3 diffuse = x_add(unit_sphere(rng), n);
                        float kk = x_mul(2.0f, x_dot(d, n));
                        V3 glossy = x_sub(d, x_scale(n, kk));
                        V3 scat = x_add(diffuse, x_scale(x_sub(glossy, diffuse), m.w));
                        V3 nd;
                        if (!x_try_normalize(scat, &nd)) nd = n;
                        o = h.p;
                        d = x_normalize_div(nd);  // Ray::new
                        path[np++] = (uint32_t)h.pid;
                        left--;
                        done = (left == 0);       // next call has depth == 0 → BLACK, no query
                        Lr = Lg = Lb = 0.0f;
                    }
                } else {  // sky (main.rs:135-144)
                    if (COUNT) ctr.v[CTR_SKY]++;
                    float rcp = x_div(1.0f, x_length(d));
                    float ny = (isfinite(rcp) && rcp > 0.0f) ? x_mul(d.y, rcp) : 0.0f;
                    float t = x_add(x_mul(ny, 0.5f), 1.0f);
                    float k1 = x_sub(1.0f, t);
                    float w = x_mul(1.0f, t);
                    Lr = x_add(w, x_mul(0.3f, k1));
                    Lg = Lr;
                    Lb = x_add(w, x_mul(0.8f, k1));
                    done = true;
                }
                if (done) {
                    // fold albedo ⊙ (albedo ⊙ (... ⊙ L)) innermost first, like the recursion unwinding
                    while (np > 0) {
                        const float4 m = __ldg(&sc.mat[path[--np]]);
                        Lr = x_mul(m.x, Lr); Lg = x_mul(m.y, Lg); Lb = x_mul(m.z, Lb);
                    }
                    sr = x_add(sr, Lr); sg = x_add(sg, Lg); sb = x_add(sb, Lb);
                    s++;
                    left = 0;
                }
            }
        }

        // ---- pixel finish + tile store: stage 96 B in smem, write three 8-byte vectors per row ----
        __syncwarp();
        uint8_t* st = stage[warp];
        st[lane * 3 + 0] = (uint8_t)quantise(sr, spp_f);
        st[lane * 3 + 1] = (uint8_t)quantise(sg, spp_f);
        st[lane * 3 + 2] = (uint8_t)quantise(sb, spp_f);
        __syncwarp();
        const uint32_t x0 = tx * TILE_W, y0 = pr.row0 + ty * TILE_H;
        const bool full = (x0 + TILE_W <= pr.width) && (y0 + TILE_H <= pr.row1) && ((pr.width & 7u) == 0) &&
                          ((reinterpret_cast<uintptr_t>(pr.out) & 7u) == 0);
        if (full) {
            if (lane < 12) {
                const uint32_t r = lane / 3, seg = lane % 3;
                const size_t off = ((size_t)(y0 + r - pr.out_row0) * pr.width + x0) * 3 + seg * 8;
                *reinterpret_cast<uint2*>(pr.out + off) = *reinterpret_cast<const uint2*>(st + r * 24 + seg * 8);
            }
        } else if (valid) {
            const size_t off = ((size_t)(y - pr.out_row0) * pr.width + x) * 3;
            pr.out[off + 0] = st[lane * 3 + 0];
            pr.out[off + 1] = st[lane * 3 + 1];
            pr.out[off + 2] = st[lane * 3 + 2];
        }
    }

    // ---- counters: warp-reduce, one atomic per warp per slot ----
    ctr.v[CTR_RAYS] = rays;
#pragma unroll
    for (int i = 0; i < NUM_COUNTERS; i++) {
        if (!COUNT && i != CTR_RAYS) continue;
        unsigned long long v = ctr.v[i];
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) v += __shfl_down_sync(0xffffffffu, v, ofs);
        if (lane == 0 && v) atomicAdd(&pr.counters[i], v);
    }
}


// ---------------------------------------------------------------------------------------------
// K2 traversal, second form: centre/half-extent slab test (9 FFMA + 4 FMNMX per box instead of 6 FFMA +
// 10 FMNMX: the first form saturated the ALU pipe at 75 % with the FMA pipe at 23 %), FMA pre-filter in
// front of the exact triangle test.  Same while-while structure, same results.
// ---------------------------------------------------------------------------------------------
constexpr int TR_DONE = (int)0x80000000;  // traversal finished (never a leaf code: first_pid < 2^26)
template <bool COUNT, bool WITH_BIG = true>
__device__ __forceinline__ void trace_bvh_ch(const DevScene& sc, const SceneView& sv, V3 o, V3 d, Hit& best, Ctr& ctr) {
    best.pid = -1;
    best.dist = 0.0f;
    // FILTER-domain ray constants; |1/d| is clamped so 0*inf never produces NaN slabs
    const float BIG = 1e30f;
    float ix = fminf(fmaxf(__frcp_rn(d.x), -BIG), BIG);
    float iy = fminf(fmaxf(__frcp_rn(d.y), -BIG), BIG);
    float iz = fminf(fmaxf(__frcp_rn(d.z), -BIG), BIG);
    if (!(fabsf(d.x) > 0.0f)) ix = BIG;
    if (!(fabsf(d.y) > 0.0f)) iy = BIG;
    if (!(fabsf(d.z) > 0.0f)) iz = BIG;
    const float ax = fabsf(ix), ay = fabsf(iy), az = fabsf(iz);
    const float qx = -o.x * ix, qy = -o.y * iy, qz = -o.z * iz;
    // rounding of the o-term: <= 3 * 2^-24 * |o*inv| per axis, in t (the c- and h-terms are padded on the host)
    const float slack = 4.8e-7f * fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + 1e-30f;
    float cull = 1001.0f;  // a hit has t < T_MAX and length(p-o) ~ t
    int stack[MAX_STACK + 1];
    stack[0] = TR_DONE;  // sentinel: popping it ends the traversal, so a pop needs no emptiness test
    int* top = stack + 1;  // next free entry
    int cur = sc.lroot;
    const int ns = (int)sc.ns;
    // split layout: the large primitives first (their hits shorten everything that follows)
    for (uint32_t i = 0; WITH_BIG && i < sc.nbig; i++) {
        const int pid = (int)sc.big_pid[i];
        if (pid < ns) {
            test_sphere<COUNT>(sc, sv.sph[pid], pid, o, d, best, ctr);
        } else {
            if (COUNT) ctr.v[CTR_TRI_TEST]++;
            if (triangle_filter(sv.tri, pid - ns, o, d, cull)) triangle_exact<COUNT>(sc, sv.tri, pid - ns, pid, o, d, best, ctr);
        }
        if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
    }
    if (!sc.ltree) return;
    for (;;) {
        while (cur >= 0) {
            const float4* nrec = sv.na + 3 * cur;
            const float4 a = nrec[0], b = nrec[1], c = nrec[2];
            const int2 ch = sv.nd[cur];
            // left box: c = (a.x,a.y,a.z) h = (a.w,b.x,b.y); right: c = (b.z,b.w,c.x) h = (c.y,c.z,c.w)
            const float lcx = fmaf(a.x, ix, qx), lcy = fmaf(a.y, iy, qy), lcz = fmaf(a.z, iz, qz);
            const float rcx = fmaf(b.z, ix, qx), rcy = fmaf(b.w, iy, qy), rcz = fmaf(c.x, iz, qz);
            const float tl = fmaxf(fmaxf(fmaf(-a.w, ax, lcx), fmaf(-b.x, ay, lcy)), fmaxf(fmaf(-b.y, az, lcz), 0.0f));
            const float fl = fminf(fminf(fmaf(a.w, ax, lcx), fmaf(b.x, ay, lcy)), fminf(fmaf(b.y, az, lcz), cull));
            const float tr = fmaxf(fmaxf(fmaf(-c.y, ax, rcx), fmaf(-c.z, ay, rcy)), fmaxf(fmaf(-c.w, az, rcz), 0.0f));
            const float fr = fminf(fminf(fmaf(c.y, ax, rcx), fmaf(c.z, ay, rcy)), fminf(fmaf(c.w, az, rcz), cull));
            const bool hl = tl <= fl + slack;
            const bool hr = tr <= fr + slack;
            if (COUNT) ctr.v[CTR_SLAB] += 2;
            // branch-light step: push and pop are short predicated blocks, the loop has one exit
            const bool swap = tr < tl;
            if (hl && hr) *top++ = swap ? ch.x : ch.y;
            int nxt = (hr && (!hl || swap)) ? ch.y : ch.x;  // the nearer (or the only) child
            if (!(hl || hr)) nxt = *--top;
            cur = nxt;
        }
        if (cur == TR_DONE) return;
        // leaf = contiguous pid range of one kind: code = ~((first << 5) | (count - 1))
        const int first = (~cur) >> 5, count = ((~cur) & 31) + 1;
        if (first < ns) {
            for (int i = 0; i < count; i++) test_sphere<COUNT>(sc, sv.sph[first + i], first + i, o, d, best, ctr);
        } else {
            for (int i = 0; i < count; i++) {
                const int pid = first + i;
                if (COUNT) ctr.v[CTR_TRI_TEST]++;
                if (triangle_filter(sv.tri, pid - ns, o, d, cull)) triangle_exact<COUNT>(sc, sv.tri, pid - ns, pid, o, d, best, ctr);
            }
        }
        if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
        cur = *--top;
        if (cur == TR_DONE) return;
    }
}


// ---------------------------------------------------------------------------------------------
// Primary-ray candidate lists.  Every camera ray of an 8x4 tile lies in a thin beam: it starts on the lens
// disc (radius aperture/2 around the camera origin, camera.rs:110-115) and passes through the focal point of a
// jittered pixel position (camera.rs:116-127), i.e. through a small patch of the sphere of radius
// focus_distance.  With X0(l) = org + l*f*d0 the beam's centre line (d0 = centre direction of the patch), any
// beam point satisfies |X - X0(l)| <= |1-l|*lr + l*Rp (Rp = patch radius).  A warp culls the scene against its
// tile's beam once and its primary rays test only the survivors — no traversal for 57 % of the rays of the
// BASELINE frames, and none at all for tiles that see only sky.  FILTER domain: the list is a conservative
// superset; hits are still decided by the exact tests and consider().
// ---------------------------------------------------------------------------------------------
constexpr int LIST_CAP = 62;       // 16-bit entries: count + 62 pids = 128 bytes per list (shared memory is tight on C3)
typedef unsigned short ListEntry;
constexpr int LIST_BAD = 0xffff;  // list[0] when the tile has too many candidates or the beam is unusable

struct Beam {
    V3 org, d0;
    float f, lr, rp, kappa;
    bool ok;
};

__device__ __forceinline__ V3 cam_dir(const DevCamera& cam, float u, float v) {
    const float tx = cam.llc[0] + u * cam.hor[0] + v * cam.ver[0] - cam.org[0];
    const float ty = cam.llc[1] + u * cam.hor[1] + v * cam.ver[1] - cam.org[1];
    const float tz = cam.llc[2] + u * cam.hor[2] + v * cam.ver[2] - cam.org[2];
    const float r = rsqrtf(tx * tx + ty * ty + tz * tz);
    return mk(tx * r, ty * r, tz * r);
}

__device__ __forceinline__ Beam tile_beam(const DevCamera& cam, const DevParams& pr, uint32_t x0, uint32_t y0) {
    Beam b;
    // pixel + jitter in [0,1): u spans [x0, x0+8]/u_den; rows y0..y0+3 ↔ y_cam = h-1-y, v spans [y_cam, y_cam+1]/v_den
    const float u0 = (float)x0 / cam.u_den, u1 = (float)(x0 + TILE_W) / cam.u_den;
    const float yc_hi = (float)(pr.height - y0), yc_lo = (float)(pr.height - y0) - (float)TILE_H;
    const float v0 = yc_lo / cam.v_den, v1 = yc_hi / cam.v_den;
    const V3 c00 = cam_dir(cam, u0, v0), c10 = cam_dir(cam, u1, v0), c01 = cam_dir(cam, u0, v1), c11 = cam_dir(cam, u1, v1);
    V3 m = mk(c00.x + c10.x + c01.x + c11.x, c00.y + c10.y + c01.y + c11.y, c00.z + c10.z + c01.z + c11.z);
    const float r = rsqrtf(m.x * m.x + m.y * m.y + m.z * m.z);
    b.d0 = mk(m.x * r, m.y * r, m.z * r);
    auto dist = [&](V3 c) {
        const float dx = c.x - b.d0.x, dy = c.y - b.d0.y, dz = c.z - b.d0.z;
        return sqrtf(dx * dx + dy * dy + dz * dz);
    };
    b.f = cam.focus;
    b.rp = fabsf(cam.focus) * (fmaxf(fmaxf(dist(c00), dist(c10)), fmaxf(dist(c01), dist(c11))) * 1.05f + 1e-6f);
    b.lr = fabsf(cam.lens_radius) * 1.01f + 1e-7f;
    b.org = mk(cam.org[0], cam.org[1], cam.org[2]);
    b.kappa = (b.lr + b.rp) / fabsf(cam.focus);
    // the bound needs a finite forward beam; anything odd (NaN, focus <= 0, aperture comparable to focus) → no list
    b.ok = (cam.focus > 0.0f) && (b.kappa < 0.45f) && (r > 0.0f) && (r < 1e30f);
    return b;
}

// can a ball (centre c, radius rad) be touched by any ray of the beam?  (conservative)
__device__ __forceinline__ bool beam_touches(const Beam& b, float cx, float cy, float cz, float rad) {
    const float wx = cx - b.org.x, wy = cy - b.org.y, wz = cz - b.org.z;
    const float sc = wx * b.d0.x + wy * b.d0.y + wz * b.d0.z;
    const float w2 = wx * wx + wy * wy + wz * wz;
    const float rho = sqrtf(fmaxf(w2 - sc * sc, 0.0f));
    const float scp = fmaxf(sc, 0.0f);
    const float lc = scp / b.f;
    const float ext = (rad + fabsf(1.0f - lc) * b.lr + lc * b.rp) * 2.0f + 1e-4f * (1.0f + sqrtf(w2));
    if (sc + ext < 0.0f) return false;  // entirely behind the lens
    const float l1 = fmaxf(scp - ext, 0.0f) / b.f, l2 = (scp + ext) / b.f;
    const float rmax = fmaxf(fabsf(1.0f - l1) * b.lr + l1 * b.rp, fabsf(1.0f - l2) * b.lr + l2 * b.rp);
    return !(rho > rad + rmax + 1e-4f * (1.0f + sqrtf(w2)));  // NaN → keep
}

// Warp-cooperative: fill list[0] = count (or -1: too many / unusable beam), list[1..] = pids.
__device__ __forceinline__ void build_tile_list(const DevScene& sc, const SceneView& sv, const DevCamera& cam,
                                                const DevParams& pr, uint32_t x0, uint32_t y0, ListEntry* list) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const Beam b = tile_beam(cam, pr, x0, y0);
    int count = b.ok ? 0 : -1;
    const int n = (int)(sc.ns + sc.nt), ns = (int)sc.ns;
    for (int base = 0; base < n && count >= 0; base += 32) {
        const int pid = base + lane;
        bool keep = false;
        if (pid < ns) {
            const float4 s = sv.sph[pid];
            keep = beam_touches(b, s.x, s.y, s.z, sqrtf(s.w) * 1.0001f + 1e-6f);
        } else if (pid < n) {
            const int t = pid - ns;
            const V3 a = ld3(sv.tri[4 * t]), ab = ld3(sv.tri[4 * t + 1]), ac = ld3(sv.tri[4 * t + 2]);
            // ball around the vertex centroid
            const float gx = a.x + (ab.x + ac.x) * (1.0f / 3.0f), gy = a.y + (ab.y + ac.y) * (1.0f / 3.0f),
                        gz = a.z + (ab.z + ac.z) * (1.0f / 3.0f);
            auto d2 = [&](float px, float py, float pz) { return (px - gx) * (px - gx) + (py - gy) * (py - gy) + (pz - gz) * (pz - gz); };
            const float r2 = fmaxf(d2(a.x, a.y, a.z), fmaxf(d2(a.x + ab.x, a.y + ab.y, a.z + ab.z), d2(a.x + ac.x, a.y + ac.y, a.z + ac.z)));
            keep = beam_touches(b, gx, gy, gz, sqrtf(r2) * 1.0001f + 1e-6f);
        }
        const unsigned m = __ballot_sync(FULL, keep);
        const int add = __popc(m);
        if (count + add > LIST_CAP) {
            count = -1;
        } else {
            if (keep) list[1 + count + __popc(m & ((1u << lane) - 1u))] = (ListEntry)pid;
            count += add;
        }
    }
    if (lane == 0) list[0] = count < 0 ? (ListEntry)LIST_BAD : (ListEntry)count;
    __syncwarp();
}

template <bool COUNT>
__device__ __forceinline__ void trace_list(const DevScene& sc, const SceneView& sv, const ListEntry* list, V3 o, V3 d,
                                           Hit& best, Ctr& ctr) {
    best.pid = -1;
    best.dist = 0.0f;
    float cull = 1001.0f;
    const int cnt = list[0], ns = (int)sc.ns;
    for (int i = 0; i < cnt; i++) {
        const int pid = list[1 + i];
        if (pid < ns) {
            test_sphere<COUNT>(sc, sv.sph[pid], pid, o, d, best, ctr);
        } else {
            if (COUNT) ctr.v[CTR_TRI_TEST]++;
            if (triangle_filter(sv.tri, pid - ns, o, d, cull)) triangle_exact<COUNT>(sc, sv.tri, pid - ns, pid, o, d, best, ctr);
        }
        if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
    }
}

// ---------------------------------------------------------------------------------------------
// The megakernel, second form: lanes are independent workers.  A lane that finishes its pixel takes the
// next pixel of the warp's current 8x4 tile (the warp pulls tiles from the global ticket counter) instead of
// idling until the slowest pixel of the tile is done; every trip of the loop is one nearest-hit query.
// ---------------------------------------------------------------------------------------------
template <int ISECT, bool SMEM, bool COUNT, bool LISTS, int MINB = 3, int TPB = THREADS>
__global__ void __launch_bounds__(TPB, MINB) render_kernel_lanes(const DevScene sc, const DevCamera cam,
                                                                const DevParams pr) {
    extern __shared__ float4 smem_dyn[];
    // per warp: candidate lists of its two most recent tiles (LISTS variant only; measured slower)
    __shared__ ListEntry s_lists[LISTS ? TPB / 32 : 1][2][LISTS ? LIST_CAP + 2 : 1];
    SceneView sv;
    if (SMEM) {
        float4* p = smem_dyn;
        float4* s_sph = p;  p += sc.ns;
        float4* s_tri = p;  p += 4 * sc.nt;
        float4* s_sph2 = p;  // brute force only: the pair-packed copy of the spheres
        if (ISECT == RT_INTERSECT_BRUTE) {
            const uint32_t ns8 = (sc.ns + 7u) & ~7u;
            for (uint32_t i = threadIdx.x; i < ns8; i += TPB) s_sph2[i] = __ldg(&sc.sph2[i]);
        }
        sv.sph2 = s_sph2;
        float4* s_na = p;   p += 3 * sc.ni;  // 48-byte node records
        int2* s_nd = reinterpret_cast<int2*>(p);
        for (uint32_t i = threadIdx.x; i < sc.ns; i += TPB) s_sph[i] = __ldg(&sc.sph[i]);
        for (uint32_t i = threadIdx.x; i < 4 * sc.nt; i += TPB) s_tri[i] = __ldg(&sc.tri[i]);
        if (ISECT == RT_INTERSECT_BVH) {
            for (uint32_t i = threadIdx.x; i < sc.lni; i += TPB) {
                s_na[3 * i] = __ldg(&sc.lnode_a[3 * i]);
                s_na[3 * i + 1] = __ldg(&sc.lnode_a[3 * i + 1]);
                s_na[3 * i + 2] = __ldg(&sc.lnode_a[3 * i + 2]);
                s_nd[i] = __ldg(&sc.lnode_d[i]);
            }
        }
        __syncthreads();
        sv.sph = s_sph; sv.tri = s_tri; sv.na = s_na; sv.nb = nullptr; sv.nc = nullptr; sv.nd = s_nd;
    } else {
        sv.sph2 = sc.sph2;
        sv.sph = sc.sph; sv.tri = sc.tri; sv.na = sc.lnode_a; sv.nb = nullptr; sv.nc = nullptr; sv.nd = sc.lnode_d;
    }

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t total_tiles = pr.tiles_x * pr.tiles_y;
    const float spp_f = (float)pr.spp;

    Ctr ctr;
#pragma unroll
    for (int i = 0; i < NUM_COUNTERS; i++) ctr.v[i] = 0;
    unsigned long long rays = 0;

    bool have_px = false, finished = false;
    uint32_t px = 0, py = 0, s = 0, left = 0, np = 0;
    float sr = 0.0f, sg = 0.0f, sb = 0.0f;
    Rng rng;
    rng.s0 = rng.s1 = rng.s2 = rng.s3 = 0;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    uint32_t path[MAX_PATH];  // pids of the non-terminal hits of the current sample

    uint32_t tile_next = TILE_W * TILE_H;  // warp-uniform tile cursor (exhausted)
    uint32_t tile_x0 = 0, tile_y0 = 0;
    bool tiles_left = true;
    // primary-ray candidate lists: only for the BVH intersector on scenes small enough to cull per tile
    const bool use_lists = LISTS && (ISECT == RT_INTERSECT_BVH) && (sc.ns + sc.nt <= (uint32_t)pr.list_max_prims);
    int cur_list = 1;       // warp-uniform: list slot of the warp's current tile
    int my_list = 0;        // per lane: list slot of this lane's pixel
    bool list_ok = false;   //           ... and whether that slot still holds this pixel's tile
    const int warp = threadIdx.x >> 5;

    for (;;) {
        // ---- hand out pixels: warp-cooperative, tile by tile ----
        unsigned want = __ballot_sync(FULL, !have_px && !finished);
        while (want) {
            if (tile_next >= (uint32_t)(TILE_W * TILE_H)) {
                unsigned int k = 0;
                if (tiles_left) {
                    if (lane == 0) k = atomicAdd(pr.tile_counter, 1u);
                    k = __shfl_sync(FULL, k, 0);
                }
                uint64_t g = (uint64_t)k * pr.tile_ranks + (pr.tile_rank + k) % pr.tile_ranks;
                if (!tiles_left || g >= total_tiles) {
                    tiles_left = false;
                    if (!have_px) finished = true;
                    break;
                }
                // hand the frame out bottom-up: rows near the ground (long paths) first, the sky rows, whose pixels are
                // short and uniform, last — they fill the tail of the launch (profiles/r1_notes.md, 8-GPU scaling)
                if (pr.tile_order_reverse) g = total_tiles - 1 - g;
                tile_x0 = (uint32_t)(g % pr.tiles_x) * TILE_W;
                tile_y0 = pr.row0 + (uint32_t)(g / pr.tiles_x) * TILE_H;
                tile_next = 0;
                if (LISTS && use_lists) {
                    cur_list ^= 1;
                    // a lane still working on a pixel of the tile that owned this slot falls back to the BVH
                    if (have_px && my_list == cur_list) list_ok = false;
                    __syncwarp();
                    build_tile_list(sc, sv, cam, pr, tile_x0, tile_y0, s_lists[warp][cur_list]);
                }
            }
            const uint32_t avail = TILE_W * TILE_H - tile_next;
            const uint32_t my = __popc(want & lt_mask);
            if (!have_px && !finished && my < avail) {
                const uint32_t j = tile_next + my;
                const uint32_t x = tile_x0 + (j & (TILE_W - 1)), y = tile_y0 + (j / TILE_W);
                if (x < pr.width && y < pr.row1) {  // tiles on the right/bottom edge are partial
                    px = x; py = y;
                    have_px = true;
                    my_list = cur_list;
                    list_ok = use_lists;
                    rng.seed_from_u64(pr.seed + ((uint64_t)y * pr.width + x));
                    sr = sg = sb = 0.0f;
                    s = 0;
                    left = 0;
                }
            }
            tile_next += min((uint32_t)__popc(want), avail);
            want = __ballot_sync(FULL, !have_px && !finished);
        }
        if (__ballot_sync(FULL, have_px) == 0) break;

        if (have_px) {
            bool primary = false;
            if (left == 0) {  // start sample s
                primary_ray(cam, px, pr.height - py - 1, rng, &o, &d);  // y_cam = h - y - 1 (main.rs:71)
                left = pr.depth;
                np = 0;
                primary = true;
            }
            // ---- one nearest-hit query (ray_color with depth > 0) ----
            rays++;
            if (COUNT) {
                ctr.v[CTR_ACTIVE_LANES]++;
                unsigned am = __activemask();
                if (lane == (__ffs(am) - 1)) ctr.v[CTR_TOTAL_LANES] += 32;
                if (pr.ray_dump) {  // measurement aid: record the query (one atomic per warp)
                    unsigned long long base = 0;
                    if (lane == (__ffs(am) - 1)) base = atomicAdd(pr.ray_dump_n, (unsigned long long)__popc(am));
                    base = __shfl_sync(am, base, __ffs(am) - 1);
                    const unsigned long long at = base + __popc(am & lt_mask);
                    if (at < pr.ray_dump_cap) {
                        pr.ray_dump[2 * at] = make_float4(o.x, o.y, o.z, d.x);
                        pr.ray_dump[2 * at + 1] = make_float4(d.y, d.z, 0.0f, 0.0f);
                    }
                }
            }
            Hit h;
            if (ISECT == RT_INTERSECT_BRUTE) {
                trace_brute<COUNT>(sc, sv, o, d, h, ctr);
            } else {
                if (LISTS) {
                    const ListEntry* list = s_lists[warp][my_list];
                    const bool by_list = primary && list_ok && list[0] != LIST_BAD;
                    if (by_list) trace_list<COUNT>(sc, sv, list, o, d, h, ctr);
                    else trace_bvh_ch<COUNT>(sc, sv, o, d, h, ctr);
                } else {
                    trace_bvh_ch<COUNT>(sc, sv, o, d, h, ctr);
                }
            }

            bool done;
            float Lr, Lg, Lb;
            if (h.pid >= 0) {
                const float e = __ldg(&sc.emis[h.pid]);
                const float4 m = __ldg(&sc.mat[h.pid]);
                if (e > 0.0f) {  // emission * albedo (main.rs:116-117)
                    Lr = x_mul(m.x, e); Lg = x_mul(m.y, e); Lb = x_mul(m.z, e);
                    done = true;
                    if (COUNT) ctr.v[CTR_EMISSIVE]++;
                } else {
                    V3 n;
                    if (COUNT) ctr.v[h.pid < (int)sc.ns ? CTR_SHADE_SPH : CTR_SHADE_TRI]++;
                    if (h.pid < (int)sc.ns) {
                        n = x_normalize_or_zero(x_sub(h.p, ld3(sv.sph[h.pid])));  // sphere.rs:49-51
                    } else {
                        n = ld3(sv.tri[4 * (h.pid - (int)sc.ns) + 3]);  // mesh.rs:163-165 (host, same ops)
                    }
                    V3 diffuse = x_add(unit_sphere(rng), n);
                    float kk = x_mul(2.0f, x_dot(d, n));
                    V3 glossy = x_sub(d, x_scale(n, kk));
                    V3 scat = x_add(diffuse, x_scale(x_sub(glossy, diffuse), m.w));
                    V3 nd;
                    if (!x_try_normalize(scat, &nd)) nd = n;
                    o = h.p;
                    d = x_normalize_div(nd);  // Ray::new
                    path[np++] = (uint32_t)h.pid;
                    left--;
                    done = (left == 0);       // next call has depth == 0 → BLACK, no query
                    Lr = Lg = Lb = 0.0f;
                }
            } else {  // sky (main.rs:135-144)
                if (COUNT) ctr.v[CTR_SKY]++;
                float rcp = x_div(1.0f, x_length(d));
                float ny = (isfinite(rcp) && rcp > 0.0f) ? x_mul(d.y, rcp) : 0.0f;
                float t = x_add(x_mul(ny, 0.5f), 1.0f);
                float k1 = x_sub(1.0f, t);
                float w = x_mul(1.0f, t);
                Lr = x_add(w, x_mul(0.3f, k1));
                Lg = Lr;
                Lb = x_add(w, x_mul(0.8f, k1));
                done = true;
            }
            if (done) {
                // fold albedo ⊙ (albedo ⊙ (... ⊙ L)) innermost first, like the recursion unwinding
                while (np > 0) {
                    const float4 m = __ldg(&sc.mat[path[--np]]);
                    Lr = x_mul(m.x, Lr); Lg = x_mul(m.y, Lg); Lb = x_mul(m.z, Lb);
                }
                sr = x_add(sr, Lr); sg = x_add(sg, Lg); sb = x_add(sb, Lb);
                s++;
                left = 0;
                if (s == pr.spp) {  // pixel finished (main.rs:78-81)
                    const size_t off = ((size_t)(py - pr.out_row0) * pr.width + px) * 3;
                    pr.out[off + 0] = (uint8_t)quantise(sr, spp_f);
                    pr.out[off + 1] = (uint8_t)quantise(sg, spp_f);
                    pr.out[off + 2] = (uint8_t)quantise(sb, spp_f);
                    have_px = false;
                }
            }
        }
    }

    ctr.v[CTR_RAYS] = rays;
#pragma unroll
    for (int i = 0; i < NUM_COUNTERS; i++) {
        if (!COUNT && i != CTR_RAYS) continue;
        unsigned long long v = ctr.v[i];
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) v += __shfl_down_sync(FULL, v, ofs);
        if (lane == 0 && v) atomicAdd(&pr.counters[i], v);
    }
}

}  // namespace rtb
#include "rt_kernel_sched.cuh"
#include "rt_kernel_deferred.cuh"
#include "rt_wavefront.cuh"
#include "rt_kernel_wq.cuh"
#include "rt_trace_bench.cuh"
namespace rtb {

// ---------------------------------------------------------------------------------------------
// FFMA-chain micro-benchmark for the FP32 roofline denominator
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
          a7 = a0 + 7;
    const float m = 0.999f + seed * 1e-6f, c = 1e-3f + seed;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ---------------------------------------------------------------------------------------------
// Launcher
// ---------------------------------------------------------------------------------------------
typedef void (*KernelFn)(const DevScene, const DevCamera, const DevParams);

template <int ISECT, bool SMEM>
static KernelFn pick_count(bool count) {
    return count ? (KernelFn)render_kernel<ISECT, SMEM, true> : (KernelFn)render_kernel<ISECT, SMEM, false>;
}
// RT_B200_BVH_KERNEL = lanes (default) | simple | pools | deferred : megakernel variant, kept selectable for A/B runs
static int bvh_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("RT_B200_BVH_KERNEL");
        v = 3;
        if (e && std::strcmp(e, "wave") == 0) v = 4;
        if (e && std::strcmp(e, "wq") == 0) v = 5;
        if (e && std::strcmp(e, "simple") == 0) v = 0;
        if (e && std::strcmp(e, "pools") == 0) v = 1;
        if (e && std::strcmp(e, "deferred") == 0) v = 2;
    }
    return v;
}
bool legacy_node_arrays_needed() { return bvh_variant() != 3 && bvh_variant() != 5; }
static int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}
#define RT_PICK_SCHED(KERNEL)                                                                                       \
    do {                                                                                                            \
        if (minb == 3) {                                                                                            \
            if (smem) return count ? (KernelFn)KERNEL<true, true, 3> : (KernelFn)KERNEL<true, false, 3>;            \
            return count ? (KernelFn)KERNEL<false, true, 3> : (KernelFn)KERNEL<false, false, 3>;                    \
        }                                                                                                           \
        if (smem) return count ? (KernelFn)KERNEL<true, true, 2> : (KernelFn)KERNEL<true, false, 2>;                \
        return count ? (KernelFn)KERNEL<false, true, 2> : (KernelFn)KERNEL<false, false, 2>;                        \
    } while (0)
static int list_max_prims() {  // RT_B200_LIST_MAX=N enables the per-tile primary-ray candidate lists up to N primitives
    static int v = -1;
    if (v < 0) v = env_int("RT_B200_LIST_MAX", 0);
    return v;
}
// RT_B200_LANES_TPB = 768 (one CTA per SM, default) | 384 (2) | 256 (3): the same 24 warps per SM; fewer CTAs keep fewer
// copies of the scene in shared memory and leave more of the 228 KB to L1, where the traversal stacks live
// (C3: 40.7 ms at 3 x 256, 39.8 at 2 x 384, 39.5 at 1 x 768; profiles/r1_notes.md)
static int lanes_tpb() {
    static int v = -1;
    if (v < 0) {
        v = env_int("RT_B200_LANES_TPB", 768);
        if (v != 384 && v != 256) v = 768;
    }
    return v;
}
template <int ISECT, bool SMEM>
static KernelFn pick_lanes(bool count) {
    if (ISECT == RT_INTERSECT_BVH && list_max_prims() > 0 && !count) return (KernelFn)render_kernel_lanes<ISECT, SMEM, false, true>;
    static int minb = -1;  // RT_B200_LANES_MINB=4: 64 registers, 32 resident warps per SM (experiment)
    if (minb < 0) minb = env_int("RT_B200_LANES_MINB", 3);
    // scene read through L1/L2 (too large for shared memory): the BVH kernel is latency bound there, and 32 warps per SM at
    // 64 registers beat 24 at 80 (65,536 spheres: 3.13 -> 2.84 ms; profiles/r1_notes.md)
    const bool wide = !SMEM && ISECT == RT_INTERSECT_BVH && minb == 3;
    if ((minb == 4 || wide) && !count) return (KernelFn)render_kernel_lanes<ISECT, SMEM, false, false, 4>;
    if (lanes_tpb() == 384 && !count) return (KernelFn)render_kernel_lanes<ISECT, SMEM, false, false, 2, 384>;
    if (lanes_tpb() == 768 && !count) return (KernelFn)render_kernel_lanes<ISECT, SMEM, false, false, 1, 768>;
    return count ? (KernelFn)render_kernel_lanes<ISECT, SMEM, true, false> : (KernelFn)render_kernel_lanes<ISECT, SMEM, false, false>;
}
static KernelFn pick_kernel(int isect, bool smem, bool count) {
    if (bvh_variant() == 3 || bvh_variant() == 5) {
        if (isect == RT_INTERSECT_BRUTE) return smem ? pick_lanes<RT_INTERSECT_BRUTE, true>(count) : pick_lanes<RT_INTERSECT_BRUTE, false>(count);
        return smem ? pick_lanes<RT_INTERSECT_BVH, true>(count) : pick_lanes<RT_INTERSECT_BVH, false>(count);
    }
    if (isect == RT_INTERSECT_BRUTE) return smem ? pick_count<RT_INTERSECT_BRUTE, true>(count) : pick_count<RT_INTERSECT_BRUTE, false>(count);
    static int minb = -1;  // RT_B200_SCHED_MINB=3 trades registers (<= 80) for 24 resident warps per SM
    if (minb < 0) minb = env_int("RT_B200_SCHED_MINB", 2) == 3 ? 3 : 2;
    const int v = bvh_variant();
    if (v == 2) RT_PICK_SCHED(render_kernel_deferred);
    if (v == 1) RT_PICK_SCHED(render_kernel_sched);
    return smem ? pick_count<RT_INTERSECT_BVH, true>(count) : pick_count<RT_INTERSECT_BVH, false>(count);
}

size_t scene_smem_bytes(const DevScene& sc, int isect) {
    size_t b = (size_t)sc.ns * 16 + (size_t)sc.nt * 64;
    if (isect == RT_INTERSECT_BVH) b += (size_t)sc.ni * (48 + 8);
    else b += (size_t)((sc.ns + 7u) & ~7u) * 16;  // pair-packed spheres
    return b;
}

cudaError_t launch_render(const DevScene& sc, const DevCamera& cam, const DevParams& pr, int isect, bool count,
                          int sm_count, int smem_optin, cudaStream_t stream, LaunchInfo* info) {
    const size_t static_smem = WARPS * TILE_W * TILE_H * 3 + 64;
    size_t need = scene_smem_bytes(sc, isect);
    // Stage the scene in shared memory only while three CTAs per SM still fit (measured on the C5 sweep: at 2048
    // spheres the 147 KB copy leaves one CTA per SM and loses to the L1/L2 path; profiles/r1_c5_sweep.log).
    static int smem_override = -2;  // RT_B200_SMEM=0|1 forces the choice (measurement only)
    if (smem_override == -2) smem_override = env_int("RT_B200_SMEM", -1);
    const bool lanes = bvh_variant() == 3 || bvh_variant() == 5;
    int threads = (lanes && !count && list_max_prims() == 0 && env_int("RT_B200_LANES_MINB", 3) != 4) ? lanes_tpb() : THREADS;
    const int ctas_target = 768 / threads;  // 24 warps per SM
    bool smem = (need + static_smem + 1024) * (size_t)ctas_target <= (size_t)smem_optin;
    if (smem_override == 0) smem = false;
    if (smem_override == 1) smem = need + static_smem + 1024 <= (size_t)smem_optin;
    if (!smem && isect == RT_INTERSECT_BVH) threads = THREADS;  // the 64-register variant runs 4 x 256 threads (pick_lanes)
    KernelFn fn = pick_kernel(isect, smem, count);
    size_t dyn = smem ? need : 0;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, dyn);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint64_t total_tiles = (uint64_t)pr.tiles_x * pr.tiles_y;
    uint64_t my_tiles = (total_tiles + pr.tile_ranks - 1) / pr.tile_ranks;
    uint64_t want_ctas = (my_tiles + (threads / 32) - 1) / (threads / 32);
    uint64_t grid = (uint64_t)sm_count * per_sm;
    if (grid > want_ctas) grid = want_ctas;
    if (grid < 1) grid = 1;
    DevParams prm = pr;
    static int w[4] = {-1, 0, 0, 0}, nnum = 1, nden = 2;
    if (w[0] < 0) {
        w[0] = env_int("RT_B200_W_NODE", 1);
        w[1] = env_int("RT_B200_W_LEAF", 1);
        w[2] = env_int("RT_B200_W_HIT", 1);
        w[3] = env_int("RT_B200_W_PRIM", 1);
        nnum = env_int("RT_B200_NODE_NUM", 1);
        nden = env_int("RT_B200_NODE_DEN", 2);
    }
    for (int i = 0; i < 4; i++) prm.sched_w[i] = w[i];
    prm.sched_node_num = nnum;
    prm.sched_node_den = nden;
    prm.list_max_prims = list_max_prims();
    static int rev = -1;  // RT_B200_TILE_ORDER=topdown restores the first hand-out order
    if (rev < 0) { const char* e = std::getenv("RT_B200_TILE_ORDER"); rev = (e && std::strcmp(e, "topdown") == 0) ? 0 : 1; }
    prm.tile_order_reverse = rev;
    fn<<<(unsigned)grid, threads, dyn, stream>>>(sc, cam, prm);
    if (info) {
        info->grid = (unsigned)grid;
        info->threads = threads;
        info->dyn_smem = dyn;
        info->ctas_per_sm = per_sm;
        info->scene_in_smem = smem;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Wavefront driver
// ---------------------------------------------------------------------------------------------
bool use_wavefront(int isect) { return isect == RT_INTERSECT_BVH && bvh_variant() == 4; }

void free_wave_buffers(WaveBuffers* wb) {
    if (wb->slots) cudaFree(wb->slots);
    if (wb->q_ray) cudaFree(wb->q_ray);
    if (wb->q_hit) cudaFree(wb->q_hit);
    if (wb->q_miss) cudaFree(wb->q_miss);
    if (wb->path_ext) cudaFree(wb->path_ext);
    if (wb->counters) cudaFree(wb->counters);
    *wb = WaveBuffers();
}

cudaError_t launch_wavefront(const DevScene& sc, const DevCamera& cam, const DevParams& pr, bool count, int sm_count,
                             int smem_optin, cudaStream_t stream, WaveBuffers* wb, LaunchInfo* info) {
    const uint64_t total_tiles = (uint64_t)pr.tiles_x * pr.tiles_y;
    const uint64_t my_tiles = (total_tiles + pr.tile_ranks - 1) / pr.tile_ranks;
    const size_t n_slots = (size_t)my_tiles * TILE_W * TILE_H;
    if (n_slots > 0xfffffff0ull || pr.spp > 65535u) return cudaErrorInvalidValue;
    const size_t ext_depth = pr.depth > 8 ? pr.depth - 8 : 0;
    cudaError_t e;
    if (wb->capacity < n_slots || wb->ext_entries < ext_depth * n_slots) {
        cudaStreamSynchronize(stream);
        free_wave_buffers(wb);
        if ((e = cudaMalloc(&wb->slots, n_slots * sizeof(WSlot))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&wb->q_ray, n_slots * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&wb->q_hit, n_slots * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&wb->q_miss, n_slots * 4)) != cudaSuccess) return e;
        if (ext_depth && (e = cudaMalloc(&wb->path_ext, ext_depth * n_slots * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&wb->counters, 2 * sizeof(WaveCounters))) != cudaSuccess) return e;
        wb->capacity = n_slots;
        wb->ext_entries = ext_depth * n_slots;
    }
    WSlot* slots = (WSlot*)wb->slots;
    WaveCounters* cnt = (WaveCounters*)wb->counters;
    if ((e = cudaMemsetAsync(cnt, 0, 2 * sizeof(WaveCounters), stream)) != cudaSuccess) return e;

    DevParams prm = pr;
    prm.sched_w[0] = env_int("RT_B200_WAVE_REFILL", 8);

    // trace kernel: scene staged in shared memory when it fits
    const size_t need = (size_t)sc.ns * 16 + (size_t)sc.nt * 64 + (size_t)sc.ni * 56;
    const bool smem = need + 1024 <= (size_t)smem_optin;
    typedef void (*TraceFn)(const DevScene, const DevParams, WSlot*, const uint32_t*, uint32_t*, uint32_t*, WaveCounters*, int);
    TraceFn tfn = smem ? (count ? (TraceFn)wave_trace<true, true> : (TraceFn)wave_trace<true, false>)
                       : (count ? (TraceFn)wave_trace<false, true> : (TraceFn)wave_trace<false, false>);
    const size_t dyn = smem ? need : 0;
    if ((e = cudaFuncSetAttribute(tfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)) != cudaSuccess) return e;
    int t_per_sm = 0, l_per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t_per_sm, tfn, 256, dyn)) != cudaSuccess) return e;
    typedef void (*LogicFn)(const DevScene, const DevCamera, const DevParams, WSlot*, const uint32_t*, const uint32_t*,
                            uint32_t*, uint32_t*, uint32_t, WaveCounters*, int);
    LogicFn lfn = count ? (LogicFn)wave_logic<true> : (LogicFn)wave_logic<false>;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&l_per_sm, lfn, 256, 0)) != cudaSuccess) return e;
    if (t_per_sm < 1) t_per_sm = 1;
    if (l_per_sm < 1) l_per_sm = 1;
    const unsigned t_grid = (unsigned)std::min<uint64_t>((uint64_t)sm_count * t_per_sm, (n_slots + 255) / 256);
    const unsigned l_grid = (unsigned)std::min<uint64_t>((uint64_t)sm_count * l_per_sm, (n_slots + 255) / 256);

    wave_init<<<(unsigned)std::min<uint64_t>((uint64_t)sm_count * 8, (n_slots + 255) / 256), 256, 0, stream>>>(
        prm, slots, (uint32_t)n_slots, wb->q_miss, cnt);
    // one round = one query of every live pixel; a pixel makes at most spp * depth queries, +1 round to finish
    const uint64_t rounds = (uint64_t)pr.spp * pr.depth + 1;
    unsigned launches = 1;
    static unsigned int* h_flag = nullptr;
    if (!h_flag) cudaMallocHost(&h_flag, sizeof(unsigned int));
    for (uint64_t r = 0; r < rounds; r++) {
        const int parity = (int)(r & 1);
        lfn<<<l_grid, 256, 0, stream>>>(sc, cam, prm, slots, wb->q_hit, wb->q_miss, wb->q_ray, wb->path_ext,
                                        (uint32_t)n_slots, cnt, parity);
        launches++;
        if (rounds > 160 && (r % 16) == 15) {  // long chains (reference defaults): stop when no ray is left
            cudaMemcpyAsync(h_flag, &cnt[parity ^ 1].n_ray, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
            if (*h_flag == 0) break;
        }
        tfn<<<t_grid, 256, dyn, stream>>>(sc, prm, slots, wb->q_ray, wb->q_hit, wb->q_miss, cnt, parity);
        launches++;
    }
    if (info) {
        info->grid = t_grid;
        info->threads = 256;
        info->dyn_smem = dyn;
        info->ctas_per_sm = t_per_sm;
        info->scene_in_smem = smem;
        info->launches = launches;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Warp-private wavefront (rt_kernel_wq.cuh)
// ---------------------------------------------------------------------------------------------
bool use_wq(int isect, const DevParams& pr) {
    return isect == RT_INTERSECT_BVH && bvh_variant() == 5 && pr.spp <= 65535u && pr.depth <= 255u;
}

void free_wq_buffers(WqBuffers* b) {
    if (b->state) cudaFree(b->state);
    *b = WqBuffers();
}

typedef void (*WqFn)(const DevScene, const DevCamera, const DevParams, const WqArgs);
template <int NW>
static WqFn pick_wq(bool smem, bool count) {
    if (smem) return count ? (WqFn)render_kernel_wq<true, true, NW> : (WqFn)render_kernel_wq<true, false, NW>;
    return count ? (WqFn)render_kernel_wq<false, true, NW> : (WqFn)render_kernel_wq<false, false, NW>;
}

cudaError_t launch_wq(const DevScene& sc, const DevCamera& cam, const DevParams& pr, bool count, int sm_count,
                      int smem_optin, cudaStream_t stream, WqBuffers* wb, LaunchInfo* info) {
    static int nw = -1, chains = -1, min_active = -1, min_node = 24;
    if (nw < 0) {
        nw = env_int("RT_B200_WQ_WARPS", 24);
        if (nw != 16 && nw != 24 && nw != 32) nw = 24;
        chains = env_int("RT_B200_WQ_CHAINS", 128);
        chains = std::max(32, std::min(WQ_MAX_CHAINS, (chains / 32) * 32));
        min_active = env_int("RT_B200_WQ_MIN_ACTIVE", 20);
        min_node = env_int("RT_B200_WQ_MIN_NODE", 24);
    }
    const size_t scene_bytes = (((size_t)sc.ns * 16 + (size_t)sc.nt * 64 + (size_t)sc.lni * 56) + 15) & ~(size_t)15;
    const size_t pool_bytes = (size_t)nw * wq_warp_smem((uint32_t)chains);
    const bool smem = scene_bytes + pool_bytes + 1024 <= (size_t)smem_optin;
    const size_t dyn = pool_bytes + (smem ? scene_bytes : 0);
    WqFn fn = nw == 16 ? pick_wq<16>(smem, count) : nw == 32 ? pick_wq<32>(smem, count) : pick_wq<24>(smem, count);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    // one CTA per SM; fewer when the share of this rank has fewer pixels than the chains of a full grid
    const uint64_t total_tiles = (uint64_t)pr.tiles_x * pr.tiles_y;
    const uint64_t my_tiles = (total_tiles + pr.tile_ranks - 1) / pr.tile_ranks;
    const uint64_t per_cta = (uint64_t)nw * chains / (TILE_W * TILE_H);  // tiles in flight per CTA
    uint64_t grid = std::min<uint64_t>((uint64_t)sm_count, (my_tiles + per_cta - 1) / per_cta);
    if (grid < 1) grid = 1;
    const size_t n = (size_t)grid * nw * chains;
    const size_t bytes = wq_state_bytes(n, pr.depth);
    if (wb->bytes < bytes) {
        cudaStreamSynchronize(stream);
        free_wq_buffers(wb);
        if ((e = cudaMalloc(&wb->state, bytes)) != cudaSuccess) return e;
        wb->bytes = bytes;
    }
    DevParams prm = pr;
    static int rev = -1;
    if (rev < 0) { const char* o = std::getenv("RT_B200_TILE_ORDER"); rev = (o && std::strcmp(o, "topdown") == 0) ? 0 : 1; }
    prm.tile_order_reverse = rev;
    WqArgs wa;
    wa.base = wb->state;
    wa.n = n;
    wa.chains = (uint32_t)chains;
    wa.min_active = (uint32_t)min_active;
    wa.min_node = (uint32_t)min_node;
    static int ww[4] = {-1, 0, 0, 0};
    if (ww[0] < 0) {
        ww[0] = env_int("RT_B200_WQ_BURST", 2);
        ww[1] = env_int("RT_B200_WQ_T_LEAF", 4);
        ww[2] = env_int("RT_B200_WQ_T_PEND", 6);
        ww[3] = env_int("RT_B200_WQ_T_FIN", 6);
    }
    wa.node_burst = ww[0]; wa.t_leaf = ww[1]; wa.t_pend = ww[2]; wa.t_fin = ww[3];
    static int cta_phases = -1, trace_budget = 0;
    if (cta_phases < 0) {
        cta_phases = env_int("RT_B200_WQ_SYNC", 0);
        trace_budget = env_int("RT_B200_WQ_BUDGET", 0);
    }
    wa.cta_phases = (uint32_t)cta_phases;
    wa.trace_budget = (uint32_t)trace_budget;
    wa.scene_bytes = (uint32_t)scene_bytes;
    fn<<<(unsigned)grid, nw * 32, dyn, stream>>>(sc, cam, prm, wa);
    if (info) {
        info->grid = (unsigned)grid;
        info->threads = nw * 32;
        info->dyn_smem = dyn;
        info->ctas_per_sm = 1;
        info->scene_in_smem = smem;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Trace-only benchmark (rt_trace_bench.cuh)
// ---------------------------------------------------------------------------------------------
cudaError_t launch_trace_bench(const DevScene& sc, int variant, bool with_big, const float4* rays, unsigned long long n,
                               unsigned long long* ticket, int2* out, int sm_count, int smem_optin, cudaStream_t stream) {
    const size_t need = (size_t)sc.ns * 16 + (size_t)sc.nt * 64 + (size_t)sc.lni * 56 + 16;
    if (need + 1024 > (size_t)smem_optin) return cudaErrorInvalidValue;  // the benchmark reads the scene from shared memory
    TbArgs a{};
    a.rays = rays;
    a.n = n;
    a.ticket = ticket;
    a.out = out;
    a.node_burst = (uint32_t)env_int("RT_B200_WQ_BURST", 4);
    a.t_leaf = (uint32_t)env_int("RT_B200_WQ_T_LEAF", 4);
    a.t_pend = (uint32_t)env_int("RT_B200_WQ_T_PEND", 6);
    a.t_fin = (uint32_t)env_int("RT_B200_WQ_T_FIN", 6);
    a.alt = (uint32_t)env_int("RT_B200_TB_ALT", 0);
    a.sstack_off = (uint32_t)(need / 4);
    const size_t need_ww = need + (a.alt ? (size_t)TB_SSTACK * 768 * 4 : 0);
    if (a.alt && (need_ww + 1024 > (size_t)smem_optin || sc.lni == 0)) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    DevScene s2 = sc;
    if (!with_big) s2.nbig = 0;  // the tree alone (a wavefront's LOGIC kernel would test the big primitives)
    if (variant == 0) {
        if ((e = cudaFuncSetAttribute(tb_ww, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need_ww)) != cudaSuccess) return e;
        tb_ww<<<sm_count, 768, need_ww, stream>>>(s2, a);
    } else {
        if ((e = cudaFuncSetAttribute(tb_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need)) != cudaSuccess) return e;
        tb_sm<<<sm_count, 768, need, stream>>>(s2, a);
    }
    return cudaGetLastError();
}

cudaError_t launch_fp32_peak(float* scratch, int sm_count, int iters, cudaStream_t stream) {
    fp32_peak_kernel<<<sm_count * 8, 256, 0, stream>>>(scratch, iters, 0.0f);
    return cudaGetLastError();
}

}  // namespace rtb
