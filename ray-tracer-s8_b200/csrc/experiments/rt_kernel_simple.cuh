// rt_kernel_simple.cuh — the first megakernel (tile per warp, lo/hi slabs over the reference-topology tree) and the
// FILTER-domain distance bounds of the deferred / wavefront variants.  Experiments build only.
#pragma once

namespace rtb {

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

}  // namespace rtb
