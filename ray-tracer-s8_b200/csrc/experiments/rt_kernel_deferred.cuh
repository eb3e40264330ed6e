// rt_kernel_deferred.cuh — K2d: ballot-scheduled BVH megakernel with DEFERRED exact arithmetic
// (included by rt_kernels.cu).
//
// Measurements that led here (profiles/r1_notes.md): the plain megakernel ran 7.9 of 32 lanes per instruction;
// scheduling the warp by pools raised that to 12.6 but not the speed, because most warp-instructions were
// then the reference's exact arithmetic (IEEE divisions, square roots, un-fused products) executed by ~6
// lanes at a time INSIDE the traversal — 0.8 exact primitive tests per ray, each a divergent excursion.
//
// This kernel keeps the traversal entirely in the FILTER domain.  A leaf test yields a conservative interval
// [lo, hi] for the reference's hit distance (sphere_bounds / triangle_bounds); the lane keeps
//     H    = the smallest `hi` of the hits that are CERTAIN to pass the exact test  (also the slab cull distance)
//     list = the (at most 3) candidates whose `lo` <= H, i.e. every primitive that can still be the nearest
// and only when the traversal is over runs the reference's arithmetic, once, on the listed candidates —
// usually exactly one, the winner, whose exact hit point the shading needs anyway.  A ray that reaches the sky
// executes no exact intersection code at all.  Decisions are still the reference's: the final nearest hit is
// chosen by consider() (exact roots, t-range, own-box slab test, min_by distance, DFS-rank ties) among a
// superset of the primitives that could win it.
//
// Lane pools: NODE (slab step) · LEAF (bounds + list insert) · HIT (resolve exactly + shade/scatter) ·
// PRIM (end path, next sample / next pixel, Camera::get_ray).  One packed REDUX vote per trip picks the
// pool with the largest weighted count.
#pragma once

namespace rtb {

enum DPool { D_NODE = 0, D_LEAF = 1, D_HIT = 2, D_PRIM = 3, D_DONE = 4 };

template <bool SMEM, bool COUNT, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) render_kernel_deferred(const DevScene sc, const DevCamera cam,
                                                                      const DevParams pr) {
    extern __shared__ float4 smem_dyn[];
    const float4 *g_sph, *g_tri, *g_na, *g_nb, *g_nc;
    const int2* g_nd;
    if (SMEM) {
        float4* p = smem_dyn;
        float4* s_sph = p;  p += sc.ns;
        float4* s_tri = p;  p += 4 * sc.nt;
        float4* s_na = p;   p += sc.ni;
        float4* s_nb = p;   p += sc.ni;
        float4* s_nc = p;   p += sc.ni;
        int2* s_nd = reinterpret_cast<int2*>(p);
        for (uint32_t i = threadIdx.x; i < sc.ns; i += THREADS) s_sph[i] = __ldg(&sc.sph[i]);
        for (uint32_t i = threadIdx.x; i < 4 * sc.nt; i += THREADS) s_tri[i] = __ldg(&sc.tri[i]);
        for (uint32_t i = threadIdx.x; i < sc.ni; i += THREADS) {
            s_na[i] = __ldg(&sc.cnode_a[i]);
            s_nb[i] = __ldg(&sc.cnode_b[i]);
            s_nc[i] = __ldg(&sc.cnode_c[i]);
            s_nd[i] = __ldg(&sc.node_d[i]);
        }
        __syncthreads();
        g_sph = s_sph; g_tri = s_tri; g_na = s_na; g_nb = s_nb; g_nc = s_nc; g_nd = s_nd;
    } else {
        g_sph = sc.sph; g_tri = sc.tri; g_na = sc.cnode_a; g_nb = sc.cnode_b; g_nc = sc.cnode_c; g_nd = sc.node_d;
    }

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t total_tiles = pr.tiles_x * pr.tiles_y;
    const float spp_f = (float)pr.spp;
    const int ns = (int)sc.ns;
    const bool check_box = (sc.ns + sc.nt) > 1;  // a single-shape world has no parent box to pass (root leaf)
    const int wN = pr.sched_w[0], wL = pr.sched_w[1], wH = pr.sched_w[2], wP = pr.sched_w[3];

    Ctr ctr;
#pragma unroll
    for (int i = 0; i < NUM_COUNTERS; i++) ctr.v[i] = 0;
    unsigned long long rays = 0;

    // ---- per-lane worker state ----
    int pool = D_PRIM;
    int endk = END_NONE;
    uint32_t px = 0, py = 0, s = 0, left = 0, np = 0;
    float sr = 0.0f, sg = 0.0f, sb = 0.0f;
    Rng rng;
    rng.s0 = rng.s1 = rng.s2 = rng.s3 = 0;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    float ix = 0, iy = 0, iz = 0, ax = 0, ay = 0, az = 0, qx = 0, qy = 0, qz = 0, slack = 0, eo = 0;
    float H = 0;                       // upper bound of the nearest certain hit (cull distance)
    int cur = 0, sp = 0;
    bool mid = false;                    // HIT phase entered from a full list, traversal not over
    int l0 = 0, l1 = 0, l2 = 0, ln = 0;  // candidate list: pids
    float f0 = 0, f1 = 0, f2 = 0;        //                 and their lower bounds
    int stack[MAX_STACK];
    uint32_t path[MAX_PATH];

    // ---- warp-uniform tile cursor ----
    uint32_t tile_next = TILE_W * TILE_H;  // exhausted
    uint32_t tile_x0 = 0, tile_y0 = 0;
    bool tiles_left = true;

    // exact arithmetic on every listed candidate → the reference's nearest hit among them.  One code instance,
    // not unrolled: the exact tests are long, and instruction-cache footprint matters more than the loop.
    auto resolve = [&](Hit& best) {
        best.pid = -1;
        best.dist = 0.0f;
        best.p = mk(0, 0, 0);
#pragma unroll 1
        for (int i = 0; i < ln; i++) {
            const int pid = i == 0 ? l0 : (i == 1 ? l1 : l2);
            float t;
            bool ok;
            if (pid < ns) {
                const float4 sp4 = g_sph[pid];
                if (COUNT) ctr.v[CTR_SPH_EXACT]++;
                ok = sphere_root_exact(d, mk(x_sub(o.x, sp4.x), x_sub(o.y, sp4.y), x_sub(o.z, sp4.z)), sp4.w, &t);
                if (COUNT && ok) ctr.v[CTR_SPH_HIT]++;
            } else {
                const int ti = pid - ns;
                int stage;
                ok = triangle_root_exact(o, d, ld3(g_tri[4 * ti]), ld3(g_tri[4 * ti + 1]), ld3(g_tri[4 * ti + 2]), &t, &stage);
                if (COUNT) {
                    if (stage >= 1) ctr.v[CTR_TRI_S1]++;
                    if (stage >= 2) ctr.v[CTR_TRI_S2]++;
                    if (stage >= 3) ctr.v[CTR_TRI_S3]++;
                    if (ok) ctr.v[CTR_TRI_HIT]++;
                }
            }
            if (ok) consider(sc, o, d, t, pid, best);
        }
        ln = 0;
    };
    auto traversal_over = [&]() {
        if (ln > 0) {
            pool = D_HIT;
        } else {
            pool = D_PRIM;
            endk = END_SKY;
        }
    };
    auto pop_or_finish = [&]() {
        if (sp > 0) {
            cur = stack[--sp];
            pool = cur >= 0 ? D_NODE : D_LEAF;
        } else {
            traversal_over();
        }
    };
    auto start_query = [&]() {
        rays++;
        ln = 0;
        H = 1001.0f;  // a hit has t < T_MAX and length(p - o) ~ t
        sp = 0;
        cur = sc.root;
        pool = cur >= 0 ? D_NODE : D_LEAF;
        // FILTER-domain ray constants; |1/d| is clamped so 0*inf never produces NaN slabs
        const float BIG = 1e30f;
        ix = fminf(fmaxf(__frcp_rn(d.x), -BIG), BIG);
        iy = fminf(fmaxf(__frcp_rn(d.y), -BIG), BIG);
        iz = fminf(fmaxf(__frcp_rn(d.z), -BIG), BIG);
        if (!(fabsf(d.x) > 0.0f)) ix = BIG;
        if (!(fabsf(d.y) > 0.0f)) iy = BIG;
        if (!(fabsf(d.z) > 0.0f)) iz = BIG;
        ax = fabsf(ix); ay = fabsf(iy); az = fabsf(iz);
        qx = -o.x * ix; qy = -o.y * iy; qz = -o.z * iz;
        // rounding of the o-term of the slab test: <= 3 * 2^-24 * |o*inv| per axis, in t
        slack = 4.8e-7f * fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + 1e-30f;
        // rounding of ray.at(t) and of length(p - o): a few ulp of |o| (the t-part is added per hit)
        eo = 1e-6f * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z));
    };

    for (;;) {
        // ---- one vote: 6-bit counters of the four live pools packed into one REDUX.SUM ----
        const unsigned contrib = pool < D_DONE ? (1u << (6 * pool)) : 0u;
        const unsigned packed = __reduce_add_sync(FULL, contrib);
        if (packed == 0) break;
        const int nN = packed & 63, nL = (packed >> 6) & 63, nH = (packed >> 12) & 63, nP = (packed >> 18) & 63;
        int phase = D_NODE, best_score = nN * wN;
        if (nL * wL > best_score) { phase = D_LEAF; best_score = nL * wL; }
        if (nH * wH > best_score) { phase = D_HIT; best_score = nH * wH; }
        if (nP * wP > best_score) { phase = D_PRIM; best_score = nP * wP; }
        if (COUNT) {
            if (pool == phase) ctr.v[CTR_ACTIVE_LANES]++;
            if (lane == 0) ctr.v[CTR_TOTAL_LANES] += 32;
        }

        if (phase == D_NODE) {
            // ================= NODE: slab steps while this pool keeps its share of the live lanes =================
            const int live = nN + nL + nH + nP;
            for (;;) {
                if (pool == D_NODE) {
                    const float4 a = g_na[cur], b = g_nb[cur], c = g_nc[cur];
                    const int2 ch = g_nd[cur];
                    // left box: c = (a.x,a.y,a.z) h = (a.w,b.x,b.y); right: c = (b.z,b.w,c.x) h = (c.y,c.z,c.w)
                    const float lcx = fmaf(a.x, ix, qx), lcy = fmaf(a.y, iy, qy), lcz = fmaf(a.z, iz, qz);
                    const float rcx = fmaf(b.z, ix, qx), rcy = fmaf(b.w, iy, qy), rcz = fmaf(c.x, iz, qz);
                    const float tl = fmaxf(fmaxf(fmaf(-a.w, ax, lcx), fmaf(-b.x, ay, lcy)), fmaxf(fmaf(-b.y, az, lcz), 0.0f));
                    const float fl = fminf(fminf(fmaf(a.w, ax, lcx), fmaf(b.x, ay, lcy)), fminf(fmaf(b.y, az, lcz), H));
                    const float tr = fmaxf(fmaxf(fmaf(-c.y, ax, rcx), fmaf(-c.z, ay, rcy)), fmaxf(fmaf(-c.w, az, rcz), 0.0f));
                    const float fr = fminf(fminf(fmaf(c.y, ax, rcx), fmaf(c.z, ay, rcy)), fminf(fmaf(c.w, az, rcz), H));
                    const bool hl = tl <= fl + slack;
                    const bool hr = tr <= fr + slack;
                    if (COUNT) ctr.v[CTR_SLAB] += 2;
                    if (hl && hr) {
                        const bool swap = tr < tl;
                        stack[sp++] = swap ? ch.x : ch.y;
                        cur = swap ? ch.y : ch.x;
                        if (cur < 0) pool = D_LEAF;
                    } else if (hl) {
                        cur = ch.x;
                        if (cur < 0) pool = D_LEAF;
                    } else if (hr) {
                        cur = ch.y;
                        if (cur < 0) pool = D_LEAF;
                    } else {
                        pop_or_finish();
                    }
                }
                const int n = __popc(__ballot_sync(FULL, pool == D_NODE));
                if (n * pr.sched_node_den < live * pr.sched_node_num) break;
                if (COUNT) {
                    if (pool == D_NODE) ctr.v[CTR_ACTIVE_LANES]++;
                    if (lane == 0) ctr.v[CTR_TOTAL_LANES] += 32;
                }
            }
        } else if (phase == D_LEAF) {
            // ================= LEAF: FILTER-domain distance bounds, candidate list =================
            if (pool == D_LEAF) {
                const int pid = ~cur;
                float lo, hi;
                int cl;
                if (pid < ns) {
                    if (COUNT) ctr.v[CTR_SPH_TEST]++;
                    cl = sphere_bounds(g_sph[pid], o, d, eo, check_box, &lo, &hi);
                } else {
                    if (COUNT) ctr.v[CTR_TRI_TEST]++;
                    cl = triangle_bounds(g_tri, pid - ns, o, d, eo, check_box, &lo, &hi);
                }
                if (cl != CL_MISS && lo <= H) {
                    if (cl == CL_SURE && hi < H) {  // a certain hit tightens the cull distance and prunes the list
                        H = hi;
                        if (ln > 2 && f2 > H) ln = 2;
                        if (ln > 1 && f1 > H) { l1 = l2; f1 = f2; ln--; }
                        if (ln > 0 && f0 > H) { l0 = l1; f0 = f1; l1 = l2; f1 = f2; ln--; }
                    }
                    if (ln == 3) {
                        // list full (rare): let the HIT phase resolve the three exactly and come back to this leaf
                        pool = D_HIT;
                        mid = true;
                    } else {
                        if (ln == 0) { l0 = pid; f0 = lo; } else if (ln == 1) { l1 = pid; f1 = lo; } else { l2 = pid; f2 = lo; }
                        ln++;
                    }
                }
                if (!mid) pop_or_finish();
            }
        } else if (phase == D_HIT) {
            // ========== HIT: the reference's arithmetic on the surviving candidates, then shade (main.rs:114-132) ==========
            if (pool == D_HIT) {
                Hit best;
                resolve(best);
                if (mid) {  // mid-traversal resolve of a full list: keep the exact winner, re-test the pending leaf
                    mid = false;
                    if (best.pid >= 0) {
                        l0 = best.pid;
                        f0 = best.dist;
                        ln = 1;
                        if (best.dist < H) H = best.dist;
                    }
                    pool = D_LEAF;
                } else if (best.pid < 0) {  // every candidate failed the exact test
                    pool = D_PRIM;
                    endk = END_SKY;
                } else {
                    const float e = __ldg(&sc.emis[best.pid]);
                    const float4 m = __ldg(&sc.mat[best.pid]);
                    if (e > 0.0f) {  // emission * albedo ends the path (main.rs:116-117)
                        if (COUNT) ctr.v[CTR_EMISSIVE]++;
                        float Lr = x_mul(m.x, e), Lg = x_mul(m.y, e), Lb = x_mul(m.z, e);
                        while (np > 0) {
                            const float4 mm = __ldg(&sc.mat[path[--np]]);
                            Lr = x_mul(mm.x, Lr); Lg = x_mul(mm.y, Lg); Lb = x_mul(mm.z, Lb);
                        }
                        sr = x_add(sr, Lr); sg = x_add(sg, Lg); sb = x_add(sb, Lb);
                        pool = D_PRIM;
                        endk = END_EMIT;  // already accumulated
                    } else {
                        V3 n;
                        if (COUNT) ctr.v[best.pid < ns ? CTR_SHADE_SPH : CTR_SHADE_TRI]++;
                        if (best.pid < ns) {
                            n = x_normalize_or_zero(x_sub(best.p, ld3(g_sph[best.pid])));  // sphere.rs:49-51
                        } else {
                            n = ld3(g_tri[4 * (best.pid - ns) + 3]);                        // mesh.rs:163-165
                        }
                        V3 diffuse = x_add(unit_sphere(rng), n);
                        float kk = x_mul(2.0f, x_dot(d, n));
                        V3 glossy = x_sub(d, x_scale(n, kk));
                        V3 scat = x_add(diffuse, x_scale(x_sub(glossy, diffuse), m.w));
                        V3 nd;
                        if (!x_try_normalize(scat, &nd)) nd = n;
                        o = best.p;
                        d = x_normalize_div(nd);  // Ray::new
                        path[np++] = (uint32_t)best.pid;
                        left--;
                        if (left == 0) {  // the recursive call has depth == 0 → BLACK, no query
                            pool = D_PRIM;
                            endk = END_BLACK;
                        } else {
                            start_query();
                        }
                    }
                }
            }
        } else {
            // ================= PRIM: end the path, next sample / next pixel, Camera::get_ray =================
            bool need_px = false, need_primary = false;
            if (pool == D_PRIM) {
                if (endk == END_NONE) {
                    need_px = true;  // a lane that has no pixel yet
                } else {
                    if (endk != END_EMIT) {  // (an emissive end was accumulated by the HIT phase)
                        float Lr = 0.0f, Lg = 0.0f, Lb = 0.0f;  // END_BLACK: ray_color(depth == 0) (main.rs:109-111)
                        if (endk == END_SKY) {               // main.rs:135-144
                            if (COUNT) ctr.v[CTR_SKY]++;
                            float rcp = x_div(1.0f, x_length(d));
                            float ny = (isfinite(rcp) && rcp > 0.0f) ? x_mul(d.y, rcp) : 0.0f;
                            float t = x_add(x_mul(ny, 0.5f), 1.0f);
                            float k1 = x_sub(1.0f, t);
                            float w = x_mul(1.0f, t);
                            Lr = x_add(w, x_mul(0.3f, k1));
                            Lg = Lr;
                            Lb = x_add(w, x_mul(0.8f, k1));
                        }
                        while (np > 0) {  // albedo ⊙ (albedo ⊙ (... ⊙ L)), innermost first
                            const float4 m = __ldg(&sc.mat[path[--np]]);
                            Lr = x_mul(m.x, Lr); Lg = x_mul(m.y, Lg); Lb = x_mul(m.z, Lb);
                        }
                        sr = x_add(sr, Lr); sg = x_add(sg, Lg); sb = x_add(sb, Lb);
                    }
                    s++;
                    if (s < pr.spp) {
                        need_primary = true;
                    } else {  // pixel finished (main.rs:78-81)
                        const size_t off = ((size_t)(py - pr.out_row0) * pr.width + px) * 3;
                        pr.out[off + 0] = (uint8_t)quantise(sr, spp_f);
                        pr.out[off + 1] = (uint8_t)quantise(sg, spp_f);
                        pr.out[off + 2] = (uint8_t)quantise(sb, spp_f);
                        need_px = true;
                        endk = END_NONE;
                    }
                }
            }
            // ---- hand out pixels: warp-cooperative, tile by tile ----
            unsigned want = __ballot_sync(FULL, need_px);
            while (want) {
                if (tile_next >= (uint32_t)(TILE_W * TILE_H)) {
                    unsigned int k = 0;
                    if (tiles_left) {
                        if (lane == 0) k = atomicAdd(pr.tile_counter, 1u);
                        k = __shfl_sync(FULL, k, 0);
                    }
                    const uint64_t g = (uint64_t)k * pr.tile_ranks + (pr.tile_rank + k) % pr.tile_ranks;
                    if (!tiles_left || g >= total_tiles) {
                        tiles_left = false;
                        if (need_px) {
                            pool = D_DONE;
                            need_px = false;
                        }
                        break;
                    }
                    tile_x0 = (uint32_t)(g % pr.tiles_x) * TILE_W;
                    tile_y0 = pr.row0 + (uint32_t)(g / pr.tiles_x) * TILE_H;
                    tile_next = 0;
                }
                const uint32_t avail = TILE_W * TILE_H - tile_next;
                const uint32_t my = __popc(want & lt_mask);
                if (need_px && my < avail) {
                    const uint32_t j = tile_next + my;
                    const uint32_t x = tile_x0 + (j & (TILE_W - 1)), y = tile_y0 + (j / TILE_W);
                    if (x < pr.width && y < pr.row1) {  // tiles on the right/bottom edge are partial
                        px = x; py = y;
                        need_px = false;
                        rng.seed_from_u64(pr.seed + ((uint64_t)y * pr.width + x));
                        sr = sg = sb = 0.0f;
                        s = 0;
                        need_primary = true;
                    }
                }
                const uint32_t served = min((uint32_t)__popc(want), avail);
                tile_next += served;
                want = __ballot_sync(FULL, need_px);
            }
            if (need_primary) {
                primary_ray(cam, px, pr.height - py - 1, rng, &o, &d);  // y_cam = h - y - 1 (main.rs:71)
                left = pr.depth;
                np = 0;
                start_query();
            }
        }
    }

    // ---- counters: warp-reduce, one atomic per warp per slot ----
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
