// rt_kernel_sched.cuh — K2s: the ballot-scheduled BVH megakernel (included by rt_kernels.cu).
//
// Why it exists (profiles/r1_notes.md): in the straightforward megakernel the issue slots were 77 % full but
// only 7.9 of 32 lanes executed the average instruction — a warp waited for its longest traversal, for the
// few lanes doing an exact leaf test, and for the few lanes shading.  Here every lane is an independent
// worker that owns one pixel chain at a time (a pixel's samples and bounces are sequential by the
// reference's RNG semantics, main.rs:69-77) and is always in exactly one pool:
//
//     NODE  : at an inner BVH node            → one two-box slab step (FILTER domain)
//     LEAF  : at a leaf                       → sphere discriminant filter / FMA triangle filter
//     EXACT : leaf passed its filter          → reference arithmetic: roots, t-range, slab check, min_by
//     SHADE : query ended on a scattering hit → normal, UnitSphere, scatter, Ray::new
//     PRIM  : path ended / no pixel yet       → sky|emission|black, fold, accumulate, next sample or pixel,
//                                               Camera::get_ray
//
// Each trip the warp takes ONE vote (a packed REDUX.SUM of 6-bit pool counters) and runs the phase of the
// largest pool, so a phase always executes with the most lanes that could execute it; the other lanes wait
// until their pool is the largest.  Pixels are handed to lanes one at a time from 8x4 tiles the warp pulls
// from the global ticket counter, so a lane that finishes its pixel takes the next one instead of idling.
//
// The slab test is in centre/half-extent form, t = (c - o)*inv -/+ h*|inv|: 9 FFMA + 2 FMNMX3 + 2 FMNMX per
// box instead of 6 FFMA + 10 FMNMX — the first kernel saturated the ALU pipe (75 %) with min/max while the
// FMA pipe sat at 23 %.
#pragma once

namespace rtb {

enum LanePool { P_NODE = 0, P_LEAF = 1, P_EXACT = 2, P_SHADE = 3, P_PRIM = 4, P_DONE = 5 };
enum EndKind { END_NONE = 0, END_SKY = 1, END_EMIT = 2, END_BLACK = 3 };

template <bool SMEM, bool COUNT, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) render_kernel_sched(const DevScene sc, const DevCamera cam,
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

    Ctr ctr;
#pragma unroll
    for (int i = 0; i < NUM_COUNTERS; i++) ctr.v[i] = 0;
    unsigned long long rays = 0;

    // ---- per-lane worker state ----
    int pool = P_PRIM;
    int endk = END_NONE;
    uint32_t px = 0, py = 0, s = 0, left = 0, np = 0;
    float sr = 0.0f, sg = 0.0f, sb = 0.0f;
    Rng rng;
    rng.s0 = rng.s1 = rng.s2 = rng.s3 = 0;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    float ix = 0, iy = 0, iz = 0, ax = 0, ay = 0, az = 0, qx = 0, qy = 0, qz = 0, slack = 0, cull = 0;
    int cur = 0, sp = 0;
    Hit best;
    best.pid = -1; best.dist = 0.0f; best.p = mk(0, 0, 0);
    int stack[MAX_STACK];
    uint32_t path[MAX_PATH];

    // ---- warp-uniform tile cursor ----
    uint32_t tile_next = TILE_W * TILE_H;  // exhausted
    uint32_t tile_x0 = 0, tile_y0 = 0;
    bool tiles_left = true;

    // the traversal of this lane's ray is over: route the lane by what the query found
    auto finish_query = [&]() {
        if (best.pid >= 0) {
            if (__ldg(&sc.emis[best.pid]) > 0.0f) {
                pool = P_PRIM;
                endk = END_EMIT;
            } else {
                pool = P_SHADE;
            }
        } else {
            pool = P_PRIM;
            endk = END_SKY;
        }
    };
    // pop the next node/leaf of this lane's traversal, or finish the query
    auto pop_or_finish = [&]() {
        if (sp > 0) {
            cur = stack[--sp];
            pool = cur >= 0 ? P_NODE : P_LEAF;
        } else {
            finish_query();
        }
    };
    // start one nearest-hit query (ray_color with depth > 0) for the ray (o, d)
    auto start_query = [&]() {
        rays++;
        best.pid = -1;
        best.dist = 0.0f;
        cull = 1001.0f;  // a hit has t < T_MAX and length(p - o) ~ t
        sp = 0;
        cur = sc.root;
        pool = cur >= 0 ? P_NODE : P_LEAF;
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
        // rounding of the o-term: <= 3 * 2^-24 * |o*inv| per axis, in t
        slack = 4.8e-7f * fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + 1e-30f;
    };

    for (;;) {
        // ---- one vote: 6-bit counters of the five live pools packed into one REDUX.SUM ----
        const unsigned contrib = pool < P_DONE ? (1u << (6 * pool)) : 0u;
        const unsigned packed = __reduce_add_sync(FULL, contrib);
        if (packed == 0) break;
        const int nN = packed & 63, nL = (packed >> 6) & 63, nE = (packed >> 12) & 63, nS = (packed >> 18) & 63,
                  nP = (packed >> 24) & 63;
        int phase = P_NODE, nmax = nN;
        if (nL > nmax) { phase = P_LEAF; nmax = nL; }
        if (nE > nmax) { phase = P_EXACT; nmax = nE; }
        if (nS > nmax) { phase = P_SHADE; nmax = nS; }
        if (nP > nmax) { phase = P_PRIM; nmax = nP; }
        if (COUNT) {
            if (pool == phase) ctr.v[CTR_ACTIVE_LANES]++;
            if (lane == 0) ctr.v[CTR_TOTAL_LANES] += 32;
        }

        if (phase == P_NODE) {
            // ================= NODE: slab steps while this pool holds at least half the live lanes =================
            const int live = nN + nL + nE + nS + nP;
            for (;;) {
                if (pool == P_NODE) {
                    const float4 a = g_na[cur], b = g_nb[cur], c = g_nc[cur];
                    const int2 ch = g_nd[cur];
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
                    if (hl && hr) {
                        const bool swap = tr < tl;
                        stack[sp++] = swap ? ch.x : ch.y;
                        cur = swap ? ch.y : ch.x;
                        if (cur < 0) pool = P_LEAF;
                    } else if (hl) {
                        cur = ch.x;
                        if (cur < 0) pool = P_LEAF;
                    } else if (hr) {
                        cur = ch.y;
                        if (cur < 0) pool = P_LEAF;
                    } else {
                        pop_or_finish();
                    }
                }
                const int n = __popc(__ballot_sync(FULL, pool == P_NODE));
                if (2 * n < live) break;
                if (COUNT) {
                    if (pool == P_NODE) ctr.v[CTR_ACTIVE_LANES]++;
                    if (lane == 0) ctr.v[CTR_TOTAL_LANES] += 32;
                }
            }
        } else if (phase == P_LEAF) {
            // ================= LEAF: cheap conservative filters =================
            if (pool == P_LEAF) {
                const int pid = ~cur;
                bool pass;
                if (pid < ns) {
                    if (COUNT) ctr.v[CTR_SPH_TEST]++;
                    pass = sphere_filter(g_sph[pid], o, d);
                } else {
                    if (COUNT) ctr.v[CTR_TRI_TEST]++;
                    pass = triangle_filter(g_tri, pid - ns, o, d, cull);
                }
                if (pass) {
                    pool = P_EXACT;
                } else {
                    pop_or_finish();
                }
            }
        } else if (phase == P_EXACT) {
            // ================= EXACT: the reference's arithmetic for this candidate =================
            if (pool == P_EXACT) {
                const int pid = ~cur;
                if (pid < ns) {
                    sphere_exact<COUNT>(sc, g_sph[pid], pid, o, d, best, ctr);
                } else {
                    triangle_exact<COUNT>(sc, g_tri, pid - ns, pid, o, d, best, ctr);
                }
                if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
                pop_or_finish();
            }
        } else if (phase == P_SHADE) {
            // ================= SHADE: scatter off a non-emissive hit (main.rs:119-132) =================
            if (pool == P_SHADE) {
                const float4 m = __ldg(&sc.mat[best.pid]);
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
                    pool = P_PRIM;
                    endk = END_BLACK;
                } else {
                    start_query();
                }
            }
        } else {
            // ================= PRIM: end the path, next sample / next pixel, Camera::get_ray =================
            bool need_px = false, need_primary = false;
            if (pool == P_PRIM) {
                if (endk == END_NONE) {
                    need_px = true;  // a lane that has no pixel yet
                } else {
                    float Lr, Lg, Lb;
                    if (endk == END_SKY) {  // main.rs:135-144
                        if (COUNT) ctr.v[CTR_SKY]++;
                        float rcp = x_div(1.0f, x_length(d));
                        float ny = (isfinite(rcp) && rcp > 0.0f) ? x_mul(d.y, rcp) : 0.0f;
                        float t = x_add(x_mul(ny, 0.5f), 1.0f);
                        float k1 = x_sub(1.0f, t);
                        float w = x_mul(1.0f, t);
                        Lr = x_add(w, x_mul(0.3f, k1));
                        Lg = Lr;
                        Lb = x_add(w, x_mul(0.8f, k1));
                    } else if (endk == END_EMIT) {  // emission * albedo (main.rs:116-117)
                        if (COUNT) ctr.v[CTR_EMISSIVE]++;
                        const float e = __ldg(&sc.emis[best.pid]);
                        const float4 m = __ldg(&sc.mat[best.pid]);
                        Lr = x_mul(m.x, e); Lg = x_mul(m.y, e); Lb = x_mul(m.z, e);
                    } else {
                        Lr = Lg = Lb = 0.0f;
                    }
                    while (np > 0) {  // albedo ⊙ (albedo ⊙ (... ⊙ L)), innermost first
                        const float4 m = __ldg(&sc.mat[path[--np]]);
                        Lr = x_mul(m.x, Lr); Lg = x_mul(m.y, Lg); Lb = x_mul(m.z, Lb);
                    }
                    sr = x_add(sr, Lr); sg = x_add(sg, Lg); sb = x_add(sb, Lb);
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
                            pool = P_DONE;
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
