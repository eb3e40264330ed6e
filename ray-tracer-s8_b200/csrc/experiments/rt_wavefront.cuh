// rt_wavefront.cuh — the WAVEFRONT form of the render path (included by rt_kernels.cu).
//
// Why: ncu on the megakernels (profiles/r1_notes.md) shows issue slots 78 % full but only ~8 of 32 lanes doing
// work — a warp waits for its longest BVH traversal, and refilling a lane inside the megakernel costs a
// 300-instruction regeneration (shade or Camera::get_ray in exact arithmetic).  Here the path state of every
// pixel lives in a 128-byte slot in HBM, and one "round" = one nearest-hit query for every live pixel:
//
//   wave_logic : one thread per finished query, hits and misses in separate lists so a warp runs one kind of
//                work: resolve the candidate list with the reference's exact arithmetic, shade/scatter, or end
//                the path (sky | emission | depth), accumulate, start the next sample (Camera::get_ray) or write
//                the pixel.  Emits the next ray of the pixel into the slot and the slot into the ray queue.
//   wave_trace : persistent warps; a lane that finishes its ray takes the next one from the queue (a 32-byte
//                load, warp-aggregated ticket).  FILTER-domain traversal only: conservative distance bounds per
//                leaf, a <= 3-entry candidate list per ray (the deferred-exact scheme of rt_kernel_deferred.cuh);
//                lanes at inner nodes and lanes at leaves are run as two ballot-scheduled pools.
//
// A pixel's samples and bounces stay sequential on its own xoshiro256++ stream (main.rs:69-77), so the number of
// rounds is the longest chain of queries of any pixel (<= spp * (max_bounces + 1)); late rounds are short.
// Results are bit-identical to the megakernels (same device functions, same decisions).
#pragma once

namespace rtb {

struct __align__(16) WSlot {  // 128 bytes = one L2 line
    uint4 rngA, rngB;         // xoshiro256++ state (s0,s1 | s2,s3) as lo/hi words
    float4 ro;                // o.xyz, d.x
    float2 rd;                // d.y, d.z
    uint32_t meta;            // s (16 bits) | left (8) | np (8)
    uint32_t pix;             // y * width + x of the global image
    float4 acc;               // sample sum r,g,b
    int4 res;                 // candidate list of the finished query: n (-1 = no query yet), pid0, pid1, pid2
    uint32_t path[8];         // pids of the scattering hits of the current sample (deeper entries: path_ext)
};
static_assert(sizeof(WSlot) == 128, "slot layout");

struct WaveCounters {  // one per round parity
    unsigned int n_ray, head, n_hit, n_miss;
};

__device__ __forceinline__ uint32_t pack_meta(uint32_t s, uint32_t left, uint32_t np) { return (s << 16) | (left << 8) | np; }

// slot index → pixel of this rank's share (same ticket map as the megakernels)
__device__ __forceinline__ bool slot_pixel(const DevParams& pr, uint32_t slot, uint32_t* x, uint32_t* y) {
    const uint32_t k = slot / (TILE_W * TILE_H), j = slot % (TILE_W * TILE_H);
    const uint64_t g = (uint64_t)k * pr.tile_ranks + (pr.tile_rank + k) % pr.tile_ranks;
    if (g >= (uint64_t)pr.tiles_x * pr.tiles_y) return false;
    *x = (uint32_t)(g % pr.tiles_x) * TILE_W + (j & (TILE_W - 1));
    *y = pr.row0 + (uint32_t)(g / pr.tiles_x) * TILE_H + (j / TILE_W);
    return *x < pr.width && *y < pr.row1;
}

// warp-aggregated append of `slot` to a queue (all 32 lanes call; `yes` selects the lanes that append)
__device__ __forceinline__ void queue_push(bool yes, uint32_t slot, uint32_t* q, unsigned int* counter) {
    const unsigned m = __ballot_sync(0xffffffffu, yes);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned int base = 0;
    if (lane == (__ffs(m) - 1)) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (yes) q[base + __popc(m & ((1u << lane) - 1u))] = slot;
}

__global__ void __launch_bounds__(256) wave_init(const DevParams pr, WSlot* slots, uint32_t n_slots, uint32_t* q_miss,
                                                 WaveCounters* cnt) {
    const uint32_t stride = gridDim.x * blockDim.x;
    // n_slots is a multiple of 32, so whole warps stay together for the aggregated push
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        uint32_t x, y;
        const bool valid = slot_pixel(pr, i, &x, &y);
        if (valid) {
            Rng rng;
            rng.seed_from_u64(pr.seed + ((uint64_t)y * pr.width + x));
            WSlot* s = &slots[i];
            s->rngA = make_uint4((uint32_t)rng.s0, (uint32_t)(rng.s0 >> 32), (uint32_t)rng.s1, (uint32_t)(rng.s1 >> 32));
            s->rngB = make_uint4((uint32_t)rng.s2, (uint32_t)(rng.s2 >> 32), (uint32_t)rng.s3, (uint32_t)(rng.s3 >> 32));
            s->meta = pack_meta(0, 0, 0);
            s->pix = y * pr.width + x;
            s->acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            s->res = make_int4(-1, 0, 0, 0);
        }
        queue_push(valid, i, q_miss, &cnt[0].n_miss);
    }
}

// The reference's exact arithmetic on a candidate list → nearest hit (consider(): roots, t-range, own-box slab
// test, min_by distance, DFS-rank ties).  One code instance, called from wave_logic and (rarely) wave_trace.
template <bool COUNT>
__device__ __noinline__ void exact_resolve(const DevScene& sc, V3 o, V3 d, int n, int p0, int p1, int p2, Hit* out,
                                           Ctr* ctr) {
    Hit best;
    best.pid = -1;
    best.dist = 0.0f;
    best.p = mk(0, 0, 0);
    const int ns = (int)sc.ns;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const int pid = i == 0 ? p0 : (i == 1 ? p1 : p2);
        float t;
        bool ok;
        if (pid < ns) {
            const float4 s = __ldg(&sc.sph[pid]);
            if (COUNT) ctr->v[CTR_SPH_EXACT]++;
            ok = sphere_root_exact(d, mk(x_sub(o.x, s.x), x_sub(o.y, s.y), x_sub(o.z, s.z)), s.w, &t);
            if (COUNT && ok) ctr->v[CTR_SPH_HIT]++;
        } else {
            const int ti = pid - ns;
            int stage;
            ok = triangle_root_exact(o, d, ld3(__ldg(&sc.tri[4 * ti])), ld3(__ldg(&sc.tri[4 * ti + 1])),
                                     ld3(__ldg(&sc.tri[4 * ti + 2])), &t, &stage);
            if (COUNT) {
                if (stage >= 1) ctr->v[CTR_TRI_S1]++;
                if (stage >= 2) ctr->v[CTR_TRI_S2]++;
                if (stage >= 3) ctr->v[CTR_TRI_S3]++;
                if (ok) ctr->v[CTR_TRI_HIT]++;
            }
        }
        if (ok) consider(sc, o, d, t, pid, best);
    }
    *out = best;
}

// ---------------------------------------------------------------------------------------------------------
// wave_logic: consume the finished queries of the previous round, emit the next ray of each live pixel
// ---------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(256, 2) wave_logic(const DevScene sc, const DevCamera cam, const DevParams pr,
                                                     WSlot* slots, const uint32_t* q_hit, const uint32_t* q_miss,
                                                     uint32_t* q_ray, uint32_t* path_ext, uint32_t n_slots,
                                                     WaveCounters* cnt, int parity) {
    const WaveCounters in = cnt[parity];
    WaveCounters* out = &cnt[parity ^ 1];
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t chunks_hit = (in.n_hit + 31) / 32, chunks_miss = (in.n_miss + 31) / 32;
    const float spp_f = (float)pr.spp;
    const int ns = (int)sc.ns;

    Ctr ctr;
#pragma unroll
    for (int i = 0; i < NUM_COUNTERS; i++) ctr.v[i] = 0;
    unsigned long long rays = 0;

    for (uint32_t chunk = warp_global; chunk < chunks_hit + chunks_miss; chunk += warps_total) {
        const bool is_hit_chunk = chunk < chunks_hit;
        const uint32_t idx = (is_hit_chunk ? chunk : chunk - chunks_hit) * 32 + lane;
        const bool live = idx < (is_hit_chunk ? in.n_hit : in.n_miss);
        uint32_t slot = 0;
        bool emit = false;
        if (live) {
            slot = is_hit_chunk ? q_hit[idx] : q_miss[idx];
            WSlot* s = &slots[slot];
            // ---- load the path state ----
            const uint4 ra = s->rngA, rb = s->rngB;
            Rng rng;
            rng.s0 = ((uint64_t)ra.y << 32) | ra.x; rng.s1 = ((uint64_t)ra.w << 32) | ra.z;
            rng.s2 = ((uint64_t)rb.y << 32) | rb.x; rng.s3 = ((uint64_t)rb.w << 32) | rb.z;
            const float4 ro = s->ro;
            const float2 rd = s->rd;
            V3 o = mk(ro.x, ro.y, ro.z), d = mk(ro.w, rd.x, rd.y);
            const uint32_t meta = s->meta, pix = s->pix;
            uint32_t smp = meta >> 16, left = (meta >> 8) & 255u, np = meta & 255u;
            const int4 res = s->res;
            auto path_ref = [&](uint32_t k) -> uint32_t* {
                return k < 8 ? &s->path[k] : &path_ext[(size_t)(k - 8) * n_slots + slot];
            };

            int endk = END_NONE;  // END_NONE here = "a new ray was produced by scattering"
            bool have_L = false;
            float Lr = 0.0f, Lg = 0.0f, Lb = 0.0f;
            if (res.x < 0) {
                endk = END_BLACK;  // no query yet: go straight to the first sample (nothing to accumulate)
            } else if (res.x == 0) {
                endk = END_SKY;
            } else {
                Hit best;
                exact_resolve<COUNT>(sc, o, d, res.x, res.y, res.z, res.w, &best, &ctr);
                if (best.pid < 0) {
                    endk = END_SKY;  // every candidate failed the exact test
                } else {
                    const float e = __ldg(&sc.emis[best.pid]);
                    const float4 m = __ldg(&sc.mat[best.pid]);
                    if (e > 0.0f) {  // emission * albedo ends the path (main.rs:116-117)
                        if (COUNT) ctr.v[CTR_EMISSIVE]++;
                        Lr = x_mul(m.x, e); Lg = x_mul(m.y, e); Lb = x_mul(m.z, e);
                        have_L = true;
                        endk = END_EMIT;
                    } else {  // scatter (main.rs:119-132)
                        V3 n;
                        if (COUNT) ctr.v[best.pid < ns ? CTR_SHADE_SPH : CTR_SHADE_TRI]++;
                        if (best.pid < ns) {
                            n = x_normalize_or_zero(x_sub(best.p, ld3(__ldg(&sc.sph[best.pid]))));  // sphere.rs:49-51
                        } else {
                            n = ld3(__ldg(&sc.tri[4 * (best.pid - ns) + 3]));                        // mesh.rs:163-165
                        }
                        V3 diffuse = x_add(unit_sphere(rng), n);
                        float kk = x_mul(2.0f, x_dot(d, n));
                        V3 glossy = x_sub(d, x_scale(n, kk));
                        V3 scat = x_add(diffuse, x_scale(x_sub(glossy, diffuse), m.w));
                        V3 nd;
                        if (!x_try_normalize(scat, &nd)) nd = n;
                        o = best.p;
                        d = x_normalize_div(nd);  // Ray::new
                        *path_ref(np) = (uint32_t)best.pid;
                        np++;
                        left--;
                        if (left == 0) {  // the recursive call has depth == 0 → BLACK, no query
                            have_L = true;
                            endk = END_BLACK;
                        } else {
                            emit = true;
                        }
                    }
                }
            }
            if (!emit) {
                // ---- the path is over (or this is the pixel's first sample) ----
                float sr, sg, sb;
                {
                    const float4 a = s->acc;
                    sr = a.x; sg = a.y; sb = a.z;
                }
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
                    have_L = true;
                }
                if (have_L) {
                    while (np > 0) {  // albedo ⊙ (albedo ⊙ (... ⊙ L)), innermost first
                        np--;
                        const float4 m = __ldg(&sc.mat[*path_ref(np)]);
                        Lr = x_mul(m.x, Lr); Lg = x_mul(m.y, Lg); Lb = x_mul(m.z, Lb);
                    }
                    sr = x_add(sr, Lr); sg = x_add(sg, Lg); sb = x_add(sb, Lb);
                    smp++;
                }
                if (smp < pr.spp) {  // next sample: Camera::get_ray
                    const uint32_t px = pix % pr.width, py = pix / pr.width;
                    primary_ray(cam, px, pr.height - py - 1, rng, &o, &d);  // y_cam = h - y - 1 (main.rs:71)
                    left = pr.depth;
                    np = 0;
                    emit = true;
                    s->acc = make_float4(sr, sg, sb, 0.0f);
                } else {  // pixel finished (main.rs:78-81)
                    const uint32_t px = pix % pr.width, py = pix / pr.width;
                    const size_t off = ((size_t)(py - pr.out_row0) * pr.width + px) * 3;
                    pr.out[off + 0] = (uint8_t)quantise(sr, spp_f);
                    pr.out[off + 1] = (uint8_t)quantise(sg, spp_f);
                    pr.out[off + 2] = (uint8_t)quantise(sb, spp_f);
                }
            }
            if (emit) {
                rays++;
                s->rngA = make_uint4((uint32_t)rng.s0, (uint32_t)(rng.s0 >> 32), (uint32_t)rng.s1, (uint32_t)(rng.s1 >> 32));
                s->rngB = make_uint4((uint32_t)rng.s2, (uint32_t)(rng.s2 >> 32), (uint32_t)rng.s3, (uint32_t)(rng.s3 >> 32));
                s->ro = make_float4(o.x, o.y, o.z, d.x);
                s->rd = make_float2(d.y, d.z);
                s->meta = pack_meta(smp, left, np);
            }
        }
        queue_push(emit, slot, q_ray, &out->n_ray);
    }

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

// ---------------------------------------------------------------------------------------------------------
// wave_trace: persistent FILTER-domain traversal with warp-aggregated refill from the ray queue
// ---------------------------------------------------------------------------------------------------------
template <bool SMEM, bool COUNT>
__global__ void __launch_bounds__(256, 3) wave_trace(const DevScene sc, const DevParams pr, WSlot* slots,
                                                     const uint32_t* q_ray, uint32_t* q_hit, uint32_t* q_miss,
                                                     WaveCounters* cnt, int parity) {
    extern __shared__ float4 smem_dyn[];
    // this launch consumes cnt[parity^1].n_ray and fills cnt[parity^1].n_hit/n_miss; cnt[parity] is free: zero it
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cnt[parity].n_ray = 0; cnt[parity].head = 0; cnt[parity].n_hit = 0; cnt[parity].n_miss = 0;
    }
    WaveCounters* c = &cnt[parity ^ 1];
    const unsigned int n_ray = c->n_ray;
    if (n_ray == 0) return;

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
        for (uint32_t i = threadIdx.x; i < sc.ns; i += blockDim.x) s_sph[i] = __ldg(&sc.sph[i]);
        for (uint32_t i = threadIdx.x; i < 4 * sc.nt; i += blockDim.x) s_tri[i] = __ldg(&sc.tri[i]);
        for (uint32_t i = threadIdx.x; i < sc.ni; i += blockDim.x) {
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
    const int ns = (int)sc.ns;
    const bool check_box = (sc.ns + sc.nt) > 1;
    const int refill_min = pr.sched_w[0] > 0 ? pr.sched_w[0] : 8;

    Ctr ctr;
#pragma unroll
    for (int i = 0; i < NUM_COUNTERS; i++) ctr.v[i] = 0;

    // lane state: 0 idle, 1 at inner node, 2 at leaf, 3 finished (result not yet published)
    int st = 0;
    uint32_t slot = 0;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    float ix = 0, iy = 0, iz = 0, ax = 0, ay = 0, az = 0, qx = 0, qy = 0, qz = 0, slack = 0, eo = 0, H = 0;
    int cur = 0, sp = 0;
    int l0 = 0, l1 = 0, l2 = 0, ln = 0;
    float f0 = 0, f1 = 0, f2 = 0;
    int stack[MAX_STACK];
    bool exhausted = false;  // warp-uniform: the queue has no more rays

    for (;;) {
        // ---- publish finished queries, refill idle lanes (warp-synchronous) ----
        const unsigned fin = __ballot_sync(FULL, st == 3);
        const unsigned idle_or_fin = __ballot_sync(FULL, st == 0 || st == 3);
        const int n_free = __popc(idle_or_fin);
        if (n_free >= refill_min || n_free == 32 || (exhausted && fin)) {
            if (fin) {
                if (st == 3) slots[slot].res = make_int4(ln, l0, l1, l2);
                queue_push(st == 3 && ln > 0, slot, q_hit, &c->n_hit);
                queue_push(st == 3 && ln == 0, slot, q_miss, &c->n_miss);
                if (st == 3) st = 0;
            }
            if (!exhausted) {
                const unsigned want = __ballot_sync(FULL, st == 0);
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(&c->head, (unsigned)__popc(want));
                base = __shfl_sync(FULL, base, 0);
                if (base + __popc(want) >= n_ray) exhausted = true;
                if (st == 0) {
                    const unsigned int idx = base + __popc(want & lt_mask);
                    if (idx < n_ray) {
                        slot = q_ray[idx];
                        const float4 ro = slots[slot].ro;
                        const float2 rd = slots[slot].rd;
                        o = mk(ro.x, ro.y, ro.z);
                        d = mk(ro.w, rd.x, rd.y);
                        ln = 0;
                        H = 1001.0f;  // a hit has t < T_MAX and length(p - o) ~ t
                        sp = 0;
                        cur = sc.root;
                        st = cur >= 0 ? 1 : 2;
                        const float BIG = 1e30f;  // |1/d| clamped: 0*inf never produces NaN slabs
                        ix = fminf(fmaxf(__frcp_rn(d.x), -BIG), BIG);
                        iy = fminf(fmaxf(__frcp_rn(d.y), -BIG), BIG);
                        iz = fminf(fmaxf(__frcp_rn(d.z), -BIG), BIG);
                        if (!(fabsf(d.x) > 0.0f)) ix = BIG;
                        if (!(fabsf(d.y) > 0.0f)) iy = BIG;
                        if (!(fabsf(d.z) > 0.0f)) iz = BIG;
                        ax = fabsf(ix); ay = fabsf(iy); az = fabsf(iz);
                        qx = -o.x * ix; qy = -o.y * iy; qz = -o.z * iz;
                        slack = 4.8e-7f * fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + 1e-30f;
                        eo = 1e-6f * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z));
                    }
                }
            }
        }
        const int nN = __popc(__ballot_sync(FULL, st == 1));
        const int nL = __popc(__ballot_sync(FULL, st == 2));
        if (nN + nL == 0) {
            if (exhausted && __ballot_sync(FULL, st == 3) == 0) break;
            continue;
        }

        if (nN >= nL) {
            // ================= NODE pool: slab steps while it holds at least half of the traversing lanes =================
            const int live = nN + nL;
            for (;;) {
                if (st == 1) {
                    const float4 a = g_na[cur], b = g_nb[cur], cc = g_nc[cur];
                    const int2 ch = g_nd[cur];
                    const float lcx = fmaf(a.x, ix, qx), lcy = fmaf(a.y, iy, qy), lcz = fmaf(a.z, iz, qz);
                    const float rcx = fmaf(b.z, ix, qx), rcy = fmaf(b.w, iy, qy), rcz = fmaf(cc.x, iz, qz);
                    const float tl = fmaxf(fmaxf(fmaf(-a.w, ax, lcx), fmaf(-b.x, ay, lcy)), fmaxf(fmaf(-b.y, az, lcz), 0.0f));
                    const float fl = fminf(fminf(fmaf(a.w, ax, lcx), fmaf(b.x, ay, lcy)), fminf(fmaf(b.y, az, lcz), H));
                    const float tr = fmaxf(fmaxf(fmaf(-cc.y, ax, rcx), fmaf(-cc.z, ay, rcy)), fmaxf(fmaf(-cc.w, az, rcz), 0.0f));
                    const float fr = fminf(fminf(fmaf(cc.y, ax, rcx), fmaf(cc.z, ay, rcy)), fminf(fmaf(cc.w, az, rcz), H));
                    const bool hl = tl <= fl + slack;
                    const bool hr = tr <= fr + slack;
                    if (COUNT) ctr.v[CTR_SLAB] += 2;
                    if (hl && hr) {
                        const bool swap = tr < tl;
                        stack[sp++] = swap ? ch.x : ch.y;
                        cur = swap ? ch.y : ch.x;
                    } else if (hl) {
                        cur = ch.x;
                    } else if (hr) {
                        cur = ch.y;
                    } else if (sp > 0) {
                        cur = stack[--sp];
                    } else {
                        st = 3;
                    }
                    if (st == 1 && cur < 0) st = 2;
                }
                const int n = __popc(__ballot_sync(FULL, st == 1));
                if (2 * n < live) break;
            }
        } else {
            // ================= LEAF pool: FILTER-domain bounds, candidate list =================
            if (st == 2) {
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
                    if (ln == 3) {  // list full (rare): resolve exactly now, keep the winner as a certain entry
                        Hit best;
                        exact_resolve<COUNT>(sc, o, d, 3, l0, l1, l2, &best, &ctr);
                        ln = 0;
                        if (best.pid >= 0) {
                            l0 = best.pid;
                            f0 = best.dist;
                            ln = 1;
                            if (best.dist < H) H = best.dist;
                        }
                        if (lo <= H) {
                            if (ln == 0) { l0 = pid; f0 = lo; } else { l1 = pid; f1 = lo; }
                            ln++;
                        }
                    } else {
                        if (ln == 0) { l0 = pid; f0 = lo; } else if (ln == 1) { l1 = pid; f1 = lo; } else { l2 = pid; f2 = lo; }
                        ln++;
                    }
                }
                if (sp > 0) {
                    cur = stack[--sp];
                    st = cur >= 0 ? 1 : 2;
                } else {
                    st = 3;
                }
            }
        }
    }

    if (COUNT) {
#pragma unroll
        for (int i = 0; i < NUM_COUNTERS; i++) {
            if (i == CTR_RAYS) continue;
            unsigned long long v = ctr.v[i];
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) v += __shfl_down_sync(FULL, v, ofs);
            if (lane == 0 && v) atomicAdd(&pr.counters[i], v);
        }
    }
}

}  // namespace rtb
