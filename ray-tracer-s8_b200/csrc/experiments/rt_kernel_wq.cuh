// rt_kernel_wq.cuh — the render path as a WARP-PRIVATE WAVEFRONT inside one persistent launch (included by
// rt_kernels.cu).
//
// Why (profiles/r1_notes.md): the lanes megakernel keeps issue slots ~78 % full but only ~8 of 32 lanes do work —
// every query makes the warp wait for its longest traversal, and between queries the lanes split into sky / shade /
// new-sample / new-pixel branches.  The first wavefront (rt_wavefront.cuh) cured the lane counts and lost to its own
// plumbing: ~94 GB of HBM traffic per frame, two launches per round, global atomics, a pool-scheduled trace kernel
// with 2.4x the thread-instructions.
//
// Here each WARP owns C pixel chains ("chains": one pixel's sequential samples and bounces on its own xoshiro256++
// stream, main.rs:69-77) and alternates between two phases with no inter-warp synchronisation and no atomics
// except the tile ticket:
//
//   TRACE  the lanes kernel's traversal arithmetic, bit for bit, as a ballot-scheduled state machine with DYNAMIC
//          FETCH.  A lane is at an inner NODE, at a LEAF (cheap FILTER test pending), PENDing an exact test, FINished,
//          or idle; each trip the warp votes and runs the one step that serves the most lanes (weighted by cost):
//          a slab step, a filter step, the reference's exact roots + min_by, or retire + refill — a finished lane
//          writes its result (HBM/L2) and its chain id (hit or end list) and takes the next ray from the warp's queue,
//          a 32-byte shared-memory read.  When the queue is empty and fewer than T lanes still traverse, the phase
//          ends; the stragglers keep their traversal state in registers and resume later.
//   LOGIC  the finished chains, one list at a time so a warp runs one kind of work with all its lanes:
//          pass A (hits): shade + scatter (main.rs:116-133) → next ray into the queue, or the chain to the end list
//          pass B (ends: sky | emissive | depth exhausted | fresh chain): radiance, albedo fold, sample sum; pixel
//          store + next pixel of the warp's tile when the pixel is finished; Camera::get_ray for the next sample.
//
// Chain state lives in a warp-private slice of one HBM buffer (L2-resident: ~110 B per chain, ~50 MB per GPU at
// C = 128); rays and the three lists live in shared memory next to the scene, which one 768-thread CTA per SM stages
// once (the lanes kernel stages it three times per SM).  Results are bit-identical to the other kernels: the same
// device functions make the same decisions; only the order in which independent chains advance differs.
#pragma once

namespace rtb {

constexpr int WQ_MAX_CHAINS = 256;            // chain ids are bytes in the shared-memory lists
constexpr uint32_t WQ_FRESH = 0xffffffffu;    // meta of a chain that holds no pixel
constexpr int WQ_FIN = (int)0x80000000;       // traversal finished (never a leaf code: first_pid < 2^26)
constexpr int WQ_BIG = (int)0x80000001;       // pseudo-leaf: the scene's `big_pid` list (split layout)

struct WqArgs {
    void* base;            // chain state, see wq_state_bytes()
    unsigned long long n;  // total chains = grid * warps per CTA * chains
    uint32_t chains;       // C, per warp (multiple of 32, <= WQ_MAX_CHAINS)
    uint32_t min_active;   // T
    uint32_t min_node;     // the node step repeats without a new vote while at least this many lanes are at inner nodes
    uint32_t node_burst;   // slab steps between two votes
    uint32_t t_leaf, t_pend, t_fin;  // a waiting state is served when this many lanes (or more lanes than at nodes) are in it
    uint32_t scene_bytes;  // shared-memory offset of the per-warp areas
    uint32_t cta_phases;   // 1: LOGIC and TRACE are CTA-wide phases separated by barriers
    uint32_t trace_budget; // CTA-wide phases: a warp leaves TRACE after this many votes (0 = no bound)
};

__host__ __device__ inline size_t wq_state_bytes(size_t n, uint32_t depth) { return n * (16 * 4 + 4 + 4 + 4 * (size_t)depth); }
__host__ __device__ inline size_t wq_warp_smem(uint32_t chains) { return ((size_t)chains * (32 + 3) + 15) & ~(size_t)15; }

__device__ __forceinline__ uint32_t wq_meta(uint32_t s, uint32_t left, uint32_t np) { return (s << 16) | (left << 8) | np; }

// The scene's `big_pid` primitives (split layout) against a freshly generated ray, run in the LOGIC passes where all
// lanes work: the result seeds the traversal's nearest hit (ray words 6, 7) and the chain's result record.
// One out-of-line copy of the reference's exact test + min_by per primitive kind: the TRACE loop and both LOGIC
// passes call the same code, which keeps the kernel's hot footprint inside the instruction cache.
template <bool COUNT>
__device__ __noinline__ void wq_exact(const DevScene& sc, const SceneView& sv, int pid, V3 o, V3 d, Hit& best, Ctr& ctr) {
    const int ns = (int)sc.ns;
    if (pid < ns) sphere_exact<COUNT>(sc, sv.sph[pid], pid, o, d, best, ctr);
    else triangle_exact<COUNT>(sc, sv.tri, pid - ns, pid, o, d, best, ctr);
}

template <bool COUNT>
__device__ __noinline__ void wq_seed_ray(const DevScene& sc, const SceneView& sv, V3 o, V3 d, float* ray, uint32_t C,
                                            uint32_t c, float4* g_res, Ctr& ctr) {
    Hit b;
    b.pid = -1;
    b.dist = 0.0f;
    b.p = mk(0, 0, 0);
    float cull = 1001.0f;
    const int ns = (int)sc.ns;
    for (uint32_t i = 0; i < sc.nbig; i++) {
        const int pid = (int)sc.big_pid[i];
        bool pass;
        if (pid < ns) {
            if (COUNT) ctr.v[CTR_SPH_TEST]++;
            pass = sphere_filter(sv.sph[pid], o, d);
        } else {
            if (COUNT) ctr.v[CTR_TRI_TEST]++;
            pass = triangle_filter(sv.tri, pid - ns, o, d, cull);
        }
        if (pass) {
            wq_exact<COUNT>(sc, sv, pid, o, d, b, ctr);
            if (b.pid >= 0) cull = fmaf(b.dist, 1.00001f, 1e-6f);
        }
    }
    ray[0 * C + c] = o.x; ray[1 * C + c] = o.y; ray[2 * C + c] = o.z;
    ray[3 * C + c] = d.x; ray[4 * C + c] = d.y; ray[5 * C + c] = d.z;
    ray[6 * C + c] = b.dist;
    ray[7 * C + c] = __int_as_float(b.pid);
    g_res[c] = make_float4(b.p.x, b.p.y, b.p.z, __int_as_float(b.pid));
}

template <bool SMEM, bool COUNT, int NW>
__global__ void __launch_bounds__(NW * 32, 1) render_kernel_wq(const DevScene sc, const DevCamera cam, const DevParams pr,
                                                               const WqArgs wa) {
    extern __shared__ float4 smem_dyn[];
    constexpr int NT = NW * 32;
    SceneView sv;
    if (SMEM) {
        float4* p = smem_dyn;
        float4* s_sph = p;  p += sc.ns;
        float4* s_tri = p;  p += 4 * sc.nt;
        float4* s_na = p;   p += 3 * sc.lni;  // 48-byte node records
        int2* s_nd = reinterpret_cast<int2*>(p);
        for (uint32_t i = threadIdx.x; i < sc.ns; i += NT) s_sph[i] = __ldg(&sc.sph[i]);
        for (uint32_t i = threadIdx.x; i < 4 * sc.nt; i += NT) s_tri[i] = __ldg(&sc.tri[i]);
        for (uint32_t i = threadIdx.x; i < sc.lni; i += NT) {
            s_na[3 * i] = __ldg(&sc.lnode_a[3 * i]);
            s_na[3 * i + 1] = __ldg(&sc.lnode_a[3 * i + 1]);
            s_na[3 * i + 2] = __ldg(&sc.lnode_a[3 * i + 2]);
            s_nd[i] = __ldg(&sc.lnode_d[i]);
        }
        __syncthreads();
        sv.sph = s_sph; sv.tri = s_tri; sv.na = s_na; sv.nb = nullptr; sv.nc = nullptr; sv.nd = s_nd;
    } else {
        sv.sph = sc.sph; sv.tri = sc.tri; sv.na = sc.lnode_a; sv.nb = nullptr; sv.nc = nullptr; sv.nd = sc.lnode_d;
    }

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t C = wa.chains;
    const uint32_t total_tiles = pr.tiles_x * pr.tiles_y;
    const float spp_f = (float)pr.spp;
    const int ns = (int)sc.ns;

    // ---- the warp's shared-memory area: rays (SoA over chains) | ray queue (ring) | hit list | end list ----
    uint8_t* area = reinterpret_cast<uint8_t*>(smem_dyn) + (SMEM ? wa.scene_bytes : 0u) + (size_t)warp * wq_warp_smem(C);
    // ray[k * C + chain], k = ox oy oz dx dy dz | distance and pid of the nearest hit among the scene's big_pid list
    float* ray = reinterpret_cast<float*>(area);
    uint8_t* q_ray = area + (size_t)C * 32;
    uint8_t* l_hit = q_ray + C;
    uint8_t* l_end = l_hit + C;

    // ---- the warp's slice of the chain state ----
    const size_t N = (size_t)wa.n;
    const size_t w0 = ((size_t)blockIdx.x * NW + warp) * C;
    uint4* g_rng0 = reinterpret_cast<uint4*>(wa.base) + w0;
    uint4* g_rng1 = reinterpret_cast<uint4*>(wa.base) + N + w0;
    float4* g_acc = reinterpret_cast<float4*>(wa.base) + 2 * N + w0;  // sample sum r, g, b
    float4* g_res = reinterpret_cast<float4*>(wa.base) + 3 * N + w0;  // hit point, pid (bits; -1 = miss)
    uint32_t* g_meta = reinterpret_cast<uint32_t*>(reinterpret_cast<float4*>(wa.base) + 4 * N) + w0;  // s | left | np
    uint32_t* g_pix = reinterpret_cast<uint32_t*>(reinterpret_cast<float4*>(wa.base) + 4 * N) + N + w0;  // y*width + x
    uint32_t* g_path = reinterpret_cast<uint32_t*>(reinterpret_cast<float4*>(wa.base) + 4 * N) + 2 * N + w0;  // [level * N]

    Ctr ctr;
#pragma unroll
    for (int i = 0; i < NUM_COUNTERS; i++) ctr.v[i] = 0;
    unsigned long long rays = 0;

    // warp-uniform bookkeeping
    uint32_t qh = 0, qn = 0, n_hit = 0, n_end = C;
    uint32_t tile_next = TILE_W * TILE_H, tile_x0 = 0, tile_y0 = 0;
    bool tiles_left = true;
    for (uint32_t i = lane; i < C; i += 32) {
        l_end[i] = (uint8_t)i;
        g_meta[i] = WQ_FRESH;
    }
    __syncwarp();

    // per-lane traversal state (lives across LOGIC phases for the stragglers)
    int chain = -1, cur = WQ_FIN, sp = 0;
    bool pend = false;
    int stack[MAX_STACK];
    float ix = 0, iy = 0, iz = 0, qx = 0, qy = 0, qz = 0, slack = 0, cull = 0;
    Hit best;
    best.pid = -1;
    best.dist = 0.0f;
    best.p = mk(0, 0, 0);

    for (;;) {
        // =====================================================================================
        // LOGIC pass A: hits → shade + scatter (main.rs:116-133)
        // =====================================================================================
        for (uint32_t base = 0; base < n_hit; base += 32) {
            const uint32_t i = base + lane;
            bool to_ray = false, to_end = false;
            uint32_t c = 0;
            if (i < n_hit) {
                c = l_hit[i];
                const float4 r = g_res[c];
                const int pid = __float_as_int(r.w);
                const float e = __ldg(&sc.emis[pid]);
                if (e > 0.0f) {
                    to_end = true;  // emission * albedo: radiance is formed in pass B
                } else {
                    const float4 m = __ldg(&sc.mat[pid]);
                    const uint32_t meta = g_meta[c];
                    uint32_t left = (meta >> 8) & 0xffu, np = meta & 0xffu;
                    Rng rng;
                    {
                        const uint4 a = g_rng0[c], b = g_rng1[c];
                        rng.s0 = ((uint64_t)a.y << 32) | a.x; rng.s1 = ((uint64_t)a.w << 32) | a.z;
                        rng.s2 = ((uint64_t)b.y << 32) | b.x; rng.s3 = ((uint64_t)b.w << 32) | b.z;
                    }
                    const V3 d = mk(ray[3 * C + c], ray[4 * C + c], ray[5 * C + c]);
                    const V3 hp = mk(r.x, r.y, r.z);
                    V3 n;
                    if (COUNT) ctr.v[pid < ns ? CTR_SHADE_SPH : CTR_SHADE_TRI]++;
                    if (pid < ns) {
                        n = x_normalize_or_zero(x_sub(hp, ld3(sv.sph[pid])));  // sphere.rs:49-51
                    } else {
                        n = ld3(sv.tri[4 * (pid - ns) + 3]);                   // mesh.rs:163-165 (host, same ops)
                    }
                    const V3 diffuse = x_add(unit_sphere(rng), n);
                    const float kk = x_mul(2.0f, x_dot(d, n));
                    const V3 glossy = x_sub(d, x_scale(n, kk));
                    const V3 scat = x_add(diffuse, x_scale(x_sub(glossy, diffuse), m.w));
                    V3 nd;
                    if (!x_try_normalize(scat, &nd)) nd = n;
                    const V3 d2 = x_normalize_div(nd);  // Ray::new
                    g_path[(size_t)np * N + c] = (uint32_t)pid;
                    np++;
                    left--;
                    g_meta[c] = (meta & 0xffff0000u) | (left << 8) | np;
                    g_rng0[c] = make_uint4((uint32_t)rng.s0, (uint32_t)(rng.s0 >> 32), (uint32_t)rng.s1, (uint32_t)(rng.s1 >> 32));
                    g_rng1[c] = make_uint4((uint32_t)rng.s2, (uint32_t)(rng.s2 >> 32), (uint32_t)rng.s3, (uint32_t)(rng.s3 >> 32));
                    if (left == 0) {
                        to_end = true;  // the next call has depth == 0 → BLACK, no query
                    } else {
                        wq_seed_ray<COUNT>(sc, sv, hp, d2, ray, C, c, g_res, ctr);
                        to_ray = true;
                        rays++;
                    }
                }
            }
            const unsigned me = __ballot_sync(FULL, to_end);
            if (to_end) l_end[n_end + __popc(me & lt_mask)] = (uint8_t)c;
            n_end += __popc(me);
            const unsigned mr = __ballot_sync(FULL, to_ray);
            if (to_ray) {
                uint32_t idx = qh + qn + __popc(mr & lt_mask);
                if (idx >= C) idx -= C;
                q_ray[idx] = (uint8_t)c;
            }
            qn += __popc(mr);
        }
        n_hit = 0;
        __syncwarp();

        // =====================================================================================
        // LOGIC pass B: sample ends (sky | emissive | exhausted) and fresh chains → next sample / next pixel
        // =====================================================================================
        for (uint32_t base = 0; base < n_end; base += 32) {
            const uint32_t i = base + lane;
            const bool on = i < n_end;
            uint32_t c = 0, s = 0, px = 0, py = 0;
            float sr = 0.0f, sg = 0.0f, sb = 0.0f;
            Rng rng;
            rng.s0 = rng.s1 = rng.s2 = rng.s3 = 0;
            bool need = false, alive = false;
            if (on) {
                c = l_end[i];
                const uint32_t meta = g_meta[c];
                if (meta != WQ_FRESH) {
                    const float4 acc = g_acc[c];
                    const uint32_t pix = g_pix[c];
                    py = pix / pr.width;
                    px = pix - py * pr.width;
                    {
                        const uint4 a = g_rng0[c], b = g_rng1[c];
                        rng.s0 = ((uint64_t)a.y << 32) | a.x; rng.s1 = ((uint64_t)a.w << 32) | a.z;
                        rng.s2 = ((uint64_t)b.y << 32) | b.x; rng.s3 = ((uint64_t)b.w << 32) | b.z;
                    }
                    const int pid = __float_as_int(g_res[c].w);
                    float Lr, Lg, Lb;
                    if (pid < 0) {  // sky (main.rs:135-144)
                        if (COUNT) ctr.v[CTR_SKY]++;
                        const V3 d = mk(ray[3 * C + c], ray[4 * C + c], ray[5 * C + c]);
                        const float rcp = x_div(1.0f, x_length(d));
                        const float ny = (isfinite(rcp) && rcp > 0.0f) ? x_mul(d.y, rcp) : 0.0f;
                        const float t = x_add(x_mul(ny, 0.5f), 1.0f);
                        const float k1 = x_sub(1.0f, t);
                        const float w = x_mul(1.0f, t);
                        Lr = x_add(w, x_mul(0.3f, k1));
                        Lg = Lr;
                        Lb = x_add(w, x_mul(0.8f, k1));
                    } else {
                        const float e = __ldg(&sc.emis[pid]);
                        if (e > 0.0f) {  // emission * albedo (main.rs:116-117)
                            if (COUNT) ctr.v[CTR_EMISSIVE]++;
                            const float4 m = __ldg(&sc.mat[pid]);
                            Lr = x_mul(m.x, e); Lg = x_mul(m.y, e); Lb = x_mul(m.z, e);
                        } else {
                            Lr = Lg = Lb = 0.0f;
                        }
                    }
                    // fold albedo ⊙ (albedo ⊙ (... ⊙ L)) innermost first, like the recursion unwinding
                    for (uint32_t k = meta & 0xffu; k-- > 0;) {
                        const float4 m = __ldg(&sc.mat[g_path[(size_t)k * N + c]]);
                        Lr = x_mul(m.x, Lr); Lg = x_mul(m.y, Lg); Lb = x_mul(m.z, Lb);
                    }
                    sr = x_add(acc.x, Lr); sg = x_add(acc.y, Lg); sb = x_add(acc.z, Lb);
                    s = (meta >> 16) + 1;
                    if (s == pr.spp) {  // pixel finished (main.rs:78-81)
                        const size_t off = ((size_t)(py - pr.out_row0) * pr.width + px) * 3;
                        pr.out[off + 0] = (uint8_t)quantise(sr, spp_f);
                        pr.out[off + 1] = (uint8_t)quantise(sg, spp_f);
                        pr.out[off + 2] = (uint8_t)quantise(sb, spp_f);
                        need = true;
                    } else {
                        alive = true;
                    }
                } else {
                    need = true;
                }
            }
            // ---- hand out pixels: warp-cooperative, tile by tile (same ticket map as the lanes kernel) ----
            unsigned want = __ballot_sync(FULL, need);
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
                        break;
                    }
                    if (pr.tile_order_reverse) g = total_tiles - 1 - g;
                    tile_x0 = (uint32_t)(g % pr.tiles_x) * TILE_W;
                    tile_y0 = pr.row0 + (uint32_t)(g / pr.tiles_x) * TILE_H;
                    tile_next = 0;
                }
                const uint32_t avail = TILE_W * TILE_H - tile_next;
                const uint32_t my = __popc(want & lt_mask);
                if (need && my < avail) {
                    const uint32_t j = tile_next + my;
                    const uint32_t x = tile_x0 + (j & (TILE_W - 1)), y = tile_y0 + (j / TILE_W);
                    if (x < pr.width && y < pr.row1) {  // tiles on the right/bottom edge are partial
                        px = x; py = y;
                        need = false;
                        alive = true;
                        rng.seed_from_u64(pr.seed + ((uint64_t)y * pr.width + x));
                        sr = sg = sb = 0.0f;
                        s = 0;
                        g_pix[c] = y * pr.width + x;
                    }
                }
                tile_next += min((uint32_t)__popc(want), avail);
                want = __ballot_sync(FULL, need);
            }
            if (alive) {  // start sample s
                V3 o, d;
                primary_ray(cam, px, pr.height - py - 1, rng, &o, &d);  // y_cam = h - y - 1 (main.rs:71)
                wq_seed_ray<COUNT>(sc, sv, o, d, ray, C, c, g_res, ctr);
                g_acc[c] = make_float4(sr, sg, sb, 0.0f);
                g_meta[c] = wq_meta(s, pr.depth, 0);
                g_rng0[c] = make_uint4((uint32_t)rng.s0, (uint32_t)(rng.s0 >> 32), (uint32_t)rng.s1, (uint32_t)(rng.s1 >> 32));
                g_rng1[c] = make_uint4((uint32_t)rng.s2, (uint32_t)(rng.s2 >> 32), (uint32_t)rng.s3, (uint32_t)(rng.s3 >> 32));
                rays++;
            }
            const unsigned mr = __ballot_sync(FULL, alive);
            if (alive) {
                uint32_t idx = qh + qn + __popc(mr & lt_mask);
                if (idx >= C) idx -= C;
                q_ray[idx] = (uint8_t)c;
            }
            qn += __popc(mr);
        }
        n_end = 0;
        __syncwarp();

        const bool warp_done = qn == 0 && __ballot_sync(FULL, chain >= 0) == 0;  // every chain of this warp is retired
        if (wa.cta_phases) {
            // CTA-wide phases: all warps run LOGIC together, then TRACE together, so each phase's code is what the SM's
            // instruction caches hold; a finished warp keeps meeting the barriers until the whole CTA is done
            if (__syncthreads_and(warp_done)) break;
        } else if (warp_done) {
            break;
        }
        uint32_t trips = 0;

        // =====================================================================================
        // TRACE: slab steps for the lanes at inner nodes; the other states are served when enough lanes wait in them
        // =====================================================================================
        for (;;) {
            // ---- up to `node_burst` slab steps between votes ----
            {
                const float ax = fabsf(ix), ay = fabsf(iy), az = fabsf(iz);
                uint32_t k = 0;
                while (__ballot_sync(FULL, cur >= 0) != 0 && k < wa.node_burst) {
                    k++;
                    if (cur >= 0) {
                        const float4* nrec = sv.na + 3 * cur;
                        const float4 a = nrec[0], b = nrec[1], c = nrec[2];
                        const int2 ch = sv.nd[cur];
                        const float lcx = fmaf(a.x, ix, qx), lcy = fmaf(a.y, iy, qy), lcz = fmaf(a.z, iz, qz);
                        const float rcx = fmaf(b.z, ix, qx), rcy = fmaf(b.w, iy, qy), rcz = fmaf(c.x, iz, qz);
                        const float tl = fmaxf(fmaxf(fmaf(-a.w, ax, lcx), fmaf(-b.x, ay, lcy)), fmaxf(fmaf(-b.y, az, lcz), 0.0f));
                        const float fl = fminf(fminf(fmaf(a.w, ax, lcx), fmaf(b.x, ay, lcy)), fminf(fmaf(b.y, az, lcz), cull));
                        const float tr = fmaxf(fmaxf(fmaf(-c.y, ax, rcx), fmaf(-c.z, ay, rcy)), fmaxf(fmaf(-c.w, az, rcz), 0.0f));
                        const float fr = fminf(fminf(fmaf(c.y, ax, rcx), fmaf(c.z, ay, rcy)), fminf(fmaf(c.w, az, rcz), cull));
                        const bool hl = tl <= fl + slack;
                        const bool hr = tr <= fr + slack;
                        if (COUNT) ctr.v[CTR_SLAB] += 2;
                        const bool swap = tr < tl;
                        const bool right_first = hr && (!hl || swap);  // right child is the nearer (or the only) one
                        if (hl && hr) stack[sp++] = swap ? ch.x : ch.y;
                        int nxt = right_first ? ch.y : ch.x;
                        const bool none = !(hl || hr);
                        if (none) nxt = WQ_FIN;
                        const bool pop = none && sp != 0;
                        sp -= pop ? 1 : 0;
                        if (pop) nxt = stack[sp];
                        cur = nxt;
                    }
                }
            }
            // ---- vote (an idle lane has cur == WQ_FIN and chain < 0) ----
            const bool has = chain >= 0;
            const unsigned nm = __ballot_sync(FULL, cur >= 0);
            const unsigned lm = __ballot_sync(FULL, cur < 0 && cur != WQ_FIN && !pend);
            const unsigned pm = __ballot_sync(FULL, pend);
            const unsigned fm = __ballot_sync(FULL, cur == WQ_FIN && (has || qn != 0));
            const uint32_t n_node = __popc(nm), n_leaf = __popc(lm), n_pend = __popc(pm), n_fin = __popc(fm);
            if (COUNT) {
                if (cur >= 0) ctr.v[CTR_ACTIVE_LANES]++;
                if (lane == 0) ctr.v[CTR_TOTAL_LANES] += 32;
            }
            if ((nm | lm | pm | fm) == 0) break;  // nothing in flight, nothing queued
            if (wa.trace_budget && ++trips > wa.trace_budget) break;  // CTA-wide phases: bound the wait at the barrier

            if (n_leaf != 0 && (n_leaf >= wa.t_leaf || n_leaf >= n_node)) {
                // ---- FILTER step on the first primitive of the leaf: code = ~((first << 5) | (count - 1)) ----
                if (cur < 0 && cur != WQ_FIN && !pend) {
                    const V3 o = mk(ray[0 * C + chain], ray[1 * C + chain], ray[2 * C + chain]);
                    const V3 d = mk(ray[3 * C + chain], ray[4 * C + chain], ray[5 * C + chain]);
                    const int first = (~cur) >> 5;
                    bool pass;
                    if (first < ns) {
                        if (COUNT) ctr.v[CTR_SPH_TEST]++;
                        pass = sphere_filter(sv.sph[first], o, d);
                    } else {
                        if (COUNT) ctr.v[CTR_TRI_TEST]++;
                        pass = triangle_filter(sv.tri, first - ns, o, d, cull);
                    }
                    if (pass) {
                        pend = true;
                    } else if (((~cur) & 31) != 0) {
                        cur -= 31;  // first + 1, count - 1
                    } else {
                        cur = sp ? stack[--sp] : WQ_FIN;
                    }
                }
            }
            if (n_pend != 0 && (n_pend >= wa.t_pend || n_pend >= n_node)) {
                // ---- the reference's exact test + min_by on the pending primitive ----
                if (pend) {
                    const V3 o = mk(ray[0 * C + chain], ray[1 * C + chain], ray[2 * C + chain]);
                    const V3 d = mk(ray[3 * C + chain], ray[4 * C + chain], ray[5 * C + chain]);
                    const int first = (~cur) >> 5;
                    wq_exact<COUNT>(sc, sv, first, o, d, best, ctr);
                    if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
                    pend = false;
                    if (((~cur) & 31) != 0) {
                        cur -= 31;
                    } else {
                        cur = sp ? stack[--sp] : WQ_FIN;
                    }
                }
            }
            if (n_fin != 0 && (n_fin >= wa.t_fin || n_fin >= n_node)) {
                // ---- retire finished rays, refill idle lanes ----
                const bool fin = has && cur == WQ_FIN;
                bool is_hit = false;
                if (fin) {
                    is_hit = best.pid >= 0;
                    // the record already holds the seed hit (wq_seed_ray); rewrite it only when the tree found a nearer one
                    if (best.pid != __float_as_int(ray[7 * C + chain]))
                        g_res[chain] = make_float4(best.p.x, best.p.y, best.p.z, __int_as_float(best.pid));
                }
                const unsigned mh = __ballot_sync(FULL, is_hit);
                if (is_hit) l_hit[n_hit + __popc(mh & lt_mask)] = (uint8_t)chain;
                n_hit += __popc(mh);
                const unsigned mm = __ballot_sync(FULL, fin && !is_hit);
                if (fin && !is_hit) l_end[n_end + __popc(mm & lt_mask)] = (uint8_t)chain;
                n_end += __popc(mm);
                if (fin) chain = -1;
                const unsigned idle = __ballot_sync(FULL, chain < 0);
                if (qn != 0) {
                    const uint32_t my = __popc(idle & lt_mask);
                    if (chain < 0 && my < qn) {
                        uint32_t idx = qh + my;
                        if (idx >= C) idx -= C;
                        chain = q_ray[idx];
                        const V3 o = mk(ray[0 * C + chain], ray[1 * C + chain], ray[2 * C + chain]);
                        const V3 d = mk(ray[3 * C + chain], ray[4 * C + chain], ray[5 * C + chain]);
                        // FILTER-domain ray constants; |1/d| is clamped so 0*inf never produces NaN slabs
                        const float BIG = 1e30f;
                        ix = fminf(fmaxf(__frcp_rn(d.x), -BIG), BIG);
                        iy = fminf(fmaxf(__frcp_rn(d.y), -BIG), BIG);
                        iz = fminf(fmaxf(__frcp_rn(d.z), -BIG), BIG);
                        if (!(fabsf(d.x) > 0.0f)) ix = BIG;
                        if (!(fabsf(d.y) > 0.0f)) iy = BIG;
                        if (!(fabsf(d.z) > 0.0f)) iz = BIG;
                        qx = -o.x * ix; qy = -o.y * iy; qz = -o.z * iz;
                        slack = 4.8e-7f * fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + 1e-30f;
                        best.dist = ray[6 * C + chain];
                        best.pid = __float_as_int(ray[7 * C + chain]);
                        cull = best.pid >= 0 ? fmaf(best.dist, 1.00001f, 1e-6f) : 1001.0f;  // a hit has t < T_MAX, length(p-o) ~ t
                        sp = 0;
                        cur = sc.ltree ? sc.lroot : WQ_FIN;
                    }
                    const uint32_t take = min((uint32_t)__popc(idle), qn);
                    qh += take;
                    if (qh >= C) qh -= C;
                    qn -= take;
                }
                // hand the finished chains to LOGIC when the queue is dry and few lanes still traverse
                const unsigned act2 = __ballot_sync(FULL, chain >= 0);
                if (qn == 0 && (uint32_t)__popc(act2) < wa.min_active && (n_hit + n_end) != 0) break;
            }
        }
        __syncwarp();
        if (wa.cta_phases) __syncthreads();
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
