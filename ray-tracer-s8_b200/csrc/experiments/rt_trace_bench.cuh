// rt_trace_bench.cuh — measurement aid (included by rt_kernels.cu): the nearest-hit query ALONE, over a recorded set of
// rays, in two forms.  It answers one planning question: how much faster than the megakernel's traversal
// could the TRACE kernel of a two-kernel wavefront be?
//
//   record   the lanes kernel (COUNT instantiation) appends (origin, direction) of every query it makes to a buffer,
//            in the order its warps make them — primary and bounce rays of neighbouring pixels interleaved, which is
//            the order a ray queue would hold them in
//   tb_ww    the megakernel's own traversal (trace_bvh_ch: while-while, 32 rays per warp at a time, warps pull
//            batches of 32 from a ticket counter): what the product does today, minus everything that is not the query
//   tb_sm    the ballot-scheduled state machine of rt_kernel_wq.cuh (slab steps in bursts; filter, exact test and
//            retire + refill served when enough lanes wait) with dynamic fetch from ONE global queue
//
// Both read the scene from shared memory (one 768-thread CTA per SM) and write (pid, length(p - o)) per ray; the host
// compares the two outputs.  The big primitives of the split layout are tested by both at the start of a ray (tb_ww: by
// all 32 lanes together; tb_sm: at refill, by the lanes that refill) or by neither (with_big = 0: the tree alone).
#pragma once

namespace rtb {

struct TbArgs {
    const float4* rays;   // 2 per ray: (o.xyz, d.x) (d.y, d.z, -, -)
    unsigned long long n;
    unsigned long long* ticket;  // zeroed before launch
    int2* out;            // pid, dist bits
    uint32_t node_burst, t_leaf, t_pend, t_fin;
    uint32_t alt;         // tb_ww: 1 = trace_bvh_smem_stack
    uint32_t sstack_off;  // ... its stack area, in ints from the start of dynamic shared memory
};

template <int NT>
__device__ __forceinline__ void tb_stage(const DevScene& sc, SceneView& sv, float4* smem) {
    float4* p = smem;
    float4* s_sph = p;  p += sc.ns;
    float4* s_tri = p;  p += 4 * sc.nt;
    float4* s_na = p;   p += 3 * sc.lni;
    int2* s_nd = reinterpret_cast<int2*>(p);
    for (uint32_t i = threadIdx.x; i < sc.ns; i += NT) s_sph[i] = __ldg(&sc.sph[i]);
    for (uint32_t i = threadIdx.x; i < 4 * sc.nt; i += NT) s_tri[i] = __ldg(&sc.tri[i]);
    for (uint32_t i = threadIdx.x; i < 3 * sc.lni; i += NT) s_na[i] = __ldg(&sc.lnode_a[i]);
    for (uint32_t i = threadIdx.x; i < sc.lni; i += NT) s_nd[i] = __ldg(&sc.lnode_d[i]);
    __syncthreads();
    sv.sph2 = sc.sph2;
    sv.sph = s_sph; sv.tri = s_tri; sv.na = s_na; sv.nb = nullptr; sv.nc = nullptr; sv.nd = s_nd;
    // the product's traversal (64-byte records, rt_device.cuh) reads its nodes from global memory here
    sv.nodes_s = 0;
    sv.nodes_g = reinterpret_cast<const char*>(sc.lnode);
}

// Experimental traversal for tb_ww (TbArgs.alt = 1): the product's traversal with the stack in shared memory
// (entry e of thread t at sstack[e * 768 + t]: conflict-free) instead of local memory.  (alt = 2 was the sphere FILTER
// inside the node loop: 32.7 ms against 30.8 ms on C3, see profiles/r1_notes.md.)
constexpr int TB_SSTACK = 24;
__device__ __forceinline__ void trace_bvh_smem_stack(const DevScene& sc, const SceneView& sv, V3 o, V3 d, Hit& best, Ctr& ctr,
                                                     int* sstack) {
    best.pid = -1;
    best.dist = 0.0f;
    const float BIG = 1e30f;
    float ix = fminf(fmaxf(__frcp_rn(d.x), -BIG), BIG);
    float iy = fminf(fmaxf(__frcp_rn(d.y), -BIG), BIG);
    float iz = fminf(fmaxf(__frcp_rn(d.z), -BIG), BIG);
    if (!(fabsf(d.x) > 0.0f)) ix = BIG;
    if (!(fabsf(d.y) > 0.0f)) iy = BIG;
    if (!(fabsf(d.z) > 0.0f)) iz = BIG;
    const float ax = fabsf(ix), ay = fabsf(iy), az = fabsf(iz);
    const float qx = -o.x * ix, qy = -o.y * iy, qz = -o.z * iz;
    const float slack = 4.8e-7f * fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + 1e-30f;
    float cull = 1001.0f;
    int* st = sstack + threadIdx.x;
    int sp = 0;  // in units of 768 ints
    int cur = sc.lroot;
    const int ns = (int)sc.ns;
    for (uint32_t i = 0; i < sc.nbig; i++) {
        const int pid = (int)sc.big_pid[i];
        if (pid < ns) test_sphere<false>(sc, sv.sph[pid], pid, o, d, best, ctr);
        else if (triangle_filter(sv.tri, pid - ns, o, d, cull)) triangle_exact<false>(sc, sv.tri, pid - ns, pid, o, d, best, ctr);
        if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
    }
    if (!sc.ltree) return;
    for (;;) {
        while (cur >= 0) {
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
            const bool swap = tr < tl;
            if (hl && hr) {
                st[sp] = swap ? ch.x : ch.y;
                sp += 768;
            }
            int nxt = (hr && (!hl || swap)) ? ch.y : ch.x;
            if (!(hl || hr)) {
                nxt = TR_DONE;
                if (sp != 0) {
                    sp -= 768;
                    nxt = st[sp];
                }
            }
            cur = nxt;
        }
        if (cur == TR_DONE) return;
        const int first = (~cur) >> 5, count = ((~cur) & 31) + 1;
        if (first < ns) {
            for (int i = 0; i < count; i++) test_sphere<false>(sc, sv.sph[first + i], first + i, o, d, best, ctr);
        } else {
            for (int i = 0; i < count; i++) {
                const int pid = first + i;
                if (triangle_filter(sv.tri, pid - ns, o, d, cull)) triangle_exact<false>(sc, sv.tri, pid - ns, pid, o, d, best, ctr);
            }
        }
        if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
        if (sp == 0) return;
        sp -= 768;
        cur = st[sp];
    }
}

// Experimental traversal for tb_ww (TbArgs.alt = 2): the 4-ary collapse of the same tree (DevScene::w4, staged in shared
// memory in place of the binary records).  One visit = four slab tests, the hit children sorted by entry distance with
// a 5-exchange network, the nearest entered, the others pushed far to near.
__device__ __forceinline__ void trace_bvh4(const DevScene& sc, const SceneView& sv, const float4* w4, V3 o, V3 d, Hit& best, Ctr& ctr) {
    best.pid = -1;
    best.dist = 0.0f;
    const float BIG = 1e30f;
    float ix = fminf(fmaxf(__frcp_rn(d.x), -BIG), BIG);
    float iy = fminf(fmaxf(__frcp_rn(d.y), -BIG), BIG);
    float iz = fminf(fmaxf(__frcp_rn(d.z), -BIG), BIG);
    if (!(fabsf(d.x) > 0.0f)) ix = BIG;
    if (!(fabsf(d.y) > 0.0f)) iy = BIG;
    if (!(fabsf(d.z) > 0.0f)) iz = BIG;
    const float ax = fabsf(ix), ay = fabsf(iy), az = fabsf(iz);
    const float qx = -o.x * ix, qy = -o.y * iy, qz = -o.z * iz;
    const float slack = 4.8e-7f * fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + 1e-30f;
    float cull = 1001.0f;
    const int ns = (int)sc.ns;
    const HitCtx hc = hit_ctx(sc);
    int stack[MAX_STACK * 3 + 1 + MAX_BIG];
    stack[0] = TR_DONE;
    int* top = stack + 1;
    if (sc.ltree) *top++ = sc.w4root;
#pragma unroll 1
    for (int i = (int)sc.nbig - 1; i >= 0; i--) *top++ = __ldg(&sc.big_code[i]);
    int cur = *--top;
    const float INF = __int_as_float(0x7f800000);
    for (;;) {
        while (cur >= 0) {
            const float4* r = w4 + 7 * cur;
            const float4 a = r[0], b = r[1], c = r[2], e = r[3], f = r[4], g = r[5];
            const int4 ch = *reinterpret_cast<const int4*>(r + 6);
            auto slab1 = [&](float cx, float cy, float cz, float hx, float hy, float hz) -> float {
                const float tx = fmaf(cx, ix, qx), ty = fmaf(cy, iy, qy), tz = fmaf(cz, iz, qz);
                const float tn = fmaxf(fmaxf(fmaf(-hx, ax, tx), fmaf(-hy, ay, ty)), fmaxf(fmaf(-hz, az, tz), 0.0f));
                const float tf = fminf(fminf(fmaf(hx, ax, tx), fmaf(hy, ay, ty)), fminf(fmaf(hz, az, tz), cull));
                return (tn <= tf + slack) ? tn : INF;
            };
            float k0 = slab1(a.x, a.y, a.z, a.w, b.x, b.y);
            float k1 = slab1(b.z, b.w, c.x, c.y, c.z, c.w);
            float k2 = slab1(e.x, e.y, e.z, e.w, f.x, f.y);
            float k3 = slab1(f.z, f.w, g.x, g.y, g.z, g.w);
            int c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
            auto cswap = [](float& ka, int& ca, float& kb, int& cb) {
                const bool sw = kb < ka;
                const float kt = sw ? kb : ka, ku = sw ? ka : kb;
                const int ct = sw ? cb : ca, cu = sw ? ca : cb;
                ka = kt; kb = ku; ca = ct; cb = cu;
            };
            cswap(k0, c0, k1, c1); cswap(k2, c2, k3, c3); cswap(k0, c0, k2, c2); cswap(k1, c1, k3, c3); cswap(k1, c1, k2, c2);
            if (k3 < INF) *top++ = c3;
            if (k2 < INF) *top++ = c2;
            if (k1 < INF) *top++ = c1;
            int nxt = c0;
            if (!(k0 < INF)) nxt = *--top;
            cur = nxt;
        }
        if (cur == TR_DONE) return;
        const int pid = (~cur) >> 5;
        float t = 0.0f;
        bool cand = false;
        if (pid < ns) {
            const float4 s = sv.sph[pid];
            const V3 oc = mk(x_sub(o.x, s.x), x_sub(o.y, s.y), x_sub(o.z, s.z));
            const float bh = fmaf(oc.z, d.z, fmaf(oc.y, d.y, oc.x * d.x));
            const float oc2 = fmaf(oc.z, oc.z, fmaf(oc.y, oc.y, oc.x * oc.x));
            const float cf = oc2 - s.w;
            const float disc = fmaf(bh, bh, -cf);
            if (!(fmaf(oc2, 2e-5f, disc) < 0.0f) && !(bh > 0.0f && cf > 1e-4f * oc2)) cand = sphere_root_exact(d, oc, s.w, &t);
        } else if (triangle_filter(sv.tri, pid - ns, o, d, cull)) {
            const V3 ta = ld3(sv.tri[4 * (pid - ns) + 0]), ab = ld3(sv.tri[4 * (pid - ns) + 1]), ac = ld3(sv.tri[4 * (pid - ns) + 2]);
            int stage;
            cand = triangle_root_exact(o, d, ta, ab, ac, &t, &stage);
        }
        if (cand) {
            consider(hc, o, d, t, pid, best);
            cull = fmaf(best.dist, 1.00001f, 1e-6f);
        }
        cur = *--top;
    }
}

__global__ void __launch_bounds__(768, 1) tb_ww(const DevScene sc, const TbArgs a) {
    extern __shared__ float4 smem_dyn[];
    SceneView sv;
    const float4* s_w4 = nullptr;
    if (a.alt == 2) {  // geometry + the 4-ary records (in place of the binary ones)
        float4* p = smem_dyn;
        float4* s_sph = p;  p += sc.ns;
        float4* s_tri = p;  p += 4 * sc.nt;
        for (uint32_t i = threadIdx.x; i < sc.ns; i += 768) s_sph[i] = __ldg(&sc.sph[i]);
        for (uint32_t i = threadIdx.x; i < 4 * sc.nt; i += 768) s_tri[i] = __ldg(&sc.tri[i]);
        for (uint32_t i = threadIdx.x; i < 7 * sc.w4n; i += 768) p[i] = __ldg(&sc.w4[i]);
        __syncthreads();
        sv.sph2 = sc.sph2; sv.sph = s_sph; sv.tri = s_tri; sv.na = nullptr; sv.nb = nullptr; sv.nc = nullptr; sv.nd = nullptr;
        s_w4 = p;
    } else {
        tb_stage<768>(sc, sv, smem_dyn);
    }
    const int lane = threadIdx.x & 31;
    Ctr ctr;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.ticket, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= a.n) break;
        const unsigned long long i = base + lane;
        if (i < a.n) {
            const float4 r0 = a.rays[2 * i], r1 = a.rays[2 * i + 1];
            Hit h;
            if (a.alt == 2) {
                trace_bvh4(sc, sv, s_w4, mk(r0.x, r0.y, r0.z), mk(r0.w, r1.x, r1.y), h, ctr);
                hit_finish(h);
            } else if (a.alt) trace_bvh_smem_stack(sc, sv, mk(r0.x, r0.y, r0.z), mk(r0.w, r1.x, r1.y), h, ctr, a.alt ? reinterpret_cast<int*>(smem_dyn) + a.sstack_off : nullptr);
            else trace_bvh_ch<false, true, false>(sc, sv, mk(r0.x, r0.y, r0.z), mk(r0.w, r1.x, r1.y), h, ctr);  // nbig == 0 skips the list
            a.out[i] = make_int2(h.pid, h.pid >= 0 ? __float_as_int(h.dist) : 0);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(768, 1) tb_sm(const DevScene sc, const TbArgs a) {
    extern __shared__ float4 smem_dyn[];
    SceneView sv;
    tb_stage<768>(sc, sv, smem_dyn);
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int ns = (int)sc.ns;
    Ctr ctr;

    // per-lane ray + traversal state
    long long ray_id = -1;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    int cur = WQ_FIN, sp = 0;
    bool pend = false;
    int stack[MAX_STACK];
    float ix = 0, iy = 0, iz = 0, qx = 0, qy = 0, qz = 0, slack = 0, cull = 0;
    Hit best;
    best.pid = -1;
    best.dist = 0.0f;
    best.p = mk(0, 0, 0);
    bool queue_dry = false;  // warp-uniform

    for (;;) {
        // ---- slab steps ----
        {
            const float ax = fabsf(ix), ay = fabsf(iy), az = fabsf(iz);
            uint32_t k = 0;
            while (__ballot_sync(FULL, cur >= 0) != 0 && k < a.node_burst) {
                k++;
                if (cur >= 0) {
                    const float4* nrec = sv.na + 3 * cur;
                    const float4 na = nrec[0], nb = nrec[1], nc = nrec[2];
                    const int2 ch = sv.nd[cur];
                    const float lcx = fmaf(na.x, ix, qx), lcy = fmaf(na.y, iy, qy), lcz = fmaf(na.z, iz, qz);
                    const float rcx = fmaf(nb.z, ix, qx), rcy = fmaf(nb.w, iy, qy), rcz = fmaf(nc.x, iz, qz);
                    const float tl = fmaxf(fmaxf(fmaf(-na.w, ax, lcx), fmaf(-nb.x, ay, lcy)), fmaxf(fmaf(-nb.y, az, lcz), 0.0f));
                    const float fl = fminf(fminf(fmaf(na.w, ax, lcx), fmaf(nb.x, ay, lcy)), fminf(fmaf(nb.y, az, lcz), cull));
                    const float tr = fmaxf(fmaxf(fmaf(-nc.y, ax, rcx), fmaf(-nc.z, ay, rcy)), fmaxf(fmaf(-nc.w, az, rcz), 0.0f));
                    const float fr = fminf(fminf(fmaf(nc.y, ax, rcx), fmaf(nc.z, ay, rcy)), fminf(fmaf(nc.w, az, rcz), cull));
                    const bool hl = tl <= fl + slack;
                    const bool hr = tr <= fr + slack;
                    const bool swap = tr < tl;
                    if (hl && hr) stack[sp++] = swap ? ch.x : ch.y;
                    int nxt = (hr && (!hl || swap)) ? ch.y : ch.x;
                    if (!(hl || hr)) {
                        nxt = WQ_FIN;
                        if (sp != 0) nxt = stack[--sp];
                    }
                    cur = nxt;
                }
            }
        }
        // ---- vote ----
        const bool has = ray_id >= 0;
        const unsigned nm = __ballot_sync(FULL, cur >= 0);
        const unsigned lm = __ballot_sync(FULL, cur < 0 && cur != WQ_FIN && !pend);
        const unsigned pm = __ballot_sync(FULL, pend);
        const unsigned fm = __ballot_sync(FULL, cur == WQ_FIN && (has || !queue_dry));
        const uint32_t n_node = __popc(nm), n_leaf = __popc(lm), n_pend = __popc(pm), n_fin = __popc(fm);
        if ((nm | lm | pm | fm) == 0) break;

        if (n_leaf != 0 && (n_leaf >= a.t_leaf || n_leaf >= n_node)) {
            if (cur < 0 && cur != WQ_FIN && !pend) {
                const int first = (~cur) >> 5;
                bool pass;
                if (first < ns) pass = sphere_filter(sv.sph[first], o, d);
                else pass = triangle_filter(sv.tri, first - ns, o, d, cull);
                if (pass) {
                    pend = true;
                } else if (((~cur) & 31) != 0) {
                    cur -= 31;
                } else {
                    cur = sp ? stack[--sp] : WQ_FIN;
                }
            }
        }
        if (n_pend != 0 && (n_pend >= a.t_pend || n_pend >= n_node)) {
            if (pend) {
                const int first = (~cur) >> 5;
                wq_exact<false>(sc, sv, first, o, d, best, ctr);
                if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
                pend = false;
                if (((~cur) & 31) != 0) {
                    cur -= 31;
                } else {
                    cur = sp ? stack[--sp] : WQ_FIN;
                }
            }
        }
        if (n_fin != 0 && (n_fin >= a.t_fin || n_fin >= n_node)) {
            // ---- retire + refill from the global queue ----
            if (has && cur == WQ_FIN) {
                a.out[ray_id] = make_int2(best.pid, best.pid >= 0 ? __float_as_int(best.dist) : 0);
                ray_id = -1;
            }
            const unsigned idle = __ballot_sync(FULL, ray_id < 0);
            if (!queue_dry) {
                unsigned long long base = 0;
                const uint32_t want = __popc(idle);
                if (lane == 0) base = atomicAdd(a.ticket, (unsigned long long)want);
                base = __shfl_sync(FULL, base, 0);
                if (base + want >= a.n) queue_dry = true;
                const unsigned long long mine = base + __popc(idle & lt_mask);
                if (ray_id < 0 && mine < a.n) {
                    ray_id = (long long)mine;
                    const float4 r0 = a.rays[2 * mine], r1 = a.rays[2 * mine + 1];
                    o = mk(r0.x, r0.y, r0.z);
                    d = mk(r0.w, r1.x, r1.y);
                    const float BIG = 1e30f;
                    ix = fminf(fmaxf(__frcp_rn(d.x), -BIG), BIG);
                    iy = fminf(fmaxf(__frcp_rn(d.y), -BIG), BIG);
                    iz = fminf(fmaxf(__frcp_rn(d.z), -BIG), BIG);
                    if (!(fabsf(d.x) > 0.0f)) ix = BIG;
                    if (!(fabsf(d.y) > 0.0f)) iy = BIG;
                    if (!(fabsf(d.z) > 0.0f)) iz = BIG;
                    qx = -o.x * ix; qy = -o.y * iy; qz = -o.z * iz;
                    slack = 4.8e-7f * fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + 1e-30f;
                    best.pid = -1;
                    best.dist = 0.0f;
                    cull = 1001.0f;
                    // the split layout's big primitives, as the wavefront's LOGIC kernel would have seeded them
                    for (uint32_t i = 0; i < sc.nbig; i++) {
                        const int pid = (int)sc.big_pid[i];
                        bool pass;
                        if (pid < ns) pass = sphere_filter(sv.sph[pid], o, d);
                        else pass = triangle_filter(sv.tri, pid - ns, o, d, cull);
                        if (pass) {
                            wq_exact<false>(sc, sv, pid, o, d, best, ctr);
                            if (best.pid >= 0) cull = fmaf(best.dist, 1.00001f, 1e-6f);
                        }
                    }
                    sp = 0;
                    pend = false;
                    cur = sc.ltree ? sc.lroot : WQ_FIN;
                }
            }
        }
    }
}

}  // namespace rtb
