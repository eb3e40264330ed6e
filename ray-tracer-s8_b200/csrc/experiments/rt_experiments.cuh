// rt_experiments.cuh — the A/B kernels of round 1 and their launchers (included by rt_kernels.cu only when built with
// -DRT_B200_EXPERIMENTS → lib/librt_b200_exp.so).  None of this is in the product library; what each variant
// measured is in DESIGN.md section 4 and profiles/r1_notes.md.
//
//   RT_B200_BVH_KERNEL = lanes (product, default) | simple | pools | deferred | wave | wq
#pragma once
#include "rt_kernel_simple.cuh"
#include "rt_kernel_sched.cuh"
#include "rt_kernel_deferred.cuh"
#include "rt_wavefront.cuh"
#include "rt_kernel_wq.cuh"
#include "rt_trace_bench.cuh"

namespace rtb {

static int x_env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}
static void read_experiment_tunables(Tunables* v) {
    const char* e = std::getenv("RT_B200_BVH_KERNEL");
    v->bvh_variant = 3;
    if (e && std::strcmp(e, "wave") == 0) v->bvh_variant = 4;
    if (e && std::strcmp(e, "wq") == 0) v->bvh_variant = 5;
    if (e && std::strcmp(e, "simple") == 0) v->bvh_variant = 0;
    if (e && std::strcmp(e, "pools") == 0) v->bvh_variant = 1;
    if (e && std::strcmp(e, "deferred") == 0) v->bvh_variant = 2;
    v->sched_minb = x_env_int("RT_B200_SCHED_MINB", 2) == 3 ? 3 : 2;
    v->w[0] = x_env_int("RT_B200_W_NODE", 1);
    v->w[1] = x_env_int("RT_B200_W_LEAF", 1);
    v->w[2] = x_env_int("RT_B200_W_HIT", 1);
    v->w[3] = x_env_int("RT_B200_W_PRIM", 1);
    v->node_num = x_env_int("RT_B200_NODE_NUM", 1);
    v->node_den = x_env_int("RT_B200_NODE_DEN", 2);
    v->wave_refill = x_env_int("RT_B200_WAVE_REFILL", 8);
    v->wq_warps = x_env_int("RT_B200_WQ_WARPS", 24);
    if (v->wq_warps != 16 && v->wq_warps != 24 && v->wq_warps != 32) v->wq_warps = 24;
    v->wq_chains = std::max(32, std::min(WQ_MAX_CHAINS, (x_env_int("RT_B200_WQ_CHAINS", 128) / 32) * 32));
    v->wq_min_active = x_env_int("RT_B200_WQ_MIN_ACTIVE", 20);
    v->wq_min_node = x_env_int("RT_B200_WQ_MIN_NODE", 24);
    v->wq_burst = x_env_int("RT_B200_WQ_BURST", 2);
    v->wq_t_leaf = x_env_int("RT_B200_WQ_T_LEAF", 4);
    v->wq_t_pend = x_env_int("RT_B200_WQ_T_PEND", 6);
    v->wq_t_fin = x_env_int("RT_B200_WQ_T_FIN", 6);
    v->wq_sync = x_env_int("RT_B200_WQ_SYNC", 0);
    v->wq_budget = x_env_int("RT_B200_WQ_BUDGET", 0);
    v->tb_alt = x_env_int("RT_B200_TB_ALT", 0);
    v->tb_burst = x_env_int("RT_B200_WQ_BURST", 4);
}
bool legacy_node_arrays_needed() { return tunables().bvh_variant != 3 && tunables().bvh_variant != 5; }

static ExperimentBuffers* g_xbuf = nullptr;
void set_experiment_buffers(ExperimentBuffers* b) { g_xbuf = b; }

typedef void (*XKernelFn)(const DevScene, const DevCamera, const DevParams);
#define RT_PICK_SCHED(KERNEL)                                                                                       \
    do {                                                                                                            \
        if (minb == 3) {                                                                                            \
            if (smem) return count ? (XKernelFn)KERNEL<true, true, 3> : (XKernelFn)KERNEL<true, false, 3>;          \
            return count ? (XKernelFn)KERNEL<false, true, 3> : (XKernelFn)KERNEL<false, false, 3>;                  \
        }                                                                                                           \
        if (smem) return count ? (XKernelFn)KERNEL<true, true, 2> : (XKernelFn)KERNEL<true, false, 2>;              \
        return count ? (XKernelFn)KERNEL<false, true, 2> : (XKernelFn)KERNEL<false, false, 2>;                      \
    } while (0)
static XKernelFn pick_experiment(int variant, int isect, bool smem, bool count) {
    const int minb = tunables().sched_minb;
    if (isect == RT_INTERSECT_BVH && variant == 2) RT_PICK_SCHED(render_kernel_deferred);
    if (isect == RT_INTERSECT_BVH && variant == 1) RT_PICK_SCHED(render_kernel_sched);
    if (isect == RT_INTERSECT_BRUTE) {
        if (smem) return count ? (XKernelFn)render_kernel<RT_INTERSECT_BRUTE, true, true> : (XKernelFn)render_kernel<RT_INTERSECT_BRUTE, true, false>;
        return count ? (XKernelFn)render_kernel<RT_INTERSECT_BRUTE, false, true> : (XKernelFn)render_kernel<RT_INTERSECT_BRUTE, false, false>;
    }
    if (smem) return count ? (XKernelFn)render_kernel<RT_INTERSECT_BVH, true, true> : (XKernelFn)render_kernel<RT_INTERSECT_BVH, true, false>;
    return count ? (XKernelFn)render_kernel<RT_INTERSECT_BVH, false, true> : (XKernelFn)render_kernel<RT_INTERSECT_BVH, false, false>;
}

cudaError_t launch_wavefront(const DevScene& sc, const DevCamera& cam, const DevParams& pr, bool count, int sm_count,
                             int smem_optin, cudaStream_t stream, WaveBuffers* wb, LaunchInfo* info);
cudaError_t launch_wq(const DevScene& sc, const DevCamera& cam, const DevParams& pr, bool count, int sm_count,
                      int smem_optin, cudaStream_t stream, WqBuffers* wb, LaunchInfo* info);

// Returns true when an A/B kernel took the launch (*err = its status); false → the product kernel runs.
static bool launch_experiment(const DevScene& sc, const DevCamera& cam, const DevParams& pr, int isect, bool count,
                              int sm_count, int smem_optin, cudaStream_t stream, LaunchInfo* info, cudaError_t* err) {
    const Tunables& tn = tunables();
    const int v = tn.bvh_variant;
    if (v == 3) return false;
    if (info) info->counts_done = false;  // these kernels do not keep the frame's completion counters
    DevParams prm = pr;
    for (int i = 0; i < 4; i++) prm.sched_w[i] = tn.w[i];
    prm.sched_node_num = tn.node_num;
    prm.sched_node_den = tn.node_den;
    prm.tile_order_reverse = tn.tile_order_reverse;
    if (v == 5 && isect == RT_INTERSECT_BVH && pr.spp <= 65535u && pr.depth <= 255u) {
        *err = g_xbuf ? launch_wq(sc, cam, prm, count, sm_count, smem_optin, stream, &g_xbuf->wq, info) : cudaErrorInvalidValue;
        return true;
    }
    if (v == 5) {
        if (info) info->counts_done = true;
        return false;
    }
    if (v == 4 && isect == RT_INTERSECT_BVH) {
        *err = g_xbuf ? launch_wavefront(sc, cam, prm, count, sm_count, smem_optin, stream, &g_xbuf->wave, info) : cudaErrorInvalidValue;
        return true;
    }
    // simple / pools / deferred: 256-thread CTAs, reference-topology node arrays
    const size_t static_smem = WARPS * TILE_W * TILE_H * 3 + 64;
    size_t need = (size_t)sc.ns * 16 + (size_t)sc.nt * 64;
    if (isect == RT_INTERSECT_BVH) need += (size_t)sc.ni * 56;
    else need += (size_t)((sc.ns + 7u) & ~7u) * 16;
    bool smem = (need + static_smem + 1024) * 3 <= (size_t)smem_optin;
    if (tn.smem_override == 0) smem = false;
    if (tn.smem_override == 1) smem = need + static_smem + 1024 <= (size_t)smem_optin;
    XKernelFn fn = pick_experiment(v, isect, smem, count);
    const size_t dyn = smem ? need : 0;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, THREADS, dyn);
    if (e != cudaSuccess) {
        *err = e;
        return true;
    }
    if (per_sm < 1) per_sm = 1;
    const uint64_t total_tiles = (uint64_t)pr.tiles_x * pr.tiles_y;
    const uint64_t my_tiles = (total_tiles + pr.tile_ranks - 1) / pr.tile_ranks;
    uint64_t grid = std::min<uint64_t>((uint64_t)sm_count * per_sm, (my_tiles + WARPS - 1) / WARPS);
    if (grid < 1) grid = 1;
    fn<<<(unsigned)grid, THREADS, dyn, stream>>>(sc, cam, prm);
    if (info) {
        info->grid = (unsigned)grid;
        info->threads = THREADS;
        info->dyn_smem = dyn;
        info->ctas_per_sm = per_sm;
        info->scene_in_smem = smem;
    }
    *err = cudaGetLastError();
    return true;
}

// ---------------------------------------------------------------------------------------------
// Wavefront driver
// ---------------------------------------------------------------------------------------------
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
    prm.sched_w[0] = tunables().wave_refill;

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
    const Tunables& tn = tunables();
    const int nw = tn.wq_warps, chains = tn.wq_chains, min_active = tn.wq_min_active, min_node = tn.wq_min_node;
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
    prm.tile_order_reverse = tn.tile_order_reverse;
    WqArgs wa;
    wa.base = wb->state;
    wa.n = n;
    wa.chains = (uint32_t)chains;
    wa.min_active = (uint32_t)min_active;
    wa.min_node = (uint32_t)min_node;
    wa.node_burst = tn.wq_burst; wa.t_leaf = tn.wq_t_leaf; wa.t_pend = tn.wq_t_pend; wa.t_fin = tn.wq_t_fin;
    wa.cta_phases = (uint32_t)tn.wq_sync;
    wa.trace_budget = (uint32_t)tn.wq_budget;
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
    const Tunables& tn = tunables();
    a.node_burst = (uint32_t)tn.tb_burst;
    a.t_leaf = (uint32_t)tn.wq_t_leaf;
    a.t_pend = (uint32_t)tn.wq_t_pend;
    a.t_fin = (uint32_t)tn.wq_t_fin;
    a.alt = (uint32_t)tn.tb_alt;
    a.sstack_off = (uint32_t)(need / 4);
    size_t need_ww = need + (a.alt == 1 ? (size_t)TB_SSTACK * 768 * 4 : 0);
    if (a.alt == 2) need_ww = (size_t)sc.ns * 16 + (size_t)sc.nt * 64 + (size_t)sc.w4n * 112 + 16;
    if (a.alt && (need_ww + 1024 > (size_t)smem_optin || sc.lni == 0)) return cudaErrorInvalidValue;
    if (a.alt == 2 && sc.w4n == 0) return cudaErrorInvalidValue;
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

void free_experiment_buffers(ExperimentBuffers* b) {
    free_wave_buffers(&b->wave);
    free_wq_buffers(&b->wq);
}

}  // namespace rtb
