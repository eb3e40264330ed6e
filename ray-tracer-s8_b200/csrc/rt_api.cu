// rt_api.cu — the C ABI of include/rt_b200.h: context, scene upload (+ host BVH build), render calls.
//
// Replaces the body of worker() in ray-tracer-slave/src/main.rs:32-106 (see the header for the mapping).
// There is no CPU fallback anywhere in this file: every render goes through launch_render().
#include <chrono>
#include <functional>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <limits>
#include <new>

#include "rt_device.cuh"
#include "rt_host.h"

using namespace rtb;

struct rt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0, clock_khz = 0, smem_optin = 0;
    char name[64] = {0};
    uint8_t* d_out = nullptr;  // band / frame staging in HBM
    size_t d_out_bytes = 0;
    unsigned long long* d_ctr = nullptr;  // NUM_COUNTERS counters + tile ticket (last slot)
    unsigned long long* h_ctr = nullptr;  // pinned mirror
    float* d_scratch = nullptr;
    WaveBuffers wave;
    WqBuffers wq;
    DeviceBuild dbuild;
    // measurement aid: when set, the instrumented render kernel records its queries here (rt_debug_trace_bench)
    float4* dump_rays = nullptr;
    unsigned long long* dump_n = nullptr;
    unsigned long long dump_cap = 0;
    // scene upload: one pinned staging buffer the blob is assembled in, and a few retired device blobs kept for the
    // next rt_scene_create (a slave makes one scene per job: cudaMalloc/cudaFree per job were a third of a small job)
    uint8_t* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    struct Retired { uint8_t* p; size_t cap; };
    std::vector<Retired> retired;
    std::string err;
};

struct rt_scene {
    uint8_t* d_blob = nullptr;
    size_t blob_bytes = 0, blob_cap = 0;
    DevScene dev{};
    uint32_t n = 0, n_nodes = 0, depth = 0;
    uint32_t device_tree_depth = 0;  // > 0: the traversal tree was built on the device (LBVH), this deep
    std::vector<uint32_t> rank_by_world;  // world position → DFS leaf rank
};

static thread_local std::string g_init_err = "";

static int set_err(rt_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_init_err = buf;
    return code;
}

#define CK(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return set_err(ctx, RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                           __FILE__, __LINE__);                                                         \
    } while (0)

extern "C" {

int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

void rt_struct_sizes(size_t out[4]) {
    out[0] = sizeof(rt_sphere);
    out[1] = sizeof(rt_triangle);
    out[2] = sizeof(rt_params);
    out[3] = sizeof(rt_stats);
}

const char* rt_last_error(const rt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_err.c_str(); }

int rt_init(int device, rt_ctx** out) {
    if (!out) return set_err(nullptr, RT_ERR_INVALID_ARG, "rt_init: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_err(nullptr, RT_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count)
        return set_err(nullptr, RT_ERR_INVALID_ARG, "device %d out of range [0,%d)", device, count);
    rt_ctx* ctx = new (std::nothrow) rt_ctx();
    if (!ctx) return set_err(nullptr, RT_ERR_INVALID_ARG, "out of host memory");
    ctx->device = device;
    cudaDeviceProp prop;
#define CKI(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            set_err(nullptr, RT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));            \
            rt_shutdown(ctx);                                                                          \
            return RT_ERR_CUDA;                                                                        \
        }                                                                                              \
    } while (0)
    CKI(cudaSetDevice(device));
    CKI(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_err(nullptr, RT_ERR_NO_DEVICE, "device %d is sm_%d%d; this build carries sm_100a code only", device,
                prop.major, prop.minor);
        rt_shutdown(ctx);
        return RT_ERR_NO_DEVICE;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    CKI(cudaDeviceGetAttribute(&ctx->clock_khz, cudaDevAttrClockRate, device));
    memcpy(ctx->name, prop.name, 63); ctx->name[63] = 0;
    CKI(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CKI(cudaEventCreate(&ctx->ev0));
    CKI(cudaEventCreate(&ctx->ev1));
    CKI(cudaMalloc(&ctx->d_ctr, (NUM_COUNTERS + 1) * sizeof(unsigned long long)));
    CKI(cudaMallocHost(&ctx->h_ctr, (NUM_COUNTERS + 1) * sizeof(unsigned long long)));
#undef CKI
    *out = ctx;
    return RT_OK;
}

void rt_shutdown(rt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_ctr) cudaFree(ctx->d_ctr);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    free_wave_buffers(&ctx->wave);
    free_wq_buffers(&ctx->wq);
    free_device_build(&ctx->dbuild);
    for (auto& r : ctx->retired) cudaFree(r.p);
    ctx->retired.clear();
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int rt_device_info(rt_ctx* ctx, int* sm_count, int* clock_khz, int* smem_optin, char name_out[64]) {
    if (!ctx) return RT_ERR_INVALID_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (clock_khz) *clock_khz = ctx->clock_khz;
    if (smem_optin) *smem_optin = ctx->smem_optin;
    if (name_out) memcpy(name_out, ctx->name, 64);
    return RT_OK;
}

void* rt_stream(rt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int rt_sync(rt_ctx* ctx) {
    if (!ctx) return RT_ERR_INVALID_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

// -------------------------------------------------------------------------------------------------
// Scene
// -------------------------------------------------------------------------------------------------
namespace {

struct PrimRef {
    uint8_t kind;  // 0 sphere, 1 triangle
    uint32_t idx;  // index in the caller's array
};

// min_by / max_by with partial_cmp().unwrap_or(Equal) (mesh.rs:46-95): min_by returns the first
// argument unless first > second; max_by returns the second unless first > second.
inline float ref_min(float x, float y) { return (x > y) ? y : x; }
inline float ref_max(float x, float y) { return (x > y) ? x : y; }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline bool finite3(const float* v) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

}  // namespace


// World order + bounds + BVH, shared by rt_scene_create and rt_bvh_build_host.  Host only.
static int build_world_bvh(const rt_sphere* spheres, uint32_t n_spheres, const rt_triangle* triangles,
                           uint32_t n_triangles, const uint32_t* world_index, std::vector<PrimRef>* world_out,
                           HostBVH* bvh, std::string* err, std::vector<Box>* boxes_out = nullptr,
                           const std::function<void(const std::vector<Box>&)>* side_job = nullptr) {
    char buf[256];
    const uint32_t n = n_spheres + n_triangles;
    std::vector<PrimRef>& world = *world_out;
    world.assign(n, PrimRef{0, 0});
    {
        std::vector<uint8_t> seen(n, 0);
        for (uint32_t i = 0; i < n; i++) {
            uint32_t pos = world_index ? world_index[i] : i;
            if (pos >= n || seen[pos]) {
                snprintf(buf, sizeof buf, "world_index is not a permutation of 0..%u", n - 1);
                *err = buf;
                return RT_ERR_INVALID_ARG;
            }
            seen[pos] = 1;
            world[pos] = i < n_spheres ? PrimRef{0, i} : PrimRef{1, i - n_spheres};
        }
    }
    // bounds per world position (Sphere::aabb sphere.rs:65-72, Triangle::aabb mesh.rs:46-95)
    std::vector<Box> boxes(n);
    for (uint32_t w = 0; w < n; w++) {
        Box& b = boxes[w];
        if (world[w].kind == 0) {
            const rt_sphere& s = spheres[world[w].idx];
            if (!finite3(s.center) || !std::isfinite(s.radius)) {
                snprintf(buf, sizeof buf, "sphere %u has non-finite geometry", world[w].idx);
                *err = buf;
                return RT_ERR_BVH;
            }
            for (int a = 0; a < 3; a++) {
                b.min[a] = s.center[a] - s.radius;
                b.max[a] = s.center[a] + s.radius;
            }
        } else {
            const rt_triangle& t = triangles[world[w].idx];
            if (!finite3(t.a) || !finite3(t.b) || !finite3(t.c)) {
                snprintf(buf, sizeof buf, "triangle %u has non-finite geometry", world[w].idx);
                *err = buf;
                return RT_ERR_BVH;
            }
            for (int a = 0; a < 3; a++) {
                b.min[a] = ref_min(ref_min(t.a[a], t.c[a]), t.b[a]);
                b.max[a] = ref_max(ref_max(t.a[a], t.c[a]), t.b[a]);
            }
        }
    }
    // an independent job on the boxes (the traversal-tree build) runs beside the reference build when it is worth a thread
    std::thread side;
    bool side_done = false;
    if (side_job && n >= 512) {
        try {
            side = std::thread([&] { (*side_job)(boxes); });
            side_done = true;
        } catch (...) {
        }
    }
    std::string berr;
    const bool built = build_bvh(boxes, bvh, &berr);
    if (side.joinable()) side.join();
    if (side_job && !side_done) (*side_job)(boxes);
    if (!built) {
        *err = "BVH build failed: " + berr;
        return RT_ERR_BVH;
    }
    if (bvh->depth > (uint32_t)MAX_STACK) {
        snprintf(buf, sizeof buf, "BVH depth %u exceeds the traversal stack (%d)", bvh->depth, MAX_STACK);
        *err = buf;
        return RT_ERR_UNSUPPORTED;
    }
    if (boxes_out) boxes_out->swap(boxes);
    return RT_OK;
}

// centre / half-extent form of a box (FILTER domain): t = (c - o)*inv -/+ h*|inv|.  h is padded for the rounding of
// the c- and h-terms (<= 2^-24 * (2|c| + h) per axis); the o-term is covered per ray by the kernels.
static void centre_half_of(const Box& b, float c[3], float h[3]) {
    double m = 0.0;
    for (int a = 0; a < 3; a++) m = std::fmax(m, std::fmax(std::fabs((double)b.min[a]), std::fabs((double)b.max[a])));
    for (int a = 0; a < 3; a++) {
        const double cc = 0.5 * ((double)b.min[a] + (double)b.max[a]);
        const double hh = 0.5 * ((double)b.max[a] - (double)b.min[a]);
        c[a] = (float)cc;
        h[a] = (float)(hh * (1.0 + 4e-6) + 2e-6 * m + 1e-30);
    }
}

int rt_bvh_build_host(const rt_sphere* spheres, uint32_t n_spheres, const rt_triangle* triangles, uint32_t n_triangles,
                      const uint32_t* world_index, uint32_t* rank_out, uint32_t* n_nodes, uint32_t* depth) {
    const uint64_t n64 = (uint64_t)n_spheres + n_triangles;
    if (n64 == 0) return RT_ERR_EMPTY_SCENE;
    if (n64 > 0x3ffffffu) return RT_ERR_UNSUPPORTED;  // leaf codes hold a 26-bit primitive id
    if ((n_spheres && !spheres) || (n_triangles && !triangles)) return RT_ERR_INVALID_ARG;
    std::vector<PrimRef> world;
    HostBVH bvh;
    std::string err;
    int rc = build_world_bvh(spheres, n_spheres, triangles, n_triangles, world_index, &world, &bvh, &err);
    if (rc) {
        g_init_err = err;
        return rc;
    }
    if (rank_out)
        for (uint32_t r = 0; r < (uint32_t)n64; r++) rank_out[bvh.leaf_order[r]] = r;
    if (n_nodes) *n_nodes = bvh.node_count;
    if (depth) *depth = bvh.depth;
    return RT_OK;
}

int rt_scene_create(rt_ctx* ctx, const rt_sphere* spheres, uint32_t n_spheres, const rt_triangle* triangles,
                    uint32_t n_triangles, const uint32_t* world_index, rt_scene** out) {
    if (!ctx) return RT_ERR_INVALID_ARG;
    if (!out) return set_err(ctx, RT_ERR_INVALID_ARG, "rt_scene_create: out is NULL");
    *out = nullptr;
    const uint64_t n64 = (uint64_t)n_spheres + n_triangles;
    if (n64 == 0)
        return set_err(ctx, RT_ERR_EMPTY_SCENE,
                       "empty world: the reference's BVHNode::build never terminates on zero shapes");
    if (n64 > 0x3ffffffu)  // leaf codes are ~((first_pid << 5) | (count - 1)) in 32 bits
        return set_err(ctx, RT_ERR_UNSUPPORTED, "too many primitives (%llu; this build holds primitive ids in 26 bits)", (unsigned long long)n64);
    if ((n_spheres && !spheres) || (n_triangles && !triangles))
        return set_err(ctx, RT_ERR_INVALID_ARG, "rt_scene_create: NULL primitive array");
    const uint32_t n = (uint32_t)n64;
    // RT_B200_TIMING=1: stage times of this call on stderr (development aid)
    static const bool timing = std::getenv("RT_B200_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[rt_scene_create n=%u] %-22s %8.3f ms\n", n, what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    std::vector<uint32_t> big_world;  // split layout: world positions of the primitives kept out of the tree
    bool ltree = true;
    // RT_B200_BUILD = auto (default) | host | device.  device: the traversal tree is an LBVH built on the GPU after the
    // upload (rt_bvh_device.cu) instead of the host's binned-SAH tree; auto picks it from kDeviceBuildMin primitives,
    // where the host build costs more than the LBVH's extra slab tests (profiles/r1_notes.md).
    static const int build_mode = [] {
        const char* e = std::getenv("RT_B200_BUILD");
        return (e && std::strcmp(e, "host") == 0) ? 0 : ((e && std::strcmp(e, "device") == 0) ? 2 : 1);
    }();
    constexpr uint32_t kDeviceBuildMin = 8192;
    bool device_tree = false;
    std::vector<float> dev_boxes;     // boxes and pids of the primitives the device-built tree covers
    std::vector<uint32_t> dev_pid;
    HostBVH sah;
    bool use_sah = false;
    std::vector<uint32_t> dev_world;  // device-built tree: world positions of the primitives it covers
    // Chooses and (on the host path) builds the traversal tree from the boxes alone, so it can run beside the reference build.
    const std::function<void(const std::vector<Box>&)> select_tree = [&](const std::vector<Box>& boxes) {
        // RT_B200_TREE = split (default) | sah | ref.  ref: the reference-topology tree itself.  sah: a 3-axis binned-SAH
        // tree over all primitives.  split: the primitives whose box area is a large share of the whole scene's go to a
        // short list tested ahead of the traversal, and the 3-axis SAH tree covers the rest (measured: profiles/).
        const char* te = std::getenv("RT_B200_TREE");
        int mode = 2;
        if (te && std::strcmp(te, "ref") == 0) mode = 0;
        if (te && std::strcmp(te, "sah") == 0) mode = 1;
        if (mode == 2 && n > 1) {
            auto area_of = [](const Box& b) {
                const double sx = (double)b.max[0] - b.min[0], sy = (double)b.max[1] - b.min[1], sz = (double)b.max[2] - b.min[2];
                return 2.0 * (sx * sy + sx * sz + sy * sz);
            };
            Box all = boxes[0];
            for (uint32_t w = 1; w < n; w++)
                for (int a = 0; a < 3; a++) {
                    all.min[a] = fminf(all.min[a], boxes[w].min[a]);
                    all.max[a] = fmaxf(all.max[a], boxes[w].max[a]);
                }
            const double thresh = area_of(all) * (1.0 / 16.0);
            std::vector<std::pair<double, uint32_t>> cand;
            for (uint32_t w = 0; w < n; w++) {
                const double a = area_of(boxes[w]);
                if (a > thresh && a > 0.0) cand.push_back({-a, w});
            }
            std::sort(cand.begin(), cand.end());
            if (cand.size() > (size_t)MAX_BIG) cand.resize(MAX_BIG);
            std::vector<uint8_t> is_big(n, 0);
            for (auto& c : cand) {
                big_world.push_back(c.second);
                is_big[c.second] = 1;
            }
            std::sort(big_world.begin(), big_world.end());
            std::vector<Box> rest;
            std::vector<uint32_t> rest_world;
            for (uint32_t w = 0; w < n; w++)
                if (!is_big[w]) {
                    rest.push_back(boxes[w]);
                    rest_world.push_back(w);
                }
            if (rest.empty()) {
                ltree = false;
            } else if (rest.size() >= 2 && (build_mode == 2 || (build_mode == 1 && rest.size() >= kDeviceBuildMin))) {
                device_tree = true;
                dev_boxes.resize(6 * rest.size());
                for (size_t i = 0; i < rest.size(); i++)
                    for (int a = 0; a < 3; a++) {
                        dev_boxes[6 * i + a] = rest[i].min[a];
                        dev_boxes[6 * i + 3 + a] = rest[i].max[a];
                    }
                dev_world = rest_world;  // pids are filled in once the reference build has assigned them
            } else {
                use_sah = build_bvh_sah(rest, &sah) && sah.depth <= (uint32_t)MAX_STACK;
                if (use_sah) {  // leaf codes: subset index → world position
                    auto remap = [&](int32_t c) { return c >= 0 ? c : ~(int32_t)rest_world[(uint32_t)~c]; };
                    for (auto& nd : sah.inner) {
                        nd.left = remap(nd.left);
                        nd.right = remap(nd.right);
                    }
                    sah.root = remap(sah.root);
                } else {
                    big_world.clear();  // fall back to the reference tree over everything
                }
            }
        } else if (mode == 1) {
            use_sah = build_bvh_sah(boxes, &sah) && sah.depth <= (uint32_t)MAX_STACK;
        }
    };
    std::vector<PrimRef> world;
    std::vector<Box> boxes;  // the reference's (unpadded) shape AABBs by world position
    HostBVH bvh;
    {
        std::string herr;
        int hrc = build_world_bvh(spheres, n_spheres, triangles, n_triangles, world_index, &world, &bvh, &herr, &boxes,
                                  &select_tree);
        if (hrc) return set_err(ctx, hrc, "%s", herr.c_str());
    }
    // pids: spheres then triangles, each in DFS leaf order
    std::vector<uint32_t> pid_of_world(n), rank_of_world(n);
    uint32_t next_s = 0, next_t = n_spheres;
    for (uint32_t r = 0; r < n; r++) {
        uint32_t w = bvh.leaf_order[r];
        rank_of_world[w] = r;
        pid_of_world[w] = world[w].kind == 0 ? next_s++ : next_t++;
    }
    const uint32_t ni = (uint32_t)bvh.inner.size();
    if (device_tree) {
        dev_pid.resize(dev_world.size());
        for (size_t i = 0; i < dev_world.size(); i++) dev_pid[i] = pid_of_world[dev_world[i]];
    }
    lap("reference-tree build");

    // blob layout (each section 256-byte aligned)
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    // The reference-topology node arrays (node_*, cnode_*) feed only the kernels kept for A/B runs
    // (RT_B200_BVH_KERNEL=simple|pools|deferred|wave); the product kernels read the lnode_* tree.
    const bool legacy = legacy_node_arrays_needed();
    const size_t nl = legacy ? (size_t)ni : 0;
    const size_t o_sph = take((size_t)n_spheres * 16), o_tri = take((size_t)n_triangles * 64);
    const size_t o_na = take(nl * 16), o_nb = take(nl * 16), o_nc = take(nl * 16);
    const size_t o_nd = take(nl * 8), o_mat = take((size_t)n * 16), o_em = take((size_t)n * 4);
    const size_t o_rank = take((size_t)n * 4), o_box = take((size_t)n * 32);
    const size_t o_ca = take(nl * 16), o_cb = take(nl * 16), o_cc = take(nl * 16);
    const size_t o_la = take((size_t)ni * 48);  // lnode_abc: three float4 per node, one 48-byte record
    const size_t o_ld = take((size_t)ni * 8);
    // brute-force kernel: spheres in pairs for the packed f32x2 filter, padded to a multiple of 8 spheres
    const uint32_t ns8 = (n_spheres + 7u) & ~7u;
    const size_t o_sph2 = take((size_t)ns8 * 16);
    struct Stage {  // the blob, assembled in the context's pinned staging buffer
        uint8_t* p;
        size_t n;
        uint8_t* data() const { return p; }
        size_t size() const { return n; }
    } blob{nullptr, off ? off : 256};
    if (ctx->h_stage_bytes < blob.n) {
        CK(ctx, cudaSetDevice(ctx->device));
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        ctx->h_stage_bytes = 0;
        size_t cap = (size_t)1 << 20;
        while (cap < blob.n) cap <<= 1;
        CK(ctx, cudaMallocHost(&ctx->h_stage, cap));
        ctx->h_stage_bytes = cap;
    }
    blob.p = ctx->h_stage;
    lap("staging buffer");
    memset(blob.p, 0, blob.n);
    float* h_sph = (float*)(blob.data() + o_sph);
    float* h_tri = (float*)(blob.data() + o_tri);
    float* h_na = (float*)(blob.data() + o_na);
    float* h_nb = (float*)(blob.data() + o_nb);
    float* h_nc = (float*)(blob.data() + o_nc);
    int32_t* h_nd = (int32_t*)(blob.data() + o_nd);
    float* h_mat = (float*)(blob.data() + o_mat);
    float* h_em = (float*)(blob.data() + o_em);
    uint32_t* h_rank = (uint32_t*)(blob.data() + o_rank);

    float* h_box = (float*)(blob.data() + o_box);
    for (uint32_t w = 0; w < n; w++) {
        const uint32_t pid = pid_of_world[w];
        h_rank[pid] = rank_of_world[w];
        for (int a = 0; a < 3; a++) {
            h_box[8 * (size_t)pid + a] = boxes[w].min[a];
            h_box[8 * (size_t)pid + 4 + a] = boxes[w].max[a];
        }
        if (world[w].kind == 0) {
            const rt_sphere& s = spheres[world[w].idx];
            float* g = h_sph + 4 * (size_t)pid;
            g[0] = s.center[0]; g[1] = s.center[1]; g[2] = s.center[2];
            g[3] = s.radius * s.radius;  // radius.powi(2)
            float* m = h_mat + 4 * (size_t)pid;
            m[0] = s.albedo[0]; m[1] = s.albedo[1]; m[2] = s.albedo[2]; m[3] = s.roughness;
            h_em[pid] = s.emission;
        } else {
            const rt_triangle& t = triangles[world[w].idx];
            float* g = h_tri + 16 * (size_t)(pid - n_spheres);
            float ab[3], ac[3], amb[3], amc[3];
            for (int a = 0; a < 3; a++) {
                ab[a] = t.b[a] - t.a[a];   // a_to_b (mesh.rs:111)
                ac[a] = t.c[a] - t.a[a];   // a_to_c
                amb[a] = t.a[a] - t.b[a];  // normal_at: (a-b).cross(a-c) (mesh.rs:164)
                amc[a] = t.a[a] - t.c[a];
            }
            // glam cross + normalize_or_zero, single f32 ops (host built with -ffp-contract=off)
            float cr[3] = {amb[1] * amc[2] - amc[1] * amb[2], amb[2] * amc[0] - amc[2] * amb[0],
                           amb[0] * amc[1] - amc[0] * amb[1]};
            float dd = (cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2];
            float rcp = 1.0f / sqrtf(dd);
            float nrm[3] = {0.0f, 0.0f, 0.0f};
            if (std::isfinite(rcp) && rcp > 0.0f) {
                nrm[0] = cr[0] * rcp; nrm[1] = cr[1] * rcp; nrm[2] = cr[2] * rcp;
            }
            for (int a = 0; a < 3; a++) {
                g[0 + a] = t.a[a];
                g[4 + a] = ab[a];
                g[8 + a] = ac[a];
                g[12 + a] = nrm[a];
            }
            float* m = h_mat + 4 * (size_t)pid;
            m[0] = t.albedo[0]; m[1] = t.albedo[1]; m[2] = t.albedo[2]; m[3] = t.roughness;
            h_em[pid] = t.emission;
        }
    }
    auto code_of = [&](int32_t c) -> int32_t { return c >= 0 ? c : ~(int32_t)pid_of_world[(uint32_t)~c]; };
    auto padded = [](const Box& b, float lo[3], float hi[3]) {
        // conservative culling: pad by a few ulp of the largest coordinate (FILTER-domain slab tests)
        float m = 0.0f;
        for (int a = 0; a < 3; a++) m = fmaxf(m, fmaxf(fabsf(b.min[a]), fabsf(b.max[a])));
        const float pad = m * 4e-6f + 1e-30f;
        for (int a = 0; a < 3; a++) {
            lo[a] = b.min[a] - pad;
            hi[a] = b.max[a] + pad;
        }
    };
    for (uint32_t i = 0; legacy && i < ni; i++) {
        const HostNode& hn = bvh.inner[i];
        float ll[3], lh[3], rl[3], rh[3];
        padded(hn.box_l, ll, lh);
        padded(hn.box_r, rl, rh);
        float* a = h_na + 4 * (size_t)i;
        float* b = h_nb + 4 * (size_t)i;
        float* c = h_nc + 4 * (size_t)i;
        a[0] = ll[0]; a[1] = ll[1]; a[2] = ll[2]; a[3] = lh[0];
        b[0] = lh[1]; b[1] = lh[2]; b[2] = rl[0]; b[3] = rl[1];
        c[0] = rl[2]; c[1] = rh[0]; c[2] = rh[1]; c[3] = rh[2];
        h_nd[2 * (size_t)i] = code_of(hn.left);
        h_nd[2 * (size_t)i + 1] = code_of(hn.right);
        // centre / half-extent form (FILTER domain): t = (c - o)*inv -/+ h*|inv|.  h is padded for the
        // rounding of the c- and h-terms (<= 2^-24 * (2|c| + h) per axis); the o-term is covered per ray.
        auto centre_half = [](const Box& b, float c[3], float h[3]) {
            double m = 0.0;
            for (int a = 0; a < 3; a++) m = std::fmax(m, std::fmax(std::fabs((double)b.min[a]), std::fabs((double)b.max[a])));
            for (int a = 0; a < 3; a++) {
                double cc = 0.5 * ((double)b.min[a] + (double)b.max[a]);
                double hh = 0.5 * ((double)b.max[a] - (double)b.min[a]);
                c[a] = (float)cc;
                h[a] = (float)(hh * (1.0 + 4e-6) + 2e-6 * m + 1e-30);
            }
        };
        float lc[3], lhh[3], rc[3], rhh[3];
        centre_half(hn.box_l, lc, lhh);
        centre_half(hn.box_r, rc, rhh);
        float* ca = (float*)(blob.data() + o_ca) + 4 * (size_t)i;
        float* cb = (float*)(blob.data() + o_cb) + 4 * (size_t)i;
        float* cc = (float*)(blob.data() + o_cc) + 4 * (size_t)i;
        ca[0] = lc[0]; ca[1] = lc[1]; ca[2] = lc[2]; ca[3] = lhh[0];
        cb[0] = lhh[1]; cb[1] = lhh[2]; cb[2] = rc[0]; cb[3] = rc[1];
        cc[0] = rc[2]; cc[1] = rhh[0]; cc[2] = rhh[1]; cc[3] = rhh[2];
    }

    lap("primitive + node arrays");
    // ---- the lanes kernel's tree.  RT_B200_TREE=sah: cull with a 3-axis binned-SAH tree instead of the reference's
    //      single-axis 6-bucket tree (ties still follow the reference tree's DFS ranks).  RT_B200_LEAF=L collapses
    //      subtrees of <= L same-kind primitives into one leaf (only with the reference tree, whose DFS order = pid order).
    uint32_t lni = 0;
    int32_t lroot = 0;
    {
        const HostBVH& T = use_sah ? sah : bvh;
        int L = 1;  // measured on C3 (reference tree): 1 → 51.8 ms, 4 → 54.6, 16 → 62.7 (profiles/)
        if (const char* e = std::getenv("RT_B200_LEAF")) L = std::atoi(e);
        if (L < 1 || use_sah) L = 1;
        if (L > 32) L = 32;
        struct Sub { uint32_t n_s, n_t, first_s, first_t; };
        auto leaf_sub = [&](int32_t code) {
            const uint32_t w = (uint32_t)~code, pid = pid_of_world[w];
            return world[w].kind == 0 ? Sub{1, 0, pid, 0} : Sub{0, 1, 0, pid};
        };
        const uint32_t tni = (uint32_t)T.inner.size();
        std::vector<Sub> sub(tni);
        auto sub_of = [&](int32_t code) { return code >= 0 ? sub[(size_t)code] : leaf_sub(code); };
        for (uint32_t i = tni; i-- > 0;) {  // pre-order: children have larger indices than their parent
            const Sub a = sub_of(T.inner[i].left), b = sub_of(T.inner[i].right);
            sub[i] = Sub{a.n_s + b.n_s, a.n_t + b.n_t, a.n_s ? a.first_s : b.first_s, a.n_t ? a.first_t : b.first_t};
        }
        float* la = (float*)(blob.data() + o_la);
        int32_t* ld = (int32_t*)(blob.data() + o_ld);
        auto leaf_code = [](uint32_t first, uint32_t count) { return ~(int32_t)((first << 5) | (count - 1)); };
        struct Item { int32_t code; uint32_t parent; int side; };
        std::vector<Item> todo;
        auto classify = [&](int32_t code, bool* is_leaf) -> int32_t {
            const Sub sb = sub_of(code);
            const uint32_t tot = sb.n_s + sb.n_t;
            if (code < 0 || (tot <= (uint32_t)L && (sb.n_s == 0 || sb.n_t == 0))) {
                *is_leaf = true;
                return leaf_code(sb.n_s ? sb.first_s : sb.first_t, tot);
            }
            *is_leaf = false;
            return 0;
        };
        bool root_leaf = true;
        if (device_tree) {  // nodes 0 .. n-2 are written on the device after the upload; the root is node 0
            lni = (uint32_t)dev_pid.size() - 1;
            lroot = 0;
        } else if (ltree) {
            lroot = classify(T.root, &root_leaf);
        }
        if (!root_leaf) {
            todo.push_back(Item{T.root, 0, -1});
            while (!todo.empty()) {
                const Item it = todo.back();
                todo.pop_back();
                const uint32_t me = lni++;
                if (it.side < 0) lroot = (int32_t)me;
                else ld[2 * (size_t)it.parent + it.side] = (int32_t)me;
                const HostNode& hn = T.inner[(size_t)it.code];
                float lcn[3], lhh[3], rcn[3], rhh[3];
                centre_half_of(hn.box_l, lcn, lhh);
                centre_half_of(hn.box_r, rcn, rhh);
                float* pa = la + 12 * (size_t)me;
                float* pb = pa + 4;
                float* pc = pa + 8;
                pa[0] = lcn[0]; pa[1] = lcn[1]; pa[2] = lcn[2]; pa[3] = lhh[0];
                pb[0] = lhh[1]; pb[1] = lhh[2]; pb[2] = rcn[0]; pb[3] = rcn[1];
                pc[0] = rcn[2]; pc[1] = rhh[0]; pc[2] = rhh[1]; pc[3] = rhh[2];
                const int32_t kids[2] = {hn.left, hn.right};
                for (int side = 1; side >= 0; side--) {  // push right first so the left subtree is numbered first
                    bool is_leaf;
                    const int32_t c = classify(kids[side], &is_leaf);
                    if (is_leaf) ld[2 * (size_t)me + side] = c;
                    else todo.push_back(Item{kids[side], me, side});
                }
            }
        }
    }

    lap("traversal tree");
    rt_scene* sc = new (std::nothrow) rt_scene();
    if (!sc) return set_err(ctx, RT_ERR_INVALID_ARG, "out of host memory");
    sc->n = n;
    sc->n_nodes = bvh.node_count;
    sc->depth = bvh.depth;
    sc->rank_by_world = rank_of_world;
    {   // pair j = spheres 2j, 2j+1: (-c0.x, -c1.x, -c0.y, -c1.y) | (-c0.z, -c1.z, r0^2, r1^2); pads can never pass
        float* h2 = (float*)(blob.data() + o_sph2);
        for (uint32_t j = 0; j < ns8 / 2; j++) {
            for (uint32_t k = 0; k < 2; k++) {
                const uint32_t pid = 2 * j + k;
                const bool real = pid < n_spheres;
                const float* g = h_sph + 4 * (size_t)pid;
                h2[8 * (size_t)j + 0 + k] = real ? -g[0] : 0.0f;
                h2[8 * (size_t)j + 2 + k] = real ? -g[1] : 0.0f;
                h2[8 * (size_t)j + 4 + k] = real ? -g[2] : 0.0f;
                h2[8 * (size_t)j + 6 + k] = real ? g[3] : -std::numeric_limits<float>::infinity();
            }
        }
    }
    sc->blob_bytes = blob.size();
    cudaError_t e = cudaSetDevice(ctx->device);
    {   // smallest retired device blob that fits, else a new allocation
        size_t pick = ctx->retired.size();
        for (size_t i = 0; i < ctx->retired.size(); i++)
            if (ctx->retired[i].cap >= blob.size() && (pick == ctx->retired.size() || ctx->retired[i].cap < ctx->retired[pick].cap)) pick = i;
        if (pick < ctx->retired.size()) {
            sc->d_blob = ctx->retired[pick].p;
            sc->blob_cap = ctx->retired[pick].cap;
            ctx->retired.erase(ctx->retired.begin() + (long)pick);
        } else if (e == cudaSuccess) {
            // power-of-two size classes (>= 256 KB): retired blobs fit the next scene of a similar size, and the driver
            // sees few distinct allocation sizes (cudaMalloc / cudaMallocHost stall for 50-200 ms now and then)
            size_t cap = (size_t)1 << 18;
            while (cap < blob.size()) cap <<= 1;
            sc->blob_cap = cap;
            e = cudaMalloc(&sc->d_blob, sc->blob_cap);
        }
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(sc->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        if (sc->d_blob) cudaFree(sc->d_blob);
        delete sc;
        return set_err(ctx, RT_ERR_CUDA, "scene upload failed: %s", cudaGetErrorString(e));
    }
    lap("upload");
    if (device_tree) {
        uint32_t depth = 0;
        e = build_lbvh_device(&ctx->dbuild, dev_boxes.data(), dev_pid.data(), (uint32_t)dev_pid.size(),
                              (float4*)(sc->d_blob + o_la), (int2*)(sc->d_blob + o_ld), &depth, ctx->stream);
        if (e != cudaSuccess || depth > (uint32_t)MAX_STACK) {
            ctx->retired.push_back({sc->d_blob, sc->blob_cap});
            delete sc;
            if (e != cudaSuccess) return set_err(ctx, RT_ERR_CUDA, "device BVH build failed: %s", cudaGetErrorString(e));
            return set_err(ctx, RT_ERR_UNSUPPORTED, "device-built tree depth %u exceeds the traversal stack (%d)", depth, MAX_STACK);
        }
        sc->device_tree_depth = depth;
        lap("device tree build");
    }
    DevScene& d = sc->dev;
    d.sph = (const float4*)(sc->d_blob + o_sph);
    d.tri = (const float4*)(sc->d_blob + o_tri);
    d.sph2 = (const float4*)(sc->d_blob + o_sph2);
    d.node_a = (const float4*)(sc->d_blob + o_na);
    d.node_b = (const float4*)(sc->d_blob + o_nb);
    d.node_c = (const float4*)(sc->d_blob + o_nc);
    d.node_d = (const int2*)(sc->d_blob + o_nd);
    d.cnode_a = (const float4*)(sc->d_blob + o_ca);
    d.cnode_b = (const float4*)(sc->d_blob + o_cb);
    d.cnode_c = (const float4*)(sc->d_blob + o_cc);
    d.lnode_a = (const float4*)(sc->d_blob + o_la);
    d.lnode_d = (const int2*)(sc->d_blob + o_ld);
    d.lni = lni;
    d.lroot = lroot;
    d.ltree = ltree ? 1 : 0;
    d.nbig = (uint32_t)big_world.size();
    for (size_t i = 0; i < (size_t)MAX_BIG; i++) d.big_pid[i] = i < big_world.size() ? pid_of_world[big_world[i]] : 0u;
    d.mat = (const float4*)(sc->d_blob + o_mat);
    d.emis = (const float*)(sc->d_blob + o_em);
    d.rank = (const uint32_t*)(sc->d_blob + o_rank);
    d.leaf_box = (const float4*)(sc->d_blob + o_box);
    d.ns = n_spheres;
    d.nt = n_triangles;
    d.ni = ni;
    d.root = code_of(bvh.root);
    *out = sc;
    return RT_OK;
}

void rt_scene_destroy(rt_ctx* ctx, rt_scene* scene) {
    if (!scene) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    if (scene->d_blob) {
        if (ctx && ctx->retired.size() < 4) ctx->retired.push_back({scene->d_blob, scene->blob_cap});
        else cudaFree(scene->d_blob);
    }
    delete scene;
}

int rt_scene_info(const rt_scene* scene, uint32_t* n_prims, uint32_t* n_nodes, uint32_t* depth, uint32_t* rank_out) {
    if (!scene) return RT_ERR_INVALID_ARG;
    if (n_prims) *n_prims = scene->n;
    if (n_nodes) *n_nodes = scene->n_nodes;
    if (depth) *depth = scene->depth;
    if (rank_out) memcpy(rank_out, scene->rank_by_world.data(), scene->n * sizeof(uint32_t));
    return RT_OK;
}

size_t rt_scene_device_bytes(const rt_scene* scene) { return scene ? scene->blob_bytes : 0; }

// -------------------------------------------------------------------------------------------------
// Render
// -------------------------------------------------------------------------------------------------
namespace {

struct Resolved {
    rt_params p;
    DevCamera cam;
    int isect;
};

// Brute force (K1) is picked only for tiny scenes: on the C5 sweep the BVH kernel already wins at 64 spheres
// (0.152 ms vs 0.247 ms at 1080p), see profiles/r1_c5_sweep.log
constexpr uint32_t kBruteMaxPrims = 16;

int resolve(rt_ctx* ctx, const rt_scene* scene, const rt_params* in, Resolved* r) {
    if (!scene || !in) return set_err(ctx, RT_ERR_INVALID_ARG, "NULL scene or params");
    rt_params p = *in;
    if (p.divisions == 0) p.divisions = 1;
    if (p.spp == 0) p.spp = 100;                 // main.rs:51
    if (p.max_bounces == 0) p.max_bounces = 10;  // main.rs:39
    if (p.aperture == 0.0f) p.aperture = 0.1f;   // main.rs:45
    if (p.focus_distance == 0.0f) p.focus_distance = 1.0f;
    if (p.field_of_view == 0.0f) p.field_of_view = 3.14159265358979323846f / 2.0f;  // PI / 2f32
    if (p.focal_length == 0.0f) p.focal_length = 1.0f;
    if (p.width == 0 || p.height == 0) return set_err(ctx, RT_ERR_INVALID_ARG, "zero width or height");
    if (p.height % p.divisions != 0)
        return set_err(ctx, RT_ERR_INVALID_ARG,
                       "height %u is not a multiple of divisions %u (the reference silently drops rows)", p.height,
                       p.divisions);
    if (p.division_no >= p.divisions) return set_err(ctx, RT_ERR_INVALID_ARG, "division_no out of range");
    if (p.max_bounces + 1 > (uint32_t)MAX_PATH)
        return set_err(ctx, RT_ERR_UNSUPPORTED, "max_bounces %u exceeds this build's limit %d", p.max_bounces, MAX_PATH - 1);
    if ((uint64_t)p.width * p.height > 0x7fffffffu) return set_err(ctx, RT_ERR_UNSUPPORTED, "frame too large");
    if (p.intersector > RT_INTERSECT_BVH) return set_err(ctx, RT_ERR_INVALID_ARG, "unknown intersector");
    r->p = p;
    r->isect = p.intersector == RT_INTERSECT_AUTO ? (scene->n <= kBruteMaxPrims ? RT_INTERSECT_BRUTE : RT_INTERSECT_BVH)
                                                  : (int)p.intersector;
    // Camera::new (camera.rs:19-47) with the arguments of main.rs:42-50, single f32 ops in order
    DevCamera& c = r->cam;
    const float aspect = (float)p.width / (float)p.height;
    const float image_height = (float)p.height;
    const float vh = 2.0f * tanf(p.field_of_view / 2.0f);
    const float vw = aspect * vh;
    const float hor[3] = {vw, 0.0f, 0.0f}, ver[3] = {0.0f, vh, 0.0f}, fl[3] = {0.0f, 0.0f, p.focal_length};
    for (int a = 0; a < 3; a++) {
        c.org[a] = p.cam_origin[a];
        c.hor[a] = hor[a];
        c.ver[a] = ver[a];
        c.llc[a] = ((p.cam_origin[a] - hor[a] / 2.0f) - ver[a] / 2.0f) - fl[a];
    }
    c.lens_radius = p.aperture / 2.0f;
    c.u_den = aspect * image_height - 1.0f;
    c.v_den = image_height - 1.0f;
    c.focus = p.focus_distance;
    return RT_OK;
}

int ensure_out(rt_ctx* ctx, size_t bytes) {
    if (ctx->d_out_bytes >= bytes) return RT_OK;
    if (ctx->d_out) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaFree(ctx->d_out));
        ctx->d_out = nullptr;
        ctx->d_out_bytes = 0;
    }
    CK(ctx, cudaMalloc(&ctx->d_out, bytes));
    ctx->d_out_bytes = bytes;
    return RT_OK;
}

// Launches the render of global rows [row0,row1) (tiles of this rank only) into `dst` whose row 0 is
// global row out_row0.  Leaves timing events recorded on the stream.
int launch(rt_ctx* ctx, const rt_scene* scene, const Resolved& r, uint32_t row0, uint32_t row1, uint32_t tile_rank,
           uint32_t tile_ranks, uint8_t* dst, uint32_t out_row0, LaunchInfo* info) {
    DevParams pr{};
    pr.width = r.p.width;
    pr.height = r.p.height;
    pr.row0 = row0;
    pr.row1 = row1;
    pr.spp = r.p.spp;
    pr.depth = r.p.max_bounces + 1;
    pr.seed = r.p.seed;
    pr.tile_rank = tile_rank;
    pr.tile_ranks = tile_ranks;
    pr.out = dst;
    pr.out_row0 = out_row0;
    pr.tiles_x = (r.p.width + TILE_W - 1) / TILE_W;
    pr.tiles_y = (row1 - row0 + TILE_H - 1) / TILE_H;
    pr.ray_dump = ctx->dump_rays;
    pr.ray_dump_n = ctx->dump_n;
    pr.ray_dump_cap = ctx->dump_cap;
    pr.counters = ctx->d_ctr;
    pr.tile_counter = reinterpret_cast<unsigned int*>(ctx->d_ctr + NUM_COUNTERS);
    CK(ctx, cudaMemsetAsync(ctx->d_ctr, 0, (NUM_COUNTERS + 1) * sizeof(unsigned long long), ctx->stream));
    CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (use_wq(r.isect, pr)) {
        CK(ctx, launch_wq(scene->dev, r.cam, pr, r.p.collect_counters != 0, ctx->sm_count, ctx->smem_optin, ctx->stream,
                          &ctx->wq, info));
    } else if (use_wavefront(r.isect)) {
        CK(ctx, launch_wavefront(scene->dev, r.cam, pr, r.p.collect_counters != 0, ctx->sm_count, ctx->smem_optin,
                                 ctx->stream, &ctx->wave, info));
    } else {
        CK(ctx, launch_render(scene->dev, r.cam, pr, r.isect, r.p.collect_counters != 0, ctx->sm_count,
                              ctx->smem_optin, ctx->stream, info));
    }
    CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    return RT_OK;
}

int finish_stats(rt_ctx* ctx, const Resolved& r, uint64_t pixels, rt_stats* st,
                 std::chrono::steady_clock::time_point t0, const LaunchInfo& li) {
    if (!st) return RT_OK;
    CK(ctx, cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, NUM_COUNTERS * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    memset(st, 0, sizeof *st);
    const unsigned long long* c = ctx->h_ctr;
    st->rays = c[CTR_RAYS];
    st->primary = pixels * r.p.spp;
    st->slab_tests = c[CTR_SLAB];
    st->sphere_tests = c[CTR_SPH_TEST];
    st->sphere_exact = c[CTR_SPH_EXACT];
    st->sphere_hits = c[CTR_SPH_HIT];
    st->tri_tests = c[CTR_TRI_TEST];
    st->tri_stage[0] = c[CTR_TRI_S1];
    st->tri_stage[1] = c[CTR_TRI_S2];
    st->tri_stage[2] = c[CTR_TRI_S3];
    st->tri_hits = c[CTR_TRI_HIT];
    st->shades_sphere = c[CTR_SHADE_SPH];
    st->shades_tri = c[CTR_SHADE_TRI];
    st->emissive = c[CTR_EMISSIVE];
    st->sky = c[CTR_SKY];
    st->active_lane_iters = c[CTR_ACTIVE_LANES];
    st->total_lane_iters = c[CTR_TOTAL_LANES];
    float ms = 0.0f;
    CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    st->kernel_ms = ms;
    st->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    st->intersector_used = (uint32_t)r.isect;
    st->kernel_launches = li.launches;
    st->grid_ctas = li.grid;
    st->cta_threads = li.threads;
    st->ctas_per_sm = (uint32_t)li.ctas_per_sm;
    st->scene_in_smem = li.scene_in_smem ? 1u : 0u;
    st->dyn_smem_bytes = (uint32_t)li.dyn_smem;
    return RT_OK;
}

int render_rows_to_host(rt_ctx* ctx, const rt_scene* scene, const Resolved& r, uint32_t row0, uint32_t row1,
                        uint8_t* out_rgb, size_t out_len, rt_stats* stats) {
    auto t0 = std::chrono::steady_clock::now();
    const size_t bytes = (size_t)(row1 - row0) * r.p.width * 3;
    if (!out_rgb) return set_err(ctx, RT_ERR_INVALID_ARG, "out_rgb is NULL");
    if (out_len != bytes)
        return set_err(ctx, RT_ERR_INVALID_ARG, "out_len %zu != (rows %u * width %u * 3) = %zu", out_len, row1 - row0,
                       r.p.width, bytes);
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_out(ctx, bytes);
    if (rc) return rc;
    LaunchInfo li;
    rc = launch(ctx, scene, r, row0, row1, 0, 1, ctx->d_out, row0, &li);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(out_rgb, ctx->d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return finish_stats(ctx, r, (uint64_t)(row1 - row0) * r.p.width, stats, t0, li);
}

}  // namespace

int rt_render_division(rt_ctx* ctx, const rt_scene* scene, const rt_params* params, uint8_t* out_rgb, size_t out_len,
                       rt_stats* stats) {
    if (!ctx) return RT_ERR_INVALID_ARG;
    Resolved r;
    int rc = resolve(ctx, scene, params, &r);
    if (rc) return rc;
    const uint32_t band = r.p.height / r.p.divisions;  // main.rs:55-56
    const uint32_t row0 = band * r.p.division_no;      // main.rs:66-68
    return render_rows_to_host(ctx, scene, r, row0, row0 + band, out_rgb, out_len, stats);
}

int rt_render_frame(rt_ctx* ctx, const rt_scene* scene, const rt_params* params, uint8_t* out_rgb, size_t out_len,
                    rt_stats* stats) {
    if (!ctx) return RT_ERR_INVALID_ARG;
    Resolved r;
    rt_params p;
    if (params) {
        p = *params;
        p.division_no = 0;
    }
    int rc = resolve(ctx, scene, params ? &p : nullptr, &r);
    if (rc) return rc;
    return render_rows_to_host(ctx, scene, r, 0, r.p.height, out_rgb, out_len, stats);
}

int rt_render_tiles_device(rt_ctx* ctx, const rt_scene* scene, const rt_params* params, uint32_t tile_rank,
                           uint32_t tile_ranks, void* frame_dev, int sync, rt_stats* stats) {
    if (!ctx) return RT_ERR_INVALID_ARG;
    auto t0 = std::chrono::steady_clock::now();
    Resolved r;
    rt_params p;
    if (params) {
        p = *params;
        p.division_no = 0;
    }
    int rc = resolve(ctx, scene, params ? &p : nullptr, &r);
    if (rc) return rc;
    if (!frame_dev) return set_err(ctx, RT_ERR_INVALID_ARG, "frame_dev is NULL");
    if (tile_ranks == 0 || tile_rank >= tile_ranks) return set_err(ctx, RT_ERR_INVALID_ARG, "bad tile_rank/tile_ranks");
    CK(ctx, cudaSetDevice(ctx->device));
    LaunchInfo li;
    rc = launch(ctx, scene, r, 0, r.p.height, tile_rank, tile_ranks, (uint8_t*)frame_dev, 0, &li);
    if (rc) return rc;
    if (sync) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        // pixel count of this rank's tiles is not needed by callers; report the frame total / ranks
        return finish_stats(ctx, r, (uint64_t)r.p.width * r.p.height / tile_ranks, stats, t0, li);
    }
    return RT_OK;
}

// -------------------------------------------------------------------------------------------------
// Memory helpers
// -------------------------------------------------------------------------------------------------
int rt_host_alloc(rt_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return RT_ERR_INVALID_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMallocHost(out, bytes ? bytes : 1));
    return RT_OK;
}
void rt_host_free(rt_ctx* ctx, void* p) {
    if (ctx) cudaSetDevice(ctx->device);
    if (p) cudaFreeHost(p);
}

int rt_frame_alloc(rt_ctx* ctx, size_t bytes, void** dev_out, uint8_t handle_out[64]) {
    if (!ctx || !dev_out) return RT_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMalloc(dev_out, bytes ? bytes : 1));
    if (handle_out) {
        cudaIpcMemHandle_t h;
        CK(ctx, cudaIpcGetMemHandle(&h, *dev_out));
        memcpy(handle_out, &h, 64);
    }
    return RT_OK;
}
int rt_frame_open(rt_ctx* ctx, const uint8_t handle[64], void** dev_out) {
    if (!ctx || !handle || !dev_out) return RT_ERR_INVALID_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(ctx, cudaIpcOpenMemHandle(dev_out, h, cudaIpcMemLazyEnablePeerAccess));
    return RT_OK;
}
int rt_frame_close(rt_ctx* ctx, void* dev) {
    if (!ctx) return RT_ERR_INVALID_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaIpcCloseMemHandle(dev));
    return RT_OK;
}
int rt_frame_free(rt_ctx* ctx, void* dev) {
    if (!ctx) return RT_ERR_INVALID_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaFree(dev));
    return RT_OK;
}
int rt_frame_download(rt_ctx* ctx, const void* frame_dev, uint8_t* out_rgb, size_t bytes) {
    if (!ctx || !frame_dev || !out_rgb) return RT_ERR_INVALID_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpyAsync(out_rgb, frame_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

// -------------------------------------------------------------------------------------------------
// FP32 roofline denominator
// -------------------------------------------------------------------------------------------------
int rt_measure_fp32_peak(rt_ctx* ctx, double* tflops_out, float* ms_out) {
    if (!ctx || !tflops_out) return RT_ERR_INVALID_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    if (!ctx->d_scratch) CK(ctx, cudaMalloc(&ctx->d_scratch, (size_t)ctx->sm_count * 8 * 256 * sizeof(float)));
    const int iters = 4096;
    float best = std::numeric_limits<float>::max();
    for (int rep = 0; rep < 4; rep++) {
        CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        CK(ctx, launch_fp32_peak(ctx->d_scratch, ctx->sm_count, iters, ctx->stream));
        CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.0f;
        CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double flops = (double)ctx->sm_count * 8 * 256 * (double)iters * 16 * 8 * 2;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return RT_OK;
}

// -------------------------------------------------------------------------------------------------
// Trace-only benchmark (csrc/rt_trace_bench.cuh): record the queries of one frame, then time the nearest-hit query
// alone in two forms over the recorded rays and compare their answers.
// -------------------------------------------------------------------------------------------------
int rt_debug_trace_bench(rt_ctx* ctx, const rt_scene* scene, const rt_params* params, uint64_t max_rays, int with_big,
                         uint64_t* n_rays_out, float* ms_while_while, float* ms_state_machine, uint64_t* mismatches_out) {
    if (!ctx || !scene || !params || !n_rays_out || !ms_while_while || !ms_state_machine || !mismatches_out)
        return RT_ERR_INVALID_ARG;
    if (max_rays == 0) return set_err(ctx, RT_ERR_INVALID_ARG, "max_rays is 0");
    rt_params p = *params;
    p.collect_counters = 1;
    p.intersector = RT_INTERSECT_BVH;
    Resolved r;
    int rc = resolve(ctx, scene, &p, &r);
    if (rc) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    rc = ensure_out(ctx, (size_t)r.p.width * r.p.height * 3);
    if (rc) return rc;
    float4* d_rays = nullptr;
    unsigned long long* d_cnt = nullptr;  // [0] recorded queries, [1] ticket
    int2 *d_out0 = nullptr, *d_out1 = nullptr;
    auto cleanup = [&] {
        ctx->dump_rays = nullptr;
        ctx->dump_n = nullptr;
        ctx->dump_cap = 0;
        cudaFree(d_rays);
        cudaFree(d_cnt);
        cudaFree(d_out0);
        cudaFree(d_out1);
    };
#define CKC(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            cleanup();                                                                                 \
            return set_err(ctx, RT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));                \
        }                                                                                              \
    } while (0)
    CKC(cudaMalloc(&d_rays, (size_t)max_rays * 32));
    CKC(cudaMalloc(&d_cnt, 16));
    CKC(cudaMemsetAsync(d_cnt, 0, 16, ctx->stream));
    ctx->dump_rays = d_rays;
    ctx->dump_n = d_cnt;
    ctx->dump_cap = max_rays;
    LaunchInfo li;
    rc = launch(ctx, scene, r, 0, r.p.height, 0, 1, ctx->d_out, 0, &li);
    ctx->dump_rays = nullptr;
    ctx->dump_n = nullptr;
    ctx->dump_cap = 0;
    if (rc) {
        cleanup();
        return rc;
    }
    unsigned long long recorded = 0;
    CKC(cudaMemcpyAsync(&recorded, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    const unsigned long long n = recorded < max_rays ? recorded : max_rays;
    if (n == 0) {
        cleanup();
        return set_err(ctx, RT_ERR_INVALID_ARG, "no query was recorded");
    }
    if (with_big >= 2) {  // with_big = 2 | 3: as 0 | 1, on the rays sorted by direction octant and origin cell
        float4* d_sorted = nullptr;
        CKC(cudaMalloc(&d_sorted, (size_t)n * 32));
        const float lo[3] = {-30.0f, -6.0f, -40.0f}, hi[3] = {30.0f, 12.0f, 2.0f};  // the BASELINE scenes' extent
        const cudaError_t se = sort_rays_device(d_rays, n, lo, hi, d_sorted, ctx->stream);
        if (se != cudaSuccess) cudaFree(d_sorted);
        CKC(se);
        cudaFree(d_rays);
        d_rays = d_sorted;
        with_big -= 2;
    }
    CKC(cudaMalloc(&d_out0, (size_t)n * 8));
    CKC(cudaMalloc(&d_out1, (size_t)n * 8));
    float best[2] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    for (int variant = 0; variant < 2; variant++)
        for (int rep = 0; rep < 3; rep++) {
            CKC(cudaEventRecord(ctx->ev0, ctx->stream));
            CKC(launch_trace_bench(scene->dev, variant, with_big != 0, d_rays, n, d_cnt + 1, variant ? d_out1 : d_out0,
                                   ctx->sm_count, ctx->smem_optin, ctx->stream));
            CKC(cudaEventRecord(ctx->ev1, ctx->stream));
            CKC(cudaStreamSynchronize(ctx->stream));
            float ms = 0.0f;
            CKC(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            if (ms < best[variant]) best[variant] = ms;
        }
    std::vector<int2> h0((size_t)n), h1((size_t)n);
    CKC(cudaMemcpy(h0.data(), d_out0, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CKC(cudaMemcpy(h1.data(), d_out1, (size_t)n * 8, cudaMemcpyDeviceToHost));
#undef CKC
    uint64_t bad = 0;
    for (size_t i = 0; i < (size_t)n; i++) bad += (h0[i].x != h1[i].x || h0[i].y != h1[i].y) ? 1 : 0;
    cleanup();
    *n_rays_out = n;
    *ms_while_while = best[0];
    *ms_state_machine = best[1];
    *mismatches_out = bad;
    return RT_OK;
}

}  // extern "C"
