// rt_host.h — host-side declarations shared by rt_api.cu, rt_kernels.cu and rt_bvh_host.cpp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/rt_b200.h"

namespace rtb {

struct DevScene;
struct DevCamera;
struct DevParams;

struct LaunchInfo {
    unsigned grid = 0, threads = 0;
    size_t dyn_smem = 0;
    int ctas_per_sm = 0;
    bool scene_in_smem = false;
    unsigned launches = 1;
};

// device buffers of the wavefront pipeline (owned by the context, grown on demand)
struct WaveBuffers {
    void* slots = nullptr;          // WSlot[capacity]
    uint32_t* q_ray = nullptr;      // [capacity] each
    uint32_t* q_hit = nullptr;
    uint32_t* q_miss = nullptr;
    uint32_t* path_ext = nullptr;   // [(ext_depth) * capacity] path entries beyond the 8 kept in the slot
    void* counters = nullptr;       // WaveCounters[2]
    size_t capacity = 0, ext_entries = 0;
};
void free_wave_buffers(WaveBuffers* wb);

// device buffer of the warp-private wavefront kernel: per-chain state (rt_kernel_wq.cuh), grown on demand
struct WqBuffers {
    void* state = nullptr;
    size_t bytes = 0;
};
void free_wq_buffers(WqBuffers* b);
bool use_wq(int isect, const DevParams& pr);
cudaError_t launch_wq(const DevScene& sc, const DevCamera& cam, const DevParams& pr, bool count, int sm_count,
                      int smem_optin, cudaStream_t stream, WqBuffers* wb, LaunchInfo* info);

// rt_bvh_device.cu: LBVH + refit of the traversal tree on the device (scratch owned by the context)
struct DeviceBuild {
    void* mem = nullptr;
    size_t bytes = 0;
};
void free_device_build(DeviceBuild* b);
cudaError_t build_lbvh_device(DeviceBuild* buf, const float* h_boxes, const uint32_t* h_pid_of, uint32_t n,
                              float4* lnode_abc, int2* lnode_d, uint32_t* depth_out, cudaStream_t stream);

cudaError_t sort_rays_device(const float4* rays, unsigned long long n, const float lo[3], const float hi[3], float4* out,
                             cudaStream_t stream);

// rt_kernels.cu
cudaError_t launch_wavefront(const DevScene& sc, const DevCamera& cam, const DevParams& pr, bool count, int sm_count,
                             int smem_optin, cudaStream_t stream, WaveBuffers* wb, LaunchInfo* info);
bool use_wavefront(int isect);
bool legacy_node_arrays_needed();  // true when RT_B200_BVH_KERNEL selects a kernel that reads node_* / cnode_*
cudaError_t launch_render(const DevScene& sc, const DevCamera& cam, const DevParams& pr, int isect, bool count,
                          int sm_count, int smem_optin, cudaStream_t stream, LaunchInfo* info);
cudaError_t launch_trace_bench(const DevScene& sc, int variant, bool with_big, const float4* rays, unsigned long long n,
                               unsigned long long* ticket, int2* out, int sm_count, int smem_optin, cudaStream_t stream);
cudaError_t launch_fp32_peak(float* scratch, int sm_count, int iters, cudaStream_t stream);
size_t scene_smem_bytes(const DevScene& sc, int isect);

// ---- rt_bvh_host.cpp: SAH BVH with the reference's topology ------------------------------------------
struct Box {
    float min[3], max[3];
};
struct HostNode {       // inner node; children are codes: >= 0 inner index, < 0 leaf of world position ~code
    Box box_l, box_r;
    int32_t left, right;
};
struct HostBVH {
    std::vector<HostNode> inner;        // DFS pre-order
    std::vector<uint32_t> leaf_order;   // rank → world position
    int32_t root = 0;                   // code
    uint32_t depth = 0;                 // deepest leaf (root = 0)
    uint32_t node_count = 0;            // inner + leaves, = 2n-1
};
// boxes[i] = AABB of the primitive at world position i.  Returns false (with a message) when the
// reference's build would panic or never terminate.
bool build_bvh(const std::vector<Box>& boxes, HostBVH* out, std::string* err);
// 3-axis, 16-bin SAH tree over the same primitives, for culling only (leaf order is NOT the reference's).
bool build_bvh_sah(const std::vector<Box>& boxes, HostBVH* out);

}  // namespace rtb
