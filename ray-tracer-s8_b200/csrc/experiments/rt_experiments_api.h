/* rt_experiments_api.h — entry points that exist only in lib/librt_b200_exp.so (built with -DRT_B200_EXPERIMENTS). */
#ifndef RT_B200_EXPERIMENTS_API_H
#define RT_B200_EXPERIMENTS_API_H
#include "../../../include/rt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Trace-only benchmark (development aid, csrc/experiments/rt_trace_bench.cuh; with_big 2 / 3 = 0 / 1 on rays sorted by
 * octant + cell): renders one frame while recording up to max_rays of its nearest-hit queries, then times the query alone
 * over the recorded rays as (a) the product kernel's while-while traversal and (b) a ballot-scheduled state machine with
 * dynamic fetch, and counts rays whose answers differ.  with_big = 0 leaves the split layout's big primitives out of
 * both.  The scene must fit shared memory. */
int rt_debug_trace_bench(rt_ctx* ctx, const rt_scene* scene, const rt_params* params, uint64_t max_rays, int with_big,
                         uint64_t* n_rays_out, float* ms_while_while, float* ms_state_machine, uint64_t* mismatches_out);
#ifdef __cplusplus
}
#endif
#endif
