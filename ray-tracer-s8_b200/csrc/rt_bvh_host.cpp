// rt_bvh_host.cpp — host-side BVH build with the reference's topology.
//
// The GPU path may traverse in any order (the BVH is a conservative cull), but two things must follow
// the reference's tree (ray-tracer-slave/local-dependencies/bvh/src/bvh/bvh_impl.rs:229-364):
//   * the DFS (left-first) leaf order, which decides exact-distance ties in the nearest-hit min_by
//     (ray-tracer-slave/src/shapes/mod.rs:177-182, bvh_impl.rs:373-398), and
//   * the N = 1 case (root is a leaf, always a candidate).
// So the builder makes the same decisions — centroid-bounds largest axis, 6 SAH buckets with
// bucket = ((c - cmin)/extent * 5.99) as usize, first strictly lower cost wins, halve the list when the
// centroid extent is < 1e-5, one primitive per leaf — evaluated with the same f32 operations, but works
// in place on one index array (stable counting sort by bucket = the reference's bucket concatenation)
// and emits only inner nodes, children encoded as codes, in DFS pre-order.
//
// Build with -ffp-contract=off (no FMA contraction on the host either).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "rt_host.h"

namespace rtb {
namespace {

constexpr int kBuckets = 6;
constexpr float kEpsilon = 0.00001f;  // bvh::EPSILON (lib.rs:80)

struct Bounds {
    float lo[3], hi[3];
    void clear() {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::numeric_limits<float>::infinity();
            hi[a] = -std::numeric_limits<float>::infinity();
        }
    }
    void join(const Box& b) {  // AABB::join — f32::min/max (NaN-ignoring)
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], b.min[a]);
            hi[a] = fmaxf(hi[a], b.max[a]);
        }
    }
    void join(const Bounds& b) {
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], b.lo[a]);
            hi[a] = fmaxf(hi[a], b.hi[a]);
        }
    }
    void grow(const float p[3]) {  // AABB::grow
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], p[a]);
            hi[a] = fmaxf(hi[a], p[a]);
        }
    }
    bool empty() const { return lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2]; }
    float area() const {  // AABB::surface_area: 2*(sx*sy + sx*sz + sy*sz)
        float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
        return 2.0f * (sx * sy + sx * sz + sy * sz);
    }
    Box box() const {
        Box b;
        for (int a = 0; a < 3; a++) {
            b.min[a] = lo[a];
            b.max[a] = hi[a];
        }
        return b;
    }
};

inline void centre_of(const Box& b, float c[3]) {  // AABB::center: min + (max - min)/2
    for (int a = 0; a < 3; a++) c[a] = b.min[a] + ((b.max[a] - b.min[a]) / 2.0f);
}

struct Builder {
    const std::vector<Box>& boxes;
    std::vector<uint32_t> idx, tmp;
    std::vector<uint8_t> bucket_of;
    std::vector<float> centre;  // AABB::center of every shape, computed once (same f32 operations every time)
    HostBVH* out;
    std::string* err;
    bool failed = false;

    Builder(const std::vector<Box>& b, HostBVH* o, std::string* e) : boxes(b), out(o), err(e) {
        idx.resize(b.size());
        tmp.resize(b.size());
        bucket_of.resize(b.size());
        centre.resize(3 * b.size());
        for (size_t i = 0; i < b.size(); i++) {
            idx[i] = (uint32_t)i;
            centre_of(b[i], &centre[3 * i]);
        }
    }

    void fail(const char* msg) {
        if (!failed && err) *err = msg;
        failed = true;
    }

    // Builds the subtree over idx[lo,hi) and returns its code.
    int32_t build(size_t lo, size_t hi, uint32_t depth) {
        if (failed) return 0;
        if (depth > 4096) {
            fail("BVH deeper than 4096 levels");
            return 0;
        }
        const size_t n = hi - lo;
        if (n == 1) {
            out->leaf_order.push_back(idx[lo]);
            if (depth > out->depth) out->depth = depth;
            return ~(int32_t)idx[lo];
        }
        Bounds all, cent;
        all.clear();
        cent.clear();
        for (size_t i = lo; i < hi; i++) {
            all.join(boxes[idx[i]]);
            cent.grow(&centre[3 * (size_t)idx[i]]);
        }
        const int32_t me = (int32_t)out->inner.size();
        out->inner.push_back(HostNode{});

        // AABB::largest_axis (aabb.rs:570-580)
        const float sx = cent.hi[0] - cent.lo[0], sy = cent.hi[1] - cent.lo[1], sz = cent.hi[2] - cent.lo[2];
        const int axis = (sx > sy && sx > sz) ? 0 : (sy > sz ? 1 : 2);
        const float extent = cent.hi[axis] - cent.lo[axis];

        size_t mid;
        Bounds bl, br;
        if (extent < kEpsilon) {
            mid = lo + n / 2;
            bl.clear();
            br.clear();
            for (size_t i = lo; i < mid; i++) bl.join(boxes[idx[i]]);
            for (size_t i = mid; i < hi; i++) br.join(boxes[idx[i]]);
        } else {
            Bounds bb[kBuckets];
            size_t cnt[kBuckets] = {0, 0, 0, 0, 0, 0};
            for (auto& b : bb) b.clear();
            for (size_t i = lo; i < hi; i++) {
                const float* c = &centre[3 * (size_t)idx[i]];
                const float rel = (c[axis] - cent.lo[axis]) / extent;
                const float scaled = rel * ((float)kBuckets - 0.01f);
                // `as usize`: truncate, saturate, NaN → 0
                size_t k = (scaled == scaled && scaled > 0.0f) ? (scaled >= 1.8e19f ? (size_t)-1 : (size_t)scaled) : 0;
                if (k >= (size_t)kBuckets) {
                    fail("bucket index out of range (non-finite bounds)");
                    return 0;
                }
                bucket_of[i] = (uint8_t)k;
                cnt[k]++;
                bb[k].join(boxes[idx[i]]);
            }
            int best = 0;
            float best_cost = std::numeric_limits<float>::infinity();
            bl.clear();
            br.clear();
            const float parent_area = all.area();
            for (int s = 0; s < kBuckets - 1; s++) {
                Bounds l, r;
                l.clear();
                r.clear();
                size_t nl = 0, nr = 0;
                for (int k = 0; k <= s; k++) {
                    l.join(bb[k]);
                    nl += cnt[k];
                }
                for (int k = s + 1; k < kBuckets; k++) {
                    r.join(bb[k]);
                    nr += cnt[k];
                }
                const float cost = ((float)nl * l.area() + (float)nr * r.area()) / parent_area;
                if (cost < best_cost) {
                    best = s;
                    best_cost = cost;
                    bl = l;
                    br = r;
                }
            }
            // stable counting sort of idx[lo,hi) by bucket == concatenating the bucket vectors
            size_t start[kBuckets], pos = 0;
            for (int k = 0; k < kBuckets; k++) {
                start[k] = pos;
                pos += cnt[k];
            }
            for (size_t i = lo; i < hi; i++) tmp[lo + start[bucket_of[i]]++] = idx[i];
            std::memcpy(&idx[lo], &tmp[lo], n * sizeof(uint32_t));
            size_t nl = 0;
            for (int k = 0; k <= best; k++) nl += cnt[k];
            mid = lo + nl;
        }
        if (bl.empty() || br.empty() || mid == lo || mid == hi) {  // reference: assert!(!child_aabb.is_empty())
            fail("degenerate split (empty child bounds)");
            return 0;
        }
        const int32_t l = build(lo, mid, depth + 1);
        const int32_t r = build(mid, hi, depth + 1);
        HostNode& nd = out->inner[me];
        nd.box_l = bl.box();
        nd.box_r = br.box();
        nd.left = l;
        nd.right = r;
        return me;
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// Traversal-quality tree: binned SAH over all three axes (16 bins), one primitive per leaf.  Used only to
// CULL on the GPU (any conservative tree gives the reference's nearest hit; exact-distance ties are broken by
// the DFS rank of the reference-topology tree, which is still built).  Leaves are codes ~index like above.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct SahBuilder {
    const std::vector<Box>& boxes;
    std::vector<uint32_t> idx;
    std::vector<float> cen;  // 3 per primitive
    HostBVH* out;
    static constexpr int kBins = 16;

    SahBuilder(const std::vector<Box>& b, HostBVH* o) : boxes(b), out(o) {
        idx.resize(b.size());
        cen.resize(3 * b.size());
        for (size_t i = 0; i < b.size(); i++) {
            idx[i] = (uint32_t)i;
            for (int a = 0; a < 3; a++) cen[3 * i + a] = 0.5f * (b[i].min[a] + b[i].max[a]);
        }
    }
    static float area(const Bounds& b) { return b.empty() ? 0.0f : b.area(); }

    int32_t build(size_t lo, size_t hi, uint32_t depth) {
        const size_t n = hi - lo;
        if (n == 1) {
            out->leaf_order.push_back(idx[lo]);
            if (depth > out->depth) out->depth = depth;
            return ~(int32_t)idx[lo];
        }
        float cmin[3], cmax[3];
        for (int a = 0; a < 3; a++) { cmin[a] = std::numeric_limits<float>::infinity(); cmax[a] = -cmin[a]; }
        for (size_t i = lo; i < hi; i++)
            for (int a = 0; a < 3; a++) {
                cmin[a] = fminf(cmin[a], cen[3 * idx[i] + a]);
                cmax[a] = fmaxf(cmax[a], cen[3 * idx[i] + a]);
            }
        int best_axis = -1, best_split = 0;
        float best_cost = std::numeric_limits<float>::infinity();
        for (int a = 0; a < 3; a++) {
            const float ext = cmax[a] - cmin[a];
            if (!(ext > 0.0f)) continue;
            Bounds bb[kBins];
            size_t cnt[kBins] = {0};
            for (auto& b : bb) b.clear();
            const float scale = (float)kBins * (1.0f - 1e-6f) / ext;
            for (size_t i = lo; i < hi; i++) {
                int k = (int)((cen[3 * idx[i] + a] - cmin[a]) * scale);
                k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
                cnt[k]++;
                bb[k].join(boxes[idx[i]]);
            }
            Bounds racc;
            racc.clear();
            float rarea[kBins];
            size_t rcnt[kBins];
            size_t c = 0;
            for (int k = kBins - 1; k > 0; k--) {
                racc.join(bb[k]);
                c += cnt[k];
                rarea[k] = area(racc);
                rcnt[k] = c;
            }
            Bounds lacc;
            lacc.clear();
            size_t lc = 0;
            for (int k = 0; k < kBins - 1; k++) {
                lacc.join(bb[k]);
                lc += cnt[k];
                if (lc == 0 || rcnt[k + 1] == 0) continue;
                const float cost = (float)lc * area(lacc) + (float)rcnt[k + 1] * rarea[k + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = k; }
            }
        }
        size_t mid;
        if (best_axis < 0) {
            mid = lo + n / 2;  // coincident centroids
        } else {
            const float ext = cmax[best_axis] - cmin[best_axis];
            const float scale = (float)kBins * (1.0f - 1e-6f) / ext;
            auto first = idx.begin() + lo, last = idx.begin() + hi;
            auto it = std::stable_partition(first, last, [&](uint32_t p) {
                int k = (int)((cen[3 * p + best_axis] - cmin[best_axis]) * scale);
                k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
                return k <= best_split;
            });
            mid = (size_t)(it - idx.begin());
            if (mid == lo || mid == hi) mid = lo + n / 2;
        }
        const int32_t me = (int32_t)out->inner.size();
        out->inner.push_back(HostNode{});
        Bounds bl, br;
        bl.clear();
        br.clear();
        for (size_t i = lo; i < mid; i++) bl.join(boxes[idx[i]]);
        for (size_t i = mid; i < hi; i++) br.join(boxes[idx[i]]);
        const int32_t l = build(lo, mid, depth + 1);
        const int32_t r = build(mid, hi, depth + 1);
        HostNode& nd = out->inner[me];
        nd.box_l = bl.box();
        nd.box_r = br.box();
        nd.left = l;
        nd.right = r;
        return me;
    }
};
}  // namespace

bool build_bvh_sah(const std::vector<Box>& boxes, HostBVH* out) {
    out->inner.clear();
    out->leaf_order.clear();
    out->depth = 0;
    if (boxes.empty()) return false;
    out->inner.reserve(boxes.size());
    out->leaf_order.reserve(boxes.size());
    SahBuilder b(boxes, out);
    out->root = b.build(0, boxes.size(), 0);
    out->node_count = (uint32_t)(out->inner.size() + out->leaf_order.size());
    return true;
}

bool build_bvh(const std::vector<Box>& boxes, HostBVH* out, std::string* err) {
    out->inner.clear();
    out->leaf_order.clear();
    out->depth = 0;
    if (boxes.empty()) {
        if (err) *err = "empty scene";
        return false;
    }
    out->inner.reserve(boxes.size());
    out->leaf_order.reserve(boxes.size());
    Builder b(boxes, out, err);
    out->root = b.build(0, boxes.size(), 0);
    out->node_count = (uint32_t)(out->inner.size() + out->leaf_order.size());
    return !b.failed;
}

}  // namespace rtb
