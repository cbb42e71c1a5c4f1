// ptcuda.cu -- libptcuda: C ABI (include/ptcuda.h) + host driver for the sm_100a path tracer.
//
// Replaces the host side of the reference's OpenCL path (reference internal/ocl/ocltracer.go):
//   Trace()            :100-226  -> ptc_open + ptc_trace + ptc_read (ptc_render does all three)
//   prepareTextures()  :228-254  -> upload_textures()
//   computeBatch()     :256-376  -> one launch per device for the whole frame instead of H/4
//                                   create/upload/launch/finish/readback rounds
// and cmd/pt/main.go:98-112 listDevices() -> ptc_device_count / ptc_device_name.
// There is no CPU path in this library: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "../../include/ptcuda.h"
#include "../../include/ptwire.h"
#include "kernels/trace.cuh"

namespace {

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

struct Error : std::exception {
    std::string msg;
    explicit Error(std::string m) : msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};
[[noreturn]] void fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(buf);
}
#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

void set_err(char* err, int errlen, const char* msg) {
    if (err && errlen > 0) std::snprintf(err, size_t(errlen), "%s", msg);
}

// ---- flattened scene (host copy), templated on the arithmetic type -----------------------------
template <typename R> struct HostScene {
    ptk::DObjHot<R> hot[ptk::kMaxObjects];
    ptk::DFast<R> fast[ptk::kFastSlots];
    ptk::DFast<R> fast2[ptk::kFastSlots];
    int fast_n[4], fast_obj[ptk::kFastSlots + 1], fast_kind[ptk::kFastSlots];
    int slow_obj[ptk::kMaxObjects], slow_kind[ptk::kMaxObjects], n_slow = 0;
    ptk::DFast<R> slow_rec[ptk::kMaxObjects], slow_rec2[ptk::kMaxObjects];
    int mesh_obj[ptk::kMaxObjects], n_mesh = 0;
    int stack_need = 0;        // deepest deferred-child stack any mesh of the scene can need
    ptk::DLight<R> light[ptk::kMaxObjects];
    int n_lights = 0;
    std::vector<ptk::DObjShade<R>> shade;
    std::vector<ptk::DMesh<R>> mesh;              // one per object
    std::vector<R> lens;       // sunflower lens points, 2 per sample (empty without depth of field)
    std::vector<ptk::V4<R>> node_lo, node_hi;     // reference BVH boxes in the reference's visiting order
    std::vector<int> node_parent;
    std::vector<ptk::V4<R>> wide;                 // rebuilt BVH: 8 children x (lo.xyz + child code, hi.xyz) per node
    std::vector<ptk::V4<R>> tri_test, tri_shade;  // 3 records per slot, slots in leaf order
    std::vector<int2> tri_info;                   // slot -> (rank in the reference's recording order, reference node)
    ptk::DCam<R> cam;
    int mesh_depth = 0;
};

// A triangle of a group object as the reference reaches it: `ref_node` is the (re-emitted) node that
// holds it, `rank` its position in the reference's recording order (root children in order, each walked
// node-first, then children[0]'s subtree, then children[1]'s: tracer.cl:624-714; a node's triangles by
// increasing offset, :637).
struct RefTri { int src, ref_node, rank; };

// Re-emit the caller's BVH below one root child in the reference's visiting order, keeping only what the
// device needs to answer "does the reference test the triangles of this node for this ray": the boxes
// (converted to R exactly as the kernel will see them) and the parent links.  `nested` is cleared when a
// child box is not contained in its parent's box -- then the device checks the whole chain per candidate.
template <typename R>
void collect_reference_nodes(const ptw_group* groups, int n_groups, int n_tris, int g, int parent, int depth, HostScene<R>& out,
                             std::vector<RefTri>& list, bool& nested) {
    if (g < 0 || g >= n_groups) fail("BVH node index %d out of range (%d groups)", g, n_groups);
    if (depth > PTW_BVH_STACK) fail("BVH deeper than %d levels (the reference's traversal stack, tracer.cl:624)", PTW_BVH_STACK);
    const ptw_group& s = groups[g];
    const int me = int(out.node_lo.size());
    out.node_lo.push_back({R(s.bb_min[0]), R(s.bb_min[1]), R(s.bb_min[2]), R(0)});
    out.node_hi.push_back({R(s.bb_max[0]), R(s.bb_max[1]), R(s.bb_max[2]), R(0)});
    out.node_parent.push_back(parent);
    if (parent >= 0) {
        const ptk::V4<R>&plo = out.node_lo[size_t(parent)], &phi = out.node_hi[size_t(parent)], &lo = out.node_lo[size_t(me)], &hi = out.node_hi[size_t(me)];
        if (!(plo.x <= lo.x && plo.y <= lo.y && plo.z <= lo.z && phi.x >= hi.x && phi.y >= hi.y && phi.z >= hi.z)) nested = false;
    }
    if (s.tri_count > 0) {
        if (s.tri_offset < 0 || s.tri_offset + s.tri_count > n_tris) fail("BVH node %d references triangles outside the buffer", g);
        for (int k = 0; k < s.tri_count; ++k) list.push_back(RefTri{s.tri_offset + k, me, int(list.size())});
    }
    for (int k = 0; k < 2; ++k)
        if (s.children[k] > 0) collect_reference_nodes(groups, n_groups, n_tris, s.children[k], me, depth + 1, out, list, nested);
}

// ---- rebuilt BVH ---------------------------------------------------------------------------------
// 8-wide BVH over the triangles of one group object, for the warp-cooperative walk of trace.cuh (8 lanes
// per ray: one child box or one leaf triangle per lane).  Built as a binary tree by the binned surface-
// area heuristic in double (leaves of <= 8 triangles), then collapsed: a wide node adopts grandchildren,
// largest box first, until it has 8 children.  It is an INDEX only -- which triangles may be hit is
// still decided by the reference's own arithmetic on the device -- so the one requirement is that a
// stored child box contains the triangles below it with room for the rounding of the device's slab
// test: boxes are padded.
struct Box3 {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    void add(const double* p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void merge(const Box3& o) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], o.lo[a]); hi[a] = std::max(hi[a], o.hi[a]); } }
    bool empty() const { return lo[0] > hi[0]; }
    double half_area() const {
        if (empty()) return 0.0;
        const double x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return x * y + y * z + z * x;
    }
};
struct BuildPrim { Box3 box; double c[3]; RefTri ref; };
struct BinNode { Box3 box; int left = -1, right = -1, begin = 0, end = 0, height = 0; };   // left < 0: leaf over prims[begin, end)

template <typename R> void padded(const Box3& b, R* lo, R* hi) {
    for (int a = 0; a < 3; ++a) {
        const double pad = 1e-4 * (b.hi[a] - b.lo[a]) + 2e-6 * std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])) + 1e-9;
        lo[a] = R(b.lo[a] - pad); hi[a] = R(b.hi[a] + pad);
        if (double(lo[a]) > b.lo[a] - 0.5 * pad) lo[a] = std::nextafter(lo[a], R(-1e30));   // rounding to R went inward
        if (double(hi[a]) < b.hi[a] + 0.5 * pad) hi[a] = std::nextafter(hi[a], R(1e30));
    }
}
inline float code_as(float, int code) { float f; std::memcpy(&f, &code, 4); return f; }   // bit pattern, never used in arithmetic
inline double code_as(double, int code) { return double(code); }                          // exact

// Threads the BVH build may use: the host's, at most 4 (a 17 k-triangle mesh gains little beyond that -- the top
// splits are serial passes -- and eight ranks of one box build at the same time); PTC_BUILD_THREADS overrides.
// The tree does not depend on it.
inline int build_threads() {
    if (const char* ov = std::getenv("PTC_BUILD_THREADS")) return std::max(1, std::min(64, std::atoi(ov)));
    const unsigned hw = std::thread::hardware_concurrency();
    return int(std::max(1u, std::min(4u, hw)));
}

template <typename R> struct BvhBuilder {
    const ptw_triangle* tris;
    HostScene<R>& out;
    std::vector<BuildPrim>& prims;
    std::vector<BinNode> bin;
    int wide_depth = 0, stack_need = 0;
    int bin_leaf = ptk::kLeafTris;      // triangles per leaf of the BINARY tree (1..kLeafTris); the optimal collapse wants a fine tree
    bool optimal = true;                // collapse by dynamic programming (see plan()); false: greedily by area
    static constexpr int kBins = 32, kSahDepth = 24;
    static constexpr int kForkMin = 1024;       // ranges of at least this many triangles may hand one half to another thread
    struct Bins { Box3 bb[3][kBins]; int bn[3][kBins] = {}; };   // split()'s scratch: clean on entry, cleaned before return
    // plan(): cost[n][i] = cheapest way to present binary subtree n as at most i+1 wide-BVH roots (i = 0..6)
    struct Plan { double cost[7]; signed char split[9]; bool leaf; };   // split[j]: roots given to the left child when n gets j (0 = "as j-1")
    std::vector<Plan> plans;

    // One step of the binary SAH build: bounds of prims[begin, end) into `box`, then either -1 (the range is a leaf)
    // or the split position, with the range reordered around it.  Touches nothing outside the range and `s`.
    int split(int begin, int end, int depth, Bins& s, Box3& box) {
        Box3 cb;
        box = Box3();
        for (int i = begin; i < end; ++i) { box.merge(prims[size_t(i)].box); cb.add(prims[size_t(i)].c); }
        const int count = end - begin;
        if (count <= bin_leaf) return -1;
        int mid = -1;
        const double ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
        if (depth < kSahDepth && (ext[0] > 0 || ext[1] > 0 || ext[2] > 0)) {
            double best = 1e300; int best_axis = -1, best_bin = -1;
            unsigned occ[3] = {0u, 0u, 0u};                        // occupied bins per axis; the bins are clean outside them
            double scale3[3];
            for (int a = 0; a < 3; ++a) scale3[a] = ext[a] > 0 ? double(kBins) / ext[a] : 0.0;
            for (int i = begin; i < end; ++i) {                    // one pass bins the three axes
                const BuildPrim& p = prims[size_t(i)];
                for (int a = 0; a < 3; ++a) {
                    int k = int((p.c[a] - cb.lo[a]) * scale3[a]);
                    k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
                    s.bb[a][k].merge(p.box); s.bn[a][k]++; occ[a] |= 1u << k;
                }
            }
            // The sweep visits occupied bins only.  Between two occupied bins both sides of the split hold the same
            // triangles, so the cost repeats and (strict <) the first of the run -- the occupied bin -- is the one a
            // sweep over all 32 bins would keep: same split, less work on the many small nodes.
            for (int a = 0; a < 3; ++a) {
                if (ext[a] > 0) {
                    double right_area[kBins]; int right_n[kBins];  // filled at occupied bins: everything from that bin up
                    Box3 acc; int n = 0;
                    for (unsigned m = occ[a]; m;) {
                        const int k = 31 - __builtin_clz(m);
                        m &= ~(1u << k);
                        acc.merge(s.bb[a][k]); n += s.bn[a][k]; right_area[k] = acc.half_area(); right_n[k] = n;
                    }
                    acc = Box3(); n = 0;
                    for (unsigned m = occ[a]; m;) {
                        const int k = __builtin_ctz(m);
                        m &= m - 1;
                        if (!m) break;                             // the last occupied bin has nothing to its right
                        const int next = __builtin_ctz(m);
                        acc.merge(s.bb[a][k]); n += s.bn[a][k];
                        // a leaf costs one cooperative step per started group of bin_leaf triangles
                        const double cost = acc.half_area() * std::ceil(n / double(bin_leaf)) + right_area[next] * std::ceil(right_n[next] / double(bin_leaf));
                        if (cost < best) { best = cost; best_axis = a; best_bin = k; }
                    }
                }
                for (unsigned m = occ[a]; m; m &= m - 1) { const int k = __builtin_ctz(m); s.bb[a][k] = Box3(); s.bn[a][k] = 0; }
            }
            if (best_axis >= 0) {
                const double scale = double(kBins) / ext[best_axis];
                const double lo = cb.lo[best_axis];
                auto it = std::partition(prims.begin() + begin, prims.begin() + end, [&](const BuildPrim& p) {
                    int k = int((p.c[best_axis] - lo) * scale);
                    k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
                    return k <= best_bin;
                });
                mid = int(it - prims.begin());
            }
        }
        if (mid <= begin || mid >= end) {                 // no usable SAH split: object median along the widest axis
            int axis = 0;
            if (ext[1] > ext[axis]) axis = 1;
            if (ext[2] > ext[axis]) axis = 2;
            mid = begin + count / 2;
            std::nth_element(prims.begin() + begin, prims.begin() + mid, prims.begin() + end,
                             [axis](const BuildPrim& x, const BuildPrim& y) { return x.c[axis] < y.c[axis] || (x.c[axis] == y.c[axis] && x.ref.rank < y.ref.rank); });
        }
        return mid;
    }

    // binary SAH tree over prims[begin, end), appended to `nodes` in preorder; returns the subtree's root index
    int build(std::vector<BinNode>& nodes, Bins& s, int begin, int end, int depth) {
        const int me = int(nodes.size());
        nodes.emplace_back();
        Box3 box;
        const int mid = split(begin, end, depth, s, box);
        nodes[size_t(me)].box = box; nodes[size_t(me)].begin = begin; nodes[size_t(me)].end = end;
        if (mid < 0) return me;
        const int l = build(nodes, s, begin, mid, depth + 1);
        const int r = build(nodes, s, mid, end, depth + 1);
        nodes[size_t(me)].left = l; nodes[size_t(me)].right = r;
        nodes[size_t(me)].height = 1 + std::max(nodes[size_t(l)].height, nodes[size_t(r)].height);
        return me;
    }

    // The same tree with the two halves of large ranges built on different threads (they touch disjoint ranges of
    // `prims`): every subtree comes back as its own preorder vector and is spliced in behind its parent, so node
    // numbering -- and everything derived from it -- does not depend on the thread count.
    std::vector<BinNode> grow(int begin, int end, int depth, int threads) {
        std::vector<BinNode> nodes;
        auto bins = std::make_unique<Bins>();
        if (threads <= 1 || end - begin < kForkMin) {
            nodes.reserve(size_t(2 * (end - begin) / std::max(1, bin_leaf) + 2));
            build(nodes, *bins, begin, end, depth);
            return nodes;
        }
        Box3 box;
        const int mid = split(begin, end, depth, *bins, box);
        nodes.emplace_back();
        nodes[0].box = box; nodes[0].begin = begin; nodes[0].end = end;
        if (mid < 0) return nodes;
        int left_threads = int(double(threads) * double(mid - begin) / double(end - begin) + 0.5);
        left_threads = std::max(1, std::min(threads - 1, left_threads));
        std::future<std::vector<BinNode>> forked;
        try {
            forked = std::async(std::launch::async, [this, begin, mid, depth, left_threads] { return grow(begin, mid, depth + 1, left_threads); });
        } catch (const std::system_error&) {}            // no thread to be had: this one does both halves
        std::vector<BinNode> right = grow(mid, end, depth + 1, threads - left_threads);
        std::vector<BinNode> left = forked.valid() ? forked.get() : grow(begin, mid, depth + 1, 1);
        auto splice = [&nodes](const std::vector<BinNode>& sub) {
            const int at = int(nodes.size());
            for (BinNode n : sub) {
                if (n.left >= 0) { n.left += at; n.right += at; }
                nodes.push_back(n);
            }
            return at;
        };
        nodes.reserve(1 + left.size() + right.size());
        const int l = splice(left), r = splice(right);
        nodes[0].left = l; nodes[0].right = r;
        nodes[0].height = 1 + std::max(nodes[size_t(l)].height, nodes[size_t(r)].height);
        return nodes;
    }

    int make_leaf(const BinNode& n) {
        const int first = int(out.tri_info.size());
        for (int i = n.begin; i < n.end; ++i) {
            const RefTri& r = prims[size_t(i)].ref;
            const ptw_triangle& t = tris[r.src];
            out.tri_test.push_back({R(t.p1[0]), R(t.p1[1]), R(t.p1[2]), R(t.e1[0])});
            out.tri_test.push_back({R(t.e1[1]), R(t.e1[2]), R(t.e2[0]), R(t.e2[1])});
            out.tri_test.push_back({R(t.e2[2]), R(0), R(0), R(0)});
            out.tri_shade.push_back({R(t.n1[0]), R(t.n1[1]), R(t.n1[2]), R(t.color[0])});
            out.tri_shade.push_back({R(t.n2[0]), R(t.n2[1]), R(t.n2[2]), R(t.color[1])});
            out.tri_shade.push_back({R(t.n3[0]), R(t.n3[1]), R(t.n3[2]), R(t.color[2])});
            out.tri_info.push_back(make_int2(r.rank, r.ref_node));
        }
        if (first >= (1 << 26)) fail("too many triangles");
        return ~((first << 4) | (n.end - n.begin));
    }

    // Optimal collapse (after Ylitie, Karras, Laine 2017, "Efficient incoherent ray traversal on GPUs through compressed
    // wide BVHs", sec. 3.1), with this walk's cost model: one cooperative step per visited wide node (kNodeCost) and per
    // visited leaf of <= kLeafTris triangles (kLeafCost), each weighted by the surface area of its box.
    static constexpr double kNodeCost = 1.0, kLeafCost = 1.25;
    void plan(int n) {
        const BinNode& b = bin[size_t(n)];
        Plan& p = plans[size_t(n)];
        const double area = b.box.half_area();
        const int count = b.end - b.begin;
        const double as_leaf = count <= ptk::kLeafTris ? area * kLeafCost : 1e300;
        std::memset(p.split, 0, sizeof p.split);
        if (b.left < 0) { for (double& c : p.cost) c = as_leaf; p.leaf = true; return; }
        plan(b.left); plan(b.right);
        const Plan &l = plans[size_t(b.left)], &r = plans[size_t(b.right)];
        auto distribute = [&](int j, signed char& k_best) {          // best split of j roots between the two children
            double best = 1e300;
            for (int k = 1; k < j; ++k) {
                const double c = l.cost[std::min(k, 7) - 1] + r.cost[std::min(j - k, 7) - 1];
                if (c < best) { best = c; k_best = (signed char)k; }
            }
            return best;
        };
        signed char k8 = 1;
        const double as_inner = area * kNodeCost + distribute(ptk::kWide, k8);
        p.split[8] = k8;
        p.leaf = as_leaf <= as_inner;
        p.cost[0] = std::min(as_leaf, as_inner);
        for (int i = 2; i <= 7; ++i) {
            signed char k = 1;
            const double d = distribute(i, k);
            if (d < p.cost[i - 2]) { p.cost[i - 1] = d; p.split[i] = k; } else { p.cost[i - 1] = p.cost[i - 2]; p.split[i] = 0; }
        }
    }
    void roots(int n, int j, std::vector<int>& list) const {         // the (at most j) roots plan() chose for subtree n
        const BinNode& b = bin[size_t(n)];
        while (j > 1 && b.left >= 0 && plans[size_t(n)].split[j] == 0) --j;
        if (j <= 1 || b.left < 0) { list.push_back(n); return; }
        const int k = plans[size_t(n)].split[j];
        roots(b.left, k, list); roots(b.right, j - k, list);
    }
    bool is_leaf(int n) const { return bin[size_t(n)].left < 0 || (optimal && plans[size_t(n)].leaf); }

    // Emits the wide node for binary node `b` (an inner node); returns its index.  `pending` = stack entries
    // the device walk may hold when it enters this node (worst case: every sibling on the way was pushed).
    int emit_wide(int b, int depth, int pending) {
        wide_depth = std::max(wide_depth, depth);
        std::vector<int> kids;
        if (optimal) {
            const int k = plans[size_t(b)].split[8];
            roots(bin[size_t(b)].left, k, kids); roots(bin[size_t(b)].right, ptk::kWide - k, kids);
            for (int kid : kids)                                        // (see below; a lopsided tree falls back to the greedy rule)
                if (pending + int(kids.size()) - 1 + (is_leaf(kid) ? 0 : bin[size_t(kid)].height) > ptk::kWideStack) { kids.clear(); break; }
        }
        const bool planned = !kids.empty();
        if (!planned) kids = {bin[size_t(b)].left, bin[size_t(b)].right};
        // Every child beyond the one being walked may sit on the device's traversal stack.  Invariant: pending +
        // height(subtree) <= kWideStack, where height is that of the BINARY subtree -- it holds at the root (the binary
        // depth is capped) and a node only adopts grandchildren while it keeps holding for every child, so even the most
        // lopsided tree fits the device stack: deep down the nodes simply get narrower.
        while (!planned && int(kids.size()) < ptk::kWide) {
            int pick = -1; double area = -1.0;
            for (size_t i = 0; i < kids.size(); ++i) {
                const BinNode& k = bin[size_t(kids[i])];
                if (is_leaf(kids[i]) || !(k.box.half_area() > area)) continue;
                int tallest = std::max(bin[size_t(k.left)].height, bin[size_t(k.right)].height);
                for (size_t q = 0; q < kids.size(); ++q) if (q != i) tallest = std::max(tallest, bin[size_t(kids[q])].height);
                if (pending + int(kids.size()) + tallest > ptk::kWideStack) continue;          // (|kids| + 1 children -> |kids| pushed)
                area = k.box.half_area(); pick = int(i);
            }
            if (pick < 0) break;
            const int k = kids[size_t(pick)];
            kids[size_t(pick)] = bin[size_t(k)].left;
            kids.push_back(bin[size_t(k)].right);
        }
        const int me = int(out.wide.size()) / (2 * ptk::kWide);
        out.wide.resize(out.wide.size() + 2 * ptk::kWide, ptk::V4<R>{R(0), R(0), R(0), R(0)});
        const int need = pending + int(kids.size()) - 1;
        stack_need = std::max(stack_need, need);
        for (int c = 0; c < ptk::kWide; ++c) {
            int code = ptk::kEmptyChild;
            R lo[3] = {R(0), R(0), R(0)}, hi[3] = {R(0), R(0), R(0)};
            if (c < int(kids.size())) {
                const BinNode& k = bin[size_t(kids[size_t(c)])];
                padded<R>(k.box, lo, hi);
                code = is_leaf(kids[size_t(c)]) ? make_leaf(k) : emit_wide(kids[size_t(c)], depth + 1, need);
            }
            const size_t at = (size_t(me) * ptk::kWide + size_t(c)) * 2;
            out.wide[at] = {lo[0], lo[1], lo[2], code_as(R(0), code)};
            out.wide[at + 1] = {hi[0], hi[1], hi[2], R(0)};
        }
        return me;
    }
};

// One group object: reference nodes, then the rebuilt BVH over all its triangles.
template <typename R>
void build_mesh(const ptw_object& s, int obj_index, const ptw_group* groups, int n_groups, const ptw_triangle* tris, int n_tris, HostScene<R>& out,
                ptk::DMesh<R>& m) {
    std::vector<RefTri> list;
    bool nested = true;
    const auto t_begin = Clock::now();
    for (int c = 0; c < s.child_count; ++c) collect_reference_nodes<R>(groups, n_groups, n_tris, s.children[c], -1, 0, out, list, nested);
    m.flags = nested ? 1 : 0;
    std::vector<BuildPrim> prims;
    prims.reserve(list.size());
    for (const RefTri& r : list) {
        const ptw_triangle& t = tris[r.src];
        BuildPrim p;
        p.ref = r;
        bool finite = true;
        double v[3][3];
        for (int a = 0; a < 3; ++a) {
            // the vertices the device test actually uses: p1, p1 + e1, p1 + e2 (tracer.cl:640-675 never reads p2 / p3)
            v[0][a] = t.p1[a]; v[1][a] = t.p1[a] + t.e1[a]; v[2][a] = t.p1[a] + t.e2[a];
            finite = finite && std::isfinite(v[0][a]) && std::isfinite(v[1][a]) && std::isfinite(v[2][a]);
        }
        // A triangle with a NaN / infinite coordinate can never be recorded with EPSILON < t < 1024: every
        // product of the test that involves it is NaN, +-inf or 0 and fails `t > EPSILON` or the u/v range.
        if (!finite) continue;
        for (int k = 0; k < 3; ++k) p.box.add(v[k]);
        for (int a = 0; a < 3; ++a) p.c[a] = 0.5 * (p.box.lo[a] + p.box.hi[a]);
        prims.push_back(p);
    }
    if (prims.empty()) { m.bvh_root = -1; return; }
    BvhBuilder<R> builder{tris, out, prims};
    out.tri_test.reserve(out.tri_test.size() + 3 * prims.size()); out.tri_shade.reserve(out.tri_shade.size() + 3 * prims.size());
    out.tri_info.reserve(out.tri_info.size() + prims.size());
    out.wide.reserve(out.wide.size() + prims.size() * 2);
    if (const char* ov = std::getenv("PTC_BVH_BIN_LEAF")) builder.bin_leaf = std::max(1, std::min(ptk::kLeafTris, std::atoi(ov)));   // tuning overrides
    if (const char* ov = std::getenv("PTC_BVH_OPTIMAL")) builder.optimal = std::atoi(ov) != 0;
    const auto t_prims = Clock::now();
    builder.bin = builder.grow(0, int(prims.size()), 0, build_threads());
    const int root = 0;
    const auto t_built = Clock::now();
    if (builder.optimal) { builder.plans.resize(builder.bin.size()); builder.plan(root); }
    const auto t_planned = Clock::now();
    if (builder.bin[size_t(root)].height > ptk::kWideStack) fail("object %d: triangle tree is %d levels deep (limit %d)", obj_index, builder.bin[size_t(root)].height, ptk::kWideStack);
    const Box3 root_box = builder.bin[size_t(root)].box;
    if (builder.is_leaf(root)) {                          // a single leaf: give it a node to hang from
        const int me = int(out.wide.size()) / (2 * ptk::kWide);
        out.wide.resize(out.wide.size() + 2 * ptk::kWide, ptk::V4<R>{R(0), R(0), R(0), code_as(R(0), ptk::kEmptyChild)});
        R lo[3], hi[3];
        padded<R>(root_box, lo, hi);
        out.wide[size_t(me) * ptk::kWide * 2] = {lo[0], lo[1], lo[2], code_as(R(0), builder.make_leaf(builder.bin[size_t(root)]))};
        out.wide[size_t(me) * ptk::kWide * 2 + 1] = {hi[0], hi[1], hi[2], R(0)};
        for (int c = 1; c < ptk::kWide; ++c) out.wide[(size_t(me) * ptk::kWide + size_t(c)) * 2 + 1] = {R(0), R(0), R(0), R(0)};
        m.bvh_root = me;
    } else {
        m.bvh_root = builder.emit_wide(root, 0, 0);
    }
    if (builder.stack_need > ptk::kWideStack) fail("object %d: rebuilt BVH needs %d traversal stack entries (limit %d)", obj_index, builder.stack_need, ptk::kWideStack);
    out.mesh_depth = std::max(out.mesh_depth, builder.wide_depth);
    out.stack_need = std::max(out.stack_need, builder.stack_need);
    R lo[3], hi[3];
    padded<R>(root_box, lo, hi);
    for (int a = 0; a < 3; ++a) { m.root_lo[a] = lo[a]; m.root_hi[a] = hi[a]; }
    if (std::getenv("PTC_DEBUG_TIMING")) {
        auto ms = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::fprintf(stderr, "[build_mesh] object %d: %zu triangles; reference nodes + boxes %.2f ms, binary SAH %.2f ms, plan %.2f ms, emit %.2f ms\n",
                     obj_index, prims.size(), ms(t_begin, t_prims), ms(t_prims, t_built), ms(t_built, t_planned), ms(t_planned, Clock::now()));
    }
}

template <typename R> void flatten(const ptc_job& job, HostScene<R>& out) {
    const auto* objs = static_cast<const ptw_object*>(job.objects);
    const auto* tris = static_cast<const ptw_triangle*>(job.triangles);
    const auto* groups = static_cast<const ptw_group*>(job.groups);
    std::memset(out.hot, 0, sizeof out.hot);
    for (int i = 0; i < job.n_objects; ++i) {
        const ptw_object& s = objs[i];
        ptk::DObjHot<R>& h = out.hot[i];
        ptk::DObjShade<R> o;
        std::memset(&o, 0, sizeof o);
        for (int k = 0; k < 12; ++k) o.inv[k] = h.inv[k] = R(s.inverse[k]);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) o.invt[r * 3 + c] = R(s.inverse_transpose[r * 4 + c]);
        for (int k = 0; k < 3; ++k) { o.color[k] = R(s.color[k]); o.emission[k] = R(s.emission[k]); }
        o.refractive_index = R(s.refractive_index); o.reflectivity = R(s.reflectivity);
        o.min_y = R(s.min_y); o.max_y = R(s.max_y);
        o.tex_sx = R(s.texture_scale_x); o.tex_sy = R(s.texture_scale_y);
        o.tex_sx_nm = R(s.texture_scale_x_nm); o.tex_sy_nm = R(s.texture_scale_y_nm);
        o.type = h.type = (s.type >= 0 && s.type <= 4) ? int(s.type) : 999;
        o.flags = (s.is_textured ? 1 : 0) | (s.is_textured_nm ? 2 : 0);
        if ((o.type >= 1 && o.type <= 3) || o.flags != 0) o.flags |= 4;      // needs the object-space hit point
        o.tex_index = s.texture_index; o.tex_index_nm = s.texture_index_nm;
        {   // world normal of an un-normal-mapped plane: normalize(inverseTranspose * (0,1,0,0)).xyz
            const double nx = s.inverse_transpose[1], ny = s.inverse_transpose[5], nz = s.inverse_transpose[9];
            const double len = std::sqrt(nx * nx + ny * ny + nz * nz);
            o.plane_n[0] = R(nx / len); o.plane_n[1] = R(ny / len); o.plane_n[2] = R(nz / len);
        }
        h.node_begin = h.node_end = 0;
        if (h.type == 2) { h.aux[0] = R(s.min_y); h.aux[1] = R(s.max_y); }
        ptk::DMesh<R> m;
        std::memset(&m, 0, sizeof m);
        m.bvh_root = -1;
        if (h.type == 4) {
            for (int k = 0; k < 3; ++k) { h.aux[k] = R(s.bb_min[k]); h.aux[3 + k] = R(s.bb_max[k]); }   // object AABB, tracer.cl:609
            if (s.child_count > 0) {
                if (s.child_count > PTW_MAX_ROOT_CHILDREN) fail("object %d: child_count %d > %d", i, s.child_count, PTW_MAX_ROOT_CHILDREN);
                if (!groups || job.n_groups <= 0) fail("object %d is a group but no BVH groups were passed", i);
                h.node_begin = int(out.node_lo.size());
                build_mesh<R>(s, i, groups, job.n_groups, tris, job.n_triangles, out, m);
                h.node_end = int(out.node_lo.size());
            }
        }
        out.mesh.push_back(m);
        out.shade.push_back(o);
    }
    // lights of the next-event estimation (tracer.cl:786-792): every object with emission.x > 0
    out.n_lights = 0;
    std::memset(out.light, 0, sizeof out.light);
    for (int i = 0; i < job.n_objects; ++i) {
        const ptw_object& s = objs[i];
        if (!(s.emission[0] > 0.0)) continue;
        ptk::DLight<R>& L = out.light[out.n_lights++];
        L.ox = R(s.transform[3]); L.oy = R(s.transform[7]); L.oz = R(s.transform[11]);
        L.scale = R(std::fmax(std::fmax(s.transform[0], s.transform[5]), s.transform[10]));
        L.t0 = R(s.transform[0]);
        L.er = R(s.emission[0]); L.eg = R(s.emission[1]); L.eb = R(s.emission[2]);
        L.obj = i;
    }
    // Intersection order.  Planes and spheres whose `inverse` is a similarity (uniform scale, any rotation: the unit
    // sphere is then a world-space sphere of radius 1/s around the transform's translation) take the kernel's unrolled
    // fast slots -- three runs "spheres, planes, spheres" filled greedily in scene order, so fast objects keep their
    // relative order; other analytic objects take the slow loop; groups with triangles the mesh walk.
    out.n_slow = out.n_mesh = 0;
    std::memset(out.slow_kind, 0, sizeof out.slow_kind); std::memset(out.slow_rec, 0, sizeof out.slow_rec); std::memset(out.slow_rec2, 0, sizeof out.slow_rec2);
    for (int k = 0; k < 4; ++k) out.fast_n[k] = 0;
    out.fast_obj[ptk::kFastSlots] = -1;
    for (int k = 0; k < ptk::kFastSlots; ++k) {
        out.fast_obj[k] = -1; out.fast_kind[k] = 0;
        out.fast[k] = out.fast2[k] = ptk::DFast<R>{R(0), R(0), R(0), R(0)};
    }
    const int run_begin[3] = {0, ptk::kFastA, ptk::kFastA + ptk::kFastB}, run_cap[3] = {ptk::kFastA, ptk::kFastB, ptk::kFastC};
    int run = 0;                                                     // run currently being filled (never goes back)
    for (int i = 0; i < job.n_objects; ++i) {
        const ptw_object& s = objs[i];
        const int type = out.hot[i].type;
        if (type == 4) { if (out.mesh[size_t(i)].bvh_root >= 0) out.mesh_obj[out.n_mesh++] = i; continue; }
        if (type < 0 || type > 3) continue;                          // unknown type: never hit (tracer.cl:549-597 has no branch for it)
        ptk::DFast<R> rec{R(0), R(0), R(0), R(0)}, rec2{R(0), R(0), R(0), R(0)};
        int want = -1, kind = 0;                                     // 0: a sphere run, 1: the plane run
        if (type == 0) {
            rec = ptk::DFast<R>{R(s.inverse[4]), R(s.inverse[5]), R(s.inverse[6]), R(s.inverse[7])};
            want = 1;
        } else if (type == 1) {
            const double* m = s.inverse;
            const double s2 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
            auto near = [&](double v, double target) { return std::fabs(v - target) <= 1e-12 * s2; };
            const bool similarity = s2 > 0.0 && std::isfinite(s2) && near(m[4] * m[4] + m[5] * m[5] + m[6] * m[6], s2) &&
                                    near(m[8] * m[8] + m[9] * m[9] + m[10] * m[10], s2) && near(m[0] * m[4] + m[1] * m[5] + m[2] * m[6], 0.0) &&
                                    near(m[0] * m[8] + m[1] * m[9] + m[2] * m[10], 0.0) && near(m[4] * m[8] + m[5] * m[9] + m[6] * m[10], 0.0) &&
                                    m[12] == 0.0 && m[13] == 0.0 && m[14] == 0.0 && m[15] == 1.0;
            if (similarity) {
                // centre = transform * (0,0,0,1) = -A^-1 t, taken from `inverse` itself so the two stay consistent: c = -(A^T t) / s^2
                const double cx = -(m[0] * m[3] + m[4] * m[7] + m[8] * m[11]) / s2, cy = -(m[1] * m[3] + m[5] * m[7] + m[9] * m[11]) / s2,
                             cz = -(m[2] * m[3] + m[6] * m[7] + m[10] * m[11]) / s2;
                rec = ptk::DFast<R>{R(cx), R(cy), R(cz), R(1.0 / s2)};
                want = 0;
            } else if (m[1] == 0.0 && m[2] == 0.0 && m[4] == 0.0 && m[6] == 0.0 && m[8] == 0.0 && m[9] == 0.0 && m[0] != 0.0 && m[5] != 0.0 &&
                       m[10] != 0.0 && std::isfinite(m[0] * m[5] * m[10]) && m[12] == 0.0 && m[13] == 0.0 && m[14] == 0.0 && m[15] == 1.0) {
                // axis-aligned ellipsoid: inverse = diag(sx, sy, sz) p + t = S (p - c) with c = -t / s
                rec = ptk::DFast<R>{R(-m[3] / m[0]), R(-m[7] / m[5]), R(-m[11] / m[10]), R(1)};
                rec2 = ptk::DFast<R>{R(m[0]), R(m[5]), R(m[10]), R(0)};
                want = 0; kind = 1;
            }
        }
        int slot = -1;
        if (want == 1 && run <= 1 && out.fast_n[1] < run_cap[1]) { run = 1; slot = run_begin[1] + out.fast_n[1]++; }
        else if (want == 0) {
            if (run == 0 && out.fast_n[0] < run_cap[0]) slot = run_begin[0] + out.fast_n[0]++;
            else if (out.fast_n[2] < run_cap[2]) { run = 2; slot = run_begin[2] + out.fast_n[2]++; }
        }
        if (slot >= 0) { out.fast[slot] = rec; out.fast2[slot] = rec2; out.fast_kind[slot] = kind; out.fast_obj[slot] = i; }
        else {
            // overflow (a full run, or an order that does not fit "spheres, planes, spheres"): fast-class objects keep
            // their record and the slots' arithmetic in the slow loop, so coincident objects still tie exactly
            out.slow_kind[out.n_slow] = want == 1 ? 0 : want == 0 ? 1 + kind : 3;
            out.slow_rec[out.n_slow] = rec; out.slow_rec2[out.n_slow] = rec2;
            out.slow_obj[out.n_slow++] = i;
        }
    }
    const auto* cam = static_cast<const ptw_camera*>(job.camera);
    out.cam.pixel_size = R(cam->pixel_size); out.cam.half_width = R(cam->half_width); out.cam.half_height = R(cam->half_height);
    out.cam.aperture = R(cam->aperture); out.cam.focal_length = R(cam->focal_length);
    for (int k = 0; k < 12; ++k) out.cam.inv[k] = R(cam->inverse[k]);
    out.cam.width = cam->width; out.cam.height = cam->height;
    if (cam->aperture != 0.0) {
        // tracer.cl:221-248 sunflower(amountPoints = samples, alpha = 2, pointNumber = n, randomize = false):
        // a function of the sample index only, evaluated here in double once instead of per path.
        const double PI = double(3.14159265359f);
        const double amount = double(job.samples);
        const double b = std::round(2.0 * std::sqrt(amount));
        const double phi = (std::sqrt(5.0) + 1.0) / 2.0;
        out.lens.resize(size_t(job.samples) * 2);
        for (int n = 0; n < job.samples; ++n) {
            const double idx = double(n);
            double r = 1.0;
            if (idx <= (amount - b)) r = std::sqrt(idx - 0.5) / std::sqrt(amount - (b + 1.0) / 2.0);   // NaN at n == 0: kept
            const double theta = 2.0 * PI * idx / (phi * phi);
            out.lens[size_t(n) * 2] = R(r * std::cos(theta));
            out.lens[size_t(n) * 2 + 1] = R(r * std::sin(theta));
        }
    }
}

struct DeviceState {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<int> rows;          // frame rows owned by this device, increasing
    std::vector<void*> allocs;      // everything allocated on this device
    cudaMemPool_t pool = nullptr;   // stream-ordered pool (single-GPU contexts), else plain cudaMalloc
    void* shade = nullptr; void* lens = nullptr;
    void* node_lo = nullptr; void* node_hi = nullptr; void* node_parent = nullptr;
    void* wide = nullptr;
    void* tri_test = nullptr; void* tri_shade = nullptr; void* tri_info = nullptr;
    void* tex[3] = {nullptr, nullptr, nullptr};
    double* seeds = nullptr;        // rows*W: the seeds of the rows this device owns
    int* row_map = nullptr;         // local row -> frame row
    int* tile_order = nullptr;      // launch position -> tile (heavy tiles first), or NULL
    unsigned* tile_cost = nullptr;  // per tile: slowest warp's clocks of the last launch
    int n_tiles = 0;
    int order_measured = 0;         // tile_order comes from: 0 the geometric estimate, 1 the cost probe, 2 the clocks of a full render
    int* out_row = nullptr;         // local row -> row of the buffer the kernel stores into (packed context rows / frame rows)
    double* out = nullptr;          // rows*W*4: this device's packed rows (unused while the kernel stores into a gather / frame buffer)
    double* acc = nullptr;          // rows*W*4 running sums of a progressive render
    uchar4* rgba8 = nullptr;        // rows*W tone-mapped bytes (ptc_read_rgba8)
    float4* f32 = nullptr;          // rows*W float RGBA (ptc_read_f32)
    int slices = 1, cluster = 1;    // sample slices per pixel; CTAs per cluster (slices = slices_per_block * cluster)
    int sm_count = 0;
    float last_ms = 0.f;
};

}  // namespace

// A whole-frame buffer on one device that kernels of several contexts -- other devices of this process, or other
// processes through a CUDA IPC mapping -- store their pixels into directly (the gather fused into the trace kernel).
struct ptc_frame {
    int device = 0, width = 0, height = 0, format = PTC_FRAME_F64;
    void* data = nullptr;
    bool owner = false;         // false: `data` is an IPC mapping of another process's allocation
    size_t bytes() const { return size_t(width) * height * (format == PTC_FRAME_F32 ? 16 : 32); }
};

struct ptc_context {
    int width = 0, height = 0, samples = 0, precision = PTC_FP32, rng_mode = PTC_RNG_PARITY;
    int n_objects = 0;
    int shard_index = 0, shard_count = 1, rows_per_tile = 4;
    std::vector<int> rows;                       // rows owned by this context (all its devices), increasing
    std::vector<DeviceState> dev;
    HostScene<float> scene32;
    HostScene<double> scene64;
    int tex_w[3] = {0, 0, 0}, tex_h[3] = {0, 0, 0}, tex_layers[3] = {0, 0, 0};
    double* gather = nullptr;                    // on dev[0]: packed rows of the whole context (n_devices > 1), written by every device's kernel
    bool peer_ok = false;                        // every device can store into dev[0]'s memory
    ptc_frame* frame = nullptr;                  // attached frame (ptc_set_frame): kernels store there, by frame row
    int nee = 0, caps = 0;                       // ptc_job.features
    ptc_stats stats{};
};

namespace {

// Device memory for single-GPU contexts comes from a per-device stream-ordered pool that keeps its
// blocks between calls: a render no longer pays cudaMalloc/cudaFree -- the latter synchronises the
// device and was measured to stall for up to 1.3 s inside a process that also hosts another CUDA
// allocator.  What the pool retains is bounded (kPoolKeepBytes: a 4K fp64 frame with its seeds and a
// mesh scene fit); anything above goes back to the driver when the context closes.  Multi-GPU contexts
// use plain cudaMalloc so peer stores need no per-pool access grants.  The pool table is the only global
// state; it is mutex-protected and holds driver handles only.  ptc_trim() returns the rest.
constexpr uint64_t kPoolKeepBytes = 768ull << 20;
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[64] = {};

cudaMemPool_t pool_for(int device) {
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[device]) {
        cudaMemPoolProps props;
        std::memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        uint64_t keep = kPoolKeepBytes;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        g_pools[device] = pool;
    }
    return g_pools[device];
}

void* dmalloc(DeviceState& d, size_t bytes) {
    void* p = nullptr;
    if (d.pool) CUDA_OK(cudaMallocFromPoolAsync(&p, bytes ? bytes : 16, d.pool, d.stream));
    else CUDA_OK(cudaMalloc(&p, bytes ? bytes : 16));
    d.allocs.push_back(p);
    return p;
}
template <typename T> void* upload(DeviceState& d, const std::vector<T>& v, int64_t& h2d) {
    void* p = dmalloc(d, v.size() * sizeof(T));
    if (!v.empty()) {
        CUDA_OK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, d.stream));
        h2d += int64_t(v.size() * sizeof(T));
    }
    return p;
}

template <typename R> void upload_scene(DeviceState& d, const HostScene<R>& s, int64_t& h2d) {
    d.shade = upload(d, s.shade, h2d);
    d.lens = s.lens.empty() ? nullptr : upload(d, s.lens, h2d);
    d.node_lo = upload(d, s.node_lo, h2d);
    d.node_hi = upload(d, s.node_hi, h2d);
    d.node_parent = upload(d, s.node_parent, h2d);
    d.wide = upload(d, s.wide, h2d);
    d.tri_test = upload(d, s.tri_test, h2d);
    d.tri_shade = upload(d, s.tri_shade, h2d);
    d.tri_info = upload(d, s.tri_info, h2d);
}

// Seeds of the rows a device owns, from the caller's full-frame array (one double per pixel, row-major): the rows
// come in runs (scanline tiles) at a regular pitch, so one strided copy moves them; ragged layouts fall back to a
// copy per run.  Only the owned rows cross PCIe (round 1 sent every device the whole frame).
void upload_seeds(DeviceState& d, const double* seeds, int width, int64_t& h2d) {
    const size_t row_bytes = size_t(width) * sizeof(double);
    struct Run { int first, count; };
    std::vector<Run> runs;
    for (int r : d.rows) {
        if (!runs.empty() && runs.back().first + runs.back().count == r) runs.back().count++;
        else runs.push_back(Run{r, 1});
    }
    size_t regular = 0;                          // leading runs of equal length at a constant pitch
    if (runs.size() >= 2) {
        const int len = runs[0].count, pitch = runs[1].first - runs[0].first;
        regular = 1;
        while (regular < runs.size() && runs[regular].count == len && runs[regular].first - runs[regular - 1].first == pitch) ++regular;
        if (regular >= 2)
            CUDA_OK(cudaMemcpy2DAsync(d.seeds, size_t(len) * row_bytes, seeds + size_t(runs[0].first) * width, size_t(pitch) * row_bytes,
                                      size_t(len) * row_bytes, regular, cudaMemcpyHostToDevice, d.stream));
        else regular = 0;
    }
    size_t local = 0;
    for (size_t k = 0; k < runs.size(); ++k) {
        if (k >= regular)
            CUDA_OK(cudaMemcpyAsync(d.seeds + local * width, seeds + size_t(runs[k].first) * width, size_t(runs[k].count) * row_bytes,
                                    cudaMemcpyHostToDevice, d.stream));
        local += size_t(runs[k].count);
    }
    h2d += int64_t(d.rows.size() * row_bytes);
}

template <typename R> ptk::Params<R> make_params(const ptc_context& c, const DeviceState& d, const HostScene<R>& s) {
    ptk::Params<R> P;
    std::memset(&P, 0, sizeof P);
    std::memcpy(P.hot, s.hot, sizeof P.hot);
    std::memcpy(P.fast, s.fast, sizeof P.fast);
    std::memcpy(P.fast_n, s.fast_n, sizeof P.fast_n);
    std::memcpy(P.fast2, s.fast2, sizeof P.fast2);
    std::memcpy(P.fast_kind, s.fast_kind, sizeof P.fast_kind);
    std::memcpy(P.fast_obj, s.fast_obj, sizeof P.fast_obj);
    std::memcpy(P.slow_obj, s.slow_obj, sizeof P.slow_obj);
    std::memcpy(P.slow_kind, s.slow_kind, sizeof P.slow_kind);
    std::memcpy(P.slow_rec, s.slow_rec, sizeof P.slow_rec);
    std::memcpy(P.slow_rec2, s.slow_rec2, sizeof P.slow_rec2);
    std::memcpy(P.mesh_obj, s.mesh_obj, sizeof P.mesh_obj);
    P.n_slow = s.n_slow; P.n_mesh = s.n_mesh;
    P.stack_entries = (s.stack_need + 1) | 1;          // odd: the four groups of a warp push to different banks
    P.lane_walk_min = ptk::kLaneWalkMin;
    P.defer_max = ptk::kDeferMax;
    P.defer_below = 4;
    if (const char* ov = std::getenv("PTC_DEFER_BELOW")) P.defer_below = std::atoi(ov);
    if (const char* ov = std::getenv("PTC_DEFER_MAX")) P.defer_max = std::atoi(ov);                // tuning override
    if (const char* ov = std::getenv("PTC_LANE_WALK_MIN")) P.lane_walk_min = std::atoi(ov);      // tuning override
    P.shade = static_cast<const ptk::DObjShade<R>*>(d.shade);
    P.lens = static_cast<const R*>(d.lens);
    P.n_objects = c.n_objects;
    P.node_lo = static_cast<const ptk::V4<R>*>(d.node_lo);
    P.node_hi = static_cast<const ptk::V4<R>*>(d.node_hi);
    P.node_parent = static_cast<const int*>(d.node_parent);
    for (int k = 0; k < c.n_objects; ++k) P.mesh[k] = s.mesh[size_t(k)];
    P.wide = static_cast<const ptk::V4<R>*>(d.wide);
    P.tri_test = static_cast<const ptk::V4<R>*>(d.tri_test);
    P.tri_shade = static_cast<const ptk::V4<R>*>(d.tri_shade);
    P.tri_info = static_cast<const int2*>(d.tri_info);
    P.cam = s.cam;
    for (int k = 0; k < 3; ++k) P.tex[k] = ptk::DTex{static_cast<const uchar4*>(d.tex[k]), c.tex_w[k], c.tex_h[k], c.tex_layers[k]};
    P.seeds = d.seeds;
    P.row_map = d.row_map;
    P.tile_order = d.tile_order;
    P.tile_cost = d.order_measured < 2 ? d.tile_cost : nullptr;
    P.n_tiles = d.n_tiles;
    P.sample_begin = 0; P.sample_end = c.samples;
    P.rows = int(d.rows.size());
    P.samples = c.samples;
    P.slices = d.slices;
    P.slices_per_block = std::min(d.slices, ptk::kBlockWarps);
    P.pi = R(double(3.14159265359f));
    P.eps = R(0.0001);
    P.nee = c.nee; P.caps = c.caps;
    std::memcpy(P.light, s.light, sizeof P.light);
    P.n_lights = s.n_lights;
    // where the pixels go: an attached frame (by frame row), the context's gather buffer on dev[0] (by position among
    // the context's rows), or this device's own packed rows
    if (c.frame) { P.out = c.frame->data; P.out_row = d.row_map; P.out_f32 = c.frame->format == PTC_FRAME_F32; }
    else if (c.gather && c.peer_ok) { P.out = c.gather; P.out_row = d.out_row; }
    else { P.out = d.out; P.out_row = nullptr; }
    return P;
}

template <typename K, typename R>
void launch_kernel(K kernel, dim3 grid, dim3 block, size_t smem, int cluster, cudaStream_t stream, const ptk::Params<R>& P) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(cluster); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = cluster > 1 ? 1 : 0;
    if (std::getenv("PTC_DEBUG_OCCUPANCY")) {
        int clusters = 0, blocks = 0;
        cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, int(block.x), smem);
        std::fprintf(stderr, "[launch] grid %u cluster %d smem %zu: max active clusters %d (= %d blocks), blocks/SM by occupancy %d\n", grid.x, cluster, smem,
                     clusters, clusters * cluster, blocks);
        cudaGetLastError();
    }
    CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, P));
}

// Renders samples [begin, end) of every pixel this device owns.  accumulate == false: a full, self-contained
// render (begin = 0, end = samples).  accumulate == true: the sums are added to the per-pixel accumulator d.acc
// (progressive rendering) and the stored pixel is acc / samples.  One launch either way: the slices of a pixel are
// reduced inside the kernel (shared memory + the cluster's distributed shared memory).
template <typename R> void launch(ptc_context& c, DeviceState& d, const HostScene<R>& s, int begin, int end, bool accumulate) {
    const int rows = int(d.rows.size());
    if (rows == 0) return;
    const size_t pixels = size_t(rows) * size_t(c.width);
    if (accumulate && !d.acc) {
        d.acc = static_cast<double*>(dmalloc(d, pixels * 4 * sizeof(double)));
        CUDA_OK(cudaMemsetAsync(d.acc, 0, pixels * 4 * sizeof(double), d.stream));
    }
    ptk::Params<R> P = make_params<R>(c, d, s);
    P.sample_begin = begin; P.sample_end = end;
    P.acc = accumulate ? reinterpret_cast<double4*>(d.acc) : nullptr;
    const int tiles_x = (c.width + ptk::kTileW - 1) / ptk::kTileW;
    const int tiles_y = (rows + ptk::kTileH - 1) / ptk::kTileH;
    const long long tiles = (long long)tiles_x * tiles_y;
    const int tiles_per_block = ptk::kBlockWarps / P.slices_per_block;
    const bool meshes = s.n_mesh > 0;
    const bool fast = c.rng_mode == PTC_RNG_FAST;
    dim3 grid((unsigned)((tiles + tiles_per_block - 1) / tiles_per_block * d.cluster), 1, 1);
    dim3 block(ptk::kBlockThreads, 1, 1);
    const size_t smem = meshes ? size_t(ptk::kBlockThreads / ptk::kWide) * size_t(P.stack_entries) * sizeof(int2) : 0;   // traversal stacks
    if (c.nee) {                  // the next-event-estimation kernels (parity RNG stream only, checked in open_impl)
        if (meshes) launch_kernel(ptk::trace_kernel<R, ptk::RNG_PARITY, true, true>, grid, block, smem, d.cluster, d.stream, P);
        else launch_kernel(ptk::trace_kernel<R, ptk::RNG_PARITY, false, true>, grid, block, smem, d.cluster, d.stream, P);
    } else if (meshes) {
        if (fast) launch_kernel(ptk::trace_kernel<R, ptk::RNG_FAST, true>, grid, block, smem, d.cluster, d.stream, P);
        else launch_kernel(ptk::trace_kernel<R, ptk::RNG_PARITY, true>, grid, block, smem, d.cluster, d.stream, P);
    } else {
        if (fast) launch_kernel(ptk::trace_kernel<R, ptk::RNG_FAST, false>, grid, block, smem, d.cluster, d.stream, P);
        else launch_kernel(ptk::trace_kernel<R, ptk::RNG_PARITY, false>, grid, block, smem, d.cluster, d.stream, P);
    }
    CUDA_OK(cudaGetLastError());
    c.stats.kernel_launches++;
}

void validate(const ptc_job& j) {
    if (j.abi_version != PTC_ABI_VERSION) fail("ptc_job.abi_version %d != %d", j.abi_version, PTC_ABI_VERSION);
    if (!j.objects || j.n_objects < 1) fail("scene has no objects");
    if (j.n_objects > PTW_MAX_OBJECTS) fail("%d objects: the kernel supports at most %d (tracer.cl:846)", j.n_objects, PTW_MAX_OBJECTS);
    if (!j.camera) fail("camera is NULL");
    if (j.n_triangles < 0 || j.n_groups < 0) fail("negative triangle/group count");
    if (j.n_triangles > 0 && !j.triangles) fail("n_triangles > 0 but triangles is NULL");
    if (j.n_groups > 0 && !j.groups) fail("n_groups > 0 but groups is NULL");
    const auto* cam = static_cast<const ptw_camera*>(j.camera);
    if (cam->width <= 0 || cam->height <= 0) fail("camera %dx%d: width and height must be positive", cam->width, cam->height);
    if ((long long)cam->width * cam->height > (1ll << 30)) fail("frame too large");
    if (j.samples < 1) fail("samples must be >= 1");
    if (!j.seeds) fail("seeds is NULL (one double per pixel)");
    if (j.precision != PTC_FP32 && j.precision != PTC_FP64) fail("unknown precision %d", j.precision);
    if (j.rng_mode != PTC_RNG_PARITY && j.rng_mode != PTC_RNG_FAST) fail("unknown rng_mode %d", j.rng_mode);
    for (int k = 0; k < 3; ++k)
        if (j.tex[k] && (j.tex_w[k] <= 0 || j.tex_h[k] <= 0 || j.tex_layers[k] <= 0)) fail("texture class %d has a non-positive size", k);
    if (j.shard_count > 1 && (j.shard_index < 0 || j.shard_index >= j.shard_count)) fail("shard_index %d outside [0,%d)", j.shard_index, j.shard_count);
    if (j.n_devices < 0 || (j.n_devices > 0 && !j.devices)) fail("bad device list");
    if (j.features & ~(PTC_FEATURE_NEE | PTC_FEATURE_CYLINDER_CAPS)) fail("unknown bits in ptc_job.features: 0x%x", j.features);
    for (int k = 0; k < 7; ++k) if (j.reserved[k] != 0) fail("ptc_job.reserved must be zero");
}

void destroy(ptc_context* c) {
    if (!c) return;
    for (DeviceState& d : c->dev) {
        cudaSetDevice(d.device);
        if (d.pool && d.stream) {
            for (void* p : d.allocs) cudaFreeAsync(p, d.stream);     // back to the pool, no device-wide sync
            cudaStreamSynchronize(d.stream);
        } else {
            for (void* p : d.allocs) cudaFree(p);
        }
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete c;
}

// Sample slices per pixel for a device that owns `px` pixels.  Small frames cannot fill 148 SMs with one thread
// per pixel, and blocks that live for all the samples of their tile leave a tail, so each pixel's samples are split
// into interleaved slices; all slices of a pixel sit in one thread-block cluster (kBlockWarps per block), which
// reduces them, so the count is a power of two <= kBlockWarps * kMaxCluster = 32.  Measured on one B200 (round 2):
//   whole 1280x960 frame, slices 1 / 2 / 4 / 8 / 16 / 32: reference scene 12.72 / 12.89 / 12.96 / 12.99 / 12.42 / 12.36,
//                                                          teapot 2.30 / 3.12 / 3.54 / 3.67 / 3.50 / 3.44 Gpaths/s;
//   one 1/8 shard of it (the 8-GPU case) with the tiles launched longest-first, 8 / 16 / 32 slices: teapot 76.5 / 77.4 /
//   78.2 ms, gopher 106.4 / 110.7 / 111.7 ms (in frame order the gopher shard needed 32 slices: 142.6 / 120.6 / 111.8).
// Eight slices = clusters of two blocks: cudaOccupancyMaxActiveClusters gives 1184 resident blocks for clusters of one or
// two, 1136 for four or eight (the GPCs' SM counts are not multiples of the cluster size), and smaller clusters keep the
// tail short only together with the launch order below.  More slices only when the frame is too small to fill the machine.
int plan_slices(long long px, int sm_count, bool meshes, int samples) {
    (void)meshes;
    const long long resident = (long long)sm_count * 1024;                        // threads at 8 blocks x 128 per SM
    long long sl = px * 8 >= 2 * resident ? 8 : (px ? (4 * resident + px - 1) / px : 1);
    if (const char* ov = std::getenv("PTC_SLICES")) sl = std::atoll(ov);    // tuning override
    const long long cap = std::min<long long>(samples, ptk::kBlockWarps * ptk::kMaxCluster);
    if (sl > cap) sl = cap;
    int pow2 = 1;
    while (2ll * pow2 <= sl) pow2 *= 2;
    return pow2;
}

// Launch order of a device's 8x4 pixel tiles.  Blocks start in launch order and live for all the samples of their tile,
// and a tile that looks at a mesh costs several times a tile that looks at a wall, so the last blocks to finish decide
// the kernel's tail -- on a 1/8-frame shard the eight shard kernels together took 10 % longer than the whole frame in one
// launch.  Longest-first scheduling: tiles whose pixels overlap the screen-space bounding rectangle of a mesh object
// (its object-space AABB through `transform` and the camera) are launched first, the cheap ones fill the end.
// Returns an empty vector when there is nothing to reorder.  Order only: every tile is rendered exactly once.
std::vector<int> plan_tile_order(const ptc_job& job, const std::vector<int>& rows) {
    const auto* objs = static_cast<const ptw_object*>(job.objects);
    const auto* cam = static_cast<const ptw_camera*>(job.camera);
    const int W = cam->width;
    // world -> view: inverse of the affine camera matrix (rows 0..2 of cam->inverse are [A | t])
    const double* m = cam->inverse;
    const double a[3][3] = {{m[0], m[1], m[2]}, {m[4], m[5], m[6]}, {m[8], m[9], m[10]}}, t[3] = {m[3], m[7], m[11]};
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (!(std::fabs(det) > 1e-300) || !(cam->pixel_size > 0.0)) return {};
    double inv[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c2 = 0; c2 < 3; ++c2) {
            const int r1 = (c2 + 1) % 3, r2 = (c2 + 2) % 3, c1 = (r + 1) % 3, c3 = (r + 2) % 3;
            inv[r][c2] = (a[r1][c1] * a[r2][c3] - a[r1][c3] * a[r2][c1]) / det;
        }
    struct Rect { double x0, y0, x1, y1; };
    std::vector<Rect> rects;
    for (int i = 0; i < job.n_objects; ++i) {
        const ptw_object& o = objs[i];
        if (o.type != 4 || o.child_count <= 0) continue;
        Rect rc{1e300, 1e300, -1e300, -1e300};
        bool whole = false;
        for (int k = 0; k < 8 && !whole; ++k) {
            const double p[3] = {k & 1 ? o.bb_max[0] : o.bb_min[0], k & 2 ? o.bb_max[1] : o.bb_min[1], k & 4 ? o.bb_max[2] : o.bb_min[2]};
            double w3[3], v[3];
            for (int r = 0; r < 3; ++r) w3[r] = o.transform[4 * r] * p[0] + o.transform[4 * r + 1] * p[1] + o.transform[4 * r + 2] * p[2] + o.transform[4 * r + 3] - t[r];
            for (int r = 0; r < 3; ++r) v[r] = inv[r][0] * w3[0] + inv[r][1] * w3[1] + inv[r][2] * w3[2];
            if (!(v[2] < -1e-6) || !std::isfinite(v[0] + v[1] + v[2])) { whole = true; break; }      // behind the camera: give up on a rectangle
            const double x = (cam->half_width - v[0] / -v[2]) / cam->pixel_size, y = (cam->half_height - v[1] / -v[2]) / cam->pixel_size;
            rc.x0 = std::min(rc.x0, x); rc.x1 = std::max(rc.x1, x); rc.y0 = std::min(rc.y0, y); rc.y1 = std::max(rc.y1, y);
        }
        if (whole) return {};
        const double margin = 8.0 + (cam->aperture != 0.0 ? 0.05 * W : 0.0);
        rects.push_back(Rect{rc.x0 - margin, rc.y0 - margin, rc.x1 + margin, rc.y1 + margin});
    }
    if (rects.empty()) return {};
    const int tiles_x = (W + ptk::kTileW - 1) / ptk::kTileW, tiles_y = (int(rows.size()) + ptk::kTileH - 1) / ptk::kTileH;
    std::vector<int> heavy, light;
    for (int ty = 0; ty < tiles_y; ++ty) {
        int r0 = 1 << 30, r1 = -1;
        for (int k = ty * ptk::kTileH; k < (ty + 1) * ptk::kTileH && k < int(rows.size()); ++k) { r0 = std::min(r0, rows[size_t(k)]); r1 = std::max(r1, rows[size_t(k)]); }
        for (int tx = 0; tx < tiles_x; ++tx) {
            const double x0 = tx * ptk::kTileW, x1 = x0 + ptk::kTileW;
            bool hit = false;
            for (const Rect& rc : rects) hit = hit || (x1 >= rc.x0 && x0 <= rc.x1 && double(r1 + 1) >= rc.y0 && double(r0) <= rc.y1);
            (hit ? heavy : light).push_back(ty * tiles_x + tx);
        }
    }
    if (heavy.empty() || light.empty()) return {};
    heavy.insert(heavy.end(), light.begin(), light.end());
    return heavy;
}

ptc_context* open_impl(const ptc_job& job) {
    validate(job);
    auto t0 = Clock::now();
    int n_cuda = 0;
    cudaError_t e = cudaGetDeviceCount(&n_cuda);
    if (e != cudaSuccess || n_cuda <= 0)
        fail("no usable CUDA device (%s); libptcuda has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");

    std::unique_ptr<ptc_context, void (*)(ptc_context*)> guard(new ptc_context, destroy);
    ptc_context& c = *guard;
    const auto* cam = static_cast<const ptw_camera*>(job.camera);
    c.width = cam->width; c.height = cam->height; c.samples = job.samples;
    c.precision = job.precision; c.rng_mode = job.rng_mode; c.n_objects = job.n_objects;
    c.nee = (job.features & PTC_FEATURE_NEE) ? 1 : 0;
    if (c.nee && job.rng_mode != PTC_RNG_PARITY) fail("PTC_FEATURE_NEE is available with PTC_RNG_PARITY only");
    c.caps = (job.features & PTC_FEATURE_CYLINDER_CAPS) ? 1 : 0;
    c.shard_count = job.shard_count > 1 ? job.shard_count : 1;
    c.shard_index = job.shard_count > 1 ? job.shard_index : 0;
    c.rows_per_tile = job.rows_per_tile > 0 ? job.rows_per_tile : 4;
    for (int k = 0; k < 3; ++k)
        if (job.tex[k]) { c.tex_w[k] = job.tex_w[k]; c.tex_h[k] = job.tex_h[k]; c.tex_layers[k] = job.tex_layers[k]; }

    std::vector<int> devices;
    if (job.n_devices > 0) devices.assign(job.devices, job.devices + job.n_devices);
    else devices.push_back(0);
    for (size_t i = 0; i < devices.size(); ++i) {
        // the reference maps a negative --device-index to 0 and aborts on an index past the end (ocltracer.go:135-141)
        if (devices[i] < 0) devices[i] = 0;
        if (devices[i] > n_cuda - 1) fail("device index %d out of bounds: highest device index: %d", devices[i], n_cuda - 1);
        for (size_t k = 0; k < i; ++k) if (devices[k] == devices[i]) fail("device %d listed twice", devices[i]);
    }
    const int nd = int(devices.size());

    // Row ownership: scanline tile k (rows_per_tile rows) -> process shard k % shard_count; within
    // the shard its tiles are dealt round-robin to the local devices.
    c.dev.resize(size_t(nd));
    const int n_tiles = (c.height + c.rows_per_tile - 1) / c.rows_per_tile;
    int local_tile = 0;
    std::vector<std::vector<int>> out_rows{size_t(nd)};          // per device: position of each of its rows among the context's rows
    for (int k = 0; k < n_tiles; ++k) {
        if (k % c.shard_count != c.shard_index) continue;
        DeviceState& d = c.dev[size_t(local_tile % nd)];
        for (int r = k * c.rows_per_tile; r < (k + 1) * c.rows_per_tile && r < c.height; ++r) {
            out_rows[size_t(local_tile % nd)].push_back(int(c.rows.size()));
            d.rows.push_back(r); c.rows.push_back(r);
        }
        ++local_tile;
    }

    // Per device, first what does not depend on the flattened scene -- stream, pool, textures, seeds, row map -- so that
    // (from pinned host memory) those copies run while the host builds the scene tables and the mesh index below.
    int64_t h2d = 0;
    for (int i = 0; i < nd; ++i) {
        DeviceState& d = c.dev[size_t(i)];
        d.device = devices[size_t(i)];
        CUDA_OK(cudaSetDevice(d.device));
        int cc_major = 0, cc_minor = 0;
        CUDA_OK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, d.device));
        CUDA_OK(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, d.device));
        if (cc_major < 10) fail("device %d is sm_%d%d; libptcuda is built for sm_100a only", d.device, cc_major, cc_minor);
        CUDA_OK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.device));
        CUDA_OK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        if (nd == 1) d.pool = pool_for(d.device);
        CUDA_OK(cudaEventCreate(&d.ev0));
        CUDA_OK(cudaEventCreate(&d.ev1));
        for (int k = 0; k < 3; ++k) {
            if (!job.tex[k]) continue;
            size_t bytes = size_t(c.tex_w[k]) * c.tex_h[k] * c.tex_layers[k] * 4;
            d.tex[k] = dmalloc(d, bytes);
            CUDA_OK(cudaMemcpyAsync(d.tex[k], job.tex[k], bytes, cudaMemcpyHostToDevice, d.stream));
            h2d += int64_t(bytes);
        }
        const size_t px = d.rows.size() * size_t(c.width);
        d.seeds = static_cast<double*>(dmalloc(d, px * sizeof(double)));
        upload_seeds(d, job.seeds, c.width, h2d);
        d.row_map = static_cast<int*>(upload(d, d.rows, h2d));
    }

    if (c.precision == PTC_FP64) flatten<double>(job, c.scene64);
    else flatten<float>(job, c.scene32);
    const bool meshes = (c.precision == PTC_FP64 ? c.scene64.n_mesh : c.scene32.n_mesh) > 0;

    for (int i = 0; i < nd; ++i) {
        DeviceState& d = c.dev[size_t(i)];
        CUDA_OK(cudaSetDevice(d.device));
        if (c.precision == PTC_FP64) upload_scene<double>(d, c.scene64, h2d);
        else upload_scene<float>(d, c.scene32, h2d);
        const size_t px = d.rows.size() * size_t(c.width);
        d.n_tiles = ((c.width + ptk::kTileW - 1) / ptk::kTileW) * ((int(d.rows.size()) + ptk::kTileH - 1) / ptk::kTileH);
        if (meshes && !std::getenv("PTC_NO_TILE_ORDER") && d.n_tiles > 0) {
            // Launch order (scenes with meshes, whose tiles differ several-fold in cost; analytic scenes lose 1-4 % when their
            // costlier tiles are bunched up and have no tail to win back): the geometric estimate first, measured clocks after.
            std::vector<int> order = plan_tile_order(job, d.rows);
            if (order.empty()) { order.resize(size_t(d.n_tiles)); for (int k = 0; k < d.n_tiles; ++k) order[size_t(k)] = k; }
            d.tile_order = static_cast<int*>(upload(d, order, h2d));
            if (!std::getenv("PTC_ORDER_GEOMETRIC_ONLY")) {
                d.tile_cost = static_cast<unsigned*>(dmalloc(d, size_t(d.n_tiles) * sizeof(unsigned)));
                CUDA_OK(cudaMemsetAsync(d.tile_cost, 0, size_t(d.n_tiles) * sizeof(unsigned), d.stream));
            }
        }
        d.out_row = nd > 1 ? static_cast<int*>(upload(d, out_rows[size_t(i)], h2d)) : nullptr;
        d.out = static_cast<double*>(dmalloc(d, px * 4 * sizeof(double)));
        d.slices = plan_slices((long long)px, d.sm_count, meshes, c.samples);
        d.cluster = std::max(1, d.slices / ptk::kBlockWarps);
    }
    if (nd > 1) {
        // Every device's kernel stores its pixels straight into one buffer on dev[0] (NVLink peer stores from the
        // kernel epilogue), so the read is a single D2H.  That needs peer access FROM each device TO dev[0].
        DeviceState& d0 = c.dev[0];
        c.peer_ok = true;
        for (int i = 1; i < nd; ++i) {
            int can = 0;
            CUDA_OK(cudaDeviceCanAccessPeer(&can, c.dev[size_t(i)].device, d0.device));
            if (!can) { c.peer_ok = false; break; }
            CUDA_OK(cudaSetDevice(c.dev[size_t(i)].device));
            cudaError_t pe = cudaDeviceEnablePeerAccess(d0.device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) c.peer_ok = false;
            cudaGetLastError();
        }
        if (c.peer_ok) {
            CUDA_OK(cudaSetDevice(d0.device));
            c.gather = static_cast<double*>(dmalloc(d0, c.rows.size() * size_t(c.width) * 4 * sizeof(double)));
        }
    }
    for (DeviceState& d : c.dev) { CUDA_OK(cudaSetDevice(d.device)); CUDA_OK(cudaStreamSynchronize(d.stream)); }
    c.stats.upload_ms = ms_since(t0);
    c.stats.h2d_bytes = h2d;
    c.stats.n_devices = nd;
    c.stats.rows = int(c.rows.size());
    c.stats.paths = int64_t(c.rows.size()) * c.width * c.samples;
    return guard.release();
}

// Adaptive launch order.  A launch that records per-tile clocks (Params::tile_cost: how long the slowest warp of each
// tile ran) lets the host sort the tiles longest-first for the launches that follow, which removes the tail that
// expensive tiles leave when they happen to start late -- on a 1/8-frame shard of the gopher scene 142.6 -> 106.4 ms
// at 8 slices, where the geometric estimate of plan_tile_order is too coarse.  Twice per context: after the 4-sample
// probe of the first render, and again after that render itself (2048 samples rank the tiles better than 4: with the
// probe's order alone two of the teapot's eight shards ran 90 ms instead of 73).
void adopt_measured_order(DeviceState& d, int level, int samples) {
    CUDA_OK(cudaSetDevice(d.device));
    std::vector<unsigned> cost(size_t(d.n_tiles));
    CUDA_OK(cudaMemcpyAsync(cost.data(), d.tile_cost, cost.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, d.stream));
    CUDA_OK(cudaStreamSynchronize(d.stream));
    std::vector<int> order(size_t(d.n_tiles));
    for (int k = 0; k < d.n_tiles; ++k) order[size_t(k)] = k;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[size_t(x)] > cost[size_t(y)]; });
    // Granularity.  A block lives for samples / slices samples of its tile, and no order can finish the launch before its
    // longest block.  Launched first, that block may take up to nearly the whole launch for free (gopher 1/8-frame shards
    // at 8 slices: longest block 88-103 ms of a 106 ms launch, and 8 slices beat 32: 106 vs 110 ms) -- but not more: on
    // two of the teapot's eight shards the costliest tile's block WAS the launch (90 ms, the other shards 72).  So after a
    // full render: if the longest block took more than the launch would need with its blocks packed perfectly (the sum of
    // the tiles' times spread over the resident warps), later launches use twice the slices and measure again.  After the probe there is only a noisy one-sample estimate, so the rule is coarse: 16 slices on a shard-sized
    // launch whose costliest 0.2 % of tiles are more than 5.5x the mean (the teapot shards that need it: 5.8-6.5, the
    // others 5.0-5.2, gopher 4.7-5.2).
    bool again = false;
    if (!std::getenv("PTC_SLICES") && samples >= 32 && !cost.empty()) {
        int sl = d.slices;
        if (level >= 2 && d.last_ms > 0.f) {
            int khz = 0;
            cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, d.device);
            const double to_ms = 256.0 / std::max(1.0, double(khz));
            const double longest_ms = double(cost[size_t(order[0])]) * to_ms;
            double total = 0.0;
            for (unsigned v : cost) total += double(v);
            // the launch if its blocks packed perfectly: every tile keeps `slices` warps busy for about its slowest warp's time
            const double balanced_ms = total * to_ms * double(d.slices) / (double(d.sm_count) * 32.0);
            if (longest_ms > 1.08 * balanced_ms && sl < ptk::kBlockWarps * ptk::kMaxCluster) { sl *= 2; again = true; }
            if (std::getenv("PTC_DEBUG_TIMING"))
                std::fprintf(stderr, "[order] full render: longest block %.1f ms, balanced %.1f ms, launch %.1f ms at %d slices -> %d slices\n", longest_ms, balanced_ms, d.last_ms, d.slices, sl);
        } else if (level == 1) {
            const size_t top = std::max<size_t>(1, cost.size() / 500);
            double sum = 0.0, sum_top = 0.0;
            for (size_t k = 0; k < cost.size(); ++k) { sum += cost[size_t(order[k])]; if (k < top) sum_top += cost[size_t(order[k])]; }
            const double r = sum > 0.0 ? (sum_top / double(top)) / (sum / double(cost.size())) : 0.0;
            const double tiles_per_warp_slot = double(cost.size()) / (double(d.sm_count) * 32.0);      // few tiles per resident warp: a shard
            if (r > 5.5 && tiles_per_warp_slot < 4.0) sl = std::max(sl, 16);
            if (std::getenv("PTC_DEBUG_TIMING")) std::fprintf(stderr, "[order] probe: r = %.2f -> %d slices\n", r, sl);
        }
        while (sl > samples) sl /= 2;
        d.slices = std::max(1, sl);
        d.cluster = std::max(1, d.slices / ptk::kBlockWarps);
    }
    if (again) level = 1;                     // keep recording: the next full render checks the new granularity
    CUDA_OK(cudaMemcpyAsync(d.tile_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, d.stream));
    CUDA_OK(cudaMemsetAsync(d.tile_cost, 0, cost.size() * sizeof(unsigned), d.stream));       // the next recording starts from zero
    CUDA_OK(cudaStreamSynchronize(d.stream));
    d.order_measured = level;
}

constexpr int kProbeSamples = 4;            // samples of the cost probe
constexpr int kProbeMinSamples = 128;       // renders shorter than this are not worth a probe

void trace_impl(ptc_context& c, int begin, int end, bool accumulate) {
    c.stats.kernel_launches = 0;
    for (DeviceState& d : c.dev) {
        CUDA_OK(cudaSetDevice(d.device));
        CUDA_OK(cudaEventRecord(d.ev0, d.stream));
    }
    // First full render of a context: a probe of kProbeSamples samples per pixel (0.2 % of a 2048-spp frame) measures
    // the tiles' relative cost, then the real launch starts the expensive ones first.  Its pixels are overwritten by
    // the render; its time is part of kernel_ms.
    if (!accumulate && end - begin >= kProbeMinSamples) {
        bool probing = false;
        for (DeviceState& d : c.dev) {
            if (!d.tile_cost || d.order_measured != 0 || d.rows.empty()) continue;
            CUDA_OK(cudaSetDevice(d.device));
            if (c.precision == PTC_FP64) launch<double>(c, d, c.scene64, begin, begin + kProbeSamples, false);
            else launch<float>(c, d, c.scene32, begin, begin + kProbeSamples, false);
            probing = true;
        }
        if (probing)
            for (DeviceState& d : c.dev)
                if (d.tile_cost && d.order_measured == 0 && !d.rows.empty()) adopt_measured_order(d, 1, c.samples);
    }
    for (DeviceState& d : c.dev) {
        CUDA_OK(cudaSetDevice(d.device));
        if (c.precision == PTC_FP64) launch<double>(c, d, c.scene64, begin, end, accumulate);
        else launch<float>(c, d, c.scene32, begin, end, accumulate);
        CUDA_OK(cudaEventRecord(d.ev1, d.stream));
    }
    double worst = 0.0;
    for (DeviceState& d : c.dev) {
        CUDA_OK(cudaSetDevice(d.device));
        CUDA_OK(cudaStreamSynchronize(d.stream));
        CUDA_OK(cudaEventElapsedTime(&d.last_ms, d.ev0, d.ev1));
        if (d.last_ms > worst) worst = d.last_ms;
    }
    c.stats.kernel_ms = worst;
    // the clocks of this launch order the next ones (a full render: final; progressive passes and short renders: until a longer one)
    for (DeviceState& d : c.dev)
        if (d.tile_cost && d.order_measured < 2 && !d.rows.empty() && end - begin >= kProbeSamples)
            adopt_measured_order(d, end - begin >= kProbeMinSamples ? 2 : 1, c.samples);
}

// The double frame of the context's rows on dev[0] when every device stored into the gather buffer; else NULL.
const double* gathered(const ptc_context& c) { return (c.dev.size() > 1 && c.gather && c.peer_ok && !c.frame) ? c.gather : nullptr; }

void read_impl(ptc_context& c, double* out) {
    auto t0 = Clock::now();
    if (c.frame) fail("ptc_read: the context renders into an attached frame; read it with ptc_frame_read");
    const size_t row_bytes = size_t(c.width) * 4 * sizeof(double);
    c.stats.d2h_bytes = 0; c.stats.p2p_bytes = 0;
    const int nd = int(c.dev.size());
    if (nd == 1 || gathered(c)) {
        DeviceState& d = c.dev[0];
        CUDA_OK(cudaSetDevice(d.device));
        CUDA_OK(cudaMemcpyAsync(out, nd == 1 ? d.out : c.gather, c.rows.size() * row_bytes, cudaMemcpyDeviceToHost, d.stream));
        CUDA_OK(cudaStreamSynchronize(d.stream));
        c.stats.d2h_bytes = int64_t(c.rows.size() * row_bytes);
        if (nd > 1) c.stats.p2p_bytes = int64_t((c.rows.size() - d.rows.size()) * row_bytes);     // stored over NVLink by the kernels
    } else {
        // No peer access: device i owns local tiles i, i+nd, i+2nd, ...: one strided 2-D copy per device places
        // them in the packed host frame (tile pitch nd*tile_bytes).
        const size_t tile_bytes = row_bytes * size_t(c.rows_per_tile);
        for (int i = 0; i < nd; ++i) {
            DeviceState& d = c.dev[size_t(i)];
            if (d.rows.empty()) continue;
            CUDA_OK(cudaSetDevice(d.device));
            const size_t full_tiles = d.rows.size() / size_t(c.rows_per_tile);
            const size_t tail_rows = d.rows.size() % size_t(c.rows_per_tile);
            char* dst_base = reinterpret_cast<char*>(out);
            if (full_tiles)
                CUDA_OK(cudaMemcpy2DAsync(dst_base + size_t(i) * tile_bytes, size_t(nd) * tile_bytes, d.out, tile_bytes, tile_bytes, full_tiles,
                                          cudaMemcpyDeviceToHost, d.stream));
            if (tail_rows)
                CUDA_OK(cudaMemcpyAsync(dst_base + (full_tiles * size_t(nd) + size_t(i)) * tile_bytes, reinterpret_cast<char*>(d.out) + full_tiles * tile_bytes,
                                        tail_rows * row_bytes, cudaMemcpyDeviceToHost, d.stream));
            c.stats.d2h_bytes += int64_t(d.rows.size() * row_bytes);
        }
        for (DeviceState& d : c.dev) { CUDA_OK(cudaSetDevice(d.device)); CUDA_OK(cudaStreamSynchronize(d.stream)); }
    }
    c.stats.read_ms = ms_since(t0);
}

template <typename F> int guarded(char* err, int errlen, F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        cudaGetLastError();
        return 1;
    }
}

}  // namespace

extern "C" {

const char* ptc_version(void) { return "libptcuda 0.2 (sm_100a) kernel " PTK_KERNEL_VERSION; }

int ptc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ptc_device_name(int index, char* buf, int buflen) {
    if (!buf || buflen <= 0) return 1;
    buf[0] = 0;
    int n = ptc_device_count();
    if (index < 0 || index >= n) return 1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, index) != cudaSuccess) { cudaGetLastError(); return 1; }
    std::snprintf(buf, size_t(buflen), "%s", prop.name);
    return 0;
}

int ptc_open(const ptc_job* job, ptc_context** ctx, char* err, int errlen) {
    if (ctx) *ctx = nullptr;
    return guarded(err, errlen, [&] {
        if (!job || !ctx) fail("ptc_open: NULL argument");
        *ctx = open_impl(*job);
    });
}

int ptc_trace(ptc_context* ctx, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!ctx) fail("ptc_trace: NULL context");
        trace_impl(*ctx, 0, ctx->samples, false);
    });
}

int ptc_trace_range(ptc_context* ctx, int32_t sample_begin, int32_t sample_end, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!ctx) fail("ptc_trace_range: NULL context");
        if (sample_begin < 0 || sample_end > ctx->samples || sample_begin >= sample_end)
            fail("ptc_trace_range: [%d, %d) is not a sub-range of [0, %d)", sample_begin, sample_end, ctx->samples);
        trace_impl(*ctx, sample_begin, sample_end, true);
    });
}

int ptc_reset(ptc_context* ctx, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!ctx) fail("ptc_reset: NULL context");
        for (DeviceState& d : ctx->dev) {
            if (!d.acc) continue;
            CUDA_OK(cudaSetDevice(d.device));
            CUDA_OK(cudaMemsetAsync(d.acc, 0, d.rows.size() * size_t(ctx->width) * 4 * sizeof(double), d.stream));
            CUDA_OK(cudaStreamSynchronize(d.stream));
        }
    });
}

// ptc_read_rgba8 / ptc_read_f32: convert on the device that holds the pixels, read back the narrow format.
// bytes_px = 4 (RGBA8, the frontend's tone step) or 16 (float RGBA).
static void read_converted(ptc_context& c, void* out, size_t bytes_px) {
    auto t0 = Clock::now();
    if (c.frame) fail("the context renders into an attached frame; read it with ptc_frame_read");
    const int nd = int(c.dev.size());
    const size_t row_bytes = size_t(c.width) * bytes_px;
    const size_t tile_bytes = row_bytes * size_t(c.rows_per_tile);
    c.stats.d2h_bytes = 0; c.stats.p2p_bytes = 0;
    auto convert = [&](DeviceState& d, const double* src, size_t pixels) -> void* {
        void*& dst = bytes_px == 4 ? reinterpret_cast<void*&>(d.rgba8) : reinterpret_cast<void*&>(d.f32);
        if (!dst) dst = dmalloc(d, std::max<size_t>(pixels, c.rows.size() * size_t(c.width)) * bytes_px);
        const unsigned blocks = (unsigned)((pixels + 255) / 256);
        if (bytes_px == 4) ptk::rgba8_kernel<<<blocks, 256, 0, d.stream>>>(reinterpret_cast<const double4*>(src), static_cast<uchar4*>(dst), int(pixels));
        else ptk::f32_kernel<<<blocks, 256, 0, d.stream>>>(reinterpret_cast<const double4*>(src), static_cast<float4*>(dst), int(pixels));
        CUDA_OK(cudaGetLastError());
        c.stats.kernel_launches++;
        return dst;
    };
    if (nd == 1 || gathered(c)) {
        DeviceState& d = c.dev[0];
        if (c.rows.empty()) { c.stats.read_ms = ms_since(t0); return; }
        CUDA_OK(cudaSetDevice(d.device));
        const size_t pixels = c.rows.size() * size_t(c.width);
        void* dst = convert(d, nd == 1 ? d.out : c.gather, pixels);
        CUDA_OK(cudaMemcpyAsync(out, dst, pixels * bytes_px, cudaMemcpyDeviceToHost, d.stream));
        CUDA_OK(cudaStreamSynchronize(d.stream));
        c.stats.d2h_bytes = int64_t(pixels * bytes_px);
    } else {
        for (int i = 0; i < nd; ++i) {
            DeviceState& d = c.dev[size_t(i)];
            if (d.rows.empty()) continue;
            CUDA_OK(cudaSetDevice(d.device));
            const size_t pixels = d.rows.size() * size_t(c.width);
            char* dst = static_cast<char*>(convert(d, d.out, pixels));
            // device i owns local tiles i, i+nd, ...: strided copy straight into the packed host frame
            const size_t full_tiles = d.rows.size() / size_t(c.rows_per_tile), tail_rows = d.rows.size() % size_t(c.rows_per_tile);
            char* o = static_cast<char*>(out);
            if (full_tiles)
                CUDA_OK(cudaMemcpy2DAsync(o + size_t(i) * tile_bytes, size_t(nd) * tile_bytes, dst, tile_bytes, tile_bytes, full_tiles, cudaMemcpyDeviceToHost, d.stream));
            if (tail_rows)
                CUDA_OK(cudaMemcpyAsync(o + (full_tiles * size_t(nd) + size_t(i)) * tile_bytes, dst + full_tiles * tile_bytes, tail_rows * row_bytes,
                                        cudaMemcpyDeviceToHost, d.stream));
            c.stats.d2h_bytes += int64_t(pixels * bytes_px);
        }
        for (DeviceState& d : c.dev) { CUDA_OK(cudaSetDevice(d.device)); CUDA_OK(cudaStreamSynchronize(d.stream)); }
    }
    c.stats.read_ms = ms_since(t0);
}

int ptc_read_rgba8(ptc_context* ctx, uint8_t* out_rgba8, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!ctx || !out_rgba8) fail("ptc_read_rgba8: NULL argument");
        read_converted(*ctx, out_rgba8, 4);
    });
}

int ptc_read_f32(ptc_context* ctx, float* out_rgba_f32, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!ctx || !out_rgba_f32) fail("ptc_read_f32: NULL argument");
        read_converted(*ctx, out_rgba_f32, 16);
    });
}

int ptc_read(ptc_context* ctx, double* out_rgba, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!ctx || !out_rgba) fail("ptc_read: NULL argument");
        read_impl(*ctx, out_rgba);
    });
}

int ptc_get_stats(const ptc_context* ctx, ptc_stats* stats) {
    if (!ctx || !stats) return 1;
    *stats = ctx->stats;
    return 0;
}

void ptc_close(ptc_context* ctx) { destroy(ctx); }

int ptc_set_seeds(ptc_context* ctx, const double* seeds, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!ctx || !seeds) fail("ptc_set_seeds: NULL argument");
        int64_t h2d = 0;
        for (DeviceState& d : ctx->dev) {
            CUDA_OK(cudaSetDevice(d.device));
            upload_seeds(d, seeds, ctx->width, h2d);
        }
        for (DeviceState& d : ctx->dev) { CUDA_OK(cudaSetDevice(d.device)); CUDA_OK(cudaStreamSynchronize(d.stream)); }
    });
}

// ---- frames: the fused gather ------------------------------------------------------------------------------------
int ptc_frame_create(int device, int32_t width, int32_t height, int32_t format, ptc_frame** frame, char* err, int errlen) {
    if (frame) *frame = nullptr;
    return guarded(err, errlen, [&] {
        if (!frame) fail("ptc_frame_create: NULL argument");
        if (width <= 0 || height <= 0 || (long long)width * height > (1ll << 30)) fail("ptc_frame_create: bad size %dx%d", width, height);
        if (format != PTC_FRAME_F64 && format != PTC_FRAME_F32) fail("ptc_frame_create: unknown format %d", format);
        if (ptc_device_count() <= 0) fail("no usable CUDA device; libptcuda has no CPU fallback");
        std::unique_ptr<ptc_frame> f(new ptc_frame);
        f->device = device < 0 ? 0 : device; f->width = width; f->height = height; f->format = format; f->owner = true;
        CUDA_OK(cudaSetDevice(f->device));
        CUDA_OK(cudaMalloc(&f->data, f->bytes()));          // plain cudaMalloc: exportable through CUDA IPC, reachable by peer stores
        CUDA_OK(cudaMemset(f->data, 0, f->bytes()));
        *frame = f.release();
    });
}

int ptc_frame_export(ptc_frame* frame, void* handle64, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!frame || !handle64) fail("ptc_frame_export: NULL argument");
        if (!frame->owner) fail("ptc_frame_export: only the process that created a frame can export it");
        static_assert(sizeof(cudaIpcMemHandle_t) == PTC_FRAME_HANDLE_BYTES, "handle size");
        cudaIpcMemHandle_t h;
        CUDA_OK(cudaSetDevice(frame->device));
        CUDA_OK(cudaIpcGetMemHandle(&h, frame->data));
        std::memcpy(handle64, &h, sizeof h);
    });
}

int ptc_frame_import(int device, const void* handle64, int32_t width, int32_t height, int32_t format, ptc_frame** frame, char* err, int errlen) {
    if (frame) *frame = nullptr;
    return guarded(err, errlen, [&] {
        if (!frame || !handle64) fail("ptc_frame_import: NULL argument");
        if (format != PTC_FRAME_F64 && format != PTC_FRAME_F32) fail("ptc_frame_import: unknown format %d", format);
        std::unique_ptr<ptc_frame> f(new ptc_frame);
        f->device = device < 0 ? 0 : device; f->width = width; f->height = height; f->format = format; f->owner = false;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, sizeof h);
        CUDA_OK(cudaSetDevice(f->device));                  // the mapping is made for the importing process's current device
        CUDA_OK(cudaIpcOpenMemHandle(&f->data, h, cudaIpcMemLazyEnablePeerAccess));
        *frame = f.release();
    });
}

int ptc_frame_read(ptc_frame* frame, void* out, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!frame || !out) fail("ptc_frame_read: NULL argument");
        CUDA_OK(cudaSetDevice(frame->device));
        CUDA_OK(cudaMemcpy(out, frame->data, frame->bytes(), cudaMemcpyDeviceToHost));
    });
}

int ptc_frame_device_pointer(ptc_frame* frame, void** dev_ptr, int64_t* bytes, int32_t* cuda_device) {
    if (!frame) return 1;
    if (dev_ptr) *dev_ptr = frame->data;
    if (bytes) *bytes = int64_t(frame->bytes());
    if (cuda_device) *cuda_device = frame->device;
    return 0;
}

void ptc_frame_destroy(ptc_frame* frame) {
    if (!frame) return;
    cudaSetDevice(frame->device);
    if (frame->data) { if (frame->owner) cudaFree(frame->data); else cudaIpcCloseMemHandle(frame->data); }
    cudaGetLastError();
    delete frame;
}

int ptc_set_frame(ptc_context* ctx, ptc_frame* frame, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!ctx) fail("ptc_set_frame: NULL context");
        if (frame) {
            if (frame->width != ctx->width || frame->height != ctx->height)
                fail("ptc_set_frame: frame is %dx%d, the context renders %dx%d", frame->width, frame->height, ctx->width, ctx->height);
            for (DeviceState& d : ctx->dev) {
                if (d.device == frame->device || !frame->owner) continue;     // an IPC mapping is already addressable from its device
                int can = 0;
                CUDA_OK(cudaDeviceCanAccessPeer(&can, d.device, frame->device));
                if (!can) fail("ptc_set_frame: device %d cannot store into device %d's memory (no peer access)", d.device, frame->device);
                CUDA_OK(cudaSetDevice(d.device));
                cudaError_t pe = cudaDeviceEnablePeerAccess(frame->device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CUDA_OK(pe);
                cudaGetLastError();
            }
        }
        ctx->frame = frame;
    });
}

int ptc_device_framebuffer(ptc_context* ctx, int local_index, void** dev_ptr, int64_t* n_doubles, int32_t* cuda_device) {
    if (!ctx || local_index < 0 || local_index >= int(ctx->dev.size())) return 1;
    DeviceState& d = ctx->dev[size_t(local_index)];
    if (gathered(*ctx)) {       // every device stored into dev[0]'s buffer: device 0 exposes all the context's rows, the others nothing
        if (dev_ptr) *dev_ptr = local_index == 0 ? ctx->gather : nullptr;
        if (n_doubles) *n_doubles = local_index == 0 ? int64_t(ctx->rows.size()) * ctx->width * 4 : 0;
    } else {
        if (dev_ptr) *dev_ptr = d.out;
        if (n_doubles) *n_doubles = int64_t(d.rows.size()) * ctx->width * 4;
    }
    if (cuda_device) *cuda_device = d.device;
    return 0;
}

int ptc_shard_rows(const ptc_context* ctx, int32_t* rows, int cap) {
    if (!ctx) return 0;
    const int n = int(ctx->rows.size());
    if (rows) for (int i = 0; i < n && i < cap; ++i) rows[i] = ctx->rows[size_t(i)];
    return n;
}

void ptc_trim(void) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (cudaMemPool_t& p : g_pools)
        if (p) cudaMemPoolTrimTo(p, 0);
}

int ptc_plan_rows(int32_t height, int32_t rows_per_tile, int32_t shard_index, int32_t shard_count, int32_t* rows, int cap) {
    if (height < 0) return -1;
    const int rpt = rows_per_tile > 0 ? rows_per_tile : 4;
    const int sc = shard_count > 1 ? shard_count : 1;
    const int si = shard_count > 1 ? shard_index : 0;
    if (si < 0 || si >= sc) return -1;
    int n = 0;
    for (int r = 0; r < height; ++r) {
        if ((r / rpt) % sc != si) continue;
        if (rows && n < cap) rows[n] = r;
        ++n;
    }
    return n;
}

int ptc_render(const ptc_job* job, double* out_rgba, char* err, int errlen) {
    ptc_context* ctx = nullptr;
    auto t0 = Clock::now();
    int rc = ptc_open(job, &ctx, err, errlen);
    if (rc != 0) return rc;
    auto t1 = Clock::now();
    rc = ptc_trace(ctx, err, errlen);
    auto t2 = Clock::now();
    if (rc == 0) rc = ptc_read(ctx, out_rgba, err, errlen);
    auto t3 = Clock::now();
    const double kernel_ms = ctx->stats.kernel_ms;
    ptc_close(ctx);
    if (std::getenv("PTC_DEBUG_TIMING"))
        std::fprintf(stderr, "[ptc_render] open %.2f ms, trace %.2f ms (kernel %.2f), read %.2f ms, close %.2f ms\n",
                     std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count(),
                     kernel_ms, std::chrono::duration<double, std::milli>(t3 - t2).count(), ms_since(t3));
    return rc;
}

int ptc_render_flat2(const void* objects, int32_t n_objects, const void* triangles, int32_t n_triangles, const void* groups,
                     int32_t n_groups, const void* camera, const uint8_t* tex_plane, const uint8_t* tex_sphere,
                     const uint8_t* tex_cube, const int32_t* tex_dims, const double* seeds, int32_t samples, int32_t precision,
                     int32_t rng_mode, int32_t features, const int32_t* devices, int32_t n_devices, double* out_rgba, char* err, int errlen) {
    ptc_job job;
    std::memset(&job, 0, sizeof job);
    job.abi_version = PTC_ABI_VERSION;
    job.objects = objects; job.n_objects = n_objects;
    job.triangles = triangles; job.n_triangles = n_triangles;
    job.groups = groups; job.n_groups = n_groups;
    job.camera = camera;
    const uint8_t* tex[3] = {tex_plane, tex_sphere, tex_cube};
    for (int k = 0; k < 3; ++k) {
        job.tex[k] = tex[k];
        if (tex[k] && tex_dims) { job.tex_w[k] = tex_dims[3 * k]; job.tex_h[k] = tex_dims[3 * k + 1]; job.tex_layers[k] = tex_dims[3 * k + 2]; }
    }
    job.seeds = seeds; job.samples = samples; job.precision = precision; job.rng_mode = rng_mode; job.features = features;
    job.devices = devices; job.n_devices = devices ? n_devices : 0;
    return ptc_render(&job, out_rgba, err, errlen);
}

int ptc_render_flat(const void* objects, int32_t n_objects, const void* triangles, int32_t n_triangles, const void* groups,
                    int32_t n_groups, const void* camera, const uint8_t* tex_plane, const uint8_t* tex_sphere,
                    const uint8_t* tex_cube, const int32_t* tex_dims, const double* seeds, int32_t samples, int32_t precision,
                    int32_t rng_mode, const int32_t* devices, int32_t n_devices, double* out_rgba, char* err, int errlen) {
    return ptc_render_flat2(objects, n_objects, triangles, n_triangles, groups, n_groups, camera, tex_plane, tex_sphere, tex_cube, tex_dims, seeds,
                            samples, precision, rng_mode, 0, devices, n_devices, out_rgba, err, errlen);
}

// Test hook (not part of the drop-in surface): evaluate noise3D on the device.
int ptc_debug_noise3d(const float* xyz, int n, int rng_mode, float* out, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!xyz || !out || n < 0) fail("ptc_debug_noise3d: bad argument");
        if (ptc_device_count() <= 0) fail("no usable CUDA device; libptcuda has no CPU fallback");
        float *dx = nullptr, *dout = nullptr;
        CUDA_OK(cudaSetDevice(0));
        CUDA_OK(cudaMalloc(&dx, size_t(n) * 3 * sizeof(float) + 16));
        CUDA_OK(cudaMalloc(&dout, size_t(n) * sizeof(float) + 16));
        CUDA_OK(cudaMemcpy(dx, xyz, size_t(n) * 3 * sizeof(float), cudaMemcpyHostToDevice));
        if (n > 0) ptk::noise3d_kernel<<<(n + 255) / 256, 256>>>(dx, dout, n, rng_mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(out, dout, size_t(n) * sizeof(float), cudaMemcpyDeviceToHost);
        cudaFree(dx); cudaFree(dout);
        CUDA_OK(e);
    });
}

// Measurement hook (not part of the drop-in surface): achieved FP32 FFMA throughput of a device in TFLOP/s
// (2 flops per FFMA), best of 5 launches of ptk::fma_peak_kernel -- calibrates the issue roofline bench.py reports against.
int ptc_debug_fma_peak(int device, double* tflops, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!tflops) fail("ptc_debug_fma_peak: tflops is NULL");
        if (ptc_device_count() <= 0) fail("no usable CUDA device; libptcuda has no CPU fallback");
        CUDA_OK(cudaSetDevice(device));
        int sms = 0;
        CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        const int blocks = sms * 8, threads = 256, iters = 4096;
        float* out = nullptr;
        CUDA_OK(cudaMalloc(&out, size_t(blocks) * threads * sizeof(float)));
        cudaEvent_t e0, e1;
        CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {                       // first launch = warm-up
            CUDA_OK(cudaEventRecord(e0));
            ptk::fma_peak_kernel<<<blocks, threads>>>(out, iters);
            CUDA_OK(cudaEventRecord(e1));
            CUDA_OK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
        CUDA_OK(cudaGetLastError());
        *tflops = 2.0 * 8 * 16 * double(iters) * blocks * threads / (best * 1e-3) / 1e12;
    });
}

#ifdef PTK_HIST
int ptc_debug_hist(unsigned long long* out40, int reset) {
    cudaDeviceSynchronize();
    if (out40) cudaMemcpyFromSymbol(out40, ptk::ptk_hist, 40 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[40] = {}; cudaMemcpyToSymbol(ptk::ptk_hist, z, sizeof z); }
    return 0;
}
#endif

// The FP64 twin: achieved DFMA throughput in TFLOP/s (the roofline of the fp64 mode).
int ptc_debug_dfma_peak(int device, double* tflops, char* err, int errlen) {
    return guarded(err, errlen, [&] {
        if (!tflops) fail("ptc_debug_dfma_peak: tflops is NULL");
        if (ptc_device_count() <= 0) fail("no usable CUDA device; libptcuda has no CPU fallback");
        CUDA_OK(cudaSetDevice(device));
        int sms = 0;
        CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        const int blocks = sms * 8, threads = 256, iters = 1024;
        double* out = nullptr;
        CUDA_OK(cudaMalloc(&out, size_t(blocks) * threads * sizeof(double)));
        cudaEvent_t e0, e1;
        CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {                       // first launch = warm-up
            CUDA_OK(cudaEventRecord(e0));
            ptk::dfma_peak_kernel<<<blocks, threads>>>(out, iters);
            CUDA_OK(cudaEventRecord(e1));
            CUDA_OK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
        CUDA_OK(cudaGetLastError());
        *tflops = 2.0 * 8 * 16 * double(iters) * blocks * threads / (best * 1e-3) / 1e12;
    });
}

// Test hook (not part of the drop-in surface; host only, no device needed): the launch plan ptc_open would make for a
// device with `sm_count` SMs that owns the job's shard -- sample slices per pixel and the initial (geometric) launch
// order of its 8x4-pixel tiles.  Returns the number of tiles; fills at most `cap` entries of `order` (identity when the
// scene has nothing to reorder); -1 with a message on a bad job.
int ptc_debug_launch_plan(const ptc_job* job, int sm_count, int32_t* slices, int32_t* order, int cap, char* err, int errlen) {
    int n_tiles = -1;
    guarded(err, errlen, [&] {
        if (!job) fail("job is NULL");
        validate(*job);
        const auto* cam = static_cast<const ptw_camera*>(job->camera);
        const int rpt = job->rows_per_tile > 0 ? job->rows_per_tile : 4, sc = job->shard_count > 1 ? job->shard_count : 1;
        const int si = job->shard_count > 1 ? job->shard_index : 0;
        std::vector<int> rows;
        for (int r = 0; r < cam->height; ++r) if ((r / rpt) % sc == si) rows.push_back(r);
        HostScene<float> hs;
        flatten<float>(*job, hs);
        if (slices) *slices = plan_slices((long long)rows.size() * cam->width, sm_count, hs.n_mesh > 0, job->samples);
        const int tiles = ((cam->width + ptk::kTileW - 1) / ptk::kTileW) * ((int(rows.size()) + ptk::kTileH - 1) / ptk::kTileH);
        std::vector<int> plan = hs.n_mesh > 0 ? plan_tile_order(*job, rows) : std::vector<int>();
        for (int k = 0; k < tiles && k < cap && order; ++k) order[k] = plan.empty() ? k : plan[size_t(k)];
        n_tiles = tiles;
    });
    return n_tiles;
}

// Test hook (not part of the drop-in surface; host only, no device needed): flattens the job's scene in
// double and copies one array of the rebuilt mesh index out, so tests can check the builder's invariants
// (every triangle in exactly one leaf, child boxes containing their triangles, depth) and replay the walk.
// what: 0 wide nodes (8 children x 2 records of 4 doubles: lo.xyz + child code, hi.xyz + 0), 4 tri_info, 5 tri_test, 6 node_lo, 7 node_hi, 8 node_parent,
// 9 mesh records (8 doubles per object: root_lo, root_hi, bvh_root, flags), 10 object node ranges (2 ints per object).
// Returns the array's size in bytes (copying at most cap_bytes), or -1 with a message.
int64_t ptc_debug_mesh_index(const ptc_job* job, int what, void* out, int64_t cap_bytes, char* err, int errlen) {
    int64_t bytes = -1;
    int rc = guarded(err, errlen, [&] {
        if (!job) fail("job is NULL");
        validate(*job);
        HostScene<double> hs;
        flatten<double>(*job, hs);
        std::vector<double> mesh;
        std::vector<int> ranges;
        for (int i = 0; i < job->n_objects; ++i) {
            const ptk::DMesh<double>& m = hs.mesh[size_t(i)];
            for (int a = 0; a < 3; ++a) mesh.push_back(m.root_lo[a]);
            for (int a = 0; a < 3; ++a) mesh.push_back(m.root_hi[a]);
            mesh.push_back(double(m.bvh_root)); mesh.push_back(double(m.flags));
            ranges.push_back(hs.hot[i].node_begin); ranges.push_back(hs.hot[i].node_end);
        }
        const void* src = nullptr;
        switch (what) {
            case 0: src = hs.wide.data(); bytes = int64_t(hs.wide.size() * sizeof(hs.wide[0])); break;
            case 4: src = hs.tri_info.data(); bytes = int64_t(hs.tri_info.size() * sizeof(int2)); break;
            case 5: src = hs.tri_test.data(); bytes = int64_t(hs.tri_test.size() * sizeof(hs.tri_test[0])); break;
            case 6: src = hs.node_lo.data(); bytes = int64_t(hs.node_lo.size() * sizeof(hs.node_lo[0])); break;
            case 7: src = hs.node_hi.data(); bytes = int64_t(hs.node_hi.size() * sizeof(hs.node_hi[0])); break;
            case 8: src = hs.node_parent.data(); bytes = int64_t(hs.node_parent.size() * sizeof(int)); break;
            case 9: src = mesh.data(); bytes = int64_t(mesh.size() * sizeof(double)); break;
            case 10: src = ranges.data(); bytes = int64_t(ranges.size() * sizeof(int)); break;
            default: fail("ptc_debug_mesh_index: unknown array %d", what);
        }
        if (out && cap_bytes > 0 && bytes > 0) std::memcpy(out, src, size_t(std::min(bytes, cap_bytes)));
    });
    return rc == 0 ? bytes : -1;
}

}  // extern "C"
