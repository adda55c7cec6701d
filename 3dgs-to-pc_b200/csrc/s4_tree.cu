// s4_tree.cu — S4: resolve the python renderer's tile quadtree on the device, then depth-sort every leaf's list.
//
// Reference semantics restated (not copied): gauss_render.py:290-344
//   BFS over tiles; a tile is skipped if w <= 1 or h <= 1 (:301), background-filled if no Gaussian overlaps it
//   (:313-315), split into TL, BL, TR, BR children if it holds more than max_gaussians_per_tile Gaussians or is wider /
//   taller than max_tile_size (:319-335), else rendered with its Gaussians ordered nearest-first (:340-344).
// The per-node overlap counts come from the preprocess kernel; one CTA walks the levels top-down (a level has at most
// a few thousand nodes), numbers the leaves in the reference's BFS order (level-major, then child-rank path order) and
// lays out the instance / pixel offsets.
//
// Depth ordering without a per-leaf sort: the Gaussians are sorted ONCE per camera by view depth (N keys), instances
// are emitted in that order, and a stable LSD radix sort on the leaf id alone (ceil(log2(#leaves)) bits, 2 passes)
// groups them by leaf while keeping the depth order inside every leaf.  The radix sorts / scan are library calls
// (cub::DeviceRadixSort, cub::DeviceScan).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include "colour_common.cuh"

namespace {

constexpr int TB = 1024;

struct TreeParams {
    QtMeta meta;
    QtTables tab;
    int32_t n1;
    const uint32_t* node_cnt;
    uint8_t* node_state;
    int32_t* leaf_of_node;
    g2pc_leaf_t* leaves;
    int32_t* seg_begin;  // max_leaves + 1
    int32_t* leaf_order; // leaves sorted by descending work (longest-processing-time-first launch order)
    int32_t max_leaves;
    const uint32_t* incl;  // inclusive scan of touched[] in depth order (last entry = instance upper bound) or null
    int64_t n;
    int32_t* header;  // G2PC_HDR_WORDS
};

__device__ __forceinline__ int off2d(int l) { return ((1 << (2 * l)) - 1) / 3; }

// block-wide exclusive scan for TB threads; returns prefix, sets total
__device__ __forceinline__ int block_scan_1024(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;      // exclusive prefix of warp totals
        if (lane == 31) s_warp[32] = winc;  // grand total
    }
    __syncthreads();
    const int res = s_warp[warp] + inc - v;
    total = s_warp[32];
    __syncthreads();
    return res;
}

// key (BFS rank within a level: child rank = 2*xbit + ybit per level, most significant first) -> (ix, iy)
__device__ __forceinline__ void deinterleave(int key, int level, int& ix, int& iy) {
    ix = 0; iy = 0;
    for (int b = 0; b < level; ++b) {
        iy |= ((key >> (2 * b)) & 1) << b;
        ix |= ((key >> (2 * b + 1)) & 1) << b;
    }
}

__global__ void __launch_bounds__(TB) tree_kernel(const TreeParams p) {
    __shared__ int s_warp[33];
    __shared__ int s_flags[2];
    const int L = p.meta.num_levels;
    if (threadIdx.x == 0) { s_flags[0] = 0; s_flags[1] = 0; }
    __syncthreads();
    int leaf_base = 0;
    for (int l = 0; l < L; ++l) {
        const int nn = 1 << (2 * l);
        const int o1 = (1 << l) - 1, o2 = off2d(l);
        for (int k0 = 0; k0 < nn; k0 += TB) {
            const int key = k0 + threadIdx.x;
            int is_leaf = 0, node = -1, ix = 0, iy = 0;
            uint32_t cnt = 0;
            if (key < nn) {
                deinterleave(key, l, ix, iy);
                node = o2 + (iy << l) + ix;
                bool exists = (l == 0);
                if (l > 0) {
                    const int pnode = off2d(l - 1) + ((iy >> 1) << (l - 1)) + (ix >> 1);
                    exists = p.node_state[pnode] == NODE_SPLIT;
                }
                uint8_t st = NODE_NONE;
                if (exists) {
                    const int fx = p.tab.xf[o1 + ix], fy = p.tab.yf[o1 + iy];
                    if (!((fx | fy) & QT_FLAG_DROPPED)) {
                        cnt = p.node_cnt[node];
                        if (cnt == 0) st = NODE_EMPTY;
                        else if (((fx | fy) & QT_FLAG_BIG) || cnt > (uint32_t)p.meta.max_gaussians_per_tile) {
                            st = NODE_SPLIT;
                            if (l == L - 1) { st = NODE_NONE; s_flags[0] = 1; }  // deeper than the tabulated levels
                        } else {
                            st = NODE_LEAF;
                            is_leaf = 1;
                        }
                    }
                }
                p.node_state[node] = st;
            }
            int tot;
            const int pre = block_scan_1024(is_leaf, s_warp, tot);
            if (is_leaf) {
                const int li = leaf_base + pre;
                if (li < p.max_leaves) {
                    g2pc_leaf_t lf;
                    lf.r0 = p.tab.ys[o1 + iy];
                    lf.c0 = p.tab.xs[o1 + ix];
                    lf.w = p.tab.xe[o1 + ix] - lf.c0 + 1;
                    lf.h = p.tab.ye[o1 + iy] - lf.r0 + 1;
                    lf.inst_begin = 0;
                    lf.inst_count = (int32_t)cnt;
                    lf.pix_offset = 0;
                    lf.node = node;
                    p.leaves[li] = lf;
                    p.leaf_of_node[node] = li;
                } else {
                    s_flags[1] = 1;
                }
            }
            leaf_base += tot;
        }
        __syncthreads();  // node_state of level l visible to level l+1
    }
    const int nl = leaf_base < p.max_leaves ? leaf_base : p.max_leaves;
    // exclusive scans of the instance counts and pixel counts over the leaves, in order
    int inst_base = 0, pix_base = 0;
    for (int k0 = 0; k0 < nl; k0 += TB) {
        const int i = k0 + threadIdx.x;
        int c = 0, a = 0;
        if (i < nl) { c = p.leaves[i].inst_count; a = p.leaves[i].w * p.leaves[i].h; }
        int tc, ta;
        const int pc = block_scan_1024(c, s_warp, tc);
        const int pa = block_scan_1024(a, s_warp, ta);
        if (i < nl) {
            p.leaves[i].inst_begin = inst_base + pc;
            p.leaves[i].pix_offset = pix_base + pa;
            p.seg_begin[i] = inst_base + pc;
        }
        inst_base += tc;
        pix_base += ta;
    }
    __syncthreads();
    // launch order for the blend: heaviest leaves first (rank by instances x pixels, ties by index)
    for (int i = threadIdx.x; i < nl; i += TB) {
        const long long wi = (long long)p.leaves[i].inst_count * (p.leaves[i].w * p.leaves[i].h);
        int rank = 0;
        for (int j = 0; j < nl; ++j) {
            const long long wj = (long long)p.leaves[j].inst_count * (p.leaves[j].w * p.leaves[j].h);
            rank += (wj > wi || (wj == wi && j < i)) ? 1 : 0;
        }
        p.leaf_order[rank] = i;
    }
    if (threadIdx.x == 0) {
        p.header[G2PC_HDR_TOTAL_UPPER] = (p.incl && p.n > 0) ? (int32_t)p.incl[p.n - 1] : 0;
        p.seg_begin[nl] = inst_base;
        p.header[G2PC_HDR_NUM_LEAVES] = leaf_base;
        p.header[G2PC_HDR_TOTAL_INST] = inst_base;
        p.header[G2PC_HDR_TOTAL_PIX] = pix_base;
        p.header[G2PC_HDR_NEED_DEEPER] = s_flags[0];
        p.header[G2PC_HDR_LEAF_OVERFLOW] = s_flags[1];
    }
}

}  // namespace

extern "C" int g2pc_build_tree(const int32_t* tables, int32_t num_levels, int32_t max_gaussians_per_tile,
                               const uint32_t* node_cnt, const uint32_t* incl, int64_t n, uint8_t* node_state,
                               int32_t* leaf_of_node, g2pc_leaf_t* leaves, int32_t* seg_begin, int32_t* leaf_order,
                               int32_t max_leaves, int32_t* header, void* stream) {
    G2PC_CHECK_ARG(tables && node_cnt && node_state && leaf_of_node && leaves && seg_begin && leaf_order && header,
                   "null pointer");
    G2PC_CHECK_ARG(num_levels >= 1 && num_levels <= G2PC_MAX_LEVELS && max_leaves >= 1, "bad sizes");
    TreeParams p;
    p.meta.num_levels = num_levels; p.meta.max_gaussians_per_tile = max_gaussians_per_tile;
    p.meta.width = 0; p.meta.height = 0;
    p.n1 = (1 << num_levels) - 1;
    p.tab.xs = tables; p.tab.xe = tables + p.n1; p.tab.xf = tables + 2 * p.n1;
    p.tab.ys = tables + 3 * p.n1; p.tab.ye = tables + 4 * p.n1; p.tab.yf = tables + 5 * p.n1;
    p.node_cnt = node_cnt; p.node_state = node_state; p.leaf_of_node = leaf_of_node; p.leaves = leaves;
    p.seg_begin = seg_begin; p.leaf_order = leaf_order; p.max_leaves = max_leaves; p.incl = incl; p.n = n;
    p.header = header;
    tree_kernel<<<1, TB, 0, (cudaStream_t)stream>>>(p);
    G2PC_CHECK_LAUNCH();
    return G2PC_OK;
}

namespace {
struct GatherTouched {
    const uint32_t* touched;
    const uint32_t* order;
    __host__ __device__ __forceinline__ uint32_t operator()(const int64_t k) const { return touched[order[k]]; }
};
__global__ void iota_kernel(uint32_t* v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

// workspace layout of g2pc_depth_order: [keys_out n u32][iota n u32][cub temp]
extern "C" int64_t g2pc_depth_order_workspace_bytes(int64_t n) {
    size_t sort_b = 0, scan_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
    GatherTouched op{nullptr, nullptr};
    auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), op);
    cub::DeviceScan::InclusiveSum(nullptr, scan_b, it, (uint32_t*)nullptr, n);
    return (int64_t)(2 * align256((size_t)n * 4) + align256(sort_b > scan_b ? sort_b : scan_b));
}

/* Sort the Gaussians by depth key (stable: ties keep index order) -> order[k]; then incl[k] = inclusive prefix sum of
 * touched[order[k]]. */
extern "C" int g2pc_depth_order(const uint32_t* depth_key, const uint32_t* touched, int64_t n, uint32_t* order,
                                uint32_t* incl, void* workspace, int64_t workspace_bytes, void* stream) {
    G2PC_CHECK_ARG(n >= 0, "n < 0");
    if (n == 0) return G2PC_OK;
    G2PC_CHECK_ARG(depth_key && touched && order && incl && workspace, "null pointer");
    G2PC_CHECK_ARG(workspace_bytes >= g2pc_depth_order_workspace_bytes(n), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    uint32_t* keys_out = (uint32_t*)ws;
    uint32_t* iota = (uint32_t*)(ws + align256((size_t)n * 4));
    void* tmp = ws + 2 * align256((size_t)n * 4);
    size_t tmp_b = (size_t)workspace_bytes - 2 * align256((size_t)n * 4);
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(iota, n);
    G2PC_CHECK_LAUNCH();
    size_t b = tmp_b;
    G2PC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b, depth_key, keys_out, (const uint32_t*)iota, order, n, 0, 32, st));
    GatherTouched op{touched, order};
    auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), op);
    b = tmp_b;
    G2PC_CUDA(cub::DeviceScan::InclusiveSum(tmp, b, it, incl, n, st));
    return G2PC_OK;
}

extern "C" int64_t g2pc_sort_instances_workspace_bytes(int64_t num_items) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    if (cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, num_items, 0, 32) != cudaSuccess) return -1;
    return (int64_t)bytes;
}

/* Stable radix sort of the (leaf id, gid) instance pairs on the low `leaf_bits` bits of the leaf id (padding entries
 * carry 0xFFFFFFFF and sort to the end).  *sorted_in_alt_host = 1 if the result ended in the *_alt buffers. */
extern "C" int g2pc_sort_instances(uint32_t* inst_leaf, uint32_t* inst_leaf_alt, uint32_t* inst_gid,
                                   uint32_t* inst_gid_alt, int64_t num_items, int32_t leaf_bits, void* workspace,
                                   int64_t workspace_bytes, int32_t* sorted_in_alt_host, void* stream) {
    G2PC_CHECK_ARG(num_items >= 0 && leaf_bits >= 1 && leaf_bits <= 32, "bad sizes");
    if (sorted_in_alt_host) *sorted_in_alt_host = 0;
    if (num_items == 0) return G2PC_OK;
    G2PC_CHECK_ARG(inst_leaf && inst_leaf_alt && inst_gid && inst_gid_alt && workspace && sorted_in_alt_host,
                   "null pointer");
    cub::DoubleBuffer<uint32_t> k(inst_leaf, inst_leaf_alt), v(inst_gid, inst_gid_alt);
    size_t bytes = (size_t)workspace_bytes;
    G2PC_CUDA(cub::DeviceRadixSort::SortPairs(workspace, bytes, k, v, num_items, 0, leaf_bits, (cudaStream_t)stream));
    *sorted_in_alt_host = (v.Current() == inst_gid_alt) ? 1 : 0;
    return G2PC_OK;
}
