// s4_tree.cu — S4: resolve the python renderer's tile quadtree on the device, then depth-sort every leaf's list.
//
// Reference semantics restated (not copied): gauss_render.py:290-344
//   BFS over tiles; a tile is skipped if w <= 1 or h <= 1 (:301), background-filled if no Gaussian overlaps it
//   (:313-315), split into TL, BL, TR, BR children if it holds more than max_gaussians_per_tile Gaussians or is wider /
//   taller than max_tile_size (:319-335), else rendered with its Gaussians ordered nearest-first (:340-344).
// The per-node overlap counts come from the preprocess kernel; one CTA walks the levels top-down (a level has at most
// a few thousand nodes), numbers the leaves in the reference's BFS order (level-major, then child-rank path order) and
// lays out the instance / pixel offsets.  The per-leaf depth sort is a library call (cub::DeviceSegmentedSort).
#include <cub/cub.cuh>
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
    int32_t max_leaves;
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
    if (threadIdx.x == 0) {
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
                               const uint32_t* node_cnt, uint8_t* node_state, int32_t* leaf_of_node,
                               g2pc_leaf_t* leaves, int32_t* seg_begin, int32_t max_leaves, int32_t* header,
                               void* stream) {
    G2PC_CHECK_ARG(tables && node_cnt && node_state && leaf_of_node && leaves && seg_begin && header, "null pointer");
    G2PC_CHECK_ARG(num_levels >= 1 && num_levels <= G2PC_MAX_LEVELS && max_leaves >= 1, "bad sizes");
    TreeParams p;
    p.meta.num_levels = num_levels; p.meta.max_gaussians_per_tile = max_gaussians_per_tile;
    p.meta.width = 0; p.meta.height = 0;
    p.n1 = (1 << num_levels) - 1;
    p.tab.xs = tables; p.tab.xe = tables + p.n1; p.tab.xf = tables + 2 * p.n1;
    p.tab.ys = tables + 3 * p.n1; p.tab.ye = tables + 4 * p.n1; p.tab.yf = tables + 5 * p.n1;
    p.node_cnt = node_cnt; p.node_state = node_state; p.leaf_of_node = leaf_of_node; p.leaves = leaves;
    p.seg_begin = seg_begin; p.max_leaves = max_leaves; p.header = header;
    tree_kernel<<<1, TB, 0, (cudaStream_t)stream>>>(p);
    G2PC_CHECK_LAUNCH();
    return G2PC_OK;
}

extern "C" int64_t g2pc_sort_workspace_bytes(int64_t num_items, int32_t num_segments) {
    size_t bytes = 0;
    cub::DoubleBuffer<unsigned long long> keys(nullptr, nullptr);
    cudaError_t e = cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, keys, num_items, num_segments,
                                                       (const int32_t*)nullptr, (const int32_t*)nullptr);
    if (e != cudaSuccess) return -1;
    return (int64_t)bytes;
}

extern "C" int g2pc_sort_leaves(uint64_t* keys, uint64_t* keys_alt, int64_t num_items, int32_t num_segments,
                                const int32_t* seg_begin, void* workspace, int64_t workspace_bytes,
                                int32_t* sorted_in_alt_host, void* stream) {
    G2PC_CHECK_ARG(num_items >= 0 && num_segments >= 0, "negative size");
    if (sorted_in_alt_host) *sorted_in_alt_host = 0;
    if (num_items == 0 || num_segments == 0) return G2PC_OK;
    G2PC_CHECK_ARG(keys && keys_alt && seg_begin && workspace && sorted_in_alt_host, "null pointer");
    cub::DoubleBuffer<unsigned long long> db((unsigned long long*)keys, (unsigned long long*)keys_alt);
    size_t bytes = (size_t)workspace_bytes;
    G2PC_CUDA(cub::DeviceSegmentedSort::SortKeys(workspace, bytes, db, num_items, num_segments, seg_begin,
                                                 seg_begin + 1, (cudaStream_t)stream));
    *sorted_in_alt_host = (db.Current() == (unsigned long long*)keys_alt) ? 1 : 0;
    return G2PC_OK;
}
