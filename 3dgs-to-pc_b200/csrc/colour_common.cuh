// colour_common.cuh — shared pieces of the colour stage (S3-S6): quadtree tables in shared memory, range queries.
#pragma once
#include "common.cuh"

#define QT_FLAG_DROPPED 1
#define QT_FLAG_BIG 2

// node states written by the tree kernel
#define NODE_NONE 0   // does not exist / dropped
#define NODE_EMPTY 1  // exists, no Gaussian overlaps it (background fill, gauss_render.py:313-315)
#define NODE_SPLIT 2  // exists and was split into 4 children (gauss_render.py:319-335)
#define NODE_LEAF 3   // exists and is rendered

// node range of a Gaussian at the first leaf-candidate level, packed 8 bits per bound
#define G2PC_RANGE_MAX_LEVEL 8
#define G2PC_RANGE_EMPTY 0x00000001u  // xlo = 1 > xhi = 0
__host__ __device__ __forceinline__ uint32_t g2pc_pack_range(int xlo, int xhi, int ylo, int yhi) {
    return (uint32_t)xlo | ((uint32_t)xhi << 8) | ((uint32_t)ylo << 16) | ((uint32_t)yhi << 24);
}
__host__ __device__ __forceinline__ void g2pc_unpack_range(uint32_t r, int& xlo, int& xhi, int& ylo, int& yhi) {
    xlo = (int)(r & 255u); xhi = (int)((r >> 8) & 255u); ylo = (int)((r >> 16) & 255u); yhi = (int)(r >> 24);
}

// Frame failure word shared by the frames in flight (uint32, 0xFFFFFFFF = none, else 1 + the LOWEST frame that did not fit):
// every kernel of frame f does nothing iff f + 1 >= *fail.  Frames of two CUDA streams may be in flight at once, so a
// later frame can fail before an earlier one has finished: the earlier one must still complete.
__device__ __forceinline__ bool g2pc_frame_skipped(const uint32_t* fail, int frame) {
    return (uint32_t)(frame + 1) >= *fail;
}

struct QtMeta {
    int32_t num_levels;  // tabulated levels 0..num_levels-1
    int32_t max_gaussians_per_tile;
    int32_t width, height;
};

// per-Gaussian projection record: 3 x float4
//   q0 = (mx, my, c00', c01')     c' = conic * (-0.5 * log2(e));  c01' = (conic01 + conic10)'
//   q1 = (c11', opacity, r, g)
//   q2 = (b, depth, radius, valid)   valid: 1.0f if in front of the camera (gauss_render.py:167), else 0
struct QtTables {
    const int32_t* xs; const int32_t* xe; const int32_t* xf;
    const int32_t* ys; const int32_t* ye; const int32_t* yf;
};

// copy the 1-D tables (6 arrays of n1 ints) into shared memory; returns pointers into smem
__device__ __forceinline__ QtTables load_tables(const QtTables g, int n1, int32_t* smem) {
    for (int i = threadIdx.x; i < n1; i += blockDim.x) {
        smem[i] = g.xs[i];
        smem[n1 + i] = g.xe[i];
        smem[2 * n1 + i] = g.xf[i];
        smem[3 * n1 + i] = g.ys[i];
        smem[4 * n1 + i] = g.ye[i];
        smem[5 * n1 + i] = g.yf[i];
    }
    __syncthreads();
    QtTables s;
    s.xs = smem; s.xe = smem + n1; s.xf = smem + 2 * n1;
    s.ys = smem + 3 * n1; s.ye = smem + 4 * n1; s.yf = smem + 5 * n1;
    return s;
}

// Members of the interval (rmin, rmax) among the 2^level nodes of one axis at `level`:
//   min(rmax, e_i) > max(rmin, s_i)   (fp32 compares, strict — gauss_render.py:308-310)
//   <=>  rmax > rmin  &&  rmax > s_i  &&  e_i > rmin  &&  e_i > s_i
// starts / ends are non-decreasing within a level, so the candidates form the index range [lo, hi]; nodes inside the
// range that are dropped or degenerate (e_i <= s_i) are filtered by the caller through axis_member().
__device__ __forceinline__ void axis_range(const int32_t* __restrict__ s, const int32_t* __restrict__ e, int level,
                                           float rmin, float rmax, float inv_step, int& lo, int& hi) {
    const int n = 1 << level;
    if (!(rmax > rmin)) { lo = 1; hi = 0; return; }
    // the nodes of a level are (nearly) uniformly spaced: start from the arithmetic guess and walk to the exact answer
    // (0-2 steps in practice) instead of a binary search.  inv_step = 2^level / extent.
    // lo = first i with e_i > rmin
    int a = min(n - 1, max(0, (int)(rmin * inv_step)));
    while (a < n && !((float)e[a] > rmin)) ++a;
    while (a > 0 && (float)e[a - 1] > rmin) --a;
    lo = a;
    // hi = last i with s_i < rmax
    a = min(n - 1, max(0, (int)(rmax * inv_step)));
    while (a >= 0 && !((float)s[a] < rmax)) --a;
    while (a + 1 < n && (float)s[a + 1] < rmax) ++a;
    hi = a;
}

// same query, started from a guess of the answer (lo_guess / hi_guess within a node or two of the exact bounds)
__device__ __forceinline__ void axis_range_from(const int32_t* __restrict__ s, const int32_t* __restrict__ e, int level,
                                                float rmin, float rmax, int lo_guess, int hi_guess, int& lo, int& hi) {
    const int n = 1 << level;
    if (!(rmax > rmin)) { lo = 1; hi = 0; return; }
    int a = min(n - 1, max(0, lo_guess));
    while (a < n && !((float)e[a] > rmin)) ++a;
    while (a > 0 && (float)e[a - 1] > rmin) --a;
    lo = a;
    a = min(n - 1, max(0, hi_guess));
    while (a >= 0 && !((float)s[a] < rmax)) --a;
    while (a + 1 < n && (float)s[a + 1] < rmax) ++a;
    hi = a;
}

__device__ __forceinline__ bool axis_member(const int32_t* __restrict__ s, const int32_t* __restrict__ e,
                                            const int32_t* __restrict__ f, int i) {
    return !(f[i] & QT_FLAG_DROPPED) && e[i] > s[i];
}

// Warp-cooperative walk over per-lane node rectangles.  Every lane brings a packed rectangle (g2pc_pack_range; empty if
// xlo > xhi) and a 32-bit payload; f(ix, iy, owner_lane, owner_payload) is called once per (lane, node).  Rectangles of up
// to WARP_SMALL_AREA nodes are walked by their own lane; larger ones (a few huge splats cover hundreds of tiles) are
// walked by the whole warp — the lanes tile the rectangle with a power-of-two number of columns, so no division is
// needed — otherwise the warp runs at the speed of its largest rectangle (ncu r02a: 6.9 of 32 threads active in the
// multisplit).  Must be called by all 32 lanes.
// (a cooperative iteration costs ~45 instructions of shuffles / loop control, a private node ~12: the break-even rectangle
// is ~10 nodes; with a threshold of 4 the typical 2x3 / 3x3 rectangles all went the slow cooperative way)
constexpr int WARP_SMALL_AREA = 12;
template <typename F>
__device__ __forceinline__ void warp_for_each_node(uint32_t range, uint32_t payload, F f) {
    const int lane = threadIdx.x & 31;
    int xlo, xhi, ylo, yhi;
    g2pc_unpack_range(range, xlo, xhi, ylo, yhi);
    const int area = (xlo > xhi || ylo > yhi) ? 0 : (xhi - xlo + 1) * (yhi - ylo + 1);
    if (area > 0 && area <= WARP_SMALL_AREA)
        for (int iy = ylo; iy <= yhi; ++iy)
            for (int ix = xlo; ix <= xhi; ++ix) f(ix, iy, lane, payload);
    unsigned big = __ballot_sync(0xffffffffu, area > WARP_SMALL_AREA);
    while (big) {
        const int b = __ffs(big) - 1;
        big &= big - 1u;
        const uint32_t br = __shfl_sync(0xffffffffu, range, b);
        const uint32_t bp = __shfl_sync(0xffffffffu, payload, b);
        int bx, bxh, by, byh;
        g2pc_unpack_range(br, bx, bxh, by, byh);
        const int bw = bxh - bx + 1, bh = byh - by + 1;
        const int sh = bw <= 1 ? 0 : 32 - __clz(bw - 1);  // ceil(log2(width))
        if (sh >= 5) {
            for (int ry = 0; ry < bh; ++ry)
                for (int rx = lane; rx < bw; rx += 32) f(bx + rx, by + ry, b, bp);
        } else {
            const int rx = lane & ((1 << sh) - 1);
            if (rx < bw)
                for (int ry = lane >> sh; ry < bh; ry += 32 >> sh) f(bx + rx, by + ry, b, bp);
        }
    }
}

// rect of a Gaussian (gauss_render.py:182-193): [mean -+ radius] clipped to [0, W-1] x [0, H-1]
__device__ __forceinline__ void gaussian_rect(float mx, float my, float radius, int W, int H, float& x0, float& x1,
                                              float& y0, float& y1) {
    const float wm = (float)W - 1.0f, hm = (float)H - 1.0f;
    x0 = fminf(fmaxf(mx - radius, 0.0f), wm);
    x1 = fminf(fmaxf(mx + radius, 0.0f), wm);
    y0 = fminf(fmaxf(my - radius, 0.0f), hm);
    y1 = fminf(fmaxf(my + radius, 0.0f), hm);
}
