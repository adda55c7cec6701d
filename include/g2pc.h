/*
 * g2pc.h — C ABI of libg2pc.so: B200-native (sm_100a) kernels for the 3DGS-to-PC hot path
 * (per-Gaussian point sampling + Mahalanobis cull, per-camera colour / visibility rasterisation).
 *
 * This is the drop-in boundary.  The reference has two native/op boundaries on this path:
 *   - pybind11 module `gaussian_pointcloud_rasterization._C`
 *       (gaussian-pointcloud-rasterization/ext.cpp:15-17, rasterize_points.cu:36-145, rasterize_points.h:18-46)
 *   - pure-torch call chains for the sampler
 *       (gauss_to_pc.py:140-275, gauss_handler.py:26-63)
 * Both are replaced by the plain-C entry points below.  Conventions (all entry points):
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`; the caller owns all memory,
 *     including scratch (sizes come from the *_bytes query functions or are stated in the comment);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no allocation, no host
 *     synchronisation and no global mutable state inside the library;
 *   - return value: 0 = G2PC_OK, otherwise an error code; g2pc_last_error() gives a thread-local message;
 *   - no C++ exception crosses the ABI.
 */
#ifndef G2PC_H
#define G2PC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2PC_OK 0
#define G2PC_ERR_INVALID 1   /* bad argument */
#define G2PC_ERR_CUDA 2      /* a CUDA runtime call / launch failed */
#define G2PC_ERR_WORKSPACE 3 /* caller-provided scratch too small */

#define G2PC_F32 0
#define G2PC_F64 1

/* ---- misc ------------------------------------------------------------------------------------ */
int g2pc_version(void);
const char* g2pc_last_error(void);

/* ---- S1: covariance build ------------------------------------------------------------------- */
/* Replaces build_rotation / build_scaling_rotation / build_covariance_from_scaling_rotation
 * (gauss_handler.py:26-63): R(q) without re-normalisation, L = R*diag(exp(mod*s)), Sigma = L*L^T.
 * scales (n,3) log-space and rots (n,4) in `in_dtype` (the ply loader feeds f64, gauss_dataloader.py:66-80;
 * the elements of R and exp(s) are formed in the input precision and then rounded to f32, as the
 * reference's slice-assignments into float tensors do).  cov: (n,3,3) f32 row-major. */
int g2pc_cov_build(const void* scales, const void* rots, int in_dtype, float scale_modifier,
                   int64_t n, float* cov, void* stream);

/* Replaces Gaussians.calculate_normals (gauss_handler.py:89-106): column argmin(scale) of R(q).
 * normals: (n,3) f32. */
int g2pc_normals(const void* scales, const void* rots, int in_dtype, int64_t n, float* normals,
                 void* stream);

/* Replaces torch.linalg.eigvals(covariances).real on the N x 3 x 3 covariance batch
 * (gauss_handler.py:112 non_posdef_covariances, :259 get_gaussian_magnitudes): eigenvalues of the symmetric part,
 * closed form in f64, rounded to f32, ascending.  eigvals: (n,3) f32.  The callers keep the reference's
 * elementwise f32 chains (ellipsoid area, <= epsilon tests) in torch. */
int g2pc_eigvals_sym3(const float* cov, int64_t n, float* eigvals, void* stream);

/* ---- S2: sampling + Mahalanobis cull --------------------------------------------------------- */
/* Replaces sample_from_multivariate_normal + mahalanobis + create_new_gaussian_points +
 * the bin loop of generate_pointcloud (gauss_to_pc.py:92-103,140-371) and
 * torch.distributions.MultivariateNormal (Cholesky + loc + L*eps).
 *
 * Work is described by host-built tables (the reference builds the same bins on the host,
 * gauss_to_pc.py:308-343):
 *   tile  = <=256/lpg consecutive Gaussians (in bin order) of one bin, all drawing k samples per attempt
 *   unit  = one contiguous span of the output, in the reference's output order:
 *           the centre points of a bin, or the samples one tile emits in one attempt.            */
typedef struct {
    int32_t j0;    /* first Gaussian (bin-order index) */
    int32_t count; /* Gaussians in the tile (<= 256/lpg) */
    int32_t k;     /* samples drawn per Gaussian per attempt = bin's points-per-Gaussian - 1 */
    int32_t lpg;   /* threads cooperating on one Gaussian: power of two, 1..256 */
} g2pc_tile_t;

typedef struct {
    int32_t attempt; /* >=0: sample unit of that attempt; -1: centre-point unit */
    int32_t j0;      /* first Gaussian (bin-order index) */
    int32_t count;   /* Gaussians covered */
    int32_t k;       /* samples per Gaussian per attempt (0 for centre units) */
} g2pc_unit_t;

#define G2PC_CULL_EPS_NORM 0 /* accept iff |eps| <= std (exact-arithmetic identity of the reference test) */
#define G2PC_CULL_EXPLICIT 1 /* accept iff sqrt(d^T Sigma^-1 d) <= std, d = mu - x, fp32 (gauss_to_pc.py:92-103) */

/* status words written by g2pc_sample_count (int32 each) */
#define G2PC_ST_OVERFLOW 0  /* !=0: some Gaussian emitted in an attempt >= attempts_stored (re-run with more) */
#define G2PC_ST_CHOLFAIL 1  /* number of Gaussians whose covariance had no Cholesky factor even after +2e-6*I */
#define G2PC_ST_CHOLREG 2   /* number of Gaussians that needed +1e-6*I or +2e-6*I (gauss_to_pc.py:147-155) */
#define G2PC_ST_WORDS 4

/* Pass 1.  For every tile: gather the tile's Gaussians through `perm`, factor Sigma (closed-form
 * Cholesky), write one 64-byte record per Gaussian in bin order, then simulate the attempt loop of
 * create_new_gaussian_points: draw k samples per unfinished Gaussian with Philox4x32-10
 * (key = seed, counter = (gid, sample, attempt, call_id)) + Box-Muller, count the accepted ones,
 * m = min(k - added, count), and write the tile-local exclusive prefix of m
 * (xl[attempt*n + j]) and the tile total (tile_totals[tile*attempts_stored + attempt]).
 *   xyz (N,3) f32 · cov (N,3,3) f32 · colours (N,3) colour_dtype · normals (N,3) f32 or NULL
 *   perm (n,) int32: bin-order index -> row of the input arrays;  the RNG is keyed by the global Gaussian id
 *   gids[row] (uint32 per input row; survives culls and sharding) or, if gids is NULL, row + gid_offset
 *   records: n*64 bytes, 16-byte aligned · xl: attempts_stored*n uint32 · tile_totals: num_tiles*attempts_stored
 *   uint32, MUST be zero-filled by the caller · status: G2PC_ST_WORDS int32, zero-filled by the caller. */
int g2pc_sample_count(const float* xyz, const float* cov, const void* colours, int colour_dtype,
                      const float* normals, const int32_t* perm, const uint32_t* gids, int64_t gid_offset, int64_t n,
                      const g2pc_tile_t* tiles, int32_t num_tiles, int32_t num_attempts,
                      int32_t attempts_stored, float mahalanobis_std, int32_t cull_mode, uint64_t seed,
                      uint32_t call_id, void* records, uint32_t* xl, uint32_t* tile_totals,
                      int32_t* status, void* stream);

/* Pass 2.  Load-balanced expansion over OUTPUT points: point p belongs to unit u (unit_base[u] <= p <
 * unit_base[u+1], unit_base = exclusive prefix sum of the unit lengths, num_units+1 int64 entries, built by the
 * caller from tile_totals); inside a sample unit the Gaussian and sample index come from a search of xl;
 * eps is regenerated from the counter and x = mu + L*eps written ("first m samples of the block",
 * gauss_to_pc.py:247-258).  Outputs: out_xyz (capacity,3) f32; out_rgb / out_nrm (capacity,3) in out_dtype
 * (out_nrm may be NULL).  The launch covers `capacity` points in chunks of C = g2pc_sample_emit_chunk_points();
 * threads beyond unit_base[num_units] exit.  chunk_unit (optional, ceil(capacity/C)+1 int32): index of the unit holding
 * output point c*C (last u with unit_base[u] <= c*C, clamped to num_units - 1) — saves the per-CTA search. */
int g2pc_sample_emit_chunk_points(void); /* output points per CTA of g2pc_sample_emit (chunk size of chunk_unit) */
int g2pc_sample_emit(const void* records, const uint32_t* xl, int64_t n, const g2pc_unit_t* units,
                     const int64_t* unit_base, const int32_t* chunk_unit, int32_t num_units, uint64_t seed,
                     uint32_t call_id,
                     float* out_xyz, void* out_rgb, void* out_nrm, int out_dtype, int64_t capacity,
                     void* stream);

/* The standard-normal draws the sampler uses: eps[s, i, :] for sample s < k of Gaussian gids[i] in `attempt`
 * — the (k, n', 3) tensor torch's MultivariateNormal.rsample would have drawn (gauss_to_pc.py:149).
 * Used to inject the kernel's random stream into the reference/oracle for parity tests. */
int g2pc_dump_eps(const int64_t* gids, int64_t n_gids, int32_t k, int32_t attempt, uint64_t seed,
                  uint32_t call_id, float* eps, void* stream);

/* ---- S3-S6: colour stage, renderer_type=python semantics (gauss_render.py:101-465) ------------------------------ */
/* Replaces GaussPythonRenderer.__call__/render (gauss_render.py:266-465) and — as the native op boundary — the role
 * of _C.rasterize_gaussians (rasterize_points.cu:36-145) in the per-camera loop of gauss_to_pc.py:437-454.
 * One camera = preprocess -> depth_order -> build_tree -> [host reads the 32-byte header] -> emit_instances ->
 * sort_instances -> blend -> accumulate (-> compose_image). */
typedef struct {
    float view[16];  /* world_view_transform, row-vector convention p_view = [p,1] * V (camera_handler.py:46), row-major */
    float proj[16];  /* projection_matrix as stored by Camera (already transposed, camera_handler.py:48), row-major */
    float campos[3]; /* camera centre (SH view directions) */
    float tan_fovx, tan_fovy, focal_x, focal_y;
    int32_t width, height;
} g2pc_camera_t;

typedef struct {
    int32_t r0, c0, w, h;    /* first row / column and size of the leaf tile in pixels */
    int32_t inst_begin;      /* offset of the leaf's list in the sorted instance arrays */
    int32_t inst_count;      /* Gaussians whose rect overlaps the leaf */
    int32_t pix_offset;      /* offset of the leaf's pixels in the concatenated leaf-colour buffer */
    int32_t node;            /* index of the quadtree node */
} g2pc_leaf_t;

#define G2PC_MAX_LEVELS 12
/* header words written by g2pc_build_tree (int32 each) */
#define G2PC_HDR_NUM_LEAVES 0
#define G2PC_HDR_TOTAL_INST 1    /* sum of the leaves' instance counts */
#define G2PC_HDR_TOTAL_PIX 2
#define G2PC_HDR_NEED_DEEPER 3   /* a tile at the deepest tabulated level still has to split: re-run with more levels */
#define G2PC_HDR_LEAF_OVERFLOW 4 /* more leaves than max_leaves */
#define G2PC_HDR_TOTAL_UPPER 5   /* instance slots emit_instances fills (>= TOTAL_INST; the rest is padding) */
#define G2PC_HDR_WORDS 8

/* Quadtree tables (host-built, g2pc/quadtree.py): `tables` = 6 int32 arrays of n1 = 2^num_levels - 1 entries each,
 * concatenated: x start, x end (inclusive), x flags, y start, y end, y flags; level l at offset 2^l - 1.
 * 2-D node index = (4^l - 1)/3 + iy * 2^l + ix. */

/* S3.  Per Gaussian: projection (projection_ndc, gauss_render.py:151-168), EWA covariance (build_covariance_2d
 * :101-148), radius / rect (:171-193), conic = inverse(cov2d) (:349) pre-scaled by -0.5*log2(e), colour (given, or SH
 * deg <= 3 evaluated towards the camera: eval_sh :43-99 + 0.5, clamped at 0), and tile-membership counting for every
 * tabulated level.  xyz (n,3) · cov (n,3,3) · opacity (n) · colours (n,3) f32 or NULL · shs (n,3,sh_stride) f32 channel-
 * major or NULL.  proj: n x 48 bytes (3 float4: {mx,my,c00',c01'} {c11',log2(opacity),r,g} {b,depth,radius,valid}).
 * node_cnt: one uint32 per 2-D node, zero-filled by the caller.  depth_key (n) uint32: bits(-z_view), 0xFFFFFFFF
 * if behind the camera.  touched (n) uint32: leaf-candidate nodes overlapped (upper bound of the instances). */
int g2pc_preprocess(const float* xyz, const float* cov, const float* opacity, const float* colours,
                    const float* shs, int32_t sh_stride, int32_t sh_degree, int64_t n,
                    const g2pc_camera_t* cam_host, const int32_t* tables, int32_t num_levels,
                    int32_t max_gaussians_per_tile, void* proj, uint32_t* node_cnt, uint32_t* depth_key,
                    uint32_t* touched, void* stream);

/* S4a.  order[k] = index of the k-th nearest Gaussian (stable radix sort of depth_key: ties keep index order, the
 * reference's torch.sort is unstable there, gauss_render.py:340-344); incl[k] = inclusive prefix sum of
 * touched[order[k]].  cub::DeviceRadixSort + cub::DeviceScan. */
int64_t g2pc_depth_order_workspace_bytes(int64_t n);
int g2pc_depth_order(const uint32_t* depth_key, const uint32_t* touched, int64_t n, uint32_t* order, uint32_t* incl,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* S4b.  Resolve the quadtree (one CTA): node states, leaves in the reference's BFS order with instance / pixel
 * offsets, seg_begin[leaf] (max_leaves + 1 entries), leaf_order (heaviest leaf first, the blend's launch order),
 * header.  node_state: uint8 per node · leaf_of_node: int32 per node.  incl/n: from g2pc_depth_order (may be NULL/0). */
int g2pc_build_tree(const int32_t* tables, int32_t num_levels, int32_t max_gaussians_per_tile,
                    const uint32_t* node_cnt, const uint32_t* incl, int64_t n, uint8_t* node_state,
                    int32_t* leaf_of_node, g2pc_leaf_t* leaves, int32_t* seg_begin, int32_t* leaf_order,
                    int32_t max_leaves, int32_t* header, void* stream);

/* S4c.  In depth order: Gaussian order[k] writes (leaf id | 0xFFFFFFFF, Gaussian id) for every leaf-candidate node it
 * overlaps at slots [incl[k] - touched, incl[k]).  level_mask: bit l set iff level l has nodes that are not forced to
 * split by their size (levels without leaf candidates are skipped). */
int g2pc_emit_instances(const void* proj, const uint32_t* order, const uint32_t* incl, const uint32_t* touched,
                        int64_t n, int32_t width, int32_t height, const int32_t* tables, int32_t num_levels,
                        uint32_t level_mask, const uint8_t* node_state, const int32_t* leaf_of_node,
                        uint32_t* inst_leaf, uint32_t* inst_gid, void* stream);

/* S4d.  Stable radix sort of the instance pairs on the low leaf_bits bits of the leaf id (cub::DeviceRadixSort): the
 * leaves' lists become contiguous ([seg_begin[l], seg_begin[l+1])) and stay depth-ordered.
 * *sorted_in_alt_host = 1 if the result ended in the *_alt buffers.  (The only entry point that writes a host word.) */
int64_t g2pc_sort_instances_workspace_bytes(int64_t num_items);
int g2pc_sort_instances(uint32_t* inst_leaf, uint32_t* inst_leaf_alt, uint32_t* inst_gid, uint32_t* inst_gid_alt,
                        int64_t num_items, int32_t leaf_bits, void* workspace, int64_t workspace_bytes,
                        int32_t* sorted_in_alt_host, void* stream);

/* S5.  Front-to-back blend of every leaf (gauss_render.py:337-369) + per-Gaussian maximum contribution / arg-max pixel
 * (:371-385) published as cam_best[g] = max((bits(contribution) << 32) | ~leaf_pixel_index).
 * max_leaf_pixels_quads: upper bound of ceil(w/4)*h over the leaves.  max_contrib (n) f32: the running maxima of the
 * earlier cameras (read-only here; contributions that cannot beat them skip the bookkeeping).  leaf_colour:
 * (total_pix,3) f32.
 * owner: uint32 per image pixel (zero-filled): 1 + index of the last leaf pixel covering it.
 * work_counter: one int32, zero-filled by the caller (persistent CTAs pull (leaf, slab) items, heaviest leaf first). */
int g2pc_blend(const g2pc_leaf_t* leaves, const int32_t* leaf_order, int32_t num_leaves, int32_t max_leaf_pixels_quads,
               const uint32_t* inst_gid, const void* proj, uint64_t* cam_best, const float* max_contrib,
               float* leaf_colour, uint32_t* owner, int32_t width, int32_t height, float background,
               int32_t* work_counter, void* stream);

/* S6.  Fold one camera into the per-Gaussian accumulators (gauss_render.py:387-395; the role of
 * GaussianRasterizer.update_max_contributions, gaussian_pointcloud_rasterization/__init__.py:142-152):
 * where the camera's best contribution beats max_contrib[g] (strict >) store it and the blended colour of the winning
 * pixel.  Clears cam_best. */
int g2pc_accumulate(uint64_t* cam_best, const float* leaf_colour, int64_t n, float* max_contrib, float* colours,
                    void* stream);

/* Rendered image (H,W,3) f32, flipped left-right like the reference (gauss_render.py:402); clears `owner`. */
int g2pc_compose_image(uint32_t* owner, const float* leaf_colour, int32_t width, int32_t height, float background,
                       float* image, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* G2PC_H */
