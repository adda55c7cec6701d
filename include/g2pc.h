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
 *   perm (n,) int32: bin-order index -> row of the input arrays;  gid_offset: added to perm[j] to form the
 *   global Gaussian id that keys the RNG (rank offset when the Gaussian array is sharded)
 *   records: n*64 bytes, 16-byte aligned · xl: attempts_stored*n uint32 · tile_totals: num_tiles*attempts_stored
 *   uint32, MUST be zero-filled by the caller · status: G2PC_ST_WORDS int32, zero-filled by the caller. */
int g2pc_sample_count(const float* xyz, const float* cov, const void* colours, int colour_dtype,
                      const float* normals, const int32_t* perm, int64_t gid_offset, int64_t n,
                      const g2pc_tile_t* tiles, int32_t num_tiles, int32_t num_attempts,
                      int32_t attempts_stored, float mahalanobis_std, int32_t cull_mode, uint64_t seed,
                      uint32_t call_id, void* records, uint32_t* xl, uint32_t* tile_totals,
                      int32_t* status, void* stream);

/* Pass 2.  Load-balanced expansion over OUTPUT points: point p belongs to unit u (unit_base[u] <= p <
 * unit_base[u+1], unit_base = exclusive prefix sum of the unit lengths, num_units+1 int64 entries, built by the
 * caller from tile_totals); inside a sample unit the Gaussian and sample index come from a search of xl;
 * eps is regenerated from the counter and x = mu + L*eps written ("first m samples of the block",
 * gauss_to_pc.py:247-258).  Outputs: out_xyz (capacity,3) f32; out_rgb / out_nrm (capacity,3) in out_dtype
 * (out_nrm may be NULL).  The launch covers `capacity` points; threads beyond unit_base[num_units] exit. */
int g2pc_sample_emit(const void* records, const uint32_t* xl, int64_t n, const g2pc_unit_t* units,
                     const int64_t* unit_base, int32_t num_units, uint64_t seed, uint32_t call_id,
                     float* out_xyz, void* out_rgb, void* out_nrm, int out_dtype, int64_t capacity,
                     void* stream);

/* The standard-normal draws the sampler uses: eps[s, i, :] for sample s < k of Gaussian gids[i] in `attempt`
 * — the (k, n', 3) tensor torch's MultivariateNormal.rsample would have drawn (gauss_to_pc.py:149).
 * Used to inject the kernel's random stream into the reference/oracle for parity tests. */
int g2pc_dump_eps(const int64_t* gids, int64_t n_gids, int32_t k, int32_t attempt, uint64_t seed,
                  uint32_t call_id, float* eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* G2PC_H */
