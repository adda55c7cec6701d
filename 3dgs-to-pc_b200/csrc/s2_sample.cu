// s2_sample.cu — S2: Cholesky + on-chip Philox multivariate-normal draw + Mahalanobis cull, two passes.
//
// Reference semantics restated (not copied):
//   gauss_to_pc.py:140-155  sample_from_multivariate_normal  (MultivariateNormal -> Cholesky, mu + L*eps,
//                           +1e-6*I retries on failure; torch multivariate_normal.py:194,251-254)
//   gauss_to_pc.py:92-103   mahalanobis (fp32 inverse + two mat-vecs)
//   gauss_to_pc.py:157-275  create_new_gaussian_points (attempt loop, accept COUNTS decide how many of the
//                           FIRST samples of each Gaussian's block are emitted, :242-258)
//   gauss_to_pc.py:324-369  bin loop: centre points first, then attempt-major / Gaussian-minor samples
//
// Pass 1 (sample_count_kernel): one CTA per tile.  Gathers the tile's Gaussians, factors Sigma, packs a 64-byte
//   record per Gaussian in bin order, simulates the attempt loop and writes per-attempt tile-local prefixes.
// Pass 2 (sample_emit_kernel): one 64-thread CTA per 256 OUTPUT points; runs are expanded into a per-point table in shared
//   memory, points are generated divergence-free (4 per thread, strided), staged in shared memory and leave as
//   16-byte coalesced stores.  RNG is regenerated, never stored.
#include "common.cuh"

namespace {

constexpr int BLOCK = 256;
constexpr unsigned FULL = 0xffffffffu;

struct CountParams {
    const float* xyz;
    const float* cov;
    const void* colours;
    int colour_dtype;
    const float* normals;
    const int32_t* perm;
    const uint32_t* gids;
    int64_t gid_offset;
    int64_t n;
    const g2pc_tile_t* tiles;
    int32_t num_attempts;
    int32_t attempts_stored;
    float std;
    uint32_t k0, k1, call_id;
    float4* records;
    uint32_t* xl;
    uint32_t* tile_totals;
    int32_t* status;
};

// closed-form Cholesky of the lower triangle with the reference's regularise-and-retry ladder
// (+1e-6*I per failed try, at most 3 tries: gauss_to_pc.py:147-155).  Returns the level used (0..2) or 3.
__device__ __forceinline__ int chol3_ladder(float a00, float a10, float a11, float a20, float a21, float a22,
                                            float L[6], float& reg) {
    reg = 0.0f;
#pragma unroll 1
    for (int lvl = 0; lvl < 3; ++lvl) {
        const float d00 = a00 + reg, d11 = a11 + reg, d22 = a22 + reg;
        bool ok = d00 > 0.0f;
        const float l00 = sqrtf(d00);
        const float l10 = a10 / l00;
        const float l20 = a20 / l00;
        const float t11 = d11 - l10 * l10;
        ok = ok && (t11 > 0.0f);
        const float l11 = sqrtf(t11);
        const float l21 = (a21 - l20 * l10) / l11;
        const float t22 = d22 - l20 * l20 - l21 * l21;
        ok = ok && (t22 > 0.0f);
        if (ok) {
            L[0] = l00; L[1] = l10; L[2] = l11; L[3] = l20; L[4] = l21; L[5] = sqrtf(t22);
            return lvl;
        }
        reg = reg + 1e-6f;  // covariances += epsilon * eye(3), cumulatively
    }
    return 3;
}

// general 3x3 inverse (adjugate / determinant), fp32 — stands in for torch.inverse (gauss_to_pc.py:99)
__device__ __forceinline__ void inv3(const float a[9], float inv[9]) {
    const float c00 = a[4] * a[8] - a[5] * a[7];
    const float c01 = a[5] * a[6] - a[3] * a[8];
    const float c02 = a[3] * a[7] - a[4] * a[6];
    const float det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const float id = 1.0f / det;
    inv[0] = c00 * id;
    inv[1] = (a[2] * a[7] - a[1] * a[8]) * id;
    inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    inv[3] = c01 * id;
    inv[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    inv[6] = c02 * id;
    inv[7] = (a[1] * a[6] - a[0] * a[7]) * id;
    inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// block-wide exclusive scan of one uint32 per thread (256 threads); returns the exclusive prefix and the total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wprefix = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) {
        const uint32_t t = s_warp[w];
        if (w < warp) wprefix += t;
        tot += t;
    }
    total = tot;
    __syncthreads();  // s_warp may be reused by the next scan
    return wprefix + inc - v;
}

template <int CULL>
__global__ void __launch_bounds__(BLOCK) sample_count_kernel(const CountParams p) {
    __shared__ uint32_t s_cnt[BLOCK];
    __shared__ uint32_t s_unf[BLOCK];
    __shared__ uint32_t s_warp[BLOCK / 32];

    const g2pc_tile_t t = p.tiles[blockIdx.x];
    const int tid = threadIdx.x;
    const int lpg = t.lpg, k = t.k;
    const int sh = 31 - __clz(lpg);
    const int g = tid >> sh, sub = tid & (lpg - 1);
    const bool active = g < t.count;

    float3 mu = make_float3(0.f, 0.f, 0.f);
    float L[6] = {0, 0, 0, 0, 0, 0};
    float inv[9];
    uint32_t gid = 0;
    int lvl = 0;
    if (active) {
        const int64_t j = (int64_t)t.j0 + g;
        const int64_t idx = p.perm[j];
        gid = p.gids ? p.gids[idx] : (uint32_t)(idx + p.gid_offset);
        mu = make_float3(p.xyz[3 * idx], p.xyz[3 * idx + 1], p.xyz[3 * idx + 2]);
        float a[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) a[c] = p.cov[9 * idx + c];
        float reg;
        lvl = chol3_ladder(a[0], a[3], a[4], a[6], a[7], a[8], L, reg);
        if (CULL == G2PC_CULL_EXPLICIT) {
            a[0] += reg; a[4] += reg; a[8] += reg;
            inv3(a, inv);
        }
        if (sub == 0) {
            float r, gg, b;
            if (p.colour_dtype == G2PC_F64) {
                const double* c = (const double*)p.colours;
                r = (float)c[3 * idx]; gg = (float)c[3 * idx + 1]; b = (float)c[3 * idx + 2];
            } else {
                const float* c = (const float*)p.colours;
                r = c[3 * idx]; gg = c[3 * idx + 1]; b = c[3 * idx + 2];
            }
            float nx = 0.f, ny = 0.f, nz = 0.f;
            if (p.normals) { nx = p.normals[3 * idx]; ny = p.normals[3 * idx + 1]; nz = p.normals[3 * idx + 2]; }
            float4* rec = p.records + 4 * j;
            rec[0] = make_float4(mu.x, mu.y, mu.z, L[0]);
            rec[1] = make_float4(L[1], L[2], L[3], L[4]);
            rec[2] = make_float4(L[5], r, gg, b);
            rec[3] = make_float4(nx, ny, nz, __uint_as_float(gid));
            if (lvl == 3) atomicAdd(&p.status[G2PC_ST_CHOLFAIL], 1);
            else if (lvl > 0) atomicAdd(&p.status[G2PC_ST_CHOLREG], 1);
        }
    }
    if (k <= 0) return;  // centre-only bin: records written, nothing to sample (uniform per CTA)

    // s_unf[g]: Gaussian g of this tile still needs points.  A covariance without a Cholesky factor never samples.
    if (sub == 0) s_unf[g] = (active && lvl < 3) ? 1u : 0u;
    uint32_t added = 0;  // owner role: thread tid tracks Gaussian tid of the tile
    __syncthreads();
    bool own_unf = (tid < t.count) && (s_unf[tid] != 0u);

    const float std_ = p.std;
    for (int a = 0; a < p.num_attempts; ++a) {
        // ---- sampler role: lanes of group g draw samples sub, sub+lpg, ... of Gaussian g ----
        uint32_t c = 0;
        const bool unf = active && (s_unf[g] != 0u);
        if (lpg > 32) s_cnt[tid] = 0;
        if (unf) {
            for (int s = sub; s < k; s += lpg) {
                const float3 e = draw_eps(gid, (uint32_t)s, (uint32_t)a, p.call_id, p.k0, p.k1);
                bool acc;
                if (CULL == G2PC_CULL_EPS_NORM) {
                    const float d = sqrtf(fmaf(e.z, e.z, fmaf(e.y, e.y, e.x * e.x)));
                    acc = d <= std_;
                } else {
                    const float3 x = mvn_point(mu, L[0], L[1], L[2], L[3], L[4], L[5], e);
                    const float dx = mu.x - x.x, dy = mu.y - x.y, dz = mu.z - x.z;
                    const float vx = fmaf(inv[2], dz, fmaf(inv[1], dy, inv[0] * dx));
                    const float vy = fmaf(inv[5], dz, fmaf(inv[4], dy, inv[3] * dx));
                    const float vz = fmaf(inv[8], dz, fmaf(inv[7], dy, inv[6] * dx));
                    const float m = fmaf(dz, vz, fmaf(dy, vy, dx * vx));
                    acc = sqrtf(m) <= std_;  // NaN (m < 0) rejects, like the reference's comparison
                }
                c += acc ? 1u : 0u;
            }
        }
        if (lpg > 32) __syncthreads();  // zeroing of s_cnt visible before the atomics
        if (lpg <= 32) {
            for (int o = lpg >> 1; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
            if (sub == 0) s_cnt[g] = c;
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
            if ((tid & 31) == 0 && c) atomicAdd(&s_cnt[g], c);
        }
        __syncthreads();
        // ---- owner role: accept bookkeeping (gauss_to_pc.py:242,262-263) ----
        uint32_t m = 0;
        if (own_unf) {
            const uint32_t cnt = s_cnt[tid];
            const uint32_t room = (uint32_t)k - added;
            m = cnt < room ? cnt : room;
            added = added + cnt > (uint32_t)k ? (uint32_t)k : added + cnt;
            own_unf = added != (uint32_t)k;
        }
        uint32_t total;
        const uint32_t x = block_excl_scan(m, s_warp, total);
        if (a < p.attempts_stored) {
            if (tid < t.count) p.xl[(int64_t)a * p.n + t.j0 + tid] = x;
            if (tid == 0) p.tile_totals[(int64_t)blockIdx.x * p.attempts_stored + a] = total;
        } else if (total > 0 && tid == 0) {
            p.status[G2PC_ST_OVERFLOW] = 1;
        }
        if (tid < t.count) s_unf[tid] = own_unf ? 1u : 0u;
        if (!__syncthreads_or(own_unf ? 1 : 0)) break;
    }
}

// ---------------------------------------------------------------------------------------------------------
struct EmitParams {
    const float4* records;
    const uint32_t* xl;
    int64_t n;
    const g2pc_unit_t* units;
    const int64_t* unit_base;
    const int32_t* chunk_unit;  // chunk_unit[c] = unit holding output point c*1024 (host-side searchsorted); may be null
    int32_t num_units;
    uint32_t k0, k1, call_id;
    float* out_xyz;
    void* out_rgb;
    void* out_nrm;
    int out_dtype;
    int64_t capacity;
};


// last u in [lo, hi] with unit_base[u] <= p   (unit_base non-decreasing; zero-length units are skipped over)
__device__ __forceinline__ int find_unit(const int64_t* __restrict__ base, int lo, int hi, int64_t p) {
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(base + mid) <= p) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// last i in [0, count) with xl[i] <= q
__device__ __forceinline__ int find_run(const uint32_t* __restrict__ xl, int count, uint32_t q) {
    int lo = 0, hi = count - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(xl + mid) <= q) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <typename OUT_T, int TILE_PTS, int BLOCK>
__device__ __forceinline__ void flush_tile(const float* s, OUT_T* out, int64_t p0, int npts, bool full) {
    if (out == nullptr) return;
    OUT_T* o = out + p0 * 3;
    if constexpr (sizeof(OUT_T) == 4) {
        if (full) {
            const float4* s4 = reinterpret_cast<const float4*>(s);
            float4* o4 = reinterpret_cast<float4*>(o);
            for (int k = threadIdx.x; k < TILE_PTS * 3 / 4; k += BLOCK) o4[k] = s4[k];
            return;
        }
    } else {
        if (full) {
            const float2* s2 = reinterpret_cast<const float2*>(s);
            double2* o2 = reinterpret_cast<double2*>(o);
            for (int k = threadIdx.x; k < TILE_PTS * 3 / 2; k += BLOCK) {
                const float2 v = s2[k];
                o2[k] = make_double2((double)v.x, (double)v.y);
            }
            return;
        }
    }
    for (int i = threadIdx.x; i < npts * 3; i += BLOCK) o[i] = (OUT_T)s[i];
}

constexpr int NU_STAGE = 4;  // units whose descriptors / prefix rows are staged in shared memory per chunk

// last i in [0, count) with xs[i] <= q, xs in shared memory
__device__ __forceinline__ int find_run_smem(const uint32_t* xs, int count, uint32_t q) {
    int lo = 0, hi = count - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (xs[mid] <= q) lo = mid; else hi = mid - 1;
    }
    return lo;
}

constexpr uint32_t CENTRE_TAG = 0xFF000000u;

// Pass 2, three phases per 256-point chunk (ETILE):
//   A  stage the descriptors / bases / prefix rows of the (few) units the chunk overlaps in shared memory;
//   B  expand runs into a per-point (record index, attempt|sample) table: one thread per run for small k (no search
//      at all), one thread per point with a shared-memory binary search for large k;
//   C  uniform, divergence-free generation: thread t handles points t, t+64, t+128, t+192 (independent record
//      loads in flight, conflict-free staging), then 16-byte coalesced stores.
constexpr int EPTS = 4;           // points per thread
constexpr int ETILE = 64 * EPTS;  // output points per emit CTA
constexpr int EB = 64;            // threads per emit CTA (2 warps: barriers are cheap, many CTAs overlap per SM)
constexpr int MAX_TILE_G = 256;   // Gaussians per tile (rows of xl staged per unit)

template <typename OUT_T, bool HAS_NRM>
__global__ void __launch_bounds__(EB, 16) sample_emit_kernel(const EmitParams p) {
    __shared__ __align__(16) float s_xyz[ETILE * 3];
    __shared__ __align__(16) float s_rgb[ETILE * 3];
    __shared__ __align__(16) float s_nrm[HAS_NRM ? ETILE * 3 : 4];
    // the per-point table lives in the xyz staging area: it is dead before the positions are written (barrier below)
    uint32_t* s_j = reinterpret_cast<uint32_t*>(s_xyz);
    uint32_t* s_s = s_j + ETILE;
    __shared__ int s_urange[2];
    __shared__ int s_nlong;
    __shared__ int4 s_long[ETILE / 8];  // runs covering more than 8 points of the chunk (at most TILE_PTS/9 of them)
    __shared__ int s_urel[NU_STAGE + 1];
    __shared__ g2pc_unit_t s_units[NU_STAGE];
    __shared__ uint32_t s_xl[NU_STAGE][MAX_TILE_G];

    const int64_t total = __ldg(p.unit_base + p.num_units);
    const int64_t p0 = (int64_t)blockIdx.x * ETILE;
    if (p0 >= total) return;
    const int64_t pend = (p0 + ETILE < total) ? p0 + ETILE : total;
    const int npts = (int)(pend - p0);
    const int tid = threadIdx.x;

    int u_lo, u_hi;
    if (p.chunk_unit) {
        // chunk c+1 starts at pend (or beyond the total): its unit bounds this chunk's range from above
        u_lo = p.chunk_unit[blockIdx.x];
        u_hi = min(p.chunk_unit[blockIdx.x + 1], p.num_units - 1);
    } else {
        if (tid < 2) {
            const int64_t q = tid == 0 ? p0 : pend - 1;
            s_urange[tid] = find_unit(p.unit_base, 0, p.num_units - 1, q);
        }
        __syncthreads();
        u_lo = s_urange[0]; u_hi = s_urange[1];
    }
    const int nu = u_hi - u_lo + 1;
    if (nu <= NU_STAGE) {  // uniform per CTA
        // ---- phase A ----  (unit bases become 32-bit offsets relative to the chunk start, clamped to the chunk)
        if (tid == 0) s_nlong = 0;
        if (tid <= nu) {
            const int64_t b = __ldg(p.unit_base + u_lo + tid) - p0;
            s_urel[tid] = (int)(b < -0x3fffffff ? -0x3fffffff : (b > 0x3fffffff ? 0x3fffffff : b));
        }
        if (tid < nu) s_units[tid] = p.units[u_lo + tid];
        __syncthreads();
        for (int ui = 0; ui < nu; ++ui) {
            const g2pc_unit_t un = s_units[ui];
            if (un.attempt >= 0)
                for (int e = tid; e < un.count; e += EB)
                    s_xl[ui][e] = __ldg(p.xl + (int64_t)un.attempt * p.n + un.j0 + e);
        }
        __syncthreads();
        // ---- phase B ----
        for (int ui = 0; ui < nu; ++ui) {
            const g2pc_unit_t un = s_units[ui];
            const int lb = s_urel[ui], le = s_urel[ui + 1];  // unit's point range relative to p0 (may be negative)
            const int cb = lb > 0 ? lb : 0, ce = le < npts ? le : npts;
            if (ce <= cb) continue;
            if (un.attempt < 0) {
                for (int l = cb + tid; l < ce; l += EB) {
                    s_j[l] = (uint32_t)(un.j0 + (l - lb));
                    s_s[l] = CENTRE_TAG;
                }
            } else for (int ri = tid; ri < un.count; ri += EB) {
                // one thread per run: short overlaps are written by the thread itself, long ones are queued and
                // filled by whole warps (lanes = consecutive points) after the barrier.  lb > -2^30 whenever the unit
                // overlaps the chunk and a unit holds < 2^30 points, so the 32-bit sums below cannot overflow.
                const int rb = lb + (int)s_xl[ui][ri];
                const int re = (ri + 1 < un.count) ? lb + (int)s_xl[ui][ri + 1] : le;
                if (rb >= ce) break;   // runs are ordered: nothing further overlaps the chunk
                const int b = rb > cb ? rb : cb, e = re < ce ? re : ce;
                const uint32_t tag = (uint32_t)un.attempt << 24;
                if (e - b > 8) {
                    const int slot = atomicAdd(&s_nlong, 1);
                    s_long[slot] = make_int4(b, e, un.j0 + ri, (int)(tag | (uint32_t)(b - rb)));
                } else {
                    for (int l = b; l < e; ++l) {
                        s_j[l] = (uint32_t)(un.j0 + ri);
                        s_s[l] = tag | (uint32_t)(l - rb);
                    }
                }
            }
        }
        __syncthreads();
        {
            const int nlong = s_nlong;
            const int lane = tid & 31, warp = tid >> 5;
            for (int r = warp; r < nlong; r += EB / 32) {
                const int4 d = s_long[r];  // (begin, end, record, tag | first sample)
                for (int l = d.x + lane; l < d.y; l += 32) {
                    s_j[l] = (uint32_t)d.z;
                    s_s[l] = (uint32_t)d.w + (uint32_t)(l - d.x);
                }
            }
        }
    } else {
        // many tiny units in one chunk (late attempts of small-k bins): per-point search in global memory
        for (int l = tid; l < npts; l += EB) {
            const int64_t pt = p0 + l;
            const int u = find_unit(p.unit_base, u_lo, u_hi, pt);
            const g2pc_unit_t un = p.units[u];
            const uint32_t q = (uint32_t)(pt - __ldg(p.unit_base + u));
            if (un.attempt < 0) {
                s_j[l] = (uint32_t)(un.j0 + q);
                s_s[l] = CENTRE_TAG;
            } else {
                const uint32_t* xrow = p.xl + (int64_t)un.attempt * p.n + un.j0;
                const int i = find_run(xrow, un.count, q);
                s_j[l] = (uint32_t)(un.j0 + i);
                s_s[l] = ((uint32_t)un.attempt << 24) | (q - __ldg(xrow + i));
            }
        }
    }
    __syncthreads();
    // ---- phase C ----
    float4 r0[EPTS], r1[EPTS];
    float l22[EPTS];
    uint32_t gidv[EPTS], ss[EPTS];
#pragma unroll
    for (int r = 0; r < EPTS; ++r) {
        const int l = tid + r * EB;
        if (l < npts) {
            const float4* rec = p.records + 4 * (int64_t)s_j[l];
            ss[r] = s_s[l];
            r0[r] = __ldg(rec); r1[r] = __ldg(rec + 1);
            const float4 r2 = __ldg(rec + 2), r3 = __ldg(rec + 3);
            l22[r] = r2.x; gidv[r] = __float_as_uint(r3.w);
            s_rgb[3 * l] = r2.y; s_rgb[3 * l + 1] = r2.z; s_rgb[3 * l + 2] = r2.w;
            if (HAS_NRM) { s_nrm[3 * l] = r3.x; s_nrm[3 * l + 1] = r3.y; s_nrm[3 * l + 2] = r3.z; }
        }
    }
    __syncthreads();  // every thread has read its table entries: the xyz staging area may be overwritten
#pragma unroll
    for (int r = 0; r < EPTS; ++r) {
        const int l = tid + r * EB;
        if (l < npts) {
            float3 x = make_float3(r0[r].x, r0[r].y, r0[r].z);
            if (ss[r] != CENTRE_TAG) {
                const float3 e = draw_eps(gidv[r], ss[r] & 0x00FFFFFFu, ss[r] >> 24, p.call_id, p.k0, p.k1);
                x = mvn_point(x, r0[r].w, r1[r].x, r1[r].y, r1[r].z, r1[r].w, l22[r], e);
            }
            s_xyz[3 * l] = x.x; s_xyz[3 * l + 1] = x.y; s_xyz[3 * l + 2] = x.z;
        }
    }
    __syncthreads();
    if (npts == ETILE && sizeof(OUT_T) == 4) {
        // full chunk, f32 outputs: one loop of 16-byte stores over the three staging areas (they are contiguous)
        constexpr int V = ETILE * 3 / 4;
        const float4* sx = reinterpret_cast<const float4*>(s_xyz);
        const float4* sr = reinterpret_cast<const float4*>(s_rgb);
        const float4* sn = reinterpret_cast<const float4*>(s_nrm);
        float4* ox = reinterpret_cast<float4*>(p.out_xyz + p0 * 3);
        float4* orgb = reinterpret_cast<float4*>((float*)p.out_rgb + p0 * 3);
        float4* on = HAS_NRM ? reinterpret_cast<float4*>((float*)p.out_nrm + p0 * 3) : nullptr;
#pragma unroll
        for (int k = tid; k < V; k += EB) {
            ox[k] = sx[k];
            orgb[k] = sr[k];
            if (HAS_NRM) on[k] = sn[k];
        }
    } else {
        const bool full = npts == ETILE;
        flush_tile<float, ETILE, EB>(s_xyz, p.out_xyz, p0, npts, full);
        flush_tile<OUT_T, ETILE, EB>(s_rgb, (OUT_T*)p.out_rgb, p0, npts, full);
        if (HAS_NRM) flush_tile<OUT_T, ETILE, EB>(s_nrm, (OUT_T*)p.out_nrm, p0, npts, full);
    }
}

__global__ void dump_eps_kernel(const int64_t* __restrict__ gids, int64_t n_gids, int32_t k, int32_t attempt,
                                uint32_t k0, uint32_t k1, uint32_t call_id, float* __restrict__ eps) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_gids * (int64_t)k) return;
    const int64_t s = t / n_gids, i = t - s * n_gids;  // eps[s, i, :]
    const float3 e = draw_eps((uint32_t)gids[i], (uint32_t)s, (uint32_t)attempt, call_id, k0, k1);
    eps[3 * t] = e.x; eps[3 * t + 1] = e.y; eps[3 * t + 2] = e.z;
}

}  // namespace

extern "C" int g2pc_sample_count(const float* xyz, const float* cov, const void* colours, int colour_dtype,
                                 const float* normals, const int32_t* perm, const uint32_t* gids, int64_t gid_offset,
                                 int64_t n,
                                 const g2pc_tile_t* tiles, int32_t num_tiles, int32_t num_attempts,
                                 int32_t attempts_stored, float mahalanobis_std, int32_t cull_mode, uint64_t seed,
                                 uint32_t call_id, void* records, uint32_t* xl, uint32_t* tile_totals,
                                 int32_t* status, void* stream) {
    G2PC_CHECK_ARG(n >= 0 && num_tiles >= 0, "negative size");
    if (n == 0 || num_tiles == 0) return G2PC_OK;
    G2PC_CHECK_ARG(xyz && cov && colours && perm && tiles && records && xl && tile_totals && status, "null pointer");
    G2PC_CHECK_ARG(colour_dtype == G2PC_F32 || colour_dtype == G2PC_F64, "bad colour_dtype");
    G2PC_CHECK_ARG(num_attempts >= 1 && attempts_stored >= 1 && attempts_stored <= num_attempts,
                   "need 1 <= attempts_stored <= num_attempts");
    // the emit pass tags every sample with (attempt << 24 | sample); tag 0xFF is the centre-point marker
    G2PC_CHECK_ARG(attempts_stored <= 255, "attempts_stored must be <= 255 (8-bit attempt tag, 0xFF reserved)");
    G2PC_CHECK_ARG(cull_mode == G2PC_CULL_EPS_NORM || cull_mode == G2PC_CULL_EXPLICIT, "bad cull_mode");
    G2PC_CHECK_ARG(((uintptr_t)records & 15) == 0, "records must be 16-byte aligned");
    CountParams p;
    p.xyz = xyz; p.cov = cov; p.colours = colours; p.colour_dtype = colour_dtype; p.normals = normals;
    p.perm = perm; p.gids = gids; p.gid_offset = gid_offset; p.n = n; p.tiles = tiles; p.num_attempts = num_attempts;
    p.attempts_stored = attempts_stored; p.std = mahalanobis_std;
    p.k0 = (uint32_t)seed; p.k1 = (uint32_t)(seed >> 32); p.call_id = call_id;
    p.records = (float4*)records; p.xl = xl; p.tile_totals = tile_totals; p.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    if (cull_mode == G2PC_CULL_EPS_NORM)
        sample_count_kernel<G2PC_CULL_EPS_NORM><<<(unsigned)num_tiles, BLOCK, 0, st>>>(p);
    else
        sample_count_kernel<G2PC_CULL_EXPLICIT><<<(unsigned)num_tiles, BLOCK, 0, st>>>(p);
    G2PC_CHECK_LAUNCH();
    return G2PC_OK;
}

extern "C" int g2pc_sample_emit_chunk_points(void) { return ETILE; }

extern "C" int g2pc_sample_emit(const void* records, const uint32_t* xl, int64_t n, const g2pc_unit_t* units,
                                const int64_t* unit_base, const int32_t* chunk_unit, int32_t num_units, uint64_t seed,
                                uint32_t call_id,
                                float* out_xyz, void* out_rgb, void* out_nrm, int out_dtype, int64_t capacity,
                                void* stream) {
    G2PC_CHECK_ARG(capacity >= 0 && num_units >= 0, "negative size");
    if (capacity == 0 || num_units == 0) return G2PC_OK;
    G2PC_CHECK_ARG(records && xl && units && unit_base && out_xyz && out_rgb, "null pointer");
    G2PC_CHECK_ARG(out_dtype == G2PC_F32 || out_dtype == G2PC_F64, "bad out_dtype");
    G2PC_CHECK_ARG((((uintptr_t)out_xyz | (uintptr_t)out_rgb | (uintptr_t)out_nrm) & 15) == 0,
                   "outputs must be 16-byte aligned");
    EmitParams p;
    p.records = (const float4*)records; p.xl = xl; p.n = n; p.units = units; p.unit_base = unit_base;
    p.chunk_unit = chunk_unit; p.num_units = num_units; p.k0 = (uint32_t)seed; p.k1 = (uint32_t)(seed >> 32); p.call_id = call_id;
    p.out_xyz = out_xyz; p.out_rgb = out_rgb; p.out_nrm = out_nrm; p.out_dtype = out_dtype; p.capacity = capacity;
    const unsigned grid = (unsigned)((capacity + ETILE - 1) / ETILE);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == G2PC_F32) {
        if (out_nrm) sample_emit_kernel<float, true><<<grid, EB, 0, st>>>(p);
        else sample_emit_kernel<float, false><<<grid, EB, 0, st>>>(p);
    } else {
        if (out_nrm) sample_emit_kernel<double, true><<<grid, EB, 0, st>>>(p);
        else sample_emit_kernel<double, false><<<grid, EB, 0, st>>>(p);
    }
    G2PC_CHECK_LAUNCH();
    return G2PC_OK;
}

extern "C" int g2pc_dump_eps(const int64_t* gids, int64_t n_gids, int32_t k, int32_t attempt, uint64_t seed,
                             uint32_t call_id, float* eps, void* stream) {
    G2PC_CHECK_ARG(n_gids >= 0 && k >= 0, "negative size");
    if (n_gids == 0 || k == 0) return G2PC_OK;
    G2PC_CHECK_ARG(gids && eps, "null pointer");
    const int64_t tot = n_gids * (int64_t)k;
    dump_eps_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        gids, n_gids, k, attempt, (uint32_t)seed, (uint32_t)(seed >> 32), call_id, eps);
    G2PC_CHECK_LAUNCH();
    return G2PC_OK;
}
