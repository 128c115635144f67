// Pair-count kernels.
//
// Replaces scipy's cKDTree.count_neighbors as called from AngularTree.count
// (reference src/yaw/catalog/trees.py:348-353) and the per-z-bin loop of
// process_patch_pair (src/yaw/correlation/measurements.py:109-124).
//
// k_count_fast  -- the production kernel.  One warp owns a register tile of YAWB_TILE
//   second-catalog points (YAWB_RPL per lane).  For every z-bin it gathers the
//   first-catalog points of the linked patch that fall into the tile's bounding box
//   grown by the bin's search radius (sky-cell rows -> contiguous runs -> per-point
//   cull), rotates them into the tile-local frame in FP64, rounds ONCE to float and
//   stages them as float4 (-2x, -2y, -2z, |s|^2 - mid) in shared memory.  The pair test is
//       u = (|r|^2 + |s|^2 - mid) - 2 r.s = d2 - mid      4 FP32 ops (FADD + 3 FFMA)
//       in  = |u| < h - eps,   maybe = |u| < h + eps      2 FSETP + 2 predicated FADD
//   on the CUDA cores; eps bounds the FP32 error of u.  Whenever a lane's counts of
//   `in` and `maybe` differ after a chunk of candidates, that lane re-evaluates the
//   chunk with the reference's exact FP64 expression ((dx*dx + dy*dy) + dz*dz, no FMA),
//   so the returned integers are bit-identical to the reference's.
//   Tensor cores are not used: K = 3 is not a dense contraction.
//
// k_count_exact -- all-pairs FP64 kernel without pruning (validation and cross-check).
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "yawb_internal.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr float EPS32 = 5.9604645e-8f;  // 2^-24
constexpr int CHUNK = 8;                // candidates between consistency checks
constexpr float FAR = 1.0e15f;          // coordinates of padding points (never in range)

struct FastParams {
    // first catalog (sky-cell index)
    const double *sx, *sy, *sz, *sw;
    const double *su, *sv, *st;  // the same rows in the frame of their own patch
    const int *cell_start;
    const SGrid *sgrid;
    const PatchFrame *sframe;
    // second catalog (register tiles)
    const double *rx, *ry, *rz, *rw;
    const Tile *tiles;
    const int *ptile_off;
    // work
    const int *pair_i, *pair_j;
    const long long *pair_item_base;
    long long n_items;
    int n_pairs, n_bins, n_edges;
    const double *r2;
    const float *r2f;
    const BinPar *binpar;
    double rmax_all;  // largest search radius over the z-bins
    unsigned long long *out_cnt;
    double *out_w;
    unsigned long long *counters;  // [0] next item, [1] tests, [2] rechecks, [3] live items
};

// The reference's comparison value: products rounded separately, summed x -> y -> z.
__device__ __forceinline__ double exact_d2(double ax, double ay, double az, double bx, double by, double bz) {
    double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// number of edges strictly below d2 (np.searchsorted(r2, d2, side="left"))
__device__ __forceinline__ int edges_below(const double *__restrict__ e, int n, double d2) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (e[mid] < d2) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// ---- per-warp shared memory -------------------------------------------------------------------
constexpr int CCAP = 128;  // (z-bin, cell-row) combinations resolved per batch

template <bool WEIGHTED>
struct WarpSmem {
    float4 *list;             // [LCAP] staged candidates (-2x, -2y, -2z, |s|^2 - mid)
    double *lw;               // [LCAP] their weights (WEIGHTED)
    unsigned long long *acc;  // [n_bins * nsub] pair counts of the current patch pair
    double *accw;             // same, weighted sums (WEIGHTED)
    double *histw;            // [nsub] scratch histogram of one z-bin (MULTI && WEIGHTED)
    float4 *binrec;           // [n_bins] half extents of the query box (x, y, z) and mid
    float2 *binthr;           // [n_bins] (h - eps, h + eps)
    int *lidx;                // [LCAP] row of the candidate in the sorted first catalog
    int *bin_iv0;             // [n_bins] first cell row of the query box
    int *bin_iu;              // [n_bins] iu0 | iu1 << 16
    int *cstart;              // [n_bins + 1] prefix of cell rows per z-bin
    int *rs0;                 // [CCAP] first candidate row of a (bin, cell-row) run
    int *rpre;                // [CCAP] inclusive prefix of run lengths
    int *cbin;                // [CCAP] z-bin of the run
    unsigned *hist;           // [nsub] (MULTI)
    unsigned short *lbin;     // [LCAP] z-bin of the candidate
};

__host__ __device__ inline size_t warp_smem_bytes(bool weighted, bool multi, int n_bins, int nsub) {
    size_t b = YAWB_LCAP * sizeof(float4);
    if (weighted) b += YAWB_LCAP * sizeof(double);
    b += (size_t)n_bins * nsub * sizeof(unsigned long long);
    if (weighted) b += (size_t)n_bins * nsub * sizeof(double);
    if (multi && weighted) b += (size_t)nsub * sizeof(double);
    b += (size_t)n_bins * (sizeof(float4) + sizeof(float2));
    b += YAWB_LCAP * sizeof(int);
    b += (size_t)(3 * n_bins + 1) * sizeof(int);
    b += 3 * CCAP * sizeof(int);
    if (multi) b += (size_t)nsub * sizeof(unsigned);
    b += YAWB_LCAP * sizeof(unsigned short);
    return (b + 15) & ~(size_t)15;
}

template <bool WEIGHTED>
__device__ __forceinline__ void carve_smem(WarpSmem<WEIGHTED> &S, unsigned char *p, bool multi, int n_bins, int nsub) {
    const size_t nacc = (size_t)n_bins * nsub;
    // 16-byte objects first, then 8-, 4- and 2-byte ones, so every array is naturally aligned
    S.list = (float4 *)p; p += YAWB_LCAP * sizeof(float4);
    S.binrec = (float4 *)p; p += (size_t)n_bins * sizeof(float4);
    S.lw = nullptr; S.accw = nullptr; S.histw = nullptr; S.hist = nullptr;
    if (WEIGHTED) { S.lw = (double *)p; p += YAWB_LCAP * sizeof(double); }
    S.acc = (unsigned long long *)p; p += nacc * sizeof(unsigned long long);
    if (WEIGHTED) { S.accw = (double *)p; p += nacc * sizeof(double); }
    if (multi && WEIGHTED) { S.histw = (double *)p; p += (size_t)nsub * sizeof(double); }
    S.binthr = (float2 *)p; p += (size_t)n_bins * sizeof(float2);
    S.lidx = (int *)p; p += YAWB_LCAP * sizeof(int);
    S.bin_iv0 = (int *)p; p += (size_t)n_bins * sizeof(int);
    S.bin_iu = (int *)p; p += (size_t)n_bins * sizeof(int);
    S.cstart = (int *)p; p += (size_t)(n_bins + 1) * sizeof(int);
    S.rs0 = (int *)p; p += CCAP * sizeof(int);
    S.rpre = (int *)p; p += CCAP * sizeof(int);
    S.cbin = (int *)p; p += CCAP * sizeof(int);
    if (multi) { S.hist = (unsigned *)p; p += (size_t)nsub * sizeof(unsigned); }
    S.lbin = (unsigned short *)p;
}

// ---- exact re-evaluation of one lane's share of a chunk, done by the whole warp ----------------
// Lane `src` saw a test inside the FP32 uncertainty band.  Its (YAWB_RPL x chunk) tests are
// re-evaluated with the reference's FP64 expression, two or more per lane, and summed.
template <bool WEIGHTED>
__device__ __forceinline__ void recheck_chunk(const FastParams &P, const WarpSmem<WEIGHTED> &S, int e0, int e1,
                                              const Tile &tl, int lane, int src, double lo, double hi,
                                              unsigned &cnt_out, double &w_out, unsigned &n_recheck) {
    unsigned cnt = 0;
    double wsum = 0.0;
    for (int t = lane; t < CHUNK * YAWB_RPL; t += 32) {
        const int e = e0 + (t & (CHUNK - 1));
        const int k = src + 32 * (t / CHUNK);
        if (e < e1 && k < tl.count) {
            const int i = S.lidx[e], j = tl.start + k;
            const double d2 = exact_d2(P.sx[i], P.sy[i], P.sz[i], P.rx[j], P.ry[j], P.rz[j]);
            if (d2 > lo && d2 <= hi) {
                cnt += 1;
                if (WEIGHTED) wsum += S.lw[e] * (P.rw ? P.rw[j] : 1.0);
            }
            n_recheck += 1;
        }
    }
    cnt_out = __reduce_add_sync(FULL, cnt);
    if (WEIGHTED) w_out = warp_sum(wsum);
}

// ---- phase 2, single sub-bin (n_edges == 2): the hot loop ----------------------------------
// entries [ea, eb) of the list belong to one z-bin with thresholds (h_in, h_out)
template <bool WEIGHTED>
__device__ __forceinline__ void phase2_single(const FastParams &P, const WarpSmem<WEIGHTED> &S, int ea, int eb,
                                              const float (&rx)[YAWB_RPL], const float (&ry)[YAWB_RPL],
                                              const float (&rz)[YAWB_RPL], const float (&rn)[YAWB_RPL],
                                              float h_in, float h_out, const Tile &tl, int lane, double lo,
                                              double hi, unsigned &cnt_total, double &w_total,
                                              unsigned &n_recheck) {
    for (int e0 = ea; e0 < eb; e0 += CHUNK) {
        const int e1 = min(e0 + CHUNK, eb);
        float c_in = 0.f, c_maybe = 0.f;
        double ws[YAWB_RPL];
        if (WEIGHTED) {
#pragma unroll
            for (int r = 0; r < YAWB_RPL; ++r) ws[r] = 0.0;
        }
#pragma unroll 4
        for (int e = e0; e < e1; ++e) {
            const float4 s = S.list[e];
            double swt = 0.0;
            if (WEIGHTED) swt = S.lw[e];
#pragma unroll
            for (int r = 0; r < YAWB_RPL; ++r) {
                float u = rn[r] + s.w;
                u = fmaf(rx[r], s.x, u);
                u = fmaf(ry[r], s.y, u);
                u = fmaf(rz[r], s.z, u);
                const float au = fabsf(u);
                const bool in = au < h_in;
                if (in) c_in += 1.f;
                if (au < h_out) c_maybe += 1.f;
                if (WEIGHTED) {
                    if (in) ws[r] += swt;
                }
            }
        }
        unsigned c = (unsigned)c_in;
        double wsum = 0.0;
        if (WEIGHTED) {
#pragma unroll
            for (int r = 0; r < YAWB_RPL; ++r) {
                const int k = lane + 32 * r;
                if (ws[r] != 0.0) wsum += ws[r] * (P.rw ? P.rw[tl.start + min(k, tl.count - 1)] : 1.0);
            }
        }
        unsigned flagged = __ballot_sync(FULL, c_in != c_maybe);
        while (flagged) {  // warp-uniform: some lane met the uncertainty band of an edge
            const int src = __ffs(flagged) - 1;
            flagged &= flagged - 1;
            unsigned cx = 0;
            double wx = 0.0;
            recheck_chunk<WEIGHTED>(P, S, e0, e1, tl, lane, src, lo, hi, cx, wx, n_recheck);
            if (lane == src) {
                c = cx;
                wsum = wx;
            }
        }
        cnt_total += c;
        if (WEIGHTED) w_total += wsum;
    }
}

// ---- phase 2, single sub-bin, unweighted: 7 FMA-pipe instructions per test ----------------------
// v = sat(C - K |u|) is a ramp through the uncertainty band: exactly 1 well inside the bin, exactly 0
// well outside, in [0.375, 0.625] wherever FP32 cannot decide (K = 1 / (8 eps), C = 1/2 + h K).
// sum(v) and sum(v*v) are accumulated; they are equal iff every v of the chunk was 0 or 1 (then
// sum(v) is the exact count); any undecidable test makes them differ by >= 0.23 and the lane's share
// of the chunk is re-evaluated in FP64.  FADD + 3 FFMA + FFMA.SAT + FADD + FFMA, no ALU-pipe work.
__device__ __forceinline__ void phase2_single_sat(const FastParams &P, const WarpSmem<false> &S, int ea, int eb,
                                                  const float (&rx)[YAWB_RPL], const float (&ry)[YAWB_RPL],
                                                  const float (&rz)[YAWB_RPL], const float (&rn)[YAWB_RPL],
                                                  float K, float C, const Tile &tl, int lane, double lo, double hi,
                                                  unsigned &cnt_total, unsigned &n_recheck) {
    const float nK = -K;
    for (int e0 = ea; e0 < eb; e0 += CHUNK) {
        const int e1 = min(e0 + CHUNK, eb);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
        for (int e = e0; e < e1; ++e) {
            const float4 s = S.list[e];
#pragma unroll
            for (int r = 0; r < YAWB_RPL; ++r) {
                float u = rn[r] + s.w;
                u = fmaf(rx[r], s.x, u);
                u = fmaf(ry[r], s.y, u);
                u = fmaf(rz[r], s.z, u);
                const float v = __saturatef(fmaf(fabsf(u), nK, C));
                s1 += v;
                s2 = fmaf(v, v, s2);
            }
        }
        unsigned c = (unsigned)(s1 + 0.5f);
        unsigned flagged = __ballot_sync(FULL, s1 != s2);
        while (flagged) {
            const int src = __ffs(flagged) - 1;
            flagged &= flagged - 1;
            unsigned cx = 0;
            double wx = 0.0;
            recheck_chunk<false>(P, S, e0, e1, tl, lane, src, lo, hi, cx, wx, n_recheck);
            if (lane == src) c = cx;
        }
        cnt_total += c;
    }
}

// ---- phase 2, several sub-bins (r-weights, multi-scale) ------------------------------------
template <bool WEIGHTED>
__device__ __forceinline__ void phase2_multi(const FastParams &P, const WarpSmem<WEIGHTED> &S, int ea, int eb,
                                             const float (&rx)[YAWB_RPL], const float (&ry)[YAWB_RPL],
                                             const float (&rz)[YAWB_RPL], const float (&rn)[YAWB_RPL],
                                             float h_out, float eps, float mid, const Tile &tl, int lane, int b,
                                             unsigned &n_recheck) {
    const int ne = P.n_edges;
    const float *ef = P.r2f + (size_t)b * ne;
    const double *ed = P.r2 + (size_t)b * ne;
    for (int e = ea; e < eb; ++e) {
        const float4 s = S.list[e];
#pragma unroll
        for (int r = 0; r < YAWB_RPL; ++r) {
            float u = rn[r] + s.w;
            u = fmaf(rx[r], s.x, u);
            u = fmaf(ry[r], s.y, u);
            u = fmaf(rz[r], s.z, u);
            if (fabsf(u) < h_out) {  // possibly inside [lo, hi]
                const float d2f = u + mid;
                int lo = 0, hi = ne;  // edges strictly below d2f (float copy of the edges)
                while (lo < hi) {
                    int m = (lo + hi) >> 1;
                    if (ef[m] < d2f) lo = m + 1; else hi = m;
                }
                int k = lo;
                // distance to the neighbouring edges decides whether FP32 was good enough
                float gap = FLT_MAX;
                if (k > 0) gap = fminf(gap, d2f - ef[k - 1]);
                if (k < ne) gap = fminf(gap, ef[k] - d2f);
                const int j = tl.start + min(lane + 32 * r, tl.count - 1);
                if (!(gap > eps)) {
                    const int i = S.lidx[e];
                    const double d2 = exact_d2(P.sx[i], P.sy[i], P.sz[i], P.rx[j], P.ry[j], P.rz[j]);
                    k = edges_below(ed, ne, d2);
                    n_recheck += 1;
                }
                if (k >= 1 && k < ne) {
                    atomicAdd(&S.hist[k - 1], 1u);
                    if (WEIGHTED) atomicAdd(&S.histw[k - 1], S.lw[e] * (P.rw ? P.rw[j] : 1.0));
                }
            }
        }
    }
}

// ---- the kernel -------------------------------------------------------------------------------
template <bool WEIGHTED, bool MULTI, bool SAT>
__global__ void __launch_bounds__(YAWB_WARPS * 32, YAWB_MIN_CTAS) k_count_fast(const FastParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nsub = P.n_edges - 1;
    const int nacc = P.n_bins * nsub;

    WarpSmem<WEIGHTED> S;
    carve_smem<WEIGHTED>(S, smem_raw + (size_t)warp * warp_smem_bytes(WEIGHTED, MULTI, P.n_bins, nsub), MULTI,
                         P.n_bins, nsub);
    for (int k = lane; k < nacc; k += 32) {
        S.acc[k] = 0ull;
        if (WEIGHTED) S.accw[k] = 0.0;
    }
    __syncwarp();

    int cur_pair = -1;
    long long cur_lo = 0, cur_hi = 0;  // item range of cur_pair
    unsigned long long n_tests = 0;
    unsigned n_recheck = 0, n_live = 0;

    auto flush_pair = [&]() {
        if (cur_pair < 0) return;
        __syncwarp();
        for (int k = lane; k < nacc; k += 32) {
            const unsigned long long c = S.acc[k];
            if (c) {
                atomicAdd(&P.out_cnt[(size_t)cur_pair * nacc + k], c);
                S.acc[k] = 0ull;
            }
            if (WEIGHTED) {
                const double w = S.accw[k];
                if (w != 0.0) {
                    atomicAdd(&P.out_w[(size_t)cur_pair * nacc + k], w);
                    S.accw[k] = 0.0;
                }
            }
        }
        __syncwarp();
    };

    while (true) {
        long long item = 0;
        if (lane == 0) item = (long long)atomicAdd(&P.counters[0], 1ull);
        item = __shfl_sync(FULL, item, 0);
        if (item >= P.n_items) break;

        if (item < cur_lo || item >= cur_hi) {  // new patch pair: flush, then locate it
            flush_pair();
            int lo = 0, hi = P.n_pairs;  // last k with base[k] <= item
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (P.pair_item_base[mid] <= item) lo = mid; else hi = mid;
            }
            cur_pair = lo;
            cur_lo = P.pair_item_base[lo];
            cur_hi = P.pair_item_base[lo + 1];
        }
        const int p1 = P.pair_i[cur_pair];
        const int p2 = P.pair_j[cur_pair];
        const Tile tl = P.tiles[P.ptile_off[p2] + (int)(item - cur_lo)];
        const PatchFrame &F = P.sframe[p1];

        // bounding-sphere rejection of the whole item (chord distances obey the triangle inequality)
        const int b_lo = tl.bin >= 0 ? tl.bin : 0;
        const int b_hi = tl.bin >= 0 ? tl.bin + 1 : P.n_bins;
        {
            const double dx = (double)tl.cx - F.c[0], dy = (double)tl.cy - F.c[1], dz = (double)tl.cz - F.c[2];
            const double rmax_all = tl.bin >= 0 ? (P.binpar[tl.bin].empty ? 0.0 : P.binpar[tl.bin].rmax) : P.rmax_all;
            const double reach = F.radius + (double)tl.rad + rmax_all + 1e-9;
            if (rmax_all == 0.0 || dx * dx + dy * dy + dz * dz > reach * reach) continue;
        }
        n_live += 1;

        // second-catalog points of this lane in the frame of patch p1: pass 1 finds the tile box,
        // pass 2 re-derives the coordinates relative to the box centre and rounds them ONCE to float
        const double c0 = F.c[0], c1 = F.c[1], c2 = F.c[2];
        const double a0 = F.e1[0], a1 = F.e1[1], a2 = F.e1[2];
        const double g0 = F.e2[0], g1 = F.e2[1], g2 = F.e2[2];
        double umin = DBL_MAX, umax = -DBL_MAX, vmin = DBL_MAX, vmax = -DBL_MAX, tmin = DBL_MAX, tmax = -DBL_MAX;
#pragma unroll
        for (int r = 0; r < YAWB_RPL; ++r) {
            const int k = lane + 32 * r;
            if (k < tl.count) {
                const int j = tl.start + k;
                const double dx = P.rx[j] - c0, dy = P.ry[j] - c1, dz = P.rz[j] - c2;
                const double lu = dx * a0 + dy * a1 + dz * a2;
                const double lv = dx * g0 + dy * g1 + dz * g2;
                const double lt = dx * c0 + dy * c1 + dz * c2;
                umin = fmin(umin, lu); umax = fmax(umax, lu);
                vmin = fmin(vmin, lv); vmax = fmax(vmax, lv);
                tmin = fmin(tmin, lt); tmax = fmax(tmax, lt);
            }
        }
        umin = warp_min(umin); umax = warp_max(umax);
        vmin = warp_min(vmin); vmax = warp_max(vmax);
        tmin = warp_min(tmin); tmax = warp_max(tmax);
        const double ou = 0.5 * (umin + umax), ov = 0.5 * (vmin + vmax), ot = 0.5 * (tmin + tmax);

        float rx[YAWB_RPL], ry[YAWB_RPL], rz[YAWB_RPL], rn[YAWB_RPL];
#pragma unroll
        for (int r = 0; r < YAWB_RPL; ++r) {
            const int k = lane + 32 * r;
            if (k < tl.count) {
                const int j = tl.start + k;
                const double dx = P.rx[j] - c0, dy = P.ry[j] - c1, dz = P.rz[j] - c2;
                rx[r] = (float)(dx * a0 + dy * a1 + dz * a2 - ou);
                ry[r] = (float)(dx * g0 + dy * g1 + dz * g2 - ov);
                rz[r] = (float)(dx * c0 + dy * c1 + dz * c2 - ot);
                rn[r] = rx[r] * rx[r] + ry[r] * ry[r] + rz[r] * rz[r];
            } else {
                rx[r] = FAR; ry[r] = FAR; rz[r] = FAR;
                rn[r] = 3.0f * FAR * FAR;
            }
        }
        const SGrid G = P.sgrid[p1];
        const double eu = 0.5 * (umax - umin), ev = 0.5 * (vmax - vmin), et = 0.5 * (tmax - tmin);

        // ---- step 1: per z-bin query box, thresholds and cell rows (one lane per z-bin) ----
        int carry = 0;
        for (int b0 = b_lo; b0 < b_hi; b0 += 32) {
            const int b = b0 + lane;
            int nrows = 0;
            if (b < b_hi) {
                const BinPar bp = P.binpar[b];
                if (!bp.empty) {
                    // query box = tile box grown by the search radius (sound: |du|,|dv|,|dt| <= chord)
                    const double qu0 = umin - bp.rmax, qu1 = umax + bp.rmax;
                    const double qv0 = vmin - bp.rmax, qv1 = vmax + bp.rmax;
                    // cell range of the box; floor((x - u0) * inv_c) is the same monotone expression the
                    // keys were made with, so a point inside the box cannot sit in a cell outside the range
                    const double fu0 = floor((qu0 - G.u0) * G.inv_c), fu1 = floor((qu1 - G.u0) * G.inv_c);
                    const double fv0 = floor((qv0 - G.v0) * G.inv_c), fv1 = floor((qv1 - G.v0) * G.inv_c);
                    if (!(fu1 < 0.0 || fv1 < 0.0 || fu0 > (double)(G.gu - 1) || fv0 > (double)(G.gv - 1))) {
                        const int iu0 = (int)fmax(fu0, 0.0), iv0 = (int)fmax(fv0, 0.0);
                        const int iu1 = (int)fmin(fu1, (double)(G.gu - 1)), iv1 = (int)fmin(fv1, (double)(G.gv - 1));
                        nrows = iv1 - iv0 + 1;
                        S.bin_iv0[b] = iv0;
                        S.bin_iu[b] = iu0 | (iu1 << 16);
                    }
                    // half extents rounded up; they bound every staged vector, hence the FP32 error of u
                    const float hx = (float)(eu + bp.rmax) * 1.000001f, hy = (float)(ev + bp.rmax) * 1.000001f,
                                hz = (float)(et + bp.rmax) * 1.000001f;
                    const float m2 = hx * hx + hy * hy + hz * hz;
                    const float eps = 64.0f * EPS32 * (m2 + bp.mid) * 1.0001f;
                    S.binrec[b] = make_float4(hx, hy, hz, bp.mid);
                    if (MULTI) {
                        S.binthr[b] = make_float2(bp.h + eps, eps + 4.0f * EPS32 * (float)bp.hi);
                    } else if (SAT) {
                        const float K = 1.0f / (8.0f * eps);
                        S.binthr[b] = make_float2(K, 0.5f + bp.h * K);
                    } else {
                        S.binthr[b] = make_float2(bp.h - eps, bp.h + eps);
                    }
                }
            }
            int incl = nrows;  // inclusive scan of the cell rows over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            if (b < b_hi) S.cstart[b + 1] = carry + incl;
            carry += __shfl_sync(FULL, incl, 31);
        }
        if (lane == 0) S.cstart[b_lo] = 0;
        const int n_combo = carry;
        __syncwarp();

        int L = 0;
        // ---- consume the staged list: one run of phase 2 per z-bin segment ----
        auto consume = [&]() {
            __syncwarp();
            int ea = 0;
            while (ea < L) {
                const int b = S.lbin[ea];
                int eb = L;  // first entry after ea that belongs to another z-bin
                for (int base = ea + 1; base < L; base += 32) {
                    const int e = base + lane;
                    const unsigned m = __ballot_sync(FULL, e < L && S.lbin[e] != b);
                    if (m) {
                        eb = base + __ffs(m) - 1;
                        break;
                    }
                }
                const float2 thr = S.binthr[b];
                if (MULTI) {
                    for (int k = lane; k < nsub; k += 32) {
                        S.hist[k] = 0u;
                        if (WEIGHTED) S.histw[k] = 0.0;
                    }
                    __syncwarp();
                    phase2_multi<WEIGHTED>(P, S, ea, eb, rx, ry, rz, rn, thr.x, thr.y, S.binrec[b].w, tl, lane, b,
                                           n_recheck);
                    __syncwarp();
                    for (int k = lane; k < nsub; k += 32) {
                        S.acc[(size_t)b * nsub + k] += S.hist[k];
                        if (WEIGHTED) S.accw[(size_t)b * nsub + k] += S.histw[k];
                    }
                    __syncwarp();
                } else {
                    unsigned cnt_total = 0;
                    double w_total = 0.0;
                    if constexpr (SAT && !WEIGHTED)
                        phase2_single_sat(P, S, ea, eb, rx, ry, rz, rn, thr.x, thr.y, tl, lane, P.binpar[b].lo,
                                          P.binpar[b].hi, cnt_total, n_recheck);
                    else
                        phase2_single<WEIGHTED>(P, S, ea, eb, rx, ry, rz, rn, thr.x, thr.y, tl, lane,
                                                P.binpar[b].lo, P.binpar[b].hi, cnt_total, w_total, n_recheck);
                    const unsigned tot = __reduce_add_sync(FULL, cnt_total);
                    double wtot = 0.0;
                    if (WEIGHTED) wtot = warp_sum(w_total);
                    if (lane == 0) {
                        S.acc[b] += tot;
                        if (WEIGHTED) S.accw[b] += wtot;
                    }
                }
                ea = eb;
            }
            n_tests += (unsigned long long)L * (unsigned long long)tl.count;
            L = 0;
            __syncwarp();
        };

        for (int cb = 0; cb < n_combo; cb += CCAP) {
            const int nb = min(CCAP, n_combo - cb);
            // ---- step 2: one lane per (z-bin, cell row): the run of candidate rows it covers ----
            int running = 0;
            for (int k0 = 0; k0 < nb; k0 += 32) {
                const int k = k0 + lane;
                int cnt = 0, s0 = 0, b = 0;
                if (k < nb) {
                    const int c = cb + k;
                    int lo = b_lo, hi = b_hi;  // last z-bin with cstart[b] <= c
                    while (hi - lo > 1) {
                        const int m = (lo + hi) >> 1;
                        if (S.cstart[m] <= c) lo = m; else hi = m;
                    }
                    b = lo;
                    const int iv = S.bin_iv0[b] + (c - S.cstart[b]);
                    const int iu = S.bin_iu[b];
                    const long long row = G.cell_base + ((long long)b * G.gv + iv) * G.gu;
                    s0 = P.cell_start[row + (iu & 0xffff)];
                    cnt = P.cell_start[row + (iu >> 16) + 1] - s0;
                }
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                if (k < nb) {
                    S.rs0[k] = s0;
                    S.rpre[k] = running + incl;
                    S.cbin[k] = b;
                }
                running += __shfl_sync(FULL, incl, 31);
            }
            const int n_cand = running;
            __syncwarp();

            // ---- step 3: flattened gather, cull against the z-bin's box, stage as float4 ----
            int cur = 0;
            for (int t0 = 0; t0 < n_cand; t0 += 32) {
                const int t = t0 + lane;
                bool ok = t < n_cand;
                int i = 0, b = 0;
                float fx = 0.f, fy = 0.f, fz = 0.f, mid = 0.f;
                if (ok) {
                    while (t >= S.rpre[cur]) ++cur;  // runs are consumed in order; empty runs are skipped
                    i = S.rs0[cur] + (t - (cur ? S.rpre[cur - 1] : 0));
                    b = S.cbin[cur];
                    fx = (float)(P.su[i] - ou);
                    fy = (float)(P.sv[i] - ov);
                    fz = (float)(P.st[i] - ot);
                    const float4 rec = S.binrec[b];
                    mid = rec.w;
                    ok = fabsf(fx) <= rec.x && fabsf(fy) <= rec.y && fabsf(fz) <= rec.z;
                }
                const unsigned m = __ballot_sync(FULL, ok);
                if (ok) {
                    const int pos = L + __popc(m & ((1u << lane) - 1u));
                    const float sn = fx * fx + fy * fy + fz * fz;
                    S.list[pos] = make_float4(-2.0f * fx, -2.0f * fy, -2.0f * fz, sn - mid);
                    S.lidx[pos] = i;
                    S.lbin[pos] = (unsigned short)b;
                    if (WEIGHTED) S.lw[pos] = P.sw ? P.sw[i] : 1.0;
                }
                L += __popc(m);
                if (L > YAWB_LCAP - 32) consume();
            }
            __syncwarp();
        }
        if (L > 0) consume();
    }
    flush_pair();
    if (lane == 0) {
        if (n_tests) atomicAdd(&P.counters[1], n_tests);
        if (n_live) atomicAdd(&P.counters[3], (unsigned long long)n_live);
    }
    const unsigned rc = __reduce_add_sync(FULL, n_recheck);
    if (lane == 0 && rc) atomicAdd(&P.counters[2], (unsigned long long)rc);
}

// ---- exact all-pairs kernel -------------------------------------------------------------------
struct ExactParams {
    const double *sx, *sy, *sz, *sw;
    const int *s_seg;  // [(P * B1) + 1]
    const double *rx, *ry, *rz, *rw;
    const int *r_seg;  // [(P * B2) + 1]
    int b1, b2;        // bins of cat1 / cat2 (b2 == 1 for unbinned)
    const int *pair_i, *pair_j;
    int n_pairs, n_bins, n_edges;
    const double *r2;
    unsigned long long *out_cnt;
    double *out_w;
    unsigned long long *counters;
};

constexpr int EX_THREADS = 256;

template <bool WEIGHTED>
__global__ void __launch_bounds__(EX_THREADS) k_count_exact(const ExactParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ne = P.n_edges, nsub = ne - 1;
    double *edges = (double *)smem_raw;
    double *tile = edges + ne;                       // [4][EX_THREADS] x, y, z, w
    double *histw = tile + 4 * EX_THREADS;           // [nsub]
    unsigned long long *hist = (unsigned long long *)(histw + nsub);  // [nsub]

    const int k = blockIdx.x / P.n_bins, b = blockIdx.x % P.n_bins;
    const int p1 = P.pair_i[k], p2 = P.pair_j[k];
    const int a0 = P.s_seg[p1 * P.b1 + b], a1 = P.s_seg[p1 * P.b1 + b + 1];
    const int c0 = P.b2 > 1 ? P.r_seg[p2 * P.b2 + b] : P.r_seg[p2];
    const int c1 = P.b2 > 1 ? P.r_seg[p2 * P.b2 + b + 1] : P.r_seg[p2 + 1];
    if (a0 >= a1 || c0 >= c1) return;
    if ((long long)blockIdx.y * EX_THREADS >= (a1 - a0)) return;

    for (int e = threadIdx.x; e < ne; e += EX_THREADS) edges[e] = P.r2[(size_t)b * ne + e];
    for (int e = threadIdx.x; e < nsub; e += EX_THREADS) { hist[e] = 0ull; histw[e] = 0.0; }
    __syncthreads();
    const double lo = edges[0], hi = edges[nsub];
    unsigned long long tests = 0;

    for (int ia = a0 + blockIdx.y * EX_THREADS; ia < a1; ia += gridDim.y * EX_THREADS) {
        const int i = ia + threadIdx.x;
        const bool live = i < a1;
        double ax = 0, ay = 0, az = 0, aw = 1.0;
        if (live) {
            ax = P.sx[i]; ay = P.sy[i]; az = P.sz[i];
            if (WEIGHTED && P.sw) aw = P.sw[i];
        }
        for (int jc = c0; jc < c1; jc += EX_THREADS) {
            const int nj = min(EX_THREADS, c1 - jc);
            __syncthreads();
            if ((int)threadIdx.x < nj) {
                const int j = jc + threadIdx.x;
                tile[threadIdx.x] = P.rx[j];
                tile[EX_THREADS + threadIdx.x] = P.ry[j];
                tile[2 * EX_THREADS + threadIdx.x] = P.rz[j];
                tile[3 * EX_THREADS + threadIdx.x] = (WEIGHTED && P.rw) ? P.rw[j] : 1.0;
            }
            __syncthreads();
            if (live) {
                for (int jj = 0; jj < nj; ++jj) {
                    const double d2 = exact_d2(ax, ay, az, tile[jj], tile[EX_THREADS + jj], tile[2 * EX_THREADS + jj]);
                    if (d2 > lo && d2 <= hi) {
                        const int kk = nsub == 1 ? 1 : edges_below(edges, ne, d2);
                        atomicAdd(&hist[kk - 1], 1ull);
                        if (WEIGHTED) atomicAdd(&histw[kk - 1], aw * tile[3 * EX_THREADS + jj]);
                    }
                }
                tests += nj;
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nsub; e += EX_THREADS) {
        const size_t o = ((size_t)k * P.n_bins + b) * nsub + e;
        if (hist[e]) atomicAdd(&P.out_cnt[o], hist[e]);
        if (WEIGHTED && histw[e] != 0.0) atomicAdd(&P.out_w[o], histw[e]);
    }
    for (int o = 16; o; o >>= 1) tests += __shfl_xor_sync(FULL, tests, o);
    if ((threadIdx.x & 31) == 0 && tests) atomicAdd(&P.counters[1], tests);
}

}  // namespace

// -----------------------------------------------------------------------------------------------
int yawb_launch_count_fast(yawb_ctx *ctx, const CountArgs &a, int *launches) {
    FastParams P{};
    P.sx = a.c1->sx; P.sy = a.c1->sy; P.sz = a.c1->sz; P.sw = a.c1->sw;
    P.su = a.c1->su; P.sv = a.c1->sv; P.st = a.c1->st;
    P.rmax_all = a.rmax_all;
    P.cell_start = a.c1->cell_start; P.sgrid = a.c1->d_sgrid; P.sframe = a.c1->d_frames;
    P.rx = a.c2->rx; P.ry = a.c2->ry; P.rz = a.c2->rz; P.rw = a.c2->rw;
    P.tiles = a.c2->d_tiles; P.ptile_off = a.c2->d_ptile_off;
    P.pair_i = a.d_pair_i; P.pair_j = a.d_pair_j; P.pair_item_base = a.d_pair_item_base;
    P.n_items = a.n_items; P.n_pairs = a.n_pairs; P.n_bins = a.n_bins; P.n_edges = a.n_edges;
    P.r2 = a.d_r2; P.r2f = a.d_r2f; P.binpar = a.d_binpar;
    P.out_cnt = a.d_out_cnt; P.out_w = a.d_out_w; P.counters = ctx->d_counters;
    if (a.n_items == 0) return 0;

    const bool multi = a.n_edges > 2;
    const int nsub = a.n_edges - 1;
    const size_t smem = YAWB_WARPS * warp_smem_bytes(a.weighted, multi, a.n_bins, nsub);
    YAWB_REQUIRE(smem <= 227 * 1024, "too many z-bins x sub-bins for the shared-memory accumulators (%zu B)", smem);
    // persistent grid: a multiple of the SM count, warps pull items from a global counter
    const long long warps_needed = a.n_items;
    int ctas = ctx->sms * YAWB_MIN_CTAS;
    ctas = (int)std::min<long long>(ctas, (warps_needed + YAWB_WARPS - 1) / YAWB_WARPS);
    ctas = std::max(ctas, 1);

#define LAUNCH(W, M, T)                                                                                   \
    do {                                                                                                  \
        YAWB_CUDA(cudaFuncSetAttribute(k_count_fast<W, M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)smem));                                                       \
        k_count_fast<W, M, T><<<ctas, YAWB_WARPS * 32, smem, ctx->stream>>>(P);                           \
    } while (0)
    // unweighted single-bin counts use the 7-instruction saturating test unless YAWB_PAIR_TEST=pred
    const char *variant = getenv("YAWB_PAIR_TEST");
    const bool sat = !(variant && strcmp(variant, "pred") == 0);
    if (a.weighted) {
        if (multi) LAUNCH(true, true, false); else LAUNCH(true, false, false);
    } else {
        if (multi) LAUNCH(false, true, false);
        else if (sat) LAUNCH(false, false, true);
        else LAUNCH(false, false, false);
    }
#undef LAUNCH
    YAWB_CUDA(cudaGetLastError());
    *launches += 1;
    return 0;
}

int yawb_launch_count_exact(yawb_ctx *ctx, const CountArgs &a, int *launches) {
    ExactParams P{};
    P.sx = a.c1->sx; P.sy = a.c1->sy; P.sz = a.c1->sz; P.sw = a.c1->sw; P.s_seg = a.c1->d_seg_off;
    P.rx = a.c2->rx; P.ry = a.c2->ry; P.rz = a.c2->rz; P.rw = a.c2->rw; P.r_seg = a.c2->d_seg_off;
    P.b1 = a.c1->n_bins; P.b2 = a.c2->n_bins;
    P.pair_i = a.d_pair_i; P.pair_j = a.d_pair_j;
    P.n_pairs = a.n_pairs; P.n_bins = a.n_bins; P.n_edges = a.n_edges;
    P.r2 = a.d_r2; P.out_cnt = a.d_out_cnt; P.out_w = a.d_out_w; P.counters = ctx->d_counters;
    if (a.n_pairs == 0) return 0;

    // rows of the largest (patch, bin) segment of cat1 decide grid.y
    int max_seg = 0;
    for (size_t s = 0; s + 1 < a.c1->h_seg_off.size(); ++s)
        max_seg = std::max(max_seg, a.c1->h_seg_off[s + 1] - a.c1->h_seg_off[s]);
    int gy = std::max(1, std::min(64, (max_seg + EX_THREADS - 1) / EX_THREADS));
    const int nsub = a.n_edges - 1;
    const size_t smem = (size_t)(a.n_edges + 4 * EX_THREADS + nsub) * sizeof(double) + (size_t)nsub * 8;
    dim3 grid((unsigned)(a.n_pairs * a.n_bins), (unsigned)gy);
    if (a.weighted)
        k_count_exact<true><<<grid, EX_THREADS, smem, ctx->stream>>>(P);
    else
        k_count_exact<false><<<grid, EX_THREADS, smem, ctx->stream>>>(P);
    YAWB_CUDA(cudaGetLastError());
    *launches += 1;
    return 0;
}
