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

#include "yawb_internal.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr float EPS32 = 5.9604645e-8f;  // 2^-24
constexpr int CHUNK = 8;                // candidates between consistency checks
constexpr float FAR = 1.0e15f;          // coordinates of padding points (never in range)

struct FastParams {
    // first catalog (sky-cell index)
    const double *sx, *sy, *sz, *sw;
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

// per-warp shared memory
template <bool WEIGHTED>
struct WarpSmem {
    float4 *list;                // [LCAP] staged candidates
    int *lidx;                   // [LCAP] their row in the sorted first catalog
    double *lw;                  // [LCAP] their weights (WEIGHTED)
    unsigned long long *acc;     // [n_bins * nsub] pair counts of the current patch pair
    double *accw;                // same, weighted sums (WEIGHTED)
    unsigned *hist;              // [nsub] scratch histogram of the current bin (MULTI)
    double *histw;               // [nsub] (MULTI && WEIGHTED)
};

__host__ __device__ inline size_t warp_smem_bytes(bool weighted, bool multi, int n_bins, int nsub) {
    size_t b = YAWB_LCAP * (sizeof(float4) + sizeof(int));
    if (weighted) b += YAWB_LCAP * sizeof(double);
    b += (size_t)n_bins * nsub * sizeof(unsigned long long);
    if (weighted) b += (size_t)n_bins * nsub * sizeof(double);
    if (multi) b += (size_t)nsub * sizeof(unsigned) + (weighted ? (size_t)nsub * sizeof(double) : 0);
    return (b + 15) & ~(size_t)15;
}

// ---- phase 2, single sub-bin (n_edges == 2): the hot loop ----------------------------------
template <bool WEIGHTED>
__device__ __forceinline__ void phase2_single(const FastParams &P, const WarpSmem<WEIGHTED> &S, int L,
                                              const float (&rx)[YAWB_RPL], const float (&ry)[YAWB_RPL],
                                              const float (&rz)[YAWB_RPL], const float (&rn)[YAWB_RPL],
                                              const double (&wr)[YAWB_RPL], float h_in, float h_out,
                                              const Tile &tl, int lane, const BinPar &bp,
                                              unsigned &cnt_total, double &w_total, unsigned &n_recheck) {
    for (int e0 = 0; e0 < L; e0 += CHUNK) {
        const int e1 = min(e0 + CHUNK, L);
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
            for (int r = 0; r < YAWB_RPL; ++r) wsum += wr[r] * ws[r];
        }
        if (c_in != c_maybe) {
            // some test of this lane fell inside the FP32 uncertainty band of an edge:
            // redo the lane's share of the chunk exactly (reference arithmetic)
            c = 0;
            wsum = 0.0;
            for (int r = 0; r < YAWB_RPL; ++r) {
                const int k = lane + 32 * r;
                if (k >= tl.count) break;
                const int j = tl.start + k;
                const double bx = P.rx[j], by = P.ry[j], bz = P.rz[j];
                double wj = 1.0;
                if (WEIGHTED) wj = wr[r];
                for (int e = e0; e < e1; ++e) {
                    const int i = S.lidx[e];
                    const double d2 = exact_d2(P.sx[i], P.sy[i], P.sz[i], bx, by, bz);
                    if (d2 > bp.lo && d2 <= bp.hi) {
                        c += 1;
                        if (WEIGHTED) wsum += S.lw[e] * wj;
                    }
                }
            }
            n_recheck += (unsigned)((e1 - e0) * YAWB_RPL);
        }
        cnt_total += c;
        if (WEIGHTED) w_total += wsum;
    }
}

// ---- phase 2, several sub-bins (r-weights, multi-scale) ------------------------------------
template <bool WEIGHTED>
__device__ __forceinline__ void phase2_multi(const FastParams &P, const WarpSmem<WEIGHTED> &S, int L,
                                             const float (&rx)[YAWB_RPL], const float (&ry)[YAWB_RPL],
                                             const float (&rz)[YAWB_RPL], const float (&rn)[YAWB_RPL],
                                             const double (&wr)[YAWB_RPL], float h_out, float eps,
                                             const Tile &tl, int lane, const BinPar &bp, int b,
                                             unsigned &n_recheck) {
    const int ne = P.n_edges;
    const float *ef = P.r2f + (size_t)b * ne;
    const double *ed = P.r2 + (size_t)b * ne;
    for (int e = 0; e < L; ++e) {
        const float4 s = S.list[e];
#pragma unroll
        for (int r = 0; r < YAWB_RPL; ++r) {
            float u = rn[r] + s.w;
            u = fmaf(rx[r], s.x, u);
            u = fmaf(ry[r], s.y, u);
            u = fmaf(rz[r], s.z, u);
            if (fabsf(u) < h_out) {  // possibly inside [lo, hi]
                const float d2f = u + bp.mid;
                int lo = 0, hi = ne;  // edges strictly below d2f (float copy of the edges)
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (ef[mid] < d2f) lo = mid + 1; else hi = mid;
                }
                int k = lo;
                // distance to the neighbouring edges decides whether FP32 was good enough
                float gap = FLT_MAX;
                if (k > 0) gap = fminf(gap, d2f - ef[k - 1]);
                if (k < ne) gap = fminf(gap, ef[k] - d2f);
                if (!(gap > eps)) {
                    const int kk = lane + 32 * r;
                    const int j = tl.start + min(kk, tl.count - 1);
                    const int i = S.lidx[e];
                    const double d2 = exact_d2(P.sx[i], P.sy[i], P.sz[i], P.rx[j], P.ry[j], P.rz[j]);
                    k = edges_below(ed, ne, d2);
                    n_recheck += 1;
                }
                if (k >= 1 && k < ne) {
                    atomicAdd(&S.hist[k - 1], 1u);
                    if (WEIGHTED) atomicAdd(&S.histw[k - 1], S.lw[e] * wr[r]);
                }
            }
        }
    }
}

// ---- the kernel -------------------------------------------------------------------------------
template <bool WEIGHTED, bool MULTI>
__global__ void __launch_bounds__(YAWB_WARPS * 32, 2) k_count_fast(const FastParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nsub = P.n_edges - 1;
    const int nacc = P.n_bins * nsub;

    WarpSmem<WEIGHTED> S;
    {
        unsigned char *p = smem_raw + (size_t)warp * warp_smem_bytes(WEIGHTED, MULTI, P.n_bins, nsub);
        S.list = (float4 *)p; p += YAWB_LCAP * sizeof(float4);
        S.lw = nullptr; S.accw = nullptr; S.hist = nullptr; S.histw = nullptr;
        if (WEIGHTED) { S.lw = (double *)p; p += YAWB_LCAP * sizeof(double); }
        S.acc = (unsigned long long *)p; p += (size_t)nacc * sizeof(unsigned long long);
        if (WEIGHTED) { S.accw = (double *)p; p += (size_t)nacc * sizeof(double); }
        if (MULTI && WEIGHTED) { S.histw = (double *)p; p += (size_t)nsub * sizeof(double); }
        S.lidx = (int *)p; p += YAWB_LCAP * sizeof(int);
        if (MULTI) { S.hist = (unsigned *)p; }
    }
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
        double rmax_all = 0.0;
        for (int b = b_lo; b < b_hi; ++b)
            if (!P.binpar[b].empty) rmax_all = fmax(rmax_all, P.binpar[b].rmax);
        {
            const double dx = (double)tl.cx - F.c[0], dy = (double)tl.cy - F.c[1], dz = (double)tl.cz - F.c[2];
            const double reach = F.radius + (double)tl.rad + rmax_all + 1e-9;
            if (rmax_all == 0.0 || dx * dx + dy * dy + dz * dz > reach * reach) continue;
        }
        n_live += 1;

        // second-catalog points of this lane, rotated into the frame of patch p1
        double lu[YAWB_RPL], lv[YAWB_RPL], lt[YAWB_RPL], wr[YAWB_RPL];
        double umin = DBL_MAX, umax = -DBL_MAX, vmin = DBL_MAX, vmax = -DBL_MAX, tmin = DBL_MAX, tmax = -DBL_MAX;
#pragma unroll
        for (int r = 0; r < YAWB_RPL; ++r) {
            const int k = lane + 32 * r;
            wr[r] = 0.0;
            if (k < tl.count) {
                const int j = tl.start + k;
                const double dx = P.rx[j] - F.c[0], dy = P.ry[j] - F.c[1], dz = P.rz[j] - F.c[2];
                lu[r] = dx * F.e1[0] + dy * F.e1[1] + dz * F.e1[2];
                lv[r] = dx * F.e2[0] + dy * F.e2[1] + dz * F.e2[2];
                lt[r] = dx * F.c[0] + dy * F.c[1] + dz * F.c[2];
                umin = fmin(umin, lu[r]); umax = fmax(umax, lu[r]);
                vmin = fmin(vmin, lv[r]); vmax = fmax(vmax, lv[r]);
                tmin = fmin(tmin, lt[r]); tmax = fmax(tmax, lt[r]);
                if (WEIGHTED) wr[r] = P.rw ? P.rw[j] : 1.0;
            } else {
                lu[r] = lv[r] = lt[r] = 0.0;
            }
        }
        umin = warp_min(umin); umax = warp_max(umax);
        vmin = warp_min(vmin); vmax = warp_max(vmax);
        tmin = warp_min(tmin); tmax = warp_max(tmax);
        const double ou = 0.5 * (umin + umax), ov = 0.5 * (vmin + vmax), ot = 0.5 * (tmin + tmax);

        float rx[YAWB_RPL], ry[YAWB_RPL], rz[YAWB_RPL], rn[YAWB_RPL];
#pragma unroll
        for (int r = 0; r < YAWB_RPL; ++r) {
            if (lane + 32 * r < tl.count) {
                rx[r] = (float)(lu[r] - ou);
                ry[r] = (float)(lv[r] - ov);
                rz[r] = (float)(lt[r] - ot);
                rn[r] = rx[r] * rx[r] + ry[r] * ry[r] + rz[r] * rz[r];
            } else {
                rx[r] = FAR; ry[r] = FAR; rz[r] = FAR;
                rn[r] = 3.0f * FAR * FAR;
            }
        }
        const SGrid G = P.sgrid[p1];

        for (int b = b_lo; b < b_hi; ++b) {
            const BinPar bp = P.binpar[b];
            if (bp.empty) continue;
            // query box = tile box grown by the search radius (sound: |du|,|dv|,|dt| <= chord)
            const double qu0 = umin - bp.rmax, qu1 = umax + bp.rmax;
            const double qv0 = vmin - bp.rmax, qv1 = vmax + bp.rmax;
            const double qt0 = tmin - bp.rmax, qt1 = tmax + bp.rmax;
            // cell range of the box; floor((x - u0) * inv_c) is the same monotone expression the keys
            // were made with, so a point inside the box can never sit in a cell outside the range
            const double fu0 = floor((qu0 - G.u0) * G.inv_c), fu1 = floor((qu1 - G.u0) * G.inv_c);
            const double fv0 = floor((qv0 - G.v0) * G.inv_c), fv1 = floor((qv1 - G.v0) * G.inv_c);
            if (fu1 < 0.0 || fv1 < 0.0 || fu0 > (double)(G.gu - 1) || fv0 > (double)(G.gv - 1)) continue;
            const int iu0 = (int)fmax(fu0, 0.0), iv0 = (int)fmax(fv0, 0.0);
            const int iu1 = (int)fmin(fu1, (double)(G.gu - 1)), iv1 = (int)fmin(fv1, (double)(G.gv - 1));

            // FP32 error bound of u for this (tile, bin): all staged vectors lie in the query box
            const float hu = (float)(0.5 * (qu1 - qu0)), hv = (float)(0.5 * (qv1 - qv0)), ht = (float)(0.5 * (qt1 - qt0));
            const float m2 = hu * hu + hv * hv + ht * ht;
            const float eps = 64.0f * EPS32 * (m2 + bp.mid) * 1.0001f;
            const float h_in = bp.h - eps, h_out = bp.h + eps;

            unsigned cnt_total = 0;
            double w_total = 0.0;
            if (MULTI) {
                for (int k = lane; k < nsub; k += 32) {
                    S.hist[k] = 0u;
                    if (WEIGHTED) S.histw[k] = 0.0;
                }
            }
            int L = 0;
            auto run_phase2 = [&]() {
                __syncwarp();
                if (MULTI)
                    phase2_multi<WEIGHTED>(P, S, L, rx, ry, rz, rn, wr, h_out, eps + 4.0f * EPS32 * (float)bp.hi, tl,
                                           lane, bp, b, n_recheck);
                else
                    phase2_single<WEIGHTED>(P, S, L, rx, ry, rz, rn, wr, h_in, h_out, tl, lane, bp, cnt_total,
                                            w_total, n_recheck);
                n_tests += (unsigned long long)L * (unsigned long long)tl.count;
                L = 0;
                __syncwarp();
            };

            const long long bin_base = G.cell_base + (long long)b * G.gu * G.gv;
            for (int iv = iv0; iv <= iv1; ++iv) {
                const long long row = bin_base + (long long)iv * G.gu;
                const int s0 = P.cell_start[row + iu0], s1 = P.cell_start[row + iu1 + 1];
                for (int base = s0; base < s1; base += 32) {
                    const int i = base + lane;
                    bool ok = i < s1;
                    double du = 0, dv = 0, dt = 0;
                    if (ok) {
                        const double dx = P.sx[i] - F.c[0], dy = P.sy[i] - F.c[1], dz = P.sz[i] - F.c[2];
                        du = dx * F.e1[0] + dy * F.e1[1] + dz * F.e1[2];
                        dv = dx * F.e2[0] + dy * F.e2[1] + dz * F.e2[2];
                        dt = dx * F.c[0] + dy * F.c[1] + dz * F.c[2];
                        ok = du >= qu0 && du <= qu1 && dv >= qv0 && dv <= qv1 && dt >= qt0 && dt <= qt1;
                    }
                    const unsigned m = __ballot_sync(FULL, ok);
                    if (ok) {
                        const int pos = L + __popc(m & ((1u << lane) - 1u));
                        const float fx = (float)(du - ou), fy = (float)(dv - ov), fz = (float)(dt - ot);
                        const double sn = (double)fx * fx + (double)fy * fy + (double)fz * fz;
                        S.list[pos] = make_float4(-2.0f * fx, -2.0f * fy, -2.0f * fz, (float)(sn - (double)bp.mid));
                        S.lidx[pos] = i;
                        if (WEIGHTED) S.lw[pos] = P.sw ? P.sw[i] : 1.0;
                    }
                    L += __popc(m);
                    if (L > YAWB_LCAP - 32) run_phase2();
                }
            }
            if (L > 0) run_phase2();

            // fold this bin into the warp's accumulators of the current patch pair
            if (MULTI) {
                __syncwarp();
                for (int k = lane; k < nsub; k += 32) {
                    S.acc[(size_t)b * nsub + k] += S.hist[k];
                    if (WEIGHTED) S.accw[(size_t)b * nsub + k] += S.histw[k];
                }
                __syncwarp();
            } else {
                const unsigned tot = __reduce_add_sync(FULL, cnt_total);
                double wtot = 0.0;
                if (WEIGHTED) wtot = warp_sum(w_total);
                if (lane == 0) {
                    S.acc[b] += tot;
                    if (WEIGHTED) S.accw[b] += wtot;
                }
            }
        }
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
    int ctas = ctx->sms * 2;
    ctas = (int)std::min<long long>(ctas, (warps_needed + YAWB_WARPS - 1) / YAWB_WARPS);
    ctas = std::max(ctas, 1);

#define LAUNCH(W, M)                                                                                   \
    do {                                                                                               \
        YAWB_CUDA(cudaFuncSetAttribute(k_count_fast<W, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)smem));                                                    \
        k_count_fast<W, M><<<ctas, YAWB_WARPS * 32, smem, ctx->stream>>>(P);                           \
    } while (0)
    if (a.weighted) {
        if (multi) LAUNCH(true, true); else LAUNCH(true, false);
    } else {
        if (multi) LAUNCH(false, true); else LAUNCH(false, false);
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
