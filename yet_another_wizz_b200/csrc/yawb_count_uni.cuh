// The production pair-count kernel k_count_uni and its building blocks (included by yawb_count.cu after
// the shared device functions).
//
// Every warp owns a shared-memory channel (control block + tables + one candidate list) and loops over
// work items (register tile of the second catalog x linked first-catalog patch):
//   step 1  ws_begin_item  one lane per z-bin: query box -> cell rows, FP32 error bound, thresholds
//   step 2  ws_fill        one lane per (z-bin, cell row): the contiguous run of candidate rows
//   step 3  ws_fill        flattened gather of the runs, cull against the z-bin's box, stage the survivors
//                          in the tile-local frame
//   tests   ws_consume     one pass of the FP32 pair test per z-bin segment of the staged list
// A warp-specialised variant (8 gather warps feeding 16 test warps per SM through double-buffered
// channels, setmaxnreg-rebalanced registers) was measured slower on B200 in round 1 (11-16 ms against
// 7 ms for C3) and removed; see DESIGN.md section 4.1 and the git history.
#pragma once

#ifndef YAWB_WS_LB
#define YAWB_WS_LB 192
#endif
constexpr int WS_LB = YAWB_WS_LB;  // entries per list buffer

struct __align__(16) ChanCtl {
    double ou, ov, ot;                            // centre of the tile box = origin of the staged vectors
    double umin, umax, vmin, vmax, tmin, tmax;    // tile box in the frame of patch p1
    int p1;
    int b_lo, b_hi;
    int st_cb, st_ncand, st_t0, st_have, st_ncombo;  // gather progress of the current item
};

template <bool WEIGHTED>
struct Channel {
    ChanCtl *ctl;
    Cand *list0;              // `nbuf` buffers of WS_LB entries each, addressed arithmetically (no
    double *lw0;              // runtime-indexed pointer arrays: those would live in local memory)
    int *lidx0;
    unsigned short *lbin0;
    __device__ __forceinline__ Cand *list(int buf) const { return list0 + buf * WS_LB; }
    __device__ __forceinline__ double *lw(int buf) const { return lw0 ? lw0 + buf * WS_LB : nullptr; }
    __device__ __forceinline__ int *lidx(int buf) const { return lidx0 + buf * WS_LB; }
    __device__ __forceinline__ unsigned short *lbin(int buf) const { return lbin0 + buf * WS_LB; }
    float4 *binrec;
    float2 *binthr;
    unsigned long long *acc;
    double *accw, *histw;
    int *bin_iv0, *bin_iu, *cstart, *rs0, *rpre, *cbin;
    unsigned *hist;
    float *cum;        // MULTI && SAT: ramp offsets of the current item, [n_bins][CUM_EDGES]
    unsigned *cumtot;  // MULTI && SAT: cumulative counts of one segment, [CUM_EDGES]
    unsigned short *seg;
};

// `acc_global`: the per-warp accumulators [n_bins][nsub] are NOT kept in shared memory (many sub-bins: they
// would cost most of the occupancy); the histogram of every z-bin segment goes straight to global atomics.
__host__ __device__ inline size_t ws_chan_bytes(bool weighted, bool multi, int n_bins, int nsub, int nbuf = 2,
                                                bool acc_global = false) {
    size_t b = sizeof(ChanCtl);
    b += (size_t)nbuf * WS_LB * sizeof(Cand) + (size_t)n_bins * sizeof(float4);
    if (weighted) b += (size_t)nbuf * WS_LB * sizeof(double);
    if (!acc_global) {
        b += (size_t)n_bins * nsub * sizeof(unsigned long long);
        if (weighted) b += (size_t)n_bins * nsub * sizeof(double);
    }
    if (multi && weighted) b += (size_t)nsub * sizeof(double);
    b += (size_t)n_bins * sizeof(float2);
    b += (size_t)nbuf * WS_LB * sizeof(int);
    b += (size_t)(3 * n_bins + 1) * sizeof(int) + 3 * CCAP * sizeof(int);
    if (multi) b += (size_t)nsub * sizeof(unsigned) + ((size_t)n_bins + 1) * CUM_EDGES * sizeof(float);
    b += (size_t)(nbuf + 1) * WS_LB * sizeof(unsigned short);
    return (b + 15) & ~(size_t)15;
}

template <bool WEIGHTED>
__device__ __forceinline__ void ws_carve(Channel<WEIGHTED> &C, unsigned char *p, bool multi, int n_bins, int nsub,
                                         int nbuf = 2, bool acc_global = false) {
    const size_t nacc = acc_global ? 0 : (size_t)n_bins * nsub;
    C.ctl = (ChanCtl *)p; p += sizeof(ChanCtl);
    C.list0 = (Cand *)p; p += (size_t)nbuf * WS_LB * sizeof(Cand);
    C.binrec = (float4 *)p; p += (size_t)n_bins * sizeof(float4);
    C.lw0 = nullptr; C.accw = nullptr; C.histw = nullptr; C.hist = nullptr;
    if (WEIGHTED) { C.lw0 = (double *)p; p += (size_t)nbuf * WS_LB * sizeof(double); }
    C.acc = (unsigned long long *)p; p += nacc * sizeof(unsigned long long);
    if (WEIGHTED) { C.accw = (double *)p; p += nacc * sizeof(double); }
    if (multi && WEIGHTED) { C.histw = (double *)p; p += (size_t)nsub * sizeof(double); }
    C.binthr = (float2 *)p; p += (size_t)n_bins * sizeof(float2);
    C.lidx0 = (int *)p; p += (size_t)nbuf * WS_LB * sizeof(int);
    C.bin_iv0 = (int *)p; p += (size_t)n_bins * sizeof(int);
    C.bin_iu = (int *)p; p += (size_t)n_bins * sizeof(int);
    C.cstart = (int *)p; p += (size_t)(n_bins + 1) * sizeof(int);
    C.rs0 = (int *)p; p += CCAP * sizeof(int);
    C.rpre = (int *)p; p += CCAP * sizeof(int);
    C.cbin = (int *)p; p += CCAP * sizeof(int);
    C.cum = nullptr; C.cumtot = nullptr;
    if (multi) {
        C.hist = (unsigned *)p; p += (size_t)nsub * sizeof(unsigned);
        C.cum = (float *)p; p += (size_t)n_bins * CUM_EDGES * sizeof(float);
        C.cumtot = (unsigned *)p; p += (size_t)CUM_EDGES * sizeof(unsigned);
    }
    C.lbin0 = (unsigned short *)p; p += (size_t)nbuf * WS_LB * sizeof(unsigned short);
    C.seg = (unsigned short *)p;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- step 1 for a fresh item ------------------------------------------------
template <bool WEIGHTED, bool MULTI, bool SAT>
__device__ __forceinline__ void ws_begin_item(const FastParams &P, const Channel<WEIGHTED> &C, int lane) {
    ChanCtl &ctl = *C.ctl;
    const int b_lo = ctl.b_lo, b_hi = ctl.b_hi;
    const SGrid G = P.sgrid[ctl.p1];
    const double umin = ctl.umin, umax = ctl.umax, vmin = ctl.vmin, vmax = ctl.vmax;
    const double eu = 0.5 * (umax - umin), ev = 0.5 * (vmax - vmin), et = 0.5 * (ctl.tmax - ctl.tmin);
    int carry = 0;
    for (int b0 = b_lo; b0 < b_hi; b0 += 32) {
        const int b = b0 + lane;
        int nrows = 0;
        if (b < b_hi) {
            const BinPar bp = P.binpar[b];
            if (!bp.empty) {
                // query box = tile box grown by the search radius (sound: |du|,|dv|,|dt| <= chord)
                const double fu0 = floor((umin - bp.rmax - G.u0) * G.inv_c), fu1 = floor((umax + bp.rmax - G.u0) * G.inv_c);
                const double fv0 = floor((vmin - bp.rmax - G.v0) * G.inv_c), fv1 = floor((vmax + bp.rmax - G.v0) * G.inv_c);
                if (!(fu1 < 0.0 || fv1 < 0.0 || fu0 > (double)(G.gu - 1) || fv0 > (double)(G.gv - 1))) {
                    const int iu0 = (int)fmax(fu0, 0.0), iv0 = (int)fmax(fv0, 0.0);
                    const int iu1 = (int)fmin(fu1, (double)(G.gu - 1)), iv1 = (int)fmin(fv1, (double)(G.gv - 1));
                    nrows = iv1 - iv0 + 1;
                    C.bin_iv0[b] = iv0;
                    C.bin_iu[b] = iu0 | (iu1 << 16);
                }
                // half extents rounded up; they bound every staged vector, hence the FP32 error of u
                const float hx = (float)(eu + bp.rmax) * 1.000001f, hy = (float)(ev + bp.rmax) * 1.000001f,
                            hz = (float)(et + bp.rmax) * 1.000001f;
                const float m2 = hx * hx + hy * hy + hz * hz;
                // |u_fp32 - (d2_ref - mid)| <= 29 eps32 M^2 + 7 eps32 mid (coordinate rounding 8, |r|^2 3,
                // |s|^2 - mid 4 + 1, first add 2 + 1, three FMAs 12 + 3, mid/h rounding 2)
                const float eps = EPS32 * (32.0f * m2 + 8.0f * bp.mid) * 1.0001f;
                C.binrec[b] = make_float4(hx, hy, hz, bp.mid);
                if (MULTI && SAT) {
                    // cumulative counts per edge: v_k = sat(K (e_k - mid - u) + 1/2); the float copy of
                    // e_k - mid adds at most eps32 |e_k - mid| to the error of u
                    const float K = 0.4f / (eps + 4.0f * EPS32 * (float)bp.hi);
                    C.binthr[b] = make_float2(-K, 0.f);
                    const int ne = P.n_edges;
                    for (int k = 0; k < CUM_EDGES; ++k)
                        C.cum[b * CUM_EDGES + k] =
                            k < ne ? fmaf(K, (float)(P.r2[(size_t)b * ne + k] - (double)bp.mid), 0.5f) : -1.0e30f;
                } else if (MULTI) {
                    C.binthr[b] = make_float2(bp.h + eps, eps + 4.0f * EPS32 * (float)bp.hi);
                } else if (SAT) {
                    const float K = 0.4f / eps;  // undecidable tests land in v = [0.1, 0.9]: v (1 - v) >= 0.09
                    C.binthr[b] = make_float2(-K, 0.5f + bp.h * K);
                } else {
                    C.binthr[b] = make_float2(bp.h - eps, bp.h + eps);
                }
            }
        }
        int incl = nrows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        if (b < b_hi) C.cstart[b + 1] = carry + incl;
        carry += __shfl_sync(FULL, incl, 31);
    }
    if (lane == 0) {
        C.cstart[b_lo] = 0;
        ctl.st_cb = 0;
        ctl.st_ncand = 0;
        ctl.st_t0 = 0;
        ctl.st_have = 0;
        ctl.st_ncombo = carry;
    }
    __syncwarp();
}

// ---- steps 2 + 3: produce one list buffer; returns L, sets `done` when the item is exhausted --------
template <bool WEIGHTED, int SUB>
__device__ __forceinline__ int ws_fill(const FastParams &P, const Channel<WEIGHTED> &C, int buf, int lane, bool &done) {
    ChanCtl &ctl = *C.ctl;
    const int b_lo = ctl.b_lo, b_hi = ctl.b_hi;
    const SGrid G = P.sgrid[ctl.p1];
    const double ou = ctl.ou, ov = ctl.ov, ot = ctl.ot;
    int cb = ctl.st_cb, n_cand = ctl.st_ncand, t0 = ctl.st_t0;
    bool have_batch = ctl.st_have != 0;
    const int n_combo = ctl.st_ncombo;
    Cand *list = C.list(buf);
    int *lidx = C.lidx(buf);
    unsigned short *lbin = C.lbin(buf);
    double *lwb = C.lw(buf);

    int cur = 0, run_lo = 0, run_hi = 0;
    auto seek = [&](int t) {  // position the per-lane run cursor for flat index t (start of a fill / batch)
        const int nb = min(CCAP, n_combo - cb);
        int lo = 0, hi = nb - 1;  // first run with rpre > t
        while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (C.rpre[m] > t) hi = m; else lo = m + 1;
        }
        cur = lo;
        run_lo = cur ? C.rpre[cur - 1] : 0;
        run_hi = C.rpre[cur];
    };
    if (have_batch) seek(min(t0 + lane, max(n_cand - 1, 0)));

    int L = 0;
    while (L <= WS_LB - 32 * SUB) {
        if (!have_batch) {
            if (cb >= n_combo) break;
            // ---- step 2: one lane per (z-bin, cell row): the run of candidate rows it covers ----
            const int nb = min(CCAP, n_combo - cb);
            int running = 0;
            for (int k0 = 0; k0 < nb; k0 += 32) {
                const int k = k0 + lane;
                int cnt = 0, s0 = 0, b = 0;
                if (k < nb) {
                    const int c = cb + k;
                    int lo = b_lo, hi = b_hi;  // last z-bin with cstart[b] <= c
                    while (hi - lo > 1) {
                        const int m = (lo + hi) >> 1;
                        if (C.cstart[m] <= c) lo = m; else hi = m;
                    }
                    b = lo;
                    const int iv = C.bin_iv0[b] + (c - C.cstart[b]);
                    const int iu = C.bin_iu[b];
                    const long long row = G.cell_base + ((long long)b * G.gv + iv) * G.gu;
                    s0 = P.cell_start[row + (iu & 0xffff)];
                    cnt = P.cell_start[row + (iu >> 16) + 1] - s0;
                    // pull the run's coordinates towards the SM while the scan below and the previous
                    // list are being worked on (runs are short: one or two 128-byte lines per array)
                    for (int q = 0; q < cnt; q += 16) {
                        prefetch_l2(P.su + s0 + q);
                        prefetch_l2(P.sv + s0 + q);
                        prefetch_l2(P.st + s0 + q);
                    }
                }
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                if (k < nb) {
                    C.rs0[k] = s0;
                    C.rpre[k] = running + incl;
                    C.cbin[k] = b;
                }
                running += __shfl_sync(FULL, incl, 31);
            }
            n_cand = running;
            t0 = 0;
            have_batch = true;
            __syncwarp();
            cur = 0;
            run_lo = 0;
            run_hi = C.rpre[0];
        }
        if (t0 >= n_cand) {
            have_batch = false;
            cb += CCAP;
            __syncwarp();
            continue;
        }
        // ---- step 3: flattened gather, cull against the z-bin's box, stage as float4 ----
        // SUB sub-batches per iteration: all their loads are in flight before the first is used
        int ci[SUB], cbn[SUB];
        bool okk[SUB];
        double lu[SUB], lv[SUB], lt[SUB];
#pragma unroll
        for (int h = 0; h < SUB; ++h) {
            const int t = t0 + 32 * h + lane;
            okk[h] = t < n_cand;
            ci[h] = 0;
            cbn[h] = 0;
            lu[h] = lv[h] = lt[h] = 0.0;
            if (okk[h]) {
                while (t >= run_hi) {  // runs are consumed in order; empty runs are skipped
                    run_lo = run_hi;
                    ++cur;
                    run_hi = C.rpre[cur];
                }
                ci[h] = C.rs0[cur] + (t - run_lo);
                cbn[h] = C.cbin[cur];
                lu[h] = P.su[ci[h]];
                lv[h] = P.sv[ci[h]];
                lt[h] = P.st[ci[h]];
            }
        }
        t0 += 32 * SUB;
#pragma unroll
        for (int h = 0; h < SUB; ++h) {
            bool ok = okk[h];
            float fx = 0.f, fy = 0.f, fz = 0.f, mid = 0.f;
            if (ok) {
                fx = (float)(lu[h] - ou);
                fy = (float)(lv[h] - ov);
                fz = (float)(lt[h] - ot);
                const float4 rec4 = C.binrec[cbn[h]];
                mid = rec4.w;
                ok = fabsf(fx) <= rec4.x && fabsf(fy) <= rec4.y && fabsf(fz) <= rec4.z;
            }
            const unsigned m = __ballot_sync(FULL, ok);
            if (ok) {
                const int pos = L + __popc(m & ((1u << lane) - 1u));
                const float sn = fx * fx + fy * fy + fz * fz;
                const float ax = -2.0f * fx, ay = -2.0f * fy, az = -2.0f * fz, aw = sn - mid;
                list[pos].a = make_float4(ax, ax, ay, ay);
                list[pos].b = make_float4(az, az, aw, aw);
                lidx[pos] = ci[h];
                lbin[pos] = (unsigned short)cbn[h];
                if (WEIGHTED) lwb[pos] = P.sw ? P.sw[ci[h]] : 1.0;
            }
            L += __popc(m);
        }
    }
    done = !have_batch && cb >= n_combo;
    if (lane == 0) {
        ctl.st_cb = cb;
        ctl.st_ncand = n_cand;
        ctl.st_t0 = t0;
        ctl.st_have = have_batch ? 1 : 0;
    }
    __syncwarp();
    return L;
}

// ---- run the pair tests on one staged list: one pass of phase 2 per z-bin segment ---------------------
template <bool WEIGHTED, bool MULTI, bool SAT>
__device__ __forceinline__ void ws_consume(const FastParams &P, const Channel<WEIGHTED> &C, int buf, int L,
                                           const float2 (&rx)[HPL], const float2 (&ry)[HPL],
                                           const float2 (&rz)[HPL], const float2 (&rn)[HPL],
                                           const Tile &tl, int lane, int nsub, unsigned &n_recheck, int cur_pair,
                                           const double (&rwt)[YAWB_RPL]) {
    WarpSmem<WEIGHTED> S;  // view of the current buffer for the shared phase-2 code
    S.list = C.list(buf); S.lw = C.lw(buf); S.lidx = C.lidx(buf); S.lbin = C.lbin(buf);
    S.hist = C.hist; S.histw = C.histw; S.acc = C.acc; S.accw = C.accw;
    S.cum = C.cum; S.cumtot = C.cumtot;
    // segment table: positions where the z-bin changes (entries arrive sorted by z-bin)
    int n_seg = 0;
    for (int base = 0; base < L; base += 32) {
        const int e = base + lane;
        const bool start = e < L && (e == 0 || S.lbin[e] != S.lbin[e - 1]);
        const unsigned ms = __ballot_sync(FULL, start);
        if (start) C.seg[n_seg + __popc(ms & ((1u << lane) - 1u))] = (unsigned short)e;
        n_seg += __popc(ms);
    }
    __syncwarp();
    for (int sg = 0; sg < n_seg; ++sg) {
        const int ea = C.seg[sg];
        const int eb = sg + 1 < n_seg ? (int)C.seg[sg + 1] : L;
        const int b = S.lbin[ea];
        const float2 thr = C.binthr[b];
        if (MULTI && SAT && !WEIGHTED) {
            if constexpr (!WEIGHTED)
                phase2_cumul(P, S, ea, eb, rx, ry, rz, rn, thr.x, tl, lane, b, n_recheck);
        } else if (MULTI) {
            for (int k = lane; k < nsub; k += 32) {
                S.hist[k] = 0u;
                if (WEIGHTED) S.histw[k] = 0.0;
            }
            __syncwarp();
            phase2_multi<WEIGHTED>(P, S, ea, eb, rx, ry, rz, rn, thr.x, thr.y, C.binrec[b].w, tl, lane, b,
                                   n_recheck);
            __syncwarp();
            if (P.acc_global) {  // straight to the result: one atomic per non-empty sub-bin of the segment
                const size_t o = ((size_t)cur_pair * P.n_bins + b) * nsub;
                for (int k = lane; k < nsub; k += 32) {
                    if (S.hist[k]) atomicAdd(&P.out_cnt[o + k], (unsigned long long)S.hist[k]);
                    if (WEIGHTED && S.histw[k] != 0.0) atomicAdd(&P.out_w[o + k], S.histw[k]);
                }
            } else {
                for (int k = lane; k < nsub; k += 32) {
                    S.acc[(size_t)b * nsub + k] += S.hist[k];
                    if (WEIGHTED) S.accw[(size_t)b * nsub + k] += S.histw[k];
                }
            }
            __syncwarp();
        } else {
            unsigned cnt_total = 0;
            double w_total = 0.0;
            phase2_single<WEIGHTED, SAT && !WEIGHTED>(P, S, ea, eb, rx, ry, rz, rn, thr.x, thr.y, tl, lane,
                                                      P.binpar[b].lo, P.binpar[b].hi, cnt_total, w_total,
                                                      n_recheck, rwt);
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

// ---- the kernel: every warp gathers for itself, then tests.  All gather state lives in the channel's
// shared-memory control block, so the registers of the hot loop are not shared with long-lived scalars
// of the gather phase.
template <bool WEIGHTED, bool MULTI, bool SAT>
__global__ void __launch_bounds__(YAWB_WARPS * 32, (WEIGHTED || (MULTI && SAT)) ? YAWB_MIN_CTAS_WEIGHTED : YAWB_MIN_CTAS)
    k_count_uni(const FastParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nsub = P.n_edges - 1;
    const bool acc_global = MULTI && P.acc_global;
    const int nacc = acc_global ? 0 : P.n_bins * nsub;  // accumulators kept in shared memory
    Channel<WEIGHTED> C;
    ws_carve<WEIGHTED>(C, smem_raw + (size_t)warp * ws_chan_bytes(WEIGHTED, MULTI, P.n_bins, nsub, 1, acc_global), MULTI,
                       P.n_bins, nsub, 1, acc_global);
    ChanCtl &ctl = *C.ctl;
    for (int k = lane; k < nacc; k += 32) {
        C.acc[k] = 0ull;
        if (WEIGHTED) C.accw[k] = 0.0;
    }
    __syncwarp();

    int cur_pair = -1;
    unsigned long long n_tests = 0;
    unsigned n_recheck = 0;
    const long long n_live = (long long)P.counters[4];  // written by k_plan

    auto flush_pair = [&]() {
        if (cur_pair < 0) return;
        __syncwarp();
        for (int k = lane; k < nacc; k += 32) {
            const unsigned long long c = C.acc[k];
            if (c) {
                atomicAdd(&P.out_cnt[(size_t)cur_pair * nacc + k], c);
                C.acc[k] = 0ull;
            }
            if (WEIGHTED) {
                const double w = C.accw[k];
                if (w != 0.0) {
                    atomicAdd(&P.out_w[(size_t)cur_pair * nacc + k], w);
                    C.accw[k] = 0.0;
                }
            }
        }
        __syncwarp();
    };

    long long grab_lo = 0, grab_hi = 0;
    while (true) {
        if (grab_lo >= grab_hi) {
            if (lane == 0) grab_lo = (long long)atomicAdd(&P.counters[0], (unsigned long long)GRAB);
            grab_lo = __shfl_sync(FULL, grab_lo, 0);
            if (grab_lo >= n_live) break;
            grab_hi = min(grab_lo + GRAB, n_live);
        }
        const int2 rec = P.live[grab_lo++];
        if (rec.x != cur_pair) {
            flush_pair();
            cur_pair = rec.x;
        }
        const int p1 = P.pair_i[cur_pair];
        const Tile tl = P.tiles[rec.y];
        const PatchFrame &F = P.sframe[p1];
        if (grab_lo < grab_hi) {  // next item of this grab: start pulling its rows into L2 now
            const Tile nx = P.tiles[P.live[grab_lo].y];
            for (int k = lane * 16; k < nx.count; k += 512) {
                prefetch_l2(P.rx + nx.start + k);
                prefetch_l2(P.ry + nx.start + k);
                prefetch_l2(P.rz + nx.start + k);
            }
        }
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
        if (lane == 0) {
            ctl.ou = ou; ctl.ov = ov; ctl.ot = ot;
            ctl.umin = umin; ctl.umax = umax; ctl.vmin = vmin; ctl.vmax = vmax; ctl.tmin = tmin; ctl.tmax = tmax;
            ctl.p1 = p1;
            ctl.b_lo = tl.bin >= 0 ? tl.bin : 0;
            ctl.b_hi = tl.bin >= 0 ? tl.bin + 1 : P.n_bins;
        }
        __syncwarp();
        float2 rx[HPL], ry[HPL], rz[HPL], rn[HPL];  // rows (2k, 2k+1) of the lane share one register pair
        double rwt[YAWB_RPL];                       // their weights (weighted kernels only)
#pragma unroll
        for (int r = 0; r < YAWB_RPL; ++r) {
            const int k = lane + 32 * r;
            float x = FAR, y = FAR, z = FAR, n = 3.0f * FAR * FAR;  // padding rows are never in range
            rwt[r] = 0.0;
            if (k < tl.count) {
                const int j = tl.start + k;
                if (WEIGHTED) rwt[r] = P.rw ? P.rw[j] : 1.0;
                const double dx = P.rx[j] - c0, dy = P.ry[j] - c1, dz = P.rz[j] - c2;
                x = (float)(dx * a0 + dy * a1 + dz * a2 - ou);
                y = (float)(dx * g0 + dy * g1 + dz * g2 - ov);
                z = (float)(dx * c0 + dy * c1 + dz * c2 - ot);
                n = x * x + y * y + z * z;
            }
            if (r & 1) { rx[r >> 1].y = x; ry[r >> 1].y = y; rz[r >> 1].y = z; rn[r >> 1].y = n; }
            else { rx[r >> 1].x = x; ry[r >> 1].x = y; rz[r >> 1].x = z; rn[r >> 1].x = n; }
        }
        ws_begin_item<WEIGHTED, MULTI, SAT>(P, C, lane);
        bool done = false;
        while (!done) {
            const int L = ws_fill<WEIGHTED, 2>(P, C, 0, lane, done);
            __syncwarp();
#ifdef YAWB_DEBUG_MODES  // development builds: YAWB_DEBUG_MODE=1 skips the pair tests, 2 runs them twice (phase timing)
            if (L > 0 && P.debug != 1) {
                ws_consume<WEIGHTED, MULTI, SAT>(P, C, 0, L, rx, ry, rz, rn, tl, lane, nsub, n_recheck, cur_pair, rwt);
                if (P.debug == 2)
                    ws_consume<WEIGHTED, MULTI, SAT>(P, C, 0, L, rx, ry, rz, rn, tl, lane, nsub, n_recheck, cur_pair, rwt);
            }
#else
            if (L > 0) ws_consume<WEIGHTED, MULTI, SAT>(P, C, 0, L, rx, ry, rz, rn, tl, lane, nsub, n_recheck, cur_pair, rwt);
#endif
            n_tests += (unsigned long long)L * (unsigned long long)tl.count;
            __syncwarp();
        }
    }
    flush_pair();
    if (lane == 0 && n_tests) atomicAdd(&P.counters[1], n_tests);
    const unsigned rc = __reduce_add_sync(FULL, n_recheck);
    if (lane == 0 && rc) atomicAdd(&P.counters[2], (unsigned long long)rc);
}
