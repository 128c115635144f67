// Probe: ceiling of the FP32 pair-test loop on B200 without any gather (candidates already staged in
// shared memory).  Variants of the loop structure around the same 7-op packed test:
//   V0  round-1 structure: chunks of 8 (unrolled) + rolled tail, ballot per chunk, REDUX + lane-0 add per segment
//   V1  segment descriptors (one LDS.128, prefetched), groups of 4 + switch tail, one check per <= 16
//       candidates, REDUX result consumed one segment later
//   V2  V1 with 16-byte candidates (x, y, z, w) duplicated into register pairs with MOVs
// Prints achieved pair tests/s for several warps-per-SM settings and segment lengths.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o test_loop test_loop.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr unsigned FULL = 0xffffffffu;
constexpr int RPL = 8, HPL = 4;
constexpr int LB = 192;

struct __align__(16) Cand {
    float4 a, b;
};
struct __align__(16) SegDesc {
    int ea, eb;
    float K, C;
};

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));               \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

__device__ __forceinline__ void test_one(const float4 ca, const float4 cb, const float2 (&rx)[HPL],
                                         const float2 (&ry)[HPL], const float2 (&rz)[HPL], const float2 (&rn)[HPL],
                                         float ta, float tb, float2 &acc_a, float2 &acc_b) {
    const float2 sx = make_float2(ca.x, ca.y), sy = make_float2(ca.z, ca.w);
    const float2 sz = make_float2(cb.x, cb.y), sw = make_float2(cb.z, cb.w);
#pragma unroll
    for (int k = 0; k < HPL; ++k) {
        float2 u = __fadd2_rn(rn[k], sw);
        u = __ffma2_rn(rx[k], sx, u);
        u = __ffma2_rn(ry[k], sy, u);
        u = __ffma2_rn(rz[k], sz, u);
        float2 v;
        v.x = __saturatef(fmaf(fabsf(u.x), ta, tb));
        v.y = __saturatef(fmaf(fabsf(u.y), ta, tb));
        acc_a = __fadd2_rn(acc_a, v);
        acc_b = __ffma2_rn(v, v, acc_b);
    }
}

// sphere form: |r|^2 folded into the candidate's operands, no FADD2
__device__ __forceinline__ void test_one_nf(const float4 ca, const float4 cb, const float2 (&rx)[HPL],
                                            const float2 (&ry)[HPL], const float2 (&rz)[HPL],
                                            float ta, float tb, float2 &acc_a, float2 &acc_b) {
    const float2 sx = make_float2(ca.x, ca.y), sy = make_float2(ca.z, ca.w);
    const float2 sz = make_float2(cb.x, cb.y), sw = make_float2(cb.z, cb.w);
#pragma unroll
    for (int k = 0; k < HPL; ++k) {
        float2 u = __ffma2_rn(rx[k], sx, sw);
        u = __ffma2_rn(ry[k], sy, u);
        u = __ffma2_rn(rz[k], sz, u);
        float2 v;
        v.x = __saturatef(fmaf(fabsf(u.x), ta, tb));
        v.y = __saturatef(fmaf(fabsf(u.y), ta, tb));
        acc_a = __fadd2_rn(acc_a, v);
        acc_b = __ffma2_rn(v, v, acc_b);
    }
}

// ALU-assisted variant: distance + ramp on the FMA pipe (4 packed + 2 scalar FFMA.SAT per row pair); the hits are
// counted by adding the bit patterns of v (0 or 0x3f800000) with IADD3, undecided tests (0 < v < 1) are detected
// with an unsigned running minimum of bits(v) - 1 (VIADDMNMX), both on the integer pipe.
__device__ __forceinline__ void test_one_alu(const float4 ca, const float4 cb, const float2 (&rx)[HPL],
                                             const float2 (&ry)[HPL], const float2 (&rz)[HPL], const float2 (&rn)[HPL],
                                             float ta, float tb, unsigned &cnt, unsigned &umin) {
    const float2 sx = make_float2(ca.x, ca.y), sy = make_float2(ca.z, ca.w);
    const float2 sz = make_float2(cb.x, cb.y), sw = make_float2(cb.z, cb.w);
#pragma unroll
    for (int k = 0; k < HPL; ++k) {
        float2 u = __fadd2_rn(rn[k], sw);
        u = __ffma2_rn(rx[k], sx, u);
        u = __ffma2_rn(ry[k], sy, u);
        u = __ffma2_rn(rz[k], sz, u);
        const unsigned bx = __float_as_uint(__saturatef(fmaf(fabsf(u.x), ta, tb)));
        const unsigned by = __float_as_uint(__saturatef(fmaf(fabsf(u.y), ta, tb)));
        cnt = cnt + bx + by;
        umin = min(umin, bx + 0xffffffffu);
        umin = min(umin, by + 0xffffffffu);
    }
}

// scalar forms: every FP32 operation is a plain FFMA / FADD (one row per instruction)
template <bool ALU>
__device__ __forceinline__ void test_one_sc(const float4 ca, const float4 cb, const float2 (&rx)[HPL],
                                            const float2 (&ry)[HPL], const float2 (&rz)[HPL], const float2 (&rn)[HPL],
                                            float ta, float tb, float2 &acc_a, float2 &acc_b, unsigned &cnt, unsigned &umin) {
    const float sx = ca.x, sy = ca.z, sz = cb.x, sw = cb.z;
#pragma unroll
    for (int k = 0; k < HPL; ++k) {
        float ux = rn[k].x + sw, uy = rn[k].y + sw;
        ux = fmaf(rx[k].x, sx, ux); uy = fmaf(rx[k].y, sx, uy);
        ux = fmaf(ry[k].x, sy, ux); uy = fmaf(ry[k].y, sy, uy);
        ux = fmaf(rz[k].x, sz, ux); uy = fmaf(rz[k].y, sz, uy);
        const float vx = __saturatef(fmaf(fabsf(ux), ta, tb)), vy = __saturatef(fmaf(fabsf(uy), ta, tb));
        if (ALU) {
            const unsigned bx = __float_as_uint(vx), by = __float_as_uint(vy);
            cnt = cnt + bx + by;
            umin = min(umin, bx + 0xffffffffu);
            umin = min(umin, by + 0xffffffffu);
        } else {
            acc_a.x += vx; acc_a.y += vy;
            acc_b.x = fmaf(vx, vx, acc_b.x); acc_b.y = fmaf(vy, vy, acc_b.y);
        }
    }
}


constexpr unsigned ONE_BITS = 0x3f800000u;
__device__ __forceinline__ unsigned lop3_or3(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xfe;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned lop3_or_and(unsigned a, unsigned b, unsigned c) {  // a | (b & c)
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xf8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// Dual-pipe variants: distance + ramp on the FMA pipe (scalar FADD + 3 FFMA + FFMA.SAT per test), everything else on
// the ALU pipe: an undecided test (0 < v < 1) is detected as min(bits, bits ^ 0x3f800000) != 0 (XOR + VIMNMX), OR-ed
// into one word (LOP3 per two tests).
//   MODE 6: hits summed on the FMA pipe (FADD per test)
//   MODE 7: hits collected as flag bits (bit 23 + row of the 7-bit exponent pattern of 1.0f, LOP3 per test), POPC + IADD per candidate
//   MODE 8: MODE 7 without the FADD (|r|^2 folded away: points on a sphere)
template <int MODE, int SLOT>
__device__ __forceinline__ void test_one_dp(const float4 q, const float2 (&rx)[HPL], const float2 (&ry)[HPL],
                                            const float2 (&rz)[HPL], const float2 (&rn)[HPL], float ta, float tb, float &sum,
                                            unsigned &det, unsigned &cnt, unsigned &w8) {
    unsigned w = 0u, zprev = 0u;
#pragma unroll
    for (int k = 0; k < 2 * HPL; ++k) {
        const float x_ = (k & 1) ? rx[k >> 1].y : rx[k >> 1].x, y_ = (k & 1) ? ry[k >> 1].y : ry[k >> 1].x;
        const float z_ = (k & 1) ? rz[k >> 1].y : rz[k >> 1].x, n_ = (k & 1) ? rn[k >> 1].y : rn[k >> 1].x;
        float u = MODE == 8 ? q.w : n_ + q.w;
        u = fmaf(x_, q.x, u);
        u = fmaf(y_, q.y, u);
        u = fmaf(z_, q.z, u);
        const float v = __saturatef(fmaf(fabsf(u), ta, tb));
        const unsigned b = __float_as_uint(v);
        const unsigned z = min(b, b ^ ONE_BITS);
        if (k & 1) det = lop3_or3(det, zprev, z); else zprev = z;
        if (MODE == 6) sum += v;
        else if (k < 7) w = lop3_or_and(w, b, 1u << (23 + k));
        else w8 = lop3_or_and(w8, b, 1u << (23 + SLOT));
    }
    if (MODE != 6) cnt += __popc(w);
}

__device__ __noinline__ unsigned slow_recheck(const float *g, int e0, int lane) {
    return (unsigned)g[e0 * 32 + lane];  // stands for the FP64 recheck (never taken in the probe)
}

template <int VARIANT>
__global__ void __launch_bounds__(128) k_probe(const float *__restrict__ init, int seg_len, int reps,
                                               unsigned long long *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr size_t per_warp = LB * sizeof(Cand) + 64 * sizeof(SegDesc) + 64 * sizeof(unsigned long long);
    unsigned char *base = smem + warp * per_warp;
    Cand *list = (Cand *)base;
    float4 *list16 = (float4 *)base;
    SegDesc *seg = (SegDesc *)(base + LB * sizeof(Cand));
    unsigned long long *acc = (unsigned long long *)(seg + 64);

    // rows of the tile: a small clump around the origin; candidates spread over a few radii
    float2 rx[HPL], ry[HPL], rz[HPL], rn[HPL];
#pragma unroll
    for (int k = 0; k < HPL; ++k) {
        const float a = init[(threadIdx.x * 8 + 2 * k) & 1023], b = init[(threadIdx.x * 8 + 2 * k + 1) & 1023];
        const float c = init[(threadIdx.x * 8 + 2 * k + 517) & 1023], d = init[(threadIdx.x * 8 + 2 * k + 700) & 1023];
        rx[k] = make_float2(a * 1e-3f, b * 1e-3f);
        ry[k] = make_float2(c * 1e-3f, d * 1e-3f);
        rz[k] = make_float2((a + c) * 1e-5f, (b - d) * 1e-5f);
        rn[k] = make_float2(rx[k].x * rx[k].x + ry[k].x * ry[k].x + rz[k].x * rz[k].x,
                            rx[k].y * rx[k].y + ry[k].y * ry[k].y + rz[k].y * rz[k].y);
    }
    const float mid = 2.0e-6f, h = 1.9e-6f, eps = 1e-11f, K = 0.4f / eps;
    for (int e = lane; e < LB; e += 32) {
        const float x = init[(e * 7 + warp) & 1023] * 2e-3f, y = init[(e * 13 + 5) & 1023] * 2e-3f, z = 1e-5f;
        const float sn = x * x + y * y + z * z;
        if (VARIANT == 2 || (VARIANT >= 6 && VARIANT != 9)) {
            list16[e] = make_float4(-2.f * x, -2.f * y, -2.f * z, sn - mid);
        } else {
            list[e].a = make_float4(-2.f * x, -2.f * x, -2.f * y, -2.f * y);
            list[e].b = make_float4(-2.f * z, -2.f * z, sn - mid, sn - mid);
        }
    }
    const int n_seg = (LB + seg_len - 1) / seg_len;
    for (int s = lane; s < n_seg; s += 32) {
        seg[s].ea = s * seg_len;
        seg[s].eb = min(LB, (s + 1) * seg_len);
        seg[s].K = -K;
        seg[s].C = 0.5f + h * K;
    }
    for (int s = lane; s < 64; s += 32) acc[s] = 0ull;
    __syncwarp();

    unsigned long long total = 0;
    const long long clk0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        if (VARIANT == 0) {
            for (int sg = 0; sg < n_seg; ++sg) {
                const int ea = seg[sg].ea, eb = seg[sg].eb;
                const float ta = seg[sg].K, tb = seg[sg].C;
                unsigned cnt_total = 0;
                for (int e0 = ea; e0 < eb; e0 += 8) {
                    const int e1 = min(e0 + 8, eb);
                    float2 acc_a = make_float2(0.f, 0.f), acc_b = make_float2(0.f, 0.f);
                    if (e1 - e0 == 8) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) test_one(list[e0 + k].a, list[e0 + k].b, rx, ry, rz, rn, ta, tb, acc_a, acc_b);
                    } else {
                        for (int e = e0; e < e1; ++e) test_one(list[e].a, list[e].b, rx, ry, rz, rn, ta, tb, acc_a, acc_b);
                    }
                    const float sa = acc_a.x + acc_a.y, sb = acc_b.x + acc_b.y;
                    unsigned c = (unsigned)(sa + 0.5f);
                    unsigned flagged = __ballot_sync(FULL, sa != sb);
                    while (flagged) {
                        const int src = __ffs(flagged) - 1;
                        flagged &= flagged - 1;
                        const unsigned cx = slow_recheck(init, e0, lane);
                        if (lane == src) c = cx;
                    }
                    cnt_total += c;
                }
                const unsigned tot = __reduce_add_sync(FULL, cnt_total);
                if (lane == 0) acc[sg & 63] += tot;
            }
        } else {
            // V1 / V2
            SegDesc d = seg[0];
            unsigned pend_tot = 0;
            int pend_sg = 0;
            for (int sg = 0; sg < n_seg; ++sg) {
                const SegDesc cur = d;
                if (sg + 1 < n_seg) d = seg[sg + 1];  // prefetch the next descriptor
                unsigned cnt_total = 0;
                for (int c0 = cur.ea; c0 < cur.eb; c0 += 16) {
                    const int c1 = min(c0 + 16, cur.eb);
                    float2 acc_a = make_float2(0.f, 0.f), acc_b = make_float2(0.f, 0.f);
                    unsigned icnt = 0u, iumin = 0xffffffffu;
                    int e = c0;
                    float dsum = 0.f;
                    unsigned ddet = 0u, dcnt = 0u, dw8 = 0u;
                    if (VARIANT >= 6 && VARIANT < 9) {
#define TD(idx, slot) test_one_dp<VARIANT, slot>(list16[idx], rx, ry, rz, rn, cur.K, cur.C, dsum, ddet, dcnt, dw8)
                        for (; e + 4 <= c1; e += 4) {
                            TD(e, 0); TD(e + 1, 1); TD(e + 2, 2); TD(e + 3, 3);
                            if (VARIANT != 6) { dcnt += __popc(dw8); dw8 = 0u; }
                        }
                        switch (c1 - e) {
                            case 3: TD(e + 2, 2);
                            case 2: TD(e + 1, 1);
                            case 1: TD(e, 0);
                            default: break;
                        }
                        if (VARIANT != 6) dcnt += __popc(dw8);
#undef TD
                        e = c1;
                    }
                    auto T = [&](int idx) {
                        if (VARIANT == 9) {
                            test_one_nf(list[idx].a, list[idx].b, rx, ry, rz, cur.K, cur.C, acc_a, acc_b);
                        } else if (VARIANT == 10) {
                            const float4 q = list16[idx];
                            test_one_nf(make_float4(q.x, q.x, q.y, q.y), make_float4(q.z, q.z, q.w, q.w), rx, ry, rz, cur.K, cur.C, acc_a, acc_b);
                        } else if (VARIANT == 4) {
                            test_one_sc<false>(list[idx].a, list[idx].b, rx, ry, rz, rn, cur.K, cur.C, acc_a, acc_b, icnt, iumin);
                        } else if (VARIANT == 5) {
                            test_one_sc<true>(list[idx].a, list[idx].b, rx, ry, rz, rn, cur.K, cur.C, acc_a, acc_b, icnt, iumin);
                        } else if (VARIANT == 3) {
                            test_one_alu(list[idx].a, list[idx].b, rx, ry, rz, rn, cur.K, cur.C, icnt, iumin);
                        } else if (VARIANT == 2) {
                            const float4 q = list16[idx];
                            test_one(make_float4(q.x, q.x, q.y, q.y), make_float4(q.z, q.z, q.w, q.w), rx, ry, rz, rn, cur.K, cur.C,
                                     acc_a, acc_b);
                        } else {
                            test_one(list[idx].a, list[idx].b, rx, ry, rz, rn, cur.K, cur.C, acc_a, acc_b);
                        }
                    };
                    for (; e + 4 <= c1; e += 4) {
                        T(e); T(e + 1); T(e + 2); T(e + 3);
                    }
                    switch (c1 - e) {
                        case 3: T(e + 2);
                        case 2: T(e + 1);
                        case 1: T(e);
                        default: break;
                    }
                    const float sa = acc_a.x + acc_a.y, sb = acc_b.x + acc_b.y;
                    unsigned c = (unsigned)(sa + 0.5f);
                    bool bad = sa != sb;
                    if (VARIANT >= 6 && VARIANT < 9) {
                        c = VARIANT == 6 ? (unsigned)(dsum + 0.5f) : dcnt;
                        bad = ddet != 0u;
                    }
                    if (VARIANT == 3 || VARIANT == 5) {
                        c = ((icnt >> 23) * 383u) & 511u;  // icnt = 127 * n << 23 (mod 2^32), 127 * 383 = 1 (mod 512)
                        bad = iumin < 0x3f7fffffu;          // some v strictly between 0 and 1
                    }
                    unsigned flagged = __ballot_sync(FULL, bad);
                    while (flagged) {
                        const int src = __ffs(flagged) - 1;
                        flagged &= flagged - 1;
                        const unsigned cx = slow_recheck(init, c0, lane);
                        if (lane == src) c = cx;
                    }
                    cnt_total += c;
                }
                if (lane == 0) acc[pend_sg & 63] += pend_tot;  // the previous segment's total (REDUX long done)
                pend_tot = __reduce_add_sync(FULL, cnt_total);
                pend_sg = sg;
            }
            if (lane == 0) acc[pend_sg & 63] += pend_tot;
        }
        total += (unsigned long long)LB * 256ull;
    }
    __syncwarp();
    unsigned long long s = 0;
    for (int k = lane; k < 64; k += 32) s += acc[k];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    const long long clk1 = clock64();
    if (lane == 0) {
        atomicAdd(&out[0], total);
        atomicAdd(&out[1], s);
        atomicMax(&out[2], (unsigned long long)(clk1 - clk0));
    }
}

template <int VARIANT>
void run(const float *d_init, unsigned long long *d_out, int sms, double peak) {
    CK(cudaFuncSetAttribute(k_probe<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (int seg_len : {192, 14}) {
        for (int ctas : {2, 3, 4, 5}) {
            // sized so that exactly `ctas` CTAs fit on an SM (an even spread over the SMs is forced)
            const size_t smem = ((size_t)(227 * 1024 / ctas) - 1024) & ~(size_t)127;
            const int reps = 400;
            CK(cudaMemset(d_out, 0, 32));
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            k_probe<VARIANT><<<sms * ctas, 128, smem>>>(d_init, seg_len, 20, d_out);  // warm-up
            CK(cudaMemset(d_out, 0, 32));
            CK(cudaEventRecord(e0));
            k_probe<VARIANT><<<sms * ctas, 128, smem>>>(d_init, seg_len, reps, d_out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            unsigned long long h[3];
            CK(cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost));
            const double rate = (double)h[0] / (ms * 1e-3);
            printf("V%d seg_len %3d warps/SM %2d: %.3f ms, %.3e tests/s = %.1f %% of roofline (pairs %llu), %.2f cycles per candidate per SMSP, SM clock %.0f MHz\n", VARIANT,
                   seg_len, 4 * ctas, ms, rate, 100.0 * rate / peak, h[1], (double)h[2] / ((double)reps * LB * ctas),
                   (double)h[2] / (ms * 1e3));
        }
    }
}

int main(int argc, char **argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const double peak = sms * 128.0 * 1.965e9 / 6.0;
    printf("%s, %d SMs, roofline %.3e tests/s\n", prop.name, sms, peak);
    std::vector<float> init(1024);
    srand(1);
    for (auto &v : init) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *d_init;
    unsigned long long *d_out;
    CK(cudaMalloc(&d_init, 4096));
    CK(cudaMalloc(&d_out, 32));
    CK(cudaMemcpy(d_init, init.data(), 4096, cudaMemcpyHostToDevice));
    if (argc > 1) {  // single launch pair for profiling: ./test_loop <variant 1|3|4>
        const int v = atoi(argv[1]);
        const size_t smem = ((size_t)(227 * 1024 / 5) - 1024) & ~(size_t)127;
        if (v == 1) { CK(cudaFuncSetAttribute(k_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); k_probe<1><<<sms * 5, 128, smem>>>(d_init, 192, 100, d_out); }
        if (v == 3) { CK(cudaFuncSetAttribute(k_probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); k_probe<3><<<sms * 5, 128, smem>>>(d_init, 192, 100, d_out); }
        if (v == 4) { CK(cudaFuncSetAttribute(k_probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); k_probe<4><<<sms * 5, 128, smem>>>(d_init, 192, 100, d_out); }
        CK(cudaDeviceSynchronize());
        return 0;
    }
    run<1>(d_init, d_out, sms, peak);
    run<2>(d_init, d_out, sms, peak);
    run<9>(d_init, d_out, sms, peak);
    run<10>(d_init, d_out, sms, peak);
    return 0;
}
