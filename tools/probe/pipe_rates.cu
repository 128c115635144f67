// Probe: issue cost (cycles per warp instruction per SM sub-partition) of the FP32 instruction forms the
// pair test can be built from, on B200.  Each kernel runs 16 independent dependency chains of one form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                \
    do {                                                                     \
        cudaError_t e = (x);                                                 \
        if (e != cudaSuccess) {                                              \
            printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); \
            exit(1);                                                         \
        }                                                                    \
    } while (0)

constexpr int CH = 16;     // independent chains per thread
constexpr int ITER = 4096;

enum Op { FFMA_RRR, FFMA_IMM, FFMA_SAT_ABS, FADD_RR, FADD_SAT_ABS, FMUL_RR, FFMA2_OP, FADD2_OP, FMUL2_OP, MIX_TEST, MIX_IMM, MIX_FADD, FMNMX_OP,
          IADD3_OP, MIX_ALU, SC_FFMA, SC_FADD, SC_ALU, PK_FADD, PK_ALU, VIADDMNMX_OP, IADD3_3 };

template <int OP>
__global__ void __launch_bounds__(128) k_rate(float *out, float a, float b, long long *cycles) {
    float x[CH];
    float2 p[CH / 2];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = a * (float)(threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < CH / 2; ++i) p[i] = make_float2(x[2 * i], x[2 * i + 1]);
    const float2 a2 = make_float2(a, a * 1.5f), b2 = make_float2(b, b * 0.5f);
    int xi[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) xi[i] = threadIdx.x + i;
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if (OP == FFMA_RRR) {
#pragma unroll
            for (int i = 0; i < CH; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        } else if (OP == FFMA_IMM) {
#pragma unroll
            for (int i = 0; i < CH; ++i) asm volatile("fma.rn.f32 %0, %0, 0f3F7FFF00, %1;" : "+f"(x[i]) : "f"(b));
        } else if (OP == FFMA_SAT_ABS) {
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = __saturatef(fmaf(fabsf(x[i]), a, b));
        } else if (OP == FADD_RR) {
#pragma unroll
            for (int i = 0; i < CH; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        } else if (OP == FADD_SAT_ABS) {
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = __saturatef(b - fabsf(x[i]));
        } else if (OP == FMUL_RR) {
#pragma unroll
            for (int i = 0; i < CH; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
        } else if (OP == FFMA2_OP) {
#pragma unroll
            for (int i = 0; i < CH / 2; ++i) p[i] = __ffma2_rn(p[i], a2, b2);
        } else if (OP == FADD2_OP) {
#pragma unroll
            for (int i = 0; i < CH / 2; ++i) p[i] = __fadd2_rn(p[i], b2);
        } else if (OP == FMUL2_OP) {
#pragma unroll
            for (int i = 0; i < CH / 2; ++i) p[i] = __fmul2_rn(p[i], a2);
        } else if (OP == FMNMX_OP) {
#pragma unroll
            for (int i = 0; i < CH; ++i) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        } else if (OP == IADD3_OP) {
#pragma unroll
            for (int i = 0; i < CH; ++i) asm volatile("add.s32 %0, %0, %1;" : "+r"(xi[i]) : "r"(it));
        } else if (OP == VIADDMNMX_OP) {
#pragma unroll
            for (int i = 0; i < CH; ++i) xi[i] = (int)min((unsigned)xi[i], (unsigned)(it + i) + 0xffffffffu);
        } else if (OP == IADD3_3) {
#pragma unroll
            for (int i = 0; i < CH; ++i) xi[i] = xi[i] + xi[(i + 1) % CH] + it;
        } else if (OP == SC_FFMA || OP == SC_FADD || OP == SC_ALU || OP == PK_FADD || OP == PK_ALU) {
            // 4 row pairs (8 rows) against one candidate (a, b stand for its broadcast operands)
            float acc1 = 0.f, acc2 = 0.f;
            float2 pa = make_float2(0.f, 0.f), pb = make_float2(0.f, 0.f);
            unsigned icnt = 0u, imin = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < CH / 2; ++i) {
                float vx, vy;
                if (OP == PK_FADD || OP == PK_ALU) {
                    float2 u = __ffma2_rn(p[i], a2, b2);  // K (rn) + sw'
                    u = __ffma2_rn(p[(i + 1) % (CH / 2)], a2, u);
                    u = __ffma2_rn(p[(i + 2) % (CH / 2)], b2, u);
                    u = __ffma2_rn(p[(i + 3) % (CH / 2)], a2, u);
                    vx = __saturatef(b - fabsf(u.x));
                    vy = __saturatef(b - fabsf(u.y));
                } else {
                    float ux, uy;
                    if (OP == SC_FFMA) {
                        ux = x[2 * i] + b;
                        uy = x[2 * i + 1] + b;
                    } else {
                        ux = fmaf(x[2 * i], a, b);
                        uy = fmaf(x[2 * i + 1], a, b);
                    }
                    ux = fmaf(x[(2 * i + 2) % CH], a, ux);
                    uy = fmaf(x[(2 * i + 3) % CH], a, uy);
                    ux = fmaf(x[(2 * i + 4) % CH], b, ux);
                    uy = fmaf(x[(2 * i + 5) % CH], b, uy);
                    ux = fmaf(x[(2 * i + 6) % CH], a, ux);
                    uy = fmaf(x[(2 * i + 7) % CH], a, uy);
                    if (OP == SC_FFMA) {
                        vx = __saturatef(fmaf(fabsf(ux), a, b));
                        vy = __saturatef(fmaf(fabsf(uy), a, b));
                    } else {
                        vx = __saturatef(b - fabsf(ux));
                        vy = __saturatef(b - fabsf(uy));
                    }
                }
                if (OP == SC_ALU || OP == PK_ALU) {
                    const unsigned bx = __float_as_uint(vx), by = __float_as_uint(vy);
                    icnt = icnt + bx + by;
                    imin = min(imin, bx + 0xffffffffu);
                    imin = min(imin, by + 0xffffffffu);
                } else if (OP == PK_FADD) {
                    const float2 v = make_float2(vx, vy);
                    pa = __fadd2_rn(pa, v);
                    pb = __ffma2_rn(v, v, pb);
                } else {
                    acc1 += vx;
                    acc1 += vy;
                    acc2 = fmaf(vx, vx, acc2);
                    acc2 = fmaf(vy, vy, acc2);
                }
            }
            x[0] += acc1 + acc2 + pa.x + pa.y + pb.x + pb.y + (float)(icnt >> 23) + (float)(imin >> 20);
            p[0].x += x[0];
        } else if (OP == MIX_TEST || OP == MIX_IMM || OP == MIX_FADD || OP == MIX_ALU) {
            // the 8-instruction pair test on CH / 2 row pairs: FADD2, 3 FFMA2, 2 scalar decide ops, FADD2, FFMA2
            float2 acc_a = make_float2(0.f, 0.f), acc_b = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < CH / 2; ++i) {
                float2 u = __fadd2_rn(p[i], b2);
                u = __ffma2_rn(p[(i + 1) % (CH / 2)], a2, u);
                u = __ffma2_rn(p[(i + 2) % (CH / 2)], b2, u);
                u = __ffma2_rn(p[(i + 3) % (CH / 2)], a2, u);
                float2 v;
                if (OP == MIX_TEST) {
                    v.x = __saturatef(fmaf(fabsf(u.x), a, b));
                    v.y = __saturatef(fmaf(fabsf(u.y), a, b));
                } else if (OP == MIX_IMM) {
                    asm volatile("fma.rn.sat.f32 %0, %1, 0fBF800000, %2;" : "=f"(v.x) : "f"(fabsf(u.x)), "f"(b));
                    asm volatile("fma.rn.sat.f32 %0, %1, 0fBF800000, %2;" : "=f"(v.y) : "f"(fabsf(u.y)), "f"(b));
                } else if (OP == MIX_FADD) {
                    v.x = __saturatef(b - fabsf(u.x));
                    v.y = __saturatef(b - fabsf(u.y));
                } else {
                    // decide on the ALU pipe: clamp with min / max (FMNMX)
                    v.x = fminf(fmaxf(b - fabsf(u.x), 0.f), 1.f);
                    v.y = fminf(fmaxf(b - fabsf(u.y), 0.f), 1.f);
                }
                acc_a = __fadd2_rn(acc_a, v);
                acc_b = __ffma2_rn(v, v, acc_b);
            }
            p[0].x += acc_a.x + acc_b.y;
            p[1].y += acc_a.y + acc_b.x;
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i] + (float)xi[i];
#pragma unroll
    for (int i = 0; i < CH / 2; ++i) s += p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char *name, int instr_per_iter, float *d_out, long long *d_cyc, int sms) {
    CK(cudaFuncSetAttribute(k_rate<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (int ctas : {2, 4, 6}) {
        // dynamic shared memory sized so that exactly `ctas` CTAs fit on an SM: an even spread is forced
        const size_t smem = ((size_t)(227 * 1024 / ctas) - 1024) & ~(size_t)127;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        k_rate<OP><<<sms * ctas, 128, smem>>>(d_out, 1.0001f, 0.5f, d_cyc);
        CK(cudaEventRecord(e0));
        k_rate<OP><<<sms * ctas, 128, smem>>>(d_out, 1.0001f, 0.5f, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        long long cyc = 0;
        CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        // warps per SMSP = ctas; instructions issued per SMSP = ctas * ITER * instr_per_iter
        printf("%-14s warps/SMSP %d: %.3f cycles per warp instruction per SMSP (block 0 clock), %.3f by events at 1.965 GHz\n",
               name, ctas, (double)cyc / ((double)ctas * ITER * instr_per_iter),
               (double)ms * 1e-3 * 1.965e9 / ((double)ctas * ITER * instr_per_iter));
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float *d_out;
    long long *d_cyc;
    CK(cudaMalloc(&d_out, sms * 4 * 128 * sizeof(float)));
    CK(cudaMalloc(&d_cyc, 8));
    run<FFMA_RRR>("FFMA rrr", CH, d_out, d_cyc, sms);
    run<FFMA_IMM>("FFMA imm", CH, d_out, d_cyc, sms);
    run<FFMA_SAT_ABS>("FFMA.SAT |a|", CH, d_out, d_cyc, sms);
    run<FADD_RR>("FADD", CH, d_out, d_cyc, sms);
    run<FADD_SAT_ABS>("FADD.SAT -|a|", CH, d_out, d_cyc, sms);
    run<FMUL_RR>("FMUL", CH, d_out, d_cyc, sms);
    run<FFMA2_OP>("FFMA2", CH / 2, d_out, d_cyc, sms);
    run<FADD2_OP>("FADD2", CH / 2, d_out, d_cyc, sms);
    run<FMUL2_OP>("FMUL2", CH / 2, d_out, d_cyc, sms);
    run<FMNMX_OP>("FMNMX", CH, d_out, d_cyc, sms);
    run<IADD3_OP>("IADD", CH, d_out, d_cyc, sms);
    run<MIX_TEST>("test FFMA.SAT", CH / 2 * 8, d_out, d_cyc, sms);
    run<MIX_IMM>("test FFMA imm", CH / 2 * 8, d_out, d_cyc, sms);
    run<MIX_FADD>("test FADD.SAT", CH / 2 * 8, d_out, d_cyc, sms);
    run<VIADDMNMX_OP>("VIADDMNMX", CH, d_out, d_cyc, sms);
    run<IADD3_3>("IADD3 3-in", CH, d_out, d_cyc, sms);
    // the following report cycles per ROW PAIR (2 tests per lane)
    run<MIX_TEST>("pair pk FFMA.SAT", CH / 2, d_out, d_cyc, sms);
    run<PK_FADD>("pair pk FADD.SAT", CH / 2, d_out, d_cyc, sms);
    run<PK_ALU>("pair pk ALU acc", CH / 2, d_out, d_cyc, sms);
    run<SC_FFMA>("pair sc FFMA.SAT", CH / 2, d_out, d_cyc, sms);
    run<SC_FADD>("pair sc FADD.SAT", CH / 2, d_out, d_cyc, sms);
    run<SC_ALU>("pair sc ALU acc", CH / 2, d_out, d_cyc, sms);
    return 0;
}
