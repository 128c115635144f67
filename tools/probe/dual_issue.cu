// Probe: can a B200 SM sub-partition issue more than one instruction per clock when the stream mixes
// instruction classes?  Each kernel interleaves two instruction forms on independent chains.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
constexpr int CH = 8, ITER = 8192;
enum { A_FFMA, A_FADD, A_FMUL, A_FMNMX, A_IADD, A_LOP, A_FFMA2, A_FADD2, A_FSAT, A_NONE, A_LOP3R, A_UMIN, A_SHF, A_PRMT, A_POPC, A_FSETSEL, A_ISETSEL, A_FFMASAT, A_LOP2R, A_SEL, A_IADD3, A_IMAD, A_FMNMXI };
template <int OP> __device__ __forceinline__ void op(float &x, int &n, float2 &p, float a, float b, float2 a2, float2 b2) {
    if (OP == A_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a), "f"(b));
    if (OP == A_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(b));
    if (OP == A_FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(a));
    if (OP == A_FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(x) : "f"(b));
    if (OP == A_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(n) : "r"(n));
    if (OP == A_LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n) : "r"(12345));
    if (OP == A_FFMA2) p = __ffma2_rn(p, a2, b2);
    if (OP == A_FADD2) p = __fadd2_rn(p, b2);
    if (OP == A_LOP3R) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n) : "r"(__float_as_int(a) + n), "r"(__float_as_int(b)));
    if (OP == A_LOP2R) asm volatile("lop3.b32 %0, %0, %1, 0x00800000, 0xf8;" : "+r"(n) : "r"(__float_as_int(b)));
    if (OP == A_UMIN) asm volatile("{ .reg .u32 t; xor.b32 t, %0, 0x3f800000; min.u32 %0, %0, t; }" : "+r"(n));
    if (OP == A_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(n) : "r"(__float_as_int(b)));
    if (OP == A_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(n) : "r"(__float_as_int(b)));
    if (OP == A_POPC) asm volatile("popc.b32 %0, %0;" : "+r"(n));
    if (OP == A_FSETSEL) asm volatile("{ .reg .pred p; .reg .f32 f; mov.b32 f, %0; setp.lt.f32 p, f, %1; selp.b32 %0, %0, 0x3f800001, p; }" : "+r"(n) : "f"(b));
    if (OP == A_ISETSEL) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.b32 %0, %0, 7, p; }" : "+r"(n) : "r"(__float_as_int(b)));
    if (OP == A_FFMASAT) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a), "f"(b));
    if (OP == A_SEL) asm volatile("slct.u32.s32 %0, %0, %1, %0;" : "+r"(n) : "r"(__float_as_int(b)));
    if (OP == A_IADD3) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(n) : "r"(__float_as_int(a)), "r"(__float_as_int(b)));
    if (OP == A_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(n) : "r"(__float_as_int(a)), "r"(__float_as_int(b)));
    if (OP == A_FMNMXI) asm volatile("min.f32 %0, %0, 0f3F000000;" : "+f"(x));
    if (OP == A_FSAT) asm volatile("add.sat.f32 %0, %0, %1;" : "+f"(x) : "f"(b));
}
template <int OP1, int OP2>
__global__ void __launch_bounds__(128) k(float *out, float a, float b) {
    float x1[CH], x2[CH]; int n1[CH], n2[CH]; float2 p1[CH], p2[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x1[i] = a * (threadIdx.x + i); x2[i] = b * (threadIdx.x + i); n1[i] = threadIdx.x + i; n2[i] = threadIdx.x * i;
        p1[i] = make_float2(x1[i], x2[i]); p2[i] = make_float2(x2[i], x1[i]); }
    const float2 a2 = make_float2(a, a * 1.5f), b2 = make_float2(b, b * 0.5f);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            op<OP1>(x1[i], n1[i], p1[i], a, b, a2, b2);
            op<OP2>(x2[i], n2[i], p2[i], a, b, a2, b2);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x1[i] + x2[i] + (float)n1[i] + (float)n2[i] + p1[i].x + p1[i].y + p2[i].x + p2[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP1, int OP2> void run(const char *name, float *d_out, int sms) {
    CK(cudaFuncSetAttribute(k<OP1, OP2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int ctas = 6;
    const size_t smem = ((size_t)(227 * 1024 / ctas) - 1024) & ~(size_t)127;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<OP1, OP2><<<sms * ctas, 128, smem>>>(d_out, 1.0001f, 0.5f);
    CK(cudaEventRecord(e0));
    k<OP1, OP2><<<sms * ctas, 128, smem>>>(d_out, 1.0001f, 0.5f);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const int n_instr = (OP2 == A_NONE ? 1 : 2) * CH;
    printf("%-16s %.3f cycles per instruction per SMSP (%.2f IPC)\n", name, ms * 1e-3 * 1.965e9 / ((double)ctas * ITER * n_instr),
           ((double)ctas * ITER * n_instr) / (ms * 1e-3 * 1.965e9));
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); const int sms = prop.multiProcessorCount;
    float *d_out; CK(cudaMalloc(&d_out, sms * 6 * 128 * 4));
    run<A_FFMA, A_NONE>("FFMA", d_out, sms);
    run<A_FFMA, A_FFMA>("FFMA+FFMA", d_out, sms);
    run<A_FFMA, A_FADD>("FFMA+FADD", d_out, sms);
    run<A_FFMA, A_FMUL>("FFMA+FMUL", d_out, sms);
    run<A_FADD, A_FADD>("FADD+FADD", d_out, sms);
    run<A_FFMA, A_FSAT>("FFMA+FADD.SAT", d_out, sms);
    run<A_FFMA, A_FMNMX>("FFMA+FMNMX", d_out, sms);
    run<A_FFMA, A_IADD>("FFMA+IADD", d_out, sms);
    run<A_FFMA, A_LOP>("FFMA+LOP", d_out, sms);
    run<A_IADD, A_LOP>("IADD+LOP", d_out, sms);
    run<A_FADD, A_IADD>("FADD+IADD", d_out, sms);
    run<A_LOP3R, A_NONE>("LOP3rrr", d_out, sms);
    run<A_UMIN, A_NONE>("XOR+UMIN (2 instr)", d_out, sms);
    run<A_SHF, A_NONE>("SHF", d_out, sms);
    run<A_PRMT, A_NONE>("PRMT", d_out, sms);
    run<A_POPC, A_NONE>("POPC", d_out, sms);
    run<A_FSETSEL, A_NONE>("FSETP+SEL (2 instr)", d_out, sms);
    run<A_ISETSEL, A_NONE>("ISETP+SEL (2 instr)", d_out, sms);
    run<A_SEL, A_NONE>("SLCT", d_out, sms);
    run<A_IMAD, A_NONE>("IMAD", d_out, sms);
    run<A_FFMA, A_LOP3R>("FFMA+LOP3rrr", d_out, sms);
    run<A_FFMA, A_LOP2R>("FFMA+LOP3rri", d_out, sms);
    run<A_FFMA, A_UMIN>("FFMA+XOR+UMIN (3 as 2)", d_out, sms);
    run<A_FFMA, A_SHF>("FFMA+SHF", d_out, sms);
    run<A_FFMA, A_PRMT>("FFMA+PRMT", d_out, sms);
    run<A_FFMA, A_POPC>("FFMA+POPC", d_out, sms);
    run<A_FFMA, A_FSETSEL>("FFMA+FSETP+SEL (3 instr as 2)", d_out, sms);
    run<A_FFMA, A_ISETSEL>("FFMA+ISETP+SEL (3 instr as 2)", d_out, sms);
    run<A_FFMA, A_SEL>("FFMA+SLCT", d_out, sms);
    run<A_FFMA, A_IADD3>("FFMA+IADD3", d_out, sms);
    run<A_FFMA, A_IMAD>("FFMA+IMAD", d_out, sms);
    run<A_FFMA, A_FMNMXI>("FFMA+FMNMXimm", d_out, sms);
    run<A_FFMASAT, A_LOP>("FFMA.SAT+LOP", d_out, sms);
    run<A_FFMASAT, A_UMIN>("FFMA.SAT+XOR+UMIN (3 as 2)", d_out, sms);
    run<A_FADD, A_LOP3R>("FADD+LOP3rrr", d_out, sms);
    run<A_FADD, A_UMIN>("FADD+XOR+UMIN (3 as 2)", d_out, sms);
    run<A_UMIN, A_LOP>("XOR+UMIN+LOP (3 as 2)", d_out, sms);
    run<A_UMIN, A_LOP3R>("XOR+UMIN+LOP3rrr (3 as 2)", d_out, sms);
    run<A_FFMA2, A_NONE>("FFMA2", d_out, sms);
    run<A_FFMA2, A_FFMA>("FFMA2+FFMA", d_out, sms);
    run<A_FFMA2, A_FADD>("FFMA2+FADD", d_out, sms);
    run<A_FFMA2, A_IADD>("FFMA2+IADD", d_out, sms);
    run<A_FFMA2, A_FADD2>("FFMA2+FADD2", d_out, sms);
    run<A_FADD2, A_NONE>("FADD2", d_out, sms);
    run<A_FADD2, A_FADD>("FADD2+FADD", d_out, sms);
    return 0;
}
