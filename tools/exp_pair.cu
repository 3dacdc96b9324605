// exp_pair.cu -- dev experiment: cost of one non-FMA instruction issued next to one packed FP32 op.
// For each (packed op P, other op O): 16 independent chains of P, each followed by one independent O.
// Reports cycles per (P + O) pair per SMSP; 2.0 = O hides completely in P's second issue cycle.
#include <cstdio>
#include <cuda_runtime.h>
#include <string>
constexpr int C = 16, INNER = 16, T = 256;
template <int P, int O> __global__ void __launch_bounds__(T) k(float* out, int iters, float a, float b, const float* sm_src) {
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = reinterpret_cast<const float4*>(sm_src)[threadIdx.x];
    __syncthreads();
    float2 v[C]; float m[C]; float4 ld = make_float4(0, 0, 0, 0); int w[C]; int wc = (int)(a * 1000.f), wd = (int)(b * 977.f);
    for (int i = 0; i < C; ++i) w[i] = threadIdx.x * 7 + i * 1000003;
    for (int i = 0; i < C; ++i) { v[i] = make_float2(threadIdx.x * 1e-3f + i, i - threadIdx.x * 1e-3f); m[i] = 3e38f - i; }
    const float c1 = a * 0.5f, c2 = b * 3.0f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < INNER; ++j) {
#pragma unroll
            for (int i = 0; i < C; ++i) {
                // packed op
                if (P == 0) asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%2}; mov.b64 rb, {%0,%1}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "+f"(v[i].x), "+f"(v[i].y) : "f"(a));
                if (P == 1) asm volatile("{.reg .b64 ra, rc; mov.b64 ra, {%0,%1}; mul.rn.f32x2 rc, ra, ra; mov.b64 {%0,%1}, rc;}" : "+f"(v[i].x), "+f"(v[i].y));
                if (P == 2) asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; fma.rn.f32x2 rc, ra, ra, rb; mov.b64 {%0,%1}, rc;}" : "+f"(v[i].x), "+f"(v[i].y) : "f"(v[(i + 1) & 15].x), "f"(v[(i + 1) & 15].y));
                if (P == 3) { asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+f"(v[i].x) : "f"(v[(i + 1) & 15].y)); asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+f"(v[i].y) : "f"(v[(i + 1) & 15].x)); }
                if (P == 4) { asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+f"(v[i].x) : "f"(v[(i + 1) & 15].y)); }
                // other op (independent of the packed chain)
                if (O == 1) asm volatile("min.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(b));
                if (O == 2) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(c1), "f"(c2));
                if (O == 3) asm volatile("min.f32 %0, %0, 0f42C80000;" : "+f"(m[i]));
                if (O == 4) { int t = __float_as_int(m[i]); asm volatile("add.s32 %0, %0, 3;" : "+r"(t)); m[i] = __int_as_float(t); }
                if (O == 5) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(m[(i + 5) & 15]), "f"(m[(i + 9) & 15]));
                if (O == 6 && (i & 3) == 0) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(ld.x), "=f"(ld.y), "=f"(ld.z), "=f"(ld.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(j + i) & 63])));
                if (O == 7) asm volatile("mov.b32 %0, %1;" : "=f"(m[i]) : "f"(m[(i + 3) & 15]));
                if (O == 8) { int t = __float_as_int(m[i]); asm volatile("min.s32 %0, %0, %1;" : "+r"(t) : "r"(__float_as_int(b))); m[i] = __int_as_float(t); }
                if (O == 9) { unsigned t = __float_as_uint(m[i]); asm volatile("min.u32 %0, %0, %1;" : "+r"(t) : "r"(__float_as_uint(b))); m[i] = __uint_as_float(t); }
                if (O == 10) { m[i] = __int_as_float(__vimin3_s32(__float_as_int(m[i]), __float_as_int(c1), __float_as_int(c2))); }
                if (O == 11) { unsigned t = __float_as_uint(m[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(t) : "r"(__float_as_uint(b)), "r"(__float_as_uint(c1))); m[i] = __uint_as_float(t); }
                if (O == 12) { int t = __float_as_int(m[i]); asm volatile("{.reg .pred p; setp.lt.s32 p, %1, %0; selp.b32 %0, %1, %0, p;}" : "+r"(t) : "r"(__float_as_int(b))); m[i] = __int_as_float(t); }
                if (O == 13) { asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %0; selp.f32 %0, %1, %0, p;}" : "+f"(m[i]) : "f"(b)); }
                if (O == 16) asm volatile("min.s32 %0, %0, %1;" : "+r"(w[i]) : "r"(wc));
                if (O == 17) w[i] = __vimin3_s32(w[i], wc, wd);
                if (O == 18) asm volatile("min.s32 %0, %0, %1;" : "+r"(w[i]) : "r"(w[(i + 5) & 15]));
                if (O == 19) asm volatile("min.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(wc));
                if (O == 20) asm volatile("add.s32 %0, %0, %1;" : "+r"(w[i]) : "r"(wc));
                if (O == 21) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(wc), "r"(wd));
                if (O == 14) { int t = __float_as_int(m[i]); asm volatile("add.s32 %0, %0, %1;" : "+r"(t) : "r"(__float_as_int(b))); m[i] = __int_as_float(t); }
                if (O == 15) { int t = __float_as_int(m[i]); asm volatile("{.reg .s32 x; sub.s32 x, %1, %0; shr.s32 x, x, 31; }" : "+r"(t) : "r"(__float_as_int(b))); m[i] = __int_as_float(t); }
            }
        }
    }
    float s = ld.x + ld.y + ld.z + ld.w;
    for (int i = 0; i < C; ++i) s += v[i].x + v[i].y + m[i] + (float)w[i];
    if (s == 123.456f) out[0] = s;
}
template <int P, int O> void run(const char* name, float* d_out, const float* d_src, int sms, double o_per_p) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, grid = sms * 4;
    k<P, O><<<grid, T>>>(d_out, 100, 1.0001f, 0.5f, d_src); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<P, O><<<grid, T>>>(d_out, iters, 1.0001f, 0.5f, d_src); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    const double p = (double)(T / 32) * 4 /*ctas per sm*/ / 4 /*smsp*/ * iters * INNER * C;   // packed instrs per SMSP
    const double cyc = best * 1e-3 * 1.965e9;
    printf("%-52s cycles per packed op = %.2f  -> cost of the other op = %.2f cycles\n", name, cyc / p, (cyc / p - 2.0) / o_per_p);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d, *src; cudaMalloc(&d, 64); cudaMalloc(&src, 1024); cudaMemset(src, 0, 1024);
    const char* pn[5] = {"FADD2(bcast,pair)", "FMUL2(pair,pair)", "FFMA2(pair,pair,pair2)", "2x scalar FFMA(r,r,r2)", "1x scalar FFMA(r,r,r2) [1 cyc]"};
#define ROW(P) \
    run<P, 0>((std::string(pn[P]) + " alone").c_str(), d, src, sms, 1); \
    run<P, 1>((std::string(pn[P]) + " + FMNMX r,r").c_str(), d, src, sms, 1); \
    run<P, 2>((std::string(pn[P]) + " + FMNMX3 r,r,r (2 const)").c_str(), d, src, sms, 1); \
    run<P, 5>((std::string(pn[P]) + " + FMNMX3 r,r,r (3 distinct)").c_str(), d, src, sms, 1); \
    run<P, 3>((std::string(pn[P]) + " + FMNMX r,imm").c_str(), d, src, sms, 1); \
    run<P, 4>((std::string(pn[P]) + " + IADD r,imm").c_str(), d, src, sms, 1); \
    run<P, 7>((std::string(pn[P]) + " + MOV r").c_str(), d, src, sms, 1); \
    run<P, 6>((std::string(pn[P]) + " + LDS.128 per 4").c_str(), d, src, sms, 0.25);
    ROW(0) ROW(1) ROW(2)
#define ROW2(P) \
    run<P, 8>((std::string(pn[P]) + " + VIMNMX.S32 r,r").c_str(), d, src, sms, 1); \
    run<P, 9>((std::string(pn[P]) + " + VIMNMX.U32 r,r").c_str(), d, src, sms, 1); \
    run<P, 10>((std::string(pn[P]) + " + VIMNMX3 r,r,r").c_str(), d, src, sms, 1); \
    run<P, 11>((std::string(pn[P]) + " + LOP3 r,r,r").c_str(), d, src, sms, 1); \
    run<P, 12>((std::string(pn[P]) + " + ISETP+SEL (int min)").c_str(), d, src, sms, 1); \
    run<P, 13>((std::string(pn[P]) + " + FSETP+FSEL (float min)").c_str(), d, src, sms, 1); \
    run<P, 14>((std::string(pn[P]) + " + IADD r,r").c_str(), d, src, sms, 1);
    ROW2(0) ROW2(2)
#define ROW3(P) \
    run<P, 16>((std::string(pn[P]) + " + VIMNMX.S32 int-reg, const-reg").c_str(), d, src, sms, 1); \
    run<P, 18>((std::string(pn[P]) + " + VIMNMX.S32 int-reg, int-reg").c_str(), d, src, sms, 1); \
    run<P, 19>((std::string(pn[P]) + " + VIMNMX.U32 int-reg, const-reg").c_str(), d, src, sms, 1); \
    run<P, 17>((std::string(pn[P]) + " + VIMNMX3 int regs").c_str(), d, src, sms, 1); \
    run<P, 20>((std::string(pn[P]) + " + IADD int-reg, reg").c_str(), d, src, sms, 1); \
    run<P, 21>((std::string(pn[P]) + " + LOP3 int regs").c_str(), d, src, sms, 1);
    ROW3(0) ROW3(2)
    return 0;
}
