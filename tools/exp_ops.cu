// exp_ops.cu -- dev experiment: issue rates of the packed FP32 ops on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o exp_ops exp_ops.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { float2 r;
    asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)); return r; }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { float2 r;
    asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)); return r; }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { float2 r;
    asm volatile("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm volatile("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
constexpr int C = 16, INNER = 32, T = 256;
// MODE 0 fma2(v,a,b)  1 sub2(v, packed c)  2 sub2(bcast scalar, v)  3 mul2(v,a)  4 scalar FFMA  5 scalar FADD  6 scalar FMUL
// 10/11 scalar tile
// 7 tile w/o min: 3 sub2(bcast)+mul2+2 fma2 on changing operands   8 tile with min3   9 fma2(v,v,c) (a==b like the tile)
template <int MODE> __global__ void __launch_bounds__(T) k(float* out, int iters, float a, float b) {
    float2 v[C]; float m[C];
    for (int i = 0; i < C; ++i) { v[i] = make_float2(threadIdx.x * 1e-3f + i, i - threadIdx.x * 1e-3f); m[i] = 3e38f; }
    float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < INNER; ++j) {
#pragma unroll
            for (int i = 0; i < C; ++i) {
                if (MODE == 0) v[i] = fma2(v[i], aa, bb);
                else if (MODE == 1) v[i] = sub2(v[i], bb);
                else if (MODE == 2) v[i] = sub2(make_float2(a, a), v[i]);
                else if (MODE == 3) v[i] = mul2(v[i], aa);
                else if (MODE == 4) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i].x) : "f"(a), "f"(b)); }
                else if (MODE == 5) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i].x) : "f"(b)); }
                else if (MODE == 6) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[i].x) : "f"(a)); }
                else if (MODE == 9) v[i] = fma2(v[i], v[i], bb);
            }
            if (MODE >= 12 && MODE <= 20) {
                // issue-port probe: 16 independent packed FMAs interleaved with R independent non-FMA ops each
#pragma unroll
                for (int i = 0; i < C; ++i) {
                    v[i] = fma2(v[i], aa, bb);
                    if (MODE == 12) { asm volatile("min.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(b)); }
                    if (MODE == 13) { asm volatile("min.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(b)); asm volatile("max.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(a)); }
                    if (MODE == 14) { int t = __float_as_int(m[i]); asm volatile("add.s32 %0, %0, 3;" : "+r"(t)); m[i] = __int_as_float(t); }
                    if (MODE == 15 && (i & 1) == 0) { asm volatile("min.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(b)); }
                    if (MODE == 16 && (i & 3) == 0) { asm volatile("min.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(b)); }
                    if (MODE == 17) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(b), "f"(a)); }
                    if (MODE == 18) { int t = __float_as_int(m[i]); asm volatile("min.s32 %0, %0, %1;" : "+r"(t) : "r"(it + i)); m[i] = __int_as_float(t); }
                    if (MODE == 19) { m[i] = __int_as_float(__vimin3_s32(__float_as_int(m[i]), it + i, j - i)); }
                    if (MODE == 20) { m[i] = __uint_as_float(__vimin3_u32(__float_as_uint(m[i]), (unsigned)(it + i), (unsigned)(j + i))); }
                }
            }
            if (MODE == 7 || MODE == 8 || MODE == 21 || MODE == 22) {
                // 8 "pred points" (scalars m-independent) x 1 half-quad whose gt operand is v[j&15] (changes every j)
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const float px = a + kk, py = b - kk, pz = a * kk;
                    float2 dx = sub2(make_float2(px, px), v[(j + 0) & 15]);
                    float2 dy = sub2(make_float2(py, py), v[(j + 1) & 15]);
                    float2 dz = sub2(make_float2(pz, pz), v[(j + 2) & 15]);
                    float2 s = mul2(dx, dx); s = fma2(dy, dy, s); s = fma2(dz, dz, s);
                    if (MODE == 8) m[kk] = min3(m[kk], s.x, s.y);
                    else if (MODE == 21) m[kk] = __int_as_float(__vimin3_s32(__float_as_int(m[kk]), __float_as_int(s.x), __float_as_int(s.y)));
                    else if (MODE == 22) m[kk] = __uint_as_float(__vimin3_u32(__float_as_uint(m[kk]), __float_as_uint(s.x), __float_as_uint(s.y)));
                    else m[kk] = s.x;
                }
                v[j & 15].x += 1.0f;  // keep the operands changing
            }
            if (MODE == 10 || MODE == 11) {
                // scalar version of the same tile: 8 preds x 2 gt points (v[.].x and v[.].y)
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const float px = a + kk, py = b - kk, pz = a * kk;
                    float s2[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float gx = h ? v[(j + 0) & 15].y : v[(j + 0) & 15].x;
                        const float gy = h ? v[(j + 1) & 15].y : v[(j + 1) & 15].x;
                        const float gz = h ? v[(j + 2) & 15].y : v[(j + 2) & 15].x;
                        float dx, dy, dz, ss;
                        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dx) : "f"(px), "f"(gx));
                        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dy) : "f"(py), "f"(gy));
                        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(dz) : "f"(pz), "f"(gz));
                        asm volatile("mul.rn.f32 %0, %1, %1;" : "=f"(ss) : "f"(dx));
                        asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(ss) : "f"(dy));
                        asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(ss) : "f"(dz));
                        s2[h] = ss;
                    }
                    if (MODE == 10) m[kk] = min3(m[kk], s2[0], s2[1]);
                    else { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(m[kk]), "f"(s2[0])); asm volatile("min.f32 %0, %1, %2;" : "=f"(m[kk]) : "f"(r), "f"(s2[1])); }
                }
                v[j & 15].x += 1.0f;
            }
        }
    }
    float s = 0.f;
    for (int i = 0; i < C; ++i) s += v[i].x + v[i].y + m[i];
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char* name, double lane_ops_per_inner, float* d_out, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000; const int grid = sms * 4;
    k<MODE><<<grid, T>>>(d_out, 100, 1.0001f, 0.5f); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<MODE><<<grid, T>>>(d_out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    // warp-instructions per SMSP per cycle
    double winst = (double)grid * (T / 32) * iters * INNER * lane_ops_per_inner;   // warp-level instrs (FMA pipe)
    double cycles = best * 1e-3 * 1.965e9;
    printf("%-44s %8.3f ms   FMA-pipe warp-instr/cycle/SMSP = %.3f  (cycles per instr %.2f)\n", name, best, winst / (sms * 4) / cycles, (sms * 4) * cycles / winst);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d; cudaMalloc(&d, 64);
    run<4>("scalar FFMA", C, d, sms); run<5>("scalar FADD", C, d, sms); run<6>("scalar FMUL", C, d, sms);
    run<0>("FFMA2 v=fma2(v,a,b)", C, d, sms); run<9>("FFMA2 v=fma2(v,v,b)", C, d, sms);
    run<1>("FADD2 packed-packed", C, d, sms); run<2>("FADD2 scalar-bcast - packed", C, d, sms); run<3>("FMUL2", C, d, sms);
    run<7>("tile 3xFADD2+FMUL2+2xFFMA2 (48 per inner)", 48, d, sms);
    run<8>("tile + FMNMX3 (48 FMA-pipe per inner)", 48, d, sms);
    run<12>("FFMA2 + 1 indep FMNMX per FFMA2", C, d, sms);
    run<13>("FFMA2 + 2 indep FMNMX per FFMA2", C, d, sms);
    run<14>("FFMA2 + 1 indep IADD per FFMA2", C, d, sms);
    run<15>("FFMA2 + 1 FMNMX per 2 FFMA2", C, d, sms);
    run<16>("FFMA2 + 1 FMNMX per 4 FFMA2", C, d, sms);
    run<17>("FFMA2 + 1 indep FMNMX3 per FFMA2", C, d, sms);
    run<18>("FFMA2 + 1 indep IMNMX (s32) per FFMA2", C, d, sms);
    run<19>("FFMA2 + 1 indep VIMNMX3 (s32) per FFMA2", C, d, sms);
    run<20>("FFMA2 + 1 indep VIMNMX3.U32 per FFMA2", C, d, sms);
    run<21>("tile + VIMNMX3 s32 (48 FMA-pipe per inner)", 48, d, sms);
    run<22>("tile + VIMNMX3 u32 (48 FMA-pipe per inner)", 48, d, sms);
    run<10>("SCALAR tile + FMNMX3 (96 FMA-pipe per inner)", 96, d, sms);
    run<11>("SCALAR tile + 2xFMNMX (96 per inner)", 96, d, sms);
    return 0;
}
