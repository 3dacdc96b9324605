// exp_scan.cu -- dev experiment: inner-loop shapes of the ADD-S scan, timed in isolation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o exp_scan exp_scan.cu
// Not part of the library; used to choose the loop shape of adds_cta_kernel.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
template <int MINMODE> __device__ __forceinline__ float min3v(float a, float b, float c) {
    float r;
    if (MINMODE == 0) asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    else if (MINMODE == 1) asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    else if (MINMODE == 2) { asm("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); asm("min.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(c)); }
    else r = a + b * 0.0f + c * 0.0f;  // MINMODE 3: no min (upper bound), keeps the dependency
    return r;
}

// MINMODE: 0 FMNMX3.NAN, 1 FMNMX3, 2 2xFMNMX, 3 none.  DEFER: 1 = mins of tile j issued after tile j+1's math.
template <int T, int K, int MINB, int U, int MINMODE, int DEFER, int SRC = 0, int UNR = 1>
__global__ void __launch_bounds__(T, MINB) scan_kernel(const float* __restrict__ g, int nquads, int reps, float* out) {
    extern __shared__ float4 sm[];
    float4* gx4 = sm;
    float4* gy4 = sm + nquads;
    float4* gz4 = sm + 2 * nquads;
    for (int i = threadIdx.x; i < 3 * nquads; i += T) sm[i] = reinterpret_cast<const float4*>(g)[i];
    __syncthreads();
    float px[K], py[K], pz[K], m[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        px[k] = g[(threadIdx.x * K + k) % (4 * nquads)];
        py[k] = g[(threadIdx.x * K + k + 7) % (4 * nquads)] * 0.5f;
        pz[k] = g[(threadIdx.x * K + k + 13) % (4 * nquads)] * 0.25f;
        m[k] = 3.0e38f;
    }
    for (int r = 0; r < reps; ++r) {
        if (!DEFER) {
            float4 RX = gx4[r & 7], RY = gy4[r & 7], RZ = gz4[r & 7];
#pragma unroll UNR
            for (int qd = 0; qd < nquads; qd += U) {
                float4 X[U], Y[U], Z[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (SRC == 0) { X[u] = gx4[qd + u]; Y[u] = gy4[qd + u]; Z[u] = gz4[qd + u]; }
                    else { RX.x += 1.0f; RY.w -= 1.0f; X[u] = RX; Y[u] = RY; Z[u] = RZ; }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        {
                            float2 dx = sub2(make_float2(px[k], px[k]), make_float2(X[u].x, X[u].y));
                            float2 dy = sub2(make_float2(py[k], py[k]), make_float2(Y[u].x, Y[u].y));
                            float2 dz = sub2(make_float2(pz[k], pz[k]), make_float2(Z[u].x, Z[u].y));
                            float2 s = mul2(dx, dx); s = fma2(dy, dy, s); s = fma2(dz, dz, s);
                            m[k] = min3v<MINMODE>(m[k], s.x, s.y);
                        }
                        {
                            float2 dx = sub2(make_float2(px[k], px[k]), make_float2(X[u].z, X[u].w));
                            float2 dy = sub2(make_float2(py[k], py[k]), make_float2(Y[u].z, Y[u].w));
                            float2 dz = sub2(make_float2(pz[k], pz[k]), make_float2(Z[u].z, Z[u].w));
                            float2 s = mul2(dx, dx); s = fma2(dy, dy, s); s = fma2(dz, dz, s);
                            m[k] = min3v<MINMODE>(m[k], s.x, s.y);
                        }
                    }
                }
            }
        } else {
            // software-pipelined: the minima of half-quad h are taken while half-quad h+1 computes
            float2 pend[K];
#pragma unroll
            for (int k = 0; k < K; ++k) pend[k] = make_float2(3.0e38f, 3.0e38f);
#pragma unroll 1
            for (int qd = 0; qd < nquads; ++qd) {
                const float4 X = gx4[qd], Y = gy4[qd], Z = gz4[qd];
                float2 s0[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float2 dx = sub2(make_float2(px[k], px[k]), make_float2(X.x, X.y));
                    float2 dy = sub2(make_float2(py[k], py[k]), make_float2(Y.x, Y.y));
                    float2 dz = sub2(make_float2(pz[k], pz[k]), make_float2(Z.x, Z.y));
                    float2 s = mul2(dx, dx); s = fma2(dy, dy, s); s0[k] = fma2(dz, dz, s);
                    m[k] = min3v<MINMODE>(m[k], pend[k].x, pend[k].y);
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float2 dx = sub2(make_float2(px[k], px[k]), make_float2(X.z, X.w));
                    float2 dy = sub2(make_float2(py[k], py[k]), make_float2(Y.z, Y.w));
                    float2 dz = sub2(make_float2(pz[k], pz[k]), make_float2(Z.z, Z.w));
                    float2 s = mul2(dx, dx); s = fma2(dy, dy, s); pend[k] = fma2(dz, dz, s);
                    m[k] = min3v<MINMODE>(m[k], s0[k].x, s0[k].y);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) m[k] = min3v<MINMODE>(m[k], pend[k].x, pend[k].y);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) s += m[k];
    if (s == 123.456f) out[0] = s;
}

template <int T, int K, int MINB, int U, int MINMODE, int DEFER, int SRC = 0, int UNR = 1>
void run(const char* name, const float* d_g, float* d_out, int nquads, int reps, int sms) {
    auto kern = scan_kernel<T, K, MINB, U, MINMODE, DEFER, SRC, UNR>;
    size_t smem = (size_t)3 * nquads * sizeof(float4);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem);
    int grid = sms * per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, T, smem>>>(d_g, nquads, reps / 4 + 1, d_out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        kern<<<grid, T, smem>>>(d_g, nquads, reps, d_out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    double pairs = (double)grid * T * K * (double)nquads * 4.0 * reps;
    double tf = pairs * 8.0 / (best * 1e-3) / 1e12;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("%-34s T=%4d K=%d ctas/sm=%d regs=%3d  %8.3f ms  %6.2f TFLOP/s  %5.1f%% of 74.45  %s\n", name, T, K, per_sm,
           fa.numRegs, best, tf, tf / 74.45 * 100.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int nquads = 512;  // 2048 gt points
    std::vector<float> h(12 * nquads);
    srand(1);
    for (auto& v : h) v = (rand() % 2000) * 1e-4f;
    float *d_g, *d_out;
    cudaMalloc(&d_g, h.size() * sizeof(float));
    cudaMalloc(&d_out, 64);
    cudaMemcpy(d_g, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
    const int reps = 64;
    run<256, 8, 2, 1, 0, 0>("base K8 T256 B2 U1 min3.NaN", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 2, 0, 0>("K8 T256 B2 U2", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1, 0, 0, 1, 1>("K8 T256 regs-src (no LDS)", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1, 0, 0, 1, 4>("K8 T256 regs-src unroll4", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1, 0, 0, 0, 2>("K8 T256 LDS pragma-unroll2", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1, 0, 0, 0, 4>("K8 T256 LDS pragma-unroll4", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1, 0, 0, 0, 8>("K8 T256 LDS pragma-unroll8", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 1, 0, 0, 0, 4>("K4 T512 LDS pragma-unroll4", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 1, 0, 0, 0, 8>("K4 T512 LDS pragma-unroll8", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 1, 0, 0>("K4 T512 B2 U1", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 2, 0, 0>("K4 T512 B2 U2", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 4, 0, 0>("K4 T512 B2 U4", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1, 1, 0>("K8 T256 min3 (no NaN)", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1, 2, 0>("K8 T256 2x FMNMX", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1, 3, 0>("K8 T256 no min (bound)", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1, 0, 1>("K8 T256 deferred mins", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 1, 0, 1>("K4 T512 deferred mins", d_g, d_out, nquads, reps, sms);
    run<128, 8, 4, 1, 0, 0>("K8 T128 B4", d_g, d_out, nquads, reps, sms);
    run<128, 16, 2, 1, 0, 0>("K16 T128 B2", d_g, d_out, nquads, reps, sms);
    run<256, 6, 2, 1, 0, 0>("K6 T256 B2", d_g, d_out, nquads, reps, sms);
    run<384, 6, 1, 1, 0, 0>("K6 T384 B1", d_g, d_out, nquads, reps, sms);
    run<1024, 2, 1, 2, 0, 0>("K2 T1024 B1 U2", d_g, d_out, nquads, reps, sms);
    run<768, 4, 1, 2, 0, 0>("K4 T768 B1 U2", d_g, d_out, nquads, reps, sms);
    run<1024, 4, 1, 2, 0, 0>("K4 T1024 B1 U2 (64 regs)", d_g, d_out, nquads, reps, sms);
    return 0;
}
