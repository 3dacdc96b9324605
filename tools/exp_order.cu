// exp_order.cu -- dev experiment: does the placement of the min instructions inside the ADD-S tile matter?
// All arithmetic is asm volatile so that program order survives into SASS as far as ptxas allows.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define SUB2(r, a, b) asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%2}; mov.b64 rb, {%3,%4}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a), "f"(b.x), "f"(b.y))
#define MUL2(r, a) asm volatile("{.reg .b64 ra, rc; mov.b64 ra, {%2,%3}; mul.rn.f32x2 rc, ra, ra; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y))
#define FMA2(r, a) asm volatile("{.reg .b64 ra, rc; mov.b64 ra, {%2,%3}; mov.b64 rc, {%0,%1}; fma.rn.f32x2 rc, ra, ra, rc; mov.b64 {%0,%1}, rc;}" : "+f"(r.x), "+f"(r.y) : "f"(a.x), "f"(a.y))
#define MIN3(m, s) asm volatile("min.NaN.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(s.x), "f"(s.y))
#define MIN2(m, v) asm volatile("min.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(v))

// ORDER 0: per pred point: sub,sub,sub,mul,fma,fma,min      (the library kernel's source order)
// ORDER 1: by plane: all subs, all muls, all fmas, all mins
// ORDER 2: deferred: subs(t), mins(t-1), mul/fma(t)
// ORDER 3: mins(t-1) interleaved one-by-one with the subs of tile t
template <int T, int K, int MINB, int ORDER>
__global__ void __launch_bounds__(T, MINB) scan_kernel(const float* __restrict__ g, int nquads, int reps, float* out) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 3 * nquads; i += T) sm[i] = reinterpret_cast<const float4*>(g)[i];
    __syncthreads();
    float px[K], py[K], pz[K], m[K], m2[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        m2[k] = 3.0e38f;
        px[k] = g[(threadIdx.x * K + k) % (4 * nquads)];
        py[k] = g[(threadIdx.x * K + k + 7) % (4 * nquads)] * 0.5f;
        pz[k] = g[(threadIdx.x * K + k + 13) % (4 * nquads)] * 0.25f;
        m[k] = 3.0e38f;
    }
    for (int r = 0; r < reps; ++r) {
        float2 pend[K];
#pragma unroll
        for (int k = 0; k < K; ++k) pend[k] = make_float2(3.0e38f, 3.0e38f);
        const float4* p = sm;
        const float4* const pe = sm + 3 * nquads;
#pragma unroll 1
        for (; p < pe; p += 3) {
            const float4 X = p[0], Y = p[1], Z = p[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 gx = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                const float2 gy = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                const float2 gz = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                float2 dx[K], dy[K], dz[K], s[K];
                if (ORDER == 0) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        SUB2(dx[k], px[k], gx); SUB2(dy[k], py[k], gy); SUB2(dz[k], pz[k], gz);
                        MUL2(s[k], dx[k]); FMA2(s[k], dy[k]); FMA2(s[k], dz[k]); MIN3(m[k], s[k]);
                    }
                } else if (ORDER == 1) {
#pragma unroll
                    for (int k = 0; k < K; ++k) { SUB2(dx[k], px[k], gx); }
#pragma unroll
                    for (int k = 0; k < K; ++k) { SUB2(dy[k], py[k], gy); }
#pragma unroll
                    for (int k = 0; k < K; ++k) { SUB2(dz[k], pz[k], gz); }
#pragma unroll
                    for (int k = 0; k < K; ++k) { MUL2(s[k], dx[k]); }
#pragma unroll
                    for (int k = 0; k < K; ++k) { FMA2(s[k], dy[k]); }
#pragma unroll
                    for (int k = 0; k < K; ++k) { FMA2(s[k], dz[k]); }
#pragma unroll
                    for (int k = 0; k < K; ++k) { MIN3(m[k], s[k]); }
                } else if (ORDER == 2) {
#pragma unroll
                    for (int k = 0; k < K; ++k) { SUB2(dx[k], px[k], gx); SUB2(dy[k], py[k], gy); SUB2(dz[k], pz[k], gz); }
#pragma unroll
                    for (int k = 0; k < K; ++k) { MIN3(m[k], pend[k]); }
#pragma unroll
                    for (int k = 0; k < K; ++k) { MUL2(s[k], dx[k]); FMA2(s[k], dy[k]); FMA2(s[k], dz[k]); pend[k] = s[k]; }
                } else if (ORDER == 4) {
                    // two 2-input minima on separate accumulators (cannot fuse into FMNMX3), each placed
                    // right behind a FADD2 of the same warp (FMNMX is free next to FADD2/FMUL2, costs a
                    // full cycle next to FFMA2: tools/exp_pair.cu)
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        SUB2(dx[k], px[k], gx); MIN2(m[k], pend[k].x); SUB2(dy[k], py[k], gy); MIN2(m2[k], pend[k].y);
                        SUB2(dz[k], pz[k], gz);
                    }
#pragma unroll
                    for (int k = 0; k < K; ++k) { MUL2(s[k], dx[k]); FMA2(s[k], dy[k]); FMA2(s[k], dz[k]); pend[k] = s[k]; }
                } else if (ORDER == 5) {
                    // same two accumulators, but in the library kernel's per-point order (no deferral)
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        SUB2(dx[k], px[k], gx); SUB2(dy[k], py[k], gy); SUB2(dz[k], pz[k], gz);
                        MUL2(s[k], dx[k]); FMA2(s[k], dy[k]); FMA2(s[k], dz[k]); MIN2(m[k], s[k].x); MIN2(m2[k], s[k].y);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) { SUB2(dx[k], px[k], gx); MIN3(m[k], pend[k]); SUB2(dy[k], py[k], gy); SUB2(dz[k], pz[k], gz); }
#pragma unroll
                    for (int k = 0; k < K; ++k) { MUL2(s[k], dx[k]); FMA2(s[k], dy[k]); FMA2(s[k], dz[k]); pend[k] = s[k]; }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) MIN3(m[k], pend[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) MIN2(m[k], m2[k]);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) s += m[k];
    if (s == 123.456f) out[0] = s;
}

template <int T, int K, int MINB, int ORDER>
void run(const char* name, const float* d_g, float* d_out, int nquads, int reps, int sms) {
    auto kern = scan_kernel<T, K, MINB, ORDER>;
    size_t smem = (size_t)3 * nquads * sizeof(float4);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem);
    int grid = sms * per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, T, smem>>>(d_g, nquads, reps / 4 + 1, d_out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0); kern<<<grid, T, smem>>>(d_g, nquads, reps, d_out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    double pairs = (double)grid * T * K * (double)nquads * 4.0 * reps;
    double tf = pairs * 8.0 / (best * 1e-3) / 1e12;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("%-28s T=%4d K=%d ctas/sm=%d regs=%3d %8.3f ms %6.2f TFLOP/s %5.1f%% of 74.45 %s\n", name, T, K, per_sm, fa.numRegs,
           best, tf, tf / 74.45 * 100.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int nquads = 512;
    std::vector<float> h(12 * nquads); srand(1);
    for (auto& v : h) v = (rand() % 2000) * 1e-4f;
    float *d_g, *d_out; cudaMalloc(&d_g, h.size() * sizeof(float)); cudaMalloc(&d_out, 64);
    cudaMemcpy(d_g, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
    const int reps = 64;
    run<512, 4, 2, 0>("K4 per-point order", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 1>("K4 plane order", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 2>("K4 deferred mins", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 3>("K4 mins among subs", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 4>("K4 2xFMNMX behind FADD2", d_g, d_out, nquads, reps, sms);
    run<512, 4, 2, 5>("K4 2xFMNMX per-point", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 4>("K8 2xFMNMX behind FADD2", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 5>("K8 2xFMNMX per-point", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 0>("K8 per-point order", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 1>("K8 plane order", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 2>("K8 deferred mins", d_g, d_out, nquads, reps, sms);
    run<256, 8, 2, 3>("K8 mins among subs", d_g, d_out, nquads, reps, sms);
    return 0;
}
