// p6d_peak.cu -- FP32 issue-rate microbenchmarks (measurement helper, sm_100a).
//
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the ADD-S kernel is bound by
// the FP32 FMA pipe, so its roofline needs a measured FP32 denominator as well:
//   kind 0  FFMA  : 16 independent scalar FMA chains per thread
//   kind 1  FFMA2 : 16 independent packed (f32x2) FMA chains per thread
//   kind 2  MIX   : the ADD-S inner tile (3 FADD2 + FMUL2 + 2 FFMA2 + FMNMX3 per two pairs)
//                   on register operands that change every tile, no shared-memory traffic:
//                   the issue-rate ceiling of the kernel's instruction mix
// FLOP accounting: FMA = 2 FLOP per lane-op; MIX = 8 FLOP per pair (3 sub, 3 mul, 2 add),
// the same convention as the kernel's algorithmic FLOPs.
#include "p6d_common.cuh"

namespace p6d {

constexpr int PK_T = 256;
constexpr int PK_CHAINS = 16;
constexpr int PK_INNER = 64;

__global__ void __launch_bounds__(PK_T) peak_ffma_kernel(float* out, int iters, float a, float b) {
    float v[PK_CHAINS];
#pragma unroll
    for (int i = 0; i < PK_CHAINS; ++i) v[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < PK_INNER; ++j) {
#pragma unroll
            for (int i = 0; i < PK_CHAINS; ++i) v[i] = __fmaf_rn(v[i], a, b);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < PK_CHAINS; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(PK_T) peak_ffma2_kernel(float* out, int iters, float a, float b) {
    float2 v[PK_CHAINS];
#pragma unroll
    for (int i = 0; i < PK_CHAINS; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, i - threadIdx.x * 1e-3f);
    const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < PK_INNER; ++j) {
#pragma unroll
            for (int i = 0; i < PK_CHAINS; ++i) v[i] = fma2(v[i], aa, bb);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < PK_CHAINS; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) out[0] = s;
}

constexpr int MIX_K = 8;
constexpr int MIX_INNER = 16;   // 16 tiles x 56 instructions = 14 KB of code: stays inside the 32 KB instruction cache

// volatile forms: keep the program order of the tile (the scheduler of the non-volatile forms
// interleaves differently and measures ~10 % lower)
__device__ __forceinline__ float2 vsub2(float a, float2 b) {
    float2 r;
    asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%2}; mov.b64 rb, {%3,%4}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
                 : "=f"(r.x), "=f"(r.y) : "f"(a), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 vmul2(float2 a) {
    float2 r;
    asm volatile("{.reg .b64 ra, rc; mov.b64 ra, {%2,%3}; mul.rn.f32x2 rc, ra, ra; mov.b64 {%0,%1}, rc;}"
                 : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ void vfma2(float2& s, float2 a) {
    asm volatile("{.reg .b64 ra, rc; mov.b64 ra, {%2,%3}; mov.b64 rc, {%0,%1}; fma.rn.f32x2 rc, ra, ra, rc; mov.b64 {%0,%1}, rc;}"
                 : "+f"(s.x), "+f"(s.y) : "f"(a.x), "f"(a.y));
}
__device__ __forceinline__ void vmin3(float& m, float2 s) {
    asm volatile("min.NaN.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(s.x), "f"(s.y));
}
// ADD-S tile on register operands that change every tile (nothing is loop-invariant, so
// the compiler cannot hoist any of the 6 packed ops): 8 "pred" points x 1 gt pair per
// tile = 48 packed FMA-pipe instructions + 8 FMNMX3 + 1 scalar add.
__global__ void __launch_bounds__(PK_T) peak_mix_kernel(float* out, int iters, float a, float b) {
    float2 v[16];
    float m[MIX_K], qx[MIX_K], qy[MIX_K], qz[MIX_K];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, i - threadIdx.x * 1e-3f);
#pragma unroll
    for (int k = 0; k < MIX_K; ++k) {
        m[k] = 3.0e38f;
        qx[k] = a + k; qy[k] = b - k; qz[k] = a * k;   // the "pred points", register-resident like in the kernel
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < MIX_INNER; ++j) {
#pragma unroll
            for (int k = 0; k < MIX_K; ++k) {
                const float px = qx[k], py = qy[k], pz = qz[k];
                const float2 dx = vsub2(px, v[(j + 0) & 15]);
                const float2 dy = vsub2(py, v[(j + 1) & 15]);
                const float2 dz = vsub2(pz, v[(j + 2) & 15]);
                float2 s = vmul2(dx);
                vfma2(s, dy);
                vfma2(s, dz);
                vmin3(m[k], s);
            }
            v[j & 15].x += 1.0f;  // keeps the "gt" operands changing
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < MIX_K; ++k) s += m[k];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i].x;
    if (s == 123.456f) out[0] = s;
}

}  // namespace p6d

using namespace p6d;

extern "C" int p6d_fp32_microbench(int kind, int device, int iters, double* tflops, double* ms) {
    if (kind < 0 || kind > 2 || iters < 1 || !tflops || !ms) { set_error("p6d_fp32_microbench: bad arguments"); return P6D_EINVAL; }
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    int sms = 0;
    P6D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float* d_out = nullptr;
    P6D_CUDA(cudaMalloc(&d_out, 64));
    cudaEvent_t e0, e1;
    P6D_CUDA(cudaEventCreate(&e0));
    P6D_CUDA(cudaEventCreate(&e1));
    const int per_sm = 4;
    const unsigned grid = (unsigned)(sms * per_sm);
    auto launch = [&](int n) {
        if (kind == 0) peak_ffma_kernel<<<grid, PK_T>>>(d_out, n, 1.0001f, 0.5f);
        else if (kind == 1) peak_ffma2_kernel<<<grid, PK_T>>>(d_out, n, 1.0001f, 0.5f);
        else peak_mix_kernel<<<grid, PK_T>>>(d_out, n, 1.0001f, 0.5f);
    };
    launch(iters / 4 + 1);  // warm-up
    P6D_CUDA(cudaDeviceSynchronize());
    P6D_CUDA(cudaEventRecord(e0));
    launch(iters);
    P6D_CUDA(cudaEventRecord(e1));
    P6D_CUDA(cudaEventSynchronize(e1));
    float t = 0.0f;
    P6D_CUDA(cudaEventElapsedTime(&t, e0, e1));
    P6D_CUDA(cudaGetLastError());
    const double threads = (double)grid * PK_T;
    double flop;
    if (kind == 0) flop = threads * iters * (double)PK_INNER * PK_CHAINS * 2.0;
    else if (kind == 1) flop = threads * iters * (double)PK_INNER * PK_CHAINS * 4.0;
    else flop = threads * iters * (double)MIX_INNER * MIX_K * 2.0 /*pairs per tile*/ * 8.0 /*FLOP per pair*/;
    *ms = t;
    *tflops = flop / (t * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return P6D_OK;
}
