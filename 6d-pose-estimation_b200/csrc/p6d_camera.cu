// p6d_camera.cu -- geometric-translation kernels (sm_100a).
//
//   pinhole_fwd / pinhole_bwd : PoseNetRGBGeometric._compute_pinhole_translation
//                               (reference models/pose_net_rgb_geometric.py:93-109)
//   depth_backproject         : PoseNetRGBDGeometric._compute_pinhole_translation
//                               (reference models/pose_net_rgbd_geometric.py:56-85)
//
// Elementwise, HBM/latency-bound: one thread per row, K read as scalars (4 of the 9
// entries), everything else coalesced.  Operation order is the reference's:
// ((u - cx) * z) / fx with an IEEE division.
#include "p6d_common.cuh"

namespace p6d {

constexpr int CAM_T = 256;

__device__ __forceinline__ void load_k(const float* __restrict__ K, int k_batched, int64_t b, float& fx,
                                       float& fy, float& cx, float& cy) {
    const float* k = K + (k_batched ? 9 * b : 0);
    fx = __ldg(k + 0);
    cx = __ldg(k + 2);
    fy = __ldg(k + 4);
    cy = __ldg(k + 5);
}

__global__ void __launch_bounds__(CAM_T) pinhole_fwd_kernel(const float* __restrict__ z, const float* __restrict__ uv,
                                                            const float* __restrict__ K, int k_batched, int64_t B,
                                                            float* __restrict__ out) {
    for (int64_t b = (int64_t)blockIdx.x * CAM_T + threadIdx.x; b < B; b += (int64_t)gridDim.x * CAM_T) {
        float fx, fy, cx, cy;
        load_k(K, k_batched, b, fx, fy, cx, cy);
        const float2 c = *reinterpret_cast<const float2*>(uv + 2 * b);
        const float zz = z[b];
        out[3 * b + 0] = __fdiv_rn(__fmul_rn(__fsub_rn(c.x, cx), zz), fx);
        out[3 * b + 1] = __fdiv_rn(__fmul_rn(__fsub_rn(c.y, cy), zz), fy);
        out[3 * b + 2] = zz;
    }
}

__global__ void __launch_bounds__(CAM_T) pinhole_bwd_kernel(const float* __restrict__ go, const float* __restrict__ uv,
                                                            const float* __restrict__ K, int k_batched, int64_t B,
                                                            float* __restrict__ gz) {
    for (int64_t b = (int64_t)blockIdx.x * CAM_T + threadIdx.x; b < B; b += (int64_t)gridDim.x * CAM_T) {
        float fx, fy, cx, cy;
        load_k(K, k_batched, b, fx, fy, cx, cy);
        const float2 c = *reinterpret_cast<const float2*>(uv + 2 * b);
        // autograd order: d/dz[((u-cx)*z)/fx] = (g/fx)*(u-cx)
        const float gx = __fmul_rn(__fdiv_rn(go[3 * b + 0], fx), __fsub_rn(c.x, cx));
        const float gy = __fmul_rn(__fdiv_rn(go[3 * b + 1], fy), __fsub_rn(c.y, cy));
        gz[b] = __fadd_rn(__fadd_rn(gx, gy), go[3 * b + 2]);
    }
}

__device__ __forceinline__ float clamp_keep_nan(float x, float lo, float hi) {
    // torch.clamp: NaN stays NaN
    if (x != x) return x;
    return x < lo ? lo : (x > hi ? hi : x);
}

__global__ void __launch_bounds__(CAM_T) depth_backproject_kernel(const float* __restrict__ depth, int H, int W,
                                                                  const float* __restrict__ uv,
                                                                  const float* __restrict__ K, int k_batched,
                                                                  int64_t B, float clamp_hi,
                                                                  float* __restrict__ out) {
    const long long hi = (long long)clamp_hi;
    for (int64_t b = (int64_t)blockIdx.x * CAM_T + threadIdx.x; b < B; b += (int64_t)gridDim.x * CAM_T) {
        float fx, fy, cx, cy;
        load_k(K, k_batched, b, fx, fy, cx, cy);
        const float2 c = *reinterpret_cast<const float2*>(uv + 2 * b);
        const float u = clamp_keep_nan(c.x, 0.0f, clamp_hi);
        const float v = clamp_keep_nan(c.y, 0.0f, clamp_hi);
        // .long(): truncation toward zero; NaN converts to INT64_MIN on the reference's x86 host
        long long ui = (u != u) ? LLONG_MIN : (long long)u;
        long long vi = (v != v) ? LLONG_MIN : (long long)v;
        ui = ui < 0 ? 0 : (ui > hi ? hi : ui);
        vi = vi < 0 ? 0 : (vi > hi ? hi : vi);
        float zz = __ldg(depth + ((int64_t)b * H + vi) * W + ui);
        zz = (zz > 0.01f) ? zz : 0.5f;  // false for NaN -> 0.5
        zz = clamp_keep_nan(zz, 0.1f, 2.0f);
        out[3 * b + 0] = __fdiv_rn(__fmul_rn(__fsub_rn(u, cx), zz), fx);
        out[3 * b + 1] = __fdiv_rn(__fmul_rn(__fsub_rn(v, cy), zz), fy);
        out[3 * b + 2] = zz;
    }
}

static int grid_for(int64_t B, int device, unsigned* grid) {
    int sms = 0;
    P6D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int64_t blocks = (B + CAM_T - 1) / CAM_T;
    const int64_t cap = (int64_t)sms * 8;
    *grid = (unsigned)(blocks > cap ? cap : blocks);
    return P6D_OK;
}

}  // namespace p6d

using namespace p6d;

extern "C" {

int p6d_pinhole_fwd(const float* z, const float* uv, const float* K, int k_batched, int64_t B, float* out,
                    int device, void* stream) {
    if (B < 0 || (B > 0 && (!z || !uv || !K || !out))) { set_error("p6d_pinhole_fwd: bad arguments"); return P6D_EINVAL; }
    if (B == 0) return P6D_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    unsigned grid;
    int rc = grid_for(B, device, &grid);
    if (rc) return rc;
    pinhole_fwd_kernel<<<grid, CAM_T, 0, static_cast<cudaStream_t>(stream)>>>(z, uv, K, k_batched, B, out);
    P6D_CUDA(cudaGetLastError());
    return P6D_OK;
}

int p6d_pinhole_bwd(const float* grad_out, const float* uv, const float* K, int k_batched, int64_t B,
                    float* grad_z, int device, void* stream) {
    if (B < 0 || (B > 0 && (!grad_out || !uv || !K || !grad_z))) { set_error("p6d_pinhole_bwd: bad arguments"); return P6D_EINVAL; }
    if (B == 0) return P6D_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    unsigned grid;
    int rc = grid_for(B, device, &grid);
    if (rc) return rc;
    pinhole_bwd_kernel<<<grid, CAM_T, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, uv, K, k_batched, B, grad_z);
    P6D_CUDA(cudaGetLastError());
    return P6D_OK;
}

int p6d_depth_backproject(const float* depth, int H, int W, const float* uv, const float* K, int k_batched,
                          int64_t B, float clamp_hi, float* out, int device, void* stream) {
    if (B < 0 || H < 1 || W < 1 || (B > 0 && (!depth || !uv || !K || !out))) {
        set_error("p6d_depth_backproject: bad arguments");
        return P6D_EINVAL;
    }
    if (!(clamp_hi >= 0.0f) || clamp_hi > (float)(H - 1) || clamp_hi > (float)(W - 1)) {
        set_error("p6d_depth_backproject: clamp_hi=%g outside the %dx%d depth map", (double)clamp_hi, H, W);
        return P6D_EINVAL;
    }
    if (B == 0) return P6D_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    unsigned grid;
    int rc = grid_for(B, device, &grid);
    if (rc) return rc;
    depth_backproject_kernel<<<grid, CAM_T, 0, static_cast<cudaStream_t>(stream)>>>(depth, H, W, uv, K, k_batched, B,
                                                                                    clamp_hi, out);
    P6D_CUDA(cudaGetLastError());
    return P6D_OK;
}

}  // extern "C"
