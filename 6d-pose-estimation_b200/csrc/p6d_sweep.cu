// p6d_sweep.cu -- the compare_all_models sweep (BASELINE config 5) as ONE native call per rank.
//
// The reference evaluates a validation set per model variant in a Python loop
// (scripts/visualization/compare_all_models.py:65-104: batches of 16 through the network, then
// ADDLoss.eval_metrics, four host syncs per pose).  Config 5 is that sweep at scale: 13 objects x
// 4 model variants x 1 M pose hypotheses.  Here the whole sweep of one rank -- hypothesis
// generation, the variant's translation kernel (d1 / d2), evaluation (a)+(b), per-variant and
// per-object accumulation -- is driven from C++ on two streams (generation of chunk k+1 overlaps
// the evaluation of chunk k, double-buffered), with no Python and no host synchronisation inside.
//
// Hypotheses are a pure function of (seed, object index, variant index, hypothesis index)
// (counter-based Philox4x32-10), so any sharding of the hypothesis axis over ranks evaluates the
// same 52 M poses and the integer hit totals are identical for every number of GPUs.
#include <vector>

#include "p6d_common.cuh"

extern "C" int p6d_pinhole_fwd(const float* z, const float* uv, const float* K, int k_batched, int64_t B, float* out,
                               int device, void* stream);
extern "C" int p6d_depth_backproject(const float* depth, int H, int W, const float* uv, const float* K, int k_batched,
                                     int64_t B, float clamp_hi, float* out, int device, void* stream);

namespace p6d {

constexpr int SYNTH_T = 256;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float u01(uint32_t x) { return static_cast<float>(x >> 8) * (1.0f / 16777216.0f); }        // [0,1)
__device__ __forceinline__ float u01_open(uint32_t x) { return (static_cast<float>(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }  // (0,1]

// two standard normals from two words (Box-Muller)
__device__ __forceinline__ float2 normal2(uint32_t a, uint32_t b) {
    const float r = sqrtf(-2.0f * logf(u01_open(a)));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    return make_float2(r * c, r * s);
}

struct SynthArgs {
    uint2 key;            // Philox key: seed mixed with (object index, variant index)
    int64_t first;        // global hypothesis index of element 0
    int64_t n;
    float rot_sigma, trans_sigma;
    int kind;             // 0 = direct xyz (rgb / rgbd), 1 = pinhole from (z, bbox centre) (rgb_geometric),
                          // 2 = depth crop under the crop-space centre (rgbd_geometric)
    float fx, fy, cx, cy;
    int64_t oid;
    float* pq; float* pt; float* gq; float* gt; int64_t* obj;
    float* z;             // kind 1: [n]
    float* uv;            // kind 1: [n,2] full-image bbox centre; kind 2: [n,2] crop-space centre
    float* kc;            // kind 2: [n,9] crop intrinsics
    float* depth;         // kind 2: [n,8,8] metres
};

// Same distributions as workloads.random_poses / sweep.variant_translation:
//   gt_q = normalize(randn4); pred_q = normalize(gt_q + rot_sigma * randn4);
//   gt_t = (U(-.2,.2), U(-.2,.2), U(.4,1.2)); pred_t = gt_t + trans_sigma * randn3
__global__ void __launch_bounds__(SYNTH_T) synth_poses_kernel(SynthArgs a) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * SYNTH_T + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * SYNTH_T) {
        const uint64_t g = static_cast<uint64_t>(a.first + i);
        const uint32_t glo = static_cast<uint32_t>(g), ghi = static_cast<uint32_t>(g >> 32);
        const uint4 r0 = philox4x32_10(make_uint4(glo, ghi, 0u, 0u), a.key);
        const uint4 r1 = philox4x32_10(make_uint4(glo, ghi, 1u, 0u), a.key);
        const uint4 r2 = philox4x32_10(make_uint4(glo, ghi, 2u, 0u), a.key);
        const uint4 r3 = philox4x32_10(make_uint4(glo, ghi, 3u, 0u), a.key);
        const float2 n0 = normal2(r0.x, r0.y), n1 = normal2(r0.z, r0.w);      // gt quaternion
        const float2 n2 = normal2(r1.x, r1.y), n3 = normal2(r1.z, r1.w);      // rotation noise
        const float2 n4 = normal2(r2.x, r2.y), n5 = normal2(r2.z, r2.w);      // translation noise (3 used)
        float gq[4] = {n0.x, n0.y, n1.x, n1.y};
        float inv = 1.0f / sqrtf(gq[0] * gq[0] + gq[1] * gq[1] + gq[2] * gq[2] + gq[3] * gq[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) gq[k] *= inv;
        float pq[4] = {gq[0] + a.rot_sigma * n2.x, gq[1] + a.rot_sigma * n2.y, gq[2] + a.rot_sigma * n3.x,
                       gq[3] + a.rot_sigma * n3.y};
        inv = 1.0f / sqrtf(pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2] + pq[3] * pq[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) pq[k] *= inv;
        const float gt[3] = {u01(r3.x) * 0.4f - 0.2f, u01(r3.y) * 0.4f - 0.2f, u01(r3.z) * 0.8f + 0.4f};
        const float pt[3] = {gt[0] + a.trans_sigma * n4.x, gt[1] + a.trans_sigma * n4.y, gt[2] + a.trans_sigma * n5.x};
        *reinterpret_cast<float4*>(a.gq + 4 * i) = make_float4(gq[0], gq[1], gq[2], gq[3]);
        *reinterpret_cast<float4*>(a.pq + 4 * i) = make_float4(pq[0], pq[1], pq[2], pq[3]);
#pragma unroll
        for (int k = 0; k < 3; ++k) a.gt[3 * i + k] = gt[k];
        a.obj[i] = a.oid;
        if (a.kind == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) a.pt[3 * i + k] = pt[k];
            continue;
        }
        // the detector's bbox centre = projection of the gt translation
        const float u = gt[0] / gt[2] * a.fx + a.cx, v = gt[1] / gt[2] * a.fy + a.cy;
        if (a.kind == 1) {
            a.z[i] = pt[2];
            a.uv[2 * i] = u;
            a.uv[2 * i + 1] = v;
        } else {
            // an 8x8 depth crop around the centre: sensor depth of the gt + 2 mm noise per pixel; the crop's
            // principal point is placed so that (u - cx) is preserved
            a.uv[2 * i] = 3.5f;
            a.uv[2 * i + 1] = 3.5f;
            float* k = a.kc + 9 * i;
            k[0] = a.fx; k[1] = 0.0f; k[2] = 3.5f - (u - a.cx);
            k[3] = 0.0f; k[4] = a.fy; k[5] = 3.5f - (v - a.cy);
            k[6] = 0.0f; k[7] = 0.0f; k[8] = 1.0f;
            float4* d = reinterpret_cast<float4*>(a.depth + 64 * i);
#pragma unroll 4
            for (int w = 0; w < 16; ++w) {
                const uint4 rr = philox4x32_10(make_uint4(glo, ghi, 4u + w, 0u), a.key);
                // zero-mean, unit-variance uniform noise: (u - 0.5) * sqrt(12)
                d[w] = make_float4(gt[2] + 0.002f * (u01(rr.x) - 0.5f) * 3.4641016f,
                                   gt[2] + 0.002f * (u01(rr.y) - 0.5f) * 3.4641016f,
                                   gt[2] + 0.002f * (u01(rr.z) - 0.5f) * 3.4641016f,
                                   gt[2] + 0.002f * (u01(rr.w) - 0.5f) * 3.4641016f);
            }
        }
    }
}

struct SweepBuf {
    float *pq, *pt, *gq, *gt, *z, *uv, *kc, *depth, *add, *adds;
    int64_t* obj;
    uint8_t *hit, *valid;
};

static size_t carve(char* base, size_t off, size_t chunk, bool need_depth, SweepBuf* b) {
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return base ? base + o : nullptr; };
    char* p;
    p = take(16 * chunk); if (b) b->pq = reinterpret_cast<float*>(p);
    p = take(16 * chunk); if (b) b->gq = reinterpret_cast<float*>(p);
    p = take(12 * chunk); if (b) b->pt = reinterpret_cast<float*>(p);
    p = take(12 * chunk); if (b) b->gt = reinterpret_cast<float*>(p);
    p = take(8 * chunk);  if (b) b->obj = reinterpret_cast<int64_t*>(p);
    p = take(4 * chunk);  if (b) b->z = reinterpret_cast<float*>(p);
    p = take(8 * chunk);  if (b) b->uv = reinterpret_cast<float*>(p);
    p = take(4 * chunk);  if (b) b->add = reinterpret_cast<float*>(p);
    p = take(4 * chunk);  if (b) b->adds = reinterpret_cast<float*>(p);
    p = take(chunk);      if (b) b->hit = reinterpret_cast<uint8_t*>(p);
    p = take(chunk);      if (b) b->valid = reinterpret_cast<uint8_t*>(p);
    if (need_depth) {
        p = take(36 * chunk);  if (b) b->kc = reinterpret_cast<float*>(p);
        p = take(256 * chunk); if (b) b->depth = reinterpret_cast<float*>(p);
    }
    return off;
}

}  // namespace p6d

using namespace p6d;

extern "C" int p6d_synth_poses(uint64_t seed, int obj_index, int variant_index, int64_t first, int64_t n,
                               float rot_sigma, float trans_sigma, int kind, const float* K_host, int64_t oid,
                               float* pq, float* pt, float* gq, float* gt, int64_t* obj, float* z, float* uv,
                               float* kc, float* depth, int device, void* stream) {
    if (n < 0 || kind < 0 || kind > 2 || (n > 0 && (!pq || !gq || !gt || !obj || !K_host)) ||
        (n > 0 && kind == 0 && !pt) || (n > 0 && kind == 1 && (!z || !uv)) || (n > 0 && kind == 2 && (!uv || !kc || !depth))) {
        set_error("p6d_synth_poses: bad arguments");
        return P6D_EINVAL;
    }
    if (n == 0) return P6D_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    SynthArgs a{};
    a.key = make_uint2(static_cast<uint32_t>(seed) + 1000003u * static_cast<uint32_t>(obj_index) +
                           10007u * static_cast<uint32_t>(variant_index),
                       static_cast<uint32_t>(seed >> 32) ^ 0x6d36u);
    a.first = first; a.n = n; a.rot_sigma = rot_sigma; a.trans_sigma = trans_sigma; a.kind = kind;
    a.fx = K_host[0]; a.cx = K_host[2]; a.fy = K_host[4]; a.cy = K_host[5];
    a.oid = oid;
    a.pq = pq; a.pt = pt; a.gq = gq; a.gt = gt; a.obj = obj; a.z = z; a.uv = uv; a.kc = kc; a.depth = depth;
    int sms = 0;
    P6D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int64_t blocks = (n + SYNTH_T - 1) / SYNTH_T;
    if (blocks > static_cast<int64_t>(sms) * 8) blocks = static_cast<int64_t>(sms) * 8;
    synth_poses_kernel<<<static_cast<unsigned>(blocks), SYNTH_T, 0, static_cast<cudaStream_t>(stream)>>>(a);
    P6D_CUDA(cudaGetLastError());
    return P6D_OK;
}

namespace p6d {

// Per-object totals of one evaluated chunk.  Inside a sweep every pose of a chunk belongs to the same
// object: letting the evaluation kernel bump the accumulators per pose means 4 atomics per pose onto
// the same 4 addresses (two of them float64), which serialise in L2 -- measured 10 % of the 500-point
// sweep.  This pass reads the chunk's 10 B per pose once and issues 4 atomics per CTA.
constexpr int RED_T = 256;

__global__ void __launch_bounds__(RED_T) chunk_totals_kernel(const float* __restrict__ add, const float* __restrict__ adds,
                                                             const uint8_t* __restrict__ hit,
                                                             const uint8_t* __restrict__ valid, int64_t n,
                                                             unsigned long long* acc_hits, unsigned long long* acc_valid,
                                                             double* acc_add, double* acc_adds) {
    unsigned h = 0, v = 0;
    double sa = 0.0, ss = 0.0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * RED_T + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * RED_T) {
        if (valid[i]) {
            ++v;
            h += hit[i];
            sa += static_cast<double>(add[i]);
            ss += static_cast<double>(adds[i]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        h += __shfl_xor_sync(0xffffffffu, h, o);
        v += __shfl_xor_sync(0xffffffffu, v, o);
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    __shared__ unsigned s_h[RED_T / 32], s_v[RED_T / 32];
    __shared__ double s_a[RED_T / 32], s_s[RED_T / 32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_h[w] = h; s_v[w] = v; s_a[w] = sa; s_s[w] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long th = 0, tv = 0;
        double ta = 0.0, ts = 0.0;
        for (int k = 0; k < RED_T / 32; ++k) { th += s_h[k]; tv += s_v[k]; ta += s_a[k]; ts += s_s[k]; }
        if (tv) {
            atomicAdd(acc_valid, tv);
            if (th) atomicAdd(acc_hits, th);
            if (acc_add) atomicAdd(acc_add, ta);
            if (acc_adds) atomicAdd(acc_adds, ts);
        }
    }
}

}  // namespace p6d

extern "C" int p6d_sweep_run(p6d_mesh_table* t, const int32_t* obj_ids, int n_obj, const int32_t* variant_kinds,
                             int n_variants, int64_t n_per_block, int64_t lo, int64_t hi, int64_t chunk,
                             uint64_t seed, const float* K_host, float rot_sigma, float trans_sigma,
                             int64_t* acc_hits, int64_t* acc_valid, double* acc_add_sum, double* acc_adds_sum,
                             int64_t check_n, float* check_pq, float* check_pt, float* check_gq, float* check_gt,
                             float* check_add, float* check_adds, uint8_t* check_hit, int* gpu_launches,
                             void* stream) {
    if (!t || !obj_ids || n_obj < 1 || !variant_kinds || n_variants < 1 || n_per_block < 0 || lo < 0 || hi < lo ||
        hi > n_per_block || chunk < 1 || !K_host || !acc_hits || !acc_valid || check_n < 0 ||
        (check_n > 0 && (!check_pq || !check_pt || !check_gq || !check_gt || !check_add || !check_adds || !check_hit))) {
        set_error("p6d_sweep_run: bad arguments");
        return P6D_EINVAL;
    }
    if (gpu_launches) *gpu_launches = 0;
    if (hi == lo) return P6D_OK;
    DeviceGuard guard(t->device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    if (chunk > hi - lo) chunk = hi - lo;
    bool need_depth = false;
    for (int v = 0; v < n_variants; ++v) {
        if (variant_kinds[v] < 0 || variant_kinds[v] > 2) { set_error("p6d_sweep_run: unknown variant kind"); return P6D_EINVAL; }
        need_depth |= variant_kinds[v] == 2;
    }
    const size_t per_buf = carve(nullptr, 0, static_cast<size_t>(chunk), need_depth, nullptr);
    const int64_t n_blocks = static_cast<int64_t>(n_obj) * n_variants;
    const int64_t cn = check_n < hi - lo ? check_n : hi - lo;        // checked poses per block
    const size_t check_bytes = static_cast<size_t>(n_blocks * cn) * (16 + 12 + 16 + 12 + 4 + 4 + 1);
    char* base = nullptr;
    cudaStream_t s_gen = nullptr, s_eval = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_gen[2] = {nullptr, nullptr}, ev_eval[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
    int rc = P6D_OK;
    int launches = 0;
    auto cleanup = [&]() {
        for (int k = 0; k < 2; ++k) {
            if (ev_gen[k]) cudaEventDestroy(ev_gen[k]);
            if (ev_eval[k]) cudaEventDestroy(ev_eval[k]);
            if (ev_join[k]) cudaEventDestroy(ev_join[k]);
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (s_gen) cudaStreamDestroy(s_gen);
        if (s_eval) cudaStreamDestroy(s_eval);
        if (base) cudaFree(base);
    };
#define SWEEP_CUDA(call)                                                   \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) { rc = cuda_fail(e__, #call); cudaDeviceSynchronize(); cleanup(); return rc; } \
    } while (0)
    SWEEP_CUDA(cudaMalloc(&base, 2 * per_buf + check_bytes + 256 + 64));
    SweepBuf buf[2];
    carve(base, 0, static_cast<size_t>(chunk), need_depth, &buf[0]);
    carve(base, per_buf, static_cast<size_t>(chunk), need_depth, &buf[1]);
    float* d_K = reinterpret_cast<float*>(base + 2 * per_buf);
    char* d_check = base + 2 * per_buf + 256;
    const size_t ncheck = static_cast<size_t>(n_blocks * cn);
    float* c_pq = reinterpret_cast<float*>(d_check);
    float* c_gq = c_pq + 4 * ncheck;
    float* c_pt = c_gq + 4 * ncheck;
    float* c_gt = c_pt + 3 * ncheck;
    float* c_add = c_gt + 3 * ncheck;
    float* c_adds = c_add + ncheck;
    uint8_t* c_hit = reinterpret_cast<uint8_t*>(c_adds + ncheck);
    SWEEP_CUDA(cudaStreamCreateWithFlags(&s_gen, cudaStreamNonBlocking));
    SWEEP_CUDA(cudaStreamCreateWithFlags(&s_eval, cudaStreamNonBlocking));
    SWEEP_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) {
        SWEEP_CUDA(cudaEventCreateWithFlags(&ev_gen[k], cudaEventDisableTiming));
        SWEEP_CUDA(cudaEventCreateWithFlags(&ev_eval[k], cudaEventDisableTiming));
        SWEEP_CUDA(cudaEventCreateWithFlags(&ev_join[k], cudaEventDisableTiming));
    }
    // fork: everything below is ordered after the work already queued on the caller's stream
    SWEEP_CUDA(cudaMemcpyAsync(d_K, K_host, 9 * sizeof(float), cudaMemcpyHostToDevice, user));
    SWEEP_CUDA(cudaEventRecord(ev_fork, user));
    SWEEP_CUDA(cudaStreamWaitEvent(s_gen, ev_fork, 0));
    SWEEP_CUDA(cudaStreamWaitEvent(s_eval, ev_fork, 0));
    const size_t ns = static_cast<size_t>(t->n_slots);
    int64_t item = 0;
    bool used[2] = {false, false};
    for (int oi = 0; oi < n_obj && rc == P6D_OK; ++oi) {
        for (int vi = 0; vi < n_variants && rc == P6D_OK; ++vi) {
            const int kind = variant_kinds[vi];
            for (int64_t c0 = lo; c0 < hi && rc == P6D_OK; c0 += chunk, ++item) {
                const int64_t n = chunk < hi - c0 ? chunk : hi - c0;
                const int k = static_cast<int>(item & 1);
                SweepBuf& b = buf[k];
                // generation + translation of this chunk overlap the evaluation of the previous one
                if (used[k]) SWEEP_CUDA(cudaStreamWaitEvent(s_gen, ev_eval[k], 0));
                rc = p6d_synth_poses(seed, oi, vi, c0, n, rot_sigma, trans_sigma, kind, K_host, obj_ids[oi], b.pq, b.pt,
                                     b.gq, b.gt, b.obj, b.z, b.uv, b.kc, b.depth, t->device, s_gen);
                if (rc) break;
                ++launches;
                if (kind == 1) {
                    rc = p6d_pinhole_fwd(b.z, b.uv, d_K, 0, n, b.pt, t->device, s_gen);
                    ++launches;
                } else if (kind == 2) {
                    rc = p6d_depth_backproject(b.depth, 8, 8, b.uv, b.kc, 1, n, 7.0f, b.pt, t->device, s_gen);
                    ++launches;
                }
                if (rc) break;
                SWEEP_CUDA(cudaEventRecord(ev_gen[k], s_gen));
                SWEEP_CUDA(cudaStreamWaitEvent(s_eval, ev_gen[k], 0));
                EvalArgs a{};
                fill_eval_args(t, a);
                a.pq = b.pq; a.pt = b.pt; a.gq = b.gq; a.gt = b.gt; a.obj = b.obj; a.B = n;
                a.add = b.add; a.adds = b.adds; a.hit = b.hit; a.valid = b.valid;
                a.has_acc = 0;                  // totals by chunk_totals_kernel below, not per pose
                rc = launch_eval(t, a, true, s_eval, &launches);
                if (rc) break;
                {
                    const int64_t slot = obj_ids[oi];
                    if (slot >= 0 && slot < t->n_slots) {
                        int64_t blocks = (n + 4 * RED_T - 1) / (4 * RED_T);
                        if (blocks > static_cast<int64_t>(t->sm_count) * 4) blocks = static_cast<int64_t>(t->sm_count) * 4;
                        const size_t at = static_cast<size_t>(vi) * ns + static_cast<size_t>(slot);
                        chunk_totals_kernel<<<static_cast<unsigned>(blocks), RED_T, 0, s_eval>>>(
                            b.add, b.adds, b.hit, b.valid, n, reinterpret_cast<unsigned long long*>(acc_hits + at),
                            reinterpret_cast<unsigned long long*>(acc_valid + at), acc_add_sum ? acc_add_sum + at : nullptr,
                            acc_adds_sum ? acc_adds_sum + at : nullptr);
                        SWEEP_CUDA(cudaGetLastError());
                        ++launches;
                    }
                }
                if (cn > 0 && c0 == lo) {
                    // keep the first cn poses of the block (inputs as evaluated + outputs) for the oracle check
                    const size_t at = static_cast<size_t>((static_cast<int64_t>(oi) * n_variants + vi) * cn);
                    const size_t m = static_cast<size_t>(cn < n ? cn : n);
                    SWEEP_CUDA(cudaMemcpyAsync(c_pq + 4 * at, b.pq, 16 * m, cudaMemcpyDeviceToDevice, s_eval));
                    SWEEP_CUDA(cudaMemcpyAsync(c_gq + 4 * at, b.gq, 16 * m, cudaMemcpyDeviceToDevice, s_eval));
                    SWEEP_CUDA(cudaMemcpyAsync(c_pt + 3 * at, b.pt, 12 * m, cudaMemcpyDeviceToDevice, s_eval));
                    SWEEP_CUDA(cudaMemcpyAsync(c_gt + 3 * at, b.gt, 12 * m, cudaMemcpyDeviceToDevice, s_eval));
                    SWEEP_CUDA(cudaMemcpyAsync(c_add + at, b.add, 4 * m, cudaMemcpyDeviceToDevice, s_eval));
                    SWEEP_CUDA(cudaMemcpyAsync(c_adds + at, b.adds, 4 * m, cudaMemcpyDeviceToDevice, s_eval));
                    SWEEP_CUDA(cudaMemcpyAsync(c_hit + at, b.hit, m, cudaMemcpyDeviceToDevice, s_eval));
                }
                SWEEP_CUDA(cudaEventRecord(ev_eval[k], s_eval));
                used[k] = true;
            }
        }
    }
    if (rc != P6D_OK) {
        cudaDeviceSynchronize();
        cleanup();
        return rc;
    }
    // join: the caller's stream continues after both internal streams
    SWEEP_CUDA(cudaEventRecord(ev_join[0], s_gen));
    SWEEP_CUDA(cudaEventRecord(ev_join[1], s_eval));
    SWEEP_CUDA(cudaStreamWaitEvent(user, ev_join[0], 0));
    SWEEP_CUDA(cudaStreamWaitEvent(user, ev_join[1], 0));
    if (cn > 0) {
        SWEEP_CUDA(cudaMemcpyAsync(check_pq, c_pq, 16 * ncheck, cudaMemcpyDeviceToHost, user));
        SWEEP_CUDA(cudaMemcpyAsync(check_gq, c_gq, 16 * ncheck, cudaMemcpyDeviceToHost, user));
        SWEEP_CUDA(cudaMemcpyAsync(check_pt, c_pt, 12 * ncheck, cudaMemcpyDeviceToHost, user));
        SWEEP_CUDA(cudaMemcpyAsync(check_gt, c_gt, 12 * ncheck, cudaMemcpyDeviceToHost, user));
        SWEEP_CUDA(cudaMemcpyAsync(check_add, c_add, 4 * ncheck, cudaMemcpyDeviceToHost, user));
        SWEEP_CUDA(cudaMemcpyAsync(check_adds, c_adds, 4 * ncheck, cudaMemcpyDeviceToHost, user));
        SWEEP_CUDA(cudaMemcpyAsync(check_hit, c_hit, ncheck, cudaMemcpyDeviceToHost, user));
    }
    // the buffers are freed below: wait for the sweep (the caller's timing events on `user` still
    // bracket exactly the device work)
    SWEEP_CUDA(cudaStreamSynchronize(user));
#undef SWEEP_CUDA
    cleanup();
    if (gpu_launches) *gpu_launches = launches;
    return P6D_OK;
}
