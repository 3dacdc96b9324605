// exp_cubin.cu -- dev harness: load a (possibly re-scheduled, csrc/sass_sched.py) cubin of
// tools/exp_block.cu through the driver API, time one scan kernel and print a checksum.
// Build: nvcc -O2 -o exp_cubin exp_cubin.cu -lcuda
// Usage: exp_cubin FILE.cubin MANGLED_KERNEL T K [label]
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); \
    printf("%s failed: %s\n", #x, s_); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 5) { printf("usage: %s cubin kernel T K [label]\n", argv[0]); return 2; }
    const char* label = argc > 5 ? argv[5] : argv[1];
    const int T = atoi(argv[3]), K = atoi(argv[4]);
    CK(cuInit(0));
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    CUcontext ctx; CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
    CUmodule mod; CK(cuModuleLoad(&mod, argv[1]));
    CUfunction fn; CK(cuModuleGetFunction(&fn, mod, argv[2]));
    int sms = 0; CK(cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
    int nquads = 512, reps = 64;
    size_t smem = (size_t)3 * nquads * 16;
    CK(cuFuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
    int per_sm = 0; CK(cuOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, T, smem));
    int grid = sms * per_sm;
    std::vector<float> h(12 * nquads); srand(1);
    for (auto& v : h) v = (rand() % 2000) * 1e-4f;
    CUdeviceptr d_g, d_out;
    CK(cuMemAlloc(&d_g, h.size() * 4)); CK(cuMemAlloc(&d_out, (size_t)grid * T * 4 + (1 << 20)));
    CK(cuMemcpyHtoD(d_g, h.data(), h.size() * 4));
    int flag = -1, r1 = 1;
    void* args1[] = {&d_g, &nquads, &r1, &d_out, &flag};
    CK(cuLaunchKernel(fn, grid, 1, 1, T, 1, 1, (unsigned)smem, 0, args1, 0));
    CK(cuCtxSynchronize());
    std::vector<float> o((size_t)grid * T);
    CK(cuMemcpyDtoH(o.data(), d_out, o.size() * 4));
    unsigned long long sum = 1469598103934665603ull;
    for (float v : o) { unsigned u; memcpy(&u, &v, 4); sum = (sum ^ u) * 1099511628211ull; }
    flag = 0;
    void* args[] = {&d_g, &nquads, &reps, &d_out, &flag};
    CUevent e0, e1; CK(cuEventCreate(&e0, 0)); CK(cuEventCreate(&e1, 0));
    CK(cuLaunchKernel(fn, grid, 1, 1, T, 1, 1, (unsigned)smem, 0, args, 0));
    CK(cuCtxSynchronize());
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cuEventRecord(e0, 0));
        CK(cuLaunchKernel(fn, grid, 1, 1, T, 1, 1, (unsigned)smem, 0, args, 0));
        CK(cuEventRecord(e1, 0)); CK(cuEventSynchronize(e1));
        float ms; CK(cuEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double pairs = (double)grid * T * K * (double)nquads * 4.0 * reps;
    double tf = pairs * 8.0 / (best * 1e-3) / 1e12;
    printf("%-40s T=%4d K=%d ctas/sm=%d %8.3f ms %6.2f TFLOP/s %5.1f%% of 74.45  checksum %016llx\n", label, T, K, per_sm,
           best, tf, tf / 74.45 * 100.0, sum);
    return 0;
}
