// Error handling, launch accounting, timing and the FP64 issue-peak microbenchmark.
#include <stdarg.h>
#include <atomic>

#include "pb2_common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static int g_timing = 0;
static double g_last_ms = 0.0;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

void pb2_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void pb2_count_launch(int n) { g_launches += n; }

int32_t pb2_check_launch(const char *what)
{
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        pb2_set_error("%s: %s", what, cudaGetErrorString(err));
        return (int32_t)err;
    }
    return 0;
}

void pb2_timing_begin(cudaStream_t s)
{
    if (!g_timing) return;
    if (!g_ev0) {
        cudaEventCreate(&g_ev0);
        cudaEventCreate(&g_ev1);
    }
    cudaEventRecord(g_ev0, s);
}

void pb2_timing_end(cudaStream_t s)
{
    if (!g_timing) return;
    cudaEventRecord(g_ev1, s);
    cudaEventSynchronize(g_ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_ev0, g_ev1);
    g_last_ms = ms;
}

// 8 independent DFMA chains per thread: the FP64 pipe (16 lanes / SM sub-partition) is the only
// limiter.  One DFMA counts as ONE op (a lane-instruction), matching SURVEY.md section 8d.
__global__ void __launch_bounds__(256) pb2_fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;
}

extern "C" {

int32_t pb2_abi_version(void) { return PB2_ABI_VERSION; }
const char *pb2_last_error(void) { return g_err; }
int32_t pb2_sizeof_params(void) { return (int32_t)sizeof(pb2_params); }
int32_t pb2_sizeof_catalog(void) { return (int32_t)sizeof(pb2_catalog); }
int32_t pb2_sizeof_pairs(void) { return (int32_t)sizeof(pb2_pairs); }
int32_t pb2_diag_lanes(void) { return PB2_DIAG_LANES; }
int64_t pb2_launch_count(void) { return g_launches.load(); }
int32_t pb2_set_timing(int32_t enable) { g_timing = enable; return 0; }
double pb2_last_kernel_ms(void) { return g_last_ms; }

int32_t pb2_fp64_peak(int32_t iters, double *ops_per_second, double *elapsed_ms)
{
    int dev = 0, sms = 0;
    PB2_CUDA(cudaGetDevice(&dev));
    PB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *d_out = nullptr;
    PB2_CUDA(cudaMalloc(&d_out, sizeof(double)));
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    PB2_CUDA(cudaEventCreate(&e0));
    PB2_CUDA(cudaEventCreate(&e1));
    pb2_fp64_peak_kernel<<<blocks, threads>>>(d_out, 64, 0.999999, 1e-9);  // warm-up
    PB2_CUDA(cudaEventRecord(e0));
    pb2_fp64_peak_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
    PB2_CUDA(cudaEventRecord(e1));
    PB2_CUDA(cudaEventSynchronize(e1));
    pb2_count_launch(2);
    float ms = 0.f;
    PB2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)blocks * threads * (double)iters * 64.0;
    *ops_per_second = ops / (ms * 1e-3);
    *elapsed_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return pb2_check_launch("pb2_fp64_peak");
}

}  // extern "C"
