// fp64_peak.cu - measured FP64 ceilings of the GPU the fit kernels run on (roofline_fp64 in bench.py):
//   dfma        warp-wide DFMA issue rate, 8 independent chains per thread
//   dmma        mma.sync.aligned.m8n8k4 f64 (the FP64 tensor-core path), 4 independent accumulators per warp
//   dfma+dmma   both interleaved: tells whether DMMA has a pipe of its own beside the FP64 ALUs
//   dfma+int    DFMA interleaved 1:1 with integer ALU work: the instruction mix of the fit kernels
// Prints one JSON object. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double *out, double a, double b)
{
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_dmma(double *out, double a, double b)
{
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_both(double *out, double a, double b)
{
    double c[8], f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i] = threadIdx.x + i; f[i] = threadIdx.x - i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            dmma884(c[2 * i], c[2 * i + 1], a, b);
            f[2 * i] = fma(f[2 * i], a, b);
            f[2 * i + 1] = fma(f[2 * i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma_int(double *out, double a, double b, int m)
{
    double c[8];
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i] = threadIdx.x + i; q[i] = threadIdx.x * (i + 1); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c[i] = fma(c[i], a, b);
            q[i] = max(q[i] ^ m, it);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i] + q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    const double warps = (double)blocks * threads / 32;
    const double ms_dfma = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    const double ms_dmma = time_ms([&] { k_dmma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    const double ms_both = time_ms([&] { k_both<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    const double ms_mix = time_ms([&] { k_dfma_int<<<blocks, threads>>>(out, 1.0000001, 1e-9, 5); });
    const double dfma_inst = warps * ITERS * 8;          // warp-level DFMA instructions
    const double dmma_inst = warps * ITERS * 4;          // warp-level DMMA instructions (256 FMA each)
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, "
           "\"dfma_T_lane_fma_per_s\": %.3f, \"dfma_warp_inst_per_clk_per_sm_at_max_clock\": %.3f, "
           "\"dmma_T_fma_per_s\": %.3f, \"dmma_ms\": %.3f, \"dfma_ms\": %.3f, "
           "\"both_ms\": %.3f, \"both_vs_sum\": %.3f, \"both_vs_max\": %.3f, "
           "\"dfma_with_int_ms\": %.3f, \"dfma_with_int_slowdown\": %.3f}\n",
           p.name, sms, clk, dfma_inst * 32 / (ms_dfma * 1e-3) / 1e12,
           dfma_inst / (ms_dfma * 1e-3) / ((double)clk * 1e3) / sms,
           dmma_inst * 256 / (ms_dmma * 1e-3) / 1e12, ms_dmma, ms_dfma, ms_both,
           ms_both / (ms_dfma + ms_dmma), ms_both / (ms_dfma > ms_dmma ? ms_dfma : ms_dmma),
           ms_mix, ms_mix / ms_dfma);
    return 0;
}
