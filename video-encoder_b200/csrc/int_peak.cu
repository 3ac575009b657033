// int_peak.cu -- microbenchmark for the roofline denominator of K1:
// the chip's sustained VABSDIFF4.U8.ACC issue rate.  MEASURED_PEAKS.json has HBM and bf16
// numbers only, so the integer-ALU peak the north_star asks for is measured here, live, by
// bench.py on the same GPU and in the same process as the kernel it is compared with.
#include "b2_common.cuh"

namespace {

constexpr int CHAINS = 16;      // independent accumulator chains per thread
constexpr int INNER = 64;       // unrolled VABSDIFF4 per chain per outer iteration

__global__ void __launch_bounds__(512, 2)
vabsdiff4_peak_kernel(uint32_t *out, int outer, uint32_t seed)
{
    uint32_t acc[CHAINS], a[CHAINS];
    const uint32_t b = seed ^ (threadIdx.x * 0x01010101u);
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { acc[c] = 0; a[c] = seed * (c + 1) + blockIdx.x; }
    for (int o = 0; o < outer; o++) {
#pragma unroll
        for (int i = 0; i < INNER; i++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) acc[c] = vsad4_acc(a[c], b, acc[c]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// variant with three distinct, never-reused source registers per instruction (operand-fetch stress)
__global__ void __launch_bounds__(512, 2)
vabsdiff4_peak3_kernel(uint32_t *out, int outer, uint32_t seed)
{
    uint32_t acc[CHAINS], a[CHAINS], b[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { acc[c] = 0; a[c] = seed * (c + 1) + blockIdx.x; b[c] = seed ^ (threadIdx.x * (c + 3)); }
    for (int o = 0; o < outer; o++) {
#pragma unroll
        for (int i = 0; i < INNER; i++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) acc[c] = vsad4_acc(a[(c + i) % CHAINS], b[(c + 2 * i + 1) % CHAINS], acc[c]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" double b2_bench_vabsdiff4_peak3(int device, int outer, int reps)
{
    if (cudaSetDevice(device) != cudaSuccess) return 0.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0.0;
    const int blocks = prop.multiProcessorCount * 2, threads = 512;
    uint32_t *d_out = nullptr;
    if (cudaMalloc(&d_out, (size_t)blocks * threads * 4) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    vabsdiff4_peak3_kernel<<<blocks, threads>>>(d_out, outer, 0x12345u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        vabsdiff4_peak3_kernel<<<blocks, threads>>>(d_out, outer, 0x12345u + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    return (double)blocks * threads * (double)outer * INNER * CHAINS / (best * 1e-3);
}

// Returns the sustained rate in VABSDIFF4 lane-instructions per second (x4 = pixel-SADs/s),
// best of `reps` timed launches of `outer` iterations each; 0 on error.
extern "C" double b2_bench_vabsdiff4_peak(int device, int outer, int reps, double *ms_best)
{
    if (cudaSetDevice(device) != cudaSuccess) return 0.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0.0;
    const int blocks = prop.multiProcessorCount * 2, threads = 512;
    uint32_t *d_out = nullptr;
    if (cudaMalloc(&d_out, (size_t)blocks * threads * 4) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    vabsdiff4_peak_kernel<<<blocks, threads>>>(d_out, outer, 0x12345u);      // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        vabsdiff4_peak_kernel<<<blocks, threads>>>(d_out, outer, 0x12345u + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    if (err != cudaSuccess) { fprintf(stderr, "b2enc: int peak bench failed: %s\n", cudaGetErrorString(err)); return 0.0; }
    if (ms_best) *ms_best = best;
    const double instr = (double)blocks * threads * (double)outer * INNER * CHAINS;
    return instr / (best * 1e-3);
}
