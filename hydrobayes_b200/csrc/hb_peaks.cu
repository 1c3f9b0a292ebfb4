// hb_peaks.cu -- fp64 pipe micro-benchmarks used as roofline denominators for the log-density kernels.
// MEASURED_PEAKS.json carries HBM and bf16 figures only; the fp64 tensor (DMMA) and DFMA peaks are measured here,
// on the box, with register-resident operands and many independent accumulator chains per warp.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hydrobayes_b200.h"

namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  const double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i * seed;
  const double a = 1.0 + seed * 1e-9, b = seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

}  // namespace

// Returns measured TFLOP/s: DMMA counts 512 flop per warp-level m8n8k4, DFMA 2 flop per lane.
extern "C" int hb_fp64_peaks(double* dmma_tflops, double* dfma_tflops) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return HB_ERR_NO_DEVICE; }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out = nullptr;
  if (cudaMalloc((void**)&out, 8) != cudaSuccess) return HB_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  float best_mma = 1e30f, best_fma = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dmma_peak_kernel<<<blocks, threads>>>(out, iters, 0.25);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_mma) best_mma = ms;
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(out, iters * 8, 0.25);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_fma) best_fma = ms;
  }
  const double warps = (double)blocks * threads / 32.0;
  if (dmma_tflops) *dmma_tflops = warps * iters * 16.0 * 512.0 / (best_mma * 1e-3) * 1e-12;
  if (dfma_tflops) *dfma_tflops = (double)blocks * threads * (iters * 8.0) * 16.0 * 2.0 / (best_fma * 1e-3) * 1e-12;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}
