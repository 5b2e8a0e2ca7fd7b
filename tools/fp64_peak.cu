// fp64 throughput micro-benchmark for the roofline denominators of the 20-/61-state paths
// ("fp64 peak is not in MEASURED_PEAKS.json -> measure once with a DFMA/DMMA
// micro-benchmark on the box").  Three kernels, register-only, no memory traffic:
//   dmma  : mma.sync.m8n8k4.f64 (SASS DMMA), 8 independent accumulator chains per warp
//   dfma  : fma.rn.f64, 8 independent chains per thread
//   both  : the two interleaved 1:1 in the same warp (do they share the fp64 units?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(256) peak_kernel(double* out, int iters, double a, double b) {
  double c[8][2], f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = i + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2) dmma(c[i][0], c[i][1], a, b);
      if (MODE == 1 || MODE == 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
  if (s == 12345.678) out[0] = s;
}

template <int MODE>
double run(int ctas, int iters) {
  double* d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  peak_kernel<MODE><<<ctas, 256>>>(d, iters, 1.0000001, 0.9999999);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  peak_kernel<MODE><<<ctas, 256>>>(d, iters, 1.0000001, 0.9999999);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(d);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const int iters = 20000;
  for (int per = 1; per <= 8; per *= 2) {
    const int ctas = sms * per;
    const double warps = (double)ctas * 8, threads = (double)ctas * 256;
    const double t0 = run<0>(ctas, iters), t1 = run<1>(ctas, iters), t2 = run<2>(ctas, iters);
    const double fl_dmma = warps * iters * 8.0 * 512.0, fl_dfma = threads * iters * 8.0 * 2.0;
    printf("{\"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, "
           "\"both_tflops\": %.2f, \"both_ms\": %.3f, \"dmma_ms\": %.3f, \"dfma_ms\": %.3f}\n",
           per, per * 8, fl_dmma / t0 * 1e-9, fl_dfma / t1 * 1e-9, (fl_dmma + fl_dfma) / t2 * 1e-9, t2, t0, t1);
  }
  return 0;
}
