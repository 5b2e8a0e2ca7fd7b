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

// DMMA with operands fetched from shared memory like the tile kernels do: `LPM4` quarter-LDS.64
// per DMMA (5 = one B load per DMMA + one A load per four: the 1.25 LDS / DMMA of gm_*_kernel)
template <int LPM4>
__global__ void __launch_bounds__(256) peak_lds_kernel(double* out, int iters) {
  __shared__ double tile[64 * 36];
  for (int i = threadIdx.x; i < 64 * 36; i += 256) tile[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const double* a = tile + (lane >> 2) * 36 + (lane & 3);
  const double* b = tile + (lane & 3) * 36 + (lane >> 2);
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
    const int kt = it & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double av = a[(kt * 4 + h * 32) % 28];
      if (LPM4 <= 4) av = 1.0000001;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        double bv = LPM4 >= 4 ? b[kt * 4 * 36 % 1000 + n * 8 + h] : 0.9999999;
        dmma(c[h * 4 + n][0], c[h * 4 + n][1], av, bv);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int LPM4>
double run_lds(int ctas, int iters) {
  double* d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  peak_lds_kernel<LPM4><<<ctas, 256>>>(d, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  peak_lds_kernel<LPM4><<<ctas, 256>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(d);
  return ms;
}

// dependent-issue latency: NACC independent accumulator chains per warp, WARPS warps per CTA
// (one CTA per SM), register operands
template <int NACC>
__global__ void chain_kernel(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16 / NACC; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
double run_chain(int ctas, int warps, int iters) {
  double* d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  chain_kernel<NACC><<<ctas, warps * 32>>>(d, iters, 1.0000001, 0.9999999);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  chain_kernel<NACC><<<ctas, warps * 32>>>(d, iters, 1.0000001, 0.9999999);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(d);
  return (double)ctas * warps * iters * 16.0 * 512.0 / ms * 1e-9;
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
  for (int warps = 4; warps <= 16; warps *= 2)
    printf("{\"warps_per_sm\": %d, \"chains_1_tflops\": %.2f, \"chains_2_tflops\": %.2f, "
           "\"chains_4_tflops\": %.2f, \"chains_8_tflops\": %.2f, \"chains_16_tflops\": %.2f}\n", warps,
           run_chain<1>(sms, warps, iters), run_chain<2>(sms, warps, iters),
           run_chain<4>(sms, warps, iters), run_chain<8>(sms, warps, iters),
           run_chain<16>(sms, warps, iters));
  for (int per = 1; per <= 8; per *= 2) {
    const int ctas = sms * per;
    const double warps = (double)ctas * 8, threads = (double)ctas * 256;
    const double t0 = run<0>(ctas, iters), t1 = run<1>(ctas, iters), t2 = run<2>(ctas, iters);
    const double fl_dmma = warps * iters * 8.0 * 512.0, fl_dfma = threads * iters * 8.0 * 2.0;
    const double l0 = run_lds<0>(ctas, iters), l4 = run_lds<4>(ctas, iters), l5 = run_lds<5>(ctas, iters);
    printf("{\"ctas_per_sm\": %d, \"dmma_regs_only_tflops\": %.2f, \"dmma_lds_b_tflops\": %.2f, "
           "\"dmma_lds_a_and_b_tflops\": %.2f}\n", per, fl_dmma / l0 * 1e-9, fl_dmma / l4 * 1e-9,
           fl_dmma / l5 * 1e-9);
    printf("{\"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, "
           "\"both_tflops\": %.2f, \"both_ms\": %.3f, \"dmma_ms\": %.3f, \"dfma_ms\": %.3f}\n",
           per, per * 8, fl_dmma / t0 * 1e-9, fl_dfma / t1 * 1e-9, (fl_dmma + fl_dfma) / t2 * 1e-9, t2, t0, t1);
  }
  return 0;
}
