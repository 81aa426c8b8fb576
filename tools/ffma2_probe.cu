// ffma2_probe.cu -- how fast can ONE warp (or a few) per SM sub-partition issue packed FFMA2?
// The warp-specialised block kernel runs 1 T-mix + 1 A-mix warp per sub-partition; this measures the issue-rate
// ceiling of that arrangement.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ffma2_probe tools/ffma2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
template <int NACC, bool SCALAR_B>
__global__ void k(float* out, int iters, long long* cyc) {
  float2 a[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  float w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = 1.0f + 1e-7f * (threadIdx.x + i);
  float2 x = make_float2(0.999f, 1.001f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        float2 b = SCALAR_B ? make_float2(w[(i + r) & 7], w[(i + r) & 7]) : make_float2(w[(i + r) & 7], w[(i + r + 1) & 7]);
        a[i] = __ffma2_rn(x, b, a[i]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC, bool SB>
void run(int warps) {
  float* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
  const int iters = 2000;
  k<NACC, SB><<<148, warps * 32>>>(d, iters, c);
  cudaDeviceSynchronize();
  long long cyc; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
  double per = double(cyc) / (double(iters) * 4 * NACC);
  printf("warps/CTA=%2d (%.1f per sub-partition) acc=%2d scalar_b=%d: %.2f cycles per FFMA2 per warp -> pipe utilisation %.0f%%\n", warps,
         warps / 4.0, NACC, int(SB), per, 100.0 * 2.0 * (warps / 4.0) / per);
  cudaFree(d); cudaFree(c);
}
int main() {
  for (int w : {4, 8, 16, 32}) { run<16, true>(w); run<16, false>(w); run<32, true>(w); }
  return 0;
}
