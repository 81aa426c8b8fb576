// tmix_probe.cu -- the T-mix inner loop of the tensor-core block kernel in isolation: ONE warp per SM sub-partition,
// weights in registers, activations streamed from shared memory with the same software pipeline.  How many cycles per
// packed FFMA2 does the loop itself need, with nothing else on the SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmix_probe tools/tmix_probe.cu && tools/bin/tmix_probe
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float4 lds4_early(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}
template <int T, int V, int QG, int TB, int MODE>  // MODE 0: as in the kernel; 1: loads replaced by register moves
__global__ void __launch_bounds__(256, 1) k(const float* wsrc, float* out, int passes, long long* cyc) {
  extern __shared__ float sm[];
  constexpr int P = T * V, NQG = (T + QG - 1) / QG;
  float* sX = sm;                 // [4 c4][P] float4
  float* sY = sm + 4 * P * 4;     // output area
  for (int i = threadIdx.x; i < 4 * P * 4; i += blockDim.x) sX[i] = 0.001f * (i % 97);
  __syncthreads();
  const int tid = threadIdx.x;
  const int rem = tid % (V * NQG) ;
  const int v = rem / NQG, qg = rem % NQG;
  float wT[T][QG];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int q = 0; q < QG; ++q) wT[t][q] = wsrc[(v * T + t) * T + (qg * QG + q) % T];
  long long t0 = clock64();
#pragma unroll 1
  for (int pass = 0; pass < passes; ++pass) {
    const int c4 = pass & 3;
    float2 a[2][QG];
#pragma unroll
    for (int q = 0; q < QG; ++q) a[0][q] = a[1][q] = make_float2(0.f, 0.f);
    const float* xp = sX + (c4 * P + v) * 4;
    float4 xc[TB], xn[TB];
#pragma unroll
    for (int i = 0; i < TB; ++i) xn[i] = lds4_early(xp + i * V * 4);
#pragma unroll
    for (int tb = 0; tb < T; tb += TB) {
#pragma unroll
      for (int i = 0; i < TB; ++i) xc[i] = xn[i];
      if (tb + TB < T) {
#pragma unroll
        for (int i = 0; i < TB; ++i) {
          if (MODE == 0) xn[i] = lds4_early(xp + (tb + TB + i) * V * 4);
          else xn[i] = make_float4(xc[i].y, xc[i].z, xc[i].w, xc[i].x);
        }
      }
#pragma unroll
      for (int i = 0; i < TB; ++i) {
        const int t = tb + i;
        const float2 xl = make_float2(xc[i].x, xc[i].y), xh = make_float2(xc[i].z, xc[i].w);
        if (MODE == 3) {
          // accumulator a[c>>1][(c&1)*(QG/2) + qp] holds outputs (q=2qp, 2qp+1) of channel c
          const float xs[4] = {xc[i].x, xc[i].y, xc[i].z, xc[i].w};
#pragma unroll
          for (int qp = 0; qp < QG / 2; ++qp) {
            const float2 wp = make_float2(wT[t][2 * qp], wT[t][2 * qp + 1]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float2& acc = a[c >> 1][(c & 1) * (QG / 2) + qp];
              acc = __ffma2_rn(wp, make_float2(xs[c], xs[c]), acc);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < QG; ++q) {
          const float2 ww = make_float2(wT[t][q], wT[t][q]);
          if (MODE == 3) {  // weights as (q, q+1) register pairs, activations as broadcast scalars (QG even): handled below
          } else if (MODE == 2) {  // plain FFMA: four scalar fused multiply-adds per (frame, output frame)
            a[0][q].x = fmaf(xl.x, ww.x, a[0][q].x); a[0][q].y = fmaf(xl.y, ww.x, a[0][q].y);
            a[1][q].x = fmaf(xh.x, ww.x, a[1][q].x); a[1][q].y = fmaf(xh.y, ww.x, a[1][q].y);
          } else {
            a[0][q] = __ffma2_rn(xl, ww, a[0][q]);
            a[1][q] = __ffma2_rn(xh, ww, a[1][q]);
          }
        }
      }
    }
    float* yp = sY + ((c4 * QG) * 136 + rem) * 4;
#pragma unroll
    for (int q = 0; q < QG; ++q) *reinterpret_cast<float4*>(yp + q * 136 * 4) = make_float4(a[0][q].x, a[0][q].y, a[1][q].x, a[1][q].y);
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (passes < 0) out[tid] = sY[tid];
}
template <int T, int V, int QG, int TB, int MODE>
void run(const char* what, int threads = 128) {
  float *w, *o; long long* c;
  cudaMalloc(&w, V * T * T * 4); cudaMemset(w, 0, V * T * T * 4); cudaMalloc(&o, 4096); cudaMalloc(&c, 8);
  constexpr int P = T * V;
  const size_t smem = (4 * P * 4 + 4 * QG * 136 * 4) * 4;
  cudaFuncSetAttribute(k<T, V, QG, TB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int passes = 4000;
  k<T, V, QG, TB, MODE><<<148, threads, smem>>>(w, o, passes, c);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
  printf("%-34s T=%d V=%d QG=%d TB=%d: %7.1f cycles per pass, %.2f cycles per FFMA2 (%s)\n", what, T, V, QG, TB, double(cyc) / passes,
         double(cyc) / passes / (T * QG * 2), cudaGetErrorString(e));
  cudaFree(w); cudaFree(o); cudaFree(c);
}

// two 4-channel groups per pass (twice the accumulators, twice the unrolled body)
template <int T, int V, int QG, int TB>
__global__ void __launch_bounds__(128, 1) k2(const float* wsrc, float* out, int passes, long long* cyc) {
  extern __shared__ float sm[];
  constexpr int P = T * V, NQG = (T + QG - 1) / QG;
  float* sX = sm;
  float* sY = sm + 4 * P * 4;
  for (int i = threadIdx.x; i < 4 * P * 4; i += 128) sX[i] = 0.001f * (i % 97);
  __syncthreads();
  const int tid = threadIdx.x;
  const int rem = tid % (V * NQG);
  const int v = rem / NQG, qg = rem % NQG;
  float wT[T][QG];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int q = 0; q < QG; ++q) wT[t][q] = wsrc[(v * T + t) * T + (qg * QG + q) % T];
  long long t0 = clock64();
#pragma unroll 1
  for (int pass = 0; pass < passes; ++pass) {
    const int c4 = (pass & 1) * 2;
    float2 a[2][2][QG];
#pragma unroll
    for (int q = 0; q < QG; ++q) a[0][0][q] = a[0][1][q] = a[1][0][q] = a[1][1][q] = make_float2(0.f, 0.f);
    const float* xp0 = sX + (c4 * P + v) * 4;
    const float* xp1 = xp0 + P * 4;
    float4 xc[TB][2], xn[TB][2];
#pragma unroll
    for (int i = 0; i < TB; ++i) { xn[i][0] = lds4_early(xp0 + i * V * 4); xn[i][1] = lds4_early(xp1 + i * V * 4); }
#pragma unroll
    for (int tb = 0; tb < T; tb += TB) {
#pragma unroll
      for (int i = 0; i < TB; ++i) { xc[i][0] = xn[i][0]; xc[i][1] = xn[i][1]; }
      if (tb + TB < T) {
#pragma unroll
        for (int i = 0; i < TB; ++i) { xn[i][0] = lds4_early(xp0 + (tb + TB + i) * V * 4); xn[i][1] = lds4_early(xp1 + (tb + TB + i) * V * 4); }
      }
#pragma unroll
      for (int i = 0; i < TB; ++i) {
        const int t = tb + i;
#pragma unroll
        for (int q = 0; q < QG; ++q) {
          const float2 ww = make_float2(wT[t][q], wT[t][q]);
          a[0][0][q] = __ffma2_rn(make_float2(xc[i][0].x, xc[i][0].y), ww, a[0][0][q]);
          a[0][1][q] = __ffma2_rn(make_float2(xc[i][0].z, xc[i][0].w), ww, a[0][1][q]);
          a[1][0][q] = __ffma2_rn(make_float2(xc[i][1].x, xc[i][1].y), ww, a[1][0][q]);
          a[1][1][q] = __ffma2_rn(make_float2(xc[i][1].z, xc[i][1].w), ww, a[1][1][q]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* yp = sY + (((c4 + h) * QG) * 136 + rem) * 4;
#pragma unroll
      for (int q = 0; q < QG; ++q) *reinterpret_cast<float4*>(yp + q * 136 * 4) = make_float4(a[h][0][q].x, a[h][0][q].y, a[h][1][q].x, a[h][1][q].y);
    }
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (passes < 0) out[tid] = sY[tid];
}
template <int T, int V, int QG, int TB>
void run2(const char* what) {
  float *w, *o; long long* c;
  cudaMalloc(&w, V * T * T * 4); cudaMemset(w, 0, V * T * T * 4); cudaMalloc(&o, 4096); cudaMalloc(&c, 8);
  constexpr int P = T * V;
  const size_t smem = (4 * P * 4 + 4 * QG * 136 * 4) * 4;
  cudaFuncSetAttribute(k2<T, V, QG, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int passes = 4000;
  k2<T, V, QG, TB><<<148, 128, smem>>>(w, o, passes, c);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
  printf("%-34s T=%d V=%d QG=%d TB=%d: %7.1f cycles per pass, %.2f cycles per FFMA2 (%s)\n", what, T, V, QG, TB, double(cyc) / passes,
         double(cyc) / passes / (T * QG * 4), cudaGetErrorString(e));
  cudaFree(w); cudaFree(o); cudaFree(c);
}
int main() {
  run2<24, 12, 3, 3>("two groups per pass (V=12)");
  run2<24, 17, 4, 3>("two groups per pass (V=17)");
  run<24, 17, 4, 6, 3>("weight pairs x scalar act (V=17)");
  run<24, 10, 4, 6, 3>("weight pairs x scalar act (V=10)");
  run<24, 12, 3, 6, 2>("plain FFMA (V=12)");
  run<24, 17, 4, 6, 2>("plain FFMA (V=17)");
  run<24, 12, 3, 6, 0>("kernel loop (V=12)");
  run<24, 12, 3, 6, 0>("2 warps per sub-partition (V=12)", 256);
  run<24, 17, 4, 6, 0>("2 warps per sub-partition (V=17)", 256);
  run<24, 12, 3, 6, 1>("no shared loads (V=12)");
  run<24, 17, 4, 6, 0>("kernel loop (V=17)");
  run<24, 17, 4, 6, 1>("no shared loads (V=17)");
  run<24, 12, 3, 3, 0>("TB=3 (V=12)");
  run<24, 12, 3, 12, 0>("TB=12 (V=12)");
  run<24, 10, 4, 6, 0>("kernel loop (V=10, QG=4)");
  return 0;
}
