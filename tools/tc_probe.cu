// tc_probe.cu -- standalone bring-up probe for the tcgen05 (5th-gen tensor core) TF32 path used by
// the 1x1 channel contraction of the ST-GCN block.  Not part of the library; run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tc_probe tools/tc_probe.cu && /tmp/tc_probe
// It validates, against a host fp64 reference:
//   (1) the shared-memory matrix descriptor + SWIZZLE_64B K-major layout written by ordinary threads,
//   (2) the instruction descriptor for kind::tf32, M=128, N in {16,32,64,128},
//   (3) the TMEM accumulator read-back mapping (tcgen05.ld 32x32b),
//   (4) how the hardware treats fp32 bit patterns fed as tf32 (truncate vs round),
//   (5) the accuracy of the split-precision schemes (1, 3 and 4 MMAs per product),
//   (6) MMA issue throughput.
// Every wait is bounded: a wedged barrier sets an error flag instead of hanging the GPU.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// SWIZZLE_64B, K-major: 64-byte rows, 16-byte chunk index XOR (row >> 1) & 3  (Swizzle<2,4,3> on byte addresses)
__device__ __forceinline__ uint32_t sw64(uint32_t byte_off) { return byte_off ^ (((byte_off >> 7) & 3u) << 4); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3FFF);
  d |= uint64_t(1) << 16;                          // LBO (unused for swizzled K-major): 1
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride between 8-row groups
  d |= uint64_t(1) << 46;                          // descriptor version (Blackwell)
  d |= uint64_t(layout_type) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int max_spins) {
  for (int i = 0; i < max_spins; ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi/lo split: hi = fp32 bit pattern (the MMA reads it as tf32), lo = x - tf32(x) rounded to tf32
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// mode: 0 = single MMA on raw fp32 bit patterns
//       1 = hi(raw)/lo(x - trunc(x), rna) 3 MMAs   2 = same, 4 MMAs
//       3 = hi(rna)/lo(x - hi, rna) 3 MMAs         4 = same, 4 MMAs
// A: [128][K] row-major fp32, B: [N][K] row-major fp32 (both "K-major"), D: [128][N]
template <int N>
__global__ void __launch_bounds__(128, 1) gemm_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                     int K, int mode, int reps, int* err, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: Ahi, Alo : [K/16][128 rows][64 B];  Bhi, Blo : [K/16][N rows][64 B]
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kc = K / 16;
  uint8_t* sAhi = base;
  uint8_t* sAlo = sAhi + kc * 128 * 64;
  uint8_t* sBhi = sAlo + kc * 128 * 64;
  uint8_t* sBlo = sBhi + kc * N * 64;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // fill operands (generic-proxy stores)
  auto fill = [&](const float* src, int rows, uint8_t* hi, uint8_t* lo) {
    for (int i = tid; i < rows * K; i += blockDim.x) {
      const int r = i / K, k = i - r * K;
      const float x = src[i];
      float h, l;
      if (mode <= 2) { h = x; l = tf32_rna(x - tf32_trunc(x)); }
      else { h = tf32_rna(x); l = tf32_rna(x - h); }
      const uint32_t off = uint32_t((k / 16) * rows * 64 + r * 64 + (k % 16) * 4);
      const uint32_t abs_hi = smem_u32(hi) + off, abs_lo = smem_u32(lo) + off;
      *reinterpret_cast<float*>(hi + (sw64(abs_hi) - smem_u32(hi))) = h;
      *reinterpret_cast<float*>(lo + (sw64(abs_lo) - smem_u32(lo))) = l;
    }
  };
  fill(A, 128, sAhi, sAlo);
  fill(B, N, sBhi, sBlo);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_slot;

  long long t0 = clock64();
  uint32_t parity = 0;
  bool ok = true;
  for (int rep = 0; rep < reps && ok; ++rep) {
    if (tid == 0) {
      const uint32_t idesc = make_idesc(128, N);
      uint32_t accum = 0;
      auto pass = [&](uint8_t* a, uint8_t* b) {
        for (int c = 0; c < kc; ++c)
          for (int h = 0; h < 2; ++h) {  // two K=8 steps per 64-byte row
            const uint64_t ad = make_desc(smem_u32(a) + c * 128 * 64 + h * 32, 512, 4);
            const uint64_t bd = make_desc(smem_u32(b) + c * N * 64 + h * 32, 512, 4);
            mma_tf32(tmem, ad, bd, idesc, accum);
            accum = 1;
          }
      };
      if (mode == 0) pass(sAhi, sBhi);
      else {
        if (mode == 2 || mode == 4) pass(sAlo, sBlo);  // smallest terms first
        pass(sAlo, sBhi);
        pass(sAhi, sBlo);
        pass(sAhi, sBhi);
      }
      mma_commit(smem_u32(&bar));
    }
    ok = mbar_wait(smem_u32(&bar), parity, 1 << 22);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  long long t1 = clock64();
  if (!ok) { if (tid == 0) *err = 1; }
  else {
    // TMEM -> registers: warp w reads lanes [32w, 32w+32); thread = one row of D
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c0, r);
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) D[(warp * 32 + (tid & 31)) * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  if (tid == 0 && cycles) *cycles = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u) : "memory");
}

static float host_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float host_rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int N>
int run(int K, int mode, int reps, bool report_timing) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N, 0.f);
  srand(1234 + N + K);
  for (auto& v : A) v = float(rand()) / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = float(rand()) / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD; int* dErr; long long* dCyc;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMalloc(&dErr, 4)); CK(cudaMalloc(&dCyc, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dErr, 0, 4)); CK(cudaMemset(dD, 0, D.size() * 4));
  const size_t smem = size_t(K / 16) * (128 + N) * 64 * 2 + 2048;
  CK(cudaFuncSetAttribute(gemm_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  gemm_probe<N><<<1, 128, smem>>>(dA, dB, dD, K, mode, reps, dErr, dCyc);
  CK(cudaDeviceSynchronize());
  int err = 0; long long cyc = 0;
  CK(cudaMemcpy(&err, dErr, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  if (err) { printf("N=%3d K=%3d mode=%d: BARRIER TIMEOUT\n", N, K, mode); return 1; }
  double e_exact = 0, e_trunc = 0, e_rna = 0, ref_max = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0, st = 0, sr = 0;
      for (int k = 0; k < K; ++k) {
        s += double(A[i * K + k]) * B[j * K + k];
        st += double(host_trunc(A[i * K + k])) * host_trunc(B[j * K + k]);
        sr += double(host_rna(A[i * K + k])) * host_rna(B[j * K + k]);
      }
      const double d = D[i * N + j];
      e_exact = fmax(e_exact, fabs(d - s)); e_trunc = fmax(e_trunc, fabs(d - st)); e_rna = fmax(e_rna, fabs(d - sr));
      ref_max = fmax(ref_max, fabs(s));
    }
  printf("N=%3d K=%3d mode=%d: max|D-exact|=%.3e  max|D-tf32trunc|=%.3e  max|D-tf32rna|=%.3e  (max|ref|=%.2f)", N, K, mode,
         e_exact, e_trunc, e_rna, ref_max);
  if (report_timing) {
    const int mmas = (K / 8) * (mode == 0 ? 1 : (mode == 2 || mode == 4 ? 4 : 3));
    printf("  %lld cycles / %d reps = %.1f cyc per rep (%d MMAs of 128x%dx8 -> %.1f cyc/MMA)", cyc, reps, double(cyc) / reps, mmas, N,
           double(cyc) / reps / mmas);
  }
  printf("\n");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dErr); cudaFree(dCyc);
  return 0;
}

int main() {
  int bad = 0;
  printf("== correctness, single MMA chain (mode 0) ==\n");
  bad += run<16>(16, 0, 1, false);
  bad += run<32>(16, 0, 1, false);
  bad += run<64>(32, 0, 1, false);
  bad += run<128>(64, 0, 1, false);
  if (bad) { printf("bring-up failed; skipping the rest\n"); return 1; }
  printf("== split-precision accuracy ==\n");
  for (int mode = 0; mode <= 4; ++mode) bad += run<128>(64, mode, 1, false);
  for (int mode = 1; mode <= 4; ++mode) bad += run<64>(128, mode, 1, false);
  for (int mode = 1; mode <= 4; ++mode) bad += run<32>(32, mode, 1, false);
  printf("== issue throughput (one CTA) ==\n");
  bad += run<128>(64, 1, 200, true);
  bad += run<128>(64, 2, 200, true);
  bad += run<64>(64, 1, 200, true);
  bad += run<32>(32, 1, 200, true);
  bad += run<16>(16, 1, 200, true);
  return bad;
}
