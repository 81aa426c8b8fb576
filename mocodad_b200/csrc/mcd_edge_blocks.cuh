// mcd_edge_blocks.cuh -- the first (2 -> 16 channels) and last (32 -> 2 channels) ST-GCN blocks of the denoiser
// (st_gcnnsp1a.0 and st_gcnnsu3.1; ST_GCNN_layer.forward, models/gcae/stsgcn.py:94-116).
//
// One side of these blocks has only the 2 coordinate channels.  The learned mixes act on positions, the 1x1
// convolution on channels, both are linear, so they commute:  W (A∘T)(X) = (A∘T)(W X).  The mixes are therefore run
// on the 2-channel side -- after the convolution in the last block, before it in the first -- which cuts their work
// 16x (8x) and leaves kernels whose cost is reading / writing the wide side once (HBM-bound).  fp32 throughout; the
// re-association changes results by rounding only (parity tests: 2e-5 against the reference per layer).
//
// One thread per (window of the tile, position (t, v)); the 2-channel planes pass through shared memory between
// the T-mix and the A-mix.  A thread's mix weights (column (t -> tq) of T[vw], column (v -> vw) of A[tq]: T + V values) do
// not depend on the tile, so the persistent loop keeps them in registers -- read through L1 per tile they were a third of the
// kernel's L1/shared-memory wavefronts, the pipe ncu shows at 80 % (profiles/r02_ncu_summary.txt).
#pragma once
#include "mcd_kernels.cuh"

namespace mcd {

template <int T_, int V_, int NW_, bool HEAD_>
struct EdgeCfg {
  static constexpr int T = T_, V = V_, NW = NW_;
  static constexpr bool HEAD = HEAD_;
  static constexpr int P = T * V;
  static constexpr int ROWS = NW * P;
  static constexpr int THREADS = (ROWS + 31) / 32 * 32;
  static constexpr int CW = HEAD ? 16 : 32;  // channels of the wide side
  static constexpr int CE = HEAD ? 16 : 2;   // output channels (embedding width)
  static constexpr int VP = (V + 3) / 4 * 4;
  static constexpr int TP4 = (T + 3) / 4 * 4;
  static constexpr int TMS = T * TP4 + 4;
  static_assert(THREADS <= 1024, "tile too large");
  // mix weights in registers (T + V per thread) -- except in the last block at long windows, whose convolution prologue and
  // fused DDPM update need the registers and the third resident CTA more (measured at T=24: 9.5 -> 10.3 ms per step with them)
  static constexpr bool REG_W = HEAD || T < 12;
  static constexpr int MIN_CTAS = (REG_W && T >= 12) ? 2 : 3;  // resident CTAs per SM the register budget is held to
};

// The block's folded 1x1 convolutions and bias travel as a kernel parameter: every use has a compile-time index, so the
// values are constant-bank operands of the FMAs instead of broadcast shared-memory loads (they were half of the kernel's
// LSU instructions: 96 of 178 per thread and tile in the first block, 128 of 210 in the last).
template <int CW, int CE>
struct EdgeConst {
  float w[2][CW][2];  // TAIL: [conv | residual conv][k][c']   HEAD: [conv | residual conv][co][k]
  float bias[CE];
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) edge_block_kernel(const BlockWeights wt, const BlockIO io,
                                                                                 const __grid_constant__ EdgeConst<Cfg::CW, Cfg::CE> ec) {
  constexpr int T = Cfg::T, V = Cfg::V, P = Cfg::P, ROWS = Cfg::ROWS, NW = Cfg::NW;
  constexpr int VP = Cfg::VP, CW = Cfg::CW, CE = Cfg::CE;
  constexpr bool HEAD = Cfg::HEAD;
  __shared__ float s0[2][ROWS];       // mix input  (2 channels)
  __shared__ float s1[2][ROWS];       // after the T-mix
  __shared__ float sEmb[NW][CE];      // Linear(SiLU(pos + cond)) of the tile's windows

  const int tid = threadIdx.x;
  const int64_t ntiles = (io.n + NW - 1) / NW;
  const bool live = tid < ROWS;
  const int wl = live ? tid / P : 0;
  const int p = tid - wl * P;
  const int tq = p / V, vw = p - tq * V;  // this thread's (frame, joint)
  const float slope = wt.prelu;
  constexpr bool REG_W = Cfg::REG_W;
  const float* tp = wt.TmE + tq * V + vw;  // [t][q][v]: the lanes of a warp (consecutive joints) read one contiguous span
  const float* ap = wt.A + tq * V * VP + vw;
  float wT[REG_W ? T : 1], wA[REG_W ? V : 1];
  if constexpr (REG_W) {
#pragma unroll
    for (int t = 0; t < T; ++t) wT[t] = live ? __ldg(tp + t * T * V) : 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) wA[v] = live ? __ldg(ap + v * VP) : 0.f;
  }

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t w = tile * NW + wl;
    const bool ok = live && w < io.n;
    float x0 = 0.f, x1 = 0.f;  // HEAD: the input coordinates; TAIL: the residual-convolution outputs
    if (ok) {
      if constexpr (HEAD) {
        const float* src = io.in + w * io.in_sn + io.in_t0 * V + p;
        x0 = __ldg(src);
        x1 = __ldg(src + io.in_sc);
        s0[0][tid] = x0;
        s0[1][tid] = x1;
      } else {
        float u0 = 0.f, u1 = 0.f;
#pragma unroll
        for (int g = 0; g < CW / 4; ++g) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(io.in + act_off(w, g, p, CW, P)));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float xk = f4get(xv, kk);
            const int k = g * 4 + kk;
            u0 = fmaf(ec.w[0][k][0], xk, u0);
            u1 = fmaf(ec.w[0][k][1], xk, u1);
            x0 = fmaf(ec.w[1][k][0], xk, x0);
            x1 = fmaf(ec.w[1][k][1], xk, x1);
          }
        }
        s0[0][tid] = u0;
        s0[1][tid] = u1;
      }
    } else if (live) {
      s0[0][tid] = 0.f;
      s0[1][tid] = 0.f;
    }
    // time/condition embedding of the tile's windows (stsgcn.py:112-114), precomputed by time_embedding_kernel
    for (int i = tid; i < NW * CE; i += Cfg::THREADS) {
      const int ewl = i / CE, co = i - ewl * CE;
      const int64_t ew = tile * NW + ewl;
      sEmb[ewl][co] = ew < io.n ? __ldg(io.emb + emb_row(io.w0, ew, io.emb_mod) * io.emb_stride + io.emb_off + co) : 0.f;
    }
    __syncthreads();

    // T-mix: y1[c][q][v] = sum_t s0[c][t][v] * Tm[v][t][q]      (this thread: q = tq, v = vw)        stsgcn.py:154
    if (live) {
      float a0 = 0.f, a1 = 0.f;
      const int base = wl * P + vw;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float k = REG_W ? wT[REG_W ? t : 0] : __ldg(tp + t * T * V);
        a0 = fmaf(s0[0][base + t * V], k, a0);
        a1 = fmaf(s0[1][base + t * V], k, a1);
      }
      s1[0][tid] = a0;
      s1[1][tid] = a1;
    }
    __syncthreads();

    // A-mix: y2[c][t][w] = sum_v s1[c][t][v] * A[t][v][w]       (this thread: t = tq, w = vw)        stsgcn.py:155
    if (live) {
      float m0 = 0.f, m1 = 0.f;
      const int base = wl * P + tq * V;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float k = REG_W ? wA[REG_W ? v : 0] : __ldg(ap + v * VP);
        m0 = fmaf(s1[0][base + v], k, m0);
        m1 = fmaf(s1[1][base + v], k, m1);
      }
      if (ok) {
        if constexpr (HEAD) {  // 1x1 conv 2 -> 16 (+ residual conv), PReLU, + emb -> planar-4
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float o[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int co = g * 4 + jj;
              float v = ec.bias[co];
              v = fmaf(ec.w[0][co][0], m0, v);
              v = fmaf(ec.w[0][co][1], m1, v);
              v = fmaf(ec.w[1][co][0], x0, v);
              v = fmaf(ec.w[1][co][1], x1, v);
              v = v > 0.f ? v : slope * v;
              o[jj] = v + sEmb[wl][co];
            }
            *reinterpret_cast<float4*>(io.out + act_off(w, g, p, 16, P)) = make_float4(o[0], o[1], o[2], o[3]);
          }
        } else {  // + residual conv, PReLU, + emb, + the U-Net's outer residual -> reference layout [n][2][P]
          float v0 = m0 + x0 + ec.bias[0], v1 = m1 + x1 + ec.bias[1];
          v0 = (v0 > 0.f ? v0 : slope * v0) + sEmb[wl][0];
          v1 = (v1 > 0.f ? v1 : slope * v1) + sEmb[wl][1];
          const int64_t e0 = (w * 2) * P + p;
          float x0v = 0.f, x1v = 0.f;
          if (io.xres != nullptr) { x0v = __ldg(io.xres + e0); x1v = __ldg(io.xres + e0 + P); v0 += x0v; v1 += x1v; }
          if (io.fuse_ddpm) {  // mocodad.py:172-178 on (x, eps = v) of this thread's two elements; same roundings as ddpm_step_kernel
            const DdpmArgs& a = io.ddpm;
            float z0 = 0.f, z1 = 0.f;
            if (a.add_noise) {
              const int64_t virt = a.virt0 + w;
              const int64_t g = virt / a.noise_B, b = virt - g * a.noise_B;
              if (a.noise != nullptr) {
                const float* np = a.noise + ((g * a.noise_slots + a.slot) * a.noise_B + b) * (2 * P);
                z0 = np[p]; z1 = np[P + p];
              } else {
                z0 = philox_normal(a.seed, uint64_t(a.first_window + b), uint32_t(g), uint32_t(a.slot), uint32_t(p));
                z1 = philox_normal(a.seed, uint64_t(a.first_window + b), uint32_t(g), uint32_t(a.slot), uint32_t(P + p));
              }
            }
            v0 = __fadd_rn(__fmul_rn(a.c1, __fsub_rn(x0v, __fmul_rn(a.c2, v0))), __fmul_rn(a.c3, z0));
            v1 = __fadd_rn(__fmul_rn(a.c1, __fsub_rn(x1v, __fmul_rn(a.c2, v1))), __fmul_rn(a.c3, z1));
          }
          io.out[e0] = v0;
          io.out[e0 + P] = v1;
        }
      }
    }
    __syncthreads();  // s0 / sEmb are rewritten by the next tile
  }
}

}  // namespace mcd
