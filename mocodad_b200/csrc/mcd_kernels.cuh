// mcd_kernels.cuh -- sm_100a device code for the MoCoDAD reverse-diffusion scoring path.
//
// Data layout in HBM
//   external tensors (x, eps, noise, data) keep the reference layout  [n][C=2][T][V]  fp32;
//   activations between denoiser blocks are "PLANAR-4"                [n][C/4][P=T*V][4] fp32:
//   16-byte elements of 4 consecutive channels, positions contiguous inside a 4-channel plane.  Channels stay
//   vectorised (the graph mixes and the 1x1 convolution work on float4 channel groups) while every warp-wide
//   global access -- one position per lane -- is a contiguous 512-byte span, for loads and stores alike.
//
// Kernels (reference code each one replaces, paths relative to the reference checkout):
//   stgcn_block_kernel   ST_GCNN_layer.forward, models/gcae/stsgcn.py:94-116 (+ :143-156):
//                        T-mix -> A-mix -> 1x1 conv (+folded BN) || residual conv -> PReLU -> +emb
//   joint_resample_kernel CNN_layer over the joint axis (+ U-Net skip add),
//                        stsgcn.py:187-199 wrapped at models/stsae/stsae_unet.py:205,213,381-394
//   bottleneck_kernel    STSE.btlnk, models/stsae/stsae.py:52-55,87
//   ddpm_step_kernel     models/mocodad.py:172-178   (optional counter-based Philox noise)
//   randn_kernel         models/mocodad.py:162
//   window_loss_kernel / best_worst_kernel   models/mocodad.py:484-485, 504-512
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcd {

constexpr int kThreads = 256;
constexpr int kMaxE = 64;  // embedding_dim upper bound (reference configs use 16)

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, bool pred) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  int sz = pred ? 16 : 0;  // src-size 0 => the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src, bool pred) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  int sz = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz) : "memory");
}
// offset (floats) of the 4-channel group g of position p of window w in a planar-4 activation tensor with C channels
__device__ __forceinline__ int64_t act_off(int64_t w, int g, int p, int C, int P) { return ((w * (C >> 2) + g) * int64_t(P) + p) * 4; }

// row of the precomputed embedding table that window w of a launch reads (BlockIO::emb_mod)
__device__ __forceinline__ int64_t emb_row(int64_t w0, int64_t w, int64_t emb_mod) { return emb_mod > 0 ? (w0 + w) % emb_mod : w; }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Blackwell packed fp32 FMA (SASS FFMA2): two lanes per issue slot, same fp32 rounding as two fmaf.
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ float f4get(const float4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// ------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.
// One call per element keeps the stream independent of thread mapping, batching and rank count:
//   counter = (window_lo, window_hi, sample g, slot << 16 | element), key = seed.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t window, uint32_t sample, uint32_t slot,
                                               uint32_t elem) {
  uint32_t c0 = static_cast<uint32_t>(window), c1 = static_cast<uint32_t>(window >> 32), c2 = sample;
  uint32_t c3 = (slot << 16) | (elem & 0xFFFFu);
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  float u1 = (static_cast<float>(c0 >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
  float u2 = static_cast<float>(c1 >> 8) * (1.0f / 16777216.0f);          // [0,1)
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ------------------------------------------------------------------------------------------
// ST_GCNN block
// ------------------------------------------------------------------------------------------
struct BlockWeights {
  const float* A;     // [T][V][VP]      gcn.A, rows zero-padded to VP = roundup4(V)
  const float* Tm;    // [V][TMS]        gcn.T as [v][t*TP4 + q], TMS = T*TP4 + 4
  const float* TmE;   // [T][T][V]       gcn.T as [t][q][v] (edge blocks: a warp's lanes = consecutive joints read one span)
  const float* Wt;    // [CINP][COUT]    BN-folded tcn conv, transposed (k-major)
  const float* Wrt;   // [CINP][COUT]    BN-folded residual conv (nullptr: identity residual)
  const float* bias;  // [COUT]          folded biases (tcn + residual)
  const float* WEt;   // [E][COUT]       emb_layer.1.weight transposed (nullptr: no embedding)
  const float* bE;    // [COUT]
  const float* Bop;   // tensor-core path: per 16-channel chunk [W hi | W lo | Wr hi | Wr lo] x [COUT][16], tf32 split,
                      // pre-arranged in the UMMA K-major SWIZZLE_64B layout (nullptr: block not packed for it)
  float prelu;        // prelu.weight[0]
};

// DDPM update x <- c1 (x - c2 eps) + c3 z (models/mocodad.py:172-178): arguments of ddpm_step_kernel, and of the last denoiser
// block when the update is fused into it
struct DdpmArgs {
  float c1, c2, c3;       // 1/sqrt(alpha_t), (1-alpha_t)/sqrt(1-alpha_hat_t), sqrt(beta_t)
  int add_noise;          // 0 at t == 1
  const float* noise;     // nullptr => Philox
  int64_t noise_B;        // windows per sample B: virtual index = g*B + b (always set)
  int32_t noise_slots;    // N-1
  int32_t slot;           // which slot this step reads / Philox slot id
  int64_t virt0;          // virtual index of window 0 of this call (for noise addressing)
  uint64_t seed;
  int64_t first_window;   // global (dataset) id of window b = 0 (Philox counter base)
};

struct BlockIO {
  const float* in;    // channel-last [n][P][CIN]  (IN_CL)  or channel-first source (IN_CF)
  float* out;         // channel-last [n][P][COUT] (OUT_CL) or [n][2][P] eps      (OUT_EPS)
  int64_t n;          // windows
  int64_t in_sn;      // IN_CF: element (w,c,t,v) = in[w*in_sn + c*in_sc + (t+in_t0)*V + v]
  int32_t in_sc, in_t0;
  const float* xres;  // OUT_EPS: eps = out + xres (the U-Net input, stsae_unet.py:401); nullptr: no add
  const float* pos;   // [E] pos_encoding(t) of this step
  const float* cond;  // [condB][E] conditioning embedding (nullptr: none); window w uses row (w0+w) % condB
  int64_t condB;
  int64_t w0;         // virtual index of window 0 of this launch
  long long* trace;   // debug: CTA 0 records (role, pair, event, clock64) quadruples here (nullptr: off)
  int trace_cap;      // capacity in records
  int32_t E;
  // denoiser blocks: time/condition embedding precomputed by time_embedding_kernel, [rows][emb_stride] floats, this block's
  // COUT values at column emb_off.  The embedding depends on the window only through its conditioning row, so the table
  // has emb_mod rows (one per conditioning row; window w reads row (w0 + w) % emb_mod) -- or, emb_mod == 0, one row per window.
  const float* emb;
  int32_t emb_stride, emb_off;
  int64_t emb_mod;
  // last denoiser block only: apply the DDPM update to its eps and write x_{t-1} to `out` (= the x buffer, in place: every
  // element is read and written by the same thread) instead of eps; saves the round trip of eps through HBM and a launch
  int32_t fuse_ddpm;
  DdpmArgs ddpm;
  // conv-first blocks with the fused up-path CNN_layer (mcd_block_cf.cuh, CfCfg::VUP > 0): the U-Net skip tensor -- must be
  // `out` itself: the resampled block output is accumulated onto it in place
  const float* skip;
};

enum { IN_CL = 0, IN_CF = 1 };
enum { OUT_CL = 0, OUT_EPS = 1 };

template <int T_, int V_, int CIN_, int COUT_, int NW_, bool EMB_, int INMODE_, int OUTMODE_>
struct BlockCfg {
  static constexpr int T = T_, V = V_, CIN = CIN_, COUT = COUT_, NW = NW_;
  static constexpr bool EMB = EMB_;
  static constexpr int INMODE = INMODE_, OUTMODE = OUTMODE_;
  static constexpr int P = T * V;
  static constexpr int ROWS = NW * P;                 // (window, frame, joint) rows per tile
  static constexpr int CINP = CIN < 4 ? 4 : CIN;      // input channels padded to a float4
  static constexpr int KC = CINP < 16 ? CINP : 16;    // channels per pipeline chunk
  static constexpr int NCHUNK = CINP / KC;
  static constexpr int C4 = KC / 4;
  static constexpr int CP = KC + 4;                   // smem row stride (floats): conflict-free conv reads
  static constexpr int VP = (V + 3) / 4 * 4;
  static constexpr int TP4 = (T + 3) / 4 * 4;
  static constexpr int TMS = T * TP4 + 4;
  static constexpr bool RESCONV = CIN != COUT;
  // 1x1 conv register tile: TPR rows x TCO output channels per thread
  static constexpr int TCO = COUT >= 128 ? 16 : (COUT >= 8 ? 8 : COUT);
  static constexpr int NCG = COUT / TCO;
  static constexpr int NPG = kThreads / NCG;
  static constexpr int TPR = (ROWS + NPG - 1) / NPG;
  // T-mix tile: 4 channels x TQ output frames;  A-mix tile: 4 channels x TW output joints
  static constexpr int TQ = TP4 <= 8 ? TP4 : (TP4 % 8 == 0 ? 8 : 4);
  static constexpr int NQT = TP4 / TQ;
  static constexpr int NWT = 2;
  static constexpr int TW = VP / NWT;
  static_assert(TP4 % TQ == 0 && VP % (2 * NWT) == 0, "tile shapes");
  static_assert(CINP % KC == 0 && COUT % TCO == 0 && kThreads % NCG == 0, "channel tiling");
  static_assert(TCO % 2 == 0 && (TCO % 4 == 0 || NCG == 1), "column tile: float4 groups, or the 2-channel output block");
  // shared memory carve-up (floats)
  static constexpr int SM_A = 0;
  static constexpr int SM_TM = SM_A + T * V * VP;
  static constexpr int SM_W = SM_TM + V * TMS;
  static constexpr int SM_WR = SM_W + CINP * COUT;
  static constexpr int SM_BIAS = SM_WR + (RESCONV ? CINP * COUT : 0);
  static constexpr int SM_EMB = SM_BIAS + COUT;
  static constexpr int SM_S = SM_EMB + NW * COUT;
  static constexpr int SM_X = (SM_S + NW * kMaxE + 3) / 4 * 4;
  static constexpr int SM_Y1 = SM_X + 2 * ROWS * CP;
  static constexpr int SM_Y2 = SM_Y1 + ROWS * CP;
  static constexpr int SM_TOTAL = SM_Y2 + ROWS * CP;
  static constexpr size_t SMEM_BYTES = size_t(SM_TOTAL) * sizeof(float);
};

template <class Cfg>
__device__ __forceinline__ void block_prefetch(const BlockIO& io, float* sXbuf, int64_t tile, int chunk, int tid) {
  constexpr int ROWS = Cfg::ROWS, CP = Cfg::CP, C4 = Cfg::C4, P = Cfg::P;
  if constexpr (Cfg::INMODE == IN_CL) {
    for (int idx = tid; idx < ROWS * C4; idx += kThreads) {
      const int j = idx / ROWS, r = idx - j * ROWS;  // consecutive threads -> consecutive positions of one 4-channel plane
      const int wl = r / P, pp = r - wl * P;
      const int64_t w = tile * Cfg::NW + wl;
      const bool ok = w < io.n;
      const float* src = ok ? io.in + act_off(w, chunk * C4 + j, pp, Cfg::CIN, P) : io.in;
      cp_async16(sXbuf + r * CP + j * 4, src, ok);
    }
  } else {  // channel-first, CIN real channels (2), padded channels stay zero (set once at start)
    for (int idx = tid; idx < ROWS * Cfg::CIN; idx += kThreads) {
      int c = idx / ROWS, r = idx - c * ROWS;
      int wl = r / P, p = r - wl * P;
      int64_t w = tile * Cfg::NW + wl;
      bool ok = w < io.n;
      const float* src = ok ? io.in + w * io.in_sn + int64_t(c) * io.in_sc + io.in_t0 * Cfg::V + p : io.in;
      cp_async4(sXbuf + r * CP + c, src, ok);
    }
  }
}

template <class Cfg>
__global__ void __launch_bounds__(kThreads, 1) stgcn_block_kernel(const BlockWeights wt, const BlockIO io) {
  constexpr int T = Cfg::T, V = Cfg::V, P = Cfg::P, ROWS = Cfg::ROWS, CP = Cfg::CP, C4 = Cfg::C4;
  constexpr int COUT = Cfg::COUT, CINP = Cfg::CINP, KC = Cfg::KC, NCHUNK = Cfg::NCHUNK;
  constexpr int VP = Cfg::VP, TP4 = Cfg::TP4, TMS = Cfg::TMS, NW = Cfg::NW;
  constexpr int TCO = Cfg::TCO, NCG = Cfg::NCG, NPG = Cfg::NPG, TPR = Cfg::TPR;
  constexpr int TQ = Cfg::TQ, NQT = Cfg::NQT, TW = Cfg::TW, NWT = Cfg::NWT;
  constexpr bool RESCONV = Cfg::RESCONV;

  extern __shared__ __align__(16) float smem[];
  float* sA = smem + Cfg::SM_A;
  float* sTm = smem + Cfg::SM_TM;
  float* sW = smem + Cfg::SM_W;
  float* sWr = smem + Cfg::SM_WR;
  float* sBias = smem + Cfg::SM_BIAS;
  float* sEmb = smem + Cfg::SM_EMB;
  float* sS = smem + Cfg::SM_S;
  float* sX = smem + Cfg::SM_X;
  float* sY1 = smem + Cfg::SM_Y1;
  float* sY2 = smem + Cfg::SM_Y2;

  const int tid = threadIdx.x;
  const int64_t ntiles = (io.n + NW - 1) / NW;
  if (int64_t(blockIdx.x) >= ntiles) return;

  // ---- once per CTA: weights -> smem (persistent over all tiles of this CTA) ----
  for (int i = tid; i < T * V * VP; i += kThreads) sA[i] = wt.A[i];
  for (int i = tid; i < V * TMS; i += kThreads) sTm[i] = wt.Tm[i];
  for (int i = tid; i < CINP * COUT; i += kThreads) sW[i] = wt.Wt[i];
  if constexpr (RESCONV)
    for (int i = tid; i < CINP * COUT; i += kThreads) sWr[i] = wt.Wrt[i];
  for (int i = tid; i < COUT; i += kThreads) sBias[i] = wt.bias[i];
  if constexpr (Cfg::INMODE == IN_CF)
    for (int i = tid; i < 2 * ROWS * CP; i += kThreads) sX[i] = 0.f;
  __syncthreads();

  // conv register tile of this thread
  const int cg = tid % NCG, pg = tid / NCG;
  constexpr int TCO2 = TCO / 2;
  float2 acc[TPR][TCO2];
#pragma unroll
  for (int i = 0; i < TPR; ++i)
#pragma unroll
    for (int j = 0; j < TCO2; ++j) acc[i][j] = make_float2(0.f, 0.f);

  // flattened (tile, chunk) pipeline: the next pair streams in while this one is computed
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t npairs = my_tiles * NCHUNK;
  block_prefetch<Cfg>(io, sX, blockIdx.x, 0, tid);
  cp_async_commit();

  for (int64_t it = 0; it < npairs; ++it) {
    const int64_t tile = blockIdx.x + (it / NCHUNK) * gridDim.x;
    const int chunk = int(it % NCHUNK);
    const int kc0 = chunk * KC;
    float* sXc = sX + (it & 1) * ROWS * CP;

    cp_async_wait_all();
    __syncthreads();  // X chunk visible; everyone is done with the previous pair's buffers
    if (it + 1 < npairs) {
      const int64_t ntile = blockIdx.x + ((it + 1) / NCHUNK) * gridDim.x;
      block_prefetch<Cfg>(io, sX + ((it + 1) & 1) * ROWS * CP, ntile, int((it + 1) % NCHUNK), tid);
      cp_async_commit();
    }

    // ---- per-tile: time/condition embedding  emb = Linear(SiLU(pos + cond)), stsgcn.py:112-114
    if constexpr (Cfg::EMB) {
      if (chunk == 0) {
        const int E = io.E;
        for (int i = tid; i < NW * E; i += kThreads) {
          int wl = i / E, j = i - wl * E;
          int64_t w = tile * NW + wl;
          float v = io.pos[j];
          if (io.cond != nullptr && w < io.n) v += io.cond[((io.w0 + w) % io.condB) * E + j];
          sS[wl * kMaxE + j] = v / (1.0f + expf(-v));
        }
      }
    }

    // ---- phase 1: T-mix   Y1[n,(q,v),c] = sum_t X[n,(t,v),c] * Tm[v][t][q]     stsgcn.py:154
    for (int task = tid; task < NQT * NW * V * C4; task += kThreads) {
      const int c4 = task % C4;
      const int col = (task / C4) % (NW * V);  // (window-in-tile, joint)
      const int qt = task / (C4 * NW * V);
      const int wl = col / V, v = col - wl * V;
      float2 a[2][TQ];  // [channel pair][output frame]
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int q = 0; q < TQ; ++q) a[c][q] = make_float2(0.f, 0.f);
      const float* xp = sXc + (wl * P + v) * CP + c4 * 4;
      const float* tp = sTm + v * TMS + qt * TQ;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float4 x = *reinterpret_cast<const float4*>(xp + t * V * CP);
        const float2 xlo = make_float2(x.x, x.y), xhi = make_float2(x.z, x.w);
#pragma unroll
        for (int q4 = 0; q4 < TQ / 4; ++q4) {
          const float4 w = *reinterpret_cast<const float4*>(tp + t * TP4 + q4 * 4);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const float wv = f4get(w, qq);
            const float2 ww = make_float2(wv, wv);
            a[0][q4 * 4 + qq] = ffma2(xlo, ww, a[0][q4 * 4 + qq]);
            a[1][q4 * 4 + qq] = ffma2(xhi, ww, a[1][q4 * 4 + qq]);
          }
        }
      }
      float* yp = sY1 + (wl * P + v) * CP + c4 * 4;
#pragma unroll
      for (int q = 0; q < TQ; ++q) {
        const int qq = qt * TQ + q;
        if (qq < T) *reinterpret_cast<float4*>(yp + qq * V * CP) = make_float4(a[0][q].x, a[0][q].y, a[1][q].x, a[1][q].y);
      }
    }
    if constexpr (Cfg::EMB) {
      if (chunk == 0) {
        __syncthreads();  // sS complete (rare path: once per tile)
        const int E = io.E;
        for (int i = tid; i < NW * COUT; i += kThreads) {
          int wl = i / COUT, co = i - wl * COUT;
          float e = wt.bE[co];
          for (int j = 0; j < E; ++j) e = fmaf(wt.WEt[j * COUT + co], sS[wl * kMaxE + j], e);
          sEmb[i] = e;
        }
      }
    }
    __syncthreads();

    // ---- phase 2: A-mix   Y2[n,(t,w),c] = sum_v Y1[n,(t,v),c] * A[t][v][w]      stsgcn.py:155
    for (int task = tid; task < NW * T * NWT * C4; task += kThreads) {
      const int c4 = task % C4;
      const int wtile = (task / C4) % NWT;
      const int row = task / (C4 * NWT);  // (window-in-tile, frame)
      const int wl = row / T, q = row - wl * T;
      float2 a[2][TW];  // [channel pair][output joint]
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < TW; ++j) a[c][j] = make_float2(0.f, 0.f);
      const float* yp = sY1 + (wl * P + q * V) * CP + c4 * 4;
      const float* ap = sA + q * V * VP + wtile * TW;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 y = *reinterpret_cast<const float4*>(yp + v * CP);
        const float2 ylo = make_float2(y.x, y.y), yhi = make_float2(y.z, y.w);
#pragma unroll
        for (int j2 = 0; j2 < TW / 2; ++j2) {
          const float2 w = *reinterpret_cast<const float2*>(ap + v * VP + j2 * 2);
          const float2 w0 = make_float2(w.x, w.x), w1 = make_float2(w.y, w.y);
          a[0][j2 * 2] = ffma2(ylo, w0, a[0][j2 * 2]);
          a[1][j2 * 2] = ffma2(yhi, w0, a[1][j2 * 2]);
          a[0][j2 * 2 + 1] = ffma2(ylo, w1, a[0][j2 * 2 + 1]);
          a[1][j2 * 2 + 1] = ffma2(yhi, w1, a[1][j2 * 2 + 1]);
        }
      }
      float* zp = sY2 + (wl * P + q * V) * CP + c4 * 4;
#pragma unroll
      for (int j = 0; j < TW; ++j) {
        const int w = wtile * TW + j;
        if (w < V) *reinterpret_cast<float4*>(zp + w * CP) = make_float4(a[0][j].x, a[0][j].y, a[1][j].x, a[1][j].y);
      }
    }
    __syncthreads();

    // ---- phase 3: 1x1 conv accumulate  acc[r][co] += Y2[r][k]*W[k][co] (+ X[r][k]*Wr[k][co])
    // Column ownership is interleaved in groups of 4: thread cg owns columns g*(NCG*4) + cg*4 + [0,4),
    // so the NCG lanes of a row group read one contiguous 16*NCG-byte span of W[k][:] (no bank conflicts).
    {
      int rows[TPR];
#pragma unroll
      for (int i = 0; i < TPR; ++i) {
        int r = pg + i * NPG;
        rows[i] = r < ROWS ? r : ROWS - 1;  // clamp: padded slots recompute the last row, never stored
      }
      auto accumulate = [&](const float* __restrict__ sIn, const float* __restrict__ sWt) {
#pragma unroll
        for (int k4 = 0; k4 < C4; ++k4) {
          float4 yv[TPR];
#pragma unroll
          for (int i = 0; i < TPR; ++i) yv[i] = *reinterpret_cast<const float4*>(sIn + rows[i] * CP + k4 * 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            float2 w[TCO2];
            const float* wp = sWt + (kc0 + k4 * 4 + kk) * COUT;
            if constexpr (TCO % 4 == 0) {
#pragma unroll
              for (int g = 0; g < TCO / 4; ++g) {
                const float4 t4 = *reinterpret_cast<const float4*>(wp + g * (NCG * 4) + cg * 4);
                w[g * 2] = make_float2(t4.x, t4.y);
                w[g * 2 + 1] = make_float2(t4.z, t4.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < TCO2; ++j) w[j] = *reinterpret_cast<const float2*>(wp + cg * TCO + j * 2);
            }
#pragma unroll
            for (int i = 0; i < TPR; ++i) {
              const float y = f4get(yv[i], kk);
              const float2 yy = make_float2(y, y);
#pragma unroll
              for (int j = 0; j < TCO2; ++j) acc[i][j] = ffma2(yy, w[j], acc[i][j]);
            }
          }
        }
      };
      accumulate(sY2, sW);
      if constexpr (RESCONV) accumulate(sXc, sWr);
      if constexpr (!RESCONV) {  // identity residual: a 4-column group lies inside exactly one K chunk
#pragma unroll
        for (int g = 0; g < (TCO + 3) / 4; ++g) {
          const int co0 = TCO % 4 == 0 ? g * (NCG * 4) + cg * 4 : cg * TCO;
          if (co0 / KC == chunk) {
#pragma unroll
            for (int i = 0; i < TPR; ++i) {
              if constexpr (TCO % 4 == 0) {
                const float4 xv = *reinterpret_cast<const float4*>(sXc + rows[i] * CP + (co0 - kc0));
                acc[i][g * 2].x += xv.x; acc[i][g * 2].y += xv.y;
                acc[i][g * 2 + 1].x += xv.z; acc[i][g * 2 + 1].y += xv.w;
              } else {
#pragma unroll
                for (int j = 0; j < TCO2; ++j) {
                  acc[i][j].x += sXc[rows[i] * CP + (co0 - kc0) + j * 2];
                  acc[i][j].y += sXc[rows[i] * CP + (co0 - kc0) + j * 2 + 1];
                }
              }
            }
          }
        }
      }

      // ---- epilogue on the last chunk: bias, PReLU, +emb, store                stsgcn.py:109-114
      if (chunk == NCHUNK - 1) {
        const float slope = wt.prelu;
#pragma unroll
        for (int i = 0; i < TPR; ++i) {
          const int r = pg + i * NPG;
          const int wl = r / P;
          const int64_t w = tile * NW + wl;
          const bool ok = (r < ROWS) && (w < io.n);
          const float* embp = sEmb + (ok ? wl : 0) * COUT;
          auto finish = [&](float v, int co) {
            v += sBias[co];
            v = v > 0.f ? v : slope * v;
            if constexpr (Cfg::EMB) v += embp[co];
            return v;
          };
          if constexpr (Cfg::OUTMODE == OUT_CL) {
            static_assert(Cfg::OUTMODE != OUT_CL || TCO % 4 == 0, "planar-4 output needs whole 4-channel groups");
            const int pp = r - wl * P;
#pragma unroll
            for (int g = 0; g < TCO / 4; ++g) {
              const int co0 = g * (NCG * 4) + cg * 4;
              const float4 o = make_float4(finish(acc[i][g * 2].x, co0), finish(acc[i][g * 2].y, co0 + 1),
                                           finish(acc[i][g * 2 + 1].x, co0 + 2), finish(acc[i][g * 2 + 1].y, co0 + 3));
              if (ok) *reinterpret_cast<float4*>(io.out + act_off(w, co0 >> 2, pp, COUT, P)) = o;
            }
          } else {  // OUT_EPS: reference layout [n][COUT][P], plus the U-Net's outer residual +X
            static_assert(Cfg::OUTMODE == OUT_CL || TCO == 2, "OUT_EPS is the 2-channel output block");
            const int p = r - wl * P;
            if (ok) {
#pragma unroll
              for (int j = 0; j < TCO2; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int co = cg * TCO + j * 2 + h;
                  const int64_t e = (w * COUT + co) * P + p;
                  const float o = finish(h == 0 ? acc[i][j].x : acc[i][j].y, co);
                  io.out[e] = io.xres != nullptr ? o + io.xres[e] : o;
                }
              }
            }
          }
#pragma unroll
          for (int j = 0; j < TCO2; ++j) acc[i][j] = make_float2(0.f, 0.f);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Joint-axis resample (CNN_layer + folded BN) with optional skip add.  HBM-bound.
//   out[n,(t,w),c] = b'[w] + sum_v Wd'[w][v] * in[n,(t,v),c]  (+ skip[n,(t,w),c])
// A planar-4 tensor is a flat array of frames (frame f = (window, 4-channel plane, t) holds V consecutive 16-byte
// elements), so a CTA works on RS_FRAMES consecutive frames: the input slab is one contiguous span that is staged
// into shared memory with coalesced 16-byte cp.async copies (rows padded to an odd number of elements: the
// per-frame reads below are then bank-conflict free), one thread computes all VOUT joints of one frame from
// registers, the results go back through the same shared-memory region and leave as one contiguous span (+ skip).
// The folded weights arrive as a kernel parameter: FFMA reads them straight from the constant bank.
// (The first version had one thread read / write its frame directly from global memory: every warp-wide access
//  touched 32 different cache lines and the kernel sat at 4.4 TB/s with the load/store pipe 65 % busy.)
// ------------------------------------------------------------------------------------------
constexpr int kRsFrames = 128;  // frames per CTA tile = threads per CTA
template <int VIN, int VOUT>
struct ResampleParams {
  float w[VOUT][VIN];
  float b[VOUT];
};
template <int VIN, int VOUT>
struct ResampleCfg {
  static constexpr int VINP = VIN | 1, VOUTP = VOUT | 1;  // padded row lengths (16-byte elements)
  static constexpr int ROWP = VINP > VOUTP ? VINP : VOUTP;
  static constexpr size_t SMEM_BYTES = size_t(kRsFrames) * ROWP * 16;
};

template <int VIN, int VOUT>
__global__ void __launch_bounds__(kRsFrames) joint_resample_kernel(const float* __restrict__ in,
                                                                    const float* __restrict__ skip,
                                                                    float* __restrict__ out,
                                                                    const __grid_constant__ ResampleParams<VIN, VOUT> prm,
                                                                    int64_t frames) {
  using Cfg = ResampleCfg<VIN, VOUT>;
  extern __shared__ float4 rs_smem[];
  const int tid = threadIdx.x;
  const int64_t ntiles = (frames + kRsFrames - 1) / kRsFrames;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t f0 = tile * kRsFrames;
    const int nf = int(frames - f0 < kRsFrames ? frames - f0 : kRsFrames);
    // 1. input slab -> shared memory [frame][VINP]
    const float* src = in + f0 * VIN * 4;
    for (int i = tid; i < nf * VIN; i += kRsFrames) {
      const int f = i / VIN, v = i - f * VIN;
      cp_async16(reinterpret_cast<float*>(rs_smem + f * Cfg::VINP + v), src + size_t(i) * 4, true);
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    // 2. one frame per thread, inputs in registers
    float4 x[VIN];
#pragma unroll
    for (int v = 0; v < VIN; ++v) x[v] = rs_smem[tid * Cfg::VINP + v];
    __syncthreads();  // everyone holds its frame: the region is reused for the outputs
#pragma unroll
    for (int w = 0; w < VOUT; ++w) {
      float4 a = make_float4(prm.b[w], prm.b[w], prm.b[w], prm.b[w]);
#pragma unroll
      for (int v = 0; v < VIN; ++v) {
        const float k = prm.w[w][v];
        a.x = fmaf(k, x[v].x, a.x); a.y = fmaf(k, x[v].y, a.y);
        a.z = fmaf(k, x[v].z, a.z); a.w = fmaf(k, x[v].w, a.w);
      }
      rs_smem[tid * Cfg::VOUTP + w] = a;
    }
    __syncthreads();
    // 3. output slab (+ skip) -> global, coalesced
    float4* dst = reinterpret_cast<float4*>(out + f0 * VOUT * 4);
    const float4* sk = skip ? reinterpret_cast<const float4*>(skip + f0 * VOUT * 4) : nullptr;
    for (int i = tid; i < nf * VOUT; i += kRsFrames) {
      const int f = i / VOUT, w = i - f * VOUT;
      float4 a = rs_smem[f * Cfg::VOUTP + w];
      if (sk) {
        const float4 sv = __ldg(sk + i);
        a.x += sv.x; a.y += sv.y; a.z += sv.z; a.w += sv.w;
      }
      dst[i] = a;
    }
    __syncthreads();  // the next tile's copies overwrite the region
  }
}

// ------------------------------------------------------------------------------------------
// Score assembly, first stage (models/mocodad.py:386-391 over compute_var_matrix, utils/eval_utils.py:27-34): every window
// spreads its loss over the frames it covers, and a frame of a (transformation, clip, person) row keeps the maximum over the
// windows that contain it; frames never covered stay 0.  One thread per (window, frame of the window); the maximum of
// non-negative floats is the maximum of their bit patterns as signed integers (negative losses lose against the 0 fill exactly
// like numpy's maximum against the zero-initialised row), so atomicMax gives the reference's float32 values bit for bit,
// in any order.  Frame numbers are 1-based; index -1 wraps to the row's last frame like numpy's fancy assignment.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) frame_scores_kernel(const float* __restrict__ loss, const int64_t* __restrict__ frames,
                                                                const int64_t* __restrict__ row, const int32_t* __restrict__ row_len,
                                                                int64_t n_items, int seg_len, int64_t stride, float* __restrict__ out) {
  const int64_t total = n_items * seg_len;
  for (int64_t i = blockIdx.x * int64_t(kThreads) + threadIdx.x; i < total; i += int64_t(gridDim.x) * kThreads) {
    const int64_t n = i / seg_len;
    const int64_t r = row[n];
    if (r < 0) continue;  // window of a clip without ground truth: not scored
    const int len = row_len[r];
    int64_t f = frames[i] - 1;
    if (f < 0) f += len;
    if (f < 0 || f >= len || f >= stride) continue;
    atomicMax(reinterpret_cast<int*>(out + r * stride + f), __float_as_int(loss[n]));
  }
}

// ------------------------------------------------------------------------------------------
// Time / condition embedding of every denoiser block, for all windows of a launch:
//   emb[w][off_b + co] = bE_b[co] + sum_j WE_b[co][j] * SiLU(pos_t[j] + cond[w][j])        stsgcn.py:112-114 (emb_layer),
//   temb = pos_encoding(t) + cond_emb                                                      stsae_unet.py:425-426
// Computed once per denoiser call (0.03 % of its traffic) instead of once per tile inside the block kernels, where
// the dependent global loads (condition row, then E rounds of weights) cost ~3.7 k cycles per tile on the epilogue
// warps -- the bottleneck of the blocks with few channel chunks (wait accounting, tools/trace_block.py).
// ------------------------------------------------------------------------------------------
constexpr int kEmbBlocks = 11;   // all ST_GCNN blocks of the denoiser
constexpr int kEmbThreads = 128;
constexpr int kEmbWin = 8;       // windows per CTA step
struct EmbTable {
  const float* WEt[kEmbBlocks];  // [E][cout]
  const float* bE[kEmbBlocks];   // [cout]
  int32_t cout[kEmbBlocks], off[kEmbBlocks];
  int32_t nblocks, total;        // total = sum of cout (row length of emb)
};
__global__ void __launch_bounds__(kEmbThreads) time_embedding_kernel(const EmbTable tb, const float* __restrict__ pos,
                                                                      const float* __restrict__ cond, int64_t condB, int64_t w0,
                                                                      int64_t n, int E, float* __restrict__ emb) {
  extern __shared__ float emb_smem[];
  float* sW = emb_smem;                       // [E][total]
  float* sB = sW + size_t(E) * tb.total;      // [total]
  float* sS = sB + tb.total;                  // [kEmbWin][E]
  const int tid = threadIdx.x, total = tb.total;
  for (int b = 0; b < tb.nblocks; ++b) {
    const int co_n = tb.cout[b], off = tb.off[b];
    for (int i = tid; i < E * co_n; i += kEmbThreads) {
      const int j = i / co_n, co = i - j * co_n;
      sW[j * total + off + co] = tb.WEt[b][i];
    }
    for (int i = tid; i < co_n; i += kEmbThreads) sB[off + i] = tb.bE[b][i];
  }
  __syncthreads();
  const int64_t ngroups = (n + kEmbWin - 1) / kEmbWin;
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int64_t wbase = g * kEmbWin;
    for (int i = tid; i < kEmbWin * E; i += kEmbThreads) {
      const int wl = i / E, j = i - wl * E;
      const int64_t w = wbase + wl;
      float v = __ldg(pos + j);
      if (cond != nullptr && w < n) v += __ldg(cond + ((w0 + w) % condB) * E + j);
      sS[i] = v / (1.0f + expf(-v));  // SiLU
    }
    __syncthreads();
    for (int o = tid; o < total; o += kEmbThreads) {
      float e[kEmbWin];
#pragma unroll
      for (int wl = 0; wl < kEmbWin; ++wl) e[wl] = sB[o];
      for (int j = 0; j < E; ++j) {
        const float k = sW[j * total + o];
#pragma unroll
        for (int wl = 0; wl < kEmbWin; ++wl) e[wl] = fmaf(k, sS[wl * E + j], e[wl]);
      }
#pragma unroll
      for (int wl = 0; wl < kEmbWin; ++wl)
        if (wbase + wl < n) emb[(wbase + wl) * total + o] = e[wl];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Dataset-item expansion: item idx of the reference's PoseDataset is the affine transform
// idx / N of base window idx % N (utils/dataset.py:67-76, apply_pose_transform utils/dataset_utils.py:273-290:
// einsum 'ktv,ck->ctv' over (x, y, 1)).  Base windows are uploaded once; the num_transform-fold dataset is
// materialised on the device, tile by tile.  Products and sums are rounded separately, in the einsum's order.
// ------------------------------------------------------------------------------------------
constexpr int kMaxTransforms = 8;
struct TransformTable {
  float m[kMaxTransforms][6];  // rows 0 and 1 of the 3x3 matrix: x' = m0 x + m1 y + m2, y' = m3 x + m4 y + m5
};
__global__ void __launch_bounds__(kThreads) expand_transforms_kernel(const float* __restrict__ base, float* __restrict__ out,
                                                                      const __grid_constant__ TransformTable tb, int64_t N,
                                                                      int64_t first_item, int64_t n_items, int plane) {
  const int64_t total = n_items * plane;  // plane = n_frames * V positions per coordinate channel
  for (int64_t i = blockIdx.x * int64_t(kThreads) + threadIdx.x; i < total; i += int64_t(gridDim.x) * kThreads) {
    const int64_t it = i / plane;
    const int p = int(i - it * plane);
    const int64_t idx = first_item + it;
    const int64_t sample = idx % N;
    const int tr = int(idx / N);
    const float x = __ldg(base + (sample * 2) * plane + p), y = __ldg(base + (sample * 2 + 1) * plane + p);
    const float* m = tb.m[tr];
    out[(it * 2) * plane + p] = __fadd_rn(__fadd_rn(__fmul_rn(x, m[0]), __fmul_rn(y, m[1])), m[2]);
    out[(it * 2 + 1) * plane + p] = __fadd_rn(__fadd_rn(__fmul_rn(x, m[3]), __fmul_rn(y, m[4])), m[5]);
  }
}

// ------------------------------------------------------------------------------------------
// Window ingest (second slice): raw trajectory rows -> dataset items, on the device.
//
// normalize_frames_kernel: Trajectory._from_image_to_centre_bounding_box (utils/data.py:165-187) with
// compute_bounding_box (utils/data.py:11-44) for every frame row [34] = (x1, y1, ..., x17, y17).  HBM-bound (272 B per
// row): rows are staged through shared memory in contiguous tiles and one thread owns one row.  The reference computes in
// float32 with every operation rounded separately (numpy scalars, Python ints for the rounded corners), so every step below
// is an explicit _rn intrinsic -- no FMA contraction.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float clip_round(float v, float hi) {  // int(round(np.clip(v, 0, hi))): half to even
  return rintf(fminf(fmaxf(v, 0.f), hi));
}
// RobustScaler.transform on one normalised coordinate (utils/data.py:345-354): zeros are missing and stay zero; sklearn's
// `X -= center_; X /= scale_` on float32 rows = one double operation each, rounded to float32 (equal to the float32 operation
// when the fitted attribute is float32).
struct ScalerTable {
  double center[34], scale[34];
};
__device__ __forceinline__ float robust_scale(float v, double center, double scale) {
  if (v == 0.f) return 0.f;
  const float r = float(double(float(double(v) - center)) / scale);
  return r != r ? 0.f : r;   // np.where(np.isnan(X_scaled), 0.0, X_scaled)
}
constexpr int kNormRows = kThreads;  // frame rows per CTA tile: one per thread
__global__ void __launch_bounds__(kThreads) normalize_frames_kernel(const float* in, float* out, int64_t F, float vid_w, float vid_h,
                                                                    const __grid_constant__ ScalerTable sc, int apply_scaler) {
  // A tile of kNormRows rows (one contiguous 34 KB span) is staged with coalesced 8-byte copies; thread r then owns row r:
  // its 17 (x, y) pairs sit at 8-byte words 17 r + j, which is conflict-free across the 16 lanes of a 64-bit shared-memory
  // phase.  Results go back through the same tile, so `out` may alias `in` (a tile is read completely before it is written
  // and no other CTA touches it).
  __shared__ __align__(16) float2 tile[kNormRows * 17];
  const float inf = __int_as_float(0x7f800000);
  const float wl = __fsub_rn(vid_w, 1.f), hl = __fsub_rn(vid_h, 1.f);
  const int64_t n_tiles = (F + kNormRows - 1) / kNormRows;
  for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
    const int64_t f0 = tl * kNormRows;
    const int nrows = int(F - f0 < kNormRows ? F - f0 : kNormRows);
    const float2* src = reinterpret_cast<const float2*>(in + f0 * 34);
    float2* dst = reinterpret_cast<float2*>(out + f0 * 34);
    for (int k = threadIdx.x; k < 17 * nrows; k += kThreads) tile[k] = src[k];
    __syncthreads();
    if (int(threadIdx.x) < nrows) {
      float2* r = tile + 17 * threadIdx.x;
      float left = inf, right = -inf, top = inf, bottom = -inf;   // over the non-zero coordinates (data.py:27-29)
#pragma unroll
      for (int j = 0; j < 17; ++j) {
        const float2 kp = r[j];
        if (kp.x != 0.f) { left = fminf(left, kp.x); right = fmaxf(right, kp.x); }
        if (kp.y != 0.f) { top = fminf(top, kp.y); bottom = fmaxf(bottom, kp.y); }
      }
      // no non-zero x or no non-zero y: np.min raises -> box (0,0,0,0) -> zero width and height -> zeros (data.py:28-32, 182-183)
      float cx = 0.f, cy = 0.f, bw = 0.f, bh = 0.f;
      if (left <= right && top <= bottom) {
        const float ew = __fmul_rn(0.1f, __fadd_rn(__fsub_rn(right, left), 1.f));   // data.py:34
        const float eh = __fmul_rn(0.1f, __fadd_rn(__fsub_rn(bottom, top), 1.f));
        const float L = clip_round(__fsub_rn(left, ew), wl), R = clip_round(__fadd_rn(right, ew), wl);   // data.py:35-41
        const float T = clip_round(__fsub_rn(top, eh), hl), B = clip_round(__fadd_rn(bottom, eh), hl);
        cx = __fmul_rn(__fadd_rn(L, R), 0.5f), cy = __fmul_rn(__fadd_rn(T, B), 0.5f);               // exact: small integers
        bw = __fsub_rn(R, L), bh = __fsub_rn(B, T);
      }
#pragma unroll 1
      for (int j = 0; j < 17; ++j) {
        const float2 kp = r[j];
        float ox = 0.f, oy = 0.f;
        if (bw != 0.f) ox = __fdiv_rn(__fsub_rn(kp.x != 0.f ? kp.x : cx, cx), bw);   // data.py:178-182
        if (bh != 0.f) oy = __fdiv_rn(__fsub_rn(kp.y != 0.f ? kp.y : cy, cy), bh);
        if (apply_scaler) {   // the scaler acts per column, so it commutes with the windowing: scale a row once, not once per item
          ox = robust_scale(ox, sc.center[2 * j], sc.scale[2 * j]);
          oy = robust_scale(oy, sc.center[2 * j + 1], sc.scale[2 * j + 1]);
        }
        r[j] = make_float2(ox, oy);
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 17 * nrows; k += kThreads) dst[k] = tile[k];
    __syncthreads();
  }
}

// build_items_kernel: dataset item idx = first_item + i is transform idx / N of window idx % N (utils/dataset.py:67-76), the
// window being rows start, start + step, ... of the normalised frame array (utils/preprocessing.py:55-86), robust-scaled
// here (apply_scaler) or already in normalize_frames_kernel, laid out [2, L, 17] (utils/dataset.py:241-256) and transformed
// like expand_transforms_kernel.  No window tensor is ever materialised: a frame row is read from L2 by the up to
// L x num_transform items that share it.  A CTA builds kItemsPerChunk consecutive items (one contiguous output span,
// coalesced stores); the 64-bit item -> (transform, window) split happens once per chunk, the rest is 32-bit arithmetic.
constexpr int kItemsPerChunk = 8;
__global__ void __launch_bounds__(kThreads) build_items_kernel(const float* __restrict__ rows, const int64_t* __restrict__ win_start,
                                                               const __grid_constant__ ScalerTable sc, int apply_scaler,
                                                               const __grid_constant__ TransformTable tb, int64_t N, int64_t first_item,
                                                               int64_t n_items, int seg_len, int row_step, float* __restrict__ out) {
  const int plane = seg_len * 17;
  const int64_t n_chunks = (n_items + kItemsPerChunk - 1) / kItemsPerChunk;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t it0 = ch * kItemsPerChunk;
    const int cnt = int(n_items - it0 < kItemsPerChunk ? n_items - it0 : kItemsPerChunk);
    const int64_t idx0 = first_item + it0;
    const int tr0 = int(idx0 / N);
    const int64_t w0 = idx0 - tr0 * N;
    float* dst = out + it0 * 2 * plane;
    // (item in the chunk, position in the item) advance incrementally: no per-element division by the runtime `plane`
    int il = 0, p = threadIdx.x;
    while (p >= plane) { p -= plane; ++il; }
    for (; il < cnt; p += kThreads) {
      while (p >= plane) { p -= plane; ++il; }
      if (il >= cnt) break;
      const int t = p / 17, v = p - t * 17;
      int64_t w = w0 + il;
      int tr = tr0;
      while (w >= N) { w -= N; ++tr; }
      const int64_t row = __ldg(win_start + w) + int64_t(t) * row_step;
      const float2 kp = __ldg(reinterpret_cast<const float2*>(rows + row * 34 + 2 * v));
      float x = kp.x, y = kp.y;
      if (apply_scaler) {
        x = robust_scale(x, sc.center[2 * v], sc.scale[2 * v]);
        y = robust_scale(y, sc.center[2 * v + 1], sc.scale[2 * v + 1]);
      }
      const float* m = tb.m[tr];
      dst[(il * 2) * plane + p] = __fadd_rn(__fadd_rn(__fmul_rn(x, m[0]), __fmul_rn(y, m[1])), m[2]);
      dst[(il * 2 + 1) * plane + p] = __fadd_rn(__fadd_rn(__fmul_rn(x, m[3]), __fmul_rn(y, m[4])), m[5]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Bottleneck linear of the conditioning encoder: emb[n][l] = b[l] + sum_k h[n][k] * Wb[k][l]
// h is planar-4 [n][C/4][P][4]; Wb was re-indexed at pack time to that order.  One warp / window.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) bottleneck_kernel(const float* __restrict__ h, const float* __restrict__ Wb,
                                                              const float* __restrict__ bb, float* __restrict__ out,
                                                              int64_t n, int K, int L) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * int64_t(kThreads) + threadIdx.x) >> 5;
  if (w >= n) return;
  float acc[kMaxE];
#pragma unroll
  for (int l = 0; l < kMaxE; ++l) acc[l] = 0.f;
  const float* hp = h + w * K;
  for (int k = lane; k < K; k += 32) {
    const float x = hp[k];
    const float* wp = Wb + int64_t(k) * L;
#pragma unroll
    for (int l = 0; l < kMaxE; ++l)
      if (l < L) acc[l] = fmaf(x, wp[l], acc[l]);
  }
#pragma unroll
  for (int l = 0; l < kMaxE; ++l) {
    if (l < L) {
      float v = acc[l];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) out[w * L + l] = v + bb[l];
    }
  }
}

// ------------------------------------------------------------------------------------------
// DDPM update (mocodad.py:178), elementwise, in place.  The arithmetic keeps the reference's
// operation order with explicit round-to-nearest ops (no FMA contraction) so that with injected
// noise the update is bit-faithful to eager PyTorch.
//   noise: pre-drawn tensor addressed as noise[(g*slots + slot)*B + b][2P] for virtual window
//   w = g*B + b (pass B = n, g = 0 for a plain [n,2P] tensor), or nullptr => Philox.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ddpm_step_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                                             int64_t n, int per_window, DdpmArgs a) {
  const int64_t total = n * per_window;
  for (int64_t i = blockIdx.x * int64_t(kThreads) + threadIdx.x; i < total; i += int64_t(gridDim.x) * kThreads) {
    const int64_t w = i / per_window;
    const int e = int(i - w * per_window);
    float z = 0.f;
    if (a.add_noise) {
      const int64_t virt = a.virt0 + w;
      if (a.noise != nullptr) {
        const int64_t g = virt / a.noise_B, b = virt - g * a.noise_B;
        z = a.noise[((g * a.noise_slots + a.slot) * a.noise_B + b) * per_window + e];
      } else {
        const int64_t g = virt / a.noise_B, b = virt - g * a.noise_B;
        z = philox_normal(a.seed, uint64_t(a.first_window + b), uint32_t(g), uint32_t(a.slot), uint32_t(e));
      }
    }
    const float t1 = __fmul_rn(a.c2, eps[i]);
    const float t2 = __fsub_rn(x[i], t1);
    const float t3 = __fmul_rn(a.c1, t2);
    const float t4 = __fmul_rn(a.c3, z);
    x[i] = __fadd_rn(t3, t4);
  }
}

// x_T: from the pre-drawn tensor (slot 0) or Philox (slot 0).  mocodad.py:162
__global__ void __launch_bounds__(kThreads) randn_kernel(float* __restrict__ x, int64_t n, int per_window, DdpmArgs a) {
  const int64_t total = n * per_window;
  for (int64_t i = blockIdx.x * int64_t(kThreads) + threadIdx.x; i < total; i += int64_t(gridDim.x) * kThreads) {
    const int64_t w = i / per_window;
    const int e = int(i - w * per_window);
    const int64_t virt = a.virt0 + w;
    float z;
    if (a.noise != nullptr) {
      const int64_t g = virt / a.noise_B, b = virt - g * a.noise_B;
      z = a.noise[((g * a.noise_slots + a.slot) * a.noise_B + b) * per_window + e];
    } else {
      const int64_t g = virt / a.noise_B, b = virt - g * a.noise_B;
      z = philox_normal(a.seed, uint64_t(a.first_window + b), uint32_t(g), uint32_t(a.slot), uint32_t(e));
    }
    x[i] = z;
  }
}

// ------------------------------------------------------------------------------------------
// Per-window loss: mean over (C,T,V) of the elementwise loss between the generated sample and
// the corrupt frames of the input window (mocodad.py:484).  One warp per virtual window.
//   x0 [nv][2][P];  data [B][2][n_frames][V], target frames start at t0;  virtual w -> b = (virt0+w) % B
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) window_loss_kernel(const float* __restrict__ x0, const float* __restrict__ data,
                                                               float* __restrict__ losses, int64_t nv, int64_t virt0,
                                                               int64_t B, int P, int V, int n_frames, int t0, int loss_fn) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * int64_t(kThreads) + threadIdx.x) >> 5;
  if (w >= nv) return;
  const int64_t b = (virt0 + w) % B;
  const float* xp = x0 + w * 2 * P;
  const float* dp = data + b * 2 * int64_t(n_frames) * V + int64_t(t0) * V;
  float s = 0.f;
  for (int e = lane; e < 2 * P; e += 32) {
    const int c = e / P, p = e - c * P;
    const float d = xp[e] - dp[int64_t(c) * n_frames * V + p];
    const float ad = fabsf(d);
    float l;
    if (loss_fn == 0) l = ad < 1.0f ? 0.5f * d * d : ad - 0.5f;  // SmoothL1, beta = 1
    else if (loss_fn == 1) l = ad;
    else l = d * d;
    s += l;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) losses[w] = s / float(2 * P);
}

// 'best' / 'worst' over the G samples (mocodad.py:504-512): strict compare against 1e10 / -1.
__global__ void __launch_bounds__(kThreads) best_worst_kernel(const float* __restrict__ losses, float* __restrict__ best,
                                                              float* __restrict__ worst, int64_t B, int G) {
  const int64_t b = blockIdx.x * int64_t(kThreads) + threadIdx.x;
  if (b >= B) return;
  float lo = 1e10f, hi = -1.0f;
  for (int g = 0; g < G; ++g) {
    const float l = losses[int64_t(g) * B + b];
    if (l < lo) lo = l;
    if (l > hi) hi = l;
  }
  if (best) best[b] = lo;
  if (worst) worst[b] = hi;
}

// planar-4 [n][C/4][P][4] -> reference layout [n][C][P]  (parity taps only)
__global__ void __launch_bounds__(kThreads) cl_to_cf_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            int64_t n, int P, int C) {
  const int64_t total = n * P * C;
  for (int64_t i = blockIdx.x * int64_t(kThreads) + threadIdx.x; i < total; i += int64_t(gridDim.x) * kThreads) {
    const int64_t w = i / (int64_t(P) * C);
    const int rem = int(i - w * P * C);
    const int c = rem / P, p = rem - c * P;
    out[i] = in[act_off(w, c >> 2, p, C, P) + (c & 3)];
  }
}

}  // namespace mcd
