// ST_GCNN_layer.forward (models/gcae/stsgcn.py:94-116) for the blocks that NARROW the channel count (Cin > Cout: the
// denoiser's 128->64 block at 10 joints and its 64->32 block at 12 joints, models/stsae/stsae_unet.py:365-403), sm_100a.
//
// The position mixes (gcn.T over frames, gcn.A over joints; stsgcn.py:143-156) act per channel, the 1x1 convolution acts
// per position: both are linear and they commute.  mcd_block_tc.cuh mixes first (Cin channels), then contracts channels.
// Here the order is swapped: the convolution runs FIRST, straight from the landed input, and the mixes run on the COUT
// side -- half the mix work (FMA issue slots and shared-memory wavefronts, the two resources the mix-first kernel is short
// of, profiles/r02_ncu_summary.txt) for these blocks:
//
//   out = PReLU( mix(W' X) + Wr' X + b ) + emb          (W', Wr': BatchNorm-folded tcn / residual 1x1 convolutions)
//
// One persistent CTA per SM, 16 warps, hand-offs through shared-memory mbarriers (same conventions as mcd_block_tc.cuh):
//   warp 13    loader: TMA copies of the next 16-channel input chunk into the X ring (planar-4 buffer) and one bulk copy per
//              chunk of its weight operands [W hi | Wr hi | W lo | Wr lo] (UMMA K-major SWIZZLE_64B)
//   warps 14,15 tf32 lo part of the landed X chunk (element-wise), even / odd chunks: with a single conversion warp and a single
//              Xlo buffer the 128->64 block ran conversion -> MMAs -> conversion serially and was slower than mix-first
//   warp 12    MMA issue: per chunk and 128-row tile  D[128 x 2*COUT] += X_hi*[W_hi|Wr_hi] + X_lo*[W_hi|Wr_hi] + X_hi*[W_lo|Wr_lo]
//              (3xTF32, fp32 accumulation in TMEM; columns [0,COUT) = convolution, [COUT,2*COUT) = residual convolution;
//              the planar X buffer is the A operand as it stands -- no-swizzle K-major descriptor)
//   warps 8-11 "drain": convolution columns TMEM -> shared memory as 16-channel planar chunks (the Z ring, two slots)
//              ... and, two chunks later, "final": mixed chunk + residual columns (still in TMEM) + bias,
//              PReLU, + embedding -> planar-4 global stores; with VUP > 0 (production calls) the finished chunk goes back to
//              shared memory instead and the same warps apply the up-path CNN_layer that follows the block (joint resample
//              V -> VUP) and accumulate the result onto the U-Net skip tensor in place (red.global.add.v4.f32) -- see upsample()
//   warps 0-3  T-mix  Y1[q,v,c] = sum_t Z[t,v,c] * Tm[v][t][q]      (register-resident weight slices, packed FFMA2)
//   warps 4-7  A-mix  Y2[t,w,c] = sum_v Y1[t,v,c] * A[t][v][w]      -> planar fp32 chunk for the final stage
// The convolution of tile i+1 (tensor pipe) overlaps the mixes of tile i (FMA pipe); TMEM holds two accumulator sets.
#pragma once
#include <type_traits>
#include "mcd_block_tc.cuh"

namespace mcd {

struct CfNoUp { int unused; };

// VUP_ > 0: the block also applies the CNN_layer that follows it on the up path (joint resample V -> VUP with folded
// BatchNorm, stsgcn.py:187-199 at stsae_unet.py:381-394) and ADDS the VUP-joint result to the U-Net skip tensor in place
template <int T_, int V_, int CIN_, int COUT_, int NW_, int VUP_ = 0>
struct CfCfg {
  using Mix = TcCfg<T_, V_, COUT_, COUT_, NW_>;  // task shapes of the two mixes (they depend on T, V and the window count only)
  static constexpr int T = T_, V = V_, CIN = CIN_, COUT = COUT_, NW = NW_, VUP = VUP_;
  using Up = std::conditional_t<(VUP_ > 0), ResampleParams<V_, (VUP_ > 0 ? VUP_ : 1)>, CfNoUp>;
  // fused resample: an epilogue thread owns UP_WG output joints (g, g + UP_NG, ...) of the frames fg, fg + UP_FG, ... of the tile
  // (12 output joints: 4 groups of 3 x 24 frames = 96 threads, one round; 17: 5 groups of 4 x 24 = 120 threads, one round)
  static constexpr int UP_WG = VUP_ > 12 ? 4 : 3;
  static constexpr int UP_NG = (VUP_ + UP_WG - 1) / UP_WG;
  static constexpr int FRAMES = NW_ * T_;
  static constexpr int UP_FG = VUP_ > 0 ? ((kTcEpilogue / (UP_NG > 0 ? UP_NG : 1)) < FRAMES ? (kTcEpilogue / (UP_NG > 0 ? UP_NG : 1)) : FRAMES) : 1;
  static constexpr int UP_ROUNDS = (FRAMES + UP_FG - 1) / UP_FG;
  static constexpr int P = T * V, ROWS = NW * P, MT = (ROWS + 127) / 128;
  static constexpr int KC = 16, C4 = 4;
  static constexpr int NCHUNK = CIN / KC;   // input chunks (convolution K steps)
  static constexpr int NCH2 = COUT / KC;    // output chunks (mix passes)
  static_assert(CIN > COUT && CIN % KC == 0 && COUT % KC == 0 && NCH2 >= 1 && NCH2 <= 4, "conv-first block: Cin > Cout, multiples of 16");
  static constexpr bool TMA_TILED = NW > 1;
  static constexpr int TCOLS = 2 * COUT;           // accumulator columns per 128-row tile: convolution | residual convolution
  static constexpr int ACC_COLS = MT * TCOLS;      // one accumulator set; TMEM holds two (tile parity)
  static_assert(TCOLS <= 256 && TCOLS % 16 == 0 && 2 * ACC_COLS <= 512, "two accumulator sets must fit TMEM");
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512);
  static_assert(ROWS % 8 == 0, "operand arrays must be whole core-matrix rows");
  static constexpr int ARR = ROWS * 16;            // one 16-channel planar chunk [c4][row][4 floats]
  static constexpr int Y1ARR = Mix::Y1ARR;
  static constexpr int WCH = 4 * COUT * 16;        // one chunk of weight operands: [W hi | Wr hi | W lo | Wr lo] x [COUT][16]
  static constexpr int UPW_FLOATS = VUP_ > 0 ? (VUP_ * (V_ + 1) + 3) / 4 * 4 : 0;  // fused resample: W[w][0..V) and b[w] per output joint
  static constexpr int SM_MISC = 2 * Y1ARR + 2 * WCH + COUT + NW * COUT + UPW_FLOATS;
  static constexpr bool fits(int arrays) { return size_t(SM_MISC + arrays * ARR) * sizeof(float) + 1024 <= 227 * 1024; }
  // arrays: X ring (NXB), Xlo (2: one per conversion warp), Z ring (NZ = 2), Y2 (2).  The X ring takes what is left, up to 4
  // slots: a slot's round trip is release (MMA commit) -> TMA issue -> landed -> conversion -> MMAs, ~4 k cycles, so a ring of 2
  // caps the convolution phase at ~2 k cycles per chunk (measured: the 128->64 block was no faster than mix-first with it)
  static constexpr int NXLO = 2, NZ = 2;
  static_assert(NCH2 % NZ == 0, "output chunks per tile must be a multiple of the Z ring depth");
  static constexpr int NXB = fits(4 + NXLO + NZ + 2) ? 4 : (fits(3 + NXLO + NZ + 2) ? 3 : 2);
  static_assert(fits(NXB + NXLO + NZ + 2), "conv-first block tile exceeds shared memory");
  static constexpr int SM_X = 0;
  static constexpr int SM_XLO = SM_X + NXB * ARR;
  static constexpr int SM_Z = SM_XLO + NXLO * ARR;
  static constexpr int SM_Y2 = SM_Z + NZ * ARR;
  static constexpr int SM_Y1 = SM_Y2 + 2 * ARR;
  static constexpr int SM_WC = SM_Y1 + 2 * Y1ARR;
  static constexpr int SM_BIAS = SM_WC + 2 * WCH;
  static constexpr int SM_EMB = SM_BIAS + COUT;
  static constexpr int SM_UP = SM_EMB + NW * COUT;
  static constexpr int SM_TOTAL = SM_UP + UPW_FLOATS;
  // the last 128-row MMA tile over-reads (MT*128 - ROWS) rows past each plane of X / Xlo: the arrays that follow absorb it
  static_assert((MT * 128 - ROWS) * 4 <= NZ * ARR, "over-read of the last MMA tile must stay inside the allocation");
  static constexpr size_t SMEM_BYTES = size_t(SM_TOTAL) * sizeof(float) + 1024;
};

enum CfBar {
  CF_X_FULL = 0,     // [4] loader -> conversion warp (-> MMA warp through CF_XLO_FULL)
  CF_X_EMPTY = 4,    // [4] tcgen05.commit -> loader
  CF_W_FULL = 8,     // [2] weight loader -> MMA warp
  CF_W_EMPTY = 10,   // [2] tcgen05.commit -> weight loader
  CF_XLO_FULL = 12,  // [2] conversion warp -> MMA warp
  CF_XLO_EMPTY = 14, // [2] tcgen05.commit -> conversion warp
  CF_ACC_FULL = 16,  // [2] tcgen05.commit -> drain
  CF_ACC_EMPTY = 18, // [2] final stage -> MMA warp
  CF_Z_FULL = 20,    // [2] drain -> T-mix (slot = output-chunk iteration % NZ); no "Z empty" barrier: see the drain
  CF_Y1_FULL = 22,   // [2] T-mix -> A-mix
  CF_Y1_EMPTY = 24,  // [2] A-mix -> T-mix
  CF_Y2_FULL = 26,   // [2] A-mix -> final stage
  CF_Y2_EMPTY = 28,  // [2] final stage -> A-mix
  CF_BAR_COUNT = 30
};

template <class Cfg>
__global__ void __launch_bounds__(kTcThreads, 1) stgcn_block_cf_kernel(const BlockWeights wt, const BlockIO io,
                                                                       const __grid_constant__ CUtensorMap tmx,
                                                                       const __grid_constant__ typename Cfg::Up up) {
  using Mix = typename Cfg::Mix;
  constexpr int T = Cfg::T, V = Cfg::V, P = Cfg::P, ROWS = Cfg::ROWS, C4 = Cfg::C4, MT = Cfg::MT;
  constexpr int CIN = Cfg::CIN, COUT = Cfg::COUT, NCHUNK = Cfg::NCHUNK, NCH2 = Cfg::NCH2, NW = Cfg::NW, NXB = Cfg::NXB;
  constexpr int VP = Mix::VP, TP4 = Mix::TP4, TMS = Mix::TMS, ARR = Cfg::ARR, WCH = Cfg::WCH;
  constexpr int QG = Mix::QG, NQG = Mix::NQG, WGS = Mix::WGS, NWG = Mix::NWG, TTP = Mix::TTP, Y1ARR = Cfg::Y1ARR;
  constexpr int VUP = Cfg::VUP;

  extern __shared__ uint8_t smem_raw[];
  float* smem = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  float* sX = smem + Cfg::SM_X;
  float* sXlo = smem + Cfg::SM_XLO;
  float* sZ = smem + Cfg::SM_Z;
  float* sY2 = smem + Cfg::SM_Y2;
  float* sY1 = smem + Cfg::SM_Y1;
  float* sWc = smem + Cfg::SM_WC;
  float* sBias = smem + Cfg::SM_BIAS;
  float* sEmb = smem + Cfg::SM_EMB;
  __shared__ __align__(8) uint64_t bars[CF_BAR_COUNT];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (io.n + NW - 1) / NW;
  if (int64_t(blockIdx.x) >= ntiles) return;  // uniform over the CTA
  const int my_tiles = int((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const int npairs1 = my_tiles * NCHUNK;  // (tile, input chunk) pairs: loaders, conversion, MMA
  const int npairs2 = my_tiles * NCH2;    // (tile, output chunk) pairs: mixes, final stage
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int slot) { return bar0 + uint32_t(slot) * 8u; };

  // ---- once per CTA ----
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(uint32_t(Cfg::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(BAR(CF_X_FULL + i), 1);   // one expect_tx arrival + the copies' bytes
      mbar_init(BAR(CF_X_EMPTY + i), 1);  // the commit of the chunk's MMAs
    }
    static_assert(Cfg::NZ == 2, "two Z ring barriers");
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(CF_Z_FULL + i), kTcEpilogue);
      mbar_init(BAR(CF_W_FULL + i), 1);
      mbar_init(BAR(CF_W_EMPTY + i), 1);
      mbar_init(BAR(CF_XLO_FULL + i), 32);
      mbar_init(BAR(CF_XLO_EMPTY + i), 1);
      mbar_init(BAR(CF_ACC_FULL + i), 1);
      mbar_init(BAR(CF_ACC_EMPTY + i), kTcEpilogue);
      mbar_init(BAR(CF_Y1_FULL + i), kTcMix);
      mbar_init(BAR(CF_Y1_EMPTY + i), kTcMix);
      mbar_init(BAR(CF_Y2_FULL + i), kTcMix);
      mbar_init(BAR(CF_Y2_EMPTY + i), kTcEpilogue);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp >= kTcEpiWarp0 && warp < kTcLoadWarp) {
    for (int i = tid - kTcEpiWarp0 * 32; i < COUT; i += kTcEpilogue) sBias[i] = wt.bias[i];
    if constexpr (VUP > 0) {
      // the fused CNN_layer's weights: the threads of a warp own different output joints, and per-lane indices into the
      // parameter (constant) bank serialise -- shared memory serves them in one wavefront
      float* sUp = smem + Cfg::SM_UP;
      for (int i = tid - kTcEpiWarp0 * 32; i < VUP * (V + 1); i += kTcEpilogue) {
        const int w = i / (V + 1), v = i - w * (V + 1);
        sUp[i] = v < V ? up.w[w][v] : up.b[w];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // =============================== T-mix warps ===============================
    reg_inc<kRegsT>();
    // Y1[n,(q,v),c] = sum_t Z[n,(t,v),c] * Tm[v][t][q]        stsgcn.py:154 (applied after the convolution)
    const int ws = tid / Mix::TT, rem_all = tid - ws * Mix::TT;
    const int cs = rem_all / Mix::TTR, rem = rem_all - cs * Mix::TTR;  // channel split, (joint, frame group)
    const int v = rem / NQG, qg = rem - v * NQG;
    const bool active = ws < Mix::WS_T;
    float wT[T][QG];  // this thread's slice of the learned time-mix matrix, resident for the whole launch
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int q = 0; q < QG; ++q)
        wT[t][q] = (active && qg * QG + q < T) ? __ldg(wt.Tm + v * TMS + t * TP4 + qg * QG + q) : 0.f;

    for (int it = 0; it < npairs2; ++it) {
      const int s = it & 1, zs = it % Cfg::NZ;
      mbar_wait(BAR(CF_Z_FULL + zs), uint32_t((it / Cfg::NZ) & 1));
      if (it >= 2) mbar_wait(BAR(CF_Y1_EMPTY + s), uint32_t((it / 2 - 1) & 1));
      if (active) {
        const float* sXc = sZ + zs * ARR;
        float* sY = sY1 + s * Y1ARR;
        for (int wl = ws; wl < NW; wl += Mix::WS_T) {
#pragma unroll 1
          for (int c4 = cs; c4 < C4; c4 += Mix::CS_T) {  // one 4-channel group per pass (see mcd_block_tc.cuh)
            float2 a[2][QG];
#pragma unroll
            for (int q = 0; q < QG; ++q) a[0][q] = a[1][q] = make_float2(0.f, 0.f);
            constexpr int TB = QG > 4 ? (T % 3 == 0 ? 3 : 2) : (T % 6 == 0 ? 6 : (T % 3 == 0 ? 3 : (T % 2 == 0 ? 2 : 1)));
            const float* xp = sXc + ((c4 * NW + wl) * P + v) * 4;  // planar chunk: [c4][window][position] 16-byte elements
            float4 xc[TB], xn[TB];
#pragma unroll
            for (int i = 0; i < TB; ++i) xn[i] = lds4_early(xp + i * V * 4);
#pragma unroll
            for (int tb = 0; tb < T; tb += TB) {
#pragma unroll
              for (int i = 0; i < TB; ++i) xc[i] = xn[i];
              if (tb + TB < T) {
#pragma unroll
                for (int i = 0; i < TB; ++i) xn[i] = lds4_early(xp + (tb + TB + i) * V * 4);
              }
#pragma unroll
              for (int i = 0; i < TB; ++i) {
                const int t = tb + i;
                const float2 xl = make_float2(xc[i].x, xc[i].y), xh = make_float2(xc[i].z, xc[i].w);
#pragma unroll
                for (int q = 0; q < QG; ++q) {
                  const float2 ww = make_float2(wT[t][q], wT[t][q]);
                  a[0][q] = ffma2(xl, ww, a[0][q]);
                  a[1][q] = ffma2(xh, ww, a[1][q]);
                }
              }
            }
            float* yp = sY + (((wl * 4 + c4) * QG) * TTP + rem) * 4;
#pragma unroll
            for (int q = 0; q < QG; ++q)
              if (T % QG == 0 || qg * QG + q < T)
                *reinterpret_cast<float4*>(yp + q * TTP * 4) = make_float4(a[0][q].x, a[0][q].y, a[1][q].x, a[1][q].y);
          }
        }
      }
      mbar_arrive(BAR(CF_Y1_FULL + s));  // (also releases Z slot zs: the drain refills it only after the final stage of `it`)
    }
  } else if (warp < 8) {
    // =============================== A-mix warps ===============================
    reg_inc<kRegsA>();
    // Y2[n,(t,w),c] = sum_v Y1[n,(t,v),c] * A[t][v][w]       stsgcn.py:155   -> planar fp32 chunk for the final stage
    const int atid = tid - kTcMix;
    const int ws = atid / Mix::TA, rem_all = atid - ws * Mix::TA;
    const int cs = rem_all / Mix::TAR, rem = rem_all - cs * Mix::TAR;  // channel split, (frame, joint group)
    constexpr int FS = (V % 8 == 1 && T % 5 != 0) ? 5 : 1;
    const int tk = rem / NWG, wg = rem - tk * NWG;
    const int t = (tk * FS) % T;
    const bool active = ws < Mix::WS_A;
    float wA[V][WGS];  // this thread's slice of the learned joint-mix matrix
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int j = 0; j < WGS; ++j)
        wA[v][j] = (active && wg + j * NWG < V) ? __ldg(wt.A + (t * V + v) * VP + wg + j * NWG) : 0.f;

    for (int it = 0; it < npairs2; ++it) {
      const int s = it & 1;
      mbar_wait(BAR(CF_Y1_FULL + s), uint32_t((it / 2) & 1));
      if (it >= 2) mbar_wait(BAR(CF_Y2_EMPTY + s), uint32_t((it / 2 - 1) & 1));
      if (active) {
        const float* sY = sY1 + s * Y1ARR;
        float* sO = sY2 + s * ARR;
        for (int wl = ws; wl < NW; wl += Mix::WS_A) {
          const int r0 = wl * P + t * V;
#pragma unroll 1
          for (int c4 = cs; c4 < C4; c4 += Mix::CS_A) {
            float2 a[2][WGS];
#pragma unroll
            for (int j = 0; j < WGS; ++j) a[0][j] = a[1][j] = make_float2(0.f, 0.f);
            const float* yp = sY + (((wl * 4 + c4) * QG + t % QG) * TTP + t / QG) * 4;  // element (v, t) at + v * NQG
            constexpr int VB = V * WGS > 60 ? 4 : (V >= 12 ? 6 : (V >= 6 ? 5 : V));
            float4 yc[VB], yn[VB];
#pragma unroll
            for (int i = 0; i < VB; ++i)
              if (i < V) yn[i] = lds4_early(yp + i * NQG * 4);
#pragma unroll
            for (int vb = 0; vb < V; vb += VB) {
#pragma unroll
              for (int i = 0; i < VB; ++i) yc[i] = yn[i];
#pragma unroll
              for (int i = 0; i < VB; ++i)
                if (vb + VB + i < V) yn[i] = lds4_early(yp + (vb + VB + i) * NQG * 4);
#pragma unroll
              for (int i = 0; i < VB; ++i) {
                const int v = vb + i;
                if (v < V) {
                  const float2 yl = make_float2(yc[i].x, yc[i].y), yh = make_float2(yc[i].z, yc[i].w);
#pragma unroll
                  for (int j = 0; j < WGS; ++j) {
                    const float2 ww = make_float2(wA[v][j], wA[v][j]);
                    a[0][j] = ffma2(yl, ww, a[0][j]);
                    a[1][j] = ffma2(yh, ww, a[1][j]);
                  }
                }
              }
            }
            // planar chunk [c4][row][4 floats]: the NWG threads of a frame store consecutive rows (16-byte elements)
#pragma unroll
            for (int j = 0; j < WGS; ++j) {
              const int w = wg + j * NWG;
              const bool okw = (V % WGS == 0 && NWG * WGS == V) || w < V;
              sts4_pred(sO + (c4 * ROWS + r0 + (okw ? w : 0)) * 4, make_float4(a[0][j].x, a[0][j].y, a[1][j].x, a[1][j].y), okw);
            }
          }
        }
      }
      mbar_arrive(BAR(CF_Y1_EMPTY + s));
      mbar_arrive(BAR(CF_Y2_FULL + s));
    }
  } else if (warp >= kTcMmaWarp) {
    reg_dec<kRegsS>();
    if (warp == kTcMmaWarp) {
      // =============================== MMA-issuing warp ===============================
      const uint32_t idesc2 = umma_idesc_tf32(2 * COUT);
      const uint64_t dW_0 = umma_desc_sw64(smem_u32(sWc)), dW_1 = umma_desc_sw64(smem_u32(sWc + WCH));
      const uint64_t dX_0 = umma_desc_planar(smem_u32(sX), ROWS * 16), dXlo_0 = umma_desc_planar(smem_u32(sXlo), ROWS * 16);
      constexpr uint64_t ARR16 = (uint64_t(ARR) * 4) >> 4;  // one chunk array, in descriptor address units (16 bytes)
      constexpr uint64_t PART = (COUT * 64) >> 4;           // one weight part, in descriptor address units
      for (int it = 0; it < npairs1; ++it) {
        const int ti = it / NCHUNK, cj = it - ti * NCHUNK;
        const int set = ti & 1, s = it & 1, b = it % NXB;
        const uint64_t bW = s ? dW_1 : dW_0;
        const uint32_t d0 = tmem + set * Cfg::ACC_COLS;
        mbar_wait(BAR(CF_W_FULL + s), uint32_t((it / 2) & 1));
        if (cj == 0 && ti >= 2) mbar_wait(BAR(CF_ACC_EMPTY + set), uint32_t((ti / 2 - 1) & 1));  // set released by the final stage
        mbar_wait(BAR(CF_XLO_FULL + s), uint32_t((it / 2) & 1));
        // the conversion warp saw the X chunk land; the issuing thread observes the copies' completion itself as well (the wait
        // is genuine: the MMAs of iteration it - 1 are issued, so every earlier phase of the slot's barrier has completed)
        mbar_wait(BAR(CF_X_FULL + b), uint32_t((it / NXB) & 1));
        tc_fence_after();
        const uint64_t xHi = dX_0 + uint64_t(b) * ARR16, xLo = dXlo_0 + uint64_t(it % Cfg::NXLO) * ARR16;
        if (elect_one()) {
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const uint32_t d = d0 + m * Cfg::TCOLS;
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // two K=8 steps per 16-channel chunk: c4 planes (2h, 2h+1)
              const uint64_t ao = uint64_t((m * 128 * 16 + h * 2 * ROWS * 16) >> 4), bo = uint64_t((h * 32) >> 4);
              const uint32_t first = (cj > 0 || h > 0) ? 1u : 0u;  // the very first MMA of a tile clears its accumulator columns
              umma_tf32(d, xHi + ao, bW + bo, idesc2, first);            // X_hi * [W_hi | Wr_hi]
              umma_tf32(d, xLo + ao, bW + bo, idesc2, 1u);               // X_lo * [W_hi | Wr_hi]
              umma_tf32(d, xHi + ao, bW + 2 * PART + bo, idesc2, 1u);    // X_hi * [W_lo | Wr_lo]
            }
          }
          umma_commit(BAR(CF_XLO_EMPTY + s));
          umma_commit(BAR(CF_X_EMPTY + b));
          umma_commit(BAR(CF_W_EMPTY + s));
          if (cj == NCHUNK - 1) umma_commit(BAR(CF_ACC_FULL + set));  // the tile's accumulators are complete
        }
        __syncwarp();
      }
    } else if (warp == kTcLoadWarp) {
      // =============================== activation loader warp ===============================
      for (int it = 0; it < npairs1; ++it) {
        const int ti = it / NCHUNK, chunk = it - ti * NCHUNK;
        const int64_t tile = blockIdx.x + int64_t(ti) * gridDim.x;
        const int b = it % NXB;
        if (it >= NXB) mbar_wait(BAR(CF_X_EMPTY + b), uint32_t((it / NXB - 1) & 1));
        constexpr uint32_t PLANE = P * 16;
        const uint32_t dst0 = smem_u32(sX + b * ARR);
        if constexpr (Cfg::TMA_TILED) {
          if (lane == 0) mbar_expect_tx(BAR(CF_X_FULL + b), uint32_t(NW) * 4u * PLANE);
          __syncwarp();
          if (lane < 4)
            tma_load_4d(dst0 + uint32_t(lane * NW) * PLANE, &tmx, 0, 0, chunk * C4 + lane, int(tile * NW), BAR(CF_X_FULL + b));
        } else {
          int64_t nvalid = io.n - tile * NW;
          if (nvalid > NW) nvalid = NW;
          if (lane == 0) mbar_expect_tx(BAR(CF_X_FULL + b), uint32_t(nvalid) * 4u * PLANE);
          __syncwarp();
          for (int k = lane; k < int(nvalid) * 4; k += 32) {
            const int wl = k >> 2, j = k & 3;
            bulk_g2s(dst0 + uint32_t(j * NW + wl) * PLANE, io.in + act_off(tile * NW + wl, chunk * C4 + j, 0, CIN, P), PLANE, BAR(CF_X_FULL + b));
          }
        }
        // the chunk's weight operands (the same commit frees the weight buffer and the X slot of iteration it - 2)
        const int s = it & 1;
        if (it >= 2) mbar_wait(BAR(CF_W_EMPTY + s), uint32_t((it / 2 - 1) & 1));
        if (lane == 0) {
          mbar_expect_tx(BAR(CF_W_FULL + s), uint32_t(WCH * 4));
          bulk_g2s(smem_u32(sWc + s * WCH), wt.Bop + size_t(chunk) * WCH, uint32_t(WCH * 4), BAR(CF_W_FULL + s));
        }
        __syncwarp();
      }
    } else {
      // =============================== conversion warps (14: even chunks, 15: odd chunks) ===============================
      static_assert(Cfg::NXLO == 2, "one Xlo buffer per conversion warp");
      const int s = warp - (kTcLoadWarp + 1);
      for (int it = s; it < npairs1; it += 2) {
        const int b = it % NXB;
        // Order matters.  A parity wait only tells "the phase with this parity is over", so it passes spuriously on a barrier
        // that is still TWO phases behind.  With an odd ring (NXB = 3) the previous phase of X_FULL[b] (iteration it - 3) was
        // awaited by the OTHER conversion warp: if that copy lands late (seen on a CTA's first tile, where three copies are
        // issued back to back: ~1 launch in 200 produced one window with a stale lo part), this warp would run through.  After the
        // XLO_EMPTY wait the MMAs of iteration it - 2 are complete, hence every copy up to it - 2 has landed and the wait below
        // is genuine.
        if (it >= 2) mbar_wait(BAR(CF_XLO_EMPTY + s), uint32_t((it / 2 - 1) & 1));  // Xlo[s] was read by the MMAs of iteration it - 2
        mbar_wait(BAR(CF_X_FULL + b), uint32_t((it / NXB) & 1));
        const float4* src = reinterpret_cast<const float4*>(sX + b * ARR);
        float4* dst = reinterpret_cast<float4*>(sXlo + s * ARR);
        constexpr int NEL = ROWS * C4, U = 4;
        int idx = lane;
        for (; idx + (U - 1) * 32 < NEL; idx += U * 32) {
          float4 x[U];
#pragma unroll
          for (int u = 0; u < U; ++u) x[u] = src[idx + u * 32];
#pragma unroll
          for (int u = 0; u < U; ++u) dst[idx + u * 32] = tf32_lo4(x[u]);
        }
        for (; idx < NEL; idx += 32) dst[idx] = tf32_lo4(src[idx]);
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor pipe
        mbar_arrive(BAR(CF_XLO_FULL + s));
      }
    }
  } else {
    // =============================== drain / final-stage warps ===============================
    reg_dec<kRegsE>();
    const float slope = wt.prelu;
    const int q = warp & 3;  // TMEM lane quarter this warp may access (warps 8-11 -> quarters 0-3)
    const int etid = tid - kTcEpiWarp0 * 32;
    constexpr int EPER = (NW * COUT + kTcEpilogue - 1) / kTcEpilogue;
    float epre[EPER];
    auto load_emb = [&](int64_t tl) {
#pragma unroll
      for (int k = 0; k < EPER; ++k) {
        const int i = etid + k * kTcEpilogue;
        const int wl = i / COUT, co = i - wl * COUT;
        const int64_t w = tl * NW + wl;
        epre[k] = (i < NW * COUT && w < io.n) ? __ldg(io.emb + emb_row(io.w0, w, io.emb_mod) * io.emb_stride + io.emb_off + co) : 0.f;
      }
    };
    // convolution columns of output-chunk iteration d: TMEM -> Z ring slot d % NZ (planar chunk, the T-mix warps' input)
    auto drain = [&](int d) {
      const int ti = d / NCH2, c = d - ti * NCH2, set = ti & 1, zs = d % Cfg::NZ;
      if (c == 0) {
        mbar_wait(BAR(CF_ACC_FULL + set), uint32_t((ti / 2) & 1));
        tc_fence_after();
      }
      // Slot zs is free without a barrier of its own: drain(d) runs after final_stage(d - NZ), which waited for Y2_FULL of that
      // iteration (A-mix done), whose A-mix waited for Y1_FULL (arrived by every T-mix thread AFTER its last read of the slot)
      float* zc = sZ + zs * ARR;
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int r = m * 128 + q * 32 + lane;
        uint32_t acc[16];
        tmem_ld16(tmem + (uint32_t(q * 32) << 16) + uint32_t(set * Cfg::ACC_COLS + m * Cfg::TCOLS + c * 16), acc);
        tmem_ld_wait16(acc);
        const bool ok = r < ROWS;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts4_pred(zc + (j * ROWS + (ok ? r : 0)) * 4,
                    make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), __uint_as_float(acc[4 * j + 2]),
                                __uint_as_float(acc[4 * j + 3])),
                    ok);
      }
      mbar_arrive(BAR(CF_Z_FULL + zs));
    };
    // mixed chunk c + residual-convolution columns + bias -> PReLU -> + embedding -> planar-4 global stores   stsgcn.py:109-114
    auto final_stage = [&](int it) {
      const int ti = it / NCH2, c = it - ti * NCH2, s = it & 1, set = ti & 1;
      const int64_t tile = blockIdx.x + int64_t(ti) * gridDim.x;
      mbar_wait(BAR(CF_Y2_FULL + s), uint32_t((it / 2) & 1));
      float* y2 = sY2 + s * ARR;
      // the chunk's bias (and, for one-window tiles, its embedding values) are the same for every row: read once per chunk
      constexpr bool EMB_HOIST = NW == 1;
      const float4* bp = reinterpret_cast<const float4*>(sBias + c * 16);
      float4 b4[4] = {bp[0], bp[1], bp[2], bp[3]}, e4[4];
      if constexpr (EMB_HOIST) {
        const float4* ep = reinterpret_cast<const float4*>(sEmb + c * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) e4[j] = ep[j];
      }
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int r = m * 128 + q * 32 + lane;
        const int wl = r / P;
        const int64_t w = tile * NW + wl;
        const bool ok = (r < ROWS) && (w < io.n);
        const int pp = r - wl * P;
        uint32_t acc[16];
        tmem_ld16(tmem + (uint32_t(q * 32) << 16) + uint32_t(set * Cfg::ACC_COLS + m * Cfg::TCOLS + COUT + c * 16), acc);
        // rows past the tile (r >= ROWS) read a dummy location nobody writes: with the fused resample the chunk is rewritten
        // in place by the rows' owners while other threads may still be loading
        float4 y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = *reinterpret_cast<const float4*>(r < ROWS ? y2 + (j * ROWS + r) * 4 : sBias + j * 4);
        if constexpr (!EMB_HOIST) {
          const float4* ep = reinterpret_cast<const float4*>(sEmb + (ok ? wl : 0) * COUT + c * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) e4[j] = ep[j];
        }
        float* op = io.out + (ok ? act_off(w, c * 4, pp, COUT, P) : 0);
        tmem_ld_wait16(acc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float o[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float x = f4get(y[j], jj) + __uint_as_float(acc[4 * j + jj]) + f4get(b4[j], jj);
            x = x > 0.f ? x : slope * x;
            o[jj] = x + f4get(e4[j], jj);
          }
          if constexpr (VUP == 0) stg4_pred(op + j * P * 4, make_float4(o[0], o[1], o[2], o[3]), ok);
          else sts4_pred(y2 + (j * ROWS + (r < ROWS ? r : 0)) * 4, make_float4(o[0], o[1], o[2], o[3]), r < ROWS);  // in place: the thread's own row
        }
      }
      if constexpr (VUP == 0) mbar_arrive(BAR(CF_Y2_EMPTY + s));  // (fused resample: released by upsample())
      if (c == NCH2 - 1) {
        tc_fence_before();  // accumulator reads ordered before the release of the set
        mbar_arrive(BAR(CF_ACC_EMPTY + set));
      }
    };

    // Fused CNN_layer + skip add (VUP > 0): the chunk's block outputs sit in Y2 slot s (written in place by final_stage);
    //   out[n, c, t, w] = skip[n, c, t, w] + (b[w] + sum_v W[w][v] * y[n, c, t, v])
    // The skip tensor is not loaded: `out` IS the skip buffer (its last use in the U-Net) and the resampled value is added
    // to it in L2 with a vector reduction (red.global.add.v4.f32: fire and forget, one per 16-byte element, so the result is
    // the same single rounding as the stand-alone kernel's a + skip and does not depend on timing).  Loading the skip values
    // into registers instead cost two thirds of this stage (by elimination: 41.8 -> 32.3 ms per step without the loads).
    // A thread owns the output joints g, g + UP_NG, ... of frames fg, fg + UP_FG, ...: consecutive threads touch consecutive
    // 16-byte elements.  Work items of a chunk: k = round * C4 + c4 (frame f = fg + round * UP_FG, 4-channel plane c4).
    const int up_fg = etid / (VUP > 0 ? Cfg::UP_NG : 1), up_g = etid - up_fg * (VUP > 0 ? Cfg::UP_NG : 1);
    auto upsample = [&](int it) {
      if constexpr (VUP > 0) {
        constexpr int WG = Cfg::UP_WG, NG = Cfg::UP_NG, K = Cfg::UP_ROUNDS * C4;
        const int ti = it / NCH2, c = it - ti * NCH2, s = it & 1;
        const int64_t tile = blockIdx.x + int64_t(ti) * gridDim.x;
        named_bar_sync(3, kTcEpilogue);  // every row of the chunk is in place
        if (up_fg < Cfg::UP_FG) {
          float rw[WG][V], rb[WG];
#pragma unroll
          for (int j = 0; j < WG; ++j) {
            const float* wp = smem + Cfg::SM_UP + (up_g + j * NG < VUP ? up_g + j * NG : 0) * (V + 1);
            rb[j] = wp[V];
#pragma unroll
            for (int v = 0; v < V; ++v) rw[j][v] = wp[v];
          }
          const float* y2 = sY2 + s * ARR;
#pragma unroll 1
          for (int k = 0; k < K; ++k) {
            const int f = up_fg + (k / C4) * Cfg::UP_FG, c4 = k % C4;
            const int wl = f / T, t = f - wl * T;
            const int64_t wdw = tile * NW + wl;
            if (f >= Cfg::FRAMES || wdw >= io.n) continue;
            float4 a[WG];
#pragma unroll
            for (int j = 0; j < WG; ++j) a[j] = make_float4(rb[j], rb[j], rb[j], rb[j]);
            const float* xp = y2 + (c4 * ROWS + f * V) * 4;
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float4 x = *reinterpret_cast<const float4*>(xp + v * 4);
#pragma unroll
              for (int j = 0; j < WG; ++j) {
                a[j].x = fmaf(rw[j][v], x.x, a[j].x); a[j].y = fmaf(rw[j][v], x.y, a[j].y);
                a[j].z = fmaf(rw[j][v], x.z, a[j].z); a[j].w = fmaf(rw[j][v], x.w, a[j].w);
              }
            }
            float* op = io.out + act_off(wdw, c * 4 + c4, t * VUP + up_g, COUT, T * VUP);
#pragma unroll
            for (int j = 0; j < WG; ++j)
              if (up_g + j * NG < VUP)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + j * NG * 4), "f"(a[j].x), "f"(a[j].y),
                             "f"(a[j].z), "f"(a[j].w)
                             : "memory");
          }
        }
        mbar_arrive(BAR(CF_Y2_EMPTY + s));
      }
    };

    load_emb(blockIdx.x);
    for (int d = 0; d < Cfg::NZ; ++d) drain(d);  // (npairs2 >= NCH2 >= NZ)
    for (int it = 0; it < npairs2; ++it) {
      if (it % NCH2 == 0) {  // a new tile: its embedding rows (loaded a tile ahead) go to shared memory
        const int64_t tile = blockIdx.x + int64_t(it / NCH2) * gridDim.x;
        named_bar_sync(2, kTcEpilogue);  // everyone is done with the previous tile's sEmb
#pragma unroll
        for (int k = 0; k < EPER; ++k)
          if (etid + k * kTcEpilogue < NW * COUT) sEmb[etid + k * kTcEpilogue] = epre[k];
        named_bar_sync(2, kTcEpilogue);
        if (it + NCH2 < npairs2) load_emb(tile + gridDim.x);
      }
      final_stage(it);
      // refill the Z slot this chunk's mixes released (the A-mix of `it` is done, hence its T-mix; NZ = 2 = the Y2 ring depth, so
      // slot (it + NZ) % NZ is the one iteration `it` used) -- possibly with the first chunks of the next tile, whose
      // convolution ran on the tensor pipe meanwhile
      if (it + Cfg::NZ < npairs2) drain(it + Cfg::NZ);
      upsample(it);  // (after the drain: the T-mix warps get their next chunk first)
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(uint32_t(Cfg::TMEM_COLS)) : "memory");
  }
}

}  // namespace mcd
