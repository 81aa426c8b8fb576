// mcd_block_tc.cuh -- the fused ST-GCN block with the 1x1 channel contraction on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), for the blocks whose channel counts
// make the contraction dense (Cin, Cout in {32, 64, 128}).
//
// Replaces ST_GCNN_layer.forward, models/gcae/stsgcn.py:94-116 (+ :143-156), like
// stgcn_block_kernel in mcd_kernels.cuh, with the same HBM layout (channel-last [n][P][C] fp32).
//
// Precision: the reference is fp32 and its DDPM chain amplifies denoiser errors ~200x,
// so a plain TF32 product (10-bit mantissa) is not admissible.  Every fp32 operand x is split as
//     x = hi + lo,   hi = tf32-truncation of x (what the MMA reads from the raw fp32 bit pattern),
//                    lo = x - hi (exact in fp32, at most 13 significant bits),
// and each product is formed by three MMAs  lo*hi + hi*lo + hi*hi  into the same fp32 TMEM
// accumulator (3xTF32).  tools/tc_probe.cu measures max |error| 3.3e-6 on |values| up to 12.7 at
// K=64 against fp64 -- the class of an fp32 FMA chain (profiles/r01_tc_probe.log).
//
// Structure of one CTA (persistent, one per SM): 8 compute warps + 1 MMA-issuing warp + 4 epilogue warps.
//   compute warps, per 16-channel chunk of the input:
//     T-mix  X -> Y1, A-mix Y1 -> Y2 (fp32 FMA pipe, packed FFMA2), writing Y2 and its lo part (and the
//     lo part of X when the block has a residual convolution) straight into the UMMA K-major
//     SWIZZLE_64B operand layout; then fence.proxy.async + named-barrier arrive;
//   MMA warp: waits on the named barrier, issues the chunk's tcgen05.mma's (A = activations rows x 16,
//     B = BN-folded weights Cout x 16, D = TMEM [128 lanes x Cout] per 128-row tile), commits to an
//     mbarrier; the tensor pipe runs while the compute warps do the T-mix of the next chunk;
//   epilogue warps (one per TMEM lane quarter): when a tile's last commit lands they read the accumulators
//     (tcgen05.ld 32x32b), apply bias, identity residual, PReLU and the time/condition embedding and store
//     channel-last, while the compute warps already mix the next tile; TMEM holds two accumulator sets.
#pragma once
#include "mcd_kernels.cuh"

namespace mcd {

constexpr int kTcCompute = 256;                     // compute threads (warps 0-7)
constexpr int kTcMmaWarp = 8;                       // the MMA-issuing warp
constexpr int kTcEpilogue = 128;                    // epilogue threads (warps 9-12: TMEM lane quarters 1,2,3,0)
constexpr int kTcThreads = kTcCompute + 32 + kTcEpilogue;
// named barriers: 1 = operands ready (compute -> MMA warp), 2 = compute-only sync, 3/4 = embedding of an even/odd tile ready (compute -> epilogue)

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_64B: 64-byte rows, 8-row groups 512 bytes apart.
// bits [0,14) address>>4 | [16,30) LBO>>4 (=1, unused) | [32,46) SBO>>4 | [46,48) version 1 | [61,64) layout 4
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  constexpr uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29);
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  return (uint64_t(hi) << 32) | lo;
}
// instruction descriptor, kind::tf32: D fp32, A/B tf32 K-major, M=128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a wedged pipeline traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < (1 << 24); ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float4 ldg_nc4(const float* p) {  // read-only global load, streaming
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg4(float* p, const float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// float index of (row r, 4-channel group c4) in a [rows][16] fp32 operand array laid out K-major SWIZZLE_64B
// (the array base is 512-byte aligned): the 16-byte chunk index is XORed with bits 1..2 of the row.
__device__ __forceinline__ int sw_off(int r, int c4) { return r * 16 + ((c4 ^ ((r >> 1) & 3)) << 2); }

__device__ __forceinline__ float4 tf32_lo4(const float4 v) {  // v - tf32_truncate(v), exact
  float4 o;
  o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  return o;
}

template <int T_, int V_, int CIN_, int COUT_, int NW_>
struct TcCfg {
  static constexpr int T = T_, V = V_, CIN = CIN_, COUT = COUT_, NW = NW_;
  static constexpr int P = T * V;
  static constexpr int ROWS = NW * P;
  static constexpr int MT = (ROWS + 127) / 128;  // 128-row MMA tiles (the last one over-reads into the next array)
  static constexpr int KC = 16;
  static constexpr int NCHUNK = CIN / KC;
  static constexpr int C4 = KC / 4;
  static constexpr bool RESCONV = CIN != COUT;
  static constexpr int NPART = RESCONV ? 4 : 2;  // weight operand parts per chunk: W hi, W lo [, Wr hi, Wr lo]
  static constexpr int VP = (V + 3) / 4 * 4;
  static constexpr int TP4 = (T + 3) / 4 * 4;
  static constexpr int TMS = T * TP4 + 4;
  static constexpr int TQ = TP4 <= 8 ? TP4 : 8;
  static constexpr int NQT = TP4 / TQ;
  static constexpr int NWT = 2;
  static constexpr int TW = VP / NWT;
  static constexpr int ACC_COLS = MT * COUT;  // one accumulator set; TMEM holds two (tile parity)
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512);
  static_assert(CIN % KC == 0 && COUT % 32 == 0 && COUT <= 256, "tensor-core block: Cin multiple of 16, Cout multiple of 32");
  static_assert(2 * MT * COUT <= 512, "two accumulator sets exceed TMEM");
  static_assert(TP4 % TQ == 0 && VP % (2 * NWT) == 0, "tile shapes");
  // shared memory carve-up, in floats from a 1024-byte aligned base
  static constexpr int ARR = (ROWS * 16 + 127) / 128 * 128;  // operand array stride: multiple of 512 bytes
  static constexpr int WCH = NPART * COUT * 16;              // one chunk of weight operands
  static constexpr int SM_X = 0;                             // 2 buffers
  static constexpr int SM_Y1 = SM_X + 2 * ARR;
  static constexpr int SM_XLO = SM_Y1 + ARR;
  static constexpr int SM_Y2 = SM_XLO + (RESCONV ? ARR : 0);
  static constexpr int SM_Y2LO = SM_Y2 + ARR;
  static constexpr int SM_WC = SM_Y2LO + ARR;                // 2 buffers; also absorbs the last tile's over-read
  static constexpr int SM_A = SM_WC + 2 * WCH;
  static constexpr int SM_TM = SM_A + T * V * VP;
  static constexpr int SM_BIAS = SM_TM + V * TMS;
  static constexpr int SM_EMB = SM_BIAS + COUT;
  static constexpr int SM_S = SM_EMB + 2 * NW * COUT;  // sEmb is double-buffered like the accumulators
  static constexpr int SM_TOTAL = SM_S + NW * kMaxE;
  static_assert((MT * 128 - ROWS) * 16 <= 2 * WCH + T * V * VP, "over-read of the last MMA tile must stay inside the allocation");
  static constexpr size_t SMEM_BYTES = size_t(SM_TOTAL) * sizeof(float) + 1024;  // + alignment slack
};

template <class Cfg>
__device__ __forceinline__ void tc_prefetch_x(const BlockIO& io, float* sXbuf, int64_t tile, int chunk, int tid) {
  const int64_t row0 = tile * Cfg::ROWS;
  const int64_t nrows = io.n * Cfg::P;
  const float* base = io.in + chunk * Cfg::KC;
  for (int idx = tid; idx < Cfg::ROWS * Cfg::C4; idx += kTcCompute) {
    const int r = idx >> 2, j = idx & 3;
    const bool ok = (row0 + r) < nrows;
    const float* src = ok ? base + (row0 + r) * Cfg::CIN + j * 4 : io.in;
    cp_async16(sXbuf + sw_off(r, j), src, ok);
  }
}
template <class Cfg>
__device__ __forceinline__ void tc_prefetch_w(const BlockWeights& wt, float* sWbuf, int chunk, int tid) {
  const float* wsrc = wt.Bop + size_t(chunk) * Cfg::WCH;
  for (int idx = tid; idx < Cfg::WCH / 4; idx += kTcCompute) cp_async16(sWbuf + idx * 4, wsrc + idx * 4, true);
}

template <class Cfg>
__global__ void __launch_bounds__(kTcThreads, 1) stgcn_block_tc_kernel(const BlockWeights wt, const BlockIO io) {
  constexpr int T = Cfg::T, V = Cfg::V, P = Cfg::P, ROWS = Cfg::ROWS, C4 = Cfg::C4, MT = Cfg::MT;
  constexpr int CIN = Cfg::CIN, COUT = Cfg::COUT, NCHUNK = Cfg::NCHUNK, NW = Cfg::NW;
  constexpr int VP = Cfg::VP, TP4 = Cfg::TP4, TMS = Cfg::TMS, ARR = Cfg::ARR, WCH = Cfg::WCH;
  constexpr int TQ = Cfg::TQ, NQT = Cfg::NQT, TW = Cfg::TW, NWT = Cfg::NWT;
  constexpr bool RESCONV = Cfg::RESCONV;

  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B with pointer arithmetic on the shared array (keeps the shared address space: LDS/STS, not generic)
  float* smem = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  float* sX = smem + Cfg::SM_X;
  float* sY1 = smem + Cfg::SM_Y1;
  float* sXlo = smem + Cfg::SM_XLO;
  float* sY2 = smem + Cfg::SM_Y2;
  float* sY2lo = smem + Cfg::SM_Y2LO;
  float* sWc = smem + Cfg::SM_WC;
  float* sA = smem + Cfg::SM_A;
  float* sTm = smem + Cfg::SM_TM;
  float* sBias = smem + Cfg::SM_BIAS;
  float* sEmb = smem + Cfg::SM_EMB;
  float* sS = smem + Cfg::SM_S;
  // mbarriers: [0] per-pair MMA commit (operand buffers free), [1],[2] accumulator set full, [3],[4] set drained
  __shared__ __align__(8) uint64_t bars[5];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (io.n + NW - 1) / NW;
  if (int64_t(blockIdx.x) >= ntiles) return;  // uniform over the CTA
  const int my_tiles = int((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const int npairs = my_tiles * NCHUNK;

  // ---- once per CTA ----
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(uint32_t(Cfg::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp < kTcMmaWarp) {
    if (tid == 0) {
      mbar_init(smem_u32(&bars[0]), 1);
      mbar_init(smem_u32(&bars[1]), 1);
      mbar_init(smem_u32(&bars[2]), 1);
      mbar_init(smem_u32(&bars[3]), kTcEpilogue);
      mbar_init(smem_u32(&bars[4]), kTcEpilogue);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    static_assert((T * V * VP) % 4 == 0 && (V * TMS) % 4 == 0 && COUT % 4 == 0, "16-byte weight copies");
    for (int i = tid; i < T * V * VP / 4; i += kTcCompute) cp_async16(sA + i * 4, wt.A + i * 4, true);
    for (int i = tid; i < V * TMS / 4; i += kTcCompute) cp_async16(sTm + i * 4, wt.Tm + i * 4, true);
    cp_async_commit();  // waited for together with the first activation chunk
  } else {
    for (int i = tid - (kTcCompute + 32); i < COUT; i += kTcEpilogue) sBias[i] = wt.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t bar_mma = smem_u32(&bars[0]);

  if (warp == kTcMmaWarp) {
    // =============================== MMA-issuing warp ===============================
    const uint32_t idesc = umma_idesc_tf32(COUT);
    for (int it = 0; it < npairs; ++it) {
      const int ti = it / NCHUNK, chunk = it - ti * NCHUNK;
      const int set = ti & 1;
      named_bar_sync(1, kTcCompute + 32);  // operands of pair `it` are in shared memory (compute warps fenced + arrived)
      if (chunk == 0 && ti >= 2) mbar_wait(smem_u32(&bars[3 + set]), uint32_t((ti / 2 - 1) & 1));  // set drained by the epilogue
      tc_fence_after();
      if (lane == 0) {
        const int buf = it & 1;
        const uint32_t aY2 = smem_u32(sY2), aY2lo = smem_u32(sY2lo);
        const uint32_t aX = smem_u32(sX + buf * ARR), aXlo = smem_u32(sXlo);
        const uint32_t bW = smem_u32(sWc + buf * WCH);
        constexpr uint32_t PART = COUT * 64;  // bytes per weight part
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint32_t d = tmem + set * Cfg::ACC_COLS + m * COUT;
          const uint32_t moff = m * 128 * 64;
          uint32_t acc = chunk > 0 ? 1u : 0u;
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // two K=8 steps per 64-byte operand row
            const uint32_t ko = h * 32;
            umma_tf32(d, umma_desc_sw64(aY2lo + moff + ko), umma_desc_sw64(bW + ko), idesc, acc);
            umma_tf32(d, umma_desc_sw64(aY2 + moff + ko), umma_desc_sw64(bW + PART + ko), idesc, 1u);
            umma_tf32(d, umma_desc_sw64(aY2 + moff + ko), umma_desc_sw64(bW + ko), idesc, 1u);
            acc = 1u;
            if constexpr (RESCONV) {
              umma_tf32(d, umma_desc_sw64(aXlo + moff + ko), umma_desc_sw64(bW + 2 * PART + ko), idesc, 1u);
              umma_tf32(d, umma_desc_sw64(aX + moff + ko), umma_desc_sw64(bW + 3 * PART + ko), idesc, 1u);
              umma_tf32(d, umma_desc_sw64(aX + moff + ko), umma_desc_sw64(bW + 2 * PART + ko), idesc, 1u);
            }
          }
        }
        umma_commit(bar_mma);
        if (chunk == NCHUNK - 1) umma_commit(smem_u32(&bars[1 + set]));  // the tile's accumulators are complete
      }
      __syncwarp();
    }
  } else if (warp > kTcMmaWarp) {
    // =============================== epilogue warps ===============================
    // TMEM -> bias, identity residual, PReLU, + emb -> channel-last store            stsgcn.py:109-114
    const float slope = wt.prelu;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int64_t tile = blockIdx.x + int64_t(ti) * gridDim.x;
      const int set = ti & 1;
      named_bar_sync(3 + set, kTcCompute + kTcEpilogue);  // sEmb[set] of this tile is written
      mbar_wait(smem_u32(&bars[1 + set]), uint32_t((ti / 2) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int r = m * 128 + q * 32 + lane;
        const int wl = r / P;
        const int64_t w = tile * NW + wl;
        const bool ok = (r < ROWS) && (w < io.n);
        const float* embp = sEmb + (set * NW + (ok ? wl : 0)) * COUT;
        const int64_t grow = tile * ROWS + r;
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 32) {
          float4 xr[8];
          if constexpr (!RESCONV) {  // identity residual: issue the loads of this column group before touching TMEM
            const float* src = io.in + (ok ? grow : 0) * CIN + c0;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) xr[j4] = ldg_nc4(src + j4 * 4);
          }
          uint32_t acc[32];
          tmem_ld32(tmem + (uint32_t(q * 32) << 16) + uint32_t(set * Cfg::ACC_COLS + m * COUT + c0), acc);
          float* dst = io.out + grow * COUT + c0;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float o[4];
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + j4 * 4);
            const float4 e4 = *reinterpret_cast<const float4*>(embp + c0 + j4 * 4);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float v = __uint_as_float(acc[j4 * 4 + jj]) + f4get(b4, jj);
              if constexpr (!RESCONV) v += f4get(xr[j4], jj);
              v = v > 0.f ? v : slope * v;
              o[jj] = v + f4get(e4, jj);
            }
            if (ok) stg4(dst + j4 * 4, make_float4(o[0], o[1], o[2], o[3]));
          }
        }
      }
      tc_fence_before();                   // accumulator reads ordered before the release of the set
      mbar_arrive(smem_u32(&bars[3 + set]));
    }
  } else {
    // =============================== compute warps ===============================
    tc_prefetch_x<Cfg>(io, sX, blockIdx.x, 0, tid);
    tc_prefetch_w<Cfg>(wt, sWc, 0, tid);
    cp_async_commit();
    int waited = 0;  // per-pair MMA commits observed so far

    for (int it = 0; it < npairs; ++it) {
      const int ti = it / NCHUNK, chunk = it - ti * NCHUNK;
      const int64_t tile = blockIdx.x + int64_t(ti) * gridDim.x;
      const int set = ti & 1;
      float* sXc = sX + (it & 1) * ARR;
      const bool more = it + 1 < npairs;
      const int nti = (it + 1) / NCHUNK, nchunk = (it + 1) - nti * NCHUNK;
      const int64_t ntile = blockIdx.x + int64_t(nti) * gridDim.x;

      cp_async_wait_all();
      named_bar_sync(2, kTcCompute);  // X / W chunk visible; the previous pair's mix finished everywhere

      // identity blocks: X is not an MMA operand, so its other buffer is free now -- prefetch a whole pair ahead
      if constexpr (!RESCONV) {
        if (more) tc_prefetch_x<Cfg>(io, sX + ((it + 1) & 1) * ARR, ntile, nchunk, tid);
      }

      // time/condition embedding input  pos + cond (stsgcn.py:112-114), once per tile: the global loads are
      // issued here and consumed after the T-mix, which hides their latency
      float temb = 0.f;
      if (chunk == 0 && tid < NW * io.E) {
        const int wl = tid / io.E, j = tid - wl * io.E;
        const int64_t w = tile * NW + wl;
        temb = __ldg(io.pos + j);
        if (io.cond != nullptr && w < io.n) temb += __ldg(io.cond + ((io.w0 + w) % io.condB) * io.E + j);
      }

      // ---- T-mix   Y1[n,(q,v),c] = sum_t X[n,(t,v),c] * Tm[v][t][q]     stsgcn.py:154
      for (int task = tid; task < NQT * NW * V * C4; task += kTcCompute) {
        const int c4 = task % C4;
        const int col = (task / C4) % (NW * V);
        const int qt = task / (C4 * NW * V);
        const int wl = col / V, v = col - wl * V;
        float2 a[2][TQ];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int q = 0; q < TQ; ++q) a[c][q] = make_float2(0.f, 0.f);
        const float* tp = sTm + v * TMS + qt * TQ;
        const int r0 = wl * P + v;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float4 x = *reinterpret_cast<const float4*>(sXc + sw_off(r0 + t * V, c4));
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
#pragma unroll
        for (int q = 0; q < TQ; ++q) {
          const int qq = qt * TQ + q;
          if (qq < T)
            *reinterpret_cast<float4*>(sY1 + sw_off(r0 + qq * V, c4)) = make_float4(a[0][q].x, a[0][q].y, a[1][q].x, a[1][q].y);
        }
      }
      if (chunk == 0 && tid < NW * io.E) {
        const int wl = tid / io.E, j = tid - wl * io.E;
        sS[wl * kMaxE + j] = temb / (1.0f + expf(-temb));  // SiLU
      }

      // the tensor pipe may still be reading Y2 / Y2lo / Xlo / X[other] / W[other] of the previous pair
      while (waited < it) { mbar_wait(bar_mma, uint32_t(waited & 1)); ++waited; }
      if (more) {
        if constexpr (RESCONV) tc_prefetch_x<Cfg>(io, sX + ((it + 1) & 1) * ARR, ntile, nchunk, tid);
        tc_prefetch_w<Cfg>(wt, sWc + ((it + 1) & 1) * WCH, nchunk, tid);
      }
      cp_async_commit();
      named_bar_sync(2, kTcCompute);  // Y1 and sS complete

      // ---- A-mix   Y2[n,(t,w),c] = sum_v Y1[n,(t,v),c] * A[t][v][w]      stsgcn.py:155   (+ tf32 lo part)
      for (int task = tid; task < NW * T * NWT * C4; task += kTcCompute) {
        const int c4 = task % C4;
        const int wtile = (task / C4) % NWT;
        const int row = task / (C4 * NWT);
        const int wl = row / T, q = row - wl * T;
        float2 a[2][TW];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j = 0; j < TW; ++j) a[c][j] = make_float2(0.f, 0.f);
        const int r0 = wl * P + q * V;
        const float* ap = sA + q * V * VP + wtile * TW;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float4 y = *reinterpret_cast<const float4*>(sY1 + sw_off(r0 + v, c4));
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
#pragma unroll
        for (int j = 0; j < TW; ++j) {
          const int w = wtile * TW + j;
          if (w < V) {
            const float4 o = make_float4(a[0][j].x, a[0][j].y, a[1][j].x, a[1][j].y);
            const int off = sw_off(r0 + w, c4);
            *reinterpret_cast<float4*>(sY2 + off) = o;
            *reinterpret_cast<float4*>(sY2lo + off) = tf32_lo4(o);
          }
        }
      }
      if constexpr (RESCONV) {  // lo part of X for the residual convolution (elementwise: the layouts coincide)
        for (int idx = tid; idx < ROWS * C4; idx += kTcCompute)
          *reinterpret_cast<float4*>(sXlo + idx * 4) = tf32_lo4(*reinterpret_cast<const float4*>(sXc + idx * 4));
      }
      if (chunk == 0) {  // emb = Linear(SiLU(pos + cond)) for the windows of this tile -> sEmb[set], read by the epilogue warps
        if (ti >= 2) mbar_wait(smem_u32(&bars[3 + set]), uint32_t((ti / 2 - 1) & 1));  // epilogue of tile ti-2 is done with it
        const int E = io.E;
        for (int i = tid; i < NW * COUT; i += kTcCompute) {
          const int wl = i / COUT, co = i - wl * COUT;
          float e = __ldg(wt.bE + co);
          for (int j = 0; j < E; ++j) e = fmaf(__ldg(wt.WEt + j * COUT + co), sS[wl * kMaxE + j], e);
          sEmb[set * NW * COUT + i] = e;
        }
        named_bar_arrive(3 + set, kTcCompute + kTcEpilogue);  // (one barrier per set: its next use is two tiles later)
      }
      fence_proxy_async();                      // generic-proxy writes (st.shared, cp.async) -> visible to the tensor pipe
      named_bar_arrive(1, kTcCompute + 32);     // hand the operands to the MMA warp
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
