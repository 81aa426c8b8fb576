// mcd_block_tc.cuh -- the fused ST-GCN block with the 1x1 channel contraction on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), for the blocks whose channel counts
// make the contraction dense (Cin, Cout in {32, 64, 128}).
//
// Replaces ST_GCNN_layer.forward, models/gcae/stsgcn.py:94-116 (+ :143-156), like
// stgcn_block_kernel in mcd_kernels.cuh, with the same HBM layout (planar-4 [n][C/4][P][4] fp32).
//
// Precision: the reference is fp32 and its DDPM chain amplifies denoiser errors ~200x,
// so a plain TF32 product (10-bit mantissa) is not admissible.  Every fp32 operand x is split as
//     x = hi + lo,   hi = tf32-truncation of x (what the MMA reads from the raw fp32 bit pattern),
//                    lo = x - hi (exact in fp32, at most 13 significant bits),
// and each product is formed by three MMAs  lo*hi + hi*lo + hi*hi  into the same fp32 TMEM
// accumulator (3xTF32).  tools/tc_probe.cu measures max |error| 3.3e-6 on |values| up to 12.7 at
// K=64 against fp64 -- the class of an fp32 FMA chain (profiles/r01_tc_probe.log).
//
// Structure of one CTA (persistent, one per SM), 4 warpgroups (register budgets re-balanced with setmaxnreg),
// every hand-off through shared-memory mbarriers:
//   warp 13     activation loader: cp.async.bulk (TMA 1-D) copies of the next 16-channel chunk X -- one per
//               (window, 4-channel plane), contiguous in the planar-4 source, issued lane-parallel;
//   warp 14     weight loader: one bulk copy of the chunk's pre-swizzled weight operands; completing on mbarriers;
//   warps 0-3   T-mix  Y1[q,v,c] = sum_t X[t,v,c] T[v,t,q]: a thread owns (joint v, a group of output frames) and
//               keeps its slice of the learned T matrix in REGISTERS for the whole launch, so the only shared
//               traffic is one 16-byte activation read per 8 packed FMAs (the v2 kernel was shared-memory bound);
//   warps 4-7   A-mix  Y2[t,w,c] = sum_v Y1[t,v,c] A[t,v,w]: a thread owns (frame t, a group of output joints) with
//               its slice of A in registers; writes Y2 and its tf32 "lo" part (and the lo part of X for blocks
//               with a residual convolution) straight into the UMMA K-major SWIZZLE_64B operand layout;
//   warp 12     MMA issue: tcgen05.mma kind::tf32, A = activations [128 rows x 8], B = BN-folded weights [Cout x 8],
//               D = TMEM [128 lanes x Cout] per 128-row tile; tcgen05.commit releases operand buffers and
//               publishes finished accumulators;
//   warps 8-11  epilogue (one per TMEM lane quarter): tcgen05.ld, bias, identity residual, PReLU, time/condition
//               embedding, channel-last store -- while the other warps already work on the next tile (TMEM holds
//               two accumulator sets).
// T-mix of chunk c+1, A-mix of chunk c, the MMAs of chunk c-1 and the epilogue of the previous tile overlap.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "mcd_kernels.cuh"

#ifndef MCD_TC_TRACE
#define MCD_TC_TRACE 0
#endif
namespace mcd {

constexpr int kTcMix = 128;        // threads per mix group (T-warps 0-3, A-warps 4-7)
constexpr int kTcEpiWarp0 = 8;     // epilogue warps 8-11 (TMEM lane quarters 0,1,2,3)
constexpr int kTcMmaWarp = 12;     // the MMA-issuing warp
constexpr int kTcEpilogue = 128;
constexpr int kTcLoadWarp = 13;    // the activation loader warp; warp 14 loads weights; warp 15 computes the lo part of X
                                   // for blocks with a residual convolution
constexpr int kTcThreads = 16 * 32;
// Register budgets per warpgroup (setmaxnreg): the T-mix threads keep 96 weights + two accumulator sets in registers.
constexpr int kRegsT = 200, kRegsA = 152, kRegsE = 104, kRegsS = 56;
static_assert(kRegsT + kRegsA + kRegsE + kRegsS <= 512, "one warp of each group shares an SM sub-partition (16K registers)");
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

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
// UMMA shared-memory matrix descriptor, K-major, no swizzle ("interleave"): 8-row x 16-byte core matrices stored as 128
// contiguous bytes; SBO = distance between core matrices adjacent in M (128 bytes: rows are contiguous 16-byte
// elements), LBO = distance between the two core matrices of a K=8 step (`lbo_bytes`: the next 4-channel plane).
__device__ __forceinline__ uint64_t umma_desc_planar(uint32_t saddr, uint32_t lbo_bytes) {
  constexpr uint32_t hi = (128u >> 4) | (1u << 14);
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
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
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
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
// cp.async completion of the executing thread counts as one (pre-counted) arrival
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// TMA tiled copy (cp.async.bulk.tensor, SASS UTMALDG) of one 4-D box; out-of-range coordinates are zero-filled and the
// whole box is counted on the mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// The wait names the loaded registers as in/out operands: every use of them is ordered after it by data dependence.
__device__ __forceinline__ void tmem_ld_wait16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
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
      : "r"(taddr));
}
// The wait names the loaded registers as in/out operands: every use of them is ordered after it by data dependence.
__device__ __forceinline__ void tmem_ld_wait(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

__device__ __forceinline__ float4 ldg_nc4(const float* p) {  // read-only global load, streaming
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// (no "memory" clobber: the kernel never reads what it stores, and a clobber would pin every shared-memory load of the
//  epilogue behind the previous store; predicated, so that a row past the end of the tensor costs no branch)
__device__ __forceinline__ void stg4_pred(float* p, const float4 v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.global.v4.f32 [%0], {%1,%2,%3,%4};\n\t}\n" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "r"(int(pred)));
}

// shared-memory 16-byte load the compiler will not sink towards its use (keeps the software pipeline's lead)
__device__ __forceinline__ float4 lds4_early(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

// predicated 16-byte shared store (a guaranteed @p STS, not a branch)
__device__ __forceinline__ void sts4_pred(float* p, const float4 v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n\t}\n" ::"r"(smem_u32(p)), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "r"(int(pred))
               : "memory");
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

// group-size chooser for the register-resident mixes: `per_window` tasks exist per window for group size g;
// efficiency of 128 threads over NW windows
constexpr int tc_ceil(int a, int b) { return (a + b - 1) / b; }
constexpr int tc_eff1000(int per_window, int nw) {
  if (per_window > kTcMix) return 0;
  const int ws = kTcMix / per_window;           // window slots
  const int used = ws < nw ? ws : nw;
  const int rounds = tc_ceil(nw, used);
  return 1000 * per_window * nw / (rounds * kTcMix);
}
constexpr int tc_pick_group(int outer, int inner, int nw) {  // tasks per window = outer * ceil(inner / g), g in {4,3,2}
  int best = 4, best_eff = -1;
  for (int g = 4; g >= 2; --g) {
    const int e = tc_eff1000(outer * tc_ceil(inner, g), nw);
    if (e > best_eff) { best_eff = e; best = g; }
  }
  return best;
}
// Task shape of a register-resident mix for one-window tiles (T=24): a thread owns (channel split cs of CS, outer index,
// group of g inner outputs), keeps len*g weights in registers (<= reg_max) and makes C4/CS passes over the 4-channel
// groups of the chunk.  Cost model (measured behaviour of the v2 kernel: both the FMA pipe of the busiest
// sub-partition and the CTA-wide shared-memory pipe matter; every warp-wide LDS.128 costs 4 wavefronts):
//   fma  = passes * len * g * 2 packed FMAs * 2 cycles        smem = active warps * passes * len loads * 4 wavefronts
struct TcShape { int g, cs; };
constexpr TcShape tc_pick_shape(int outer, int inner, int len, int reg_max, int g_max, int c4) {
  TcShape best{2, 1};
  int best_cost = 1 << 30;
  for (int g = 2; g <= g_max; ++g) {
    if (len * g > reg_max) continue;
    for (int cs = 1; cs <= c4; cs *= 2) {
      const int tasks = cs * outer * tc_ceil(inner, g);
      if (tasks > kTcMix) continue;
      const int passes = c4 / cs, warps = tc_ceil(tasks, 32);
      const int cost = passes * len * g * 2 * 2 + warps * passes * len * 4;
      if (cost < best_cost) { best_cost = cost; best = TcShape{g, cs}; }
    }
  }
  return best;
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
  // Identity residual through the tensor pipe (short windows, narrow levels): the block is run as if it had a residual
  // convolution whose weight is the identity (X_hi*I + X_lo*I = X exactly).  The epilogue then never re-reads X from global
  // memory -- at T=3 those loads were 40 % of the epilogue warps' stall samples (profiles/r02_ncu_T3_summary.txt) -- at the
  // price of tensor-pipe reads of X / Xlo, which these blocks have room for (shared memory 30 % busy).  Kept off where the
  // extra Xlo buffer would cost the second Y2 buffer (V=17) and at long windows (measured at T=24, V=12: no change -- the T-mix
  // warps, not the epilogue, are the critical role there).
  static constexpr bool IDRES_MMA = CIN == COUT && T <= 4 && V <= 12;
  static constexpr bool RESCONV = CIN != COUT || IDRES_MMA;
  static constexpr int NPART = RESCONV ? 4 : 2;  // weight operand parts per chunk: W hi, W lo [, Wr hi, Wr lo]
  // Multi-window tiles (T <= 12) are loaded with ONE tiled TMA copy per 4-channel plane (box = [frame row, T, 1 plane, NW
  // windows]) instead of one bulk copy per (window, plane): at T=3 the 32 small copies per chunk cost the loader warp ~3 k
  // cycles of issue and made it the bottleneck of the V=12 / V=10 blocks (profiles/r02_waits_T3.log)
  static constexpr bool TMA_TILED = NW > 1;
  static constexpr bool RES_AHEAD = true;        // residual MMAs one chunk ahead of the convolution MMAs (see the MMA warp)
  static constexpr int VP = (V + 3) / 4 * 4;
  static constexpr int TP4 = (T + 3) / 4 * 4;
  static constexpr int TMS = T * TP4 + 4;
  // T-mix: thread = (window slot, joint v, group of QG output frames);  A-mix: thread = (window slot, frame t, WGS joints)
  //        (one-window tiles additionally split the chunk's four 4-channel groups over CS thread sets, tc_pick_shape)
  static constexpr bool SHAPED = NW == 1 && T > 4;
  // V=12, T=24: groups of 6 output frames x 2 channel splits (144 weights per thread, no spills under the 200-register
  // budget) instead of the model's 3 x 1: the same FMA count on the busiest sub-partition and half the X re-reads (each X
  // element is read by T/QG threads) -- the V=12 blocks run 3-4 % faster (round 2, gpurun e1: 27.1 -> 26.1 ms per 27 launches
  // of the 64->64 blocks); groups of 5 x 2, which the model rates equal, measured in between
  static constexpr TcShape SH_T = (V == 12 && T == 24) ? TcShape{6, 2} : tc_pick_shape(V, T, T, 100, 8, C4),
                           SH_A = tc_pick_shape(T, V, V, 76, 5, C4);
  // (A-mix groups of 6 joints at V=12 measured slower: two joint groups per frame make the operand stores conflict)
  static constexpr int QG = T <= 4 ? T : (SHAPED ? SH_T.g : tc_pick_group(V, T, NW));
  static constexpr int CS_T = SHAPED ? SH_T.cs : 1;
  static constexpr int NQG = tc_ceil(T, QG);
  static constexpr int TTR = V * NQG;                                  // (joint, frame group) pairs
  static constexpr int TT = CS_T * TTR;                                // T-mix tasks per window
  static constexpr int WS_T = (kTcMix / TT) < NW ? (kTcMix / TT) : NW; // window slots
  static constexpr int WGS = SHAPED ? SH_A.g : tc_pick_group(T, V, NW);
  static constexpr int CS_A = SHAPED ? SH_A.cs : 1;
  static constexpr int NWG = tc_ceil(V, WGS);
  static constexpr int TAR = T * NWG;                                  // (frame, joint group) pairs
  static constexpr int TA = CS_A * TAR;
  static constexpr int WS_A = (kTcMix / TA) < NW ? (kTcMix / TA) : NW;
  static_assert(TT <= kTcMix && TA <= kTcMix && WS_T >= 1 && WS_A >= 1, "mix task mapping");
  // "N-merge" (blocks with a residual convolution whose accumulators fit twice): the hi and lo weight parts sit back to back
  // in the operand array, so ONE MMA with N = 2*COUT forms act_hi*W_hi (columns [0,COUT)) and act_hi*W_lo (columns
  // [COUT,2*COUT)) from a single read of the activation operand; a second MMA adds act_lo*W_hi.  Two MMAs and 14 KB of
  // operand reads per product instead of three and 18 KB (COUT=64) -- the V=10 blocks are bound by exactly that traffic.
  // The epilogue adds the two column halves.
  static constexpr bool NMERGE = RESCONV && 2 * MT * 2 * COUT <= 512 && 2 * COUT <= 256;
  static constexpr int TCOLS = NMERGE ? 2 * COUT : COUT;  // accumulator columns per 128-row tile
  static constexpr int ACC_COLS = MT * TCOLS;  // one accumulator set; TMEM holds two (tile parity)
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512);
  static_assert(CIN % KC == 0 && COUT % 32 == 0 && COUT <= 256, "tensor-core block: Cin multiple of 16, Cout multiple of 32");
  static_assert(2 * ACC_COLS <= 512, "two accumulator sets exceed TMEM");
  static_assert(ROWS % 8 == 0, "operand arrays must be whole swizzle atoms");
  // shared memory carve-up, in floats from a 1024-byte aligned base
  static constexpr int ARR = ROWS * 16;          // one operand array [ROWS][16] (a multiple of 512 bytes)
  // Y1 is not an MMA operand, so its layout is chosen for the two mixes: 16-byte (4-channel) elements indexed
  // [window][c4][q within the frame group][T-mix thread (v, qg)] with the thread pitch TTP = 2 (mod 8): the T-mix
  // threads of a warp store consecutive elements, the A-mix threads of a warp (consecutive frames) hit distinct banks
  static constexpr int TTP = T >= 8 ? (TTR + 5) / 8 * 8 + 2 : (TTR | 1);
  static constexpr int Y1ARR = ((NW * 4 * QG * TTP * 4) + 127) / 128 * 128;
  static constexpr int WCH = NPART * COUT * 16;  // one chunk of weight operands
  // residual convolution: the planar X buffer itself is the "hi" operand (no-swizzle K-major descriptor); only its
  // tf32 lo part is materialised, in the same planar layout, by the conversion warp (2 buffers)
  // Buffer counts by what fits into 227 KB: the X ring (planar [c4][window][position] 16-byte elements) is 3 deep where
  // possible -- the refill of a slot (release -> bulk copy issue -> ~1.5 k cycles of latency) then hides behind two mix
  // iterations instead of one; the widest tiles (V=17 with a residual convolution) fall back to single Xlo / Y2 buffers.
  static constexpr int SM_MISC = 2 * Y1ARR + 2 * WCH + COUT + NW * COUT;
  static constexpr bool tc_fits(int arrays) { return size_t(SM_MISC + arrays * ARR) * sizeof(float) + 1024 <= 227 * 1024; }
  // One-window tiles (T=24): Xlo / Y2 double buffers first, then the third X slot (tuned on the T=24 wait accounting).
  // Multi-window tiles (short windows): the X ring first -- a chunk is little work there, and with the tiled TMA loader the
  // round trip "slot released -> copy issued -> data landed" (~2-3 k cycles) is what the T-mix warps wait for (they sat 50-60 %
  // in x_full with a 2-deep ring while the loader idled, profiles/r02_waits_T3.log); up to 4 slots, Xlo single-buffered.
  static constexpr int tc_budget() { int a = 0; while (a < 16 && tc_fits(a + 1)) ++a; return a; }
  static constexpr int BUDGET = tc_budget();
  static constexpr bool DEEP_X = NW > 1;
  static constexpr int NXLO0 = !RESCONV ? 0 : (tc_fits(2 + 2 + 4) ? 2 : 1);
  static constexpr int NY2_0 = tc_fits(2 + NXLO0 + 4) ? 2 : 1;
  static constexpr int NXB0 = tc_fits(3 + NXLO0 + 2 * NY2_0) ? 3 : 2;
  static constexpr int NY2_D = BUDGET >= 2 + (RESCONV ? 1 : 0) + 4 ? 2 : 1;
  static constexpr int NXB_D0 = BUDGET - (RESCONV ? 1 : 0) - 2 * NY2_D;
  static constexpr int NXB_D = NXB_D0 > 4 ? 4 : (NXB_D0 < 2 ? 2 : NXB_D0);
  static constexpr int NXLO_D = !RESCONV ? 0 : (BUDGET - NXB_D - 2 * NY2_D >= 2 ? 2 : 1);
  static constexpr int NXLO = DEEP_X ? NXLO_D : NXLO0;
  static constexpr int NY2 = DEEP_X ? NY2_D : NY2_0;
  static constexpr int NXB = DEEP_X ? NXB_D : NXB0;
  static_assert(tc_fits(NXB + NXLO + 2 * NY2), "tensor-core block tile exceeds shared memory");
  static constexpr int SM_X = 0;                 // NXB buffers
  static constexpr int SM_Y1 = SM_X + NXB * ARR;  // 2
  static constexpr int SM_XLO = SM_Y1 + 2 * Y1ARR;
  static constexpr int SM_Y2 = SM_XLO + NXLO * ARR;           // NY2
  static constexpr int SM_Y2LO = SM_Y2 + NY2 * ARR;           // NY2
  static constexpr int SM_WC = SM_Y2LO + NY2 * ARR;           // 2; also absorbs the last tile's over-read
  static constexpr int SM_BIAS = SM_WC + 2 * WCH;
  static constexpr int SM_EMB = SM_BIAS + COUT;
  static constexpr int SM_TOTAL = SM_EMB + NW * COUT;
  static_assert((MT * 128 - ROWS) * 16 <= 2 * WCH, "over-read of the last MMA tile must stay inside the allocation");
  static constexpr size_t SMEM_BYTES = size_t(SM_TOTAL) * sizeof(float) + 1024;  // + alignment slack
};

// barrier slots in shared memory
enum TcBar {
  BAR_X_FULL = 0,                 // [4] loader -> T-warps (and A-warps / MMA with a residual convolution)
  BAR_X_EMPTY = 4,                // [4] consumers -> loader
  BAR_W_FULL = 8,                 // [2] loader (bulk copy) -> MMA warp
  BAR_Y1_FULL = 10,               // [2] T-warps -> A-warps
  BAR_Y1_EMPTY = 12,              // [2] A-warps -> T-warps
  BAR_OPS_FULL = 14,              // [2] A-warps -> MMA warp
  BAR_MMA_DONE = 16,              // [2] tcgen05.commit -> A-warps, loader (operand buffers free)
  BAR_ACC_FULL = 18,              // [2] tcgen05.commit -> epilogue
  BAR_ACC_EMPTY = 20,             // [2] epilogue -> MMA warp
  BAR_XLO_FULL = 22,              // [2] conversion warp -> MMA warp (residual convolution)
  BAR_RES_DONE = 24,              // [2] tcgen05.commit -> conversion warp (Xlo buffer free)
  BAR_COUNT = 26
};
constexpr int kTcMaxXSlots = 4;

template <class Cfg>
__global__ void __launch_bounds__(kTcThreads, 1) stgcn_block_tc_kernel(const BlockWeights wt, const BlockIO io,
                                                                       const __grid_constant__ CUtensorMap tmx) {
  constexpr int T = Cfg::T, V = Cfg::V, P = Cfg::P, ROWS = Cfg::ROWS, C4 = Cfg::C4, MT = Cfg::MT;
  constexpr int CIN = Cfg::CIN, COUT = Cfg::COUT, NCHUNK = Cfg::NCHUNK, NW = Cfg::NW, NXB = Cfg::NXB;
  constexpr int VP = Cfg::VP, TP4 = Cfg::TP4, TMS = Cfg::TMS, ARR = Cfg::ARR, WCH = Cfg::WCH;
  constexpr int QG = Cfg::QG, NQG = Cfg::NQG, WGS = Cfg::WGS, NWG = Cfg::NWG, TTP = Cfg::TTP, Y1ARR = Cfg::Y1ARR;
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
  float* sBias = smem + Cfg::SM_BIAS;
  float* sEmb = smem + Cfg::SM_EMB;
  __shared__ __align__(8) uint64_t bars[BAR_COUNT];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (io.n + NW - 1) / NW;
  if (int64_t(blockIdx.x) >= ntiles) return;  // uniform over the CTA
  const int my_tiles = int((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const int npairs = my_tiles * NCHUNK;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int slot) { return bar0 + uint32_t(slot) * 8u; };
  // debug timeline (mcd_debug_trace_next): lane 0 of each role's first warp in CTA 0 appends (role, pair, event, clock).
  // Compiled in only with -DMCD_TC_TRACE=1 (tools/trace_block.py builds that variant): the inline trace code costs
  // instruction fetches on every loop iteration of every role.
#if MCD_TC_TRACE
  __shared__ int trace_n;
  if (tid == 0) trace_n = 0;
  auto TRACE = [&](int role, int it, int ev) {
    if (io.trace != nullptr && io.trace_cap > 0 && blockIdx.x == 0 && lane == 0) {  // trace_cap < 0: wait accounting only
      const int k = atomicAdd(&trace_n, 1);
      if (k < io.trace_cap) {
        io.trace[4 * k + 0] = role; io.trace[4 * k + 1] = it; io.trace[4 * k + 2] = ev; io.trace[4 * k + 3] = clock64();
      }
    }
  };
  // wait accounting: cycles this thread spent in each of its (up to 4) barrier waits, reported at the end of the role
  long long wacc[4] = {0, 0, 0, 0};
  auto WAIT = [&](int slot, uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    wacc[slot] += clock64() - t0;
  };
  const long long role_t0 = clock64();
  auto PHASE = [&](int slot, long long t0) { wacc[slot] += clock64() - t0; };  // free-form phase timer (epilogue)
  auto WAIT_REPORT = [&](int role) {  // records (role, 100 + slot, 0, cycles waited) and (role, 99, 0, cycles in the role)
    if (io.trace != nullptr && blockIdx.x == 0 && lane == 0) {
      for (int k = 0; k < 5; ++k) {
        const int r = atomicAdd(&trace_n, 1);
        if (r < (io.trace_cap < 0 ? -io.trace_cap : io.trace_cap)) {
          io.trace[4 * r + 0] = role; io.trace[4 * r + 1] = k < 4 ? 100 + k : 99; io.trace[4 * r + 2] = 0;
          io.trace[4 * r + 3] = k < 4 ? wacc[k] : clock64() - role_t0;
        }
      }
    }
  };
#else
  auto TRACE = [](int, int, int) {};
  auto WAIT = [&](int, uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); };
  auto WAIT_REPORT = [](int) {};
  auto PHASE = [](int, long long) {};
#endif
#if MCD_TC_TRACE
#define MCD_CLOCK() clock64()
#else
#define MCD_CLOCK() 0ll
#endif

  // ---- once per CTA ----
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(uint32_t(Cfg::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (tid == 0) {
    for (int i = 0; i < kTcMaxXSlots; ++i) {
      mbar_init(BAR(BAR_X_FULL + i), 1);  // one expect_tx arrival + the bulk copies' bytes
      mbar_init(BAR(BAR_X_EMPTY + i), RESCONV ? kTcMix + 1 : kTcMix);  // T-warps (+ the commit of the residual MMAs)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(BAR_W_FULL + i), 1);
      mbar_init(BAR(BAR_Y1_FULL + i), kTcMix);
      mbar_init(BAR(BAR_Y1_EMPTY + i), kTcMix);
      mbar_init(BAR(BAR_OPS_FULL + i), kTcMix);
      mbar_init(BAR(BAR_MMA_DONE + i), 1);
      mbar_init(BAR(BAR_ACC_FULL + i), 1);
      mbar_init(BAR(BAR_ACC_EMPTY + i), kTcEpilogue);
      mbar_init(BAR(BAR_XLO_FULL + i), 32);
      mbar_init(BAR(BAR_RES_DONE + i), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp >= kTcEpiWarp0 && warp < kTcLoadWarp)
    for (int i = tid - kTcEpiWarp0 * 32; i < COUT; i += kTcEpilogue) sBias[i] = wt.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // =============================== T-mix warps ===============================
    reg_inc<kRegsT>();
    // Y1[n,(q,v),c] = sum_t X[n,(t,v),c] * Tm[v][t][q]        stsgcn.py:154
    const int ws = tid / Cfg::TT, rem_all = tid - ws * Cfg::TT;
    const int cs = rem_all / Cfg::TTR, rem = rem_all - cs * Cfg::TTR;  // channel split, (joint, frame group)
    const int v = rem / NQG, qg = rem - v * NQG;
    const bool active = ws < Cfg::WS_T;
    float wT[T][QG];  // this thread's slice of the learned time-mix matrix, resident for the whole launch
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int q = 0; q < QG; ++q)
        wT[t][q] = (active && qg * QG + q < T) ? __ldg(wt.Tm + v * TMS + t * TP4 + qg * QG + q) : 0.f;

    for (int it = 0; it < npairs; ++it) {
      const int b = it % NXB, s = it & 1;
      if (warp == 0) TRACE(0, it, 0);
      WAIT(0, BAR(BAR_X_FULL + b), uint32_t((it / NXB) & 1));
      if (warp == 0) TRACE(0, it, 1);
      if (it >= 2) WAIT(1, BAR(BAR_Y1_EMPTY + s), uint32_t((it / 2 - 1) & 1));
      if (warp == 0) TRACE(0, it, 2);
      if (active) {
        const float* sXc = sX + b * ARR;
        float* sY = sY1 + s * Y1ARR;
        for (int wl = ws; wl < NW; wl += Cfg::WS_T) {
          // One 4-channel group per pass: the fully unrolled body (T*QG*2 packed FMAs -- the weights live in registers, so
          // the frame loop cannot be rolled) must stay small enough for the ~6 KB L0 instruction cache it shares with the
          // A-mix warp of the same sub-partition (ncu: 28 % of the mix warps' stall samples were instruction fetches with
          // two groups per pass).  Latency is covered by a software pipeline over blocks of TB frames instead.
#pragma unroll 1
          for (int c4 = cs; c4 < C4; c4 += Cfg::CS_T) {
            float2 a[2][QG];
#pragma unroll
            for (int q = 0; q < QG; ++q) a[0][q] = a[1][q] = make_float2(0.f, 0.f);
            constexpr int TB = QG > 4 ? (T % 3 == 0 ? 3 : 2) : (T % 6 == 0 ? 6 : (T % 3 == 0 ? 3 : (T % 2 == 0 ? 2 : 1)));
            const float* xp = sXc + ((c4 * NW + wl) * P + v) * 4;  // planar X: [c4][window][position] 16-byte elements
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
      if (warp == 0) TRACE(0, it, 3);
      mbar_arrive(BAR(BAR_Y1_FULL + s));
      mbar_arrive(BAR(BAR_X_EMPTY + b));
    }
    if (warp == 0) WAIT_REPORT(0);
  } else if (warp < 8) {
    // =============================== A-mix warps ===============================
    reg_inc<kRegsA>();
    // Y2[n,(t,w),c] = sum_v Y1[n,(t,v),c] * A[t][v][w]       stsgcn.py:155   (+ tf32 lo parts for the tensor pipe)
    const int atid = tid - kTcMix;
    const int ws = atid / Cfg::TA, rem_all = atid - ws * Cfg::TA;
    const int cs = rem_all / Cfg::TAR, rem = rem_all - cs * Cfg::TAR;  // channel split, (frame, joint group)
    // thread wg of a frame owns output joints wg, wg+NWG, ...; frames are dealt to consecutive thread groups with a
    // stride FS coprime to T chosen so that the 8 lanes of a quarter warp store to 8 different bank groups
    constexpr int FS = (V % 8 == 1 && T % 5 != 0) ? 5 : 1;
    const int tk = rem / NWG, wg = rem - tk * NWG;
    const int t = (tk * FS) % T;
    const bool active = ws < Cfg::WS_A;          // (so the NWG threads of a frame store consecutive operand rows)
    float wA[V][WGS];  // this thread's slice of the learned joint-mix matrix
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int j = 0; j < WGS; ++j)
        wA[v][j] = (active && wg + j * NWG < V) ? __ldg(wt.A + (t * V + v) * VP + wg + j * NWG) : 0.f;

    for (int it = 0; it < npairs; ++it) {
      const int s = it & 1;
      if (warp == 4) TRACE(1, it, 0);
      WAIT(0, BAR(BAR_Y1_FULL + s), uint32_t((it / 2) & 1));
      if (warp == 4) TRACE(1, it, 1);
      constexpr int NY2 = Cfg::NY2;  // Y2 / Y2lo buffer it % NY2 was last read by the MMAs of iteration it - NY2
      if (it >= NY2) WAIT(1, BAR(BAR_MMA_DONE + ((it - NY2) & 1)), uint32_t(((it - NY2) / 2) & 1));
      if (warp == 4) TRACE(1, it, 2);
      if (active) {
        const float* sY = sY1 + s * Y1ARR;
        float* sZ = sY2 + (it % NY2) * ARR;
        float* sZlo = sY2lo + (it % NY2) * ARR;
        for (int wl = ws; wl < NW; wl += Cfg::WS_A) {
          const int r0 = wl * P + t * V;
#pragma unroll 1
          for (int c4 = cs; c4 < C4; c4 += Cfg::CS_A) {  // one 4-channel group per pass (small unrolled body, see the T-mix)
            float2 a[2][WGS];
#pragma unroll
            for (int j = 0; j < WGS; ++j) a[0][j] = a[1][j] = make_float2(0.f, 0.f);
            const float* yp = sY + (((wl * 4 + c4) * QG + t % QG) * TTP + t / QG) * 4;  // element (v, t) at + v * NQG
            constexpr int VB = V * WGS > 60 ? 4 : (V >= 12 ? 6 : (V >= 6 ? 5 : V));  // joints per pipeline block (the last may be partial)
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
#pragma unroll
            for (int j = 0; j < WGS; ++j) {  // predicated stores (no branches: every taken branch costs an instruction fetch)
              const int w = wg + j * NWG;
              const bool okw = (V % WGS == 0 && NWG * WGS == V) || w < V;
              const float4 o = make_float4(a[0][j].x, a[0][j].y, a[1][j].x, a[1][j].y);
              const int off = sw_off(r0 + (okw ? w : 0), c4);
              sts4_pred(sZ + off, o, okw);
              sts4_pred(sZlo + off, tf32_lo4(o), okw);
            }
          }
        }
      }
      if (warp == 4) TRACE(1, it, 3);
      mbar_arrive(BAR(BAR_Y1_EMPTY + s));
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor pipe
      if (warp == 4) TRACE(1, it, 4);
      mbar_arrive(BAR(BAR_OPS_FULL + s));
    }
    if (warp == 4) WAIT_REPORT(1);
  } else if (warp >= kTcMmaWarp) {
    reg_dec<kRegsS>();
    if (warp == kTcMmaWarp) {
    // =============================== MMA-issuing warp ===============================
    const uint32_t idesc = umma_idesc_tf32(COUT), idesc2 = umma_idesc_tf32(2 * COUT);
    constexpr bool NMERGE = Cfg::NMERGE;
    constexpr int TCOLS = Cfg::TCOLS;
    // operand descriptors of every buffer, computed once (warp-uniform): inside the loop a descriptor is base + constant
    const uint64_t dY2_0 = umma_desc_sw64(smem_u32(sY2)), dY2lo_0 = umma_desc_sw64(smem_u32(sY2lo));
    const uint64_t dW_0 = umma_desc_sw64(smem_u32(sWc)), dW_1 = umma_desc_sw64(smem_u32(sWc + WCH));
    // residual convolution: X (planar, no swizzle) and its lo part; descriptor of buffer k = base + k * ARR bytes
    const uint64_t dX_0 = umma_desc_planar(smem_u32(sX), ROWS * 16), dXlo_0 = umma_desc_planar(smem_u32(sXlo), ROWS * 16);
    constexpr uint64_t ARR16 = (uint64_t(ARR) * 4) >> 4;   // one operand array, in descriptor address units (16 bytes)
    constexpr uint64_t PART = (COUT * 64) >> 4;            // one weight part, in descriptor address units
    // residual 1x1 convolution of iteration j, straight from the landed X buffer: issued as soon as the lo part exists,
    // i.e. long before the mixes of the chunk finish, so the X ring slot is released early
    auto issue_res = [&](int j) {
      const int tj = j / NCHUNK, cj = j - tj * NCHUNK;
      const int set = tj & 1, s = j & 1, b = j % NXB;
      const uint64_t bW = s ? dW_1 : dW_0;
      const uint32_t d0 = tmem + set * Cfg::ACC_COLS;
      WAIT(0, BAR(BAR_W_FULL + s), uint32_t((j / 2) & 1));
      if (cj == 0 && tj >= 2) WAIT(1, BAR(BAR_ACC_EMPTY + set), uint32_t((tj / 2 - 1) & 1));  // set drained by the epilogue
      WAIT(2, BAR(BAR_XLO_FULL + s), uint32_t((j / 2) & 1));
      tc_fence_after();
      const uint64_t xHi = dX_0 + uint64_t(b) * ARR16, xLo = dXlo_0 + uint64_t(j % (Cfg::NXLO > 0 ? Cfg::NXLO : 1)) * ARR16;
      if (elect_one()) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint32_t d = d0 + m * TCOLS;
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // two K=8 steps per 16-channel chunk: c4 planes (2h, 2h+1)
            const uint64_t ao = uint64_t((m * 128 * 16 + h * 2 * ROWS * 16) >> 4), bo = uint64_t((h * 32) >> 4);
            const uint32_t first = (cj > 0 || h > 0) ? 1u : 0u;  // the very first MMA of a tile clears its accumulator columns
            if constexpr (NMERGE) {
              umma_tf32(d, xHi + ao, bW + 2 * PART + bo, idesc2, first);  // X_hi * [Wr_hi | Wr_lo]
              umma_tf32(d, xLo + ao, bW + 2 * PART + bo, idesc, 1u);      // X_lo * Wr_hi
            } else {
              umma_tf32(d, xLo + ao, bW + 2 * PART + bo, idesc, first);
              umma_tf32(d, xHi + ao, bW + 3 * PART + bo, idesc, 1u);
              umma_tf32(d, xHi + ao, bW + 2 * PART + bo, idesc, 1u);
            }
          }
        }
        umma_commit(BAR(BAR_RES_DONE + s));  // Xlo buffer free again
        umma_commit(BAR(BAR_X_EMPTY + b));   // ... and the X ring slot (together with the T-warps' arrivals)
      }
      __syncwarp();
    };
    int res_issued = 0;  // residual MMAs are in the pipe for iterations < res_issued
    for (int it = 0; it < npairs; ++it) {
      const int ti = it / NCHUNK, chunk = it - ti * NCHUNK;
      const int set = ti & 1, s = it & 1;
      const uint64_t bW = s ? dW_1 : dW_0;
      const uint32_t d0 = tmem + set * Cfg::ACC_COLS;
      TRACE(2, it, 0);
      if constexpr (RESCONV) {
        if (res_issued <= it) { issue_res(it); res_issued = it + 1; }
        // One chunk ahead inside a tile: the residual MMAs of chunk c+1 enter the pipe before the (later) convolution
        // MMAs of chunk c and free their X slot a whole iteration earlier.  Never across tiles (that would make this
        // tile's last convolution wait for the epilogue of the previous tile), and unconditionally inside one, so
        // the accumulation order -- hence every bit of the result -- does not depend on timing.
        if (Cfg::RES_AHEAD && chunk + 1 < NCHUNK) { issue_res(it + 1); res_issued = it + 2; }
      } else {
        WAIT(0, BAR(BAR_W_FULL + s), uint32_t((it / 2) & 1));
        if (chunk == 0 && ti >= 2) WAIT(1, BAR(BAR_ACC_EMPTY + set), uint32_t((ti / 2 - 1) & 1));  // set drained by the epilogue
      }
      TRACE(2, it, 1);
      WAIT(3, BAR(BAR_OPS_FULL + s), uint32_t((it / 2) & 1));
      TRACE(2, it, 2);
      tc_fence_after();
      const uint64_t aHi = dY2_0 + uint64_t(it % Cfg::NY2) * ARR16, aLo = dY2lo_0 + uint64_t(it % Cfg::NY2) * ARR16;
      const uint32_t acc0 = (RESCONV || chunk > 0) ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint32_t d = d0 + m * TCOLS;
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // two K=8 steps per 64-byte operand row
            const uint64_t ao = uint64_t((m * 128 * 64 + h * 32) >> 4), bo = uint64_t((h * 32) >> 4);
            if constexpr (NMERGE) {
              umma_tf32(d, aHi + ao, bW + bo, idesc2, 1u);  // Y2_hi * [W_hi | W_lo]   (the residual MMAs cleared the tile)
              umma_tf32(d, aLo + ao, bW + bo, idesc, 1u);   // Y2_lo * W_hi
            } else {
              umma_tf32(d, aLo + ao, bW + bo, idesc, h == 0 ? acc0 : 1u);
              umma_tf32(d, aHi + ao, bW + PART + bo, idesc, 1u);
              umma_tf32(d, aHi + ao, bW + bo, idesc, 1u);
            }
          }
        }
        umma_commit(BAR(BAR_MMA_DONE + s));
        if (chunk == NCHUNK - 1) umma_commit(BAR(BAR_ACC_FULL + set));  // the tile's accumulators are complete
      }
      __syncwarp();
      TRACE(2, it, 3);
    }
    WAIT_REPORT(2);
    } else if (warp == kTcLoadWarp) {
    // =============================== activation loader warp ===============================
    for (int it = 0; it < npairs; ++it) {
      const int ti = it / NCHUNK, chunk = it - ti * NCHUNK;
      const int64_t tile = blockIdx.x + int64_t(ti) * gridDim.x;
      const int b = it % NXB;
      TRACE(3, it, 0);
      if (it >= NXB) WAIT(0, BAR(BAR_X_EMPTY + b), uint32_t((it / NXB - 1) & 1));
      TRACE(3, it, 1);
      // one bulk copy per (window, 4-channel plane): P x 16 contiguous bytes of the planar-4 source land as one
      // plane of the planar X buffer [c4][window][position] (rows = (window, position) are contiguous inside a
      // 4-channel plane, which makes the buffer a valid no-swizzle K-major UMMA operand as it stands); windows past
      // the end of the tensor are skipped (their rows are never stored).
      // The copies are issued by different lanes (a single lane needed ~1-2.6 k cycles for the 4*NW issues).
      constexpr uint32_t PLANE = P * 16;
      const uint32_t dst0 = smem_u32(sX + b * ARR);
      if constexpr (Cfg::TMA_TILED) {
        // tensor [4V floats | T | CIN/4 planes | n windows] (mcd_api.cu: make_x_tensor_map); windows past the end arrive as zeros
        if (lane == 0) mbar_expect_tx(BAR(BAR_X_FULL + b), uint32_t(NW) * 4u * PLANE);
        __syncwarp();
        if (lane < 4)
          tma_load_4d(dst0 + uint32_t(lane * NW) * PLANE, &tmx, 0, 0, chunk * C4 + lane, int(tile * NW), BAR(BAR_X_FULL + b));
      } else {
        int64_t nvalid = io.n - tile * NW;
        if (nvalid > NW) nvalid = NW;
        if (lane == 0) mbar_expect_tx(BAR(BAR_X_FULL + b), uint32_t(nvalid) * 4u * PLANE);
        __syncwarp();
        for (int k = lane; k < int(nvalid) * 4; k += 32) {
          const int wl = k >> 2, j = k & 3;
          bulk_g2s(dst0 + uint32_t(j * NW + wl) * PLANE, io.in + act_off(tile * NW + wl, chunk * C4 + j, 0, CIN, P), PLANE, BAR(BAR_X_FULL + b));
        }
      }
      TRACE(3, it, 2);
      __syncwarp();
    }
    WAIT_REPORT(3);
    } else if (warp == kTcLoadWarp + 1) {
    // =============================== weight loader warp ===============================
    // its own warp, so that waiting for a free weight buffer (MMAs of iteration it-2) never delays the activations
    for (int it = 0; it < npairs; ++it) {
      const int chunk = it % NCHUNK, s = it & 1;
      if (it >= 2) WAIT(0, BAR(BAR_MMA_DONE + s), uint32_t((it / 2 - 1) & 1));  // W[s] free again
      if (lane == 0) {
        mbar_expect_tx(BAR(BAR_W_FULL + s), uint32_t(WCH * 4));
        bulk_g2s(smem_u32(sWc + s * WCH), wt.Bop + size_t(chunk) * WCH, uint32_t(WCH * 4), BAR(BAR_W_FULL + s));
      }
      __syncwarp();
    }
    WAIT_REPORT(5);
    } else if constexpr (RESCONV) {
    // =============================== conversion warp (15) ===============================
    // tf32 lo part of the landed X chunk, same planar layout (an element-wise pass), for the residual convolution
    for (int it = 0; it < npairs; ++it) {
      constexpr int NXLO = Cfg::NXLO > 0 ? Cfg::NXLO : 1;
      const int b = it % NXB;
      WAIT(0, BAR(BAR_X_FULL + b), uint32_t((it / NXB) & 1));
      // Xlo buffer it % NXLO was last read by the residual MMAs of iteration it - NXLO
      if (it >= NXLO) WAIT(1, BAR(BAR_RES_DONE + ((it - NXLO) & 1)), uint32_t(((it - NXLO) / 2) & 1));
      const float4* src = reinterpret_cast<const float4*>(sX + b * ARR);
      float4* dst = reinterpret_cast<float4*>(sXlo + (it % NXLO) * ARR);
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
      mbar_arrive(BAR(BAR_XLO_FULL + (it & 1)));
    }
    WAIT_REPORT(6);
    }
  } else {
    // =============================== epilogue warps ===============================
    reg_dec<kRegsE>();
    // TMEM -> bias, identity residual, PReLU, + Linear(SiLU(pos + cond)) -> channel-last store      stsgcn.py:109-114
    const float slope = wt.prelu;
    const int q = warp & 3;  // TMEM lane quarter this warp may access (warps 8-11 -> quarters 0-3)
    const int etid = tid - kTcEpiWarp0 * 32;
    // Linear(SiLU(pos + cond)) of the tile's windows comes precomputed (time_embedding_kernel); each thread carries its
    // share of the NEXT tile's values in registers so the global latency never sits on the per-tile critical path
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
    load_emb(blockIdx.x);
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int64_t tile = blockIdx.x + int64_t(ti) * gridDim.x;
      const int set = ti & 1;
      const long long t_emb = MCD_CLOCK();
      named_bar_sync(2, kTcEpilogue);  // everyone is done with the previous tile's sEmb
#pragma unroll
      for (int k = 0; k < EPER; ++k)
        if (etid + k * kTcEpilogue < NW * COUT) sEmb[etid + k * kTcEpilogue] = epre[k];
      named_bar_sync(2, kTcEpilogue);
      if (ti + 1 < my_tiles) load_emb(tile + gridDim.x);  // the next tile's values travel during this tile's stores
      PHASE(1, t_emb);
      if (warp == kTcEpiWarp0) TRACE(4, ti, 0);
      WAIT(0, BAR(BAR_ACC_FULL + set), uint32_t((ti / 2) & 1));
      if (warp == kTcEpiWarp0) TRACE(4, ti, 1);
      tc_fence_after();
      if constexpr (NW == 1) {
        // Column groups of 16 channels outermost, 128-row tiles inside: the group's bias (and, for one-window tiles, its
        // embedding values -- the same for every row) are read from shared memory ONCE per tile instead of once per row tile.
        // With the row tiles outermost those broadcast loads were 14-20 % of the kernel's shared-memory wavefronts, the pipe
        // the mixes are bound by (ncu, 64->64 block at 12 joints: 384 of 2 740 wavefronts per chunk).
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 16) {
          const float4* bp = reinterpret_cast<const float4*>(sBias + c0);
          const float4* ep = reinterpret_cast<const float4*>(sEmb + c0);
          const float4 b4[4] = {bp[0], bp[1], bp[2], bp[3]}, e4[4] = {ep[0], ep[1], ep[2], ep[3]};
#pragma unroll 1
          for (int m = 0; m < MT; ++m) {
            const int r = m * 128 + q * 32 + lane;
            const int wl = r / P;
            const int64_t w = tile * NW + wl;
            const bool ok = (r < ROWS) && (w < io.n);
            const int pp = r - wl * P;  // consecutive lanes -> consecutive positions: planar-4 loads / stores are contiguous
            float4 xr[4];
            const float* ip = io.in + (ok ? act_off(w, c0 >> 2, pp, CIN, P) : 0);
            float* op = io.out + (ok ? act_off(w, c0 >> 2, pp, COUT, P) : 0);
            if constexpr (!RESCONV) {  // identity residual: issue the loads of this column group before touching TMEM
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) xr[j4] = ldg_nc4(ip + j4 * P * 4);  // 4-channel planes are P elements apart
            }
            uint32_t acc[16];
            const long long t_ld = MCD_CLOCK();
            const uint32_t tcol = tmem + (uint32_t(q * 32) << 16) + uint32_t(set * Cfg::ACC_COLS + m * Cfg::TCOLS + c0);
            tmem_ld16(tcol, acc);
            uint32_t acc2[Cfg::NMERGE ? 16 : 1];
            if constexpr (Cfg::NMERGE) tmem_ld16(tcol + COUT, acc2);  // the act_hi * W_lo partial sums
            tmem_ld_wait16(acc);
            if constexpr (Cfg::NMERGE) {
              tmem_ld_wait16(acc2);
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(acc2[i]));
            }
            PHASE(2, t_ld);
            const long long t_st = MCD_CLOCK();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              float o[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                float v = __uint_as_float(acc[j4 * 4 + jj]) + f4get(b4[j4], jj);
                if constexpr (!RESCONV) v += f4get(xr[j4], jj);
                v = v > 0.f ? v : slope * v;
                o[jj] = v + f4get(e4[j4], jj);
              }
              stg4_pred(op + j4 * P * 4, make_float4(o[0], o[1], o[2], o[3]), ok);
            }
            PHASE(3, t_st);
          }
        }
      } else {
        // several windows per tile (short windows): the embedding row depends on the row's window, so nothing but the bias
        // could be hoisted, and 32-channel groups keep twice the identity-residual loads in flight per thread -- with
        // 16-channel groups the 32->32 blocks at 17 joints ran 36 % slower at T=3 (the epilogue's load latency chain)
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
          const int r = m * 128 + q * 32 + lane;
          const int wl = r / P;
          const int64_t w = tile * NW + wl;
          const bool ok = (r < ROWS) && (w < io.n);
          const float* embp = sEmb + (ok ? wl : 0) * COUT;
          const int pp = r - wl * P;  // consecutive lanes -> consecutive positions: planar-4 loads / stores are contiguous
#pragma unroll 1
          for (int c0 = 0; c0 < COUT; c0 += 32) {
            float4 xr[8];
            const float* ip = io.in + (ok ? act_off(w, c0 >> 2, pp, CIN, P) : 0);
            float* op = io.out + (ok ? act_off(w, c0 >> 2, pp, COUT, P) : 0);
            if constexpr (!RESCONV) {  // identity residual: issue the loads of this column group before touching TMEM
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) xr[j4] = ldg_nc4(ip + j4 * P * 4);  // 4-channel planes are P elements apart
            }
            uint32_t acc[32];
            const long long t_ld = MCD_CLOCK();
            const uint32_t tcol = tmem + (uint32_t(q * 32) << 16) + uint32_t(set * Cfg::ACC_COLS + m * Cfg::TCOLS + c0);
            tmem_ld32(tcol, acc);
            uint32_t acc2[Cfg::NMERGE ? 32 : 1];
            if constexpr (Cfg::NMERGE) tmem_ld32(tcol + COUT, acc2);  // the act_hi * W_lo partial sums
            // bias / embedding of the first two 4-channel groups travel while the TMEM load is in flight; the rest is
            // fetched two groups ahead (shared-memory latency is ~100 cycles with the mixes and the tensor pipe on the port)
            const float4* bp = reinterpret_cast<const float4*>(sBias + c0);
            const float4* ep = reinterpret_cast<const float4*>(embp + c0);
            float4 b4[2] = {bp[0], bp[1]}, e4[2] = {ep[0], ep[1]};
            tmem_ld_wait(acc);
            if constexpr (Cfg::NMERGE) {
              tmem_ld_wait(acc2);
#pragma unroll
              for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(acc2[i]));
            }
            PHASE(2, t_ld);
            const long long t_st = MCD_CLOCK();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bc = b4[j4 & 1], ec = e4[j4 & 1];
              if (j4 + 2 < 8) { b4[j4 & 1] = bp[j4 + 2]; e4[j4 & 1] = ep[j4 + 2]; }
              float o[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                float v = __uint_as_float(acc[j4 * 4 + jj]) + f4get(bc, jj);
                if constexpr (!RESCONV) v += f4get(xr[j4], jj);
                v = v > 0.f ? v : slope * v;
                o[jj] = v + f4get(ec, jj);
              }
              stg4_pred(op + j4 * P * 4, make_float4(o[0], o[1], o[2], o[3]), ok);
            }
            PHASE(3, t_st);
          }
        }
      }
      tc_fence_before();  // accumulator reads ordered before the release of the set
      if (warp == kTcEpiWarp0) TRACE(4, ti, 2);
      mbar_arrive(BAR(BAR_ACC_EMPTY + set));
    }
    if (warp == kTcEpiWarp0) WAIT_REPORT(4);
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(uint32_t(Cfg::TMEM_COLS)) : "memory");
  }
}

}  // namespace mcd
