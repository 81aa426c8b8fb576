// mcd_latent.cuh -- the latent-space variant of the scoring loop (MoCoDADlatent, stage 'diffusion').
//
// Replaces, in the reference tree:
//   models/mocodad_latent.py:100-127     per generated sample: x_T ~ N(0,1) [B, latent], then for t = N-1..1 one call of the MLP
//                                        denoiser and the DDPM update on vectors; loss against the latent code
//   models/common/components.py:264-291  Denoiser.forward: per layer  x = act(BN(Linear(x))) + Linear_c(pos(t) + cond),
//                                        act = ReLU, no BN / ReLU on the last layer (build_model :231-245)
// (the latent code itself -- STSE_Unet at the constant step t = -1, stsae_unet.py:222-246 -- is the down half of the denoiser
//  and runs on the block kernels; see latent_encode_impl in mcd_api.cu.)
//
// One persistent CTA per SM keeps the whole network (BatchNorm1d folded into the linear layers, weights transposed to
// [in][out]) in shared memory and carries a tile of kLatVec latent vectors through ALL noise_steps-1 denoiser calls: the
// vectors never leave the SM between x_T and the loss, so HBM sees the conditioning rows, the latent code and (when noise is
// injected) the noise tensor only.  fp32 FMA throughout; the DDPM update uses the same separately rounded operations as
// ddpm_step_kernel.
#pragma once
#include "mcd_kernels.cuh"

namespace mcd {

constexpr int kLatMaxLayers = 8;
constexpr int kLatVec = 32;       // latent vectors per CTA tile
constexpr int kLatThreads = 256;

struct LatentNet {
  int32_t n_layers, L, E, maxw, total;   // latent width, embedding width, widest layer, floats in the packed image
  int32_t in[kLatMaxLayers], out[kLatMaxLayers], relu[kLatMaxLayers];
  int32_t w_off[kLatMaxLayers];    // Wt  [in][out]   BN-folded
  int32_t b_off[kLatMaxLayers];    // b   [out]       BN-folded
  int32_t wc_off[kLatMaxLayers];   // Wct [E][out]    cond_layers[i].weight transposed
  int32_t bc_off[kLatMaxLayers];   // bc  [out]
};

struct LatentArgs {
  const float* img;        // packed network image (LatentNet offsets)
  const float* pos;        // [N + 1][E] pos_encoding table (row N = step -1)
  const float* coef;       // [N][3] DDPM coefficients c1, c2, c3 of step t
  const float* cond;       // [B][E] conditioning embedding or nullptr
  const float* code;       // [B][L] latent code (loss target); nullptr in single-call mode
  const float* noise;      // [G][N-1][B][L] injected noise or nullptr (Philox)
  const float* x_in;       // single-call mode: [n][L] input vectors
  float* x_out;            // full mode: x_0 [nv][L] or nullptr;  single-call mode: predicted noise [n][L]
  float* losses;           // [nv] per-vector loss (full mode)
  int64_t nv, B;           // virtual vectors (index g*B + b), vectors per sample
  int64_t first_window;    // Philox counter base
  uint64_t seed;
  int32_t N;               // noise_steps
  int32_t single_t;        // >= 0: single-call mode at this step
  int32_t loss_fn;
};

// One dense layer for the tile: hout[v][o] = act(sum_k Wt[k][o] hin[v][k] + b[o]) + sum_e Wct[e][o] cs[v][e] + bc[o]
// A thread computes 4 vectors x 4 outputs; consecutive lanes own consecutive output groups (conflict-free 16-byte weight
// reads), the activation reads of a warp are broadcasts.
__device__ __forceinline__ void latent_layer(const float* __restrict__ Wt, const float* __restrict__ b, const float* __restrict__ Wct,
                                             const float* __restrict__ bc, const float* __restrict__ cs, const float* hin, int hin_stride,
                                             float* hout, int hout_stride, int in, int out, int E, bool relu, int tid) {
  const int ogs = out >> 2;
  for (int task = tid; task < ogs * (kLatVec / 4); task += kLatThreads) {
    const int og = task % ogs, vg = task / ogs;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[j][o] = 0.f;
    for (int k = 0; k < in; k += 4) {
      float4 hv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) hv[j] = *reinterpret_cast<const float4*>(hin + (vg * 4 + j) * hin_stride + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w = *reinterpret_cast<const float4*>(Wt + (k + kk) * out + og * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float h = f4get(hv[j], kk);
          acc[j][0] = fmaf(h, w.x, acc[j][0]); acc[j][1] = fmaf(h, w.y, acc[j][1]);
          acc[j][2] = fmaf(h, w.z, acc[j][2]); acc[j][3] = fmaf(h, w.w, acc[j][3]);
        }
      }
    }
    const float4 b4 = *reinterpret_cast<const float4*>(b + og * 4), bc4 = *reinterpret_cast<const float4*>(bc + og * 4);
    float cadd[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { cadd[j][0] = bc4.x; cadd[j][1] = bc4.y; cadd[j][2] = bc4.z; cadd[j][3] = bc4.w; }
    for (int e = 0; e < E; ++e) {
      const float4 w = *reinterpret_cast<const float4*>(Wct + e * out + og * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float c = cs[(vg * 4 + j) * E + e];
        cadd[j][0] = fmaf(c, w.x, cadd[j][0]); cadd[j][1] = fmaf(c, w.y, cadd[j][1]);
        cadd[j][2] = fmaf(c, w.z, cadd[j][2]); cadd[j][3] = fmaf(c, w.w, cadd[j][3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = acc[j][q] + f4get(b4, q);
        if (relu) v = fmaxf(v, 0.f);
        o[q] = v + cadd[j][q];
      }
      *reinterpret_cast<float4*>(hout + (vg * 4 + j) * hout_stride + og * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

__global__ void __launch_bounds__(kLatThreads, 1) latent_diffusion_kernel(const __grid_constant__ LatentNet net,
                                                                          const __grid_constant__ LatentArgs a) {
  extern __shared__ __align__(16) float lat_smem[];
  float* sImg = lat_smem;                               // [total]
  float* hA = sImg + ((net.total + 3) & ~3);            // [kLatVec][maxw]
  float* hB = hA + kLatVec * net.maxw;
  float* xs = hB + kLatVec * net.maxw;                  // [kLatVec][L]  the vectors being denoised
  float* cs = xs + kLatVec * net.L;                     // [kLatVec][E]  pos(t) + cond
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = net.L, E = net.E;
  for (int i = tid; i < net.total; i += kLatThreads) sImg[i] = a.img[i];
  __syncthreads();

  const int64_t ntiles = (a.nv + kLatVec - 1) / kLatVec;
  const int n_slots = a.N - 1 > 1 ? a.N - 1 : 1;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t v0 = tile * kLatVec;
    // x_T (mocodad_latent.py:104) -- or the given vectors in single-call mode; rows past the end stay zero
    for (int i = tid; i < kLatVec * L; i += kLatThreads) {
      const int vl = i / L, o = i - vl * L;
      const int64_t v = v0 + vl;
      float x = 0.f;
      if (v < a.nv) {
        if (a.single_t >= 0) {
          x = a.x_in[v * L + o];
        } else {
          const int64_t g = v / a.B, b = v - g * a.B;
          x = a.noise != nullptr ? a.noise[((g * n_slots + 0) * a.B + b) * L + o]
                                 : philox_normal(a.seed, uint64_t(a.first_window + b), uint32_t(g), 0u, uint32_t(o));
        }
      }
      xs[i] = x;
    }
    int slot = 0;
    const int t_hi = a.single_t >= 0 ? a.single_t : a.N - 1, t_lo = a.single_t >= 0 ? a.single_t : 1;
    for (int t = t_hi; t >= t_lo; --t) {
      ++slot;
      for (int i = tid; i < kLatVec * E; i += kLatThreads) {   // components.py:276-283: cond = pos(t) + cond
        const int vl = i / E, e = i - vl * E;
        const int64_t v = v0 + vl;
        float c = a.pos[size_t(t < 0 ? a.N : t) * E + e];
        if (a.cond != nullptr && v < a.nv) c += a.cond[(v % a.B) * E + e];
        cs[i] = c;
      }
      __syncthreads();
      const float* hin = xs;
      int hin_stride = L;
      for (int l = 0; l < net.n_layers; ++l) {
        float* hout = (l & 1) ? hB : hA;
        latent_layer(sImg + net.w_off[l], sImg + net.b_off[l], sImg + net.wc_off[l], sImg + net.bc_off[l], cs, hin, hin_stride, hout,
                     net.maxw, net.in[l], net.out[l], E, net.relu[l] != 0, tid);
        __syncthreads();
        hin = hout;
        hin_stride = net.maxw;
      }
      if (a.single_t >= 0) {   // parity tap: the predicted noise itself
        for (int i = tid; i < kLatVec * L; i += kLatThreads) {
          const int vl = i / L, o = i - vl * L;
          if (v0 + vl < a.nv) a.x_out[(v0 + vl) * L + o] = hin[vl * hin_stride + o];
        }
        break;
      }
      // DDPM update, mocodad_latent.py:111-119 (same roundings as ddpm_step_kernel)
      const float c1 = a.coef[t * 3 + 0], c2 = a.coef[t * 3 + 1], c3 = a.coef[t * 3 + 2];
      for (int i = tid; i < kLatVec * L; i += kLatThreads) {
        const int vl = i / L, o = i - vl * L;
        const int64_t v = v0 + vl;
        float z = 0.f;
        if (t > 1 && v < a.nv) {
          const int64_t g = v / a.B, b = v - g * a.B;
          z = a.noise != nullptr ? a.noise[((g * n_slots + slot) * a.B + b) * L + o]
                                 : philox_normal(a.seed, uint64_t(a.first_window + b), uint32_t(g), uint32_t(slot), uint32_t(o));
        }
        const float eps = hin[vl * hin_stride + o];
        xs[i] = __fadd_rn(__fmul_rn(c1, __fsub_rn(xs[i], __fmul_rn(c2, eps))), __fmul_rn(c3, z));
      }
      __syncthreads();
    }
    if (a.single_t < 0) {
      // per-vector loss against the latent code (mocodad.py:484 on [B, latent]); one warp per vector
      for (int vl = warp; vl < kLatVec; vl += kLatThreads / 32) {
        const int64_t v = v0 + vl;
        if (v >= a.nv) continue;
        const int64_t b = v % a.B;
        float s = 0.f;
        for (int o = lane; o < L; o += 32) {
          const float x = xs[vl * L + o];
          if (a.x_out != nullptr) a.x_out[v * L + o] = x;
          const float d = x - a.code[b * L + o], ad = fabsf(d);
          s += a.loss_fn == 0 ? (ad < 1.0f ? 0.5f * d * d : ad - 0.5f) : (a.loss_fn == 1 ? ad : d * d);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) a.losses[v] = s / float(L);
      }
    }
    __syncthreads();   // xs / cs are rewritten by the next tile
  }
}

}  // namespace mcd
