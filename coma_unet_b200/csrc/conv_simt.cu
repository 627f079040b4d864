// Direct (gather) 3-D convolution on CUDA cores, NDHWC, fp32 accumulate.
//
// Role on the hot path: (1) the exact-fp32 path (north_star's <=1e-4 tolerance), (2) shapes the
// tcgen05 implicit GEMM does not take (Cin = 1 head conv, per-sample expert-mixed 1x1x1
// `reduce_channels`, 1-channel heads), (3) on-device cross-check for the tensor-core kernels.
// Replaces cuDNN fprop/dgrad/wgrad behind torch.nn.Conv3d / ConvTranspose3d
// (reference call sites: attn_unet_data_parallel.py:126,285-306,442,495-497,546-558).
#include <algorithm>

#include "common.cuh"

namespace coma {

constexpr int kSimtThreads = 128;

template <typename T, int COT, int VEC>
__global__ void __launch_bounds__(kSimtThreads) conv_simt_kernel(coma_conv_args a, int chunks) {
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
  const int64_t v = (int64_t)chunk * kSimtThreads + threadIdx.x;
  const bool valid = v < Vo;
  const int co0 = blockIdx.y * COT;
  const int K = a.ksize;
  float acc[COT];
#pragma unroll
  for (int j = 0; j < COT; ++j) acc[j] = 0.f;

  if (valid) {
    const int ow = (int)(v % a.Wo), oh = (int)((v / a.Wo) % a.Ho), od = (int)(v / ((int64_t)a.Wo * a.Ho));
    const T* xb = static_cast<const T*>(a.x) + (int64_t)b * a.Di * a.Hi * a.Wi * a.x_cs + a.x_co;
    const T* wb = static_cast<const T*>(a.w) + (int64_t)b * a.w_bstride;
    const float* isc = a.in_scale ? a.in_scale + (int64_t)b * a.Cin : nullptr;       // input prologue (reference implementation)
    const float* ish = a.in_scale ? a.in_shift + (int64_t)b * a.Cin : nullptr;
    const float ineg = a.in_act == COMA_ACT_NONE ? 1.f : (a.in_act == COMA_ACT_RELU ? 0.f : (a.in_slope ? __ldg(a.in_slope) : 0.f));
    for (int kd = 0; kd < K; ++kd) {
      int id;
      if (!a.transposed) {
        id = od * a.stride + kd - a.pad;
      } else {
        const int t = od + a.pad - kd;
        if (t < 0 || (t % a.stride) != 0) continue;
        id = t / a.stride;
      }
      if (id < 0 || id >= a.Di) continue;
      for (int kh = 0; kh < K; ++kh) {
        int ih;
        if (!a.transposed) {
          ih = oh * a.stride + kh - a.pad;
        } else {
          const int t = oh + a.pad - kh;
          if (t < 0 || (t % a.stride) != 0) continue;
          ih = t / a.stride;
        }
        if (ih < 0 || ih >= a.Hi) continue;
        for (int kw = 0; kw < K; ++kw) {
          int iw;
          if (!a.transposed) {
            iw = ow * a.stride + kw - a.pad;
          } else {
            const int t = ow + a.pad - kw;
            if (t < 0 || (t % a.stride) != 0) continue;
            iw = t / a.stride;
          }
          if (iw < 0 || iw >= a.Wi) continue;
          const int tap = (kd * K + kh) * K + kw;
          const T* xp = xb + (((int64_t)id * a.Hi + ih) * a.Wi + iw) * a.x_cs;
          const T* wp = wb + ((int64_t)tap * a.Cout + co0) * a.Cin;
          if (VEC == 8) {
            for (int ci = 0; ci < a.Cin; ci += 8) {
              float xv[8];
              load8(xp + ci, xv);
              if (isc) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float u = fmaf(__ldg(isc + ci + e), xv[e], __ldg(ish + ci + e));
                  xv[e] = fmaf(ineg, fminf(u, 0.f), fmaxf(u, 0.f));
                }
              }
#pragma unroll
              for (int j = 0; j < COT; ++j) {
                if (co0 + j < a.Cout) {
                  float wv[8];
                  load8(wp + (int64_t)j * a.Cin + ci, wv);
#pragma unroll
                  for (int e = 0; e < 8; ++e) acc[j] = fmaf(xv[e], wv[e], acc[j]);
                }
              }
            }
          } else {
            for (int ci = 0; ci < a.Cin; ++ci) {
              float xs = Elem<T>::ld(xp + ci);
              if (isc) {
                const float u = fmaf(__ldg(isc + ci), xs, __ldg(ish + ci));
                xs = fmaf(ineg, fminf(u, 0.f), fmaxf(u, 0.f));
              }
#pragma unroll
              for (int j = 0; j < COT; ++j)
                if (co0 + j < a.Cout) acc[j] = fmaf(xs, Elem<T>::ld(wp + (int64_t)j * a.Cin + ci), acc[j]);
            }
          }
        }
      }
    }
  }

  // ---- epilogue: bias, statistics of the raw output, per-(sample,channel) affine + activation ----
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  __shared__ float red[kSimtThreads / 32][COT][2];
  T* yp = static_cast<T*>(a.y) + ((int64_t)b * Vo + v) * a.y_cs + a.y_co;
#pragma unroll
  for (int j = 0; j < COT; ++j) {
    const int co = co0 + j;
    float val = 0.f;
    if (co < a.Cout && valid) {
      val = acc[j] + (a.bias ? __ldg(a.bias + (int64_t)b * a.bias_bstride + co) : 0.f);
      if (co < a.y_cn) {
        float u = val;
        if (a.scale) u = fmaf(__ldg(a.scale + (int64_t)b * a.Cout + co), val, __ldg(a.shift + (int64_t)b * a.Cout + co));
        Elem<T>::st(yp + co, act_fwd(a.act, u, slope));
      }
    }
    if (a.stats) {
      const float s1 = warp_sum(val), s2 = warp_sum(val * val);
      if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5][j][0] = s1;
        red[threadIdx.x >> 5][j][1] = s2;
      }
    }
  }
  if (a.stats) {
    __syncthreads();
    if (threadIdx.x < COT * 2) {
      const int j = threadIdx.x >> 1, q = threadIdx.x & 1, co = co0 + j;
      if (co < a.Cout) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kSimtThreads / 32; ++w) s += red[w][j][q];
        a.stats[(((int64_t)b * chunks + chunk) * a.Cout + co) * 2 + q] = s;
      }
    }
  }
}

static int simt_chunks(const coma_conv_args& a) {
  const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
  return (int)((Vo + kSimtThreads - 1) / kSimtThreads);
}

template <typename T>
static int launch_simt_t(const coma_conv_args& a, cudaStream_t stream) {
  const int chunks = simt_chunks(a);
  const bool vec = (a.Cin % 8 == 0) && (a.x_cs % 8 == 0) && (a.x_co % 8 == 0) && (a.w_bstride % 8 == 0) &&
                   (reinterpret_cast<uintptr_t>(a.x) % 32 == 0) && (reinterpret_cast<uintptr_t>(a.w) % 32 == 0);
  const bool wide = a.Cout >= 8;
  dim3 grid((unsigned)(a.B * chunks), (unsigned)((a.Cout + (wide ? 8 : 1) - 1) / (wide ? 8 : 1)));
  if (wide && vec) conv_simt_kernel<T, 8, 8><<<grid, kSimtThreads, 0, stream>>>(a, chunks);
  else if (wide) conv_simt_kernel<T, 8, 1><<<grid, kSimtThreads, 0, stream>>>(a, chunks);
  else if (vec) conv_simt_kernel<T, 1, 8><<<grid, kSimtThreads, 0, stream>>>(a, chunks);
  else conv_simt_kernel<T, 1, 1><<<grid, kSimtThreads, 0, stream>>>(a, chunks);
  COMA_CHECK_LAUNCH("conv_simt");
  return COMA_OK;
}


// ------------------------------------------------------------------------------------------------
// Pointwise heads with one output channel (reduce_channels 32 -> 1 with per-sample expert-mixed weights, the 2 -> 1
// final_pred_head, the projection heads): pure HBM streaming, so one thread walks many voxels with the weights in
// registers and all loads of a voxel group in flight, instead of the generic one-voxel-per-thread gather above.
// ------------------------------------------------------------------------------------------------
constexpr int kPwThreads = 256;

template <typename T, int CIN, bool VEC>
__global__ void __launch_bounds__(kPwThreads) conv_pw1_kernel(coma_conv_args a, int chunks) {
  // VEC: LPV = CIN/8 lanes share a voxel (16 bytes each), so a warp load covers 512 contiguous bytes; the partial dot
  // products meet through shuffles.  Scalar path (tiny CIN): one lane per voxel.
  constexpr int LPV = VEC ? CIN / 8 : 1, CPL = VEC ? 8 : CIN;       // lanes per voxel, channels per lane
  constexpr int U = (VEC || CIN <= 4) ? 4 : 1;                      // voxels in flight per lane
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
  const int64_t per = (Vo + chunks - 1) / chunks;
  const int64_t v0 = (int64_t)chunk * per, v1 = min(Vo, v0 + per);
  const int part = threadIdx.x % LPV, vl = threadIdx.x / LPV;
  constexpr int VPB = kPwThreads / LPV;                              // voxels per block and pass
  const T* xb = static_cast<const T*>(a.x) + (int64_t)b * Vo * a.x_cs + a.x_co + part * CPL;
  const T* wb = static_cast<const T*>(a.w) + (int64_t)b * a.w_bstride + part * CPL;
  T* yb = static_cast<T*>(a.y) + (int64_t)b * Vo * a.y_cs + a.y_co;
  float w[CPL], isc[CPL], ish[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    w[c] = Elem<T>::ld(wb + c);
    isc[c] = a.in_scale ? __ldg(a.in_scale + (int64_t)b * a.Cin + part * CPL + c) : 1.f;
    ish[c] = a.in_scale ? __ldg(a.in_shift + (int64_t)b * a.Cin + part * CPL + c) : 0.f;
  }
  const bool pro = a.in_scale != nullptr;
  const float ineg = a.in_act == COMA_ACT_NONE ? 1.f : (a.in_act == COMA_ACT_RELU ? 0.f : (a.in_slope ? __ldg(a.in_slope) : 0.f));
  const float bias = a.bias ? __ldg(a.bias + (int64_t)b * a.bias_bstride) : 0.f;
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  const float sc = a.scale ? __ldg(a.scale + (int64_t)b * a.Cout) : 1.f, sh = a.scale ? __ldg(a.shift + (int64_t)b * a.Cout) : 0.f;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t vb = v0 + vl; vb - vl < v1; vb += (int64_t)VPB * U) {      // warp-uniform trip count (shuffles inside)
    float xv[U][CPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = vb + (int64_t)u * VPB;
#pragma unroll
      for (int c = 0; c < CPL; ++c) xv[u][c] = 0.f;
      if (v < v1) {
        const T* xp = xb + v * a.x_cs;
        if (VEC) {
          float t8[8];
          load8(xp, t8);
#pragma unroll
          for (int c = 0; c < CPL; ++c) xv[u][c] = t8[c % 8];
        } else {
#pragma unroll
          for (int c = 0; c < CPL; ++c) xv[u][c] = Elem<T>::ld(xp + c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = vb + (int64_t)u * VPB;
      float acc = 0.f;
      if (pro) {        // producer's norm + FiLM + activation applied on load
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const float t = fmaf(isc[c], xv[u][c], ish[c]);
          acc = fmaf(fmaf(ineg, fminf(t, 0.f), fmaxf(t, 0.f)), w[c], acc);
        }
      } else {
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc = fmaf(xv[u][c], w[c], acc);
      }
#pragma unroll
      for (int o = 1; o < LPV; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (v < v1 && part == 0) {
        const float val = acc + bias;
        s1 += val;
        s2 = fmaf(val, val, s2);
        Elem<T>::st(yb + v * a.y_cs, act_fwd(a.act, fmaf(sc, val, sh), slope));
      }
    }
  }
  if (a.stats) {
    __shared__ float red[kPwThreads / 32][2];
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s1; red[threadIdx.x >> 5][1] = s2; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float t = 0.f;
#pragma unroll
      for (int wi = 0; wi < kPwThreads / 32; ++wi) t += red[wi][threadIdx.x];
      a.stats[((int64_t)b * chunks + chunk) * 2 + threadIdx.x] = t;
    }
  }
}


// Pointwise conv from ONE input channel (the data gradient of the one-channel heads: dx[v][c] = dy[v] * w[c], per-sample
// weights for the expert-mixed reduce_channels): an outer product, pure HBM write traffic.  One thread = one voxel x 8 channels.
template <typename T>
__global__ void __launch_bounds__(256) conv_pw_from1_kernel(coma_conv_args a) {
  const int b = blockIdx.y;
  const int CV = a.Cout >> 3, cvec = threadIdx.x % CV, vlane = threadIdx.x / CV, lanes = 256 / CV;
  const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
  const T* wb = static_cast<const T*>(a.w) + (int64_t)b * a.w_bstride + cvec * 8;      // packed w[tap=0][Cout][Cin=1]
  float w8[8], s8[8], h8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int co = cvec * 8 + e;
    const float bias = a.bias ? __ldg(a.bias + (int64_t)b * a.bias_bstride + co) : 0.f;
    const float sc = a.scale ? __ldg(a.scale + (int64_t)b * a.Cout + co) : 1.f;
    w8[e] = Elem<T>::ld(wb + e) * sc;
    h8[e] = fmaf(sc, bias, a.scale ? __ldg(a.shift + (int64_t)b * a.Cout + co) : 0.f);
    s8[e] = 0.f;
  }
  (void)s8;
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  const T* xb = static_cast<const T*>(a.x) + (int64_t)b * Vo * a.x_cs + a.x_co;
  T* yb = static_cast<T*>(a.y) + (int64_t)b * Vo * a.y_cs + a.y_co + cvec * 8;
  const int64_t step = (int64_t)gridDim.x * lanes;
  int64_t v = (int64_t)blockIdx.x * lanes + vlane;
  for (; v + 3 * step < Vo; v += 4 * step) {          // four voxels in flight per thread: the loads first, then the four stores
    float xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xv[u] = Elem<T>::ld(xb + (v + u * step) * a.x_cs);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = act_fwd(a.act, fmaf(xv[u], w8[e], h8[e]), slope);
      store8(yb + (v + u * step) * a.y_cs, o);
    }
  }
  for (; v < Vo; v += step) {
    const float xv = Elem<T>::ld(xb + v * a.x_cs);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = act_fwd(a.act, fmaf(xv, w8[e], h8[e]), slope);
    store8(yb + v * a.y_cs, o);
  }
}

static bool pw_from1_applicable(const coma_conv_args& a) {
  const int esz = a.dtype == COMA_BF16 ? 2 : 4;
  return a.ksize == 1 && a.stride == 1 && !a.transposed && a.Cin == 1 && a.Cout % 8 == 0 && a.Cout <= 512 && 256 % (a.Cout / 8) == 0 &&
         a.y_cn == a.Cout && !a.stats && !a.in_scale && a.y_cs % 8 == 0 && a.y_co % 8 == 0 &&
         reinterpret_cast<uintptr_t>(a.y) % (8 * esz > 16 ? 16 : 8 * esz) == 0;
}

static bool pw1_applicable(const coma_conv_args& a) {
  return a.ksize == 1 && a.stride == 1 && !a.transposed && a.Cout == 1 && a.y_cn >= 1 &&
         (a.Cin == 1 || a.Cin == 2 || a.Cin == 3 || a.Cin == 8 || a.Cin == 16 || a.Cin == 32 || a.Cin == 64);
}
static int pw1_chunks(const coma_conv_args& a) {
  const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
  const int64_t by_work = (Vo + kPwThreads * 8 - 1) / (kPwThreads * 8);
  const int64_t by_sms = (8 * (int64_t)num_sms() + a.B - 1) / a.B;
  return (int)std::max<int64_t>(1, std::min(by_work, by_sms));
}

template <typename T>
static int launch_pw1_t(const coma_conv_args& a, cudaStream_t stream) {
  const int chunks = pw1_chunks(a);
  const bool vec = (a.Cin % 8 == 0) && (a.x_cs % 8 == 0) && (a.x_co % 8 == 0) && (reinterpret_cast<uintptr_t>(a.x) % 32 == 0);
  dim3 grid((unsigned)chunks, (unsigned)a.B);
#define COMA_PW_CASE(C)                                                                                 \
  if (a.Cin == C) {                                                                                     \
    if (vec && C % 8 == 0) conv_pw1_kernel<T, C, (C % 8 == 0)><<<grid, kPwThreads, 0, stream>>>(a, chunks); \
    else conv_pw1_kernel<T, C, false><<<grid, kPwThreads, 0, stream>>>(a, chunks);                      \
  }
  COMA_PW_CASE(1) COMA_PW_CASE(2) COMA_PW_CASE(3) COMA_PW_CASE(8) COMA_PW_CASE(16) COMA_PW_CASE(32) COMA_PW_CASE(64)
#undef COMA_PW_CASE
  COMA_CHECK_LAUNCH("conv_pw1");
  return COMA_OK;
}


// ------------------------------------------------------------------------------------------------
// Few-channel pointwise convolutions (the gate's W_g / W_x 1x1x1 convs in training mode and their data gradients, Cin, Cout <= 32):
// 0.4 .. 0.8 GB of traffic and ~0.5 kFMA per voxel -- HBM streaming, not a GEMM.  On the 128-voxel-tile tcgen05 kernel they ran at
// 15 % of the HBM peak (one K = 16/32 MMA per tile cannot hide the per-tile TMA / TMEM / epilogue round trip).  Here a thread
// owns whole voxels: CIN values in registers, fp32 weights broadcast from shared memory, COUT accumulators, the same fused
// prologue / epilogue (input affine + activation, bias, statistics partials, output affine + activation) as the other kernels.
// ------------------------------------------------------------------------------------------------
constexpr int kPwgThreads = 256, kPwgVox = 8;      // voxels per thread; a block (= one statistics chunk) covers 2048 voxels

template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(kPwgThreads) conv_pwg_kernel(coma_conv_args a, int chunks) {
  __shared__ __align__(16) float ws[CIN][COUT];
  __shared__ float red[kPwgThreads / 32][COUT][2];
  __shared__ float cb[COUT], cs[COUT], ch[COUT];      // bias, output scale / shift of this sample (defaults 0, 1, 0)
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
  const T* wb = static_cast<const T*>(a.w) + (int64_t)b * a.w_bstride;
  for (int i = threadIdx.x; i < CIN * COUT; i += kPwgThreads) ws[i % CIN][i / CIN] = Elem<T>::ld(wb + i);     // w[co][ci] -> ws[ci][co]
  __shared__ float isc[CIN], ish[CIN];               // input prologue coefficients (broadcast reads; registers go to the accumulators)
  const bool pro = a.in_scale != nullptr;
  const float ineg = a.in_act == COMA_ACT_NONE ? 1.f : (a.in_act == COMA_ACT_RELU ? 0.f : (a.in_slope ? __ldg(a.in_slope) : 0.f));
  if (pro && threadIdx.x < CIN) {
    isc[threadIdx.x] = __ldg(a.in_scale + (int64_t)b * CIN + threadIdx.x);
    ish[threadIdx.x] = __ldg(a.in_shift + (int64_t)b * CIN + threadIdx.x);
  }
  if (threadIdx.x < COUT) {
    cb[threadIdx.x] = a.bias ? __ldg(a.bias + (int64_t)b * a.bias_bstride + threadIdx.x) : 0.f;
    cs[threadIdx.x] = a.scale ? __ldg(a.scale + (int64_t)b * COUT + threadIdx.x) : 1.f;
    ch[threadIdx.x] = a.scale ? __ldg(a.shift + (int64_t)b * COUT + threadIdx.x) : 0.f;
  }
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  __syncthreads();
  const T* xb = static_cast<const T*>(a.x) + (int64_t)b * Vo * a.x_cs + a.x_co;
  T* yb = static_cast<T*>(a.y) + (int64_t)b * Vo * a.y_cs + a.y_co;
  float s1[COUT], s2[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) s1[j] = s2[j] = 0.f;
  const bool simple = a.act == COMA_ACT_NONE || a.act == COMA_ACT_RELU || a.act == COMA_ACT_LEAKY;      // act(u) = max(u, 0) + aneg * min(u, 0)
  const float aneg = a.act == COMA_ACT_NONE ? 1.f : (a.act == COMA_ACT_RELU ? 0.f : slope);
  const int64_t v0 = (int64_t)chunk * (kPwgThreads * kPwgVox);
#pragma unroll 2
  for (int it = 0; it < kPwgVox; ++it) {
    const int64_t v = v0 + it * kPwgThreads + threadIdx.x;
    if (v - (threadIdx.x & 31) >= Vo) break;                        // warp-uniform: the staged store below is warp-collective
    const bool live = v < Vo;
    float xv[CIN];
#pragma unroll
    for (int c = 0; c < CIN; c += 8) load8(xb + (live ? v : Vo - 1) * a.x_cs + c, *reinterpret_cast<float(*)[8]>(&xv[c]));
    if (pro) {
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float u = fmaf(isc[c], xv[c], ish[c]);
        xv[c] = fmaf(ineg, fminf(u, 0.f), fmaxf(u, 0.f));
      }
    }
    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int j = 0; j < COUT; j += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(&ws[c][j]);
        acc[j] = fmaf(xv[c], w4.x, acc[j]);
        acc[j + 1] = fmaf(xv[c], w4.y, acc[j + 1]);
        acc[j + 2] = fmaf(xv[c], w4.z, acc[j + 2]);
        acc[j + 3] = fmaf(xv[c], w4.w, acc[j + 3]);
      }
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
      const float val = live ? acc[j] + cb[j] : 0.f;
      s1[j] += val;
      s2[j] = fmaf(val, val, s2[j]);
      const float u = fmaf(cs[j], val, ch[j]);
      acc[j] = simple ? fmaf(aneg, fminf(u, 0.f), fmaxf(u, 0.f)) : act_fwd(a.act, u, slope);
    }
#pragma unroll
    for (int j = 0; j < COUT; j += 8)
      if (live) store8(yb + v * a.y_cs + j, *reinterpret_cast<const float(*)[8]>(&acc[j]));
  }
  if (a.stats) {
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
      const float t1 = warp_sum(s1[j]), t2 = warp_sum(s2[j]);
      if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][j][0] = t1; red[threadIdx.x >> 5][j][1] = t2; }
    }
    __syncthreads();
    if (threadIdx.x < COUT * 2) {
      const int j = threadIdx.x >> 1, q = threadIdx.x & 1;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kPwgThreads / 32; ++w) t += red[w][j][q];
      a.stats[(((int64_t)b * chunks + chunk) * COUT + j) * 2 + q] = t;
    }
  }
}


// ------------------------------------------------------------------------------------------------
// The same few-channel pointwise convolutions on the warp-level tensor-core path, for the plain cases (bf16, no output affine /
// activation / input prologue: the gate's W_g / W_x convs before their BatchNorm, and their data gradients): the CUDA-core kernel
// above is instruction-bound (0.5 kFMA per voxel), while as m16n8k16 tiles the whole problem is four to eight mma.sync per 16
// voxels and the A fragments are plain 4-byte loads straight from the NDHWC rows (row = voxel, two consecutive channels per register)
// -- no shared-memory staging at all; the weights sit in registers as B fragments.  HBM streaming is what is left.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_pw_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(kPwgThreads) conv_pwm_kernel(coma_conv_args a, int chunks) {
  constexpr int KS = CIN / 16, NT = COUT / 8, TILES = kPwgVox * 32 / 16;      // 16-voxel tiles per warp and chunk
  __shared__ float red[kPwgThreads / 32][COUT][2];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
  const __nv_bfloat16* wb = static_cast<const __nv_bfloat16*>(a.w) + (int64_t)b * a.w_bstride;
  uint32_t bf[KS][NT][2];                       // B fragments: w[co = n0 + g][ci = k0 + 2t (+8)], two channels per register
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const __nv_bfloat16* wp = wb + (int64_t)(nt * 8 + g) * CIN + ks * 16 + 2 * t;
      bf[ks][nt][0] = __ldg(reinterpret_cast<const uint32_t*>(wp));
      bf[ks][nt][1] = __ldg(reinterpret_cast<const uint32_t*>(wp + 8));
    }
  float bias[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) bias[nt][j] = a.bias ? __ldg(a.bias + (int64_t)b * a.bias_bstride + nt * 8 + 2 * t + j) : 0.f;
  float s1[NT][2], s2[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f;

  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(a.x) + (int64_t)b * Vo * a.x_cs + a.x_co + 2 * t;
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(a.y) + (int64_t)b * Vo * a.y_cs + a.y_co + 2 * t;
  const int64_t w0 = (int64_t)chunk * (kPwgThreads * kPwgVox) + (int64_t)warp * (TILES * 16);
#pragma unroll 2
  for (int tile = 0; tile < TILES; ++tile) {
    const int64_t v0 = w0 + tile * 16;
    if (v0 >= Vo) break;
    const int64_t r0 = v0 + g, r1 = v0 + g + 8;
    const bool ok0 = r0 < Vo, ok1 = r1 < Vo;
    const __nv_bfloat16* x0 = xb + (ok0 ? r0 : Vo - 1) * a.x_cs;
    const __nv_bfloat16* x1 = xb + (ok1 ? r1 : Vo - 1) * a.x_cs;
    uint32_t af[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      af[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(x0 + ks * 16));
      af[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(x1 + ks * 16));
      af[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(x0 + ks * 16 + 8));
      af[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(x1 + ks * 16 + 8));
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) mma_pw_bf16(acc, af[ks], bf[ks][nt][0], bf[ks][nt][1]);
      const float v00 = acc[0] + bias[nt][0], v01 = acc[1] + bias[nt][1], v10 = acc[2] + bias[nt][0], v11 = acc[3] + bias[nt][1];
      if (ok0) {
        s1[nt][0] += v00; s1[nt][1] += v01; s2[nt][0] = fmaf(v00, v00, s2[nt][0]); s2[nt][1] = fmaf(v01, v01, s2[nt][1]);
        *reinterpret_cast<uint32_t*>(yb + r0 * a.y_cs + nt * 8) = pack_bf16x2(v00, v01);
      }
      if (ok1) {
        s1[nt][0] += v10; s1[nt][1] += v11; s2[nt][0] = fmaf(v10, v10, s2[nt][0]); s2[nt][1] = fmaf(v11, v11, s2[nt][1]);
        *reinterpret_cast<uint32_t*>(yb + r1 * a.y_cs + nt * 8) = pack_bf16x2(v10, v11);
      }
    }
  }
  if (a.stats) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float u1 = s1[nt][j], u2 = s2[nt][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { u1 += __shfl_xor_sync(0xffffffffu, u1, o); u2 += __shfl_xor_sync(0xffffffffu, u2, o); }
        if (g == 0) { red[warp][nt * 8 + 2 * t + j][0] = u1; red[warp][nt * 8 + 2 * t + j][1] = u2; }
      }
    __syncthreads();
    if (threadIdx.x < COUT * 2) {
      const int j = threadIdx.x >> 1, q = threadIdx.x & 1;
      float u = 0.f;
#pragma unroll
      for (int w = 0; w < kPwgThreads / 32; ++w) u += red[w][j][q];
      a.stats[(((int64_t)b * chunks + chunk) * COUT + j) * 2 + q] = u;
    }
  }
}

static bool pwm_applicable(const coma_conv_args& a) {
  static const bool off = [] { const char* e = getenv("COMA_DISABLE_PWM"); return e && e[0] == '1'; }();
  // 16 / 32 channels: the top-level gate convs (W_g, W_x: 32 -> 16) and their data gradients; 64 <-> 32: the same at the second level
  // (on the 128-voxel-tile tcgen05 kernel these streamed at 30-46 "TFLOP/s", i.e. a fifth of the HBM rate)
  const bool small = (a.Cin == 16 || a.Cin == 32) && (a.Cout == 16 || a.Cout == 32);
  const bool level2 = (a.Cin == 64 && a.Cout == 32) || (a.Cin == 32 && a.Cout == 64);
  return !off && a.dtype == COMA_BF16 && a.ksize == 1 && a.stride == 1 && !a.transposed && (small || level2) && a.y_cn == a.Cout && !a.scale && !a.in_scale && a.act == COMA_ACT_NONE &&
         a.x_cs % 2 == 0 && a.x_co % 2 == 0 && a.y_cs % 2 == 0 && a.y_co % 2 == 0 && a.w_bstride % 2 == 0 &&
         reinterpret_cast<uintptr_t>(a.x) % 4 == 0 && reinterpret_cast<uintptr_t>(a.y) % 4 == 0 && reinterpret_cast<uintptr_t>(a.w) % 4 == 0;
}
static int launch_pwm(const coma_conv_args& a, cudaStream_t stream) {
  const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
  const int chunks = (int)((Vo + kPwgThreads * kPwgVox - 1) / (kPwgThreads * kPwgVox));
  dim3 grid((unsigned)chunks, (unsigned)a.B);
  if (a.Cin == 16 && a.Cout == 16) conv_pwm_kernel<16, 16><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  else if (a.Cin == 32 && a.Cout == 16) conv_pwm_kernel<32, 16><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  else if (a.Cin == 16 && a.Cout == 32) conv_pwm_kernel<16, 32><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  else if (a.Cin == 64 && a.Cout == 32) conv_pwm_kernel<64, 32><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  else if (a.Cin == 32 && a.Cout == 64) conv_pwm_kernel<32, 64><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  else conv_pwm_kernel<32, 32><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  COMA_CHECK_LAUNCH("conv_pwm");
  return COMA_OK;
}

static bool pwg_applicable(const coma_conv_args& a) {
  static const bool off = [] { const char* e = getenv("COMA_DISABLE_PWG"); return e && e[0] == '1'; }();
  const int esz = a.dtype == COMA_BF16 ? 2 : 4;
  const uintptr_t al = 8 * esz > 16 ? 16 : 8 * esz;
  return !off && a.ksize == 1 && a.stride == 1 && !a.transposed && (a.Cin == 16 || a.Cin == 32) && a.Cout == 16 &&          // measured: 16 -> 32 is faster on the tcgen05 tile kernel (0.49 vs 0.78 ms)
         a.y_cn == a.Cout && a.x_cs % 8 == 0 && a.x_co % 8 == 0 && a.y_cs % 8 == 0 && a.y_co % 8 == 0 &&
         reinterpret_cast<uintptr_t>(a.x) % al == 0 && reinterpret_cast<uintptr_t>(a.y) % al == 0;
}
static int pwg_chunks(const coma_conv_args& a) {
  const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
  return (int)((Vo + kPwgThreads * kPwgVox - 1) / (kPwgThreads * kPwgVox));
}
template <typename T>
static int launch_pwg_t(const coma_conv_args& a, cudaStream_t stream) {
  const int chunks = pwg_chunks(a);
  dim3 grid((unsigned)chunks, (unsigned)a.B);
  if (a.Cin == 16 && a.Cout == 16) conv_pwg_kernel<T, 16, 16><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  else if (a.Cin == 32 && a.Cout == 16) conv_pwg_kernel<T, 32, 16><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  else if (a.Cin == 16 && a.Cout == 32) conv_pwg_kernel<T, 16, 32><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  else conv_pwg_kernel<T, 32, 32><<<grid, kPwgThreads, 0, stream>>>(a, chunks);
  COMA_CHECK_LAUNCH("conv_pwg");
  return COMA_OK;
}
// the few-channel pointwise problems are HBM streaming: IMPL_AUTO prefers this kernel over the tcgen05 tile kernel (api.cu)
bool conv_simt_preferred(const coma_conv_args& a) { return pwm_applicable(a) || pwg_applicable(a); }

bool conv_simt_prologue_fused(const coma_conv_args& a) { return pw1_applicable(a) || pwg_applicable(a); }   // (pwm takes no prologue)

int conv_simt_stat_chunks(const coma_conv_args& a) { return pw1_applicable(a) ? pw1_chunks(a) : ((pwm_applicable(a) || pwg_applicable(a)) ? pwg_chunks(a) : simt_chunks(a)); }

int conv_simt_launch(const coma_conv_args& a, cudaStream_t stream) {
  if (pw_from1_applicable(a)) {
    const int64_t Vo = (int64_t)a.Do * a.Ho * a.Wo;
    const int lanes = 256 / (a.Cout / 8);
    const int64_t want = (Vo + lanes * 4 - 1) / (lanes * 4), cap = std::max<int64_t>(1, (int64_t)num_sms() * 8 / a.B);
    dim3 grid((unsigned)std::max<int64_t>(1, std::min(want, cap)), (unsigned)a.B);
    if (a.dtype == COMA_BF16) conv_pw_from1_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a);
    else conv_pw_from1_kernel<float><<<grid, 256, 0, stream>>>(a);
    COMA_CHECK_LAUNCH("conv_pw_from1");
    return COMA_OK;
  }
  if (pw1_applicable(a)) return a.dtype == COMA_BF16 ? launch_pw1_t<__nv_bfloat16>(a, stream) : launch_pw1_t<float>(a, stream);
  if (pwm_applicable(a)) return launch_pwm(a, stream);
  if (pwg_applicable(a)) return a.dtype == COMA_BF16 ? launch_pwg_t<__nv_bfloat16>(a, stream) : launch_pwg_t<float>(a, stream);
  if (a.dtype == COMA_BF16) return launch_simt_t<__nv_bfloat16>(a, stream);
  return launch_simt_t<float>(a, stream);
}

// ------------------------------------------------------------------------------------------------
// Weight gradient: dw[tap][cg][cx] += sum_o g[o][cg] * x[o*stride + k - pad][cx]   (split over voxels)
// ------------------------------------------------------------------------------------------------
constexpr int kWgTile = 64, kWgVox = 16;

template <typename T>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(coma_wgrad_args a, int64_t vchunk, int cx_tiles) {
  __shared__ float sg[kWgVox][kWgTile + 4];
  __shared__ float sx[kWgVox][kWgTile + 4];
  const int K = a.ksize;
  const int tap = blockIdx.y;
  const int kd = tap / (K * K), kh = (tap / K) % K, kw = tap % K;
  const int cg0 = (blockIdx.z / cx_tiles) * kWgTile, cx0 = (blockIdx.z % cx_tiles) * kWgTile;
  const int64_t Vg = (int64_t)a.Dg * a.Hg * a.Wg, total = (int64_t)a.B * Vg;
  const int64_t begin = (int64_t)blockIdx.x * vchunk, end = min(begin + vchunk, total);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const T* gp = static_cast<const T*>(a.g);
  const T* xp = static_cast<const T*>(a.x);
  for (int64_t base = begin; base < end; base += kWgVox) {
    // 16 voxels x 64 channels per operand, 4 elements per thread each
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = r * 256 + threadIdx.x;
      const int vv = idx >> 6, c = idx & 63;
      const int64_t o = base + vv;
      float gval = 0.f, xval = 0.f;
      if (o < end) {
        if (cg0 + c < a.Cg) gval = Elem<T>::ld(gp + o * a.g_cs + a.g_co + cg0 + c);
        if (cx0 + c < a.Cx) {
          const int64_t bb = o / Vg, rem = o % Vg;
          const int ow = (int)(rem % a.Wg), oh = (int)((rem / a.Wg) % a.Hg), od = (int)(rem / ((int64_t)a.Wg * a.Hg));
          const int id = od * a.stride + kd - a.pad, ih = oh * a.stride + kh - a.pad, iw = ow * a.stride + kw - a.pad;
          if (id >= 0 && id < a.Dx && ih >= 0 && ih < a.Hx && iw >= 0 && iw < a.Wx)
            xval = Elem<T>::ld(xp + (((bb * a.Dx + id) * a.Hx + ih) * a.Wx + iw) * a.x_cs + a.x_co + cx0 + c);
        }
      }
      sg[vv][c] = gval;
      sx[vv][c] = xval;
    }
    __syncthreads();
#pragma unroll
    for (int vv = 0; vv < kWgVox; ++vv) {
      float gv[4], xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) gv[i] = sg[vv][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = sx[vv][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], xv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cg = cg0 + ty * 4 + i;
    if (cg >= a.Cg) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cx = cx0 + tx * 4 + j;
      if (cx < a.Cx) atomicAdd(a.dw + ((int64_t)tap * a.Cg + cg) * a.Cx + cx, acc[i][j]);
    }
  }
}

// Cg == 1 (gradients of the 1-channel heads: psi, 16->1 / 8->1 modulator convs, projection heads): a per-voxel
// scaled accumulation of x, one tap per blockIdx.y; HBM-bound (reads g and the shifted x once per tap).
template <typename T>
__global__ void __launch_bounds__(256) wgrad_cg1_kernel(coma_wgrad_args a, int64_t vchunk) {
  constexpr int CXMAX = 64;
  __shared__ float red[8][CXMAX];
  const int K = a.ksize;
  const int tap = blockIdx.y;
  const int kd = tap / (K * K), kh = (tap / K) % K, kw = tap % K;
  const int64_t Vg = (int64_t)a.Dg * a.Hg * a.Wg, total = (int64_t)a.B * Vg;
  const int64_t begin = (int64_t)blockIdx.x * vchunk, end = min(begin + vchunk, total);
  const T* gp = static_cast<const T*>(a.g) + a.g_co;
  const T* xp = static_cast<const T*>(a.x) + a.x_co;
  const bool vec = (a.Cx % 8 == 0) && (a.x_cs % 8 == 0) && (a.x_co % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
  float acc[CXMAX];
#pragma unroll
  for (int c = 0; c < CXMAX; ++c) acc[c] = 0.f;
  for (int64_t o = begin + threadIdx.x; o < end; o += 256) {
    const float gv = Elem<T>::ld(gp + o * a.g_cs);
    const int64_t bb = o / Vg, rem = o - bb * Vg;
    const int ow = (int)(rem % a.Wg), t2 = (int)(rem / a.Wg);
    const int oh = t2 % a.Hg, od = t2 / a.Hg;
    const int id = od * a.stride + kd - a.pad, ih = oh * a.stride + kh - a.pad, iw = ow * a.stride + kw - a.pad;
    if (id < 0 || id >= a.Dx || ih < 0 || ih >= a.Hx || iw < 0 || iw >= a.Wx) continue;
    const T* xr = xp + (((bb * a.Dx + id) * a.Hx + ih) * a.Wx + iw) * a.x_cs;
    if (vec) {
#pragma unroll
      for (int c = 0; c < CXMAX; c += 8) {
        if (c < a.Cx) {
          float xv[8];
          load8(xr + c, xv);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[c + e] = fmaf(gv, xv[e], acc[c + e]);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < CXMAX; ++c)
        if (c < a.Cx) acc[c] = fmaf(gv, Elem<T>::ld(xr + c), acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < CXMAX; ++c) {
    if (c < a.Cx) {
      const float s = warp_sum(acc[c]);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < a.Cx) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(a.dw + (int64_t)tap * a.Cx + threadIdx.x, s);
  }
}


// Cg == 1, 1x1x1 (psi of the four gates, the 1-output pointwise heads): dw[c] = sum_v g[v] * x[v][c], one streaming pass over x
// and g.  A thread owns one 8-channel vector lane and every (256 / CV)-th voxel, four voxels in flight; no per-voxel index
// arithmetic (wgrad_cg1_kernel decodes (b, d, h, w) with 64-bit divisions per voxel: 215 us for 0.3 GB at batch 4 x 128^3).
template <typename T>
__global__ void __launch_bounds__(256) wgrad_cg1_k1_kernel(coma_wgrad_args a, int64_t vchunk) {
  __shared__ float red[8][64];
  const int CV = a.Cx >> 3, cvec = threadIdx.x % CV, vlane = threadIdx.x / CV, lanes = 256 / CV;
  const int64_t total = (int64_t)a.B * a.Dg * a.Hg * a.Wg;
  const int64_t begin = (int64_t)blockIdx.x * vchunk, end = min(begin + vchunk, total);
  const T* gp = static_cast<const T*>(a.g) + a.g_co;
  const T* xp = static_cast<const T*>(a.x) + a.x_co + cvec * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  int64_t o = begin + vlane;
  for (; o + 3 * lanes < end; o += 4 * lanes) {
    float xv[4][8], gv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      load8_stream(xp + (o + (int64_t)u * lanes) * a.x_cs, xv[u]);
      gv[u] = Elem<T>::ld(gp + (o + (int64_t)u * lanes) * a.g_cs);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(gv[u], xv[u][e], acc[e]);
  }
  for (; o < end; o += lanes) {
    float xv[8];
    load8_stream(xp + o * a.x_cs, xv);
    const float gv = Elem<T>::ld(gp + o * a.g_cs);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(gv, xv[e], acc[e]);
  }
  // lanes of a warp with the same cvec: xor strides CV .. 16 (CV is a power of two <= 8)
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float v = acc[e];
    for (int st = 16; st >= CV; st >>= 1) v += __shfl_xor_sync(0xffffffffu, v, st);
    if ((threadIdx.x & 31) < CV) red[threadIdx.x >> 5][(threadIdx.x & 31) * 8 + e] = v;
  }
  __syncthreads();
  if (threadIdx.x < a.Cx) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(a.dw + threadIdx.x, t);
  }
}

// The same for a dense x with 1, 2 or 4 channels (final_pred_head 2 -> 1): a 16-byte load covers 8 / CX voxels.
template <typename T, int CX>
__global__ void __launch_bounds__(256) wgrad_cg1_k1_few_kernel(coma_wgrad_args a, int64_t vchunk) {
  constexpr int VPT = 8 / CX;
  __shared__ float red[8][CX];
  const int64_t total = (int64_t)a.B * a.Dg * a.Hg * a.Wg;
  const int64_t begin = (int64_t)blockIdx.x * vchunk, end = min(begin + vchunk, total);     // vchunk is a multiple of 8
  const T* gp = static_cast<const T*>(a.g) + a.g_co;
  const T* xp = static_cast<const T*>(a.x);
  float acc[CX];
#pragma unroll
  for (int c = 0; c < CX; ++c) acc[c] = 0.f;
  int64_t o = begin + (int64_t)threadIdx.x * VPT;
  constexpr int64_t STEP = 256 * VPT;
  for (; o + 3 * STEP + VPT <= end; o += 4 * STEP) {
    float xv[4][8], gv[4][VPT];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      load8_stream(xp + (o + u * STEP) * CX, xv[u]);
#pragma unroll
      for (int j = 0; j < VPT; ++j) gv[u][j] = Elem<T>::ld(gp + (o + u * STEP + j) * a.g_cs);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < VPT; ++j)
#pragma unroll
        for (int c = 0; c < CX; ++c) acc[c] = fmaf(gv[u][j], xv[u][j * CX + c], acc[c]);
  }
  for (; o < end; o += STEP)
    for (int j = 0; j < VPT && o + j < end; ++j) {
      const float gv = Elem<T>::ld(gp + (o + j) * a.g_cs);
#pragma unroll
      for (int c = 0; c < CX; ++c) acc[c] = fmaf(gv, Elem<T>::ld(xp + (o + j) * CX + c), acc[c]);
    }
#pragma unroll
  for (int c = 0; c < CX; ++c) {
    const float v = warp_sum(acc[c]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < CX) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(a.dw + threadIdx.x, t);
  }
}

// One-channel gradient (the 16 -> 1 modulator heads), k3 s1: walk the x voxels ONCE per kd plane instead of once per tap.
// A thread owns one 8-channel vector of an x voxel and accumulates its contribution to the nine (kh,kw) taps of its kd:
// dw[kd,kh,kw][c] += x[i][c] * g[i - k + 1]; the 72 partial sums are reduced per block and added atomically.
template <typename T>
__global__ void __launch_bounds__(256) wgrad_cg1_k3s1_kernel(coma_wgrad_args a, int64_t vchunk) {
  // vchunk counts x LINES here: a block walks whole W lines, so the (b, d, h) decode and the validity of the three g lines
  // are per line, not per voxel (64-bit div / mod per voxel made the sweep instruction-bound: 1.3 ms for 0.3 GB)
  const int kd = blockIdx.y;
  const int CV = a.Cx >> 3, cvec = threadIdx.x % CV, vlane = threadIdx.x / CV, lanes = 256 / CV;
  const int64_t lines = (int64_t)a.B * a.Dx * a.Hx;
  const int64_t begin = (int64_t)blockIdx.x * vchunk, end = min(begin + vchunk, lines);
  const T* gp = static_cast<const T*>(a.g) + a.g_co;
  const T* xp = static_cast<const T*>(a.x) + a.x_co + cvec * 8;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[t][e] = 0.f;
  for (int64_t line = begin; line < end; ++line) {
    const int ih = (int)(line % a.Hx);
    const int64_t t2 = line / a.Hx;
    const int id = (int)(t2 % a.Dx);
    const int64_t bb = t2 / a.Dx;
    const int od = id - kd + 1;
    if (od < 0 || od >= a.Dg) continue;
    const T* xline = xp + line * a.Wx * a.x_cs;
    const T* gplane = gp + ((bb * a.Dg + od) * a.Hg) * (int64_t)a.Wg * a.g_cs;
    for (int iw = vlane; iw < a.Wx; iw += lanes) {
      float xv[8];
      load8(xline + (int64_t)iw * a.x_cs, xv);
      float gv[9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int oh = ih - kh + 1;
        const bool hok = oh >= 0 && oh < a.Hg;
        const T* grow = gplane + (int64_t)(hok ? oh : 0) * a.Wg * a.g_cs;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int ow = iw - kw + 1;
          const bool ok = hok && ow >= 0 && ow < a.Wg;
          gv[kh * 3 + kw] = ok ? Elem<T>::ld(grow + (int64_t)ow * a.g_cs) : 0.f;
        }
      }
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] = fmaf(gv[t], xv[e], acc[t][e]);
    }
  }
  // reduce over the voxel lanes that share a channel vector: lanes with equal (threadIdx.x % CV); CV divides 32
  __shared__ float red[8][9][8 * 8];        // [warp][tap][cvec * 8 + e], CV <= 8
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = acc[t][e];
      for (int o = 16; o >= CV; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) < CV) red[threadIdx.x >> 5][t][cvec * 8 + e] = v;
    }
  __syncthreads();
  for (int j = threadIdx.x; j < 9 * a.Cx; j += 256) {
    const int t = j / a.Cx, c = j % a.Cx;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][t][c];
    atomicAdd(a.dw + (int64_t)(kd * 9 + t) * a.Cx + c, sum);
  }
}

int wgrad_simt_launch(const coma_wgrad_args& a, cudaStream_t stream) {
  const int64_t total = (int64_t)a.B * a.Dg * a.Hg * a.Wg;
  if (a.Cg == 1 && a.ksize == 3 && a.stride == 1 && a.Cx % 8 == 0 && a.Cx <= 64 && 32 % (a.Cx / 8) == 0 && a.x_cs % 8 == 0 && a.x_co % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(a.x) & 15) == 0) {
    const int64_t totx = (int64_t)a.B * a.Dx * a.Hx;          // x lines: the kernel's work unit
    int64_t want = (int64_t)num_sms() * 8 / 3;
    int64_t vchunk = std::max<int64_t>((totx + want - 1) / want, 8);
    const int64_t nch = (totx + vchunk - 1) / vchunk;
    dim3 grid((unsigned)nch, 3);
    if (a.dtype == COMA_BF16) wgrad_cg1_k3s1_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a, vchunk);
    else wgrad_cg1_k3s1_kernel<float><<<grid, 256, 0, stream>>>(a, vchunk);
    COMA_CHECK_LAUNCH("wgrad_cg1_k3s1");
    return COMA_OK;
  }
  if (a.Cg == 1 && a.ksize == 1 && a.stride == 1 && (a.Cx == 8 || a.Cx == 16 || a.Cx == 32 || a.Cx == 64) && a.x_cs % 8 == 0 &&
      a.x_co % 8 == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && a.Dg == a.Dx && a.Hg == a.Hx && a.Wg == a.Wx) {
    const int64_t want = (int64_t)num_sms() * 8;
    int64_t vchunk = (total + want - 1) / want;
    if (vchunk < 2048) vchunk = 2048;
    const unsigned nch = (unsigned)((total + vchunk - 1) / vchunk);
    if (a.dtype == COMA_BF16) wgrad_cg1_k1_kernel<__nv_bfloat16><<<nch, 256, 0, stream>>>(a, vchunk);
    else wgrad_cg1_k1_kernel<float><<<nch, 256, 0, stream>>>(a, vchunk);
    COMA_CHECK_LAUNCH("wgrad_cg1_k1");
    return COMA_OK;
  }
  if (a.Cg == 1 && a.ksize == 1 && a.stride == 1 && (a.Cx == 1 || a.Cx == 2 || a.Cx == 4) && a.x_cs == a.Cx && a.x_co == 0 &&
      (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && a.Dg == a.Dx && a.Hg == a.Hx && a.Wg == a.Wx) {
    const int64_t want = (int64_t)num_sms() * 8;
    int64_t vchunk = ((total + want - 1) / want + 7) / 8 * 8;
    if (vchunk < 4096) vchunk = 4096;
    const unsigned nch = (unsigned)((total + vchunk - 1) / vchunk);
#define COMA_CG1_FEW(T)                                                                          \
    do {                                                                                         \
      if (a.Cx == 1) wgrad_cg1_k1_few_kernel<T, 1><<<nch, 256, 0, stream>>>(a, vchunk);          \
      else if (a.Cx == 2) wgrad_cg1_k1_few_kernel<T, 2><<<nch, 256, 0, stream>>>(a, vchunk);     \
      else wgrad_cg1_k1_few_kernel<T, 4><<<nch, 256, 0, stream>>>(a, vchunk);                    \
    } while (0)
    if (a.dtype == COMA_BF16) COMA_CG1_FEW(__nv_bfloat16);
    else COMA_CG1_FEW(float);
#undef COMA_CG1_FEW
    COMA_CHECK_LAUNCH("wgrad_cg1_k1_few");
    return COMA_OK;
  }
  if (a.Cg == 1 && a.Cx <= 64) {
    const int taps = a.ksize * a.ksize * a.ksize;
    int64_t want = (int64_t)num_sms() * 8 / taps;
    if (want < 1) want = 1;
    int64_t vchunk = (total + want - 1) / want;
    if (vchunk < 2048) vchunk = 2048;
    const int64_t nch = (total + vchunk - 1) / vchunk;
    dim3 grid((unsigned)nch, (unsigned)taps);
    if (a.dtype == COMA_BF16) wgrad_cg1_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a, vchunk);
    else wgrad_cg1_kernel<float><<<grid, 256, 0, stream>>>(a, vchunk);
    COMA_CHECK_LAUNCH("wgrad_cg1");
    return COMA_OK;
  }
  int64_t nchunks = (total + 255) / 256;
  if (nchunks > 512) nchunks = 512;
  int64_t vchunk = (total + nchunks - 1) / nchunks;
  vchunk = (vchunk + kWgVox - 1) / kWgVox * kWgVox;
  nchunks = (total + vchunk - 1) / vchunk;
  const int cg_tiles = (a.Cg + kWgTile - 1) / kWgTile, cx_tiles = (a.Cx + kWgTile - 1) / kWgTile;
  dim3 grid((unsigned)nchunks, (unsigned)(a.ksize * a.ksize * a.ksize), (unsigned)(cg_tiles * cx_tiles));
  if (a.dtype == COMA_BF16) wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a, vchunk, cx_tiles);
  else wgrad_simt_kernel<float><<<grid, 256, 0, stream>>>(a, vchunk, cx_tiles);
  COMA_CHECK_LAUNCH("wgrad_simt");
  return COMA_OK;
}

}  // namespace coma
