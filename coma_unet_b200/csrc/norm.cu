// Normalisation statistics, FiLM/affine + activation apply, and their backward.  HBM-bound kernels:
// every pass is one coalesced, 16-byte-vectorised sweep over an NDHWC tensor with warp-shuffle /
// shared-memory reductions.  Replaces the BatchNorm3d / InstanceNorm3d / PReLU / ReLU / LeakyReLU
// kernel chains of MONAI's ADN (reference: attn_unet_data_parallel.py:285-306,495-497,558; MONAI
// blocks/acti_norm.py) plus the missing CondConv modulation (oracle/cond_conv.py).
//
// Algorithmic bytes per element (s = sizeof(T)):  stats: s read;  apply fwd: s read + s write;
// bwd reduce: 2s read;  bwd apply: 2s read + s write.
#include "common.cuh"

namespace coma {

constexpr int kThreads = 256;

static int stats_chunks(int64_t V) {
  int64_t c = (V + 2047) / 2048;
  return (int)(c < 1 ? 1 : (c > 256 ? 256 : c));
}

// ---- per-(b, chunk, c) partial sums ------------------------------------------------------------
// NQ = 2: (sum x, sum x^2).   NQ = 3 (backward): (sum dz, sum dz*xhat, sum_{u<0} dy*u)
struct BwdCtx {
  const void* dy; const float* A; const float* S; const float* mean; const float* rstd; const float* slope;
  int dy_cs, dy_co, act;
  const void* r; int r_cs;   // optional residual added before the activation: u = A*x + S + r
};

// SIMPLE: act in {none, relu, leaky/prelu}: d act/du = u > 0 ? 1 : gneg, d act/d slope = leaky ? min(u, 0) : 0 -- no per-element
// dispatch on the activation code (the switch made these sweeps instruction- rather than HBM-bound)
template <typename T, int NQ, bool SIMPLE>
__device__ __forceinline__ void accumulate8(const float (&xv)[8], const float (&dyv)[8], const float (&rv)[8], const float* A8,
                                            const float* S8, const float* M8, const float* R8, int act, float slope,
                                            float (&acc)[8][NQ]) {
  const float gneg = act == COMA_ACT_NONE ? 1.f : (act == COMA_ACT_RELU ? 0.f : slope);
  const float sflag = act == COMA_ACT_LEAKY ? 1.f : 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (NQ == 2) {
      acc[e][0] += xv[e];
      acc[e][1] += xv[e] * xv[e];
    } else {
      const float u = fmaf(A8[e], xv[e], S8[e]) + rv[e];
      const float dz = dyv[e] * (SIMPLE ? (u > 0.f ? 1.f : gneg) : act_grad(act, u, slope));
      acc[e][0] += dz;
      acc[e][1] += dz * (xv[e] - M8[e]) * R8[e];
      if (NQ > 2) acc[e][NQ - 1] += SIMPLE ? sflag * dyv[e] * fminf(u, 0.f) : dyv[e] * act_slope_grad(act, u, slope);
    }
  }
}

template <typename T, int NQ, bool SIMPLE>
__global__ void __launch_bounds__(kThreads) reduce_vec_kernel(const T* __restrict__ x, int64_t V, int C, int cs, int co,
                                                              int chunks, float* __restrict__ partial, BwdCtx ctx) {
  extern __shared__ float red[];  // [kThreads][8*NQ]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int CV = C >> 3, lanes = kThreads / CV;
  const int cvec = threadIdx.x % CV, vlane = threadIdx.x / CV;
  const int64_t per = (V + chunks - 1) / chunks, v0 = (int64_t)chunk * per, v1 = min(v0 + per, V);
  float acc[8][NQ];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[e][q] = 0.f;
  float A8[8], S8[8], M8[8], R8[8];
  float slope = 0.f;
  if (NQ == 3) {
    const int64_t o = (int64_t)b * C + cvec * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      A8[e] = ctx.A[o + e]; S8[e] = ctx.S[o + e]; M8[e] = ctx.mean[o + e]; R8[e] = ctx.rstd[o + e];
    }
    slope = ctx.slope ? __ldg(ctx.slope) : 0.f;
  }
  const T* xb = x + (int64_t)b * V * cs + co + cvec * 8;
  const T* dyb = NQ == 3 ? static_cast<const T*>(ctx.dy) + (int64_t)b * V * ctx.dy_cs + ctx.dy_co + cvec * 8 : nullptr;
  const T* rb = (NQ == 3 && ctx.r) ? static_cast<const T*>(ctx.r) + (int64_t)b * V * ctx.r_cs + cvec * 8 : nullptr;
  if (vlane < lanes) {
    int64_t v = v0 + vlane;
    // two voxels per trip: all loads of both are issued before the first accumulate (the sweep is latency- not issue-bound)
    for (; v + lanes < v1; v += 2 * lanes) {
      float xv[8], dyv[8], rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float xw[8], dyw[8], rw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      load8(xb + v * cs, xv);
      load8(xb + (v + lanes) * cs, xw);
      if (NQ == 3) load8(dyb + v * ctx.dy_cs, dyv);
      if (NQ == 3) load8(dyb + (v + lanes) * ctx.dy_cs, dyw);
      if (NQ == 3 && rb) load8(rb + v * ctx.r_cs, rv);
      if (NQ == 3 && rb) load8(rb + (v + lanes) * ctx.r_cs, rw);
      accumulate8<T, NQ, SIMPLE>(xv, dyv, rv, A8, S8, M8, R8, ctx.act, slope, acc);
      accumulate8<T, NQ, SIMPLE>(xw, dyw, rw, A8, S8, M8, R8, ctx.act, slope, acc);
    }
    if (v < v1) {
      float xv[8], dyv[8], rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      load8(xb + v * cs, xv);
      if (NQ == 3) load8(dyb + v * ctx.dy_cs, dyv);
      if (NQ == 3 && rb) load8(rb + v * ctx.r_cs, rv);
      accumulate8<T, NQ, SIMPLE>(xv, dyv, rv, A8, S8, M8, R8, ctx.act, slope, acc);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int q = 0; q < NQ; ++q) red[threadIdx.x * 8 * NQ + e * NQ + q] = acc[e][q];
  __syncthreads();
  for (int i = threadIdx.x; i < C * NQ; i += kThreads) {
    const int c = i / NQ, q = i % NQ;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[(l * CV + (c >> 3)) * 8 * NQ + (c & 7) * NQ + q];
    partial[(((int64_t)b * chunks + chunk) * C + c) * NQ + q] = s;
  }
}

// scalar variant for C < 8 (1-channel volumes of the modulator stacks and heads)
template <typename T, int NQ>
__global__ void __launch_bounds__(kThreads) reduce_small_kernel(const T* __restrict__ x, int64_t V, int C, int cs, int co,
                                                                int chunks, float* __restrict__ partial, BwdCtx ctx) {
  __shared__ float red[kThreads / 32][8][NQ];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (V + chunks - 1) / chunks, v0 = (int64_t)chunk * per, v1 = min(v0 + per, V);
  float acc[8][NQ];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[e][q] = 0.f;
  const float slope = (NQ == 3 && ctx.slope) ? __ldg(ctx.slope) : 0.f;
  const T* xb = x + (int64_t)b * V * cs + co;
  const T* dyb = NQ == 3 ? static_cast<const T*>(ctx.dy) + (int64_t)b * V * ctx.dy_cs + ctx.dy_co : nullptr;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < C) {
        const float xv = Elem<T>::ld(xb + v * cs + c);
        if (NQ == 2) {
          acc[c][0] += xv;
          acc[c][1] += xv * xv;
        } else {
          const int64_t o = (int64_t)b * C + c;
          const float u = fmaf(ctx.A[o], xv, ctx.S[o]) +
                          (ctx.r ? Elem<T>::ld(static_cast<const T*>(ctx.r) + ((int64_t)b * V + v) * ctx.r_cs + c) : 0.f);
          const float dyv = Elem<T>::ld(dyb + v * ctx.dy_cs + c);
          const float dz = dyv * act_grad(ctx.act, u, slope);
          acc[c][0] += dz;
          acc[c][1] += dz * (xv - ctx.mean[o]) * ctx.rstd[o];
          acc[c][NQ - 1] += dyv * act_slope_grad(ctx.act, u, slope);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float s = warp_sum(acc[c][q]);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c][q] = s;
    }
  __syncthreads();
  if (threadIdx.x < C * NQ) {
    const int c = threadIdx.x / NQ, q = threadIdx.x % NQ;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w][c][q];
    partial[(((int64_t)b * chunks + chunk) * C + c) * NQ + q] = s;
  }
}

// dense ONE-channel volumes (psi of the four gates, the 16 -> 1 / 2 -> 1 modulator heads): 8 voxels per 16-byte load instead of
// one 2-byte load per thread (reduce_small_kernel ran at 1.2 TB/s: 27 us for 34 MB at batch 4 x 128^3)
template <typename T, int NQ>
__global__ void __launch_bounds__(kThreads) reduce_c1_kernel(const T* __restrict__ x, int64_t V, int chunks, float* __restrict__ partial,
                                                             BwdCtx ctx) {
  __shared__ float red[kThreads / 32][NQ];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = V / chunks, v0 = (int64_t)chunk * per, v1 = v0 + per;      // per is a multiple of 8 (checked by the launcher)
  float acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
  const float slope = (NQ == 3 && ctx.slope) ? __ldg(ctx.slope) : 0.f;
  const float A = NQ == 3 ? ctx.A[b] : 0.f, S = NQ == 3 ? ctx.S[b] : 0.f, mean = NQ == 3 ? ctx.mean[b] : 0.f,
              rstd = NQ == 3 ? ctx.rstd[b] : 0.f;
  const T* xb = x + (int64_t)b * V;
  const T* dyb = NQ == 3 ? static_cast<const T*>(ctx.dy) + (int64_t)b * V : nullptr;
  const T* rb = (NQ == 3 && ctx.r) ? static_cast<const T*>(ctx.r) + (int64_t)b * V : nullptr;
  for (int64_t v = v0 + threadIdx.x * 8; v < v1; v += kThreads * 8) {
    float xv[8], dyv[8], rv[8];
    load8_stream(xb + v, xv);
    if (NQ == 3) {
      load8_stream(dyb + v, dyv);
      if (rb) load8_stream(rb + v, rv);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (NQ == 2) {
        acc[0] += xv[e];
        acc[1] += xv[e] * xv[e];
      } else {
        const float u = fmaf(A, xv[e], S) + (rb ? rv[e] : 0.f);
        const float dz = dyv[e] * act_grad(ctx.act, u, slope);
        acc[0] += dz;
        acc[1] += dz * (xv[e] - mean) * rstd;
        acc[NQ - 1] += dyv[e] * act_slope_grad(ctx.act, u, slope);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float t = warp_sum(acc[q]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = t;
  }
  __syncthreads();
  if (threadIdx.x < NQ) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += red[w][threadIdx.x];
    partial[((int64_t)b * chunks + chunk) * NQ + threadIdx.x] = t;
  }
}

static bool c1_ok(const void* p, int C, int cs, int co, int64_t V, int chunks, int dtype) {
  const int esz = dtype == COMA_BF16 ? 2 : 4;
  return C == 1 && cs == 1 && co == 0 && V % ((int64_t)chunks * 8) == 0 && (V * esz) % 16 == 0 &&
         reinterpret_cast<uintptr_t>(p) % 16 == 0;
}

static bool vec_ok(const void* p, int C, int cs, int co, int dtype) {
  const int esz = dtype == COMA_BF16 ? 2 : 4;
  return C >= 8 && (C % 8 == 0) && (kThreads % (C / 8) == 0) && (cs % 8 == 0) && (co % 8 == 0) &&
         (reinterpret_cast<uintptr_t>(p) % (8 * esz > 16 ? 16 : 8 * esz) == 0);
}

template <typename T, int NQ>
static int launch_reduce(const void* x, int B, int64_t V, int C, int cs, int co, int dtype, int chunks, float* partial,
                         const BwdCtx& ctx, cudaStream_t stream) {
  dim3 grid((unsigned)chunks, (unsigned)B);
  bool v = vec_ok(x, C, cs, co, dtype);
  if (NQ == 3) v = v && vec_ok(ctx.dy, C, ctx.dy_cs, ctx.dy_co, dtype) && (!ctx.r || vec_ok(ctx.r, C, ctx.r_cs, 0, dtype));
  if (v) {
    const size_t smem = (size_t)kThreads * 8 * NQ * sizeof(float);
    const bool simple = NQ == 3 && (ctx.act == COMA_ACT_NONE || ctx.act == COMA_ACT_RELU || ctx.act == COMA_ACT_LEAKY);
    if (simple) reduce_vec_kernel<T, NQ, true><<<grid, kThreads, smem, stream>>>(static_cast<const T*>(x), V, C, cs, co, chunks, partial, ctx);
    else reduce_vec_kernel<T, NQ, false><<<grid, kThreads, smem, stream>>>(static_cast<const T*>(x), V, C, cs, co, chunks, partial, ctx);
  } else if (c1_ok(x, C, cs, co, V, chunks, dtype) &&
             (NQ != 3 || (c1_ok(ctx.dy, C, ctx.dy_cs, ctx.dy_co, V, chunks, dtype) && (!ctx.r || c1_ok(ctx.r, C, ctx.r_cs, 0, V, chunks, dtype))))) {
    reduce_c1_kernel<T, NQ><<<grid, kThreads, 0, stream>>>(static_cast<const T*>(x), V, chunks, partial, ctx);
  } else {
    COMA_CHECK_ARG(C <= 8, "norm reduce: C=%d must be a power-of-two multiple of 8 (aligned) or <= 8", C);
    reduce_small_kernel<T, NQ><<<grid, kThreads, 0, stream>>>(static_cast<const T*>(x), V, C, cs, co, chunks, partial, ctx);
  }
  COMA_CHECK_LAUNCH("norm_reduce");
  return COMA_OK;
}

// ---- forward finalize ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) finalize_kernel(coma_norm_finalize_args a) {
  // INSTANCE: one block per (b, c).  BATCH: one block per c.  NONE / GIVEN: one thread per (b, c).
  __shared__ double sh[2][kThreads / 32];
  const int B = a.B, C = a.C;
  if (a.mode == COMA_NORM_NONE || a.mode == COMA_NORM_GIVEN) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= B * C) return;
    const int c = i % C;
    float mean = 0.f, rstd = 1.f;
    if (a.mode == COMA_NORM_GIVEN) {
      mean = a.given_mean[c];
      rstd = rsqrtf(a.given_var[c] + a.eps);
    }
    const float g = a.g ? a.g[i] : 1.f, h = a.h ? a.h[i] : 0.f;
    a.A[i] = g * rstd;
    a.S[i] = h - g * rstd * mean;
    a.mean[i] = mean;
    a.rstd[i] = rstd;
    return;
  }
  const bool batch = a.mode == COMA_NORM_BATCH;
  const int c = batch ? blockIdx.x : blockIdx.x % C;
  const int b0 = batch ? 0 : blockIdx.x / C, b1 = batch ? B : b0 + 1;
  double s1 = 0.0, s2 = 0.0;
  const int n = (b1 - b0) * a.chunks;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const int b = b0 + i / a.chunks, ch = i % a.chunks;
    const float* p = a.partial + (((int64_t)b * a.chunks + ch) * C + c) * 2;
    s1 += p[0];
    s2 += p[1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s1;
    sh[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s1 = 0.0; s2 = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { s1 += sh[0][w]; s2 += sh[1][w]; }
    const double cnt = (double)(b1 - b0) * (double)a.V;
    const double mean = s1 / cnt;
    double var = s2 / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
    for (int b = b0; b < b1; ++b) {
      const int i = b * C + c;
      const float g = a.g ? a.g[i] : 1.f, h = a.h ? a.h[i] : 0.f;
      a.A[i] = g * rstd;
      a.S[i] = h - g * rstd * (float)mean;
      a.mean[i] = (float)mean;
      a.rstd[i] = rstd;
    }
    if (batch && a.running_mean) {
      const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
      float rm = a.running_mean[c], rv = a.running_var[c];
      for (int u = 0; u < a.n_updates; ++u) {
        rm = (1.f - a.momentum) * rm + a.momentum * (float)mean;
        rv = (1.f - a.momentum) * rv + a.momentum * (float)unbiased;
      }
      a.running_mean[c] = rm;
      a.running_var[c] = rv;
    }
  }
}

// ---- apply: y = act(A*x + S) ---------------------------------------------------------------------
// SIMPLE: act in {none, relu, leaky/prelu} evaluated as max(u,0) + neg * min(u,0) (no per-element dispatch on the act code)
template <typename T, int U, bool SIMPLE, bool HASR>
__global__ void __launch_bounds__(kThreads) affine_act_vec_kernel(coma_affine_act_args a, int chunks, int creal) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int CV = a.C >> 3, lanes = kThreads / CV;
  const int cvec = threadIdx.x % CV, vlane = threadIdx.x / CV;
  if (vlane >= lanes) return;
  const int64_t per = (a.V + chunks - 1) / chunks, v0 = (int64_t)chunk * per, v1 = min(v0 + per, a.V);
  float A8[8], S8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    // creal > 0: a dense tensor with fewer than 8 channels viewed as 8-wide vectors (creal divides 8)
    const int64_t ci = creal > 0 ? (int64_t)b * creal + (e % creal) : (int64_t)b * a.C + cvec * 8 + e;
    A8[e] = a.A[ci];
    S8[e] = a.S[ci];
  }
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  const float neg = a.act == COMA_ACT_NONE ? 1.f : (a.act == COMA_ACT_RELU ? 0.f : slope);
  const T* xb = static_cast<const T*>(a.x) + (int64_t)b * a.V * a.x_cs + a.x_co + cvec * 8;
  T* yb = static_cast<T*>(a.y) + (int64_t)b * a.V * a.y_cs + a.y_co + cvec * 8;
  const T* rb = HASR ? static_cast<const T*>(a.r) + (int64_t)b * a.V * a.r_cs + cvec * 8 : nullptr;
  // four voxels per thread and iteration, every load issued before the first use: HBM streaming wants bytes in flight
  for (int64_t vb = v0 + vlane; vb < v1; vb += (int64_t)lanes * U) {
    Raw8<T> xr[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = vb + (int64_t)u * lanes;
      if (v < v1) {
        ldraw_stream(xb + v * a.x_cs, xr[u]);
        if (HASR) ldraw_stream(rb + v * a.r_cs, rr[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = vb + (int64_t)u * lanes;
      if (v < v1) {
        float xv[8], rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        unpack8(xr[u], xv);
        if (HASR) unpack8(rr[u], rv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float u = HASR ? fmaf(A8[e], xv[e], S8[e]) + rv[e] : fmaf(A8[e], xv[e], S8[e]);
          xv[e] = SIMPLE ? fmaf(neg, fminf(u, 0.f), fmaxf(u, 0.f)) : act_fwd(a.act, u, slope);
        }
        store8(yb + v * a.y_cs, xv);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) affine_act_small_kernel(coma_affine_act_args a) {
  const int b = blockIdx.y;
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  const T* xb = static_cast<const T*>(a.x) + (int64_t)b * a.V * a.x_cs + a.x_co;
  T* yb = static_cast<T*>(a.y) + (int64_t)b * a.V * a.y_cs + a.y_co;
  for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < a.V; v += (int64_t)gridDim.x * kThreads)
    for (int c = 0; c < a.C; ++c) {
      const float u = fmaf(a.A[(int64_t)b * a.C + c], Elem<T>::ld(xb + v * a.x_cs + c), a.S[(int64_t)b * a.C + c]) +
                      (a.r ? Elem<T>::ld(static_cast<const T*>(a.r) + ((int64_t)b * a.V + v) * a.r_cs + c) : 0.f);
      Elem<T>::st(yb + v * a.y_cs + c, act_fwd(a.act, u, slope));
    }
}

// ---- backward finalize: partial sums -> dg, dh, dslope and dx = P*dz + Q + R*x coefficients -------
__global__ void __launch_bounds__(kThreads) bwd_finalize_kernel(coma_affine_act_bwd_args a, int chunks) {
  // one block per channel c.  The (b, chunk) partials of FB batch entries are summed at a time -- 3 * FB independent reductions
  // behind ONE barrier pair (this kernel sits between the two sweeps of every norm backward: its latency is on the step's
  // critical path 39 times per training step) -- then thread b turns the sums of batch entry b into its coefficients.
  constexpr int FB = 4, NW = kThreads / 32;
  __shared__ float sh[NW][FB * 3];
  __shared__ float red[3];                 // batch mode: sum_b g*dh, sum_b g*dg; always: sum_b dslope terms
  const int c = blockIdx.x, B = a.B, C = a.C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m1 = 0.f, m2 = 0.f, ds = 0.f;      // meaningful in thread 0
  for (int b0 = 0; b0 < B; b0 += FB) {
    float s[FB][3];
#pragma unroll
    for (int j = 0; j < FB; ++j) s[j][0] = s[j][1] = s[j][2] = 0.f;
    for (int ch = threadIdx.x; ch < chunks; ch += kThreads) {
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        if (b0 + j < B) {
          const float* p = a.partial + (((int64_t)(b0 + j) * chunks + ch) * C + c) * 3;
          s[j][0] += p[0]; s[j][1] += p[1]; s[j][2] += p[2];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < FB; ++j) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float v = warp_sum(s[j][q]);
        if (lane == 0) sh[warp][j * 3 + q] = v;
      }
    }
    __syncthreads();
    if (threadIdx.x < FB * 3) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += sh[w][threadIdx.x];
      sh[0][threadIdx.x] = t;              // only thread x reads / writes column x of row 0 here
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int j = 0; j < FB && b0 + j < B; ++j) {
        const int i = (b0 + j) * C + c;
        const float g = a.g ? a.g[i] : 1.f;
        a.dh[i] = sh[0][j * 3 + 0];
        a.dg[i] = sh[0][j * 3 + 1];
        m1 += g * sh[0][j * 3 + 0];
        m2 += g * sh[0][j * 3 + 1];
        ds += sh[0][j * 3 + 2];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (a.dslope && (a.act == COMA_ACT_LEAKY || a.act == COMA_ACT_LEAKY_RELU)) atomicAdd(a.dslope, ds);
    red[0] = m1; red[1] = m2;
  }
  __syncthreads();
  m1 = red[0]; m2 = red[1];
  const float invBV = 1.f / ((float)B * (float)a.V), invV = 1.f / (float)a.V;
  for (int b = threadIdx.x; b < B; b += kThreads) {
    const int i = b * C + c;
    const float P = a.A[i], rstd = a.rstd[i], mean = a.mean[i];
    float R = 0.f, Q = 0.f;
    if (a.mode == COMA_NORM_INSTANCE) {
      R = -P * rstd * a.dg[i] * invV;      // dg / dh of entry b were written by thread 0 of this block before the barrier above
      Q = -P * a.dh[i] * invV - R * mean;
    } else if (a.mode == COMA_NORM_BATCH) {
      R = -rstd * rstd * m2 * invBV;
      Q = -rstd * m1 * invBV - R * mean;
    }
    a.coef[i * 3 + 0] = P;
    a.coef[i * 3 + 1] = Q;
    a.coef[i * 3 + 2] = R;
  }
}

template <typename T, bool SIMPLE>
__global__ void __launch_bounds__(kThreads) bwd_apply_vec_kernel(coma_affine_act_bwd_args a, int chunks) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int CV = a.C >> 3, lanes = kThreads / CV;
  const int cvec = threadIdx.x % CV, vlane = threadIdx.x / CV;
  if (vlane >= lanes) return;
  const int64_t per = (a.V + chunks - 1) / chunks, v0 = (int64_t)chunk * per, v1 = min(v0 + per, a.V);
  float A8[8], S8[8], P8[8], Q8[8], R8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int64_t i = (int64_t)b * a.C + cvec * 8 + e;
    A8[e] = a.A[i]; S8[e] = a.S[i];
    P8[e] = a.coef[i * 3]; Q8[e] = a.coef[i * 3 + 1]; R8[e] = a.coef[i * 3 + 2];
  }
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  const float gneg = a.act == COMA_ACT_NONE ? 1.f : (a.act == COMA_ACT_RELU ? 0.f : slope);
  const T* xb = static_cast<const T*>(a.x) + (int64_t)b * a.V * a.x_cs + a.x_co + cvec * 8;
  const T* dyb = static_cast<const T*>(a.dy) + (int64_t)b * a.V * a.dy_cs + a.dy_co + cvec * 8;
  T* dxb = static_cast<T*>(a.dx) + (int64_t)b * a.V * a.dx_cs + a.dx_co + cvec * 8;
  const T* rb = a.r ? static_cast<const T*>(a.r) + (int64_t)b * a.V * a.r_cs + cvec * 8 : nullptr;
  T* drb = a.dr ? static_cast<T*>(a.dr) + (int64_t)b * a.V * a.dr_cs + cvec * 8 : nullptr;
  for (int64_t v = v0 + vlane; v < v1; v += lanes) {
    float xv[8], dyv[8], rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    load8(xb + v * a.x_cs, xv);
    load8(dyb + v * a.dy_cs, dyv);
    if (rb) load8(rb + v * a.r_cs, rv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float u = fmaf(A8[e], xv[e], S8[e]) + rv[e];
      const float dz = dyv[e] * (SIMPLE ? (u > 0.f ? 1.f : gneg) : act_grad(a.act, u, slope));
      xv[e] = fmaf(P8[e], dz, fmaf(R8[e], xv[e], Q8[e]));
      rv[e] = dz;
    }
    store8(dxb + v * a.dx_cs, xv);
    if (drb) store8(drb + v * a.dr_cs, rv);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) bwd_apply_small_kernel(coma_affine_act_bwd_args a) {
  const int b = blockIdx.y;
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  const T* xb = static_cast<const T*>(a.x) + (int64_t)b * a.V * a.x_cs + a.x_co;
  const T* dyb = static_cast<const T*>(a.dy) + (int64_t)b * a.V * a.dy_cs + a.dy_co;
  T* dxb = static_cast<T*>(a.dx) + (int64_t)b * a.V * a.dx_cs + a.dx_co;
  for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < a.V; v += (int64_t)gridDim.x * kThreads)
    for (int c = 0; c < a.C; ++c) {
      const int64_t i = (int64_t)b * a.C + c;
      const float xv = Elem<T>::ld(xb + v * a.x_cs + c);
      const float u = fmaf(a.A[i], xv, a.S[i]) +
                      (a.r ? Elem<T>::ld(static_cast<const T*>(a.r) + ((int64_t)b * a.V + v) * a.r_cs + c) : 0.f);
      const float dz = Elem<T>::ld(dyb + v * a.dy_cs + c) * act_grad(a.act, u, slope);
      Elem<T>::st(dxb + v * a.dx_cs + c, fmaf(a.coef[i * 3], dz, fmaf(a.coef[i * 3 + 2], xv, a.coef[i * 3 + 1])));
      if (a.dr) Elem<T>::st(static_cast<T*>(a.dr) + ((int64_t)b * a.V + v) * a.dr_cs + c, dz);
    }
}


template <typename T>
__global__ void __launch_bounds__(kThreads) bwd_apply_c1_kernel(coma_affine_act_bwd_args a) {
  // dense one-channel volume: 8 voxels per thread and step (see reduce_c1_kernel)
  const int b = blockIdx.y;
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  const float A = a.A[b], S = a.S[b], P = a.coef[b * 3], Q = a.coef[b * 3 + 1], R = a.coef[b * 3 + 2];
  const T* xb = static_cast<const T*>(a.x) + (int64_t)b * a.V;
  const T* dyb = static_cast<const T*>(a.dy) + (int64_t)b * a.V;
  const T* rb = a.r ? static_cast<const T*>(a.r) + (int64_t)b * a.V : nullptr;
  T* dxb = static_cast<T*>(a.dx) + (int64_t)b * a.V;
  T* drb = a.dr ? static_cast<T*>(a.dr) + (int64_t)b * a.V : nullptr;
  for (int64_t v = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 8; v < a.V; v += (int64_t)gridDim.x * kThreads * 8) {
    float xv[8], dyv[8], rv[8], dx[8], dz[8];
    load8_stream(xb + v, xv);
    load8_stream(dyb + v, dyv);
    if (rb) load8_stream(rb + v, rv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float u = fmaf(A, xv[e], S) + (rb ? rv[e] : 0.f);
      dz[e] = dyv[e] * act_grad(a.act, u, slope);
      dx[e] = fmaf(P, dz[e], fmaf(R, xv[e], Q));
    }
    store8(dxb + v, dx);
    if (drb) store8(drb + v, dz);
  }
}

// ---- bulk-copy streaming variants of the two backward sweeps (bf16, contiguous tensors) ------------------------------------
// The register-staged sweeps above top out near half of the HBM peak: bytes in flight cost registers (126 regs, two blocks per
// SM in the reduce sweep).  Here a ring of shared-memory stages is filled by cp.async.bulk (one elected thread issues 16 KB
// copies that complete on an mbarrier), so ~100 KB per SM are in flight whatever the compute part needs, and the 256 threads
// only read shared memory.  A tile is kTileBytes of each operand = 1024 16-byte vectors; thread t owns vectors t, t+256, ...:
// kThreads is a multiple of the C/8 vectors of a voxel, so a thread always sees the same 8 channels (coefficients in registers).
constexpr int kTileBytes = 16384;
constexpr int kTileVecs = kTileBytes / 16;
constexpr int kBulkStages = 3;

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bulk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_addr_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();   // watchdog: a lost copy must not hang the GPU
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr_u32(bar))
               : "memory");
}
__device__ __forceinline__ void unpack8_u4(const uint4& r, float (&v)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

struct BulkRing {
  const char* src[3];      // x, dy, residual (contiguous [V, C] bf16 of sample b, offset to the block's first voxel)
  int nsrc;
  int64_t bytes;           // bytes of each operand this block sweeps
};

// Issue the copies of tile `t` into stage `t % kBulkStages` (nsrc operand slots per stage).  Called by one thread.
__device__ __forceinline__ void bulk_issue(const BulkRing& ring, char* stage_base, uint64_t* full, int t) {
  const int s = t % kBulkStages;
  const int64_t off = (int64_t)t * kTileBytes;
  const uint32_t n = (uint32_t)min((int64_t)kTileBytes, ring.bytes - off);
  bulk_mbar_expect_tx(&full[s], n * ring.nsrc);
  for (int i = 0; i < ring.nsrc; ++i)
    bulk_load(stage_base + ((size_t)s * ring.nsrc + i) * kTileBytes, ring.src[i] + off, n, &full[s]);
}

// MODE 0: reduce sweep (partial sums of dz, dz*xhat, dslope terms).  MODE 1: apply sweep (dx = P*dz + Q + R*x, optional dr = dz).
template <int MODE, bool SIMPLE, bool HASR>
__global__ void __launch_bounds__(kThreads, 2) bwd_bulk_kernel(coma_affine_act_bwd_args a, int chunks) {
  extern __shared__ __align__(128) char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);                 // [kBulkStages]
  constexpr int NS = HASR ? 3 : 2;
  char* stages = smem + 128;                                           // [kBulkStages][NS][kTileBytes]
  const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int C = a.C, CV = C >> 3;
  const int cvec = tid % CV;
  // chunk boundaries in voxels, rounded so that every chunk starts on a tile boundary of whole voxels
  const int tile_vox = kTileVecs / CV;
  const int64_t tiles_total = (a.V + tile_vox - 1) / tile_vox;
  const int64_t tiles_per = (tiles_total + chunks - 1) / chunks;
  const int64_t v0 = min((int64_t)chunk * tiles_per * tile_vox, a.V), v1 = min(v0 + tiles_per * tile_vox, a.V);
  const int ntiles = (int)((v1 - v0 + tile_vox - 1) / tile_vox);
  BulkRing ring;
  const int64_t row = (int64_t)C * 2;
  ring.src[0] = static_cast<const char*>(a.x) + ((int64_t)b * a.V + v0) * row;
  ring.src[1] = static_cast<const char*>(a.dy) + ((int64_t)b * a.V + v0) * row;
  ring.src[2] = HASR ? static_cast<const char*>(a.r) + ((int64_t)b * a.V + v0) * row : nullptr;
  ring.nsrc = NS;
  ring.bytes = (v1 - v0) * row;
  if (tid == 0) {
    for (int s = 0; s < kBulkStages; ++s) bulk_mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
    for (int t = 0; t < min(kBulkStages, ntiles); ++t) bulk_issue(ring, stages, full, t);

  float A8[8], S8[8], M8[8], R8[8], P8[8], Q8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int64_t i = (int64_t)b * C + cvec * 8 + e;
    A8[e] = a.A[i]; S8[e] = a.S[i];
    if (MODE == 0) { M8[e] = a.mean[i]; R8[e] = a.rstd[i]; }
    else { P8[e] = a.coef[i * 3]; Q8[e] = a.coef[i * 3 + 1]; R8[e] = a.coef[i * 3 + 2]; }
  }
  const float slope = a.slope ? __ldg(a.slope) : 0.f;
  const float gneg = a.act == COMA_ACT_NONE ? 1.f : (a.act == COMA_ACT_RELU ? 0.f : slope);
  const float sflag = a.act == COMA_ACT_LEAKY ? 1.f : 0.f;
  float acc[8][3];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e][0] = acc[e][1] = acc[e][2] = 0.f;
  __nv_bfloat16* dxb = MODE == 1 ? static_cast<__nv_bfloat16*>(a.dx) + ((int64_t)b * a.V + v0) * C : nullptr;
  __nv_bfloat16* drb = (MODE == 1 && a.dr) ? static_cast<__nv_bfloat16*>(a.dr) + ((int64_t)b * a.V + v0) * C : nullptr;

  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kBulkStages;
    bulk_mbar_wait(&full[s], (uint32_t)(t / kBulkStages) & 1u);
    const int64_t off = (int64_t)t * kTileBytes;
    const int nvec = (int)(min((int64_t)kTileBytes, ring.bytes - off) >> 4);
    const uint4* sx = reinterpret_cast<const uint4*>(stages + ((size_t)s * NS + 0) * kTileBytes);
    const uint4* sd = reinterpret_cast<const uint4*>(stages + ((size_t)s * NS + 1) * kTileBytes);
    const uint4* sr = reinterpret_cast<const uint4*>(stages + ((size_t)s * NS + (HASR ? 2 : 0)) * kTileBytes);
#pragma unroll
    for (int j = 0; j < kTileVecs / kThreads; ++j) {
      const int vi = tid + j * kThreads;
      if (vi < nvec) {
        float xv[8], dyv[8], rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        unpack8_u4(sx[vi], xv);
        unpack8_u4(sd[vi], dyv);
        if (HASR) unpack8_u4(sr[vi], rv);
        if (MODE == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float u = fmaf(A8[e], xv[e], S8[e]) + rv[e];
            const float dz = dyv[e] * (SIMPLE ? (u > 0.f ? 1.f : gneg) : act_grad(a.act, u, slope));
            acc[e][0] += dz;
            acc[e][1] += dz * (xv[e] - M8[e]) * R8[e];
            acc[e][2] += SIMPLE ? sflag * dyv[e] * fminf(u, 0.f) : dyv[e] * act_slope_grad(a.act, u, slope);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float u = fmaf(A8[e], xv[e], S8[e]) + rv[e];
            const float dz = dyv[e] * (SIMPLE ? (u > 0.f ? 1.f : gneg) : act_grad(a.act, u, slope));
            xv[e] = fmaf(P8[e], dz, fmaf(R8[e], xv[e], Q8[e]));
            rv[e] = dz;
          }
          const int64_t eoff = (off >> 1) + (int64_t)vi * 8;          // element offset inside the block's range
          store8(dxb + eoff, xv);
          if (drb) store8(drb + eoff, rv);
        }
      }
    }
    __syncthreads();                                                   // every thread is done with stage s
    if (tid == 0 && t + kBulkStages < ntiles) bulk_issue(ring, stages, full, t + kBulkStages);
  }
  if (MODE == 0) {
    // block reduction over the threads that share a channel vector (the stage buffers are free now)
    float* red = reinterpret_cast<float*>(stages);                     // [kThreads][24]
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int q = 0; q < 3; ++q) red[tid * 24 + e * 3 + q] = acc[e][q];
    __syncthreads();
    const int lanes = kThreads / CV;
    for (int i = tid; i < C * 3; i += kThreads) {
      const int c = i / 3, q = i % 3;
      float sum = 0.f;
      for (int l = 0; l < lanes; ++l) sum += red[(l * CV + (c >> 3)) * 24 + (c & 7) * 3 + q];
      a.partial[(((int64_t)b * chunks + chunk) * C + c) * 3 + q] = sum;
    }
  }
}

constexpr size_t bulk_smem(bool hasr) { return 128 + (size_t)kBulkStages * (hasr ? 3 : 2) * kTileBytes; }   // 96 KB: two blocks per SM

static bool bulk_ok(const coma_affine_act_bwd_args* a) {
  static const bool off = [] { const char* e = getenv("COMA_DISABLE_NORM_BULK"); return e && atoi(e) != 0; }();
  if (off || a->dtype != COMA_BF16) return false;
  const int C = a->C;
  if (C < 8 || (C & (C - 1)) != 0 || C > 512) return false;          // C/8 must divide kThreads
  if (a->x_cs != C || a->x_co != 0 || a->dy_cs != C || a->dy_co != 0 || a->dx_cs != C || a->dx_co != 0) return false;
  if (a->r && a->r_cs != C) return false;
  if (a->dr && a->dr_cs != C) return false;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(a->x) || !al(a->dy) || !al(a->dx) || (a->r && !al(a->r)) || (a->dr && !al(a->dr))) return false;
  return a->V * C >= 4 * kTileVecs * 8;                                // small tensors: the plain sweeps are latency-bound anyway
}

template <int MODE>
static int launch_bulk(const coma_affine_act_bwd_args* a, int chunks, cudaStream_t stream) {
  const bool simple = a->act == COMA_ACT_NONE || a->act == COMA_ACT_RELU || a->act == COMA_ACT_LEAKY;
  dim3 grid((unsigned)chunks, (unsigned)a->B);
#define COMA_BULK_LAUNCH(S, R)                                                                                          \
  do {                                                                                                                   \
    static bool attr = false;                                                                                            \
    if (!attr) {                                                                                                         \
      cudaFuncSetAttribute(bwd_bulk_kernel<MODE, S, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem(R));  \
      attr = true;                                                                                                       \
    }                                                                                                                    \
    bwd_bulk_kernel<MODE, S, R><<<grid, kThreads, bulk_smem(R), stream>>>(*a, chunks);                                    \
  } while (0)
  if (simple && !a->r) COMA_BULK_LAUNCH(true, false);
  else if (simple) COMA_BULK_LAUNCH(true, true);
  else if (!a->r) COMA_BULK_LAUNCH(false, false);
  else COMA_BULK_LAUNCH(false, true);
#undef COMA_BULK_LAUNCH
  COMA_CHECK_LAUNCH(MODE == 0 ? "norm_bwd_reduce_bulk" : "norm_bwd_apply_bulk");
  return COMA_OK;
}

}  // namespace coma

using namespace coma;

extern "C" int coma_norm_stats_chunks(int64_t V) { return stats_chunks(V); }

extern "C" int coma_norm_stats(const void* x, int32_t B, int64_t V, int32_t C, int32_t cs, int32_t co, int32_t dtype,
                               float* partial, coma_stream_t stream) {
  COMA_CHECK_ARG(x && partial && B > 0 && V > 0 && C > 0, "coma_norm_stats: bad arguments");
  BwdCtx ctx{};
  const int chunks = stats_chunks(V);
  if (dtype == COMA_BF16) return launch_reduce<__nv_bfloat16, 2>(x, B, V, C, cs, co, dtype, chunks, partial, ctx, stream);
  return launch_reduce<float, 2>(x, B, V, C, cs, co, dtype, chunks, partial, ctx, stream);
}

extern "C" int coma_gate_stats(const void* x, int32_t B, int64_t V, int32_t C, int32_t cs, int32_t co, int32_t dtype,
                               float* partial, coma_stream_t stream) {
  return coma_norm_stats(x, B, V, C, cs, co, dtype, partial, stream);
}

extern "C" int coma_norm_stats_finalize(const coma_norm_finalize_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->A && a->S && a->mean && a->rstd && a->B > 0 && a->C > 0, "coma_norm_stats_finalize: bad arguments");
  unsigned blocks;
  if (a->mode == COMA_NORM_NONE || a->mode == COMA_NORM_GIVEN) {
    COMA_CHECK_ARG(a->mode == COMA_NORM_NONE || (a->given_mean && a->given_var), "finalize: GIVEN needs mean/var");
    blocks = (unsigned)((a->B * a->C + kThreads - 1) / kThreads);
  } else {
    COMA_CHECK_ARG(a->partial && a->chunks > 0 && a->V > 0, "finalize: partial sums missing");
    blocks = (unsigned)(a->mode == COMA_NORM_BATCH ? a->C : a->B * a->C);
  }
  finalize_kernel<<<blocks, kThreads, 0, stream>>>(*a);
  COMA_CHECK_LAUNCH("norm_finalize");
  return COMA_OK;
}

extern "C" int coma_norm_film_act_fwd(const coma_affine_act_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->x && a->y && a->A && a->S && a->B > 0 && a->C > 0 && a->V > 0, "coma_norm_film_act_fwd: bad arguments");
  bool vec = vec_ok(a->x, a->C, a->x_cs, a->x_co, a->dtype) && vec_ok(a->y, a->C, a->y_cs, a->y_co, a->dtype) &&
             (!a->r || vec_ok(a->r, a->C, a->r_cs, 0, a->dtype));
  coma_affine_act_args v = *a;
  int creal = 0;
  if (!vec && a->C < 8 && 8 % a->C == 0 && a->x_cs == a->C && a->y_cs == a->C && a->x_co == 0 && a->y_co == 0 &&
      (!a->r || a->r_cs == a->C) && (a->V * a->C) % 8 == 0) {
    // dense tensor with 1 / 2 / 4 channels (the modulator heads): sweep it as 8-wide vectors, coefficients repeat every C
    v.C = 8; v.V = a->V * a->C / 8; v.x_cs = v.y_cs = 8; v.r_cs = a->r ? 8 : 0;
    if (vec_ok(v.x, 8, 8, 0, v.dtype) && vec_ok(v.y, 8, 8, 0, v.dtype) && (!v.r || vec_ok(v.r, 8, 8, 0, v.dtype))) {
      vec = true;
      creal = a->C;
    } else {
      v = *a;
    }
  }
  if (vec) {
    static const int env_chunks = [] { const char* e = getenv("COMA_AFFINE_CHUNKS"); return e ? atoi(e) : 0; }();
    int chunks = (int)std::min<int64_t>(std::max<int64_t>((v.V * (v.C / 8) + kThreads * 8 - 1) / (kThreads * 8), 1), 4096);
    if (env_chunks > 0) chunks = env_chunks;
    dim3 grid((unsigned)chunks, (unsigned)a->B);
    const bool simple = a->act == COMA_ACT_NONE || a->act == COMA_ACT_RELU || a->act == COMA_ACT_LEAKY;
#define COMA_AFFINE_LAUNCH(T, U)                                                                                     \
    do {                                                                                                              \
      if (simple && !v.r) affine_act_vec_kernel<T, U, true, false><<<grid, kThreads, 0, stream>>>(v, chunks, creal);   \
      else if (simple) affine_act_vec_kernel<T, U, true, true><<<grid, kThreads, 0, stream>>>(v, chunks, creal);        \
      else if (!v.r) affine_act_vec_kernel<T, U, false, false><<<grid, kThreads, 0, stream>>>(v, chunks, creal);       \
      else affine_act_vec_kernel<T, U, false, true><<<grid, kThreads, 0, stream>>>(v, chunks, creal);                   \
    } while (0)
    if (a->dtype == COMA_BF16) COMA_AFFINE_LAUNCH(__nv_bfloat16, 2);
    else COMA_AFFINE_LAUNCH(float, 1);
#undef COMA_AFFINE_LAUNCH
  } else {
    COMA_CHECK_ARG(a->C <= 64, "coma_norm_film_act_fwd: unaligned C=%d too large for the scalar path", a->C);
    dim3 grid((unsigned)std::min<int64_t>((a->V + kThreads - 1) / kThreads, 2048), (unsigned)a->B);
    if (a->dtype == COMA_BF16) affine_act_small_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(*a);
    else affine_act_small_kernel<float><<<grid, kThreads, 0, stream>>>(*a);
  }
  COMA_CHECK_LAUNCH("affine_act_fwd");
  return COMA_OK;
}

extern "C" int coma_norm_film_act_bwd(const coma_affine_act_bwd_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->x && a->dy && a->dx && a->A && a->S && a->mean && a->rstd && a->partial && a->dg && a->dh && a->coef,
                 "coma_norm_film_act_bwd: bad arguments");
  const int chunks = stats_chunks(a->V);
  if (bulk_ok(a)) {
    int rc = launch_bulk<0>(a, chunks, stream);
    if (rc) return rc;
    bwd_finalize_kernel<<<(unsigned)a->C, kThreads, 0, stream>>>(*a, chunks);
    COMA_CHECK_LAUNCH("norm_bwd_finalize");
    // apply sweep: enough blocks for ~4 waves of two blocks per SM
    const int ach = std::max(1, (4 * 2 * num_sms() + a->B - 1) / a->B);
    return launch_bulk<1>(a, ach, stream);
  }
  BwdCtx ctx{a->dy, a->A, a->S, a->mean, a->rstd, a->slope, a->dy_cs, a->dy_co, a->act, a->r, a->r_cs};
  int rc;
  if (a->dtype == COMA_BF16)
    rc = launch_reduce<__nv_bfloat16, 3>(a->x, a->B, a->V, a->C, a->x_cs, a->x_co, a->dtype, chunks, a->partial, ctx, stream);
  else
    rc = launch_reduce<float, 3>(a->x, a->B, a->V, a->C, a->x_cs, a->x_co, a->dtype, chunks, a->partial, ctx, stream);
  if (rc) return rc;
  bwd_finalize_kernel<<<(unsigned)a->C, kThreads, 0, stream>>>(*a, chunks);
  COMA_CHECK_LAUNCH("norm_bwd_finalize");
  const bool vec = vec_ok(a->x, a->C, a->x_cs, a->x_co, a->dtype) && vec_ok(a->dy, a->C, a->dy_cs, a->dy_co, a->dtype) &&
                   vec_ok(a->dx, a->C, a->dx_cs, a->dx_co, a->dtype) && (!a->r || vec_ok(a->r, a->C, a->r_cs, 0, a->dtype)) &&
                   (!a->dr || vec_ok(a->dr, a->C, a->dr_cs, 0, a->dtype));
  if (vec) {
    const int ach = (int)std::min<int64_t>(std::max<int64_t>((a->V * (a->C / 8) + kThreads * 8 - 1) / (kThreads * 8), 1), 4096);
    dim3 grid((unsigned)ach, (unsigned)a->B);
    const bool simple = a->act == COMA_ACT_NONE || a->act == COMA_ACT_RELU || a->act == COMA_ACT_LEAKY;
    if (a->dtype == COMA_BF16) {
      if (simple) bwd_apply_vec_kernel<__nv_bfloat16, true><<<grid, kThreads, 0, stream>>>(*a, ach);
      else bwd_apply_vec_kernel<__nv_bfloat16, false><<<grid, kThreads, 0, stream>>>(*a, ach);
    } else {
      if (simple) bwd_apply_vec_kernel<float, true><<<grid, kThreads, 0, stream>>>(*a, ach);
      else bwd_apply_vec_kernel<float, false><<<grid, kThreads, 0, stream>>>(*a, ach);
    }
  } else if (c1_ok(a->x, a->C, a->x_cs, a->x_co, a->V, 1, a->dtype) && c1_ok(a->dy, a->C, a->dy_cs, a->dy_co, a->V, 1, a->dtype) &&
             c1_ok(a->dx, a->C, a->dx_cs, a->dx_co, a->V, 1, a->dtype) && (!a->r || c1_ok(a->r, a->C, a->r_cs, 0, a->V, 1, a->dtype)) &&
             (!a->dr || c1_ok(a->dr, a->C, a->dr_cs, 0, a->V, 1, a->dtype))) {
    dim3 grid((unsigned)std::min<int64_t>((a->V / 8 + kThreads - 1) / kThreads, 2048), (unsigned)a->B);
    if (a->dtype == COMA_BF16) bwd_apply_c1_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(*a);
    else bwd_apply_c1_kernel<float><<<grid, kThreads, 0, stream>>>(*a);
  } else {
    COMA_CHECK_ARG(a->C <= 64, "coma_norm_film_act_bwd: unaligned C=%d too large for the scalar path", a->C);
    dim3 grid((unsigned)std::min<int64_t>((a->V + kThreads - 1) / kThreads, 2048), (unsigned)a->B);
    if (a->dtype == COMA_BF16) bwd_apply_small_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(*a);
    else bwd_apply_small_kernel<float><<<grid, kThreads, 0, stream>>>(*a);
  }
  COMA_CHECK_LAUNCH("affine_act_bwd");
  return COMA_OK;
}
