// ROI painting, prompt packing and the RoiMSE loss: the HBM-bound glue around the modulator stacks.
//  - coma_roi_paint     replaces the B x 36 masked index_put_ loop + .item() syncs of
//                       forward_modulator_with_uq (attn_unet_data_parallel.py:632-649)
//  - coma_pack2_*       replace `general_prompt + ...` and the torch.cat calls (:651,654)
//  - coma_roi_mse_*     replace RoiMSE.forward (criterions.py:181-211; ~40 ATen kernels + 2 syncs)
#include <algorithm>

#include "common.cuh"

namespace coma {

constexpr int kMaxRoi = 64;

// 16 bf16 channels of one voxel, [c0, c1, c2, 0, ...]: one 256-bit store = one full 32-byte sector (sm_100 st.global.v8)
__device__ __forceinline__ void store16_bf16_head(__nv_bfloat16* o, float c0, float c1, float c2) {
  const __nv_bfloat162 p01 = __floats2bfloat162_rn(c0, c1), p23 = __floats2bfloat162_rn(c2, 0.f);
  const uint32_t w0 = *reinterpret_cast<const uint32_t*>(&p01), w1 = *reinterpret_cast<const uint32_t*>(&p23);
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %3, %3, %3, %3, %3};" ::"l"(o), "r"(w0), "r"(w1), "r"(0u) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(256) roi_paint_kernel(coma_roi_paint_args a) {
  __shared__ float ids[kMaxRoi];
  __shared__ float lut[kMaxRoi * 2];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < a.n_roi; i += 256) {
    ids[i] = (float)a.roi_ids[i];
    lut[2 * i] = a.lut[((int64_t)b * a.n_roi + i) * 2];
    lut[2 * i + 1] = a.lut[((int64_t)b * a.n_roi + i) * 2 + 1];
  }
  __syncthreads();
  const bool pos = a.is_pos[b] == 1.0f;
  const float* prompt = pos ? a.pos_prompt : a.neg_prompt;
  T* ob = static_cast<T*>(a.out) + (int64_t)b * a.V * a.out_cs;
  for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < a.V; v += (int64_t)gridDim.x * 256) {
    const float label = __ldg(a.roi + (int64_t)b * a.V + v);
    float loc = 0.f, sd = 0.f;
    if (!(__ldg(a.mri + (int64_t)b * a.V + v) < 1e-4f)) {
      for (int i = 0; i < a.n_roi; ++i)
        if (label == ids[i]) { loc = lut[2 * i]; sd = lut[2 * i + 1]; }
    }
    T* o = ob + v * a.out_cs;
    if (sizeof(T) == 2 && a.out_cs == 16 && (reinterpret_cast<uintptr_t>(o) & 31) == 0) {
      store16_bf16_head(reinterpret_cast<__nv_bfloat16*>(o), __ldg(prompt + v), sd, loc);
    } else if ((a.out_cs & 7) == 0) {   // 16-byte vector stores: [prompt, saliency, suvr, 0, ...]
      float vals[8] = {__ldg(prompt + v), sd, loc, 0.f, 0.f, 0.f, 0.f, 0.f};
      store8(o, vals);
      const float zeros[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int c = 8; c < a.out_cs; c += 8) store8(o + c, zeros);
    } else if (sizeof(T) == 2 && a.out_cs == 4) {   // [prompt, saliency, suvr, 0] as one 8-byte store (tap-packed consumer conv)
      const __nv_bfloat162 p01 = __floats2bfloat162_rn(__ldg(prompt + v), sd), p23 = __floats2bfloat162_rn(loc, 0.f);
      uint2 q;
      q.x = *reinterpret_cast<const uint32_t*>(&p01);
      q.y = *reinterpret_cast<const uint32_t*>(&p23);
      *reinterpret_cast<uint2*>(o) = q;
    } else {
      Elem<T>::st(o, __ldg(prompt + v));
      if (a.out_cs > 1) Elem<T>::st(o + 1, sd);
      if (a.out_cs > 2) Elem<T>::st(o + 2, loc);
      for (int c = 3; c < a.out_cs; ++c) Elem<T>::st(o + c, 0.f);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) roi_paint_bwd_kernel(const T* __restrict__ dbuf, const float* __restrict__ is_pos,
                                                            float* __restrict__ dpos, float* __restrict__ dneg, int B,
                                                            int64_t V, int cs) {
  for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < V; v += (int64_t)gridDim.x * 256) {
    float p = 0.f, n = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = Elem<T>::ld(dbuf + ((int64_t)b * V + v) * cs);
      if (is_pos[b] == 1.0f) p += d; else n += d;
    }
    dpos[v] = p;
    dneg[v] = n;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pack2_kernel(coma_pack2_args a) {
  const int b = blockIdx.y;                       // one sample per grid row: no 64-bit modulo per voxel
  const T* pa = static_cast<const T*>(a.a) + (int64_t)b * a.V;
  const T* pb = a.b ? static_cast<const T*>(a.b) + (int64_t)b * a.V : nullptr;
  T* d = static_cast<T*>(a.dst) + (int64_t)b * a.V * a.dst_cs;
  const bool vec = (a.dst_cs & 7) == 0;
  constexpr int U = 4;
  for (int64_t vb = (int64_t)blockIdx.x * 256 + threadIdx.x; vb < a.V; vb += (int64_t)gridDim.x * 256 * U) {
    float va[U], vbv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = vb + (int64_t)u * gridDim.x * 256;
      va[u] = vbv[u] = 0.f;
      if (v < a.V) {
        va[u] = Elem<T>::ld(pa + v) + (a.a_add ? __ldg(a.a_add + v) : 0.f);
        if (pb) vbv[u] = Elem<T>::ld(pb + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = vb + (int64_t)u * gridDim.x * 256;
      if (v >= a.V) continue;
      T* o = d + v * a.dst_cs;
      if (sizeof(T) == 2 && a.dst_cs == 16 && (reinterpret_cast<uintptr_t>(o) & 31) == 0) {
        store16_bf16_head(reinterpret_cast<__nv_bfloat16*>(o), va[u], vbv[u], 0.f);
      } else if (vec) {
        float vals[8] = {va[u], vbv[u], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store8(o, vals);
        const float zeros[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c = 8; c < a.dst_cs; c += 8) store8(o + c, zeros);
      } else if (sizeof(T) == 2 && a.dst_cs == 2) {
        const __nv_bfloat162 pr = __floats2bfloat162_rn(va[u], vbv[u]);
        *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<const uint32_t*>(&pr);
      } else {
        Elem<T>::st(o, va[u]);
        Elem<T>::st(o + 1, vbv[u]);
        for (int c = 2; c < a.dst_cs; ++c) Elem<T>::st(o + c, 0.f);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) unpack2_kernel(coma_unpack2_args a) {
  const T* d = static_cast<const T*>(a.ddst);
  T* da = static_cast<T*>(a.da);
  T* db = static_cast<T*>(a.db);
  for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < a.V; v += (int64_t)gridDim.x * 256) {
    float s = 0.f;
    for (int b = 0; b < a.B; ++b) {
      const int64_t i = (int64_t)b * a.V + v;
      const float g0 = Elem<T>::ld(d + i * a.dst_cs), g1 = db ? Elem<T>::ld(d + i * a.dst_cs + 1) : 0.f;
      if (da) Elem<T>::st(da + i, g0);
      if (db) Elem<T>::st(db + i, g1);
      s += g0;
    }
    if (a.d_a_add) a.d_a_add[v] = s;
  }
}

// ---- RoiMSE ------------------------------------------------------------------------------------
static int mse_chunks(int64_t V) {
  int64_t c = (V + 4095) / 4096;
  return (int)(c < 1 ? 1 : (c > 512 ? 512 : c));
}

template <typename T>
__global__ void __launch_bounds__(256) roi_mse_partial_kernel(coma_roi_mse_args a, int chunks) {
  __shared__ float ids[kMaxRoi], wts[kMaxRoi];
  __shared__ float red[2][8];
  const int b = blockIdx.y, chunk = blockIdx.x;
  for (int i = threadIdx.x; i < a.n_roi; i += 256) {
    ids[i] = (float)a.roi_ids[i];
    wts[i] = a.roi_w[i];
  }
  __syncthreads();
  const int64_t per = (a.V + chunks - 1) / chunks, v0 = (int64_t)chunk * per, v1 = min(v0 + per, a.V);
  const T* pred = static_cast<const T*>(a.pred) + (int64_t)b * a.V;
  float sq = 0.f, mk = 0.f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += 256) {
    const float d = Elem<T>::ld(pred + v) - __ldg(a.gt + (int64_t)b * a.V + v);
    sq = fmaf(d, d, sq);
    const float label = __ldg(a.roi + (int64_t)b * a.V + v);
    float w = 0.f;
    for (int i = 0; i < a.n_roi; ++i)
      if (label == ids[i]) w = wts[i];
    mk += w;
  }
  sq = warp_sum(sq);
  mk = warp_sum(mk);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sq; red[1][threadIdx.x >> 5] = mk; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    a.partial[((int64_t)b * chunks + chunk) * 2 + threadIdx.x] = s;
  }
}

__global__ void roi_mse_finalize_kernel(coma_roi_mse_args a, int chunks) {
  const int b = blockIdx.x;
  double sq = 0.0, mk = 0.0;
  for (int i = threadIdx.x; i < chunks; i += 32) {
    sq += a.partial[((int64_t)b * chunks + i) * 2];
    mk += a.partial[((int64_t)b * chunks + i) * 2 + 1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    mk += __shfl_xor_sync(0xffffffffu, mk, o);
  }
  if (threadIdx.x == 0) {
    a.sums[2 * b] = (float)sq;
    a.sums[2 * b + 1] = (float)mk;
    a.loss[b] = (float)((mk / (double)a.V) * (sq / (double)a.V));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) roi_mse_bwd_kernel(coma_roi_mse_args a) {
  const int b = blockIdx.y;
  const float invV = 1.f / (float)a.V;
  const float k = a.dloss[b] * a.sums[2 * b + 1] * invV * 2.f * invV;
  const T* pred = static_cast<const T*>(a.pred) + (int64_t)b * a.V;
  T* dp = static_cast<T*>(a.dpred) + (int64_t)b * a.V;
  for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < a.V; v += (int64_t)gridDim.x * 256)
    Elem<T>::st(dp + v, k * (Elem<T>::ld(pred + v) - __ldg(a.gt + (int64_t)b * a.V + v)));
}

static unsigned sweep_blocks(int64_t n) { return (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 8); }

}  // namespace coma

using namespace coma;

extern "C" int coma_roi_paint(const coma_roi_paint_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->roi && a->mri && a->lut && a->roi_ids && a->is_pos && a->pos_prompt && a->neg_prompt && a->out,
                 "coma_roi_paint: null argument");
  COMA_CHECK_ARG(a->n_roi > 0 && a->n_roi <= kMaxRoi && a->out_cs >= 3, "coma_roi_paint: n_roi in 1..64, out_cs >= 3");
  dim3 grid(sweep_blocks(a->V), (unsigned)a->B);
  if (a->dtype == COMA_BF16) roi_paint_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(*a);
  else roi_paint_kernel<float><<<grid, 256, 0, stream>>>(*a);
  COMA_CHECK_LAUNCH("roi_paint");
  return COMA_OK;
}

extern "C" int coma_roi_paint_bwd(const void* dbuf, const float* is_pos, float* dpos, float* dneg, int32_t B, int64_t V,
                                  int32_t cs, int32_t dtype, coma_stream_t stream) {
  COMA_CHECK_ARG(dbuf && is_pos && dpos && dneg && B > 0 && V > 0, "coma_roi_paint_bwd: bad arguments");
  if (dtype == COMA_BF16)
    roi_paint_bwd_kernel<__nv_bfloat16><<<sweep_blocks(V), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dbuf), is_pos, dpos, dneg, B, V, cs);
  else
    roi_paint_bwd_kernel<float><<<sweep_blocks(V), 256, 0, stream>>>(static_cast<const float*>(dbuf), is_pos, dpos, dneg, B, V, cs);
  COMA_CHECK_LAUNCH("roi_paint_bwd");
  return COMA_OK;
}

extern "C" int coma_pack2_fwd(const coma_pack2_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->a && a->dst && a->dst_cs >= 2, "coma_pack2_fwd: bad arguments");   // b may be NULL (zeros)
  const int64_t want = (a->V + 1023) / 1024, cap = std::max<int64_t>(1, (int64_t)num_sms() * 8 / std::max(1, a->B));
  const dim3 grid((unsigned)std::max<int64_t>(1, std::min(want, cap)), (unsigned)a->B);
  if (a->dtype == COMA_BF16) pack2_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(*a);
  else pack2_kernel<float><<<grid, 256, 0, stream>>>(*a);
  COMA_CHECK_LAUNCH("pack2");
  return COMA_OK;
}

extern "C" int coma_pack2_bwd(const coma_unpack2_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->ddst && a->dst_cs >= 2, "coma_pack2_bwd: bad arguments");
  if (a->dtype == COMA_BF16) unpack2_kernel<__nv_bfloat16><<<sweep_blocks(a->V), 256, 0, stream>>>(*a);
  else unpack2_kernel<float><<<sweep_blocks(a->V), 256, 0, stream>>>(*a);
  COMA_CHECK_LAUNCH("unpack2");
  return COMA_OK;
}

extern "C" int coma_roi_mse_chunks(int64_t V) { return mse_chunks(V); }

extern "C" int coma_roi_mse_fwd(const coma_roi_mse_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->pred && a->gt && a->roi && a->roi_ids && a->roi_w && a->partial && a->loss && a->sums,
                 "coma_roi_mse_fwd: null argument");
  COMA_CHECK_ARG(a->n_roi > 0 && a->n_roi <= kMaxRoi, "coma_roi_mse_fwd: n_roi in 1..64");
  const int chunks = mse_chunks(a->V);
  dim3 grid((unsigned)chunks, (unsigned)a->B);
  if (a->dtype == COMA_BF16) roi_mse_partial_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(*a, chunks);
  else roi_mse_partial_kernel<float><<<grid, 256, 0, stream>>>(*a, chunks);
  COMA_CHECK_LAUNCH("roi_mse_partial");
  roi_mse_finalize_kernel<<<(unsigned)a->B, 32, 0, stream>>>(*a, chunks);
  COMA_CHECK_LAUNCH("roi_mse_finalize");
  return COMA_OK;
}

extern "C" int coma_roi_mse_bwd(const coma_roi_mse_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->pred && a->gt && a->sums && a->dloss && a->dpred, "coma_roi_mse_bwd: null argument");
  dim3 grid(sweep_blocks(a->V), (unsigned)a->B);
  if (a->dtype == COMA_BF16) roi_mse_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(*a);
  else roi_mse_bwd_kernel<float><<<grid, 256, 0, stream>>>(*a);
  COMA_CHECK_LAUNCH("roi_mse_bwd");
  return COMA_OK;
}

// ---- fused evaluation metrics (include/coma_b200.h) ------------------------------------------------------------------
constexpr int kEvalLut = 4096, kEvalQ = 8;

__global__ void __launch_bounds__(256) eval_metrics_kernel(coma_eval_metrics_args a) {
  __shared__ int8_t lut[kEvalLut];                    // label value -> ROI slot (-1: not a listed ROI)
  __shared__ double acc[(kMaxRoi + 1) * kEvalQ];
  const int b = blockIdx.y, R = a.n_roi;
  for (int i = threadIdx.x; i < kEvalLut; i += 256) lut[i] = -1;
  for (int i = threadIdx.x; i < (R + 1) * kEvalQ; i += 256) acc[i] = 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < R; i += 256) {
    const int id = a.roi_ids[i];
    if (id >= 0 && id < kEvalLut) lut[id] = (int8_t)i;
  }
  __syncthreads();
  const float* pp = a.pred + (int64_t)b * a.V;
  const float* tp = a.tau + (int64_t)b * a.V;
  const float* rp = a.roi + (int64_t)b * a.V;
  float g_ad = 0.f, g_d2 = 0.f, g_t = 0.f, g_t2 = 0.f, g_r = 0.f, g_nan = 0.f, g_m = 0.f, g_n = 0.f;
  for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < a.V; v += (int64_t)gridDim.x * 256) {
    const float p = __ldg(pp + v), t = __ldg(tp + v), lab = __ldg(rp + v);
    const float d = p - t, ad = fabsf(d), d2 = d * d, ratio = fabsf(d / t);
    const bool isn = ratio != ratio;
    g_n += 1.f; g_ad += ad; g_d2 += d2; g_t += t; g_t2 = fmaf(t, t, g_t2);
    if (isn) g_nan += 1.f; else g_r += ratio;
    if (fabsf(t) > 1e-8f) g_m += 100.f * ratio;
    const int li = (int)lab;
    if (lab >= 0.f && lab < (float)kEvalLut && (float)li == lab) {
      const int slot = lut[li];
      if (slot >= 0) {
        double* s = acc + slot * kEvalQ;
        atomicAdd(s + 0, 1.0); atomicAdd(s + 1, (double)ad); atomicAdd(s + 2, (double)d2); atomicAdd(s + 3, (double)t);
        atomicAdd(s + 4, (double)t * t);
        if (isn) atomicAdd(s + 6, 1.0); else atomicAdd(s + 5, (double)ratio);
      }
    }
  }
  // all-voxel slot: registers -> warp -> shared
  const float gq[kEvalQ] = {g_n, g_ad, g_d2, g_t, g_t2, g_r, g_nan, g_m};
#pragma unroll
  for (int q = 0; q < kEvalQ; ++q) {
    const float w = warp_sum(gq[q]);
    if ((threadIdx.x & 31) == 0 && w != 0.f) atomicAdd(acc + R * kEvalQ + q, (double)w);
  }
  __syncthreads();
  double* ob = a.out + (int64_t)b * (R + 1) * kEvalQ;
  for (int i = threadIdx.x; i < (R + 1) * kEvalQ; i += 256)
    if (acc[i] != 0.0) atomicAdd(ob + i, acc[i]);
}

extern "C" int coma_eval_metrics(const coma_eval_metrics_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->pred && a->tau && a->roi && a->roi_ids && a->out && a->B > 0 && a->V > 0, "coma_eval_metrics: bad arguments");
  COMA_CHECK_ARG(a->n_roi >= 0 && a->n_roi <= kMaxRoi, "coma_eval_metrics: at most %d ROIs", kMaxRoi);
  const int64_t want = (a->V + 256 * 16 - 1) / (256 * 16), cap = std::max<int64_t>(1, (int64_t)num_sms() * 4 / a->B);
  const dim3 grid((unsigned)std::max<int64_t>(1, std::min(want, cap)), (unsigned)a->B);
  eval_metrics_kernel<<<grid, 256, 0, stream>>>(*a);
  COMA_CHECK_LAUNCH("eval_metrics");
  return COMA_OK;
}

// ---- GPU-side input preparation (include/coma_b200.h) -------------------------------------------------------------------
__device__ __forceinline__ float nan_to_num(float v) {     // torch.nan_to_num defaults
  if (v != v) return 0.f;
  if (v == __int_as_float(0x7f800000)) return 3.402823466e+38f;
  if (v == __int_as_float(0xff800000)) return -3.402823466e+38f;
  return v;
}

__global__ void __launch_bounds__(256) prepare_volumes_kernel(coma_prepare_args a) {
  const int64_t total = (int64_t)a.out_size[0] * a.out_size[1] * a.out_size[2];
  for (int64_t o = (int64_t)blockIdx.x * 256 + threadIdx.x; o < total; o += (int64_t)gridDim.x * 256) {
    const int ox = (int)(o % a.out_size[2]), t = (int)(o / a.out_size[2]);
    const int oc[3] = {t / a.out_size[1], t % a.out_size[1], ox};
    bool padded = false, inside = true;
    int64_t src = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      const int r = oc[ax] - a.pad_before[ax];
      if (r < 0 || r >= a.res_size[ax]) padded = true;
      const double c = (double)r * a.ratio[ax];             // ITK maps indices through physical space in double
      if (!(c >= -0.5 && c < (double)a.in_size[ax] - 0.5)) inside = false;
      int n = (int)floor(c + 0.5);
      n = n < 0 ? 0 : (n >= a.in_size[ax] ? a.in_size[ax] - 1 : n);
      src = src * a.in_size[ax] + n;
    }
    float m = 0.f, tv = 0.f, rv = 0.f;
    if (!padded) {
      if (a.mri) m = nan_to_num(inside ? __ldg(a.mri + src) : a.default_value);
      if (a.tau) tv = nan_to_num(inside ? __ldg(a.tau + src) : a.default_value);
      if (a.roi) rv = nan_to_num(inside ? __ldg(a.roi + src) : a.default_value);
    }
    if (a.mri && a.roi && rv == 0.f) m = 0.f;              // mri[roi == 0] = 0
    if (a.mri_out) a.mri_out[o] = m;
    if (a.tau_out) a.tau_out[o] = tv;
    if (a.roi_out) a.roi_out[o] = rv;
  }
}

extern "C" int coma_prepare_volumes(const coma_prepare_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && (a->mri || a->tau || a->roi), "coma_prepare_volumes: no input");
  COMA_CHECK_ARG((!a->mri || a->mri_out) && (!a->tau || a->tau_out) && (!a->roi || a->roi_out), "coma_prepare_volumes: missing output");
  for (int ax = 0; ax < 3; ++ax)
    COMA_CHECK_ARG(a->in_size[ax] > 0 && a->res_size[ax] > 0 && a->out_size[ax] > 0 && a->pad_before[ax] >= 0 && a->ratio[ax] > 0.0,
                   "coma_prepare_volumes: bad geometry on axis %d", ax);
  const int64_t total = (int64_t)a->out_size[0] * a->out_size[1] * a->out_size[2];
  prepare_volumes_kernel<<<sweep_blocks(total), 256, 0, stream>>>(*a);
  COMA_CHECK_LAUNCH("prepare_volumes");
  return COMA_OK;
}

namespace coma {
struct SmallUpload { float v[960]; };
__global__ void upload_small_kernel(float* __restrict__ dst, const SmallUpload s, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = s.v[i];
}
}  // namespace coma

extern "C" int coma_upload_small(float* dst, const float* host_values, int64_t n, coma_stream_t stream) {
  COMA_CHECK_ARG(dst && host_values && n >= 0, "coma_upload_small: bad arguments");
  for (int64_t off = 0; off < n; off += 960) {
    coma::SmallUpload s;
    const int m = (int)std::min<int64_t>(960, n - off);
    memcpy(s.v, host_values + off, (size_t)m * sizeof(float));
    coma::upload_small_kernel<<<(m + 255) / 256, 256, 0, stream>>>(dst + off, s, m);
    COMA_CHECK_LAUNCH("upload_small");
  }
  return COMA_OK;
}
