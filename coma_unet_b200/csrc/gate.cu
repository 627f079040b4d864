// Additive attention gate (reference: ObservableAttentionBlock.forward, attn_unet_data_parallel.py:139-150;
// MONAI attentionunet.AttentionBlock):  out = x * sigmoid(psi(relu(W_g g + W_x x))).
//
// coma_gate_fwd: ONE kernel for folded BatchNorm (eval): reads g and x once, writes out once
// (3*C*s bytes per voxel, HBM-bound) instead of ~10 library kernels.  The two 1x1x1 convolutions are
// C x C/2 mat-vecs per voxel kept in registers; weights are broadcast from shared memory.
// coma_gate_apply_fwd / coma_gate_bwd: out = x * p with p one value per voxel, and its backward.
#include "common.cuh"

namespace coma {

template <typename T, int C>
__global__ void __launch_bounds__(128) gate_fused_kernel(coma_gate_args a) {
  constexpr int F = C / 2;
  __shared__ __align__(16) float swg[F * C];
  __shared__ __align__(16) float swx[F * C];
  __shared__ float sb[F], sp[F];
  for (int i = threadIdx.x; i < F * C; i += 128) {
    swg[i] = a.wg[i];
    swx[i] = a.wx[i];
  }
  for (int i = threadIdx.x; i < F; i += 128) {
    sb[i] = a.bsum[i];
    sp[i] = a.wpsi[i];
  }
  __syncthreads();
  const float bpsi = a.bpsi_ptr ? __ldg(a.bpsi_ptr) : a.bpsi;
  const int64_t total = (int64_t)a.B * a.V;
  const T* gp = static_cast<const T*>(a.g) + a.g_co;
  const T* xp = static_cast<const T*>(a.x) + a.x_co;
  T* op = static_cast<T*>(a.out) + a.out_co;
  for (int64_t v = (int64_t)blockIdx.x * 128 + threadIdx.x; v < total; v += (int64_t)gridDim.x * 128) {
    float gv[C], xv[C];
#pragma unroll
    for (int c = 0; c < C; c += 8) {
      load8(gp + v * a.g_cs + c, *reinterpret_cast<float(*)[8]>(&gv[c]));
      load8(xp + v * a.x_cs + c, *reinterpret_cast<float(*)[8]>(&xv[c]));
    }
    float q = bpsi;
#pragma unroll 2
    for (int j = 0; j < F; ++j) {
      float t0 = sb[j], t1 = 0.f;
      const float4* wg4 = reinterpret_cast<const float4*>(&swg[j * C]);
      const float4* wx4 = reinterpret_cast<const float4*>(&swx[j * C]);
#pragma unroll
      for (int c = 0; c < C / 4; ++c) {
        const float4 wa = wg4[c], wb = wx4[c];
        t0 = fmaf(wa.x, gv[4 * c], t0); t1 = fmaf(wb.x, xv[4 * c], t1);
        t0 = fmaf(wa.y, gv[4 * c + 1], t0); t1 = fmaf(wb.y, xv[4 * c + 1], t1);
        t0 = fmaf(wa.z, gv[4 * c + 2], t0); t1 = fmaf(wb.z, xv[4 * c + 2], t1);
        t0 = fmaf(wa.w, gv[4 * c + 3], t0); t1 = fmaf(wb.w, xv[4 * c + 3], t1);
      }
      q = fmaf(sp[j], fmaxf(t0 + t1, 0.f), q);
    }
    const float att = 1.f / (1.f + __expf(-q));
#pragma unroll
    for (int c = 0; c < C; ++c) xv[c] *= att;
#pragma unroll
    for (int c = 0; c < C; c += 8) store8(op + v * a.out_cs + c, *reinterpret_cast<float(*)[8]>(&xv[c]));
    if (a.psi_out) Elem<T>::st(static_cast<T*>(a.psi_out) + v, att);
  }
}

template <typename T>
static int launch_gate(const coma_gate_args& a, cudaStream_t stream) {
  const int64_t total = (int64_t)a.B * a.V;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 127) / 128, (int64_t)num_sms() * 16);
  switch (a.C) {
    case 8: gate_fused_kernel<T, 8><<<blocks, 128, 0, stream>>>(a); break;
    case 16: gate_fused_kernel<T, 16><<<blocks, 128, 0, stream>>>(a); break;
    case 32: gate_fused_kernel<T, 32><<<blocks, 128, 0, stream>>>(a); break;
    case 64: gate_fused_kernel<T, 64><<<blocks, 128, 0, stream>>>(a); break;
    default:
      set_error("coma_gate_fwd: fused kernel takes C in {8,16,32,64}, got %d (compose it from the k=1 conv path)", a.C);
      return COMA_ERR_UNSUPPORTED;
  }
  COMA_CHECK_LAUNCH("gate_fused");
  return COMA_OK;
}

// out = x * p  (p: one value per voxel)
template <typename T>
__global__ void __launch_bounds__(256) bcast_mul_fwd_kernel(coma_bcast_mul_args a) {
  const int CV = a.C >> 3;
  const int64_t total = (int64_t)a.B * a.V * CV;
  const T* xp = static_cast<const T*>(a.x) + a.x_co;
  const T* pp = static_cast<const T*>(a.p);
  T* op = static_cast<T*>(a.out) + a.out_co;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t v = i / CV;
    const int c = (int)(i % CV) * 8;
    float xv[8];
    load8(xp + v * a.x_cs + c, xv);
    const float p = Elem<T>::ld(pp + v);
#pragma unroll
    for (int e = 0; e < 8; ++e) xv[e] *= p;
    store8(op + v * a.out_cs + c, xv);
  }
}

// dx = dout * p ; dp = sum_c dout * x.   CV (= C/8, power of two <= 32) adjacent lanes share a voxel.
template <typename T>
__global__ void __launch_bounds__(256) bcast_mul_bwd_kernel(coma_bcast_mul_args a) {
  const int CV = a.C >> 3;
  const int64_t total = (int64_t)a.B * a.V * CV;
  const int64_t padded = (total + 31) / 32 * 32;
  const T* xp = static_cast<const T*>(a.x) + a.x_co;
  const T* pp = static_cast<const T*>(a.p);
  const T* dop = static_cast<const T*>(a.dout) + a.out_co;
  T* dxp = static_cast<T*>(a.dx) + a.x_co;
  T* dpp = static_cast<T*>(a.dp);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < padded; i += (int64_t)gridDim.x * 256) {
    float part = 0.f;
    const bool on = i < total;
    const int64_t v = on ? i / CV : 0;
    if (on) {
      const int c = (int)(i % CV) * 8;
      float xv[8], dv[8];
      load8(xp + v * a.x_cs + c, xv);
      load8(dop + v * a.out_cs + c, dv);
      const float p = Elem<T>::ld(pp + v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        part = fmaf(dv[e], xv[e], part);
        dv[e] *= p;
      }
      store8(dxp + v * a.x_cs + c, dv);
    }
    for (int o = CV >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (on && (i % CV) == 0) Elem<T>::st(dpp + v, part);
  }
}

}  // namespace coma

using namespace coma;

extern "C" int coma_gate_fwd(const coma_gate_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->g && a->x && a->out && a->wg && a->wx && a->bsum && a->wpsi, "coma_gate_fwd: null argument");
  COMA_CHECK_ARG(a->F * 2 == a->C, "coma_gate_fwd: F must be C/2");
  COMA_CHECK_ARG(a->g_cs % 8 == 0 && a->g_co % 8 == 0 && a->x_cs % 8 == 0 && a->x_co % 8 == 0 && a->out_cs % 8 == 0 &&
                     a->out_co % 8 == 0, "coma_gate_fwd: channel strides/offsets must be multiples of 8");
  if (a->dtype == COMA_BF16) return launch_gate<__nv_bfloat16>(*a, stream);
  return launch_gate<float>(*a, stream);
}

static int check_bcast(const coma_bcast_mul_args* a, const char* who) {
  COMA_CHECK_ARG(a && a->x && a->p, "%s: null argument", who);
  const int CV = a->C / 8;
  COMA_CHECK_ARG(a->C % 8 == 0 && CV >= 1 && CV <= 32 && (CV & (CV - 1)) == 0, "%s: C=%d must be 8*2^k <= 256", who, a->C);
  COMA_CHECK_ARG(a->x_cs % 8 == 0 && a->x_co % 8 == 0 && a->out_cs % 8 == 0 && a->out_co % 8 == 0, "%s: unaligned channel view", who);
  return COMA_OK;
}

extern "C" int coma_gate_apply_fwd(const coma_bcast_mul_args* a, coma_stream_t stream) {
  if (int rc = check_bcast(a, "coma_gate_apply_fwd")) return rc;
  COMA_CHECK_ARG(a->out, "coma_gate_apply_fwd: null out");
  const int64_t total = (int64_t)a->B * a->V * (a->C / 8);
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
  if (a->dtype == COMA_BF16) bcast_mul_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(*a);
  else bcast_mul_fwd_kernel<float><<<blocks, 256, 0, stream>>>(*a);
  COMA_CHECK_LAUNCH("gate_apply_fwd");
  return COMA_OK;
}

extern "C" int coma_gate_bwd(const coma_bcast_mul_args* a, coma_stream_t stream) {
  if (int rc = check_bcast(a, "coma_gate_bwd")) return rc;
  COMA_CHECK_ARG(a->dout && a->dx && a->dp, "coma_gate_bwd: null gradient buffer");
  const int64_t total = (int64_t)a->B * a->V * (a->C / 8);
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
  if (a->dtype == COMA_BF16) bcast_mul_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(*a);
  else bcast_mul_bwd_kernel<float><<<blocks, 256, 0, stream>>>(*a);
  COMA_CHECK_LAUNCH("gate_bwd");
  return COMA_OK;
}
