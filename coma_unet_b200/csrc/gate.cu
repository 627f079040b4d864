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


// ---- bf16 tensor-core variant (mma.sync m16n8k16): the two 1x1x1 convolutions of the gate are a [32 voxels x 2C] x [2C x C/2]
// GEMM per warp iteration, so the kernel becomes HBM-bound (3*C*2 bytes per voxel) instead of FMA/LDS-bound. ----
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// three blocks (24 warps) per SM: each warp's load -> smem -> mma -> sigmoid -> store chain is latency-bound, HBM wants them all
template <int C>
__global__ void __launch_bounds__(256, 3) gate_mma_kernel(coma_gate_args a) {
  constexpr int F = C / 2, LD = C + 8, NTILES = F / 8, CV = C / 8;
  extern __shared__ __align__(16) uint8_t gsm[];
  __nv_bfloat16* swg = reinterpret_cast<__nv_bfloat16*>(gsm);            // [F][LD]
  __nv_bfloat16* swx = swg + F * LD;                                      // [F][LD]
  __nv_bfloat16* tiles = swx + F * LD;                                    // [8 warps][2][32][LD]
  float* sb = reinterpret_cast<float*>(tiles + 8 * 2 * 32 * LD);          // [F]
  float* sp = sb + F;                                                     // [F]
  float* satt = sp + F;                                                   // [8][32]
  for (int i = threadIdx.x; i < F * C; i += 256) {
    swg[(i / C) * LD + (i % C)] = __float2bfloat16_rn(a.wg[i]);
    swx[(i / C) * LD + (i % C)] = __float2bfloat16_rn(a.wx[i]);
  }
  for (int i = threadIdx.x; i < F; i += 256) { sb[i] = a.bsum[i]; sp[i] = a.wpsi[i]; }
  __syncthreads();
  const float bpsi = a.bpsi_ptr ? __ldg(a.bpsi_ptr) : a.bpsi;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* tg = tiles + (size_t)warp * 2 * 32 * LD;
  __nv_bfloat16* tx = tg + 32 * LD;
  float* att = satt + warp * 32;
  const int64_t total = (int64_t)a.B * a.V, ntiles = (total + 31) / 32;
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(a.g) + a.g_co;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(a.x) + a.x_co;
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(a.out) + a.out_co;
  for (int64_t tile = (int64_t)blockIdx.x * 8 + warp; tile < ntiles; tile += (int64_t)gridDim.x * 8) {
    const int64_t v0 = tile * 32;
#pragma unroll
    for (int i = 0; i < CV; ++i) {
      const int idx = lane + 32 * i, row = idx / CV, vec = (idx % CV) * 8;
      const int64_t v = v0 + row;
      uint4 gv = make_uint4(0, 0, 0, 0), xv = make_uint4(0, 0, 0, 0);
      if (v < total) {
        gv = __ldg(reinterpret_cast<const uint4*>(gp + v * a.g_cs + vec));
        xv = __ldg(reinterpret_cast<const uint4*>(xp + v * a.x_cs + vec));
      }
      *reinterpret_cast<uint4*>(tg + row * LD + vec) = gv;
      *reinterpret_cast<uint4*>(tx + row * LD + vec) = xv;
    }
    __syncwarp();
    float acc[2][NTILES][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < NTILES; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[m][n][e] = 0.f;
    const int mat = lane >> 3, r = lane & 7;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const __nv_bfloat16* ta = half ? tx : tg;
      const __nv_bfloat16* tw = half ? swx : swg;
#pragma unroll
      for (int k0 = 0; k0 < C; k0 += 16) {
        uint32_t af[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) ldsm_x4(af[m], ta + (m * 16 + (mat & 1) * 8 + r) * LD + k0 + (mat >> 1) * 8);
        if constexpr (NTILES >= 2) {
#pragma unroll
          for (int n = 0; n < NTILES; n += 2) {
            uint32_t bf[4];
            ldsm_x4(bf, tw + (n * 8 + (mat >> 1) * 8 + r) * LD + k0 + (mat & 1) * 8);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              mma16816(acc[m][n], af[m], bf[0], bf[1]);
              mma16816(acc[m][n + 1], af[m], bf[2], bf[3]);
            }
          }
        } else {
          uint32_t bf[2];
          ldsm_x2(bf, tw + ((lane & 7)) * LD + k0 + ((lane >> 3) & 1) * 8);
#pragma unroll
          for (int m = 0; m < 2; ++m) mma16816(acc[m][0], af[m], bf[0], bf[1]);
        }
      }
    }
    // q = sum_f wpsi[f] * relu(t[f] + b[f]);  accumulator: (e&1) -> column 2*(lane%4)+(e&1), (e>>1) -> row lane/4 + 8*(e>>1)
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int n = 0; n < NTILES; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = n * 8 + (lane & 3) * 2 + (e & 1);
          const float t = fmaxf(acc[m][n][e] + sb[col], 0.f) * sp[col];
          if (e < 2) q0 += t; else q1 += t;
        }
      q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
      if ((lane & 3) == 0) {
        att[m * 16 + (lane >> 2)] = 1.f / (1.f + __expf(-(q0 + bpsi)));
        att[m * 16 + (lane >> 2) + 8] = 1.f / (1.f + __expf(-(q1 + bpsi)));
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < CV; ++i) {
      const int idx = lane + 32 * i, row = idx / CV, vec = (idx % CV) * 8;
      const int64_t v = v0 + row;
      if (v < total) {
        float xv[8];
        const uint4 raw = *reinterpret_cast<const uint4*>(tx + row * LD + vec);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        const float s = att[row];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xv[2 * j] = __uint_as_float(w[j] << 16) * s;
          xv[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u) * s;
        }
        store8(op + v * a.out_cs + vec, xv);
      }
    }
    if (a.psi_out && lane < 32 && v0 + lane < total)
      static_cast<__nv_bfloat16*>(a.psi_out)[v0 + lane] = __float2bfloat16_rn(att[lane]);
    __syncwarp();
  }
}

template <int C>
static int launch_gate_mma(const coma_gate_args& a, cudaStream_t stream) {
  constexpr int F = C / 2, LD = C + 8;
  const size_t smem = (size_t)(2 * F * LD + 8 * 2 * 32 * LD) * sizeof(__nv_bfloat16) + (size_t)(2 * F + 8 * 32) * sizeof(float);
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(gate_mma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = true; }
  const int64_t ntiles = ((int64_t)a.B * a.V + 31) / 32;
  const unsigned blocks = (unsigned)std::min<int64_t>((ntiles + 7) / 8, (int64_t)num_sms() * 2);
  gate_mma_kernel<C><<<blocks, 256, smem, stream>>>(a);
  COMA_CHECK_LAUNCH("gate_mma");
  return COMA_OK;
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
  if (a->dtype == COMA_BF16) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(a->g) | reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->out)) & 15) == 0;
    if (aligned && a->C == 16) return launch_gate_mma<16>(*a, stream);
    if (aligned && a->C == 32) return launch_gate_mma<32>(*a, stream);
    if (aligned && a->C == 64) return launch_gate_mma<64>(*a, stream);
    return launch_gate<__nv_bfloat16>(*a, stream);
  }
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
