// All FiLM MLPs of a model (Linear(n, 64) -> ReLU -> Linear(64, 2C) per conditioned layer) in one launch each way: one block per
// layer, everything in fp32.  The work is a few hundred KFLOP per layer; what it replaces is ~12 framework launches per layer and
// training step (cuBLAS gemv / sgemm on [B x 6] .. [B x 512] matrices, ReLU, bias reductions, chunk / cat), ~185 launches per step
// of the north-star model.
#include "common.cuh"

namespace coma {
namespace {
constexpr int FH = COMA_FILM_HIDDEN;
constexpr int kFilmThreads = 256;

__global__ void __launch_bounds__(kFilmThreads) film_fwd_kernel(const __grid_constant__ coma_film_args a) {
  extern __shared__ float fsm[];
  const int l = blockIdx.x, B = a.B, n = a.n_cov[l], C = a.C[l];
  float* hid = fsm;                                   // [B][64]
  const float* W1 = a.W1[l];
  const float* b1 = a.b1[l];
  for (int i = threadIdx.x; i < B * FH; i += kFilmThreads) {
    const int b = i / FH, j = i - b * FH;
    float s = __ldg(b1 + j);
    for (int k = 0; k < n; ++k) s = fmaf(__ldg(a.cov + (size_t)b * a.cov_stride + k), __ldg(W1 + j * n + k), s);
    s = fmaxf(s, 0.f);
    hid[i] = s;
    a.hid[((size_t)l * B + b) * FH + j] = s;
  }
  __syncthreads();
  const float* W2 = a.W2[l];
  const float* b2 = a.b2[l];
  float* out = a.out[l];
  for (int o = threadIdx.x; o < 2 * C; o += kFilmThreads) {
    float w[FH];
#pragma unroll
    for (int j = 0; j < FH; j += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(W2 + (size_t)o * FH + j));
      w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
    }
    const float bias = __ldg(b2 + o);
    const int half = o >= C, c = o - half * C;
    for (int b = 0; b < B; ++b) {
      float s = bias;
#pragma unroll
      for (int j = 0; j < FH; ++j) s = fmaf(hid[b * FH + j], w[j], s);
      out[((size_t)half * B + b) * C + c] = s;
    }
  }
}

__global__ void __launch_bounds__(kFilmThreads) film_bwd_kernel(const __grid_constant__ coma_film_args a) {
  extern __shared__ float fsm[];
  const int l = blockIdx.x, B = a.B, n = a.n_cov[l], C = a.C[l];
  float* hid = fsm;                                   // [B][64]
  float* dhid = hid + B * FH;                         // [B][64]
  float* dout = dhid + B * FH;                        // [B][2C]: d_dgamma | d_beta
  for (int i = threadIdx.x; i < B * FH; i += kFilmThreads) hid[i] = a.hid[(size_t)l * B * FH + i];
  const float* dg = a.d_dgamma[l];
  const float* db = a.d_beta[l];
  for (int i = threadIdx.x; i < B * 2 * C; i += kFilmThreads) {
    const int b = i / (2 * C), o = i - b * 2 * C;
    const float* src = o < C ? dg : db;
    dout[i] = src ? __ldg(src + (size_t)b * C + (o < C ? o : o - C)) : 0.f;
  }
  __syncthreads();
  // second Linear: db2[o] = sum_b dout[b][o], dW2[o][j] = sum_b dout[b][o] hid[b][j]
  for (int o = threadIdx.x; o < 2 * C; o += kFilmThreads) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dout[b * 2 * C + o];
    a.db2[l][o] = s;
  }
  for (int i = threadIdx.x; i < 2 * C * FH; i += kFilmThreads) {
    const int o = i / FH, j = i - o * FH;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dout[b * 2 * C + o], hid[b * FH + j], s);
    a.dW2[l][i] = s;
  }
  // through the ReLU: dhid[b][j] = (hid > 0) * sum_o dout[b][o] W2[o][j].  W2 (up to 128 KB) was last touched in the forward pass,
  // a whole step ago: the loads come from HBM, so they are spread over the warps (warp w owns the rows o = w, w + 8, ..., a lane
  // the columns lane and lane + 32) and unrolled 8 deep -- 16 loads in flight per thread instead of a 512-long chain per thread
  // (157 us for the 256-channel layer before).
  const float* W2 = a.W2[l];
  float* part = dout + B * 2 * C;                     // [8 warps][4][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b0 = 0; b0 < B; b0 += 4) {
    float acc[4][2];
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) acc[bb][0] = acc[bb][1] = 0.f;
#pragma unroll 8
    for (int o = warp; o < 2 * C; o += 8) {
      const float wa = __ldg(W2 + (size_t)o * FH + lane), wb = __ldg(W2 + (size_t)o * FH + lane + 32);
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const float d = b0 + bb < B ? dout[(b0 + bb) * 2 * C + o] : 0.f;
        acc[bb][0] = fmaf(d, wa, acc[bb][0]);
        acc[bb][1] = fmaf(d, wb, acc[bb][1]);
      }
    }
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      part[(warp * 4 + bb) * FH + lane] = acc[bb][0];
      part[(warp * 4 + bb) * FH + lane + 32] = acc[bb][1];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * FH; i += kFilmThreads) {
      const int bb = i / FH, j = i - bb * FH;
      if (b0 + bb < B) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[(w * 4 + bb) * FH + j];
        dhid[(b0 + bb) * FH + j] = hid[(b0 + bb) * FH + j] > 0.f ? t : 0.f;
      }
    }
    __syncthreads();
  }
  // first Linear: db1[j] = sum_b dhid[b][j], dW1[j][k] = sum_b dhid[b][j] cov[b][k]
  for (int j = threadIdx.x; j < FH; j += kFilmThreads) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dhid[b * FH + j];
    a.db1[l][j] = s;
  }
  for (int i = threadIdx.x; i < FH * n; i += kFilmThreads) {
    const int j = i / n, k = i - j * n;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dhid[b * FH + j], __ldg(a.cov + (size_t)b * a.cov_stride + k), s);
    a.dW1[l][i] = s;
  }
}

int check_film(const coma_film_args* a, bool bwd, size_t* smem) {
  COMA_CHECK_ARG(a && a->cov && a->hid && a->n_layers > 0 && a->n_layers <= COMA_FILM_MAX_LAYERS && a->B > 0 && a->B <= 64,
                 "coma_film_mlp: bad arguments (1..32 layers, batch <= 64)");
  int cmax = 0;
  for (int l = 0; l < a->n_layers; ++l) {
    COMA_CHECK_ARG(a->W1[l] && a->b1[l] && a->W2[l] && a->b2[l] && a->C[l] > 0 && a->n_cov[l] > 0 && a->n_cov[l] <= a->cov_stride,
                   "coma_film_mlp: layer %d: null parameter or bad extents", l);
    COMA_CHECK_ARG(reinterpret_cast<uintptr_t>(a->W2[l]) % 16 == 0, "coma_film_mlp: layer %d: W2 must be 16-byte aligned", l);
    if (bwd) COMA_CHECK_ARG(a->dW1[l] && a->db1[l] && a->dW2[l] && a->db2[l], "coma_film_mlp_bwd: layer %d: null gradient output", l);
    else COMA_CHECK_ARG(a->out[l] != nullptr, "coma_film_mlp_fwd: layer %d: null output", l);
    cmax = a->C[l] > cmax ? a->C[l] : cmax;
  }
  *smem = (size_t)a->B * FH * sizeof(float) * (bwd ? 2 : 1) + (bwd ? ((size_t)a->B * 2 * cmax + 8 * 4 * FH) * sizeof(float) : 0);
  COMA_CHECK_ARG(*smem <= 200 * 1024, "coma_film_mlp: batch x channels too large for one block's shared memory");
  return COMA_OK;
}
}  // namespace
}  // namespace coma

extern "C" int coma_film_mlp_fwd(const coma_film_args* a, coma_stream_t stream) {
  using namespace coma;
  size_t smem = 0;
  if (int rc = check_film(a, false, &smem)) return rc;
  film_fwd_kernel<<<(unsigned)a->n_layers, kFilmThreads, smem, (cudaStream_t)stream>>>(*a);
  COMA_CHECK_LAUNCH("film_mlp_fwd");
  return COMA_OK;
}

extern "C" int coma_film_mlp_bwd(const coma_film_args* a, coma_stream_t stream) {
  using namespace coma;
  size_t smem = 0;
  if (int rc = check_film(a, true, &smem)) return rc;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(film_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("coma_film_mlp_bwd: cannot get %zu bytes of shared memory", smem); return COMA_ERR_CUDA; }
  }
  film_bwd_kernel<<<(unsigned)a->n_layers, kFilmThreads, smem, (cudaStream_t)stream>>>(*a);
  COMA_CHECK_LAUNCH("film_mlp_bwd");
  return COMA_OK;
}
