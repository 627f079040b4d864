// Weight layout transforms: the framework keeps a convolution weight as [A][B][T] fp32 (T = k^3 taps fastest: Conv3d
// [Cout][Cin][k][k][k], ConvTranspose3d [Cin][Cout][k][k][k]); every conv kernel of this library reads [T][rows][cols] with the
// reduction channel fastest, zero padded to the kernel's channel granularity, in the compute dtype.  One launch per transform
// replaces the permute / zero-fill / slice-assign / cast / flip / transpose chain of ATen kernels (~7 launches per conv and
// training step, ~350 per step of the north-star model).
//
// A block moves a 16 x 16 tile of (a, b) pairs with all their taps through shared memory, so that both the parameter side
// (runs of 16 * T consecutive floats) and the packed side (runs of 16 consecutive channels per tap) are accessed in full sectors.
#include "common.cuh"

namespace coma {

constexpr int WT = 16;
constexpr int WT_MAX_TAPS = 27;

template <typename PT>
__global__ void __launch_bounds__(256) weight_pack_kernel(const float* __restrict__ param, PT* __restrict__ packed, int A, int B,
                                                          int T, int R_pad, int C_pad, int swap, int flip) {
  __shared__ float tile[WT * WT * WT_MAX_TAPS];
  const int r0 = blockIdx.y * WT, c0 = blockIdx.x * WT;
  const int a0 = swap ? c0 : r0, b0 = swap ? r0 : c0;
  const int n = WT * WT * T;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int ai = i / (WT * T), rem = i - ai * (WT * T);
    const int bi = rem / T, t = rem - bi * T;
    const int a = a0 + ai, b = b0 + bi;
    tile[i] = (a < A && b < B) ? __ldg(param + ((size_t)a * B + b) * T + t) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * WT * WT; i += 256) {
    const int t = i / (WT * WT), rc = i - t * (WT * WT);
    const int ri = rc / WT, ci = rc - ri * WT;
    const int r = r0 + ri, c = c0 + ci;
    if (r >= R_pad || c >= C_pad) continue;
    const int ai = swap ? ci : ri, bi = swap ? ri : ci;
    const int ts = flip ? T - 1 - t : t;
    Elem<PT>::st(packed + ((size_t)t * R_pad + r) * C_pad + c, tile[(ai * WT + bi) * T + ts]);
  }
}

// packed fp32 [T][R_pad][C_pad] (a weight gradient as the kernels produce it) -> parameter layout [A][B][T], (r, c) = (a, b)
__global__ void __launch_bounds__(256) weight_unpack_kernel(const float* __restrict__ packed, float* __restrict__ param, int A, int B,
                                                            int T, int R_pad, int C_pad) {
  __shared__ float tile[WT * WT * WT_MAX_TAPS];
  const int a0 = blockIdx.y * WT, b0 = blockIdx.x * WT;
  for (int i = threadIdx.x; i < T * WT * WT; i += 256) {
    const int t = i / (WT * WT), rc = i - t * (WT * WT);
    const int ai = rc / WT, bi = rc - ai * WT;
    const int a = a0 + ai, b = b0 + bi;
    tile[(ai * WT + bi) * T + t] = (a < A && b < B && a < R_pad && b < C_pad) ? __ldg(packed + ((size_t)t * R_pad + a) * C_pad + b) : 0.f;
  }
  __syncthreads();
  const int n = WT * WT * T;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int ai = i / (WT * T), rem = i - ai * (WT * T);
    const int bi = rem / T, t = rem - bi * T;
    const int a = a0 + ai, b = b0 + bi;
    if (a < A && b < B) param[((size_t)a * B + b) * T + t] = tile[i];
  }
}

}  // namespace coma

extern "C" int coma_weight_layout(const coma_weight_layout_args* a, coma_stream_t stream) {
  using namespace coma;
  COMA_CHECK_ARG(a && a->param && a->packed, "coma_weight_layout: null pointer");
  COMA_CHECK_ARG(a->A > 0 && a->B > 0 && a->T > 0 && a->T <= WT_MAX_TAPS, "coma_weight_layout: bad extents (taps <= 27)");
  COMA_CHECK_ARG(a->R_pad > 0 && a->C_pad > 0, "coma_weight_layout: bad packed extents");
  cudaStream_t s = (cudaStream_t)stream;
  if (a->unpack) {
    COMA_CHECK_ARG(a->dtype == COMA_F32 && !a->swap && !a->flip, "coma_weight_layout: unpack takes an fp32 packed tensor, no swap / flip");
    dim3 grid((a->B + WT - 1) / WT, (a->A + WT - 1) / WT);
    weight_unpack_kernel<<<grid, 256, 0, s>>>((const float*)a->packed, (float*)a->param, a->A, a->B, a->T, a->R_pad, a->C_pad);
  } else {
    dim3 grid((a->C_pad + WT - 1) / WT, (a->R_pad + WT - 1) / WT);
    if (a->dtype == COMA_BF16)
      weight_pack_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const float*)a->param, (__nv_bfloat16*)a->packed, a->A, a->B, a->T,
                                                              a->R_pad, a->C_pad, a->swap, a->flip);
    else if (a->dtype == COMA_F32)
      weight_pack_kernel<float><<<grid, 256, 0, s>>>((const float*)a->param, (float*)a->packed, a->A, a->B, a->T, a->R_pad,
                                                     a->C_pad, a->swap, a->flip);
    else
      COMA_CHECK_ARG(false, "coma_weight_layout: dtype must be COMA_BF16 or COMA_F32");
  }
  COMA_CHECK_LAUNCH("weight_layout");
  return COMA_OK;
}
