// Weight layout transforms: the framework keeps a convolution weight as [A][B][T] fp32 (T = k^3 taps fastest: Conv3d
// [Cout][Cin][k][k][k], ConvTranspose3d [Cin][Cout][k][k][k]); every conv kernel of this library reads [T][rows][cols] with the
// reduction channel fastest, zero padded to the kernel's channel granularity, in the compute dtype.  One launch per transform
// replaces the permute / zero-fill / slice-assign / cast / flip / transpose chain of ATen kernels (~7 launches per conv and
// training step, ~350 per step of the north-star model).
//
// A block moves a 16 x 16 tile of (a, b) pairs with a group of taps (9 of 27) through shared memory, so that both the parameter
// side (runs of consecutive taps) and the packed side (runs of 16 consecutive channels per tap) are accessed in whole sectors.
#include "common.cuh"

namespace coma {

constexpr int WT = 16;
constexpr int WT_MAX_TAPS = 27;

// TZ taps per block (blockIdx.z selects the group; TZ = 0: all T taps, run-time trip counts).  With a compile-time TZ both loops
// unroll completely, so a block has all its loads in flight at once: these kernels run ~125 times per training step on a few KB
// each and their LATENCY is what the step pays (8 us per launch with the rolled loops, see profiles/r02_ncu_launch_list_train_b4).
template <typename PT, int TZ>
__global__ void __launch_bounds__(256) weight_pack_kernel(const float* __restrict__ param, PT* __restrict__ packed, int A, int B,
                                                          int T, int R_pad, int C_pad, int swap, int flip) {
  __shared__ float tile[WT * WT * (TZ ? TZ : WT_MAX_TAPS)];
  const int tz = TZ ? TZ : T;
  const int r0 = blockIdx.y * WT, c0 = blockIdx.x * WT;
  const int a0 = swap ? c0 : r0, b0 = swap ? r0 : c0;
  const int t0 = blockIdx.z * tz;                                  // packed taps [t0, t0 + tz)
  const int s0 = flip ? T - t0 - tz : t0;                          // the source taps they come from (mirrored: same set, reversed)
  constexpr int ITER = TZ ? TZ : WT_MAX_TAPS;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int i = threadIdx.x + k * 256;
    if (!TZ && i >= WT * WT * tz) break;
    const int ai = i / (WT * tz), rem = i - ai * (WT * tz);
    const int bi = rem / tz, tt = rem - bi * tz;
    const int a = a0 + ai, b = b0 + bi;
    tile[i] = (a < A && b < B) ? __ldg(param + ((size_t)a * B + b) * T + s0 + tt) : 0.f;
  }
  __syncthreads();
  const int rc = threadIdx.x, ri = rc / WT, ci = rc - ri * WT;
  const int r = r0 + ri, c = c0 + ci;
  if (r >= R_pad || c >= C_pad) return;
  const int ai = swap ? ci : ri, bi = swap ? ri : ci;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    if (!TZ && k >= tz) break;
    const int tt = flip ? tz - 1 - k : k;
    Elem<PT>::st(packed + ((size_t)(t0 + k) * R_pad + r) * C_pad + c, tile[(ai * WT + bi) * tz + tt]);
  }
}

// packed fp32 [T][R_pad][C_pad] (a weight gradient as the kernels produce it) -> parameter layout [A][B][T], (r, c) = (a, b)
template <int TZ>
__global__ void __launch_bounds__(256) weight_unpack_kernel(const float* __restrict__ packed, float* __restrict__ param, int A, int B,
                                                            int T, int R_pad, int C_pad) {
  __shared__ float tile[WT * WT * (TZ ? TZ : WT_MAX_TAPS)];
  const int tz = TZ ? TZ : T;
  const int a0 = blockIdx.y * WT, b0 = blockIdx.x * WT, t0 = blockIdx.z * tz;
  constexpr int ITER = TZ ? TZ : WT_MAX_TAPS;
  {
    const int rc = threadIdx.x, ai = rc / WT, bi = rc - ai * WT;
    const int a = a0 + ai, b = b0 + bi;
    const bool ok = a < A && b < B && a < R_pad && b < C_pad;
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
      if (!TZ && k >= tz) break;
      tile[(ai * WT + bi) * tz + k] = ok ? __ldg(packed + ((size_t)(t0 + k) * R_pad + a) * C_pad + b) : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int i = threadIdx.x + k * 256;
    if (!TZ && i >= WT * WT * tz) break;
    const int ai = i / (WT * tz), rem = i - ai * (WT * tz);
    const int bi = rem / tz, tt = rem - bi * tz;
    const int a = a0 + ai, b = b0 + bi;
    if (a < A && b < B) param[((size_t)a * B + b) * T + t0 + tt] = tile[i];
  }
}

template <typename PT>
static void launch_pack(const coma_weight_layout_args* a, cudaStream_t s) {
  const float* src = static_cast<const float*>(a->param);
  PT* dst = static_cast<PT*>(a->packed);
  const unsigned gx = (a->C_pad + WT - 1) / WT, gy = (a->R_pad + WT - 1) / WT;
  if (a->T == 27)
    weight_pack_kernel<PT, 9><<<dim3(gx, gy, 3), 256, 0, s>>>(src, dst, a->A, a->B, a->T, a->R_pad, a->C_pad, a->swap, a->flip);
  else if (a->T == 1)
    weight_pack_kernel<PT, 1><<<dim3(gx, gy, 1), 256, 0, s>>>(src, dst, a->A, a->B, a->T, a->R_pad, a->C_pad, a->swap, a->flip);
  else
    weight_pack_kernel<PT, 0><<<dim3(gx, gy, 1), 256, 0, s>>>(src, dst, a->A, a->B, a->T, a->R_pad, a->C_pad, a->swap, a->flip);
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
    const float* src = static_cast<const float*>(a->packed);
    float* dst = static_cast<float*>(a->param);
    const unsigned gx = (a->B + WT - 1) / WT, gy = (a->A + WT - 1) / WT;
    if (a->T == 27)
      weight_unpack_kernel<9><<<dim3(gx, gy, 3), 256, 0, s>>>(src, dst, a->A, a->B, a->T, a->R_pad, a->C_pad);
    else if (a->T == 1)
      weight_unpack_kernel<1><<<dim3(gx, gy, 1), 256, 0, s>>>(src, dst, a->A, a->B, a->T, a->R_pad, a->C_pad);
    else
      weight_unpack_kernel<0><<<dim3(gx, gy, 1), 256, 0, s>>>(src, dst, a->A, a->B, a->T, a->R_pad, a->C_pad);
  } else if (a->dtype == COMA_BF16) {
    launch_pack<__nv_bfloat16>(a, s);
  } else if (a->dtype == COMA_F32) {
    launch_pack<float>(a, s);
  } else {
    COMA_CHECK_ARG(false, "coma_weight_layout: dtype must be COMA_BF16 or COMA_F32");
  }
  COMA_CHECK_LAUNCH("weight_layout");
  return COMA_OK;
}
