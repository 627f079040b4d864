// C-ABI entry points that dispatch between kernel implementations (include/coma_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace coma {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int conv_simt_stat_chunks(const coma_conv_args& a);
int conv_simt_launch(const coma_conv_args& a, cudaStream_t stream);
int wgrad_simt_launch(const coma_wgrad_args& a, cudaStream_t stream);
bool wgrad_mma_supported(const coma_wgrad_args& a);
int wgrad_mma_launch(const coma_wgrad_args& a, cudaStream_t stream);
bool wgrad_tc_supported(const coma_wgrad_args& a);
int64_t wgrad_tc_workspace(const coma_wgrad_args& a);
bool wgrad_c1k3_supported(const coma_wgrad_args& a);
int64_t wgrad_c1k3_workspace(const coma_wgrad_args& a);
int wgrad_c1k3_launch(const coma_wgrad_args& a, cudaStream_t stream);
int wgrad_tc_launch(const coma_wgrad_args& a, cudaStream_t stream);
bool conv_tc_supported(const coma_conv_args& a);
int conv_tc_stat_chunks(const coma_conv_args& a);
int conv_tc_launch(const coma_conv_args& a, cudaStream_t stream);
bool conv_tc_prologue_supported(const coma_conv_args& a);
bool conv_simt_prologue_fused(const coma_conv_args& a);
bool conv_simt_preferred(const coma_conv_args& a);

static int check_conv(const coma_conv_args* a, const char* who) {
  COMA_CHECK_ARG(a && a->x && a->w && a->y, "%s: null tensor", who);
  COMA_CHECK_ARG(a->B > 0 && a->Cin > 0 && a->Cout > 0 && a->Di > 0 && a->Hi > 0 && a->Wi > 0, "%s: bad extents", who);
  COMA_CHECK_ARG(a->ksize == 1 || a->ksize == 3, "%s: kernel size %d not on the hot path (1 or 3)", who, a->ksize);
  COMA_CHECK_ARG(a->stride == 1 || a->stride == 2, "%s: stride %d not on the hot path (1 or 2)", who, a->stride);
  COMA_CHECK_ARG(a->pad == (a->ksize - 1) / 2, "%s: padding must be (k-1)/2", who);
  COMA_CHECK_ARG(a->dtype == COMA_F32 || a->dtype == COMA_BF16 || a->dtype == COMA_BF16_F32OUT, "%s: bad dtype", who);
  COMA_CHECK_ARG(a->dtype != COMA_BF16_F32OUT || (!a->in_scale && a->w_bstride == 0), "%s: the bf16-in / fp32-out mode takes no prologue and no per-sample weights", who);
  COMA_CHECK_ARG((a->scale == nullptr) == (a->shift == nullptr), "%s: scale and shift go together", who);
  COMA_CHECK_ARG(a->y_cn > 0 && a->y_cn <= a->Cout && a->y_co + a->y_cn <= a->y_cs, "%s: bad output channel view", who);
  COMA_CHECK_ARG(a->x_co + a->Cin <= a->x_cs, "%s: bad input channel view", who);
  COMA_CHECK_ARG((a->in_scale == nullptr) == (a->in_shift == nullptr), "%s: in_scale and in_shift go together", who);
  COMA_CHECK_ARG(!a->in_scale || a->in_act == COMA_ACT_NONE || a->in_act == COMA_ACT_RELU || (a->in_act == COMA_ACT_LEAKY && a->in_slope),
                 "%s: the input prologue takes act none / relu / leaky(+slope)", who);
  COMA_CHECK_ARG(!a->in_scale || !a->transposed, "%s: no input prologue on transposed convolutions", who);
  if (!a->transposed) {
    COMA_CHECK_ARG(a->Do == (a->Di + 2 * a->pad - a->ksize) / a->stride + 1 && a->Ho == (a->Hi + 2 * a->pad - a->ksize) / a->stride + 1 &&
                       a->Wo == (a->Wi + 2 * a->pad - a->ksize) / a->stride + 1, "%s: output extents do not match the conv geometry", who);
  } else {
    COMA_CHECK_ARG(a->Do == a->Di * a->stride && a->Ho == a->Hi * a->stride && a->Wo == a->Wi * a->stride,
                   "%s: transposed conv output must be stride x input (output_padding = stride-1)", who);
  }
  return COMA_OK;
}

// COMA_DISABLE_TCGEN05=1 routes IMPL_AUTO to the CUDA-core kernels (A/B debugging of the tensor-core path)
static bool tc_enabled() {
  static const bool on = [] { const char* e = getenv("COMA_DISABLE_TCGEN05"); return !(e && e[0] == '1'); }();
  return on;
}
static int pick_impl(const coma_conv_args& a) {
  if (a.dtype == COMA_BF16_F32OUT) return COMA_IMPL_TCGEN05;   // only the per-tap tcgen05 kernel stores fp32 (run_conv rejects what it cannot take)
  if (a.impl != COMA_IMPL_AUTO) return a.impl;
  if (conv_simt_preferred(a)) return COMA_IMPL_SIMT;        // few-channel pointwise: HBM streaming kernel
  return (tc_enabled() && conv_tc_supported(a)) ? COMA_IMPL_TCGEN05 : COMA_IMPL_SIMT;
}

static int run_conv(const coma_conv_args* a, cudaStream_t stream, const char* who) {
  if (int rc = check_conv(a, who)) return rc;
  const int impl = pick_impl(*a);
  if (impl == COMA_IMPL_TCGEN05) {
    if (!conv_tc_supported(*a)) {
      set_error("%s: problem not supported by the tcgen05 path", who);
      return COMA_ERR_UNSUPPORTED;
    }
    return conv_tc_launch(*a, stream);
  }
  return conv_simt_launch(*a, stream);
}

}  // namespace coma

using namespace coma;

extern "C" int coma_version(void) { return 100; }
extern "C" const char* coma_last_error(void) { return g_error; }

extern "C" int coma_conv3d_stat_chunks(const coma_conv_args* a) {
  if (!a) return 0;
  const int impl = pick_impl(*a);
  return impl == COMA_IMPL_TCGEN05 ? conv_tc_stat_chunks(*a) : conv_simt_stat_chunks(*a);
}
extern "C" int coma_conv3d_tcgen05_supported(const coma_conv_args* a) { return a && conv_tc_supported(*a) ? 1 : 0; }
extern "C" int coma_conv3d_impl(const coma_conv_args* a) { return a ? pick_impl(*a) : COMA_IMPL_SIMT; }
extern "C" int coma_conv3d_wgrad_tcgen05_supported(const coma_wgrad_args* a) {
  return a && a->impl != COMA_IMPL_SIMT && wgrad_tc_supported(*a) ? 1 : 0;
}
extern "C" int64_t coma_conv3d_wgrad_workspace_size(const coma_wgrad_args* a) {
  if (!a || a->impl == COMA_IMPL_SIMT) return 0;
  if (wgrad_tc_supported(*a)) return wgrad_tc_workspace(*a);
  if (wgrad_c1k3_supported(*a)) return wgrad_c1k3_workspace(*a);
  return 0;
}
extern "C" int coma_conv3d_prologue_supported(const coma_conv_args* a) {
  if (!a || a->transposed) return 0;
  const int impl = pick_impl(*a);
  return (impl == COMA_IMPL_TCGEN05 ? conv_tc_prologue_supported(*a) : conv_simt_prologue_fused(*a)) ? 1 : 0;
}

extern "C" int coma_conv3d_fprop(const coma_conv_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && !a->transposed, "coma_conv3d_fprop: use coma_convT3d_fprop for transposed convolutions");
  return run_conv(a, stream, "coma_conv3d_fprop");
}
extern "C" int coma_convT3d_fprop(const coma_conv_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && a->transposed, "coma_convT3d_fprop: transposed flag not set");
  return run_conv(a, stream, "coma_convT3d_fprop");
}
// dgrad(conv, stride 1) = conv with flipped kernel; dgrad(conv, stride s>1) = transposed conv
extern "C" int coma_conv3d_dgrad(const coma_conv_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && (a->transposed || a->stride == 1), "coma_conv3d_dgrad: the adjoint of a strided conv is a transposed conv");
  return run_conv(a, stream, "coma_conv3d_dgrad");
}
// dgrad(convT, stride s) = conv with stride s
extern "C" int coma_convT3d_dgrad(const coma_conv_args* a, coma_stream_t stream) {
  COMA_CHECK_ARG(a && !a->transposed, "coma_convT3d_dgrad: the adjoint of a transposed conv is a strided conv");
  return run_conv(a, stream, "coma_convT3d_dgrad");
}

static int check_wgrad(const coma_wgrad_args* a, const char* who) {
  COMA_CHECK_ARG(a && a->g && a->x && a->dw, "%s: null tensor", who);
  COMA_CHECK_ARG(a->ksize == 1 || a->ksize == 3, "%s: kernel size must be 1 or 3", who);
  COMA_CHECK_ARG(a->pad == (a->ksize - 1) / 2 && (a->stride == 1 || a->stride == 2), "%s: bad geometry", who);
  COMA_CHECK_ARG(a->Dg == (a->Dx + 2 * a->pad - a->ksize) / a->stride + 1 && a->Hg == (a->Hx + 2 * a->pad - a->ksize) / a->stride + 1 &&
                     a->Wg == (a->Wx + 2 * a->pad - a->ksize) / a->stride + 1, "%s: extents do not match the conv geometry", who);
  return COMA_OK;
}
static int run_wgrad(const coma_wgrad_args* a, cudaStream_t stream) {
  if (a->impl != COMA_IMPL_SIMT && wgrad_tc_supported(*a)) return wgrad_tc_launch(*a, stream);      // tcgen05 (k3, stride 1 / 2)
  if (a->impl != COMA_IMPL_SIMT && wgrad_c1k3_supported(*a)) return wgrad_c1k3_launch(*a, stream);  // one-channel gradient, k3 s1
  if (a->impl != COMA_IMPL_SIMT && wgrad_mma_supported(*a)) return wgrad_mma_launch(*a, stream);
  return wgrad_simt_launch(*a, stream);
}
extern "C" int coma_conv3d_wgrad(const coma_wgrad_args* a, coma_stream_t stream) {
  if (int rc = check_wgrad(a, "coma_conv3d_wgrad")) return rc;
  return run_wgrad(a, stream);
}
extern "C" int coma_convT3d_wgrad(const coma_wgrad_args* a, coma_stream_t stream) {
  if (int rc = check_wgrad(a, "coma_convT3d_wgrad")) return rc;
  return run_wgrad(a, stream);
}
