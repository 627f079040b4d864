// 3-D convolution as an implicit GEMM on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, tiled 5-D boxes with hardware zero-fill for the halo)  ->  128B/64B/32B-swizzled
//   shared memory  ->  tcgen05.mma (cta_group::1, kind::f16, bf16 x bf16 -> fp32 in TMEM)  ->  tcgen05.ld epilogue.
//
// Replaces cuDNN fprop/dgrad behind torch.nn.Conv3d / ConvTranspose3d on the reference's hot path
// (attn_unet_data_parallel.py:126,285-306,495-497; MONAI ConvBlock / UpConv / AttentionLayer.merge).
//
// GEMM view: M = 128 output voxels (an 8 x 4 x 4 box in W,H,D), N = Cout tile (<= 256), K = taps x Cin.
// One k-step = one filter tap x one channel chunk KC (16/32/64): the A operand is the activation box shifted
// by the tap (a fresh TMA load; out-of-bounds voxels arrive as zeros = the conv padding), the B operand is the
// [Cout x KC] slice of the packed weights w[tap][Cout][Cin].  Stride-2 convolutions use the TMA element stride.
// The transposed convolution (k3, s2, p1, op1) runs as 8 output-parity classes, each an ordinary stride-1
// gather over the INPUT grid with 1..8 taps, scattered to the strided output positions by the epilogue.
//
// Warp roles (192 threads, 1 CTA / SM, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer +
// TMEM allocator, warps 2..5 = epilogue (bias, per-channel statistics partials, per-(sample,channel) affine,
// activation, bf16 store).  Two TMEM accumulators let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cstdio>
#include <type_traits>
#include <cuda.h>

#include <mutex>
#include <unordered_map>
#include <string>
#include <cstring>
#include <cstdlib>

#include "common.cuh"

namespace coma {

namespace {

constexpr int TW = 8, TH = 4, TD = 4;      // M tile = 128 voxels
constexpr int kTcThreads = 192;
constexpr int kMaxStages = 8;

struct TcParams {
  int B, Di, Hi, Wi, Do, Ho, Wo, Cin, Cout;
  int ksize, stride, pad, transposed;
  int KC, kchunks, NT, n_tiles;
  int tiles_w, tiles_h, tiles_d, classes, total_tiles;
  int stages;
  uint32_t a_bytes, b_bytes, stage_bytes;
  int swz;                  // swizzle span in bytes: 32 / 64 / 128
  __nv_bfloat16* y;
  int y_cs, y_cn;
  const float* bias; const float* scale; const float* shift; const float* slope;
  float* stats;
  int act, stat_chunks;
  uint32_t tmem_cols;
  int out_f32;              // COMA_BF16_F32OUT: y is a float tensor (y, y_cs, y_cn in float elements)
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();   // watchdog: a lost TMA / MMA completion must not hang the GPU
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// one lane of a converged warp (all 32 lanes must call it)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------- epilogue helpers
// y = act(cA[c] * acc + cS[c]) with cA = scale (or 1) and cS = scale*bias + shift staged in shared memory per (sample,
// N tile); activation as hi + neg*lo (NONE: neg=1, RELU: neg=0, LEAKY/PReLU: neg=slope; clamp0 = ReLU(PReLU(.))).
// The epilogue warps are the critical path once the MMA side is lean, so nothing but 1 FMA + 3 ALU ops per element remains.
__device__ __forceinline__ void epi_stage_coef(const float* bias, const float* scale, const float* shift, int64_t bc_off, int n0,
                                               int nt, float* cA, float* cS, int tid128, uint32_t bar_id = 1u) {
  float* cB = cS + nt;    // plain bias: the statistics are those of conv + bias, before any scale / shift
  for (int j = tid128; j < nt; j += 128) {
    const float a = scale ? __ldg(scale + bc_off + n0 + j) : 1.f;
    const float bb = bias ? __ldg(bias + n0 + j) : 0.f;
    cA[j] = a;
    cS[j] = fmaf(a, bb, scale ? __ldg(shift + bc_off + n0 + j) : 0.f);
    cB[j] = bb;
  }
  asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
}

// explicit shared-space 128-bit load (the coefficient pointers reach the epilogue as generic pointers, which compiled to LD.E)
__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

// One 16-channel chunk of an accumulator row: v = cA * acc + cS (cS already holds cA * bias + shift), statistics of acc + bias,
// activation, bf16 pack, one 256-bit store.  The epilogue warps are instruction-issue-bound on the small-channel and transposed
// kernels, so the common cases skip work per element: `unit` (no scale: cA == 1 and cS == bias, so v doubles as the statistics
// operand), identity activation (InstanceNorm layers: raw output + statistics), and ReLU as one packed max on the bf16 pairs
// (relu(round(u)) == round(relu(u))); only PReLU / LeakyReLU take the fp32 min/max/fma path.
__device__ __forceinline__ void epi_chunk16(const uint32_t (&raw)[16], const float* cA, const float* cS, const float* cB, bool valid,
                                            float neg, bool clamp0, float slope, bool stats, float* s1, float* s2,
                                            __nv_bfloat16* yrow, int c_abs, int y_cn, int y_cs, bool unit) {
  float v[16];
  if (unit) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 sh = lds128(cS + 4 * q4);
      v[4 * q4 + 0] = __uint_as_float(raw[4 * q4 + 0]) + sh.x;
      v[4 * q4 + 1] = __uint_as_float(raw[4 * q4 + 1]) + sh.y;
      v[4 * q4 + 2] = __uint_as_float(raw[4 * q4 + 2]) + sh.z;
      v[4 * q4 + 3] = __uint_as_float(raw[4 * q4 + 3]) + sh.w;
    }
    if (stats && valid) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s1[j] += v[j];
        s2[j] = fmaf(v[j], v[j], s2[j]);
      }
    }
  } else {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 a = lds128(cA + 4 * q4), sh = lds128(cS + 4 * q4);
      v[4 * q4 + 0] = fmaf(a.x, __uint_as_float(raw[4 * q4 + 0]), sh.x);
      v[4 * q4 + 1] = fmaf(a.y, __uint_as_float(raw[4 * q4 + 1]), sh.y);
      v[4 * q4 + 2] = fmaf(a.z, __uint_as_float(raw[4 * q4 + 2]), sh.z);
      v[4 * q4 + 3] = fmaf(a.w, __uint_as_float(raw[4 * q4 + 3]), sh.w);
    }
    if (stats && valid) {     // statistics of conv + bias (the input of the following Instance/BatchNorm)
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 bb = lds128(cB + 4 * q4);
        const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float t = __uint_as_float(raw[4 * q4 + e]) + bv[e];
          s1[4 * q4 + e] += t;
          s2[4 * q4 + e] = fmaf(t, t, s2[4 * q4 + e]);
        }
      }
    }
  }
  if (!valid) return;
  const bool ident = neg == 1.f && !clamp0, relu = neg == 0.f && !clamp0;
  if (!ident && !relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float lo = fminf(v[j], 0.f), hi = fmaxf(v[j], 0.f);
      v[j] = hi + (clamp0 ? fmaxf(slope * lo, 0.f) : neg * lo);
    }
  }
  if (c_abs + 16 <= y_cn && (y_cs & 7) == 0) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm("max.bf16x2 %0, %1, %2;" : "=r"(w[j]) : "r"(w[j]), "r"(0u));
    }
    if ((reinterpret_cast<uintptr_t>(yrow) & 31) == 0) {
      // one 256-bit store = one full 32-byte sector per voxel (two 128-bit stores cost the LSU two half-sector writes)
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   ::"l"(yrow), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    } else {
      reinterpret_cast<uint4*>(yrow)[0] = make_uint4(w[0], w[1], w[2], w[3]);
      reinterpret_cast<uint4*>(yrow)[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c_abs + j < y_cn) yrow[j] = __float2bfloat16_rn(relu ? fmaxf(v[j], 0.f) : v[j]);
  }
}

__device__ __forceinline__ float act_neg(int act, float slope) { return act == COMA_ACT_NONE ? 1.f : (act == COMA_ACT_RELU ? 0.f : slope); }

// fp32-store variant of epi_chunk16 (COMA_BF16_F32OUT, per-tap kernel only): same arithmetic, the result is not rounded to bf16
__device__ __forceinline__ void epi_chunk16_f32(const uint32_t (&raw)[16], const float* cA, const float* cS, const float* cB, bool valid,
                                                float neg, bool clamp0, float slope, bool stats, float* s1, float* s2,
                                                float* yrow, int c_abs, int y_cn, bool unit) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float r = __uint_as_float(raw[j]);
    v[j] = unit ? r + cS[j] : fmaf(cA[j], r, cS[j]);
    if (stats && valid) {
      const float t = unit ? v[j] : r + cB[j];
      s1[j] += t;
      s2[j] = fmaf(t, t, s2[j]);
    }
  }
  if (!valid) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float lo = fminf(v[j], 0.f), hi = fmaxf(v[j], 0.f);
    v[j] = hi + (clamp0 ? fmaxf(slope * lo, 0.f) : neg * lo);
  }
  if (c_abs + 16 <= y_cn && (reinterpret_cast<uintptr_t>(yrow) & 15) == 0) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) reinterpret_cast<float4*>(yrow)[q4] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c_abs + j < y_cn) yrow[j] = v[j];
  }
}

// taps of a tile: ordinary conv -> all k^3 taps; transposed conv -> the taps that hit output parity class `cls`
struct Tap { int dd, dh, dw, widx; };
__device__ __forceinline__ int num_taps(const TcParams& p, int cls) {
  if (!p.transposed) return p.ksize * p.ksize * p.ksize;
  return (1 + (cls & 1)) * (1 + ((cls >> 1) & 1)) * (1 + ((cls >> 2) & 1));
}
__device__ __forceinline__ Tap get_tap(const TcParams& p, int cls, int t) {
  Tap r;
  if (!p.transposed) {
    const int K = p.ksize;
    const int kw = t % K, kh = (t / K) % K, kd = t / (K * K);
    r.dw = kw - p.pad; r.dh = kh - p.pad; r.dd = kd - p.pad; r.widx = t;
    return r;
  }
  // per dim: parity 0 -> (k=1, shift 0); parity 1 -> (k=0, shift +1), (k=2, shift 0)        [o = 2*i - 1 + k]
  const int pw = cls & 1, ph = (cls >> 1) & 1, pd = (cls >> 2) & 1;
  const int nw = 1 + pw, nh = 1 + ph;
  const int iw = t % nw, ih = (t / nw) % nh, id = t / (nw * nh);
  const int kw = pw ? (iw ? 2 : 0) : 1, kh = ph ? (ih ? 2 : 0) : 1, kd = pd ? (id ? 2 : 0) : 1;
  r.dw = (pw && !iw) ? 1 : 0; r.dh = (ph && !ih) ? 1 : 0; r.dd = (pd && !id) ? 1 : 0;
  r.widx = (kd * 3 + kh) * 3 + kw;
  return r;
}

struct TileCoord { int n_tile, cls, b, d0, h0, w0; };
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t) {
  TileCoord c;
  c.n_tile = t % p.n_tiles; t /= p.n_tiles;
  c.w0 = (t % p.tiles_w) * TW; t /= p.tiles_w;
  c.h0 = (t % p.tiles_h) * TH; t /= p.tiles_h;
  c.d0 = (t % p.tiles_d) * TD; t /= p.tiles_d;
  c.cls = t % p.classes; t /= p.classes;
  c.b = t;
  return c;
}

__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* tfull = empty + kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sstat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));      // [4 warps][NT][2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer =================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        const int nt = num_taps(p, tc.cls);
        const int es = p.transposed ? 1 : p.stride;
        for (int tap = 0; tap < nt; ++tap) {
          const Tap tp = get_tap(p, tc.cls, tap);
          const int cw = tc.w0 * es + tp.dw, ch = tc.h0 * es + tp.dh, cd = tc.d0 * es + tp.dd;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* a_dst = smem + (size_t)stage * p.stage_bytes;
            uint8_t* b_dst = a_dst + p.a_bytes;
            mbar_expect_tx(&full[stage], p.a_bytes + p.b_bytes);
            tma_load_5d(a_dst, &tmA, &full[stage], kc * p.KC, cw, ch, cd, tc.b);
            tma_load_2d(b_dst, &tmB, &full[stage], kc * p.KC, tp.widx * p.Cout + tc.n_tile * p.NT);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===================================
    {
      // warp-uniform loop (descriptors in uniform registers), one elected lane issues
      const bool leader = elect_one();
      // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), K-major A and B, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t layout = p.swz == 128 ? 2u : (p.swz == 64 ? 4u : 6u);
      const uint32_t hi = ((8u * (uint32_t)p.swz) >> 4) | (1u << 14) | (layout << 29);
      const uint32_t base_lo = ((smem_u32(smem) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t stage16 = p.stage_bytes >> 4, a16 = p.a_bytes >> 4;
      const int kk_n = p.KC >> 4;
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local) {
        // k-steps of this tile: taps x channel chunks (the parity class is the slowest-but-one tile index)
        int ksteps;
        if (!p.transposed) {
          ksteps = p.ksize * p.ksize * p.ksize * p.kchunks;
        } else {
          const int cls = (t / (p.n_tiles * p.tiles_w * p.tiles_h * p.tiles_d)) % p.classes;
          ksteps = num_taps(p, cls) * p.kchunks;
        }
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.NT);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = base_lo + (uint32_t)stage * stage16;
          const uint32_t b_lo = a_lo + a16;
          if (leader) {
            for (int kk = 0; kk < kk_n; ++kk) {
              // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in the (addr >> 4) field
              asm volatile(
                  "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                  "setp.ne.b32 p, %5, 0;\n\t"
                  "mov.b64 da, {%1, %3};\n\t"
                  "mov.b64 db, {%2, %3};\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                  ::"r"(d_tmem), "r"(a_lo + (uint32_t)(kk * 2)), "r"(b_lo + (uint32_t)(kk * 2)), "r"(hi), "r"(idesc),
                    "r"((ks | kk) ? 1u : 0u)
                  : "memory");
            }
            tc_commit(&empty[stage]);          // frees the smem stage when these MMAs retire
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (leader) tc_commit(&tfull[acc]);    // accumulator complete -> epilogue
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue (warps 2..5) =========================
    const int q = warp & 3;                  // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;           // GEMM row = voxel inside the tile
    const int lw = row % TW, lh = (row / TW) % TH, ld = row / (TW * TH);
    const float slope = p.slope ? __ldg(p.slope) : 0.f;
    const float neg = act_neg(p.act, slope);
    const bool clamp0 = p.act == COMA_ACT_LEAKY_RELU, do_stats = p.stats != nullptr;
    float* wstat = sstat + (size_t)(warp - 2) * p.NT * 2;
    float* cA = sstat + 8 * p.NT;
    float* cS = cA + p.NT;
    int local = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++local) {
      const TileCoord tc = decode_tile(p, t);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      // output voxel of this row
      int od, oh, ow;
      bool valid;
      if (!p.transposed) {
        od = tc.d0 + ld; oh = tc.h0 + lh; ow = tc.w0 + lw;
        valid = od < p.Do && oh < p.Ho && ow < p.Wo;
      } else {
        const int id = tc.d0 + ld, ih = tc.h0 + lh, iw = tc.w0 + lw;
        valid = id < p.Di && ih < p.Hi && iw < p.Wi;
        od = 2 * id + ((tc.cls >> 2) & 1); oh = 2 * ih + ((tc.cls >> 1) & 1); ow = 2 * iw + (tc.cls & 1);
      }
      const int n0 = tc.n_tile * p.NT;
      const int64_t yoff = ((((int64_t)tc.b * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.y_cs + n0;
      __nv_bfloat16* yrow = p.y + yoff;
      float* yrow_f = reinterpret_cast<float*>(p.y) + yoff;
      epi_stage_coef(p.bias, p.scale, p.shift, (int64_t)tc.b * p.Cout, n0, p.NT, cA, cS, (int)threadIdx.x - 64);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      for (int c0 = 0; c0 < p.NT; c0 += 16) {
        uint32_t raw[16];
        tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT + c0), raw);
        float s1c[16], s2c[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { s1c[j] = 0.f; s2c[j] = 0.f; }
        if (p.out_f32) epi_chunk16_f32(raw, cA + c0, cS + c0, cS + p.NT + c0, valid, neg, clamp0, slope, do_stats, s1c, s2c, yrow_f + c0, n0 + c0, p.y_cn, p.scale == nullptr);
        else epi_chunk16(raw, cA + c0, cS + c0, cS + p.NT + c0, valid, neg, clamp0, slope, do_stats, s1c, s2c, yrow + c0, n0 + c0, p.y_cn, p.y_cs, p.scale == nullptr);
        if (do_stats) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a1 = warp_sum(s1c[j]), a2 = warp_sum(s2c[j]);
            if (lane == 0) { wstat[(c0 + j) * 2] = a1; wstat[(c0 + j) * 2 + 1] = a2; }
          }
        }
      }
      // accumulator drained -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (!p.stats) asm volatile("bar.sync 1, 128;" ::: "memory");   // coefficients of this tile are no longer read
      if (p.stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // chunk index of this tile inside its sample
        int tt = t / p.n_tiles;
        const int per_sample = p.tiles_w * p.tiles_h * p.tiles_d * p.classes;
        const int chunk = tt % per_sample;
        for (int i = threadIdx.x - 64; i < p.NT * 2; i += 128) {
          const float s = sstat[i] + sstat[p.NT * 2 + i] + sstat[p.NT * 4 + i] + sstat[p.NT * 6 + i];
          p.stats[(((int64_t)tc.b * p.stat_chunks + chunk) * p.Cout + n0 + (i >> 1)) * 2 + (i & 1)] = s;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}


// ================================================================================================
// v2: halo-reuse kernel for 3x3x3 stride-1 convolutions with few channels (Cin, Cout <= 64).
//
// The per-tap kernel above re-reads every activation tile 27x from L2 and pays one TMA + two mbarrier round
// trips per tap; at 128^3 with 16..64 channels that, not the tensor core, sets the time.  Here a CTA sweeps a
// column of the volume plane by plane: M tile = 8(w) x 16(h) voxels of one d-plane, each input plane is loaded ONCE
// as a 10 x 18 halo slab (one TMA box, zero-filled outside the volume), kept in a ring of shared-memory slots,
// and used by all 27 taps of three consecutive output planes through shifted UMMA descriptors
// (start = slab + (kh*10 + kw) rows, 8-row groups 10 rows apart).  tcgen05 applies the 128B/64B/32B swizzle to
// absolute shared-memory address bits, so such unaligned starts read exactly what TMA wrote
// (profiles/r01_probe_umma_descriptor_addressing.log).  All 27 weight tiles stay resident in shared memory.
// ================================================================================================
constexpr int HW_T = 8, HH_T = 16, HALO_W = HW_T + 2, HALO_H = HH_T + 2, HALO_ROWS = HALO_W * HALO_H;
constexpr int kMaxSlabs = 8;

struct HaloParams {
  int B, D, H, W, Cin, Cout;
  int KC;
  int cols_w, cols_h, segs_d, DS, total_segs;
  int nslab;
  uint32_t rowb, slab_bytes, slab_tx, w_tile_bytes, w_bytes;
  uint32_t chunk_bytes, chunk_tx;   // per 64-channel chunk of a slab (v3 with Cin = 128: two chunks)
  __nv_bfloat16* y;
  int y_cs, y_cn;
  const float* bias; const float* scale; const float* shift; const float* slope;
  float* stats;
  int act, stat_chunks;
  uint32_t tmem_cols;
  const float* in_scale; const float* in_shift;   // input prologue (v3 stride-1 kernel): x' = max(u,0) + in_neg*min(u,0), u = in_scale*x + in_shift
  const float* in_slope;
  int in_act;
};

struct SegCoord { int b, d0, nd, h0, w0, chunk; };
__device__ __forceinline__ SegCoord decode_seg(const HaloParams& p, int t) {
  SegCoord c;
  const int per_sample = p.cols_w * p.cols_h * p.segs_d;
  c.b = t / per_sample;
  c.chunk = t % per_sample;
  int r = c.chunk;
  c.w0 = (r % p.cols_w) * HW_T; r /= p.cols_w;
  c.h0 = (r % p.cols_h) * HH_T; r /= p.cols_h;
  c.d0 = r * p.DS;
  c.nd = min(p.DS, p.D - c.d0);
  return c;
}

template <int NT, int KC>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wreg = smem;                                                   // 27 weight tiles
  uint8_t* slabs = smem + ((p.w_bytes + 1023u) & ~1023u);                 // nslab input-plane slabs
  uint64_t* sfull = reinterpret_cast<uint64_t*>(slabs + (size_t)p.nslab * p.slab_bytes);
  uint64_t* sempty = sfull + kMaxSlabs;
  uint64_t* wfull = sempty + kMaxSlabs;
  uint64_t* tfull = wfull + 1;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sstat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));                 // [4 warps][NT][2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslab; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 1); }
    mbar_init(wfull, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer =================================
    if (lane == 0) {
      mbar_expect_tx(wfull, p.w_bytes);
      for (int tap = 0; tap < 27; ++tap) tma_load_2d(wreg + (size_t)tap * p.w_tile_bytes, &tmB, wfull, 0, tap * p.Cout);
      uint32_t slot = 0, ph = 0;
      for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
        const SegCoord sc = decode_seg(p, t);
        for (int pi = 0; pi < sc.nd + 2; ++pi) {
          mbar_wait(&sempty[slot], ph ^ 1u);
          mbar_expect_tx(&sfull[slot], p.slab_tx);
          tma_load_5d(slabs + (size_t)slot * p.slab_bytes, &tmA, &sfull[slot], 0, sc.w0 - 1, sc.h0 - 1, sc.d0 - 1 + pi, sc.b);
          if (++slot == (uint32_t)p.nslab) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===================================
    // One thread, and its instruction stream is the critical path at small N (a 128xNx16 MMA occupies the tensor
    // pipe for only ~64 cycles): everything but two 32-bit adds per MMA is hoisted out of the fully unrolled tap loop.
    {
      // the whole warp runs the (warp-uniform) loop so descriptors live in uniform registers; one elected lane issues
      const bool leader = elect_one();
      constexpr uint32_t ROWB = KC * 2u;
      constexpr uint32_t LAYOUT = ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u);
      constexpr uint32_t A_HI = ((HALO_W * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);   // SBO = 10 rows, version 1, swizzle
      constexpr uint32_t B_HI = ((8u * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);        // SBO = 8 rows
      constexpr uint32_t W_TILE16 = (NT * ROWB) >> 4;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t w_lo = ((smem_u32(wreg) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t s_lo = ((smem_u32(slabs) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t slab16 = p.slab_bytes >> 4;
      const uint32_t nslab = (uint32_t)p.nslab;
      mbar_wait(wfull, 0);
      uint32_t slot0 = 0;                  // ring slot of the oldest plane (d-1) of the current output plane
      uint32_t wslot = 0, wph = 0;         // next plane whose "full" barrier has not been waited on yet
      uint32_t ahead = 0;                  // planes already waited on that belong to the current/future output planes
      int local = 0;
      for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
        const SegCoord sc = decode_seg(p, t);
        for (int i = 0; i < sc.nd; ++i, ++local) {
          while (ahead < 3u) {
            mbar_wait(&sfull[wslot], wph);
            if (++wslot == nslab) { wslot = 0; wph ^= 1u; }
            ++ahead;
          }
          const int acc = local & 1;
          mbar_wait(&tempty[acc], (((uint32_t)local >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NT);
          uint32_t sl = slot0;
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const uint32_t a_kd = s_lo + sl * slab16;
            if (++sl == nslab) sl = 0;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  const uint32_t a_lo = a_kd + (uint32_t)(((kh * HALO_W + kw) * ROWB + kk * 32u) >> 4);
                  const uint32_t b_lo = w_lo + (uint32_t)(((kd * 3 + kh) * 3 + kw) * W_TILE16 + kk * 2);
                  if (leader) asm volatile(
                      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                      "setp.ne.b32 p, %6, 0;\n\t"
                      "mov.b64 da, {%1, %2};\n\t"
                      "mov.b64 db, {%3, %4};\n\t"
                      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                      ::"r"(d_tmem), "r"(a_lo), "r"(A_HI), "r"(b_lo), "r"(B_HI), "r"(idesc), "r"((kd | kh | kw | kk) ? 1u : 0u)
                      : "memory");
                }
              }
            }
          }
          if (leader) {
            tc_commit(&tfull[acc]);
            tc_commit(&sempty[slot0]);                  // the oldest plane is no longer needed
          }
          __syncwarp();
          if (++slot0 == nslab) slot0 = 0;
          --ahead;
        }
        // the last two planes of the segment are not shared with the next segment
        if (leader) tc_commit(&sempty[slot0]);
        if (++slot0 == nslab) slot0 = 0;
        if (leader) tc_commit(&sempty[slot0]);
        if (++slot0 == nslab) slot0 = 0;
        __syncwarp();
        ahead -= 2u;
      }
    }
  } else {
    // ================================ epilogue (warps 2..5) =========================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lw = row % HW_T, lh = row / HW_T;
    const float slope = p.slope ? __ldg(p.slope) : 0.f;
    const float neg = act_neg(p.act, slope);
    const bool clamp0 = p.act == COMA_ACT_LEAKY_RELU, do_stats = p.stats != nullptr;
    float* cA = sstat + 8 * NT;
    float* cS = cA + NT;
    int local = 0;
    for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
      const SegCoord sc = decode_seg(p, t);
      const int oh = sc.h0 + lh, ow = sc.w0 + lw;
      const bool valid = oh < p.H && ow < p.W;
      epi_stage_coef(p.bias, p.scale, p.shift, (int64_t)sc.b * p.Cout, 0, NT, cA, cS, (int)threadIdx.x - 64);
      float s1[NT], s2[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
      for (int i = 0; i < sc.nd; ++i, ++local) {
        const int acc = local & 1;
        __nv_bfloat16* yrow = p.y + ((((int64_t)sc.b * p.D + (sc.d0 + i)) * p.H + oh) * p.W + ow) * p.y_cs;
        mbar_wait(&tfull[acc], ((uint32_t)local >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t raw[16];
          tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT + c0), raw);
          epi_chunk16(raw, cA + c0, cS + c0, cS + NT + c0, valid, neg, clamp0, slope, do_stats, s1 + c0, s2 + c0, yrow + c0, c0, p.y_cn, p.y_cs, p.scale == nullptr);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // all epilogue warps are done with this segment's coefficients
      if (p.stats) {   // one partial per segment: reduce the per-row running sums over the 128 rows
        float* wstat = sstat + (size_t)(warp - 2) * NT * 2;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float a = warp_sum(s1[j]), b2 = warp_sum(s2[j]);
          if (lane == 0) { wstat[j * 2] = a; wstat[j * 2 + 1] = b2; }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = threadIdx.x - 64; i < NT * 2; i += 128) {
          const float s = sstat[i] + sstat[NT * 2 + i] + sstat[NT * 4 + i] + sstat[NT * 6 + i];
          p.stats[(((int64_t)sc.b * p.stat_chunks + sc.chunk) * p.Cout + (i >> 1)) * 2 + (i & 1)] = s;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}


// ================================================================================================
// v3: halo-reuse kernel with the kd taps fused into the MMA N dimension ("input-plane driven").
//
// Measured (profiles/r01_probe_mma_rate_ss_vs_ts.log): a 128 x N x 16 SS-mode MMA costs 40 / 48 / 64 cycles at
// N = 32 / 64 / 128 -- the A-operand (activation) read from shared memory is the floor, so widening N is almost free.
// An input plane p, shifted by (kh,kw), contributes to THREE output planes: p+1 through W[kd=0], p through W[kd=1],
// p-1 through W[kd=2].  With the weight tiles stored [kh][kw][kd][co][ci], one MMA with N = 3*Cout and the same A
// descriptor produces all three contributions into three neighbouring accumulator blocks of a TMEM ring
// (8 blocks of Cout columns, block(d_out) = (-d_out) mod 8), i.e. a third of the A reads and MMA issues of v2.
// The first contribution to an output plane is always its kd=0 block, issued separately with accumulate=0.
// ================================================================================================
constexpr int kRing = 8;

// All 27 taps of one input plane for the steady state (every kd valid).  The ring position S only matters through the
// TMEM column of block S and through whether the three target blocks wrap around the ring (WRAP = 0: S <= 5, no wrap;
// 1: S = 6; 2: S = 7), so three instantiations cover all eight positions; the tap loop stays rolled over (kh,kw) to keep
// the single issuing warp's code inside the instruction cache (a fully unrolled 8-variant version thrashed it at Cin = 128).
template <int NT, int KC, int KCH, int WRAP>
__device__ __forceinline__ void halo3_issue_steady(bool leader, uint32_t tmem_base, uint32_t S, uint32_t a_pl, uint32_t w_lo, uint32_t chunk16) {
  constexpr uint32_t ROWB = KC * 2u;
  constexpr uint32_t LAYOUT = ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u);
  constexpr uint32_t A_HI = ((HALO_W * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
  constexpr uint32_t B_HI = ((8u * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
  constexpr uint32_t W_TILE16 = (NT * ROWB) >> 4;
  constexpr uint32_t IDESC0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
  constexpr int nA = 3 - WRAP, nB = WRAP;                         // kd 0..2 from block S: nA blocks, then nB from block 0
  constexpr int nF = WRAP == 1 ? 1 : 2, nG = 2 - nF;              // kd 1..2 from block S+1 (S = 7: block 0, no wrap)
  const uint32_t colS = tmem_base + S * NT;
  const uint32_t colF = tmem_base + ((S + 1u) & (kRing - 1)) * NT;
#define COMA_MMA(COLADDR, NBLK, ALO, BLO, ACC)                                                                     \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\t"     \
               "mov.b64 db, {%3, %4};\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"             \
               ::"r"(COLADDR), "r"(ALO), "r"(A_HI), "r"(BLO), "r"(B_HI),                                           \
                 "r"(IDESC0 | ((((uint32_t)(NBLK) * NT) >> 3) << 17)), "r"((uint32_t)(ACC)) : "memory")
  // first MMA of the plane: the kd = 0 block starts a new output plane (accumulate = 0), kd = 1,2 keep accumulating
  if (leader) {
    COMA_MMA(colS, 1, a_pl, w_lo, 0);
    COMA_MMA(colF, nF, a_pl, w_lo + W_TILE16, 1);
    if (nG) COMA_MMA(tmem_base, nG, a_pl, w_lo + (1 + nF) * W_TILE16, 1);
  }
  constexpr int UNROLL_KH = (KCH * (KC / 16) <= 2) ? 3 : 1;       // tiny bodies: unroll all 9 taps; big ones: keep the code small
#pragma unroll UNROLL_KH
  for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const uint32_t a_t9 = a_pl + (uint32_t)(kh * HALO_W + kw) * (ROWB >> 4);
      const uint32_t b_t9 = w_lo + (uint32_t)((kh * 3 + kw) * KCH * 3) * W_TILE16;
#pragma unroll
      for (int kc = 0; kc < KCH; ++kc) {
#pragma unroll
        for (int kk = 0; kk < KC / 16; ++kk) {
          const uint32_t a_lo = a_t9 + (uint32_t)kc * chunk16 + (uint32_t)(kk * 2);
          const uint32_t b_lo = b_t9 + (uint32_t)(kc * 3) * W_TILE16 + (uint32_t)(kk * 2);
          if (leader && (kh | kw | kc | kk) != 0) {                // the very first MMA was issued above
            COMA_MMA(colS, nA, a_lo, b_lo, 1);
            if (nB) COMA_MMA(tmem_base, nB, a_lo, b_lo + nA * W_TILE16, 1);
          }
        }
      }
    }
  }
#undef COMA_MMA
}

// Epilogue shared by the plane-ring kernels (v3 and the stride-2 kernel): output plane i of a segment sits in ring block
// (s_seg - i) mod 8 and the ring position advances by nd + SEG_EXTRA per segment (SEG_EXTRA = the extra input planes the
// issuer steps through per segment: 2 for v3, 0 for stride 2, which counts output planes).
template <int NT, int EPI, int SEG_EXTRA>
__device__ __forceinline__ void ring_epilogue(const HaloParams& p, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty, float* sstat,
                                              int n0, int warp, int lane) {
  // ================================ epilogue (warps 2..5 [, 6..9]) =========================
  const int q = warp & 3;
  const int grp = (warp - 2) >> 2;                       // epilogue warpgroup: drains output planes with ocount % EPI == grp
  const int tid128 = ((int)threadIdx.x - 64) & 127;
  const uint32_t bar_id = 1u + (uint32_t)grp;
  const int row = q * 32 + lane;
  const int lw = row % HW_T, lh = row / HW_T;
  const float slope = p.slope ? __ldg(p.slope) : 0.f;
  const float neg = act_neg(p.act, slope);
  const bool clamp0 = p.act == COMA_ACT_LEAKY_RELU, do_stats = p.stats != nullptr;
  float* gstat = sstat + (size_t)grp * 11 * NT;          // per group: [4 warps][NT][2] partial sums, then cA, cS, cB
  float* cA = gstat + 8 * NT;
  float* cS = cA + NT;
  uint32_t ocount = 0;
  uint32_t s_seg = 0, ebits = 0;
  for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
    const SegCoord sc = decode_seg(p, t);
    const int oh = sc.h0 + lh, ow = sc.w0 + lw;
    const bool valid = oh < p.H && ow < p.W;
    epi_stage_coef(p.bias, p.scale, p.shift, (int64_t)sc.b * p.Cout, n0, NT, cA, cS, tid128, bar_id);
    float s1[NT], s2[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    for (int i = 0; i < sc.nd; ++i, ++ocount) {
      const uint32_t blk = (s_seg + (uint32_t)(kRing * 64 - i)) & (kRing - 1);
      if (EPI > 1 && (ocount % EPI) != (uint32_t)grp) {    // the other warpgroup's plane: only track the barrier phase
        ebits ^= 1u << blk;
        continue;
      }
      __nv_bfloat16* yrow = p.y + ((((int64_t)sc.b * p.D + (sc.d0 + i)) * p.H + oh) * p.W + ow) * p.y_cs + n0;
      mbar_wait(&tfull[blk], (ebits >> blk) & 1u);
      ebits ^= 1u << blk;
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < NT; c0 += 16) {
        uint32_t raw[16];
        tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + blk * NT + (uint32_t)c0, raw);
        epi_chunk16(raw, cA + c0, cS + c0, cS + NT + c0, valid, neg, clamp0, slope, do_stats, s1 + c0, s2 + c0, yrow + c0, n0 + c0, p.y_cn, p.y_cs, p.scale == nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[blk]);
    }
    s_seg = (s_seg + (uint32_t)(kRing * 64 - (sc.nd + SEG_EXTRA))) & (kRing - 1);
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // this group is done with the segment's coefficients
    if (p.stats) {
      float* wstat = gstat + (size_t)q * NT * 2;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float a = warp_sum(s1[j]), b2 = warp_sum(s2[j]);
        if (lane == 0) { wstat[j * 2] = a; wstat[j * 2 + 1] = b2; }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      for (int i = tid128; i < NT * 2; i += 128) {
        const float sm = gstat[i] + gstat[NT * 2 + i] + gstat[NT * 4 + i] + gstat[NT * 6 + i];
        p.stats[(((int64_t)sc.b * p.stat_chunks + sc.chunk * EPI + grp) * p.Cout + n0 + (i >> 1)) * 2 + (i & 1)] = sm;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    }
  }

}

// EPI epilogue warpgroups (4 warps each) drain alternate output planes: at N <= 32 the epilogue, not the MMA side, was the limit.
// CPS = CTAs per SM: with few channels the single MMA-issuing warp's instruction latency (not the tensor pipe, not the epilogue) sets
// the pace, so two small CTAs per SM (EPI = 1, half the shared memory and TMEM each) give two independent issue streams.
template <int NT, int KC, int KCH, int EPI, int CPS>
__global__ void __launch_bounds__(64 + 128 * EPI, CPS)
conv_halo3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wreg = smem;                                                   // 9 x [3 kd][NT][KC] weight tiles
  uint8_t* slabs = smem + ((p.w_bytes + 1023u) & ~1023u);
  uint64_t* sfull = reinterpret_cast<uint64_t*>(slabs + (size_t)p.nslab * p.slab_bytes);
  uint64_t* sempty = sfull + kMaxSlabs;
  uint64_t* sready = sempty + kMaxSlabs;                                 // prologue: slab transformed in place, MMA may read it
  uint64_t* wfull = sready + kMaxSlabs;
  uint64_t* tfull = wfull + 1;
  uint64_t* tempty = tfull + kRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kRing);
  float* sstat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
  float* xcoef = sstat + 22 * NT;                                         // [2][64]: prologue scale, shift of the current sample
  constexpr int XW = 1;                                                   // transform warps (a second one cost more in registers than it gained)
  const bool pro = KCH == 1 && CPS > 1 && KC <= 32 && p.in_scale != nullptr;     // only the small-channel (multi-CTA) variants carry the prologue

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * NT;          // this CTA's slice of the output channels (Cout split when the weights do not fit)
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslab; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 1); mbar_init(&sready[s], XW); }
    mbar_init(wfull, 1);
    for (int a = 0; a < kRing; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

// ---- input prologue (transform warps: warp 0 and, in the multi-CTA variants, one helper warp after the epilogue warps)
auto prologue_walk = [&](const int widx) {
    // Each transform warp walks the plane sequence; warp 0 also runs LAG planes ahead issuing the TMA loads (lane 0).  The warps
    // rewrite an arrived slab in place (x' = act(scale*x + shift) on in-bounds voxels; the zero-filled halo outside the
    // volume IS the padding of the activated tensor and stays zero) and hand it to the MMA warp through sready.
    constexpr uint32_t ROWB = KC * 2u, CPR = ROWB / 16u, SWM = CPR - 1u;     // chunks (16 B) per row; swizzle mask on the chunk index
    struct Cur { int t, pi; SegCoord sc; uint32_t slot, ph; bool valid; };
    auto cur_init = [&](Cur& c) { c.t = blockIdx.x; c.pi = 0; c.slot = 0; c.ph = 0; c.valid = c.t < p.total_segs; if (c.valid) c.sc = decode_seg(p, c.t); };
    auto cur_next = [&](Cur& c) {
      if (++c.slot == (uint32_t)p.nslab) { c.slot = 0; c.ph ^= 1u; }
      if (++c.pi == c.sc.nd + 2) { c.pi = 0; c.t += gridDim.x; c.valid = c.t < p.total_segs; if (c.valid) c.sc = decode_seg(p, c.t); }
    };
    const float in_neg = p.in_act == COMA_ACT_NONE ? 1.f : (p.in_act == COMA_ACT_RELU ? 0.f : __ldg(p.in_slope));
    const int lag = min(p.nslab - 1, 4);
    Cur is, xf;
    cur_init(is);
    cur_init(xf);
    int ahead = 0, staged_b = -1;
    uint32_t av2[4] = {0, 0, 0, 0}, bv2[4] = {0, 0, 0, 0};      // packed bf16x2 scale / shift of this lane's channel group
    const uint32_t neg2 = pack_bf16x2(in_neg, in_neg);
    while (xf.valid) {
      while (is.valid && ahead < lag) {
        if (widx == 0 && lane == 0) {
          mbar_wait(&sempty[is.slot], is.ph ^ 1u);
          mbar_expect_tx(&sfull[is.slot], p.slab_tx);
          tma_load_5d(slabs + (size_t)is.slot * p.slab_bytes, &tmA, &sfull[is.slot], 0, is.sc.w0 - 1, is.sc.h0 - 1, is.sc.d0 - 1 + is.pi, is.sc.b);
        }
        cur_next(is);
        ++ahead;
      }
      __syncwarp();
      // lane L always meets the same logical 8-channel group: chunk c = L + 32 k has c % CPR = L % CPR and swizzle term
      // (c >> 3) % CPR = (L >> 3) % CPR for CPR <= 4, so its 8 scales / shifts live in registers per sample
      if (xf.sc.b != staged_b) {
        staged_b = xf.sc.b;
        const uint32_t jl = ((uint32_t)lane & SWM) ^ (((uint32_t)lane >> 3) & SWM);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* sc2 = p.in_scale + (int64_t)staged_b * p.Cin + jl * 8 + 2 * j;
          const float* sh2 = p.in_shift + (int64_t)staged_b * p.Cin + jl * 8 + 2 * j;
          av2[j] = pack_bf16x2(__ldg(sc2), __ldg(sc2 + 1));
          bv2[j] = pack_bf16x2(__ldg(sh2), __ldg(sh2 + 1));
        }
      }
      mbar_wait(&sfull[xf.slot], xf.ph);
      const int d = xf.sc.d0 - 1 + xf.pi;
      if (d >= 0 && d < p.D) {
        uint8_t* slab = slabs + (size_t)xf.slot * p.slab_bytes;
        constexpr uint32_t NCH = (uint32_t)HALO_ROWS * CPR;
        for (uint32_t c0 = lane + 32u * widx; c0 < NCH; c0 += 128u * XW) {       // four chunks per lane in flight
          uint4 q[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t c = c0 + 32u * XW * u, r = c / CPR;
            const int gh = xf.sc.h0 - 1 + (int)(r / HALO_W), gw = xf.sc.w0 - 1 + (int)(r % HALO_W);
            ok[u] = c < NCH && gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
            if (ok[u]) q[u] = *reinterpret_cast<const uint4*>(slab + c * 16u);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            uint32_t wd[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // packed bf16x2: u = fma(scale, x, shift) rounded once, then max(u,0) + neg * min(u,0)
              uint32_t uu, lo, hi;
              asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(uu) : "r"(av2[j]), "r"(wd[j]), "r"(bv2[j]));
              asm("min.bf16x2 %0, %1, %2;" : "=r"(lo) : "r"(uu), "r"(0u));
              asm("max.bf16x2 %0, %1, %2;" : "=r"(hi) : "r"(uu), "r"(0u));
              asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(wd[j]) : "r"(neg2), "r"(lo), "r"(hi));
            }
            *reinterpret_cast<uint4*>(slab + (c0 + 32u * XW * u) * 16u) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&sready[xf.slot]);
      cur_next(xf);
      --ahead;
    }
  };

  if (warp == 0) {
    // ================================ TMA producer =================================
    if (lane == 0) {
      mbar_expect_tx(wfull, p.w_bytes);
      for (int t9 = 0; t9 < 9; ++t9)
        for (int kc = 0; kc < KCH; ++kc)
          for (int kd = 0; kd < 3; ++kd)
            tma_load_2d(wreg + (size_t)((t9 * KCH + kc) * 3 + kd) * p.w_tile_bytes, &tmB, wfull, kc * KC, (kd * 9 + t9) * p.Cout + n0);
      if (!pro) {
        uint32_t slot = 0, ph = 0;
        for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
          const SegCoord sc = decode_seg(p, t);
          for (int pi = 0; pi < sc.nd + 2; ++pi) {
            mbar_wait(&sempty[slot], ph ^ 1u);
            mbar_expect_tx(&sfull[slot], p.slab_tx);
            for (int kc = 0; kc < KCH; ++kc)
              tma_load_5d(slabs + (size_t)slot * p.slab_bytes + (size_t)kc * p.chunk_bytes, &tmA, &sfull[slot], kc * KC, sc.w0 - 1, sc.h0 - 1,
                          sc.d0 - 1 + pi, sc.b);
            if (++slot == (uint32_t)p.nslab) { slot = 0; ph ^= 1u; }
          }
        }
      }
    }
    if (pro) prologue_walk(0);
  } else if (warp == 1) {
    // ================================ MMA issuer (warp-uniform loop, one elected lane issues) ============
    const bool leader = elect_one();
    constexpr uint32_t ROWB = KC * 2u;
    constexpr uint32_t LAYOUT = ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u);
    constexpr uint32_t A_HI = ((HALO_W * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
    constexpr uint32_t B_HI = ((8u * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
    constexpr uint32_t W_TILE16 = (NT * ROWB) >> 4;
    constexpr uint32_t IDESC0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t w_lo = ((smem_u32(wreg) & 0x3FFFFu) >> 4) | 0x10000u;
    const uint32_t s_lo = ((smem_u32(slabs) & 0x3FFFFu) >> 4) | 0x10000u;
    const uint32_t slab16 = p.slab_bytes >> 4, chunk16 = p.chunk_bytes >> 4;
    const uint32_t nslab = (uint32_t)p.nslab;
    mbar_wait(wfull, 0);
    uint32_t s = 0;            // ring block of the kd=0 target (output plane d0+pi) of the current input plane
    uint32_t pbits = 0;        // per-block phase of tempty
    uint32_t sslot = 0, sph = 0;
    for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
      const SegCoord sc = decode_seg(p, t);
      for (int pi = 0; pi < sc.nd + 2; ++pi) {
        mbar_wait(pro ? &sready[sslot] : &sfull[sslot], sph);
        const int kd_lo = max(0, pi - sc.nd + 1), kd_hi = min(2, pi);
        if (kd_lo == 0) {       // block s starts a new output plane: its previous tenant must have been drained
          mbar_wait(&tempty[s], ((pbits >> s) & 1u) ^ 1u);
          pbits ^= 1u << s;
        }
        tc_fence_after();
        if (kd_lo == 0 && kd_hi == 2) {       // steady state: dispatch on the ring position, everything else is immediates
          const uint32_t a_pl0 = s_lo + sslot * slab16;
          if (s <= 5u) halo3_issue_steady<NT, KC, KCH, 0>(leader, tmem_base, s, a_pl0, w_lo, chunk16);
          else if (s == 6u) halo3_issue_steady<NT, KC, KCH, 1>(leader, tmem_base, s, a_pl0, w_lo, chunk16);
          else halo3_issue_steady<NT, KC, KCH, 2>(leader, tmem_base, s, a_pl0, w_lo, chunk16);
          if (leader) {
            tc_commit(&tfull[(s + 2u) & (kRing - 1)]);
            tc_commit(&sempty[sslot]);
          }
          __syncwarp();
          if (++sslot == nslab) { sslot = 0; sph ^= 1u; }
          s = (s + kRing - 1) & (kRing - 1);
          continue;
        }
        // runs of contiguous ring blocks: [ga .. ga+na) then (after the wrap) [0 .. nb)
        auto make_run = [&](int k0, int k1, uint32_t& colA, uint32_t& nA, uint32_t& kA, uint32_t& nB, uint32_t& kB) {
          const int n = k1 - k0 + 1;
          const uint32_t slot_a = (s + (uint32_t)k0) & (kRing - 1);
          colA = slot_a * NT; kA = (uint32_t)k0;
          nA = n > 0 ? (uint32_t)min(n, (int)(kRing - slot_a)) : 0u;
          nB = n > 0 ? (uint32_t)n - nA : 0u;
          kB = (uint32_t)k0 + nA;
        };
        uint32_t colA, nA, kA, nB, kB;                  // general steps: all valid kd, accumulate
        make_run(kd_lo, kd_hi, colA, nA, kA, nB, kB);
        uint32_t colF, nF, kF, nG, kG;                  // first step: kd >= 1 part (kd = 0 goes alone with accumulate = 0)
        make_run(max(kd_lo, 1), kd_hi, colF, nF, kF, nG, kG);
        const uint32_t a_pl = s_lo + sslot * slab16;
        auto mma = [&](uint32_t col, uint32_t nblk, uint32_t a_lo, uint32_t b_lo, uint32_t accum) {
          const uint32_t idesc = IDESC0 | (((nblk * NT) >> 3) << 17);
          asm volatile(
              "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
              "setp.ne.b32 p, %6, 0;\n\t"
              "mov.b64 da, {%1, %2};\n\t"
              "mov.b64 db, {%3, %4};\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
              ::"r"(tmem_base + col), "r"(a_lo), "r"(A_HI), "r"(b_lo), "r"(B_HI), "r"(idesc), "r"(accum)
              : "memory");
        };
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) {
#pragma unroll
          for (int kck = 0; kck < KCH * (KC / 16); ++kck) {
            const int kc = kck / (KC / 16), kk = kck % (KC / 16);
            const uint32_t a_lo = a_pl + (uint32_t)kc * chunk16 + (uint32_t)((((t9 / 3) * HALO_W + (t9 % 3)) * ROWB + kk * 32u) >> 4);
            const uint32_t b_t9 = w_lo + (uint32_t)((t9 * KCH + kc) * 3) * W_TILE16 + (uint32_t)(kk * 2);
            if (leader) {
              if (t9 == 0 && kck == 0 && kd_lo == 0) {
                mma(s * NT, 1u, a_lo, b_t9, 0u);
                if (nF) mma(colF, nF, a_lo, b_t9 + kF * W_TILE16, 1u);
                if (nG) mma(0u, nG, a_lo, b_t9 + kG * W_TILE16, 1u);
              } else {
                mma(colA, nA, a_lo, b_t9 + kA * W_TILE16, 1u);
                if (nB) mma(0u, nB, a_lo, b_t9 + kB * W_TILE16, 1u);
              }
            }
          }
        }
        if (leader) {
          if (pi >= 2) tc_commit(&tfull[(s + 2u) & (kRing - 1)]);   // output plane d0+pi-2 has all 27 taps
          tc_commit(&sempty[sslot]);
        }
        __syncwarp();
        if (++sslot == nslab) { sslot = 0; sph ^= 1u; }
        s = (s + kRing - 1) & (kRing - 1);
      }
    }
  } else if (warp < 2 + 4 * EPI) {
    ring_epilogue<NT, EPI, 2>(p, tmem_base, tfull, tempty, sstat, n0, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ================================================================================================
// v2-T: halo-reuse kernel for the transposed convolution (k3, s2, p1, op1), Cin, Cout <= 64.
// M tile = 8(w) x 16(h) voxels of one INPUT plane; the 9 x 17 slab of that plane (+1 halo on the high side) and of
// the next plane serve all 27 taps; the 8 output-parity classes accumulate in 8 TMEM accumulators
// (class (pd,ph,pw): per dim parity 0 -> tap k=1 shift 0, parity 1 -> taps k=0 shift +1 and k=2 shift 0) and the
// epilogue scatters each to its strided output voxels (2d+pd, 2h+ph, 2w+pw).
// ================================================================================================
constexpr int HT_W = HW_T + 1, HT_H = HH_T + 1, HT_ROWS = HT_W * HT_H;

// EPI epilogue warpgroups split the 8 parity classes of every tile (group g drains classes with pd = g): the 8x larger output makes
// the epilogue, not the MMA side, the limit of this kernel.
template <int NT, int KC, int EPI>
__global__ void __launch_bounds__(64 + 128 * EPI, 1)
convT_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  constexpr int NACC = (8 * NT * 2 <= 512) ? 2 : 1;     // accumulator sets (8 classes each)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wreg = smem;
  uint8_t* slabs = smem + ((p.w_bytes + 1023u) & ~1023u);
  uint64_t* sfull = reinterpret_cast<uint64_t*>(slabs + (size_t)p.nslab * p.slab_bytes);
  uint64_t* sempty = sfull + kMaxSlabs;
  uint64_t* wfull = sempty + kMaxSlabs;
  uint64_t* tfull = wfull + 1;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sstat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslab; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 1); }
    mbar_init(wfull, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4 * EPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, p.w_bytes);
      for (int tap = 0; tap < 27; ++tap) tma_load_2d(wreg + (size_t)tap * p.w_tile_bytes, &tmB, wfull, 0, tap * p.Cout);
      uint32_t slot = 0, ph = 0;
      for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
        const SegCoord sc = decode_seg(p, t);
        for (int pi = 0; pi < sc.nd + 1; ++pi) {
          mbar_wait(&sempty[slot], ph ^ 1u);
          mbar_expect_tx(&sfull[slot], p.slab_tx);
          tma_load_5d(slabs + (size_t)slot * p.slab_bytes, &tmA, &sfull[slot], 0, sc.w0, sc.h0, sc.d0 + pi, sc.b);
          if (++slot == (uint32_t)p.nslab) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = elect_one();
      constexpr uint32_t ROWB = KC * 2u;
      constexpr uint32_t LAYOUT = ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u);
      constexpr uint32_t A_HI = ((HT_W * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
      constexpr uint32_t B_HI = ((8u * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
      constexpr uint32_t W_TILE16 = (NT * ROWB) >> 4;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t w_lo = ((smem_u32(wreg) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t s_lo = ((smem_u32(slabs) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t slab16 = p.slab_bytes >> 4;
      const uint32_t nslab = (uint32_t)p.nslab;
      mbar_wait(wfull, 0);
      uint32_t slot0 = 0, wslot = 0, wph = 0, ahead = 0;
      int local = 0;
      for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
        const SegCoord sc = decode_seg(p, t);
        for (int i = 0; i < sc.nd; ++i, ++local) {
          while (ahead < 2u) {
            mbar_wait(&sfull[wslot], wph);
            if (++wslot == nslab) { wslot = 0; wph ^= 1u; }
            ++ahead;
          }
          const int acc = NACC == 2 ? (local & 1) : 0;
          const uint32_t par = NACC == 2 ? (((uint32_t)local >> 1) & 1u) : ((uint32_t)local & 1u);
          mbar_wait(&tempty[acc], par ^ 1u);
          tc_fence_after();
          const uint32_t d_base = tmem_base + (uint32_t)(acc * 8 * NT);
          const uint32_t a_pl0 = s_lo + slot0 * slab16;
          const uint32_t a_pl1 = s_lo + ((slot0 + 1u == nslab) ? 0u : slot0 + 1u) * slab16;
#pragma unroll
          for (int cls = 0; cls < 8; ++cls) {
            const int pw = cls & 1, ph = (cls >> 1) & 1, pd = (cls >> 2) & 1;
            bool first = true;
#pragma unroll
            for (int id = 0; id <= pd; ++id) {
#pragma unroll
              for (int ih = 0; ih <= ph; ++ih) {
#pragma unroll
                for (int iw = 0; iw <= pw; ++iw) {
                  const int kd = pd ? (id ? 2 : 0) : 1, kh = ph ? (ih ? 2 : 0) : 1, kw = pw ? (iw ? 2 : 0) : 1;
                  const int sd = (pd && !id) ? 1 : 0, sh = (ph && !ih) ? 1 : 0, sw = (pw && !iw) ? 1 : 0;
                  const uint32_t a_pl = sd ? a_pl1 : a_pl0;
#pragma unroll
                  for (int kk = 0; kk < KC / 16; ++kk) {
                    const uint32_t a_lo = a_pl + (uint32_t)(((sh * HT_W + sw) * ROWB + kk * 32u) >> 4);
                    const uint32_t b_lo = w_lo + (uint32_t)(((kd * 3 + kh) * 3 + kw) * W_TILE16 + kk * 2);
                    if (leader) asm volatile(
                        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                        "setp.ne.b32 p, %6, 0;\n\t"
                        "mov.b64 da, {%1, %2};\n\t"
                        "mov.b64 db, {%3, %4};\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                        ::"r"(d_base + (uint32_t)(cls * NT)), "r"(a_lo), "r"(A_HI), "r"(b_lo), "r"(B_HI), "r"(idesc),
                          "r"((first && kk == 0) ? 0u : 1u)
                        : "memory");
                  }
                  first = false;
                }
              }
            }
          }
          if (leader) {
            tc_commit(&tfull[acc]);
            tc_commit(&sempty[slot0]);
          }
          __syncwarp();
          if (++slot0 == nslab) slot0 = 0;
          --ahead;
        }
        if (leader) tc_commit(&sempty[slot0]);       // the segment's last plane (only its "+1" role was used)
        __syncwarp();
        if (++slot0 == nslab) slot0 = 0;
        --ahead;
      }
    }
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int tid128 = ((int)threadIdx.x - 64) & 127;
    const uint32_t bar_id = 1u + (uint32_t)grp;
    const int row = q * 32 + lane;
    const int lw = row % HW_T, lh = row / HW_T;
    const float slope = p.slope ? __ldg(p.slope) : 0.f;
    const float neg = act_neg(p.act, slope);
    const bool clamp0 = p.act == COMA_ACT_LEAKY_RELU, do_stats = p.stats != nullptr;
    float* gstat = sstat + (size_t)grp * 11 * NT;
    float* cA = gstat + 8 * NT;
    float* cS = cA + NT;
    const int Do = 2 * p.D, Ho = 2 * p.H, Wo = 2 * p.W;
    int local = 0;
    for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
      const SegCoord sc = decode_seg(p, t);
      const int ih = sc.h0 + lh, iw = sc.w0 + lw;
      const bool valid = ih < p.H && iw < p.W;
      epi_stage_coef(p.bias, p.scale, p.shift, (int64_t)sc.b * p.Cout, 0, NT, cA, cS, tid128, bar_id);
      float s1[NT], s2[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
      for (int i = 0; i < sc.nd; ++i, ++local) {
        const int acc = NACC == 2 ? (local & 1) : 0;
        const uint32_t par = NACC == 2 ? (((uint32_t)local >> 1) & 1u) : ((uint32_t)local & 1u);
        mbar_wait(&tfull[acc], par);
        tc_fence_after();
#pragma unroll 1
        for (int cls = grp * (8 / EPI); cls < (grp + 1) * (8 / EPI); ++cls) {
          const int od = 2 * (sc.d0 + i) + ((cls >> 2) & 1), oh = 2 * ih + ((cls >> 1) & 1), ow = 2 * iw + (cls & 1);
          __nv_bfloat16* yrow = p.y + ((((int64_t)sc.b * Do + od) * Ho + oh) * Wo + ow) * p.y_cs;
#pragma unroll
          for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t raw[16];
            tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 8 * NT + cls * NT + c0), raw);
            epi_chunk16(raw, cA + c0, cS + c0, cS + NT + c0, valid, neg, clamp0, slope, do_stats, s1 + c0, s2 + c0, yrow + c0, c0, p.y_cn, p.y_cs, p.scale == nullptr);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // this group is done with the segment's coefficients
      if (p.stats) {
        float* wstat = gstat + (size_t)q * NT * 2;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float a = warp_sum(s1[j]), b2 = warp_sum(s2[j]);
          if (lane == 0) { wstat[j * 2] = a; wstat[j * 2 + 1] = b2; }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        for (int i = tid128; i < NT * 2; i += 128) {
          const float s = gstat[i] + gstat[NT * 2 + i] + gstat[NT * 4 + i] + gstat[NT * 6 + i];
          p.stats[(((int64_t)sc.b * p.stat_chunks + sc.chunk * EPI + grp) * p.Cout + (i >> 1)) * 2 + (i & 1)] = s;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ================================================================================================
// v3-S2: plane-ring kernel for the stride-2 (k3, p1) down-sampling convolutions.
// M tile = 8(w) x 16(h) OUTPUT voxels.  An input plane z is staged as four (h,w)-parity sub-slabs of 9 x 17 voxels
// (TMA element strides 2 on W and H; sub-slab (ph,pw) starts at input (2 h0 - ph, 2 w0 - pw), out-of-range rows are
// zero-filled = the padding), so tap kh reads parity ph = (kh != 1) at row shift (kh == 2), same along w: every tap
// is again a row-shifted window of a resident slab.  Along depth the issuer walks input planes 2 d0 - 1 .. 2 d0 + 2 nd - 1:
// an even plane 2 d feeds output d through kd = 1; an odd plane 2 d + 1 feeds output d + 1 through kd = 0 and output d
// through kd = 2 with ONE MMA of N = 2 Cout (weight tiles stored [kh][kw][kd = 0, 2, 1], ring block(d) = (-d) mod 8),
// so each plane is loaded and read once.
// ================================================================================================
constexpr int S2_ROWS = HT_ROWS;      // 9 x 17 voxels per parity sub-slab

template <int NT, int KC, int EPI>
__global__ void __launch_bounds__(64 + 128 * EPI, 1)
conv_halo_s2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wreg = smem;                                                   // 9 x [kd = 0, 2, 1][NT][KC] weight tiles
  uint8_t* slabs = smem + ((p.w_bytes + 1023u) & ~1023u);
  uint64_t* sfull = reinterpret_cast<uint64_t*>(slabs + (size_t)p.nslab * p.slab_bytes);
  uint64_t* sempty = sfull + kMaxSlabs;
  uint64_t* wfull = sempty + kMaxSlabs;
  uint64_t* tfull = wfull + 1;
  uint64_t* tempty = tfull + kRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kRing);
  float* sstat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * NT;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslab; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 1); }
    mbar_init(wfull, 1);
    for (int a = 0; a < kRing; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer =================================
    if (lane == 0) {
      mbar_expect_tx(wfull, p.w_bytes);
      for (int t9 = 0; t9 < 9; ++t9)
        for (int j = 0; j < 3; ++j) {
          const int kd = j == 0 ? 0 : (j == 1 ? 2 : 1);
          tma_load_2d(wreg + (size_t)(t9 * 3 + j) * p.w_tile_bytes, &tmB, wfull, 0, (kd * 9 + t9) * p.Cout + n0);
        }
      uint32_t slot = 0, ph = 0;
      for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
        const SegCoord sc = decode_seg(p, t);
        for (int pi = 0; pi <= 2 * sc.nd; ++pi) {
          mbar_wait(&sempty[slot], ph ^ 1u);
          mbar_expect_tx(&sfull[slot], p.slab_tx);
          for (int sub = 0; sub < 4; ++sub)
            tma_load_5d(slabs + (size_t)slot * p.slab_bytes + (size_t)sub * p.chunk_bytes, &tmA, &sfull[slot], 0, 2 * sc.w0 - (sub & 1),
                        2 * sc.h0 - (sub >> 1), 2 * sc.d0 - 1 + pi, sc.b);
          if (++slot == (uint32_t)p.nslab) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (warp-uniform loop, one elected lane issues) ============
    const bool leader = elect_one();
    constexpr uint32_t ROWB = KC * 2u;
    constexpr uint32_t LAYOUT = ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u);
    constexpr uint32_t A_HI = ((HT_W * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
    constexpr uint32_t B_HI = ((8u * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
    constexpr uint32_t W_TILE16 = (NT * ROWB) >> 4;
    constexpr uint32_t IDESC0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t w_lo = ((smem_u32(wreg) & 0x3FFFFu) >> 4) | 0x10000u;
    const uint32_t s_lo = ((smem_u32(slabs) & 0x3FFFFu) >> 4) | 0x10000u;
    const uint32_t slab16 = p.slab_bytes >> 4, sub16 = p.chunk_bytes >> 4;
    const uint32_t nslab = (uint32_t)p.nslab;
    mbar_wait(wfull, 0);
    uint32_t s = 0;            // ring block of output plane d0 + pi/2 (the kd = 0 target of an odd input plane)
    uint32_t pbits = 0;        // per-block phase of tempty
    uint32_t sslot = 0, sph = 0;
    auto mma = [&](uint32_t col, uint32_t nblk, uint32_t a_lo, uint32_t b_lo, uint32_t accum) {
      const uint32_t idesc = IDESC0 | (((nblk * NT) >> 3) << 17);
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
          "setp.ne.b32 p, %6, 0;\n\t"
          "mov.b64 da, {%1, %2};\n\t"
          "mov.b64 db, {%3, %4};\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
          ::"r"(tmem_base + col), "r"(a_lo), "r"(A_HI), "r"(b_lo), "r"(B_HI), "r"(idesc), "r"(accum)
          : "memory");
    };
    for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
      const SegCoord sc = decode_seg(p, t);
      for (int pi = 0; pi <= 2 * sc.nd; ++pi) {
        mbar_wait(&sfull[sslot], sph);
        const bool odd_plane = (pi & 1) == 0;               // input plane 2 d0 - 1 + pi
        const bool has0 = odd_plane && pi < 2 * sc.nd;      // kd = 0 -> output pi/2 (block s) starts here
        const bool has2 = odd_plane && pi > 0;              // kd = 2 -> output pi/2 - 1 (block s + 1)
        if (has0) {
          mbar_wait(&tempty[s], ((pbits >> s) & 1u) ^ 1u);
          pbits ^= 1u << s;
        }
        tc_fence_after();
        const uint32_t a_pl = s_lo + sslot * slab16;
        const uint32_t col0 = s * NT, col2 = ((s + 1u) & (kRing - 1)) * NT;
        const bool fused = has0 && has2 && s != (uint32_t)(kRing - 1);
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) {
          const int kh = t9 / 3, kw = t9 % 3;
          const int sub = (kh != 1 ? 2 : 0) + (kw != 1 ? 1 : 0);
          const uint32_t a_t9 = a_pl + (uint32_t)sub * sub16 + (uint32_t)((((kh == 2 ? HT_W : 0) + (kw == 2 ? 1 : 0)) * ROWB) >> 4);
          const uint32_t b_t9 = w_lo + (uint32_t)(t9 * 3) * W_TILE16;
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk) {
            const uint32_t a_lo = a_t9 + (uint32_t)(kk * 2), b_lo = b_t9 + (uint32_t)(kk * 2);
            if (leader) {
              if (!odd_plane) {
                mma(col2, 1u, a_lo, b_lo + 2u * W_TILE16, 1u);            // kd = 1 into output (pi-1)/2 = block s + 1
              } else if (t9 == 0 && kk == 0) {
                if (has0) mma(col0, 1u, a_lo, b_lo, 0u);                   // first contribution to a new output plane
                if (has2) mma(col2, 1u, a_lo, b_lo + W_TILE16, 1u);
              } else if (fused) {
                mma(col0, 2u, a_lo, b_lo, 1u);
              } else {
                if (has0) mma(col0, 1u, a_lo, b_lo, 1u);
                if (has2) mma(col2, 1u, a_lo, b_lo + W_TILE16, 1u);
              }
            }
          }
        }
        if (leader) {
          if (has2) tc_commit(&tfull[(s + 1u) & (kRing - 1)]);   // output pi/2 - 1 has all 27 taps
          tc_commit(&sempty[sslot]);
        }
        __syncwarp();
        if (++sslot == nslab) { sslot = 0; sph ^= 1u; }
        if (has0) s = (s + kRing - 1) & (kRing - 1);   // the next odd plane starts the next output plane (also across segments)
      }
    }
  } else {
    ring_epilogue<NT, EPI, 0>(p, tmem_base, tfull, tempty, sstat, n0, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}


// ================================================================================================
// Tap-packed kernel for the few-channel first layers (head conv on the 1-channel MRI, the modulator stacks on the 3-channel
// painted prompt and the 2-channel [prompt | backbone] pair; attn_unet_data_parallel.py:285-286,553-556,651).
// Padding 1-3 channels to 16 makes every tap a K = 16 MMA that is 6-19 % real work and forces a 16-channel copy of the
// input through HBM.  Here K runs over (tap, channel): k = tap * Cin + ci, K = 27 * Cin padded to a multiple of 32, so a
// 128-voxel tile costs 2-8 MMAs instead of 10-12 and the input is read in its own few-channel layout.
// Input planes arrive by TMA as un-swizzled 18 x (24 Cin) slabs (the (w, c) axes merged so a row is >= 48 bytes; zero fill
// outside the volume = padding) into a ring that keeps three planes live.  Eight builder warps (two threads per tile row:
// taps 0..15 / 16..26) gather each voxel's neighbourhood from the slabs with constant-offset shared-memory loads and write
// the A tile [128 rows][32 k] per K chunk in the 64B-swizzled K-major layout of the UMMA descriptor, fence the generic-proxy
// writes and hand the stage to the MMA warp.  The B tiles are built once per CTA from the standard packed weights
// w[tap][Cout][Cin].  Accumulator ring, segment walk and epilogue are the plane-ring ones (ring_epilogue).
// Measured (8 x 128^3): 2->16 0.60 ms, 4->16 0.67 ms against 0.38 ms for the zero-padded 16->16 plane-ring kernel plus ~0.15 ms
// of extra pack traffic: the ~1200 builder warp-instructions per tile make the SM issue-bound, so this kernel serves callers
// that hold few-channel tensors (no 16-channel copy in HBM) while the model keeps padding its small inputs (model.slim_inputs).
// ================================================================================================
constexpr int kTapStages = 4, kTapBuilderWarps = 8, kTapSlabs = 8, kTapEpi = 1;
constexpr int kTapThreads = 64 + 128 * kTapEpi + 32 * kTapBuilderWarps;

template <int NT, int CIN>
__global__ void __launch_bounds__(kTapThreads, 1)
conv_taps_kernel(const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ w, const HaloParams p) {
  constexpr int K = 27 * CIN, CH = (K + 31) / 32;                        // K chunks of 32 elements (64-byte rows)
  constexpr uint32_t A_CHUNK = 128u * 64u, A_STAGE = A_CHUNK * CH, B_CHUNK = (uint32_t)NT * 64u;
  // slab row: 24 voxels x CIN bf16 starting at w0 - 8 (TMA needs a 16-byte aligned start in the innermost dimension, so the
  // one-voxel halo on the low side costs seven unused voxels), un-swizzled
  constexpr uint32_t SROW = 48u * CIN, SLAB = ((uint32_t)HALO_H * SROW + 127u) & ~127u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wreg = smem;                                                   // [CH][NT rows][64 B], swizzled
  uint8_t* astg = smem + ((B_CHUNK * CH + 1023u) & ~1023u);               // [kTapStages][CH][128 rows][64 B], swizzled
  uint8_t* slabs = astg + (size_t)kTapStages * A_STAGE;                   // [kTapSlabs][18 rows][SROW]
  uint64_t* afull = reinterpret_cast<uint64_t*>(slabs + (size_t)kTapSlabs * SLAB);
  uint64_t* aempty = afull + kTapStages;
  uint64_t* sfull = aempty + kTapStages;
  uint64_t* sempty = sfull + kTapSlabs;
  uint64_t* tfull = sempty + kTapSlabs;
  uint64_t* tempty = tfull + kRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kRing);
  float* sstat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTapStages; ++s) { mbar_init(&afull[s], kTapBuilderWarps); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < kTapSlabs; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], kTapBuilderWarps); }
    for (int a = 0; a < kRing; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B tiles: element (co, k) of chunk c sits at row co, byte (k%32)*2, 16-byte piece index XOR ((co >> 1) & 3)
  for (int i = threadIdx.x; i < CH * NT * 32; i += blockDim.x) {
    const int k = (i / (NT * 32)) * 32 + (i % 32), co = (i / 32) % NT;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (k < K && co < p.Cout) v = w[((size_t)(k / CIN) * p.Cout + co) * CIN + (k % CIN)];     // standard packed w[tap][Cout][Cin]
    const uint32_t kb = (uint32_t)(k % 32) * 2u;
    const uint32_t off = (uint32_t)(k / 32) * B_CHUNK + (uint32_t)co * 64u + ((((kb >> 4) ^ ((uint32_t)co >> 1)) & 3u) << 4) + (kb & 15u);
    *reinterpret_cast<__nv_bfloat16*>(wreg + off) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer: planes d0-1 .. d0+nd of every segment =================================
    if (lane == 0) {
      uint32_t slot = 0, ph = 0;
      for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
        const SegCoord sc = decode_seg(p, t);
        for (int pi = 0; pi < sc.nd + 2; ++pi) {
          mbar_wait(&sempty[slot], ph ^ 1u);
          mbar_expect_tx(&sfull[slot], (uint32_t)HALO_H * SROW);
#ifndef TAPS_NO_TMA
          tma_load_4d(slabs + (size_t)slot * SLAB, &tmX, &sfull[slot], (sc.w0 - 8) * CIN, sc.h0 - 1, sc.d0 - 1 + pi, sc.b);
#else
          asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&sfull[slot])), "r"((uint32_t)HALO_H * SROW) : "memory");
#endif
          if (++slot == (uint32_t)kTapSlabs) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =================================
    const bool leader = elect_one();
    constexpr uint32_t D_HI = ((8u * 64u) >> 4) | (1u << 14) | (4u << 29);          // SBO = 8 rows of 64 B, version 1, 64B swizzle
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t w_lo = ((smem_u32(wreg) & 0x3FFFFu) >> 4) | 0x10000u;
    const uint32_t a_lo0 = ((smem_u32(astg) & 0x3FFFFu) >> 4) | 0x10000u;
    uint32_t stage = 0, sph = 0, s_seg = 0, pbits = 0;
    for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
      const SegCoord sc = decode_seg(p, t);
      for (int i = 0; i < sc.nd; ++i) {
        const uint32_t blk = (s_seg + (uint32_t)(kRing * 64 - i)) & (kRing - 1);
        mbar_wait(&afull[stage], sph);
        mbar_wait(&tempty[blk], ((pbits >> blk) & 1u) ^ 1u);
        pbits ^= 1u << blk;
        tc_fence_after();
        const uint32_t a_st = a_lo0 + stage * (A_STAGE >> 4);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t a_lo = a_st + (uint32_t)c * (A_CHUNK >> 4) + (uint32_t)kk * 2u;
            const uint32_t b_lo = w_lo + (uint32_t)c * (B_CHUNK >> 4) + (uint32_t)kk * 2u;
            const uint32_t accum = (c | kk) != 0;
            if (leader)
              asm volatile(
                  "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                  "setp.ne.b32 p, %6, 0;\n\t"
                  "mov.b64 da, {%1, %2};\n\t"
                  "mov.b64 db, {%3, %4};\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                  ::"r"(tmem_base + blk * NT), "r"(a_lo), "r"(D_HI), "r"(b_lo), "r"(D_HI), "r"(idesc), "r"(accum)
                  : "memory");
          }
        }
        if (leader) {
          tc_commit(&tfull[blk]);
          tc_commit(&aempty[stage]);
        }
        __syncwarp();
        if (++stage == (uint32_t)kTapStages) { stage = 0; sph ^= 1u; }
      }
      s_seg = (s_seg + (uint32_t)(kRing * 64 - sc.nd)) & (kRing - 1);
    }
  } else if (warp < 2 + 4 * kTapEpi) {
    ring_epilogue<NT, kTapEpi, 0>(p, tmem_base, tfull, tempty, sstat, 0, warp, lane);
  } else {
    // ================================ A-tile builders: 256 threads, two per tile row ==============
    const int bt = (int)threadIdx.x - (64 + 128 * kTapEpi);
    const int row = bt & 127, half = bt >> 7;                                // half 0: taps 0..15, half 1: taps 16..26 (+ zero padding)
    const int lw = row % HW_T, lh = row / HW_T;
    constexpr int WPT = CIN >= 2 ? CIN / 2 : 1;                              // 32-bit words per tap (CIN = 1: two taps share a word)
    constexpr int NW = CIN == 1 ? 8 : 16 * WPT;                              // words this thread writes (taps 0..15 of its half)
    const uint32_t voff = (uint32_t)lh * SROW + (uint32_t)lw * CIN * 2u;     // this row's voxel inside a slab (tap kh = kw = 0)
    const uint32_t sw = ((uint32_t)row >> 1) & 3u;
    uint32_t slot0 = 0, ph0 = 0;                                             // slab of input plane d0 - 1 + i (oldest of the three)
    uint32_t stage = 0, aph = 0;
    for (int t = blockIdx.x; t < p.total_segs; t += gridDim.x) {
      const SegCoord sc = decode_seg(p, t);
      for (int i = 0; i < sc.nd; ++i) {
        uint32_t sl[3], sp[3];
        sl[0] = slot0; sp[0] = ph0;
#pragma unroll
        for (int q = 1; q < 3; ++q) { sl[q] = sl[q - 1] + 1; sp[q] = sp[q - 1]; if (sl[q] == (uint32_t)kTapSlabs) { sl[q] = 0; sp[q] ^= 1u; } }
#pragma unroll
        for (int q = 0; q < 3; ++q) mbar_wait(&sfull[sl[q]], sp[q]);
        uint32_t wd[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) wd[q] = 0u;
        auto gather = [&](auto hsel) {            // compile-time half: every tap offset and word index is an immediate
          constexpr int T0 = decltype(hsel)::value * 16, T1 = decltype(hsel)::value ? 27 : 16;
#pragma unroll
          for (int tap = T0; tap < T1; ++tap) {
            const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3, lt = tap - T0;
            const uint8_t* src = slabs + (size_t)sl[kd] * SLAB + voff + (uint32_t)kh * SROW + (uint32_t)(kw + 7) * CIN * 2u;
            if (CIN == 1) {
              const uint32_t v = *reinterpret_cast<const uint16_t*>(src);
              wd[lt >> 1] |= v << (16 * (lt & 1));
            } else if (CIN == 2) {
              wd[lt] = *reinterpret_cast<const uint32_t*>(src);
            } else {
              const uint2 v = *reinterpret_cast<const uint2*>(src);
              wd[2 * lt] = v.x;
              wd[2 * lt + 1] = v.y;
            }
          }
        };
        if (half == 0) gather(std::integral_constant<int, 0>{});
        else gather(std::integral_constant<int, 1>{});
        mbar_wait(&aempty[stage], aph ^ 1u);
        uint8_t* ast = astg + (size_t)stage * A_STAGE + (size_t)row * 64u;
        if (CIN == 1) {          // one chunk: half 0 writes 16-byte pieces 0,1 (taps 0..15), half 1 pieces 2,3
#pragma unroll
          for (int j = 0; j < 2; ++j)
            *reinterpret_cast<uint4*>(ast + ((((uint32_t)(2 * half + j)) ^ sw) << 4)) = make_uint4(wd[4 * j], wd[4 * j + 1], wd[4 * j + 2], wd[4 * j + 3]);
        } else {                 // CIN/2 chunks per half (16 taps x CIN elements = CIN/2 x 32)
#pragma unroll
          for (int c = 0; c < CIN / 2; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(ast + (size_t)(half * (CIN / 2) + c) * A_CHUNK + (((uint32_t)j ^ sw) << 4)) =
                  make_uint4(wd[c * 16 + j * 4], wd[c * 16 + j * 4 + 1], wd[c * 16 + j * 4 + 2], wd[c * 16 + j * 4 + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&afull[stage]);
          // the slab reads fed the stores above, so they are complete: the oldest plane is not needed after this tile,
          // the last tile of a segment also releases the other two
          mbar_arrive(&sempty[sl[0]]);
          if (i == sc.nd - 1) { mbar_arrive(&sempty[sl[1]]); mbar_arrive(&sempty[sl[2]]); }
        }
        if (++stage == (uint32_t)kTapStages) { stage = 0; aph ^= 1u; }
        if (++slot0 == (uint32_t)kTapSlabs) { slot0 = 0; ph0 ^= 1u; }
      }
      // the segment consumed nd + 2 slabs
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (++slot0 == (uint32_t)kTapSlabs) { slot0 = 0; ph0 ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ================================================================================================
// CTA-pair plane kernel (cta_group::2) for k3 s1 convolutions with 64 / 128 input channels.
//
// With Cin = 128 the 27 weight tiles of even 32 output channels (221 KB) do not fit next to two slabs, so the single-CTA
// kernel splits Cout four ways and issues N = 3*16 = 48 MMAs that are bound by the A-operand read (44 cycles for 24 cycles
// of math).  Two SMs of a TPC can share ONE weight set: each CTA keeps its own 128-voxel tile (own slabs, own TMEM
// accumulators, own epilogue) and HALF of the rows of every weight tile; the leader CTA issues tcgen05.mma.cta_group::2
// with M = 256, N = 3*32 = 96 (measured 49 cycles, profiles/r01_probe_cta_pair_mma.log: 1.8x the work per SM-cycle).
//
// Because the two CTAs hold fixed halves of each 96-row weight tile, every MMA must be the full N = 96: there are no
// partial-N steps for ring wrap-around, segment edges or "first tap" initialisation.  Instead
//   * TMEM is a LINEAR array of 16 blocks per segment (<= 12 output planes): input plane pi always adds into blocks
//     [pi, pi+1, pi+2] through W[kd = 2, 1, 0]; blocks 0, 1 and nd+2, nd+3 only ever receive out-of-segment garbage;
//   * every MMA accumulates; the epilogue warps clear a block (tcgen05.st) after draining it, and all blocks once at start.
// Cross-CTA plumbing: both producers signal the LEADER's slab barrier (cp.async.bulk.tensor.cta_group::2 + remote
// expect_tx), both epilogues arrive on the leader's accumulator-free barriers through shared::cluster addresses, and the
// leader's tcgen05.commit multicasts "accumulator ready" / "slab free" to both CTAs.
// ================================================================================================
constexpr int kPairBlocks = 16, kPairMaxND = kPairBlocks - 4;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(const void* local, uint32_t rank) {        // shared::cluster address of `local` in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: what must be ordered before this arrive are tcgen05 accesses (tcgen05.fence::before_thread_sync), not generic
  // memory -- the default .release.cluster costs a GPU-scope MEMBAR per call (40 % of all stall samples in the epilogue)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {      // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_st16_zero(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
               ::"r"(taddr), "r"(0u) : "memory");
}

// EPI epilogue warpgroups per CTA, CPS CTAs per SM (few channels: more issuing warps and epilogue warps per SM)
template <int NT, int KC, int KCH, int EPI, int CPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * EPI, CPS)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  constexpr uint32_t ROWB = KC * 2u, HALF = 3u * NT / 2u;                  // weight rows per CTA and (tap, K chunk): half of [kd2|kd1|kd0] x NT
  constexpr uint32_t LAYOUT = ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wreg = smem;                                                    // [9 (kh,kw)][KCH][HALF rows][ROWB]
  uint8_t* slabs = smem + ((p.w_bytes + 1023u) & ~1023u);
  uint64_t* sfull = reinterpret_cast<uint64_t*>(slabs + (size_t)p.nslab * p.slab_bytes);
  uint64_t* sempty = sfull + kMaxSlabs;
  uint64_t* wfull = sempty + kMaxSlabs;
  uint64_t* tfull = wfull + 1;
  uint64_t* tempty = tfull + kPairBlocks;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kPairBlocks);
  float* sstat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n0 = blockIdx.y * NT;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1, units = p.total_segs >> 1;    // a unit = two w-adjacent segments
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslab; ++s) { mbar_init(&sfull[s], 2); mbar_init(&sempty[s], 1); }      // leader: both CTAs' slabs
    mbar_init(wfull, 1);
    for (int a = 0; a < kPairBlocks; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }   // leader: 4 warps of each CTA
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    // this CTA's half of every weight tile: rows [HALF*rank, HALF*rank + HALF) of [kd=2 | kd=1 | kd=0] x NT, in 8-row boxes
    mbar_expect_tx(wfull, p.w_bytes);
    for (int t9 = 0; t9 < 9; ++t9)
      for (int kc = 0; kc < KCH; ++kc)
        for (uint32_t q = 0; q < HALF / 8u; ++q) {
          const uint32_t j0 = HALF * rank + 8u * q;
          const int kd = 2 - (int)(j0 / NT), co0 = (int)(j0 % NT);
          tma_load_2d(wreg + (size_t)((t9 * KCH + kc) * HALF + 8u * q) * ROWB, &tmB, wfull, kc * KC, (kd * 9 + t9) * p.Cout + n0 + co0);
        }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(kPairBlocks * NT)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  mbar_wait(wfull, 0);
  cluster_sync_all();                       // both CTAs: barriers initialised, weights resident, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs; completion is signalled on the LEADER's barrier) ==========
    if (lane == 0) {
      uint32_t slot = 0, ph = 0;
      for (int u = pair; u < units; u += npairs) {
        const SegCoord sc = decode_seg(p, 2 * u + (int)rank);
        for (int pi = 0; pi < sc.nd + 2; ++pi) {
          mbar_wait(&sempty[slot], ph ^ 1u);
          const uint32_t lead_full = mapa_rank(&sfull[slot], 0);
          mbar_expect_tx_cluster(lead_full, p.slab_tx);
          for (int kc = 0; kc < KCH; ++kc)
            tma_load_5d_pair(slabs + (size_t)slot * p.slab_bytes + (size_t)kc * p.chunk_bytes, &tmA, lead_full, kc * KC, sc.w0 - 1, sc.h0 - 1,
                             sc.d0 - 1 + pi, sc.b);
          if (++slot == (uint32_t)p.nslab) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer: leader CTA only, M = 256 over both SMs ============
    if (rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t A_HI = ((HALO_W * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
      constexpr uint32_t B_HI = ((8u * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
      constexpr uint32_t W_TILE16 = (HALF * ROWB) >> 4;
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (((3u * NT) >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t w_lo = ((smem_u32(wreg) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t s_lo = ((smem_u32(slabs) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t slab16 = p.slab_bytes >> 4, chunk16 = p.chunk_bytes >> 4;
      uint32_t slot = 0, sph = 0, pbits = 0;
      for (int u = pair; u < units; u += npairs) {
        const SegCoord sc = decode_seg(p, 2 * u);
        for (int pi = 0; pi < sc.nd + 2; ++pi) {
          mbar_wait(&sfull[slot], sph);
          // blocks this plane touches for the first time in this segment must have been drained and cleared
          for (int b = (pi == 0 ? 0 : pi + 2); b <= pi + 2; ++b) {
            mbar_wait(&tempty[b], (pbits >> b) & 1u);
            pbits ^= 1u << b;
          }
          tc_fence_after();
          const uint32_t a_pl = s_lo + slot * slab16;
          const uint32_t dcol = tmem_base + (uint32_t)pi * NT;
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
              for (int kc = 0; kc < KCH; ++kc) {
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  const uint32_t a_lo = a_pl + (uint32_t)kc * chunk16 + (uint32_t)((((kh * HALO_W + kw) * ROWB) >> 4) + kk * 2);
                  const uint32_t b_lo = w_lo + (uint32_t)(((kh * 3 + kw) * KCH + kc)) * W_TILE16 + (uint32_t)(kk * 2);
                  if (leader)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                        "setp.ne.b32 p, %6, 0;\n\t"
                        "mov.b64 da, {%1, %2};\n\t"
                        "mov.b64 db, {%3, %4};\n\t"
                        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
                        ::"r"(dcol), "r"(a_lo), "r"(A_HI), "r"(b_lo), "r"(B_HI), "r"(IDESC), "r"(1u)
                        : "memory");
                }
              }
            }
          }
          if (leader) {
            tc_commit_pair(&tfull[pi]);                       // block pi has received its last contribution
            if (pi == sc.nd + 1) { tc_commit_pair(&tfull[pi + 1]); tc_commit_pair(&tfull[pi + 2]); }
            tc_commit_pair(&sempty[slot]);
          }
          __syncwarp();
          if (++slot == (uint32_t)p.nslab) { slot = 0; sph ^= 1u; }
        }
      }
    }
  } else {
    // ================================ epilogue (both CTAs): drain, store, clear ======================
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const int tid128 = ((int)threadIdx.x - 64) & 127;
    const uint32_t bar_id = 1u + (uint32_t)grp;
    const int row = q * 32 + lane;
    const int lw = row % HW_T, lh = row / HW_T;
    const float slope = p.slope ? __ldg(p.slope) : 0.f;
    const float neg = act_neg(p.act, slope);
    const bool clamp0 = p.act == COMA_ACT_LEAKY_RELU, do_stats = p.stats != nullptr;
    float* gstat = sstat + (size_t)grp * 11 * NT;
    float* cA = gstat + 8 * NT;
    float* cS = cA + NT;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t lead_tempty = mapa_rank(&tempty[0], 0);
    auto release = [&](int b) {                // clear block b, then tell the leader's MMA warp it is free
#pragma unroll
      for (int c0 = 0; c0 < NT; c0 += 16) tc_st16_zero(lane_base + (uint32_t)(b * NT + c0));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_tempty + (uint32_t)b * 8u);
    };
    for (int b = grp; b < kPairBlocks; b += EPI) release(b);
    uint32_t bcount = 0, ebits = 0;
    for (int u = pair; u < units; u += npairs) {
      const SegCoord sc = decode_seg(p, 2 * u + (int)rank);
      const int oh = sc.h0 + lh, ow = sc.w0 + lw;
      const bool valid = oh < p.H && ow < p.W;
      epi_stage_coef(p.bias, p.scale, p.shift, (int64_t)sc.b * p.Cout, n0, NT, cA, cS, tid128, bar_id);
      float s1[NT], s2[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
      for (int b = 0; b < sc.nd + 4; ++b, ++bcount) {
        if (EPI > 1 && (bcount % EPI) != (uint32_t)grp) { ebits ^= 1u << b; continue; }
        mbar_wait(&tfull[b], (ebits >> b) & 1u);
        ebits ^= 1u << b;
        tc_fence_after();
        if (b >= 2 && b < sc.nd + 2) {
          __nv_bfloat16* yrow = p.y + ((((int64_t)sc.b * p.D + (sc.d0 + b - 2)) * p.H + oh) * p.W + ow) * p.y_cs + n0;
#pragma unroll
          for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t raw[16];
            tc_ld16(lane_base + (uint32_t)(b * NT + c0), raw);
            epi_chunk16(raw, cA + c0, cS + c0, cS + NT + c0, valid, neg, clamp0, slope, do_stats, s1 + c0, s2 + c0, yrow + c0, n0 + c0, p.y_cn, p.y_cs, p.scale == nullptr);
          }
        }
        release(b);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (p.stats) {
        float* wstat = gstat + (size_t)q * NT * 2;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float a = warp_sum(s1[j]), b2 = warp_sum(s2[j]);
          if (lane == 0) { wstat[j * 2] = a; wstat[j * 2 + 1] = b2; }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        for (int i = tid128; i < NT * 2; i += 128) {
          const float sm = gstat[i] + gstat[NT * 2 + i] + gstat[NT * 4 + i] + gstat[NT * 6 + i];
          p.stats[(((int64_t)sc.b * p.stat_chunks + sc.chunk * EPI + grp) * p.Cout + n0 + (i >> 1)) * 2 + (i & 1)] = sm;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer may still be reading this CTA's weights / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(kPairBlocks * NT)) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

std::mutex g_map_mutex;
std::unordered_map<std::string, CUtensorMap> g_map_cache;

bool make_map(CUtensorMap* out, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box, const cuuint32_t* estr, int swz) {
  std::string key(reinterpret_cast<const char*>(&base), sizeof(base));
  key.append(reinterpret_cast<const char*>(dims), sizeof(cuuint64_t) * rank);
  key.append(reinterpret_cast<const char*>(strides_bytes), sizeof(cuuint64_t) * (rank - 1));
  key.append(reinterpret_cast<const char*>(box), sizeof(cuuint32_t) * rank);
  key.append(reinterpret_cast<const char*>(estr), sizeof(cuuint32_t) * rank);
  key.push_back((char)swz);
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) { *out = it->second; return true; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return false; }
  const CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swz == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_32B));
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return false; }
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache.emplace(key, *out);
  return true;
}

int pick_kc(int cin) { return cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : (cin % 16 == 0 ? 16 : 0)); }   // channels per K chunk (= swizzle span / 2)
int pick_nt(int cout) {
  if (cout <= 256) return cout;
  for (int nt = 256; nt >= 16; nt -= 16)
    if (cout % nt == 0) return nt;
  return 0;
}

// ---- halo-reuse (v2) planning ----
// barriers (sfull, sempty, sready per slab; wfull; tfull, tempty per ring block), the TMEM slot, two epilogue groups of
// [4 warps][NT][2] stat partials + cA, cS, cB, and the prologue coefficients (scale, shift for up to 64 input channels)
static size_t halo_tail_bytes(int nt) { return (3 * kMaxSlabs + 1 + 2 * kRing) * 8 + 16 + (size_t)22 * nt * sizeof(float) + 64 + 2 * 64 * sizeof(float); }

struct HaloPlan { bool ok, s2; int ctas; int cols_w, cols_h, segs_d, DS, nslab, NT, KCH; uint32_t rowb, slab_bytes, chunk_bytes, w_tile_bytes, w_bytes; size_t smem; };

HaloPlan plan_halo(const coma_conv_args& a) {
  HaloPlan h{};
  h.ok = false;
  static const bool disabled = [] { const char* e = getenv("COMA_DISABLE_HALO"); return e && e[0] == '1'; }();
  if (disabled || a.ksize != 3) return h;
  static const bool v3 = [] { const char* e = getenv("COMA_DISABLE_HALO3"); return !(e && e[0] == '1'); }();
  static const bool s2_off = [] { const char* e = getenv("COMA_DISABLE_HALO_S2"); return e && e[0] == '1'; }();
  h.s2 = !a.transposed && a.stride == 2 && a.pad == 1 && v3 && !s2_off;
  if (a.transposed ? a.stride != 2 : (a.stride != 1 && !h.s2)) return h;
  const bool use_v3 = v3 && !a.transposed;
  const bool cin_ok = a.Cin == 16 || a.Cin == 32 || a.Cin == 64 || (use_v3 && !h.s2 && a.Cin == 128);
  if (!cin_ok || a.Cout % 16 != 0 || a.Cout > 256) return h;
  const int gw = a.transposed ? a.Wi : a.Wo, gh = a.transposed ? a.Hi : a.Ho, gd = a.transposed ? a.Di : a.Do;
  if (gw < HW_T || gh < HH_T) return h;                  // tiny planes: the per-tap kernel wastes less
  h.KCH = a.Cin == 128 ? 2 : 1;
  h.rowb = (uint32_t)(a.Cin / h.KCH) * 2u;
  h.chunk_bytes = ((uint32_t)(a.transposed || h.s2 ? HT_ROWS : HALO_ROWS) * h.rowb + 1023u) & ~1023u;
  h.slab_bytes = h.chunk_bytes * (uint32_t)(h.s2 ? 4 : h.KCH);      // stride 2: four (h,w)-parity sub-slabs per input plane
  size_t budget = 222 * 1024;
  h.ctas = 1;
  // v2 / transposed keep three input planes live (+1 in flight); v3 consumes each plane once (+1 in flight)
  const size_t min_slabs = use_v3 ? 2 : 4;
  // N per CTA: all of Cout when the 27 weight tiles fit, otherwise (v3 only) split Cout over grid.y
  h.NT = 0;
  for (int nt : {64, 32, 16}) {
    if (a.Cout % nt != 0) continue;
    if (nt != a.Cout && !use_v3) continue;
    const size_t tail_nt = halo_tail_bytes(nt);
    const size_t fixed_nt = 1024 + ((27u * h.KCH * nt * h.rowb + 1023u) & ~1023u) + tail_nt;
    if (fixed_nt + min_slabs * (size_t)h.slab_bytes <= budget) { h.NT = nt; break; }
  }
  if (h.NT == 0) return h;
  if ((h.KCH > 1 || h.s2) && a.Cout / h.NT > 4) return h;          // too many re-reads of A: the per-tap kernel does better
  {
    // two CTAs per SM (stride-1 v3 kernel): everything of one CTA must fit half an SM (shared memory, 256 TMEM columns)
    static const bool two_off = [] { const char* e = getenv("COMA_DISABLE_HALO_2CTA"); return e && e[0] == '1'; }();
    const size_t half = (h.NT == 16 ? 72 : 110) * 1024;      // NT = 16: three CTAs per SM
    const size_t tail2 = halo_tail_bytes(h.NT);
    const size_t fixed2 = 1024 + ((27u * h.KCH * h.NT * h.rowb + 1023u) & ~1023u) + tail2;
    if (!two_off && use_v3 && !h.s2 && h.KCH == 1 && h.NT <= 32 && kRing * h.NT <= 256 && fixed2 + 4 * (size_t)h.slab_bytes <= half) {
      h.ctas = 2;
      budget = half;
    }
  }
  h.w_tile_bytes = (uint32_t)h.NT * h.rowb;
  h.w_bytes = 27u * (uint32_t)h.KCH * h.w_tile_bytes;
  const size_t tail = halo_tail_bytes(h.NT);
  const size_t fixed = 1024 + ((h.w_bytes + 1023u) & ~1023u) + tail;
  int nslab = (int)((budget - fixed) / h.slab_bytes);
  h.nslab = nslab > kMaxSlabs ? kMaxSlabs : nslab;
  h.smem = fixed + (size_t)h.nslab * h.slab_bytes;
  h.cols_w = (gw + HW_T - 1) / HW_T;
  h.cols_h = (gh + HH_T - 1) / HH_T;
  const int ncols = a.B * h.cols_w * h.cols_h;
  int segs = (4 * num_sms() * h.ctas * (h.ctas == 2 && h.NT == 16 ? 3 : 2) / 2 + ncols - 1) / ncols;
  const int max_segs = (gd + 3) / 4;
  if (segs > max_segs) segs = max_segs;
  if (segs < 1) segs = 1;
  h.DS = (gd + segs - 1) / segs;
  h.segs_d = (gd + h.DS - 1) / h.DS;
  h.ok = true;
  return h;
}

template <int NT, int KC, int KCH = 1>
int launch_halo(const coma_conv_args& a, const HaloPlan& h, const CUtensorMap& tmA, const CUtensorMap& tmB, cudaStream_t stream) {
  HaloParams p{};
  const bool tr = a.transposed != 0;
  p.B = a.B; p.D = tr ? a.Di : a.Do; p.H = tr ? a.Hi : a.Ho; p.W = tr ? a.Wi : a.Wo; p.Cin = a.Cin; p.Cout = a.Cout; p.KC = KC;
  p.cols_w = h.cols_w; p.cols_h = h.cols_h; p.segs_d = h.segs_d; p.DS = h.DS;
  p.total_segs = a.B * h.cols_w * h.cols_h * h.segs_d;
  p.nslab = h.nslab; p.rowb = h.rowb; p.slab_bytes = h.slab_bytes;
  p.slab_tx = h.s2 ? 4u * (uint32_t)S2_ROWS * h.rowb : (uint32_t)(tr ? HT_ROWS : HALO_ROWS) * h.rowb * (uint32_t)KCH;
  p.chunk_bytes = h.chunk_bytes; p.chunk_tx = (uint32_t)(tr ? HT_ROWS : HALO_ROWS) * h.rowb;
  p.w_tile_bytes = h.w_tile_bytes; p.w_bytes = h.w_bytes;
  p.y = static_cast<__nv_bfloat16*>(a.y) + a.y_co; p.y_cs = a.y_cs; p.y_cn = a.y_cn;
  p.bias = a.bias; p.scale = a.scale; p.shift = a.shift; p.slope = a.slope; p.stats = a.stats; p.act = a.act;
  p.stat_chunks = h.cols_w * h.cols_h * h.segs_d;
  p.in_scale = a.in_scale; p.in_shift = a.in_shift; p.in_slope = a.in_slope; p.in_act = a.in_act;
  static const bool v3 = [] { const char* e = getenv("COMA_DISABLE_HALO3"); return !(e && e[0] == '1'); }();
  uint32_t cols = 32;
  const uint32_t need = tr ? ((8u * NT * 2u <= 512u) ? 16u * NT : 8u * NT) : (v3 ? (uint32_t)kRing * NT : 2u * NT);
  while (cols < need) cols <<= 1;
  p.tmem_cols = cols;
  static bool attr_set = false;
  if (!attr_set) {
    if (KCH == 1) {
      cudaFuncSetAttribute(conv_halo_kernel<NT, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(convT_halo_kernel<NT, KC, (NT <= 32 ? 2 : 1)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    cudaFuncSetAttribute(conv_halo3_kernel<NT, KC, KCH, (NT <= 32 ? 2 : 1), 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if constexpr (KCH == 1 && NT <= 32) cudaFuncSetAttribute(conv_halo3_kernel<NT, KC, 1, 1, (NT == 16 ? 3 : 2)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (KCH == 1) cudaFuncSetAttribute(conv_halo_s2_kernel<NT, KC, (NT <= 32 ? 2 : 1)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  const int nsplit = a.Cout / NT;
  int grid = num_sms() / nsplit;
  if (grid < 1) grid = 1;
  if (grid > p.total_segs) grid = p.total_segs;
  if (KCH == 1 && h.s2) {
    constexpr int EPI = NT <= 32 ? 2 : 1;
    p.stat_chunks *= EPI;
    conv_halo_s2_kernel<NT, KC, EPI><<<dim3((unsigned)grid, (unsigned)nsplit), 64 + 128 * EPI, h.smem, stream>>>(tmA, tmB, p);
  }
  else if (KCH == 1 && tr) {
    constexpr int EPI = NT <= 32 ? 2 : 1;
    p.stat_chunks *= EPI;
    convT_halo_kernel<NT, KC, EPI><<<grid, 64 + 128 * EPI, h.smem, stream>>>(tmA, tmB, p);
  }
  else if (KCH == 1 && NT <= 32 && h.ctas == 2) {
   if constexpr (KCH == 1 && NT <= 32) {
    constexpr int CPS = NT == 16 ? 3 : 2;
    int grid2 = CPS * num_sms() / nsplit;
    if (grid2 < 1) grid2 = 1;
    if (grid2 > p.total_segs) grid2 = p.total_segs;
    conv_halo3_kernel<NT, KC, 1, 1, CPS><<<dim3((unsigned)grid2, (unsigned)nsplit), 64 + 128, h.smem, stream>>>(tmA, tmB, p);
   }
  }
  else if (v3 || KCH > 1) {
    constexpr int EPI = NT <= 32 ? 2 : 1;
    p.stat_chunks *= EPI;
    conv_halo3_kernel<NT, KC, KCH, EPI, 1><<<dim3((unsigned)grid, (unsigned)nsplit), 64 + 128 * EPI, h.smem, stream>>>(tmA, tmB, p);
  }
  else conv_halo_kernel<NT, KC><<<grid, kTcThreads, h.smem, stream>>>(tmA, tmB, p);
  COMA_CHECK_LAUNCH("conv_halo");
  return COMA_OK;
}

}  // namespace

// cached cuTensorMapEncodeTiled for bf16 tensors, for the other translation units (wgrad_tc.cu)
bool tensor_map_bf16(CUtensorMap* out, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, const cuuint32_t* estr, int swz) {
  return make_map(out, base, rank, dims, strides_bytes, box, estr, swz);
}

// ---- CTA-pair kernel planning ----
struct PairPlan { bool ok; int cols_w, cols_h, segs_d, DS, nslab, KCH, KC, NT, EPI, CPS; uint32_t rowb, slab_bytes, chunk_bytes, w_bytes; size_t smem; };

static PairPlan plan_pair(const coma_conv_args& a) {
  PairPlan t{};
  static const bool off = [] { const char* e = getenv("COMA_DISABLE_PAIR"); return e && e[0] == '1'; }();
  static const bool all64 = [] { const char* e = getenv("COMA_PAIR_CIN64"); return !(e && e[0] == '0'); }();   // Cin = 64 too (1.1-1.3x)
  static const bool small = [] { const char* e = getenv("COMA_PAIR_SMALL"); return e && e[0] == '1'; }();       // 16 -> 16 (experimental)
  if (off || a.transposed || a.ksize != 3 || a.stride != 1 || a.dtype != COMA_BF16 || a.w_bstride != 0 || a.bias_bstride != 0 || a.in_scale) return t;
  if (a.act == COMA_ACT_SIGMOID) return t;
  if (a.Cin == 128 || (all64 && a.Cin == 64)) {
    if (a.Cout % 32 != 0 || a.Cout / 32 > 4) return t;
    t.KC = 64; t.NT = 32; t.EPI = 2; t.CPS = 1;
  } else if (small && a.Cin == 16 && a.Cout == 16) {
    t.KC = 16; t.NT = 16; t.EPI = 1; t.CPS = 2;
  } else {
    return t;
  }
  if (a.x_cs % 8 != 0 || a.x_co % 8 != 0 || (reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.w) & 15)) return t;
  if (a.Wo < 2 * HW_T || a.Ho < HH_T) return t;
  t.cols_w = (a.Wo + HW_T - 1) / HW_T;
  if (t.cols_w % 2 != 0) return t;                       // the two CTAs of a pair take w-adjacent columns
  t.cols_h = (a.Ho + HH_T - 1) / HH_T;
  t.KCH = a.Cin / t.KC;
  t.rowb = (uint32_t)t.KC * 2u;
  t.chunk_bytes = ((uint32_t)HALO_ROWS * t.rowb + 1023u) & ~1023u;
  t.slab_bytes = t.chunk_bytes * (uint32_t)t.KCH;
  t.w_bytes = 9u * (uint32_t)t.KCH * (3u * (uint32_t)t.NT / 2u) * t.rowb;       // HALF = 3 * NT / 2 rows per (tap, chunk)
  const size_t tail = (2 * kMaxSlabs + 1 + 2 * kPairBlocks) * 8 + 16 + (size_t)22 * t.NT * sizeof(float) + 64;
  const size_t fixed = 1024 + ((t.w_bytes + 1023u) & ~1023u) + tail;
  const size_t budget = (t.CPS == 1 ? 222 : 110) * 1024;
  if (fixed + 2 * (size_t)t.slab_bytes > budget) return t;
  int nslab = (int)((budget - fixed) / t.slab_bytes);
  t.nslab = nslab > kMaxSlabs ? kMaxSlabs : nslab;
  t.smem = fixed + (size_t)t.nslab * t.slab_bytes;
  int segs = (a.Do + kPairMaxND - 1) / kPairMaxND;
  // enough units (pairs of segments) to give every SM pair a few
  const int ncols = a.B * t.cols_w * t.cols_h / 2;
  const int want = (2 * num_sms() * t.CPS + ncols - 1) / ncols;
  if (segs < want) segs = want;
  if (segs > a.Do) segs = a.Do;
  t.DS = (a.Do + segs - 1) / segs;
  t.segs_d = (a.Do + t.DS - 1) / t.DS;
  t.ok = true;
  return t;
}

template <int NT, int KC, int KCH, int EPI, int CPS>
static int launch_pair(const coma_conv_args& a, const PairPlan& t, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)a.Cin, (cuuint64_t)a.Wi, (cuuint64_t)a.Hi, (cuuint64_t)a.Di, (cuuint64_t)a.B};
    cuuint64_t strides[4] = {(cuuint64_t)a.x_cs * 2, (cuuint64_t)a.Wi * a.x_cs * 2, (cuuint64_t)a.Hi * a.Wi * a.x_cs * 2,
                             (cuuint64_t)a.Di * a.Hi * a.Wi * a.x_cs * 2};
    cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)HALO_W, (cuuint32_t)HALO_H, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = const_cast<void*>(static_cast<const void*>(static_cast<const __nv_bfloat16*>(a.x) + a.x_co));
    if (!make_map(&tmA, base, 5, dims, strides, box, estr, (int)t.rowb)) return COMA_ERR_CUDA;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cin, (cuuint64_t)27 * a.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, 8};
    cuuint32_t estr[2] = {1, 1};
    if (!make_map(&tmB, const_cast<void*>(a.w), 2, dims, strides, box, estr, (int)t.rowb)) return COMA_ERR_CUDA;
  }
  HaloParams p{};
  p.B = a.B; p.D = a.Do; p.H = a.Ho; p.W = a.Wo; p.Cin = a.Cin; p.Cout = a.Cout; p.KC = KC;
  p.cols_w = t.cols_w; p.cols_h = t.cols_h; p.segs_d = t.segs_d; p.DS = t.DS;
  p.total_segs = a.B * t.cols_w * t.cols_h * t.segs_d;
  p.nslab = t.nslab; p.rowb = t.rowb; p.slab_bytes = t.slab_bytes; p.slab_tx = (uint32_t)HALO_ROWS * t.rowb * (uint32_t)KCH;
  p.chunk_bytes = t.chunk_bytes; p.chunk_tx = (uint32_t)HALO_ROWS * t.rowb;
  p.w_tile_bytes = (3u * NT / 2u) * t.rowb; p.w_bytes = t.w_bytes;
  p.y = static_cast<__nv_bfloat16*>(a.y) + a.y_co; p.y_cs = a.y_cs; p.y_cn = a.y_cn;
  p.bias = a.bias; p.scale = a.scale; p.shift = a.shift; p.slope = a.slope; p.stats = a.stats; p.act = a.act;
  p.stat_chunks = t.cols_w * t.cols_h * t.segs_d * EPI;
  p.tmem_cols = kPairBlocks * NT;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(conv_pair_kernel<NT, KC, KCH, EPI, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (CPS == 1 ? 227 : 112) * 1024); attr_set = true; }
  const int nsplit = a.Cout / NT;
  int npairs = num_sms() * CPS / 2 / nsplit;
  if (npairs < 1) npairs = 1;
  if (npairs > p.total_segs / 2) npairs = p.total_segs / 2;
  conv_pair_kernel<NT, KC, KCH, EPI, CPS><<<dim3((unsigned)(2 * npairs), (unsigned)nsplit), 64 + 128 * EPI, t.smem, stream>>>(tmA, tmB, p);
  COMA_CHECK_LAUNCH("conv_pair");
  return COMA_OK;
}

static int conv_pair_launch(const coma_conv_args& a, const PairPlan& t, cudaStream_t stream) {
  if (t.KC == 64 && t.KCH == 2) return launch_pair<32, 64, 2, 2, 1>(a, t, stream);
  if (t.KC == 64 && t.KCH == 1) return launch_pair<32, 64, 1, 2, 1>(a, t, stream);
  if (t.KC == 16) return launch_pair<16, 16, 1, 1, 2>(a, t, stream);
  set_error("conv_pair: unsupported channel combination");
  return COMA_ERR_UNSUPPORTED;
}

// ---- tap-packed kernel (few input channels) ----
struct TapsPlan { bool ok; int cols_w, cols_h, segs_d, DS; size_t smem; };

static TapsPlan plan_taps(const coma_conv_args& a) {
  TapsPlan t{};
  static const bool off = [] { const char* e = getenv("COMA_DISABLE_TAPS"); return e && e[0] == '1'; }();
  if (off || a.transposed || a.ksize != 3 || a.stride != 1 || a.dtype != COMA_BF16 || a.w_bstride != 0 || a.bias_bstride != 0) return t;
  if (!(a.Cin == 1 || a.Cin == 2 || a.Cin == 4) || a.x_cs != a.Cin || a.x_co != 0 || a.in_scale) return t;
  if (!(a.Cout == 16 || a.Cout == 32) || a.act == COMA_ACT_SIGMOID) return t;
  if (a.Wo < HW_T || a.Ho < HH_T) return t;
  if (reinterpret_cast<uintptr_t>(a.x) % 16 != 0 || (a.Wi * a.Cin) % 8 != 0) return t;      // TMA: 16-byte base and row pitch
  const int ch = (27 * a.Cin + 31) / 32;
  const size_t slab = ((size_t)HALO_H * 48 * a.Cin + 127) & ~(size_t)127;
  t.smem = 1024 + (((size_t)a.Cout * 64 * ch + 1023) & ~(size_t)1023) + (size_t)kTapStages * 128 * 64 * ch + kTapSlabs * slab +
           (2 * kTapStages + 2 * kTapSlabs + 2 * kRing) * 8 + 16 + (size_t)22 * a.Cout * sizeof(float) + 64;
  if (t.smem > 222 * 1024) return t;
  t.cols_w = (a.Wo + HW_T - 1) / HW_T;
  t.cols_h = (a.Ho + HH_T - 1) / HH_T;
  const int ncols = a.B * t.cols_w * t.cols_h;
  int segs = (4 * num_sms() + ncols - 1) / ncols;
  const int max_segs = (a.Do + 3) / 4;
  if (segs > max_segs) segs = max_segs;
  if (segs < 1) segs = 1;
  t.DS = (a.Do + segs - 1) / segs;
  t.segs_d = (a.Do + t.DS - 1) / t.DS;
  t.ok = true;
  return t;
}

template <int NT, int CIN>
static int launch_taps(const coma_conv_args& a, const TapsPlan& t, cudaStream_t stream) {
  HaloParams p{};
  p.B = a.B; p.D = a.Do; p.H = a.Ho; p.W = a.Wo; p.Cin = a.Cin; p.Cout = a.Cout;
  p.cols_w = t.cols_w; p.cols_h = t.cols_h; p.segs_d = t.segs_d; p.DS = t.DS;
  p.total_segs = a.B * t.cols_w * t.cols_h * t.segs_d;
  p.y = static_cast<__nv_bfloat16*>(a.y) + a.y_co; p.y_cs = a.y_cs; p.y_cn = a.y_cn;
  p.bias = a.bias; p.scale = a.scale; p.shift = a.shift; p.slope = a.slope; p.stats = a.stats; p.act = a.act;
  p.stat_chunks = t.cols_w * t.cols_h * t.segs_d * kTapEpi;
  uint32_t cols = 32;
  while (cols < (uint32_t)kRing * NT) cols <<= 1;
  p.tmem_cols = cols;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(conv_taps_kernel<NT, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_set = true; }
  int grid = num_sms();
  if (grid > p.total_segs) grid = p.total_segs;
  CUtensorMap tmX;
  {   // (w, c) merged so that a slab row is >= 32 bytes: [B, D, H, W*Cin], box 16 voxels x 18 rows of one plane
    cuuint64_t dims[4] = {(cuuint64_t)a.Wi * CIN, (cuuint64_t)a.Hi, (cuuint64_t)a.Di, (cuuint64_t)a.B};
    cuuint64_t strides[3] = {(cuuint64_t)a.Wi * CIN * 2, (cuuint64_t)a.Hi * a.Wi * CIN * 2, (cuuint64_t)a.Di * a.Hi * a.Wi * CIN * 2};
    cuuint32_t box[4] = {24u * CIN, (cuuint32_t)HALO_H, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!make_map(&tmX, const_cast<void*>(a.x), 4, dims, strides, box, estr, 0)) return COMA_ERR_CUDA;
  }
  conv_taps_kernel<NT, CIN><<<grid, kTapThreads, t.smem, stream>>>(tmX, static_cast<const __nv_bfloat16*>(a.w), p);
  COMA_CHECK_LAUNCH("conv_taps");
  return COMA_OK;
}

static int conv_taps_launch(const coma_conv_args& a, const TapsPlan& t, cudaStream_t stream) {
#define COMA_TAPS_CASE(NTV, CV) if (a.Cout == NTV && a.Cin == CV) return launch_taps<NTV, CV>(a, t, stream);
  COMA_TAPS_CASE(16, 1) COMA_TAPS_CASE(16, 2) COMA_TAPS_CASE(16, 4)
  COMA_TAPS_CASE(32, 1) COMA_TAPS_CASE(32, 2) COMA_TAPS_CASE(32, 4)
#undef COMA_TAPS_CASE
  set_error("conv_taps: unsupported channel combination");
  return COMA_ERR_UNSUPPORTED;
}

// the input prologue lives in the stride-1 plane-ring kernel (one K chunk): everything else declines it
bool conv_tc_prologue_supported(const coma_conv_args& a) {
  if (!a.in_scale) return true;
  if (a.transposed || a.stride != 1 || a.ksize != 3 || a.dtype != COMA_BF16) return false;
  static const bool v3 = [] { const char* e = getenv("COMA_DISABLE_HALO3"); return !(e && e[0] == '1'); }();
  coma_conv_args b = a;
  b.in_scale = b.in_shift = nullptr;
  const HaloPlan h = plan_halo(b);
  return v3 && h.ok && !h.s2 && h.KCH == 1 && h.ctas == 2 && a.Cin <= 32;
}

bool conv_tc_supported(const coma_conv_args& a) {
  if (a.in_scale && !conv_tc_prologue_supported(a)) return false;
  if (plan_taps(a).ok) return true;
  if ((a.dtype != COMA_BF16 && a.dtype != COMA_BF16_F32OUT) || a.w_bstride != 0 || a.bias_bstride != 0) return false;
  if (a.dtype == COMA_BF16_F32OUT && a.in_scale) return false;
  if (a.act == COMA_ACT_SIGMOID) return false;   // epilogue: relu-family activations only
  if (pick_kc(a.Cin) == 0 || a.Cout % 16 != 0 || pick_nt(a.Cout) == 0) return false;
  if (a.x_cs % 8 != 0 || a.x_co % 8 != 0 || (reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.w) & 15)) return false;
  if (a.transposed) return a.ksize == 3 && a.stride == 2;
  return (a.ksize == 1 || a.ksize == 3) && (a.stride == 1 || a.stride == 2);
}

static void tile_counts(const coma_conv_args& a, int& tw, int& th, int& td, int& classes) {
  const int gw = a.transposed ? a.Wi : a.Wo, gh = a.transposed ? a.Hi : a.Ho, gd = a.transposed ? a.Di : a.Do;
  tw = (gw + TW - 1) / TW; th = (gh + TH - 1) / TH; td = (gd + TD - 1) / TD;
  classes = a.transposed ? 8 : 1;
}

int conv_tc_stat_chunks(const coma_conv_args& a) {
  {
    const PairPlan pp = plan_pair(a);
    if (pp.ok) return pp.cols_w * pp.cols_h * pp.segs_d * pp.EPI;
  }
  {
    const TapsPlan t = plan_taps(a);
    if (t.ok) return t.cols_w * t.cols_h * t.segs_d * kTapEpi;
  }
  HaloPlan h = plan_halo(a);
  if (a.dtype != COMA_BF16) h.ok = false;          // fp32-out mode: per-tap kernel only
  if (h.ok) {
    static const bool v3 = [] { const char* e = getenv("COMA_DISABLE_HALO3"); return !(e && e[0] == '1'); }();
    const bool dual = (a.transposed || v3 || h.KCH > 1) && h.NT <= 32 && (h.ctas == 1 || a.transposed || h.s2);   // two epilogue warpgroups -> two partials per segment
    return h.cols_w * h.cols_h * h.segs_d * (dual ? 2 : 1);
  }
  int tw, th, td, cl;
  tile_counts(a, tw, th, td, cl);
  return tw * th * td * cl;
}

static int conv_halo_launch(const coma_conv_args& a, const HaloPlan& h, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)a.Cin, (cuuint64_t)a.Wi, (cuuint64_t)a.Hi, (cuuint64_t)a.Di, (cuuint64_t)a.B};
    cuuint64_t strides[4] = {(cuuint64_t)a.x_cs * 2, (cuuint64_t)a.Wi * a.x_cs * 2, (cuuint64_t)a.Hi * a.Wi * a.x_cs * 2,
                             (cuuint64_t)a.Di * a.Hi * a.Wi * a.x_cs * 2};
    const cuuint32_t es = h.s2 ? 2 : 1;                // stride 2: one parity class per load (9 x 17 voxels out of an 18 x 34 window)
    cuuint32_t box[5] = {(cuuint32_t)(a.Cin / h.KCH), (cuuint32_t)(a.transposed || h.s2 ? HT_W : HALO_W) * es,
                         (cuuint32_t)(a.transposed || h.s2 ? HT_H : HALO_H) * es, 1, 1};
    cuuint32_t estr[5] = {1, es, es, 1, 1};
    void* base = const_cast<void*>(static_cast<const void*>(static_cast<const __nv_bfloat16*>(a.x) + a.x_co));
    if (!make_map(&tmA, base, 5, dims, strides, box, estr, (int)h.rowb)) return COMA_ERR_CUDA;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.Cin, (cuuint64_t)27 * a.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)(a.Cin / h.KCH), (cuuint32_t)h.NT};
    cuuint32_t estr[2] = {1, 1};
    if (!make_map(&tmB, const_cast<void*>(a.w), 2, dims, strides, box, estr, (int)h.rowb)) return COMA_ERR_CUDA;
  }
#define COMA_HALO_CASE(NTV, KCV) if (h.NT == NTV && a.Cin == KCV) return launch_halo<NTV, KCV>(a, h, tmA, tmB, stream);
  COMA_HALO_CASE(16, 16) COMA_HALO_CASE(16, 32) COMA_HALO_CASE(16, 64)
  COMA_HALO_CASE(32, 16) COMA_HALO_CASE(32, 32) COMA_HALO_CASE(32, 64)
  COMA_HALO_CASE(64, 16) COMA_HALO_CASE(64, 32) COMA_HALO_CASE(64, 64)
#undef COMA_HALO_CASE
  if (a.Cin == 128 && h.KCH == 2) {
    if (h.NT == 16) return launch_halo<16, 64, 2>(a, h, tmA, tmB, stream);
    if (h.NT == 32) return launch_halo<32, 64, 2>(a, h, tmA, tmB, stream);
  }
  set_error("conv_halo: unsupported channel combination");
  return COMA_ERR_UNSUPPORTED;
}

int conv_tc_launch(const coma_conv_args& a, cudaStream_t stream) {
  if (a.dtype == COMA_BF16) {      // the plane-ring / CTA-pair / tap-packed kernels store bf16 only
    {
      const PairPlan pp = plan_pair(a);
      if (pp.ok) return conv_pair_launch(a, pp, stream);
    }
    {
      const TapsPlan t = plan_taps(a);
      if (t.ok) return conv_taps_launch(a, t, stream);
    }
    {
      const HaloPlan h = plan_halo(a);
      if (h.ok) return conv_halo_launch(a, h, stream);
    }
  }
  TcParams p{};
  p.B = a.B; p.Di = a.Di; p.Hi = a.Hi; p.Wi = a.Wi; p.Do = a.Do; p.Ho = a.Ho; p.Wo = a.Wo; p.Cin = a.Cin; p.Cout = a.Cout;
  p.ksize = a.ksize; p.stride = a.stride; p.pad = a.pad; p.transposed = a.transposed;
  p.KC = pick_kc(a.Cin); p.kchunks = a.Cin / p.KC;
  p.NT = pick_nt(a.Cout); p.n_tiles = a.Cout / p.NT;
  tile_counts(a, p.tiles_w, p.tiles_h, p.tiles_d, p.classes);
  p.total_tiles = a.B * p.classes * p.tiles_d * p.tiles_h * p.tiles_w * p.n_tiles;
  p.swz = p.KC * 2;
  p.a_bytes = 128u * p.KC * 2u;
  p.b_bytes = (uint32_t)p.NT * p.KC * 2u;
  p.stage_bytes = p.a_bytes + ((p.b_bytes + 1023u) & ~1023u);
  p.out_f32 = a.dtype == COMA_BF16_F32OUT;
  p.y = p.out_f32 ? reinterpret_cast<__nv_bfloat16*>(static_cast<float*>(a.y) + a.y_co) : static_cast<__nv_bfloat16*>(a.y) + a.y_co;
  p.y_cs = a.y_cs; p.y_cn = a.y_cn;
  p.bias = a.bias; p.scale = a.scale; p.shift = a.shift; p.slope = a.slope; p.stats = a.stats; p.act = a.act;
  p.stat_chunks = p.tiles_w * p.tiles_h * p.tiles_d * p.classes;
  uint32_t cols = 32;
  while (cols < 2u * p.NT) cols <<= 1;
  p.tmem_cols = cols;

  const size_t tail = 2 * kMaxStages * 8 + 4 * 8 + 16 + (size_t)11 * p.NT * sizeof(float) + 64;
  const size_t budget = 200 * 1024;
  int stages = (int)((budget - tail - 1024) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) { set_error("conv_tc: stage of %u bytes does not fit", p.stage_bytes); return COMA_ERR_UNSUPPORTED; }
  p.stages = stages;
  const size_t smem_bytes = 1024 + (size_t)stages * p.stage_bytes + tail;

  // A: activations [B, D, H, W, C] (innermost first for TMA), element stride = conv stride on the spatial dims
  CUtensorMap tmA, tmB;
  {
    const int es = a.transposed ? 1 : a.stride;
    cuuint64_t dims[5] = {(cuuint64_t)a.Cin, (cuuint64_t)a.Wi, (cuuint64_t)a.Hi, (cuuint64_t)a.Di, (cuuint64_t)a.B};
    cuuint64_t strides[4] = {(cuuint64_t)a.x_cs * 2, (cuuint64_t)a.Wi * a.x_cs * 2, (cuuint64_t)a.Hi * a.Wi * a.x_cs * 2,
                             (cuuint64_t)a.Di * a.Hi * a.Wi * a.x_cs * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.KC, (cuuint32_t)(TW * es), (cuuint32_t)(TH * es), (cuuint32_t)(TD * es), 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)es, (cuuint32_t)es, (cuuint32_t)es, 1};
    void* base = const_cast<void*>(static_cast<const void*>(static_cast<const __nv_bfloat16*>(a.x) + a.x_co));
    if (!make_map(&tmA, base, 5, dims, strides, box, estr, p.swz)) return COMA_ERR_CUDA;
  }
  {
    const int taps = a.ksize * a.ksize * a.ksize;
    cuuint64_t dims[2] = {(cuuint64_t)a.Cin, (cuuint64_t)taps * a.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)a.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.KC, (cuuint32_t)p.NT};
    cuuint32_t estr[2] = {1, 1};
    if (!make_map(&tmB, const_cast<void*>(a.w), 2, dims, strides, box, estr, p.swz)) return COMA_ERR_CUDA;
  }

  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  int grid = num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  conv_tc_kernel<<<grid, kTcThreads, smem_bytes, stream>>>(tmA, tmB, p);
  COMA_CHECK_LAUNCH("conv_tc");
  return COMA_OK;
}

}  // namespace coma
