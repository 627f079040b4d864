// Weight gradient on the warp-level tensor-core path (mma.sync m16n8k16, bf16 x bf16 -> fp32):
//     dw[tap][cg][cx] += sum_o g[o][cg] * x[o*stride + k - pad][cx]
// Replaces cuDNN wgrad behind loss.backward() (attn_unet_data_parallel.py:884) for every Conv3d / ConvTranspose3d.
// GEMM view per tap: M = Cg (tile 64), N = Cx (tile 64), K = voxels (split over CTAs, fp32 atomics at the end).
// Operands are staged voxel-major ([K][channels], the NDHWC layout as is) with cp.async (zero-fill outside the volume
// = conv padding) and read with ldmatrix.trans.  This is the interim wgrad: the tcgen05 version (MN-major operands out
// of the same halo slabs as the forward kernel) is described in DESIGN.md "next".
#include <cstdlib>

#include "common.cuh"

namespace coma {

namespace {
constexpr int WT = 64;        // output tile (cg x cx)
constexpr int KS = 32;        // voxels per stage
constexpr int LDS = WT + 8;   // padded smem row (elements): 144 B rows keep ldmatrix conflict-free

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) wgrad_mma_kernel(coma_wgrad_args a, int64_t vchunk, int cx_tiles) {
  __shared__ __align__(16) __nv_bfloat16 sg[2][KS][LDS];
  __shared__ __align__(16) __nv_bfloat16 sx[2][KS][LDS];
  const int K = a.ksize;
  const int tap = blockIdx.y;
  const int kd = tap / (K * K), kh = (tap / K) % K, kw = tap % K;
  const int cg0 = (blockIdx.z / cx_tiles) * WT, cx0 = (blockIdx.z % cx_tiles) * WT;
  const int64_t Vg = (int64_t)a.Dg * a.Hg * a.Wg, total = (int64_t)a.B * Vg;
  const int64_t begin = (int64_t)blockIdx.x * vchunk, end = min(begin + vchunk, total);
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(a.g) + a.g_co;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(a.x) + a.x_co;

  const int lrow = threadIdx.x >> 3, lvec = (threadIdx.x & 7) * 8;       // this thread's 16-byte piece of a stage
  const bool g_on = cg0 + lvec < a.Cg, x_on = cx0 + lvec < a.Cx;

  auto load_stage = [&](int buf, int64_t base) {
    const int64_t o = base + lrow;
    const bool in = o < end;
    const __nv_bfloat16* gs = gp;
    const __nv_bfloat16* xs = xp;
    bool gv = in && g_on, xv = false;
    if (gv) gs = gp + o * a.g_cs + cg0 + lvec;
    if (in && x_on) {
      const int64_t bb = o / Vg, rem = o - bb * Vg;
      const int ow = (int)(rem % a.Wg), t2 = (int)(rem / a.Wg);
      const int oh = t2 % a.Hg, od = t2 / a.Hg;
      const int id = od * a.stride + kd - a.pad, ih = oh * a.stride + kh - a.pad, iw = ow * a.stride + kw - a.pad;
      if (id >= 0 && id < a.Dx && ih >= 0 && ih < a.Hx && iw >= 0 && iw < a.Wx) {
        xv = true;
        xs = xp + (((bb * a.Dx + id) * a.Hx + ih) * a.Wx + iw) * a.x_cs + cx0 + lvec;
      }
    }
    cp_async16(&sg[buf][lrow][lvec], gs, gv);
    cp_async16(&sx[buf][lrow][lvec], xs, xv);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t nstages = (end - begin + KS - 1) / KS;
  if (nstages > 0) load_stage(0, begin);
  for (int64_t s = 0; s < nstages; ++s) {
    const int buf = (int)(s & 1);
    if (s + 1 < nstages) {
      load_stage(buf ^ 1, begin + (s + 1) * KS);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < KS; k0 += 16) {
      // A fragment: rows = cg (m0..m0+15), cols = voxels (k0..k0+15), storage [voxel][cg] -> ldmatrix.trans
      uint32_t af[4];
      {
        const int mat = lane >> 3, r = lane & 7;
        ldsm_x4_t(af, &sg[buf][k0 + (mat >> 1) * 8 + r][m0 + (mat & 1) * 8]);
      }
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        uint32_t bf[4];   // (k 0-7, n), (k 8-15, n), (k 0-7, n+8), (k 8-15, n+8)
        const int mat = lane >> 3, r = lane & 7;
        ldsm_x4_t(bf, &sx[buf][k0 + (mat & 1) * 8 + r][n0 + nb * 16 + (mat >> 1) * 8]);
        mma_bf16(acc[nb * 2], af, bf[0], bf[1]);
        mma_bf16(acc[nb * 2 + 1], af, bf[2], bf[3]);
      }
    }
    __syncthreads();
  }
  // accumulator fragment: c0,c1 -> row lane/4, cols 2*(lane%4)+{0,1}; c2,c3 -> row + 8
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int cg = cg0 + m0 + (lane >> 2) + (e >> 1) * 8;
      const int cx = cx0 + n0 + nt * 8 + (lane & 3) * 2 + (e & 1);
      if (cg < a.Cg && cx < a.Cx) atomicAdd(a.dw + ((int64_t)tap * a.Cg + cg) * a.Cx + cx, acc[nt][e]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Halo variant for 3x3x3 stride-1 layers with few channels (the 128^3 / 64^3 levels): the per-tap kernel above re-reads
// g and x once per tap (27x) and is bound by that; here a CTA stages one 8x8x4 voxel block of g and the 10x10x6 halo
// block of x in shared memory ONCE (cp.async, double buffered) and accumulates all 27 taps from it, one
// [CgT x CxT] accumulator per tap held in registers across the whole persistent loop (27 x 4 registers per thread).
// ------------------------------------------------------------------------------------------------
// S = conv stride: stride 1 stages 8x8x4 output voxels + a 10x10x6 halo; stride 2 (x index = 2 o + k - 1) stages 8x4x2 output
// voxels + the 17x9x5 input voxels they touch, and the B fragments step two halo rows per voxel
template <int S>
struct HaloGeo {
  static constexpr int VW = 8, VH = S == 1 ? 8 : 4, VD = S == 1 ? 4 : 2, VOX = VW * VH * VD;
  static constexpr int E = S == 1 ? 2 : 1;
  static constexpr int XW = S * VW + E, XH = S * VH + E, XD = S * VD + E, XROWS = XW * XH * XD;
};

__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}

template <int CGT, int CXT, int S>
__global__ void __launch_bounds__(256, 1) wgrad_halo_kernel(coma_wgrad_args a, int cx_tiles, int nbw, int nbh, int nbd) {
  using G = HaloGeo<S>;
  constexpr int VW = G::VW, VH = G::VH, VD = G::VD, VOX = G::VOX, XW = G::XW, XH = G::XH, XROWS = G::XROWS;
  constexpr int LDG = CGT + 8, LDX = CXT + 8;
  constexpr int MT = CGT / 16, NTL = CXT / 8, T = MT * NTL, WT_ = 8 / T;   // tiles, warps per tile (k split)
  static_assert(T <= 8 && 8 % T == 0, "tile layout");
  extern __shared__ __align__(16) uint8_t wsm[];
  __nv_bfloat16* sg[2];
  __nv_bfloat16* sx[2];
  sg[0] = reinterpret_cast<__nv_bfloat16*>(wsm);
  sx[0] = sg[0] + VOX * LDG;
  sg[1] = sx[0] + XROWS * LDX;
  sx[1] = sg[1] + VOX * LDG;

  const int cg0 = (blockIdx.y / cx_tiles) * CGT, cx0 = (blockIdx.y % cx_tiles) * CXT;
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(a.g) + a.g_co + cg0;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(a.x) + a.x_co + cx0;
  const int nblocks = a.B * nbd * nbh * nbw;

  auto load_block = [&](int buf, int blk) {
    int t = blk;
    const int wb = t % nbw; t /= nbw;
    const int hb = t % nbh; t /= nbh;
    const int db = t % nbd; t /= nbd;
    const int64_t b = t;
    const int w0 = wb * VW, h0 = hb * VH, d0 = db * VD;
    constexpr int GV = CGT / 8, XV = CXT / 8;
    for (int i = threadIdx.x; i < VOX * GV; i += 256) {
      const int row = i / GV, vec = (i % GV) * 8;
      const int w = w0 + (row % VW), h = h0 + (row / VW) % VH, d = d0 + row / (VW * VH);
      const bool ok = w < a.Wg && h < a.Hg && d < a.Dg;
      const __nv_bfloat16* src = ok ? gp + (((b * a.Dg + d) * a.Hg + h) * a.Wg + w) * a.g_cs + vec : gp;
      cp_async16(sg[buf] + row * LDG + vec, src, ok);
    }
    for (int i = threadIdx.x; i < XROWS * XV; i += 256) {
      const int row = i / XV, vec = (i % XV) * 8;
      const int w = S * w0 - 1 + (row % XW), h = S * h0 - 1 + (row / XW) % XH, d = S * d0 - 1 + row / (XW * XH);
      const bool ok = w >= 0 && w < a.Wx && h >= 0 && h < a.Hx && d >= 0 && d < a.Dx;
      const __nv_bfloat16* src = ok ? xp + (((b * a.Dx + d) * a.Hx + h) * a.Wx + w) * a.x_cs + vec : xp;
      cp_async16(sx[buf] + row * LDX + vec, src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = warp % T, kslice = warp / T;
  const int m0 = (tile % MT) * 16, n0 = (tile / MT) * 8;
  float acc[27][4];
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;

  int blk = blockIdx.x, it = 0;
  if (blk < nblocks) load_block(0, blk);
  for (; blk < nblocks; blk += gridDim.x, ++it) {
    const int buf = it & 1;
    const int next = blk + gridDim.x;
    if (next < nblocks) {
      load_block(buf ^ 1, next);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const __nv_bfloat16* g_s = sg[buf];
    const __nv_bfloat16* x_s = sx[buf];
#pragma unroll 1
    for (int ks = kslice; ks < VOX / 16; ks += WT_) {
      const int d = ks / (VH / 2), hp = ks % (VH / 2);           // 16 voxels = two h-lines (2hp, 2hp+1) of plane d
      uint32_t af[4];
      {
        const int mat = lane >> 3, r = lane & 7;
        ldsm_x4_t(af, g_s + ((d * VH + 2 * hp) * VW + (mat >> 1) * 8 + r) * LDG + m0 + (mat & 1) * 8);
      }
      const int lrow = ((lane >> 3) & 1) * (S * XW) + (lane & 7) * S;      // second 8 voxels are the next h-line: + S * XW halo rows
#pragma unroll
      for (int kd = 0; kd < 3; ++kd)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            uint32_t bf[2];
            ldsm_x2_t(bf, x_s + (((S * d + kd) * XH + S * 2 * hp + kh) * XW + kw + lrow) * LDX + n0);
            mma_bf16(acc[(kd * 3 + kh) * 3 + kw], af, bf[0], bf[1]);
          }
    }
    __syncthreads();
  }
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int cg = cg0 + m0 + (lane >> 2) + (e >> 1) * 8;
      const int cx = cx0 + n0 + (lane & 3) * 2 + (e & 1);
      atomicAdd(a.dw + ((int64_t)t * a.Cg + cg) * a.Cx + cx, acc[t][e]);
    }
}


// ------------------------------------------------------------------------------------------------
// Second halo variant: the kernel above gives every warp one m16 x n8 tile for all 27 taps, so each x fragment it reads
// from shared memory feeds ONE mma (about 15 flop per shared-memory byte: ldmatrix bandwidth, not the tensor pipe, bounds it).
// Here the 27 taps are dealt round-robin to TG = 8 / (CXT / 16) warp groups and a warp owns the whole CGT x 16 tile of its taps:
// one x fragment (ldmatrix.x4, n16 x k16) feeds CGT / 8 mmas and the g fragments are shared by all taps of the warp
// (about 26 flop per byte at 32 x 32).  Staging, geometry and the atomic epilogue are those of the kernel above.
// ------------------------------------------------------------------------------------------------
template <int CGT, int CXT, int S>
__global__ void __launch_bounds__(256, (CGT + CXT <= 32 ? 2 : 1)) wgrad_halo_taps_kernel(coma_wgrad_args a, int cx_tiles, int nbw, int nbh, int nbd) {
  using G = HaloGeo<S>;
  constexpr int VW = G::VW, VH = G::VH, VD = G::VD, VOX = G::VOX, XW = G::XW, XH = G::XH, XROWS = G::XROWS;
  constexpr int LDG = CGT + 8, LDX = CXT + 8;
  constexpr int MI = CGT / 16, NH = CXT / 16, TG = 8 / NH, TPG = (27 + TG - 1) / TG;
  static_assert(NH == 1 || NH == 2, "warp layout");
  extern __shared__ __align__(16) uint8_t wsm[];
  __nv_bfloat16* sg[2];
  __nv_bfloat16* sx[2];
  sg[0] = reinterpret_cast<__nv_bfloat16*>(wsm);
  sx[0] = sg[0] + VOX * LDG;
  sg[1] = sx[0] + XROWS * LDX;
  sx[1] = sg[1] + VOX * LDG;

  const int cg0 = (blockIdx.y / cx_tiles) * CGT, cx0 = (blockIdx.y % cx_tiles) * CXT;
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(a.g) + a.g_co + cg0;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(a.x) + a.x_co + cx0;
  const int nblocks = a.B * nbd * nbh * nbw;

  auto load_block = [&](int buf, int blk) {
    int t = blk;
    const int wb = t % nbw; t /= nbw;
    const int hb = t % nbh; t /= nbh;
    const int db = t % nbd; t /= nbd;
    const int64_t b = t;
    const int w0 = wb * VW, h0 = hb * VH, d0 = db * VD;
    constexpr int GV = CGT / 8, XV = CXT / 8;
    for (int i = threadIdx.x; i < VOX * GV; i += 256) {
      const int row = i / GV, vec = (i % GV) * 8;
      const int w = w0 + (row % VW), h = h0 + (row / VW) % VH, d = d0 + row / (VW * VH);
      const bool ok = w < a.Wg && h < a.Hg && d < a.Dg;
      const __nv_bfloat16* src = ok ? gp + (((b * a.Dg + d) * a.Hg + h) * a.Wg + w) * a.g_cs + vec : gp;
      cp_async16(sg[buf] + row * LDG + vec, src, ok);
    }
    for (int i = threadIdx.x; i < XROWS * XV; i += 256) {
      const int row = i / XV, vec = (i % XV) * 8;
      const int w = S * w0 - 1 + (row % XW), h = S * h0 - 1 + (row / XW) % XH, d = S * d0 - 1 + row / (XW * XH);
      const bool ok = w >= 0 && w < a.Wx && h >= 0 && h < a.Hx && d >= 0 && d < a.Dx;
      const __nv_bfloat16* src = ok ? xp + (((b * a.Dx + d) * a.Hx + h) * a.Wx + w) * a.x_cs + vec : xp;
      cp_async16(sx[buf] + row * LDX + vec, src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = (warp / NH) & (TG - 1), n0 = (warp % NH) * 16;   // tap group, cx offset of this warp's 16 columns
  const bool last_tap = q + (TPG - 1) * TG < 27;                // only the last tap slot of a group can be empty
  static_assert((TPG - 2) * TG + TG - 1 < 27, "tap slots");
  int tap_off[TPG];                                          // halo-row offset of each of this warp's taps (t = q + j * TG)
#pragma unroll
  for (int j = 0; j < TPG; ++j) {
    const int t = min(q + j * TG, 26);
    tap_off[j] = ((t / 9) * XH + (t / 3) % 3) * XW + t % 3;
  }
  float acc[TPG][MI][2][4];
#pragma unroll
  for (int j = 0; j < TPG; ++j)
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][mi][ni][e] = 0.f;

  const int mat = lane >> 3, r = lane & 7;
  const int g_lane = ((mat >> 1) * 8 + r) * LDG + (mat & 1) * 8;                           // A: (m 0-7|8-15) x (k 0-7|8-15)
  const int x_lane = ((mat & 1) * (S * XW) + r * S) * LDX + n0 + (mat >> 1) * 8;           // B: (k 0-7, n) (k 8-15, n) (k 0-7, n+8) (k 8-15, n+8)

  int blk = blockIdx.x, it = 0;
  if (blk < nblocks) load_block(0, blk);
  for (; blk < nblocks; blk += gridDim.x, ++it) {
    const int buf = it & 1;
    const int next = blk + gridDim.x;
    if (next < nblocks) {
      load_block(buf ^ 1, next);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const __nv_bfloat16* g_s = sg[buf] + g_lane;
    const __nv_bfloat16* x_s = sx[buf] + x_lane;
    // fragments of k-step ks + 1 are fetched while the mmas of k-step ks issue (register double buffer, two k-steps per trip)
    uint32_t af[2][MI][4], bf[2][TPG][4];
    auto fetch = [&](int ks, uint32_t (&fa)[MI][4], uint32_t (&fb)[TPG][4]) {
      const int d = ks / (VH / 2), hp = ks % (VH / 2);           // 16 voxels = two h-lines (2hp, 2hp+1) of plane d
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) ldsm_x4_t(fa[mi], g_s + ((d * VH + 2 * hp) * VW) * LDG + mi * 16);
      const __nv_bfloat16* x_k = x_s + ((S * d * XH + S * 2 * hp) * XW) * LDX;
#pragma unroll
      for (int j = 0; j < TPG; ++j)
        if (j < TPG - 1 || last_tap) ldsm_x4_t(fb[j], x_k + tap_off[j] * LDX);
    };
    auto issue = [&](const uint32_t (&fa)[MI][4], const uint32_t (&fb)[TPG][4]) {
#pragma unroll
      for (int j = 0; j < TPG; ++j)
        if (j < TPG - 1 || last_tap) {
#pragma unroll
          for (int mi = 0; mi < MI; ++mi) {
            mma_bf16(acc[j][mi][0], fa[mi], fb[j][0], fb[j][1]);
            mma_bf16(acc[j][mi][1], fa[mi], fb[j][2], fb[j][3]);
          }
        }
    };
    static_assert((VOX / 16) % 2 == 0, "two k-steps per trip");
    fetch(0, af[0], bf[0]);
#pragma unroll 1
    for (int ks = 0; ks < VOX / 16; ks += 2) {
      fetch(ks + 1, af[1], bf[1]);
      issue(af[0], bf[0]);
      if (ks + 2 < VOX / 16) fetch(ks + 2, af[0], bf[0]);
      issue(af[1], bf[1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < TPG; ++j) {
    const int t = q + j * TG;
    if (t < 27) {
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cg = cg0 + mi * 16 + (lane >> 2) + (e >> 1) * 8;
            const int cx = cx0 + n0 + ni * 8 + (lane & 3) * 2 + (e & 1);
            atomicAdd(a.dw + ((int64_t)t * a.Cg + cg) * a.Cx + cx, acc[j][mi][ni][e]);
          }
    }
  }
}

template <int CGT, int CXT, int S = 1>
int launch_wgrad_halo(const coma_wgrad_args& a, cudaStream_t stream) {
  using G = HaloGeo<S>;
  constexpr int VW = G::VW, VH = G::VH, VD = G::VD, VOX = G::VOX, XROWS = G::XROWS;
  constexpr int LDG = CGT + 8, LDX = CXT + 8;
  const size_t smem = (size_t)2 * (VOX * LDG + XROWS * LDX) * sizeof(__nv_bfloat16);
  static const bool v1 = [] { const char* e = getenv("COMA_WGRAD_HALO_V1"); return e && e[0] == '1'; }();   // the one-tile-per-warp kernel
  static bool set = false;
  static int resident = 1;       // CTAs of the tap-split kernel that fit one SM (the narrow tiles leave room for two)
  if (!set) {
    cudaFuncSetAttribute(wgrad_halo_kernel<CGT, CXT, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(wgrad_halo_taps_kernel<CGT, CXT, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, wgrad_halo_taps_kernel<CGT, CXT, S>, 256, smem) != cudaSuccess || resident < 1)
      resident = 1;
    if (resident > 2) resident = 2;
    set = true;
  }
  const int nbw = (a.Wg + VW - 1) / VW, nbh = (a.Hg + VH - 1) / VH, nbd = (a.Dg + VD - 1) / VD;
  const int nblocks = a.B * nbw * nbh * nbd;
  const int cg_tiles = a.Cg / CGT, cx_tiles = a.Cx / CXT;
  int gx = num_sms() * (v1 ? 1 : resident) / (cg_tiles * cx_tiles);
  if (gx < 1) gx = 1;
  if (gx > nblocks) gx = nblocks;
  dim3 grid((unsigned)gx, (unsigned)(cg_tiles * cx_tiles));
  if (v1) wgrad_halo_kernel<CGT, CXT, S><<<grid, 256, smem, stream>>>(a, cx_tiles, nbw, nbh, nbd);
  else wgrad_halo_taps_kernel<CGT, CXT, S><<<grid, 256, smem, stream>>>(a, cx_tiles, nbw, nbh, nbd);
  COMA_CHECK_LAUNCH("wgrad_halo");
  return COMA_OK;
}

// ------------------------------------------------------------------------------------------------
// Pointwise (k = 1) weight gradient with few channels (the gate's W_g / W_x: Cg = 16, Cx = 32 at 128^3): dw[cg][cx] = sum_v g[v][cg] x[v][cx]
// is one pass over 0.8 GB, but on the 64 x 64-tile per-tap kernel above it took 0.68 ms (7/8 of every staged tile and mma is padding).
// Here the tile is exactly Cg x Cx, the eight warps of a block split the VOXELS of a 256-voxel stage (32 each, two k-steps) and
// their accumulators are summed through shared memory before one set of atomics per block.
// ------------------------------------------------------------------------------------------------
template <int CG, int CX>
__global__ void __launch_bounds__(256) wgrad_pw_kernel(coma_wgrad_args a, int64_t vchunk) {
  constexpr int VS = 256;                                  // voxels per stage
  constexpr int LDG = CG + 8, LDX = CX + 8;                // padded rows keep ldmatrix conflict-free
  constexpr int MT = CG / 16, NB = CX / 16;                // m16 tiles, n16 blocks
  extern __shared__ __align__(16) uint8_t psm[];
  __nv_bfloat16* sg[2];
  __nv_bfloat16* sx[2];
  sg[0] = reinterpret_cast<__nv_bfloat16*>(psm);
  sx[0] = sg[0] + VS * LDG;
  sg[1] = sx[0] + VS * LDX;
  sx[1] = sg[1] + VS * LDG;
  const int64_t total = (int64_t)a.B * a.Dg * a.Hg * a.Wg;
  const int64_t begin = (int64_t)blockIdx.x * vchunk, end = min(begin + vchunk, total);
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(a.g) + a.g_co;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(a.x) + a.x_co;

  auto load_stage = [&](int buf, int64_t base) {
    constexpr int GV = CG / 8, XV = CX / 8;
    for (int i = threadIdx.x; i < VS * GV; i += 256) {
      const int row = i / GV, vec = (i % GV) * 8;
      const bool ok = base + row < end;
      cp_async16(sg[buf] + row * LDG + vec, ok ? gp + (base + row) * a.g_cs + vec : gp, ok);
    }
    for (int i = threadIdx.x; i < VS * XV; i += 256) {
      const int row = i / XV, vec = (i % XV) * 8;
      const bool ok = base + row < end;
      cp_async16(sx[buf] + row * LDX + vec, ok ? xp + (base + row) * a.x_cs + vec : xp, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mat = lane >> 3, r = lane & 7;
  float acc[MT][NB][2][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mi][nb][h][e] = 0.f;

  const int64_t nstages = (end - begin + VS - 1) / VS;
  if (nstages > 0) load_stage(0, begin);
  for (int64_t s = 0; s < nstages; ++s) {
    const int buf = (int)(s & 1);
    if (s + 1 < nstages) {
      load_stage(buf ^ 1, begin + (s + 1) * VS);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 16) {
      const int row0 = warp * 32 + k0;                      // this warp's 16 voxels of the k-step
      uint32_t af[MT][4];
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) ldsm_x4_t(af[mi], sg[buf] + (row0 + (mat >> 1) * 8 + r) * LDG + mi * 16 + (mat & 1) * 8);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        uint32_t bf[4];   // (k 0-7, n), (k 8-15, n), (k 0-7, n+8), (k 8-15, n+8)
        ldsm_x4_t(bf, sx[buf] + (row0 + (mat & 1) * 8 + r) * LDX + nb * 16 + (mat >> 1) * 8);
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
          mma_bf16(acc[mi][nb][0], af[mi], bf[0], bf[1]);
          mma_bf16(acc[mi][nb][1], af[mi], bf[2], bf[3]);
        }
      }
    }
    __syncthreads();
  }
  // sum the eight warps' partial tiles in shared memory (the stage buffers are free now), then one set of atomics per block
  float* red = reinterpret_cast<float*>(psm);               // [CG][CX]
  for (int i = threadIdx.x; i < CG * CX; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cg = mi * 16 + (lane >> 2) + (e >> 1) * 8;
          const int cx = nb * 16 + h * 8 + (lane & 3) * 2 + (e & 1);
          atomicAdd(red + cg * CX + cx, acc[mi][nb][h][e]);
        }
  __syncthreads();
  for (int i = threadIdx.x; i < CG * CX; i += 256) atomicAdd(a.dw + (int64_t)(i / CX) * a.Cx + (i % CX), red[i]);
}

template <int CG, int CX>
int launch_wgrad_pw(const coma_wgrad_args& a, cudaStream_t stream) {
  constexpr size_t smem = (size_t)2 * 256 * ((CG + 8) + (CX + 8)) * sizeof(__nv_bfloat16);
  static_assert(smem >= (size_t)CG * CX * sizeof(float), "reduction buffer aliases the stages");
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(wgrad_pw_kernel<CG, CX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = true; }
  const int64_t total = (int64_t)a.B * a.Dg * a.Hg * a.Wg;
  int64_t want = (int64_t)num_sms() * 4;
  int64_t vchunk = (total + want - 1) / want;
  if (vchunk < 2048) vchunk = 2048;
  vchunk = (vchunk + 255) / 256 * 256;
  const int64_t nchunks = (total + vchunk - 1) / vchunk;
  wgrad_pw_kernel<CG, CX><<<(unsigned)nchunks, 256, smem, stream>>>(a, vchunk);
  COMA_CHECK_LAUNCH("wgrad_pw");
  return COMA_OK;
}
}  // namespace


// ---- one-channel gradient, 3x3x3, stride 1 (the 16 -> 1 / 8 -> 1 heads of the modulator stacks, attn_unet_data_parallel.py:495-497) ----
//     dw[tap][0][cx] = sum_u x[u][cx] * g[u - (tap - 1)]
// HBM-bound by nature (x: 32 B per voxel, g: 2 B), but 27 x 16 FMA per voxel is too much for the CUDA cores at that rate and
// zero-padding g to a 16-channel row for the tcgen05 kernel cost 0.45 ms per layer (pad + 16-row instance) for 0.05 ms of traffic.
// Here M = 27 taps (two m16 tiles), N = 16 x-channels (two n8 tiles), K = voxels: the B fragments come straight from the x tile
// (rows = voxels, ldmatrix.trans), the A fragments are GATHERED from a three-plane halo tile of the scalar field g.
// A block owns 8 lines x 32 voxels of a plane (one line per warp, two k16 steps) and marches along depth.
namespace {
constexpr int C1_BH = 8, C1_BW = 32, C1_HP = C1_BH + 2, C1_WP = C1_BW + 2;

__global__ void __launch_bounds__(256) wgrad_c1k3_kernel(coma_wgrad_args a, int DC, int nbd, int nbh, int nbw, float* __restrict__ ws) {
  __shared__ __align__(16) __nv_bfloat16 xs[C1_BH * C1_BW * 16];      // [line][voxel][16 ch]: 32-byte rows
  __shared__ __nv_bfloat16 gs[3 * C1_HP * C1_WP];                      // g planes d-1, d, d+1 with an H / W halo
  __shared__ float red[8][32 * 16];
  int it = blockIdx.x;
  const int wb = it % nbw; it /= nbw;
  const int hb = it % nbh; it /= nbh;
  const int db = it % nbd; it /= nbd;
  const int b = it;
  const int d0 = db * DC, d1 = min(d0 + DC, a.Dg), h0 = hb * C1_BH, w0 = wb * C1_BW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* xg = static_cast<const __nv_bfloat16*>(a.x) + a.x_co;
  const __nv_bfloat16* gg = static_cast<const __nv_bfloat16*>(a.g) + a.g_co;
  const int64_t plane = (int64_t)a.Hg * a.Wg;
  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
  // tap offsets of this thread's four A rows (m = lane / 4 + {0, 8} in each of the two m16 tiles); taps >= 27 read tap 26 and are dropped
  int goff[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int tap = (r >> 1) * 16 + (r & 1) * 8 + (lane >> 2);
    tap = tap < 27 ? tap : 26;
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    // g[u - (k - 1)]: plane (2 - kd), line + (2 - kh), column + (2 - kw) inside the halo tile (u itself sits at +1, +1)
    goff[r] = ((2 - kd) * C1_HP + (2 - kh)) * C1_WP + (2 - kw);
  }
  for (int d = d0; d < d1; ++d) {
    __syncthreads();
    // x tile: 8 lines x 32 voxels x 32 bytes = 512 16-byte vectors
    for (int v = tid; v < C1_BH * C1_BW * 2; v += 256) {
      const int line = v / (C1_BW * 2), rem = v % (C1_BW * 2), vox = rem >> 1, half = rem & 1;
      const int64_t gi = (((int64_t)b * a.Dx + d) * a.Hx + h0 + line) * a.Wx + w0 + vox;
      *reinterpret_cast<uint4*>(&xs[(line * C1_BW + vox) * 16 + half * 8]) = __ldg(reinterpret_cast<const uint4*>(xg + gi * a.x_cs + half * 8));
    }
    for (int e = tid; e < 3 * C1_HP * C1_WP; e += 256) {
      const int pz = e / (C1_HP * C1_WP), r = e % (C1_HP * C1_WP), py = r / C1_WP, px = r % C1_WP;
      const int dz = d - 1 + pz, hy = h0 - 1 + py, wx = w0 - 1 + px;
      const bool in = dz >= 0 && dz < a.Dg && hy >= 0 && hy < a.Hg && wx >= 0 && wx < a.Wg;
      gs[e] = in ? gg[(((int64_t)b * a.Dg + dz) * plane + (int64_t)hy * a.Wg + wx) * a.g_cs] : __float2bfloat16(0.f);
    }
    __syncthreads();
    const unsigned short* g16 = reinterpret_cast<const unsigned short*>(gs);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      // B fragments: x[k = voxel][n = cx], two n8 tiles; rows of 32 bytes, matrices 8 voxels x 8 channels
      uint32_t bfr[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const __nv_bfloat16* p = &xs[(warp * C1_BW + ks * 16 + (lane & 15)) * 16 + nt * 8];
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(p);
        asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(bfr[nt][0]), "=r"(bfr[nt][1]) : "r"(sa));
      }
      // A fragments: a0 = (row m, k 2c..2c+1), a1 = (row m+8, same k), a2 = (row m, k+8..), a3 = (row m+8, k+8..); c = lane % 4
      const int kbase = warp * C1_WP + ks * 16 + (lane & 3) * 2;        // line `warp`, column of voxel k inside the halo tile (before +off)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t af[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int o = goff[mt * 2 + (q & 1)] + kbase + (q >> 1) * 8;
          af[q] = (uint32_t)g16[o] | ((uint32_t)g16[o + 1] << 16);
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) mma_bf16(acc[mt][nt], af, bfr[nt][0], bfr[nt][1]);
      }
    }
  }
  // block reduction over the eight warps, then one partial block [27][16] per CTA (deterministic finish) or atomics
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = mt * 16 + (lane >> 2) + (q >> 1) * 8, n = nt * 8 + (lane & 3) * 2 + (q & 1);
        red[warp][m * 16 + n] = acc[mt][nt][q];
      }
  __syncthreads();
  for (int e = tid; e < 27 * 16; e += 256) {
    float s2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s2 += red[w][e];
    if (ws) ws[(int64_t)blockIdx.x * (27 * 16) + e] = s2;
    else atomicAdd(a.dw + e, s2);
  }
}

// sum the per-CTA blocks in CTA order (eight interleaved partial sums, then a fixed-order combine) and store dw
__global__ void __launch_bounds__(256) wgrad_c1k3_finish_kernel(const float* __restrict__ ws, float* __restrict__ dw, int nblocks) {
  __shared__ float part[8][32];
  const int row = threadIdx.x >> 5, e = blockIdx.x * 32 + (threadIdx.x & 31);
  float s2 = 0.f;
  if (e < 27 * 16)
    for (int c = row; c < nblocks; c += 8) s2 += ws[(int64_t)c * (27 * 16) + e];
  part[row][threadIdx.x & 31] = s2;
  __syncthreads();
  if (row == 0 && e < 27 * 16) {
    float t = part[0][threadIdx.x];
#pragma unroll
    for (int r = 1; r < 8; ++r) t += part[r][threadIdx.x];
    dw[e] = t;
  }
}

struct C1Plan { int DC, nbd, nbh, nbw; int64_t blocks; };
C1Plan plan_c1k3(const coma_wgrad_args& a) {
  C1Plan p;
  p.nbh = a.Hg / C1_BH; p.nbw = a.Wg / C1_BW;
  p.DC = a.Dg;
  while (p.DC > 8 && (int64_t)a.B * ((a.Dg + p.DC - 1) / p.DC) * p.nbh * p.nbw < (int64_t)8 * num_sms()) p.DC = (p.DC + 1) / 2;
  p.nbd = (a.Dg + p.DC - 1) / p.DC;
  p.blocks = (int64_t)a.B * p.nbd * p.nbh * p.nbw;
  return p;
}
}  // namespace

bool wgrad_c1k3_supported(const coma_wgrad_args& a) {
  static const bool off = [] { const char* e = getenv("COMA_DISABLE_WGRAD_C1K3"); return e && e[0] == '1'; }();
  return !off && a.dtype == COMA_BF16 && a.Cg == 1 && a.Cx == 16 && a.ksize == 3 && a.stride == 1 && a.pad == 1 && a.Hg % C1_BH == 0 &&
         a.Wg % C1_BW == 0 && a.x_cs % 8 == 0 && a.x_co % 8 == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
}
int64_t wgrad_c1k3_workspace(const coma_wgrad_args& a) { return plan_c1k3(a).blocks * 27 * 16 * 4; }
int wgrad_c1k3_launch(const coma_wgrad_args& a, cudaStream_t stream) {
  const C1Plan p = plan_c1k3(a);
  float* ws = (a.workspace && a.workspace_bytes >= p.blocks * 27 * 16 * 4) ? static_cast<float*>(a.workspace) : nullptr;
  wgrad_c1k3_kernel<<<(unsigned)p.blocks, 256, 0, stream>>>(a, p.DC, p.nbd, p.nbh, p.nbw, ws);
  COMA_CHECK_LAUNCH("wgrad_c1k3");
  if (ws) {
    wgrad_c1k3_finish_kernel<<<(27 * 16 + 31) / 32, 256, 0, stream>>>(ws, a.dw, (int)p.blocks);
    COMA_CHECK_LAUNCH("wgrad_c1k3_finish");
  }
  return COMA_OK;
}

bool wgrad_mma_supported(const coma_wgrad_args& a) {
  return a.dtype == COMA_BF16 && a.Cg % 8 == 0 && a.Cx % 8 == 0 && a.g_cs % 8 == 0 && a.g_co % 8 == 0 && a.x_cs % 8 == 0 &&
         a.x_co % 8 == 0 && (reinterpret_cast<uintptr_t>(a.g) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
}

int wgrad_mma_launch(const coma_wgrad_args& a, cudaStream_t stream) {
  if (a.ksize == 1 && a.stride == 1 && (a.Cg == 16 || a.Cg == 32) && (a.Cx == 16 || a.Cx == 32 || a.Cx == 64)) {
#define COMA_WGPW_CASE(GV, XV) if (a.Cg == GV && a.Cx == XV) return launch_wgrad_pw<GV, XV>(a, stream);
    COMA_WGPW_CASE(16, 16) COMA_WGPW_CASE(16, 32) COMA_WGPW_CASE(16, 64) COMA_WGPW_CASE(32, 16) COMA_WGPW_CASE(32, 32) COMA_WGPW_CASE(32, 64)
#undef COMA_WGPW_CASE
  }
  static const bool halo_off = [] { const char* e = getenv("COMA_DISABLE_WGRAD_HALO"); return e && e[0] == '1'; }();
  if (!halo_off && a.ksize == 3 && a.stride == 1 && a.Cg % 16 == 0 && a.Cx % 16 == 0 && a.Cg <= 256 && a.Cx <= 256 &&
      (int64_t)a.Dg * a.Hg * a.Wg >= 16 * 16 * 16) {
    const bool g32 = a.Cg % 32 == 0, x32 = a.Cx % 32 == 0;
    if (g32 && x32) return launch_wgrad_halo<32, 32>(a, stream);
    if (g32) return launch_wgrad_halo<32, 16>(a, stream);
    if (x32) return launch_wgrad_halo<16, 32>(a, stream);
    return launch_wgrad_halo<16, 16>(a, stream);
  }
  if (!halo_off && a.ksize == 3 && a.stride == 2 && a.Cg % 16 == 0 && a.Cx % 16 == 0 && a.Cg <= 512 && a.Cx <= 512 &&
      (int64_t)a.Dg * a.Hg * a.Wg >= 8 * 8 * 8) {
    // strided layers (down-sampling convs and, with the roles of x and dy swapped, the transposed convs)
    const bool g32 = a.Cg % 32 == 0, x32 = a.Cx % 32 == 0;
    if (g32 && x32) return launch_wgrad_halo<32, 32, 2>(a, stream);
    if (g32) return launch_wgrad_halo<32, 16, 2>(a, stream);
    if (x32) return launch_wgrad_halo<16, 32, 2>(a, stream);
    return launch_wgrad_halo<16, 16, 2>(a, stream);
  }
  const int64_t total = (int64_t)a.B * a.Dg * a.Hg * a.Wg;
  const int taps = a.ksize * a.ksize * a.ksize;
  const int cg_tiles = (a.Cg + WT - 1) / WT, cx_tiles = (a.Cx + WT - 1) / WT;
  // enough CTAs for ~8 waves, but chunks no shorter than 1024 voxels (bounds the atomic traffic)
  int64_t want = (int64_t)num_sms() * 8 / ((int64_t)taps * cg_tiles * cx_tiles);
  if (want < 1) want = 1;
  int64_t vchunk = (total + want - 1) / want;
  if (vchunk < 1024) vchunk = 1024;
  vchunk = (vchunk + KS - 1) / KS * KS;
  const int64_t nchunks = (total + vchunk - 1) / vchunk;
  dim3 grid((unsigned)nchunks, (unsigned)taps, (unsigned)(cg_tiles * cx_tiles));
  wgrad_mma_kernel<<<grid, 256, 0, stream>>>(a, vchunk, cx_tiles);
  COMA_CHECK_LAUNCH("wgrad_mma");
  return COMA_OK;
}

}  // namespace coma
