// Weight gradient on the warp-level tensor-core path (mma.sync m16n8k16, bf16 x bf16 -> fp32):
//     dw[tap][cg][cx] += sum_o g[o][cg] * x[o*stride + k - pad][cx]
// Replaces cuDNN wgrad behind loss.backward() (attn_unet_data_parallel.py:884) for every Conv3d / ConvTranspose3d.
// GEMM view per tap: M = Cg (tile 64), N = Cx (tile 64), K = voxels (split over CTAs, fp32 atomics at the end).
// Operands are staged voxel-major ([K][channels], the NDHWC layout as is) with cp.async (zero-fill outside the volume
// = conv padding) and read with ldmatrix.trans.  This is the interim wgrad: the tcgen05 version (MN-major operands out
// of the same halo slabs as the forward kernel) is described in DESIGN.md "next".
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
}  // namespace

bool wgrad_mma_supported(const coma_wgrad_args& a) {
  return a.dtype == COMA_BF16 && a.Cg % 8 == 0 && a.Cx % 8 == 0 && a.g_cs % 8 == 0 && a.g_co % 8 == 0 && a.x_cs % 8 == 0 &&
         a.x_co % 8 == 0 && (reinterpret_cast<uintptr_t>(a.g) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
}

int wgrad_mma_launch(const coma_wgrad_args& a, cudaStream_t stream) {
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
