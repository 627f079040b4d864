// Weight gradient of the 3x3x3 stride-1 convolutions on the 5th-generation tensor cores (tcgen05 / TMEM, sm_100a):
//     dw[kd,kh,kw][cg][cx] += sum_v g[v][cg] * x[v + (kd-1, kh-1, kw-1)][cx]
// Replaces cuDNN wgrad behind loss.backward() (attn_unet_data_parallel.py:884) for the Conv3d layers whose channels are
// multiples of 16 (ConvBlock convs, AttentionLayer.merge, the modulator stacks; :285-306,495-497 and MONAI's attentionunet).
// The text below describes the 32-channel / 32-voxel-line instance; the kernel is a template over the channels per row of
// either operand (32 -> 64B swizzle, 16 -> 32B swizzle with eight row-shifted chunks in M, three of them useful) and over the
// line length (W % 32 == 0: 8 lines x 32 voxels per plane; W == 16: 16 lines x 16 voxels).
//
// GEMM view.  K = voxels.  Both operands are read straight out of the NDHWC activations, i.e. MN-major (a shared-memory
// row = one voxel, 32 channels = 64 bytes, 64B swizzle as written by TMA): profiles/r01_probe_mn_major_operand.log shows
// that an MN-major descriptor may start at any row, that LBO strides the 32-channel chunks of the M / N extent and SBO the
// 8-row K groups, and that the chunk stride may be a single row.  So ONE tcgen05.mma (M = 128, N = 96, K = 16 voxels of a
// W line) produces nine filter taps at once:
//     A = x slab, 4 chunks one row apart       -> M = (kw = 0, 1, 2, [unused]) x 32 input channels
//     B = g slab, 3 chunks one H line apart    -> N = (kh = 2, 1, 0) x 32 output-gradient channels
// (substituting u = v + (0, kh-1, 0) moves the kh shift onto g: dw = sum_u g[u - (0, kh-1, 0)] * x[u + (kd-1, 0, kw-1)]),
// and kd selects one of three TMEM accumulators fed from the x planes d-1, d, d+1.  128 x 96 x 16 issues in 56 cycles
// (profiles/r01_probe_mma_rate_layout_sbo.log), 3/4 of the M rows are useful.
//
// A CTA owns one (32 input channels) x (32 gradient channels) pair and marches along D through [8 lines x 32 voxels]
// columns of the volume: per plane one TMA box of x (8 lines x 36 rows, W halo, zero fill = conv padding) and one of g
// (10 lines x 32 rows, H halo) go into a ring of four slots; plane d's MMAs read the x slots d-1, d, d+1, so every plane
// is fetched once.  The three [128 x 96] fp32 accumulators stay in TMEM for the CTA's whole life and are added into dw
// with fp32 atomics at the end (the same contract as the mma.sync kernels in wgrad_mma.cu).
// Warp roles (192 threads, 1 CTA / SM): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include <cuda.h>
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace coma {

bool tensor_map_bf16(CUtensorMap* out, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, const cuuint32_t* estr, int swz);

namespace {

// CHX / CHG: channels per x / g row (32 -> 64B swizzle, 16 -> 32B swizzle); BW: voxels per line of a column (W % BW == 0)
template <int CHX, int CHG, int BW>
struct Geo {
  static constexpr int BH = 256 / BW;                         // lines per column: 256 voxels per plane
  static constexpr int MCH = 128 / CHX;                       // row-shifted chunks in M (3 useful)
  static constexpr int XR = BW + MCH;                         // x rows per line: W halo + the rows the unused chunks touch
  static constexpr int GL = BH + 2;                           // g lines per plane: H halo
  static constexpr int ROWBX = CHX * 2, ROWBG = CHG * 2;      // bytes per row
  static constexpr uint32_t X_BYTES = BH * XR * ROWBX, G_BYTES = GL * BW * ROWBG, SLOT_BYTES = X_BYTES + G_BYTES;
  static constexpr int ACC_COLS = 3 * CHG;                    // N = (kh = 2, 1, 0) x CHG
  static constexpr int KSTEPS = 256 / 16, SEGS = BW / 16;
  static_assert(X_BYTES % 1024 == 0 && G_BYTES % 1024 == 0, "slots keep the swizzle phase");
};
constexpr int RING = 4;
constexpr uint32_t TMEM_COLS = 512;
constexpr int kThreads = 192;

struct WgParams {
  float* dw;
  float* ws;                                   // deterministic path: per-CTA partial blocks [gridDim.x][pairs][27][CHG][CHX], or nullptr
  int B, D, H, W, Cg, Cx;
  int cg_tiles, cx_slabs;
  int DC, nd, nh, nw, items;                    // depth chunk, chunks / columns per axis, work items per channel pair
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = s32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();            // a lost TMA / MMA completion must not hang the GPU
  }
}
__device__ __forceinline__ void tma_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void commit_to(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_lane() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// MN-major operand descriptor (rows of ROWB bytes = the swizzle span): LBO = stride between chunks, SBO = stride between 8-row K groups
template <int ROWB>
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
  d |= (uint64_t)(ROWB == 64 ? 4 : 6) << 61;    // 64B / 32B swizzle
  return d;
}

struct Item { int b, d0, nd, h0, w0; };
template <int BH, int BW>
__device__ __forceinline__ Item decode_item(const WgParams& p, int it) {
  Item r;
  const int wb = it % p.nw; it /= p.nw;
  const int hb = it % p.nh; it /= p.nh;
  const int db = it % p.nd; it /= p.nd;
  r.b = it;
  r.d0 = db * p.DC;
  r.nd = min(p.DC, p.D - r.d0);
  r.h0 = hb * BH;
  r.w0 = wb * BW;
  return r;
}

template <int CHX, int CHG, int BW>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                                                               const WgParams p) {
  using G = Geo<CHX, CHG, BW>;
  constexpr int BH = G::BH, XR = G::XR, ROWBX = G::ROWBX, ROWBG = G::ROWBG, ACC_COLS = G::ACC_COLS;
  constexpr uint32_t X_BYTES = G::X_BYTES, SLOT_BYTES = G::SLOT_BYTES;
  extern __shared__ uint8_t wg_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wg_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + RING * SLOT_BYTES);
  uint64_t* empty = full + RING;
  uint64_t* done = empty + RING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y, cg0 = (pair / p.cx_slabs) * CHG, cx0 = (pair % p.cx_slabs) * CHX;
  const bool has_work = (int)blockIdx.x < p.items;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { bar_init(full + i, 1); bar_init(empty + i, 1); }
    bar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: planes d0-1 .. d0+nd of every item
    if (elect_lane()) {
      uint32_t n = 0;
      for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
        const Item c = decode_item<BH, BW>(p, it);
        for (int pl = -1; pl <= c.nd; ++pl, ++n) {
          const int slot = n % RING;
          bar_wait(empty + slot, ((n / RING) & 1) ^ 1);
          uint8_t* dst = smem + slot * SLOT_BYTES;
          const bool with_g = pl >= 0 && pl < c.nd;
          bar_expect_tx(full + slot, with_g ? SLOT_BYTES : X_BYTES);
          tma_5d(dst, &tmX, full + slot, cx0, c.w0 - 1, c.h0, c.d0 + pl, c.b);
          if (with_g) tma_5d(dst + X_BYTES, &tmG, full + slot, cg0, c.w0, c.h0 - 1, c.d0 + pl, c.b);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_lane()) {
      // D f32, A / B bf16, both MN-major (bits 15, 16), N = 3 * CHG, M = 128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (((uint32_t)ACC_COLS >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t base = s32(smem);
      uint32_t n0 = 0, started = 0;
      for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
        const Item c = decode_item<BH, BW>(p, it);
        for (int t = 0; t < c.nd; ++t) {
          // loads n0 + t, n0 + t + 1, n0 + t + 2 hold the x planes d-1, d, d+1; n0 + t + 1 also holds g(d)
          if (t == 0) {
            bar_wait(full + (n0 % RING), (n0 / RING) & 1);
            bar_wait(full + ((n0 + 1) % RING), ((n0 + 1) / RING) & 1);
          }
          bar_wait(full + ((n0 + t + 2) % RING), ((n0 + t + 2) / RING) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t gbase = base + ((n0 + t + 1) % RING) * SLOT_BYTES + X_BYTES;
#pragma unroll 1
          for (int ks = 0; ks < G::KSTEPS; ++ks) {
            const int hl = ks / G::SEGS, seg = ks % G::SEGS;
            const uint64_t bdesc = mn_desc<ROWBG>(gbase + (uint32_t)(hl * BW + seg * 16) * ROWBG, BW * ROWBG, 8 * ROWBG);
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
              const uint32_t xbase = base + ((n0 + t + kd) % RING) * SLOT_BYTES;
              const uint64_t adesc = mn_desc<ROWBX>(xbase + (uint32_t)(hl * XR + seg * 16) * ROWBX, ROWBX, 8 * ROWBX);
              mma_bf16_ss(tmem + kd * ACC_COLS, adesc, bdesc, idesc, started);
            }
            started = 1;
          }
          commit_to(empty + ((n0 + t) % RING));                    // plane d-1 is not needed again
        }
        commit_to(empty + ((n0 + c.nd) % RING));                   // the last two planes of the item
        commit_to(empty + ((n0 + c.nd + 1) % RING));
        n0 += c.nd + 2;
      }
      commit_to(done);
    }
  } else if (has_work) {
    // ---------------------------------------------------------------- epilogue: TMEM -> fp32 atomics into dw[tap][Cg][Cx]
    bar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = (warp & 3) * 32 + lane;                           // TMEM lane = M row = kw * CHX + input channel
    const int kw = m / CHX, cxl = m % CHX;
#pragma unroll 1
    float* wsb = p.ws ? p.ws + ((int64_t)blockIdx.x * gridDim.y + pair) * (27 * CHG * CHX) : nullptr;
#pragma unroll 1
    for (int kd = 0; kd < 3; ++kd)
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        const int tap = (kd * 3 + (2 - j)) * 3 + (kw < 3 ? kw : 0);
        float* dst = p.dw + ((int64_t)tap * p.Cg + cg0) * p.Cx + cx0 + cxl;
        float* wdst = wsb ? wsb + (int64_t)tap * (CHG * CHX) + cxl : nullptr;
#pragma unroll
        for (int part = 0; part < CHG / 16; ++part) {
          uint32_t v[16];
          tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + kd * ACC_COLS + j * CHG + part * 16, v);   // warp-collective
          if (kw < 3) {
            if (wdst) {
#pragma unroll
              for (int i = 0; i < 16; ++i) wdst[(part * 16 + i) * CHX] = __uint_as_float(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) atomicAdd(dst + (int64_t)(part * 16 + i) * p.Cx, __uint_as_float(v[i]));
            }
          }
        }
      }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}


// Deterministic finish: dw[tap][cg][cx] = sum over the CTAs of a channel pair of their partial blocks, in CTA order (a plain
// store: dw need not be zeroed).  ws: [gx][pairs][27][CHG][CHX].
template <int ROWS>
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ ws, float* __restrict__ dw, int gx, int pairs,
                                                           int cx_slabs, int CHG, int CHX, int Cg, int Cx) {
  // A block covers 1024 / ROWS consecutive elements as float4 lanes x ROWS rows; row r sums the CTA blocks r, r + ROWS, ...
  // (16-byte loads, all in flight at once: the kernel is a latency chain over 16 MB at most), then a fixed tree over the rows --
  // one association for a given grid, whatever order the CTAs finished in.  ROWS follows gx (many channel pairs = few CTAs per
  // pair: with 8 rows for gx = 1 seven of eight threads idled, 120 us for the 512 -> 256 merge conv).
  constexpr int LANES = 256 / ROWS;
  __shared__ float4 part[ROWS][LANES];
  const int per = 27 * CHG * CHX;                                  // a multiple of 32
  const int pair = blockIdx.y;
  const int cg0 = (pair / cx_slabs) * CHG, cx0 = (pair % cx_slabs) * CHX;
  const int lane = threadIdx.x % LANES, row = threadIdx.x / LANES, e = (blockIdx.x * LANES + lane) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e < per) {
#pragma unroll 8
    for (int c = row; c < gx; c += ROWS) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ws + ((int64_t)c * pairs + pair) * per + e));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  if (ROWS > 1) {
    part[row][lane] = s;
    __syncthreads();
#pragma unroll
    for (int st = ROWS / 2; st >= 1; st >>= 1) {
      if (row < st) {
        const float4 o = part[row + st][lane];
        float4& m = part[row][lane];
        m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
      }
      __syncthreads();
    }
    s = part[0][lane];
  }
  if (row == 0 && e < per) {
    const int tap = e / (CHG * CHX), r2 = e % (CHG * CHX);         // the four elements share a tap and a gradient channel
    float* o = dw + ((int64_t)tap * Cg + cg0 + r2 / CHX) * Cx + cx0 + r2 % CHX;
    o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w;
  }
}

static int wgrad_finish(const WgParams& p, const coma_wgrad_args& a, int gx, int pairs, int CHG, int CHX, cudaStream_t stream) {
  const int vecs = 27 * CHG * CHX / 4;                             // CHG * CHX is a multiple of 256
#define COMA_FINISH(ROWS)                                                                                                 \
  wgrad_finish_kernel<ROWS><<<dim3((unsigned)((vecs + 256 / ROWS - 1) / (256 / ROWS)), (unsigned)pairs), 256, 0, stream>>>( \
      p.ws, p.dw, gx, pairs, p.cx_slabs, CHG, CHX, a.Cg, a.Cx)
  if (gx >= 24) COMA_FINISH(32);
  else if (gx >= 4) COMA_FINISH(4);
  else COMA_FINISH(1);
#undef COMA_FINISH
  COMA_CHECK_LAUNCH("wgrad_finish");
  return COMA_OK;
}

// =================================================================================================================================
// Stride-2 weight gradient (the four down-sampling convs and, with the roles of x and dy swapped, the four transposed convs):
//     dw[kd,kh,kw][cg][cx] += sum_o g[o][cg] * x[2 o + (kd-1, kh-1, kw-1)][cx]          (g on the coarse grid, x on the fine one)
// The stride breaks the stride-1 trick of moving the kh shift onto g (o -> o + 1 moves x by two voxels), so here
//   * the M = 128 rows of one MMA are (kh = 0, 1, 2, [unused]) x 32 x-channels: the fine lines 2 oh - 1, 2 oh, 2 oh + 1 of a
//     slab are consecutive, so the four M chunks are one line pitch apart (LBO = line pitch, any multiple of 16 bytes is legal);
//   * kw is the PARITY of the fine voxel: TMA element strides (2 along W) deliver every fine plane as an even slab E (fine
//     w = 2 ow, BW rows per line) and an odd slab O (fine w = 2 ow - 1, BW + 1 rows per line); kw = 1 reads E, kw = 0 reads O
//     from row 0, kw = 2 reads O from row 1 -- three MMAs (N = 32 gradient channels, 128 x 32 x 16: 40 cycles) per 16 coarse voxels;
//   * kd is the parity of the fine PLANE: plane 2 od is the kd = 1 operand of coarse plane od, plane 2 od + 1 the kd = 2 operand of
//     od and the kd = 0 operand of od + 1.  Fine planes are consumed in arrival order against one or two resident g planes,
//     so a plane is fetched once (odd W voxels once, not twice) and only a short ring of slots is needed.
// Nine [128 x 32] fp32 accumulators (kd, kw) live in TMEM for the CTA's life (288 columns) and leave through fp32 atomics.
// Peak of the scheme: 96 x 32 x 16 x 2 useful FLOP per 40-cycle MMA = 690 TFLOP/s at 1.9 GHz (the mma.sync kernel it replaces
// runs at ~130); TMA traffic 83 KB per 2880-cycle coarse plane = 29 B/clk/SM.
template <int BW>
struct GeoS2 {
  static constexpr int CH = 32, ROWB = 64;
  static constexpr int BH = 128 / BW;                           // K = 128 coarse voxels per plane: 4 x 32 or 8 x 16
  static constexpr int L = 2 * BH + 1;                          // fine lines per plane
  static constexpr uint32_t E_BYTES = L * BW * ROWB;            // even slab
  static constexpr uint32_t O_RAW = L * (BW + 1) * ROWB;        // odd slab as TMA writes it
  static constexpr uint32_t O_BYTES = (O_RAW + 1023u) & ~1023u;
  static constexpr uint32_t F_BYTES = E_BYTES + O_BYTES;        // one fine-plane slot
  static constexpr uint32_t G_BYTES = BH * BW * ROWB;           // one coarse plane of g
  static constexpr int KSTEPS = 128 / 16, SEGS = BW / 16;
  static_assert(E_BYTES % 1024 == 0 && G_BYTES % 1024 == 0, "slabs keep the swizzle phase");
};
constexpr int FRING = 4, GRING = 3;

template <int BW>
__global__ void __launch_bounds__(kThreads, 1) wgrad_s2_tc_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmO,
                                                                  const __grid_constant__ CUtensorMap tmG, const WgParams p) {
  using G = GeoS2<BW>;
  constexpr int BH = G::BH;
  extern __shared__ uint8_t wg_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wg_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* gsm = smem + FRING * G::F_BYTES;
  uint64_t* ffull = reinterpret_cast<uint64_t*>(gsm + GRING * G::G_BYTES + 4096);     // 4 KB that the unused M chunk may read
  uint64_t* fempty = ffull + FRING;
  uint64_t* gfull = fempty + FRING;
  uint64_t* gempty = gfull + GRING;
  uint64_t* done = gempty + GRING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y, cg0 = (pair / p.cx_slabs) * 32, cx0 = (pair % p.cx_slabs) * 32;
  const bool has_work = (int)blockIdx.x < p.items;

  if (threadIdx.x == 0) {
    for (int i = 0; i < FRING; ++i) { bar_init(ffull + i, 1); bar_init(fempty + i, 1); }
    for (int i = 0; i < GRING; ++i) { bar_init(gfull + i, 1); bar_init(gempty + i, 1); }
    bar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: fine planes 2 d0 - 1 .. 2 (d0 + nd) - 1, g planes d0 ..
    if (elect_lane()) {
      uint32_t nf = 0, ng = 0;
      for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
        const Item c = decode_item<BH, BW>(p, it);
        for (int q = 0; q <= 2 * c.nd; ++q) {
          if ((q & 1) == 0 && (q >> 1) < c.nd) {                       // g[t] is first needed with fine plane 2 t
            const int gs = ng % GRING;
            bar_wait(gempty + gs, ((ng / GRING) & 1) ^ 1);
            bar_expect_tx(gfull + gs, G::G_BYTES);
            tma_5d(gsm + gs * G::G_BYTES, &tmG, gfull + gs, cg0, c.w0, c.h0, c.d0 + (q >> 1), c.b);
            ++ng;
          }
          const int fs = nf % FRING;
          bar_wait(fempty + fs, ((nf / FRING) & 1) ^ 1);
          uint8_t* dst = smem + fs * G::F_BYTES;
          bar_expect_tx(ffull + fs, G::E_BYTES + G::O_RAW);
          const int dz = 2 * c.d0 - 1 + q;
          tma_5d(dst, &tmE, ffull + fs, cx0, 2 * c.w0, 2 * c.h0 - 1, dz, c.b);
          tma_5d(dst + G::E_BYTES, &tmO, ffull + fs, cx0, 2 * c.w0 - 1, 2 * c.h0 - 1, dz, c.b);
          ++nf;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_lane()) {
      // D f32, A / B bf16, both MN-major, N = 32, M = 128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t base = s32(smem), gbase0 = s32(gsm);
      uint32_t nf = 0, ng0 = 0, started = 0;                           // started: bit kd set once the kd accumulators hold data
      for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
        const Item c = decode_item<BH, BW>(p, it);
        for (int q = 0; q <= 2 * c.nd; ++q, ++nf) {
          const int fs = nf % FRING;
          const int t = q >> 1;
          if ((q & 1) == 0 && t < c.nd) bar_wait(gfull + ((ng0 + t) % GRING), ((ng0 + t) / GRING) & 1);
          bar_wait(ffull + fs, (nf / FRING) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t fbase = base + fs * G::F_BYTES;
          // the (kd, g plane) uses of this fine plane
          int kds[2], gts[2], nuse = 0;
          if (q & 1) { kds[0] = 1; gts[0] = t; nuse = 1; }
          else {
            if (t < c.nd) { kds[nuse] = 0; gts[nuse] = t; ++nuse; }
            if (t >= 1) { kds[nuse] = 2; gts[nuse] = t - 1; ++nuse; }
          }
          for (int u = 0; u < nuse; ++u) {
            const int kd = kds[u];
            const uint32_t gb = gbase0 + ((ng0 + gts[u]) % GRING) * G::G_BYTES;
            const uint32_t first = (started >> kd) & 1u;
#pragma unroll 1
            for (int ks = 0; ks < G::KSTEPS; ++ks) {
              const int hl = ks / G::SEGS, seg = ks % G::SEGS;
              const uint64_t bdesc = mn_desc<64>(gb + (uint32_t)(hl * BW + seg * 16) * 64u, 64u, 512u);
              const uint32_t acc = first | (ks > 0 ? 1u : 0u);
              // kw = 1: even slab; kw = 0 / 2: odd slab from row 0 / row 1.  M chunks = the fine lines 2 hl, 2 hl + 1, 2 hl + 2 (kh).
              const uint64_t a1 = mn_desc<64>(fbase + (uint32_t)((2 * hl) * BW + seg * 16) * 64u, (uint32_t)BW * 64u, 512u);
              const uint32_t ob = fbase + G::E_BYTES + (uint32_t)((2 * hl) * (BW + 1) + seg * 16) * 64u;
              const uint64_t a0 = mn_desc<64>(ob, (uint32_t)(BW + 1) * 64u, 512u);
              const uint64_t a2 = mn_desc<64>(ob + 64u, (uint32_t)(BW + 1) * 64u, 512u);
              mma_bf16_ss(tmem + (kd * 3 + 0) * 32, a0, bdesc, idesc, acc);
              mma_bf16_ss(tmem + (kd * 3 + 1) * 32, a1, bdesc, idesc, acc);
              mma_bf16_ss(tmem + (kd * 3 + 2) * 32, a2, bdesc, idesc, acc);
            }
            started |= 1u << kd;
          }
          commit_to(fempty + fs);
          if ((q & 1) == 0 && t >= 1) commit_to(gempty + ((ng0 + t - 1) % GRING));      // g[t-1] has seen its kd = 0, 1, 2 planes
        }
        ng0 += c.nd;
      }
      commit_to(done);
    }
  } else if (has_work) {
    // ---------------------------------------------------------------- epilogue: TMEM -> fp32 atomics into dw[tap][Cg][Cx]
    bar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = (warp & 3) * 32 + lane;                           // TMEM lane = M row = kh * 32 + x channel
    const int kh = m / 32, cxl = m % 32;
    float* wsb = p.ws ? p.ws + ((int64_t)blockIdx.x * gridDim.y + pair) * (27 * 32 * 32) : nullptr;
#pragma unroll 1
    for (int a = 0; a < 9; ++a) {                                   // accumulator a = kd * 3 + kw
      const int kd = a / 3, kw = a % 3;
      const int tap = (kd * 3 + (kh < 3 ? kh : 0)) * 3 + kw;
      float* dst = p.dw + ((int64_t)tap * p.Cg + cg0) * p.Cx + cx0 + cxl;
      float* wdst = wsb ? wsb + (int64_t)tap * (32 * 32) + cxl : nullptr;
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + a * 32 + part * 16, v);   // warp-collective
        if (kh < 3) {
          if (wdst) {
#pragma unroll
            for (int i = 0; i < 16; ++i) wdst[(part * 16 + i) * 32] = __uint_as_float(v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(dst + (int64_t)(part * 16 + i) * p.Cx, __uint_as_float(v[i]));
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <int BW>
int launch_wgrad_s2_tc(const coma_wgrad_args& a, cudaStream_t stream) {
  using G = GeoS2<BW>;
  CUtensorMap tmE, tmO, tmG;
  {
    cuuint64_t dims[5] = {(cuuint64_t)a.Cx, (cuuint64_t)a.Wx, (cuuint64_t)a.Hx, (cuuint64_t)a.Dx, (cuuint64_t)a.B};
    cuuint64_t strides[4] = {(cuuint64_t)a.x_cs * 2, (cuuint64_t)a.Wx * a.x_cs * 2, (cuuint64_t)a.Hx * a.Wx * a.x_cs * 2,
                             (cuuint64_t)a.Dx * a.Hx * a.Wx * a.x_cs * 2};
    cuuint32_t estr[5] = {1, 2, 1, 1, 1};                       // every other fine voxel along W
    void* base = const_cast<void*>(static_cast<const void*>(static_cast<const __nv_bfloat16*>(a.x) + a.x_co));
    cuuint32_t boxE[5] = {32, 2 * BW, (cuuint32_t)G::L, 1, 1};
    cuuint32_t boxO[5] = {32, 2 * (BW + 1), (cuuint32_t)G::L, 1, 1};
    if (!tensor_map_bf16(&tmE, base, 5, dims, strides, boxE, estr, 64)) return COMA_ERR_CUDA;
    if (!tensor_map_bf16(&tmO, base, 5, dims, strides, boxO, estr, 64)) return COMA_ERR_CUDA;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)a.Cg, (cuuint64_t)a.Wg, (cuuint64_t)a.Hg, (cuuint64_t)a.Dg, (cuuint64_t)a.B};
    cuuint64_t strides[4] = {(cuuint64_t)a.g_cs * 2, (cuuint64_t)a.Wg * a.g_cs * 2, (cuuint64_t)a.Hg * a.Wg * a.g_cs * 2,
                             (cuuint64_t)a.Dg * a.Hg * a.Wg * a.g_cs * 2};
    cuuint32_t box[5] = {32, BW, (cuuint32_t)G::BH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = const_cast<void*>(static_cast<const void*>(static_cast<const __nv_bfloat16*>(a.g) + a.g_co));
    if (!tensor_map_bf16(&tmG, base, 5, dims, strides, box, estr, 64)) return COMA_ERR_CUDA;
  }
  WgParams p{};
  p.dw = a.dw;
  p.B = a.B; p.D = a.Dg; p.H = a.Hg; p.W = a.Wg; p.Cg = a.Cg; p.Cx = a.Cx;
  p.cg_tiles = a.Cg / 32;
  p.cx_slabs = a.Cx / 32;
  p.nh = a.Hg / G::BH;
  p.nw = a.Wg / BW;
  const int pairs = p.cg_tiles * p.cx_slabs;
  int gx = num_sms() / pairs;
  if (gx < 1) gx = 1;
  p.DC = a.Dg;
  while (p.DC > 4 && (int64_t)a.B * ((a.Dg + p.DC - 1) / p.DC) * p.nh * p.nw < (int64_t)4 * gx) p.DC = (p.DC + 1) / 2;
  p.nd = (a.Dg + p.DC - 1) / p.DC;
  p.items = a.B * p.nd * p.nh * p.nw;
  if (gx > p.items) gx = p.items;
  const size_t smem = (size_t)FRING * G::F_BYTES + (size_t)GRING * G::G_BYTES + 4096 + 1024 + 256;
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(wgrad_s2_tc_kernel<BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = true; }
  dim3 grid((unsigned)gx, (unsigned)pairs);
  COMA_CHECK_ARG(reinterpret_cast<uintptr_t>(a.workspace) % 16 == 0, "coma_conv3d_wgrad: workspace must be 16-byte aligned");
  p.ws = (a.workspace && a.workspace_bytes >= (int64_t)gx * pairs * 27 * 32 * 32 * 4) ? static_cast<float*>(a.workspace) : nullptr;
  wgrad_s2_tc_kernel<BW><<<grid, kThreads, smem, stream>>>(tmE, tmO, tmG, p);
  COMA_CHECK_LAUNCH("wgrad_s2_tc");
  if (p.ws) return wgrad_finish(p, a, gx, pairs, 32, 32, stream);
  return COMA_OK;
}

static bool wgrad_s2_tc_ok(const coma_wgrad_args& a) {
  static const bool off = [] { const char* e = getenv("COMA_DISABLE_WGRAD_S2_TC"); return e && e[0] == '1'; }();
  const bool lines = (a.Wg % 32 == 0 && a.Hg % 4 == 0) || (a.Wg % 16 == 0 && a.Hg % 8 == 0);
  return !off && a.dtype == COMA_BF16 && a.ksize == 3 && a.stride == 2 && a.pad == 1 && a.Cg % 32 == 0 && a.Cx % 32 == 0 && lines &&
         a.Dg >= 2 && a.g_cs % 8 == 0 && a.g_co % 8 == 0 && a.x_cs % 8 == 0 && a.x_co % 8 == 0 &&
         ((reinterpret_cast<uintptr_t>(a.g) | reinterpret_cast<uintptr_t>(a.x)) & 15) == 0;
}

}  // namespace

namespace {
template <int CHX, int CHG, int BW>
int launch_wgrad_tc(const coma_wgrad_args& a, cudaStream_t stream) {
  using G = Geo<CHX, CHG, BW>;
  CUtensorMap tmX, tmG;
  {
    cuuint64_t dims[5] = {(cuuint64_t)a.Cx, (cuuint64_t)a.Wx, (cuuint64_t)a.Hx, (cuuint64_t)a.Dx, (cuuint64_t)a.B};
    cuuint64_t strides[4] = {(cuuint64_t)a.x_cs * 2, (cuuint64_t)a.Wx * a.x_cs * 2, (cuuint64_t)a.Hx * a.Wx * a.x_cs * 2,
                             (cuuint64_t)a.Dx * a.Hx * a.Wx * a.x_cs * 2};
    cuuint32_t box[5] = {CHX, G::XR, G::BH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = const_cast<void*>(static_cast<const void*>(static_cast<const __nv_bfloat16*>(a.x) + a.x_co));
    if (!tensor_map_bf16(&tmX, base, 5, dims, strides, box, estr, G::ROWBX)) return COMA_ERR_CUDA;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)a.Cg, (cuuint64_t)a.Wg, (cuuint64_t)a.Hg, (cuuint64_t)a.Dg, (cuuint64_t)a.B};
    cuuint64_t strides[4] = {(cuuint64_t)a.g_cs * 2, (cuuint64_t)a.Wg * a.g_cs * 2, (cuuint64_t)a.Hg * a.Wg * a.g_cs * 2,
                             (cuuint64_t)a.Dg * a.Hg * a.Wg * a.g_cs * 2};
    cuuint32_t box[5] = {CHG, BW, G::GL, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = const_cast<void*>(static_cast<const void*>(static_cast<const __nv_bfloat16*>(a.g) + a.g_co));
    if (!tensor_map_bf16(&tmG, base, 5, dims, strides, box, estr, G::ROWBG)) return COMA_ERR_CUDA;
  }
  WgParams p{};
  p.dw = a.dw;
  p.B = a.B; p.D = a.Dg; p.H = a.Hg; p.W = a.Wg; p.Cg = a.Cg; p.Cx = a.Cx;
  p.cg_tiles = a.Cg / CHG;
  p.cx_slabs = a.Cx / CHX;
  p.nh = a.Hg / G::BH;
  p.nw = a.Wg / BW;
  const int pairs = p.cg_tiles * p.cx_slabs;
  int gx = num_sms() / pairs;
  if (gx < 1) gx = 1;
  // depth chunks: long marches amortise the two halo planes; shorter ones balance the CTAs of a channel pair
  p.DC = a.Dg;
  while (p.DC > 8 && (int64_t)a.B * ((a.Dg + p.DC - 1) / p.DC) * p.nh * p.nw < (int64_t)4 * gx) p.DC = (p.DC + 1) / 2;
  p.nd = (a.Dg + p.DC - 1) / p.DC;
  p.items = a.B * p.nd * p.nh * p.nw;
  if (gx > p.items) gx = p.items;
  const size_t smem = (size_t)RING * G::SLOT_BYTES + 1024 + 256;
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(wgrad_tc_kernel<CHX, CHG, BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = true; }
  dim3 grid((unsigned)gx, (unsigned)pairs);
  COMA_CHECK_ARG(reinterpret_cast<uintptr_t>(a.workspace) % 16 == 0, "coma_conv3d_wgrad: workspace must be 16-byte aligned");
  p.ws = (a.workspace && a.workspace_bytes >= (int64_t)gx * pairs * 27 * CHG * CHX * 4) ? static_cast<float*>(a.workspace) : nullptr;
  wgrad_tc_kernel<CHX, CHG, BW><<<grid, kThreads, smem, stream>>>(tmX, tmG, p);
  COMA_CHECK_LAUNCH("wgrad_tc");
  if (p.ws) return wgrad_finish(p, a, gx, pairs, CHG, CHX, stream);
  return COMA_OK;
}
}  // namespace

bool wgrad_tc_supported(const coma_wgrad_args& a) {
  if (a.stride == 2) return wgrad_s2_tc_ok(a);
  static const bool off = [] { const char* e = getenv("COMA_DISABLE_WGRAD_TC"); return e && e[0] == '1'; }();
  const bool lines = (a.Wg % 32 == 0 && a.Hg % 8 == 0) || (a.Wg == 16 && a.Hg % 16 == 0);
  return !off && a.dtype == COMA_BF16 && a.ksize == 3 && a.stride == 1 && a.pad == 1 && a.Cg % 16 == 0 && a.Cx % 16 == 0 && lines &&
         a.Dg >= 4 && a.g_cs % 8 == 0 && a.g_co % 8 == 0 && a.x_cs % 8 == 0 && a.x_co % 8 == 0 &&
         ((reinterpret_cast<uintptr_t>(a.g) | reinterpret_cast<uintptr_t>(a.x)) & 15) == 0;
}

int64_t wgrad_tc_workspace(const coma_wgrad_args& a) {
  // one [27][CHG][CHX] block per CTA; gx * pairs <= max(#SMs, pairs)
  const int chg = a.stride == 2 ? 32 : (a.Cg % 32 == 0 ? 32 : 16), chx = a.stride == 2 ? 32 : (a.Cx % 32 == 0 ? 32 : 16);
  const int64_t pairs = (int64_t)(a.Cg / chg) * (a.Cx / chx);
  const int64_t ctas = std::max<int64_t>(num_sms(), pairs);
  return ctas * 27 * chg * chx * 4;
}

int wgrad_tc_launch(const coma_wgrad_args& a, cudaStream_t stream) {
  if (a.stride == 2) return a.Wg % 32 == 0 && a.Hg % 4 == 0 ? launch_wgrad_s2_tc<32>(a, stream) : launch_wgrad_s2_tc<16>(a, stream);
  const bool x32 = a.Cx % 32 == 0, g32 = a.Cg % 32 == 0, w32 = a.Wg % 32 == 0;
#define COMA_WGTC_CASE(XV, GV, WV) if (x32 == (XV == 32) && g32 == (GV == 32) && w32 == (WV == 32)) return launch_wgrad_tc<XV, GV, WV>(a, stream);
  COMA_WGTC_CASE(32, 32, 32) COMA_WGTC_CASE(32, 16, 32) COMA_WGTC_CASE(16, 32, 32) COMA_WGTC_CASE(16, 16, 32)
  COMA_WGTC_CASE(32, 32, 16) COMA_WGTC_CASE(32, 16, 16) COMA_WGTC_CASE(16, 32, 16) COMA_WGTC_CASE(16, 16, 16)
#undef COMA_WGTC_CASE
  set_error("wgrad_tc: unsupported shape");
  return COMA_ERR_UNSUPPORTED;
}

}  // namespace coma
