// (probe_commit.cu: same harness; MODE 2 adds one / two tcgen05.commit after every 9 MMAs to measure what a commit costs the MMA stream.)
// Throughput probe (one CTA): cycles per SS-mode tcgen05.mma (M=128, K=16) as a function of N, the swizzle mode / row bytes of the
// K-major operands and the A stride-byte-offset, with the halo kernel's tap-shifted A start addresses.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp256(uint32_t t, uint64_t a) { asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(t), "l"(a) : "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int MODE>
__device__ __forceinline__ void body(bool leader, uint32_t tmem, uint32_t a_lo0, uint32_t b_lo0, uint32_t a_hi, uint32_t b_hi,
                                     uint32_t idesc, int N, int iters, uint32_t rowb, uint32_t halo_w, uint32_t bar2) {
  for (int it = 0; it < iters / 18; ++it) {
#pragma unroll
    for (int u = 0; u < 18; ++u) {
      // tap (kh,kw) = (u%9/3, u%3): A start shifted by kh halo rows + kw voxels, exactly like the halo kernel
      const uint32_t a_off = MODE == 2 ? 0u : (uint32_t)(((u % 9) / 3) * halo_w + (u % 3)) * rowb;
      const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo0 + (a_off >> 4));
      const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo0 + (uint32_t)(((u % 3) * N * rowb) >> 4));
      if (leader) {
        if (MODE == 1) mma_ss(tmem + (uint32_t)((u % 3) * N), ad, bd, idesc, 1);
        else mma_ss(tmem, ad, bd, idesc, 1);
        if (MODE >= 1 && (u % 9) == 8) {
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2) : "memory");
          if (MODE == 2) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2 + 8) : "memory");
        }
      }
    }
  }
}

__global__ void rate(long long* out, int N, int iters, uint32_t layout, uint32_t rowb, uint32_t halo_w) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 2))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 3))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x < 32) {
    const bool leader = elect_one();
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
    const uint32_t a_hi = ((halo_w * rowb) >> 4) | (1u << 14) | (layout << 29), b_hi = ((8u * rowb) >> 4) | (1u << 14) | (layout << 29);
    const uint32_t a_lo0 = ((sa & 0x3FFFFu) >> 4) | 0x10000u, b_lo0 = ((sb & 0x3FFFFu) >> 4) | 0x10000u;
    uint32_t parity = 0;
#define RUN(M)                                                                                                   \
    {                                                                                                            \
      long long t0 = clock64();                                                                                  \
      body<M>(leader, tmem, a_lo0, b_lo0, a_hi, b_hi, idesc, N, iters, rowb, halo_w, smem_u32(bar + 2));                                          \
      if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); \
      __syncwarp();                                                                                              \
      uint32_t done = 0;                                                                                         \
      while (!done)                                                                                              \
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory"); \
      parity ^= 1;                                                                                               \
      if (leader) out[M] = clock64() - t0;                                                                       \
    }
    RUN(0) RUN(1) RUN(2)
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const char* names[3] = {"no commit", "3 acc + 1 commit / 9 MMAs", "2 commits / 9 MMAs"};
  struct Cfg { uint32_t layout, rowb, halo_w; const char* name; };
  const Cfg cfgs[] = {{6, 32, 10, "SW32 rows 32B SBO 320 (halo, 16 ch)"}, {2, 128, 10, "SW128 rows 128B SBO 1280 (halo, 64 ch)"}};
  for (const Cfg& c : cfgs) {
    printf("%s\n", c.name);
    for (int N : {16, 32, 48, 64, 96, 128}) {
      const int iters = 5400;   // multiple of 18
      rate<<<1, 128, 100 * 1024>>>(d, N, iters, c.layout, c.rowb, c.halo_w);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[3]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("  N=%3d:", N);
      for (int m = 0; m < 3; ++m) printf("  %s %.1f cyc/op", names[m], (double)h[m] / iters);
      printf("\n");
    }
  }
  return 0;
}
