// Semantics probe for CTA pairs: tcgen05.alloc / mma / commit with cta_group::2 (M = 256 over two SMs, each CTA holding its own
// 128 rows of A and HALF of the N rows of B), checked against a CPU GEMM, plus the MMA rate at N = 96 / 192.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)layout << 61;
  return d;
}

// A: [256 rows][64 k] (rank r holds rows 128 r ..), B: [N rows][64 k] (rank r holds rows N/2 r ..), D = A B^T: [256][N]
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, long long* cyc, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                       // 128 rows x 128 B, 128B swizzle
  uint8_t* sb = smem + 16384;               // N/2 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const uint32_t rank = cta_rank();
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    const uint32_t off = r * 128 + ((((k * 2) >> 4) ^ (r & 7)) << 4) + ((k * 2) & 15);
    *reinterpret_cast<__nv_bfloat16*>(sa + off) = A[(size_t)(rank * 128 + r) * 64 + k];
  }
  for (int i = threadIdx.x; i < (N / 2) * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    const uint32_t off = r * 128 + ((((k * 2) >> 4) ^ (r & 7)) << 4) + ((k * 2) & 15);
    *reinterpret_cast<__nv_bfloat16*>(sb + off) = B[(size_t)(rank * (N / 2) + r) * 64 + k];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                           // both CTAs' operands and barriers are in place
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 at bit 17, M>>4 at bit 24 with M = 256
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
  if (warp == 0 && rank == 0) {
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = make_desc(smem_u32(sa) + kk * 32, 1024, 2), bd = make_desc(smem_u32(sb) + kk * 32, 1024, 2);
        const uint32_t acc = (it | kk) != 0;
        if (leader)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    if (leader)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    __syncwarp();
    if (leader && cyc) *cyc = clock64() - t0;
  }
  // every CTA waits on its OWN barrier (the commit was multicast to both)
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
  if (warp == 0 && rank == 0 && cyc && elect_one()) cyc[1] = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 of its own CTA = rows rank*128 + 32 w + lane
  const int row = rank * 128 + warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

template <int N>
int run(int iters) {
  std::vector<__nv_bfloat16> hA(256 * 64), hB(N * 64);
  std::vector<float> fA(256 * 64), fB(N * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 17) - 8) / 8.f; hA[i] = __float2bfloat16(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 17) - 8) / 8.f; hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD; long long* dc;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 256 * N * 4); cudaMalloc(&dc, 16);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 256 * N * 4);
  cudaFuncSetAttribute(pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  pair_kernel<N><<<2, 128, 48 * 1024>>>(dA, dB, dD, dc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
  std::vector<float> hD(256 * N);
  long long hc[2];
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)fA[m * 64 + k] * fB[n * 64 + k];
      s *= iters;
      const double d = fabs(s - hD[m * N + n]);
      if (d > maxerr) maxerr = d;
    }
  printf("N=%3d iters=%d: max |D - ref| = %.3g  (%s)   issue+commit %lld cycles = %.1f cyc per 256xNx16 MMA\n", N, iters, maxerr,
         maxerr < 1e-3 * iters ? "OK" : "MISMATCH", hc[0], (double)hc[0] / (4.0 * iters));
  return maxerr < 1e-3 * iters ? 0 : 1;
}

int main() {
  int rc = 0;
  rc |= run<32>(1);
  rc |= run<96>(1);
  rc |= run<96>(2000);
  rc |= run<192>(2000);
  rc |= run<64>(2000);
  return rc;
}
