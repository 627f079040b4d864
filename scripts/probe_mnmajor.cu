// Probe: tcgen05.mma with an MN-major A operand read straight out of a voxel-major (NDHWC) shared-memory slab -- what the
// weight-gradient GEMM needs (K = voxels is the slab's row index, M = channels is contiguous inside a row).
//   * which of the descriptor's LBO / SBO fields strides the 8-row K groups and which strides the MN chunks
//     (a chunk = one swizzle row: 32 bf16 at 64B swizzle, 64 bf16 at 128B swizzle)?
//   * may the descriptor start at any slab row, and may the MN-chunk stride be a single row (64 / 128 bytes)?  Then ONE MMA with
//     M = 128 covers 4 (2) row-shifted copies of a 32 (64) channel slab: the kw = 0, 1, 2 filter taps of the weight gradient.
// B is a K-major 16x16 permutation (n -> k = (n + 1) % 16) at 32B swizzle (known-good from probe_umma), so D[m][n] = A[m][(n+1)%16].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_mnmajor probe_mnmajor.cu && ./probe_mnmajor
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

struct Case { int row_shift, chunk_rows, kgroup_rows, swapped; };   // swapped: LBO <- K-group stride, SBO <- chunk stride

__global__ void probe(const Case* cases, int ncases, float* out /* [ncases][128][16] */, int swz /*128 or 64*/) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int rowb = swz;                         // bytes per slab row (one voxel)
  const int slab_rows = 512;
  uint8_t* bsm = smem + slab_rows * rowb;       // B: 16x16 permutation, K-major, 32B rows, SW32
  uint64_t* bar = reinterpret_cast<uint64_t*>(bsm + 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int elems = rowb / 2;
  const int mask = swz == 128 ? 7 : 3;
  for (int i = threadIdx.x; i < slab_rows * elems; i += blockDim.x) {
    const int R = i / elems, c = i % elems;
    uint32_t off = R * rowb + c * 2;
    off ^= ((off >> 7) & mask) << 4;
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = __float2bfloat16((float)((R * 7 + c * 3) % 251));
  }
  for (int i = threadIdx.x; i < 16 * 16; i += blockDim.x) {
    const int n = i / 16, k = i % 16;
    uint32_t off = n * 32 + k * 2;
    off ^= ((off >> 7) & 1) << 4;
    *reinterpret_cast<__nv_bfloat16*>(bsm + off) = __float2bfloat16(k == (n + 1) % 16 ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  // D f32, A bf16, B bf16, A MN-major (bit 15), B K-major, N = 16, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t parity = 0;
  for (int c = 0; c < ncases; ++c) {
    const Case cs = cases[c];
    if (threadIdx.x == 0) {
      const uint32_t start = smem_u32(smem) + cs.row_shift * rowb;
      const uint32_t chunk = (uint32_t)cs.chunk_rows * rowb, kgroup = (uint32_t)cs.kgroup_rows * rowb;
      const uint64_t ad = cs.swapped ? make_desc(start, kgroup, chunk, swz == 128 ? 2 : 4) : make_desc(start, chunk, kgroup, swz == 128 ? 2 : 4);
      const uint64_t bd = make_desc(smem_u32(bsm), 16, 256, 6);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x < 128) {
      uint32_t v[16];
      const uint32_t taddr = tmem + (((uint32_t)(threadIdx.x / 32) * 32) << 16);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) out[((size_t)c * 128 + threadIdx.x) * 16 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  for (int swz : {128, 64}) {
    const int ce = swz / 2;                      // channels per MN chunk
    std::vector<Case> cases;
    for (int swapped : {0, 1})
      for (int chunk : {64, 16, 18, 2, 1})
        for (int kg : {8, 24})
          for (int shift : {0, 1, 3, 9}) cases.push_back({shift, chunk, kg, swapped});
    Case* dc;
    float* dout;
    cudaMalloc(&dc, cases.size() * sizeof(Case));
    cudaMalloc(&dout, cases.size() * 128 * 16 * sizeof(float));
    cudaMemcpy(dc, cases.data(), cases.size() * sizeof(Case), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    probe<<<1, 128, 80 * 1024, 0>>>(dc, (int)cases.size(), dout, swz);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> out(cases.size() * 128 * 16);
    cudaMemcpy(out.data(), dout, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    printf("== %dB swizzle, MN-major A: mismatches vs 'A[m][k] = slab[shift + (m / %d) * chunk_rows + (k / 8) * kgroup_rows + k %% 8][m %% %d]'\n", swz, ce, ce);
    for (size_t c = 0; c < cases.size(); ++c) {
      const Case cs = cases[c];
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
          const int k = (n + 1) % 16;
          const int R = cs.row_shift + (m / ce) * cs.chunk_rows + (k / 8) * cs.kgroup_rows + k % 8;
          const float want = (float)((R * 7 + (m % ce) * 3) % 251);
          if (out[(c * 128 + m) * 16 + n] != want) { if (first < 0) first = m; ++bad; }
        }
      printf("%s chunk stride %2d rows, K-group stride %2d rows, start row %d : %s (bad %d, first row %d)\n",
             cs.swapped ? "LBO=K-group SBO=chunk" : "LBO=chunk SBO=K-group", cs.chunk_rows, cs.kgroup_rows, cs.row_shift,
             bad ? "MISMATCH" : "ok", bad, first);
    }
    cudaFree(dc);
    cudaFree(dout);
  }
  return 0;
}
