// Micro-benchmark: cycles per tcgen05.mma (kind::tf32, M=128, K=8) as a function of N,
// shared-memory layout (no swizzle vs 128B swizzle) and accumulator dependency.
// One CTA, operands are whatever is in smem (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

struct Cfg { int N, layout, a_lbo, a_sbo, b_lbo, b_sbo, nacc, count, a_step, kind16; };

__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t sb = smem_u32(smem);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 49152; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  if (threadIdx.x == 0) {
    const uint32_t id = idesc_tf32(128, c.N);
    const uint32_t a0 = sb, b0 = sb + 96 * 1024;
    // descriptors are precomputed; the timed loop is 8 MMAs of straight-line code
    uint64_t ad[8];
    for (int i = 0; i < 8; ++i) ad[i] = desc(a0 + i * c.a_step, c.a_lbo, c.a_sbo, c.layout);
    const uint64_t bd = desc(b0, c.b_lbo, c.b_sbo, c.layout);
    uint32_t dcol[8];
    for (int i = 0; i < 8; ++i) dcol[i] = tm + (i % c.nacc) * c.N;
    for (int i = 0; i < 8; ++i) umma(dcol[i], ad[i], bd, id, 0u);
    long long t0 = clock64();
    for (int it = 0; it < c.count / 8; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) umma(dcol[i], ad[i], bd, id, 1u);
    }
    long long t1 = clock64();
    commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%-44s %6s %10s %10s\n", "config", "N", "issue/mma", "total/mma");
  const int Ns[] = {16, 32, 48, 96, 128, 256};
  for (int layout : {0, 2}) {            // 0 = no swizzle, 2 = 128B swizzle
    for (int nacc : {1, 2, 4}) {
      for (int N : Ns) {
        if (nacc * N > 512) continue;
        Cfg c;
        c.N = N; c.layout = layout; c.nacc = nacc; c.count = 256; c.kind16 = 0;
        if (layout == 0) { c.a_lbo = 160; c.a_sbo = 8 * 160; c.b_lbo = N * 16; c.b_sbo = 128; c.a_step = 16; }
        else { c.a_lbo = 16; c.a_sbo = 1024; c.b_lbo = 16; c.b_sbo = 1024; c.a_step = 32; }
        k<<<1, 128, 200 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        char name[64];
        snprintf(name, sizeof name, "%s nacc=%d", layout == 0 ? "no-swizzle(LBO160,SBO1280)" : "swizzle128B(SBO1024)", nacc);
        printf("%-44s %6d %10.1f %10.1f %s\n", name, N, (double)h[0] / c.count, (double)h[1] / c.count, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  }
  // no-swizzle with a dense A (rows contiguous groups: LBO=128*16? SBO=128) to see if strides matter
  for (int N : {16, 48, 96}) {
    Cfg c; c.N = N; c.layout = 0; c.nacc = 1; c.count = 256; c.kind16 = 0;
    c.a_lbo = 2048; c.a_sbo = 128; c.b_lbo = N * 16; c.b_sbo = 128; c.a_step = 16;
    k<<<1, 128, 200 * 1024>>>(c, d);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-44s %6d %10.1f %10.1f\n", "no-swizzle dense(LBO2048,SBO128) nacc=1", N, (double)h[0] / c.count, (double)h[1] / c.count);
  }
  return 0;
}
