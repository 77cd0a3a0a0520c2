// Experiment: can a tcgen05 A operand be a SHIFTED VIEW into a voxel-major, 128B-swizzled brick
// ([18 rows][10 voxels][32 ch] = what one TMA tiled load with SWIZZLE_128B would write), i.e.
// start address not 1024-aligned and SBO (10 voxels = 1280 B) not a multiple of the swizzle
// atom?  The brick is written with the swizzle computed from ABSOLUTE smem address bits
// (chunk ^= (addr >> 7) & 7), the MMA is issued for several (kh,kw) shifts with and without the
// descriptor's base_offset field, and D is compared with a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_swizzle_view umma_swizzle_view.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
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
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

#ifndef CCH
#define CCH 32
#endif
constexpr int BR = 18, BW = 10, C = CCH, N = 16;
constexpr int ROWB = C * 4;                       // bytes per voxel row: 128 / 64 / 32
constexpr int SWMASK = ROWB / 16 - 1;             // 7 / 3 / 1: Swizzle<B,4,3> xors addr bits [7,7+B) into [4,4+B)
constexpr int LAYOUT = ROWB == 128 ? 2 : ROWB == 64 ? 4 : 6;

// mode bit0: use base_offset = (start >> 7) & 7 ; shift given by kh,kw
__global__ void __launch_bounds__(128, 1) k(const float* brick, const float* wts, float* out, int kh, int kw, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  float* sA = reinterpret_cast<float*>(smem);                 // 18*10*128 B = 23040 B
  float* sB = reinterpret_cast<float*>(smem + 24576);         // no-swizzle [cq][n][4]
  const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // brick: voxel v = hh*10+ww at byte offset v*128, 16B chunk j stored at j ^ ((abs_addr >> 7) & 7)
  for (int i = threadIdx.x; i < BR * BW * (ROWB / 16); i += 128) {
    const int v = i / (ROWB / 16), j = i % (ROWB / 16);
    const uint32_t addr = a_base + v * ROWB + j * 16;          // logical byte address of the chunk
    const uint32_t paddr = addr ^ (((addr >> 7) & SWMASK) << 4); // swizzled (absolute address bits)
    const float4 val = *reinterpret_cast<const float4*>(brick + v * C + j * 4);
    *reinterpret_cast<float4*>(smem + (paddr - a_base)) = val;
  }
  // B: [cq][n][4] no swizzle: LBO = N*16, SBO = 128
  for (int i = threadIdx.x; i < C * N; i += 128) {
    const int jq = i & 3, n = (i >> 2) % N, cq = (i >> 2) / N;
    sB[i] = wts[n * C + cq * 4 + jq];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  if (threadIdx.x == 0) {
    const uint32_t id = idesc_tf32(128, N);
    const uint32_t a_start = a_base + (kh * BW + kw) * ROWB;
    const uint32_t boff = (mode & 1) ? ((a_start >> 7) & 7) : 0;
    for (int k8 = 0; k8 < C / 8; ++k8) {
      const uint64_t ad = desc(a_start + k8 * 32, 16, BW * ROWB, LAYOUT, boff);
      const uint64_t bd = desc(b_base + k8 * 2 * N * 16, N * 16, 128, 0, 0);
      umma(tm, ad, bd, id, k8 > 0);
    }
    commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r[16];
    const uint32_t taddr = tm + ((uint32_t)((threadIdx.x / 32) * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm) : "memory");
}

int main() {
  float *hb = new float[BR * BW * C], *hw = new float[N * C], *ho = new float[128 * 16];
  srand(1);
  for (int i = 0; i < BR * BW * C; ++i) hb[i] = (float)(rand() % 17 - 8) / 8.f;   // exact in tf32
  for (int i = 0; i < N * C; ++i) hw[i] = (float)(rand() % 9 - 4) / 4.f;
  float *db, *dw, *dout;
  cudaMalloc(&db, BR * BW * C * 4); cudaMalloc(&dw, N * C * 4); cudaMalloc(&dout, 128 * 16 * 4);
  cudaMemcpy(db, hb, BR * BW * C * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw, N * C * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("C=%d rowbytes=%d layout_type=%d\n", C, ROWB, LAYOUT);
  for (int mode = 0; mode < 1; ++mode)
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        k<<<1, 128, 64 * 1024>>>(db, dw, dout, kh, kw, mode);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ho, dout, 128 * 16 * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < 128; ++m) {
          const int hh = m / 8 + kh, ww = m % 8 + kw;
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int c = 0; c < C; ++c) ref += (double)hb[(hh * BW + ww) * C + c] * hw[n * C + c];
            const double err = fabs(ref - ho[m * 16 + n]);
            if (err > maxerr) maxerr = err;
            if (err > 1e-3) ++bad;
          }
        }
        printf("base_offset=%s kh=%d kw=%d: max|err| %.4g, %d/2048 wrong %s\n", mode ? "set" : "0  ", kh, kw, maxerr, bad, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
