// TMA fetch rate as a function of the box shape (sm_100a).  Question behind it: the Cout <= 8
// convolution kernel reads channels-last bricks of 8-channel voxels, i.e. TMA boxes whose inner
// dimension is 32 bytes; is the L2 -> shared-memory rate a function of the inner box size?
// Every variant moves the same 6 x 32 x 8-float (6 KB) brick per request from an 84 MB volume
// (L2-flushed before each run), with the ring depth and CTA count of the real kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_box_rate tma_box_rate.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

constexpr int kSlots = 4, kBrick = 6 * 32 * 8 * 4;
// mode 0: {C, W, H, D} box {8,32,6,1}; 1: planar {W, H, D, C} box {32,6,1,8};
// 2: merged {W*C, H, D} box {256,6,1}; 3: planar, six requests of box {32,1,1,8} per brick
// (the [row][channel][column] order a MN-major UMMA operand needs)
__global__ void __launch_bounds__(32, 1)
fetch_kernel(const __grid_constant__ CUtensorMap map, int mode, int tiles_w, int tiles_h, int D, int items, unsigned long long* cycles) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kSlots * kBrick;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned long long t0 = clock64();
    uint32_t g = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int tw = item % tiles_w, th = (item / tiles_w) % tiles_h;
      for (int d = 0; d < D; ++d, ++g) {
        const int slot = g % kSlots;
        if (g >= kSlots) mbar_wait(bars + 8 * slot, ((g / kSlots) - 1) & 1);   // previous load of the slot landed
        mbar_expect_tx(bars + 8 * slot, kBrick);
        const uint32_t dst = base + slot * kBrick;
        if (mode == 0) tma_load_4d(dst, &map, bars + 8 * slot, 0, tw * 30 - 1, th * 4 - 1, d);
        else if (mode == 1) tma_load_4d(dst, &map, bars + 8 * slot, tw * 28 - 4, th * 4 - 1, d, 0);   // box start must be 16 B aligned
        else if (mode == 3) {
          for (int r = 0; r < 6; ++r) tma_load_4d(dst + r * 1024, &map, bars + 8 * slot, tw * 28 - 4, th * 4 - 1 + r, d, 0);
        } else tma_load_3d(dst, &map, bars + 8 * slot, (tw * 30 - 1) * 8, th * 4 - 1, d);
      }
    }
    for (uint32_t k = (g > kSlots ? g - kSlots : 0); k < g; ++k) mbar_wait(bars + 8 * (k % kSlots), (k / kSlots) & 1);
    if (cycles) cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int D = 8, H = 512, W = 640, C = 8;
  const size_t n = (size_t)D * H * W * C;
  float* x;
  CK(cudaMalloc(&x, n * 4));
  CK(cudaMemset(x, 0, n * 4));
  unsigned char* flush;
  CK(cudaMalloc(&flush, 256u << 20));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn encode = (EncodeFn)fn;
  int sms = 0, khz = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const int tiles_w = (W + 29) / 30, tiles_h = H / 4, items = tiles_w * tiles_h;
  const int smem = kSlots * kBrick + 64 + 1024;
  CK(cudaFuncSetAttribute(fetch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const char* names[4] = {"channels-last {C8,W,H,D} box {8,32,6,1} swizzle 32B (inner 32 B)",
                          "planar        {W,H,D,C} box {32,6,1,8} swizzle 128B (inner 128 B)",
                          "merged        {W*C,H,D} box {256,6,1} no swizzle   (inner 1024 B)",
                          "planar        {W,H,D,C} 6 x box {32,1,1,8} swizzle 128B (inner 128 B)"};
  for (int mode = 0; mode < 4; ++mode) {
    CUtensorMap map;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    if (mode == 0) {
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D};
      cuuint64_t str[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
      cuuint32_t box[4] = {8, 32, 6, 1};
      r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (mode == 1 || mode == 3) {
      cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C};
      cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4};
      cuuint32_t box[4] = {32, (cuuint32_t)(mode == 1 ? 6 : 1), 1, 8};
      r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[3] = {(cuuint64_t)W * C, (cuuint64_t)H, (cuuint64_t)D};
      cuuint64_t str[2] = {(cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
      cuuint32_t box[3] = {256, 6, 1};
      r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode failed for mode %d: %d\n", mode, (int)r); continue; }
    for (int per_sm = 1; per_sm <= 8; per_sm *= 2) {
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemsetAsync(flush, rep, 256u << 20));
        CK(cudaEventRecord(e0));
        fetch_kernel<<<sms * per_sm, 32, smem>>>(map, mode, tiles_w, tiles_h, D, items, nullptr);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      const double bytes = (double)items * D * kBrick;
      printf("%-70s CTAs/SM %d  %7.1f us  %6.2f TB/s fetched  %5.1f B/clk/SM (nominal %d MHz)\n", names[mode], per_sm,
             best * 1e3, bytes / best / 1e9, bytes / (best * 1e-3) / sms / (khz * 1e3), khz / 1000);
    }
  }
  return 0;
}
