// PTX wrappers shared by the tcgen05 convolution kernels (sm_100a only): mbarrier,
// cp.async, tcgen05 alloc/fence/commit/mma/ld/st, UMMA descriptors.
#pragma once
#include "common.cuh"

namespace casmvs {
namespace tc {

constexpr int kTileW = 8, kTileH = 16;           // M = 128
constexpr int kHaloW = kTileW + 2, kHaloH = kTileH + 2;
constexpr int kSlots = 4;
// warps 0-3 epilogue (TMEM lane quadrants), kProducerWarps producer warps, last warp MMA issuer
constexpr int kProducerWarps = 4;   // 8 was slower: 416 threads x 64-80 regs leave one CTA per SM
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kMmaWarp = 4 + kProducerWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar, uint32_t elected) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar), "r"(elected)
      : "memory");
}
// One lane of a converged warp (the MMA warp runs warp-uniform code so that the
// descriptors stay in uniform registers; only the issue itself is predicated).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}
// D[tmem] += A[smem desc] * B[smem desc], tf32 operands, fp32 accumulate; issued by the
// lane whose `elected` is non-zero
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                          uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t elected) {
  // descriptors travel as (lo,hi) words: only the low word (start address) changes per tap
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "setp.ne.b32 p, %6, 0xffffffff;\n\t"          // always true: accumulate
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(elected)
      : "memory");
}
// shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor):
// [0,14) start>>4 | [16,30) leading byte offset>>4 | [32,46) stride byte offset>>4 |
// [46,48) version = 1 (Blackwell) | [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4,
// a/b format TF32 (2) @7/@10, a/b major K (0), N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <int NPAD>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[NPAD]) {
  static_assert(NPAD == 16 || NPAD == 32, "NPAD");
  uint32_t r[NPAD];
  if constexpr (NPAD == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  } else {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < NPAD; ++i) v[i] = __uint_as_float(r[i]);
}

// zero 16 consecutive fp32 columns of this warp's 32 TMEM lanes
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}


// Cooperative smem fill of the pre-built B operand image (global, 16 B aligned, bytes % 16 == 0)
__device__ __forceinline__ void load_image_async(uint32_t dst_smem, const float* src, int bytes) {
  for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16)
    cp_async16(dst_smem + i, reinterpret_cast<const char*>(src) + i, 16u);
  cp_async_commit();
  cp_async_wait<0>();
}
// barrier init spread over the first threads of the CTA (one mbarrier each)
__device__ __forceinline__ void init_barriers(uint32_t bar_full, uint32_t bar_empty,
                                              uint32_t bar_tfull, int slots) {
  const int t = threadIdx.x;
  if (t < slots) mbar_init(bar_full + 8 * t, kProducerThreads);
  else if (t < 2 * slots) mbar_init(bar_empty + 8 * (t - slots), 1);
  else if (t < 2 * slots + 32) mbar_init(bar_tfull + 8 * (t - 2 * slots), 1);
  if (t < 2 * slots + 32) fence_barrier_init();
}

// Persistent cache of built B-operand images keyed by (packed-weight pointer, kernel tag,
// size) -- weight_image.cu.  `hit`: the image exists (no build needed; if it was built on a
// different stream this stream has been ordered after the build).  `settled`: the build is
// known to have completed, so a programmatic-dependent-launch prologue may read the image.
// On a miss the caller launches its builder kernel on `st` and then calls image_cache_built.
struct ImageRef { float* img; bool hit; bool settled; };
ImageRef image_cache_get(const void* wpk, int tag, size_t bytes, cudaStream_t st);
void image_cache_built(const float* img, cudaStream_t st);
// B operand image of the stride-1 kernels: [chunk][kh][kw][CIN/4][3*GW][4], tf32-rounded,
// column group g holds the weights of kd = 2 - g (weight_image.cu)
int build_stride1_image(const float* wpk, float* img, int CIN, int GW, int chunk, int cout_total,
                        cudaStream_t st);

}  // namespace tc
}  // namespace casmvs
