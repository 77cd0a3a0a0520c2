// K2 (tensor-core variant, TMA producer, narrow outputs) — stride-1 3x3x3 convolution with
// Cout <= 8: conv0 of every stage and the `prob` head, the layers that carry most of
// CostRegNet's bytes.  sm_100a only.
//
// Replaces (reference, relative to /root/reference): ConvBnReLU3D conv0
// (models/mvsnet.py:63, models/modules.py:21-31) and the prob head (models/mvsnet.py:89,103).
//
// profiles/microbench/umma_rate.cu: a tf32 M=128 MMA costs 44.6 cycles for any N <= ~89, and
// with the TMA producer conv3d_tma.cu is bound by exactly that (9*Cin/8 MMAs of N = 48 per
// input slice keep the tensor pipe 100 % busy).  Here BOTH the kd and the kw taps are folded
// into N, which cuts the MMA count per input slice to 3*Cin/8 (N = 80):
//   * GEMM rows: M = 128 = 4 image rows x 32 brick columns.  The TMA brick of a slice is
//     [6 rows][32 columns][CB channels] (voxel-major, swizzled by the row size), so the sixteen
//     8-voxel row groups of the A operand are evenly spaced (SBO = 8 voxels) and the kh tap is a
//     start-address shift of one brick row;
//   * GEMM columns: 9 groups of 8 = (kd, kw) x Cout: D[m][(kd,kw,co)] = sum_{kh,ci}
//     x[s][h+kh-1][i][ci] * W[kd][kh][kw][ci][co] for brick column i;
//   * the kw taps are recombined in the epilogue: output column c of the tile needs
//     D[c][kw=0] + D[c+1][kw=1] + D[c+2][kw=2], i.e. a shift of 1 and 2 TMEM lanes INSIDE one
//     warp (a warp owns one image row of 32 brick columns): two __shfl_down per value, no
//     shared memory and no cross-warp barrier.  30 of the 32 brick columns produce outputs.
// Everything else is as in conv3d_tma.cu: persistent CTAs, 4 x SETS epilogue warps + TMA
// producer + MMA issuer (SETS = 2 at Cin = 8: the two sets drain alternate output slices),
// linear TMEM accumulators (24 columns per output slice) that every MMA accumulates into and
// the epilogue re-zeroes after reading, full/empty ring barriers and tfull/tempty accumulator
// barriers.  What bounds it at the full-resolution Cin = 8 layers is the TMA brick-fetch rate
// times the halo amplification (profiles/r2_tma_box_rate.txt, profiles/r2_k2_n8_stalls.txt).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace casmvs {
namespace tma8 {

using namespace casmvs::tc;
using casmvs::tma::mbar_expect_tx;
using casmvs::tma::tma_load_5d;

// warps: SETS x 4 epilogue warps (set k drains the output slices j with j % SETS == k; warp w of
// any set reads TMEM lanes 32*(w % 4)..), then the TMA producer warp, then the MMA issuer
constexpr int kRowsOut = 4;     // image rows per tile (= epilogue warps)
constexpr int kBH = 6;          // brick rows (with halo)
constexpr int kBW = 32;         // brick columns (with halo) = lanes of a warp
constexpr int kColsOut = 30;    // output columns per tile
constexpr int kG = 24;          // accumulator columns per output slice: 3 kw x 8 channels
constexpr int kBRows = 80;      // B rows per kh: 3 kd x 24 + 8 zero rows (N must be % 16)

struct Params {
  const float* bimg;   // [kh][CIN/4][80][4], tf32-rounded
  const float* scale;  // [Cout] or null
  const float* shift;  // [Cout] or null
  const float* skip;   // (B,D,H,W,Cout) or null
  float* y;            // (B,D,H,W,Cout)
  float slope;
  int B, D, H, W, Cout;
  int tiles_w, tiles_h, nchunks, dchunk;
  int round_out;
  int planar;          // 1x3x3 kernel: input slice s feeds output slice s only (kd = 1)
  FastDiv fd_tw, fd_th, fd_ck;   // item -> (tile column, tile row, depth chunk, batch)
  long long* dbg;
};
// per-role clock64 timeline of CTA 0 (profiles/tc_timeline.py): compiled in only with
// `make TIMELINE=1` -- each stamp costs ~6 instructions in loops whose roles are bound by
// their own scalar instruction stream (profiles/r2_k2_n8_stalls.txt)
#ifdef CASMVS_TIMELINE
#define N8_STAMP(role, idx, k)                                                                  \
  do {                                                                                          \
    if (p.dbg && blockIdx.x == 0 && (idx) < 64) p.dbg[((role) * 64 + (idx)) * 4 + (k)] = clock64(); \
  } while (0)
#else
#define N8_STAMP(role, idx, k) do { } while (0)
#endif

template <int CIN, int SLOTS_>
struct Smem {
  static constexpr int SLOTS = SLOTS_;
  static constexpr int ROWB = CIN * 4;                              // bytes per voxel = swizzle span
  static constexpr int kSlotBytes = kBH * kBW * ROWB;               // 6 KB x CIN/8: 1024-aligned
  static constexpr int kWBytes = 3 * CIN * kBRows * 4;              // [kh][cq][80][4]
  static constexpr int kRingOff = 0;
  static constexpr int kWOff = SLOTS * kSlotBytes;
  static constexpr int kParamOff = kWOff + kWBytes;                 // scale/shift [2][8]
  static constexpr int kBarOff = kParamOff + 2 * 8 * 4;
  // barriers: full[8] @0, empty[8] @64, tmem ptr @128, tfull[32] @192, tempty[32] @448
  static constexpr int kTotal = kBarOff + 192 + 32 * 8 + 32 * 8 + 1024;   // + alignment slack
  static constexpr uint32_t kLayout = ROWB == 128 ? 2u : ROWB == 64 ? 4u : 6u;
};

__host__ __device__ constexpr int tmem_cols_for8(int n) {
  return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

// the three kw groups of one output slice (columns +0, +8, +16) in one statement: one address
// register, one wait
__device__ __forceinline__ void tmem_ld3x8(uint32_t taddr, float (&a)[8], float (&b)[8],
                                           float (&c)[8]) {
  uint32_t r[24];
  asm volatile(
      "{\n\t.reg .b32 t1, t2;\n\t"
      "add.u32 t1, %24, 8;\n\t"
      "add.u32 t2, %24, 16;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [t1];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [t2];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t}"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(r[8 + i]);
    c[i] = __uint_as_float(r[16 + i]);
  }
}
__device__ __forceinline__ void tmem_zero3x8(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t.reg .b32 t1, t2;\n\t"
      "add.u32 t1, %0, 8;\n\t"
      "add.u32 t2, %0, 16;\n\t"
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n\t"
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [t1], {%1,%1,%1,%1,%1,%1,%1,%1};\n\t"
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [t2], {%1,%1,%1,%1,%1,%1,%1,%1};\n\t"
      "tcgen05.wait::st.sync.aligned;\n\t}"
      ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_ld3x1(uint32_t taddr, float& a, float& b, float& c) {
  uint32_t r0, r1, r2;
  asm volatile(
      "{\n\t.reg .b32 t1, t2;\n\t"
      "add.u32 t1, %3, 8;\n\t"
      "add.u32 t2, %3, 16;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%3];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%1}, [t1];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%2}, [t2];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t}"
      : "=r"(r0), "=r"(r1), "=r"(r2)
      : "r"(taddr)
      : "memory");
  a = __uint_as_float(r0); b = __uint_as_float(r1); c = __uint_as_float(r2);
}
__device__ __forceinline__ void tmem_zero3x1(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t.reg .b32 t1, t2;\n\t"
      "add.u32 t1, %0, 8;\n\t"
      "add.u32 t2, %0, 16;\n\t"
      "tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n\t"
      "tcgen05.st.sync.aligned.32x32b.x1.b32 [t1], {%1};\n\t"
      "tcgen05.st.sync.aligned.32x32b.x1.b32 [t2], {%1};\n\t"
      "tcgen05.wait::st.sync.aligned;\n\t}"
      ::"r"(taddr), "r"(z)
      : "memory");
}
// round to tf32, to nearest with ties away (== cvt.rna.tf32.f32 for every finite input and inf;
// the compiler expands the cvt into this plus an inf/NaN guard)
__device__ __forceinline__ float round_tf32_bits(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ void tmem_zero8(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};"
               ::"r"(taddr), "r"(z)
               : "memory");
}

// CO: output channels known at compile time (8: conv0 and the planar FeatureNet layers, 1: the
// prob head), 0 = any Cout <= 8 (per-channel stores).  The specialised epilogues are 120 (Cout 8:
// packed add / fma pairs, 2-instruction tf32 rounding) and 50 (Cout 1: one TMEM column per kw
// group) instructions per output slice instead of 160.  On their own they did not change the step
// time -- the Cin = 8 layers are bound by the TMA brick-fetch rate, profiles/r2_k2_n8_stalls.txt --
// but they are what made a second set of epilogue warps (SETS = 2) pay: the drain -> re-issue
// round trip of an accumulator group shortens, 3.4 % of the cfg2 step.
template <int CIN, int SLOTS_, int SETS, int CO>
// two sets at Cin = 8: 320 threads; four resident CTAs need <= 51 registers (48 used, no spills)
__global__ void __launch_bounds__((4 * SETS + 2) * 32, (SETS == 2 && CIN == 8) ? 4 : 1)
conv3d_tma_n8_kernel(const __grid_constant__ CUtensorMap xmap, const Params p) {
  using S = Smem<CIN, SLOTS_>;
  constexpr int SLOTS = S::SLOTS;
  constexpr int kThreads8 = (4 * SETS + 2) * 32;
  constexpr int kProdWarp = 4 * SETS, kIssueWarp = 4 * SETS + 1;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t s_raw = smem_u32(smem_raw);
  const uint32_t s_base = (s_raw + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (s_base - s_raw);
  const uint32_t s_ring = s_base + S::kRingOff, s_w = s_base + S::kWOff,
                 s_bar = s_base + S::kBarOff;
  float* s_param = reinterpret_cast<float*>(smem + S::kParamOff);
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 64, bar_tfull = s_bar + 192,
                 bar_tempty = s_bar + 448, bar_w = s_bar + 136;   // bar_w: weight image landed
  volatile uint32_t* s_tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + S::kBarOff + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // dchunk groups + 8 spill columns (N is rounded up to a multiple of 16: the extra 8 columns of
  // a 24- or 72-column MMA land on the next group's first 8 columns)
  const uint32_t tmem_cols = tmem_cols_for8(p.dchunk * kG + 8);
  const int total_items = p.B * p.nchunks * p.tiles_h * p.tiles_w;

  // ---- one-time setup ----
  if (threadIdx.x == 0) N8_STAMP(3, 0, 0);
  {
    const int t = threadIdx.x;
    if (t < SLOTS) mbar_init(bar_full + 8 * t, 1);
    else if (t < 2 * SLOTS) mbar_init(bar_empty + 8 * (t - SLOTS), 1);
    else if (t >= 32 && t < 64) mbar_init(bar_tfull + 8 * (t - 32), 1);
    else if (t >= 64 && t < 96) mbar_init(bar_tempty + 8 * (t - 64), 128);
    else if (t == 96) mbar_init(bar_w, 1);
    if (t < 97) fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32((const void*)s_tmem_ptr), tmem_cols);
  for (int i = threadIdx.x; i < 8; i += kThreads8) {
    s_param[i] = (i < p.Cout) ? (p.scale ? __ldg(p.scale + i) : 1.f) : 0.f;
    s_param[8 + i] = (i < p.Cout) ? (p.shift ? __ldg(p.shift + i) : 0.f) : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  if (threadIdx.x == 0) tma::load_image_bulk(s_w, p.bimg, S::kWBytes, bar_w);
  if (warp < 4) {
    for (int c = 0; c < p.dchunk * kG + 8; c += 8)
      tmem_zero8(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) N8_STAMP(3, 0, 1);
  // nothing above depends on the previous kernel of the stream (see tma_common.cuh)
  tma::pdl_trigger();
  tma::pdl_wait();
  bool w_ready = false;                             // MMA issuer: weight image has landed
  uint32_t gs = 0;                                  // slices processed before this item (all roles)
  int ep = 0;                                       // items processed by this CTA
  for (int item0 = blockIdx.x; item0 < total_items; item0 += gridDim.x, ++ep) {
    uint32_t utw, uth, uck;
    const int b = (int)fastdivmod(
        fastdivmod(fastdivmod((uint32_t)item0, p.fd_tw, utw), p.fd_th, uth), p.fd_ck, uck);
    const int tw = (int)utw, th = (int)uth, ck = (int)uck;
    const int w0 = tw * kColsOut, h0 = th * kRowsOut;
    const int d0 = ck * p.dchunk, d1 = min(p.D, d0 + p.dchunk);
    const int nd = d1 - d0;
    const int halo = p.planar ? 0 : 1;
    const int nslices = nd + 2 * halo;              // input slices d0-halo .. d1-1+halo

    if (warp == kProdWarp) {
      // ===================== producer: one TMA load per slice =====================
      if (lane == 0) {
        for (int it = 0; it < nslices; ++it) {
          const uint32_t g = gs + it;
          const int slot = g % SLOTS;
          N8_STAMP(0, g, 0);
          if (g >= (uint32_t)SLOTS) mbar_wait(bar_empty + 8 * slot, ((g / SLOTS) - 1) & 1);
          N8_STAMP(0, g, 1);
          mbar_expect_tx(bar_full + 8 * slot, S::kSlotBytes);
          tma_load_5d(s_ring + slot * S::kSlotBytes, &xmap, bar_full + 8 * slot, 0, w0 - 1,
                      h0 - 1, d0 - halo + it, b);
          N8_STAMP(0, g, 2);
        }
      }
      __syncwarp();
    } else if (warp == kIssueWarp) {
      // ===================== MMA issuer =====================
      constexpr uint32_t a_lbo = 16, a_sbo = 8 * S::ROWB;               // 8-voxel group stride
      constexpr uint32_t b_lbo = kBRows * 16, b_sbo = 128;
      const uint32_t elected = elect_one();
      const uint64_t a_desc0 = make_desc(s_ring, a_lbo, a_sbo) | ((uint64_t)S::kLayout << 61);
      const uint64_t b_desc0 = make_desc(s_w, b_lbo, b_sbo);
      const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
      int waited = 0;                               // groups whose tempty has been consumed
      for (int it = 0; it < nslices; ++it) {
        const uint32_t g = gs + it;
        // input slice `it` feeds output slices j = it - kd, kd = 0,1,2, clipped to [0,nd):
        // columns [j_lo*24, (j_hi+1)*24) (+8 pad columns when the count is odd),
        // B rows [(2-kd_hi)*24, ...)
        const int kd_lo = p.planar ? 1 : max(0, it - (nd - 1));
        const int kd_hi = p.planar ? 1 : min(2, it);
        const int cnt = kd_hi - kd_lo + 1;
        const int j_lo = p.planar ? it : it - kd_hi;
        const uint32_t idesc = make_idesc(128, cnt == 1 ? 32 : cnt == 2 ? 48 : 80);
        const uint32_t acc = tmem_base + j_lo * kG;
        if (lane == 0) N8_STAMP(1, g, 0);
        mbar_wait(bar_full + 8 * (g % SLOTS), (g / SLOTS) & 1);
        if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }
        if (lane == 0) N8_STAMP(1, g, 1);
        // Every group this slice touches -- including the pad columns on group it+1 -- must
        // have been drained and re-zeroed by the epilogue of the previous item.
        if (ep > 0) {
          const int need = min(it + 1, p.dchunk - 1);
          for (; waited <= need; ++waited) mbar_wait(bar_tempty + 8 * waited, (ep - 1) & 1);
        }
        if (lane == 0) N8_STAMP(1, g, 2);
        tc_fence_after();
        const uint32_t a_lo0 = (uint32_t)a_desc0 + (((g % SLOTS) * S::kSlotBytes) >> 4);
        const uint32_t b_lo0 = (uint32_t)b_desc0 + (((2 - kd_hi) * kG * 16) >> 4);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int k8 = 0; k8 < CIN / 8; ++k8) {
            const uint32_t a_off = (kh * kBW * S::ROWB + k8 * 32) >> 4;
            const uint32_t b_off = (kh * (CIN * kBRows * 4) + k8 * 2 * kBRows * 16) >> 4;
            umma_tf32(acc, a_lo0 + a_off, a_hi, b_lo0 + b_off, b_hi, idesc, elected);
          }
        }
        if (p.planar) umma_commit(bar_tfull + 8 * it, elected);        // slice it complete
        else if (it >= 2) umma_commit(bar_tfull + 8 * (it - 2), elected);   // slice it-2 complete
        umma_commit(bar_empty + 8 * (g % SLOTS), elected);             // smem slot free
        if (lane == 0) N8_STAMP(1, g, 3);
      }
      // groups this (short) chunk did not use go through the same handshake (empty -> full) so
      // that every barrier sees exactly one completion per item; the commit (not a plain
      // arrive) orders the epilogue's clean-up of the pad columns after the last MMA
      for (int j = nd; j < p.dchunk; ++j) {
        if (ep > 0) {
          for (; waited <= j; ++waited) mbar_wait(bar_tempty + 8 * waited, (ep - 1) & 1);
        }
        umma_commit(bar_tfull + 8 * j, elected);
      }
    } else {
      // ============ epilogue warps: one image row each, set (warp / 4) of SETS ============
      const int q = warp & 3, set = warp >> 2;
      const int oh = h0 + q, ow = w0 + lane;
      const bool writes = lane < kColsOut && oh < p.H && ow < p.W;
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
      // this lane's voxel in output slice d0; one slice further = slice_stride floats
      const size_t o0 = ((((size_t)b * p.D + d0) * p.H + oh) * p.W + ow) * p.Cout;
      const size_t slice_stride = (size_t)p.H * p.W * p.Cout;
      size_t o = o0 + (size_t)set * slice_stride;
      for (int j = set; j < p.dchunk; j += SETS, o += SETS * slice_stride) {
        if (threadIdx.x == 0) N8_STAMP(2, ep * p.dchunk + j, 0);
        mbar_wait(bar_tfull + 8 * j, ep & 1);
        if (threadIdx.x == 0) N8_STAMP(2, ep * p.dchunk + j, 1);
        tc_fence_after();
        const uint32_t tg = lane_base + j * kG;
        if (j >= nd) {
          // unused group: only its first 8 columns can have been written (pad columns of the
          // last slice's MMA, non-zero weights) -- clean them for the next item
          if (j == nd) {
            tmem_zero8(tg);
            tmem_wait_st();
            tc_fence_before();
          }
          mbar_arrive(bar_tempty + 8 * j);
          continue;
        }
        if constexpr (CO == 1) {
          // channel 0 of the three kw groups; every other column of the group only ever
          // accumulates x * 0 (zero rows of the operand image) and is never read
          float a0, a1, a2;
          tmem_ld3x1(tg, a0, a1, a2);
          tmem_zero3x1(tg);
          tc_fence_before();
          mbar_arrive(bar_tempty + 8 * j);         // group j drained and zero again
          if (threadIdx.x == 0) N8_STAMP(2, ep * p.dchunk + j, 2);
          const float s1 = __shfl_down_sync(0xffffffffu, a1, 1);
          const float s2 = __shfl_down_sync(0xffffffffu, a2, 2);
          float t = fmaf(a0 + s1 + s2, s_param[0], s_param[8]);
          t = t >= 0.f ? t : t * p.slope;
          if (writes) {
            if (p.skip) t += __ldg(p.skip + o);
            p.y[o] = p.round_out ? round_tf32_bits(t) : t;
          }
        } else {
          float a0[8], a1[8], a2[8];
          tmem_ld3x8(tg, a0, a1, a2);
          tmem_zero3x8(tg);
          tc_fence_before();
          mbar_arrive(bar_tempty + 8 * j);         // group j drained and zero again
          if (threadIdx.x == 0) N8_STAMP(2, ep * p.dchunk + j, 2);
          // out[c] = D_kw0[c] + D_kw1[c+1] + D_kw2[c+2]   (c = brick column = lane), as packed
          // pairs of channels (add.f32x2 / fma.f32x2: the same IEEE operations, half the issues)
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            const float s1a = __shfl_down_sync(0xffffffffu, a1[k], 1);
            const float s1b = __shfl_down_sync(0xffffffffu, a1[k + 1], 1);
            const float s2a = __shfl_down_sync(0xffffffffu, a2[k], 2);
            const float s2b = __shfl_down_sync(0xffffffffu, a2[k + 1], 2);
            const unsigned long long t2 =
                fma2(add2(add2(pk2(a0[k], a0[k + 1]), pk2(s1a, s1b)), pk2(s2a, s2b)),
                     pk2(s_param[k], s_param[k + 1]), pk2(s_param[8 + k], s_param[9 + k]));
            float ta, tb;
            unpk2(t2, ta, tb);
            v[k] = ta >= 0.f ? ta : ta * p.slope;
            v[k + 1] = tb >= 0.f ? tb : tb * p.slope;
          }
          if (writes) {
            if constexpr (CO == 8) {
              if (p.skip) {
                const float4 s0 = ldg4(p.skip + o), s1 = ldg4(p.skip + o + 4);
                v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
                v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
              }
              if (p.round_out) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = round_tf32_bits(v[k]);
              }
              st4(p.y + o, make_float4(v[0], v[1], v[2], v[3]));
              st4(p.y + o + 4, make_float4(v[4], v[5], v[6], v[7]));
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                if (k < p.Cout) {
                  float tt = v[k];
                  if (p.skip) tt += __ldg(p.skip + o + k);
                  p.y[o + k] = p.round_out ? round_tf32_bits(tt) : tt;
                }
              }
            }
          }
        }
      }
    }
    gs += nslices;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// [kh][cq][row = g*24 + kw*8 + co][4], g = 2 - kd; rows 72..79 zero; tf32-rounded
__global__ void build_image_n8_kernel(const float* __restrict__ wpk, float* __restrict__ img,
                                      int CIN, int Cout, int planar) {
  const int CQ = CIN / 4;
  const int total = 3 * CIN * kBRows;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int jq = i & 3;
    const int row = (i >> 2) % kBRows;
    const int r = (i >> 2) / kBRows;       // kh*CQ + cq
    const int cq = r % CQ, kh = r / CQ;
    const int ci = cq * 4 + jq;
    float v = 0.f;
    if (row < 72) {
      const int g = row / kG, kw = (row % kG) / 8, co = row % 8;
      const int kd = 2 - g;
      // planar: only the centre plane (the pad columns of its N = 32 MMA fall on the next
      // output slice's accumulator, so the rows after group 1 MUST be zero)
      if (co < Cout && (!planar || kd == 1))
        v = to_tf32(__ldg(wpk + ((size_t)((kd * 3 + kh) * 3 + kw) * CIN + ci) * Cout + co));
    }
    img[i] = v;
  }
}

template <int CIN, int SLOTS, int SETS, int CO>
static int launch8co(const float* x, const float* wpk, Params p, cudaStream_t st) {
  using S = Smem<CIN, SLOTS>;
  constexpr int kThreads8 = (4 * SETS + 2) * 32;
  auto kfn = conv3d_tma_n8_kernel<CIN, SLOTS, SETS, CO>;
  static std::atomic<bool> attr_set[kMaxDevices];
  if (int rc = opt_in_smem(kfn, S::kTotal, attr_set, "conv3d_tma_n8")) return rc;
  const CUtensorMap* map = tma::input_map(x, p.B, p.D, p.H, p.W, CIN, CIN, kBW, kBH);
  if (!map) return -2;
  static int per_sm_env = -1, dchunk_env = -1;
  if (per_sm_env < 0) {
    const char* e = getenv("CASMVS_N8_PER_SM");
    per_sm_env = e ? atoi(e) : 0;
    const char* d = getenv("CASMVS_N8_DCHUNK");
    dchunk_env = d ? atoi(d) : 0;
  }
  // resident CTAs per SM: as many as shared memory allows (1 KB per CTA is reserved by the
  // system), at most 4.  Measured on cfg2 (profiles/r1_n8_per_sm.txt): more, shorter-chunk CTAs
  // beat fewer, longer ones for Cin = 8 and 16 (the epilogue's store latency is what has to be
  // hidden); Cin = 32 fits two.
  // resident CTAs by shared memory (1 KB per CTA is reserved by the system) and registers
  static int reg_limit = 0;
  if (!reg_limit) {
    cudaFuncAttributes fa;
    reg_limit = 4;
    if (cudaFuncGetAttributes(&fa, kfn) == cudaSuccess && fa.numRegs > 0)
      reg_limit = 65536 / (((fa.numRegs + 7) / 8 * 8) * kThreads8);
    if (reg_limit < 1) reg_limit = 1;
  }
  int smem_limit = (228 * 1024) / (S::kTotal + 1024);
  if (smem_limit > reg_limit) smem_limit = reg_limit;
  int per_sm = smem_limit < 4 ? smem_limit : 4;
  if (per_sm_env > 0 && per_sm_env < per_sm) per_sm = per_sm_env;
  if (per_sm < 1) per_sm = 1;
  int cap = (tma::pow2_floor(512 / per_sm) - 8) / kG;
  if (cap > 21) cap = 21;
  if (dchunk_env > 0 && dchunk_env < cap) cap = dchunk_env;
  p.tiles_w = (p.W + kColsOut - 1) / kColsOut;
  p.tiles_h = (p.H + kRowsOut - 1) / kRowsOut;
  const long cols = (long)p.B * p.tiles_w * p.tiles_h;
  const int dchunk =
      tma::pick_dchunk(p.D, cap, cols, (long)num_sms() * per_sm, 1, p.planar ? 0 : 2);
  p.dchunk = dchunk;
  p.nchunks = (p.D + dchunk - 1) / dchunk;
  const ImageRef ir = image_cache_get(wpk, 8000 + CIN + (p.planar ? 500 : 0), (size_t)S::kWBytes, st);
  float* img = ir.img;
  if (!img) return -2;
  if (!ir.hit) {
    build_image_n8_kernel<<<32, 256, 0, st>>>(wpk, img, CIN, p.Cout, p.planar);
    if (int rc = after_launch("conv3d_tma_n8/build_image")) return rc;
    image_cache_built(img, st);
  }
  p.bimg = img;
  const long items = (long)p.B * p.nchunks * p.tiles_h * p.tiles_w;
  const uint32_t dmax = (uint32_t)std::max(p.tiles_w, std::max(p.tiles_h, p.nchunks));
  if (!fastdiv_ok((uint64_t)items, dmax)) return 1;          // left to the other kernels
  p.fd_tw = make_fastdiv((uint32_t)p.tiles_w);
  p.fd_th = make_fastdiv((uint32_t)p.tiles_h);
  p.fd_ck = make_fastdiv((uint32_t)p.nchunks);
  const long resident = (long)num_sms() * per_sm;
  const long gx = items < resident ? items : resident;
  tma::launch_pdl(ir.settled, kfn, dim3((unsigned)gx), kThreads8, S::kTotal, st, *map, p);
  return after_launch("conv3d_tma_n8");
}

template <int CIN, int SLOTS, int SETS>
static int launch8(const float* x, const float* wpk, const Params& p, cudaStream_t st) {
  if (p.Cout == 8) return launch8co<CIN, SLOTS, SETS, 8>(x, wpk, p, st);
  if (p.Cout == 1) return launch8co<CIN, SLOTS, SETS, 1>(x, wpk, p, st);
  return launch8co<CIN, SLOTS, SETS, 0>(x, wpk, p, st);
}

}  // namespace tma8

// Returns 0 when handled, 1 when the layer shape is left to the other kernels.
int conv3d_tma_n8(const float* x, const float* wpk, const float* scale, const float* shift,
                  float slope, const float* skip, float* y, int B, int Cin, int Cout, int D,
                  int h, int w, int kind, int stride, int precision_flags, cudaStream_t st) {
  const int precision = precision_flags & 0xff;
  static int enabled = -1, round_out = 1;
  static long long* dbg = nullptr;
  if (enabled < 0) {
    const char* e = getenv("CASMVS_N8");
    enabled = e ? atoi(e) : 1;
    if (const char* s = getenv("CASMVS_TC_ROUND")) round_out = atoi(s);
    if (const char* s = getenv("CASMVS_TC_DBG")) dbg = (long long*)strtoull(s, nullptr, 0);
  }
  if (!enabled || precision != CASMVS_TF32) return 1;
  if ((kind != CASMVS_CONV && kind != CASMVS_CONV_PLANAR) || stride != 1) return 1;
  if (!(Cin == 8 || Cin == 16 || Cin == 32) || Cout > 8) return 1;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return 1;
  tma8::Params p;
  p.scale = scale; p.shift = shift; p.skip = skip; p.y = y;
  p.slope = slope; p.B = B; p.D = D; p.H = h; p.W = w; p.Cout = Cout;
  p.dbg = dbg;
  p.planar = kind == CASMVS_CONV_PLANAR ? 1 : 0;
  // the prob head feeds the softmax: keep fp32; callers can ask for unrounded outputs
  p.round_out = (round_out && Cout > 1 && !(precision_flags & CASMVS_KEEP_FP32_OUT)) ? 1 : 0;
  // second set of epilogue warps (alternate output slices): bit 0 -> Cin = 8, bit 1 -> Cin = 16,
  // bit 2 -> Cin = 32 (CASMVS_N8_SETS2_MASK; measured per bit in profiles/r2_switch_ab.txt)
  static int sets2 = -1;
  if (sets2 < 0) {
    const char* e = getenv("CASMVS_N8_SETS2_MASK");
    sets2 = e ? atoi(e) : 1;     // cfg2 step: 1.0957 ms with 0, 1.0589 with 1, 1.0580 with 3, 1.0620 with 7 (three sets at Cin = 8: 1.1111)
  }
  if (Cin == 8 && (sets2 & 1)) return tma8::launch8<8, 4, 2>(x, wpk, p, st);
  if (Cin == 16 && (sets2 & 2)) return tma8::launch8<16, 4, 2>(x, wpk, p, st);
  if (Cin == 32 && (sets2 & 4)) return tma8::launch8<32, 3, 2>(x, wpk, p, st);
  static int slots8 = -1;
  if (slots8 < 0) {
    const char* e = getenv("CASMVS_N8_SLOTS8");       // ring depth of the Cin = 8 kernels (4 or 6)
    slots8 = e ? atoi(e) : 4;
  }
  if (Cin == 8 && slots8 == 6) return tma8::launch8<8, 6, 1>(x, wpk, p, st);
  if (Cin == 8) return tma8::launch8<8, 4, 1>(x, wpk, p, st);
  if (Cin == 16) return tma8::launch8<16, 4, 1>(x, wpk, p, st);
  return tma8::launch8<32, 3, 1>(x, wpk, p, st);
}

}  // namespace casmvs
