// K2 (tensor-core variant, part 3) — stride-1 3x3x3 convolution with Cout <= 8 (conv0 of every
// stage and the `prob` head: the layers that carry most of CostRegNet's bytes) with BOTH the
// kd and the kh taps folded into the MMA's N dimension.
//
// profiles/microbench/umma_rate.cu: a tf32 M=128 MMA costs 44.6 cycles for any N <= ~89, so
// these small-Cout layers are bound by the NUMBER of MMAs.  conv3d_tc.cu issues 9*Cin/8 per
// input slice (kd folded, N = 48).  Here the A operand is the resident brick shifted in w
// only (kw); the three kh taps become three 8-column accumulator groups whose rows are
// recombined with a row shift in the epilogue:
//     out[t] = acc_kh0[t] + acc_kh1[t+1] + acc_kh2[t+2]        (t = output row in the tile)
// => 3*Cin/8 MMAs of N = 80 (9 groups x 8 + 8 zero-weight pad columns) per input slice,
// 14 valid output rows per 16-row tile.  The row shift crosses warps, so the epilogue
// exchanges the kh=1,2 partials through shared memory (double buffered, one named barrier
// per slice).  Everything else (producers, ring, linear TMEM accumulators zeroed once,
// single-use barriers, cached B image) is as in conv3d_tc.cu.
//
// Replaces (reference, relative to /root/reference): ConvBnReLU3D conv0
// (models/mvsnet.py:63, models/modules.py:21-31) and the prob head (models/mvsnet.py:89,103).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace casmvs {
namespace tc3 {

using namespace casmvs::tc;

constexpr int kRows = 16;       // brick rows = GEMM row groups
constexpr int kOutRows = 14;    // valid output rows per tile
constexpr int kBW = 10;         // brick column positions (8 + halo)
constexpr int kG = 24;          // accumulator columns per output slice: 3 kh x 8 channels
constexpr int kBRows = 80;      // B rows per kw: 3 kd x 24 + 8 zero rows (N must be % 16)

struct Params {
  const float* x;      // (B,D,H,W,CIN)
  const float* bimg;   // [kw][CIN/4][80][4], tf32-rounded
  const float* scale;  // [Cout] or null
  const float* shift;  // [Cout] or null
  const float* skip;   // (B,D,H,W,Cout) or null
  float* y;            // (B,D,H,W,Cout)
  float slope;
  int B, D, H, W, Cout;
  int tiles_w, tiles_h, nchunks, dchunk;
  int round_out;
};

template <int CIN>
struct Smem {
  static constexpr int CQ = CIN / 4;
  static constexpr int kSlotBytes = kRows * CQ * kBW * 16;
  static constexpr int kWBytes = 3 * CIN * kBRows * 4;
  static constexpr int kRingOff = kWBytes;
  static constexpr int kXchgOff = kRingOff + kSlots * kSlotBytes;   // [2 buffers][2 kh][128][8]
  static constexpr int kParamOff = kXchgOff + 2 * 2 * 128 * 8 * 4;
  static constexpr int kBarOff = kParamOff + 2 * 8 * 4;
  static constexpr int kTotal = kBarOff + 128 + 32 * 8;
};

__host__ __device__ constexpr int tmem_cols_for3(int n) {
  return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
                 "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

template <int CIN>
__global__ void __launch_bounds__(kThreads, 1) conv3d_tc3_kernel(const Params p) {
  using S = Smem<CIN>;
  constexpr int CQ = S::CQ;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_w = s_base, s_ring = s_base + S::kRingOff, s_bar = s_base + S::kBarOff;
  float* s_xchg = reinterpret_cast<float*>(smem + S::kXchgOff);
  float* s_param = reinterpret_cast<float*>(smem + S::kParamOff);
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 32, bar_tfull = s_bar + 128;
  volatile uint32_t* s_tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + S::kBarOff + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int item = blockIdx.x;
  const int tw = item % p.tiles_w; item /= p.tiles_w;
  const int th = item % p.tiles_h; item /= p.tiles_h;
  const int ck = item % p.nchunks;
  const int b = item / p.nchunks;
  const int w0 = tw * kTileW, h0 = th * kOutRows;
  const int d0 = ck * p.dchunk, d1 = min(p.D, d0 + p.dchunk);
  const int nd = d1 - d0;
  const int nslices = nd + 2;
  const uint32_t tmem_cols = tmem_cols_for3(p.dchunk * kG + 32);

  init_barriers(bar_full, bar_empty, bar_tfull, kSlots);
  if (warp == 0) tmem_alloc(smem_u32((const void*)s_tmem_ptr), tmem_cols);
  load_image_async(s_w, p.bimg, S::kWBytes);
  for (int i = threadIdx.x; i < 8; i += kThreads) {
    s_param[i] = (i < p.Cout) ? (p.scale ? __ldg(p.scale + i) : 1.f) : 0.f;
    s_param[8 + i] = (i < p.Cout) ? (p.shift ? __ldg(p.shift + i) : 0.f) : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  if (warp < 4) {
    // zero the chunk's accumulators and the 32 spill columns after them
    for (int c = 0; c < nd * kG + 32; c += 16) {
      if (c + 16 <= (int)tmem_cols) tmem_zero16(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= 4 && warp < kMmaWarp) {
    // ===================== producers: 16 x 10 voxel bricks =====================
    const int ptid = threadIdx.x - 128;
    for (int it = 0; it < nslices; ++it) {
      const int s = d0 - 1 + it;
      const int slot = it & (kSlots - 1);
      if (it >= kSlots) mbar_wait(bar_empty + 8 * slot, ((it >> 2) - 1) & 1);
      const uint32_t dst0 = s_ring + slot * S::kSlotBytes;
      const bool s_ok = (s >= 0) && (s < p.D);
      const float* xs = p.x + (((size_t)b * p.D + (s_ok ? s : 0)) * p.H) * (size_t)p.W * CIN;
      for (int c = ptid; c < kRows * kBW * CQ; c += kProducerThreads) {
        const int cq = c % CQ;
        const int vox = c / CQ;
        const int ww = vox % kBW, hh = vox / kBW;
        const int ih = h0 - 1 + hh, iw = w0 - 1 + ww;
        const bool ok = s_ok && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
        const float* src = ok ? xs + ((size_t)ih * p.W + iw) * CIN + cq * 4 : p.x;
        cp_async16(dst0 + ((hh * CQ + cq) * kBW + ww) * 16, src, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (it >= 1) {
        cp_async_wait<1>();
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * ((it - 1) & (kSlots - 1)));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    mbar_arrive(bar_full + 8 * ((nslices - 1) & (kSlots - 1)));
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    constexpr uint32_t a_lbo = kBW * 16, a_sbo = CQ * kBW * 16;
    constexpr uint32_t b_lbo = kBRows * 16, b_sbo = 128;
    const uint32_t elected = elect_one();
    const uint64_t a_desc0 = make_desc(s_ring, a_lbo, a_sbo);
    const uint64_t b_desc0 = make_desc(s_w, b_lbo, b_sbo);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
    for (int it = 0; it < nslices; ++it) {
      const int kd_lo = max(0, it - (nd - 1)), kd_hi = min(2, it);
      const int cnt = kd_hi - kd_lo + 1;
      const int j_lo = it - kd_hi;
      // N must be a multiple of 16: 24 -> 32 and 72 -> 80; the extra columns either meet the
      // zero pad rows or land in accumulator columns that are never read (phantom slice)
      const uint32_t idesc = make_idesc(128, cnt == 1 ? 32 : cnt == 2 ? 48 : 80);
      const uint32_t acc = tmem_base + j_lo * kG;
      mbar_wait(bar_full + 8 * (it & (kSlots - 1)), (it >> 2) & 1);
      tc_fence_after();
      const uint32_t a_lo0 = (uint32_t)a_desc0 + (((it & (kSlots - 1)) * S::kSlotBytes) >> 4);
      const uint32_t b_lo0 = (uint32_t)b_desc0 + (((2 - kd_hi) * kG * 16) >> 4);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
        for (int k8 = 0; k8 < CIN / 8; ++k8) {
          const uint32_t a_off = (kw * 16 + k8 * 2 * kBW * 16) >> 4;
          const uint32_t b_off = (kw * (CIN * kBRows * 4) + k8 * 2 * kBRows * 16) >> 4;
          umma_tf32(acc, a_lo0 + a_off, a_hi, b_lo0 + b_off, b_hi, idesc, elected);
        }
      }
      if (it >= 2) umma_commit(bar_tfull + 8 * (it - 2), elected);
      umma_commit(bar_empty + 8 * (it & (kSlots - 1)), elected);
    }
  } else {
    // ===================== epilogue warps 0..3 =====================
    const int m = warp * 32 + lane;                  // GEMM row = TMEM lane = (brick row, w)
    const int t = m >> 3;                            // output row in the tile == brick row of kh=0
    const int oh = h0 + t, ow = w0 + (m & 7);
    const bool writes = t < kOutRows && oh < p.H && ow < p.W;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int j = 0; j < nd; ++j) {
      mbar_wait(bar_tfull + 8 * j, 0);
      tc_fence_after();
      float a0[8], a1[8], a2[8];
      tmem_ld8(lane_base + j * kG, a0);
      tmem_ld8(lane_base + j * kG + 8, a1);
      tmem_ld8(lane_base + j * kG + 16, a2);
      // row-shifted recombination through shared memory: out[t] = a0[t] + a1[t+1] + a2[t+2]
      float* xb = s_xchg + (j & 1) * (2 * 128 * 8);
      *reinterpret_cast<float4*>(xb + m * 8) = make_float4(a1[0], a1[1], a1[2], a1[3]);
      *reinterpret_cast<float4*>(xb + m * 8 + 4) = make_float4(a1[4], a1[5], a1[6], a1[7]);
      *reinterpret_cast<float4*>(xb + 1024 + m * 8) = make_float4(a2[0], a2[1], a2[2], a2[3]);
      *reinterpret_cast<float4*>(xb + 1024 + m * 8 + 4) = make_float4(a2[4], a2[5], a2[6], a2[7]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (writes) {
        const float4 b0 = *reinterpret_cast<const float4*>(xb + (m + 8) * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(xb + (m + 8) * 8 + 4);
        const float4 c0 = *reinterpret_cast<const float4*>(xb + 1024 + (m + 16) * 8);
        const float4 c1 = *reinterpret_cast<const float4*>(xb + 1024 + (m + 16) * 8 + 4);
        float v[8] = {a0[0] + b0.x + c0.x, a0[1] + b0.y + c0.y, a0[2] + b0.z + c0.z,
                      a0[3] + b0.w + c0.w, a0[4] + b1.x + c1.x, a0[5] + b1.y + c1.y,
                      a0[6] + b1.z + c1.z, a0[7] + b1.w + c1.w};
        const size_t o = ((((size_t)b * p.D + (d0 + j)) * p.H + oh) * p.W + ow) * p.Cout;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float tt = fmaf(v[k], s_param[k], s_param[8 + k]);
          v[k] = tt >= 0.f ? tt : tt * p.slope;
        }
        if (p.Cout == 8) {
          if (p.skip) {
            const float4 s0 = ldg4(p.skip + o), s1 = ldg4(p.skip + o + 4);
            v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
            v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
          }
          if (p.round_out) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = to_tf32(v[k]);
          }
          st4(p.y + o, make_float4(v[0], v[1], v[2], v[3]));
          st4(p.y + o + 4, make_float4(v[4], v[5], v[6], v[7]));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k < p.Cout) {
              float tt = v[k];
              if (p.skip) tt += __ldg(p.skip + o + k);
              p.y[o + k] = p.round_out ? to_tf32(tt) : tt;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// [kw][cq][row = g*24 + kh*8 + co][4], g = 2 - kd; rows 72..79 zero; tf32-rounded
__global__ void build_image3_kernel(const float* __restrict__ wpk, float* __restrict__ img,
                                    int CIN, int Cout) {
  const int CQ = CIN / 4;
  const int total = 3 * CIN * kBRows;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int jq = i & 3;
    const int row = (i >> 2) % kBRows;
    const int r = (i >> 2) / kBRows;       // kw*CQ + cq
    const int cq = r % CQ, kw = r / CQ;
    const int ci = cq * 4 + jq;
    float v = 0.f;
    if (row < 72) {
      const int g = row / kG, kh = (row % kG) / 8, co = row % 8;
      const int kd = 2 - g;
      if (co < Cout)
        v = to_tf32(__ldg(wpk + ((size_t)((kd * 3 + kh) * 3 + kw) * CIN + ci) * Cout + co));
    }
    img[i] = v;
  }
}

template <int CIN>
static int launch3(Params p, const float* wpk, cudaStream_t st) {
  using S = Smem<CIN>;
  auto kfn = conv3d_tc3_kernel<CIN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::kTotal);
    if (e != cudaSuccess) {
      set_error("conv3d_tc3: cannot opt in to %d B of shared memory: %s", S::kTotal,
                cudaGetErrorString(e));
      return -2;
    }
    attr_set = true;
  }
  p.tiles_w = (p.W + kTileW - 1) / kTileW;
  p.tiles_h = (p.H + kOutRows - 1) / kOutRows;
  // TMEM: dchunk*24 + 32 columns per CTA, rounded to a power of two; keep as many CTAs
  // resident as shared memory allows
  // (2 CTAs/SM at most: 4 would leave 4-slice chunks whose 2-slice halo costs 50 % extra work)
  const int per_sm = S::kTotal * 2 <= 227 * 1024 ? 2 : 1;
  const int cap = per_sm == 1 ? 20 : 9;
  int dchunk = p.D < cap ? p.D : cap;
  const long cols = (long)p.B * p.tiles_w * p.tiles_h;
  while (dchunk > 4 && cols * ((p.D + dchunk - 1) / dchunk) < (long)num_sms() * 3 * per_sm)
    dchunk = (dchunk + 1) / 2;
  p.dchunk = dchunk;
  p.nchunks = (p.D + dchunk - 1) / dchunk;
  bool hit = false;
  float* img = image_cache_lookup(wpk, 3000 + CIN, (size_t)S::kWBytes, &hit);
  if (!img) { set_error("conv3d_tc3: cannot allocate the weight image"); return -2; }
  if (!hit) {
    build_image3_kernel<<<64, 256, 0, st>>>(wpk, img, CIN, p.Cout);
    if (int rc = after_launch("conv3d_tc3/build_image")) return rc;
  }
  p.bimg = img;
  const long items = cols * p.nchunks;
  kfn<<<(unsigned)items, kThreads, S::kTotal, st>>>(p);
  return after_launch("conv3d_tc3");
}

}  // namespace tc3

// Returns 0 when handled, 1 when the layer is left to conv3d_tc / the CUDA-core kernel.
int conv3d_tc3(const float* x, const float* wpk, const float* scale, const float* shift,
               float slope, const float* skip, float* y, int B, int Cin, int Cout, int D, int h,
               int w, int kind, int stride, int precision, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) {
    // Off by default: measured 2-4 % SLOWER end to end than conv3d_tc.cu on cfg2 (round 1).
    // The MMA count drops 3x but the tiles lose 2 of 16 rows, the epilogue gains a shared-
    // memory exchange + named barrier per slice, and TMEM (24 columns per slice + spill)
    // limits a CTA to 9 slices at 2 CTAs/SM, so per-CTA setup and pipeline fill dominate.
    // It becomes interesting once CTAs are persistent.  CASMVS_TC3=1 enables it.
    const char* e = getenv("CASMVS_TC3");
    enabled = e ? atoi(e) : 0;
  }
  if (!enabled || precision != CASMVS_TF32) return 1;
  if (kind != CASMVS_CONV || stride != 1 || Cout > 8) return 1;
  if (!(Cin == 8 || Cin == 16 || Cin == 32)) return 1;
  tc3::Params p;
  p.x = x; p.scale = scale; p.shift = shift; p.skip = skip; p.y = y; p.slope = slope;
  p.B = B; p.D = D; p.H = h; p.W = w; p.Cout = Cout;
  p.round_out = Cout > 1 ? 1 : 0;                  // the prob head feeds the softmax: keep fp32
  if (Cin == 8) return tc3::launch3<8>(p, wpk, st);
  if (Cin == 16) return tc3::launch3<16>(p, wpk, st);
  return tc3::launch3<32>(p, wpk, st);
}

}  // namespace casmvs
