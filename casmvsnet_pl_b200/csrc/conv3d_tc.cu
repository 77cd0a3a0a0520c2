// K2 (tensor-core variant) — 3x3x3 stride-1 convolution as an im2col-free implicit
// GEMM on tcgen05 (kind::tf32, fp32 accumulators in TMEM) with the norm-act (+skip)
// epilogue fused.  sm_100a only.
//
// Replaces (reference, relative to /root/reference): ConvBnReLU3D
// (models/modules.py:21-31) and the `prob` head (models/mvsnet.py:89,103) for the
// stride-1 layers of CostRegNet (conv0, conv2, conv4, prob: ~85 % of its FLOPs).
//
// GEMM view per CTA:  D[128 voxels x N] += A_tap[128 x Cin] * W_tap[Cin x N], 27 taps.
//   M = 128 output voxels = an 8(w) x 16(h) patch of one depth slice,
//   K = Cin per tap (Cin/8 MMAs of K=8),
//   N = 3 x GW: the three kd taps that share one A operand (input slice s shifted by
//   kh,kw) are issued as ONE MMA whose column groups are the accumulators of the
//   output slices s+1, s, s-1 (GW = Cout padded to 16).  With Cout <= 32 the MMA is
//   bound by the shared-memory read of A (4 KB per K=8 step), so tripling N per A read
//   is what takes the layer from 13x off the HBM roofline to ~1.3x.
// The CTA marches along depth.  Each input depth slice (with its 1-voxel halo,
// 18 x 10 voxels) is brought into shared memory ONCE by 4 producer warps
// (cp.async, zero-fill = the conv's zero padding) into a 4-slot ring and is used
// by the 27 taps of the three output slices it touches: the A operand of a tap is
// just a shifted *view* of the resident brick, expressed in the UMMA shared-memory
// descriptor (no-swizzle K-major layout: [h][Cin/4][w][4 floats]; 8 consecutive w
// form a core matrix, LBO = one channel-quad plane, SBO = one brick row).
// Nothing is re-read from L2 per tap, which is what makes the layer HBM-bound
// instead of L2-bound (a classic implicit GEMM would fetch every voxel 27 times).
//
// Warp roles (9 warps): 0-3 epilogue (TMEM -> regs -> scale/shift/LeakyReLU/+skip
// -> global), 4-7 producers, 8 MMA issuer (one elected thread).  Pipelines: smem
// ring full/empty mbarriers, double-buffered TMEM accumulator full/empty mbarriers.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace casmvs {

namespace tc {

struct Params {
  const float* x;       // (B,D,H,W,CIN)
  const float* wpk;     // [27][CIN][Cout]  (direct-kernel packing)
  const float* bimg;    // pre-built B operand image [kh][kw][CIN/4][3*GW][4] (tf32-rounded)
  const float* scale;   // [Cout] or null
  const float* shift;   // [Cout] or null
  const float* skip;    // (B,D,H,W,Cout) or null
  float* y;             // (B,D,H,W,Cout)
  float slope;
  int B, D, H, W, Cout;   // Cout = channels handled by one CTA (<= GW)
  int cout_total;         // channel count of the output tensor; blockIdx.y selects the chunk
  int tiles_w, tiles_h, nchunks, dchunk;
  int round_out;        // round the stored activations to tf32 (unbiased next-layer operand)
  long long* dbg;       // optional timeline of CTA 0: [role][slice][4] clock64 stamps
};
#define TC_STAMP(role, idx, k)                                                        \
  do {                                                                                \
    if (p.dbg && blockIdx.x == 0) p.dbg[((role) * 64 + (idx)) * 4 + (k)] = clock64(); \
  } while (0)

template <int CIN, int GW>
struct Smem {
  static constexpr int CQ = CIN / 4;
  // 64-channel (deep) layers: one CTA handles a 16-channel slice of Cout with a 2-slot ring
  static constexpr int SLOTS = CIN > 32 ? 2 : 4;
  static constexpr int kSlotBytes = kHaloH * CQ * kHaloW * 16;
  static constexpr int kWBytes = 9 * CIN * 3 * GW * 4;               // [kh][kw][cq][3*GW][4]
  static constexpr int kRingOff = kWBytes;
  static constexpr int kParamOff = kRingOff + SLOTS * kSlotBytes;   // scale/shift [2][GW]
  static constexpr int kBarOff = kParamOff + 2 * GW * 4;
  static constexpr int kTotal = kBarOff + 128 + 32 * 8;
};

// Largest power of two >= n (>= 32): TMEM allocations must be a power of two.
__host__ __device__ constexpr int tmem_cols_for(int n) {
  return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

template <int CIN, int GW>
__global__ void __launch_bounds__(kThreads, 1) conv3d_tc_kernel(const Params p) {
  using S = Smem<CIN, GW>;
  constexpr int CQ = S::CQ;
  constexpr int SLOTS = S::SLOTS;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_w = s_base, s_ring = s_base + S::kRingOff, s_bar = s_base + S::kBarOff;
  float* s_param = reinterpret_cast<float*>(smem + S::kParamOff);
  // barriers: full[4] @0, empty[4] @32, tmem ptr @64, tmem_full[32] @128.  Every output
  // slice has its OWN single-use accumulator-ready barrier: nothing throttles the MMA warp
  // against the epilogue (the accumulators are not recycled), so a recycled barrier could
  // be lapped twice and its parity would alias (seen as a hang with Cin = 8, 16+ slices).
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 32, bar_tfull = s_bar + 128;
  volatile uint32_t* s_tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + S::kBarOff + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work item ----
  int item = blockIdx.x;
  const int tw = item % p.tiles_w; item /= p.tiles_w;
  const int th = item % p.tiles_h; item /= p.tiles_h;
  const int ck = item % p.nchunks;
  const int b = item / p.nchunks;
  const int w0 = tw * kTileW, h0 = th * kTileH;
  const int d0 = ck * p.dchunk, d1 = min(p.D, d0 + p.dchunk);
  const int nd = d1 - d0;                         // <= 512 / GW (host guarantees)
  const int nslices = nd + 2;                     // input slices d0-1 .. d1
  const uint32_t tmem_cols = tmem_cols_for(p.dchunk * GW);

  // ---- one-time setup ----
  if (threadIdx.x == 0) TC_STAMP(3, 0, 0);
  init_barriers(bar_full, bar_empty, bar_tfull, SLOTS);
  if (warp == 0) tmem_alloc(smem_u32((const void*)s_tmem_ptr), tmem_cols);
  // B operand image (built once per launch by build_image_kernel) -> smem
  const int co_base = blockIdx.y * p.Cout;
  load_image_async(s_w, p.bimg + (size_t)blockIdx.y * (S::kWBytes / 4), S::kWBytes);
  for (int i = threadIdx.x; i < GW; i += kThreads) {
    s_param[i] = (i < p.Cout) ? (p.scale ? __ldg(p.scale + co_base + i) : 1.f) : 0.f;
    s_param[GW + i] = (i < p.Cout) ? (p.shift ? __ldg(p.shift + co_base + i) : 0.f) : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  // One accumulator (GW columns) per output slice of the chunk, laid out linearly, so
  // the three slices an input slice feeds are always adjacent columns.  They are zeroed
  // here once; every MMA then accumulates (the accumulate flag is per instruction, not
  // per column, so a first-touch overwrite is not expressible for one group of three).
  if (warp < 4) {
    for (int c = 0; c < nd * GW; c += 16)
      tmem_zero16(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) TC_STAMP(3, 0, 1);

  if (warp >= 4 && warp < kMmaWarp) {
    // ===================== producers: input bricks -> smem ring =====================
    const int ptid = threadIdx.x - 128;
    for (int it = 0; it < nslices; ++it) {
      const int s = d0 - 1 + it;
      const int slot = it % SLOTS;
      if (ptid == 0) TC_STAMP(0, it, 0);
      if (it >= SLOTS) mbar_wait(bar_empty + 8 * slot, ((it / SLOTS) - 1) & 1);
      if (ptid == 0) TC_STAMP(0, it, 1);
      const uint32_t dst0 = s_ring + slot * S::kSlotBytes;
      const bool s_ok = (s >= 0) && (s < p.D);
      const float* xs = p.x + (((size_t)b * p.D + (s_ok ? s : 0)) * p.H) * (size_t)p.W * CIN;
      for (int c = ptid; c < kHaloH * kHaloW * CQ; c += kProducerThreads) {
        const int cq = c % CQ;            // channel quad fastest: coalesced global reads
        const int vox = c / CQ;
        const int ww = vox % kHaloW, hh = vox / kHaloW;
        const int ih = h0 - 1 + hh, iw = w0 - 1 + ww;
        const bool ok = s_ok && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
        const float* src = ok ? xs + ((size_t)ih * p.W + iw) * CIN + cq * 4 : p.x;
        cp_async16(dst0 + ((hh * CQ + cq) * kHaloW + ww) * 16, src, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (ptid == 0) TC_STAMP(0, it, 2);
      if (it >= 1) {
        cp_async_wait<1>();               // slice it-1 has landed
        if (ptid == 0) TC_STAMP(0, it, 3);
        fence_proxy_async();              // generic-proxy writes -> visible to the MMA (async proxy)
        mbar_arrive(bar_full + 8 * ((it - 1) % SLOTS));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    mbar_arrive(bar_full + 8 * ((nslices - 1) % SLOTS));
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (single thread) =====================
    // The issuing thread is a scalar loop: anything computed per MMA costs ~5 cycles per
    // dependent instruction, and the tensor pipe retires a small MMA in ~45 cycles
    // (profiles/microbench/umma_rate.cu), so descriptors are base + compile-time offset.
    {
      constexpr uint32_t a_lbo = kHaloW * 16, a_sbo = CQ * kHaloW * 16;   // cq plane / brick row
      constexpr uint32_t b_lbo = 3 * GW * 16, b_sbo = 128;
      const uint32_t elected = elect_one();
      const uint64_t a_desc0 = make_desc(s_ring, a_lbo, a_sbo);
      const uint64_t b_desc0 = make_desc(s_w, b_lbo, b_sbo);
      for (int it = 0; it < nslices; ++it) {
        // input slice `it` feeds output slices j = it - kd, kd = 0,1,2, clipped to [0,nd):
        // columns [j_lo*GW, (j_hi+1)*GW), B rows [(2-kd_hi)*GW, (3-kd_lo)*GW)
        const int kd_lo = max(0, it - (nd - 1)), kd_hi = min(2, it);
        const int j_lo = it - kd_hi;
        const uint32_t idesc = make_idesc(128, (kd_hi - kd_lo + 1) * GW);
        const uint32_t acc = tmem_base + j_lo * GW;
        if (lane == 0) TC_STAMP(1, it, 0);
        mbar_wait(bar_full + 8 * (it % SLOTS), (it / SLOTS) & 1);
        if (lane == 0) TC_STAMP(1, it, 1);
        tc_fence_after();
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t a_lo0 = (uint32_t)a_desc0 + (((it % SLOTS) * S::kSlotBytes) >> 4);
        const uint32_t b_lo0 = (uint32_t)b_desc0 + (((2 - kd_hi) * GW * 16) >> 4);
#pragma unroll
        for (int khw = 0; khw < 9; ++khw) {
          const int kh = khw / 3, kw = khw % 3;
#pragma unroll
          for (int k8 = 0; k8 < CIN / 8; ++k8) {
            const uint32_t a_off = ((kh * CQ * kHaloW + kw) * 16 + k8 * 2 * kHaloW * 16) >> 4;
            const uint32_t b_off = (khw * (CIN * 3 * GW * 4) + k8 * 2 * 3 * GW * 16) >> 4;
            umma_tf32(acc, a_lo0 + a_off, a_hi, b_lo0 + b_off, b_hi, idesc, elected);
          }
        }
        if (it >= 2) umma_commit(bar_tfull + 8 * (it - 2), elected);         // slice it-2 complete
        umma_commit(bar_empty + 8 * (it % SLOTS), elected);                  // smem slot free
        if (lane == 0) TC_STAMP(1, it, 3);
      }
    }
  } else {
    // ===================== epilogue warps 0..3 =====================
    const int m = warp * 32 + lane;              // GEMM row = TMEM lane
    const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
    const bool in_range = oh < p.H && ow < p.W;
    for (int j = 0; j < nd; ++j) {
      if (threadIdx.x == 0) TC_STAMP(2, j, 0);
      mbar_wait(bar_tfull + 8 * j, 0);
      if (threadIdx.x == 0) TC_STAMP(2, j, 1);
      tc_fence_after();
      float acc[GW];
      tmem_ld<GW>(tmem_base + ((uint32_t)(warp * 32) << 16) + j * GW, acc);
      if (threadIdx.x == 0) TC_STAMP(2, j, 2);
      if (in_range) {
        const size_t o =
            ((((size_t)b * p.D + (d0 + j)) * p.H + oh) * p.W + ow) * p.cout_total + co_base;
        if (p.Cout % 4 == 0) {
#pragma unroll
          for (int c = 0; c < GW; c += 4) {
            if (c < p.Cout) {
              float v[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float t = fmaf(acc[c + k], s_param[c + k], s_param[GW + c + k]);
                v[k] = t >= 0.f ? t : t * p.slope;
              }
              if (p.skip) {
                const float4 s4 = ldg4(p.skip + o + c);
                v[0] += s4.x; v[1] += s4.y; v[2] += s4.z; v[3] += s4.w;
              }
              if (p.round_out) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = to_tf32(v[k]);
              }
              st4(p.y + o + c, make_float4(v[0], v[1], v[2], v[3]));
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < GW; ++c) {
            if (c < p.Cout) {
              float t = fmaf(acc[c], s_param[c], s_param[GW + c]);
              t = t >= 0.f ? t : t * p.slope;
              if (p.skip) t += __ldg(p.skip + o + c);
              p.y[o + c] = t;
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

// [kh][kw][cq][n = g*GW + co][4], g = 2 - kd; tf32-rounded (to nearest); zero rows for co >= Cout
// One image per Cout chunk (chunk = blockIdx.y of the conv kernel), `chunk` channels each.
__global__ void build_image_kernel(const float* __restrict__ wpk, float* __restrict__ img,
                                   int CIN, int GW, int chunk, int cout_total) {
  const int CQ = CIN / 4;
  const int per = 9 * CIN * 3 * GW;
  const int total = per * (cout_total / chunk);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int ck = t / per, i = t - ck * per;
    const int j = i & 3;
    const int n = (i >> 2) % (3 * GW);
    const int r = (i >> 2) / (3 * GW);    // (kh*3+kw)*CQ + cq
    const int cq = r % CQ, khw = r / CQ;
    const int g = n / GW, co = n % GW;
    const int kd = 2 - g;
    const int ci = cq * 4 + j;
    float v = 0.f;
    if (co < chunk)
      v = to_tf32(__ldg(wpk + ((size_t)(kd * 9 + khw) * CIN + ci) * cout_total + ck * chunk + co));
    img[t] = v;
  }
}

int build_stride1_image(const float* wpk, float* img, int CIN, int GW, int chunk, int cout_total,
                        cudaStream_t st) {
  build_image_kernel<<<64, 256, 0, st>>>(wpk, img, CIN, GW, chunk, cout_total);
  return after_launch("conv3d_tc/build_image");
}

float* image_scratch(size_t bytes) {
  constexpr int kRing = 8;
  constexpr size_t kSlot = 512 * 1024;
  static float* base = nullptr;
  static int next = 0;
  if (bytes > kSlot) return nullptr;
  if (!base) {
    if (cudaMalloc(&base, kRing * kSlot) != cudaSuccess) { base = nullptr; return nullptr; }
  }
  float* p = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + (size_t)next * kSlot);
  next = (next + 1) % kRing;
  return p;
}

struct CacheEntry { const void* key; int tag; size_t bytes; float* img; };
static CacheEntry g_cache[256];
static int g_cache_n = 0;

float* image_cache_lookup(const void* wpk, int tag, size_t bytes, bool* hit) {
  for (int i = 0; i < g_cache_n; ++i)
    if (g_cache[i].key == wpk && g_cache[i].tag == tag && g_cache[i].bytes == bytes) {
      *hit = true;
      return g_cache[i].img;
    }
  *hit = false;
  if (g_cache_n >= 256) return image_scratch(bytes);       // cache full: fall back to the ring
  float* img = nullptr;
  if (cudaMalloc(&img, bytes) != cudaSuccess) return nullptr;
  g_cache[g_cache_n++] = CacheEntry{wpk, tag, bytes, img};
  return img;
}

void image_cache_clear() {
  for (int i = 0; i < g_cache_n; ++i) cudaFree(g_cache[i].img);
  g_cache_n = 0;
}

template <int CIN, int GW>
static int launch(Params p, cudaStream_t st) {
  using S = Smem<CIN, GW>;
  auto kfn = conv3d_tc_kernel<CIN, GW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::kTotal);
    if (e != cudaSuccess) {
      set_error("conv3d_tc: cannot opt in to %d B of shared memory: %s", S::kTotal,
                cudaGetErrorString(e));
      return -2;
    }
    attr_set = true;
  }
  // depth chunk: one TMEM accumulator group (GW columns) per output slice, and as many
  // CTAs per SM as shared memory allows must be able to hold their TMEM at once
  const int per_sm = S::kTotal * 4 <= 227 * 1024 ? 4 : S::kTotal * 2 <= 227 * 1024 ? 2 : 1;
  const int cap = (512 / per_sm) / GW;
  int dchunk = p.D < cap ? p.D : cap;
  const long cols = (long)p.B * p.tiles_w * p.tiles_h;
  while (dchunk > 4 && cols * ((p.D + dchunk - 1) / dchunk) < (long)num_sms() * 3 * per_sm)
    dchunk = (dchunk + 1) / 2;
  p.dchunk = dchunk;
  p.nchunks = (p.D + dchunk - 1) / dchunk;
  const int nco = p.cout_total / p.Cout;
  bool hit = false;
  float* img = image_cache_lookup(p.wpk, 1000 + CIN * 100 + GW, (size_t)S::kWBytes * nco, &hit);
  if (!img) { set_error("conv3d_tc: cannot allocate the weight image"); return -2; }
  if (!hit) {
    if (int rc = build_stride1_image(p.wpk, img, CIN, GW, p.Cout, p.cout_total, st)) return rc;
  }
  p.bimg = img;
  const long items = (long)p.B * p.nchunks * p.tiles_h * p.tiles_w;
  static int extra = -1;
  if (extra < 0) {
    const char* e = getenv("CASMVS_TC_EXTRA_SMEM");
    extra = e ? atoi(e) : 0;
    if (extra > 0)
      cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal + extra);
  }
  kfn<<<dim3((unsigned)items, (unsigned)nco), kThreads, S::kTotal + extra, st>>>(p);
  return after_launch("conv3d_tc");
}

}  // namespace tc

// Returns 0 when handled, 1 when the layer shape is left to the CUDA-core kernel.
int conv3d_tc(const float* x, const float* wpk, const float* scale, const float* shift,
              float slope, const float* skip, float* y, int B, int Cin, int Cout, int D, int h,
              int w, int kind, int stride, int precision, cudaStream_t st) {
  static int enabled = -1, round_out = 1;
  static long long* dbg = nullptr;
  if (enabled < 0) {
    const char* e = getenv("CASMVS_TC");
    enabled = e ? atoi(e) : 1;
    if (const char* s = getenv("CASMVS_TC_ROUND")) round_out = atoi(s);
    if (const char* s = getenv("CASMVS_TC_DBG")) dbg = (long long*)strtoull(s, nullptr, 0);
  }
  if (!enabled || precision != CASMVS_TF32) return 1;
  if (kind != CASMVS_CONV || stride != 1) return 1;
  const bool deep = Cin == 64 && Cout == 64;          // conv6: 16-channel Cout slices
  if (!deep && (!(Cin == 8 || Cin == 16 || Cin == 32) || Cout > 32)) return 1;
  tc::Params p;
  p.x = x; p.wpk = wpk; p.scale = scale; p.shift = shift; p.skip = skip; p.y = y;
  p.slope = slope; p.B = B; p.D = D; p.H = h; p.W = w;
  p.Cout = deep ? 16 : Cout; p.cout_total = Cout;
  p.tiles_w = (w + tc::kTileW - 1) / tc::kTileW;
  p.tiles_h = (h + tc::kTileH - 1) / tc::kTileH;
  const int npad = p.Cout <= 16 ? 16 : 32;
  p.dbg = dbg;
  p.round_out = (round_out && Cout > 1) ? 1 : 0;   // the prob head feeds the softmax: keep fp32
#define TC_CASE(CI, NP) if (Cin == CI && npad == NP) return tc::launch<CI, NP>(p, st);
  TC_CASE(8, 16) TC_CASE(8, 32) TC_CASE(16, 16) TC_CASE(16, 32) TC_CASE(32, 16) TC_CASE(32, 32)
  TC_CASE(64, 16)
#undef TC_CASE
  return 1;
}

}  // namespace casmvs
