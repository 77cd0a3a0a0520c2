// K2 (tensor-core variant, part 2) — the stride-2 convolutions (conv1, conv3) and the
// transposed convolutions (conv9, conv11) of CostRegNet on tcgen05, with the same
// "resident brick + shifted-view UMMA descriptors" scheme as conv3d_tc.cu.
//
// Replaces (reference, relative to /root/reference):
//   ConvBnReLU3D(stride=2)                         models/modules.py:21-31, mvsnet.py:65,68
//   ConvTranspose3d(k3,s2,p1,op1) + norm_act + skip models/mvsnet.py:79-87,99-101
//
// MODE_S2 (stride 2): M = 8(w) x 16(h) OUTPUT voxels of one output slice od.  Input row
//   ih = 2*oh + kh - 1 => consecutive GEMM row groups are two brick rows apart (SBO = 2
//   rows); input column iw = 2*ow + kw - 1 => the brick stores its 17 columns
//   de-interleaved (8 even | 9 odd) so that 8 consecutive ow are again 16 B apart.
//   Odd input slices (s = 2a+1) feed outputs a (kd=2) and a+1 (kd=0) in ONE MMA of
//   N = 2*GW; even slices feed output a (kd=1).
// MODE_T (transposed, output = 2x input): M = 8 x 16 INPUT voxels j of one input slice.
//   Output voxel o = 2j + p (p in {0,1}^3, 8 parity classes); class p reads input j + s
//   with tap k:  p=0 -> (s=0,k=1);  p=1 -> (s=0,k=2) and (s=1,k=0).  For each of the 4
//   in-plane shifts (sh,sw) the A view is shared by every (class, kd) it reaches, so one
//   MMA of N = 12*Cout covers [kd=0 -> slice jd-1, pd=1 classes | kd=1 -> slice jd, pd=0 |
//   kd=2 -> slice jd, pd=1] with zero weight rows for unreachable classes (the MMA count,
//   not N, is what costs at these sizes: profiles/microbench/umma_rate.cu).
// Accumulators of a depth chunk live linearly in TMEM and are zeroed once (see
// conv3d_tc.cu); one single-use mbarrier per output group.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace casmvs {
namespace tc2 {

using namespace casmvs::tc;

enum { MODE_S2 = 0, MODE_T = 1 };

struct Params {
  const float* x;      // (B,Di,Hi,Wi,CIN)
  const float* wpk;    // [27][CIN][Cout]
  const float* bimg;   // pre-built B operand image [tap][CIN/4][BROWS][4] (tf32-rounded)
  const float* scale;  // [Cout]
  const float* shift;  // [Cout]
  const float* skip;   // output-shaped or null
  float* y;            // (B,Do,Ho,Wo,Cout)
  float slope;
  int B, Di, Hi, Wi, Do, Ho, Wo, Cout;     // Cout = channel count of the output tensor;
                                           // a CTA handles the COUT-channel chunk blockIdx.y
  int tiles_w, tiles_h, nchunks, dchunk;   // tiles over the M space (output for S2, input for T)
  int round_out;
};

template <int MODE, int CIN, int COUT>
struct Cfg {
  static constexpr int CQ = CIN / 4;
  // brick geometry
  static constexpr int BR = MODE == MODE_S2 ? 33 : 17;      // rows
  static constexpr int BW = MODE == MODE_S2 ? 17 : 9;       // column positions
  static constexpr int kSlotBytes = BR * CQ * BW * 16;
  // accumulator group (columns per output group) and B image rows per tap
  static constexpr int GW = MODE == MODE_S2 ? (COUT <= 16 ? 16 : 32) : 8 * COUT;
  static constexpr int BROWS = MODE == MODE_S2 ? 3 * GW : 12 * COUT;
  static constexpr int NTAP = MODE == MODE_S2 ? 9 : 4;      // A views per input slice
  static constexpr int kWBytes = NTAP * CIN * BROWS * 4;
  static constexpr int kRingOff = kWBytes;
  // deep layers (big bricks): 2-slot ring, one CTA per COUT-channel chunk of the output
  static constexpr int SLOTS = (kWBytes + 4 * kSlotBytes + 1024 <= 227 * 1024) ? 4 : 2;
  static constexpr int kParamOff = kRingOff + SLOTS * kSlotBytes;   // scale/shift [2][COUT pad 32]
  static constexpr int kBarOff = kParamOff + 2 * 32 * 4;
  static constexpr int kTotal = kBarOff + 128 + 32 * 8;
  static constexpr int kMaxGroups = 512 / GW;               // TMEM capacity
};

__host__ __device__ constexpr int tmem_cols_for2(int n) {
  return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

template <int MODE, int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1) conv3d_tc2_kernel(const Params p) {
  using C = Cfg<MODE, CIN, COUT>;
  constexpr int CQ = C::CQ, BR = C::BR, BW = C::BW, GW = C::GW, BROWS = C::BROWS;
  constexpr int SLOTS = C::SLOTS;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_w = s_base, s_ring = s_base + C::kRingOff, s_bar = s_base + C::kBarOff;
  float* s_param = reinterpret_cast<float*>(smem + C::kParamOff);
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 32, bar_tfull = s_bar + 128;
  volatile uint32_t* s_tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + C::kBarOff + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work item: (b, chunk of groups along depth, tile_h, tile_w) over the M space ----
  int item = blockIdx.x;
  const int tw = item % p.tiles_w; item /= p.tiles_w;
  const int th = item % p.tiles_h; item /= p.tiles_h;
  const int ck = item % p.nchunks;
  const int b = item / p.nchunks;
  const int w0 = tw * kTileW, h0 = th * kTileH;              // M-space origin of the tile
  const int Dm = MODE == MODE_S2 ? p.Do : p.Di;              // M-space depth
  const int g0 = ck * p.dchunk, g1 = min(Dm, g0 + p.dchunk);
  const int ng = g1 - g0;                                    // accumulator groups in this CTA
  // input slices walked: S2: s = 2*g0-1 .. 2*g1-1  (2*ng+1);  T: s = g0 .. g1  (ng+1)
  const int nslices = MODE == MODE_S2 ? 2 * ng + 1 : ng + 1;
  const int s_first = MODE == MODE_S2 ? 2 * g0 - 1 : g0;
  const uint32_t tmem_cols = tmem_cols_for2(p.dchunk * GW);

  init_barriers(bar_full, bar_empty, bar_tfull, SLOTS);
  if (warp == 0) tmem_alloc(smem_u32((const void*)s_tmem_ptr), tmem_cols);

  // B operand image (built once per launch by build_image2_kernel) -> smem
  const int co_base = blockIdx.y * COUT;
  load_image_async(s_w, p.bimg + (size_t)blockIdx.y * (C::kWBytes / 4), C::kWBytes);
  for (int i = threadIdx.x; i < 32; i += kThreads) {
    s_param[i] = (i < COUT) ? (p.scale ? __ldg(p.scale + co_base + i) : 1.f) : 0.f;
    s_param[32 + i] = (i < COUT) ? (p.shift ? __ldg(p.shift + co_base + i) : 0.f) : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  if (warp < 4) {
    for (int c = 0; c < ng * GW; c += 16)
      tmem_zero16(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= 4 && warp < kMmaWarp) {
    // ===================== producers =====================
    const int ptid = threadIdx.x - 128;
    for (int it = 0; it < nslices; ++it) {
      const int s = s_first + it;
      const int slot = it % SLOTS;
      if (it >= SLOTS) mbar_wait(bar_empty + 8 * slot, ((it / SLOTS) - 1) & 1);
      const uint32_t dst0 = s_ring + slot * C::kSlotBytes;
      const bool s_ok = (s >= 0) && (s < p.Di);
      const float* xs = p.x + (((size_t)b * p.Di + (s_ok ? s : 0)) * p.Hi) * (size_t)p.Wi * CIN;
      for (int c = ptid; c < BR * BW * CQ; c += kProducerThreads) {
        const int cq = c % CQ;
        const int vox = c / CQ;
        const int wl = vox % BW, r = vox / BW;     // wl: position in global-column order
        int ih, iw, wp;
        if (MODE == MODE_S2) {
          ih = 2 * h0 - 1 + r;
          iw = 2 * w0 - 1 + wl;                    // wl even -> odd plane, wl odd -> even plane
          wp = (wl & 1) ? (wl >> 1) : 8 + (wl >> 1);
        } else {
          ih = h0 + r; iw = w0 + wl; wp = wl;
        }
        const bool ok = s_ok && ih >= 0 && ih < p.Hi && iw >= 0 && iw < p.Wi;
        const float* src = ok ? xs + ((size_t)ih * p.Wi + iw) * CIN + cq * 4 : p.x;
        cp_async16(dst0 + ((r * CQ + cq) * BW + wp) * 16, src, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (it >= 1) {
        cp_async_wait<1>();
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * ((it - 1) % SLOTS));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    mbar_arrive(bar_full + 8 * ((nslices - 1) % SLOTS));
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (warp-uniform, elect-predicated) =====================
    constexpr uint32_t a_lbo = BW * 16;
    constexpr uint32_t a_sbo = (MODE == MODE_S2 ? 2 : 1) * CQ * BW * 16;
    constexpr uint32_t b_lbo = BROWS * 16, b_sbo = 128;
    const uint32_t elected = elect_one();
    const uint64_t a_desc0 = make_desc(s_ring, a_lbo, a_sbo);
    const uint64_t b_desc0 = make_desc(s_w, b_lbo, b_sbo);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
    for (int it = 0; it < nslices; ++it) {
      // which accumulator columns / B rows this input slice feeds
      int col, row0, ncols, done = -1;             // done: group index completed by this slice
      bool kd1_slice = false;
      if (MODE == MODE_S2) {
        if (it & 1) {                              // even input slice s = 2(g0+a): kd=1 -> group a
          const int a = (it - 1) >> 1;
          col = a * GW; row0 = 0; ncols = GW; kd1_slice = true;
        } else {                                   // odd slice: kd=2 -> a-1, kd=0 -> a
          const int a = it >> 1;
          const int lo = max(a - 1, 0), hi = min(a, ng - 1);
          col = lo * GW;
          row0 = (a - 1 >= 0) ? GW : 2 * GW;       // rows [W1 | W2 | W0]
          ncols = (hi - lo + 1) * GW;
          if (a >= 1) done = a - 1;
        }
      } else {
        // T: slice it -> group it (blocks kd=1,kd=2; if it < ng) and group it-1 (block kd=0)
        const bool cur = it < ng, prev = it >= 1;
        col = prev ? (it - 1) * GW + 4 * COUT : 0;
        row0 = prev ? 0 : 4 * COUT;
        ncols = (prev ? 4 * COUT : 0) + (cur ? 8 * COUT : 0);
        if (prev) done = it - 1;
      }
      (void)kd1_slice;
      const uint32_t idesc = make_idesc(128, ncols);
      const uint32_t acc = tmem_base + col;
      mbar_wait(bar_full + 8 * (it % SLOTS), (it / SLOTS) & 1);
      tc_fence_after();
      const uint32_t a_lo0 = (uint32_t)a_desc0 + (((it % SLOTS) * C::kSlotBytes) >> 4);
      const uint32_t b_lo0 = (uint32_t)b_desc0 + ((row0 * 16) >> 4);
#pragma unroll
      for (int tap = 0; tap < C::NTAP; ++tap) {
        int a_tap;
        if (MODE == MODE_S2) {
          const int kh = tap / 3, kw = tap % 3;
          const int wpos = kw == 1 ? 0 : (kw == 0 ? 8 : 9);   // even plane @0, odd plane @8
          a_tap = (kh * CQ * BW + wpos) * 16;
        } else {
          const int sh = tap >> 1, sw = tap & 1;
          a_tap = (sh * CQ * BW + sw) * 16;
        }
#pragma unroll
        for (int k8 = 0; k8 < CIN / 8; ++k8) {
          const uint32_t a_off = (a_tap + k8 * 2 * BW * 16) >> 4;
          const uint32_t b_off = (tap * (CIN * BROWS * 4) + k8 * 2 * BROWS * 16) >> 4;
          umma_tf32(acc, a_lo0 + a_off, a_hi, b_lo0 + b_off, b_hi, idesc, elected);
        }
      }
      if (done >= 0) umma_commit(bar_tfull + 8 * done, elected);
      umma_commit(bar_empty + 8 * (it % SLOTS), elected);
    }
  } else {
    // ===================== epilogue warps 0..3 =====================
    const int m = warp * 32 + lane;
    const int mh = h0 + (m >> 3), mw = w0 + (m & 7);          // M-space voxel of this thread
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < ng; ++g) {
      mbar_wait(bar_tfull + 8 * g, 0);
      tc_fence_after();
      if constexpr (MODE == MODE_S2) {
        float acc[GW];
        tmem_ld<GW>(lane_base + g * GW, acc);
        if (mh < p.Ho && mw < p.Wo) {
          const size_t o =
              ((((size_t)b * p.Do + (g0 + g)) * p.Ho + mh) * p.Wo + mw) * p.Cout + co_base;
#pragma unroll
          for (int c = 0; c < COUT; c += 4) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float t = fmaf(acc[c + k], s_param[c + k], s_param[32 + c + k]);
              t = t >= 0.f ? t : t * p.slope;
              v[k] = p.round_out ? to_tf32(t) : t;
            }
            st4(p.y + o + c, make_float4(v[0], v[1], v[2], v[3]));
          }
        }
      } else {
        // 8 classes x COUT columns: [pd][ph][pw][co]
#pragma unroll
        for (int cls = 0; cls < 8; ++cls) {
          constexpr int CW = COUT <= 16 ? 16 : 32;            // tcgen05.ld width
          float acc[CW];
          if constexpr (COUT == 8) {
            if (cls & 1) continue;                            // classes are read in pairs (16 cols)
            tmem_ld<16>(lane_base + g * GW + cls * COUT, acc);
          } else {
            tmem_ld<CW>(lane_base + g * GW + cls * COUT, acc);
          }
          constexpr int NC = COUT == 8 ? 2 : 1;               // classes held in acc[]
#pragma unroll
          for (int q = 0; q < NC; ++q) {
            const int cc = cls + q;
            const int pd = cc >> 2, ph = (cc >> 1) & 1, pw = cc & 1;
            const int od = 2 * (g0 + g) + pd, oh = 2 * mh + ph, ow = 2 * mw + pw;
            if (mh < p.Hi && mw < p.Wi) {
              const size_t o =
                  ((((size_t)b * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.Cout + co_base;
#pragma unroll
              for (int c = 0; c < COUT; c += 4) {
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float t = fmaf(acc[q * COUT + c + k], s_param[c + k], s_param[32 + c + k]);
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

// B operand image [tap][cq][row][4], tf32-rounded; see the header comment for the row order
template <int MODE, int CIN, int COUT>
__global__ void build_image2_kernel(const float* __restrict__ wpk, float* __restrict__ img,
                                    int cout_total) {
  using C = Cfg<MODE, CIN, COUT>;
  constexpr int CQ = C::CQ, GW = C::GW, BROWS = C::BROWS;
  constexpr int per = C::NTAP * CIN * BROWS;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per * (cout_total / COUT);
       t += gridDim.x * blockDim.x) {
    const int ck = t / per, i = t - ck * per;
    const int jq = i & 3;
    const int row = (i >> 2) % BROWS;
    const int r = (i >> 2) / BROWS;          // tap*CQ + cq
    const int cq = r % CQ, tap = r / CQ;
    const int ci = cq * 4 + jq;
    int kd = -1, kh = -1, kw = -1, co = -1;
    if (MODE == MODE_S2) {
      // rows [W(kd=1) | W(kd=2) | W(kd=0)], tap = kh*3+kw
      const int g = row / GW;
      co = row % GW;
      kd = g == 0 ? 1 : g == 1 ? 2 : 0;
      kh = tap / 3; kw = tap % 3;
      if (co >= COUT) kd = -1;
    } else {
      // tap = sh*2+sw; rows: block 0 kd=0 (pd=1), block 1 kd=1 (pd=0), block 2 kd=2 (pd=1);
      // inside a block: class (ph,pw) = 2*ph+pw, then co
      const int sh = tap >> 1, sw = tap & 1;
      const int blk = row / (4 * COUT);
      const int cls = (row / COUT) & 3;
      co = row % COUT;
      const int ph = cls >> 1, pw = cls & 1;
      kd = blk;
      kh = sh == 0 ? (ph == 0 ? 1 : 2) : (ph == 1 ? 0 : -1);
      kw = sw == 0 ? (pw == 0 ? 1 : 2) : (pw == 1 ? 0 : -1);
      if (kh < 0 || kw < 0) kd = -1;
    }
    float v = 0.f;
    if (kd >= 0)
      v = to_tf32(__ldg(wpk + ((size_t)((kd * 3 + kh) * 3 + kw) * CIN + ci) * cout_total +
                        ck * COUT + co));
    img[t] = v;
  }
}

template <int MODE, int CIN, int COUT>
static int launch2(Params p, cudaStream_t st) {
  using C = Cfg<MODE, CIN, COUT>;
  static_assert(C::kTotal <= 227 * 1024, "shared memory budget");
  auto kfn = conv3d_tc2_kernel<MODE, CIN, COUT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kTotal);
    if (e != cudaSuccess) {
      set_error("conv3d_tc2: cannot opt in to %d B of shared memory: %s", C::kTotal,
                cudaGetErrorString(e));
      return -2;
    }
    attr_set = true;
  }
  const int Dm = MODE == MODE_S2 ? p.Do : p.Di;
  const int Hm = MODE == MODE_S2 ? p.Ho : p.Hi, Wm = MODE == MODE_S2 ? p.Wo : p.Wi;
  p.tiles_w = (Wm + kTileW - 1) / kTileW;
  p.tiles_h = (Hm + kTileH - 1) / kTileH;
  int dchunk = Dm < C::kMaxGroups ? Dm : C::kMaxGroups;
  const long cols = (long)p.B * p.tiles_w * p.tiles_h;
  while (dchunk > 2 && cols * ((Dm + dchunk - 1) / dchunk) < (long)num_sms() * 3)
    dchunk = (dchunk + 1) / 2;
  p.dchunk = dchunk;
  p.nchunks = (Dm + dchunk - 1) / dchunk;
  const long items = cols * p.nchunks;
  const int nco = p.Cout / COUT;
  bool hit = false;
  float* img = image_cache_lookup(p.wpk, 2000 + MODE * 10000 + CIN * 100 + COUT,
                                  (size_t)C::kWBytes * nco, &hit);
  if (!img) { set_error("conv3d_tc2: cannot allocate the weight image"); return -2; }
  if (!hit) {
    build_image2_kernel<MODE, CIN, COUT><<<64, 256, 0, st>>>(p.wpk, img, p.Cout);
    if (int rc = after_launch("conv3d_tc2/build_image")) return rc;
  }
  p.bimg = img;
  kfn<<<dim3((unsigned)items, (unsigned)nco), kThreads, C::kTotal, st>>>(p);
  return after_launch("conv3d_tc2");
}

}  // namespace tc2

// Returns 0 when handled, 1 when the layer shape is left to the CUDA-core kernel.
int conv3d_tc2(const float* x, const float* wpk, const float* scale, const float* shift,
               float slope, const float* skip, float* y, int B, int Cin, int Cout, int D, int h,
               int w, int kind, int stride, int precision, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("CASMVS_TC2");
    enabled = e ? atoi(e) : 1;
  }
  if (!enabled || precision != CASMVS_TF32) return 1;
  tc2::Params p;
  p.x = x; p.wpk = wpk; p.scale = scale; p.shift = shift; p.skip = skip; p.y = y;
  p.slope = slope; p.B = B; p.Di = D; p.Hi = h; p.Wi = w; p.Cout = Cout; p.round_out = 1;
  if (kind == CASMVS_CONV && stride == 2) {
    if (skip) return 1;
    p.Do = (D - 1) / 2 + 1; p.Ho = (h - 1) / 2 + 1; p.Wo = (w - 1) / 2 + 1;
    if (Cin == 8 && Cout == 16) return tc2::launch2<tc2::MODE_S2, 8, 16>(p, st);
    if (Cin == 16 && Cout == 32) return tc2::launch2<tc2::MODE_S2, 16, 32>(p, st);
    if (Cin == 32 && Cout == 64) return tc2::launch2<tc2::MODE_S2, 32, 16>(p, st);   // conv5: 4 chunks
    return 1;
  }
  if (kind == CASMVS_CONV_TRANSPOSE) {
    p.Do = 2 * D; p.Ho = 2 * h; p.Wo = 2 * w;
    if (Cin == 16 && Cout == 8) return tc2::launch2<tc2::MODE_T, 16, 8>(p, st);
    if (Cin == 32 && Cout == 16) return tc2::launch2<tc2::MODE_T, 32, 16>(p, st);
    if (Cin == 64 && Cout == 32) return tc2::launch2<tc2::MODE_T, 64, 8>(p, st);     // conv7: 4 chunks
    return 1;
  }
  return 1;
}

}  // namespace casmvs
