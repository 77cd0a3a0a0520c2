// K2 (tensor-core variant, TMA producer) — 3x3x3 stride-1 convolution as an im2col-free
// implicit GEMM on tcgen05 (kind::tf32, fp32 accumulators in TMEM) with the norm-act (+skip)
// epilogue fused.  sm_100a only.
//
// Replaces (reference, relative to /root/reference): ConvBnReLU3D
// (models/modules.py:21-31) and the `prob` head (models/mvsnet.py:89,103) for the
// stride-1 layers of CostRegNet (conv0, conv2, conv4, conv6, prob).
//
// The GEMM: M = 128 voxels = 8(w) x 16(h) of one depth slice, K = Cin per tap, N = 3 x GW (the
// three kd taps share one A operand: an input slice feeds up to three output slices in ONE
// MMA).  The input brick of a depth slice (18 x 10 voxels with halo, up to 32 channels) is
// brought in by ONE TMA tiled load (cp.async.bulk.tensor.5d over x viewed as {C, W, H, D, B};
// out-of-bounds elements are zero-filled by the TMA unit = the conv's zero padding, in all
// three spatial dimensions) -- the first generation of this kernel issued 180-1440 16-byte
// cp.async per slice from four producer warps and was bound by them.  The brick
// is voxel-major [18][10][CB] with the TMA swizzle matching the row size (CB*4 = 128/64/32
// bytes -> SWIZZLE_128B/64B/32B); the A operand of tap (kh,kw) is a SHIFTED VIEW of it:
// descriptor start = brick + (kh*10 + kw)*rowbytes (+32 B per K=8 step), stride between
// 8-voxel groups (SBO) = one brick row of 10 voxels.  The hardware swizzle is a function of
// absolute shared-memory address bits, so a start address that is not atom-aligned reads
// consistently what the TMA wrote (profiles/microbench/umma_swizzle_view.cu: exact for all
// nine shifts and all three swizzle modes with base_offset = 0).
//
// Persistent CTAs (one launch wave), 6 warps: 0-3 epilogue, 4 TMA producer (one elected
// thread), 5 MMA issuer.  Ring full barriers are armed with expect_tx and completed by the
// TMA unit; ring empty barriers by tcgen05.commit; per output slice tfull (MMA -> epilogue)
// and tempty (epilogue -> MMA, accumulator read and re-zeroed) barriers.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include <mutex>

#include "tma_common.cuh"

namespace casmvs {

namespace tma {

using namespace tc;

constexpr int kThreadsTma = 6 * 32;
constexpr int kProdWarp = 4, kIssueWarp = 5;

struct Params {
  const float* bimg;    // pre-built B operand image [chunk][kh][kw][CIN/4][3*GW][4] (tf32-rounded)
  const float* scale;   // [Cout] or null
  const float* shift;   // [Cout] or null
  const float* skip;    // (B,D,H,W,Cout) or null
  float* y;             // (B,D,H,W,Cout)
  float slope;
  int B, D, H, W, Cout;   // Cout = channels handled by one CTA (<= GW)
  int cout_total;         // channel count of the output tensor; blockIdx.y selects the chunk
  int tiles_w, tiles_h, nchunks, dchunk;
  int round_out;        // round the stored activations to tf32 (unbiased next-layer operand)
  int planar;           // 1x3x3 kernel: input slice s feeds output slice s only (kd = 1)
  long long* dbg;       // optional timeline of CTA 0: [role][slice][4] clock64 stamps
};
// per-role clock64 timeline of CTA 0 (profiles/tc_timeline.py): compiled in only with
// `make TIMELINE=1` -- each stamp costs ~6 instructions in loops whose roles are bound by
// their own scalar instruction stream (profiles/r2_k2_n8_stalls.txt)
#ifdef CASMVS_TIMELINE
#define TMA_STAMP(role, idx, k)                                                                 \
  do {                                                                                          \
    if (p.dbg && blockIdx.x == 0 && (idx) < 64) p.dbg[((role) * 64 + (idx)) * 4 + (k)] = clock64(); \
  } while (0)
#else
#define TMA_STAMP(role, idx, k) do { } while (0)
#endif

template <int CIN, int GW, int SLOTS_>
struct Smem {
  static constexpr int SLOTS = SLOTS_;
  static constexpr int CB = CIN > 32 ? 32 : CIN;                    // channels per brick
  static constexpr int NB = CIN / CB;                               // bricks per slice
  static constexpr int ROWB = CB * 4;                               // bytes per voxel = swizzle span
  static constexpr int kBrickData = kHaloH * kHaloW * ROWB;         // bytes one TMA load writes
  static constexpr int kBrickBytes = (kBrickData + 1023) / 1024 * 1024;
  static constexpr int kSlotBytes = NB * kBrickBytes;
  static constexpr int kWBytes = 9 * CIN * 3 * GW * 4;              // [kh][kw][cq][3*GW][4]
  static constexpr int kRingOff = 0;                                // 1024-aligned (swizzle atoms)
  static constexpr int kWOff = SLOTS * kSlotBytes;
  static constexpr int kParamOff = kWOff + kWBytes;                 // scale/shift [2][GW]
  static constexpr int kBarOff = kParamOff + 2 * GW * 4;
  // barriers: full[8] @0, empty[8] @64, tmem ptr @128, tfull[32] @192, tempty[32] @448
  static constexpr int kTotal = kBarOff + 192 + 32 * 8 + 32 * 8 + 1024;   // + alignment slack
  // UMMA layout type of the A operand: SWIZZLE_128B = 2, 64B = 4, 32B = 6
  static constexpr uint32_t kLayout = ROWB == 128 ? 2u : ROWB == 64 ? 4u : 6u;
};

__host__ __device__ constexpr int tmem_cols_for(int n) {
  return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

template <int CIN, int GW, int SLOTS_>
__global__ void __launch_bounds__(kThreadsTma, 1)
conv3d_tma_kernel(const __grid_constant__ CUtensorMap xmap, const Params p) {
  using S = Smem<CIN, GW, SLOTS_>;
  constexpr int CQ = CIN / 4;
  constexpr int SLOTS = S::SLOTS;
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
  const uint32_t tmem_cols = tmem_cols_for(p.dchunk * GW);
  const int total_items = p.B * p.nchunks * p.tiles_h * p.tiles_w;

  // ---- one-time setup ----
  if (threadIdx.x == 0) TMA_STAMP(3, 0, 0);
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
  const int co_base = blockIdx.y * p.Cout;
  for (int i = threadIdx.x; i < GW; i += kThreadsTma) {
    s_param[i] = (i < p.Cout) ? (p.scale ? __ldg(p.scale + co_base + i) : 1.f) : 0.f;
    s_param[GW + i] = (i < p.Cout) ? (p.shift ? __ldg(p.shift + co_base + i) : 0.f) : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  if (threadIdx.x == 0) load_image_bulk(s_w, p.bimg + (size_t)blockIdx.y * (S::kWBytes / 4), S::kWBytes, bar_w);
  // One accumulator (GW columns) per output slice of a chunk, laid out linearly, so the three
  // slices an input slice feeds are always adjacent columns.  Zeroed here, and re-zeroed by the
  // epilogue after every read: all MMAs accumulate (the accumulate flag is per instruction, not
  // per column, so a first-touch overwrite is not expressible for one group of three).
  if (warp < 4) {
    for (int c = 0; c < p.dchunk * GW; c += 16)
      tmem_zero16(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) TMA_STAMP(3, 0, 1);
  // nothing above depends on the previous kernel of the stream (see tma_common.cuh)
  tma::pdl_trigger();
  tma::pdl_wait();
  bool w_ready = false;                             // MMA issuer: weight image has landed
  uint32_t gs = 0;                                  // slices processed before this item (all roles)
  int ep = 0;                                       // items processed by this CTA
  for (int item0 = blockIdx.x; item0 < total_items; item0 += gridDim.x, ++ep) {
    int item = item0;
    const int tw = item % p.tiles_w; item /= p.tiles_w;
    const int th = item % p.tiles_h; item /= p.tiles_h;
    const int ck = item % p.nchunks;
    const int b = item / p.nchunks;
    const int w0 = tw * kTileW, h0 = th * kTileH;
    const int d0 = ck * p.dchunk, d1 = min(p.D, d0 + p.dchunk);
    const int nd = d1 - d0;
    const int halo = p.planar ? 0 : 1;
    const int nslices = nd + 2 * halo;              // input slices d0-halo .. d1-1+halo

    if (warp == kProdWarp) {
      // ===================== producer: one TMA load per brick =====================
      if (lane == 0) {
        for (int it = 0; it < nslices; ++it) {
          const uint32_t g = gs + it;
          const int slot = g % SLOTS;
          TMA_STAMP(0, g, 0);
          if (g >= (uint32_t)SLOTS) mbar_wait(bar_empty + 8 * slot, ((g / SLOTS) - 1) & 1);
          TMA_STAMP(0, g, 1);
          const uint32_t dst = s_ring + slot * S::kSlotBytes;
          mbar_expect_tx(bar_full + 8 * slot, S::NB * S::kBrickData);
#pragma unroll
          for (int nb = 0; nb < S::NB; ++nb)
            tma_load_5d(dst + nb * S::kBrickBytes, &xmap, bar_full + 8 * slot, nb * S::CB,
                        w0 - 1, h0 - 1, d0 - halo + it, b);
          TMA_STAMP(0, g, 2);
        }
      }
      __syncwarp();
    } else if (warp == kIssueWarp) {
      // ===================== MMA issuer =====================
      // Warp-uniform code with elect-predicated issue (descriptors stay in uniform registers);
      // everything per MMA is base + compile-time offset.
      constexpr uint32_t a_lbo = 16, a_sbo = kHaloW * S::ROWB;          // 8-voxel group stride
      constexpr uint32_t b_lbo = 3 * GW * 16, b_sbo = 128;
      constexpr int KPB = S::CB / 8;                                      // K=8 steps per brick
      const uint32_t elected = elect_one();
      const uint64_t a_desc0 = make_desc(s_ring, a_lbo, a_sbo) | ((uint64_t)S::kLayout << 61);
      const uint64_t b_desc0 = make_desc(s_w, b_lbo, b_sbo);
      const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
      for (int it = 0; it < nslices; ++it) {
        const uint32_t g = gs + it;
        // input slice `it` feeds output slices j = it - kd, kd = 0,1,2, clipped to [0,nd):
        // columns [j_lo*GW, (j_hi+1)*GW), B rows [(2-kd_hi)*GW, (3-kd_lo)*GW)
        const int kd_lo = p.planar ? 1 : max(0, it - (nd - 1));
        const int kd_hi = p.planar ? 1 : min(2, it);
        const int j_lo = p.planar ? it : it - kd_hi;
        const uint32_t idesc = make_idesc(128, (kd_hi - kd_lo + 1) * GW);
        const uint32_t acc = tmem_base + j_lo * GW;
        if (lane == 0) TMA_STAMP(1, g, 0);
        mbar_wait(bar_full + 8 * (g % SLOTS), (g / SLOTS) & 1);
        if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }
        if (lane == 0) TMA_STAMP(1, g, 1);
        // first touch of group `it` in this item: the epilogue must have drained + re-zeroed it
        if (ep > 0 && it < nd) mbar_wait(bar_tempty + 8 * it, (ep - 1) & 1);   // (it < nd always when planar)
        if (lane == 0) TMA_STAMP(1, g, 2);
        tc_fence_after();
        const uint32_t a_lo0 = (uint32_t)a_desc0 + (((g % SLOTS) * S::kSlotBytes) >> 4);
        const uint32_t b_lo0 = (uint32_t)b_desc0 + (((2 - kd_hi) * GW * 16) >> 4);
#pragma unroll
        for (int khw = 0; khw < 9; ++khw) {
          const int kh = khw / 3, kw = khw % 3;
#pragma unroll
          for (int k8 = 0; k8 < CIN / 8; ++k8) {
            const uint32_t a_off = ((kh * kHaloW + kw) * S::ROWB + (k8 % KPB) * 32 +
                                    (k8 / KPB) * S::kBrickBytes) >> 4;
            const uint32_t b_off = (khw * (CIN * 3 * GW * 4) + k8 * 2 * 3 * GW * 16) >> 4;
            umma_tf32(acc, a_lo0 + a_off, a_hi, b_lo0 + b_off, b_hi, idesc, elected);
          }
        }
        if (p.planar) umma_commit(bar_tfull + 8 * it, elected);        // slice it complete
        else if (it >= 2) umma_commit(bar_tfull + 8 * (it - 2), elected);   // slice it-2 complete
        umma_commit(bar_empty + 8 * (g % SLOTS), elected);             // smem slot free
        if (lane == 0) TMA_STAMP(1, g, 3);
      }
      // groups this (short) chunk did not use go through the same handshake (empty -> full)
      // so that every barrier sees exactly one completion per item and no phase can alias
      for (int j = nd; j < p.dchunk; ++j) {
        if (ep > 0) mbar_wait(bar_tempty + 8 * j, (ep - 1) & 1);
        if (elected) mbar_arrive(bar_tfull + 8 * j);
        __syncwarp();
      }
    } else {
      // ===================== epilogue warps 0..3 =====================
      const int m = warp * 32 + lane;              // GEMM row = TMEM lane
      const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
      const bool in_range = oh < p.H && ow < p.W;
      const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int j = 0; j < p.dchunk; ++j) {
        if (threadIdx.x == 0) TMA_STAMP(2, ep * p.dchunk + j, 0);
        mbar_wait(bar_tfull + 8 * j, ep & 1);
        if (threadIdx.x == 0) TMA_STAMP(2, ep * p.dchunk + j, 1);
        if (j >= nd) {                             // unused group: handshake only
          mbar_arrive(bar_tempty + 8 * j);
          continue;
        }
        tc_fence_after();
        float acc[GW];
        tmem_ld<GW>(lane_base + j * GW, acc);
#pragma unroll
        for (int c = 0; c < GW; c += 16) tmem_zero16(lane_base + j * GW + c);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar_tempty + 8 * j);           // group j drained and zero again
        if (threadIdx.x == 0) TMA_STAMP(2, ep * p.dchunk + j, 2);
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
    gs += nslices;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- host side ----

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// Generic tiled-map encoder for other kernels of the library (K1's feature boxes): fp32
// elements, rank <= 5, zero fill out of bounds.  0 on success.
int encode_tiled(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  auto enc = encode_fn();
  if (!enc) { set_error("tma: cuTensorMapEncodeTiled is not available"); return -2; }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                         const_cast<void*>(base), gdim, gstr, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tma: cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
    return -2;
  }
  return 0;
}

// Tensor maps are pure functions of (pointer, shape, box): memoised, since inference calls
// every layer with the same workspace pointers each step.
struct MapEntry { const void* x; int B, D, H, W, C, CB, bw, bh, sw; CUtensorMap map; };
static MapEntry g_maps[128];
static int g_maps_n = 0, g_maps_next = 0;
static std::mutex g_maps_mu;

const CUtensorMap* input_map(const float* x, int B, int D, int H, int W, int C, int CB, int box_w,
                             int box_h, int stride_w) {
  // the returned map is a per-thread copy: ring slots may be recycled by other threads
  static thread_local CUtensorMap t_ret;
  std::lock_guard<std::mutex> lock(g_maps_mu);
  for (int i = 0; i < g_maps_n; ++i) {
    const MapEntry& e = g_maps[i];
    if (e.x == x && e.B == B && e.D == D && e.H == H && e.W == W && e.C == C && e.CB == CB &&
        e.bw == box_w && e.bh == box_h && e.sw == stride_w) {
      t_ret = e.map;
      return &t_ret;
    }
  }
  auto enc = encode_fn();
  if (!enc) { set_error("conv3d_tma: cuTensorMapEncodeTiled is not available"); return nullptr; }
  MapEntry& e = g_maps[g_maps_next];
  g_maps_next = (g_maps_next + 1) % 128;
  if (g_maps_n < 128) ++g_maps_n;
  const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D,
                              (cuuint64_t)B};
  const cuuint64_t gstr[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4,
                              (cuuint64_t)D * H * W * C * 4};
  const cuuint32_t box[5] = {(cuuint32_t)CB, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
  const cuuint32_t estr[5] = {1, (cuuint32_t)stride_w, 1, 1, 1};
  const CUtensorMapSwizzle sw = CB * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : CB * 4 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                               : CU_TENSOR_MAP_SWIZZLE_32B;
  const CUresult r = enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(x), gdim,
                         gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    e.x = nullptr;
    set_error("conv3d_tma: cuTensorMapEncodeTiled failed (%d) for C=%d W=%d H=%d D=%d B=%d",
              (int)r, C, W, H, D, B);
    return nullptr;
  }
  e.x = x; e.B = B; e.D = D; e.H = H; e.W = W; e.C = C; e.CB = CB; e.bw = box_w; e.bh = box_h; e.sw = stride_w;
  t_ret = e.map;
  return &t_ret;
}


template <int CIN, int GW, int SLOTS>
static int launch(const float* x, const float* wpk, Params p, cudaStream_t st) {
  using S = Smem<CIN, GW, SLOTS>;
  auto kfn = conv3d_tma_kernel<CIN, GW, SLOTS>;
  static std::atomic<bool> attr_set[kMaxDevices];
  if (int rc = opt_in_smem(kfn, S::kTotal, attr_set, "conv3d_tma")) return rc;
  const CUtensorMap* map = input_map(x, p.B, p.D, p.H, p.W, CIN, S::CB, kHaloW, kHaloH);
  if (!map) return -2;
  // resident CTAs per SM by shared memory (1 KB per CTA is reserved by the system); the TMEM
  // of all of them must fit in 512 columns: one accumulator group (GW columns) per output slice
  static int per_sm_env = -1, dchunk_env = -1;
  if (per_sm_env < 0) {
    const char* e = getenv("CASMVS_TMA_PER_SM");
    per_sm_env = e ? atoi(e) : 0;
    const char* d = getenv("CASMVS_TMA_DCHUNK");
    dchunk_env = d ? atoi(d) : 0;
  }
  int per_sm = (228 * 1024) / (S::kTotal + 1024);
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  if (per_sm_env > 0 && per_sm_env < per_sm) per_sm = per_sm_env;
  const int nco = p.cout_total / p.Cout;
  int cap = pow2_floor(512 / per_sm) / GW;
  if (cap > 32) cap = 32;
  const long cols = (long)p.B * p.tiles_w * p.tiles_h;
  int dchunk = pick_dchunk(p.D, cap, cols, (long)num_sms() * per_sm / nco, 1, p.planar ? 0 : 2);
  if (dchunk_env > 0 && dchunk_env <= cap) dchunk = dchunk_env < p.D ? dchunk_env : p.D;
  p.dchunk = dchunk;
  p.nchunks = (p.D + dchunk - 1) / dchunk;
  const ImageRef ir = image_cache_get(wpk, 1000 + CIN * 100 + GW, (size_t)S::kWBytes * nco, st);
  if (!ir.img) return -2;
  if (!ir.hit) {
    if (int rc = build_stride1_image(wpk, ir.img, CIN, GW, p.Cout, p.cout_total, st)) return rc;
    image_cache_built(ir.img, st);
  }
  p.bimg = ir.img;
  const long items = (long)p.B * p.nchunks * p.tiles_h * p.tiles_w;
  int resident = num_sms() * per_sm / nco;
  if (resident < 1) resident = 1;
  const long gx = items < resident ? items : resident;
  launch_pdl(ir.settled, kfn, dim3((unsigned)gx, (unsigned)nco), kThreadsTma, S::kTotal, st, *map, p);
  return after_launch("conv3d_tma");
}

}  // namespace tma

// Returns 0 when handled, 1 when the layer shape is left to the other kernels.
int conv3d_tma(const float* x, const float* wpk, const float* scale, const float* shift,
               float slope, const float* skip, float* y, int B, int Cin, int Cout, int D, int h,
               int w, int kind, int stride, int precision_flags, cudaStream_t st) {
  const int precision = precision_flags & 0xff;
  static int enabled = -1, round_out = 1;
  static long long* dbg = nullptr;
  if (enabled < 0) {
    const char* e = getenv("CASMVS_TMA");
    enabled = e ? atoi(e) : 1;
    if (const char* s = getenv("CASMVS_TC_ROUND")) round_out = atoi(s);
    if (const char* s = getenv("CASMVS_TC_DBG")) dbg = (long long*)strtoull(s, nullptr, 0);
  }
  if (!enabled || precision != CASMVS_TF32) return 1;
  if ((kind != CASMVS_CONV && kind != CASMVS_CONV_PLANAR) || stride != 1) return 1;
  const bool deep = Cin == 64 && Cout == 64;          // conv6: 16-channel Cout slices
  if (!deep && (!(Cin == 8 || Cin == 16 || Cin == 32) || Cout > 32)) return 1;
  // the TMA global strides must be multiples of 16 B and the base 16 B aligned
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return 1;
  tma::Params p;
  p.scale = scale; p.shift = shift; p.skip = skip; p.y = y;
  p.slope = slope; p.B = B; p.D = D; p.H = h; p.W = w;
  p.Cout = deep ? 16 : Cout; p.cout_total = Cout;
  p.tiles_w = (w + tc::kTileW - 1) / tc::kTileW;
  p.tiles_h = (h + tc::kTileH - 1) / tc::kTileH;
  const int npad = p.Cout <= 16 ? 16 : 32;
  p.dbg = dbg;
  p.planar = kind == CASMVS_CONV_PLANAR ? 1 : 0;
  // the prob head feeds the softmax: keep fp32; callers can ask for unrounded outputs
  p.round_out = (round_out && Cout > 1 && !(precision_flags & CASMVS_KEEP_FP32_OUT)) ? 1 : 0;
#define TMA_CASE(CI, NP, SL) \
  if (Cin == CI && npad == NP) return tma::launch<CI, NP, SL>(x, wpk, p, st);
  TMA_CASE(8, 16, 4) TMA_CASE(8, 32, 4) TMA_CASE(16, 16, 4) TMA_CASE(16, 32, 4)
  TMA_CASE(32, 16, 4) TMA_CASE(32, 32, 4) TMA_CASE(64, 16, 2)
#undef TMA_CASE
  return 1;
}

}  // namespace casmvs
