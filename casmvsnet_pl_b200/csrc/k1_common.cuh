// Device helpers shared by the K1 kernels (warp_cost.cu: gather-from-L1 generation and the
// generic / group-wise-correlation variants; warp_cost_smem.cu: TMA-staged generation).
#pragma once
#include "common.cuh"

namespace casmvs {

constexpr int kCPT = 8;          // channels per thread

// ---- packed fp32x2 helpers (Blackwell FFMA2 / 256-bit LDG, STG) -----------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpk2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
struct Tex8 { u64 v[4]; };   // 8 channels of one texel, as 4 packed pairs
__device__ __forceinline__ Tex8 ldg256(const float* p) {   // 32-byte aligned
  Tex8 t;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];"
               : "=l"(t.v[0]), "=l"(t.v[1]), "=l"(t.v[2]), "=l"(t.v[3]) : "l"(p));
  return t;
}
__device__ __forceinline__ void stg256(float* p, const u64 (&v)[4]) {
  asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(v[2]),
               "l"(v[3]) : "memory");
}

__device__ __forceinline__ float round_tf32_f(float x) {
  uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r);
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}


// The kernel is instruction-issue bound (ncu: issue ~50 %, DRAM ~10 % in the first
// version), so the sampler is written for instruction count:
//  * the 2x2 window is addressed as ONE base (clamped to [0,w-2]x[0,h-2]) plus
//    compile-time offsets {0, C, w*C, w*C + C}; the zero-padding rule of
//    grid_sample becomes a remap of the four weights at the image border,
//  * reciprocals are MUFU.RCP (1 ulp) instead of the IEEE sequence,
//  * the blend runs on packed FFMA2 with 256-bit texel loads.
struct Window {
  Tex8 t00, t01, t10, t11;
};

// Branch-free: a sample that contributes nothing (behind the camera / fully outside
// the source image) gets four zero weights and a clamped, always-valid address, so
// the loop body is straight-line code and ptxas can keep the loads of all views in
// flight at once.  CT = compile-time channel count (0: use C).
template <int CT>
__device__ __forceinline__ void sample_view(const float* __restrict__ vbase, float qx, float qy,
                                            float qz, int h, int w, int C, int row_floats,
                                            Window& win, float& w00, float& w01, float& w10,
                                            float& w11) {
  const float rz = rcp_approx(qz);
  const float u = qx * rz, v = qy * rz;
  const float x0f = floorf(u), y0f = floorf(v);
  // float->int saturates, so huge |u| fails the range test like ATen's within_bounds;
  // q_z <= 1e-7 is mapped to (w,h) = fully outside by the reference (modules.py:76-79)
  const int x0 = __float2int_rd(u), y0 = __float2int_rd(v);
  const bool valid = (qz > 1e-7f) && (unsigned)(x0 + 1) <= (unsigned)w &&
                     (unsigned)(y0 + 1) <= (unsigned)h;
  const float fx = u - x0f, fy = v - y0f;
  float wxa = 1.f - fx, wxb = fx, wya = 1.f - fy, wyb = fy;
  // border: texel x0 (or x0+1) is outside => its weight is dropped; the pair
  // (xs, xs+1) stays inside the image and the surviving weight moves to its slot
  if (x0 < 0) { wxa = wxb; wxb = 0.f; }
  if (x0 > w - 2) { wxb = wxa; wxa = 0.f; }
  if (y0 < 0) { wya = wyb; wyb = 0.f; }
  if (y0 > h - 2) { wyb = wya; wya = 0.f; }
  if (!valid) { wxa = 0.f; wxb = 0.f; }
  const int xs = min(max(x0, 0), w - 2), ys = min(max(y0, 0), h - 2);
  w00 = wxa * wya; w01 = wxb * wya; w10 = wxa * wyb; w11 = wxb * wyb;
  const int cc = CT > 0 ? CT : C;
  const unsigned off = (unsigned)(ys * row_floats + xs * cc);
  const float* p = vbase + off;
  win.t00 = ldg256(p);
  win.t01 = ldg256(p + cc);
  win.t10 = ldg256(p + row_floats);
  win.t11 = ldg256(p + row_floats + cc);
}


// Where a pixel's depth hypotheses come from: the (B,D,h,w) tensor of the public API, or -- in
// the cascade -- the ladder first + step*d that get_depth_values / the initial planes define
// (models/modules.py:44-48, models/mvsnet.py:215-229), generated in the kernel with the SAME two
// roundings (multiply, then add) so that the (B,D,h,w) tensor never has to exist in HBM.
struct Hyp {
  const float* dv;         // (B,D,h,w) or null => ladder
  const float* first_map;  // (B,h,w) first hypothesis per pixel, or null
  const float* first_b;    // (B) first hypothesis per batch item, or null
  const float* step_b;     // (B) plane spacing per batch item, or null
  float first, step;       // scalars used when the pointers above are null
};
struct HypPix {
  const float* p;
  size_t hw;
  float first, step;
  __device__ __forceinline__ HypPix(const Hyp& h, int b, int D, size_t hw_, int pix) : hw(hw_) {
    p = h.dv ? h.dv + (size_t)b * D * hw_ + pix : nullptr;
    first = h.first_map ? __ldg(h.first_map + (size_t)b * hw_ + pix) : h.first_b ? __ldg(h.first_b + b) : h.first;
    step = h.step_b ? __ldg(h.step_b + b) : h.step;
  }
  __device__ __forceinline__ float at(int d) const {
    return p ? __ldg(p + (size_t)d * hw) : __fadd_rn(first, __fmul_rn(step, (float)d));
  }
};

}  // namespace casmvs
