// Input pipeline, device side: uint8 RGB images -> normalised planar fp32.
//
// Replaces T.ToTensor() + T.Normalize(mean, std) of the reference's data sets
// (datasets/dtu.py:130-137) for images that are uploaded as BYTES: (N,H,W,3) uint8 ->
// (N,3,H,W) float32 with  y = ((float)x / 255 - mean[c]) / std[c]  in exactly that operation
// order (torchvision: .div(255), .sub_(mean), .div_(std)), so results are bit-identical to the
// host path while the H2D copy carries a quarter of the bytes.  HBM-bound: 3 B read + 12 B
// written per pixel; the 256 x 3 possible outputs are tabulated once per block.
#include "common.cuh"

namespace casmvs {

__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, size_t hw,
                    float m0, float m1, float m2, float s0, float s1, float s2) {
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8, v = i & 255;
    const float mean = c == 0 ? m0 : c == 1 ? m1 : m2, sd = c == 0 ? s0 : c == 1 ? s1 : s2;
    lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), mean), sd);
  }
  __syncthreads();
  const size_t n = blockIdx.y;
  const uint8_t* ip = in + n * hw * 3;
  float* op = out + n * hw * 3;
  // 4 pixels (12 bytes = 3 words) per thread per step when the planes of every image stay
  // 16 B aligned (hw % 4 == 0); otherwise (N == 1, odd sizes) the scalar path does everything
  const size_t quads = (hw % 4 == 0) ? hw / 4 : 0;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads;
       q += (size_t)gridDim.x * blockDim.x) {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(ip + q * 12);
    const uint32_t a = __ldg(wp), b = __ldg(wp + 1), c = __ldg(wp + 2);
    // bytes: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
    float4 r, g, bl;
    r.x = lut[0][a & 255];          g.x = lut[1][(a >> 8) & 255];   bl.x = lut[2][(a >> 16) & 255];
    r.y = lut[0][a >> 24];          g.y = lut[1][b & 255];          bl.y = lut[2][(b >> 8) & 255];
    r.z = lut[0][(b >> 16) & 255];  g.z = lut[1][b >> 24];          bl.z = lut[2][c & 255];
    r.w = lut[0][(c >> 8) & 255];   g.w = lut[1][(c >> 16) & 255];  bl.w = lut[2][c >> 24];
    st4(op + q * 4, r);
    st4(op + hw + q * 4, g);
    st4(op + 2 * hw + q * 4, bl);
  }
  if (blockIdx.x == 0) {
    for (size_t p = quads * 4 + threadIdx.x; p < hw; p += blockDim.x)
      for (int c = 0; c < 3; ++c) op[c * hw + p] = lut[c][ip[p * 3 + c]];
  }
}

}  // namespace casmvs

using namespace casmvs;

extern "C" int casmvs_normalize_u8_fwd(const uint8_t* images, float* out, int N, int H, int W,
                                       const float* mean3, const float* std3, void* stream) {
  CASMVS_REQUIRE(images && out && mean3 && std3, "normalize_u8: null pointer");
  CASMVS_REQUIRE(N >= 0 && N <= 65535 && H > 0 && W > 0, "normalize_u8: bad dims");
  CASMVS_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "normalize_u8: zero std");
  if (N == 0) return 0;
  const size_t hw = (size_t)H * W;
  // word loads / float4 stores need hw % 4 == 0 for every image after the first; otherwise
  // (never the case for the reference's 32-divisible image sizes) fall to one image per launch
  CASMVS_REQUIRE(hw % 4 == 0 || N == 1, "normalize_u8: H*W must be a multiple of 4 for N > 1");
  CASMVS_REQUIRE((reinterpret_cast<uintptr_t>(images) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 15) == 0, "normalize_u8: misaligned");
  const unsigned bx = (unsigned)((hw / 4 + 255) / 256 < 1 ? 1 : (hw / 4 + 255) / 256);
  dim3 grd(bx < 2048 ? bx : 2048, (unsigned)N);
  normalize_u8_kernel<<<grd, 256, 0, as_stream(stream)>>>(images, out, hw, mean3[0], mean3[1],
                                                          mean3[2], std3[0], std3[1], std3[2]);
  return after_launch("normalize_u8");
}
