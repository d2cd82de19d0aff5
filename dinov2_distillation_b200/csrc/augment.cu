// GPU input pipeline: the per-image part of datasets/augmentations.py:24-78 that is pure pixel arithmetic,
// fused into one pass over the output batch:
//   RandomResizedCrop(size, scale, BICUBIC)   (augmentations.py:36-40)  crop box -> size x size, PIL's antialiased bicubic
//   RandomHorizontalFlip(0.5)                 (:41)
//   ToTensor + Normalize(IMAGENET mean/std)   (:60-66)
//   RandomErasing(p=0.25, value=0)            (:44-49)
// The random draws (crop box, flip, erase box) are made on the host by torchvision's own get_params, so the distributions
// -- and, under the same seed, the draws -- are the reference's; this code only applies them. RandAugment (:52-58: nine
// PIL ops per image) is not covered: with it the reference pipeline stays on the host.
// Input: decoded uint8 RGB images, HWC, any sizes, packed back to back in one buffer (offsets[b] = first byte of image b).
// Output: fp32 NCHW [B, 3, S, S], what DINOv2ViT.forward / the student consume.
//
// The reference resizes a PIL image, so parity means PIL's arithmetic (Pillow libImaging/Resample.c, 8 bits per channel),
// restated here integer for integer -- the result is BIT-EXACT with the reference transforms (tests/test_oracle_augment.py
// pins the oracle on PIL, the GPU test pins this file on the oracle):
//   per axis: scale = in / out, support = 2 * max(scale, 1); for output i: centre = (i + 0.5) * scale,
//   taps [int(centre - support + 0.5), int(centre + support + 0.5)) clipped to the crop, weights = Keys cubic (a = -0.5) at
//   (x - centre + 0.5) / max(scale, 1) in double, normalised by their sequential sum, rounded to 22-bit fixed point;
//   horizontal pass first: h = clip8((2^21 + sum p * kx) >> 22) -- an 8-bit intermediate image -- then the vertical pass on
//   h the same way. Kernel 1 builds the two coefficient tables per image; kernel 2 computes, per output pixel, the
//   horizontally resampled 8-bit value of each source row it needs and the vertical sum over them (the intermediate image
//   never exists in memory), then flip (by reading the mirrored column's table row), 1/255, normalisation and erasing.
#include "common.cuh"
#include "../../include/b200_distill.h"

namespace b200 {

constexpr int AUG_PRECISION_BITS = 32 - 8 - 2;   // Pillow: PRECISION_BITS

__device__ __forceinline__ double pil_bicubic(double x) {   // Pillow bicubic_filter, a = -0.5; no FMA contraction
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(1.5, x), -2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(x, -5.0), x), 8.0), x), -4.0), -0.5);
  return 0.0;
}

struct AugParams {
  const uint8_t* pixels;
  const long long* offsets;   // [B] byte offset of each image
  const int* hw;              // [B][2] source height, width
  const int* crop;            // [B][4] top, left, height, width
  const int* flip;            // [B]
  const int* erase;           // [B][4] top, left, height, width in the OUTPUT (height 0: none)
  float* out;                 // [B, 3, S, S]
  int* bounds;                // [B][2 axes][S][2]  first tap, tap count        (axis 0 = x / horizontal, 1 = y)
  int* kk;                    // [B][2 axes][S][KS] fixed-point weights
  int B, S, KS;
  float mean[3], std[3];
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc for one output index of one axis of one image
__global__ void __launch_bounds__(128) augment_coeffs_kernel(const AugParams p) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x, axis = blockIdx.y, b = blockIdx.z;
  if (i >= p.S) return;
  const int in_size = p.crop[4 * b + (axis == 0 ? 3 : 2)];
  const double scale = (double)in_size / (double)p.S;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)i, 0.5), scale);
  const double ss = 1.0 / filterscale;
  int xmin = (int)__dadd_rn(__dadd_rn(center, -support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  int n = xmax - xmin;
  if (n > p.KS) n = p.KS;   // (the host sized KS from the largest support: cannot happen)
  const long long row = ((long long)(b * 2 + axis) * p.S + i);
  int* k = p.kk + row * p.KS;
  double ww = 0.0;
  for (int x = 0; x < n; ++x) ww = __dadd_rn(ww, pil_bicubic(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss)));
  for (int x = 0; x < n; ++x) {
    double w = pil_bicubic(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss));
    if (ww != 0.0) w = w / ww;
    const double f = __dmul_rn(w, (double)(1 << AUG_PRECISION_BITS));
    k[x] = w < 0.0 ? (int)__dadd_rn(-0.5, f) : (int)__dadd_rn(0.5, f);
  }
  p.bounds[row * 2] = xmin;
  p.bounds[row * 2 + 1] = n;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= AUG_PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) augment_resample_kernel(const AugParams p) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.z;
  const int ox = blockIdx.x * 16 + (threadIdx.x & 15), oy = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (ox >= p.S || oy >= p.S) return;
  float* o = p.out + ((long long)b * 3 * p.S + oy) * p.S + ox;
  const long long plane = (long long)p.S * p.S;
  const int ei = p.erase[4 * b], ej = p.erase[4 * b + 1], eh = p.erase[4 * b + 2], ew = p.erase[4 * b + 3];
  if (eh > 0 && oy >= ei && oy < ei + eh && ox >= ej && ox < ej + ew) {   // RandomErasing(value=0) acts after Normalize
    o[0] = 0.f; o[plane] = 0.f; o[2 * plane] = 0.f;
    return;
  }
  const int W = p.hw[2 * b + 1];
  const int ct = p.crop[4 * b], cl = p.crop[4 * b + 1];
  const int sx = p.flip[b] ? p.S - 1 - ox : ox;   // hflip of the resized crop
  const long long rx = (long long)(b * 2 + 0) * p.S + sx, ry = (long long)(b * 2 + 1) * p.S + oy;
  const int x0 = p.bounds[rx * 2], xn = p.bounds[rx * 2 + 1];
  const int y0 = p.bounds[ry * 2], yn = p.bounds[ry * 2 + 1];
  const int* kx = p.kk + rx * p.KS;
  const int* ky = p.kk + ry * p.KS;
  const uint8_t* img = p.pixels + p.offsets[b];
  const int half = 1 << (AUG_PRECISION_BITS - 1);
  int a0 = half, a1 = half, a2 = half;
  for (int y = 0; y < yn; ++y) {
    const uint8_t* row = img + ((long long)(ct + y0 + y) * W + cl + x0) * 3;
    int h0 = half, h1 = half, h2 = half;
    for (int x = 0; x < xn; ++x) {
      const int k = __ldg(kx + x);
      h0 += (int)__ldg(row + 3 * x) * k;
      h1 += (int)__ldg(row + 3 * x + 1) * k;
      h2 += (int)__ldg(row + 3 * x + 2) * k;
    }
    const int k = __ldg(ky + y);
    a0 += clip8(h0) * k;   // (the 8-bit intermediate image of Pillow's two-pass resize)
    a1 += clip8(h1) * k;
    a2 += clip8(h2) * k;
  }
  // ToTensor: uint8 -> float / 255; Normalize: (x - mean) / std, both IEEE (no reciprocal, no contraction)
  const float v0 = __fdiv_rn((float)clip8(a0), 255.f), v1 = __fdiv_rn((float)clip8(a1), 255.f), v2 = __fdiv_rn((float)clip8(a2), 255.f);
  o[0] = __fdiv_rn(__fsub_rn(v0, p.mean[0]), p.std[0]);
  o[plane] = __fdiv_rn(__fsub_rn(v1, p.mean[1]), p.std[1]);
  o[2 * plane] = __fdiv_rn(__fsub_rn(v2, p.mean[2]), p.std[2]);
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_augment_ws_bytes(int B, int S, int max_taps) {
  if (B <= 0 || S <= 0 || max_taps <= 0) return 0;
  return (size_t)B * 2 * S * (2 + (size_t)max_taps) * sizeof(int);
}

extern "C" int b200_augment_batch(const unsigned char* pixels, const long long* offsets, const int* hw, const int* crop,
                                  const int* flip, const int* erase, float* out, int B, int S, int max_taps,
                                  const float* mean, const float* std, void* ws, size_t ws_bytes, void* stream) {
  B200_CHECK_ARG(pixels && offsets && hw && crop && flip && erase && out && mean && std && ws, "null argument");
  B200_CHECK_ARG(B > 0 && B <= 65535 && S > 0 && S <= 4096, "batch / output size out of range");
  B200_CHECK_ARG(max_taps >= 5 && max_taps <= 1024, "max_taps = 2 * ceil(2 * max(largest crop side / S, 1)) + 1, at most 1024");
  B200_CHECK_ARG(ws_bytes >= b200_augment_ws_bytes(B, S, max_taps), "workspace too small (b200_augment_ws_bytes)");
  AugParams p{};
  p.pixels = pixels; p.offsets = offsets; p.hw = hw; p.crop = crop; p.flip = flip; p.erase = erase; p.out = out;
  p.B = B; p.S = S; p.KS = max_taps;
  p.bounds = static_cast<int*>(ws);
  p.kk = p.bounds + (size_t)B * 2 * S * 2;
  for (int c = 0; c < 3; ++c) {
    B200_CHECK_ARG(std[c] > 0.f, "std must be positive");
    p.mean[c] = mean[c];
    p.std[c] = std[c];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CUDA_OK(launch_pdl(augment_coeffs_kernel, dim3((unsigned)cdiv(S, 128), 2, (unsigned)B), dim3(128), 0, st, p));
  B200_LAUNCH_OK();
  B200_CUDA_OK(launch_pdl(augment_resample_kernel, dim3((unsigned)cdiv(S, 16), (unsigned)cdiv(S, 16), (unsigned)B), dim3(256), 0, st, p));
  B200_LAUNCH_OK();
  return 0;
}
