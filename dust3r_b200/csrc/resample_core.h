// Per-thread bodies of the image preprocessing kernels (csrc/image_ops.cu), written once for device AND host: the CUDA kernels
// call them with t = blockIdx.x * blockDim.x + threadIdx.x, tests/native/resample_host.cpp compiles this very header with g++ and
// calls them in a loop over t, so the index arithmetic and the integer resampling that run on the GPU are checked bit for bit
// against Pillow on machines without a GPU.
//
// What is computed (dust3r/utils/image.py:62-71 `_resize_pil_image`, :101-124 crop + ImgNorm, i.e. Pillow's Image.resize for
// 8-bit images = two separable passes of src/libImaging/Resample.c, Image.crop, torchvision ToTensor + Normalize(0.5, 0.5)):
//   horizontal pass   tmp[y][x][c] = clip8((2^21 + sum_i src[y][lo_x + i][c] * kx[x][i]) >> 22)      uint8 -> uint8
//   vertical pass     out[c][y][x] = lut[clip8((2^21 + sum_i tmp[lo_y + i][x][c] * ky[y][i]) >> 22)]   uint8 -> fp32 CHW
// with 22-bit fixed-point coefficient tables built on the host (dust3r_b200/utils/image.py: Pillow's precompute_coeffs /
// normalize_coeffs_8bpc in double precision) and lut[v] = (v / 255 - 0.5) / 0.5 as torch computes it on the CPU.  Only the
// rows / columns of the intermediate image that the cropped output reads are produced.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define D3R_IMG_HD __host__ __device__ __forceinline__
#else
#define D3R_IMG_HD inline
#endif

namespace d3r {
namespace image {

constexpr int kPrecisionBits = 22;   // Resample.c PRECISION_BITS = 32 - 8 - 2

// Resample.c clip8: arithmetic shift, then clamp to [0, 255]
D3R_IMG_HD uint8_t clip8(uint32_t acc) {
  const int32_t v = (int32_t)acc >> kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

struct HorizontalArgs {
  const uint8_t* src;      // [H0][W0][3] decoded RGB
  int32_t W0;              // source row length in pixels
  int32_t row0, rows;      // source rows the vertical pass will read: [row0, row0 + rows)
  int32_t col0, cols;      // resized columns that survive the crop: [col0, col0 + cols)
  const int32_t* bounds;   // [W1][2] first source column, tap count
  const int32_t* coefs;    // [W1][ksize]
  int32_t ksize;
  uint8_t* tmp;            // [rows][cols][3]
};

// thread t -> (row yi, column xi, channel c) of tmp, c fastest: neighbouring threads write neighbouring bytes
D3R_IMG_HD void horizontal_body(long long t, const HorizontalArgs& a) {
  const long long total = (long long)a.rows * a.cols * 3;
  if (t >= total) return;
  const int c = (int)(t % 3);
  const long long r = t / 3;
  const int xi = (int)(r % a.cols);
  const int yi = (int)(r / a.cols);
  const int x = a.col0 + xi;
  const int lo = a.bounds[2 * x], cnt = a.bounds[2 * x + 1];
  const int32_t* k = a.coefs + (long long)x * a.ksize;
  const uint8_t* p = a.src + ((long long)(a.row0 + yi) * a.W0 + lo) * 3 + c;
  uint32_t acc = 1u << (kPrecisionBits - 1);
  for (int i = 0; i < cnt; ++i) acc += (uint32_t)((int32_t)p[3 * i] * k[i]);
  a.tmp[t] = clip8(acc);
}

struct VerticalArgs {
  const uint8_t* tmp;      // [rows][cols][3] written by the horizontal pass
  int32_t row0, cols;      // as above
  const int32_t* bounds;   // [H1][2] first source row, tap count
  const int32_t* coefs;    // [H1][ksize]
  int32_t ksize;
  int32_t crop_y0;         // first resized row of the crop
  int32_t H2, W2;          // output size (W2 == cols)
  const float* lut;        // [256]
  float* out;              // [3][H2][W2]
};

// thread t -> (channel c, row y2, column x2) of out, x2 fastest: coalesced fp32 stores, byte loads 3 apart
D3R_IMG_HD void vertical_body(long long t, const VerticalArgs& a) {
  const long long total = 3ll * a.H2 * a.W2;
  if (t >= total) return;
  const int x2 = (int)(t % a.W2);
  const long long r = t / a.W2;
  const int y2 = (int)(r % a.H2);
  const int c = (int)(r / a.H2);
  const int y1 = a.crop_y0 + y2;
  const int lo = a.bounds[2 * y1] - a.row0, cnt = a.bounds[2 * y1 + 1];
  const int32_t* k = a.coefs + (long long)y1 * a.ksize;
  const long long pitch = (long long)a.cols * 3;
  const uint8_t* p = a.tmp + (long long)lo * pitch + (long long)x2 * 3 + c;
  uint32_t acc = 1u << (kPrecisionBits - 1);
  for (int i = 0; i < cnt; ++i) acc += (uint32_t)((int32_t)p[i * pitch] * k[i]);
  a.out[t] = a.lut[clip8(acc)];
}

}  // namespace image
}  // namespace d3r
