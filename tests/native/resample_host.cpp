// Host harness of the image preprocessing kernels: compiles dust3r_b200/csrc/resample_core.h -- the very per-thread bodies the
// CUDA kernels of csrc/image_ops.cu call -- with g++ and runs them for every thread index a launch would cover (plus one
// ragged block of out-of-range indices, as the rounded-up grid produces).  tests/test_image_preprocess.py compares the result
// with Pillow / torchvision bit for bit, so the GPU code's index arithmetic and integer resampling are verified on machines
// without a GPU.  Same argument list as d3r_image_resize_crop_normalize minus the stream; pointers are HOST pointers here.
#include "../../dust3r_b200/csrc/resample_core.h"

using namespace d3r::image;

extern "C" int resample_host(const uint8_t* src, int32_t H0, int32_t W0, int32_t H1, int32_t W1, const int32_t* xbounds,
                             const int32_t* xcoefs, int32_t kx, const int32_t* ybounds, const int32_t* ycoefs, int32_t ky,
                             int32_t row0, int32_t rows, int32_t crop_x0, int32_t crop_y0, int32_t H2, int32_t W2, const float* lut,
                             uint8_t* tmp, float* out) {
  (void)H0; (void)kx; (void)ky;
  const int kThreads = 256;
  const long long n_tmp = (long long)rows * W2, n_out = (long long)H2 * W2;
  HorizontalArgs h{src, W0, row0, rows, crop_x0, W2, W1, xbounds, xcoefs, tmp};
  for (long long t = 0; t < (n_tmp + kThreads - 1) / kThreads * kThreads; ++t) horizontal_body(t, h);
  VerticalArgs v{tmp, row0, W2, H1, ybounds, ycoefs, crop_y0, H2, W2, lut, out};
  for (long long t = 0; t < (n_out + kThreads - 1) / kThreads * kThreads; ++t) vertical_body(t, v);
  return 0;
}
