// Image preprocessing of load_images on the GPU (dust3r/utils/image.py:62-71, 101-124): the decoded 8-bit RGB
// image is uploaded ONCE as bytes (3 B / source pixel instead of PIL resizing on a host core and 12 B / output pixel going up) and
// Pillow's two-pass fixed-point resampling, the centre crop and ImgNorm run in HBM.  Integer / byte work, HBM-bound and tiny
// (36 MB for a 12 Mpx photo): one thread per output PIXEL (three accumulators share every coefficient load), consecutive threads
// on consecutive columns so that a warp's loads of a tap cover one contiguous stretch of a row (overlapping windows of neighbouring
// outputs are served by L1), tap-major coefficient tables (coalesced / broadcast loads), coalesced stores; the 4.6 MB intermediate
// image never leaves the 126 MB L2.  No tensor cores, no shared-memory staging needed at this size.  The per-thread bodies live in
// resample_core.h so that the host test harness runs the same code (bit-exact against Pillow without a GPU).
#include "d3r_common.cuh"
#include "prof.h"
#include "resample_core.h"

namespace d3r {
namespace image {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) horizontal_kernel(HorizontalArgs a) {
  horizontal_body((long long)blockIdx.x * blockDim.x + threadIdx.x, a);
}

__global__ void __launch_bounds__(kThreads) vertical_kernel(VerticalArgs a) {
  vertical_body((long long)blockIdx.x * blockDim.x + threadIdx.x, a);
}

}  // namespace image
}  // namespace d3r

using namespace d3r;
using namespace d3r::image;

extern "C" int d3r_image_resize_crop_normalize(const uint8_t* src_dev, int32_t H0, int32_t W0, int32_t H1, int32_t W1,
                                               const int32_t* xbounds_dev, const int32_t* xcoefs_dev, int32_t kx,
                                               const int32_t* ybounds_dev, const int32_t* ycoefs_dev, int32_t ky, int32_t row0,
                                               int32_t rows, int32_t crop_x0, int32_t crop_y0, int32_t H2, int32_t W2,
                                               const float* lut_dev, uint8_t* tmp_dev, float* out_dev, void* stream) {
  D3R_CHECK_ARG(src_dev && xbounds_dev && xcoefs_dev && ybounds_dev && ycoefs_dev && lut_dev && tmp_dev && out_dev,
                "d3r_image_resize_crop_normalize: null pointer");
  D3R_CHECK_ARG(H0 > 0 && W0 > 0 && H1 > 0 && W1 > 0 && kx > 0 && ky > 0 && H2 > 0 && W2 > 0,
                "d3r_image_resize_crop_normalize: sizes must be positive");
  D3R_CHECK_ARG(crop_x0 >= 0 && crop_y0 >= 0 && crop_x0 + W2 <= W1 && crop_y0 + H2 <= H1,
                "d3r_image_resize_crop_normalize: crop (%d, %d) + %d x %d leaves the resized image %d x %d", crop_x0, crop_y0, W2, H2,
                W1, H1);
  D3R_CHECK_ARG(row0 >= 0 && rows > 0 && row0 + rows <= H0, "d3r_image_resize_crop_normalize: source rows [%d, %d) outside [0, %d)",
                row0, row0 + rows, H0);
  const long long n_tmp = (long long)rows * W2, n_out = (long long)H2 * W2;    // pixels = threads of the two launches
  D3R_CHECK_ARG(n_tmp < (1ll << 31) * kThreads && n_out < (1ll << 31) * kThreads, "d3r_image_resize_crop_normalize: image too large");
  cudaStream_t st = (cudaStream_t)stream;
  {
    HorizontalArgs a{src_dev, W0, row0, rows, crop_x0, W2, W1, xbounds_dev, xcoefs_dev, tmp_dev};
    prof::Scope scope("image_resample_h", st, 0.0, 3.0 * (double(rows) * W0 + double(n_tmp)), 1);
    horizontal_kernel<<<(unsigned)((n_tmp + kThreads - 1) / kThreads), kThreads, 0, st>>>(a);
    D3R_LAUNCH_CHECK();
  }
  {
    VerticalArgs a{tmp_dev, row0, W2, H1, ybounds_dev, ycoefs_dev, crop_y0, H2, W2, lut_dev, out_dev};
    prof::Scope scope("image_resample_v", st, 0.0, 3.0 * double(n_tmp) + 12.0 * double(n_out), 1);
    vertical_kernel<<<(unsigned)((n_out + kThreads - 1) / kThreads), kThreads, 0, st>>>(a);
    D3R_LAUNCH_CHECK();
  }
  return D3R_OK;
}
