// Host-side helpers of the C ABI: version, error text, Gaussian taps.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "lhn_common.cuh"

namespace lhn {

static thread_local char g_err[256] = "";

const char* set_last_error(cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return g_err;
}

int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error(e); return LHN_ECUDA; }
  return LHN_OK;
}

}  // namespace lhn

extern "C" int lhn_version(void) { return 100; }

extern "C" const char* lhn_last_cuda_error(void) { return lhn::g_err; }

// cv2.getGaussianKernel(ksize, sigma<=0, CV_64F): fixed tables for ksize<=7, else
// sigma = 0.3*((ksize-1)*0.5-1)+0.8, taps exp(-(i-c)^2/(2 sigma^2)) normalised in double.
extern "C" int lhn_gaussian_taps(int ksize, double* taps) {
  if (!taps || ksize < 1 || ksize > LHN_MAX_TAPS || (ksize & 1) == 0) return LHN_EINVAL;
  static const double small_tab[4][7] = {
      {1.0},
      {0.25, 0.5, 0.25},
      {0.0625, 0.25, 0.375, 0.25, 0.0625},
      {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125}};
  if (ksize <= 7) {
    for (int i = 0; i < ksize; ++i) taps[i] = small_tab[ksize >> 1][i];
    return LHN_OK;
  }
  const double sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8;
  const double scale2x = -0.5 / (sigma * sigma);
  const double c = (ksize - 1) * 0.5;
  double sum = 0.0;
  for (int i = 0; i < ksize; ++i) { double x = i - c; taps[i] = exp(scale2x * x * x); sum += taps[i]; }
  for (int i = 0; i < ksize; ++i) taps[i] = taps[i] / sum;
  return LHN_OK;
}
