// hb_outlier.cu -- device side of check_outlier (R/internals.R:71-87) for one large population:
//   proxy -> sorted copy (CUB radix sort, a utility: this runs at most adapt*t/updateInterval times per run)
//   -> type-7 quartiles (stats::quantile default) -> threshold Q1 - 2|Q3 - Q1| -> ascending list of outliers.
// Only the count (and, when it is non-zero, the short list) travels to the host, which draws the replacement
// chains without replacement (tape or Philox) exactly like R's sample().
#include <cub/cub.cuh>

#include "hb_kernels.h"

namespace {

__global__ void hb_quartile_thr_kernel(const double* __restrict__ s, long long n, double* thr) {
  // stats::quantile type 7 (recollection): index = 1 + (n-1) p ; (1-h) x[lo] + h x[hi]
  double q[2];
  const double pr[2] = {0.25, 0.75};
  for (int i = 0; i < 2; ++i) {
    const double index = 1 + (double)(n > 1 ? n - 1 : 0) * pr[i];
    const double lo = floor(index), hi = ceil(index);
    double qs = s[(long long)lo - 1];
    const double xh = s[(long long)hi - 1];
    if (index > lo && xh != qs) { const double h = index - lo; qs = (1 - h) * qs + h * xh; }
    q[i] = qs;
  }
  thr[0] = q[0] - 2 * fabs(q[1] - q[0]);   // R/internals.R:76-78
}

struct below_thr {
  const double* proxy;
  const double* thr;
  __device__ bool operator()(const int& j) const { return proxy[j] < *thr; }
};

}  // namespace

size_t hb_outlier_workspace_bytes(long long nc) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, a, (const double*)nullptr, (double*)nullptr, (int)nc);
  cub::CountingInputIterator<int> it(0);
  below_thr pred{nullptr, nullptr};
  cub::DeviceSelect::If(nullptr, b, it, (int*)nullptr, (int*)nullptr, (int)nc, pred);
  return a > b ? a : b;
}

cudaError_t hb_outlier_device(const double* proxy, long long nc, void* tmp, size_t tmp_bytes, double* sorted,
                              double* thr, int* list, int* count, cudaStream_t st) {
  cudaError_t e = cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, proxy, sorted, (int)nc, 0, 64, st);
  if (e != cudaSuccess) return e;
  hb_quartile_thr_kernel<<<1, 1, 0, st>>>(sorted, nc, thr);
  cub::CountingInputIterator<int> it(0);
  below_thr pred{proxy, thr};
  e = cub::DeviceSelect::If(tmp, tmp_bytes, it, list, count, (int)nc, pred, st);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}
