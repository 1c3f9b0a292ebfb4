// A third-party device target for the plugin boundary of include/hydrobayes_b200.h (hb_target_vtable +
// hb_register_target): what a user of the reference writes instead of an R function passed by name to
// dream(fun = "...", ...) -- get(fun)(x, ...) of R/dream_parallel.R:279-283.  TEST INFRASTRUCTURE: built by
// tests/test_plugin_target.py with nvcc, linked against libhydrobayes_b200.so only for hb_register_target.
//
// Target "gauss_shifted": x -> sum_k log N(x_k; mu_k, 1), mu = the `...` of the R call (a vector of length d; zeros when
// absent).  lik = 2: the function returns a log-likelihood; lik = 1: it returns a likelihood (calc_ll takes the log,
// R/internals.R:4-7).  One warp per point, lanes over the dimensions.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>

#include "hydrobayes_b200.h"

namespace {

struct gauss_ctx {
  int d = 0, lik = 2;
  double* mu_dev = nullptr;
};

__global__ void gauss_kernel(const double* __restrict__ x, long long n, int d, int ldx, const double* __restrict__ mu,
                             int lik, double* __restrict__ ll, double* __restrict__ fx) {
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= n) return;
  double s = 0.0;
  for (int k = lane; k < d; k += 32) { const double z = x[p * ldx + k] - mu[k]; s += z * z; }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const double v = -0.5 * s - 0.5 * d * 1.8378770664093453;   // log(2 pi)
    ll[p] = v;                                                   // before the floor: the sampler applies it (R/internals.R:19)
    if (fx) fx[p] = lik == 1 ? exp(v) : v;                       // the raw output of the target function
  }
}

int gauss_init(void** ctx, int d, int lik, const double* data, const int64_t* dims, int ndims, const double* obs, int64_t n_obs,
               double glue_shape, char* err, int errlen) {
  (void)obs; (void)n_obs; (void)glue_shape;
  if (lik != 1 && lik != 2) { snprintf(err, (size_t)errlen, "gauss_shifted: lik has to be 1 or 2"); return HB_ERR_INVALID; }
  if (ndims > 0 && (ndims != 1 || dims[0] != d)) { snprintf(err, (size_t)errlen, "gauss_shifted: mu must have length d"); return HB_ERR_INVALID; }
  gauss_ctx* c = new gauss_ctx();
  c->d = d; c->lik = lik;
  if (cudaMalloc((void**)&c->mu_dev, sizeof(double) * (size_t)d) != cudaSuccess) { delete c; snprintf(err, (size_t)errlen, "cudaMalloc failed"); return HB_ERR_CUDA; }
  cudaError_t e = ndims > 0 ? cudaMemcpy(c->mu_dev, data, sizeof(double) * (size_t)d, cudaMemcpyHostToDevice)
                            : cudaMemset(c->mu_dev, 0, sizeof(double) * (size_t)d);
  if (e != cudaSuccess) { cudaFree(c->mu_dev); delete c; snprintf(err, (size_t)errlen, "upload of mu failed"); return HB_ERR_CUDA; }
  *ctx = c;
  return HB_OK;
}

int gauss_launch(void* ctx, const double* x_dev, int64_t n, int d, int ldx, double* ll_dev, double* fx_dev, void* stream) {
  const gauss_ctx* c = (const gauss_ctx*)ctx;
  if (n <= 0) return HB_OK;
  const int threads = 256;                                      // 8 points per block
  const long long blocks = (n * 32 + threads - 1) / threads;
  gauss_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(x_dev, n, d, ldx, c->mu_dev, c->lik, ll_dev, fx_dev);
  return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

void gauss_destroy(void* ctx) {
  gauss_ctx* c = (gauss_ctx*)ctx;
  if (!c) return;
  cudaFree(c->mu_dev);
  delete c;
}

void gauss_work(void* ctx, int d, double* flops, double* bytes) { (void)ctx; *flops = 3.0 * d; *bytes = 8.0 * d + 8.0; }

}  // namespace

extern "C" int hb_gauss_plugin_register(void) {
  static const hb_target_vtable vt = {"gauss_shifted", gauss_init, gauss_launch, gauss_destroy, gauss_work, nullptr};
  return hb_register_target(&vt);
}
