// hb_rstat.cu -- K4: the reference's (non-textbook) Gelman-Rubin diagnostic, R/gelman_rubin.R:21-56.
//
//   y        <- x[ceiling(t/2):t, k, ]                         (:37)
//   x_mean_c <- 2/(t-2) * colSums(y)        per chain          (:31,42)  (NOT the true mean)
//   W_k      <- 2/(n(t-2)) * sum_rows sum_c (y - x_mean_c)^2   (:33,37)
//   B_k      <- 1/(2(n-1)) * sum_c (x_mean_c - mean(x_mean))^2 (:44-46)
//   sigma    <- (t-2)/t * W + 2 B ;  R <- sqrt((n+1)/n * sigma/W - (t-2)/(n t))   (:51,54)
//
// Pass 1 (HBM-bound, 2 * 8 * (floor(t/2)+1) * d * n bytes): one thread per (chain, parameter) walks the window
// twice (sum, then squared deviations from the biased mean) with Kahan compensation; consecutive threads read
// consecutive addresses of the device history [row][chain][d].  Pass 2: one block per parameter reduces over chains
// with warp shuffles in a fixed order.
#include "hb_kernels.h"

namespace {

__device__ __forceinline__ void kahan_add(double& s, double& c, double v) {
  const double y = v - c;
  const double t = s + y;
  c = (t - s) - y;
  s = t;
}

// history layout: element (it, chain c, k) at (it * nct_stride + c0 + c) * d + k
__global__ void hb_rstat_pass1_hist(const double* __restrict__ h, long long t_rows, long long nct_stride, long long c0,
                                    long long n, int d, double* __restrict__ xm, double* __restrict__ ss) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * d) return;
  const long long r0 = (t_rows + 1) / 2 - 1;                // ceiling(t/2), 0-based
  const long long row_stride = nct_stride * d;
  const double* base = h + c0 * d + e;
  double s = 0, c = 0;
  for (long long it = r0; it < t_rows; ++it) kahan_add(s, c, base[it * row_stride]);
  const double m = 2.0 / (double)(t_rows - 2) * s;
  double q = 0, cq = 0;
  for (long long it = r0; it < t_rows; ++it) { const double v = base[it * row_stride] - m; kahan_add(q, cq, v * v); }
  xm[e] = m;                                                // [chain][k]
  ss[e] = q;
}

// R layout x[it + t*(k + d*c)]: one warp per (k, c) column, lanes stride over rows, warp-shuffle reduction
__global__ void hb_rstat_pass1_colmajor(const double* __restrict__ x, long long t, int d, long long n,
                                        double* __restrict__ xm, double* __restrict__ ss) {
  const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // = k + d*c
  const int lane = threadIdx.x & 31;
  if (col >= n * d) return;
  const long long r0 = (t + 1) / 2 - 1;
  const double* base = x + col * t;
  double s = 0, c = 0;
  for (long long it = r0 + lane; it < t; it += 32) kahan_add(s, c, base[it]);
  s = hb_warp_sum(s);
  const double m = 2.0 / (double)(t - 2) * s;
  double q = 0, cq = 0;
  for (long long it = r0 + lane; it < t; it += 32) { const double v = base[it] - m; kahan_add(q, cq, v * v); }
  q = hb_warp_sum(q);
  if (lane == 0) {
    const long long c_ = col / d; const int k = (int)(col % d);
    xm[c_ * d + k] = m;
    ss[c_ * d + k] = q;
  }
}

__device__ double block_sum(double v, double* red) {
  v = hb_warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double s = 0;
  if (threadIdx.x == 0) for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  if (threadIdx.x == 0) red[32] = s;
  __syncthreads();
  return red[32];
}

// one block per parameter k: W, B, R
__global__ void hb_rstat_pass2(const double* __restrict__ xm, const double* __restrict__ ss, long long t, long long n,
                               int d, double* __restrict__ out) {
  __shared__ double red[33];
  const int k = blockIdx.x;
  double w = 0, m = 0;
  for (long long c = threadIdx.x; c < n; c += blockDim.x) { w += ss[c * d + k]; m += xm[c * d + k]; }
  const double wsum = block_sum(w, red);
  const double xmm = block_sum(m, red) / (double)n;
  double b = 0;
  for (long long c = threadIdx.x; c < n; c += blockDim.x) { const double e = xm[c * d + k] - xmm; b += e * e; }
  const double bsum = block_sum(b, red);
  if (threadIdx.x == 0) {
    const double W = 2.0 / ((double)n * (double)(t - 2)) * wsum;
    const double B = 1.0 / (2.0 * (double)(n - 1)) * bsum;
    const double sigma = (double)(t - 2) / (double)t * W + 2.0 * B;
    out[k] = sqrt((double)(n + 1) / (double)n * sigma / W - (double)(t - 2) / ((double)n * (double)t));
  }
}

// one warp per (run, parameter): W, B, R over the run's nc chains, fixed shuffle order
__global__ void hb_rstat_pass2_runs(const double* __restrict__ xm, const double* __restrict__ ss, long long t, long long nc,
                                    long long n_runs, int d, double* __restrict__ out) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= n_runs * d) return;
  const long long run = wid / d; const int k = (int)(wid % d);
  const double* xmr = xm + run * nc * d; const double* ssr = ss + run * nc * d;
  double w = 0, m = 0;
  for (long long c = lane; c < nc; c += 32) { w += ssr[c * d + k]; m += xmr[c * d + k]; }
  const double wsum = hb_warp_sum(w);
  const double xmm = hb_warp_sum(m) / (double)nc;
  double b = 0;
  for (long long c = lane; c < nc; c += 32) { const double e = xmr[c * d + k] - xmm; b += e * e; }
  const double bsum = hb_warp_sum(b);
  if (lane == 0) {
    const double n = (double)nc;
    const double W = 2.0 / (n * (double)(t - 2)) * wsum;
    const double B = 1.0 / (2.0 * (n - 1)) * bsum;
    const double sigma = (double)(t - 2) / (double)t * W + 2.0 * B;
    out[wid] = sqrt((n + 1) / n * sigma / W - (double)(t - 2) / (n * (double)t));
  }
}

// ---- sharded population: the chains of one R_stat live on several GPUs --------------------------------------------
// local sums per parameter: out[k] = sum_c ss, out[d + k] = sum_c x_mean_c  (all-reduced by the caller)
__global__ void hb_rstat_shard_sums(const double* __restrict__ xm, const double* __restrict__ ss, long long n, int d,
                                    double* __restrict__ out) {
  __shared__ double red[33];
  const int k = blockIdx.x;
  double w = 0, m = 0;
  for (long long c = threadIdx.x; c < n; c += blockDim.x) { w += ss[c * d + k]; m += xm[c * d + k]; }
  const double wsum = block_sum(w, red);
  const double msum = block_sum(m, red);
  if (threadIdx.x == 0) { out[k] = wsum; out[d + k] = msum; }
}
// local sum of squared deviations from the GLOBAL mean of the chain means (second pass, all-reduced by the caller)
__global__ void hb_rstat_shard_dev(const double* __restrict__ xm, long long n, int d, long long n_total,
                                   const double* __restrict__ sums, double* __restrict__ out) {
  __shared__ double red[33];
  const int k = blockIdx.x;
  const double xmm = sums[d + k] / (double)n_total;
  double b = 0;
  for (long long c = threadIdx.x; c < n; c += blockDim.x) { const double e = xm[c * d + k] - xmm; b += e * e; }
  const double bsum = block_sum(b, red);
  if (threadIdx.x == 0) out[k] = bsum;
}
__global__ void hb_rstat_shard_final(const double* __restrict__ sums, const double* __restrict__ bsum, long long t,
                                     long long n, int d, double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d) return;
  const double W = 2.0 / ((double)n * (double)(t - 2)) * sums[k];
  const double B = 1.0 / (2.0 * (double)(n - 1)) * bsum[k];
  const double sigma = (double)(t - 2) / (double)t * W + 2.0 * B;
  out[k] = sqrt((double)(n + 1) / (double)n * sigma / W - (double)(t - 2) / ((double)n * (double)t));
}

}  // namespace

cudaError_t hb_launch_rstat_shard_pass1(const double* hist_x, long long t_rows, long long n_local, int d, double* xm,
                                        double* ss, double* sums2d, cudaStream_t st) {
  const long long tot = n_local * d;
  hb_rstat_pass1_hist<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(hist_x, t_rows, n_local, 0, n_local, d, xm, ss);
  hb_rstat_shard_sums<<<d, 256, 0, st>>>(xm, ss, n_local, d, sums2d);
  return cudaGetLastError();
}
cudaError_t hb_launch_rstat_shard_dev(const double* xm, long long n_local, int d, long long n_total, const double* sums2d,
                                      double* bsum, cudaStream_t st) {
  hb_rstat_shard_dev<<<d, 256, 0, st>>>(xm, n_local, d, n_total, sums2d, bsum);
  return cudaGetLastError();
}
cudaError_t hb_launch_rstat_shard_final(const double* sums2d, const double* bsum, long long t_rows, long long n_total, int d,
                                        double* out_d, cudaStream_t st) {
  hb_rstat_shard_final<<<(d + 127) / 128, 128, 0, st>>>(sums2d, bsum, t_rows, n_total, d, out_d);
  return cudaGetLastError();
}

cudaError_t hb_launch_rstat_runs(const double* hist_x, long long t_rows, long long nct, long long nc, long long n_runs,
                                 int d, double* xm_scratch, double* ss_scratch, double* out_runs_d, cudaStream_t st) {
  const long long tot = nct * d;
  hb_rstat_pass1_hist<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(hist_x, t_rows, nct, 0, nct, d, xm_scratch, ss_scratch);
  const long long warps = n_runs * d;
  hb_rstat_pass2_runs<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(xm_scratch, ss_scratch, t_rows, nc, n_runs, d, out_runs_d);
  return cudaGetLastError();
}

cudaError_t hb_launch_rstat(const double* hist_x, long long t_rows, long long nct_stride, long long c0, long long n,
                            int d, double* xm_scratch, double* ss_scratch, double* out_d, cudaStream_t st) {
  const long long tot = n * d;
  hb_rstat_pass1_hist<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(hist_x, t_rows, nct_stride, c0, n, d, xm_scratch,
                                                                      ss_scratch);
  hb_rstat_pass2<<<d, 256, 0, st>>>(xm_scratch, ss_scratch, t_rows, n, d, out_d);
  return cudaGetLastError();
}

cudaError_t hb_launch_rstat_colmajor(const double* x, long long t, int d, long long n, double* xm_scratch,
                                     double* ss_scratch, double* out_d, cudaStream_t st) {
  const long long warps = n * d;
  hb_rstat_pass1_colmajor<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(x, t, d, n, xm_scratch, ss_scratch);
  hb_rstat_pass2<<<d, 256, 0, st>>>(xm_scratch, ss_scratch, t, n, d, out_d);
  return cudaGetLastError();
}
