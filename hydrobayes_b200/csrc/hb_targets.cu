// hb_targets.cu -- K2: batched per-chain log-density plugins and the name-keyed registry that replaces
// get(fun)(x, ...) (R/dream_parallel.R:219-227, 278-287) for the package's own test targets.
//
//   "mixture"         R/test_mixture.R:16            fused elementwise kernel (density, then calc_ll lik = 1 -> log)
//   "t_distribution"  R/test_t_distribution.R:19-46  Z = Xp * W (W = chol(C)^-1, upper triangular) on fp64 tensor
//                                                    cores (mma.sync m8n8k4 f64 -> SASS DMMA), rss = rowsum(Z o Z),
//                                                    const - (df+d)/2 * log1p(rss/df) in the epilogue; Z never stored
//   "HydroModel"      R/test_HydroModel.R:17-36      one chain per thread, forcing + obs staged into shared memory by
//                                                    a 1-D bulk TMA copy (cp.async.bulk + mbarrier), true fp64
//                                                    division per step, GLUE likelihood (R/internals.R:13-14) folded
//
// Plugins return the value BEFORE the -708.396 floor; the sampler (K3) applies calc_ll's floor and finiteness check.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/hydrobayes_b200.h"
#include "hb_common.cuh"
#include "hb_kernels.h"

namespace {

void set_err(char* err, int errlen, const char* msg) {
  if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", msg);
}

// =========================================== mixture ======================================================
// R's dnorm4 (nmath/dnorm.c, recollection): plain formula for |z| < 5, split exponent beyond, 0 past ~38.57
__device__ __forceinline__ double dnorm_r(double x, double mu) {
  const double M_1_SQRT_2PI_ = 0.398942280401432677939946059934;
  double z = fabs(x - mu);
  if (!isfinite(z)) return 0.0;
  if (z < 5.0) return M_1_SQRT_2PI_ * exp(-0.5 * z * z);
  if (z > 38.56804181549334) return 0.0;   // sqrt(-2*M_LN2*(DBL_MIN_EXP + 1 - DBL_MANT_DIG))
  const double x1 = rint(z * 65536.0) * (1.0 / 65536.0);
  const double x2 = z - x1;
  return M_1_SQRT_2PI_ * (exp(-0.5 * x1 * x1) * exp((-0.5 * x2 - x1) * x2));
}

__global__ void hb_mixture_kernel(const double* __restrict__ x, long long n, int ldx, int lik, double* __restrict__ ll,
                                  double* __restrict__ fx) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double v = x[j * ldx];
  const double dens = 1.0 / 6.0 * dnorm_r(v, -8.0) + 5.0 / 6.0 * dnorm_r(v, 10.0);   // R/test_mixture.R:16
  if (fx) fx[j] = dens;
  ll[j] = (lik == 1) ? log(dens) : dens;                                               // R/internals.R:4-7
}

struct mixture_ctx { int lik; };

int mixture_init(void** ctx, int d, int lik, const double*, const int64_t*, int, const double*, int64_t, double,
                 char* err, int errlen) {
  if (d != 1) { set_err(err, errlen, "mixture() is a 1-d target (R/test_mixture.R:15)"); return HB_ERR_INVALID; }
  if (lik != 1 && lik != 2) { set_err(err, errlen, "mixture: lik must be 1 (density) or 2"); return HB_ERR_UNSUPPORTED; }
  auto* c = new mixture_ctx{lik};
  *ctx = c;
  return HB_OK;
}
int mixture_launch(void* ctx, const double* x, int64_t n, int d, int ldx, double* ll, double* fx, void* stream) {
  auto* c = (mixture_ctx*)ctx;
  if (n <= 0) return HB_OK;
  hb_mixture_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, ldx, c->lik, ll, fx);
  return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}
void mixture_destroy(void* ctx) { delete (mixture_ctx*)ctx; }
void mixture_work(void*, int, double* flops, double* bytes) { *flops = 40; *bytes = 16; }

// =========================================== t_distribution ===============================================
// KB k-blocks of 4, NB = ceil(KB/2) n-blocks of 8.  One warp owns 8 chains per iteration:
//   A (8x4, row): lane holds X[row = lane>>2][k = 4*kb + (lane&3)]   -- all KB fragments loaded up front (MLP = KB)
//   B (4x8, col): lane holds W[k = 4*kb + (lane&3)][n = 8*nb + (lane>>2)] from shared memory
//   C (8x8): lane holds row lane>>2, two columns -> squared and reduced over the 4 lanes of a row group.
// W is upper triangular (z_n = sum_{k<=n} x_k W[k][n]), so n-block nb only needs k-blocks kb <= 2*nb+1: the fully
// unrolled loop nest drops the others at compile time (about d^2/2 multiply-adds executed, as backsolve does).
// W is stored BLOCK-TRIANGULAR in fragment order: only the (nb, kb) blocks that are used, 32 doubles each, element
// `lane` of a block is the B fragment of that lane -- a warp's B load is one contiguous, conflict-free 256-byte line, and
// the whole operand is 46 KB instead of 91 KB at d = 100, which is what lets three CTAs (24 warps) share an SM: the
// kernel is bound by the latency of the 25 global loads that open every tile, hidden only by other warps.
template <int KB>
__global__ void __launch_bounds__(256, (KB > 25 ? 2 : 3)) hb_tdist_dmma_kernel(const double* __restrict__ X, long long n, int d, int ldx,
                                                            const double* __restrict__ Wblk, double konst,
                                                            double df, int lik, double* __restrict__ ll,
                                                            double* __restrict__ fx) {
  constexpr int NB = (KB + 1) / 2;
  extern __shared__ __align__(16) double Ws[];
  __shared__ __align__(8) uint64_t wbar;
  // the block-ordered W (a multiple of 256 bytes) is staged by bulk TMA copies of at most 32 KB each
  constexpr uint32_t w_bytes = (uint32_t)tdist_n_blocks(KB) * 32u * 8u;
  if (threadIdx.x == 0) {
    mbar_init(&wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&wbar, w_bytes);
    for (uint32_t off = 0; off < w_bytes; off += 32768u) {
      const uint32_t nbytes = (w_bytes - off) < 32768u ? (w_bytes - off) : 32768u;
      tma_load_1d(reinterpret_cast<char*>(Ws) + off, reinterpret_cast<const char*>(Wblk) + off, nbytes, &wbar);
    }
  }
  mbar_wait(&wbar, 0);
  const int lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long n_tiles = (n + 7) / 8;
  for (long long tile = warp_id; tile < n_tiles; tile += n_warps) {
    const long long row = tile * 8 + gid;
    const bool rv = row < n;
    const double* xr = X + (rv ? row : 0) * (long long)ldx;
    double a[KB];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      const int k = 4 * kb + tig;
      a[kb] = (rv && k < d) ? xr[k] : 0.0;
    }
    double rss = 0.0;
    // four n-blocks at a time: four independent accumulator chains keep the DMMA pipe busy
#pragma unroll
    for (int g0 = 0; g0 < NB; g0 += 4) {
      double c[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int nb = g0 + i;
          if (nb < NB && kb <= 2 * nb + 1) {
            const double b = Ws[(tdist_blk_base(nb, KB) + kb) * 32 + lane];
            dmma884(c[i][0], c[i][1], a[kb], b);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rss = fma(c[i][0], c[i][0], rss);
        rss = fma(c[i][1], c[i][1], rss);
      }
    }
    rss += __shfl_xor_sync(0xffffffffu, rss, 1);
    rss += __shfl_xor_sync(0xffffffffu, rss, 2);
    if (rv && tig == 0) {
      const double v = konst - 0.5 * (df + (double)d) * log1p(rss / df);   // dmvt, R/test_t_distribution.R:43
      if (fx) fx[row] = v;
      ll[row] = (lik == 1) ? log(v) : v;
    }
  }
}

// Structured evaluation (the default).  The reference's covariance is C = S (I/2 + 11'/2) S with S = diag(sqrt(i))
// (R/test_t_distribution.R:32-40: C[i,j] = A[i,j] sqrt(i j), A = 0.5 I + 0.5), so by Sherman-Morrison
//   C^-1 = 2 S^-1 (I - 11'/(d + 1)) S^-1   and   x' C^-1 x = 2 (sum u_i^2 - (sum u_i)^2 / (d + 1)),  u_i = x_i / sqrt(i):
// the quadratic form dmvt obtains from chol(C) and a triangular solve (d^2 + 3 d flops per point) costs 4 d flops, and the
// kernel is a stream over the proposals (808 bytes per point) instead of a d x d product.  Same value to rounding: against
// the long-double restatement of dmvt the log-density agrees to 5e-16 relative at d = 100 (tests: every plugin test runs both
// evaluations).  Cancellation is bounded: (sum u)^2 <= d sum u^2, so the result is at least 2 sum u^2 / (d + 1).
// rs[i] = 1 / sqrt(i + 1), correctly rounded on the host.  One warp per point; hb_tdist_acc / hb_tdist_logdens (hb_common.cuh)
// are shared with the proposal kernels that fold the density into their own sweep (hb_k1_propose.cu) and fix the order of
// the sums, so the value does not depend on the kernel that produced it.
__global__ void hb_tdist_struct_kernel(const double* __restrict__ X, long long n, int d, int ldx, const double* __restrict__ rs,
                                       double konst, double df, int lik, int batch, double* __restrict__ ll, double* __restrict__ fx) {
  // a warp takes `batch` (<= 32) consecutive points: the sums of point c stay with lane c, so the log1p of the batch runs
  // lane-parallel and ll leaves as one coalesced store (one point per warp would spend most of its time in lane 0's log1p)
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long base = warp_id * batch; base < n; base += n_warps * batch) {
    const int nb = (int)((n - base) < batch ? (n - base) : batch);
    double m1 = 0.0, m2 = 0.0;
    for (int c = 0; c < nb; ++c) {
      const double* xr = X + (base + c) * (long long)ldx;
      double s1 = 0.0, s2 = 0.0;
      for (int k0 = 2 * lane; k0 < d; k0 += 64) {                          // the fixed order of hb_tdist_acc: pairs 2 (32 m + lane)
        const bool has1 = k0 + 1 < d;
        hb_tdist_acc(s1, s2, xr[k0], has1 ? xr[k0 + 1] : 0.0, rs[k0], has1 ? rs[k0 + 1] : 0.0);
      }
      hb_warp_sum2(s1, s2, lane);
      if (lane == c) { m1 = s1; m2 = s2; }
    }
    if (lane < nb) {
      const double v = hb_tdist_logdens(m1, m2, d, konst, df);             // dmvt, R/test_t_distribution.R:43
      if (fx) fx[base + lane] = v;
      ll[base + lane] = (lik == 1) ? log(v) : v;
    }
  }
}

// generic FMA path for d > 128 (and the cross-check of the DMMA kernel in the tests): one warp per chain
__global__ void hb_tdist_fma_kernel(const double* __restrict__ X, long long n, int d, int ldx,
                                    const double* __restrict__ Wg, int ldw, double konst, double df, int lik,
                                    double* __restrict__ ll, double* __restrict__ fx) {
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long row = warp_id; row < n; row += n_warps) {
    const double* xr = X + row * (long long)ldx;
    double rss = 0.0;
    for (int nn = lane; nn < d; nn += 32) {
      double z = 0.0;
      for (int k = 0; k <= nn; ++k) z = fma(xr[k], Wg[k * ldw + nn], z);
      rss = fma(z, z, rss);
    }
    rss = hb_warp_sum(rss);
    if (lane == 0) {
      const double v = konst - 0.5 * (df + (double)d) * log1p(rss / df);
      if (fx) fx[row] = v;
      ll[row] = (lik == 1) ? log(v) : v;
    }
  }
}

struct tdist_ctx {
  int d, lik, KB, ldw;
  double konst, df;
  double* W_dev;       // [kp][ldw] dense upper triangle (FMA path)
  double* Wblk_dev;    // block-triangular fragment order (DMMA path), tdist_n_blocks(KB) x 32 doubles
  double* rs_dev;      // [d] 1 / sqrt(i): the structured evaluation
  int sm_count;
  int force_fma;
  int dense;           // HB_TDIST_DENSE=1: dmvt's own route (triangular product on the fp64 tensor cores) instead of the structured one
};

// C = (0.5 I + 0.5) o sqrt(i j); R = chol(C) (upper); W = R^-1; all in long double, rounded once
int tdist_build(int d, std::vector<double>& W, int kp, int ldw, double* konst) {
  const double df = 60.0;
  std::vector<long double> C((size_t)d * d), R((size_t)d * d, 0.0L), Wi((size_t)d * d, 0.0L);
  for (int i = 1; i <= d; ++i)
    for (int j = 1; j <= d; ++j) C[(size_t)(i - 1) * d + (j - 1)] = (i == j ? 1.0L : 0.5L) * sqrtl((long double)i * j);
  for (int j = 0; j < d; ++j) {
    long double s = C[(size_t)j * d + j];
    for (int k = 0; k < j; ++k) s -= R[(size_t)k * d + j] * R[(size_t)k * d + j];
    if (!(s > 0)) return -1;
    const long double rjj = sqrtl(s);
    R[(size_t)j * d + j] = rjj;
    for (int i = j + 1; i < d; ++i) {
      long double v = C[(size_t)j * d + i];
      for (int k = 0; k < j; ++k) v -= R[(size_t)k * d + j] * R[(size_t)k * d + i];
      R[(size_t)j * d + i] = v / rjj;
    }
  }
  // W = R^-1 (upper): back substitution column by column
  for (int j = 0; j < d; ++j) {
    Wi[(size_t)j * d + j] = 1.0L / R[(size_t)j * d + j];
    for (int i = j - 1; i >= 0; --i) {
      long double s = 0;
      for (int k = i + 1; k <= j; ++k) s += R[(size_t)i * d + k] * Wi[(size_t)k * d + j];
      Wi[(size_t)i * d + j] = -s / R[(size_t)i * d + i];
    }
  }
  W.assign((size_t)kp * ldw, 0.0);
  for (int k = 0; k < d; ++k)
    for (int nn = k; nn < d; ++nn) W[(size_t)k * ldw + nn] = (double)Wi[(size_t)k * d + nn];
  long double sl = 0;
  for (int i = 0; i < d; ++i) sl += logl(R[(size_t)i * d + i]);
  *konst = lgamma((d + df) / 2) - (lgamma(df / 2) + (double)sl + d / 2.0 * log(3.141592653589793238462643383279502884 * df));
  return 0;
}

const int kKB[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 13, 16, 20, 25, 32};

int tdist_init(void** ctx, int d, int lik, const double*, const int64_t*, int, const double*, int64_t, double, char* err,
               int errlen) {
  if (lik != 2 && lik != 1) { set_err(err, errlen, "t_distribution returns a log-density: use lik = 2"); return HB_ERR_UNSUPPORTED; }
  auto* c = new tdist_ctx();
  c->d = d; c->lik = lik; c->df = 60.0; c->force_fma = 0;
  const int kb_need = (d + 3) / 4;
  c->KB = 0;
  for (int v : kKB) if (v >= kb_need) { c->KB = v; break; }
  const int kp = c->KB ? 4 * c->KB : d;
  int np = c->KB ? 8 * ((c->KB + 1) / 2) : d;
  // B-fragment read of a half-warp (64-bit shared loads are served 16 lanes at a time): lane (gid, tig), gid < 4, reads
  // double (4kb + tig)*ldw + 8nb + gid -> 16 distinct 8-byte banks iff ldw == 4 (mod 16).  (ldw == 8 (mod 16), the
  // first layout, put tig = 0, 2 and tig = 1, 3 on the same banks: ncu counted a 2-way conflict on every B load.)
  c->ldw = np + ((4 - np % 16) + 16) % 16;
  if (!c->KB) c->ldw = d;
  std::vector<double> W;
  if (tdist_build(d, W, kp, c->ldw, &c->konst)) { delete c; set_err(err, errlen, "chol(C) failed"); return HB_ERR_INVALID; }
  if (cudaMalloc(&c->W_dev, W.size() * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(c->W_dev, W.data(), W.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete c; set_err(err, errlen, "cudaMalloc/cudaMemcpy failed for W"); return HB_ERR_CUDA;
  }
  c->Wblk_dev = nullptr;
  if (c->KB) {   // block-triangular fragment order: block (nb, kb <= 2nb+1), element lane = W[4kb + (lane&3)][8nb + (lane>>2)]
    const int KB = c->KB, NB = (KB + 1) / 2;
    std::vector<double> Wb;
    for (int nb = 0; nb < NB; ++nb)
      for (int kb = 0; kb < KB && kb <= 2 * nb + 1; ++kb)
        for (int lane = 0; lane < 32; ++lane) {
          const int k = 4 * kb + (lane & 3), nn = 8 * nb + (lane >> 2);
          Wb.push_back((k < kp && nn < c->ldw) ? W[(size_t)k * c->ldw + nn] : 0.0);
        }
    if (cudaMalloc(&c->Wblk_dev, Wb.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(c->Wblk_dev, Wb.data(), Wb.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
      delete c; set_err(err, errlen, "cudaMalloc/cudaMemcpy failed for W blocks"); return HB_ERR_CUDA;
    }
  }
  int dev = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev);
  const char* f = getenv("HB_TDIST_FORCE_FMA");
  if (f && f[0] == '1') c->force_fma = 1;
  f = getenv("HB_TDIST_DENSE");
  c->dense = (f && f[0] == '1') || c->force_fma;
  {
    std::vector<double> rs((size_t)d);
    for (int i = 0; i < d; ++i) rs[(size_t)i] = 1.0 / sqrt((double)(i + 1));      // IEEE sqrt and division: correctly rounded twice
    c->rs_dev = nullptr;
    if (cudaMalloc(&c->rs_dev, rs.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(c->rs_dev, rs.data(), rs.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(c->W_dev); cudaFree(c->Wblk_dev); cudaFree(c->rs_dev); delete c; set_err(err, errlen, "cudaMalloc/cudaMemcpy failed for the scale vector"); return HB_ERR_CUDA;
    }
  }
  *ctx = c;
  return HB_OK;
}

template <int KB>
int tdist_launch_kb(tdist_ctx* c, const double* x, int64_t n, int ldx, double* ll, double* fx, cudaStream_t st) {
  const size_t smem = (size_t)tdist_n_blocks(KB) * 32 * sizeof(double);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    if (cudaFuncSetAttribute(hb_tdist_dmma_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return HB_ERR_CUDA;
    attr_set = true;
  }
  const long long tiles = (n + 7) / 8;
  const int threads = tiles >= (long long)c->sm_count * 16 ? 256 : 128;   // small populations: spread over more SMs
  const int wpb = threads / 32;
  long long blocks = (tiles + wpb - 1) / wpb;
  long long per_sm = (long long)(224 * 1024 / (smem + 1024));
  if (per_sm > (KB > 25 ? 2 : 3)) per_sm = (KB > 25 ? 2 : 3);  // __launch_bounds__ of the kernel
  if (per_sm < 1) per_sm = 1;
  const long long cap = (long long)c->sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  hb_tdist_dmma_kernel<KB><<<(unsigned)blocks, threads, smem, st>>>(x, n, c->d, ldx, c->Wblk_dev, c->konst, c->df,
                                                                c->lik, ll, fx);
  return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

int tdist_launch(void* ctx, const double* x, int64_t n, int d, int ldx, double* ll, double* fx, void* stream) {
  auto* c = (tdist_ctx*)ctx;
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return HB_OK;
  if (!c->dense) {
    const long long resident_warps = (long long)c->sm_count * 64;            // points per warp: as many as keep every warp of the device busy
    long long batch = n / resident_warps;
    batch = batch < 1 ? 1 : batch > 32 ? 32 : batch;
    const long long warps = (n + batch - 1) / batch;
    long long blocks = (warps + 7) / 8;
    const long long cap = (long long)c->sm_count * 8;
    if (blocks > cap) blocks = cap;
    hb_tdist_struct_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, c->d, ldx, c->rs_dev, c->konst, c->df, c->lik, (int)batch, ll, fx);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
  }
  if (!c->KB || c->force_fma) {
    long long blocks = (n + 3) / 4;
    const long long cap = (long long)c->sm_count * 16;
    if (blocks > cap) blocks = cap;
    hb_tdist_fma_kernel<<<(unsigned)blocks, 128, 0, st>>>(x, n, c->d, ldx, c->W_dev, c->ldw, c->konst, c->df, c->lik, ll, fx);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
  }
  switch (c->KB) {
#define HB_KB_CASE(K) case K: return tdist_launch_kb<K>(c, x, n, ldx, ll, fx, st);
    HB_KB_CASE(1) HB_KB_CASE(2) HB_KB_CASE(3) HB_KB_CASE(4) HB_KB_CASE(5) HB_KB_CASE(6) HB_KB_CASE(7) HB_KB_CASE(8)
    HB_KB_CASE(10) HB_KB_CASE(13) HB_KB_CASE(16) HB_KB_CASE(20) HB_KB_CASE(25) HB_KB_CASE(32)
#undef HB_KB_CASE
  }
  return HB_ERR_INVALID;
}
void tdist_destroy(void* ctx) {
  auto* c = (tdist_ctx*)ctx;
  if (c) { cudaFree(c->W_dev); cudaFree(c->Wblk_dev); cudaFree(c->rs_dev); delete c; }
}
void tdist_work(void* ctx, int d, double* flops, double* bytes) {
  const auto* c = (const tdist_ctx*)ctx;
  *flops = (c && !c->dense) ? 4.0 * d : (double)d * d + 3.0 * d;     // structured: u, sum u, sum u^2 ; dense: triangular solve as backsolve
  *bytes = 8.0 * d + 8;
}

}  // namespace

// operands of the proposal kernels that fold the log-density into their sweep (hb_k1_propose.cu) when the handle's target is
// this plugin in its structured evaluation
bool hb_tdist_fused_info(const hb_target_vtable* vt, void* ctx, hb_fused_tdist* out) {
  if (!vt || !ctx || vt->launch != tdist_launch) return false;
  const auto* c = (const tdist_ctx*)ctx;
  if (c->dense) return false;                           // HB_TDIST_DENSE=1: dmvt's own route, stand-alone kernels only
  out->rs = c->rs_dev; out->lik = c->lik; out->konst = c->konst; out->df = c->df;
  return true;
}

namespace {

// =========================================== HydroModel ====================================================
// a / b for a fixed divisor b with y = RN(1/b) precomputed by a true IEEE division: q0 = RN(a y), r = a - b q0
// (exact, one FMA), q = RN(q0 + r y).  By Markstein's theorem (correctly rounded reciprocal, faithful q0) q is the
// correctly rounded quotient, i.e. bit-identical to R's `/` -- the bit-exact series test against the oracle and the
// randomised check in tests/test_gpu_parity.py::test_division_by_reciprocal_is_exact pin this.
__device__ __forceinline__ double div_by_recip(double a, double b, double y) {
  const double q0 = __dmul_rn(a, y);
  const double r = __fma_rn(-b, q0, a);
  return __fma_rn(r, y, q0);
}

// Sum of squared residuals over the series.  ncu (profiles/r02_c3_hydro_full.txt) shows the HydroModel kernels bound by
// the fp64 pipe (74 % of its cycles active, 11 fp64 instructions per time step), so the sum is kept cheap: FOUR running
// sums, term t goes to sum t mod 4 as one FMA (e*e + s), folded at the end as (s0 + s1) + (s2 + s3).  A Kahan sum cost four
// dependent fp64 instructions per term; four plain sums of ~900 positive terms each stay within ~1e-14 relative of the
// oracle's long-double sum (the tests hold 1e-12).  Every HydroModel kernel uses the same slots, so the single-run and the
// batched kernels return identical values for the same series.
struct hydro_sum {
  double s[4];
  __device__ __forceinline__ hydro_sum() { s[0] = 0.0; s[1] = 0.0; s[2] = 0.0; s[3] = 0.0; }
  template <int R> __device__ __forceinline__ void add_sq(double e) { s[R] = __fma_rn(e, e, s[R]); }
  __device__ __forceinline__ double total() const { return (s[0] + s[1]) + (s[2] + s[3]); }
};

// One chain per thread.  P and obs (T_pad doubles each, T_pad even so that the byte count is a multiple of 16) are
// staged once per CTA.  Per step (R/test_HydroModel.R:30-32): S = (P_i*Dt + S)/(1 + K*Dt), Y_i = K*S with Dt = 1, a
// true IEEE division.  The likelihood flag is folded into the sweep as an epilogue (calc_ll, R/internals.R:3-23):
//   31  glue_shape*log(max(1 - sse/sst, 0))                                   (:13-14)   sse: hydro_sum
//   21  -m/2 log(2 pi) - sum(log(abc_e)) - 0.5 sum((abc_rho(sim,obs)/abc_e)^2)  (:8-10)  abc_rho = "abc_distfun" = |sim-obs|
//   22  min(abc_e - abc_rho(sim,obs))                                         (:11-12)
//   99  lik_fun(sim, obs), lik_fun = "fun_lik" = -n/2 log(sum((sim-obs)^2))    (:15-16; examples/script_examples.R:503-507)
// EVEC: abc_e is a vector of length T (staged as a third series) instead of a recycled scalar.
template <int LIK, bool EVEC>
__global__ void __launch_bounds__(256) hb_hydro_kernel(const double* __restrict__ X, long long n, int ldx,
                                                       const double* __restrict__ forcing,
                                                       const double* __restrict__ obs, const double* __restrict__ abc_e,
                                                       int T, int T_pad, double sst, double glue_shape, double e_scalar,
                                                       double konst, double* __restrict__ ll, unsigned* err) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t bar;
  double* sP = sm;
  double* sO = sm + T_pad;
  double* sE = sm + 2 * T_pad;
  const uint32_t bytes = (uint32_t)T_pad * 8u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, (EVEC ? 3 : 2) * bytes);
    tma_load_1d(sP, forcing, bytes, &bar);
    tma_load_1d(sO, obs, bytes, &bar);
    if (EVEC) tma_load_1d(sE, abc_e, bytes, &bar);
  }
  mbar_wait(&bar, 0);
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double K = X[j * ldx];
  if (K < 0) { atomicOr(err, HB_FLAG_TARGET_DOMAIN); ll[j] = NAN; return; }   // R/test_HydroModel.R:18-19
  const double den = 1.0 + K * 1.0;
  const double rden = __ddiv_rn(1.0, den);
  double S = 1.0, vmin = INFINITY;
  hydro_sum acc;
  auto step = [&](int i, auto slot) {
    constexpr int R = decltype(slot)::value;
    S = div_by_recip(__dadd_rn(sP[i], S), den, rden);         // (forcing[i]*Dt + S0) / (1 + param*Dt), Dt = 1
    const double e = __fma_rn(K, S, -sO[i]);                  // Y[i] - obs[i]  (one rounding instead of two: <= 1 ulp of Y)
    if (LIK == 22) {
      const double v = (EVEC ? sE[i] : e_scalar) - fabs(e);   // abc_e - rho
      if (v < vmin || v != v) vmin = v;
    } else if (LIK == 21) {
      acc.template add_sq<R>(__ddiv_rn(fabs(e), EVEC ? sE[i] : e_scalar));   // (rho/abc_e)^2
    } else {
      acc.template add_sq<R>(e);
    }
  };
  int i = 0;
  for (; i + 4 <= T; i += 4) {
    step(i, std::integral_constant<int, 0>()); step(i + 1, std::integral_constant<int, 1>());
    step(i + 2, std::integral_constant<int, 2>()); step(i + 3, std::integral_constant<int, 3>());
  }
  if (i < T) step(i, std::integral_constant<int, 0>());
  if (i + 1 < T) step(i + 1, std::integral_constant<int, 1>());
  if (i + 2 < T) step(i + 2, std::integral_constant<int, 2>());
  const double sse = acc.total();
  if (LIK == 31) { const double nse = 1.0 - sse / sst; ll[j] = glue_shape * log(nse > 0.0 ? nse : 0.0); }
  else if (LIK == 21) ll[j] = konst - 0.5 * sse;              // konst = -m/2 log(2 pi) - sum(log(abc_e)), host long double
  else if (LIK == 22) ll[j] = vmin;
  else ll[j] = -(double)T / 2 * log(sse);                     // fun_lik
}

// full series for parity tests of the model itself: out [n][T]
__global__ void hb_hydro_series_kernel(const double* __restrict__ param, long long n, const double* __restrict__ forcing,
                                       int T, double* __restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double K = param[j];
  const double den = 1.0 + K * 1.0;
  const double rden = __ddiv_rn(1.0, den);
  double S = 1.0;
  for (int i = 0; i < T; ++i) {
    S = div_by_recip(__dadd_rn(forcing[i], S), den, rden);
    out[j * (long long)T + i] = __dmul_rn(K, S);
  }
}

// batched independent runs: every run has its own forcing / obs series (BASELINE C5: distinct catchments).
// The series are stored tiled and interleaved: PO[run / 32][t][run % 32] = {P_t, obs_t}.  A warp that owns 32
// consecutive runs (sequential sweep: thread i is run i) reads one contiguous 512-byte line per time step, and a
// block of time steps of its tile is one contiguous piece of memory -- a single 1-D bulk TMA copy per pipeline stage.
// (Run-major storage made every thread stream its own 58 KB row: one 32-byte sector per 8 useful bytes.)
__device__ __forceinline__ long long hydro_po_index(long long run, int t, int T) {
  return ((run >> 5) * (long long)T + t) * 32 + (run & 31);
}

template <int R>
__device__ __forceinline__ void hydro_step(double2 v, double K, double den, double rden, double& S, hydro_sum& acc) {
  S = div_by_recip(__dadd_rn(v.x, S), den, rden);                      // R/test_HydroModel.R:30
  acc.add_sq<R>(__fma_rn(K, S, -v.y));                                 // (Y_i - obs_i)^2, R/test_HydroModel.R:32
}
// n consecutive time steps t0 .. t0 + n - 1 (t0 a multiple of 4) of one run; `at(t)` returns {P_t, obs_t}
template <typename F>
__device__ __forceinline__ void hydro_steps(F at, int n, double K, double den, double rden, double& S, hydro_sum& acc) {
  int t = 0;
#pragma unroll 2
  for (; t + 4 <= n; t += 4) {
    hydro_step<0>(at(t), K, den, rden, S, acc); hydro_step<1>(at(t + 1), K, den, rden, S, acc);
    hydro_step<2>(at(t + 2), K, den, rden, S, acc); hydro_step<3>(at(t + 3), K, den, rden, S, acc);
  }
  if (t < n) hydro_step<0>(at(t), K, den, rden, S, acc);
  if (t + 1 < n) hydro_step<1>(at(t + 1), K, den, rden, S, acc);
  if (t + 2 < n) hydro_step<2>(at(t + 2), K, den, rden, S, acc);
}

// generic batched path (any item -> run mapping): direct loads from the tiled layout
__global__ void __launch_bounds__(128) hb_hydro_items_kernel(const double* __restrict__ X, long long n, int ldx,
                                                             const double2* __restrict__ PO,
                                                             const double* __restrict__ sst, int T,
                                                             double glue_shape, long long chain0, long long stride,
                                                             long long nc, double* __restrict__ ll, unsigned* err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long run = (chain0 + i * stride) / nc;
  const double K = X[i * ldx];
  if (K < 0) { atomicOr(err, HB_FLAG_TARGET_DOMAIN); ll[i] = NAN; return; }
  const double den = 1.0 + K * 1.0;
  const double rden = __ddiv_rn(1.0, den);
  double S = 1.0;
  hydro_sum acc;
  hydro_steps([&](int t) { return PO[hydro_po_index(run, t, T)]; }, T, K, den, rden, S, acc);
  const double nse = 1.0 - acc.total() / sst[run];
  ll[i] = glue_shape * log(nse > 0.0 ? nse : 0.0);
}

// sequential sweep (item i is chain j of run i): one warp per tile of 32 runs; the tile's series stream through a
// two-stage shared-memory ring filled by 1-D bulk TMA copies of HYDRO_TT time steps (32 KB) each, so the recurrence
// never waits on HBM latency -- what is left is the dependent fp64 chain of the backward-Euler step.
constexpr int HYDRO_TT = 64;
__global__ void __launch_bounds__(32) hb_hydro_runs_tma_kernel(const double* __restrict__ X, long long n, int ldx,
                                                               const double2* __restrict__ PO,
                                                               const double* __restrict__ sst, int T, double glue_shape,
                                                               double* __restrict__ ll, unsigned* err) {
  extern __shared__ __align__(16) unsigned char hy_smem[];
  double2* ring = reinterpret_cast<double2*>(hy_smem);                                   // [2][HYDRO_TT][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(hy_smem + 2 * HYDRO_TT * 32 * sizeof(double2));
  const int lane = threadIdx.x;
  const long long tile = blockIdx.x;
  const long long i = tile * 32 + lane;                                                  // item == run
  const double2* __restrict__ src = PO + tile * (long long)T * 32;
  const int n_stages = (T + HYDRO_TT - 1) / HYDRO_TT;
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int s) {
    const int t0 = s * HYDRO_TT;
    const int nt = (T - t0) < HYDRO_TT ? (T - t0) : HYDRO_TT;
    const uint32_t bytes = (uint32_t)nt * 32u * (uint32_t)sizeof(double2);
    mbar_expect_tx(&bars[s & 1], bytes);
    tma_load_1d(ring + (size_t)(s & 1) * HYDRO_TT * 32, src + (long long)t0 * 32, bytes, &bars[s & 1]);
  };
  if (lane == 0) { issue(0); if (n_stages > 1) issue(1); }
  const bool valid = i < n;
  const double K = valid ? X[i * ldx] : 1.0;
  const bool bad = valid && K < 0;
  if (bad) atomicOr(err, HB_FLAG_TARGET_DOMAIN);
  const double den = 1.0 + K * 1.0;
  const double rden = __ddiv_rn(1.0, den);
  static_assert(HYDRO_TT % 4 == 0, "a stage must hold whole groups of the four running sums");
  double S = 1.0;
  hydro_sum acc;
  for (int s = 0; s < n_stages; ++s) {
    mbar_wait(&bars[s & 1], (s >> 1) & 1);
    const double2* __restrict__ buf = ring + (size_t)(s & 1) * HYDRO_TT * 32 + lane;
    const int t0 = s * HYDRO_TT;
    const int nt = (T - t0) < HYDRO_TT ? (T - t0) : HYDRO_TT;
    hydro_steps([&](int t) { return buf[t * 32]; }, nt, K, den, rden, S, acc);
    __syncwarp();                                                                        // every lane is done with the stage
    if (lane == 0 && s + 2 < n_stages) issue(s + 2);
  }
  if (valid) {
    const double nse = 1.0 - acc.total() / sst[i];
    ll[i] = bad ? NAN : glue_shape * log(nse > 0.0 ? nse : 0.0);
  }
}

struct hydro_ctx {
  int T, T_pad, lik;
  double sst, glue_shape;
  double* forcing_dev;
  double* obs_dev;
  double* abc_e_dev = nullptr;   // lik 21 / 22 with a vector abc_e of length T
  double e_scalar = 0.0, konst21 = 0.0;
  bool have_abc = false;
  unsigned* err_dev;
  double* sst_dev = nullptr;   // batched: [n_runs]
  long long n_runs = 0;
};

int hydro_init(void** ctx, int d, int lik, const double* data, const int64_t* dims, int ndims, const double* obs,
               int64_t n_obs, double glue_shape, char* err, int errlen) {
  if (d != 1) { set_err(err, errlen, "HydroModel(forcing, param) has one parameter (R/test_HydroModel.R:17)"); return HB_ERR_INVALID; }
  if (lik != 31 && lik != 21 && lik != 22 && lik != 99) { set_err(err, errlen, "HydroModel returns the simulated series: use a likelihood flag that folds it (lik = 21, 22, 31 or 99)"); return HB_ERR_UNSUPPORTED; }
  if (!data || ndims < 1 || dims[0] < 1) { set_err(err, errlen, "HydroModel needs the forcing series"); return HB_ERR_INVALID; }
  const int64_t T = dims[0];
  if (!obs || n_obs != T) { set_err(err, errlen, "lik = 21/22/31/99 on HydroModel need obs of the same length as forcing"); return HB_ERR_INVALID; }
  for (int64_t i = 0; i < T; ++i)
    if (data[i] < 0) { set_err(err, errlen, "Values of argument 'forcing' need to be >= zero!"); return HB_ERR_INVALID; }  // :20-21
  if ((size_t)(T + 1) * 24 > 220 * 1024) { set_err(err, errlen, "forcing series too long for shared-memory staging"); return HB_ERR_UNSUPPORTED; }
  auto* c = new hydro_ctx();
  c->T = (int)T; c->T_pad = (int)((T + 1) & ~1LL); c->lik = lik; c->glue_shape = glue_shape;
  // sum((obs - mean(obs))^2) with R's long-double mean refinement (summary.c, recollection)
  long double s = 0;
  for (int64_t i = 0; i < T; ++i) s += obs[i];
  s /= T;
  long double r = 0;
  for (int64_t i = 0; i < T; ++i) r += (obs[i] - s);
  s += r / T;
  const double mo = (double)s;
  long double q = 0;
  for (int64_t i = 0; i < T; ++i) { const double e = obs[i] - mo; q += (long double)(e * e); }
  c->sst = (double)q;
  std::vector<double> buf((size_t)c->T_pad, 0.0);
  bool ok = cudaMalloc(&c->forcing_dev, sizeof(double) * c->T_pad) == cudaSuccess &&
            cudaMalloc(&c->obs_dev, sizeof(double) * c->T_pad) == cudaSuccess &&
            cudaMalloc(&c->err_dev, sizeof(unsigned)) == cudaSuccess;
  if (ok) {
    memcpy(buf.data(), data, sizeof(double) * T);
    ok = cudaMemcpy(c->forcing_dev, buf.data(), sizeof(double) * c->T_pad, cudaMemcpyHostToDevice) == cudaSuccess;
    memcpy(buf.data(), obs, sizeof(double) * T);
    ok = ok && cudaMemcpy(c->obs_dev, buf.data(), sizeof(double) * c->T_pad, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemset(c->err_dev, 0, sizeof(unsigned)) == cudaSuccess;
  }
  if (!ok) { set_err(err, errlen, "cudaMalloc/cudaMemcpy failed for forcing/obs"); delete c; return HB_ERR_CUDA; }
  *ctx = c;
  return HB_OK;
}

template <int LIK, bool EVEC>
int hydro_launch_t(hydro_ctx* c, const double* x, int64_t n, int ldx, double* ll, cudaStream_t st) {
  const size_t smem = (size_t)(EVEC ? 3 : 2) * c->T_pad * sizeof(double);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    if (cudaFuncSetAttribute(hb_hydro_kernel<LIK, EVEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) return HB_ERR_CUDA;
    attr_set = true;
  }
  const int threads = n >= 65536 ? 256 : 128;
  hb_hydro_kernel<LIK, EVEC><<<(unsigned)((n + threads - 1) / threads), threads, smem, st>>>(
      x, n, ldx, c->forcing_dev, c->obs_dev, c->abc_e_dev, c->T, c->T_pad, c->sst, c->glue_shape, c->e_scalar, c->konst21, ll, c->err_dev);
  return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}
int hydro_launch(void* ctx, const double* x, int64_t n, int d, int ldx, double* ll, double* fx, void* stream) {
  auto* c = (hydro_ctx*)ctx;
  (void)fx;  // the raw series (3653 values per chain) is folded on the device, SURVEY 7.2 item 7
  if (n <= 0) return HB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if ((c->lik == 21 || c->lik == 22) && !c->have_abc) return HB_ERR_STATE;   // hb_set_lik_args was not called
  const bool ev = c->abc_e_dev != nullptr;
  switch (c->lik) {
    case 31: return hydro_launch_t<31, false>(c, x, n, ldx, ll, st);
    case 99: return hydro_launch_t<99, false>(c, x, n, ldx, ll, st);
    case 21: return ev ? hydro_launch_t<21, true>(c, x, n, ldx, ll, st) : hydro_launch_t<21, false>(c, x, n, ldx, ll, st);
    case 22: return ev ? hydro_launch_t<22, true>(c, x, n, ldx, ll, st) : hydro_launch_t<22, false>(c, x, n, ldx, ll, st);
  }
  return HB_ERR_UNSUPPORTED;
}
int hydro_launch_items(void* ctx, const double* x, int64_t n, int d, int ldx, double* ll, double* fx, int64_t chain0,
                       int64_t stride, int64_t nc, void* stream) {
  auto* c = (hydro_ctx*)ctx;
  if (!c->sst_dev) return hydro_launch(ctx, x, n, d, ldx, ll, fx, stream);   // one series shared by all runs
  if (n <= 0) return HB_OK;
  const double2* po = reinterpret_cast<const double2*>(c->forcing_dev);
  if (stride == nc && chain0 < nc && n == c->n_runs) {   // sequential sweep: item i is chain chain0 of run i
    const size_t smem = 2 * HYDRO_TT * 32 * sizeof(double2) + 16;
    static bool attr_set = false;
    if (!attr_set) {
      if (cudaFuncSetAttribute(hb_hydro_runs_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return HB_ERR_CUDA;
      attr_set = true;
    }
    hb_hydro_runs_tma_kernel<<<(unsigned)((n + 31) / 32), 32, smem, (cudaStream_t)stream>>>(x, n, ldx, po, c->sst_dev, c->T, c->glue_shape, ll, c->err_dev);
  } else {
    hb_hydro_items_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        x, n, ldx, po, c->sst_dev, c->T, c->glue_shape, chain0, stride, nc, ll, c->err_dev);
  }
  return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}
void hydro_destroy(void* ctx) {
  auto* c = (hydro_ctx*)ctx;
  if (c) { cudaFree(c->forcing_dev); cudaFree(c->obs_dev); cudaFree(c->abc_e_dev); cudaFree(c->err_dev); cudaFree(c->sst_dev); delete c; }
}
void hydro_work(void* ctx, int, double* flops, double* bytes) {
  auto* c = (hydro_ctx*)ctx;
  *flops = 5.0 * (c ? c->T : 0); *bytes = 16;
}

// =========================================== registry =====================================================
std::mutex g_mu;
std::vector<hb_target_vtable>& registry() {
  static std::vector<hb_target_vtable> r = {
      {"mixture", mixture_init, mixture_launch, mixture_destroy, mixture_work, nullptr},
      {"t_distribution", tdist_init, tdist_launch, tdist_destroy, tdist_work, nullptr},
      {"HydroModel", hydro_init, hydro_launch, hydro_destroy, hydro_work, hydro_launch_items},
      // the wrapper of the reference's own example script, fun_wrap_glue(x, forc) = HydroModel(forcing = forc, param = x["K"])
      // (examples/script_examples.R:495-501): registered under its name so that the example's dream(fun = "fun_wrap_glue",
      // lik = 99 | 31, forc = ...) call resolves like abc_rho = "abc_distfun" and lik_fun = "fun_lik" do
      {"fun_wrap_glue", hydro_init, hydro_launch, hydro_destroy, hydro_work, hydro_launch_items},
  };
  return r;
}
std::vector<std::string>& owned_names() { static std::vector<std::string> v; return v; }

}  // namespace

const hb_target_vtable* hb_find_target(const char* name) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& v : registry())
    if (name && strcmp(v.name, name) == 0) return &v;
  return nullptr;
}

// per-run forcing/obs for batched independent runs: data [n_runs][T], obs [n_runs][T] (row-major)
int hb_hydro_init_batched(void** ctx, int d, int lik, const double* data, int64_t T, const double* obs, int64_t n_obs,
                          long long n_runs, double glue_shape, char* err, int errlen) {
  if (d != 1) { set_err(err, errlen, "HydroModel(forcing, param) has one parameter (R/test_HydroModel.R:17)"); return HB_ERR_INVALID; }
  if (lik != 31) { set_err(err, errlen, "HydroModel on the device folds the GLUE likelihood: use lik = 31"); return HB_ERR_UNSUPPORTED; }
  if (!data || !obs || T < 1 || n_obs != T || n_runs < 1) { set_err(err, errlen, "batched HydroModel needs forcing and obs of equal length per run"); return HB_ERR_INVALID; }
  for (int64_t i = 0; i < T * n_runs; ++i)
    if (data[i] < 0) { set_err(err, errlen, "Values of argument 'forcing' need to be >= zero!"); return HB_ERR_INVALID; }
  auto* c = new hydro_ctx();
  c->T = (int)T; c->T_pad = (int)((T + 1) & ~1LL); c->lik = lik; c->glue_shape = glue_shape; c->n_runs = n_runs; c->sst = 0;
  const long long n_tiles = (n_runs + 31) / 32;
  std::vector<double> F((size_t)n_tiles * 32 * T * 2, 0.0), sst((size_t)n_runs);   // PO[run / 32][t][run % 32] = {P, obs}
  for (long long r = 0; r < n_runs; ++r) {
    const double* o = obs + r * T;
    const double* f = data + r * T;
    for (int64_t i = 0; i < T; ++i) {
      const size_t e = (((size_t)(r >> 5) * T + i) * 32 + (r & 31)) * 2;
      F[e] = f[i]; F[e + 1] = o[i];
    }
    long double s = 0;
    for (int64_t i = 0; i < T; ++i) s += o[i];
    s /= T;
    long double rr = 0;
    for (int64_t i = 0; i < T; ++i) rr += (o[i] - s);
    s += rr / T;
    const double mo = (double)s;
    long double q = 0;
    for (int64_t i = 0; i < T; ++i) { const double e = o[i] - mo; q += (long double)(e * e); }
    sst[(size_t)r] = (double)q;
  }
  c->obs_dev = nullptr;
  bool ok = cudaMalloc(&c->forcing_dev, F.size() * 8) == cudaSuccess &&
            cudaMalloc(&c->sst_dev, sst.size() * 8) == cudaSuccess && cudaMalloc(&c->err_dev, sizeof(unsigned)) == cudaSuccess;
  ok = ok && cudaMemcpy(c->forcing_dev, F.data(), F.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(c->sst_dev, sst.data(), sst.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemset(c->err_dev, 0, sizeof(unsigned)) == cudaSuccess;
  if (!ok) { set_err(err, errlen, "cudaMalloc/cudaMemcpy failed for batched forcing/obs"); delete c; return HB_ERR_CUDA; }
  *ctx = c;
  return HB_OK;
}

// lik 21 / 22: abc_e (n_e == 1 recycled, or n_e == T) for a HydroModel context created with that flag
int hb_hydro_set_abc(void* ctx, const double* abc_e, int64_t n_e, char* err, int errlen) {
  auto* c = (hydro_ctx*)ctx;
  if (!c || !abc_e || (n_e != 1 && n_e != c->T)) { set_err(err, errlen, "abc_e must have length 1 (recycled) or length(obs)"); return HB_ERR_INVALID; }
  if (c->sst_dev) { set_err(err, errlen, "batched HydroModel runs support lik = 31 only"); return HB_ERR_UNSUPPORTED; }
  long double sl = 0;
  for (int64_t i = 0; i < n_e; ++i) sl += logl((long double)abc_e[i]);      // sum(log(abc_e)) over abc_e AS GIVEN (R/internals.R:10)
  c->konst21 = -(double)c->T / 2 * log(2.0 * 3.14159265358979323846) - (double)sl;
  c->e_scalar = abc_e[0];
  if (n_e > 1) {
    std::vector<double> buf((size_t)c->T_pad, 1.0);
    memcpy(buf.data(), abc_e, sizeof(double) * (size_t)n_e);
    if (!c->abc_e_dev && cudaMalloc(&c->abc_e_dev, sizeof(double) * c->T_pad) != cudaSuccess) { set_err(err, errlen, "cudaMalloc failed for abc_e"); return HB_ERR_CUDA; }
    if (cudaMemcpy(c->abc_e_dev, buf.data(), sizeof(double) * c->T_pad, cudaMemcpyHostToDevice) != cudaSuccess) { set_err(err, errlen, "cudaMemcpy failed for abc_e"); return HB_ERR_CUDA; }
  }
  c->have_abc = true;
  return HB_OK;
}

// HydroModel's domain check lives in its own context (the sampler polls it through this hook)
unsigned hb_hydro_poll_err(void* ctx) {
  auto* c = (hydro_ctx*)ctx;
  unsigned v = 0;
  if (c && cudaMemcpy(&v, c->err_dev, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return v;
}

extern "C" int hb_register_target(const hb_target_vtable* vt) {
  if (!vt || !vt->name || !vt->init || !vt->launch || !vt->destroy) return HB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& v : registry())
    if (strcmp(v.name, vt->name) == 0) return HB_ERR_INVALID;
  owned_names().reserve(64);
  owned_names().emplace_back(vt->name);
  hb_target_vtable c = *vt;
  c.name = owned_names().back().c_str();
  registry().push_back(c);
  return HB_OK;
}

extern "C" int hb_target_registered(const char* name) { return hb_find_target(name) != nullptr; }

__global__ void hb_divcheck_kernel(const double* a, const double* b, long long n, unsigned long long* bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double y = __ddiv_rn(1.0, b[i]);
  if (div_by_recip(a[i], b[i], y) != __ddiv_rn(a[i], b[i])) atomicAdd(bad, 1ull);
}

// test hook: counts pairs where the reciprocal-based quotient differs from the IEEE division (must be 0)
extern "C" int hb_divcheck(const double* a, const double* b, int64_t n, int64_t* n_bad) {
  double *da = nullptr, *db = nullptr; unsigned long long* dbad = nullptr; unsigned long long hb = 0;
  bool ok = cudaMalloc(&da, n * 8) == cudaSuccess && cudaMalloc(&db, n * 8) == cudaSuccess && cudaMalloc(&dbad, 8) == cudaSuccess;
  ok = ok && cudaMemcpy(da, a, n * 8, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(db, b, n * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemset(dbad, 0, 8) == cudaSuccess;
  if (ok) { hb_divcheck_kernel<<<(unsigned)((n + 255) / 256), 256>>>(da, db, n, dbad); ok = cudaMemcpy(&hb, dbad, 8, cudaMemcpyDeviceToHost) == cudaSuccess; }
  cudaFree(da); cudaFree(db); cudaFree(dbad);
  if (!ok) return HB_ERR_CUDA;
  *n_bad = (int64_t)hb;
  return HB_OK;
}

extern "C" int hb_hydromodel_series(const double* forcing, int64_t T, const double* param, int64_t n, double* out,
                                    char* err, int errlen) {
  if (!forcing || !param || !out || T < 1 || n < 1) { set_err(err, errlen, "bad arguments"); return HB_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err(err, errlen, "no CUDA device (there is no CPU fallback)"); return HB_ERR_NO_DEVICE; }
  for (int64_t i = 0; i < n; ++i) if (param[i] < 0) { set_err(err, errlen, "Argument 'param' has to be >= zero!"); return HB_ERR_INVALID; }
  for (int64_t i = 0; i < T; ++i) if (forcing[i] < 0) { set_err(err, errlen, "Values of argument 'forcing' need to be >= zero!"); return HB_ERR_INVALID; }
  double *f = nullptr, *p = nullptr, *o = nullptr;
  bool ok = cudaMalloc(&f, sizeof(double) * T) == cudaSuccess && cudaMalloc(&p, sizeof(double) * n) == cudaSuccess &&
            cudaMalloc(&o, sizeof(double) * n * T) == cudaSuccess;
  ok = ok && cudaMemcpy(f, forcing, sizeof(double) * T, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(p, param, sizeof(double) * n, cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok) {
    hb_hydro_series_kernel<<<(unsigned)((n + 127) / 128), 128>>>(p, n, f, (int)T, o);
    ok = cudaGetLastError() == cudaSuccess &&
         cudaMemcpy(out, o, sizeof(double) * n * T, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  cudaFree(f); cudaFree(p); cudaFree(o);
  if (!ok) { set_err(err, errlen, cudaGetErrorString(cudaGetLastError())); return HB_ERR_CUDA; }
  return HB_OK;
}
