// hb_mt.cu -- MT-DREAM (mt > 1): candidate selection and the multi-try acceptance rule.
//
// Replaces R/dream_parallel.R:295-341 with metropolis_acceptance's mt branch (R/internals.R:227-236).  Per generation
// the host enqueues (hb_api.cu, run_generation_mt):
//   K1 x mt (proposal sets around xt, try-major)  ->  K2 over the mt*nc proposals  ->  hb_mt_select_kernel
//   K1 x (mt-1) (proposal sets around the selected candidates zt)  ->  K2 over the (mt-1)*nc reference proposals
//   ->  hb_mt_accept_kernel  ->  K3 with the candidates as "proposals" and the decisions handed in.
#include "hb_kernels.h"

namespace {

__device__ __forceinline__ double floored_ll(double llr, unsigned& errbits) {
  if (!(llr >= HB_FLOOR)) llr = (llr != llr) ? llr : HB_FLOOR;    // calc_ll floor, R/internals.R:19
  if (!isfinite(llr)) errbits |= HB_FLAG_LL_NONFINITE;            // :21
  return llr;
}

// one warp per chain: selection among the chain's mt proposals with probability exp(lpost_m) / sum (uniform when the sum
// is 0), R/dream_parallel.R:297-307, then the copy of the selected proposal into the candidate population zt (:308-314)
__global__ void __launch_bounds__(256) hb_mt_select_kernel(const hb_mt_params p) {
  const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= p.n_local) return;
  unsigned errbits = 0;
  int pick = 0;
  if (lane == 0) {
    double psum = 0.0;
    for (int m = 0; m < p.mt; ++m) {
      const double ll = floored_ll(p.ll_raw[(long long)m * p.n_local + j], errbits);
      psum += exp(p.lp_xp[(long long)m * p.n_local + j] + ll);
    }
    p.p1[j] = psum;                                               // sum(exp(p_xp)) of R/internals.R:229
    if (p.tape_pick) pick = p.tape_pick[j];
    else {
      const hb_phx ph(p.seed, (uint64_t)(p.gchain0 + j), p.gen);
      const double u = hb_u53(hb_pair64(ph.slot(7), 0));          // slot 7: the multi-try selection uniform
      double cum = 0.0;
      pick = p.mt - 1;
      for (int m = 0; m < p.mt; ++m) {
        const double ll = floored_ll(p.ll_raw[(long long)m * p.n_local + j], errbits);
        cum += (psum == 0.0) ? 1.0 / p.mt : exp(p.lp_xp[(long long)m * p.n_local + j] + ll) / psum;   // :301-305
        if (u < cum) { pick = m; break; }
      }
    }
    if (pick < 0 || pick >= p.mt) { errbits |= HB_FLAG_BAD_TAPE; pick = 0; }
    const long long sel = (long long)pick * p.n_local + j;
    p.lp_zt[j] = p.lp_xp[sel];
    p.ll_zt[j] = floored_ll(p.ll_raw[sel], errbits);
    if (p.fx_zt && p.fx_xp) p.fx_zt[j] = p.fx_xp[sel];
    p.id_zt[j] = p.id[sel];
  }
  pick = __shfl_sync(0xffffffffu, pick, 0);
  const double* __restrict__ src = p.XP + ((long long)pick * p.n_local + j) * p.ldx;
  double* __restrict__ dst = p.ZT + j * (long long)p.ldx;
  for (int k = lane; k < p.ldx; k += 32) dst[k] = src[k];
  if (errbits) atomicOr(p.err, errbits);
}

// thread per chain: p2 = sum over the mt-1 reference proposals of exp(lpost_zp) + exp(lpost of the current state)
// (:335), accept <- min(1, p1/p2) > runif(1), reject when p2 == 0  (R/internals.R:229-236)
__global__ void __launch_bounds__(256) hb_mt_accept_kernel(const hb_mt_params p) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.n_local) return;
  unsigned errbits = 0;
  double p2 = 0.0;
  for (int m = 0; m < p.mt - 1; ++m) {
    const double ll = floored_ll(p.ll_raw[(long long)m * p.n_local + j], errbits);
    p2 += exp(p.lp_xp[(long long)m * p.n_local + j] + ll);
  }
  p2 += exp(p.lpost[j]);
  int a = 0;
  if (p2 != 0.0) {
    const double r = p.p1[j] / p2;
    const double p_acc = r < 1.0 ? r : 1.0;
    double u;
    if (p.u_accept) u = p.u_accept[j];
    else { const hb_phx ph(p.seed, (uint64_t)(p.gchain0 + j), p.gen); u = hb_u53(hb_pair64(ph.slot(HB_SLOT_ACCEPT), 0)); }
    a = p_acc > u;
  }
  p.accept_out[j] = (uint8_t)a;
  if (errbits) atomicOr(p.err, errbits);
}

}  // namespace

cudaError_t hb_launch_mt_select(const hb_mt_params& p, cudaStream_t st) {
  if (p.n_local <= 0) return cudaSuccess;
  hb_mt_select_kernel<<<(unsigned)((p.n_local * 32 + 255) / 256), 256, 0, st>>>(p);
  return cudaGetLastError();
}
cudaError_t hb_launch_mt_accept(const hb_mt_params& p, cudaStream_t st) {
  if (p.n_local <= 0) return cudaSuccess;
  hb_mt_accept_kernel<<<(unsigned)((p.n_local + 255) / 256), 256, 0, st>>>(p);
  return cudaGetLastError();
}
