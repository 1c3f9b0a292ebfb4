// hb_delta.cu -- HB_EXCHANGE_DELTA: the chain-state exchange of a population sharded over the GPUs of one node, over
// peer-mapped memory (CUDA IPC / NVLink), no NCCL on this path.
//
// Why there is an exchange at all: calc_prop draws the 2D partner rows of a proposal from the WHOLE population
// (R/internals.R:142), so every rank needs every row that changed (R/dream_parallel.R:357-358) before the next
// generation's proposals.  Only accepted rows change (acceptance rate 0.16-0.45), so only they travel.
//
// Stage-and-apply protocol (one generation, shard split into P pieces):
//   K3 of piece p appends the rows it accepted to this rank's DELTA LOG (row data + chain index); X is not touched, so
//   the proposal kernels of the later pieces -- here and on every peer -- keep reading the generation-start state
//   (the synchronous sweep of R/dream_parallel.R:269) and no barrier is needed between proposals and the exchange;
//   hb_delta_send_kernel (a few CTAs on a high-priority side stream, concurrent with K1/K2/K3 of piece p + 1) copies the
//   log of piece p into box (parity, self) of every peer's arena with plain posted stores over NVLink, then raises the
//   peers' piece flag;  hb_delta_apply_kernel waits for the piece flags of every peer and copies all logged rows (its
//   own and the peers') into the local replica of X.  Boxes are double buffered by generation parity: a peer can only
//   overwrite box (parity, r) two generations later, i.e. after it has seen this rank's flags of the generation in
//   between, which this rank raises after its own apply.
//
// Every spin has a wall-clock bound (globaltimer, HB_PEER_TIMEOUT_S, default 120 s) and gives up at once when the handle's
// sticky HB_FLAG_PEER_TIMEOUT bit is already set, so a dead peer costs one timeout, not one per barrier.
#include <cstdlib>
#include "hb_kernels.h"

namespace {

__device__ __forceinline__ unsigned long long hb_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// spin until *f >= value (wrap-safe); false on timeout / sticky abort
__device__ __forceinline__ bool hb_spin_ge(const volatile unsigned* f, unsigned value, unsigned* err, unsigned long long timeout_ns) {
  if ((int)(*f - value) >= 0) return true;
  if (*reinterpret_cast<volatile unsigned*>(err) & HB_FLAG_PEER_TIMEOUT) return false;
  const unsigned long long t0 = hb_now_ns();
  unsigned it = 0;
  while ((int)(*f - value) < 0) {
    __nanosleep(32);
    if ((++it & 1023u) == 0) {
      if (hb_now_ns() - t0 > timeout_ns || (*reinterpret_cast<volatile unsigned*>(err) & HB_FLAG_PEER_TIMEOUT)) {
        atomicOr(err, HB_FLAG_PEER_TIMEOUT);
        return false;
      }
    }
  }
  return true;
}

// flag words live at the start of every rank's arena: flags[which * 8 + source rank], values are monotonic
__device__ __forceinline__ void hb_flag_signal_all(hb_arena_peers peers, int which, int self, int world, unsigned value) {
  const int r = threadIdx.x;
  if (r < world && r != self && peers.p[r]) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(reinterpret_cast<unsigned*>(peers.p[r]) + which * 8 + self) = value;
  }
}
__device__ __forceinline__ bool hb_flag_wait_all(const unsigned char* own, int which, int self, int world, unsigned value,
                                                 unsigned* err, unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  bool ok = true;
  if (r < world && r != self) {
    ok = hb_spin_ge(reinterpret_cast<const unsigned*>(own) + which * 8 + r, value, err, timeout_ns);
    __threadfence_system();
  }
  return ok;
}

__global__ void hb_flag_signal_kernel(hb_arena_peers peers, int which, int self, int world, unsigned value) {
  hb_flag_signal_all(peers, which, self, world, value);
}
__global__ void hb_flag_wait_kernel(const unsigned char* own, int which, int self, int world, unsigned value, unsigned* err,
                                    unsigned long long timeout_ns) {
  hb_flag_wait_all(own, which, self, world, value, err, timeout_ns);
}

// ---- small collectives ------------------------------------------------------------------------------------------
__global__ void hb_peer_allreduce_post_kernel(hb_arena_peers peers, size_t slot_off, hb_peer_reduce_desc d, int self,
                                              int world, unsigned seq) {
  double* slot = reinterpret_cast<double*>(peers.p[self] + slot_off);
  int o = 0;
  for (int v = 0; v < 2; ++v) { for (int e = threadIdx.x; e < d.n_f64[v]; e += blockDim.x) slot[o + e] = d.f64[v][e]; o += d.n_f64[v]; }
  unsigned long long* us = reinterpret_cast<unsigned long long*>(slot + o);
  for (int e = threadIdx.x; e < d.n_u64; e += blockDim.x) us[e] = d.u64[e];
  __threadfence_system();
  __syncthreads();
  hb_flag_signal_all(peers, 2, self, world, seq);
}
__global__ void hb_peer_allreduce_finish_kernel(hb_arena_peers peers, size_t slot_off, hb_peer_reduce_desc d, int self,
                                                int world, unsigned seq, unsigned* err, unsigned long long timeout_ns) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (!hb_flag_wait_all(peers.p[self], 2, self, world, seq, err, timeout_ns)) s_ok = 0;
  __syncthreads();
  if (!s_ok) return;
  int o = 0;
  for (int v = 0; v < 2; ++v) {
    for (int e = threadIdx.x; e < d.n_f64[v]; e += blockDim.x) {
      double s = 0.0;
      for (int r = 0; r < world; ++r) s += reinterpret_cast<const volatile double*>(peers.p[r] + slot_off)[o + e];   // rank order
      d.f64[v][e] = s;
    }
    o += d.n_f64[v];
  }
  for (int e = threadIdx.x; e < d.n_u64; e += blockDim.x) {
    unsigned long long s = 0;
    for (int r = 0; r < world; ++r) s += reinterpret_cast<const volatile unsigned long long*>(peers.p[r] + slot_off + (size_t)o * 8)[e];
    d.u64[e] = s;
  }
}

__global__ void hb_peer_allgather_post_kernel(hb_arena_peers peers, size_t area_off, const double* a, const double* b,
                                              long long n_local, int self, int world, unsigned seq, unsigned* done) {
  double* area = reinterpret_cast<double*>(peers.p[self] + area_off);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  for (long long e = tid; e < n_local; e += nt) { area[e] = a[e]; area[n_local + e] = b[e]; }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;     // the last block signals
  __syncthreads();
  if (last) { if (threadIdx.x == 0) *done = 0; __threadfence_system(); hb_flag_signal_all(peers, 3, self, world, seq); }
}
__global__ void hb_peer_allgather_finish_kernel(hb_arena_peers peers, size_t area_off, double* full_a, double* full_b,
                                                long long n_local, int self, int world, unsigned seq, unsigned* err,
                                                unsigned long long timeout_ns) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (!hb_flag_wait_all(peers.p[self], 3, self, world, seq, err, timeout_ns)) s_ok = 0;   // every block waits for itself
  __syncthreads();
  if (!s_ok) return;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < world; ++r) {
    const volatile double* area = reinterpret_cast<const volatile double*>(peers.p[r] + area_off);
    for (long long e = tid; e < n_local; e += nt) { full_a[r * n_local + e] = area[e]; full_b[r * n_local + e] = area[n_local + e]; }
  }
}

// bitwise OR of the ranks' error words (every rank returns the same status from hb_run): own word -> own arena slot,
// flag 4, then OR over all arenas into the local control block
__global__ void hb_peer_err_post_kernel(hb_arena_peers peers, size_t slot_off, const unsigned* err, unsigned extra, int self,
                                        int world, unsigned seq) {
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned*>(peers.p[self] + slot_off + (size_t)(seq & 1u) * 4) = *err | extra;
    __threadfence_system();
  }
  __syncthreads();
  hb_flag_signal_all(peers, 4, self, world, seq);
}
__global__ void hb_peer_err_finish_kernel(hb_arena_peers peers, size_t slot_off, unsigned* err, int self, int world,
                                          unsigned seq, unsigned long long timeout_ns) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (!hb_flag_wait_all(peers.p[self], 4, self, world, seq, err, timeout_ns)) s_ok = 0;
  __syncthreads();
  if (!s_ok) return;
  if (threadIdx.x == 0) {
    unsigned m = 0;
    for (int r = 0; r < world; ++r) m |= *reinterpret_cast<const volatile unsigned*>(peers.p[r] + slot_off + (size_t)(seq & 1u) * 4);
    if (m) atomicOr(err, m);
  }
}

// NCCL exchange modes: the error mask travels as one word per bit through a sum all-reduce
__global__ void hb_err_expand_kernel(const unsigned* err, unsigned extra, unsigned long long* bits) {
  if (threadIdx.x < 16) bits[threadIdx.x] = ((*err | extra) >> threadIdx.x) & 1u;
}
__global__ void hb_err_collapse_kernel(unsigned* err, const unsigned long long* bits) {
  if (threadIdx.x < 16 && bits[threadIdx.x]) atomicOr(err, 1u << threadIdx.x);
}

// ---- delta log: send ------------------------------------------------------------------------------------------------
// Copies piece `piece` of this rank's log (rows + chain indices) into box (parity, self) of every peer.  The log is a
// dense array, so this is a flat vector copy: every 16-byte vector is loaded once from local HBM and stored to each peer
// (posted NVLink writes).  The last CTA to finish publishes the row count in every box header (its own included) and
// raises the peers' piece flag.
template <typename V, int U>
__device__ __forceinline__ void delta_copy(const hb_delta_send_params& p, size_t src_off, size_t n_vec) {
  const V* __restrict__ src = reinterpret_cast<const V*>(p.box[p.self] + src_off);
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
  size_t i = tid;
  for (; i + (U - 1) * nt < n_vec; i += U * nt) {               // U independent loads in flight per thread
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = src[i + u * nt];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < p.world && r != p.self) {
        V* dst = reinterpret_cast<V*>(p.box[r] + src_off);
#pragma unroll
        for (int u = 0; u < U; ++u) dst[i + u * nt] = v[u];
      }
  }
  for (; i < n_vec; i += nt) {
    const V v = src[i];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < p.world && r != p.self) reinterpret_cast<V*>(p.box[r] + src_off)[i] = v;
  }
}

__global__ void __launch_bounds__(128) hb_delta_send_kernel(const hb_delta_send_params p) {
  const unsigned n = *reinterpret_cast<const volatile unsigned*>(p.log_count);
  const size_t rows_off = p.rows_off + (size_t)p.piece_off * p.ldx * 8;
  const size_t idx_off = p.idx_off + (size_t)p.piece_off * 4;
  if ((p.ldx & 1) == 0) delta_copy<double2, 8>(p, rows_off, (size_t)n * p.ldx / 2);
  else delta_copy<double, 4>(p, rows_off, (size_t)n * p.ldx);
  delta_copy<int, 2>(p, idx_off, (size_t)n);
  __threadfence_system();                                       // this thread's remote stores are performed ...
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(p.done_ctr, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  const int r = threadIdx.x;
  if (r < p.world) {                                            // ... before the count and the flag that publish them
    reinterpret_cast<volatile unsigned*>(p.box[r])[p.piece] = n;
    __threadfence_system();
    if (r != p.self) *reinterpret_cast<volatile unsigned*>(reinterpret_cast<unsigned*>(p.arena[r]) + 1 * 8 + p.self) = p.flag_value;
  }
  if (threadIdx.x == 0) { *p.done_ctr = 0u; *p.log_count = 0u; }   // the piece's counter is free for generation + 2
}

// ---- delta log: apply -----------------------------------------------------------------------------------------------
// x[accepted, ] <- xp[accepted, ] (R/dream_parallel.R:358) for the rows every rank accepted in pieces [p0, p1) of this
// generation: waits for the peers' piece flags, then one warp per logged row copies it into the replica.
__global__ void __launch_bounds__(256) hb_delta_apply_kernel(const hb_delta_apply_params p) {
  __shared__ unsigned s_pref[8 * HB_MAX_PIECES + 1];
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  {
    const int r = threadIdx.x;
    if (r < p.world && r != p.self) {
      if (!hb_spin_ge(p.flags + 1 * 8 + r, p.flag_value, p.err, p.timeout_ns)) s_ok = 0;
      __threadfence_system();
    }
  }
  __syncthreads();
  if (!s_ok) return;
  const int np = p.p1 - p.p0, n_seg = p.world * np;
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    for (int sgi = 0; sgi < n_seg; ++sgi) {
      s_pref[sgi] = acc;
      const int r = sgi / np, pc = p.p0 + sgi % np;
      unsigned c = reinterpret_cast<const volatile unsigned*>(p.box[r])[pc];
      if ((long long)c > p.piece_cap) { c = 0; atomicOr(p.err, HB_FLAG_PEER_TIMEOUT); }   // corrupt header: never write wild
      acc += c;
    }
    s_pref[n_seg] = acc;
  }
  __syncthreads();
  const unsigned total = s_pref[n_seg];
  const int lane = threadIdx.x & 31;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int ldx = p.ldx;
  auto locate = [&](long long i, const double*& src, double*& dst) -> bool {
    int sgi = 0;
    while (sgi + 1 < n_seg && s_pref[sgi + 1] <= (unsigned)i) ++sgi;
    const int r = sgi / np, pc = p.p0 + sgi % np;
    const long long slot = (long long)pc * p.piece_cap + (i - s_pref[sgi]);
    const int chain = reinterpret_cast<const int*>(p.box[r] + p.idx_off)[slot];
    if (chain < 0 || chain >= p.nc) { if (lane == 0) atomicOr(p.err, HB_FLAG_PEER_TIMEOUT); return false; }
    src = reinterpret_cast<const double*>(p.box[r] + p.rows_off) + slot * ldx;
    dst = p.X + (long long)chain * ldx;
    return true;
  };
  if ((ldx & 1) == 0 && ldx <= 128) {          // rows of up to 64 16-byte vectors: two rows (four loads) in flight per warp
    const int nv = ldx / 2;
    for (long long i = warp_id; i < (long long)total; i += 2 * n_warps) {
      const double* sa = nullptr; double* da = nullptr; const double* sb = nullptr; double* db = nullptr;
      const bool oka = locate(i, sa, da);
      const bool okb = (i + n_warps < (long long)total) && locate(i + n_warps, sb, db);
      double2 a0, a1, b0, b1;
      if (oka) { if (lane < nv) a0 = reinterpret_cast<const double2*>(sa)[lane]; if (lane + 32 < nv) a1 = reinterpret_cast<const double2*>(sa)[lane + 32]; }
      if (okb) { if (lane < nv) b0 = reinterpret_cast<const double2*>(sb)[lane]; if (lane + 32 < nv) b1 = reinterpret_cast<const double2*>(sb)[lane + 32]; }
      if (oka) { if (lane < nv) reinterpret_cast<double2*>(da)[lane] = a0; if (lane + 32 < nv) reinterpret_cast<double2*>(da)[lane + 32] = a1; }
      if (okb) { if (lane < nv) reinterpret_cast<double2*>(db)[lane] = b0; if (lane + 32 < nv) reinterpret_cast<double2*>(db)[lane + 32] = b1; }
    }
  } else {
    for (long long i = warp_id; i < (long long)total; i += n_warps) {
      const double* src = nullptr; double* dst = nullptr;
      if (!locate(i, src, dst)) continue;
      if ((ldx & 1) == 0) {
        const double2* s2 = reinterpret_cast<const double2*>(src);
        double2* d2 = reinterpret_cast<double2*>(dst);
        for (int k = lane; k < ldx / 2; k += 32) d2[k] = s2[k];
      } else {
        for (int k = lane; k < ldx; k += 32) dst[k] = src[k];
      }
    }
  }
}

// ---- delta log: direct push of the last piece ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hb_delta_push_kernel(const hb_delta_push_params p) {
  __shared__ int s_ok;
  __shared__ bool last;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  {
    const int r = threadIdx.x;
    if (r < p.world && r != p.self) {
      if (!hb_spin_ge(p.flags + 5 * 8 + r, p.gen, p.err, p.timeout_ns)) s_ok = 0;
      __threadfence_system();
    }
  }
  __syncthreads();
  if (s_ok) {
    const unsigned n = *reinterpret_cast<const volatile unsigned*>(p.log_count);
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int ldx = p.ldx;
    const int* idx = reinterpret_cast<const int*>(p.box + p.idx_off) + p.piece_off;
    const double* rows = reinterpret_cast<const double*>(p.box + p.rows_off) + p.piece_off * ldx;
    for (long long i = warp_id; i < (long long)n; i += n_warps) {
      const int chain = idx[i];
      if (chain < 0 || chain >= p.nc) { if (lane == 0) atomicOr(p.err, HB_FLAG_PEER_TIMEOUT); continue; }
      const double* __restrict__ src = rows + i * ldx;
      const long long o = (long long)chain * ldx;
      if ((ldx & 1) == 0 && ldx <= 128) {
        const int nv = ldx / 2;
        double2 a0 = make_double2(0, 0), a1 = make_double2(0, 0);
        if (lane < nv) a0 = reinterpret_cast<const double2*>(src)[lane];
        if (lane + 32 < nv) a1 = reinterpret_cast<const double2*>(src)[lane + 32];
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < p.world) {
            double2* dst = reinterpret_cast<double2*>(p.X[r] + o);
            if (lane < nv) dst[lane] = a0;
            if (lane + 32 < nv) dst[lane + 32] = a1;
          }
      } else {
        for (int k = lane; k < ldx; k += 32) {
          const double v = src[k];
          for (int r = 0; r < p.world; ++r) p.X[r][o + k] = v;
        }
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(p.done_ctr, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  const int r = threadIdx.x;
  if (r < p.world && r != p.self)
    *reinterpret_cast<volatile unsigned*>(reinterpret_cast<unsigned*>(p.arena[r]) + 6 * 8 + p.self) = p.gen;
  if (threadIdx.x == 0) { *p.done_ctr = 0u; *p.log_count = 0u; }
}

}  // namespace

unsigned long long hb_peer_timeout_ns() {
  static const unsigned long long v = []() {
    double s = 120.0;
    if (const char* e = getenv("HB_PEER_TIMEOUT_S")) { const double x = atof(e); if (x > 0) s = x; }
    return (unsigned long long)(s * 1e9);
  }();
  return v;
}

cudaError_t hb_launch_flag_signal(hb_arena_peers peers, int which, int self, int world, unsigned value, cudaStream_t st) {
  hb_flag_signal_kernel<<<1, 32, 0, st>>>(peers, which, self, world, value);
  return cudaGetLastError();
}
cudaError_t hb_launch_flag_wait(const unsigned char* own_arena, int which, int self, int world, unsigned value, unsigned* err,
                                cudaStream_t st) {
  hb_flag_wait_kernel<<<1, 32, 0, st>>>(own_arena, which, self, world, value, err, hb_peer_timeout_ns());
  return cudaGetLastError();
}

cudaError_t hb_launch_peer_allreduce_post(hb_arena_peers peers, size_t slot_off, hb_peer_reduce_desc desc, int self, int world,
                                          unsigned seq, cudaStream_t st) {
  hb_peer_allreduce_post_kernel<<<1, 256, 0, st>>>(peers, slot_off, desc, self, world, seq);
  return cudaGetLastError();
}
cudaError_t hb_launch_peer_allreduce_finish(hb_arena_peers peers, size_t slot_off, hb_peer_reduce_desc desc, int self, int world,
                                            unsigned seq, unsigned* err, cudaStream_t st) {
  hb_peer_allreduce_finish_kernel<<<1, 256, 0, st>>>(peers, slot_off, desc, self, world, seq, err, hb_peer_timeout_ns());
  return cudaGetLastError();
}

static int allgather_blocks(long long n_local) {
  const long long b = (n_local + 1023) / 1024;
  return (int)(b < 1 ? 1 : b > 64 ? 64 : b);
}
cudaError_t hb_launch_peer_allgather_post(hb_arena_peers peers, size_t area_off, const double* local_a, const double* local_b,
                                          long long n_local, int self, int world, unsigned seq, cudaStream_t st) {
  unsigned* done = reinterpret_cast<unsigned*>(peers.p[self]) + 63;       // last word of the flag block: block counter
  hb_peer_allgather_post_kernel<<<allgather_blocks(n_local), 256, 0, st>>>(peers, area_off, local_a, local_b, n_local, self, world, seq, done);
  return cudaGetLastError();
}
cudaError_t hb_launch_peer_allgather_finish(hb_arena_peers peers, size_t area_off, double* full_a, double* full_b,
                                            long long n_local, int self, int world, unsigned seq, unsigned* err, cudaStream_t st) {
  hb_peer_allgather_finish_kernel<<<allgather_blocks(n_local), 256, 0, st>>>(peers, area_off, full_a, full_b, n_local, self, world, seq, err,
                                                                           hb_peer_timeout_ns());
  return cudaGetLastError();
}

cudaError_t hb_launch_peer_err_post(hb_arena_peers peers, size_t slot_off, const unsigned* err, unsigned extra, int self, int world,
                                    unsigned seq, cudaStream_t st) {
  hb_peer_err_post_kernel<<<1, 32, 0, st>>>(peers, slot_off, err, extra, self, world, seq);
  return cudaGetLastError();
}
cudaError_t hb_launch_peer_err_finish(hb_arena_peers peers, size_t slot_off, unsigned* err, int self, int world, unsigned seq,
                                      cudaStream_t st) {
  hb_peer_err_finish_kernel<<<1, 32, 0, st>>>(peers, slot_off, err, self, world, seq, hb_peer_timeout_ns());
  return cudaGetLastError();
}

cudaError_t hb_launch_err_expand(const unsigned* err, unsigned extra, unsigned long long* bits16, cudaStream_t st) {
  hb_err_expand_kernel<<<1, 32, 0, st>>>(err, extra, bits16);
  return cudaGetLastError();
}
cudaError_t hb_launch_err_collapse(unsigned* err, const unsigned long long* bits16, cudaStream_t st) {
  hb_err_collapse_kernel<<<1, 32, 0, st>>>(err, bits16);
  return cudaGetLastError();
}

cudaError_t hb_launch_delta_send(const hb_delta_send_params& p, int ctas, cudaStream_t st) {
  hb_delta_send_kernel<<<ctas < 1 ? 1 : ctas, 128, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t hb_launch_delta_push(const hb_delta_push_params& p, int sm_count, cudaStream_t st) {
  if (p.world > 8) return cudaErrorInvalidValue;
  hb_delta_push_kernel<<<sm_count * 2, 256, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t hb_launch_delta_apply(const hb_delta_apply_params& p, int sm_count, cudaStream_t st) {
  if (p.p1 - p.p0 < 1 || p.p1 - p.p0 > HB_MAX_PIECES || p.world > 8) return cudaErrorInvalidValue;
  hb_delta_apply_kernel<<<sm_count * 4, 256, 0, st>>>(p);
  return cudaGetLastError();
}
