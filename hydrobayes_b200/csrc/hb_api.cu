// hb_api.cu -- the C ABI (include/hydrobayes_b200.h): handle life cycle, the generation loop of the synchronous
// sweep (R/dream_parallel.R:257-435 as K1 -> K2 plugin -> K3 -> finalize [-> exchange] [-> outlier check]),
// result fetch in R's layouts, and the stand-alone entry points.
#include <execinfo.h>
#include <math.h>
#include <signal.h>
#include <stdarg.h>
#include <unistd.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <mutex>
#include <atomic>
#include <chrono>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "hb_internal.h"

int hb_seq_supported(const hb_config* cfg, std::string& why);   // hb_seq.cu
int hb_seq_generation(hb_handle* h, long long i);               // hb_seq.cu
int hb_seq_allocate(hb_handle* h);
int hb_seq_initialise(hb_handle* h);
void hb_seq_destroy(hb_handle* h);
int hb_seq_pull_outliers(hb_handle* h);
int hb_hydro_init_batched(void** ctx, int d, int lik, const double* data, int64_t T, const double* obs, int64_t n_obs,
                          long long n_runs, double glue_shape, char* err, int errlen);

namespace {

// HB_BACKTRACE=1: print the native stack on SIGSEGV / SIGABRT (debugging aid; the offsets resolve with addr2line -e <.so>)
void hb_crash_handler(int sig) {
  const char msg[] = "[hb] fatal signal, native backtrace:\n";
  if (write(2, msg, sizeof msg - 1) < 0) {}
  void* frames[64];
  const int n = backtrace(frames, 64);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}
struct hb_crash_init {
  hb_crash_init() {
    const char* e = getenv("HB_BACKTRACE");
    if (!(e && e[0] == '1')) return;
    void* warm[4]; backtrace(warm, 4);            // loads libgcc now, not inside the handler
    static char altstack[1 << 16];
    stack_t ss{}; ss.ss_sp = altstack; ss.ss_size = sizeof altstack; sigaltstack(&ss, nullptr);
    struct sigaction sa{}; sa.sa_handler = hb_crash_handler; sa.sa_flags = SA_ONSTACK; sigemptyset(&sa.sa_mask);
    sigaction(SIGSEGV, &sa, nullptr); sigaction(SIGABRT, &sa, nullptr); sigaction(SIGBUS, &sa, nullptr);
  }
} g_crash_init;

thread_local std::string g_create_err;   // last hb_create failure of the calling thread (hb_last_error(NULL))

int fail(hb_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (h) h->err = buf; else g_create_err = buf;
  return code;
}

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return fail(h, HB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// Deferred release of a destroyed handle's device buffers.  cudaFree of the multi-GB history buffers takes 15-350 ms (it
// unmaps and synchronises), and it sits on the caller's critical path -- dream() returns only after hb_destroy.  The
// frees therefore run on a reaper thread; an allocation that fails for lack of memory first waits for the reapers.
void reaper_at_exit();
struct reaper_t {
  std::mutex mu;
  std::vector<std::thread> threads;
  bool at_exit_set = false;
  void add(std::thread&& t) {
    bool prune = false;
    {
      std::lock_guard<std::mutex> lk(mu);
      threads.push_back(std::move(t));
      prune = threads.size() > 8;
      // registered after the CUDA runtime's own exit handler (a CUDA call always precedes the first destroy), so it runs
      // BEFORE the runtime is torn down: no cudaFree is in flight during teardown
      if (!at_exit_set) { at_exit_set = true; atexit(reaper_at_exit); }
    }
    if (prune) join_all();   // bounds the number of pending reapers (test suites create hundreds of handles)
  }
  void join_all() {
    std::vector<std::thread> ts;
    { std::lock_guard<std::mutex> lk(mu); ts.swap(threads); }
    for (auto& t : ts) if (t.joinable()) t.join();
  }
  ~reaper_t() { join_all(); }
} g_reaper;
void reaper_at_exit() { g_reaper.join_all(); }

template <typename T>
cudaError_t dalloc_raw(T** p, size_t n);

template <typename T>
cudaError_t dalloc(T** p, size_t n) {
  cudaError_t e = dalloc_raw(p, n);
  if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); g_reaper.join_all(); e = dalloc_raw(p, n); }
  return e;
}

template <typename T>
cudaError_t dalloc_raw(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  return cudaMalloc((void**)p, n * sizeof(T));
}

void compute_dims(hb_handle* h) {
  const hb_config& c = h->cfg;
  if (c.burnin > 0 || c.thin > 1) h->out_t = (long long)((1 - c.burnin) * (double)c.t / c.thin + 1);   // :195-198
  else h->out_t = c.t;
  if (c.sweep == HB_SWEEP_SYNCHRONOUS) {
    h->ar_rows = c.t / c.update_interval + 1; h->ar_cols = 2;                                         // :205
    h->cr_rows = h->ar_rows; h->cr_cols = c.nCR + 1;                                                  // :204
  } else {
    h->ar_rows = h->out_t; h->ar_cols = c.nCR; h->cr_rows = h->out_t; h->cr_cols = c.nCR;             // R/dream.R:162
  }
  h->H = (long long)floor(c.adapt * (double)c.t) + 1;
  h->rstat_rows = (c.check_convergence && h->out_t > 50) ? h->out_t - 50 : 0;
}

// host-side Philox for the outlier replacement draws (slot map in hb_common.cuh)
void sample_replacements(const hb_handle* h, long long gen, const std::vector<long long>& outl,
                         std::vector<long long>& news) {
  const long long nc = h->nc;
  news.resize(outl.size());
  std::vector<long long> cand; cand.reserve((size_t)nc);
  size_t o = 0;
  for (long long j = 0; j < nc; ++j) { if (o < outl.size() && outl[o] == j) { ++o; continue; } cand.push_back(j); }
  long long n_rem = (long long)cand.size();
  const hb_phx ph(h->cfg.seed, ((uint64_t)0xFFFFFFFFu << 32) | 0u, (uint32_t)gen);
  uint4 w = make_uint4(0, 0, 0, 0);
  for (size_t q = 0; q < outl.size(); ++q) {
    if ((q & 1) == 0) w = ph.slot((uint32_t)(q >> 1));
    const unsigned __int128 prod = (unsigned __int128)hb_pair64(w, (int)(q & 1)) * (unsigned __int128)(uint64_t)n_rem;
    const long long r = (long long)(uint64_t)(prod >> 64);
    news[q] = cand[(size_t)r];
    cand[(size_t)r] = cand[(size_t)--n_rem];
  }
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// HB_TIMING=1: one stderr line per host-side phase (seconds, and GB/s when a byte count is given)
void timing_note(const char* what, double seconds, double bytes = 0) {
  static const bool on = getenv("HB_TIMING") && getenv("HB_TIMING")[0] == '1';
  if (!on) return;
  if (bytes > 0) fprintf(stderr, "[hb timing] %-14s %8.1f ms  %6.2f GB  %6.1f GB/s\n", what, seconds * 1e3, bytes / 1e9, bytes / 1e9 / seconds);
  else fprintf(stderr, "[hb timing] %-14s %8.1f ms\n", what, seconds * 1e3);
}

// Copy into the caller's (pageable, pre-faulted) result array with non-temporal stores: the destination is written once and
// not read again here, so the read-for-ownership of every destination line a plain memcpy of this size would do (a thread's
// share of a chunk is below glibc's non-temporal threshold) is a third of the copy's memory traffic -- and the fetch of a
// 9.5 GB result is bound by the host's memory bandwidth (DMA into the ring + this copy), not by PCIe.
void stream_copy(char* dst, const char* src, size_t n) {
#if defined(__SSE2__)
  while (n >= 8 && (reinterpret_cast<uintptr_t>(dst) & 15)) { memcpy(dst, src, 8); dst += 8; src += 8; n -= 8; }
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32)), d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a); _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c); _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
    }
    _mm_sfence();
    dst += i; src += i; n -= i;
  }
#endif
  if (n) memcpy(dst, src, n);
}

// Pinned host ring of the upload / fetch pipelines.  cudaHostAlloc of 128 MB costs ~0.1 s, more than the transfer of a
// typical initial population, so the ring is cached per process: a destroyed handle parks its ring here and the next
// handle (the next dream() call of the same R / Python session) picks it up.  Freed when the library is unloaded.
struct pin_cache_t {
  std::mutex mu; void* p = nullptr; size_t bytes = 0;
  ~pin_cache_t() { if (p) cudaFreeHost(p); }
} g_pin_cache;
int pin_acquire(hb_handle* h, size_t bytes) {
  if (h->pin_bytes >= bytes) return HB_OK;
  {
    std::lock_guard<std::mutex> lk(g_pin_cache.mu);
    if (g_pin_cache.p && g_pin_cache.bytes >= bytes) {
      if (h->pin) cudaFreeHost(h->pin);
      h->pin = g_pin_cache.p; h->pin_bytes = g_pin_cache.bytes; g_pin_cache.p = nullptr; g_pin_cache.bytes = 0;
      return HB_OK;
    }
  }
  if (h->pin) { cudaFreeHost(h->pin); h->pin = nullptr; h->pin_bytes = 0; }
  CK(cudaHostAlloc((void**)&h->pin, bytes, cudaHostAllocDefault));
  h->pin_bytes = bytes;
  return HB_OK;
}
void pin_release(hb_handle* h) {
  if (!h->pin) return;
  std::lock_guard<std::mutex> lk(g_pin_cache.mu);
  if (!g_pin_cache.p || g_pin_cache.bytes < h->pin_bytes) {
    if (g_pin_cache.p) cudaFreeHost(g_pin_cache.p);
    g_pin_cache.p = h->pin; g_pin_cache.bytes = h->pin_bytes;
  } else cudaFreeHost(h->pin);
  h->pin = nullptr; h->pin_bytes = 0;
}

struct prof_scope {
  hb_handle* h; const char* name; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
  prof_scope(hb_handle* h_, const char* n, cudaStream_t st_ = nullptr) : h(h_), name(n), st(st_ ? st_ : h_->stream) {
    if (h->profile) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
  }
  void done(int launches = 1) {
    h->launches += launches;
    if (h->profile) {
      cudaEventRecord(b, st); cudaEventSynchronize(b);
      float ms = 0; cudaEventElapsedTime(&ms, a, b);
      auto& k = h->kstats[name]; k.ms += ms; k.launches += launches;
      cudaEventDestroy(a); cudaEventDestroy(b);
    }
  }
};

int check_flags(hb_handle* h) {
  hb_ctrl c;
  CK(cudaMemcpyAsync(&c, h->ctrl, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  unsigned e = c.err;
  if (h->target_is_hydro) e |= hb_hydro_poll_err(h->tctx);
  if (!e) return HB_OK;
  if (e & HB_FLAG_BAD_TAPE) return fail(h, HB_ERR_INVALID, "invalid tape record (D, partner or id out of range)");
  if (e & HB_FLAG_PROP_NONFINITE) return fail(h, HB_ERR_NONFINITE, "Calculated proposal is not finite!");
  if (e & HB_FLAG_LP_NONFINITE) return fail(h, HB_ERR_NONFINITE, "Calculated log-prior is not finite");
  if (e & HB_FLAG_LL_NONFINITE) return fail(h, HB_ERR_NONFINITE, "Calculated log-likelihood is not finite!");
  if (e & HB_FLAG_TARGET_DOMAIN) return fail(h, HB_ERR_INVALID, "Argument 'param' has to be >= zero!");
  if (e & HB_FLAG_PEER_TIMEOUT) return fail(h, HB_ERR_NCCL, "HB_EXCHANGE_DELTA: a peer GPU did not reach the generation barrier (timeout)");
  if (e & HB_FLAG_J_NAN) return fail(h, HB_ERR_NONFINITE, "missing value where TRUE/FALSE needed (J is NaN: zero std_x)");
  return fail(h, HB_ERR_INVALID, "device error flags 0x%x", e);
}

int allocate(hb_handle* h) {
  const hb_config& c = h->cfg;
  const long long nl = h->n_local, nc = h->modeb ? h->nct : h->nc;   // batched runs: all runs' chains, run-major
  const int d = h->d, ldx = h->ldx;
  if (h->peer_mode) {   // this rank's shard only, double buffered; X becomes the virtual full-population base
    for (int b = 0; b < 2; ++b) {
      CK(dalloc(&h->Xbuf[b], (size_t)nl * ldx));
      CK(cudaMemsetAsync(h->Xbuf[b], 0, (size_t)nl * ldx * sizeof(double), h->stream));
    }
    h->cur = 0;
    h->X = h->Xbuf[0] - h->chain0 * ldx;
  } else {
    CK(dalloc(&h->X, (size_t)nc * ldx));
    CK(cudaMemsetAsync(h->X, 0, (size_t)nc * ldx * sizeof(double), h->stream));
  }
  if (c.past_sample) { CK(dalloc(&h->Z, (size_t)h->m_cap * ldx)); }
  h->wide = g_hb_force_wide || nl >= (long long)h->sm_count * 16 * 32 / 2;   // the threshold of hb_launch_k1 / hb_launch_k3
  // diagnostics: HB_FORCE_DELTA=1 runs the stage-and-apply pipeline (pieces, delta log, send, apply) on ONE GPU with no
  // peers, so that its kernel and launch overheads can be measured against the plain path without a multi-GPU box
  if (h->world == 1 && !h->modeb && c.mt <= 1 && !h->peer_mode && getenv("HB_FORCE_DELTA") && getenv("HB_FORCE_DELTA")[0] == '1') {
    h->delta_mode = true; h->peers_ready = true; h->rank = 0;
  }
  h->pieces = 1; h->piece_n = nl;
  if (h->delta_mode) {
    // pieces per generation: the send of piece p overlaps the kernels of piece p + 1, so only the last piece's transfer is
    // exposed; a piece must still fill the device (>= 16 384 chains).  HB_PIECES overrides (tests run tiny populations).
    int P = 1;
    if (h->wide) { P = 2; while (P > 1 && nl / P < 32768) P >>= 1; }
    if (const char* e = getenv("HB_PIECES")) { const int v = atoi(e); if (v >= 1 && v <= HB_MAX_PIECES) P = v; }
    if ((long long)P > nl) P = (int)nl;
    h->pieces = P;
    h->piece_n = (nl + P - 1) / P;
    if (P > 1) h->piece_n = (h->piece_n + 31) & ~31LL;
    // the send kernel (128 threads, ~80 registers: 10k registers per CTA) runs next to the persistent proposal grid: the
    // proposal kernel leaves `reserve_ctas` of its 2-per-SM CTA slots free, each of which hosts three send CTAs
    h->send_ctas = 48;
    if (const char* e = getenv("HB_SEND_CTAS")) { const int v = atoi(e); if (v >= 1 && v <= 1024) h->send_ctas = v; }
    h->reserve_ctas = h->wide ? (h->send_ctas + 2) / 3 : 0;
    if (const char* e = getenv("HB_RESERVE_CTAS")) { const int v = atoi(e); if (v >= 0 && v <= 128) h->reserve_ctas = v; }
    // arena mapped by the peers: flags, error slots, all-reduce slots, all-gather area, delta boxes (no NCCL on this path)
    const size_t E = (size_t)(2 + c.nCR) * d;
    h->arena_red_bytes = ((E * (size_t)c.update_interval + HB_MAX_NCR + 1 + 3 * (size_t)d) * 8 + 255) & ~(size_t)255;
    h->arena_gather_off = HB_ARENA_RED_OFF + 2 * h->arena_red_bytes;
    h->arena_box_off = (h->arena_gather_off + 2 * (size_t)nl * sizeof(double) + 255) & ~(size_t)255;
    const long long cap = h->piece_n * P;                                       // log slots per box
    h->box_idx_off = 256;                                                       // counts[HB_MAX_PIECES] live in the first 64 bytes
    h->box_rows_off = (h->box_idx_off + (size_t)cap * 4 + 255) & ~(size_t)255;
    h->box_bytes = (h->box_rows_off + (size_t)cap * ldx * 8 + 255) & ~(size_t)255;
    CK(cudaMalloc((void**)&h->arena, h->arena_box_off));
    CK(cudaMemsetAsync(h->arena, 0, h->arena_box_off, h->stream));
    for (int r = 0; r < h->world; ++r) {
      CK(cudaMalloc((void**)&h->boxmem[r], 2 * h->box_bytes));
      for (int b = 0; b < 2; ++b) CK(cudaMemsetAsync(h->boxmem[r] + (size_t)b * h->box_bytes, 0, 256, h->stream));   // the piece counts
    }
    h->peer_box[h->rank] = h->boxmem[h->rank];
    h->arena_peer[h->rank] = h->arena;
    CK(dalloc(&h->log_ctr, (size_t)3 * HB_MAX_PIECES));
    CK(cudaMemsetAsync(h->log_ctr, 0, sizeof(unsigned) * 3 * HB_MAX_PIECES, h->stream));
    CK(dalloc(&h->ibuf, E * (size_t)c.update_interval));
    if (!h->side) {
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));
      CK(cudaStreamCreateWithPriority(&h->side2, cudaStreamNonBlocking, hi));
    }
    for (int q = 0; q < P; ++q) CK(cudaEventCreateWithFlags(&h->ev_piece[q], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_sent, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_k1last, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_applied, cudaEventDisableTiming));
  } else if (h->world > 1) {
    CK(dalloc(&h->errbits_dev, 16));
  }
  const size_t mtn = (size_t)(c.mt > 1 ? c.mt : 1) * (size_t)nl;   // MT-DREAM: mt proposal sets live side by side, try-major
  CK(dalloc(&h->XP, mtn * ldx));
  CK(cudaMemsetAsync(h->XP, 0, mtn * ldx * sizeof(double), h->stream));
  CK(dalloc(&h->id, mtn));
  CK(dalloc(&h->lp_xp, mtn)); CK(dalloc(&h->ll_raw, mtn)); CK(dalloc(&h->fx_xp, mtn));
  if (c.mt > 1) {
    // candidates zt: the reference proposals draw their partners from the WHOLE candidate population (R/dream_parallel.R:316),
    // so a sharded run keeps a full-size copy (this rank's rows at chain0, the others all-gathered every generation)
    const size_t zrows = h->world > 1 ? (size_t)h->nc : (size_t)nl;
    CK(dalloc(&h->ZT, zrows * ldx));
    CK(cudaMemsetAsync(h->ZT, 0, zrows * ldx * sizeof(double), h->stream));
    CK(dalloc(&h->lp_zt, (size_t)nl)); CK(dalloc(&h->ll_zt, (size_t)nl)); CK(dalloc(&h->fx_zt, (size_t)nl));
    CK(dalloc(&h->id_zt, (size_t)nl)); CK(dalloc(&h->mt_p1, (size_t)nl)); CK(dalloc(&h->mt_accept, (size_t)nl));
  }
  CK(dalloc(&h->lp, (size_t)nl)); CK(dalloc(&h->ll, (size_t)nl)); CK(dalloc(&h->lpost, (size_t)nl));
  CK(dalloc(&h->fx, (size_t)nl));
  CK(dalloc(&h->lpost_hist, (size_t)h->H * nl));
  CK(dalloc(&h->hist_x, (size_t)h->out_t * nl * d));
  CK(dalloc(&h->hist_l, (size_t)h->out_t * nl * 3));
  if (c.store_fx) CK(dalloc(&h->hist_fx, (size_t)h->out_t * nl));
  if (c.record_accept) { CK(dalloc(&h->accept_rec, (size_t)(h->t - 1) * nl)); CK(cudaMemsetAsync(h->accept_rec, 0, (size_t)(h->t - 1) * nl, h->stream)); }
  CK(dalloc(&h->ctrl, (size_t)h->cfg.n_runs));
  CK(dalloc(&h->shift, (size_t)d)); CK(dalloc(&h->colsum, (size_t)2 * d)); CK(dalloc(&h->qsum, (size_t)c.nCR * d));
  size_t smem = 0;
  h->k3_blocks_piece = hb_k3_grid(h->piece_n, d, c.nCR, h->sm_count, h->wide ? 1 : 0, &smem);
  if (h->k3_blocks_piece < 1) return fail(h, HB_ERR_UNSUPPORTED, "d = %d with nCR = %d exceeds the shared-memory budget of K3", d, c.nCR);
  h->k3_blocks = h->k3_blocks_piece * h->pieces;        // rows the finalize kernel reduces (rows a short last piece leaves unwritten stay 0)
  CK(dalloc(&h->partials, (size_t)h->k3_blocks * (c.nCR + 2) * d));
  CK(cudaMemsetAsync(h->partials, 0, sizeof(double) * (size_t)h->k3_blocks * (c.nCR + 2) * d, h->stream));
  const long long nr = h->cfg.n_runs;
  CK(dalloc(&h->out_ar, (size_t)nr * h->ar_rows * h->ar_cols)); CK(dalloc(&h->out_cr, (size_t)nr * h->cr_rows * h->cr_cols));
  CK(hb_launch_fill(h->out_ar, nr * h->ar_rows * h->ar_cols, NAN, h->stream));
  CK(hb_launch_fill(h->out_cr, nr * h->cr_rows * h->cr_cols, NAN, h->stream));
  CK(dalloc(&h->proxy, (size_t)nl)); CK(dalloc(&h->proxy_full, (size_t)nc)); CK(dalloc(&h->lpost_full, (size_t)nc));
  CK(dalloc(&h->outl_dev, (size_t)nc)); CK(dalloc(&h->news_dev, (size_t)nc));
  if (!h->modeb) {
    h->outl_tmp_bytes = hb_outlier_workspace_bytes(h->nc);
    CK(cudaMalloc(&h->outl_tmp, h->outl_tmp_bytes ? h->outl_tmp_bytes : 16));
    CK(dalloc(&h->outl_sorted, (size_t)h->nc)); CK(dalloc(&h->outl_thr, 1)); CK(dalloc(&h->outl_count, 1));
  }
  if (h->rstat_rows > 0) { CK(dalloc(&h->rstat_dev, (size_t)h->rstat_rows * nr * d)); CK(hb_launch_fill(h->rstat_dev, h->rstat_rows * nr * d, NAN, h->stream)); }   // [row][run][d]
  CK(dalloc(&h->rstat_xm, (size_t)nl * d)); CK(dalloc(&h->rstat_ss, (size_t)nl * d));
  {  // gamma_d table, R/internals.R:190
    std::vector<double> g((size_t)c.delta * d);
    for (int D = 1; D <= c.delta; ++D)
      for (int ds = 1; ds <= d; ++ds) g[(size_t)(D - 1) * d + (ds - 1)] = 2.38 / sqrt(2.0 * (double)D * (double)ds);
    CK(dalloc(&h->gamma_tab, g.size()));
    CK(cudaMemcpyAsync(h->gamma_tab, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  // control block: pCR = 1/nCR (R/dream_parallel.R:192), counters zero, i_ar = 1 (:250)
  hb_ctrl hc; memset(&hc, 0, sizeof hc);
  for (int q = 0; q < c.nCR; ++q) hc.pCR[q] = 1.0 / c.nCR;
  hc.i_ar = 1; hc.cum_eval = (double)h->nc;
  std::vector<hb_ctrl> hcs((size_t)nr, hc);
  CK(cudaMemcpyAsync(h->ctrl, hcs.data(), sizeof(hb_ctrl) * nr, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->modeb) return hb_seq_allocate(h);
  return HB_OK;
}

// initial population: prior, target, lpost[1,], output row 1   (R/dream_parallel.R:211-250)
int initialise(hb_handle* h) {
  const hb_config& c = h->cfg;
  const long long nl = h->n_local;
  const int d = h->d, ldx = h->ldx;
  const double* Xl = h->X + h->chain0 * ldx;
  unsigned* errp = &h->ctrl->err;
  if (c.prior == HB_PRIOR_NORMAL && c.lik != 22) CK(hb_launch_prior_normal(Xl, ldx, d, nl, h->prior_W, h->prior_mu, h->prior_const, h->lp_xp, errp, h->stream));
  else CK(hb_launch_prior(Xl, ldx, d, nl, c.lik == 22 ? HB_PRIOR_FLAT : c.prior, h->pmin, h->pmax, h->lp_xp, errp, h->stream));
  h->launches++;
  if (int rc = hb_target_launch(h, Xl, nl, h->chain0, 1)) return fail(h, rc, "target launch failed");
  h->launches++;
  CK(hb_launch_init_state(h->lp_xp, h->ll_raw, h->fx_xp, h->lp, h->ll, h->lpost, h->fx, h->lpost_hist, nl, errp, h->stream));
  CK(hb_launch_store_row(Xl, ldx, d, nl, h->lp, h->ll, h->lpost, h->fx, h->hist_x, h->hist_l, h->hist_fx, h->stream));
  h->launches += 2;
  // AR/CR row 1: out_AR[1,1] = out_CR[1,1] = nc ; out_AR[1,2] = NA ; out_CR[1,-1] = pCR   (:245-247)
  const long long nr = c.n_runs;
  const size_t ars = (size_t)h->ar_rows * h->ar_cols, crs = (size_t)h->cr_rows * h->cr_cols;
  std::vector<double> ar(ars * nr, NAN), cr(crs * nr, NAN);
  for (long long r = 0; r < nr; ++r) {
    if (c.sweep == HB_SWEEP_SYNCHRONOUS) {
      ar[r * ars] = (double)h->nc; cr[r * crs] = (double)h->nc;
      for (int q = 0; q < c.nCR; ++q) cr[r * crs + (size_t)h->cr_rows * (q + 1)] = 1.0 / c.nCR;
    } else {   // out_AR[1,] <- rep(0, nCR) ; out_CR[1,] <- pCR   (R/dream.R:193-194)
      for (int q = 0; q < c.nCR; ++q) { ar[r * ars + (size_t)h->ar_rows * q] = 0.0; cr[r * crs + (size_t)h->cr_rows * q] = 1.0 / c.nCR; }
    }
  }
  CK(cudaMemcpyAsync(h->out_ar, ar.data(), ar.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->out_cr, cr.data(), cr.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->gen = 1; h->ind_out = 1; h->initialised = true;
  // sharded population: a failure of this rank's initial evaluation is reported at the end of the slice, after the error
  // masks were OR-ed over the ranks -- returning here would leave the peers waiting at their first barrier
  return h->world > 1 ? HB_OK : check_flags(h);
}

// One generation is built as a PROGRAM: a list of stages, each a list of host steps that enqueue device work.  A stage
// boundary (cut) sits wherever the next step waits for something the PEERS signal (a delta piece flag, the flag of a small
// collective).  Real ranks simply run all stages in order.  A rank emulation on one device (hb_group_run) runs stage s of
// every rank before stage s + 1 of any rank, on one stream, so every signal is enqueued before the wait that needs it and
// no kernel ever spins on another kernel of the same device.
struct gen_prog {
  std::vector<std::vector<std::function<int()>>> stages;
  gen_prog() { stages.emplace_back(); }
  void then(std::function<int()> f) { stages.back().push_back(std::move(f)); }
  void cut() { stages.emplace_back(); }
};

hb_arena_peers arena_peers(const hb_handle* h) {
  hb_arena_peers ap; for (int r = 0; r < 8; ++r) ap.p[r] = h->arena_peer[r];
  return ap;
}

// sum `n` doubles over the ranks, in place (HB_EXCHANGE_DELTA: peer-memory all-reduce in rank order; else NCCL)
int allreduce_post(hb_handle* h, hb_peer_reduce_desc rd) {
  const size_t need = ((size_t)rd.n_f64[0] + rd.n_f64[1] + rd.n_u64) * 8;
  if (need > h->arena_red_bytes) return fail(h, HB_ERR_UNSUPPORTED, "all-reduce of %zu bytes exceeds the arena slot", need);
  ++h->red_seq;
  CK(hb_launch_peer_allreduce_post(arena_peers(h), HB_ARENA_RED_OFF + (size_t)(h->red_seq & 1u) * h->arena_red_bytes, rd, h->rank,
                                   h->world, h->red_seq, h->stream));
  h->launches++;
  return HB_OK;
}
int allreduce_finish(hb_handle* h, hb_peer_reduce_desc rd) {
  CK(hb_launch_peer_allreduce_finish(arena_peers(h), HB_ARENA_RED_OFF + (size_t)(h->red_seq & 1u) * h->arena_red_bytes, rd, h->rank,
                                     h->world, h->red_seq, &h->ctrl->err, h->stream));
  h->launches++;
  return HB_OK;
}
// appended to a program: buf[0..n) summed over the ranks (no-op on one GPU)
void prog_allreduce_f64(hb_handle* h, gen_prog& g, double* buf, int n) {
  if (h->world <= 1) return;
  if (h->delta_mode) {
    hb_peer_reduce_desc rd{};
    rd.f64[0] = buf; rd.n_f64[0] = n; rd.f64[1] = buf; rd.n_f64[1] = 0; rd.u64 = nullptr; rd.n_u64 = 0;
    g.then([=]() -> int { return allreduce_post(h, rd); });
    g.cut();
    g.then([=]() -> int { return allreduce_finish(h, rd); });
  } else {
    g.then([=]() -> int { return hb_nccl_allreduce_f64(h->comm, buf, (size_t)n, h->stream, h->err); });
  }
}

// R_stat(chain[1:t_rows, 1:d, ]) of the whole population into out_dev[d]; on several GPUs the per-chain statistics of
// the shards are combined with two small all-reduces (sums, then squared deviations from the global mean)
void prog_rstat_population(hb_handle* h, gen_prog& g, long long t_rows, double* out_dev) {
  const long long nl = h->n_local; const int d = h->d;
  if (h->world <= 1) {
    g.then([=]() -> int { CK(hb_launch_rstat(h->hist_x, t_rows, nl, 0, nl, d, h->rstat_xm, h->rstat_ss, out_dev, h->stream)); h->launches += 2; return HB_OK; });
    return;
  }
  g.then([=]() -> int {
    if (!h->rstat_sums) CK(dalloc(&h->rstat_sums, (size_t)3 * d));
    CK(hb_launch_rstat_shard_pass1(h->hist_x, t_rows, nl, d, h->rstat_xm, h->rstat_ss, h->rstat_sums, h->stream));
    h->launches++;
    return HB_OK;
  });
  // rstat_sums is allocated by the first step; the all-reduce steps read the pointer when they run
  if (h->delta_mode) {
    g.then([=]() -> int { hb_peer_reduce_desc rd{}; rd.f64[0] = h->rstat_sums; rd.n_f64[0] = 2 * d; rd.f64[1] = h->rstat_sums; return allreduce_post(h, rd); });
    g.cut();
    g.then([=]() -> int { hb_peer_reduce_desc rd{}; rd.f64[0] = h->rstat_sums; rd.n_f64[0] = 2 * d; rd.f64[1] = h->rstat_sums; return allreduce_finish(h, rd); });
  } else g.then([=]() -> int { return hb_nccl_allreduce_f64(h->comm, h->rstat_sums, (size_t)2 * d, h->stream, h->err); });
  g.then([=]() -> int { CK(hb_launch_rstat_shard_dev(h->rstat_xm, nl, d, h->nc, h->rstat_sums, h->rstat_sums + 2 * d, h->stream)); h->launches++; return HB_OK; });
  if (h->delta_mode) {
    g.then([=]() -> int { hb_peer_reduce_desc rd{}; rd.f64[0] = h->rstat_sums + 2 * d; rd.n_f64[0] = d; rd.f64[1] = h->rstat_sums; return allreduce_post(h, rd); });
    g.cut();
    g.then([=]() -> int { hb_peer_reduce_desc rd{}; rd.f64[0] = h->rstat_sums + 2 * d; rd.n_f64[0] = d; rd.f64[1] = h->rstat_sums; return allreduce_finish(h, rd); });
  } else g.then([=]() -> int { return hb_nccl_allreduce_f64(h->comm, h->rstat_sums + 2 * d, (size_t)d, h->stream, h->err); });
  g.then([=]() -> int { CK(hb_launch_rstat_shard_final(h->rstat_sums, h->rstat_sums + 2 * d, t_rows, h->nc, d, out_dev, h->stream)); h->launches++; return HB_OK; });
}

// check_outlier (R/internals.R:71-87) at generation i: proxies, order statistics, the outlier list and the row copies on
// the device; only the without-replacement draw of the replacement chains runs on the host (it runs at most
// adapt*t/updateInterval times per run)
int outlier_select_and_reset(hb_handle* h, long long i) {
  const long long nc = h->nc, nl = h->n_local;
  const double* proxy_all = h->world > 1 ? h->proxy_full : h->proxy;
  {  // quartiles, threshold and the ascending outlier list on the device (R/internals.R:75-78)
    prof_scope ps(h, "OUTLIER");
    CK(hb_outlier_device(proxy_all, nc, h->outl_tmp, h->outl_tmp_bytes, h->outl_sorted, h->outl_thr, h->outl_dev,
                         h->outl_count, h->stream));
    ps.done(3);
  }
  int n_out = 0;
  CK(cudaMemcpyAsync(&n_out, h->outl_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (n_out == 0) return HB_OK;
  std::vector<int> o32((size_t)n_out);
  CK(cudaMemcpy(o32.data(), h->outl_dev, sizeof(int) * n_out, cudaMemcpyDeviceToHost));
  std::vector<long long> outl(o32.begin(), o32.end());
  std::vector<long long> news;
  if (h->cfg.rng_mode == HB_RNG_TAPE) {
    const long long e = i / h->cfg.update_interval - 1;
    if (e < 0 || e >= h->tape_events) return fail(h, HB_ERR_INVALID, "tape has no outlier event %lld", e);
    news.resize(outl.size());
    for (size_t q = 0; q < outl.size(); ++q) {
      news[q] = h->tape_outlier_new[(size_t)e * h->nct + q];
      if (news[q] < 0 || news[q] >= nc) return fail(h, HB_ERR_INVALID, "tape outlier replacement out of range");
    }
  } else {
    sample_replacements(h, i, outl, news);                                        // :81
  }
  std::vector<int> n32(news.begin(), news.end());
  CK(cudaMemcpyAsync(h->news_dev, n32.data(), sizeof(int) * n32.size(), cudaMemcpyHostToDevice, h->stream));
  double* hrow = (i <= h->H) ? h->lpost_hist + (i - 1) * nl : nullptr;
  {
    prof_scope ps(h, "OUTLIER");
    // the snapshot of lpost (all chains, pre-reset) keeps the copies independent of the order they run in
    if (h->world == 1) CK(cudaMemcpyAsync(h->lpost_full, h->lpost, sizeof(double) * nc, cudaMemcpyDeviceToDevice, h->stream));
    if (h->peer_mode) {
      hb_peer_ptrs pp;
      for (int r = 0; r < 8; ++r) pp.p[r] = h->peer_ptr[h->cur][r];
      CK(hb_launch_outlier_reset_peer(h->Xbuf[h->cur], pp, nl, h->ldx, h->d, h->lpost, hrow, h->lpost_full, h->chain0, nl,
                                      h->outl_dev, h->news_dev, (int)o32.size(), h->stream));
    } else
    CK(hb_launch_outlier_reset(h->X, h->ldx, h->d, h->lpost, hrow, h->lpost_full, h->chain0, nl, h->outl_dev, h->news_dev,
                               (int)o32.size(), h->stream));
    ps.done();
  }
  if (h->peer_mode) { if (int rc = hb_nccl_allreduce_u64(h->comm, h->ctrl->n_id_gen, 1, h->stream, h->err)) return rc; }
  CK(cudaStreamSynchronize(h->stream));
  for (long long j : outl) { h->outlier_gen.push_back(i); h->outlier_chain.push_back(j); }   // outl[[i]]  R/dream_parallel.R:418
  return HB_OK;
}

void prog_outlier_check(hb_handle* h, gen_prog& g, long long i) {
  const long long nl = h->n_local;
  const long long r0 = (i + 1) / 2;   // ceiling(i/2), 1-based
  g.then([=]() -> int {
    prof_scope ps(h, "OUTLIER");
    CK(hb_launch_proxy(h->lpost_hist, nl, r0 - 1, i - 1, nl, h->proxy, h->stream));
    ps.done();
    return HB_OK;
  });
  if (h->world > 1 && h->delta_mode) {
    g.then([=]() -> int {
      CK(hb_launch_peer_allgather_post(arena_peers(h), h->arena_gather_off, h->proxy, h->lpost, nl, h->rank, h->world, ++h->gather_seq, h->stream));
      h->launches++;
      return HB_OK;
    });
    g.cut();
    g.then([=]() -> int {
      CK(hb_launch_peer_allgather_finish(arena_peers(h), h->arena_gather_off, h->proxy_full, h->lpost_full, nl, h->rank, h->world,
                                         h->gather_seq, &h->ctrl->err, h->stream));
      h->launches++;
      return HB_OK;
    });
  } else if (h->world > 1) {
    g.then([=]() -> int {
      if (int rc = hb_nccl_allgather_f64(h->comm, h->proxy, h->proxy_full, (size_t)nl, h->stream, h->err)) return rc;
      return hb_nccl_allgather_f64(h->comm, h->lpost, h->lpost_full, (size_t)nl, h->stream, h->err);
    });
  }
  g.then([=]() -> int { return outlier_select_and_reset(h, i); });
}

}  // namespace

// K1 / K3 parameter blocks for a launch over items i -> chain first + i*stride (see hb_k1_params)
hb_k1_params hb_make_k1(hb_handle* h, long long i, long long first, long long stride, long long n_items) {
  const hb_config& c = h->cfg;
  hb_k1_params p{};
  p.X = h->X; p.Z = h->Z; p.m_arch = h->m_arch; p.XP = h->XP; p.id_out = h->id; p.lp_xp = h->lp_xp;
  p.ctrl = h->ctrl; p.err = &h->ctrl->err; p.nc = h->nc; p.chain0 = first; p.stride = stride; p.n_local = n_items;
  p.n_runs = h->cfg.n_runs; p.gchain0 = (long long)c.run_offset * h->nc;
  p.d = h->d; p.ldx = h->ldx; p.delta = c.delta; p.nCR = c.nCR; p.c_val = c.c_val; p.c_star = c.c_star; p.p_g = c.p_g;
  p.beta0 = c.beta0; p.psnooker = c.psnooker; p.past_sample = c.past_sample;
  p.bound = h->have_min_max ? c.bound : 0; p.prior = (c.lik == 22) ? HB_PRIOR_FLAT : c.prior;   // prior_pdf: lik == 22 -> 0, R/internals.R:50-51
  p.pmin = h->pmin; p.pmax = h->pmax;
  p.seed = c.seed; p.gen = (uint32_t)i; p.tape = h->tape; p.tape_g = i - 2; p.gamma_tab = h->gamma_tab;
  p.n_peers = h->peer_mode ? h->world : 1; p.peer_nl = h->n_local;
  for (int r = 0; r < 8; ++r) p.peer[r] = h->peer_mode ? h->peer_ptr[h->cur][r] : nullptr;
  p.force_wide = (h->delta_mode && h->wide) ? 1 : 0; p.batch = 32; p.reserve_ctas = 0;
  hb_philox_key_schedule(c.seed, p.ks);
  for (int q = 0; q < c.nCR; ++q) p.thr16[q] = hb_rng_zt_threshold16((double)(q + 1) / (double)c.nCR);
  return p;
}

hb_k3_params hb_make_k3(hb_handle* h, long long i, long long first, long long stride, long long n_items, bool store,
                        bool in_adapt, bool write_dx) {
  const hb_config& c = h->cfg;
  const long long nl = h->n_local;
  const int d = h->d;
  hb_k3_params p{};
  p.X = h->X; p.Xout = h->peer_mode ? h->Xbuf[h->cur ^ 1] - h->chain0 * h->ldx : h->X; p.XP = h->XP; p.XP_rw = h->XP; p.id = h->id; p.lp_xp = h->lp_xp; p.ll_raw = h->ll_raw;
  p.fx_xp = c.store_fx ? h->fx_xp : nullptr;
  p.lp = h->lp; p.ll = h->ll; p.lpost = h->lpost; p.fx = c.store_fx ? h->fx : nullptr;
  p.lpost_hist_row = (i <= h->H) ? h->lpost_hist + (i - 1) * nl : nullptr;
  if (store) {
    p.hist_x = h->hist_x + (size_t)(h->ind_out - 1) * nl * d;
    p.hist_l = h->hist_l + (size_t)(h->ind_out - 1) * nl * 3;
    p.hist_fx = h->hist_fx ? h->hist_fx + (size_t)(h->ind_out - 1) * nl : nullptr;
  }
  p.accept_rec = h->accept_rec ? h->accept_rec + (size_t)(i - 2) * nl : nullptr;
  p.ctrl = h->ctrl; p.shift = h->shift; p.partials = h->partials;
  p.chain0 = first; p.stride = stride; p.n_local = n_items; p.nc = h->nc; p.n_runs = h->cfg.n_runs;
  p.state0 = h->chain0; p.gchain0 = (long long)c.run_offset * h->nc;
  p.force_wide = (h->delta_mode && h->wide) ? 1 : 0;
  p.d = d; p.ldx = h->ldx; p.nCR = c.nCR;
  p.adapt = in_adapt; p.write_dx = write_dx; p.lik22 = (c.lik == 22); p.seed = c.seed; p.gen = (uint32_t)i;
  p.u_accept = (c.rng_mode == HB_RNG_TAPE) ? h->tape.u_accept + (size_t)(i - 2) * h->tape.nct : nullptr;
  return p;
}

// `off`: the launch covers proposals XP[off ...] of the shard (one piece); ll_raw / fx_xp are written at the same offset
int hb_target_launch(hb_handle* h, const double* x, long long n, long long first, long long stride, long long off) {
  if (h->vt->launch_items && h->cfg.n_runs > 1)
    return h->vt->launch_items(h->tctx, x, n, h->d, h->ldx, h->ll_raw + off, h->fx_xp + off, first, stride, h->nc, h->stream);
  return h->vt->launch(h->tctx, x, n, h->d, h->ldx, h->ll_raw + off, h->fx_xp + off, h->stream);
}

// prior = "normal": the proposal kernel leaves lp_xp = 0; the quadratic form runs as its own launch over XP
int hb_prior_after_k1(hb_handle* h, long long n_items, long long off) {
  if (h->cfg.prior != HB_PRIOR_NORMAL || h->cfg.lik == 22) return HB_OK;
  if (cudaError_t e = hb_launch_prior_normal(h->XP + off * h->ldx, h->ldx, h->d, n_items, h->prior_W, h->prior_mu, h->prior_const,
                                             h->lp_xp + off, &h->ctrl->err, h->stream); e != cudaSuccess) { h->err = cudaGetErrorString(e); return HB_ERR_CUDA; }
  h->launches++;
  return HB_OK;
}

int hb_fail(hb_handle* h, int code, const char* msg) { if (h) h->err = msg; return code; }

namespace {

// The fused generation kernel applies to: dream_parallel's synchronous sweep of one population on one GPU, the built-in
// t_distribution target with lik = 2, flat / uniform prior, no archive, no multi-try, populations below the size where the
// large-population kernels take over (hb_fused_small_eligible).  HB_FUSED=0 switches it off.
bool fused_available(hb_handle* h) {
  const hb_config& c = h->cfg;
  if (h->fused_state == 0) {
    h->fused_state = -1;
    const char* e = getenv("HB_FUSED");                        // "1" on, "0" off, unset: kFusedDefault
    constexpr bool kFusedDefault = true;
    const bool want = e ? (e[0] == '1') : kFusedDefault;
    if (want && c.sweep == HB_SWEEP_SYNCHRONOUS && !h->modeb && c.mt <= 1 && h->world == 1 && !h->peer_mode &&
        !h->delta_mode && !c.past_sample && c.lik == 2 && c.prior != HB_PRIOR_NORMAL && h->X &&
        hb_tdist_fused_info(h->vt, h->tctx, &h->fused_td)) {
      if (cudaMalloc((void**)&h->Xalt, (size_t)h->nc * h->ldx * sizeof(double)) == cudaSuccess &&
          cudaMemsetAsync(h->Xalt, 0, (size_t)h->nc * h->ldx * sizeof(double), h->stream) == cudaSuccess)
        h->fused_state = 1;
      else cudaGetLastError();
    }
  }
  return h->fused_state == 1 && !h->profile && hb_fused_small_eligible(h->n_local, h->d, h->sm_count);
}

// The built-in t_distribution target in its structured evaluation: hb_k1_pair_kernel (large populations, native draws) can
// fold the log-density into the proposal sweep.  HB_K1_DENSITY=0 keeps the plugin launch (the cross-check of the tests).
bool k1_density_available(hb_handle* h) {
  const hb_config& c = h->cfg;
  if (h->k1_td_state == 0) {
    h->k1_td_state = -1;
    const char* e = getenv("HB_K1_DENSITY");
    const bool want = e ? (e[0] == '1') : true;
    if (want && c.sweep == HB_SWEEP_SYNCHRONOUS && !h->modeb && c.mt <= 1 && c.n_runs <= 1 && c.rng_mode != HB_RNG_TAPE &&
        hb_tdist_fused_info(h->vt, h->tctx, &h->k1_td))
      h->k1_td_state = 1;
  }
  return h->k1_td_state == 1;
}

// Persistent generation loop (hb_persist_small_kernel): generations [i0, i1) of a small population, all behind the adaptation
// period, in ONE cooperative launch.  Returns HB_OK with *taken = false when the handle cannot use it (the caller then runs
// the generations one by one).
int run_persistent(hb_handle* h, long long i0, long long i1, bool* taken, bool adapt = false) {
  const hb_config& c = h->cfg;
  *taken = false;
  if (h->persist_state < 0 || i1 <= i0 || c.check_convergence) return HB_OK;
  if (h->persist_state == 0) {
    const char* e = getenv("HB_PERSIST");                      // "0" off
    if (e && e[0] == '0') { h->persist_state = -1; return HB_OK; }
    if (cudaMalloc((void**)&h->grid_bar, 2 * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); h->persist_state = -1; return HB_OK; }
    h->persist_state = 1;
  }
  const long long nl = h->n_local;
  long long n_store = 0;
  for (long long i = i0; i < i1; ++i) if (((double)i > c.burnin * (double)h->t) && (i % c.thin == 0)) ++n_store;   // :263, :386
  if (h->ind_out + n_store > h->out_t) return fail(h, HB_ERR_INVALID, "subscript out of bounds (non-integer out_t, R/dream_parallel.R:195)");
  const bool tape_mode = (c.rng_mode == HB_RNG_TAPE);
  const hb_k1_params p = hb_make_k1(h, i0, h->chain0, 1, nl);
  hb_k3_params q = hb_make_k3(h, i0, h->chain0, 1, nl, false, adapt, false);
  q.Xout = h->Xalt;
  hb_persist_params pp{};
  if (adapt) {
    const size_t E = (size_t)(c.nCR + 2) * h->d;
    if (h->islot + (i1 - i0) > c.update_interval) return fail(h, HB_ERR_STATE, "persistent adaptation launch crosses an updateInterval boundary");
    if (!h->persist_partials) CK(dalloc(&h->persist_partials, (size_t)c.update_interval * ((nl + 7) / 8) * E));
    if (!h->persist_ibuf) CK(dalloc(&h->persist_ibuf, (size_t)c.update_interval * E));
    pp.adapt = 1; pp.slot0 = h->islot; pp.partials = h->persist_partials; pp.shift = h->shift;
  }
  pp.i0 = i0; pp.i1 = i1; pp.XA = h->X; pp.XB = h->Xalt; pp.t = h->t; pp.burnin = c.burnin; pp.thin = c.thin;
  pp.update_interval = c.update_interval; pp.ind_out0 = h->ind_out;
  pp.hist_x = h->hist_x; pp.hist_l = h->hist_l; pp.hist_fx = h->hist_fx; pp.lpost_hist = h->lpost_hist; pp.H = h->H;
  pp.accept_rec = h->accept_rec; pp.u_accept = tape_mode ? h->tape.u_accept : nullptr; pp.tape_nct = h->tape.nct;
  pp.pending0 = h->pending_evals; pp.out_ar = h->out_ar; pp.out_cr = h->out_cr; pp.ar_rows = h->ar_rows; pp.bar = h->grid_bar;
  { const char* e = getenv("HB_PERSIST_TRACE"); pp.trace = (e && e[0] == '1') ? 1 : 0; }
  CK(cudaMemsetAsync(h->grid_bar, 0, 2 * sizeof(unsigned), h->stream));
  const cudaError_t e = hb_launch_persist_small(p, q, h->fused_td, pp, tape_mode, h->sm_count, h->stream);
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorInvalidValue || e == cudaErrorNotSupported) {
    cudaGetLastError(); h->persist_state = -1; return HB_OK;   // population too large for one co-resident grid: per-generation launches
  }
  if (e != cudaSuccess) return fail(h, HB_ERR_CUDA, "persistent generation kernel: %s", cudaGetErrorString(e));
  *taken = true;
  h->launches++;
  h->ind_out += n_store;
  long long last_upd = ((i1 - 1) / c.update_interval) * c.update_interval;       // largest multiple of updateInterval below i1
  if (adapt) { h->islot += (int)(i1 - i0); h->pending_evals += h->nc * (i1 - i0); }   // folded by persist_fold
  else h->pending_evals = (last_upd >= i0) ? h->nc * (i1 - 1 - last_upd) : h->pending_evals + h->nc * (i1 - i0);
  if ((i1 - i0) & 1) std::swap(h->X, h->Xalt);
  h->gen = i1 - 1;
  return HB_OK;
}

// the finalize step for the adaptation generations a persistent launch has parked (their per-CTA statistic sums sit in
// slots 0 .. islot-1): reduce every slot, then J / n_id / n_acc in generation order, pCR and the AR / CR rows if the last
// one is an updateInterval boundary (hb_finalize_kernel's interval mode, shared with the multi-GPU path)
int persist_fold(hb_handle* h, bool upd) {
  if (h->islot <= 0) return HB_OK;
  const hb_config& c = h->cfg;
  const size_t E = (size_t)(c.nCR + 2) * h->d;
  hb_fin_params f{};
  f.ctrl = h->ctrl; f.partials = h->persist_partials; f.nblocks = (int)((h->n_local + 7) / 8); f.reduce_slots = h->islot;
  f.shift = h->shift; f.colsum = h->persist_ibuf; f.qsum = h->persist_ibuf + 2 * h->d; f.nc = h->nc; f.d = h->d; f.nCR = c.nCR;
  f.adapt = 1; f.do_pcr = upd; f.do_arcr = upd; f.out_ar = h->out_ar; f.out_cr = h->out_cr; f.ar_rows = h->ar_rows;
  f.n_eval_inc = h->pending_evals; f.n_slots = h->islot; f.slot_stride = (long long)E; f.phase = 2;
  CK(hb_launch_finalize(f, h->stream));
  h->launches += 2;
  h->islot = 0; h->pending_evals = 0;
  return HB_OK;
}

// MT-DREAM, R/dream_parallel.R:269-341 (hb_mt.cu): the whole generation as one host step
int mt_generation(hb_handle* h, long long i, bool store, bool in_adapt, bool tape_mode) {
  const hb_config& c = h->cfg;
  const long long nl = h->n_local;
  const int ldx = h->ldx;
  const int mt = c.mt, n_sets = 2 * mt - 1;
  auto k1_set = [&](int set, int slot, const double* pop) -> int {   // proposal set `set` into try slot `slot`
    hb_k1_params p = hb_make_k1(h, i, h->chain0, 1, nl);
    p.X = pop;
    p.XP = h->XP + (size_t)slot * nl * ldx; p.id_out = h->id + (size_t)slot * nl; p.lp_xp = h->lp_xp + (size_t)slot * nl;
    p.gen = (uint32_t)i | ((uint32_t)(set + 1) << 24);              // Philox substream of the set
    p.tape_g = (i - 2) * n_sets + set;                              // tape row g*n_sets + set
    prof_scope ps(h, "K1");
    CK(hb_launch_k1(p, tape_mode, h->sm_count, h->stream));
    ps.done();
    return HB_OK;
  };
  double* const ZTl = h->world > 1 ? h->ZT + (size_t)h->chain0 * ldx : h->ZT;               // this rank's rows of the candidate population
  const double* const ZTpop = h->world > 1 ? h->ZT : h->ZT - (size_t)h->chain0 * ldx;     // indexed by the chain number, like X
  for (int m = 0; m < mt; ++m) if (int rc = k1_set(m, m, h->X)) return rc;                 // :269  around xt
  if (int rc = hb_prior_after_k1(h, (long long)mt * nl)) return rc;
  { prof_scope ps(h, "K2"); if (int rc = hb_target_launch(h, h->XP, (long long)mt * nl, h->chain0, 1)) return fail(h, rc, "target launch failed"); ps.done(); }
  hb_mt_params mp{};
  mp.mt = mt; mp.n_local = nl; mp.ldx = ldx; mp.XP = h->XP; mp.lp_xp = h->lp_xp; mp.ll_raw = h->ll_raw;
  mp.fx_xp = c.store_fx ? h->fx_xp : nullptr; mp.id = h->id; mp.ZT = ZTl; mp.lp_zt = h->lp_zt; mp.ll_zt = h->ll_zt;
  mp.fx_zt = c.store_fx ? h->fx_zt : nullptr; mp.id_zt = h->id_zt; mp.p1 = h->mt_p1; mp.lpost = h->lpost;
  mp.accept_out = h->mt_accept; mp.seed = c.seed; mp.gen = (uint32_t)i; mp.gchain0 = (long long)c.run_offset * h->nc + h->chain0;
  mp.err = &h->ctrl->err;
  mp.tape_pick = tape_mode ? h->tape_mt_pick + (size_t)(i - 2) * h->tape.nct + h->chain0 : nullptr;   // indexed by the local chain
  mp.u_accept = tape_mode ? h->tape.u_accept + (size_t)(i - 2) * h->tape.nct + h->chain0 : nullptr;
  { prof_scope ps(h, "MT"); CK(hb_launch_mt_select(mp, h->stream)); ps.done(); }           // :297-314
  if (h->world > 1) {   // every rank needs every candidate row: partners of the reference proposals come from all of zt
    prof_scope ps(h, "XCHG");
    if (int rc = hb_nccl_allgather_f64(h->comm, ZTl, h->ZT, (size_t)nl * ldx, h->stream, h->err)) return rc;
    ps.done();
  }
  for (int m = 0; m < mt - 1; ++m) if (int rc = k1_set(mt + m, m, ZTpop)) return rc;       // :316  around zt
  if (int rc = hb_prior_after_k1(h, (long long)(mt - 1) * nl)) return rc;
  { prof_scope ps(h, "K2"); if (int rc = hb_target_launch(h, h->XP, (long long)(mt - 1) * nl, h->chain0, 1)) return fail(h, rc, "target launch failed"); ps.done(); }
  { prof_scope ps(h, "MT"); CK(hb_launch_mt_accept(mp, h->stream)); ps.done(); }           // :340, R/internals.R:229-236
  hb_k3_params p = hb_make_k3(h, i, h->chain0, 1, nl, store, in_adapt, false);
  p.XP = ZTl; p.XP_rw = ZTl; p.id = h->id_zt; p.lp_xp = h->lp_zt; p.ll_raw = h->ll_zt;   // :349-355 accept the candidates
  p.fx_xp = c.store_fx ? h->fx_zt : nullptr; p.accept_in = h->mt_accept;
  prof_scope ps(h, "K3");
  int nb = 0;
  CK(hb_launch_k3(p, h->sm_count, &nb, nullptr, h->stream));
  ps.done();
  return HB_OK;
}

// K1 -> K2 -> K3 over chains [off, off + n) of the shard (the whole shard, or one piece of it in delta mode)
int piece_kernels(hb_handle* h, long long i, int piece, long long off, long long n, bool store, bool in_adapt, bool tape_mode,
                  cudaEvent_t ev_after_k1 = nullptr) {
  const hb_config& c = h->cfg;
  const int d = h->d, ldx = h->ldx;
  bool density_done = false;
  {  // K1
    hb_k1_params p = hb_make_k1(h, i, h->chain0 + off, 1, n);
    p.XP += off * ldx; p.id_out += off; p.lp_xp += off;
    p.reserve_ctas = h->reserve_ctas;
    if (k1_density_available(h)) {   // built-in t_distribution, structured evaluation: the large-population proposal kernel writes ll too
      p.td_rs = h->k1_td.rs; p.td_konst = h->k1_td.konst; p.td_df = h->k1_td.df; p.td_lik = h->k1_td.lik;
      p.td_ll = h->ll_raw + off; p.td_fx = h->fx_xp + off;
    }
    prof_scope ps(h, "K1");
    CK(hb_launch_k1(p, tape_mode, h->sm_count, h->stream, &density_done));
    ps.done();
    if (ev_after_k1) CK(cudaEventRecord(ev_after_k1, h->stream));
    if (int rc = hb_prior_after_k1(h, n, off)) return rc;
  }
  if (!density_done) {  // K2: the target plugin
    prof_scope ps(h, "K2");
    if (int rc = hb_target_launch(h, h->XP + off * ldx, n, h->chain0 + off, 1, off)) return fail(h, rc, "target launch failed");
    ps.done();
  }
  {  // K3
    hb_k3_params p = hb_make_k3(h, i, h->chain0 + off, 1, n, store, in_adapt, false);
    p.XP += off * ldx; p.XP_rw += off * ldx; p.id += off; p.lp_xp += off; p.ll_raw += off;
    if (p.fx_xp) p.fx_xp += off;
    p.partials = h->partials + (size_t)piece * h->k3_blocks_piece * (c.nCR + 2) * d;
    if (h->delta_mode) {   // accepted rows go to this rank's delta log: box (parity, self), slots of this piece
      unsigned char* box = h->boxmem[h->rank] + (size_t)(i & 1) * h->box_bytes;
      p.log_rows = reinterpret_cast<double*>(box + h->box_rows_off) + (size_t)piece * h->piece_n * ldx;
      p.log_idx = reinterpret_cast<int*>(box + h->box_idx_off) + (size_t)piece * h->piece_n;
      p.log_count = h->log_ctr + (size_t)(i & 1) * HB_MAX_PIECES + piece;
    }
    prof_scope ps(h, "K3");
    int nb = 0;
    CK(hb_launch_k3(p, h->sm_count, &nb, nullptr, h->stream));
    ps.done();
  }
  return HB_OK;
}

// ship piece `piece` of generation i's delta log to every peer (side stream: overlaps the next piece's kernels)
int piece_send(hb_handle* h, long long i, int piece) {
  hb_delta_send_params sp{};
  for (int r = 0; r < h->world; ++r) {
    sp.arena[r] = h->arena_peer[r];
    sp.box[r] = h->peer_box[r] + (size_t)(i & 1) * h->box_bytes;        // rank r's box for this rank's log
  }
  sp.idx_off = h->box_idx_off; sp.rows_off = h->box_rows_off; sp.piece_off = (long long)piece * h->piece_n;
  sp.piece = piece; sp.ldx = h->ldx; sp.self = h->rank; sp.world = h->world;
  sp.log_count = h->log_ctr + (size_t)(i & 1) * HB_MAX_PIECES + piece;
  sp.done_ctr = h->log_ctr + 2 * HB_MAX_PIECES + piece;
  sp.flag_value = (unsigned)i * HB_MAX_PIECES + (unsigned)piece + 1u;
  if (h->side != h->stream) {
    CK(cudaEventRecord(h->ev_piece[piece], h->stream));
    CK(cudaStreamWaitEvent(h->side, h->ev_piece[piece], 0));
  }
  prof_scope ps(h, "XCHG", h->side);
  CK(hb_launch_delta_send(sp, h->send_ctas, h->side));
  ps.done();
  return HB_OK;
}

// fold the delta logs of pieces [p0, p1) of generation i -- this rank's and every peer's -- into the replica
int delta_apply(hb_handle* h, long long i, int p0, int p1, cudaStream_t st) {
  hb_delta_apply_params ap{};
  for (int r = 0; r < h->world; ++r) ap.box[r] = h->boxmem[r] + (size_t)(i & 1) * h->box_bytes;
  ap.idx_off = h->box_idx_off; ap.rows_off = h->box_rows_off; ap.piece_cap = h->piece_n; ap.p0 = p0; ap.p1 = p1;
  ap.ldx = h->ldx; ap.self = h->rank; ap.world = h->world; ap.nc = h->nc; ap.X = h->X;
  ap.flags = reinterpret_cast<const unsigned*>(h->arena); ap.flag_value = (unsigned)i * HB_MAX_PIECES + (unsigned)p1;
  ap.err = &h->ctrl->err; ap.timeout_ns = hb_peer_timeout_ns();
  prof_scope ps(h, "APPLY", st);
  CK(hb_launch_delta_apply(ap, h->sm_count, st));
  ps.done();
  return HB_OK;
}

// store the rows of piece `piece` of this rank's log straight into every replica (the last piece of a generation)
int delta_push(hb_handle* h, long long i, int piece) {
  hb_delta_push_params pp{};
  pp.box = h->boxmem[h->rank] + (size_t)(i & 1) * h->box_bytes;
  pp.idx_off = h->box_idx_off; pp.rows_off = h->box_rows_off; pp.piece_off = (long long)piece * h->piece_n;
  pp.ldx = h->ldx; pp.self = h->rank; pp.world = h->world; pp.nc = h->nc;
  for (int r = 0; r < h->world; ++r) { pp.X[r] = (r == h->rank) ? h->X : h->peerX[r]; pp.arena[r] = h->arena_peer[r]; }
  pp.log_count = h->log_ctr + (size_t)(i & 1) * HB_MAX_PIECES + piece;
  pp.done_ctr = h->log_ctr + 2 * HB_MAX_PIECES + piece;
  pp.flags = reinterpret_cast<const unsigned*>(h->arena); pp.gen = (unsigned)i;
  pp.err = &h->ctrl->err; pp.timeout_ns = hb_peer_timeout_ns();
  prof_scope ps(h, "XCHG");
  CK(hb_launch_delta_push(pp, h->sm_count, h->stream));
  ps.done();
  return HB_OK;
}

hb_fin_params make_fin(hb_handle* h, bool in_adapt, bool upd) {
  const hb_config& c = h->cfg;
  hb_fin_params f{};
  f.ctrl = h->ctrl; f.partials = h->partials; f.nblocks = h->k3_blocks; f.shift = h->shift; f.colsum = h->colsum;
  f.qsum = h->qsum; f.nc = h->nc; f.d = h->d; f.nCR = c.nCR; f.adapt = in_adapt; f.do_pcr = in_adapt && upd;
  f.do_arcr = upd; f.out_ar = h->out_ar; f.out_cr = h->out_cr; f.ar_rows = h->ar_rows;
  f.n_eval_inc = h->pending_evals; f.n_slots = 1; f.slot_stride = 0;
  return f;
}

// the program of generation i of the synchronous sweep (R/dream_parallel.R:257-435)
int build_generation(hb_handle* h, long long i, gen_prog& g) {
  const hb_config& c = h->cfg;
  const long long nl = h->n_local;
  const int d = h->d, ldx = h->ldx;
  const bool store = ((double)i > c.burnin * (double)h->t) && (i % c.thin == 0);       // :263, :386
  if (store) {
    h->ind_out++;
    if (h->ind_out > h->out_t) return fail(h, HB_ERR_INVALID, "subscript out of bounds (non-integer out_t, R/dream_parallel.R:195)");
  }
  const bool in_adapt = ((double)i <= c.adapt * (double)h->t);                         // :403
  const bool upd = (i % c.update_interval == 0);
  const bool tape_mode = (c.rng_mode == HB_RNG_TAPE);

  if (h->modeb) { g.then([=]() -> int { return hb_seq_generation(h, i); }); return HB_OK; }
  if (c.mt > 1) {
    g.then([=]() -> int { return mt_generation(h, i, store, in_adapt, tape_mode); });
  } else if (!in_adapt && fused_available(h)) {
    // small population outside the adaptation period: proposal + log-density + accept in one launch, the post-generation
    // rows go to the second buffer (no grid-wide barrier between the partner reads and the state update)
    g.then([=]() -> int {
      const hb_k1_params p = hb_make_k1(h, i, h->chain0, 1, nl);
      hb_k3_params q = hb_make_k3(h, i, h->chain0, 1, nl, store, false, false);
      q.Xout = h->Xalt;
      CK(hb_launch_fused_small(p, q, h->fused_td, tape_mode, h->sm_count, h->stream));
      h->launches++;
      std::swap(h->X, h->Xalt);
      return HB_OK;
    });
  } else if (h->delta_mode) {
    // sharded population, stage-and-apply exchange (hb_delta.cu): piece p's log travels while piece p + 1 computes
    const int P = h->pieces;
    // Pieces 0 .. P-2: K3 logs the accepted rows, the send kernel ships the log while the next piece computes, and as soon
    // as this rank's LAST proposal kernel is done (nothing reads the generation-start state any more) the logs that have
    // arrived are folded into the replica on a second side stream, behind the last piece's K2 / K3.
    // Last piece: flag 5 tells the peers that this rank's proposals are done; the push kernel waits for every peer's flag 5
    // and stores the rows of the last piece straight into every replica (no log on the receiving side, no apply pass),
    // then raises flag 6; the generation ends when every peer's flag 6 has arrived.
    const bool conc = h->side != h->stream;                 // real ranks; a rank emulation runs everything on one stream
    g.then([=]() -> int {
      for (int p = 0; p < P; ++p) {
        const long long off = (long long)p * h->piece_n, n = std::min(h->piece_n, nl - off);
        const bool last = (p == P - 1);
        if (n > 0) { if (int rc = piece_kernels(h, i, p, off, n, store, in_adapt, tape_mode, (conc && last) ? h->ev_k1last : nullptr)) return rc; }
        else if (conc && last) CK(cudaEventRecord(h->ev_k1last, h->stream));
        if (!last) {
          if (int rc = piece_send(h, i, p)) return rc;      // an empty piece still raises its flag
          if (conc && p == P - 2) CK(cudaEventRecord(h->ev_sent, h->side));
        } else {
          if (conc) CK(cudaStreamWaitEvent(h->side2, h->ev_k1last, 0));   // every proposal of this rank has read its partner rows
          CK(hb_launch_flag_signal(arena_peers(h), 5, h->rank, h->world, (unsigned)i, h->side2));
          h->launches++;
          if (conc && P > 1) {
            CK(cudaStreamWaitEvent(h->side2, h->ev_sent, 0));            // this rank's own logs (and counts) of pieces < P - 1
            if (int rc = delta_apply(h, i, 0, P - 1, h->side2)) return rc;
          }
          if (conc) CK(cudaEventRecord(h->ev_applied, h->side2));
        }
      }
      return HB_OK;
    });
    g.cut();
    g.then([=]() -> int {
      if (!conc && P > 1) { if (int rc = delta_apply(h, i, 0, P - 1, h->stream)) return rc; }
      if (int rc = delta_push(h, i, P - 1)) return rc;
      if (conc) CK(cudaStreamWaitEvent(h->stream, h->ev_applied, 0));
      return HB_OK;
    });
    g.cut();
    g.then([=]() -> int {
      CK(hb_launch_flag_wait(h->arena, 6, h->rank, h->world, (unsigned)i, &h->ctrl->err, h->stream));
      h->launches++;
      return HB_OK;
    });
  } else {
    g.then([=]() -> int { return piece_kernels(h, i, 0, 0, nl, store, in_adapt, tape_mode); });
  }
  h->pending_evals += h->nc;

  // finalize: J / n_id / n_acc, pCR adaptation, AR/CR rows
  if (h->delta_mode) {
    // the adaptation sums of every generation are parked in a slot of the interval buffer; one all-reduce per
    // updateInterval carries them (and the integer counters), then they are folded in generation order
    const int E = (c.nCR + 2) * d;
    if (in_adapt) {
      const int slot = h->islot++;
      g.then([=]() -> int {
        hb_fin_params f = make_fin(h, true, false);
        f.phase = 0; f.colsum = h->ibuf + (size_t)slot * E; f.qsum = f.colsum + 2 * d;
        prof_scope ps(h, "FIN");
        CK(hb_launch_finalize(f, h->stream));
        ps.done();
        return HB_OK;
      });
    }
    if (upd) {
      const int n_slots = in_adapt ? h->islot : 0;
      h->islot = 0;
      hb_fin_params f = make_fin(h, in_adapt, true);
      f.phase = 1; f.colsum = h->ibuf; f.qsum = h->ibuf + 2 * d; f.n_slots = n_slots; f.slot_stride = E;
      h->pending_evals = 0;
      hb_peer_reduce_desc rd{};
      rd.f64[0] = h->ibuf; rd.n_f64[0] = n_slots * E; rd.f64[1] = h->ibuf; rd.n_f64[1] = 0;
      rd.u64 = h->ctrl->n_id_gen; rd.n_u64 = HB_MAX_NCR + 1;
      g.then([=]() -> int { return allreduce_post(h, rd); });
      g.cut();
      g.then([=]() -> int {
        if (int rc = allreduce_finish(h, rd)) return rc;
        prof_scope ps(h, "FIN");
        CK(hb_launch_finalize(f, h->stream));
        ps.done();
        return HB_OK;
      });
    }
  } else if (in_adapt || upd || h->peer_mode) {   // (peer mode: every generation, its counter all-reduce is the generation barrier)
    hb_fin_params f = make_fin(h, in_adapt, upd);
    h->pending_evals = 0;
    g.then([=]() -> int {
      hb_fin_params ff = f;
      prof_scope ps(h, "FIN");
      if (h->world > 1) {
        ff.phase = 0; CK(hb_launch_finalize(ff, h->stream));
        if (in_adapt) {
          if (int rc = hb_nccl_allreduce_f64(h->comm, h->colsum, (size_t)2 * d, h->stream, h->err)) return rc;
          if (int rc = hb_nccl_allreduce_f64(h->comm, h->qsum, (size_t)c.nCR * d, h->stream, h->err)) return rc;
        }
        if (int rc = hb_nccl_allreduce_u64(h->comm, h->ctrl->n_id_gen, HB_MAX_NCR + 1, h->stream, h->err)) return rc;
        ff.phase = 1; CK(hb_launch_finalize(ff, h->stream));
        ps.done(2);
      } else {
        ff.phase = 2; CK(hb_launch_finalize(ff, h->stream));
        ps.done();
      }
      return HB_OK;
    });
  }
  if (h->peer_mode) {   // K3 wrote every row of the shard into the other buffer; peers read it next generation
    g.then([=]() -> int { h->cur ^= 1; h->X = h->Xbuf[h->cur] - h->chain0 * ldx; return HB_OK; });
  } else if (h->world > 1 && !h->delta_mode) {  // chain-state exchange: every rank needs every row for the next generation's partner draws
    g.then([=]() -> int {
      prof_scope ps(h, "XCHG");
      if (int rc = hb_nccl_allgather_f64(h->comm, h->X + h->chain0 * ldx, h->X, (size_t)nl * ldx, h->stream, h->err)) return rc;
      ps.done();
      return HB_OK;
    });
  }
  if (c.past_sample && h->archive_update > 0 && (i % h->archive_update == 0)) {     // :382-383  z <- rbind(z, xt)
    g.then([=]() -> int {
      if (h->m_arch + h->nc > h->m_cap) return fail(h, HB_ERR_STATE, "archive capacity exceeded");
      CK(cudaMemcpyAsync(h->Z + (size_t)h->m_arch * ldx, h->X, sizeof(double) * (size_t)h->nc * ldx, cudaMemcpyDeviceToDevice, h->stream));
      h->m_arch += h->nc;
      return HB_OK;
    });
  }
  if (c.check_convergence && store && h->ind_out > 50)   // R_stat(chain[1:ind_out, 1:d, ]) as R/dream_test.R:344
    prog_rstat_population(h, g, h->ind_out, h->rstat_dev + (size_t)(h->ind_out - 51) * d);
  if (in_adapt && upd && !c.past_sample && c.outlier_check)                          // :414-419
    prog_outlier_check(h, g, i);
  g.then([=]() -> int { h->gen = i; return HB_OK; });
  return HB_OK;
}

// end of a slice on several GPUs: OR the ranks' error masks so that every rank returns the same status (a shard-local
// condition -- a non-finite density, a bad tape record -- must not leave the other ranks waiting in the next slice)
void prog_err_reduce(hb_handle* h, gen_prog& g) {
  if (h->world <= 1) return;
  if (h->delta_mode) {
    g.then([=]() -> int {
      unsigned extra = 0u;
      if (h->target_is_hydro) { CK(cudaStreamSynchronize(h->stream)); extra = hb_hydro_poll_err(h->tctx); }
      CK(hb_launch_peer_err_post(arena_peers(h), HB_ARENA_ERR_OFF, &h->ctrl->err, extra, h->rank, h->world, ++h->err_seq, h->stream));
      return HB_OK;
    });
    g.cut();
    g.then([=]() -> int { CK(hb_launch_peer_err_finish(arena_peers(h), HB_ARENA_ERR_OFF, &h->ctrl->err, h->rank, h->world, h->err_seq, h->stream)); return HB_OK; });
  } else {
    g.then([=]() -> int {
      unsigned extra = 0u;
      if (h->target_is_hydro) { CK(cudaStreamSynchronize(h->stream)); extra = hb_hydro_poll_err(h->tctx); }
      CK(hb_launch_err_expand(&h->ctrl->err, extra, h->errbits_dev, h->stream));
      if (int rc = hb_nccl_allreduce_u64(h->comm, h->errbits_dev, 16, h->stream, h->err)) return rc;
      CK(hb_launch_err_collapse(&h->ctrl->err, h->errbits_dev, h->stream));
      return HB_OK;
    });
  }
}

// start of a slice in delta mode: one flag barrier absorbs the host-side skew between the ranks (allocation, page
// faults, garbage collection between slices), so the waits inside the slice only ever see in-slice skew
void prog_slice_barrier(hb_handle* h, gen_prog& g) {
  if (h->world <= 1 || !h->delta_mode) return;
  g.then([=]() -> int { CK(hb_launch_flag_signal(arena_peers(h), 0, h->rank, h->world, ++h->slice_seq, h->stream)); h->launches++; return HB_OK; });
  g.cut();
  g.then([=]() -> int { CK(hb_launch_flag_wait(h->arena, 0, h->rank, h->world, h->slice_seq, &h->ctrl->err, h->stream)); h->launches++; return HB_OK; });
}

// runs a set of programs (one per rank handle) stage by stage: stage s of every rank before stage s + 1 of any
int run_programs(std::vector<hb_handle*>& hs, std::vector<gen_prog>& gs) {
  const size_t n_st = gs[0].stages.size();
  for (size_t r = 1; r < gs.size(); ++r)
    if (gs[r].stages.size() != n_st) return fail(hs[r], HB_ERR_STATE, "rank programs diverged (%zu vs %zu stages)", gs[r].stages.size(), n_st);
  for (size_t st = 0; st < n_st; ++st)
    for (size_t r = 0; r < gs.size(); ++r)
      for (auto& f : gs[r].stages[st]) if (int rc = f()) return rc;
  return HB_OK;
}

int upload_tape(hb_handle* h, const hb_tape* tp) {
  const long long n_gen = tp->n_gen, nct = tp->nct;
  const int d = tp->d, delta = tp->delta;
  auto up = [&](const void* src, size_t bytes, const void** dst) -> int {
    *dst = nullptr;
    if (!src) return HB_OK;
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes ? bytes : 1));
    h->tape_allocs.push_back(p);
    CK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
    *dst = p;
    return HB_OK;
  };
  const size_t n_sets = h->cfg.mt > 1 ? (size_t)(2 * h->cfg.mt - 1) : 1;
  const size_t n1 = (size_t)n_gen * nct;         // per generation and chain: u_accept, mt_pick
  const size_t n = n1 * n_sets;                  // per proposal set
  int rc = HB_OK;
  if ((rc = up(tp->u_snooker, n * 8, (const void**)&h->tape.u_snooker))) return rc;
  if ((rc = up(tp->D, n, (const void**)&h->tape.D))) return rc;
  if ((rc = up(tp->partners, n * 2 * delta * 4, (const void**)&h->tape.partners))) return rc;
  if ((rc = up(tp->id, n, (const void**)&h->tape.id))) return rc;
  if ((rc = up(tp->zt, n * d * 8, (const void**)&h->tape.zt))) return rc;
  if ((rc = up(tp->lambda, n * d * 8, (const void**)&h->tape.lambda))) return rc;
  if ((rc = up(tp->g_is_one, n, (const void**)&h->tape.g_is_one))) return rc;
  if ((rc = up(tp->zeta, n * d * 8, (const void**)&h->tape.zeta))) return rc;
  if ((rc = up(tp->g_snooker, n * 8, (const void**)&h->tape.g_snooker))) return rc;
  if ((rc = up(tp->u_accept, n1 * 8, (const void**)&h->tape.u_accept))) return rc;
  if (h->cfg.mt > 1 && (rc = up(tp->mt_pick, n1, (const void**)&h->tape_mt_pick))) return rc;
  h->tape.nct = nct;
  h->tape_n_gen = n_gen;
  h->tape_events = tp->n_events;
  if (tp->outlier_new && tp->n_events > 0) {
    h->tape_outlier_new.assign(tp->outlier_new, tp->outlier_new + (size_t)tp->n_events * nct);
    if ((rc = up(tp->outlier_new, (size_t)tp->n_events * nct * 4, (const void**)&h->tape_outlier_dev))) return rc;
  }
  h->have_tape = true;
  return HB_OK;
}

}  // namespace

// =========================================== C ABI ========================================================
extern "C" {

void hb_config_default(hb_config* c) {
  memset(c, 0, sizeof *c);
  c->abi_version = HB_ABI_VERSION; c->sweep = HB_SWEEP_SYNCHRONOUS; c->n_runs = 1;
  c->burnin = 0; c->adapt = 0.1; c->update_interval = 10; c->delta = 3; c->c_val = 0.1; c->c_star = 1e-12; c->nCR = 3;
  c->thin = 1; c->p_g = 0.2; c->beta0 = 1; c->outlier_check = 1; c->prior = HB_PRIOR_FLAT; c->bound = HB_BOUND_NONE;
  c->mt = 1; c->rng_mode = HB_RNG_PHILOX; c->seed = 1312;
}

void hb_wait_idle(void) { g_reaper.join_all(); }

int hb_test_force_wide(int on) { const int prev = g_hb_force_wide; g_hb_force_wide = on ? 1 : 0; return prev; }

int hb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* hb_last_error(const hb_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int hb_create(const hb_config* cfg, int device, hb_handle** out) {
  hb_handle* h = nullptr;
  if (!cfg || !out) return fail(h, HB_ERR_INVALID, "hb_create: null argument");
  *out = nullptr;
  if (cfg->abi_version != HB_ABI_VERSION) return fail(h, HB_ERR_INVALID, "hb_config.abi_version %d != %d", cfg->abi_version, HB_ABI_VERSION);
  // argument checks, R/dream.R:136-137 and R/dream_parallel.R:166-175
  if (cfg->nc < 1 || cfg->t < 2 || cfg->d < 1) return fail(h, HB_ERR_INVALID, "nc, t, d must be positive (t >= 2)");
  if (!cfg->past_sample && cfg->nc <= 2 * (long long)cfg->delta) return fail(h, HB_ERR_INVALID, "Argument 'nc' must be > 'delta' * 2!");
  if (cfg->delta < 1 || cfg->delta > HB_MAX_DELTA) return fail(h, HB_ERR_UNSUPPORTED, "delta must be in 1..%d on the device", HB_MAX_DELTA);
  if (cfg->nCR < 1 || cfg->nCR > HB_MAX_NCR) return fail(h, HB_ERR_UNSUPPORTED, "nCR must be in 1..%d on the device", HB_MAX_NCR);
  if (cfg->mt > 1) {   // MT-DREAM exists in dream_parallel only (R/dream_parallel.R:295)
    if (cfg->sweep != HB_SWEEP_SYNCHRONOUS || cfg->n_runs > 1) return fail(h, HB_ERR_UNSUPPORTED, "mt > 1 (MT-DREAM) needs the synchronous sweep of dream_parallel and a single run");
    if (cfg->mt > 64) return fail(h, HB_ERR_UNSUPPORTED, "mt must be <= 64 on the device");
    if (cfg->t >= (1 << 24)) return fail(h, HB_ERR_UNSUPPORTED, "mt > 1 needs t < 2^24 (the Philox substream lives in the top byte of the generation word)");
  }
  if ((cfg->mt > 1 || cfg->psnooker > 0) && (cfg->lik == 21 || cfg->lik == 22))
    return fail(h, HB_ERR_INVALID, "ABC method cannot be combined with MT-DREAM or snooker updating, see Sadegh and Vrugt (2014), WRR!");
  if (cfg->lik != 1 && cfg->lik != 2 && cfg->lik != 31 && cfg->lik != 21 && cfg->lik != 22 && cfg->lik != 99)
    return fail(h, HB_ERR_INVALID, "Argument 'lik' has to be one of {1,2,21,22}!");
  if ((cfg->lik == 21 || cfg->lik == 22 || cfg->lik == 99) && cfg->n_runs > 1)
    return fail(h, HB_ERR_UNSUPPORTED, "batched independent runs support lik = 1, 2 and 31");
  if (cfg->thin < 1 || cfg->update_interval < 1) return fail(h, HB_ERR_INVALID, "thin and updateInterval must be >= 1");
  if (cfg->psnooker > 0 && !cfg->past_sample) return fail(h, HB_ERR_INVALID, "psnooker > 0 can only be applied if past_sample == TRUE");
  if (cfg->d > 1024) return fail(h, HB_ERR_UNSUPPORTED, "d > 1024 is not supported on the device");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(h, HB_ERR_NO_DEVICE, "no CUDA device is available and there is no CPU fallback"); }
  if (device < 0 || device >= ndev) return fail(h, HB_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
  h = new hb_handle();
  h->cfg = *cfg;
  if (h->cfg.n_runs < 1) h->cfg.n_runs = 1;
  h->device = device;
  h->nc = cfg->nc; h->t = cfg->t; h->d = cfg->d; h->nct = h->cfg.n_runs * cfg->nc;
  h->ldx = cfg->d == 1 ? 1 : (cfg->d + 3) & ~3;
  h->m0 = cfg->m0 > 0 ? cfg->m0 : (cfg->past_sample ? 10LL * cfg->d : cfg->nc);          // :168-171
  h->archive_update = cfg->archive_update > 0 ? cfg->archive_update : (cfg->past_sample ? 10 : 0);   // :172-173
  if (h->m0 < h->nc) { std::string e = "m0 must be >= nc"; delete h; g_create_err = e; return HB_ERR_INVALID; }
  h->m_arch = h->m0;
  h->m_cap = h->m0 + (h->archive_update ? h->nc * (h->t / h->archive_update + 1) : 0);
  h->n_local = h->nc; h->chain0 = 0;
  compute_dims(h);
  if (cfg->sweep == HB_SWEEP_SEQUENTIAL || h->cfg.n_runs > 1) {
    std::string why;
    if (hb_seq_supported(&h->cfg, why) != HB_OK) { g_create_err = why; delete h; return HB_ERR_UNSUPPORTED; }
    h->modeb = true;
    h->n_local = h->nct;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  if (e != cudaSuccess) { g_create_err = std::string("CUDA initialisation failed: ") + cudaGetErrorString(e); delete h; return HB_ERR_CUDA; }
  *out = h;
  return HB_OK;
}

void hb_destroy(hb_handle* h) {
  if (!h) return;
  const double t_begin = now_s();
  cudaSetDevice(h->device);
  if (h->group_member) cudaDeviceSynchronize();   // the shared stream of a rank emulation may already be gone with rank 0's handle
  else if (h->stream) cudaStreamSynchronize(h->stream);
  hb_seq_destroy(h);
  if (h->vt && h->tctx) h->vt->destroy(h->tctx);
  hb_nccl_destroy(h->comm);
  std::vector<void*> ipc_close;       // peers' buffers mapped here, and this rank's exported buffers: released by the reaper thread
  std::vector<void*> exported;        // (unmapping / freeing 1-3 GB takes 0.1-0.2 s and sat on the caller's critical path)
  if (h->delta_mode) {
    for (int r = 0; r < 8; ++r) {
      if (h->peerX_opened[r]) ipc_close.push_back(h->peerX_opened[r]);
      if (h->arena_opened[r]) ipc_close.push_back(h->arena_opened[r]);
      if (h->peer_box_opened[r]) ipc_close.push_back(h->peer_box_opened[r]);
      if (h->boxmem[r]) exported.push_back(h->boxmem[r]);
    }
    if (h->arena) exported.push_back(h->arena);
    if (h->X) exported.push_back(h->X);
    h->X = nullptr;
    if (h->side && h->side != h->stream) cudaStreamDestroy(h->side);
    if (h->side2 && h->side2 != h->stream) cudaStreamDestroy(h->side2);
    for (int q = 0; q < HB_MAX_PIECES; ++q) if (h->ev_piece[q]) cudaEventDestroy(h->ev_piece[q]);
    if (h->ev_sent) cudaEventDestroy(h->ev_sent);
    if (h->ev_k1last) cudaEventDestroy(h->ev_k1last);
    if (h->ev_applied) cudaEventDestroy(h->ev_applied);
  }
  if (h->group_member && h->own_stream) {   // rank emulation: the shared stream belongs to rank 0's handle
    h->stream = h->own_stream; h->own_stream = nullptr;
  }
  if (h->peer_mode) {
    h->X = nullptr;
    for (int b = 0; b < 2; ++b) for (int r = 0; r < 8; ++r) if (h->peer_opened[b][r]) cudaIpcCloseMemHandle(h->peer_opened[b][r]);
    cudaFree(h->Xbuf[0]); cudaFree(h->Xbuf[1]);
  }
  void* ptrs[] = {h->X, h->Xalt, h->persist_partials, h->persist_ibuf, h->Z, h->XP, h->id, h->lp_xp, h->ll_raw, h->fx_xp, h->lp, h->ll, h->lpost, h->fx, h->lpost_hist,
                  h->hist_x, h->hist_l, h->hist_fx, h->accept_rec, h->ctrl, h->shift, h->colsum, h->qsum, h->partials,
                  h->out_ar, h->out_cr, h->proxy, h->proxy_full, h->lpost_full, h->outl_dev, h->news_dev, h->rstat_dev,
                  h->rstat_xm, h->rstat_ss, h->stage, h->log_ctr, h->ibuf, h->errbits_dev, h->grid_bar, h->pmin, h->pmax, h->prior_W, h->prior_mu, h->ZT, h->lp_zt, h->ll_zt, h->fx_zt, h->id_zt, h->mt_p1, h->mt_accept, h->rstat_sums, h->gamma_tab, h->outl_tmp, h->outl_sorted, h->outl_thr,
                  h->outl_count};
  {
    std::vector<void*> dead;
    for (void* p : ptrs) if (p) dead.push_back(p);
    for (void* p : h->tape_allocs) dead.push_back(p);
    const int dev = h->device;
    g_reaper.add(std::thread([dead, ipc_close, exported, dev]() {
      cudaSetDevice(dev);
      for (void* p : ipc_close) cudaIpcCloseMemHandle(p);   // this rank's imports first, then its own exported buffers
      for (void* p : exported) cudaFree(p);
      for (void* p : dead) cudaFree(p);
    }));
  }
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  pin_release(h);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  timing_note("destroy", now_s() - t_begin);
}

int hb_set_bounds(hb_handle* h, const double* par_min, const double* par_max, int n) {
  if (!h) return HB_ERR_INVALID;
  if (!par_min || !par_max) return fail(h, HB_ERR_INVALID, "hb_set_bounds: min and max must both be given (R/internals.R:210)");
  if (n != 1 && n != h->d) return fail(h, HB_ERR_INVALID, "hb_set_bounds: n must be 1 or d");
  for (int k = 0; k < n; ++k) {
    if (!(par_min[k] <= par_max[k])) return fail(h, HB_ERR_INVALID, "hb_set_bounds: min must be <= max (parameter %d: min = %g, max = %g)", k + 1, par_min[k], par_max[k]);
    if (par_min[k] == par_max[k] && (h->cfg.bound == HB_BOUND_REFLECT || h->cfg.bound == HB_BOUND_FOLD))
      return fail(h, HB_ERR_INVALID, "hb_set_bounds: bound = 'reflect' / 'fold' need min < max (parameter %d): R's while loop would not end", k + 1);
  }
  cudaSetDevice(h->device);
  h->pmin_h.resize((size_t)h->d); h->pmax_h.resize((size_t)h->d);
  for (int k = 0; k < h->d; ++k) { h->pmin_h[k] = par_min[n > 1 ? k : 0]; h->pmax_h[k] = par_max[n > 1 ? k : 0]; }   // :211-212
  if (!h->pmin) { CK(dalloc(&h->pmin, (size_t)h->d)); CK(dalloc(&h->pmax, (size_t)h->d)); }
  CK(cudaMemcpy(h->pmin, h->pmin_h.data(), sizeof(double) * h->d, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->pmax, h->pmax_h.data(), sizeof(double) * h->d, cudaMemcpyHostToDevice));
  h->have_min_max = true;
  return HB_OK;
}

// par.info$mu, par.info$cov for prior = "normal" (R/internals.R:58-59): chol(cov) and its inverse in long double on the host
int hb_set_prior_normal(hb_handle* h, const double* mu, const double* cov_colmajor) {
  if (!h) return HB_ERR_INVALID;
  if (!mu || !cov_colmajor) return fail(h, HB_ERR_INVALID, "hb_set_prior_normal: mu and cov must both be given");
  cudaSetDevice(h->device);
  const int d = h->d;
  std::vector<long double> C((size_t)d * d), Rm((size_t)d * d, 0.0L), Wi((size_t)d * d, 0.0L);
  for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) C[(size_t)i * d + j] = cov_colmajor[i + (size_t)d * j];
  for (int j = 0; j < d; ++j) {                                  // upper Cholesky factor: cov = R^T R
    long double s = C[(size_t)j * d + j];
    for (int k = 0; k < j; ++k) s -= Rm[(size_t)k * d + j] * Rm[(size_t)k * d + j];
    if (!(s > 0)) return fail(h, HB_ERR_INVALID, "'cov' is not positive definite (chol failed)");
    const long double rjj = sqrtl(s);
    Rm[(size_t)j * d + j] = rjj;
    for (int i = j + 1; i < d; ++i) {
      long double v = C[(size_t)j * d + i];
      for (int k = 0; k < j; ++k) v -= Rm[(size_t)k * d + j] * Rm[(size_t)k * d + i];
      Rm[(size_t)j * d + i] = v / rjj;
    }
  }
  for (int j = 0; j < d; ++j) {                                  // W = R^-1 (upper triangular), column by column
    Wi[(size_t)j * d + j] = 1.0L / Rm[(size_t)j * d + j];
    for (int i = j - 1; i >= 0; --i) {
      long double v = 0;
      for (int k = i + 1; k <= j; ++k) v += Rm[(size_t)i * d + k] * Wi[(size_t)k * d + j];
      Wi[(size_t)i * d + j] = -v / Rm[(size_t)i * d + i];
    }
  }
  long double sl = 0;
  for (int i = 0; i < d; ++i) sl += logl(Rm[(size_t)i * d + i]);
  h->prior_const = -(double)sl - 0.5 * d * log(2.0 * 3.14159265358979323846);
  std::vector<double> W((size_t)d * d);
  for (size_t e = 0; e < W.size(); ++e) W[e] = (double)Wi[e];
  if (!h->prior_W) { CK(dalloc(&h->prior_W, (size_t)d * d)); CK(dalloc(&h->prior_mu, (size_t)d)); }
  CK(cudaMemcpy(h->prior_W, W.data(), sizeof(double) * W.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->prior_mu, mu, sizeof(double) * d, cudaMemcpyHostToDevice));
  h->have_prior_normal = true;
  return HB_OK;
}

int hb_set_lik_args(hb_handle* h, const char* abc_rho, const double* abc_e, int64_t n_abc_e, const char* lik_fun) {
  if (!h) return HB_ERR_INVALID;
  if (abc_rho && strcmp(abc_rho, "abc_distfun") != 0)
    return fail(h, HB_ERR_UNKNOWN_TARGET, "object '%s' of mode 'function' was not found in the device registry of abc_rho functions (registered: abc_distfun)", abc_rho);
  if (lik_fun && strcmp(lik_fun, "fun_lik") != 0)
    return fail(h, HB_ERR_UNKNOWN_TARGET, "object '%s' of mode 'function' was not found in the device registry of lik_fun functions (registered: fun_lik)", lik_fun);
  if (abc_e && n_abc_e < 1) return fail(h, HB_ERR_INVALID, "abc_e must not be empty");
  h->abc_rho = abc_rho ? abc_rho : "";
  h->lik_fun = lik_fun ? lik_fun : "";
  h->abc_e_h.clear();
  if (abc_e) h->abc_e_h.assign(abc_e, abc_e + n_abc_e);
  return HB_OK;
}

int hb_set_obs(hb_handle* h, const double* obs, int64_t n) {
  if (!h || !obs || n < 1) return fail(h, HB_ERR_INVALID, "hb_set_obs: bad arguments");
  h->obs_h.assign(obs, obs + n);
  return HB_OK;
}

int hb_set_target(hb_handle* h, const char* name, const double* data, const int64_t* dims, int ndims) {
  if (!h || !name) return HB_ERR_INVALID;
  cudaSetDevice(h->device);
  const hb_target_vtable* vt = hb_find_target(name);
  if (!vt) return fail(h, HB_ERR_UNKNOWN_TARGET, "object '%s' of mode 'function' was not found in the device plugin registry (arbitrary R closures stay off the GPU path)", name);
  if (h->vt && h->tctx) { h->vt->destroy(h->tctx); h->tctx = nullptr; }
  char err[256] = {0};
  void* ctx = nullptr;
  const int rc = vt->init(&ctx, h->d, h->cfg.lik, data, dims, ndims, h->obs_h.empty() ? nullptr : h->obs_h.data(),
                          (int64_t)h->obs_h.size(), h->cfg.glue_shape, err, (int)sizeof err);
  if (rc != HB_OK) return fail(h, rc, "%s", err[0] ? err : "target init failed");
  h->vt = vt; h->tctx = ctx; h->target_name = name; h->target_is_hydro = (h->target_name == "HydroModel" || h->target_name == "fun_wrap_glue");
  const int lik = h->cfg.lik;
  if (lik == 21 || lik == 22 || lik == 99) {   // the closures calc_ll would get() by name (R/internals.R:10,12,16)
    if (!h->target_is_hydro) return fail(h, HB_ERR_UNSUPPORTED, "lik = %d is a device epilogue of HydroModel only", lik);
    if (lik == 99) {
      if (h->lik_fun.empty()) return fail(h, HB_ERR_INVALID, "lik = 99 needs lik_fun (hb_set_lik_args)");
    } else {
      if (h->abc_rho.empty() || h->abc_e_h.empty()) return fail(h, HB_ERR_INVALID, "lik = %d needs abc_rho and abc_e (hb_set_lik_args)", lik);
      if (int rc2 = hb_hydro_set_abc(ctx, h->abc_e_h.data(), (int64_t)h->abc_e_h.size(), err, (int)sizeof err)) return fail(h, rc2, "%s", err);
    }
  }
  return HB_OK;
}

int hb_set_target_batched(hb_handle* h, const char* name, const double* data, int64_t len, const double* obs,
                          int64_t n_obs) {
  if (!h || !name) return HB_ERR_INVALID;
  cudaSetDevice(h->device);
  if (std::string(name) != "HydroModel") return fail(h, HB_ERR_UNSUPPORTED, "only HydroModel takes per-run data; use hb_set_target for '%s'", name);
  const hb_target_vtable* vt = hb_find_target(name);
  if (h->vt && h->tctx) { h->vt->destroy(h->tctx); h->tctx = nullptr; }
  char err[256] = {0};
  void* ctx = nullptr;
  const int rc = hb_hydro_init_batched(&ctx, h->d, h->cfg.lik, data, len, obs, n_obs, h->cfg.n_runs, h->cfg.glue_shape, err, (int)sizeof err);
  if (rc != HB_OK) return fail(h, rc, "%s", err[0] ? err : "target init failed");
  h->vt = vt; h->tctx = ctx; h->target_name = name; h->target_is_hydro = true;
  return HB_OK;
}

int hb_comm_unique_id(void* id128) {
  std::string err;
  const int rc = hb_nccl_unique_id(id128, err);
  if (rc) g_create_err = err;
  return rc;
}

int hb_comm_init(hb_handle* h, int rank, int world, const void* id128) {
  if (!h) return HB_ERR_INVALID;
  if (h->have_initial) return fail(h, HB_ERR_STATE, "hb_comm_init must precede hb_set_initial");
  if (world < 1 || rank < 0 || rank >= world) return fail(h, HB_ERR_INVALID, "bad rank/world");
  if (h->nc % world) return fail(h, HB_ERR_INVALID, "nc must be divisible by the number of GPUs");
  if (h->cfg.sweep != HB_SWEEP_SYNCHRONOUS || h->cfg.n_runs > 1) return fail(h, HB_ERR_UNSUPPORTED, "only the synchronous single-population sweep shards across GPUs");
  if (h->cfg.mt > 1 && world > 1 && h->cfg.exchange != HB_EXCHANGE_ALLGATHER)
    return fail(h, HB_ERR_UNSUPPORTED, "mt > 1 (MT-DREAM) on several GPUs runs over HB_EXCHANGE_ALLGATHER (the candidate population zt is all-gathered too)");
  cudaSetDevice(h->device);
  const double t_begin = now_s();
  // HB_EXCHANGE_DELTA needs no NCCL communicator: its barriers and small collectives run over peer-mapped memory
  if (world > 1 && h->cfg.exchange != HB_EXCHANGE_DELTA) {
    if (!id128) return fail(h, HB_ERR_INVALID, "hb_comm_init: this exchange mode needs the communicator id of hb_comm_unique_id");
    if (int rc = hb_nccl_init(&h->comm, rank, world, id128, h->err)) return rc;
    timing_note("nccl_init", now_s() - t_begin);
  }
  h->rank = rank; h->world = world;
  h->n_local = h->nc / world; h->chain0 = rank * h->n_local;
  if (world > 1 && h->cfg.exchange == HB_EXCHANGE_DELTA) {
    if (world > 8) return fail(h, HB_ERR_UNSUPPORTED, "delta exchange supports up to 8 GPUs of one node");
    h->delta_mode = true;
  }
  if (world > 1 && h->cfg.exchange == HB_EXCHANGE_PEER) {
    if (world > 8) return fail(h, HB_ERR_UNSUPPORTED, "peer exchange supports up to 8 GPUs of one node");
    if (h->cfg.past_sample) return fail(h, HB_ERR_UNSUPPORTED, "peer exchange does not cover the DREAM-ZS archive; use HB_EXCHANGE_ALLGATHER");
    if (h->d <= 16) return fail(h, HB_ERR_UNSUPPORTED, "peer exchange needs d > 16; use HB_EXCHANGE_ALLGATHER");
    h->peer_mode = true;
  }
  return HB_OK;
}

// HB_EXCHANGE_DELTA with a sharded upload: complete the peers' replicas with this rank's rows (NVLink peer copies), then
// arrive at flag 1 with value 1; the first generation waits for every peer's arrival
static int complete_replicas(hb_handle* h) {
  if (!h->initial_sharded) return HB_OK;
  const size_t off = (size_t)h->chain0 * h->ldx, bytes = (size_t)h->n_local * h->ldx * sizeof(double);
  for (int q = 1; q < h->world; ++q) {
    const int r = (h->rank + q) % h->world;
    CK(cudaMemcpyAsync(h->peerX[r] + off, h->X + off, bytes, cudaMemcpyDefault, h->stream));
  }
  hb_arena_peers ap; for (int r = 0; r < 8; ++r) ap.p[r] = h->arena_peer[r];
  CK(hb_launch_flag_signal(ap, 1, h->rank, h->world, 1u, h->stream));
  return HB_OK;
}

int hb_peer_export(hb_handle* h, void* out128) {
  if (!h || !out128) return HB_ERR_INVALID;
  if (!h->peer_mode && !h->delta_mode) return fail(h, HB_ERR_STATE, "hb_peer_export: the handle is not in HB_EXCHANGE_PEER / HB_EXCHANGE_DELTA mode");
  if (!h->have_initial) return fail(h, HB_ERR_STATE, "hb_peer_export must follow hb_set_initial");
  cudaSetDevice(h->device);
  cudaIpcMemHandle_t hd[HB_PEER_EXPORT_HANDLES];
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memset(hd, 0, sizeof hd);
  if (h->delta_mode) {   // X | arena | box of source rank 0 .. 7
    CK(cudaIpcGetMemHandle(&hd[0], h->X));
    CK(cudaIpcGetMemHandle(&hd[1], h->arena));
    for (int r = 0; r < h->world; ++r) if (r != h->rank) CK(cudaIpcGetMemHandle(&hd[2 + r], h->boxmem[r]));
    memcpy(out128, hd, sizeof hd);
    return HB_OK;
  }
  CK(cudaIpcGetMemHandle(&hd[0], h->Xbuf[0]));
  CK(cudaIpcGetMemHandle(&hd[1], h->Xbuf[1]));
  memcpy(out128, hd, sizeof hd);
  return HB_OK;
}

int hb_peer_import(hb_handle* h, const void* all_handles) {
  if (!h || !all_handles) return HB_ERR_INVALID;
  if (!h->peer_mode && !h->delta_mode) return fail(h, HB_ERR_STATE, "hb_peer_import: the handle is not in HB_EXCHANGE_PEER / HB_EXCHANGE_DELTA mode");
  cudaSetDevice(h->device);
  const double t_begin = now_s();
  struct note_at_exit { double t0; ~note_at_exit() { timing_note("peer_import", now_s() - t0); } } note_guard{t_begin};
  const cudaIpcMemHandle_t* hd = (const cudaIpcMemHandle_t*)all_handles;
  if (h->delta_mode) {
    for (int r = 0; r < h->world; ++r) {
      if (r == h->rank) continue;
      void* px = nullptr; void* pf = nullptr; void* pb = nullptr;
      const cudaIpcMemHandle_t* hr = hd + (size_t)r * HB_PEER_EXPORT_HANDLES;
      CK(cudaIpcOpenMemHandle(&px, hr[0], cudaIpcMemLazyEnablePeerAccess));
      CK(cudaIpcOpenMemHandle(&pf, hr[1], cudaIpcMemLazyEnablePeerAccess));
      CK(cudaIpcOpenMemHandle(&pb, hr[2 + h->rank], cudaIpcMemLazyEnablePeerAccess));   // only the box this rank writes
      h->peerX_opened[r] = px; h->peerX[r] = (double*)px;
      h->arena_opened[r] = pf; h->arena_peer[r] = (unsigned char*)pf;
      h->peer_box_opened[r] = pb; h->peer_box[r] = (unsigned char*)pb;
    }
    h->arena_peer[h->rank] = h->arena;
    if (int rc = complete_replicas(h)) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->peers_ready = true;
    return HB_OK;
  }
  for (int r = 0; r < h->world; ++r)
    for (int b = 0; b < 2; ++b) {
      if (r == h->rank) { h->peer_ptr[b][r] = h->Xbuf[b]; continue; }
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, hd[(size_t)r * HB_PEER_EXPORT_HANDLES + b], cudaIpcMemLazyEnablePeerAccess));
      h->peer_opened[b][r] = p;
      h->peer_ptr[b][r] = (const double*)p;
    }
  h->peers_ready = true;
  return HB_OK;
}

// pageable host -> device through the pinned ring: a few threads fill pinned[k & 1] while the DMA of chunk k-1 runs
static int upload_pipelined(hb_handle* h, void* dst_dev, const void* src_host, size_t bytes) {
  const size_t want = std::min<size_t>((size_t)64 << 20, std::max<size_t>(bytes, 4096));
  if (bytes < ((size_t)16 << 20)) { CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, h->stream)); return HB_OK; }
  if (int rc = pin_acquire(h, 2 * ((size_t)64 << 20) > 2 * want ? 2 * ((size_t)64 << 20) : 2 * want)) return rc;
  const size_t half = h->pin_bytes / 2;
  const size_t n_chunks = (bytes + half - 1) / half;
  unsigned hw = std::thread::hardware_concurrency();
  const int n_thr = (int)std::min<unsigned>(8u, hw > 2 ? hw - 1 : 1);
  std::vector<cudaEvent_t> ev(n_chunks, nullptr);
  cudaError_t e = cudaSuccess;
  for (size_t k = 0; k < n_chunks && e == cudaSuccess; ++k) {
    if (k >= 2) e = cudaEventSynchronize(ev[k - 2]);               // pinned[k & 1] has been consumed by the DMA
    if (e != cudaSuccess) break;
    const size_t off = k * half, n = std::min(half, bytes - off);
    char* pb = reinterpret_cast<char*>(h->pin) + (k & 1) * half;
    const char* sb = reinterpret_cast<const char*>(src_host) + off;
    std::vector<std::thread> th;
    for (int t = 1; t < n_thr; ++t) th.emplace_back([=]() { const size_t lo = n / n_thr * t, hi = (t == n_thr - 1) ? n : n / n_thr * (t + 1); memcpy(pb + lo, sb + lo, hi - lo); });
    memcpy(pb, sb, n_thr > 1 ? n / n_thr : n);
    for (auto& x : th) x.join();
    e = cudaMemcpyAsync(reinterpret_cast<char*>(dst_dev) + off, pb, n, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev[k], h->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  for (auto& x : ev) if (x) cudaEventDestroy(x);
  if (e != cudaSuccess) return fail(h, HB_ERR_CUDA, "initial-state upload failed: %s", cudaGetErrorString(e));
  return HB_OK;
}

static int set_initial_impl(hb_handle* h, const double* x0, bool row_major) {
  if (!h || !x0) return HB_ERR_INVALID;
  if (h->have_initial) return fail(h, HB_ERR_STATE, "hb_set_initial was already called");
  cudaSetDevice(h->device);
  const double t_begin = now_s();
  if (int rc = allocate(h)) return rc;
  timing_note("allocate", now_s() - t_begin);
  const double t_up = now_s();
  const long long m0 = h->modeb ? h->nct : h->m0, nc = h->modeb ? h->nct : h->nc;
  const int d = h->d, ldx = h->ldx;
  // HB_EXCHANGE_DELTA, numpy layout, no archive: every rank uploads only the rows of ITS chains over PCIe; the other
  // (G-1)/G of the replica arrives over NVLink in hb_peer_import (each rank copies its shard into every peer's X)
  const bool shard_upload = h->delta_mode && h->world > 1 && row_major && !h->Z && m0 == nc;
  h->initial_sharded = shard_upload;
  const long long up_row0 = shard_upload ? h->chain0 : 0, up_rows = shard_upload ? h->n_local : m0;
  double* tmp = nullptr;
  CK(dalloc(&tmp, (size_t)up_rows * d));
  if (int rc = upload_pipelined(h, tmp, x0 + (size_t)up_row0 * d, sizeof(double) * (size_t)up_rows * d)) { cudaFree(tmp); return rc; }
  // xt <- z[(m0-nc+1):m0, ]  (R/dream_parallel.R:213)
  if (row_major) {   // rows are already contiguous: strided device copies add the ldx padding
    const size_t sp = sizeof(double) * d, dp = sizeof(double) * ldx;
    if (h->peer_mode) CK(cudaMemcpy2DAsync(h->Xbuf[0], dp, tmp + (m0 - nc + h->chain0) * d, sp, sp, (size_t)h->n_local, cudaMemcpyDeviceToDevice, h->stream));
    else if (shard_upload) CK(cudaMemcpy2DAsync(h->X + h->chain0 * ldx, dp, tmp, sp, sp, (size_t)h->n_local, cudaMemcpyDeviceToDevice, h->stream));
    else CK(cudaMemcpy2DAsync(h->X, dp, tmp + (m0 - nc) * d, sp, sp, (size_t)nc, cudaMemcpyDeviceToDevice, h->stream));
    if (h->Z) CK(cudaMemcpy2DAsync(h->Z, dp, tmp, sp, sp, (size_t)m0, cudaMemcpyDeviceToDevice, h->stream));
  } else {
    if (h->peer_mode) CK(hb_launch_colmajor_to_rows(tmp, m0, m0 - nc + h->chain0, h->n_local, d, ldx, h->Xbuf[0], h->stream));
    else CK(hb_launch_colmajor_to_rows(tmp, m0, m0 - nc, nc, d, ldx, h->X, h->stream));
    if (h->Z) CK(hb_launch_colmajor_to_rows(tmp, m0, 0, m0, d, ldx, h->Z, h->stream));
  }
  {  // shift of the one-pass column variance: chain 0 of the run (identical on every rank)
    std::vector<double> sh((size_t)d);
    for (int k = 0; k < d; ++k) sh[(size_t)k] = row_major ? x0[(m0 - nc) * (long long)d + k] : x0[(m0 - nc) + m0 * (long long)k];
    CK(cudaMemcpyAsync(h->shift, sh.data(), sizeof(double) * d, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(tmp);
  h->have_initial = true;
  timing_note("set_initial", now_s() - t_up, sizeof(double) * (double)up_rows * d);
  return HB_OK;
}

int hb_set_initial(hb_handle* h, const double* x0) { return set_initial_impl(h, x0, false); }
int hb_set_initial_rows(hb_handle* h, const double* x0_rowmajor) { return set_initial_impl(h, x0_rowmajor, true); }

int hb_set_tape(hb_handle* h, const hb_tape* tape) {
  if (!h || !tape) return HB_ERR_INVALID;
  cudaSetDevice(h->device);
  if (tape->nct != h->nct || tape->d != h->d || tape->delta != h->cfg.delta || tape->n_gen < h->t - 1)
    return fail(h, HB_ERR_INVALID, "tape geometry does not match the configuration");
  if (!tape->D || !tape->partners || !tape->id || !tape->zt || !tape->lambda || !tape->g_is_one || !tape->zeta || !tape->u_accept)
    return fail(h, HB_ERR_INVALID, "tape is missing a required array");
  if (h->cfg.mt > 1 && (tape->n_sets != 2 * h->cfg.mt - 1 || !tape->mt_pick))
    return fail(h, HB_ERR_INVALID, "mt > 1 needs a tape with n_sets = 2*mt - 1 proposal sets per generation and mt_pick");
  return upload_tape(h, tape);
}

// One slice of generations for a set of rank handles driven in lockstep: a single handle (hb_run: one process per GPU) or
// the ranks of an emulation on one device (hb_group_run).
static int run_slice(std::vector<hb_handle*>& hs, int64_t n_generations) {
  for (hb_handle* h : hs) {
    cudaSetDevice(h->device);
    if (!h->vt) return fail(h, HB_ERR_STATE, "hb_run: no target set");
    if (!h->have_initial) return fail(h, HB_ERR_STATE, "hb_run: hb_set_initial has not been called");
    if (h->cfg.rng_mode == HB_RNG_TAPE && !h->have_tape) return fail(h, HB_ERR_STATE, "hb_run: rng_mode is TAPE but no tape was set");
    if (h->cfg.prior == HB_PRIOR_UNIFORM && !h->have_min_max) return fail(h, HB_ERR_INVALID, "prior = 'uniform' needs min and max");
    if (h->cfg.prior == HB_PRIOR_NORMAL && !h->have_prior_normal) return fail(h, HB_ERR_INVALID, "prior = 'normal' needs mu and cov (hb_set_prior_normal)");
    if ((h->peer_mode || h->delta_mode) && !h->peers_ready) return fail(h, HB_ERR_STATE, "hb_run: HB_EXCHANGE_PEER / HB_EXCHANGE_DELTA need hb_peer_export / hb_peer_import after hb_set_initial");
    if (h->group_member && hs.size() == 1 && h->world > 1) return fail(h, HB_ERR_STATE, "hb_run: the handle belongs to a rank emulation, use hb_group_run");
    if (h->gen != hs[0]->gen || h->t != hs[0]->t) return fail(h, HB_ERR_STATE, "hb_group_run: the handles are not at the same generation");
  }
  for (hb_handle* h : hs) {
    if (h->initialised) continue;
    if (h->initial_sharded)   // every peer has copied its shard of the initial population into this replica
      CK(hb_launch_flag_wait(h->arena, 1, h->rank, h->world, 1u, &h->ctrl->err, h->stream));
    if (int rc = initialise(h)) return rc;
    if (h->modeb) { if (int rc2 = hb_seq_initialise(h)) return rc2; }
  }
  {
    std::vector<gen_prog> gs(hs.size());
    for (size_t r = 0; r < hs.size(); ++r) prog_slice_barrier(hs[r], gs[r]);
    if (int rc = run_programs(hs, gs)) return rc;
  }
  long long last = hs[0]->gen + n_generations;
  if (last > hs[0]->t) last = hs[0]->t;
  for (hb_handle* h : hs) CK(cudaEventRecord(h->ev0, h->stream));
  const double t_enq0 = now_s();
  for (long long i = hs[0]->gen + 1; i <= last; ++i) {
    if (hs.size() == 1) {   // small population behind the adaptation period: the rest of the slice in one persistent launch
      hb_handle* h = hs[0];
      const bool in_adapt = ((double)i <= h->cfg.adapt * (double)h->t);
      if (!in_adapt && !h->modeb && h->cfg.mt <= 1 && fused_available(h)) {
        if (!h->delta_mode) { if (int rc = persist_fold(h, false)) return rc; }
        bool taken = false;
        if (int rc = run_persistent(h, i, last + 1, &taken)) return rc;
        if (taken) break;
      }
      if (in_adapt && !h->modeb && h->cfg.mt <= 1 && h->persist_adapt_ok >= 0 && fused_available(h)) {
        // adaptation period: persistent launches that end at the next updateInterval boundary (pCR changes there and the
        // outlier check of R/dream_parallel.R:414-419 runs between the launches)
        if (h->persist_adapt_ok == 0) { const char* e = getenv("HB_PERSIST_ADAPT"); h->persist_adapt_ok = (e && e[0] == '0') ? -1 : 1; }
        const long long ui = h->cfg.update_interval;
        long long i1 = (i / ui + 1) * ui + 1;                      // one past the next multiple of updateInterval (>= i)
        if (i % ui == 0) i1 = i + 1;
        while ((double)(i1 - 1) > h->cfg.adapt * (double)h->t) --i1;  // generations of the slice all lie inside the period
        if (i1 > last + 1) i1 = last + 1;
        bool taken = false;
        if (h->persist_adapt_ok > 0 && i1 > i) { if (int rc = run_persistent(h, i, i1, &taken, true)) return rc; }
        if (taken) {
          const long long ie = i1 - 1;
          if (ie % ui == 0) { if (int rc = persist_fold(h, true)) return rc; }
          if (ie % ui == 0 && !h->cfg.past_sample && h->cfg.outlier_check) {
            std::vector<gen_prog> gs(1);
            prog_outlier_check(h, gs[0], ie);
            if (int rc = run_programs(hs, gs)) return rc;
          }
          i = ie;
          continue;
        }
      }
      if (!h->delta_mode) { if (int rc = persist_fold(h, false)) return rc; }   // parked generations precede this one
    }
    std::vector<gen_prog> gs(hs.size());
    for (size_t r = 0; r < hs.size(); ++r) if (int rc = build_generation(hs[r], i, gs[r])) return rc;
    if (int rc = run_programs(hs, gs)) return rc;
  }
  {
    std::vector<gen_prog> gs(hs.size());
    for (size_t r = 0; r < hs.size(); ++r) prog_err_reduce(hs[r], gs[r]);
    if (int rc = run_programs(hs, gs)) return rc;
  }
  timing_note("enqueue (host)", now_s() - t_enq0);
  int rc_all = HB_OK;
  for (hb_handle* h : hs) {
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->loop_ms += ms;
    timing_note("run (slice)", ms * 1e-3);
    const int rc = check_flags(h);
    if (rc && !rc_all) rc_all = rc;
  }
  return rc_all;
}

int hb_run(hb_handle* h, int64_t n_generations, int64_t* done_total) {
  if (!h) return HB_ERR_INVALID;
  std::vector<hb_handle*> hs{h};
  const int rc = run_slice(hs, n_generations);
  if (done_total) *done_total = h->gen;
  return rc;
}

// ---- rank emulation on ONE device (testing hook; the driver's one-GPU test box exercises the sharded path with it) ----
// n handles created with hb_comm_init(rank r, world n), HB_EXCHANGE_DELTA, after hb_set_initial: wires the peers' buffers
// in-process (plain device pointers instead of CUDA IPC mappings) and moves all handles onto the first handle's stream.
int hb_group_attach(hb_handle** hs, int n) {
  if (!hs || n < 1 || n > 8) return HB_ERR_INVALID;
  for (int r = 0; r < n; ++r) {
    hb_handle* h = hs[r];
    if (!h) return HB_ERR_INVALID;
    if (!h->delta_mode || h->world != n || h->rank != r) return fail(h, HB_ERR_STATE, "hb_group_attach: handle %d must be rank %d of %d in HB_EXCHANGE_DELTA mode", r, r, n);
    if (!h->have_initial) return fail(h, HB_ERR_STATE, "hb_group_attach must follow hb_set_initial");
    if (h->device != hs[0]->device) return fail(h, HB_ERR_INVALID, "hb_group_attach: all handles must live on one device");
    if (h->peers_ready) return fail(h, HB_ERR_STATE, "hb_group_attach: the handle is already attached");
  }
  cudaSetDevice(hs[0]->device);
  for (int r = 0; r < n; ++r) {   // the uploads of hb_set_initial ran on the handles' own streams
    hb_handle* h = hs[r];
    CK(cudaStreamSynchronize(h->stream));
  }
  for (int r = 0; r < n; ++r) {
    hb_handle* h = hs[r];
    h->group_member = true;
    if (r > 0) { h->own_stream = h->stream; h->stream = hs[0]->stream; }
    if (h->side) cudaStreamDestroy(h->side);
    if (h->side2) cudaStreamDestroy(h->side2);
    h->side = h->stream; h->side2 = h->stream;             // one stream: the emulation is sequential by construction
    for (int q = 0; q < n; ++q) { h->arena_peer[q] = hs[q]->arena; h->peerX[q] = (q == r) ? nullptr : hs[q]->X; h->peer_box[q] = hs[q]->boxmem[r]; }
  }
  for (int r = 0; r < n; ++r) if (int rc = complete_replicas(hs[r])) return rc;
  for (int r = 0; r < n; ++r) hs[r]->peers_ready = true;
  return HB_OK;
}

int hb_group_run(hb_handle** hs, int n, int64_t n_generations, int64_t* done_total) {
  if (!hs || n < 1 || n > 8) return HB_ERR_INVALID;
  std::vector<hb_handle*> v(hs, hs + n);
  for (hb_handle* h : v) if (!h || !h->group_member) return h ? fail(h, HB_ERR_STATE, "hb_group_run: call hb_group_attach first") : HB_ERR_INVALID;
  const int rc = run_slice(v, n_generations);
  if (done_total) *done_total = v[0]->gen;
  return rc;
}

int hb_sync(hb_handle* h) {
  if (!h) return HB_ERR_INVALID;
  cudaSetDevice(h->device);
  CK(cudaStreamSynchronize(h->stream));
  return HB_OK;
}

int hb_out_dims(const hb_handle* h, int64_t* out_t, int64_t* ar_rows, int64_t* ar_cols, int64_t* cr_rows, int64_t* cr_cols,
                int64_t* fx_m, int64_t* rstat_rows) {
  if (!h) return HB_ERR_INVALID;
  if (out_t) *out_t = h->out_t;
  if (ar_rows) *ar_rows = h->ar_rows;
  if (ar_cols) *ar_cols = h->ar_cols;
  if (cr_rows) *cr_rows = h->cr_rows;
  if (cr_cols) *cr_cols = h->cr_cols;
  if (fx_m) *fx_m = h->cfg.store_fx ? 1 : 0;
  if (rstat_rows) *rstat_rows = h->rstat_rows;
  return HB_OK;
}

int64_t hb_ind_out(const hb_handle* h) { return h ? h->ind_out : 0; }

// chain[nrows, d+3, n_local] (R layout, generation fastest) -> caller's host array, as a three-stage pipeline:
//   device   tiled transposition of a chunk of chains from the [generation][chain][column] history into the R layout
//   DMA      chunk -> pinned host ring (async copy on the handle's stream, full PCIe rate)
//   host     a few threads copy the chunk from the pinned ring into the caller's (pageable, usually never touched:
//            R allocVector / numpy empty) array -- the page faults of a fresh 10 GB result are spread over the threads
//            and overlap the DMA of the next chunk instead of serialising in front of one copying thread.
int hb_fetch_chain_rows(hb_handle* h, int64_t row0, int64_t nrows, double* out) {
  if (!h || !out) return HB_ERR_INVALID;
  if (!h->initialised) return fail(h, HB_ERR_STATE, "nothing has been run yet");
  if (row0 < 0 || nrows < 1 || row0 + nrows > h->out_t) return fail(h, HB_ERR_INVALID, "row range out of bounds");
  cudaSetDevice(h->device);
  const double t_begin = now_s();
  const long long nl = h->n_local;
  const int dd = h->d + 3;
  const size_t per_chain = (size_t)nrows * dd * sizeof(double);
  const size_t total = per_chain * (size_t)nl;
  size_t want = std::min<size_t>((size_t)64 << 20, total);
  if (want < per_chain) want = per_chain;
  if (h->stage_bytes < 2 * want) { if (h->stage) cudaFree(h->stage); h->stage = nullptr; CK(cudaMalloc((void**)&h->stage, 2 * want)); h->stage_bytes = 2 * want; }
  {  // a page-locked destination (hb_host_register, cudaHostAlloc): the transposed chunks go straight into it by DMA at
     // PCIe rate, no pinned ring and no copier threads (which the ranks of a node would have to share)
    cudaPointerAttributes at{};
    const bool pinned_dst = cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned_dst) {
      const size_t half = h->stage_bytes / 2;
      const long long chunk = (long long)(half / per_chain);
      const long long n_chunks = (nl + chunk - 1) / chunk;
      cudaEvent_t ev[2] = {nullptr, nullptr};
      cudaError_t e = cudaSuccess;
      for (long long k = 0; k < n_chunks && e == cudaSuccess; ++k) {
        const long long c0 = k * chunk, ncs = std::min(chunk, nl - c0);
        char* dst = reinterpret_cast<char*>(h->stage) + (size_t)(k & 1) * half;
        if (ev[k & 1]) e = cudaStreamWaitEvent(h->stream, ev[k & 1], 0);   // stream order already covers it; kept for clarity
        if (e == cudaSuccess) e = hb_launch_chain_transpose(h->hist_x, h->hist_l, h->out_t, row0, nrows, h->ind_out, nl, h->d, c0, ncs,
                                                            reinterpret_cast<double*>(dst), h->stream);
        h->launches++;
        if (e == cudaSuccess) e = cudaMemcpyAsync(reinterpret_cast<char*>(out) + (size_t)c0 * per_chain, dst, per_chain * (size_t)ncs, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && !ev[k & 1]) e = cudaEventCreateWithFlags(&ev[k & 1], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(ev[k & 1], h->stream);
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      for (auto& x : ev) if (x) cudaEventDestroy(x);
      if (e != cudaSuccess) return fail(h, HB_ERR_CUDA, "chain fetch (direct) failed: %s", cudaGetErrorString(e));
      timing_note("fetch (direct)", now_s() - t_begin, (double)total);
      return HB_OK;
    }
  }
  if (int rc = pin_acquire(h, 2 * ((size_t)64 << 20) > 2 * want ? 2 * ((size_t)64 << 20) : 2 * want)) return rc;
  const size_t half = std::min(h->stage_bytes, h->pin_bytes) / 2;
  const long long chunk = (long long)(half / per_chain);
  const long long n_chunks = (nl + chunk - 1) / chunk;
  unsigned hw = std::thread::hardware_concurrency();
  // the ranks of one node share the host cores: each takes its share of them
  const unsigned share = (hw > 2 ? hw - 1 : 1) / (unsigned)std::max(1, h->world);
  int n_thr = total < ((size_t)32 << 20) ? 1 : (int)std::min<unsigned>(16u, std::max(2u, share));
  // ready[k] = the DMA of chunk k has landed in pinned buffer k & 1 ; done[k] = copier threads finished with chunk k
  std::vector<std::atomic<int>> ready((size_t)n_chunks), done((size_t)n_chunks);
  for (auto& a : ready) a.store(0);
  for (auto& a : done) a.store(0);
  std::atomic<int> abort_flag{0};
  char* pin = reinterpret_cast<char*>(h->pin);
  char* outb = reinterpret_cast<char*>(out);
  std::vector<std::thread> copiers;
  for (int t = 0; t < n_thr; ++t)
    copiers.emplace_back([&, t]() {
      for (long long k = 0; k < n_chunks; ++k) {
        while (!ready[(size_t)k].load(std::memory_order_acquire)) { if (abort_flag.load()) return; std::this_thread::yield(); }
        const long long c0 = k * chunk, ncs = std::min(chunk, nl - c0);
        const size_t bytes = per_chain * (size_t)ncs;
        const size_t lo = bytes / n_thr * t & ~(size_t)63, hi = (t == n_thr - 1) ? bytes : (bytes / n_thr * (t + 1) & ~(size_t)63);
        if (hi > lo) stream_copy(outb + (size_t)c0 * per_chain + lo, pin + (size_t)(k & 1) * half + lo, hi - lo);
        done[(size_t)k].fetch_add(1, std::memory_order_release);
      }
    });
  std::vector<cudaEvent_t> ev((size_t)n_chunks, nullptr);
  cudaError_t e = cudaSuccess;
  auto enqueue = [&](long long k) -> cudaError_t {   // transpose chunk k on the device, then DMA it into pinned[k & 1]
    const long long c0 = k * chunk, ncs = std::min(chunk, nl - c0);
    char* dst = reinterpret_cast<char*>(h->stage) + (size_t)(k & 1) * half;
    cudaError_t r = hb_launch_chain_transpose(h->hist_x, h->hist_l, h->out_t, row0, nrows, h->ind_out, nl, h->d, c0, ncs,
                                              reinterpret_cast<double*>(dst), h->stream);
    h->launches++;
    if (r == cudaSuccess) r = cudaMemcpyAsync(pin + (size_t)(k & 1) * half, dst, per_chain * (size_t)ncs, cudaMemcpyDeviceToHost, h->stream);
    if (r == cudaSuccess) r = cudaEventCreateWithFlags(&ev[(size_t)k], cudaEventDisableTiming);
    if (r == cudaSuccess) r = cudaEventRecord(ev[(size_t)k], h->stream);
    return r;
  };
  for (long long k = 0; k < n_chunks && e == cudaSuccess; ++k) {
    if (k >= 2) while (done[(size_t)(k - 2)].load(std::memory_order_acquire) < n_thr) std::this_thread::yield();   // pinned[k & 1] is free
    e = enqueue(k);
    if (k >= 1 && e == cudaSuccess) { e = cudaEventSynchronize(ev[(size_t)(k - 1)]); ready[(size_t)(k - 1)].store(1, std::memory_order_release); }
  }
  if (e == cudaSuccess && n_chunks >= 1) { e = cudaEventSynchronize(ev[(size_t)(n_chunks - 1)]); ready[(size_t)(n_chunks - 1)].store(1, std::memory_order_release); }
  if (e != cudaSuccess) abort_flag.store(1);
  for (auto& t : copiers) t.join();
  for (auto& x : ev) if (x) cudaEventDestroy(x);
  if (e != cudaSuccess) return fail(h, HB_ERR_CUDA, "chain fetch failed: %s", cudaGetErrorString(e));
  timing_note("fetch_chain", now_s() - t_begin, (double)total);
  return HB_OK;
}

int hb_host_prefault(void* p, size_t bytes, int n_threads) {
  if (!p) return HB_ERR_INVALID;
  unsigned hw = std::thread::hardware_concurrency();
  const int nt = n_threads > 0 ? n_threads : (int)std::max(1u, std::min(8u, hw / 2));
  char* base = reinterpret_cast<char*>(p);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([=]() {
      const size_t lo = bytes / nt * t, hi = (t == nt - 1) ? bytes : bytes / nt * (t + 1);
      for (size_t o = lo; o < hi; o += 4096) reinterpret_cast<volatile char*>(base)[o] = 0;
    });
  for (auto& x : th) x.join();
  return HB_OK;
}

// Page-lock a host result array (after hb_host_prefault, from the same helper thread, while hb_run is busy) so that
// hb_fetch_chain moves it by direct DMA.  Returns HB_ERR_CUDA when the pages cannot be locked (RLIMIT_MEMLOCK): the
// caller simply keeps the pageable array and the fetch takes the pinned-ring path.
int hb_host_register(void* p, size_t bytes) {
  if (!p || !bytes) return HB_ERR_INVALID;
  const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) { cudaGetLastError(); return HB_ERR_CUDA; }
  return HB_OK;
}
int hb_host_unregister(void* p) {
  if (!p) return HB_ERR_INVALID;
  const cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) { cudaGetLastError(); return HB_ERR_CUDA; }
  return HB_OK;
}

int hb_fetch_chain(hb_handle* h, double* chain) { return h ? hb_fetch_chain_rows(h, 0, h->out_t, chain) : HB_ERR_INVALID; }

int hb_fetch_fx(hb_handle* h, double* fx) {
  if (!h || !fx) return HB_ERR_INVALID;
  if (!h->cfg.store_fx || !h->hist_fx) return fail(h, HB_ERR_STATE, "store_fx was not enabled");
  cudaSetDevice(h->device);
  const long long nl = h->n_local;
  std::vector<double> tmp((size_t)h->out_t * nl);
  CK(cudaMemcpy(tmp.data(), h->hist_fx, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (long long it = 0; it < h->out_t; ++it)
    for (long long c = 0; c < nl; ++c) fx[it + h->out_t * c] = it < h->ind_out ? tmp[(size_t)it * nl + c] : NAN;
  return HB_OK;
}

int hb_fetch_ar(hb_handle* h, double* ar) {
  if (!h || !ar) return HB_ERR_INVALID;
  cudaSetDevice(h->device);
  CK(cudaMemcpy(ar, h->out_ar, sizeof(double) * (size_t)h->cfg.n_runs * h->ar_rows * h->ar_cols, cudaMemcpyDeviceToHost));
  return HB_OK;
}
int hb_fetch_cr(hb_handle* h, double* cr) {
  if (!h || !cr) return HB_ERR_INVALID;
  cudaSetDevice(h->device);
  CK(cudaMemcpy(cr, h->out_cr, sizeof(double) * (size_t)h->cfg.n_runs * h->cr_rows * h->cr_cols, cudaMemcpyDeviceToHost));
  return HB_OK;
}

int hb_fetch_rstat(hb_handle* h, double* rstat) {
  if (!h || !rstat) return HB_ERR_INVALID;
  if (h->rstat_rows < 1) return fail(h, HB_ERR_STATE, "check_convergence was not enabled (or out_t <= 50)");
  cudaSetDevice(h->device);
  const long long nr = h->cfg.n_runs, rows = h->rstat_rows;
  std::vector<double> tmp((size_t)rows * nr * h->d);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(tmp.data(), h->rstat_dev, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (long long run = 0; run < nr; ++run)                     // run-major blocks of column-major [rstat_rows, d]
    for (long long r = 0; r < rows; ++r)
      for (int k = 0; k < h->d; ++k)
        rstat[(size_t)run * rows * h->d + r + rows * k] = tmp[((size_t)r * nr + run) * h->d + k];
  return HB_OK;
}

int hb_fetch_state(hb_handle* h, double* xt, double* lp, double* ll, double* lpost) {
  if (!h) return HB_ERR_INVALID;
  if (!h->have_initial) return fail(h, HB_ERR_STATE, "no state yet");
  cudaSetDevice(h->device);
  const long long nc = h->peer_mode ? h->n_local : (h->modeb ? h->nct : h->nc), nl = h->n_local;
  if (xt) {   // peer mode: the rank's shard [n_local, d]; otherwise the whole population [nc, d]
    double* tmp = nullptr;
    CK(dalloc(&tmp, (size_t)nc * h->d));
    CK(hb_launch_rows_to_colmajor(h->peer_mode ? h->Xbuf[h->cur] : h->X, h->ldx, h->d, nc, tmp, h->stream));
    CK(cudaMemcpyAsync(xt, tmp, sizeof(double) * (size_t)nc * h->d, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(tmp);
  }
  if (lp) CK(cudaMemcpy(lp, h->lp, sizeof(double) * nl, cudaMemcpyDeviceToHost));
  if (ll) CK(cudaMemcpy(ll, h->ll, sizeof(double) * nl, cudaMemcpyDeviceToHost));
  if (lpost) CK(cudaMemcpy(lpost, h->lpost, sizeof(double) * nl, cudaMemcpyDeviceToHost));
  return HB_OK;
}

int hb_fetch_accept(hb_handle* h, uint8_t* accept) {
  if (!h || !accept) return HB_ERR_INVALID;
  if (!h->accept_rec) return fail(h, HB_ERR_STATE, "record_accept was not enabled");
  cudaSetDevice(h->device);
  CK(cudaMemcpy(accept, h->accept_rec, (size_t)(h->t - 1) * h->n_local, cudaMemcpyDeviceToHost));
  return HB_OK;
}

int hb_fetch_outliers(hb_handle* h, int64_t* n_pairs, int64_t* gen, int64_t* chain, int64_t cap) {
  if (!h || !n_pairs) return HB_ERR_INVALID;
  if (h->modeb) { if (int rc = hb_seq_pull_outliers(h)) return rc; }
  *n_pairs = (int64_t)h->outlier_gen.size();
  if (gen && chain)
    for (int64_t q = 0; q < *n_pairs && q < cap; ++q) { gen[q] = h->outlier_gen[(size_t)q]; chain[q] = h->outlier_chain[(size_t)q]; }
  return HB_OK;
}

int hb_rstat(hb_handle* h, double* out_d) {
  if (!h || !out_d) return HB_ERR_INVALID;
  if (!h->initialised) return fail(h, HB_ERR_STATE, "nothing has been run yet");
  if (h->ind_out < 3) return fail(h, HB_ERR_INVALID, "R_stat needs at least 3 stored generations");
  cudaSetDevice(h->device);
  double* o = nullptr;
  CK(dalloc(&o, (size_t)h->d));
  if (h->group_member && h->world > 1) { cudaFree(o); return fail(h, HB_ERR_UNSUPPORTED, "hb_rstat on a rank emulation: use check_convergence (the collective needs all ranks in lockstep)"); }
  {
    prof_scope ps(h, "K4");
    gen_prog g;
    prog_rstat_population(h, g, h->ind_out, o);   // collective on a sharded population
    std::vector<hb_handle*> hs{h}; std::vector<gen_prog> gs{g};
    if (int rc = run_programs(hs, gs)) { cudaFree(o); return rc; }
    ps.done(0);
  }
  CK(cudaMemcpyAsync(out_d, o, sizeof(double) * h->d, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(o);
  return HB_OK;
}

int hb_rstat_runs(hb_handle* h, double* out) {
  if (!h || !out) return HB_ERR_INVALID;
  if (!h->initialised) return fail(h, HB_ERR_STATE, "nothing has been run yet");
  if (h->ind_out < 3) return fail(h, HB_ERR_INVALID, "R_stat needs at least 3 stored generations");
  if (h->world > 1) return fail(h, HB_ERR_UNSUPPORTED, "hb_rstat_runs on a sharded population is not wired yet");
  cudaSetDevice(h->device);
  const long long nr = h->cfg.n_runs, nct = h->modeb ? h->nct : h->nc;
  double* o = nullptr;
  CK(dalloc(&o, (size_t)nr * h->d));
  prof_scope ps(h, "K4");
  CK(hb_launch_rstat_runs(h->hist_x, h->ind_out, nct, h->nc, nr, h->d, h->rstat_xm, h->rstat_ss, o, h->stream));
  ps.done(2);
  CK(cudaMemcpyAsync(out, o, sizeof(double) * (size_t)nr * h->d, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(o);
  return HB_OK;
}

int hb_get_timing(const hb_handle* h, double* loop_ms, int64_t* kernel_launches) {
  if (!h) return HB_ERR_INVALID;
  if (loop_ms) *loop_ms = h->loop_ms;
  if (kernel_launches) *kernel_launches = h->launches;
  return HB_OK;
}

int hb_profile(hb_handle* h, int enable) {
  if (!h) return HB_ERR_INVALID;
  h->profile = enable != 0;
  return HB_OK;
}

int hb_get_kernel_ms(const hb_handle* h, const char* kernel, double* total_ms, int64_t* launches) {
  if (!h || !kernel) return HB_ERR_INVALID;
  auto it = h->kstats.find(kernel);
  if (total_ms) *total_ms = it == h->kstats.end() ? 0.0 : it->second.ms;
  if (launches) *launches = it == h->kstats.end() ? 0 : it->second.launches;
  return HB_OK;
}

// ---- stand-alone entry points ---------------------------------------------------------------------------
static void set_msg(char* err, int errlen, const char* m) { if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", m); }

int hb_rstat_array(const double* x, int64_t t, int32_t d, int64_t n, double* out, char* err, int errlen) {
  if (!x || !out || t < 3 || d < 1 || n < 2) { set_msg(err, errlen, "R_stat: need x[t >= 3, d >= 1, n >= 2]"); return HB_ERR_INVALID; }
  if (hb_device_count() < 1) { set_msg(err, errlen, "no CUDA device is available and there is no CPU fallback"); return HB_ERR_NO_DEVICE; }
  double *xd = nullptr, *xm = nullptr, *ss = nullptr, *od = nullptr;
  const size_t tot = (size_t)t * d * n;
  bool ok = cudaMalloc((void**)&xd, tot * 8) == cudaSuccess && cudaMalloc((void**)&xm, (size_t)n * d * 8) == cudaSuccess &&
            cudaMalloc((void**)&ss, (size_t)n * d * 8) == cudaSuccess && cudaMalloc((void**)&od, (size_t)d * 8) == cudaSuccess;
  ok = ok && cudaMemcpy(xd, x, tot * 8, cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && hb_launch_rstat_colmajor(xd, t, d, n, xm, ss, od, nullptr) == cudaSuccess;
  ok = ok && cudaMemcpy(out, od, (size_t)d * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  cudaFree(xd); cudaFree(xm); cudaFree(ss); cudaFree(od);
  if (!ok) { set_msg(err, errlen, cudaGetErrorString(cudaGetLastError())); return HB_ERR_CUDA; }
  return HB_OK;
}

int hb_logdensity(const char* name, int32_t lik, const double* data, const int64_t* dims, int ndims, const double* obs,
                  int64_t n_obs, double glue_shape, const double* x, int64_t n, int32_t d, double* ll_out, double* fx_out,
                  char* err, int errlen) {
  if (!name || !x || !ll_out || n < 1 || d < 1) { set_msg(err, errlen, "hb_logdensity: bad arguments"); return HB_ERR_INVALID; }
  if (hb_device_count() < 1) { set_msg(err, errlen, "no CUDA device is available and there is no CPU fallback"); return HB_ERR_NO_DEVICE; }
  const hb_target_vtable* vt = hb_find_target(name);
  if (!vt) { set_msg(err, errlen, "unknown target"); return HB_ERR_UNKNOWN_TARGET; }
  void* ctx = nullptr;
  int rc = vt->init(&ctx, d, lik, data, dims, ndims, obs, n_obs, glue_shape, err, errlen);
  if (rc) return rc;
  const int ldx = d == 1 ? 1 : (d + 3) & ~3;
  double *xc = nullptr, *xr = nullptr, *ll = nullptr, *fx = nullptr;
  bool ok = cudaMalloc((void**)&xc, (size_t)n * d * 8) == cudaSuccess && cudaMalloc((void**)&xr, (size_t)n * ldx * 8) == cudaSuccess &&
            cudaMalloc((void**)&ll, (size_t)n * 8) == cudaSuccess && cudaMalloc((void**)&fx, (size_t)n * 8) == cudaSuccess;
  ok = ok && cudaMemcpy(xc, x, (size_t)n * d * 8, cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaMemset(fx, 0, (size_t)n * 8) == cudaSuccess;
  ok = ok && hb_launch_colmajor_to_rows(xc, n, 0, n, d, ldx, xr, nullptr) == cudaSuccess;
  if (ok) rc = vt->launch(ctx, xr, n, d, ldx, ll, fx, nullptr);
  ok = ok && rc == HB_OK && cudaDeviceSynchronize() == cudaSuccess;
  ok = ok && cudaMemcpy(ll_out, ll, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (ok && fx_out) ok = cudaMemcpy(fx_out, fx, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  unsigned dom = 0;
  if (ok && (strcmp(name, "HydroModel") == 0 || strcmp(name, "fun_wrap_glue") == 0)) dom = hb_hydro_poll_err(ctx);
  vt->destroy(ctx);
  cudaFree(xc); cudaFree(xr); cudaFree(ll); cudaFree(fx);
  if (!ok) { set_msg(err, errlen, rc ? "target launch failed" : cudaGetErrorString(cudaGetLastError())); return rc ? rc : HB_ERR_CUDA; }
  if (dom) { set_msg(err, errlen, "Argument 'param' has to be >= zero!"); return HB_ERR_INVALID; }
  for (int64_t i = 0; i < n; ++i) {                         // calc_ll floor and check, R/internals.R:19-21
    double v = ll_out[i];
    if (!(v >= HB_FLOOR)) v = (v != v) ? v : HB_FLOOR;
    if (!std::isfinite(v)) { set_msg(err, errlen, "Calculated log-likelihood is not finite!"); return HB_ERR_NONFINITE; }
    ll_out[i] = v;
  }
  return HB_OK;
}

// ---- log-density session: the same evaluation as hb_logdensity with the plugin context and the device buffers kept ----
// between calls (rwm() / am() / am_sd() evaluate the target once per iteration, R/rwm.R:42: per-call setup would dominate).
struct hb_density {
  const hb_target_vtable* vt = nullptr;
  void* ctx = nullptr;
  bool hydro = false;
  int32_t d = 0;
  int ldx = 0;
  int64_t cap = 0;
  double *xc = nullptr, *xr = nullptr, *ll = nullptr, *fx = nullptr;
};

int hb_density_open(const char* name, int32_t lik, const double* data, const int64_t* dims, int ndims, const double* obs,
                    int64_t n_obs, double glue_shape, int32_t d, int64_t max_n, hb_density** out, char* err, int errlen) {
  if (!name || !out || d < 1 || max_n < 1) { set_msg(err, errlen, "hb_density_open: bad arguments"); return HB_ERR_INVALID; }
  *out = nullptr;
  if (hb_device_count() < 1) { set_msg(err, errlen, "no CUDA device is available and there is no CPU fallback"); return HB_ERR_NO_DEVICE; }
  const hb_target_vtable* vt = hb_find_target(name);
  if (!vt) { set_msg(err, errlen, "unknown target"); return HB_ERR_UNKNOWN_TARGET; }
  hb_density* s = new hb_density();
  s->vt = vt; s->d = d; s->cap = max_n; s->ldx = d == 1 ? 1 : (d + 3) & ~3;
  s->hydro = strcmp(name, "HydroModel") == 0 || strcmp(name, "fun_wrap_glue") == 0;
  int rc = vt->init(&s->ctx, d, lik, data, dims, ndims, obs, n_obs, glue_shape, err, errlen);
  if (rc) { delete s; return rc; }
  const bool ok = cudaMalloc((void**)&s->xc, (size_t)max_n * d * 8) == cudaSuccess &&
                  cudaMalloc((void**)&s->xr, (size_t)max_n * s->ldx * 8) == cudaSuccess &&
                  cudaMalloc((void**)&s->ll, (size_t)max_n * 8) == cudaSuccess && cudaMalloc((void**)&s->fx, (size_t)max_n * 8) == cudaSuccess;
  if (!ok) { set_msg(err, errlen, cudaGetErrorString(cudaGetLastError())); hb_density_close(s); return HB_ERR_CUDA; }
  *out = s;
  return HB_OK;
}

int hb_density_eval(hb_density* s, const double* x, int64_t n, double* ll_out, double* fx_out, char* err, int errlen) {
  if (!s || !x || !ll_out || n < 1 || n > s->cap) { set_msg(err, errlen, "hb_density_eval: bad arguments (1 <= n <= max_n of the session)"); return HB_ERR_INVALID; }
  const int32_t d = s->d;
  int rc = HB_OK;
  bool ok = cudaMemcpy(s->xc, x, (size_t)n * d * 8, cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaMemset(s->fx, 0, (size_t)n * 8) == cudaSuccess;
  ok = ok && hb_launch_colmajor_to_rows(s->xc, n, 0, n, d, s->ldx, s->xr, nullptr) == cudaSuccess;
  if (ok) rc = s->vt->launch(s->ctx, s->xr, n, d, s->ldx, s->ll, s->fx, nullptr);
  ok = ok && rc == HB_OK && cudaDeviceSynchronize() == cudaSuccess;
  ok = ok && cudaMemcpy(ll_out, s->ll, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (ok && fx_out) ok = cudaMemcpy(fx_out, s->fx, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (!ok) { set_msg(err, errlen, rc ? "target launch failed" : cudaGetErrorString(cudaGetLastError())); return rc ? rc : HB_ERR_CUDA; }
  if (s->hydro && hb_hydro_poll_err(s->ctx)) { set_msg(err, errlen, "Argument 'param' has to be >= zero!"); return HB_ERR_INVALID; }
  for (int64_t i = 0; i < n; ++i) {                         // calc_ll floor and check, R/internals.R:19-21
    double v = ll_out[i];
    if (!(v >= HB_FLOOR)) v = (v != v) ? v : HB_FLOOR;
    if (!std::isfinite(v)) { set_msg(err, errlen, "Calculated log-likelihood is not finite!"); return HB_ERR_NONFINITE; }
    ll_out[i] = v;
  }
  return HB_OK;
}

int32_t hb_density_dim(const hb_density* s) { return s ? s->d : 0; }

void hb_density_close(hb_density* s) {
  if (!s) return;
  if (s->ctx) s->vt->destroy(s->ctx);
  cudaFree(s->xc); cudaFree(s->xr); cudaFree(s->ll); cudaFree(s->fx);
  delete s;
}

}  // extern "C"
