// hb_internal.h -- the handle behind the C ABI (host side, C++).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/hydrobayes_b200.h"
#include "hb_kernels.h"

const hb_target_vtable* hb_find_target(const char* name);
unsigned hb_hydro_poll_err(void* ctx);
int hb_hydro_set_abc(void* ctx, const double* abc_e, int64_t n_e, char* err, int errlen);

struct hb_nccl;  // hb_comm.cu

struct hb_kstat { double ms = 0; long long launches = 0; };

struct hb_handle {
  hb_config cfg{};
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string err;

  // derived sizes (R/dream_parallel.R:166-205)
  long long nc = 0, t = 0, nct = 0;
  int d = 0, ldx = 0;
  long long m0 = 0, m_arch = 0, m_cap = 0;
  int archive_update = 0;
  long long out_t = 0, ar_rows = 0, ar_cols = 0, cr_rows = 0, cr_cols = 0, H = 0;
  long long rstat_rows = 0;

  // multi-GPU shard of the population
  int rank = 0, world = 1;
  long long chain0 = 0, n_local = 0;
  hb_nccl* comm = nullptr;

  // HB_EXCHANGE_PEER: double-buffered shard + peers' buffers (CUDA IPC); X is then the VIRTUAL base
  // Xbuf[cur] - chain0*ldx so that X + chain*ldx addresses this rank's rows only
  bool peer_mode = false, peers_ready = false;
  double* Xbuf[2] = {nullptr, nullptr};
  int cur = 0;
  const double* peer_ptr[2][8] = {};
  void* peer_opened[2][8] = {};

  // HB_EXCHANGE_DELTA (hb_delta.cu): every rank keeps a full replica of X.  The shard is processed in `pieces` pieces per
  // generation; K3 appends the rows it accepted to this rank's delta log, a send kernel on the side stream ships every
  // piece's log into the peers' arenas while the next piece computes, and an apply kernel folds all logs into the replica.
  bool delta_mode = false;
  bool initial_sharded = false;   // set_initial uploaded this rank's rows only; peer_import completes the replicas
  double* peerX[8] = {};          // peer r's replica base (null for self); only used to complete the initial replicas
  void* peerX_opened[8] = {};
  unsigned char* arena = nullptr; // [flags | err slots | reduce slot x2 | gather area | delta boxes], see hb_kernels.h
  unsigned char* arena_peer[8] = {};   // every rank's arena as mapped here (own entry = arena)
  void* arena_opened[8] = {};
  size_t arena_red_bytes = 0, arena_gather_off = 0, arena_box_off = 0;
  size_t box_bytes = 0, box_idx_off = 0, box_rows_off = 0;
  // delta boxes: boxmem[src] = [2 parities][box_bytes], the log this rank RECEIVES from rank src (boxmem[rank]: its own log).
  // One allocation per source, so that a peer maps only the box it writes (CUDA IPC maps whole allocations, and mapping
  // costs time in proportion to the size: 0.35 s per rank at 8 GPUs when every peer mapped all eight boxes).
  unsigned char* boxmem[8] = {};
  unsigned char* peer_box[8] = {};     // rank r's boxmem[this rank], as mapped here (own entry = boxmem[rank])
  void* peer_box_opened[8] = {};
  unsigned red_seq = 0, gather_seq = 0, err_seq = 0, slice_seq = 0;
  int pieces = 1;                 // pieces per generation (1 unless delta_mode)
  long long piece_n = 0;          // chains per piece (the last piece may be shorter)
  int k3_blocks_piece = 0;        // partial-sum rows K3 writes per piece
  bool wide = false;              // the shard is large enough for the large-population kernels (decided on n_local, not per piece)
  int reserve_ctas = 0;           // CTAs the persistent proposal grid leaves free for the concurrent send kernel
  int send_ctas = 0;
  unsigned* log_ctr = nullptr;    // [2 parities][HB_MAX_PIECES] slot counters + [HB_MAX_PIECES] CTA completion counters
  double* ibuf = nullptr;         // [update_interval][(2 + nCR) d] per-generation adaptation sums of the current interval
  int islot = 0;
  cudaStream_t side = nullptr;    // high-priority stream of the send kernels (== stream in a rank emulation)
  cudaStream_t side2 = nullptr;   // high-priority stream of the early apply (pieces < P - 1, behind the last piece's K2 / K3)
  cudaEvent_t ev_piece[HB_MAX_PIECES] = {};
  cudaEvent_t ev_sent = nullptr, ev_k1last = nullptr, ev_applied = nullptr;
  unsigned long long* errbits_dev = nullptr;   // NCCL modes: error mask expanded to one word per bit for the sum all-reduce
  // rank emulation on one device (hb_group_attach): the handles share one stream and reference each other's buffers
  bool group_member = false;
  cudaStream_t own_stream = nullptr;

  // MT-DREAM (mt > 1): candidate population and its scalars, multi-try sums, decisions
  double* ZT = nullptr; double* lp_zt = nullptr; double* ll_zt = nullptr; double* fx_zt = nullptr; int* id_zt = nullptr;
  double* mt_p1 = nullptr; uint8_t* mt_accept = nullptr;
  double* rstat_sums = nullptr;   // [3d] scratch of the sharded R_stat (sums, then deviations)
  const uint8_t* tape_mt_pick = nullptr;
  // likelihood closures named by the R call: abc_rho / abc_e (lik 21, 22), lik_fun (lik 99)
  std::string abc_rho, lik_fun;
  std::vector<double> abc_e_h;
  // bounds / prior
  bool have_prior_normal = false;     // prior = "normal": dmvnorm(x, mu, cov, log = TRUE), R/internals.R:59
  double* prior_W = nullptr;          // [d][d] row-major, W = chol(cov)^-1 (upper triangular)
  double* prior_mu = nullptr;         // [d]
  double prior_const = 0.0;           // -sum(log(diag(chol(cov)))) - d/2 log(2 pi)
  bool have_min_max = false;
  std::vector<double> pmin_h, pmax_h;
  double *pmin = nullptr, *pmax = nullptr;
  std::vector<double> obs_h;

  // target plugin
  const hb_target_vtable* vt = nullptr;
  void* tctx = nullptr;
  std::string target_name;
  bool target_is_hydro = false;

  // device state
  double *X = nullptr, *Z = nullptr, *XP = nullptr;
  // fused small-population generation (hb_launch_fused_small): second population buffer, swapped with X every generation
  double* Xalt = nullptr;
  int fused_state = 0;                    // 0 undecided, 1 available, -1 not applicable to this handle
  int persist_state = 0;                  // persistent generation loop (one cooperative launch per slice): same convention
  unsigned* grid_bar = nullptr;           // [2] grid barrier words of the persistent kernel
  hb_fused_tdist fused_td{};
  int k1_td_state = 0;                    // large populations: log-density folded into the proposal kernel; 0 undecided, 1 on, -1 off
  hb_fused_tdist k1_td{};
  double* persist_partials = nullptr;     // [update_interval][ceil(n_local / 8)][(nCR + 2) d]: per-CTA statistic sums parked by the persistent kernel
  double* persist_ibuf = nullptr;         // [update_interval][(nCR + 2) d]: their per-generation reductions
  int persist_adapt_ok = 0;               // 0 undecided, 1 on, -1 off (HB_PERSIST_ADAPT=0)
  int* id = nullptr;
  double *lp_xp = nullptr, *ll_raw = nullptr, *fx_xp = nullptr;
  double *lp = nullptr, *ll = nullptr, *lpost = nullptr, *fx = nullptr;
  double* lpost_hist = nullptr;           // [H][n_local]
  double *hist_x = nullptr, *hist_l = nullptr, *hist_fx = nullptr;   // [out_t][n_local][d|3|1]
  uint8_t* accept_rec = nullptr;          // [t-1][n_local]
  hb_ctrl* ctrl = nullptr;
  double *shift = nullptr, *colsum = nullptr, *qsum = nullptr, *partials = nullptr;
  int k3_blocks = 0;
  double *out_ar = nullptr, *out_cr = nullptr;
  double *proxy = nullptr, *proxy_full = nullptr, *lpost_full = nullptr;
  int *outl_dev = nullptr, *news_dev = nullptr;
  double *rstat_dev = nullptr, *rstat_xm = nullptr, *rstat_ss = nullptr;
  double* stage = nullptr; size_t stage_bytes = 0;
  void* pin = nullptr; size_t pin_bytes = 0;      // pinned host ring of the chain fetch / initial-state upload
  double* gamma_tab = nullptr;
  void* outl_tmp = nullptr; size_t outl_tmp_bytes = 0;
  double *outl_sorted = nullptr, *outl_thr = nullptr; int* outl_count = nullptr;

  // sequential sweep / batched independent runs (hb_seq.cu)
  bool modeb = false;
  unsigned long long* outl_count_dev = nullptr;
  long long *outl_gen_dev = nullptr, *outl_chain_dev = nullptr;
  long long outl_cap = 0;
  const int32_t* tape_outlier_dev = nullptr;

  // tape
  bool have_tape = false;
  hb_tape_dev tape{};
  std::vector<void*> tape_allocs;
  std::vector<int32_t> tape_outlier_new;  // host copy [n_events][nct]
  long long tape_events = 0, tape_n_gen = 0;

  // progress
  bool have_initial = false, initialised = false;
  long long gen = 1;                      // last completed generation (1 = initial state)
  long long ind_out = 1;
  long long pending_evals = 0;
  std::vector<long long> outlier_gen, outlier_chain;

  // timing
  double loop_ms = 0;
  long long launches = 0;
  bool profile = false;
  std::map<std::string, hb_kstat> kstats;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

hb_k1_params hb_make_k1(hb_handle* h, long long i, long long first, long long stride, long long n_items);
hb_k3_params hb_make_k3(hb_handle* h, long long i, long long first, long long stride, long long n_items, bool store,
                        bool in_adapt, bool write_dx);
int hb_prior_after_k1(hb_handle* h, long long n_items, long long off = 0);
int hb_target_launch(hb_handle* h, const double* x, long long n, long long first, long long stride, long long off = 0);
int hb_fail(hb_handle* h, int code, const char* msg);
// hb_targets.cu: true when vt/ctx is the built-in t_distribution plugin; fills the operands of the fused kernel
bool hb_tdist_fused_info(const hb_target_vtable* vt, void* ctx, hb_fused_tdist* out);

// hb_comm.cu: NCCL loaded lazily with dlopen (no link-time dependency; single-GPU use needs no NCCL)
int hb_nccl_unique_id(void* id128, std::string& err);
int hb_nccl_init(hb_nccl** out, int rank, int world, const void* id128, std::string& err);
void hb_nccl_destroy(hb_nccl* c);
int hb_nccl_allgather_f64(hb_nccl* c, const double* send, double* recv, size_t count, cudaStream_t st, std::string& err);
int hb_nccl_allreduce_f64(hb_nccl* c, double* buf, size_t count, cudaStream_t st, std::string& err);
int hb_nccl_allreduce_u64(hb_nccl* c, unsigned long long* buf, size_t count, cudaStream_t st, std::string& err);
