// hb_kernels.h -- host-side launch interface of the population (synchronous sweep) kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hb_common.cuh"

extern int g_hb_force_wide;   // hb_test_force_wide(): select the large-population kernels regardless of the population size

// device-resident copy of an hb_tape (all pointers are device pointers; rows of generation g start at g*nct)
struct hb_tape_dev {
  const double* u_snooker;
  const uint8_t* D;
  const int32_t* partners;
  const uint8_t* id;
  const double* zt;
  const double* lambda;
  const uint8_t* g_is_one;
  const double* zeta;
  const double* g_snooker;
  const double* u_accept;
  long long nct;
};

// ---- K1: proposal (calc_prop + bound_par + prior_pdf), R/internals.R:115-219, 89-113, 49-69 -------------
struct hb_k1_params {
  const double* X;        // population [nc][ldx]: partner rows (and own row) are read from here
  const double* Z;        // archive [m_arch][ldx] when past_sample, else nullptr
  long long m_arch;
  double* XP;             // proposals of the local shard [n_local][ldx]
  int* id_out;            // [n_local] crossover index
  double* lp_xp;          // [n_local] log prior of the proposal (floored)
  const hb_ctrl* ctrl;    // pCR
  unsigned* err;
  long long nc;           // chains per run
  long long chain0;       // item i of this launch is chain chain0 + i*stride of the handle (run-major index);
  long long stride;       //   synchronous sweep: (shard start, 1) ; sequential sweep, chain j of every run: (j, nc)
  long long n_local;      // items handled by this launch
  long long n_runs;       // independent runs batched in the handle (ctrl is an array of n_runs blocks)
  long long gchain0;      // offset added to the handle's chain index to form the Philox counter (run_offset * nc)
  int d, ldx;
  int delta, nCR;
  double c_val, c_star, p_g, beta0, psnooker;
  int past_sample;
  int bound;              // HB_BOUND_* ; 0 when min/max unset
  int prior;              // HB_PRIOR_*
  const double* pmin;     // [d] device, or nullptr
  const double* pmax;
  uint64_t seed;
  uint32_t gen;           // R's generation index i (2..t)
  hb_tape_dev tape;
  long long tape_g;       // generation row of the tape (i - 2)
  const double* gamma_tab; // [delta][d]: 2.38/sqrt(2 D d*), built on the host with IEEE sqrt and division
  const double* peer[8];  // HB_EXCHANGE_PEER: rank r's shard (current buffer), chains [r*peer_nl, (r+1)*peer_nl)
  int n_peers; long long peer_nl;
  int force_wide;         // host decision: use the large-population kernels (the launch may be one PIECE of a big shard)
  int batch;              // wide kernels: chains a warp takes per batch (32, or 16 / 8 for pieces with fewer batches than warps)
  int reserve_ctas;       // wide kernels: CTAs left out of the persistent grid (room for the concurrent send kernel)
  // the built-in t_distribution target in its structured evaluation (hb_targets.cu): the large-population proposal kernel
  // folds the log-density of every proposal into its own sweep (4 d flops on values it holds in registers) and the plugin
  // launch -- a second pass over the proposals -- is skipped.  td_rs == nullptr: not requested.
  const double* td_rs;    // [d] 1 / sqrt(i)
  double td_konst, td_df; int td_lik;
  double* td_ll; double* td_fx;   // [n_local] the plugin's outputs (raw log-likelihood; target value or nullptr)
  uint32_t ks[20];        // Philox round keys of `seed` (constant bank operands in the kernel)
  uint32_t thr16[HB_MAX_NCR];   // hb_rng_zt_threshold16(CR_q) = ceil(CR_q 2^16 - 0.5): exact integer form of zt < CR[id]
};
// density_done (may be nullptr): set to true when the launch also wrote td_ll / td_fx (the caller then skips the plugin)
cudaError_t hb_launch_k1(const hb_k1_params& p, bool tape_mode, int sm_count, cudaStream_t st, bool* density_done = nullptr);

struct hb_k3_params;
// ---- fused generation for small populations (hb_k1_propose.cu): K1 + t-distribution density + K3 in one launch ----
struct hb_fused_tdist {
  const double* rs;              // [d] 1 / sqrt(i): the structured evaluation of the target (hb_targets.cu)
  int lik; double konst, df;
};
bool hb_fused_small_eligible(long long n_local, int d, int sm_count);
cudaError_t hb_launch_fused_small(const hb_k1_params& p, const hb_k3_params& q, const hb_fused_tdist& td, bool tape_mode,
                                  int sm_count, cudaStream_t st);

// persistent variant: ONE cooperative launch runs generations [i0, i1) (all behind the adaptation period), grid barrier per
// generation; the kernel derives the per-generation pointers itself
struct hb_persist_params {
  long long i0, i1;
  double* XA; double* XB;   // generation i reads XA when (i - i0) is even, XB otherwise, and writes the other one
  long long t; double burnin; int thin; int update_interval;
  long long ind_out0;       // stored rows before generation i0
  double* hist_x; double* hist_l; double* hist_fx;   // bases of the history arrays
  double* lpost_hist; long long H;
  uint8_t* accept_rec;      // base [t-1][n_local] or nullptr
  const double* u_accept;   // tape base [n_gen][nct] or nullptr
  long long tape_nct;
  long long pending0;       // evaluations since the last AR / CR fold
  double* out_ar; double* out_cr; long long ar_rows;
  unsigned* bar;            // [2], zero at launch: arrivals, released round
  int adapt;                // generations inside the adaptation period: every CTA parks its std_x / jump statistic sums of
                            // generation i in slot slot0 + (i - i0); the finalize kernels fold the slots after the launch, which
                            // must END at the next updateInterval boundary (pCR changes there; the next generation's draws,
                            // computed under the barrier, read it)
  int slot0;
  double* partials;         // adapt: [slots][grid][(nCR + 2) * d] per-CTA sums
  const double* shift;      // adapt: [d] column shift of the one-pass variance, constant over the interval
  int trace;                // HB_PERSIST_TRACE=1: per-phase clock counts of a few generations on stdout (diagnostics)
};
cudaError_t hb_launch_persist_small(const hb_k1_params& p, const hb_k3_params& q, const hb_fused_tdist& td,
                                    const hb_persist_params& pp, bool tape_mode, int sm_count, cudaStream_t st);

// ---- MT-DREAM (mt > 1): candidate selection and multi-try acceptance, R/dream_parallel.R:295-341 -----------------
struct hb_mt_params {
  int mt; long long n_local; int ldx;
  const double* XP;        // [mt][n_local][ldx] proposals, try-major
  const double* lp_xp;     // [mt][n_local]
  const double* ll_raw;    // [mt][n_local] plugin output before the floor
  const double* fx_xp;     // [mt][n_local] or nullptr
  const int* id;           // [mt][n_local]
  double* ZT;              // [n_local][ldx] selected candidates
  double* lp_zt; double* ll_zt; double* fx_zt; int* id_zt;   // [n_local]
  double* p1;              // [n_local] sum(exp(lpost_xp)) over the mt tries (written by select, read by accept)
  const double* lpost;     // [n_local] current log posterior
  uint8_t* accept_out;     // [n_local] decisions handed to K3
  const uint8_t* tape_pick;   // replay: [n_local] recorded sample(1:mt, 1, prob) decisions of this generation, else nullptr
  const double* u_accept;     // replay: [n_local], else nullptr
  uint64_t seed; uint32_t gen; long long gchain0;
  unsigned* err;
};
cudaError_t hb_launch_mt_select(const hb_mt_params& p, cudaStream_t st);
cudaError_t hb_launch_mt_accept(const hb_mt_params& p, cudaStream_t st);

// ---- K3: Metropolis accept + update + history + adaptation statistics ----------------------------------
struct hb_k3_params {
  double* X;              // [nc][ldx] updated in place (rows of the local shard)
  double* Xout;           // == X (in place) or the next buffer of a double-buffered shard (every row is written)
  const double* XP;       // [n_local][ldx]
  double* XP_rw;          // same buffer, writable: receives dx when write_dx
  const int* id;          // [n_local]
  const double* lp_xp;    // [n_local]
  const double* ll_raw;   // [n_local] plugin output before the floor
  const double* fx_xp;    // [n_local] or nullptr
  double* lp; double* ll; double* lpost; double* fx;   // [n_local] current values
  double* lpost_hist_row; // [n_local] row i of the lpost history, or nullptr (i > adapt*t)
  double* hist_x;         // [n_local][d] slot of stored generation, or nullptr when this generation is not stored
  double* hist_l;         // [n_local][3]
  double* hist_fx;        // [n_local] or nullptr
  uint8_t* accept_rec;    // [n_local] or nullptr
  hb_ctrl* ctrl;
  const double* shift;    // [d] column shift for the one-pass variance (previous generation's mean)
  double* partials;       // [nblocks][(nCR+2)*d] block partial sums (adaptation only) or nullptr
  long long chain0, n_local, gchain0;   // item i is chain chain0 + i*stride of the handle (see hb_k1_params)
  long long stride, nc, n_runs;
  long long state0;       // lp/ll/lpost/history arrays hold chains [state0, ...): index = chain - state0
  int d, ldx, nCR;
  int adapt;              // generation is inside the adaptation period: accumulate std / jump statistics
  // HB_EXCHANGE_DELTA (stage-and-apply): accepted rows are NOT written into X but appended to this rank's delta log
  // (row data + chain index, slot taken from a device counter); hb_delta_send_kernel ships the log to every peer's inbox
  // and hb_delta_apply_kernel copies all logs into the replica once every rank's proposals of the generation are done.
  double* log_rows;       // [cap][ldx] or nullptr (= update X in place)
  int* log_idx;           // [cap] chain index (within the handle) of every logged row
  unsigned* log_count;    // device counter of the piece
  int force_wide;         // host decision: use the large-population kernel (the launch may be one piece of a big shard)
  const uint8_t* accept_in;   // MT-DREAM: the decisions were taken by hb_mt_accept_kernel (nullptr otherwise)
  int lik22;              // lik == 22: deterministic ABC acceptance rule (R/internals.R:222-225), no uniform drawn
  int write_dx;           // per-run statistics path: overwrite XP[item] with the realised jump dx (0 if rejected)
  uint64_t seed;
  uint32_t gen;
  const double* u_accept; // tape row (device) or nullptr in Philox mode
};
// returns the grid size used (partials rows) through *nblocks_out; query with p.X == nullptr to size buffers
cudaError_t hb_launch_k3(const hb_k3_params& p, int sm_count, int* nblocks_out, size_t* smem_out, cudaStream_t st);
int hb_k3_grid(long long n_local, int d, int nCR, int sm_count, int force_wide, size_t* smem_out);

// ---- finalize: fold partials into J / n_id / n_acc, refresh shift, pCR update, AR/CR rows ---------------
struct hb_fin_params {
  hb_ctrl* ctrl;
  const double* partials; int nblocks;
  int reduce_slots;       // phase 0: partials holds this many [nblocks][E] blocks; block s is reduced into colsum/qsum + s*slot_stride (0 = 1)
  double* shift;          // [d]
  double* colsum;         // [2*d] reduced (sum, sumsq) of the shifted columns -- exposed for multi-GPU all-reduce
  double* qsum;           // [nCR*d]
  long long nc;
  int d, nCR;
  int adapt;              // accumulate J this generation
  int do_pcr;             // i <= adapt*t && i %% updateInterval == 0  (R/dream_parallel.R:403-411)
  int do_arcr;            // i %% updateInterval == 0                  (R/dream_parallel.R:425-433)
  double* out_ar; double* out_cr; long long ar_rows;   // column-major [ar_rows,2] / [ar_rows,nCR+1]
  int phase;              // 0: reduce partials -> colsum/qsum ; 1: consume colsum/qsum ; 2: both (single GPU)
  int n_slots;            // phase 1: number of per-generation (colsum, qsum) slots to fold, slot s at + s*slot_stride (0 = 1)
  long long slot_stride;
  long long n_eval_inc;   // nc (all ranks)
};
cudaError_t hb_launch_finalize(const hb_fin_params& p, cudaStream_t st);
#ifdef __CUDACC__
// the scalar part of the finalize step (one thread): fold the per-generation integer counters into n_id / n_acc / n_eval
// (R/dream_parallel.R:368-375), the pCR update (:407-411) and the AR / CR rows (:425-433)
__device__ __forceinline__ void hb_fin_counters(const hb_fin_params& p) {
  hb_ctrl* c = p.ctrl;
  const int nCR = p.nCR;
  for (int q = 0; q < nCR; ++q) { c->n_id[q] += (double)c->n_id_gen[q]; c->n_id_gen[q] = 0ull; }   // :368-369
  c->n_acc += (double)c->n_acc_gen; c->n_acc_gen = 0ull;                                          // :374
  c->n_eval += (double)p.n_eval_inc;                                                              // :375
  if (p.do_pcr) {                                         // R/dream_parallel.R:407-411
    bool anypos = false, anynan = false;
    for (int q = 0; q < nCR; ++q) { if (c->J[q] != c->J[q]) anynan = true; if (c->J[q] > 0) anypos = true; }
    if (anynan) atomicOr(&c->err, HB_FLAG_J_NAN);
    else if (anypos) {
      double s = 0.0;
      for (int q = 0; q < nCR; ++q) {
        double v = c->J[q] / c->n_id[q];
        if (v != v) v = 1.0 / nCR;                        // n_id == 0 -> NaN -> 1/nCR
        c->pCR[q] = v; s += v;
      }
      for (int q = 0; q < nCR; ++q) c->pCR[q] = c->pCR[q] / s;
    }
  }
  if (p.do_arcr) {                                        // :425-433
    c->i_ar += 1;
    const long long r = c->i_ar - 1;
    c->cum_eval += c->n_eval;
    if (r < p.ar_rows) {
      p.out_ar[r + p.ar_rows] = c->n_acc / c->n_eval;
      p.out_ar[r] = c->cum_eval;
      p.out_cr[r] = c->cum_eval;
      for (int q = 0; q < nCR; ++q) p.out_cr[r + p.ar_rows * (q + 1)] = c->pCR[q];
    }
    c->n_acc = 0; c->n_eval = 0;
  }
}
#endif

// ---- per-run bookkeeping (sequential sweep and batched independent runs): one CTA per run ---------------
struct hb_runfin_params {
  hb_ctrl* ctrl;          // [n_runs]
  double* X; int ldx, d, nCR;
  long long nc, nct, state0;
  const double* DX;       // XP buffer after K3 (write_dx): realised jumps of this round's items
  const int* id;          // [items]
  long long item0, item_run_stride; int items_per_run;   // items of run r: item0 + r*item_run_stride + (0..items_per_run-1)
  long long n_eval_inc;   // evaluations per run since the last fold
  int adapt, do_pcr, do_outlier, do_arcr;
  long long seq_row;      // >= 0: write the sequential sweep's AR/CR row (R/dream.R:266-267) before adapting
  double* lpost; double* lpost_hist; long long hist_ld, hist_rows;
  long long gen; int update_interval;
  const int32_t* tape_outlier;   // device copy of hb_tape::outlier_new, or nullptr (Philox)
  uint64_t seed; long long run_offset;
  double* out_ar; double* out_cr; long long ar_rows; int ar_cols, cr_cols;
  unsigned long long* outlier_count; long long* outlier_gen; long long* outlier_chain; long long outlier_cap;
  int npow2;              // filled by the launcher
};
size_t hb_runfin_smem(int d, long long nc, int* npow2_out);
cudaError_t hb_launch_runfin(const hb_runfin_params& p, long long n_runs, cudaStream_t st);

// ---- outlier check pieces (R/internals.R:71-87) ---------------------------------------------------------
cudaError_t hb_launch_proxy(const double* lpost_hist, long long nc_local, long long row0, long long row1,
                            long long ld, double* proxy, cudaStream_t st);
cudaError_t hb_launch_outlier_reset(double* X, int ldx, int d, double* lpost, double* lpost_hist_row,
                                    const double* lpost_src, long long chain0, long long n_local, const int* outl,
                                    const int* news, int n_out, cudaStream_t st);
struct hb_peer_ptrs { const double* p[8]; };
// ---- HB_EXCHANGE_DELTA over peer-mapped memory (hb_delta.cu; no NCCL on this path) -----------------------------------
// Every rank owns an "arena" that its peers map through CUDA IPC:
//   [0, 256)     flags: 64 x u32, flags[which * 8 + source rank], monotonic values (never reset)
//                which = 0 slice barrier, 1 delta pieces, 2 all-reduce, 3 all-gather, 4 error reduce; word 63 = block counter
//   [256, 512)   error-reduce slots (two parities)
//   reduce slots (two parities) | gather area | delta boxes [parity][source rank] (see hb_delta.cu)
#define HB_MAX_PIECES 16
#define HB_ARENA_ERR_OFF 256
#define HB_ARENA_RED_OFF 512
struct hb_arena_peers { unsigned char* p[8]; };
unsigned long long hb_peer_timeout_ns();   // HB_PEER_TIMEOUT_S (default 120): wall-clock bound of every device-side wait
// signal = store `value` into flags[which*8 + self] of every peer; wait = spin (bounded) until the local
// flags[which*8 + r] of every peer r reached `value`
cudaError_t hb_launch_flag_signal(hb_arena_peers peers, int which, int self, int world, unsigned value, cudaStream_t st);
cudaError_t hb_launch_flag_wait(const unsigned char* own_arena, int which, int self, int world, unsigned value, unsigned* err,
                                cudaStream_t st);
// small collectives, split into a POST and a FINISH launch (a rank emulation on one device runs every rank's post before
// any finish; real ranks just launch them back to back):
//   all-reduce: post = copy the local vectors into the own reduce slot, then signal flag 2; finish = wait for every
//   peer's flag, then sum the slots of all ranks IN RANK ORDER (deterministic, identical on every rank).
//   all-gather: post = copy the local arrays into the own gather area, signal flag 3; finish = wait, then read every
//   peer's area into the full-length local arrays.
struct hb_peer_reduce_desc {
  double* f64[2]; int n_f64[2];           // two double vectors, summed as doubles
  unsigned long long* u64; int n_u64;     // one integer vector (per-generation counters), summed as integers
};
cudaError_t hb_launch_peer_allreduce_post(hb_arena_peers peers, size_t slot_off, hb_peer_reduce_desc desc, int self, int world,
                                          unsigned seq, cudaStream_t st);
cudaError_t hb_launch_peer_allreduce_finish(hb_arena_peers peers, size_t slot_off, hb_peer_reduce_desc desc, int self, int world,
                                            unsigned seq, unsigned* err, cudaStream_t st);
cudaError_t hb_launch_peer_allgather_post(hb_arena_peers peers, size_t area_off, const double* local_a, const double* local_b,
                                          long long n_local, int self, int world, unsigned seq, cudaStream_t st);
cudaError_t hb_launch_peer_allgather_finish(hb_arena_peers peers, size_t area_off, double* full_a, double* full_b,
                                            long long n_local, int self, int world, unsigned seq, unsigned* err, cudaStream_t st);
// bitwise OR of the ranks' error words into every rank's control block (hb_run returns the same status everywhere)
cudaError_t hb_launch_peer_err_post(hb_arena_peers peers, size_t slot_off, const unsigned* err, unsigned extra, int self, int world,
                                    unsigned seq, cudaStream_t st);
cudaError_t hb_launch_peer_err_finish(hb_arena_peers peers, size_t slot_off, unsigned* err, int self, int world, unsigned seq,
                                      cudaStream_t st);
// NCCL exchange modes: error mask <-> one word per bit (summed by ncclAllReduce)
cudaError_t hb_launch_err_expand(const unsigned* err, unsigned extra, unsigned long long* bits16, cudaStream_t st);
cudaError_t hb_launch_err_collapse(unsigned* err, const unsigned long long* bits16, cudaStream_t st);
// delta log (stage-and-apply exchange of the accepted rows, see hb_delta.cu)
struct hb_delta_send_params {
  unsigned char* box[8];   // box (parity, self) inside every rank's arena; box[self] is local and holds the log K3 wrote
  unsigned char* arena[8]; // every rank's arena base (flags)
  size_t idx_off, rows_off;   // offsets of idx[cap] (i32) and rows[cap][ldx] (f64) inside a box; counts[HB_MAX_PIECES] at 0
  long long piece_off;     // first log slot of the piece
  int piece, ldx, self, world;
  unsigned* log_count;     // device counter K3 allocated the piece's slots from (reset by the kernel)
  unsigned* done_ctr;      // CTA completion counter (zero between launches)
  unsigned flag_value;     // generation * HB_MAX_PIECES + piece + 1
};
cudaError_t hb_launch_delta_send(const hb_delta_send_params& p, int ctas, cudaStream_t st);
struct hb_delta_apply_params {
  const unsigned char* box[8];   // the LOCAL boxes (parity, r), r = 0..world-1
  size_t idx_off, rows_off;
  long long piece_cap;     // log slots per piece
  int p0, p1;              // pieces [p0, p1)
  int ldx, self, world;
  long long nc;
  double* X;               // this rank's replica
  const unsigned* flags;   // own arena
  unsigned flag_value;     // generation * HB_MAX_PIECES + p1: every peer has shipped its pieces < p1
  unsigned* err; unsigned long long timeout_ns;
};
cudaError_t hb_launch_delta_apply(const hb_delta_apply_params& p, int sm_count, cudaStream_t st);
// the LAST piece of a generation needs no log on the receiving side: once every rank's proposals of the generation are done
// (flag 5, raised after each rank's last proposal kernel) nothing reads the generation-start state any more, so the rows
// of this rank's log are stored straight into every replica of X (its own included); the last CTA raises flag 6
struct hb_delta_push_params {
  const unsigned char* box;   // this rank's box (parity, self)
  size_t idx_off, rows_off;
  long long piece_off;        // first log slot of the piece
  int ldx, self, world;
  long long nc;
  double* X[8];               // every rank's replica (X[self] is local)
  unsigned char* arena[8];    // every rank's arena (flags)
  unsigned* log_count;        // slot counter of the piece (reset by the kernel)
  unsigned* done_ctr;
  const unsigned* flags;      // own arena
  unsigned gen;               // flag value: wait flags[5*8 + r] >= gen, then raise flags[6*8 + self] = gen on every peer
  unsigned* err; unsigned long long timeout_ns;
};
cudaError_t hb_launch_delta_push(const hb_delta_push_params& p, int sm_count, cudaStream_t st);
cudaError_t hb_launch_outlier_reset_peer(double* Xlocal, hb_peer_ptrs peers, long long peer_nl, int ldx, int d,
                                         double* lpost, double* lpost_hist_row, const double* lpost_src,
                                         long long chain0, long long n_local, const int* outl, const int* news,
                                         int n_out, cudaStream_t st);

size_t hb_outlier_workspace_bytes(long long nc);
cudaError_t hb_outlier_device(const double* proxy, long long nc, void* tmp, size_t tmp_bytes, double* sorted,
                              double* thr, int* list, int* count, cudaStream_t st);

// ---- history transposition into R's chain[out_t, d+3, nc] layout ----------------------------------------
cudaError_t hb_launch_chain_transpose(const double* hist_x, const double* hist_l, long long out_t_alloc,
                                      long long row0, long long nrows, long long rows_filled, long long nct, int d,
                                      long long c0, long long ncs, double* out, cudaStream_t st);

// ---- K4: Gelman-Rubin (R/gelman_rubin.R:21-56) on history [rows][nct][d] ---------------------------------
cudaError_t hb_launch_rstat(const double* hist_x, long long t_rows, long long nct_stride, long long c0, long long n,
                            int d, double* xm_scratch, double* ss_scratch, double* out_d, cudaStream_t st);
// same on an R-layout array x[t, d, n] (column-major)
// sharded population (multi-GPU): pass 1 + local sums [2d] -> (all-reduce) -> squared deviations from the global mean [d]
// -> (all-reduce) -> R_stat
cudaError_t hb_launch_rstat_shard_pass1(const double* hist_x, long long t_rows, long long n_local, int d, double* xm,
                                        double* ss, double* sums2d, cudaStream_t st);
cudaError_t hb_launch_rstat_shard_dev(const double* xm, long long n_local, int d, long long n_total, const double* sums2d,
                                      double* bsum, cudaStream_t st);
cudaError_t hb_launch_rstat_shard_final(const double* sums2d, const double* bsum, long long t_rows, long long n_total, int d,
                                        double* out_d, cudaStream_t st);
// per-run variant: out[run*d + k] for runs of nc consecutive chains (pass 1 over all chains, pass 2 one warp per (run, k))
cudaError_t hb_launch_rstat_runs(const double* hist_x, long long t_rows, long long nct, long long nc, long long n_runs,
                                 int d, double* xm_scratch, double* ss_scratch, double* out_runs_d, cudaStream_t st);
cudaError_t hb_launch_rstat_colmajor(const double* x, long long t, int d, long long n, double* xm_scratch,
                                     double* ss_scratch, double* out_d, cudaStream_t st);

// initial evaluation helpers
cudaError_t hb_launch_init_state(const double* lp_in, const double* ll_raw, const double* fx_in, double* lp, double* ll,
                                 double* lpost, double* fx, double* lpost_hist_row, long long n, unsigned* err,
                                 cudaStream_t st);
cudaError_t hb_launch_prior(const double* X, int ldx, int d, long long n, int prior, const double* pmin,
                            const double* pmax, double* lp, unsigned* err, cudaStream_t st);
// prior = "normal" (R/internals.R:59): lp[j] = konst - 0.5 * || (x_j - mu) W ||^2, floored (:65); one warp per row
cudaError_t hb_launch_prior_normal(const double* X, int ldx, int d, long long n, const double* W, const double* mu,
                                   double konst, double* lp, unsigned* err, cudaStream_t st);
cudaError_t hb_launch_store_row(const double* X, int ldx, int d, long long n, const double* lp, const double* ll,
                                const double* lpost, const double* fx, double* hist_x, double* hist_l, double* hist_fx,
                                cudaStream_t st);
cudaError_t hb_launch_colmajor_to_rows(const double* in_colmajor, long long n_rows_total, long long row0, long long n,
                                       int d, int ldx, double* out, cudaStream_t st);
cudaError_t hb_launch_rows_to_colmajor(const double* in, int ldx, int d, long long n, double* out_colmajor,
                                       cudaStream_t st);
cudaError_t hb_launch_fill(double* p, long long n, double v, cudaStream_t st);
