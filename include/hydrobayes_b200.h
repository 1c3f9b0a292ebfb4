/*
 * hydrobayes_b200.h -- C ABI of libhydrobayes_b200.so
 *
 * The drop-in boundary for the DREAM generation step of tpilz/HydroBayes.
 * The reference has no FFI at all (it is 100 % interpreted R); these entry
 * points are what a thin `.Call` glue behind R/dream.R, R/dream_parallel.R and
 * R/gelman_rubin.R binds (INTEGRATION.md shows the glue).  Every function
 * returns an int status (HB_OK == 0); the message of the last failure is
 * available through hb_last_error().  The wording of the messages follows the
 * reference's stop() sites (R/internals.R:17,21,67,209,243, R/dream.R:137,
 * R/dream_parallel.R:167,175).
 *
 * Layout convention: every host array crossing this boundary uses R's
 * column-major layout, so that the R glue can pass REAL(x) directly:
 *   x0      [n0, d]            x0[j + n0*k]
 *   chain   [out_t, d+3, nc]   chain[it + out_t*(k + (d+3)*c)]      (R/dream_parallel.R:201)
 *   AR/CR   [rows, cols]       a[r + rows*c]                        (R/dream_parallel.R:204-205)
 *   R_stat  x[t, d, n]         x[it + t*(k + d*c)]                  (R/gelman_rubin.R:21)
 * No torch types, no CUDA types: plain pointers and sizes only.
 *
 * Threading: call from one host thread per handle; the library never calls
 * back into the caller.  There is NO CPU fallback: without a CUDA device
 * hb_create() fails with HB_ERR_NO_DEVICE.
 */
#ifndef HYDROBAYES_B200_H
#define HYDROBAYES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 1

/* status codes */
enum {
  HB_OK = 0,
  HB_ERR_INVALID = 1,      /* bad argument / configuration (message says which) */
  HB_ERR_NO_DEVICE = 2,    /* no CUDA device / driver: there is no CPU path */
  HB_ERR_CUDA = 3,         /* CUDA runtime error (message carries cudaGetErrorString) */
  HB_ERR_NONFINITE = 4,    /* "Calculated proposal is not finite!" etc. (R/internals.R:21,67,209,243) */
  HB_ERR_UNKNOWN_TARGET = 5, /* name not in the device plugin registry (get(fun) would fail in R) */
  HB_ERR_STATE = 6,        /* call order (e.g. hb_run before hb_set_initial) */
  HB_ERR_NCCL = 7,
  HB_ERR_UNSUPPORTED = 8   /* a combination that is not on the device path (e.g. mt > 1 with the delta exchange) */
};

/* sweep: which reference driver's data dependence to reproduce */
enum {
  HB_SWEEP_SYNCHRONOUS = 0, /* R/dream_parallel.R:257-435 : all proposals from generation-start state */
  HB_SWEEP_SEQUENTIAL = 1   /* R/dream.R:203-295        : chain j+1 sees chain j's update            */
};

/* par.info$prior (R/internals.R:49-69) */
enum { HB_PRIOR_FLAT = 0, HB_PRIOR_UNIFORM = 1, HB_PRIOR_NORMAL = 2 };
/* par.info$bound (R/internals.R:89-113); NONE == NULL in R */
enum { HB_BOUND_NONE = 0, HB_BOUND_BOUND = 1, HB_BOUND_REFLECT = 2, HB_BOUND_FOLD = 3 };
/* random numbers: decision tape (record/replay) or Philox4x32-10 keyed by (seed; chain, generation, slot) */
enum { HB_RNG_TAPE = 0, HB_RNG_PHILOX = 1 };
/* multi-GPU chain-state exchange */
enum { HB_EXCHANGE_ALLGATHER = 0 /* ncclAllGather of the shard's rows */,
       HB_EXCHANGE_PEER = 1      /* proposal kernel gathers partner rows through NVLink peer pointers */,
       HB_EXCHANGE_DELTA = 2     /* only ACCEPTED rows travel: every rank keeps a full replica; the accept kernel appends
                                    the rows it accepted to a delta log, a send kernel ships the log of one piece of the
                                    shard into the peers' arenas over NVLink (posted stores) while the next piece
                                    computes, an apply kernel folds all logs into the replica (traffic x acceptance
                                    rate, no barrier between proposals and exchange) */ };

/*
 * Mirrors the formals of dream() (R/dream.R:126-133) and dream_parallel()
 * (R/dream_parallel.R:155-163).  Fill with hb_config_default() first.
 */
typedef struct hb_config {
  int32_t abi_version;     /* HB_ABI_VERSION */
  int32_t sweep;           /* HB_SWEEP_* */
  int64_t nc;              /* chains per run */
  int64_t t;               /* samples per chain (generation 1 is the initial state) */
  int32_t d;               /* parameters */
  int32_t lik;             /* likelihood flag of calc_ll (R/internals.R:3-23): 1, 2, 21, 22, 31, 99 */
  int64_t n_runs;          /* independent runs batched in one handle (BASELINE C5); 1 = the reference call */
  double burnin;           /* 0   */
  double adapt;            /* 0.1 */
  int32_t update_interval; /* 10  */
  int32_t delta;           /* 3   */
  double c_val;            /* 0.1 */
  double c_star;           /* 1e-12 */
  int32_t nCR;             /* 3   */
  int32_t thin;            /* 1   */
  double p_g;              /* 0.2 */
  double beta0;            /* 1   */
  int32_t outlier_check;   /* TRUE */
  int32_t prior;           /* HB_PRIOR_*  (dream: "uniform", dream_parallel: "flat") */
  int32_t bound;           /* HB_BOUND_* ; only effective when min and max are set (R/internals.R:210) */
  int32_t past_sample;     /* DREAM(ZS) archive sampling (dream_parallel only) */
  int64_t m0;              /* 0 = reference default (10*d if past_sample else nc) */
  int32_t archive_update;  /* 0 = reference default (10 if past_sample else never) */
  int32_t mt;              /* multi-try (MT-DREAM, R/dream_parallel.R:295-341); synchronous sweep; several GPUs: HB_EXCHANGE_ALLGATHER */
  double psnooker;         /* 0 */
  double glue_shape;       /* lik == 31 */
  int32_t rng_mode;        /* HB_RNG_* */
  int32_t record_accept;   /* keep the accept mask of every generation (parity tests) */
  uint64_t seed;           /* Philox key */
  int32_t check_convergence; /* R_stat rows of the stored chain as dream_test does (R/dream_test.R:344); SURVEY F7 */
  int32_t store_fx;        /* keep raw target output per stored generation (only targets with m == 1) */
  int32_t exchange;        /* HB_EXCHANGE_* (multi-GPU only) */
  int32_t run_offset;      /* global index of this handle's first run: Philox chain id = (run_offset + run)*nc + j */
} hb_config;

/*
 * Decision-level tape.  One record per generation
 * i = 2..t (g = i-2) and global chain c = run*nc + j (nct = n_runs*nc).  It
 * records decisions (D, partner ids, crossover id, gamma choice) and the
 * values of the continuous draws as R returned them, so that it does not
 * depend on R's sample() internals.
 */
typedef struct hb_tape {
  int64_t n_gen;             /* t - 1 */
  int64_t nct;               /* n_runs * nc */
  int32_t d;
  int32_t delta;
  const double*  u_snooker;  /* [n_gen][nct]            R/internals.R:119 ; may be NULL when psnooker == 0 */
  const uint8_t* D;          /* [n_gen][nct]  1..delta  R/internals.R:123 */
  const int32_t* partners;   /* [n_gen][nct][2*delta]   0-based ids within the run (or archive rows if past_sample);
                                first D = a, next D = b (R/internals.R:143-144); snooker: 3 rows (:127) */
  const uint8_t* id;         /* [n_gen][nct]  0-based crossover index   R/internals.R:173 */
  const double*  zt;         /* [n_gen][nct][d]         R/internals.R:175 */
  const double*  lambda;     /* [n_gen][nct][d]         first d* used, ascending order of A (R/internals.R:188) */
  const uint8_t* g_is_one;   /* [n_gen][nct]            R/internals.R:192 */
  const double*  zeta;       /* [n_gen][nct][d]         first d* used (snooker: all d)  R/internals.R:194,158 */
  const double*  g_snooker;  /* [n_gen][nct]            R/internals.R:156 ; may be NULL */
  const double*  u_accept;   /* [n_gen][nct]            R/internals.R:241 */
  const int32_t* outlier_new;/* [n_events][nct] replacement chains (0-based, within run) for the k-th outlier of the
                                run in ascending order; event e is generation (e+1)*update_interval (R/internals.R:81) */
  int64_t n_events;
  /* MT-DREAM (mt > 1, R/dream_parallel.R:295-336): every generation draws n_sets = 2*mt - 1 proposal sets -- mt sets
   * around xt, then mt - 1 around the selected candidates zt.  The per-proposal arrays above then carry a set
   * dimension, [n_gen][n_sets][nct]... (set s of generation g is row g*n_sets + s); u_accept stays [n_gen][nct].
   * mt_pick is the decision of sample(1:mt, 1, prob = p) (:306), 0-based try index.  n_sets <= 1 and mt_pick NULL
   * for mt == 1. */
  int32_t n_sets;
  int32_t reserved;
  const uint8_t* mt_pick;    /* [n_gen][nct] */
} hb_tape;

typedef struct hb_handle hb_handle;

/* ---- device log-density plugins (the get(fun)(x, ...) boundary, R/dream_parallel.R:279-283) ---------- */
/*
 * A plugin evaluates calc_ll(get(fun)(x, ...)) for a batch of points resident
 * on the device: x_dev is row-major [n][ldx] (ldx >= d), ll_dev [n] receives
 * the value BEFORE the -708.396 floor of R/internals.R:19 (the sampler applies
 * the floor and the finiteness check), fx_dev (may be NULL) receives the raw
 * scalar target output when the target has m == 1.  `stream` is a cudaStream_t.
 */
typedef struct hb_target_vtable {
  const char* name;
  /* data/dims are the `...` of the R call (e.g. forcing for HydroModel); obs/glue_shape come from the sampler */
  int (*init)(void** ctx, int d, int lik, const double* data, const int64_t* dims, int ndims,
              const double* obs, int64_t n_obs, double glue_shape, char* err, int errlen);
  int (*launch)(void* ctx, const double* x_dev, int64_t n, int d, int ldx, double* ll_dev, double* fx_dev,
                void* stream);
  void (*destroy)(void* ctx);
  /* algorithmic work per evaluated point, for roofline reports */
  void (*work)(void* ctx, int d, double* flops, double* bytes);
  /* optional (may be NULL): targets with per-run data in batched independent runs.  Point i is chain
   * chain0 + i*stride of the handle, i.e. it belongs to run (chain0 + i*stride) / nc. */
  int (*launch_items)(void* ctx, const double* x_dev, int64_t n, int d, int ldx, double* ll_dev, double* fx_dev,
                      int64_t chain0, int64_t stride, int64_t nc, void* stream);
} hb_target_vtable;

int hb_register_target(const hb_target_vtable* vt);       /* user plugins; built-ins are pre-registered */
int hb_target_registered(const char* name);               /* 1 / 0 */

/* ---- sampler ------------------------------------------------------------------------------------------ */
void hb_config_default(hb_config* cfg);                    /* reference defaults of dream_parallel() */
int hb_device_count(void);                                 /* CUDA devices visible; 0 => nothing will run */
int hb_create(const hb_config* cfg, int device, hb_handle** out);
void hb_destroy(hb_handle* h);                             /* device buffers are released by a background thread */
void hb_wait_idle(void);                                   /* blocks until the releases of destroyed handles are done */
const char* hb_last_error(const hb_handle* h);             /* h may be NULL: last error of a failed hb_create */

int hb_set_bounds(hb_handle* h, const double* par_min, const double* par_max, int n); /* n == 1 (recycled) or d */
int hb_set_prior_normal(hb_handle* h, const double* mu, const double* cov_colmajor); /* par.info$mu, $cov */
int hb_set_obs(hb_handle* h, const double* obs, int64_t n);                           /* obs= */
/* abc_rho=, abc_e= (lik 21 / 22) and lik_fun= (lik 99) of dream()/dream_parallel() (R/dream.R:131-132; calc_ll,
 * R/internals.R:8-16).  The R API passes these functions BY NAME; like targets they resolve in a device registry:
 * abc_rho "abc_distfun" = abs(sim - obs) and lik_fun "fun_lik" = -n/2 log(sum((sim-obs)^2)), the two closures of the
 * reference's example script (examples/script_examples.R:494, 503-507).  Other names -> HB_ERR_UNKNOWN_TARGET (arbitrary
 * R closures stay off the GPU path).  abc_e: n_abc_e == 1 (recycled) or length(obs).  Call before hb_set_target. */
int hb_set_lik_args(hb_handle* h, const char* abc_rho, const double* abc_e, int64_t n_abc_e, const char* lik_fun);
int hb_set_target(hb_handle* h, const char* name, const double* data, const int64_t* dims, int ndims);
/* per-run target data for batched independent runs: data is [n_runs][len], obs is [n_runs][n_obs] (row-major) */
int hb_set_target_batched(hb_handle* h, const char* name, const double* data, int64_t len,
                          const double* obs, int64_t n_obs);
/* x0: [n_runs*n0, d] column-major, n0 = m0 if past_sample else nc; prior_sample() stays on the host side */
int hb_set_initial(hb_handle* h, const double* x0_colmajor);
/* same, for hosts whose matrices are row-major (C / numpy): x0[j*d + k]; no host-side transposition needed */
int hb_set_initial_rows(hb_handle* h, const double* x0_rowmajor);
int hb_set_tape(hb_handle* h, const hb_tape* tape);        /* rng_mode == HB_RNG_TAPE */

/* multi-GPU: one process per GPU; id is the ncclUniqueId (128 bytes) created by hb_comm_unique_id on rank 0 */
int hb_comm_unique_id(void* id128);
int hb_comm_init(hb_handle* h, int rank, int world, const void* id128);
/* HB_EXCHANGE_PEER: instead of all-gathering the chain state every generation, the proposal kernel fetches the partner
 * rows it needs straight from the owning GPU through NVLink-mapped peer pointers (bulk TMA reads), and the only
 * per-generation collective left is a 17-word all-reduce that doubles as the generation barrier.  After
 * hb_set_initial every rank exports HB_PEER_EXPORT_BYTES bytes of CUDA IPC handles, the host gathers them from all ranks
 * (rank order) and hands them to hb_peer_import. */
#define HB_PEER_EXPORT_HANDLES 10                      /* 64-byte CUDA IPC handles per rank */
#define HB_PEER_EXPORT_BYTES (64 * HB_PEER_EXPORT_HANDLES)
/* out: HB_PEER_EXPORT_BYTES bytes (HB_EXCHANGE_DELTA: replica, arena, one delta box per source rank; HB_EXCHANGE_PEER: the two
 * shard buffers); all_handles: world * HB_PEER_EXPORT_BYTES bytes in rank order */
int hb_peer_export(hb_handle* h, void* out);
int hb_peer_import(hb_handle* h, const void* all_handles);

/* Testing hook -- rank emulation on ONE device: hs[r] is rank r of n (hb_comm_init(hs[r], r, n, NULL), HB_EXCHANGE_DELTA,
 * hb_set_initial done).  hb_group_attach wires the handles to each other in-process (plain device pointers where real ranks
 * use CUDA IPC mappings) and puts them on one stream; hb_group_run drives them in lockstep: the same kernels and the same
 * arena protocol as n processes on n GPUs, with every rank's signalling step enqueued before any rank's waiting step, so
 * a one-GPU box can check that the shards of a sharded run equal the single-GPU run bit for bit. */
int hb_group_attach(hb_handle** hs, int n);
int hb_group_run(hb_handle** hs, int n, int64_t n_generations, int64_t* generations_done_total);

/* Runs up to n_generations further generations (resumable slices: the R driver ticks its progress bar and polls
 * interrupts between slices).  The first call also evaluates the initial population (R/dream_parallel.R:215-250). */
int hb_run(hb_handle* h, int64_t n_generations, int64_t* generations_done_total);
int hb_sync(hb_handle* h);

/* results; caller allocates (R: allocVector(REALSXP, ...)) */
int hb_out_dims(const hb_handle* h, int64_t* out_t, int64_t* ar_rows, int64_t* ar_cols, int64_t* cr_rows,
                int64_t* cr_cols, int64_t* fx_m, int64_t* rstat_rows);
int hb_fetch_chain(hb_handle* h, double* chain);           /* [out_t, d+3, nct]; rows not yet written are NaN (R: NA) */
int hb_fetch_chain_rows(hb_handle* h, int64_t row0, int64_t nrows, double* out); /* [nrows, d+3, nct] */
/* Touch every page of a freshly allocated host result array with a few threads (blocking; thread-safe; no device work).
 * The reference allocates `chain <- array(NA, ...)` BEFORE the generation loop (R/dream_parallel.R:201-203), i.e. its
 * result pages are faulted in up front; a host that allocates uninitialised memory (numpy.empty) can call this from a
 * helper thread while hb_run is busy, so that hb_fetch_chain copies into resident pages.  n_threads <= 0: default. */
int hb_host_prefault(void* p, size_t bytes, int n_threads);
/* Page-lock / release a host result array (cudaHostRegister): hb_fetch_chain then writes it by direct DMA at PCIe rate
 * instead of through the pinned ring and copier threads.  Optional; meant to run in the same helper thread as
 * hb_host_prefault while hb_run is busy.  HB_ERR_CUDA = the pages could not be locked (keep the pageable array). */
int hb_host_register(void* p, size_t bytes);
int hb_host_unregister(void* p);
int hb_fetch_fx(hb_handle* h, double* fx);                 /* [out_t, nct, m] */
int hb_fetch_ar(hb_handle* h, double* ar);                 /* run-major blocks of [ar_rows, ar_cols] */
int hb_fetch_cr(hb_handle* h, double* cr);
int hb_fetch_rstat(hb_handle* h, double* rstat);           /* run-major blocks of [rstat_rows, d] (either sweep) */
int hb_fetch_state(hb_handle* h, double* xt_colmajor, double* lp, double* ll, double* lpost);
int hb_fetch_accept(hb_handle* h, uint8_t* accept);        /* [t-1][nct] if record_accept */
/* outlier list (R/dream_parallel.R:418): call with buf == NULL to get the count of (generation, chain) pairs */
int hb_fetch_outliers(hb_handle* h, int64_t* n_pairs, int64_t* gen, int64_t* chain, int64_t cap);
/* Gelman-Rubin of the stored chain so far, rows 1..ind_out (R/gelman_rubin.R:21-56 on chain[1:ind_out,1:d,]) */
int hb_rstat(hb_handle* h, double* out_d);
/* the same per independent run of a batched handle: out [n_runs][d] row-major, one launch for all runs.  This is the
 * convergence monitor of BASELINE config 5 ("each run to R-hat < 1.2", SURVEY 8d metric 2) */
int hb_rstat_runs(hb_handle* h, double* out_runs_d);
int64_t hb_ind_out(const hb_handle* h);
/* device time (ms) spent inside the generation loop so far and number of kernels launched by this library */
int hb_get_timing(const hb_handle* h, double* loop_ms, int64_t* kernel_launches);
/* per-kernel device time of the last profiled slice: names K1,K2,K3,K4,XCHG ; enable with hb_profile(h,1) */
int hb_profile(hb_handle* h, int enable);
int hb_get_kernel_ms(const hb_handle* h, const char* kernel, double* total_ms, int64_t* launches);

/* testing hook: the kernels written for large populations (warp-per-32-chains proposal and accept kernels) are
 * normally selected only from ~38 000 chains per GPU upwards; on = 1 selects them for every d > 16 so that the parity
 * tests can compare them with the oracle at sizes the oracle finishes in seconds.  Returns the previous setting. */
int hb_test_force_wide(int on);

/* ---- stand-alone device entry points (unit parity tests and drop-ins for exported R functions) ------- */
/* R_stat(x), x [t, d, n] column-major host array -> out[d]          (R/gelman_rubin.R:21) */
int hb_rstat_array(const double* x, int64_t t, int32_t d, int64_t n, double* out, char* err, int errlen);
/* calc_ll(get(name)(x_i, data), lik, obs, glue_shape) floored, for n host points x [n, d] column-major */
int hb_logdensity(const char* name, int32_t lik, const double* data, const int64_t* dims, int ndims,
                  const double* obs, int64_t n_obs, double glue_shape, const double* x_colmajor, int64_t n,
                  int32_t d, double* ll_out, double* fx_out, char* err, int errlen);
/* The same evaluation as a session: the plugin context (covariance factor, forcing, observations) and the device buffers
 * are set up once and kept, hb_density_eval() evaluates up to max_n points per call.  For callers that evaluate the
 * target once per iteration: pdf(xp) of rwm() / am() / am_sd() (R/rwm.R:42, R/am.R:51, R/am_sd.R:59) and of de_mc()
 * (R/de_mc.R:56), get(fun) + calc_ll of game() (R/game.R:163-176).  Errors as hb_logdensity. */
typedef struct hb_density hb_density;
int hb_density_open(const char* name, int32_t lik, const double* data, const int64_t* dims, int ndims,
                    const double* obs, int64_t n_obs, double glue_shape, int32_t d, int64_t max_n, hb_density** out,
                    char* err, int errlen);
int hb_density_eval(hb_density* s, const double* x_colmajor, int64_t n, double* ll_out, double* fx_out, char* err,
                    int errlen);
int32_t hb_density_dim(const hb_density* s);   /* d the session was opened for (0 for NULL): callers check ncol(x) against it */
void hb_density_close(hb_density* s);
/* HydroModel(forcing, param) full series for n parameters: out [n][T] row-major  (R/test_HydroModel.R:17-36) */
int hb_hydromodel_series(const double* forcing, int64_t T, const double* param, int64_t n, double* out, char* err,
                         int errlen);

/* test hook: number of pairs for which HydroModel's reciprocal-based division differs from IEEE a/b (must be 0) */
int hb_divcheck(const double* a, const double* b, int64_t n, int64_t* n_bad);

/* fp64 pipe micro-benchmarks (TFLOP/s): roofline denominators for the log-density kernels (no reference analogue) */
int hb_fp64_peaks(double* dmma_tflops, double* dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* HYDROBAYES_B200_H */
