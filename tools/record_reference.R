# record_reference.R -- records a decision-level tape (SURVEY Appendix D) and the outputs of the REAL HydroBayes
# package, so that oracle <-> reference parity can be pinned by anyone who has R (R is absent from the build image AND
# from the GPU boxes: bench.py prints "rscript": "absent ..." in its JSON line).
#
#   Rscript tools/record_reference.R /path/to/HydroBayes tests/golden [case ...]
#
# cases (default: all three):
#   c1     BASELINE C1: dream_parallel, mixture, lik = 1, nc = 10, t = 5000, d = 1              -> reference_c1_*.csv
#   tdist  t_distribution d = 100, lik = 2, nc = 64, t = 200 (needs mvtnorm: pins dmvt)        -> reference_tdist_*.csv
#   hydro  HydroModel, GLUE lik = 31, glue_shape = 50, uniform prior, bound = "bound", nc = 10, t = 300, forcing / obs
#          read from tests/golden/reference_hydro_inputs.csv (written by the Python harness)   -> reference_hydro_*.csv
# It sources the reference's R files, masks runif / rnorm / sample inside that environment with recording wrappers,
# runs dream_parallel() and writes <case>_draws.csv (call, n, values) and <case>_chain.csv.  tools/reference_tape.py
# turns the draw log into an hb_tape; tests/test_reference_tape.py (skipped while the files are absent) replays it
# through the oracle and compares with the recorded chain at 1e-12: that test is what pins the oracle to the real R package.
args <- commandArgs(trailingOnly = TRUE)
ref <- args[1]; outdir <- args[2]
cases <- if (length(args) > 2) args[-(1:2)] else c("c1", "tdist", "hydro")
phi <- 0.6180339887498949
frac <- function(x) x - floor(x)
for (case in cases) {
  env <- new.env()
  for (f in c("internals.R", "dream_parallel.R", "dream.R", "gelman_rubin.R", "test_mixture.R", "test_t_distribution.R", "test_HydroModel.R"))
    sys.source(file.path(ref, "R", f), envir = env)
  log <- list()
  rec <- function(name, val) { log[[length(log) + 1]] <<- list(name = name, val = val); val }
  env$runif <- function(n, min = 0, max = 1) rec("runif", stats::runif(n, min, max))
  env$rnorm <- function(n, mean = 0, sd = 1) rec("rnorm", stats::rnorm(n, mean, sd))
  env$sample <- function(x, size, replace = FALSE, prob = NULL) rec("sample", base::sample(x, size, replace, prob))
  set.seed(1312)
  if (case == "c1") {
    x0 <- matrix(-20 + 40 * frac((1:10) * phi), ncol = 1)
    res <- env$dream_parallel(fun = "mixture", lik = 1, par.info = list(initial = "user", val_ini = x0, prior = "flat"),
                              nc = 10, t = 5000, d = 1, verbose = FALSE)
  } else if (case == "tdist") {
    nc <- 64; d <- 100
    x0 <- outer(1:nc, 1:d, function(j, k) -5 + 20 * frac((j * d + k) * phi))
    res <- env$dream_parallel(fun = "t_distribution", lik = 2, par.info = list(initial = "user", val_ini = x0, prior = "flat"),
                              nc = nc, t = 200, d = d, verbose = FALSE)
  } else if (case == "hydro") {
    io <- read.csv(file.path(outdir, "reference_hydro_inputs.csv"))
    x0 <- matrix(0.01 + 1.99 * frac((1:10) * phi), ncol = 1)
    # the reference calls get(fun)(x, ...), HydroModel's formals are (forcing, param): wrap it as the example script does
    # (examples/script_examples.R, fun_wrap_glue)
    env$fun_wrap_glue <- function(x, forcing) env$HydroModel(forcing, x)
    environment(env$fun_wrap_glue) <- env
    res <- env$dream_parallel(fun = "fun_wrap_glue", forcing = io$forcing, lik = 31, obs = io$obs, glue_shape = 50,
                              par.info = list(initial = "user", val_ini = x0, prior = "uniform", min = 0.01, max = 2, bound = "bound"),
                              nc = 10, t = 300, d = 1, verbose = FALSE)
  } else stop("unknown case ", case)
  con <- file(file.path(outdir, paste0("reference_", case, "_draws.csv")), "w")
  for (e in log) writeLines(paste(e$name, length(e$val), paste(format(e$val, digits = 17), collapse = " "), sep = ","), con)
  close(con)
  dm <- dim(res$chain)
  write.csv(matrix(res$chain, nrow = dm[1]), file.path(outdir, paste0("reference_", case, "_chain.csv")), row.names = FALSE)
}
