# record_reference.R -- records a decision-level tape (SURVEY Appendix D) and the outputs of the REAL HydroBayes
# package, so that oracle <-> reference parity can be pinned by anyone who has R (R is absent from the build image).
#
#   Rscript tools/record_reference.R /path/to/HydroBayes tests/golden/reference_c1
#
# It sources the reference's R files, masks runif / rnorm / sample inside that environment with recording wrappers,
# runs dream_parallel() on the BASELINE C1 call, and writes <out_prefix>_draws.csv (call, n, values) and
# <out_prefix>_chain.csv.  tools/reference_tape.py turns the draw log into an hb_tape; tests/test_reference_tape.py
# (skipped while the two files are absent) replays it through the oracle and compares with the recorded chain: that
# test is what pins the oracle to the real R package.
args <- commandArgs(trailingOnly = TRUE)
ref <- args[1]; out <- args[2]
env <- new.env()
for (f in c("internals.R", "dream_parallel.R", "dream.R", "gelman_rubin.R", "test_mixture.R", "test_HydroModel.R"))
  sys.source(file.path(ref, "R", f), envir = env)
log <- list()
rec <- function(name, val) { log[[length(log) + 1]] <<- list(name = name, val = val); val }
env$runif <- function(n, min = 0, max = 1) rec("runif", stats::runif(n, min, max))
env$rnorm <- function(n, mean = 0, sd = 1) rec("rnorm", stats::rnorm(n, mean, sd))
env$sample <- function(x, size, replace = FALSE, prob = NULL) rec("sample", base::sample(x, size, replace, prob))
set.seed(1312)
phi <- 0.6180339887498949
x0 <- matrix(-20 + 40 * ((1:10) * phi) %% 1, ncol = 1)
res <- env$dream_parallel(fun = "mixture", lik = 1, par.info = list(initial = "user", val_ini = x0, prior = "flat"),
                          nc = 10, t = 5000, d = 1, verbose = FALSE)
con <- file(paste0(out, "_draws.csv"), "w")
for (e in log) writeLines(paste(e$name, length(e$val), paste(format(e$val, digits = 17), collapse = " "), sep = ","), con)
close(con)
write.csv(res$chain[, , ], paste0(out, "_chain.csv"), row.names = FALSE)
