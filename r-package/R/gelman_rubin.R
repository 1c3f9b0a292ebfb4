# R_stat(x), x[samples, parameters, chains]: the device reduction replacing R/gelman_rubin.R:21-56 (same quirky
# normalisation, SURVEY F8).  Like the reference it returns its value invisibly (the last expression is an assignment).
R_stat <- function(x) invisible(.Call(hbR_rstat_array, x))
