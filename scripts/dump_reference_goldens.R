#!/usr/bin/env Rscript
# dump_reference_goldens.R -- pin the CPU oracle (and through it the CUDA path) to the REAL reference.
#
# Nothing in this repository was ever produced by airGR / hydroGOF / TauDEM / R's round(): none of them exists
# in the build image, so oracle/ restates them from their published algorithms ("parity unpinned").  Run this
# script ONCE on any machine that has R (>= 4.0) with airGR and hydroGOF installed -- optionally also the
# reference package GR2MSemiDistr itself and the TauDEM executables -- and commit what it writes:
#
#     Rscript scripts/dump_reference_goldens.R            # from the repository root
#     git add tests/golden/ref && git commit -m "reference goldens from R <version> / airGR <version>"
#
# It reads the committed inputs tests/golden/pin_inputs/*.csv (the bundled example decoded by
# tests/golden/make_fixtures.py: forcing table, subbasin attributes, parameter sets, rounding cases) and writes
# tests/golden/ref/*.csv with 17 significant digits.  tests/test_reference_pin.py consumes those files when they
# exist (oracle on the CPU, CUDA path with -m gpu) and reports "parity unpinned" while they do not.
#
# Tiers (each writes its own files; later tiers are skipped when their packages are missing):
#   A  airGR + hydroGOF called exactly as the reference calls them
#        R/Run_GR2MSemiDistr.R:96-107,133-187,223-229   forcing correction, airGR chain, mm -> m3/s
#        R/Optim_GR2MSemiDistr.R:186-216                 outlet sum, warm-up, na.omit, KGE / NSE / rmse
#      -> ref_airgr_set<k>_{PR,AE,SM,RU,QS}.csv  (ntime x nsub, unrounded), ref_airgr_endstate.csv,
#         ref_objective.csv, ref_sink_set<k>.csv, ref_gof_degenerate.csv, ref_round.csv, ref_versions.csv
#      This tier also settles the one decisive from-memory detail: whether airGR's percolation exponent behaves
#      as the exact 1/3 (library default) or as the single-precision literal (flag GR2M_PERC_F32_EXPONENT).
#   B  the reference package itself, if installed: GR2MSemiDistr::Run_GR2MSemiDistr on the same table
#      -> ref_pkg_run_{QS,RU,PR,AE,SM,SINK}.csv  (rounded, warm-up dropped, as the function returns them)
#   C  TauDEM, if `aread8` (or `AreaD8`) and `mpiexec` are on the PATH and the raster package is installed:
#      AreaD8 -wg on the committed flow-direction grid, exactly as R/Routing_GR2MSemiDistr.R:109-118 runs it
#      -> ref_ad8_month<k>.csv.gz  (float32 accumulation grid, nodata as NA)

args <- commandArgs(trailingOnly = TRUE)
root <- if (length(args) >= 1) args[1] else "."
inp  <- file.path(root, "tests", "golden", "pin_inputs")
out  <- file.path(root, "tests", "golden", "ref")
if (!file.exists(file.path(inp, "data.csv"))) stop("run from the repository root (tests/golden/pin_inputs/data.csv not found)")
dir.create(out, recursive = TRUE, showWarnings = FALSE)

suppressPackageStartupMessages({ library(airGR); library(hydroGOF) })

wr <- function(x, name) {                       # 17 significant digits: doubles survive the round trip
  x <- as.data.frame(x)
  x[] <- lapply(x, function(v) if (is.numeric(v)) ifelse(is.na(v), "NA", sprintf("%.17g", v)) else v)
  utils::write.csv(x, file.path(out, name), row.names = FALSE, quote = FALSE)
}

Data  <- utils::read.csv(file.path(inp, "data.csv"), stringsAsFactors = FALSE)
Sub   <- utils::read.csv(file.path(inp, "subbasins.csv"), stringsAsFactors = FALSE)
Sets  <- as.matrix(utils::read.csv(file.path(inp, "param_sets.csv")))          # one row per set: 4 * nreg values
Cfg   <- utils::read.csv(file.path(inp, "config.csv"), stringsAsFactors = FALSE) # key,value
cfg   <- function(k) Cfg$value[Cfg$key == k]
nsub  <- nrow(Sub); area <- Sub$Area; region <- Sub$Region
Zone  <- sort(unique(region)); nreg <- length(Zone)

# ---- tier A ----------------------------------------------------------------------------------------
# Run_GR2MSemiDistr.R:74-79 date handling
Data$DatesR <- as.POSIXct(paste0(Data$DatesR, " 00:00:00"), "GMT", tryFormats = c("%Y-%m-%d", "%Y/%m/%d", "%d-%m-%Y", "%d/%m/%Y"))
numberOfDays <- function(date) {               # Run_GR2MSemiDistr.R:108-117
  m <- format(date, format = "%m")
  while (format(date, format = "%m") == m) date <- date + 1
  as.integer(format(date - 1, format = "%d"))
}

simulate <- function(Database, Parameters) {   # Run_GR2MSemiDistr.R:82-107,133-187 without the PSOCK cluster
  ntime <- nrow(Database)
  X1 <- Parameters[1:nreg]; X2 <- Parameters[(nreg + 1):(2 * nreg)]
  Fpp <- Parameters[(2 * nreg + 1):(3 * nreg)]; Fpe <- Parameters[(3 * nreg + 1):length(Parameters)]
  lapply(1:nsub, function(i) {
    r  <- match(region[i], Zone)
    Fix <- data.frame(DatesR = Database[, 1],
                      P = round(Fpp[r] * Database[, i + 1], 1),
                      E = round(Fpe[r] * Database[, i + 1 + nsub], 1))
    InputsModel <- CreateInputsModel(FUN_MOD = RunModel_GR2M, DatesR = Fix$DatesR, Precip = Fix$P, PotEvap = Fix$E)
    RunOptions  <- CreateRunOptions(FUN_MOD = RunModel_GR2M, InputsModel = InputsModel, IndPeriod_Run = 1:ntime,
                                    verbose = FALSE, warnings = FALSE)
    RunModel(InputsModel = InputsModel, RunOptions = RunOptions, Param = c(X1[r], X2[r]), FUN = RunModel_GR2M)
  })
}

objective <- function(Database, Parameters, WarmUp) {   # Optim_GR2MSemiDistr.R:186-212
  nDays <- sapply(as.Date(Database$DatesR), numberOfDays)
  Res <- simulate(Database, Parameters)
  if (nsub == 1) {
    qsink <- (area[1] * Res[[1]]$Qsim) / (86.4 * nDays)
  } else {
    Qlist <- lapply(1:nsub, function(w) (area[w] * Res[[w]]$Qsim) / (86.4 * nDays))
    qsink <- round(apply(do.call(cbind, Qlist), 1, FUN = sum), 2)
  }
  Qobs <- Database$Q; Qsim <- qsink
  if (WarmUp > 0) { Qobs <- Qobs[-WarmUp:-1]; Qsim <- Qsim[-WarmUp:-1] }
  res.df <- na.omit(data.frame(sim = Qsim, obs = Qobs))
  kge <- suppressWarnings(KGE(res.df$sim, res.df$obs)); nse <- suppressWarnings(NSE(res.df$sim, res.df$obs))
  rm_ <- suppressWarnings(rmse(res.df$sim, res.df$obs))
  list(sink = qsink, crit = c(kge, nse, rm_), of = c(1 - round(kge, 3), 1 - round(nse, 3), round(rm_, 3)))
}

full <- Data                                                        # Run example: 01/1981..12/2016
ntime <- nrow(full)
nDays <- sapply(as.Date(full$DatesR), numberOfDays)
nsim  <- min(nrow(Sets), as.integer(cfg("n_sim_sets")))
ends  <- NULL
for (k in seq_len(nsim)) {
  Res <- simulate(full, Sets[k, ])
  tag <- sprintf("ref_airgr_set%02d_", k - 1)
  wr(do.call(cbind, lapply(Res, function(o) o$Precip)), paste0(tag, "PR.csv"))
  wr(do.call(cbind, lapply(Res, function(o) o$AE)),     paste0(tag, "AE.csv"))
  wr(do.call(cbind, lapply(Res, function(o) o$Prod)),   paste0(tag, "SM.csv"))
  wr(do.call(cbind, lapply(Res, function(o) o$Qsim)),   paste0(tag, "RU.csv"))
  wr(do.call(cbind, lapply(seq_len(nsub), function(w) (area[w] * Res[[w]]$Qsim) / (86.4 * nDays))), paste0(tag, "QS.csv"))
  ends <- rbind(ends, data.frame(set = k - 1, subbasin = seq_len(nsub) - 1,
                                 Prod = sapply(Res, function(o) o$StateEnd$Store$Prod),
                                 Rout = sapply(Res, function(o) o$StateEnd$Store$Rout)))
}
wr(ends, "ref_airgr_endstate.csv")
wr(data.frame(ndays = nDays), "ref_ndays.csv")

# calibration period of the Optim example (Optim_GR2MSemiDistr.R:26-32): 01/1981..12/2002, WarmUp 36
i1 <- which(format(Data$DatesR, "%m/%Y") == cfg("optim_end"))
cal <- Data[1:i1, ]; warm <- as.integer(cfg("optim_warmup"))
ob <- NULL
for (k in seq_len(nrow(Sets))) {
  o <- objective(cal, Sets[k, ], warm)
  ob <- rbind(ob, data.frame(set = k - 1, KGE = o$crit[1], NSE = o$crit[2], RMSE = o$crit[3],
                             of_KGE = o$of[1], of_NSE = o$of[2], of_RMSE = o$of[3]))
  if (k <= nsim) wr(data.frame(sink = o$sink), sprintf("ref_sink_set%02d.csv", k - 1))
}
wr(ob, "ref_objective.csv")

# hydroGOF on degenerate series (which criteria are NA)
gofrow <- function(name, sim, obs) {
  d <- na.omit(data.frame(sim = sim, obs = obs))
  f <- function(expr) tryCatch(suppressWarnings(expr), error = function(e) NA_real_)
  data.frame(case = name, KGE = f(KGE(d$sim, d$obs)), NSE = f(NSE(d$sim, d$obs)), RMSE = f(rmse(d$sim, d$obs)))
}
obs5 <- c(3, 5, 4, 9, 1)
wr(rbind(gofrow("regular",        c(1, 2, 3, 4, 6), c(1, 3, 2, 5, 5)),
         gofrow("constant_sim",   rep(2, 5), obs5),
         gofrow("constant_obs",   obs5, rep(2, 5)),
         gofrow("zero_obs",       obs5, rep(0, 5)),
         gofrow("one_pair",       c(1, 2, 3), c(NA, 2.5, NA)),
         gofrow("two_pairs",      c(1, 2, 3), c(NA, 2.5, 2.0)),
         gofrow("no_pair",        c(1, 2), c(NA_real_, NA_real_)),
         gofrow("identical",      obs5, obs5)), "ref_gof_degenerate.csv")

# R's round() on the committed cases
rc <- utils::read.csv(file.path(inp, "round_cases.csv"))
wr(data.frame(x = rc$x, r1 = round(rc$x, 1), r2 = round(rc$x, 2), r3 = round(rc$x, 3)), "ref_round.csv")

ver <- data.frame(what = c("R", "airGR", "hydroGOF", "platform", "date"),
                  version = c(paste(R.version$major, R.version$minor, sep = "."), as.character(utils::packageVersion("airGR")),
                              as.character(utils::packageVersion("hydroGOF")), R.version$platform, format(Sys.time(), "%Y-%m-%d")))

# ---- tier B: the reference package itself ------------------------------------------------------------
if (requireNamespace("GR2MSemiDistr", quietly = TRUE)) {
  suppressPackageStartupMessages(library(GR2MSemiDistr))
  # Run_GR2MSemiDistr only touches Subbasins@data$Area, @data$Region and Subbasins$COMID (Run:68-71)
  methods::setClass("PinSubbasins", representation(data = "data.frame"))
  methods::setMethod("$", "PinSubbasins", function(x, name) x@data[[name]])
  roi <- methods::new("PinSubbasins", data = Sub)
  raw <- utils::read.csv(file.path(inp, "data.csv"), stringsAsFactors = FALSE)
  ans <- Run_GR2MSemiDistr(Data = raw, Subbasins = roi, RunIni = cfg("run_ini"), RunEnd = cfg("run_end"),
                           WarmUp = as.integer(cfg("run_warmup")), Parameters = Sets[1, ])
  for (k in c("QS", "RU", "PR", "AE", "SM")) wr(ans[[k]], paste0("ref_pkg_run_", k, ".csv"))
  wr(ans$SINK, "ref_pkg_run_SINK.csv")
  ver <- rbind(ver, data.frame(what = "GR2MSemiDistr", version = as.character(utils::packageVersion("GR2MSemiDistr"))))
} else message("tier B skipped: package GR2MSemiDistr is not installed")

# ---- tier C: TauDEM AreaD8 -wg -------------------------------------------------------------------------
exe <- Sys.which(c("aread8", "AreaD8")); exe <- exe[nzchar(exe)][1]
if (!is.na(exe) && nzchar(Sys.which("mpiexec")) && requireNamespace("raster", quietly = TRUE)) {
  fd  <- as.matrix(utils::read.csv(gzfile(file.path(inp, "flowdir.csv.gz")), header = FALSE))
  wg  <- utils::read.csv(file.path(inp, "route_weights.csv"))      # columns: cell (0-based row-major), m0, m1, ...
  tmp <- tempfile("pin_taudem"); dir.create(tmp); old <- setwd(tmp)
  mk <- function(m) raster::raster(m, xmn = 0, xmx = ncol(m), ymn = 0, ymx = nrow(m), crs = "+proj=longlat +datum=WGS84")
  p <- mk(fd); p[p < 1] <- NA
  raster::writeRaster(p, "FlowDir.tif", datatype = "INT2S", overwrite = TRUE)
  for (m in seq_len(ncol(wg) - 1)) {
    w <- mk(matrix(0, nrow(fd), ncol(fd)))
    w[wg$cell + 1] <- wg[[m + 1]]                                   # Routing_GR2MSemiDistr.R:109-113
    raster::writeRaster(w, "Weights.tif", overwrite = TRUE)        # default FLT4S, as Routing:114
    system(paste("mpiexec -n 8", exe, "-p FlowDir.tif -wg Weights.tif -ad8 FlowAcum.tif"))   # Routing:117
    a <- raster::as.matrix(raster::raster("FlowAcum.tif"))
    gz <- gzfile(file.path(out, sprintf("ref_ad8_month%02d.csv.gz", m - 1)), "w")
    txt <- matrix(ifelse(is.na(a), "NA", sprintf("%.9g", a)), nrow(a), ncol(a))
    utils::write.table(txt, gz, sep = ",", row.names = FALSE, col.names = FALSE, quote = FALSE)
    close(gz)
    file.remove("Weights.tif", "FlowAcum.tif")
  }
  setwd(old)
  ver <- rbind(ver, data.frame(what = "TauDEM", version = exe))
} else message("tier C skipped: aread8 / mpiexec / raster not available")

wr(ver, "ref_versions.csv")
message("reference goldens written to ", out)
