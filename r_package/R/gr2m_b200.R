# GR2MSemiDistr front end over libgr2m_b200.so.  Never run under R (no R in the build image): the .Call glue it uses is compiled and
# executed against a mock R API (tests/test_r_shim.py); the R code itself is checked there for routine names / argument counts.
# Exported signatures and return structures follow GR2MSemiDistr 4.2.1
# (R/Run_GR2MSemiDistr.R:48, R/Optim_GR2MSemiDistr.R:41, R/Routing_GR2MSemiDistr.R:35);
# the arithmetic is done by .Call into the C ABI, the wrappers keep only what must stay in R:
# date handling, the parameter table, R's own round(), list assembly, file output.

# flags for the C ABI: options(gr2m.perc_f32_exponent = TRUE) selects the percolation exponent gfortran
# gives airGR's literal (1./3.) -- see DESIGN.md section 2; default is the exact cube root.
.gr2m_flags <- function() if (isTRUE(getOption("gr2m.perc_f32_exponent"))) 1L else 0L

.gr2m_dates <- function(x) as.POSIXct(paste0(x, " 00:00:00"), "GMT",
                                      tryFormats = c("%Y-%m-%d", "%Y/%m/%d", "%d-%m-%Y", "%d/%m/%Y"))

.gr2m_setup <- function(Data, Subbasins, RunIni, RunEnd) {
  d     <- .gr2m_dates(Data$DatesR)
  key   <- format(d, "%m/%Y")
  rows  <- seq(which(key == RunIni), which(key == RunEnd))
  nsub  <- nrow(Subbasins@data)
  zone  <- sort(unique(Subbasins@data$Region))
  db    <- Data[rows, ]
  list(rows = rows, dates = as.Date(d[rows]), nsub = nsub, zone = zone,
       region_id = as.integer(match(Subbasins@data$Region, zone) - 1L),
       area = as.numeric(Subbasins@data$Area), comid = as.vector(Subbasins$COMID),
       P = as.matrix(db[, 1 + seq_len(nsub), drop = FALSE]) * 1.0,
       E = as.matrix(db[, 1 + nsub + seq_len(nsub), drop = FALSE]) * 1.0,
       ndays = as.integer(lubridate::days_in_month(as.Date(d[rows]))),
       Q = db$Q)
}

Run_GR2MSemiDistr <- function(Data, Subbasins, RunIni, RunEnd, WarmUp = NULL, Parameters,
                              IniState = NULL, Save = FALSE, Update = FALSE) {
  tictoc::tic()
  s <- .gr2m_setup(Data, Subbasins, RunIni, RunEnd)
  stopifnot(length(Parameters) == 4 * length(s$zone))
  S0 <- R0 <- NULL
  if (!is.null(IniState)) {
    S0 <- vapply(IniState, function(z) z$Store$Prod, 0)
    R0 <- vapply(IniState, function(z) z$Store$Rout, 0)
  }
  message(paste("Running GR2M model for", s$nsub, "subbasins"))
  o <- .Call(C_gr2m_run, s$P, s$E, s$area, s$ndays, s$region_id, length(s$zone),
             as.numeric(Parameters), S0, R0, .gr2m_flags())
  EndState <- lapply(seq_len(s$nsub), function(i)
    list(Store = list(Prod = o$S_end[i], Rout = o$R_end[i], Exp = NA), UH = list(UH1 = NA, UH2 = NA)))
  pr <- round(o$PR, 1); ae <- round(o$AE, 1); sm <- round(o$SM, 1); ru <- round(o$RU, 1)
  if (s$nsub == 1) {
    qs <- round((s$area[1] * ru) / (86.4 * s$ndays), 1); qt <- as.vector(qs)
  } else {
    qs <- round(o$QS, 1); qt <- round(rowSums(qs), 2)
  }
  drop <- if (is.null(WarmUp)) integer(0) else seq_len(WarmUp)
  keep <- setdiff(seq_along(s$rows), drop)
  pr <- pr[keep, , drop = s$nsub == 1]; ae <- ae[keep, , drop = s$nsub == 1]
  sm <- sm[keep, , drop = s$nsub == 1]; ru <- ru[keep, , drop = s$nsub == 1]
  qs <- qs[keep, , drop = s$nsub == 1]; qt <- qt[keep]; Dates <- s$dates[keep]
  Ans <- list(QS = qs, RU = ru, PR = pr, AE = ae, SM = sm, Dates = Dates, COMID = s$comid, EndState = EndState)
  if (!is.null(s$Q)) {
    sink <- data.frame(sim = round(qt, 3), obs = round(s$Q[keep], 3)); rownames(sink) <- Dates
    Ans <- c(list(SINK = sink), Ans)
  }
  if (Save) .gr2m_save(Ans, c("PR", "AE", "SM", "RU"), RunEnd, Update)
  message("Done!"); tictoc::toc()
  Ans
}

# Batched OFUN: `parameter_sets` is a matrix with one candidate per column (4*nreg rows).
gr2m_objective_batch <- function(handle, parameter_sets, Qobs, WarmUp = 0L,
                                 Optimization = c("KGE", "NSE", "RMSE")) {
  which <- match(match.arg(Optimization), c("KGE", "NSE", "RMSE")) - 1L
  .Call(C_gr2m_objective, handle, as.matrix(parameter_sets) * 1.0, as.numeric(Qobs),
        as.integer(WarmUp), which, .gr2m_flags())
}

# Month-invariant inputs of the routing seam (Routing:77-81, 119-121), shared by Routing_ and the routed calibration
.gr2m_routing_geometry <- function(Subbasins, Dem, FlowDir) {
  roi  <- sf::st_as_sf(Subbasins)
  cen  <- methods::as(sf::st_as_sf(geos::geos_point_on_surface(roi)), "Spatial")
  src  <- raster::cellFromXY(Dem, sp::coordinates(cen)) - 1L        # 0-based row-major
  cov  <- exactextractr::exact_extract(Dem, roi, include_cell = TRUE, progress = FALSE)
  inner <- lapply(cov, function(d) d$cell[d$coverage_fraction == 1] - 1L)
  if (is.null(FlowDir)) {           # no TauDEM: pit filling + D8 codes on the GPU (own flat rule, see DESIGN.md)
    fd <- .Call(C_wfa_flowdir, raster::as.matrix(Dem) * 1.0)
  } else {                          # FlowDir.tif of "D8Flowdir -p" as a RasterLayer
    fd <- raster::as.matrix(FlowDir); storage.mode(fd) <- "integer"
  }
  list(fd = fd, src = as.integer(src), off = as.integer(c(0, cumsum(lengths(inner)))), cells = as.integer(unlist(inner)),
       comid = as.vector(roi$COMID))
}

# Gauge / Dem / FlowDir are additions after the reference's arguments: with Gauge (a COMID) every candidate is
# simulated, routed over the DEM and scored on the routed discharge of that subbasin's polygon in one device call.
Optim_GR2MSemiDistr <- function(Data, Subbasins, RunIni, RunEnd, WarmUp = NULL, Parameters,
                                Parameters.Min, Parameters.Max, Max.Functions = 5000,
                                Optimization = "NSE", No.Optim = NULL, Gauge = NULL, Dem = NULL, FlowDir = NULL) {
  tictoc::tic()
  s <- .gr2m_setup(Data, Subbasins, RunIni, RunEnd)
  nreg  <- length(s$zone)
  vpar  <- rep(s$zone, 4)
  fixed <- if (is.null(No.Optim)) rep(FALSE, 4 * nreg) else vpar %in% No.Optim
  nopt  <- nreg - sum(s$zone %in% No.Optim)
  h <- .Call(C_gr2m_basin, s$P, s$E, s$area, s$ndays, s$region_id, nreg)
  warm <- if (is.null(WarmUp)) 0L else as.integer(WarmUp)
  routing <- NULL; gauge <- 0L
  if (!is.null(Gauge)) {
    g <- .gr2m_routing_geometry(Subbasins, Dem, FlowDir)
    routing <- .Call(C_wfa_router, g$fd, g$src, g$off, g$cells)
    gauge <- match(Gauge, g$comid) - 1L
  }
  which <- match(Optimization, c("KGE", "NSE", "RMSE")) - 1L
  OFUN <- function(parameter_set) {            # drop-in body: one candidate, forcing already on the GPU
    full <- as.numeric(Parameters); full[!fixed] <- parameter_set
    if (is.null(routing)) gr2m_objective_batch(h, matrix(full, ncol = 1), s$Q, warm, Optimization)
    else .Call(C_gr2m_objective_routed, h, routing, matrix(full, ncol = 1), as.numeric(s$Q), gauge, warm, which, .gr2m_flags())
  }
  message(paste("Optimizing", Optimization, "with SCE-UA"))
  cal <- rtop::sceua(OFUN, pars = Parameters[!fixed], lower = rep(Parameters.Min, each = nopt),
                     upper = rep(Parameters.Max, each = nopt), maxn = Max.Functions)
  fo <- if (Optimization == "RMSE") round(cal$value, 3) else round(1 - cal$value, 3)
  par <- as.numeric(Parameters); par[!fixed] <- round(cal$par, 3)
  Ans <- list(Param = par, Value = fo)
  print(paste0(rep(c("X1=", "X2=", "fpr=", "fpe="), each = nreg), Ans$Param))
  print(paste0(Optimization, "=", Ans$Value))
  message("Done!"); tictoc::toc()
  Ans
}

Routing_GR2MSemiDistr <- function(Model, Subbasins, Dem, AcumIni = NULL, AcumEnd = NULL,
                                  Save = FALSE, Update = FALSE, FlowDir = NULL) {
  tictoc::tic()
  QS <- as.matrix(Model$QS); Dates <- Model$Dates
  if (!(is.null(AcumIni) && is.null(AcumEnd)) && nrow(QS) > 1) {
    k <- format(as.Date(Dates), "%d/%m/%Y")
    i <- seq(which(k == paste0("01/", AcumIni)), which(k == paste0("01/", AcumEnd)))
    QS <- QS[i, , drop = FALSE]; Dates <- Dates[i]
  }
  g <- .gr2m_routing_geometry(Subbasins, Dem, FlowDir)
  message(paste0("Routing streamflows for ", ncol(QS), " sub-basins"))
  qr <- round(.Call(C_wfa_route, g$fd, QS * 1.0, g$src, g$off, g$cells, 0L), 1)
  Ans <- list(QR = qr, Dates = Dates, COMID = g$comid)
  if (Save) .gr2m_save(c(Ans, list(Dates = Dates)), "QR", format(utils::tail(Dates, 1), "%m/%Y"), Update)
  message("Done!"); tictoc::toc()
  Ans
}

# Create_Forcing_Inputs (R/Create_Forcing_Inputs.R:42): areal-mean P / PE table in airGR layout.  Host-side and once
# per study, so it stays in R on raster + exactextractr like the reference; differences: the whole brick goes through
# exact_extract in one call (no PSOCK cluster per variable, Create:101-108, 149-157, the first of which is never
# stopped), and the default Members = NULL works (the reference's `if (Members == 1 | is.null(Members))`, Create:82,
# is an error for NULL).  Same crop (extent * Buffer), same nearest-neighbour resample to `Resolution`, same
# round(, 1) / round(Q, 3), same column names.
.gr2m_areal_mean <- function(brick, roi, ind, Buffer, Resolution, Update) {
  x <- if (Update) brick[[raster::nlayers(brick)]] else brick[[ind]]
  buf <- raster::crop(x, raster::extent(roi) * Buffer)
  fine <- raster::raster(raster::extent(buf)); raster::crs(fine) <- raster::crs(buf); raster::res(fine) <- Resolution
  m <- exactextractr::exact_extract(raster::resample(buf, fine, method = "ngb"), sf::st_as_sf(roi), fun = "mean",
                                    progress = FALSE)
  round(t(as.matrix(m)), 1)                      # layers x subbasins
}

Create_Forcing_Inputs <- function(Subbasins, Precip, PotEvap, Qobs = NULL, DateIni, DateEnd,
                                  IniGrids = "01/1981", Save = FALSE, Update = FALSE, Resolution = 0.01,
                                  Buffer = 1.1, Members = NULL) {
  tictoc::tic()
  comid <- as.vector(Subbasins$COMID)
  first <- function(x) as.Date(paste0("01/", x), "%d/%m/%Y")
  index <- function(brick) {
    d <- seq(first(IniGrids), by = "month", length.out = raster::nlayers(brick))
    if (is.null(Members) || Members == 1) seq(which(d == first(DateIni)), which(d == first(DateEnd)))
    else seq_len(raster::nlayers(brick))
  }
  cols <- list()
  if (!is.null(Precip)) {
    message("Calculating monthly mean-areal precipitation [mm]")
    pr <- .gr2m_areal_mean(Precip, Subbasins, index(Precip), Buffer, Resolution, Update)
    colnames(pr) <- paste0("P_", comid); cols <- c(cols, list(pr))
  }
  if (!is.null(PotEvap)) {
    message("Calculating monthly mean-areal pot. evapotranspiration [mm]")
    pe <- .gr2m_areal_mean(PotEvap, Subbasins, index(PotEvap), Buffer, Resolution, Update)
    colnames(pe) <- paste0("E_", comid); cols <- c(cols, list(pe))
  }
  DatesMonths <- if (Update) first(DateEnd) else {
    d <- seq(first(DateIni), first(DateEnd), by = "month")
    if (is.null(Members)) d else rep(d, times = Members)
  }
  Ans <- data.frame(DatesR = DatesMonths, do.call(cbind, cols), check.names = FALSE)
  if (!is.null(Qobs)) Ans$Q <- round(Qobs[, 2], 3)
  if (Save) {
    dir.create("./Inputs", showWarnings = FALSE)
    utils::write.table(Ans, file = "./Inputs/Inputs_model.txt", sep = "\t", col.names = TRUE, row.names = FALSE)
  }
  message("Done!"); tictoc::toc()
  Ans
}

.gr2m_save <- function(Ans, keys, RunEnd, Update) {
  dir.create("./Outputs", recursive = TRUE, showWarnings = FALSE)
  for (k in keys) {
    new <- as.data.frame(Ans[[k]]); colnames(new) <- paste0(k, "_", Ans$COMID); rownames(new) <- Ans$Dates
    if (Update) {
      first <- as.Date(format(Sys.Date(), "%Y-%m-01"))      # month arithmetic without attaching lubridate
      m1 <- format(seq(first, by = "-2 month", length.out = 2)[2], "%Y%m")
      m2 <- format(seq(first, by = "-1 month", length.out = 2)[2], "%Y%m")
      f1 <- file.path("Outputs", paste0(k, "_GR2MSemiDistr_", m1, ".txt"))
      old <- utils::read.table(f1, header = TRUE, sep = "\t")
      all <- rbind(as.matrix(old), as.matrix(new)); rownames(all) <- c(rownames(old), rownames(new))
      utils::write.table(all, file.path("Outputs", paste0(k, "_GR2MSemiDistr_", m2, ".txt")), sep = "\t")
      file.remove(f1)
    } else {
      mn <- format(as.Date(paste0("01/", RunEnd), "%d/%m/%Y"), "%Y%m")
      utils::write.table(new, file.path("Outputs", paste0(k, "_GR2MSemiDistr_", mn, ".txt")), sep = "\t")
    }
  }
}
