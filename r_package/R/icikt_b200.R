# Host layer of the B200 back end.  Function names, arguments, defaults, return values and
# warning/error texts follow the reference package (paths relative to its repository):
#   ici_kendalltau         R/kendalltau.R:96-179   (setup_comparisons :181-278, scale_and_reshape :357-421)
#   ici_kt                 R/RcppExports.R:62-64   -> src/kendallc.cpp:166-366
#   kt_fast                R/kendalltau.R:448-545  (kt_split :310-354)
#   pairwise_completeness  R/kendalltau.R:563-629
# What changed: the furrr pair loop and the per-pair .Call into Rcpp are ONE batched .Call into
# libicikt_b200.so.  Everything that is not pair arithmetic (argument checks, pair planning,
# scaling, matrix fill) stays in R.  There is no CPU fallback.

.warn_status = function(status) {
  # the reference warns once per pair from C++ (src/kendallc.cpp:225,238,292); batched here
  texts = c(`2` = "Warning: The vectors only have a single value, NA returned!",
            `3` = "Warning: Either 'X' or 'Y' have only a single unique value, NA returned!",
            `4` = "Warning: Ties equal the total, NA returned!")
  for (code in names(texts)) {
    n_hit = sum(status == as.integer(code))
    if (n_hit == 1) warning(texts[[code]], call. = FALSE)
    if (n_hit > 1) warning(sprintf("%s (%d pairs)", texts[[code]], n_hit), call. = FALSE)
  }
  invisible(NULL)
}

.as_numeric_matrix = function(x, arg) {
  if (is.null(colnames(x))) stop(sprintf("Colnames of `%s` must be be specified.", arg), call. = FALSE)
  if (is.data.frame(x)) {
    message(sprintf("`%s` is a data.frame, converting to matrix ...", arg))
    x = as.matrix(x)
  }
  if (!(is.double(x) || is.integer(x))) stop(sprintf("`%s` must be a numeric type.", arg), call. = FALSE)
  storage.mode(x) = "double"
  x
}

.missing_matrix = function(data_matrix, global_na) {
  # R/utils.R:1-23: NA and Inf entries of global_na select classes, the rest are literals
  out = matrix(FALSE, nrow(data_matrix), ncol(data_matrix), dimnames = dimnames(data_matrix))
  if (any(is.na(global_na))) out[is.na(data_matrix)] = TRUE
  if (any(is.infinite(global_na))) out[is.infinite(data_matrix)] = TRUE
  for (v in global_na[is.finite(global_na)]) out[which(data_matrix == v)] = TRUE
  out
}

# All pairs in utils::combn(n, 2) order as 1-based column indices, vectorised (utils::combn is an
# R-level loop over 12.5 M pairs at 5 000 samples); the diagonal is appended when !diag_good.
.all_pair_indices = function(n_sample, diag_good) {
  cnt = if (n_sample > 1) (n_sample - 1L):1L else integer(0)
  i = rep.int(seq_len(n_sample - 1L), cnt)
  j = i + sequence(cnt)
  if (!diag_good) { i = c(i, seq_len(n_sample)); j = c(j, seq_len(n_sample)) }
  list(i = as.integer(i), j = as.integer(j))
}

# Pair plan in utils::combn order (+ the diagonal when !diag_good) as 1-based column indices.
# `all_pairs` tells the caller that the list is exactly what icikt_all_pairs enumerates itself.
# need_indices = FALSE: the caller hands the all-pairs case to the library without ever looking at
# the index vectors (matrix outputs), so none are built -- only their number, `n_todo`.
.plan_pairs = function(samples, include_only = NULL, diag_good = TRUE, include_arg = "include_only",
                       need_indices = TRUE) {
  n_sample = length(samples)
  if (is.null(include_only) && !need_indices) {
    n_todo = n_sample * (n_sample - 1) / 2 + if (diag_good) 0 else n_sample
    if (n_todo == 0) {
      stop(sprintf("No comparisons to do. Check the list of column names in `%s` vs those in the samples.",
                   include_arg), call. = FALSE)
    }
    return(list(i = NULL, j = NULL, all_pairs = TRUE, n_todo = n_todo))
  }
  idx = .all_pair_indices(n_sample, diag_good)
  i = idx$i; j = idx$j
  all_pairs = TRUE
  if (!is.null(include_only)) {
    all_pairs = FALSE
    if (is.data.frame(include_only)) include_only = as.list(include_only)
    if (is.list(include_only)) {
      if (length(include_only) != 2) {
        stop(sprintf(paste0("`%s` must be a vector, a data.frame with two columns, or list of two vectors. ",
                            "Currently, `length(%s)` returns %d"), include_arg, include_arg, length(include_only)),
             call. = FALSE)
      }
      want = paste0(include_only[[1]], ".", include_only[[2]])   # paste0 recycles like the reference
      fwd = paste0(samples[i], ".", samples[j]); rev = paste0(samples[j], ".", samples[i])
      keep = (fwd %in% want) | (rev %in% want)
    } else {
      keep = (samples[i] %in% include_only) | (samples[j] %in% include_only)
    }
    i = i[keep]; j = j[keep]
  }
  if (length(i) == 0) {
    stop(sprintf("No comparisons to do. Check the list of column names in `%s` vs those in the samples.",
                 include_arg), call. = FALSE)
  }
  list(i = as.integer(i), j = as.integer(j), all_pairs = all_pairs, n_todo = length(i))
}

.run_pairs = function(data, global_na, plan, include_diag, perspective, alternative, continuity, device) {
  na_inf = any(is.infinite(global_na))
  literals = as.double(global_na[is.finite(global_na)])
  if (!any(is.na(global_na)) && length(global_na) > 0) {
    # the library always treats NaN as missing (the pair kernel needs a missing marker); a
    # global_na without NA on data that holds NA is outside what the reference tests
    if (anyNA(data)) warning("global_na has no NA but the data does: NA entries are still treated as missing")
  }
  if (plan$all_pairs) {
    .Call(C_icikt_all_pairs, data, literals, perspective, alternative, continuity, include_diag, na_inf,
          as.integer(device))
  } else {
    .Call(C_icikt_pair_list, data, literals, plan$i, plan$j, perspective, alternative, continuity, na_inf,
          as.integer(device))
  }
}

.fill_matrices = function(samples, i, j, columns) {
  lapply(columns, function(v) {
    m = matrix(0, length(samples), length(samples), dimnames = list(samples, samples))
    m[cbind(i, j)] = v
    m[cbind(j, i)] = v
    m
  })
}

icikt_device_count = function() .Call(C_icikt_device_count)

ici_kt = function(x, y, perspective = "local", alternative = "two.sided", continuity = FALSE,
                  output = "simple", device = 0L) {
  if (length(x) != length(y)) stop("'X' and 'Y' are not the same length!")   # src/kendallc.cpp:168-170
  out_names = c("tau", "pvalue", "tau_max", "completeness")
  if (length(x) == 0) return(stats::setNames(rep(NA_real_, 4), out_names))
  data = cbind(as.double(x), as.double(y))
  r = .Call(C_icikt_pair_list, data, double(0), 1L, 2L, perspective, alternative, continuity, FALSE,
            as.integer(device))
  .warn_status(r$status)
  res = stats::setNames(c(r$raw, r$pvalue, r$taumax, r$completeness), out_names)
  if (output != "simple") print(res)
  res
}

# `device`: one CUDA ordinal, or a vector of ordinals -- the pair order is then sliced over those
# GPUs inside the one library call (the role of the reference's furrr workers).
ici_kendalltau = function(data_matrix, global_na = c(NA, Inf, 0), perspective = "global", scale_max = TRUE,
                          diag_good = TRUE, include_only = NULL, alternative = "two.sided",
                          continuity = FALSE, check_timing = FALSE, return_matrix = TRUE, device = 0L) {
  arg = deparse(substitute(data_matrix)); include_arg = deparse(substitute(include_only))
  data_matrix = .as_numeric_matrix(data_matrix, arg)
  samples = colnames(data_matrix)
  exclude_loc = .missing_matrix(data_matrix, global_na)
  # the all-pairs matrix path never needs the 1-based index vectors (12.5 M pairs at 5 000 samples)
  # (several devices: all pairs only -- every GPU returns its own block of columns of each matrix)
  on_device = return_matrix && (length(device) == 1L || is.null(include_only))
  plan = .plan_pairs(samples, include_only, diag_good, include_arg, need_indices = !on_device || check_timing)
  n_todo = plan$n_todo

  if (check_timing) {   # R/kendalltau.R:141-148,633-669: time five random pairs, extrapolate
    pick = sample(n_todo, min(5L, n_todo))
    sub = list(i = plan$i[pick], j = plan$j[pick], all_pairs = FALSE)
    t0 = Sys.time()
    .run_pairs(data_matrix, global_na, sub, FALSE, perspective, alternative, continuity, device)
    t_total = as.numeric(difftime(Sys.time(), t0, units = "secs"))
    t_all = t_total / length(pick) * n_todo
    return(data.frame(which = c("n_tested", "n_todo", "time_tested", "time_single", "time_all",
                                "time_across_cores", "time_minutes", "time_hours", "time_days"),
                      value = c(length(pick), n_todo, t_total, t_total / length(pick), t_all, t_all,
                                t_all / 60, t_all / 3600, t_all / 216000)))
  }

  if (on_device) {
    # scale_and_reshape (R/kendalltau.R:357-421) on the device: the five matrices come back filled,
    # degenerate pairs already NA; only the dimnames and `keep` are added here
    t1 = Sys.time()
    r = .Call(C_icikt_matrices, data_matrix, as.double(global_na[is.finite(global_na)]),
              if (plan$all_pairs) NULL else plan$i, if (plan$all_pairs) NULL else plan$j,
              perspective, alternative, continuity, any(is.infinite(global_na)), as.integer(device),
              scale_max, diag_good, as.integer(colSums(!exclude_loc)))
    run_time = as.numeric(difftime(Sys.time(), t1, units = "secs"))
    .warn_status(rep(seq_along(r$status_counts) - 1L, times = r$status_counts))
    mats = lapply(r[c("cor", "raw", "pvalue", "taumax", "completeness")],
                  function(m) { dimnames(m) = list(samples, samples); m })
    return(c(mats, list(keep = t(!exclude_loc), run_time = run_time)))
  }

  t1 = Sys.time()
  r = .run_pairs(data_matrix, global_na, plan, !diag_good, perspective, alternative, continuity, device)
  run_time = as.numeric(difftime(Sys.time(), t1, units = "secs"))
  .warn_status(r$status)

  # R/kendalltau.R:368-372: scale by the largest tau_max over the pairs actually computed
  cor = if (scale_max) r$raw / r$max_taumax else r$raw
  all_cor = data.frame(s1 = samples[plan$i], s2 = samples[plan$j], core = 0, raw = r$raw, pvalue = r$pvalue,
                       taumax = r$taumax, completeness = r$completeness, cor = cor)
  i = plan$i; j = plan$j
  n_good = colSums(!exclude_loc)
  if (diag_good) {      # :374-386, appended after scaling and never scaled
    all_cor = rbind(all_cor, data.frame(s1 = samples, s2 = samples, core = 0, raw = n_good / max(n_good),
                                        pvalue = 0, taumax = 1, completeness = n_good / nrow(exclude_loc),
                                        cor = n_good / max(n_good)))
    i = c(i, seq_along(samples)); j = c(j, seq_along(samples))
  }
  rownames(all_cor) = NULL
  if (!return_matrix) return(list(cor = all_cor, run_time = run_time))
  mats = .fill_matrices(samples, i, j, all_cor[c("cor", "raw", "pvalue", "taumax", "completeness")])
  c(mats, list(keep = t(!exclude_loc), run_time = run_time))
}

kt_fast = function(x, y = NULL, use = "everything", alternative = "two.sided", continuity = FALSE,
                   return_matrix = TRUE, device = 0L) {
  na_method = match.arg(use, c("all.obs", "complete.obs", "pairwise.complete.obs", "everything", "na.or.complete"))
  if (na_method == "na.or.complete") {
    stop("'na.or.complete' is not a supported value for `use`. Please use one of all.obs complete.obs pairwise.complete everthing.")
  }
  if (is.null(y)) {
    if (is.null(dim(x))) stop("`x` and `y` should both be provided as vectors, or `x` should be matrix-like. `x` is a single vector, and `y` is `NULL`.")
    data = .as_numeric_matrix(x, deparse(substitute(x)))
  } else {
    if (!is.null(dim(x)) || !is.null(dim(y))) stop("Both `x` and `y` must be vectors.")
    data = cbind(x = as.double(x), y = as.double(y))
  }
  samples = colnames(data)
  plan = .plan_pairs(samples, NULL, diag_good = FALSE)          # (i,i) pairs are computed, :479
  n_pair = length(plan$i)
  tau = rep(NA_real_, n_pair); pvalue = rep(NA_real_, n_pair)
  any_na = anyNA(data)
  do_it = !(na_method %in% c("everything", "all.obs") && any_na)
  if (na_method == "complete.obs") {
    ok_rows = rowSums(is.na(data)) == 0
    if (!any(ok_rows)) do_it = FALSE else data = data[ok_rows, , drop = FALSE]
  }
  t1 = Sys.time()
  if (do_it) {
    # kt_split calls ici_kt(tmp_x, tmp_y) with its defaults (:341): local, two.sided, no continuity
    if (na_method == "pairwise.complete.obs" && anyNA(data)) {
      # rows missing in either column are dropped per pair (:323-331) -- on the device
      # (complete-observations mode); status 9 marks the pairs the device cannot take
      # (a column whose missing rows tie with its minimum in fp64), those are filtered here
      r = .Call(C_icikt_all_pairs, data, double(0), "complete", "two.sided", FALSE, TRUE, FALSE, as.integer(device))
      tau = r$raw; pvalue = r$pvalue
      redo = which(r$status == 9L)
      .warn_status(r$status[r$status != 9L])
      for (k in redo) {
        good = !is.na(data[, plan$i[k]]) & !is.na(data[, plan$j[k]])
        tau[k] = NA_real_; pvalue[k] = NA_real_
        if (!any(good)) next
        r2 = .Call(C_icikt_pair_list, cbind(data[good, plan$i[k]], data[good, plan$j[k]]), double(0), 1L, 2L,
                   "local", "two.sided", FALSE, FALSE, as.integer(device))
        .warn_status(r2$status); tau[k] = r2$raw; pvalue[k] = r2$pvalue
      }
    } else {
      r = .Call(C_icikt_all_pairs, data, double(0), "local", "two.sided", FALSE, TRUE, FALSE, as.integer(device))
      .warn_status(r$status); tau = r$raw; pvalue = r$pvalue
    }
  }
  run_time = as.numeric(difftime(Sys.time(), t1, units = "secs"))
  if (!return_matrix) {
    return(list(tau = data.frame(s1 = samples[plan$i], s2 = samples[plan$j], tau = tau, pvalue = pvalue),
                run_time = run_time))
  }
  c(.fill_matrices(samples, plan$i, plan$j, list(tau = tau, pvalue = pvalue)), list(run_time = run_time))
}

pairwise_completeness = function(data_matrix, global_na = c(NA, Inf, 0), include_only = NULL,
                                 return_matrix = TRUE, device = 0L) {
  # 1 - (rows missing in either sample) / n, incl. (i,i) (R/kendalltau.R:563-629): missing-row bit
  # masks and popc(x | y) on the device, no pair kernel
  data_matrix = .as_numeric_matrix(data_matrix, deparse(substitute(data_matrix)))
  samples = colnames(data_matrix)
  plan = .plan_pairs(samples, include_only, diag_good = FALSE, deparse(substitute(include_only)),
                     need_indices = !return_matrix)
  r = .Call(C_icikt_pairwise_completeness, data_matrix, as.double(global_na),
            if (plan$all_pairs) NULL else plan$i, if (plan$all_pairs) NULL else plan$j,
            as.integer(device), return_matrix && plan$all_pairs)
  if (!return_matrix) {
    return(data.frame(s1 = samples[plan$i], s2 = samples[plan$j], core = 0,
                      missingness = as.numeric(r$missingness), completeness = r$completeness))
  }
  if (plan$all_pairs) {
    m = r$matrix
    dimnames(m) = list(samples, samples)
    return(m)
  }
  .fill_matrices(samples, plan$i, plan$j, list(completeness = r$completeness))$completeness
}

.onUnload = function(libpath) {
  .Call(C_icikt_release)
  library.dynam.unload("ICIKendallTauB200", libpath)
}

# Result formats either side of the path (reference: R/reshaping.R:16-68), same names and columns.
cor_matrix_2_long_df = function(in_matrix) {
  if (is.null(rownames(in_matrix)) || is.null(colnames(in_matrix))) stop("`in_matrix` needs row and column names")
  data.frame(s1 = rep(rownames(in_matrix), times = ncol(in_matrix)),
             s2 = rep(colnames(in_matrix), each = nrow(in_matrix)),
             cor = as.vector(in_matrix), stringsAsFactors = FALSE)
}

long_df_2_cor_matrix = function(long_df, is_square = TRUE) {
  if (!all(c("s1", "s2", "cor") %in% names(long_df))) {
    stop("The data.frame must contain the names 's1', 's2', and 'cor'.")
  }
  s1 = as.character(long_df[["s1"]]); s2 = as.character(long_df[["s2"]])
  if (is_square) { rows = cols = sort(unique(c(s1, s2))) } else { rows = sort(unique(s1)); cols = sort(unique(s2)) }
  out = matrix(NA_real_, length(rows), length(cols), dimnames = list(rows, cols))
  out[cbind(s1, s2)] = long_df[["cor"]]
  if (is_square && nrow(long_df) != length(out)) out[cbind(s2, s1)] = long_df[["cor"]]
  out
}
