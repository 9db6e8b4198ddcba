# Known answers the reference's own suite pins for this path (tests/testthat/test-kendall-tau.R
# and _snaps/kendall-tau.md), restated against the B200 back end.  Needs R + a B200; it has NOT
# been run in the build image (no R there) -- the same expectations are executed through the
# Python host layer in tests/test_gpu_parity.py.
test_that("basic ici_kt matches the asymptotic Kendall test", {
  x = seq(1, 10); y = seq(1, 10); y[2] = 15
  t1 = ici_kt(x, y, "global")
  c1 = cor.test(x, y, method = "kendall", exact = FALSE)
  expect_equal(t1[["tau"]], unname(c1$estimate))
  expect_equal(t1[["pvalue"]], c1$p.value, tolerance = 1.5e-8)
  expect_equal(t1[["completeness"]], 1)
  x[2] = NA
  expect_equal(ici_kt(x, y, "global")[["completeness"]], 0.9)
  y[3] = NA
  expect_equal(ici_kt(x, y, "global")[["completeness"]], 0.8)
})

test_that("degenerate inputs give NA with the reference's warnings", {
  expect_warning(r1 <- ici_kt(c(1, NA), c(NA, 2), "local"), "single value")
  expect_true(all(is.na(r1)))
  expect_warning(r2 <- ici_kt(rep(1, 10), seq(1, 10)), "single unique value")
  expect_true(all(is.na(r2)))
  expect_error(ici_kt(1:3, 1:4), "not the same length")
})

test_that("large vectors reproduce the reference snapshot", {
  set.seed(1234)
  x = rnorm(50000); y = rnorm(50000)
  v = ici_kt(x, y, perspective = "global")
  expect_equal(v[["tau"]], -0.00123518, tolerance = 1e-6)   # _snaps/kendall-tau.md:1-7
  expect_equal(v[["pvalue"]], 0.67867094, tolerance = 1e-7)
  expect_equal(v[["tau_max"]], 1)
  expect_equal(v[["completeness"]], 1)
})

test_that("matrix interface agrees with the pair interface", {
  set.seed(1234)
  x = matrix(rnorm(2000), 100, 20); x[sample(length(x), 300)] = NA
  colnames(x) = paste0("s", seq_len(ncol(x)))
  m = ici_kendalltau(x, global_na = c(NA), scale_max = FALSE, diag_good = FALSE)
  p = ici_kt(x[, 1], x[, 2], "global")
  expect_equal(m$raw[2, 1], p[["tau"]])
  expect_equal(m$pvalue[2, 1], p[["pvalue"]])
  expect_equal(dim(m$cor), c(20L, 20L))
  only = ici_kendalltau(x, include_only = "s1", global_na = c(NA))
  expect_equal(sum(only$cor == 0), 20 * 20 - (2 * 19 + 20))
})

test_that("device-filled matrices agree with the long format and the host scatter", {
  set.seed(42)
  x = matrix(rnorm(3000), 150, 20); x[sample(length(x), 400)] = NA
  x[, 7] = 1                                           # a constant column: NA entries and one warning class
  colnames(x) = paste0("s", seq_len(ncol(x)))
  suppressWarnings({
    m = ici_kendalltau(x, global_na = c(NA))           # C_icikt_matrices
    l = ici_kendalltau(x, global_na = c(NA), return_matrix = FALSE)$cor
  })
  for (what in c("cor", "raw", "pvalue", "taumax", "completeness")) {
    expect_identical(m[[what]][cbind(l$s1, l$s2)], l[[what]])
    expect_identical(m[[what]][cbind(l$s2, l$s1)], l[[what]])
  }
  expect_true(is.na(m$raw["s7", "s1"]) && !is.nan(m$raw["s7", "s1"]))   # NA_real_, not NaN
  expect_equal(unname(diag(m$taumax)), rep(1, 20))
})

test_that("pairwise_completeness matches the mask arithmetic of the reference", {
  set.seed(7)
  x = matrix(rpois(4000, 1), 200, 20); storage.mode(x) = "double"
  colnames(x) = paste0("s", seq_len(ncol(x)))
  excl = x == 0
  m = pairwise_completeness(x)
  expect_equal(m["s2", "s5"], 1 - sum(excl[, 2] | excl[, 5]) / 200)
  expect_equal(unname(diag(m)), unname(1 - colSums(excl) / 200))
  l = pairwise_completeness(x, include_only = "s3", return_matrix = FALSE)
  expect_equal(l$missingness, vapply(seq_len(nrow(l)), function(k) sum(excl[, l$s1[k]] | excl[, l$s2[k]]), numeric(1)))
})
