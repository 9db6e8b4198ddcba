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
