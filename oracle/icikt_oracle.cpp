// oracle/icikt_oracle.cpp
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of the reference's ICI-Kendall-tau pair kernel.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product path (icikendalltau_b200/) never does.
//
// The reference itself (/root/reference/src/kendallc.cpp) cannot be compiled here:
// it includes <Rcpp.h> and calls R's nmath pnorm, and neither R nor Rcpp exist in
// this image.  Every function below therefore restates the reference's algorithm
// step by step and cites the reference lines it follows (paths relative to
// /root/reference).  Third-party arithmetic that is not vendored in the reference:
//   * R nmath pnorm5/pnorm_both (R >= 3.5, version unpinned by DESCRIPTION:50) --
//     restated below from the published Cody (1969) ANORM algorithm that R uses.
//   * Rcpp sugar (unique/duplicated/table/cumsum/diff/sum/min; version unpinned,
//     DESCRIPTION:25-26) -- restated with libstdc++ containers; its IntegerVector
//     arithmetic is int32 and is reproduced by the `emulate_int32` switch.
//
// Parity pinning: tests/test_oracle_golden.py checks this file against the
// reference's own snapshot values (tests/testthat/_snaps/kendall-tau.md) using an
// emulation of R's Mersenne-Twister/inversion RNG, against the deterministic
// known answers of tests/testthat/test-kendall-tau.R:5-59, and against
// scipy.stats.kendalltau as an independent tau/p check.
//
// Build: see oracle/Makefile (g++ -O2, R's default optimisation level).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <numeric>
#include <thread>
#include <unordered_set>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------
// R nmath pnorm_both / pnorm5 (third party; restated from Cody's algorithm as used
// by R's src/nmath/pnorm.c).  Reached from src/kendallc.cpp:324-330 through Rcpp
// sugar pnorm(z, 0, 1[, lower, log]).
// ---------------------------------------------------------------------------------
const double kSqrt32 = 5.656854249492380195206754896838;      // M_SQRT_32
const double k1Sqrt2Pi = 0.398942280401432677939946059934;    // M_1_SQRT_2PI

void pnorm_both(double x, double* cum, double* ccum, int i_tail) {
  static const double a[5] = {2.2352520354606839287, 161.02823106855587881,
                              1067.6894854603709582, 18154.981253343561249,
                              0.065682337918207449113};
  static const double b[4] = {47.20258190468824187, 976.09855173777669322,
                              10260.932208618978205, 45507.789335026729956};
  static const double c[9] = {0.39894151208813466764, 8.8831497943883759412,
                              93.506656132177855979,  597.27027639480026226,
                              2494.5375852903726711,  6848.1904505362823326,
                              11602.651437647350124,  9842.7148383839780218,
                              1.0765576773720192317e-8};
  static const double d[8] = {22.266688044328115691, 235.38790178262499861,
                              1519.377599407554805,  6485.558298266760755,
                              18615.571640885098091, 34900.952721145977266,
                              38912.003286093271411, 19685.429676859990727};
  static const double p[6] = {0.21589853405795699,     0.1274011611602473639,
                              0.022235277870649807,    0.001421619193227893466,
                              2.9112874951168792e-5,   0.02307344176494017303};
  static const double q[5] = {1.28426009614491121,    0.468238212480865118,
                              0.0659881378689285515,  0.00378239633202758244,
                              7.29751555083966205e-5};
  double xden, xnum, temp, del, eps, xsq, y;
  int i;
  const bool lower = i_tail != 1, upper = i_tail != 0;

  if (std::isnan(x)) { *cum = *ccum = NAN; return; }
  eps = 2.220446049250313e-16 * 0.5;
  y = std::fabs(x);
  if (y <= 0.67448975) {
    if (y > eps) {
      xsq = x * x;
      xnum = a[4] * xsq;
      xden = xsq;
      for (i = 0; i < 3; ++i) { xnum = (xnum + a[i]) * xsq; xden = (xden + b[i]) * xsq; }
    } else {
      xnum = xden = 0.0;
    }
    temp = x * (xnum + a[3]) / (xden + b[3]);
    if (lower) *cum = 0.5 + temp;
    if (upper) *ccum = 0.5 - temp;
  } else if (y <= kSqrt32) {
    xnum = c[8] * y;
    xden = y;
    for (i = 0; i < 7; ++i) { xnum = (xnum + c[i]) * y; xden = (xden + d[i]) * y; }
    temp = (xnum + c[7]) / (xden + d[7]);
    xsq = std::trunc(y * 16) / 16;
    del = (y - xsq) * (y + xsq);
    *cum = std::exp(-xsq * xsq * 0.5) * std::exp(-del * 0.5) * temp;
    *ccum = 1.0 - *cum;
    if (x > 0.) { temp = *cum; if (lower) *cum = *ccum; *ccum = temp; }
  } else if ((lower && -37.5193 < x && x < 8.2924) || (upper && -8.2924 < x && x < 37.5193)) {
    xsq = 1.0 / (x * x);
    xnum = p[5] * xsq;
    xden = xsq;
    for (i = 0; i < 4; ++i) { xnum = (xnum + p[i]) * xsq; xden = (xden + q[i]) * xsq; }
    temp = xsq * (xnum + p[4]) / (xden + q[4]);
    temp = (k1Sqrt2Pi - temp) / y;
    xsq = std::trunc(x * 16) / 16;
    del = (x - xsq) * (x + xsq);
    *cum = std::exp(-xsq * xsq * 0.5) * std::exp(-del * 0.5) * temp;
    *ccum = 1.0 - *cum;
    if (x > 0.) { temp = *cum; if (lower) *cum = *ccum; *ccum = temp; }
  } else {
    if (x > 0) { *cum = 1.; *ccum = 0.; } else { *cum = 0.; *ccum = 1.; }
  }
}

// pnorm5(x, 0, 1, lower_tail, log_p = FALSE)
double pnorm_std(double x, bool lower_tail) {
  if (std::isnan(x)) return NAN;
  if (std::isinf(x)) return (x < 0) == lower_tail ? 0.0 : 1.0;
  double p = 0, cp = 0;
  pnorm_both(x, &p, &cp, lower_tail ? 0 : 1);
  return lower_tail ? p : cp;
}

// ---------------------------------------------------------------------------------
// Helpers of src/kendallc.cpp:5-129
// ---------------------------------------------------------------------------------

// sortedIndex, src/kendallc.cpp:6-12 : stable argsort with comparator x[i] < x[j]
std::vector<int> sortedIndex(const std::vector<double>& x) {
  std::vector<int> idx(x.size());
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int i, int j) { return x[i] < x[j]; });
  return idx;
}

// compare_self, src/kendallc.cpp:15-31
std::vector<int> compare_self(const std::vector<double>& x) {
  const int n = (int)x.size();
  std::vector<int> m(n);
  m[0] = 1;
  for (int i = 1; i < n; i++) m[i] = (x[i] != x[i - 1]) ? 1 : 0;
  return m;
}

// compare_both, src/kendallc.cpp:34-51 (note the trailing 1 that is pushed back)
std::vector<int> compare_both(const std::vector<int>& x, const std::vector<int>& y) {
  const int n = (int)x.size();
  std::vector<int> m(n);
  m[0] = 1;
  for (int i = 1; i < n; i++) m[i] = ((x[i] != x[i - 1]) || (y[i] != y[i - 1])) ? 1 : 0;
  m.push_back(1);
  return m;
}

// which_notzero, src/kendallc.cpp:54-67
std::vector<int> which_notzero(const std::vector<int>& x) {
  std::vector<int> nz;
  nz.reserve(x.size());
  for (int i = 0; i < (int)x.size(); i++)
    if (x[i] != 0) nz.push_back(i);
  return nz;
}

// kendall_discordant, src/kendallc.cpp:70-100.  The reference accumulates into
// `int dis` (int32) and returns int; `wrap32` reproduces that, otherwise int64.
int64_t kendall_discordant(const std::vector<int>& x, const std::vector<int>& y, bool wrap32) {
  const int sup = 1 + *std::max_element(y.begin(), y.end());
  std::vector<int> arr(sup, 0);
  int64_t i = 0, k = 0;
  const int64_t n = (int64_t)x.size();
  int idx = 0;
  int64_t dis = 0;
  while (i < n) {
    while ((k < n) && (x[i] == x[k])) {
      dis = dis + i;
      idx = y[k];
      while (idx != 0) {
        dis = dis - arr[idx];
        idx = idx & (idx - 1);
      }
      if (wrap32) dis = (int64_t)(int32_t)(uint32_t)(uint64_t)dis;
      k++;
    }
    while (i < k) {
      idx = y[i];
      while (idx < sup) {
        arr[idx] = arr[idx] + 1;
        idx = idx + (idx & (-1 * idx));
      }
      i++;
    }
  }
  return dis;
}

inline int64_t wrap_i32(int64_t v) { return (int64_t)(int32_t)(uint32_t)(uint64_t)v; }

// count_rank_tie, src/kendallc.cpp:103-118: duplicated() + table() + three sums.
// Rcpp evaluates the products and sums on IntegerVector (int32); wrap32 reproduces
// the two's-complement wrap, otherwise the sums are exact int64.
void count_rank_tie(const std::vector<int>& ranks, bool wrap32, double out[3]) {
  std::unordered_set<int> seen;
  std::map<int, int> tab;  // table() is ordered
  for (int r : ranks) {
    if (!seen.insert(r).second) tab[r] += 1;  // ranks[duplicated(ranks)] -> table
  }
  int64_t s0 = 0, s1 = 0, s2 = 0;
  for (auto& kv : tab) {
    const int64_t t = (int64_t)kv.second + 1;  // number_tied = table(ranks2) + 1
    if (wrap32) {
      const int64_t a = wrap_i32(t * (t - 1));
      s0 = wrap_i32(s0 + a);
      s1 = wrap_i32(s1 + wrap_i32(a * (t - 2)));
      s2 = wrap_i32(s2 + wrap_i32(a * wrap_i32(2 * t + 5)));
    } else {
      s0 += t * (t - 1);
      s1 += t * (t - 1) * (t - 2);
      s2 += t * (t - 1) * (2 * t + 5);
    }
  }
  out[0] = (double)(s0 / 2);  // integer division as in sum(...) / 2 on int
  out[1] = (double)(s1 / 2);
  out[2] = (double)s2;
}

inline double signC(double x) { return x > 0 ? 1.0 : (x == 0 ? 0.0 : -1.0); }  // :121-129

}  // namespace

extern "C" {

// Result of one pair.  status: 0 ok, 1 all-NA (silent NA, :190-199), 2 n<2
// (:224-231), 3 single unique value (:234-244), 4 ties == total (:291-298).
struct icikt_oracle_result {
  double tau, pvalue, tau_max, completeness;
  int64_t dis, ntie, xtie, ytie, tot, n_entry, x0, y0, x1, y1, n_matching_na;
  double z, var;
  int32_t status;
};

// ici_kt, src/kendallc.cpp:166-366.
// perspective: 1 = "local", anything else behaves as global (:180).
// alternative: 0 two.sided, 1 less, 2 greater, other -> p stays 0 (:323-332).
// emulate_int32: reproduce the Rcpp int32 arithmetic (forensic only).
// Returns 0, or -1 for the length-mismatch stop() (:168-170) which the caller
// signals by passing nx != ny.
int icikt_oracle_ici_kt(const double* xin, int64_t nx, const double* yin, int64_t ny,
                        int perspective, int alternative, int continuity, int emulate_int32,
                        icikt_oracle_result* r) {
  const double NA = NAN;
  std::memset(r, 0, sizeof(*r));
  r->tau = r->pvalue = r->tau_max = r->completeness = NA;
  r->z = r->var = NA;
  if (nx != ny) return -1;
  const bool wrap32 = emulate_int32 != 0;

  std::vector<double> x(xin, xin + nx), y(yin, yin + ny);
  // diagnostic only (not part of the reference's outputs): rows missing in both vectors
  for (size_t i = 0; i < x.size(); i++) r->n_matching_na += (std::isnan(x[i]) && std::isnan(y[i]));
  if (perspective == 1) {  // :180-185
    std::vector<double> fx, fy;
    fx.reserve(x.size());
    fy.reserve(y.size());
    for (size_t i = 0; i < x.size(); i++) {
      if (!(std::isnan(x[i]) && std::isnan(y[i]))) { fx.push_back(x[i]); fy.push_back(y[i]); }
    }
    x.swap(fx);
    y.swap(fy);
  }
  std::vector<double> x2(x), y2(y);  // :187-188

  int64_t n_na_x = 0, n_na_y = 0;  // :190-191
  for (double v : x) n_na_x += std::isnan(v);
  for (double v : y) n_na_y += std::isnan(v);
  if ((n_na_x == (int64_t)x.size()) || (n_na_y == (int64_t)y.size())) {  // :193-199
    r->status = 1;
    return 0;
  }

  // completeness :205-212
  int64_t missingness = 0;
  for (size_t i = 0; i < x.size(); i++) missingness += (std::isnan(x[i]) || std::isnan(y[i]));
  const long double either_na_length = (long double)x.size();
  const long double completeness = 1 - (missingness / either_na_length);

  // :214-219  NA -> (min of the non-missing values of that vector) - 0.1
  double min_x = INFINITY, min_y = INFINITY;
  for (double v : x2) if (!std::isnan(v) && v < min_x) min_x = v;
  for (double v : y2) if (!std::isnan(v) && v < min_y) min_y = v;
  min_x -= 0.1;
  min_y -= 0.1;
  for (size_t i = 0; i < x2.size(); i++) if (std::isnan(x[i])) x2[i] = min_x;
  for (size_t i = 0; i < y2.size(); i++) if (std::isnan(y[i])) y2[i] = min_y;

  const int64_t n_entry = (int64_t)x2.size();  // :221
  r->n_entry = n_entry;
  if (n_entry < 2) {  // :224-231
    r->status = 2;
    return 0;
  }
  {  // :234-244 unique() == 1
    bool ux = false, uy = false;
    for (int64_t i = 1; i < n_entry; i++) { ux |= (x2[i] != x2[0]); uy |= (y2[i] != y2[0]); }
    if (!ux || !uy) { r->status = 3; return 0; }
  }

  // :247-251
  std::vector<int> perm_y = sortedIndex(y2);
  {
    std::vector<double> tx(n_entry), ty(n_entry);
    for (int64_t i = 0; i < n_entry; i++) { tx[i] = x2[perm_y[i]]; ty[i] = y2[perm_y[i]]; }
    x2.swap(tx);
    y2.swap(ty);
  }
  std::vector<int> y3 = compare_self(y2);
  std::vector<int> y4(n_entry);
  std::partial_sum(y3.begin(), y3.end(), y4.begin());

  // :254-258
  std::vector<int> perm_x = sortedIndex(x2);
  {
    std::vector<double> tx(n_entry);
    std::vector<int> ty(n_entry);
    for (int64_t i = 0; i < n_entry; i++) { tx[i] = x2[perm_x[i]]; ty[i] = y4[perm_x[i]]; }
    x2.swap(tx);
    y4.swap(ty);
  }
  std::vector<int> x3 = compare_self(x2);
  std::vector<int> x4(n_entry);
  std::partial_sum(x3.begin(), x3.end(), x4.begin());

  // :261-267
  std::vector<int> obs = compare_both(x4, y4);
  std::vector<int> nz = which_notzero(obs);
  int64_t dis = kendall_discordant(x4, y4, wrap32);
  int64_t ntie_i = 0;
  for (size_t i = 1; i < nz.size(); i++) {
    const int64_t cnt = nz[i] - nz[i - 1];
    if (wrap32) ntie_i = wrap_i32(ntie_i + wrap_i32(cnt * (cnt - 1)) / 2);
    else ntie_i += (cnt * (cnt - 1)) / 2;
  }
  const long double ntie = (long double)ntie_i;

  // :270-278
  double xc[3], yc[3];
  count_rank_tie(x4, wrap32, xc);
  count_rank_tie(y4, wrap32, yc);
  const double xtie = xc[0], x0 = xc[1], x1 = xc[2];
  const double ytie = yc[0], y0 = yc[1], y1 = yc[2];

  const int64_t tot = (n_entry * (n_entry - 1)) / 2;  // :280
  r->dis = dis; r->ntie = ntie_i; r->xtie = (int64_t)xtie; r->ytie = (int64_t)ytie;
  r->tot = tot; r->x0 = (int64_t)x0; r->y0 = (int64_t)y0; r->x1 = (int64_t)x1; r->y1 = (int64_t)y1;

  if ((xtie == tot) || (ytie == tot)) {  // :291-298
    r->status = 4;
    return 0;
  }

  // :300-308 (long double = x87 80-bit here exactly as in the reference build)
  long double con_minus_dis = tot - xtie - ytie + ntie - 2 * dis;
  long double tau = con_minus_dis / std::sqrt((tot - xtie) * (tot - ytie));
  long double con_plus_dis = tot - xtie - ytie + ntie;
  long double tau_max = con_plus_dis / std::sqrt((tot - xtie) * (tot - ytie));
  if (tau > 1) tau = 1;
  else if (tau < -1) tau = -1;

  // :310-321
  const int64_t m = n_entry * (n_entry - 1);
  long double var = ((m * (2 * n_entry + 5) - x1 - y1) / 18 + (2 * xtie * ytie) / m +
                     x0 * y0 / (9 * m * (n_entry - 2)));
  long double s_adjusted = tau * std::sqrt(((m / 2) - xtie) * ((m / 2) - ytie));
  if (continuity) {
    long double adj_s2 = signC((double)s_adjusted) * (std::abs(s_adjusted) - 1);
    s_adjusted = adj_s2;
  }
  const double z_b = (double)(s_adjusted / std::sqrt(var));

  // :323-332
  double pval = 0.0;
  if (alternative == 1) {
    pval = pnorm_std(z_b, true);
  } else if (alternative == 2) {
    pval = pnorm_std(z_b, false);
  } else if (alternative == 0) {
    const double p0 = pnorm_std(z_b, true), p1 = pnorm_std(z_b, false);
    // Rcpp sugar min(): returns NA/NaN as soon as one is met
    double mn = p0;
    if (!std::isnan(mn)) { if (std::isnan(p1)) mn = p1; else if (p1 < mn) mn = p1; }
    pval = 2 * mn;
  }
  r->tau = (double)tau;
  r->pvalue = pval;
  r->tau_max = (double)tau_max;
  r->completeness = (double)completeness;
  r->z = z_b;
  r->var = (double)var;
  r->status = 0;
  return 0;
}

// ici_kt_pairs, src/kendallc.cpp:370-549: the O(n^2) cross-check the reference's own
// tests use (tests/testthat/test-kendall-tau.R:34-40,80-89).  out = {tau, pvalue}.
// Differences from ici_kt kept on purpose: joint minimum for the NA value (:431-437),
// continuity correction always applied (:504), un-halved t0 term (:498).
int icikt_oracle_ici_kt_pairs(const double* xin, int64_t nx, const double* yin, int64_t ny,
                              int perspective, int alternative, double out[2]) {
  out[0] = out[1] = 0.0;
  if (nx != ny) return -1;
  std::vector<double> x(xin, xin + nx), y(yin, yin + ny);
  if (perspective == 1) {
    std::vector<double> fx, fy;
    for (size_t i = 0; i < x.size(); i++)
      if (!(std::isnan(x[i]) && std::isnan(y[i]))) { fx.push_back(x[i]); fy.push_back(y[i]); }
    x.swap(fx);
    y.swap(fy);
  }
  int64_t n_na_x = 0, n_na_y = 0;
  for (double v : x) n_na_x += std::isnan(v);
  for (double v : y) n_na_y += std::isnan(v);
  if ((n_na_x == (int64_t)x.size()) || (n_na_y == (int64_t)y.size())) return 0;  // returns 0.0
  double mn = INFINITY;
  for (double v : x) if (!std::isnan(v) && v < mn) mn = v;
  for (double v : y) if (!std::isnan(v) && v < mn) mn = v;
  const double na_value = mn - 0.1;
  std::vector<double> x2(x), y2(y);
  for (auto& v : x2) if (std::isnan(v)) v = na_value;
  for (auto& v : y2) if (std::isnan(v)) v = na_value;
  const double n_entry = (double)x2.size();
  if (n_entry < 2) return 0;
  double sum_concordant = 0, sum_discordant = 0;
  const int64_t n = (int64_t)x2.size();
  for (int64_t i = 0; i < n - 1; i++)
    for (int64_t j = i + 1; j < n; j++) {
      const double s = signC(x2[i] - x2[j]) * signC(y2[i] - y2[j]);
      sum_concordant += s > 0;
      sum_discordant += s < 0;
    }
  const double k_numerator = sum_concordant - sum_discordant;
  auto tied = [](const std::vector<double>& v) {
    std::map<double, int> cnt;
    for (double d : v) cnt[d] += 1;
    std::vector<double> t;
    for (auto& kv : cnt) if (kv.second > 1) t.push_back((double)kv.second);
    return t;
  };
  std::vector<double> t1 = tied(x2), t2 = tied(y2);
  auto sum_f = [](const std::vector<double>& t, int which) {
    double s = 0;
    for (double v : t) s += which == 0 ? v * (v - 1) : which == 1 ? v * (v - 1) * (2 * v + 5) : v * (v - 1) * (v - 2);
    return s;
  };
  const double t_0 = n_entry * (n_entry - 1) / 2;
  const double x_tied_sum_t1 = sum_f(t1, 0) / 2, y_tied_sum_t2 = sum_f(t2, 0) / 2;
  const double k_denominator = std::sqrt((t_0 - x_tied_sum_t1) * (t_0 - y_tied_sum_t2));
  const double k_tau = k_denominator == 0 ? 0 : k_numerator / k_denominator;
  const double s_adjusted = k_tau * std::sqrt((t_0 - x_tied_sum_t1) * (t_0 - y_tied_sum_t2));
  const double v_0_sum = n_entry * (n_entry - 1) * (2 * n_entry + 5);
  const double v_t_sum = sum_f(t1, 1), v_u_sum = sum_f(t2, 1);
  const double v_t1_sum = sum_f(t1, 0) * sum_f(t2, 0);
  const double v_t2_sum = sum_f(t1, 2) * sum_f(t2, 2);
  const double s_var = (v_0_sum - v_t_sum - v_u_sum) / 18 + v_t1_sum / (2 * n_entry * (n_entry - 1)) +
                       v_t2_sum / (9 * n_entry * (n_entry - 1) * (n_entry - 2));
  const double s_adjusted2 = signC(s_adjusted) * (std::fabs(s_adjusted) - 1);
  const double z_b = s_adjusted2 / std::sqrt(s_var);
  double pval = 0;
  if (alternative == 1) pval = pnorm_std(z_b, true);
  else if (alternative == 2) pval = pnorm_std(z_b, false);
  else if (alternative == 0) {
    const double p0 = pnorm_std(z_b, true), p1 = pnorm_std(z_b, false);
    double m2 = p0;
    if (!std::isnan(m2)) { if (std::isnan(p1)) m2 = p1; else if (p1 < m2) m2 = p1; }
    pval = 2 * m2;
  }
  out[0] = k_tau;
  out[1] = pval;
  return 0;
}

double icikt_oracle_pnorm(double x, int lower_tail) { return pnorm_std(x, lower_tail != 0); }

// The reference's pair loop, R/kendalltau.R:280-308 (ici_split) run by
// computation$split_fun over contiguous chunks of ceil(P/ncore) pairs
// (R/kendalltau.R:250-253).  data is the column-major n x C `exclude_data`
// (missing already marked NaN, R/kendalltau.R:119-121).  One std::thread per chunk
// stands in for one furrr worker.  pi/pj are 0-based column indices.  Outputs have
// length P; counts (7 int64 per pair: dis, ntie, xtie, ytie, tot, n_entry, b) may be NULL.
// z != NULL: also the normal deviate of every pair (src/kendallc.cpp:321), which the parity
// tests need for the z^2-aware p-value tolerance
int icikt_oracle_pair_loop_z(const double* data, int64_t n, int64_t C, const int32_t* pi,
                             const int32_t* pj, int64_t P, int perspective, int alternative,
                             int continuity, int emulate_int32, int ncore, double* raw,
                             double* pvalue, double* taumax, double* completeness, int32_t* status,
                             int64_t* counts, double* z) {
  (void)C;
  if (ncore < 1) ncore = 1;
  const int64_t n_each = (P + ncore - 1) / ncore;
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t k = lo; k < hi; k++) {
      icikt_oracle_result r;
      icikt_oracle_ici_kt(data + (int64_t)pi[k] * n, n, data + (int64_t)pj[k] * n, n, perspective,
                          alternative, continuity, emulate_int32, &r);
      raw[k] = r.tau;
      pvalue[k] = r.pvalue;
      taumax[k] = r.tau_max;
      completeness[k] = r.completeness;
      if (status) status[k] = r.status;
      if (z) z[k] = r.z;
      if (counts) {
        int64_t* c = counts + 7 * k;
        c[0] = r.dis; c[1] = r.ntie; c[2] = r.xtie; c[3] = r.ytie; c[4] = r.tot; c[5] = r.n_entry;
        c[6] = r.n_matching_na;
      }
    }
  };
  if (ncore == 1) {
    work(0, P);
    return 0;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < ncore; t++) {
    const int64_t lo = t * n_each, hi = std::min<int64_t>(P, lo + n_each);
    if (lo >= hi) break;
    th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
  return 0;
}

int icikt_oracle_pair_loop(const double* data, int64_t n, int64_t C, const int32_t* pi,
                           const int32_t* pj, int64_t P, int perspective, int alternative,
                           int continuity, int emulate_int32, int ncore, double* raw,
                           double* pvalue, double* taumax, double* completeness, int32_t* status,
                           int64_t* counts) {
  return icikt_oracle_pair_loop_z(data, n, C, pi, pj, P, perspective, alternative, continuity, emulate_int32,
                                  ncore, raw, pvalue, taumax, completeness, status, counts, nullptr);
}

}  // extern "C"
