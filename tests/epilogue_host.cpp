// Test-only host build of the fp64 epilogue (icikt_common.cuh is __host__ __device__), so the
// formulas the K3 kernel runs can be checked against the oracle without a GPU.
#include "../icikendalltau_b200/csrc/icikt_common.cuh"

extern "C" int epilogue_host(long long n, const long long* xs, const long long* ys, long long dis,
                             long long ntie, long long b, long long g00, int perspective, int alternative,
                             int continuity, double* out4, long long* out_counts) {
  // xs/ys: n_na, n_groups, g0extra, s2o, s3o, s5o
  icikt::ColStats X{}, Y{};
  X.n_na = (int)xs[0]; X.n_groups = (int)xs[1]; X.g0extra = (int)xs[2]; X.s2o = xs[3]; X.s3o = xs[4]; X.s5o = xs[5];
  Y.n_na = (int)ys[0]; Y.n_groups = (int)ys[1]; Y.g0extra = (int)ys[2]; Y.s2o = ys[3]; Y.s3o = ys[4]; Y.s5o = ys[5];
  icikt::PairOut o;
  icikt::pair_epilogue(n, X, Y, dis, ntie, b, g00, perspective, alternative, continuity, o);
  out4[0] = o.tau; out4[1] = o.pvalue; out4[2] = o.taumax; out4[3] = o.completeness;
  out_counts[0] = o.xtie; out_counts[1] = o.ytie; out_counts[2] = o.tot; out_counts[3] = o.n_entry;
  return o.status;
}

extern "C" double one_minus_ratio_host(unsigned long long m, unsigned long long n) {
  return icikt::one_minus_ratio_x87(m, n);
}
