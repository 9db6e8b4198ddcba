// icikt_capi.cu -- host side of libicikt_b200.so: the C ABI declared in include/icikt_b200.h.
// Owns device memory, the streams, the pair-unit table and the CUDA-event timings; all
// arithmetic lives in icikt_columns.cu / icikt_pairs.cu.  There is no CPU fallback.
// Large all-pairs jobs of the one-shot entry points are pipelined (build_stage_table, run_staged):
// column chunks upload behind the pair launches of earlier chunks, finished row blocks of the
// results leave behind later ones.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <condition_variable>
#include <thread>
#include <vector>

#include "../../include/icikt_b200.h"
#include "icikt_internal.h"

using namespace icikt;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return ICIKT_ERR_CUDA;
}
// failure of one of the launchers in icikt_columns.cu / icikt_pairs.cu / icikt_reshape.cu
int launch_fail(const char* what) {
  cudaError_t e = g_launch_error;
  if (e == cudaSuccess) e = cudaGetLastError();
  g_err = std::string(what) + ": " + (g_launch_note ? g_launch_note : "") + ": " + cudaGetErrorString(e);
  g_launch_error = cudaSuccess;
  g_launch_note = "";
  return ICIKT_ERR_CUDA;
}
#define CK(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
  } while (0)

template <typename T>
cudaError_t dmalloc(T** p, size_t count) {
  return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T));
}

int64_t tri_pairs(int64_t C) { return C * (C - 1) / 2; }
int64_t row_start(int64_t i, int64_t C) { return i * (2 * C - i - 1) / 2; }

}  // namespace

// result sets up to stage_all() bytes: one pinned copy of everything; above: two pinned chunks of
// stage_chunk() bytes (the environment overrides exist for the tests)
static size_t env_bytes(const char* name, size_t dflt) {
  const char* e = std::getenv(name);
  const long long v = e ? std::atoll(e) : 0;
  return v > 0 ? (size_t)v : dflt;
}
static size_t stage_all() { return env_bytes("ICIKT_STAGE_ALL", 96u << 20); }
static size_t stage_chunk() { return env_bytes("ICIKT_STAGE_CHUNK", 32u << 20); }

// One launch of the pipelined one-shot call.
struct StageLaunch {
  int64_t col_lo = 0, col_hi = 0;    // columns uploaded and preprocessed right before it (empty: none)
  int64_t unit_lo = 0, unit_hi = 0;  // its units
  int64_t slot_lo = 0, slot_hi = 0;  // results complete once it has run: a contiguous range of the pair order
  int max_unit_pairs = 1;
};

// The launches of the pipelined one-shot call over all pairs of C columns (combn order, then the C diagonal
// pairs if include_diag).  Columns arrive in chunks [0, f), [f, 2f), [2f, 4f) ... (f ~ C/16): the launch of a
// chunk covers every pair (i, j), i < j, whose LATER column j lies in the chunk, so it needs nothing that has
// not been uploaded, and the upload of the next chunk hides behind it (the work of a chunk grows with the
// square of the columns, the upload linearly).  The last chunk holds most of the pairs; it is cut into
// `n_blocks` row blocks of about equal work: once the block of rows [r0, r1) has run, every pair of those rows
// is done -- slots [row_start(r0), row_start(r1)) of the pair order -- and can be copied out while the next
// block runs.  Units never straddle rows; their length tapers off towards the end of a launch (long units
// while much work is left, single pairs at the end) so that the persistent CTAs run dry together.
void build_stage_table(int64_t C, bool include_diag, int64_t cta_slots, int n_blocks, std::vector<PairUnit>& U,
                       std::vector<StageLaunch>& S) {
  U.clear();
  S.clear();
  const int64_t ptri = C * (C - 1) / 2;
  auto rstart = [C](int64_t i) { return i * (2 * C - i - 1) / 2; };
  std::vector<int64_t> cb{0};
  int64_t f = std::max<int64_t>(64, ((C / 16) + 7) & ~int64_t(7));
  for (; 4 * f <= 3 * C; f *= 2) cb.push_back(f);
  cb.push_back(C);
  // pairs (i, j) with r0 <= i < r1 and max(i + 1, lo) <= j < hi, as units appended to U
  auto emit = [&](int64_t r0, int64_t r1, int64_t lo, int64_t hi, StageLaunch& st) {
    int64_t total = 0;
    for (int64_t i = r0; i < r1; ++i) total += std::max<int64_t>(0, hi - std::max(i + 1, lo));
    const int64_t ulen_max = std::max<int64_t>(1, std::min<int64_t>(16, total / (16 * cta_slots)));
    int64_t left = total;
    st.unit_lo = (int64_t)U.size();
    for (int64_t i = r0; i < r1; ++i) {
      const int64_t ja = std::max(i + 1, lo);
      for (int64_t j = ja; j < hi;) {
        const int64_t want = std::max<int64_t>(1, std::min<int64_t>(ulen_max, left / (4 * cta_slots)));
        const int64_t len = std::min<int64_t>(want, hi - j);
        PairUnit u;
        u.slot0 = rstart(i) + (j - i - 1);
        u.col = (int32_t)i;
        u.j0 = (int32_t)j;
        u.count = (int32_t)len;
        u.j_explicit = 0;
        U.push_back(u);
        st.max_unit_pairs = std::max(st.max_unit_pairs, (int)len);
        j += len;
        left -= len;
      }
    }
    st.unit_hi = (int64_t)U.size();
  };
  const size_t nchunks = cb.size() - 1;
  for (size_t s = 0; s + 1 < nchunks; ++s) {
    StageLaunch st;
    st.col_lo = cb[s];
    st.col_hi = cb[s + 1];
    emit(0, cb[s + 1] - 1, cb[s], cb[s + 1], st);
    S.push_back(st);
  }
  // the last chunk in row blocks of about equal work
  const int64_t lo = cb[nchunks - 1];
  int64_t total = 0;
  for (int64_t i = 0; i + 1 < C; ++i) total += C - std::max(i + 1, lo);
  const int B = (int)std::max<int64_t>(1, std::min<int64_t>(n_blocks, total / std::max<int64_t>(1, 64 * cta_slots)));
  int64_t r0 = 0, acc = 0;
  for (int b = 0; b < B; ++b) {
    int64_t r1 = r0;
    const int64_t goal = total * (b + 1) / B;
    if (b == B - 1) {
      r1 = std::max<int64_t>(C - 1, 0);
      acc = total;
    } else {
      while (r1 + 1 < C && acc < goal) {
        acc += C - std::max(r1 + 1, lo);
        ++r1;
      }
    }
    StageLaunch st;
    if (b == 0) {
      st.col_lo = lo;
      st.col_hi = C;
    }
    emit(r0, r1, lo, C, st);
    st.slot_lo = rstart(r0);
    st.slot_hi = (b == B - 1) ? ptri : rstart(r1);
    if (b == B - 1 && include_diag) {
      for (int64_t i = 0; i < C; ++i) {
        PairUnit u;
        u.slot0 = ptri + i;
        u.col = (int32_t)i;
        u.j0 = (int32_t)i;
        u.count = 1;
        u.j_explicit = 0;
        U.push_back(u);
      }
      st.unit_hi = (int64_t)U.size();
      st.slot_hi = ptri + C;
    }
    S.push_back(st);
    r0 = r1;
  }
}

struct icikt_plan {
  int device = 0;
  int n_sm = 0;
  cudaStream_t stream = nullptr;
  int64_t n = 0, C = 0, P = 0;
  icikt_opts opts{};
  bool want_counts = false;

  double* d_data_own = nullptr;
  const double* d_data = nullptr;
  int64_t ld = 0;
  double* d_global_na = nullptr;

  ColumnTables tab;
  ColumnWork wk;
  bool columns_done = false;
  // launch shapes of the pair kernel, one per tier (see tiled_shape): no large tie groups /
  // large tie groups sorted in place / pass B.  All are enqueued; the device-side maxima over the
  // columns pick the one that runs (no host round trip).
  TiledShape shape, shape_mid, shape_heavy;
  unsigned char* d_scratch = nullptr;  // global-memory variant of the pair kernel
  size_t scratch_bytes = 0;

  std::vector<PairUnit> units;
  int max_unit_pairs = 1;  // longest unit
  PairUnit* d_units = nullptr;
  int32_t* d_pj = nullptr;
  PairRaw* d_raw = nullptr;
  PairComplete* d_pw = nullptr;  // complete-observations mode
  // results, packed in one allocation so that one copy brings them to the host:
  // [tau P][pvalue P][taumax P][completeness P] doubles, [max taumax bits] u64, [status P] int32
  unsigned char* d_res = nullptr;
  unsigned char* h_res = nullptr;  // pinned staging copy (small and medium result sets only)
  unsigned char* h_stage[2] = {nullptr, nullptr};  // large result sets: two pinned chunks, copy-out pipelined
  size_t stage_bytes = 0;
  cudaEvent_t stage_ev[2]{};
  bool stage_busy[2] = {false, false};  // an upload out of the chunk is in flight (staged_copy_in)
  size_t res_bytes = 0;
  double *d_tau = nullptr, *d_p = nullptr, *d_tm = nullptr, *d_comp = nullptr;
  unsigned long long* d_maxbits = nullptr;
  int32_t* d_status = nullptr;
  int64_t* d_counts = nullptr;
  unsigned long long* d_scalars = nullptr;  // [0] unit counter, [1] max taumax bits
  uint32_t* d_naive = nullptr;
  int64_t naive_threads = 0;
  // matrix output (icikt_plan_download_matrices), allocated on first use
  double* d_mat = nullptr;          // 5 x [C][C]: cor, raw, pvalue, taumax, completeness (multi-GPU: 5 blocks of columns)
  size_t mat_elems = 0;             // doubles allocated at d_mat
  unsigned char* h_mat = nullptr;   // pinned copy of all five (small C only)
  unsigned long long* d_hist = nullptr;  // [16] pairs per status class
  int32_t* d_ngood = nullptr;       // [C] caller-supplied n_good
  bool pairs_done = false;

  // Pipelined one-shot call (staged plans only, see build_stage_table / run_staged): the pair order is cut
  // into launches that need only the columns uploaded so far, then into row blocks whose results are
  // contiguous in the pair order and leave while the next block is computed.
  bool staged = false;
  bool staged_run = false;               // the last run went through run_staged (selects the timing sums)
  std::vector<StageLaunch> stages;
  cudaStream_t copy_stream = nullptr;    // uploads and downloads of a staged run
  std::vector<cudaEvent_t> sync_ev;      // per launch: [2l] its columns have landed, [2l+1] its results are ready
  std::vector<cudaEvent_t> time_ev;      // per launch: K1 start/end, K2 start, K2 end, K3 end

  cudaEvent_t ev[9]{};
  icikt_timings tm{};
};

namespace {

void free_plan(icikt_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaFree(p->d_data_own);
  cudaFree(p->d_global_na);
  cudaFree(p->tab.perm);
  cudaFree(p->tab.rank);
  cudaFree(p->tab.trow);
  cudaFree(p->tab.trun);
  cudaFree(p->tab.tend);
  cudaFree(p->tab.tord);
  cudaFree(p->tab.nabits);
  cudaFree(p->tab.firstbits);
  cudaFree(p->tab.gstart);
  cudaFree(p->tab.lgrp);
  cudaFree(p->tab.stats);
  cudaFree(p->tab.max_tied);
  cudaFree(p->wk.keys_in);
  cudaFree(p->wk.keys_out);
  cudaFree(p->wk.vals_in);
  cudaFree(p->wk.gpos);
  cudaFree(p->d_units);
  cudaFree(p->d_pj);
  cudaFree(p->d_raw);
  cudaFree(p->d_pw);
  cudaFree(p->d_res);
  if (p->h_res) cudaFreeHost(p->h_res);
  for (int i = 0; i < 2; ++i) {
    if (p->h_stage[i]) cudaFreeHost(p->h_stage[i]);
    if (p->stage_ev[i]) cudaEventDestroy(p->stage_ev[i]);
  }
  cudaFree(p->d_counts);
  cudaFree(p->d_scalars);
  cudaFree(p->d_naive);
  cudaFree(p->d_scratch);
  cudaFree(p->d_mat);
  if (p->h_mat) cudaFreeHost(p->h_mat);
  cudaFree(p->d_hist);
  cudaFree(p->d_ngood);
  for (auto& e : p->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : p->sync_ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : p->time_ev)
    if (e) cudaEventDestroy(e);
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

// Split the pair order into units: consecutive pairs that share their first column.
int build_units(icikt_plan* p, const int32_t* pi, const int32_t* pj, int64_t P_list) {
  const int64_t C = p->C;
  std::vector<PairUnit>& U = p->units;
  U.clear();
  const int64_t slots = (int64_t)p->n_sm * 2;
  if (!pi) {
    const int64_t ptri = tri_pairs(C);
    const int64_t ptot = ptri + (p->opts.include_diag ? C : 0);
    int64_t lo = p->opts.pair_lo, hi = p->opts.pair_hi;
    if (lo == 0 && hi == 0) hi = ptot;
    if (lo < 0 || hi > ptot || lo > hi) return fail(ICIKT_ERR_BAD_ARG, "pair_lo/pair_hi out of range");
    p->P = hi - lo;
    int64_t ulen = p->P / (16 * slots);
    ulen = std::max<int64_t>(1, std::min<int64_t>(16, ulen));
    for (int64_t i = 0; i + 1 < C; ++i) {
      const int64_t rs = row_start(i, C), re = rs + (C - 1 - i);
      const int64_t a = std::max(rs, lo), b = std::min(re, hi);
      for (int64_t s = a; s < b; s += ulen) {
        PairUnit u;
        u.slot0 = s - lo;
        u.col = (int32_t)i;
        u.j0 = (int32_t)(i + 1 + (s - rs));
        u.count = (int32_t)std::min<int64_t>(ulen, b - s);
        u.j_explicit = 0;
        U.push_back(u);
      }
      if (re >= hi) break;
    }
    if (p->opts.include_diag) {
      for (int64_t i = 0; i < C; ++i) {
        const int64_t s = ptri + i;
        if (s < lo || s >= hi) continue;
        PairUnit u;
        u.slot0 = s - lo;
        u.col = (int32_t)i;
        u.j0 = (int32_t)i;
        u.count = 1;
        u.j_explicit = 0;
        U.push_back(u);
      }
    }
  } else {
    p->P = P_list;
    for (int64_t k = 0; k < P_list; ++k)
      if (pi[k] < 0 || pi[k] >= C || pj[k] < 0 || pj[k] >= C)
        return fail(ICIKT_ERR_BAD_ARG, "pair index out of range");
    int64_t ulen = P_list / (16 * slots);
    ulen = std::max<int64_t>(1, std::min<int64_t>(16, ulen));
    int64_t k = 0;
    while (k < P_list) {
      int64_t e = k + 1;
      while (e < P_list && e - k < ulen && pi[e] == pi[k]) ++e;
      PairUnit u;
      u.slot0 = k;
      u.col = pi[k];
      u.j0 = pj[k];
      u.count = (int32_t)(e - k);
      u.j_explicit = 1;
      U.push_back(u);
      k = e;
    }
  }
  return ICIKT_OK;
}

int select_device(int device) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt < 1) {
    cudaGetLastError();
    return fail(ICIKT_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= cnt) return fail(ICIKT_ERR_BAD_ARG, "device ordinal out of range");
  CK(cudaSetDevice(device));
  return ICIKT_OK;
}

// memcpy into the caller's (usually freshly allocated, not yet faulted-in) array: one thread moves
// about 5 GB/s there, a PCIe 5 link delivers ten times that, so large blocks are split over a few
// threads (ICIKT_COPY_THREADS, default min(8, hardware threads); 1 = plain memcpy)
void host_copy(void* dst, const void* src, size_t bytes) {
  static const int n_threads = [] {
    const char* e = std::getenv("ICIKT_COPY_THREADS");
    int t = e ? std::atoi(e) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    return std::max(1, std::min(t, 64));
  }();
  constexpr size_t kMinPart = 2u << 20;
  const int parts = (int)std::min<size_t>((size_t)n_threads, bytes / kMinPart);
  if (parts <= 1) {
    std::memcpy(dst, src, bytes);
    return;
  }
  const size_t per = ((bytes / parts) + 4095) & ~size_t(4095);
  std::vector<std::thread> th;
  th.reserve((size_t)parts - 1);
  for (int k = 1; k < parts; ++k) {
    const size_t o = per * k;
    if (o >= bytes) break;
    const size_t l = std::min(per, bytes - o);
    th.emplace_back([=] { std::memcpy(static_cast<unsigned char*>(dst) + o, static_cast<const unsigned char*>(src) + o, l); });
  }
  std::memcpy(dst, src, std::min(per, bytes));
  for (auto& t : th) t.join();
}

// Device -> pageable host memory through two pinned chunks: the copy of chunk k+1 runs while chunk k
// is moved from the staging buffer into the caller's array (cudaMemcpyAsync straight into pageable
// memory is staged by the driver at a fraction of the link rate).
int staged_copy_out(icikt_plan* p, void* dst, const void* d_src, size_t bytes, cudaStream_t stream) {
  unsigned char* out = static_cast<unsigned char*>(dst);
  const unsigned char* src = static_cast<const unsigned char*>(d_src);
  size_t off[2] = {0, 0}, len[2] = {0, 0};
  int k = 0;
  p->stage_busy[0] = p->stage_busy[1] = false;  // uploads out of the chunks precede these copies in stream order
  for (size_t o = 0; o < bytes || len[0] || len[1]; ++k) {
    const int cur = k & 1;
    if (len[cur]) {  // the chunk issued two steps ago has landed in h_stage[cur]
      CK(cudaEventSynchronize(p->stage_ev[cur]));
      host_copy(out + off[cur], p->h_stage[cur], len[cur]);
      len[cur] = 0;
    }
    if (o < bytes) {
      const size_t l = std::min(p->stage_bytes, bytes - o);
      CK(cudaMemcpyAsync(p->h_stage[cur], src + o, l, cudaMemcpyDeviceToHost, stream));
      CK(cudaEventRecord(p->stage_ev[cur], stream));
      off[cur] = o;
      len[cur] = l;
      o += l;
    }
  }
  return ICIKT_OK;
}

int ensure_stage(icikt_plan* p) {
  if (p->h_stage[0]) return ICIKT_OK;
  p->stage_bytes = stage_chunk();
  for (int i = 0; i < 2; ++i) {
    CK(cudaMallocHost(reinterpret_cast<void**>(&p->h_stage[i]), p->stage_bytes));
    CK(cudaEventCreateWithFlags(&p->stage_ev[i], cudaEventDisableTiming));
  }
  return ICIKT_OK;
}

// Pageable host memory -> device through the plan's two pinned chunks: cudaMemcpyAsync from pageable
// memory is staged by the driver on one thread (~10 GB/s, a fraction of the link); here chunk k + 1 is
// copied into pinned memory by several host threads while chunk k is on the wire.  The source is fully
// consumed when this returns (same guarantee as the plain call).
int staged_copy_in(icikt_plan* p, void* d_dst, const void* src, size_t bytes, cudaStream_t stream) {
  int rc = ensure_stage(p);
  if (rc != ICIKT_OK) return rc;
  unsigned char* dst = static_cast<unsigned char*>(d_dst);
  const unsigned char* in = static_cast<const unsigned char*>(src);
  int k = 0;
  for (size_t o = 0; o < bytes; ++k) {
    const int cur = k & 1;
    if (p->stage_busy[cur]) {  // the chunk sent two steps ago (possibly by the previous call) has left
      CK(cudaEventSynchronize(p->stage_ev[cur]));
      p->stage_busy[cur] = false;
    }
    const size_t l = std::min(p->stage_bytes, bytes - o);
    host_copy(p->h_stage[cur], in + o, l);
    CK(cudaMemcpyAsync(dst + o, p->h_stage[cur], l, cudaMemcpyHostToDevice, stream));
    CK(cudaEventRecord(p->stage_ev[cur], stream));
    p->stage_busy[cur] = true;
    o += l;
  }
  return ICIKT_OK;
}

bool host_is_pinned(const void* ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// host -> device copy of columns [c_lo, c_hi) into the plan's own matrix buffer, on `stream`
int copy_columns_in(icikt_plan* p, const double* data, int64_t ld, int64_t c_lo, int64_t c_hi, cudaStream_t stream) {
  if (!p->d_data_own) CK(dmalloc(&p->d_data_own, (size_t)p->n * p->C));
  if (c_hi <= c_lo) return ICIKT_OK;
  const size_t bytes = sizeof(double) * (size_t)p->n * (size_t)(c_hi - c_lo);
  if (ld == p->n && bytes >= (8u << 20) && !std::getenv("ICIKT_NO_STAGED_UPLOAD") && !host_is_pinned(data + (size_t)c_lo * ld))
    return staged_copy_in(p, p->d_data_own + (size_t)c_lo * p->n, data + (size_t)c_lo * ld, bytes, stream);
  CK(cudaMemcpy2DAsync(p->d_data_own + (size_t)c_lo * p->n, sizeof(double) * p->n, data + (size_t)c_lo * ld,
                       sizeof(double) * ld, sizeof(double) * p->n, (size_t)(c_hi - c_lo), cudaMemcpyHostToDevice, stream));
  return ICIKT_OK;
}

int upload_columns(icikt_plan* p, const double* data, int64_t ld, int64_t c_lo, int64_t c_hi) {
  CK(cudaSetDevice(p->device));
  CK(cudaEventRecord(p->ev[0], p->stream));
  const int rc = copy_columns_in(p, data, ld, c_lo, c_hi, p->stream);
  if (rc != ICIKT_OK) return rc;
  CK(cudaEventRecord(p->ev[1], p->stream));
  p->d_data = p->d_data_own;
  p->ld = p->n;
  p->columns_done = false;
  p->pairs_done = false;
  p->staged_run = false;
  return ICIKT_OK;
}

// reusable barrier for the worker threads of icikt_all_pairs_multi
class HostBarrier {
 public:
  explicit HostBarrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    const int gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen != gen_; });
    }
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int n_, count_ = 0, gen_ = 0;
};

// elapsed time between two completed events; 0 if one of them was never recorded
float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) ms = 0.f;
  return ms;
}

}  // namespace

extern "C" {

void icikt_default_opts(icikt_opts* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->perspective = ICIKT_PERSPECTIVE_GLOBAL;
  o->alternative = ICIKT_ALT_TWO_SIDED;
  o->kernel = ICIKT_KERNEL_TILED;
}

int icikt_abi_version(void) { return ICIKT_ABI_VERSION; }

int icikt_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return cnt;
}

int64_t icikt_max_n(void) { return tiled_max_n(); }

const char* icikt_last_error(void) { return g_err.c_str(); }

// `staged`: the unit table of the pipelined one-shot call (all pairs only), see build_stage_table
static int plan_create_impl(icikt_plan** out, int64_t n, int64_t C, const int32_t* pi, const int32_t* pj,
                            int64_t P, const icikt_opts* opts_in, bool staged) {
  if (!out) return fail(ICIKT_ERR_BAD_ARG, "plan pointer is NULL");
  *out = nullptr;
  if (n < 1 || C < 1) return fail(ICIKT_ERR_BAD_ARG, "n and C must be >= 1");
  if ((pi == nullptr) != (pj == nullptr)) return fail(ICIKT_ERR_BAD_ARG, "pi and pj must both be given");
  if (pi && P < 1) return fail(ICIKT_ERR_BAD_ARG, "empty pair list");
  if (n > tiled_max_n()) return fail(ICIKT_ERR_TOO_LONG, "n exceeds icikt_max_n()");
  if (C > 2147483647LL) return fail(ICIKT_ERR_BAD_ARG, "too many columns");
  icikt_opts o;
  if (opts_in) o = *opts_in; else icikt_default_opts(&o);
  int rc = select_device(o.device);
  if (rc != ICIKT_OK) return rc;

  icikt_plan* p = new (std::nothrow) icikt_plan();
  if (!p) return fail(ICIKT_ERR_ALLOC, "out of host memory");
  p->device = o.device;
  p->opts = o;
  p->want_counts = o.want_counts != 0;
  p->n = n;
  p->C = C;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, o.device) != cudaSuccess) { free_plan(p); return fail(ICIKT_ERR_CUDA, "cudaGetDeviceProperties failed"); }
  p->n_sm = prop.multiProcessorCount;
  if (prop.major < 10) {
    free_plan(p);
    return fail(ICIKT_ERR_NO_DEVICE, "device is not sm_100-class; this library is built for sm_100a only");
  }
  if (staged) {
    if (pi || o.pair_lo != 0 || o.pair_hi != 0) { free_plan(p); return fail(ICIKT_ERR_BAD_ARG, "staged plans cover all pairs"); }
    p->staged = true;
    p->P = tri_pairs(C) + (o.include_diag ? C : 0);
    // row blocks of the last chunk: the copy-out of the final block is all that stays exposed
    int blocks = (size_t)p->P * (4 * sizeof(double) + sizeof(int32_t)) > stage_all() ? 16 : 8;
    if (const char* e = std::getenv("ICIKT_PIPELINE_BLOCKS")) blocks = std::max(1, std::min(64, std::atoi(e)));
    build_stage_table(C, o.include_diag != 0, (int64_t)p->n_sm * 2, blocks, p->units, p->stages);
  } else {
    rc = build_units(p, pi, pj, P);
    if (rc != ICIKT_OK) { free_plan(p); return rc; }
  }
  for (const PairUnit& u : p->units) p->max_unit_pairs = std::max(p->max_unit_pairs, (int)u.count);

#define PCK(call)                                                \
  do {                                                           \
    cudaError_t e_ = (call);                                     \
    if (e_ != cudaSuccess) {                                     \
      const int c_ = (e_ == cudaErrorMemoryAllocation) ? fail(ICIKT_ERR_ALLOC, "device allocation failed: " #call) \
                                                       : cuda_fail(e_, #call);                     \
      free_plan(p);                                              \
      return c_;                                                 \
    }                                                            \
  } while (0)

  PCK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  for (auto& e : p->ev) PCK(cudaEventCreate(&e));
  if (staged) {
    PCK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    p->sync_ev.assign(2 * p->stages.size(), nullptr);
    p->time_ev.assign(5 * p->stages.size(), nullptr);
    for (auto& e : p->sync_ev) PCK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : p->time_ev) PCK(cudaEventCreate(&e));
  }
  ColumnTables& t = p->tab;
  t.n = n;
  t.C = C;
  t.nstride = (n + 63) & ~63LL;
  const int64_t nwords = ((n + 31) & ~31LL) / 32;
  t.wstride = (nwords + 3) & ~3LL;
  const size_t ne = (size_t)t.nstride * C, nw = (size_t)t.wstride * C;
  PCK(dmalloc(&t.perm, ne));
  PCK(dmalloc(&t.rank, ne));
  PCK(dmalloc(&t.trow, ne));
  PCK(dmalloc(&t.trun, ne));
  PCK(dmalloc(&t.tend, ne));
  PCK(dmalloc(&t.tord, ne));
  PCK(dmalloc(&t.nabits, nw));
  PCK(dmalloc(&t.firstbits, nw));
  t.gstride = t.nstride + 64;
  PCK(dmalloc(&t.gstart, (size_t)t.gstride * C));
  PCK(dmalloc(&t.lgrp, (size_t)kLargeStride * C));
  PCK(dmalloc(&t.stats, (size_t)C));
  PCK(dmalloc(&t.max_tied, 4));
  // padding words of the bit arrays (beyond n32/32) are never written by the kernels
  PCK(cudaMemsetAsync(t.nabits, 0, sizeof(uint32_t) * nw, p->stream));
  PCK(cudaMemsetAsync(t.firstbits, 0, sizeof(uint32_t) * nw, p->stream));
  PCK(cudaMemsetAsync(t.max_tied, 0, 4 * sizeof(int32_t), p->stream));
  p->shape = tiled_shape(n, 0, t.wstride, p->n_sm);
  if (!columns_fused(n)) {  // scratch of the multi-kernel column path (long columns)
    PCK(dmalloc(&p->wk.keys_in, ne));
    PCK(dmalloc(&p->wk.keys_out, ne));
    PCK(dmalloc(&p->wk.vals_in, ne));
    PCK(dmalloc(&p->wk.gpos, (size_t)(t.nstride + 64) * C));
  }
  PCK(dmalloc(&p->d_global_na, 64));

  const size_t np = (size_t)std::max<int64_t>(p->P, 1);
  PCK(dmalloc(&p->d_units, p->units.size()));
  PCK(cudaMemcpyAsync(p->d_units, p->units.data(), sizeof(PairUnit) * p->units.size(),
                      cudaMemcpyHostToDevice, p->stream));
  if (pi) {
    PCK(dmalloc(&p->d_pj, np));
    PCK(cudaMemcpyAsync(p->d_pj, pj, sizeof(int32_t) * np, cudaMemcpyHostToDevice, p->stream));
  }
  PCK(dmalloc(&p->d_raw, np));
  if (o.perspective == ICIKT_PERSPECTIVE_COMPLETE) PCK(dmalloc(&p->d_pw, np));
  p->res_bytes = np * (4 * sizeof(double) + sizeof(int32_t)) + sizeof(unsigned long long);
  PCK(cudaMalloc(reinterpret_cast<void**>(&p->d_res), p->res_bytes));
  p->d_tau = reinterpret_cast<double*>(p->d_res);
  p->d_p = p->d_tau + np;
  p->d_tm = p->d_p + np;
  p->d_comp = p->d_tm + np;
  p->d_maxbits = reinterpret_cast<unsigned long long*>(p->d_comp + np);
  p->d_status = reinterpret_cast<int32_t*>(p->d_maxbits + 1);
  if (p->res_bytes <= stage_all()) {
    PCK(cudaMallocHost(reinterpret_cast<void**>(&p->h_res), p->res_bytes));
  } else {
    p->stage_bytes = stage_chunk();
    for (int i = 0; i < 2; ++i) {
      PCK(cudaMallocHost(reinterpret_cast<void**>(&p->h_stage[i]), p->stage_bytes));
      PCK(cudaEventCreateWithFlags(&p->stage_ev[i], cudaEventDisableTiming));
    }
  }
  if (p->want_counts) PCK(dmalloc(&p->d_counts, np * ICIKT_NCOUNTS));
  PCK(dmalloc(&p->d_scalars, 2));  // [0] unit counter
  if (o.kernel == ICIKT_KERNEL_NAIVE) {
    p->naive_threads = std::min<int64_t>((int64_t)p->units.size(), (int64_t)p->n_sm * 256);
    p->naive_threads = std::max<int64_t>(p->naive_threads, 1);
    PCK(cudaMalloc(reinterpret_cast<void**>(&p->d_naive), naive_scratch_bytes(n, p->naive_threads)));
  }
  PCK(cudaStreamSynchronize(p->stream));  // pj / units are read from caller memory
#undef PCK
  *out = p;
  return ICIKT_OK;
}

int icikt_plan_create(icikt_plan** out, int64_t n, int64_t C, const int32_t* pi, const int32_t* pj,
                      int64_t P, const icikt_opts* opts_in) {
  return plan_create_impl(out, n, C, pi, pj, P, opts_in, false);
}

int64_t icikt_plan_num_pairs(const icikt_plan* p) { return p ? p->P : 0; }

int icikt_plan_upload(icikt_plan* p, const double* data, int64_t ld) {
  if (!p || !data || ld < p->n) return fail(ICIKT_ERR_BAD_ARG, "bad upload arguments");
  return upload_columns(p, data, ld, 0, p->C);
}

int icikt_plan_set_device_matrix(icikt_plan* p, const double* d_data, int64_t ld) {
  if (!p || !d_data || ld < p->n) return fail(ICIKT_ERR_BAD_ARG, "bad device matrix arguments");
  p->d_data = d_data;
  p->ld = ld;
  p->columns_done = false;
  p->pairs_done = false;
  CK(cudaSetDevice(p->device));
  CK(cudaEventRecord(p->ev[0], p->stream));
  CK(cudaEventRecord(p->ev[1], p->stream));
  return ICIKT_OK;
}

// K1 over columns [col_lo, col_hi); `finish`: the statistics of every column are in place afterwards
// (a full run), so the pair kernel may be launched
// `reset`: a partial run that starts a sequence of partial runs over all columns on this device (the
// column kernels then raise the tier maxima themselves); e0/e1: timing events (default ev[2], ev[3])
static int plan_columns_impl(icikt_plan* p, const double* global_na, int32_t n_global_na, int64_t col_lo,
                             int64_t col_hi, bool finish, bool reset = false, cudaEvent_t e0 = nullptr,
                             cudaEvent_t e1 = nullptr) {
  if (!p || !p->d_data) return fail(ICIKT_ERR_BAD_ARG, "no matrix set on the plan");
  if (n_global_na < 0 || (n_global_na > 0 && !global_na)) return fail(ICIKT_ERR_BAD_ARG, "bad global_na");
  if (col_lo < 0 || col_hi > p->C || col_lo > col_hi) return fail(ICIKT_ERR_BAD_ARG, "column range out of bounds");
  CK(cudaSetDevice(p->device));
  // R/utils.R:6-15: NA and Inf entries of global_na select classes, the rest are literals
  double lit[64];
  int nlit = 0, na_inf = p->opts.na_inf;
  for (int i = 0; i < n_global_na; ++i) {
    const double v = global_na[i];
    if (std::isnan(v)) continue;
    if (std::isinf(v)) { na_inf = 1; continue; }
    if (nlit == 64) return fail(ICIKT_ERR_BAD_ARG, "more than 64 global_na literals");
    lit[nlit++] = v;
  }
  if (nlit) CK(cudaMemcpyAsync(p->d_global_na, lit, sizeof(double) * nlit, cudaMemcpyHostToDevice, p->stream));
  if (reset) CK(cudaMemsetAsync(p->tab.max_tied, 0, 4 * sizeof(int32_t), p->stream));
  CK(cudaEventRecord(e0 ? e0 : p->ev[2], p->stream));
  // the largest tier decides whether the global scratch may be needed at all
  const int64_t n_units = (int64_t)p->units.size();
  const bool inplace_ok = p->opts.perspective != ICIKT_PERSPECTIVE_COMPLETE;  // that mode has no in-place kernel
  p->shape = tiled_shape(p->n, 0, p->tab.wstride, p->n_sm, n_units, inplace_ok);
  p->shape_mid = tiled_shape(p->n, 1, p->tab.wstride, p->n_sm, n_units, inplace_ok);
  p->shape_heavy = tiled_shape(p->n, 2, p->tab.wstride, p->n_sm, n_units, inplace_ok);
  const TiledShape& worst = p->shape_heavy;
  const int slot_bytes = std::max(worst.region_bytes, worst.const_region_bytes);
  const int slot_ctas = std::max(std::max(worst.max_ctas, p->shape.max_ctas), std::max(p->shape_mid.max_ctas, p->n_sm * 2));
  const bool any_gmem = worst.gmem || worst.const_gmem || p->shape.gmem || p->shape_mid.gmem;
  const size_t need = (size_t)slot_ctas * (size_t)slot_bytes;
  if (any_gmem && need > p->scratch_bytes) {
    // the shape depends on the perspective class and on tuning knobs in the environment, so a cached
    // plan may need a larger scratch than the one it was given first
    CK(cudaStreamSynchronize(p->stream));
    if (p->d_scratch) { cudaFree(p->d_scratch); p->d_scratch = nullptr; p->scratch_bytes = 0; }
    if (cudaMalloc(reinterpret_cast<void**>(&p->d_scratch), need) != cudaSuccess) {
      cudaGetLastError();
      return fail(ICIKT_ERR_ALLOC, "device allocation of the pair kernel's global scratch failed");
    }
    p->scratch_bytes = need;
  }
  for (TiledShape* sh : {&p->shape, &p->shape_mid, &p->shape_heavy}) {
    sh->scratch_stride = p->d_scratch ? slot_bytes : 0;
    sh->scratch_ctas = p->d_scratch ? (int)std::min<size_t>((size_t)slot_ctas, p->scratch_bytes / (size_t)slot_bytes) : 0;
  }
  const int l = launch_columns(p->d_data, p->ld, p->d_global_na, nlit, na_inf, p->tab, p->wk, p->shape,
                               p->d_scratch, p->stream, col_lo, col_hi);
  if (l < 0) return launch_fail("column kernels");
  CK(cudaEventRecord(e1 ? e1 : p->ev[3], p->stream));
  // no synchronisation: `lit` is pageable host memory, so the copy above was staged before
  // cudaMemcpyAsync returned
  p->tm.n_launches = reset || !e0 ? l : p->tm.n_launches + l;
  p->columns_done = finish;
  p->pairs_done = false;
  return ICIKT_OK;
}

int icikt_plan_columns(icikt_plan* p, const double* global_na, int32_t n_global_na) {
  if (!p) return fail(ICIKT_ERR_BAD_ARG, "plan is NULL");
  return plan_columns_impl(p, global_na, n_global_na, 0, p->C, true);
}

int icikt_plan_columns_range(icikt_plan* p, const double* global_na, int32_t n_global_na, int64_t col_lo,
                             int64_t col_hi) {
  if (!p) return fail(ICIKT_ERR_BAD_ARG, "plan is NULL");
  return plan_columns_impl(p, global_na, n_global_na, col_lo, col_hi, false);
}

int icikt_plan_columns_finish(icikt_plan* p) {
  if (!p || !p->d_data) return fail(ICIKT_ERR_BAD_ARG, "no matrix set on the plan");
  CK(cudaSetDevice(p->device));
  const int l = launch_max_tied(p->tab, p->stream);
  if (l < 0) return launch_fail("tier maxima kernel");
  CK(cudaEventRecord(p->ev[3], p->stream));  // the exchange of the tables counts as column time
  p->tm.n_launches += l;
  p->columns_done = true;
  return ICIKT_OK;
}

int icikt_plan_upload_columns(icikt_plan* p, const double* data, int64_t ld, int64_t col_lo, int64_t col_hi) {
  if (!p || !data || ld < p->n || col_lo < 0 || col_hi > p->C || col_lo > col_hi)
    return fail(ICIKT_ERR_BAD_ARG, "bad upload arguments");
  return upload_columns(p, data, ld, col_lo, col_hi);
}

int icikt_plan_tables(icikt_plan* p, icikt_table* out, int32_t cap) {
  if (!p) return fail(ICIKT_ERR_BAD_ARG, "plan is NULL");
  const ColumnTables& t = p->tab;
  const bool complete = p->opts.perspective == ICIKT_PERSPECTIVE_COMPLETE;
  // the sorted positions of the tied rows are read only by the pair kernel's variants for long vectors
  // (rank table not staged: in place / global scratch)
  bool positions = false;
  for (int tier = 0; tier < 3; ++tier) {
    const TiledShape sh = tiled_shape(p->n, tier, t.wstride, p->n_sm, 0, !complete);
    positions = positions || sh.gmem || sh.inplace_kk != 0;
  }
  const icikt_table all[] = {
      {t.perm, (int64_t)(2 * t.nstride)},       {t.rank, (int64_t)(2 * t.nstride)},
      {t.trow, (int64_t)(2 * t.nstride)},       {t.trun, (int64_t)(2 * t.nstride)},
      {t.tord, (int64_t)(4 * t.nstride)},       {t.nabits, (int64_t)(4 * t.wstride)},
      {t.firstbits, (int64_t)(4 * t.wstride)},  {t.lgrp, (int64_t)(2 * kLargeStride)},
      {t.stats, (int64_t)sizeof(ColStats)},     {t.tend, (int64_t)(2 * t.nstride)},
      {t.gstart, (int64_t)(2 * t.gstride)},
  };
  int n_out = 0;
  for (const icikt_table& e : all) {
    if (e.ptr == (void*)t.tend && !positions) continue;
    if (e.ptr == (void*)t.gstart && !complete) continue;  // that mode only
    if (out && n_out < cap) out[n_out] = e;
    ++n_out;
  }
  return n_out;
}

// K2 (every tier's shape; the device-side maxima pick the one that runs) + K3 over units [unit_lo, unit_hi);
// e_start / e_mid / e_end bracket K2 and K3 on the plan's stream
static int launch_pair_group(icikt_plan* p, int64_t unit_lo, int64_t unit_hi, int max_unit_pairs, cudaEvent_t e_start,
                             cudaEvent_t e_mid, cudaEvent_t e_end, int* launches_out) {
  CK(cudaMemsetAsync(p->d_scalars, 0, 2 * sizeof(unsigned long long), p->stream));
  CK(cudaEventRecord(e_start, p->stream));
  PairLaunch pl;
  pl.tab = &p->tab;
  pl.units = p->d_units + unit_lo;
  pl.n_units = unit_hi - unit_lo;
  pl.pj_list = p->d_pj;
  pl.raw = p->d_raw;
  pl.pw = (p->opts.perspective == ICIKT_PERSPECTIVE_COMPLETE) ? p->d_pw : nullptr;
  pl.unit_counter = p->d_scalars;
  pl.scratch = p->d_scratch;
  int launches = 0;
  if (pl.n_units <= 0) {
    CK(cudaEventRecord(e_mid, p->stream));
  } else {
    int l;
    if (pl.pw && p->opts.kernel == ICIKT_KERNEL_NAIVE)
      return fail(ICIKT_ERR_BAD_ARG, "the complete-observations mode needs the tiled kernel");
    if (p->opts.kernel == ICIKT_KERNEL_NAIVE)
      l = launch_pairs_naive(pl, p->P, p->d_naive, p->naive_threads, p->stream);
    else {
      l = launch_pairs_tiled(pl, p->shape, p->n_sm, 0, p->stream);
      if (l > 0) {
        const int l2 = launch_pairs_tiled(pl, p->shape_mid, p->n_sm, 1, p->stream);
        l = l2 < 0 ? l2 : l + l2;
      }
      if (l > 0) {
        const int l3 = launch_pairs_tiled(pl, p->shape_heavy, p->n_sm, 2, p->stream);
        l = l3 < 0 ? l3 : l + l3;
      }
    }
    if (l == -2)
      return fail(ICIKT_ERR_TOO_LONG, "the tied-value lists of this matrix do not fit the pair kernel's "
                                      "shared memory (long vectors with heavy ties)");
    if (l < 0) return launch_fail("pair kernel");
    launches += l;
    CK(cudaEventRecord(e_mid, p->stream));
    EpilogueLaunch el;
    el.tab = &p->tab;
    el.units = pl.units;
    el.n_units = pl.n_units;
    el.max_unit_pairs = max_unit_pairs;
    el.pj_list = p->d_pj;
    el.raw = p->d_raw;
    el.pw = pl.pw;
    el.perspective = p->opts.perspective;
    el.alternative = p->opts.alternative;
    el.continuity = p->opts.continuity;
    el.tau = p->d_tau;
    el.pvalue = p->d_p;
    el.taumax = p->d_tm;
    el.completeness = p->d_comp;
    el.status = p->d_status;
    el.counts = p->d_counts;
    el.max_taumax_bits = p->d_maxbits;
    l = launch_epilogue(el, p->stream);
    if (l < 0) return launch_fail("epilogue kernel");
    launches += l;
  }
  CK(cudaEventRecord(e_end, p->stream));
  *launches_out += launches;
  return ICIKT_OK;
}

int icikt_plan_pairs(icikt_plan* p) {
  if (!p || !p->columns_done) return fail(ICIKT_ERR_BAD_ARG, "icikt_plan_columns has not run");
  CK(cudaSetDevice(p->device));
  CK(cudaMemsetAsync(p->d_maxbits, 0, sizeof(unsigned long long), p->stream));
  int launches = 0;
  const int rc = launch_pair_group(p, 0, (int64_t)p->units.size(), p->max_unit_pairs, p->ev[4], p->ev[8], p->ev[5], &launches);
  if (rc != ICIKT_OK) return rc;
  p->tm.n_launches = (p->tm.n_launches & 0xffff) | (launches << 16);
  p->pairs_done = true;
  p->staged_run = false;
  return ICIKT_OK;
}

int icikt_plan_sync(icikt_plan* p) {
  if (!p) return fail(ICIKT_ERR_BAD_ARG, "plan is NULL");
  CK(cudaSetDevice(p->device));
  CK(cudaStreamSynchronize(p->stream));
  return ICIKT_OK;
}

int icikt_plan_download(icikt_plan* p, double* raw, double* pvalue, double* taumax, double* completeness,
                        int32_t* status, int64_t* counts, double* max_taumax) {
  if (!p) return fail(ICIKT_ERR_BAD_ARG, "plan is NULL");
  if (counts && !p->d_counts) return fail(ICIKT_ERR_BAD_ARG, "plan was created without want_counts");
  CK(cudaSetDevice(p->device));
  const size_t np = (size_t)p->P;
  CK(cudaEventRecord(p->ev[6], p->stream));
  unsigned long long bits = 0;
  if (p->h_res) {
    // one device-to-host copy into pinned memory, then plain host copies into the caller's arrays
    CK(cudaMemcpyAsync(p->h_res, p->d_res, p->res_bytes, cudaMemcpyDeviceToHost, p->stream));
    if (counts && np) CK(cudaMemcpyAsync(counts, p->d_counts, sizeof(int64_t) * np * ICIKT_NCOUNTS, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaEventRecord(p->ev[7], p->stream));
    CK(cudaStreamSynchronize(p->stream));
    const size_t npad = std::max<size_t>(np, 1);
    const double* h = reinterpret_cast<const double*>(p->h_res);
    if (np) {
      if (raw) host_copy(raw, h, sizeof(double) * np);
      if (pvalue) host_copy(pvalue, h + npad, sizeof(double) * np);
      if (taumax) host_copy(taumax, h + 2 * npad, sizeof(double) * np);
      if (completeness) host_copy(completeness, h + 3 * npad, sizeof(double) * np);
      if (status) host_copy(status, p->h_res + 4 * npad * sizeof(double) + sizeof(unsigned long long), sizeof(int32_t) * np);
    }
    std::memcpy(&bits, h + 4 * npad, sizeof(bits));
  } else {
    if (np) {
      int rc = ICIKT_OK;
      if (raw) rc = staged_copy_out(p, raw, p->d_tau, sizeof(double) * np, p->stream);
      if (rc == ICIKT_OK && pvalue) rc = staged_copy_out(p, pvalue, p->d_p, sizeof(double) * np, p->stream);
      if (rc == ICIKT_OK && taumax) rc = staged_copy_out(p, taumax, p->d_tm, sizeof(double) * np, p->stream);
      if (rc == ICIKT_OK && completeness) rc = staged_copy_out(p, completeness, p->d_comp, sizeof(double) * np, p->stream);
      if (rc == ICIKT_OK && status) rc = staged_copy_out(p, status, p->d_status, sizeof(int32_t) * np, p->stream);
      if (rc == ICIKT_OK && counts) rc = staged_copy_out(p, counts, p->d_counts, sizeof(int64_t) * np * ICIKT_NCOUNTS, p->stream);
      if (rc != ICIKT_OK) return rc;
    }
    CK(cudaMemcpyAsync(&bits, p->d_maxbits, sizeof(bits), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaEventRecord(p->ev[7], p->stream));
    CK(cudaStreamSynchronize(p->stream));
  }
  if (max_taumax) {
    double v;
    std::memcpy(&v, &bits, sizeof(v));
    *max_taumax = bits ? v : std::nan("");
  }
  return ICIKT_OK;
}

int icikt_plan_download_matrices(icikt_plan* p, int32_t scale_max, int32_t diag_good, const int32_t* n_good,
                                 double* cor, double* raw, double* pvalue, double* taumax, double* completeness,
                                 int64_t* status_counts, double* max_taumax) {
  if (!p || !p->pairs_done) return fail(ICIKT_ERR_BAD_ARG, "icikt_plan_pairs has not run");
  CK(cudaSetDevice(p->device));
  const size_t C = (size_t)p->C, cc = C * C;
  double* outs[5] = {cor, raw, pvalue, taumax, completeness};
  if (p->mat_elems < 5 * cc) {
    cudaFree(p->d_mat);
    p->d_mat = nullptr;
    p->mat_elems = 0;
    if (dmalloc(&p->d_mat, 5 * cc) != cudaSuccess) {
      cudaGetLastError();
      return fail(ICIKT_ERR_ALLOC, "device allocation of the result matrices failed");
    }
    p->mat_elems = 5 * cc;
  }
  if (!p->d_hist) CK(dmalloc(&p->d_hist, 16));
  if (!p->d_ngood) CK(dmalloc(&p->d_ngood, C));
  CK(cudaEventRecord(p->ev[6], p->stream));
  // entries no pair writes stay 0 (R/kendalltau.R:389-396 start from matrix(0, ...)); with every pair
  // and the diagonal present nothing needs clearing
  const int64_t ptot = tri_pairs(p->C) + (p->opts.include_diag ? p->C : 0);
  const bool covered = !p->d_pj && p->P == ptot && (p->opts.include_diag || diag_good);
  MatrixFill mf{};
  for (int k = 0; k < 5; ++k) {
    mf.m[k] = outs[k] ? p->d_mat + (size_t)k * cc : nullptr;
    if (outs[k] && !covered) CK(cudaMemsetAsync(mf.m[k], 0, sizeof(double) * cc, p->stream));
  }
  CK(cudaMemsetAsync(p->d_hist, 0, 16 * sizeof(unsigned long long), p->stream));
  if (n_good) CK(cudaMemcpyAsync(p->d_ngood, n_good, sizeof(int32_t) * C, cudaMemcpyHostToDevice, p->stream));
  mf.units = p->d_units;
  mf.n_units = (int64_t)p->units.size();
  mf.max_unit_pairs = p->max_unit_pairs;
  mf.pj_list = p->d_pj;
  mf.tau = p->d_tau;
  mf.pvalue = p->d_p;
  mf.taumax = p->d_tm;
  mf.completeness = p->d_comp;
  mf.status = p->d_status;
  mf.max_taumax_bits = p->d_maxbits;
  mf.stats = p->tab.stats;
  mf.n_good = n_good ? p->d_ngood : nullptr;
  mf.n = p->n;
  mf.C = p->C;
  mf.scale_max = scale_max != 0;
  mf.diag_good = diag_good != 0;
  mf.hist = p->d_hist;
  if (p->P <= 0) mf.n_units = 0;
  if (launch_matrix_fill(mf, p->stream) < 0) return launch_fail("matrix fill kernel");
  const size_t all_bytes = 5 * cc * sizeof(double);
  if (all_bytes <= stage_all()) {
    if (!p->h_mat) CK(cudaMallocHost(reinterpret_cast<void**>(&p->h_mat), all_bytes));
    for (int k = 0; k < 5; ++k)
      if (outs[k])
        CK(cudaMemcpyAsync(p->h_mat + (size_t)k * cc * sizeof(double), mf.m[k], sizeof(double) * cc,
                           cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int k = 0; k < 5; ++k)
      if (outs[k]) host_copy(outs[k], p->h_mat + (size_t)k * cc * sizeof(double), sizeof(double) * cc);
  } else {
    {
      const int rc = ensure_stage(p);
      if (rc != ICIKT_OK) return rc;
    }
    for (int k = 0; k < 5; ++k)
      if (outs[k]) {
        const int rc = staged_copy_out(p, outs[k], mf.m[k], sizeof(double) * cc, p->stream);
        if (rc != ICIKT_OK) return rc;
      }
  }
  unsigned long long hist[16], bits = 0;
  CK(cudaMemcpyAsync(hist, p->d_hist, sizeof(hist), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaMemcpyAsync(&bits, p->d_maxbits, sizeof(bits), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaEventRecord(p->ev[7], p->stream));
  CK(cudaStreamSynchronize(p->stream));
  if (status_counts) {
    int64_t bad = 0;
    for (int k = 1; k < ICIKT_NSTATUS; ++k) { status_counts[k] = (int64_t)hist[k]; bad += status_counts[k]; }
    status_counts[0] = p->P - bad;
  }
  if (max_taumax) {
    double v;
    std::memcpy(&v, &bits, sizeof(v));
    *max_taumax = bits ? v : std::nan("");
  }
  return ICIKT_OK;
}

int icikt_plan_column_info(icikt_plan* p, int32_t* n_na) {
  if (!p || !p->columns_done) return fail(ICIKT_ERR_BAD_ARG, "icikt_plan_columns has not run");
  CK(cudaSetDevice(p->device));
  std::vector<ColStats> s((size_t)p->C);
  CK(cudaMemcpyAsync(s.data(), p->tab.stats, sizeof(ColStats) * s.size(), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  if (n_na)
    for (int64_t c = 0; c < p->C; ++c) n_na[c] = s[(size_t)c].n_na;
  return ICIKT_OK;
}

void* icikt_plan_stream(icikt_plan* p) { return p ? (void*)p->stream : nullptr; }

int icikt_plan_timings(icikt_plan* p, icikt_timings* t) {
  if (!p || !t) return fail(ICIKT_ERR_BAD_ARG, "NULL argument");
  CK(cudaSetDevice(p->device));
  CK(cudaStreamSynchronize(p->stream));
  const int launches = (p->tm.n_launches & 0xffff) + (p->tm.n_launches >> 16);
  std::memset(t, 0, sizeof(*t));
  t->n_launches = launches;
  if (p->copy_stream) CK(cudaStreamSynchronize(p->copy_stream));
  // the streams are idle, so every recorded event has completed
  t->h2d_ms = ev_ms(p->ev[0], p->ev[1]);
  if (p->staged_run) {
    // pipelined call: kernel times are sums over the launches; h2d = start until the last column has landed,
    // d2h = first result copy until the last one (both overlap the kernels, so the parts exceed total_ms)
    for (size_t l = 0; l < p->stages.size(); ++l) {
      const cudaEvent_t* e = &p->time_ev[5 * l];
      if (p->stages[l].col_hi > p->stages[l].col_lo) t->columns_ms += ev_ms(e[0], e[1]);
      t->pairs_ms += ev_ms(e[2], e[3]);
      t->epilogue_ms += ev_ms(e[3], e[4]);
    }
  } else {
    t->columns_ms = ev_ms(p->ev[2], p->ev[3]);
    t->pairs_ms = ev_ms(p->ev[4], p->ev[8]);
    t->epilogue_ms = ev_ms(p->ev[8], p->ev[5]);
  }
  t->d2h_ms = ev_ms(p->ev[6], p->ev[7]);
  t->total_ms = ev_ms(p->ev[0], p->ev[7]);
  cudaGetLastError();
  return ICIKT_OK;
}

void icikt_plan_destroy(icikt_plan* p) { free_plan(p); }

// The pipelined one-shot call on a staged plan (build_stage_table): uploads on the copy stream chunk by chunk,
// K1 of a chunk and the pair launches that need nothing beyond it on the compute stream, the results of
// a finished row block back on the copy stream while the next block runs.  `download` false: the results
// stay on the device (matrix output follows).  Same results as upload + columns + pairs + download: only
// the order in which the pairs are computed differs.
static int run_staged(icikt_plan* p, const double* data, int64_t ld, const double* global_na, int32_t n_global_na,
                      bool download, double* raw, double* pvalue, double* taumax, double* completeness,
                      int32_t* status, int64_t* counts, double* max_taumax) {
  if (!p || !p->staged) return fail(ICIKT_ERR_BAD_ARG, "not a staged plan");
  if (!data || ld < p->n) return fail(ICIKT_ERR_BAD_ARG, "bad upload arguments");
  if (counts && !p->d_counts) return fail(ICIKT_ERR_BAD_ARG, "plan was created without want_counts");
  CK(cudaSetDevice(p->device));
  if (!p->d_data_own) CK(dmalloc(&p->d_data_own, (size_t)p->n * p->C));
  p->d_data = p->d_data_own;
  p->ld = p->n;
  p->columns_done = false;
  p->pairs_done = false;
  p->staged_run = true;
  CK(cudaEventRecord(p->ev[0], p->stream));
  CK(cudaStreamWaitEvent(p->copy_stream, p->ev[0], 0));
  CK(cudaMemsetAsync(p->d_maxbits, 0, sizeof(unsigned long long), p->stream));
  int pair_launches = 0;
  const size_t L = p->stages.size();
  for (size_t l = 0; l < L; ++l) {
    const StageLaunch& st = p->stages[l];
    cudaEvent_t* te = &p->time_ev[5 * l];
    if (st.col_hi > st.col_lo) {
      int rc = copy_columns_in(p, data, ld, st.col_lo, st.col_hi, p->copy_stream);
      if (rc != ICIKT_OK) return rc;
      CK(cudaEventRecord(p->sync_ev[2 * l], p->copy_stream));
      if (st.col_hi == p->C) CK(cudaEventRecord(p->ev[1], p->copy_stream));
      CK(cudaStreamWaitEvent(p->stream, p->sync_ev[2 * l], 0));
      rc = plan_columns_impl(p, global_na, n_global_na, st.col_lo, st.col_hi, st.col_hi == p->C, st.col_lo == 0, te[0], te[1]);
      if (rc != ICIKT_OK) return rc;
    }
    const int rc = launch_pair_group(p, st.unit_lo, st.unit_hi, st.max_unit_pairs, te[2], te[3], te[4], &pair_launches);
    if (rc != ICIKT_OK) return rc;
    CK(cudaEventRecord(p->sync_ev[2 * l + 1], p->stream));
  }
  p->tm.n_launches = (p->tm.n_launches & 0xffff) | (pair_launches << 16);
  p->pairs_done = true;
  bool d2h_started = false;  // ev[6]: the first result copy may start
  unsigned long long bits = 0;
  if (download) {
    const size_t np = (size_t)p->P, npad = std::max<size_t>(np, 1);
    struct Out { void* host; const unsigned char* dev; size_t elem; size_t res_off; };
    const Out outs[6] = {
        {raw, reinterpret_cast<const unsigned char*>(p->d_tau), sizeof(double), 0},
        {pvalue, reinterpret_cast<const unsigned char*>(p->d_p), sizeof(double), npad * sizeof(double)},
        {taumax, reinterpret_cast<const unsigned char*>(p->d_tm), sizeof(double), 2 * npad * sizeof(double)},
        {completeness, reinterpret_cast<const unsigned char*>(p->d_comp), sizeof(double), 3 * npad * sizeof(double)},
        {status, reinterpret_cast<const unsigned char*>(p->d_status), sizeof(int32_t),
         4 * npad * sizeof(double) + sizeof(unsigned long long)},
        {counts, reinterpret_cast<const unsigned char*>(p->d_counts), sizeof(int64_t) * ICIKT_NCOUNTS, 0},
    };
    if (p->h_res) {
      // every block's copies into the pinned mirror are enqueued at once (each waits for its block), the
      // host then follows block by block and moves the ranges into the caller's arrays
      for (size_t l = 0; l < L; ++l) {
        const StageLaunch& st = p->stages[l];
        if (st.slot_hi <= st.slot_lo) continue;
        CK(cudaStreamWaitEvent(p->copy_stream, p->sync_ev[2 * l + 1], 0));
        if (!d2h_started) CK(cudaEventRecord(p->ev[6], p->copy_stream));
        d2h_started = true;
        for (int k = 0; k < 6; ++k) {
          if (!outs[k].host) continue;
          const size_t o = (size_t)st.slot_lo * outs[k].elem, b = (size_t)(st.slot_hi - st.slot_lo) * outs[k].elem;
          void* dst = (k == 5) ? static_cast<void*>(static_cast<unsigned char*>(outs[k].host) + o)
                               : static_cast<void*>(p->h_res + outs[k].res_off + o);
          CK(cudaMemcpyAsync(dst, outs[k].dev + o, b, cudaMemcpyDeviceToHost, p->copy_stream));
        }
        CK(cudaEventRecord(p->sync_ev[2 * l], p->copy_stream));  // free since the block's columns were waited for
      }
      for (size_t l = 0; l < L; ++l) {
        const StageLaunch& st = p->stages[l];
        if (st.slot_hi <= st.slot_lo) continue;
        CK(cudaEventSynchronize(p->sync_ev[2 * l]));
        for (int k = 0; k < 5; ++k) {
          if (!outs[k].host) continue;
          const size_t o = (size_t)st.slot_lo * outs[k].elem, b = (size_t)(st.slot_hi - st.slot_lo) * outs[k].elem;
          host_copy(static_cast<unsigned char*>(outs[k].host) + o, p->h_res + outs[k].res_off + o, b);
        }
      }
    } else {
      for (size_t l = 0; l < L; ++l) {
        const StageLaunch& st = p->stages[l];
        if (st.slot_hi <= st.slot_lo) continue;
        CK(cudaStreamWaitEvent(p->copy_stream, p->sync_ev[2 * l + 1], 0));
        if (!d2h_started) CK(cudaEventRecord(p->ev[6], p->copy_stream));
        d2h_started = true;
        for (int k = 0; k < 6; ++k) {
          if (!outs[k].host) continue;
          const size_t o = (size_t)st.slot_lo * outs[k].elem, b = (size_t)(st.slot_hi - st.slot_lo) * outs[k].elem;
          const int rc = staged_copy_out(p, static_cast<unsigned char*>(outs[k].host) + o, outs[k].dev + o, b, p->copy_stream);
          if (rc != ICIKT_OK) return rc;
        }
      }
    }
  }
  CK(cudaStreamWaitEvent(p->copy_stream, p->sync_ev[2 * L - 1], 0));
  if (!d2h_started) CK(cudaEventRecord(p->ev[6], p->copy_stream));
  CK(cudaMemcpyAsync(&bits, p->d_maxbits, sizeof(bits), cudaMemcpyDeviceToHost, p->copy_stream));
  CK(cudaEventRecord(p->ev[7], p->copy_stream));
  CK(cudaStreamSynchronize(p->copy_stream));
  CK(cudaStreamSynchronize(p->stream));
  if (max_taumax) {
    double v;
    std::memcpy(&v, &bits, sizeof(v));
    *max_taumax = bits ? v : std::nan("");
  }
  return ICIKT_OK;
}

// all pairs of a matrix this large go through the pipelined call (ICIKT_PIPELINE_MIN_BYTES, default 32 MB of
// input; ICIKT_NO_PIPELINE=1 switches it off)
static bool want_staged(int64_t n, int64_t C, const icikt_opts& o, bool all_pairs) {
  if (!all_pairs || o.pair_lo != 0 || o.pair_hi != 0 || o.kernel != ICIKT_KERNEL_TILED) return false;
  if (std::getenv("ICIKT_NO_PIPELINE")) return false;
  const size_t min_bytes = env_bytes("ICIKT_PIPELINE_MIN_BYTES", 32u << 20);
  return C >= 128 && sizeof(double) * (size_t)n * (size_t)C >= min_bytes;
}

// One cached plan for the one-shot entry points (all-pairs shape only).
static icikt_plan* g_cached = nullptr;
static icikt_plan* g_cached_multi[64] = {};  // one slot per worker of icikt_all_pairs_multi
static std::mutex g_cache_mu;

static bool cache_matches(const icikt_plan* p, int64_t n, int64_t C, const icikt_opts& o, bool staged) {
  return p && p->staged == staged && p->n == n && p->C == C && p->opts.device == o.device && p->opts.kernel == o.kernel &&
         p->opts.include_diag == o.include_diag && p->opts.pair_lo == o.pair_lo &&
         p->opts.pair_hi == o.pair_hi && p->want_counts == (o.want_counts != 0);
}

struct MatrixOut {  // the download step of one_shot as matrices instead of per-pair arrays
  int32_t scale_max, diag_good;
  const int32_t* n_good;
  double* cor;
  int64_t* status_counts;
};

static int one_shot(const double* data, int64_t n, int64_t C, int64_t ld, const double* global_na,
                    int32_t n_global_na, const int32_t* pi, const int32_t* pj, int64_t P,
                    const icikt_opts* opts, double* raw, double* pvalue, double* taumax,
                    double* completeness, int32_t* status, int64_t* counts, double* max_taumax,
                    icikt_timings* timings, const MatrixOut* mat = nullptr) {
  if (!data || (!raw && !mat)) return fail(ICIKT_ERR_BAD_ARG, "data and raw must not be NULL");
  icikt_opts o;
  if (opts) o = *opts; else icikt_default_opts(&o);
  o.want_counts = counts ? 1 : 0;
  if (!pi && o.pair_lo == 0 && o.pair_hi == tri_pairs(C) + (o.include_diag ? C : 0)) o.pair_hi = 0;  // the whole order
  std::lock_guard<std::mutex> lock(g_cache_mu);
  icikt_plan* p = nullptr;
  const bool cacheable = (pi == nullptr);
  const bool staged = want_staged(n, C, o, pi == nullptr);
  int rc = ICIKT_OK;
  if (cacheable && cache_matches(g_cached, n, C, o, staged) &&
      (o.perspective != ICIKT_PERSPECTIVE_COMPLETE || g_cached->d_pw)) {
    p = g_cached;
    p->opts.perspective = o.perspective;
    p->opts.alternative = o.alternative;
    p->opts.continuity = o.continuity;
    p->opts.na_inf = o.na_inf;
  } else {
    if (cacheable && g_cached) { icikt_plan_destroy(g_cached); g_cached = nullptr; }
    rc = plan_create_impl(&p, n, C, pi, pj, P, &o, staged);
    if (rc != ICIKT_OK) return rc;
    if (cacheable) g_cached = p;
  }
  if (staged) {
    rc = run_staged(p, data, ld, global_na, n_global_na, mat == nullptr, raw, pvalue, taumax, completeness, status,
                    counts, max_taumax);
  } else {
    rc = icikt_plan_upload(p, data, ld);
    if (rc == ICIKT_OK) rc = icikt_plan_columns(p, global_na, n_global_na);
    if (rc == ICIKT_OK) rc = icikt_plan_pairs(p);
  }
  if (rc == ICIKT_OK && mat)
    rc = icikt_plan_download_matrices(p, mat->scale_max, mat->diag_good, mat->n_good, mat->cor, raw, pvalue, taumax,
                                      completeness, mat->status_counts, max_taumax);
  else if (rc == ICIKT_OK && !staged)
    rc = icikt_plan_download(p, raw, pvalue, taumax, completeness, status, counts, max_taumax);
  if (rc == ICIKT_OK && timings) rc = icikt_plan_timings(p, timings);
  const std::string keep = g_err;
  if (!cacheable) {
    icikt_plan_destroy(p);
  } else if (rc != ICIKT_OK) {
    icikt_plan_destroy(g_cached);
    g_cached = nullptr;
  }
  if (rc != ICIKT_OK) g_err = keep;
  return rc;
}

int icikt_pair_from_index(int64_t C, int32_t include_diag, int64_t index, int32_t* i, int32_t* j) {
  if (C < 1 || !i || !j || index < 0) return fail(ICIKT_ERR_BAD_ARG, "bad pair index arguments");
  const int64_t ptri = tri_pairs(C);
  if (index >= ptri) {
    if (!include_diag || index >= ptri + C) return fail(ICIKT_ERR_BAD_ARG, "pair index out of range");
    *i = *j = (int32_t)(index - ptri);
    return ICIKT_OK;
  }
  // row r starts at r(2C-r-1)/2: invert with a float guess, then fix up exactly
  const double b = 2.0 * (double)C - 1.0;
  int64_t r = (int64_t)((b - std::sqrt(b * b - 8.0 * (double)index)) / 2.0);
  r = std::max<int64_t>(0, std::min<int64_t>(r, C - 2));
  while (r > 0 && row_start(r, C) > index) --r;
  while (r + 1 < C - 1 && row_start(r + 1, C) <= index) ++r;
  *i = (int32_t)r;
  *j = (int32_t)(r + 1 + (index - row_start(r, C)));
  return ICIKT_OK;
}

void icikt_release_workspace(void) {
  std::lock_guard<std::mutex> lock(g_cache_mu);
  if (g_cached) { icikt_plan_destroy(g_cached); g_cached = nullptr; }
  for (icikt_plan*& s : g_cached_multi)
    if (s) { icikt_plan_destroy(s); s = nullptr; }
}

int64_t icikt_stage_table(int64_t C, int32_t include_diag, int32_t cta_slots, int32_t n_blocks, int64_t cap_units,
                          int64_t* units, int64_t cap_launches, int64_t* launches, int64_t* n_launches) {
  if (C < 1 || cta_slots < 1 || n_blocks < 1) return fail(ICIKT_ERR_BAD_ARG, "bad stage table arguments");
  std::vector<PairUnit> U;
  std::vector<StageLaunch> S;
  build_stage_table(C, include_diag != 0, cta_slots, n_blocks, U, S);
  for (size_t k = 0; units && k < U.size() && (int64_t)k < cap_units; ++k) {
    units[4 * k + 0] = U[k].slot0;
    units[4 * k + 1] = U[k].col;
    units[4 * k + 2] = U[k].j0;
    units[4 * k + 3] = U[k].count;
  }
  for (size_t l = 0; launches && l < S.size() && (int64_t)l < cap_launches; ++l) {
    const int64_t row[6] = {S[l].col_lo, S[l].col_hi, S[l].unit_lo, S[l].unit_hi, S[l].slot_lo, S[l].slot_hi};
    std::memcpy(launches + 6 * l, row, sizeof(row));
  }
  if (n_launches) *n_launches = (int64_t)S.size();
  return (int64_t)U.size();
}

int icikt_launch_shape(int64_t n, int32_t tier, int32_t n_sm, int32_t complete_obs, int32_t* out) {
  if (n < 1 || n > tiled_max_n() || tier < 0 || tier > 2 || n_sm < 1 || !out)
    return fail(ICIKT_ERR_BAD_ARG, "bad launch shape arguments");
  const int64_t nwords = ((n + 31) & ~31LL) / 32, wstride = (nwords + 3) & ~3LL;
  const TiledShape sh = tiled_shape(n, tier, wstride, n_sm, 0, complete_obs == 0);
  const int cap = sh.warps * sh.kk * 256;
  out[0] = sh.warps;
  out[1] = sh.kk;
  out[2] = sh.region_bytes;
  out[3] = sh.inplace_kk != 0 ? 1 : (sh.gmem ? 2 : 0);
  out[4] = cap;
  // in place: rows of the other column's rank table one staging part holds (the counter area behind the sequence)
  out[5] = sh.inplace_kk != 0 ? (int)std::min<int64_t>((n + 63) & ~63LL, ((sh.region_bytes - 2 * cap) >> 1) & ~63) : 0;
  return ICIKT_OK;
}

int icikt_measure_smem_bandwidth(int32_t device, double* g32, double* g128) {
  int rc = select_device(device);
  if (rc != ICIKT_OK) return rc;
  double a = 0, b = 0;
  if (measure_smem_bandwidth(&a, &b) < 0) return launch_fail("smem bandwidth kernel");
  if (g32) *g32 = a;
  if (g128) *g128 = b;
  return ICIKT_OK;
}

int icikt_measure_issue_rate(int32_t device, double* alu, double* fma, double* mixed) {
  int rc = select_device(device);
  if (rc != ICIKT_OK) return rc;
  double a = 0, f = 0, m = 0;
  if (measure_issue_rate(&a, &f, &m) < 0) return launch_fail("issue-rate kernel");
  if (alu) *alu = a;
  if (fma) *fma = f;
  if (mixed) *mixed = m;
  return ICIKT_OK;
}

int icikt_all_pairs(const double* data, int64_t n, int64_t C, int64_t ld, const double* global_na,
                    int32_t n_global_na, const icikt_opts* opts, double* raw, double* pvalue,
                    double* taumax, double* completeness, int32_t* status, int64_t* counts,
                    double* max_taumax, icikt_timings* timings) {
  return one_shot(data, n, C, ld, global_na, n_global_na, nullptr, nullptr, 0, opts, raw, pvalue, taumax,
                  completeness, status, counts, max_taumax, timings);
}

int icikt_matrices(const double* data, int64_t n, int64_t C, int64_t ld, const double* global_na,
                   int32_t n_global_na, const int32_t* pi, const int32_t* pj, int64_t P, const icikt_opts* opts,
                   int32_t scale_max, int32_t diag_good, const int32_t* n_good, double* cor, double* raw,
                   double* pvalue, double* taumax, double* completeness, int64_t* status_counts,
                   double* max_taumax, icikt_timings* timings) {
  if ((pi == nullptr) != (pj == nullptr)) return fail(ICIKT_ERR_BAD_ARG, "pi and pj must both be given");
  icikt_opts o;
  if (opts) o = *opts; else icikt_default_opts(&o);
  if (!pi && (o.pair_lo != 0 || o.pair_hi != 0))
    return fail(ICIKT_ERR_BAD_ARG, "matrix output needs the whole pair order on one device");
  if (!pi) o.include_diag = diag_good ? 0 : 1;  // setup_comparisons, R/kendalltau.R:191-194
  const MatrixOut mo{scale_max, diag_good, n_good, cor, status_counts};
  return one_shot(data, n, C, ld, global_na, n_global_na, pi, pj, P, &o, raw, pvalue, taumax, completeness,
                  nullptr, nullptr, max_taumax, timings, &mo);
}

int icikt_pairwise_completeness(const double* data, int64_t n, int64_t C, int64_t ld, const double* global_na,
                                int32_t n_global_na, int32_t device, const int32_t* pi, const int32_t* pj,
                                int64_t P, int32_t* missing, double* completeness, double* matrix) {
  if (!data || n < 1 || C < 1 || ld < n) return fail(ICIKT_ERR_BAD_ARG, "bad matrix arguments");
  if ((pi == nullptr) != (pj == nullptr)) return fail(ICIKT_ERR_BAD_ARG, "pi and pj must both be given");
  if (n_global_na < 0 || (n_global_na > 0 && !global_na)) return fail(ICIKT_ERR_BAD_ARG, "bad global_na");
  if (matrix && pi) return fail(ICIKT_ERR_BAD_ARG, "the matrix output covers all pairs; pass pi = pj = NULL");
  if (!pi) P = tri_pairs(C) + C;
  if (pi)
    for (int64_t k = 0; k < P; ++k)
      if (pi[k] < 0 || pi[k] >= C || pj[k] < 0 || pj[k] >= C) return fail(ICIKT_ERR_BAD_ARG, "pair index out of range");
  int rc = select_device(device);
  if (rc != ICIKT_OK) return rc;
  // R/utils.R:6-15: NA and Inf entries of global_na select classes, the rest are literals
  double lit[64];
  int nlit = 0, na_nan = 0, na_inf = 0;
  for (int i = 0; i < n_global_na; ++i) {
    const double v = global_na[i];
    if (std::isnan(v)) { na_nan = 1; continue; }
    if (std::isinf(v)) { na_inf = 1; continue; }
    if (nlit == 64) return fail(ICIKT_ERR_BAD_ARG, "more than 64 global_na literals");
    lit[nlit++] = v;
  }
  const int64_t words = (n + 31) / 32;
  const bool per_pair = missing || completeness;
  struct Buffers {
    double *data = nullptr, *lit = nullptr, *comp = nullptr, *mat = nullptr;
    uint32_t* bits = nullptr;
    int32_t *pi = nullptr, *pj = nullptr, *miss = nullptr;
    cudaStream_t stream = nullptr;
    ~Buffers() {
      cudaFree(data); cudaFree(lit); cudaFree(comp); cudaFree(mat); cudaFree(bits);
      cudaFree(pi); cudaFree(pj); cudaFree(miss);
      if (stream) cudaStreamDestroy(stream);
    }
  } b;
  CK(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
  CK(dmalloc(&b.data, (size_t)n * C));
  CK(dmalloc(&b.lit, 64));
  CK(dmalloc(&b.bits, (size_t)words * C));
  CK(cudaMemcpy2DAsync(b.data, sizeof(double) * n, data, sizeof(double) * ld, sizeof(double) * n, (size_t)C,
                       cudaMemcpyHostToDevice, b.stream));
  if (nlit) CK(cudaMemcpyAsync(b.lit, lit, sizeof(double) * nlit, cudaMemcpyHostToDevice, b.stream));
  if (launch_missing_bits(b.data, n, n, C, b.lit, nlit, na_nan, na_inf, b.bits, words, b.stream) < 0)
    return launch_fail("missing-mask kernel");
  if (per_pair) {
    if (pi) {
      CK(dmalloc(&b.pi, (size_t)P));
      CK(dmalloc(&b.pj, (size_t)P));
      CK(cudaMemcpyAsync(b.pi, pi, sizeof(int32_t) * P, cudaMemcpyHostToDevice, b.stream));
      CK(cudaMemcpyAsync(b.pj, pj, sizeof(int32_t) * P, cudaMemcpyHostToDevice, b.stream));
    }
    if (missing) CK(dmalloc(&b.miss, (size_t)P));
    if (completeness) CK(dmalloc(&b.comp, (size_t)P));
    if (launch_pair_missing(b.bits, words, n, C, b.pi, b.pj, P, b.miss, b.comp, b.stream) < 0)
      return launch_fail("pair completeness kernel");
    if (missing) CK(cudaMemcpyAsync(missing, b.miss, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, b.stream));
    if (completeness) CK(cudaMemcpyAsync(completeness, b.comp, sizeof(double) * P, cudaMemcpyDeviceToHost, b.stream));
  }
  if (matrix) {
    CK(dmalloc(&b.mat, (size_t)C * C));
    if (launch_missing_matrix(b.bits, words, n, C, b.mat, b.stream) < 0)
      return launch_fail("completeness matrix kernel");
    CK(cudaMemcpyAsync(matrix, b.mat, sizeof(double) * C * C, cudaMemcpyDeviceToHost, b.stream));
  }
  CK(cudaStreamSynchronize(b.stream));
  return ICIKT_OK;
}

}  // extern "C"

namespace {

// One multi-device job: worker k (one host thread per device) owns the k-th contiguous slice of the
// pair order.  Sharded preprocessing: if every pair of devices can reach each other directly (NVLink /
// NVSwitch), device k uploads only its slice of the columns over its own PCIe link, runs K1 on that
// slice and pulls the other slices of the per-column TABLES from its peers (an all-gather by peer
// copies: K1 is not replicated and the raw matrix never travels between devices); otherwise every
// device uploads the whole matrix and preprocesses all columns.
struct MultiJob {
  const double* data;
  int64_t n, C, ld;
  const double* global_na;
  int32_t n_global_na;
  icikt_opts base;
  const int32_t* devices;
  int n_devices;
  int64_t ptot;
  bool peers_ok = false;  // every pair of distinct ordinals has peer access
  bool gather = false;    // sharded K1 + peer gather of the tables
  HostBarrier barrier;
  std::vector<icikt_plan*> peer_plan;
  explicit MultiJob(int nd) : n_devices(nd), barrier(nd), peer_plan((size_t)nd, nullptr) {}
  int device_of(int k) const { return devices ? devices[k] : k; }
  int64_t pair_lo(int k) const { return ptot * k / n_devices; }  // slices differ by at most one pair
  int64_t col_lo(int k) const { return C * k / n_devices; }
};

void multi_probe_peers(MultiJob& J) {
  bool distinct = true;
  J.peers_ok = true;
  for (int a = 0; a < J.n_devices; ++a)
    for (int b = 0; b < J.n_devices; ++b) {
      if (a == b) continue;
      const int da = J.device_of(a), db = J.device_of(b);
      if (da == db) { distinct = false; continue; }
      int ok = 0;
      if (cudaDeviceCanAccessPeer(&ok, da, db) != cudaSuccess || !ok) J.peers_ok = false;
    }
  cudaGetLastError();
  J.gather = J.n_devices >= 2 && distinct && J.peers_ok && J.C >= J.n_devices && J.ptot >= J.n_devices &&
             !std::getenv("ICIKT_NO_PEER_GATHER");
}

int enable_peers(const MultiJob& J, int k) {
  for (int j = 0; j < J.n_devices; ++j) {
    if (J.device_of(j) == J.device_of(k)) continue;
    const cudaError_t e = cudaDeviceEnablePeerAccess(J.device_of(j), 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
    cudaGetLastError();
  }
  return ICIKT_OK;
}

// Worker k up to and including the pair kernel + epilogue (results stay on its device).  In gather mode
// every worker passes both barriers whatever happened to it, so nobody is left waiting.
int multi_worker_pairs(MultiJob& J, int k, icikt_plan** out) {
  int rc = ICIKT_OK;
  icikt_opts o = J.base;
  o.device = J.device_of(k);
  o.pair_lo = J.pair_lo(k);
  o.pair_hi = J.pair_lo(k + 1);
  const int64_t cnt = o.pair_hi - o.pair_lo;
  // one cached plan per slot, like the one-shot entry points: repeated calls of the same shape do not
  // pay the device allocations again
  icikt_plan*& slot = g_cached_multi[k];
  icikt_plan* p = nullptr;
  if (cache_matches(slot, J.n, J.C, o, false) && (o.perspective != ICIKT_PERSPECTIVE_COMPLETE || slot->d_pw)) {
    p = slot;
    p->opts.perspective = o.perspective;
    p->opts.alternative = o.alternative;
    p->opts.continuity = o.continuity;
    p->opts.na_inf = o.na_inf;
  } else {
    if (slot) { icikt_plan_destroy(slot); slot = nullptr; }
    rc = icikt_plan_create(&p, J.n, J.C, nullptr, nullptr, 0, &o);
    if (rc == ICIKT_OK) slot = p;
  }
  *out = p;
  if (!J.gather) {
    if (cnt <= 0) return rc;
    if (rc == ICIKT_OK) rc = icikt_plan_upload(p, J.data, J.ld);
    if (rc == ICIKT_OK) rc = icikt_plan_columns(p, J.global_na, J.n_global_na);
  } else {
    if (rc == ICIKT_OK) rc = upload_columns(p, J.data, J.ld, J.col_lo(k), J.col_lo(k + 1));
    if (rc == ICIKT_OK) rc = icikt_plan_columns_range(p, J.global_na, J.n_global_na, J.col_lo(k), J.col_lo(k + 1));
    if (rc == ICIKT_OK && cudaStreamSynchronize(p->stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "column slice");
    J.peer_plan[(size_t)k] = (rc == ICIKT_OK) ? p : nullptr;
    J.barrier.wait();  // every slice of the tables is complete on its device
    bool all_ok = true;
    for (int j = 0; j < J.n_devices; ++j) all_ok = all_ok && J.peer_plan[(size_t)j] != nullptr;
    if (rc == ICIKT_OK && !all_ok) rc = fail(ICIKT_ERR_CUDA, "a peer device failed to preprocess its slice");
    if (rc == ICIKT_OK) rc = enable_peers(J, k);
    if (rc == ICIKT_OK) {
      icikt_table mine[ICIKT_MAX_TABLES], theirs[ICIKT_MAX_TABLES];
      const int nt = icikt_plan_tables(p, mine, ICIKT_MAX_TABLES);
      for (int j = 0; j < J.n_devices && rc == ICIKT_OK; ++j) {
        if (j == k) continue;
        icikt_plan_tables(J.peer_plan[(size_t)j], theirs, ICIKT_MAX_TABLES);
        for (int t = 0; t < nt && rc == ICIKT_OK; ++t) {
          const size_t off = (size_t)J.col_lo(j) * (size_t)mine[t].bytes_per_column;
          const size_t len = (size_t)(J.col_lo(j + 1) - J.col_lo(j)) * (size_t)mine[t].bytes_per_column;
          if (len && cudaMemcpyPeerAsync(static_cast<unsigned char*>(mine[t].ptr) + off, o.device,
                                         static_cast<const unsigned char*>(theirs[t].ptr) + off, J.device_of(j), len,
                                         p->stream) != cudaSuccess)
            rc = cuda_fail(cudaGetLastError(), "cudaMemcpyPeerAsync");
        }
      }
      if (rc == ICIKT_OK) rc = icikt_plan_columns_finish(p);
      if (rc == ICIKT_OK && cudaStreamSynchronize(p->stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "peer gather");
    }
    J.barrier.wait();  // nobody's tables are touched (or freed) before every pull has finished
    if (cnt <= 0) return rc;
  }
  if (rc == ICIKT_OK) rc = icikt_plan_pairs(p);
  return rc;
}

struct MultiWork {
  int rc = ICIKT_OK;
  std::string err;
  double mx = std::nan("");
  icikt_timings tm{};
  unsigned long long hist[16] = {};
};

int multi_collect(const std::vector<MultiWork>& work, double* max_taumax, icikt_timings* timings) {
  double mx = std::nan("");
  icikt_timings slowest{};
  for (const MultiWork& w : work) {
    if (w.rc != ICIKT_OK) return fail(w.rc, w.err);
    if (w.mx == w.mx && !(mx >= w.mx)) mx = w.mx;
    if (w.tm.total_ms >= slowest.total_ms) {
      const int launches = slowest.n_launches + w.tm.n_launches;
      slowest = w.tm;
      slowest.n_launches = launches;
    } else {
      slowest.n_launches += w.tm.n_launches;
    }
  }
  if (max_taumax) *max_taumax = mx;
  if (timings) *timings = slowest;
  return ICIKT_OK;
}

}  // namespace

extern "C" {

int icikt_all_pairs_multi(const double* data, int64_t n, int64_t C, int64_t ld, const double* global_na,
                          int32_t n_global_na, const icikt_opts* opts, const int32_t* devices,
                          int32_t n_devices, double* raw, double* pvalue, double* taumax,
                          double* completeness, int32_t* status, int64_t* counts, double* max_taumax,
                          icikt_timings* timings) {
  if (!data || !raw) return fail(ICIKT_ERR_BAD_ARG, "data and raw must not be NULL");
  if (n_devices < 1 || n_devices > 64) return fail(ICIKT_ERR_BAD_ARG, "n_devices must be 1..64");
  if (n < 1 || C < 1) return fail(ICIKT_ERR_BAD_ARG, "n and C must be >= 1");
  MultiJob J(n_devices);
  J.data = data; J.n = n; J.C = C; J.ld = ld; J.global_na = global_na; J.n_global_na = n_global_na;
  J.devices = devices;
  if (opts) J.base = *opts; else icikt_default_opts(&J.base);
  J.base.want_counts = counts ? 1 : 0;
  J.ptot = tri_pairs(C) + (J.base.include_diag ? C : 0);
  std::lock_guard<std::mutex> lock(g_cache_mu);  // the cached plans are not re-entrant
  multi_probe_peers(J);
  std::vector<MultiWork> work((size_t)n_devices);
  std::vector<std::thread> threads;
  for (int k = 0; k < n_devices; ++k) {
    threads.emplace_back([&, k]() {
      MultiWork& w = work[(size_t)k];
      icikt_plan* p = nullptr;
      w.rc = multi_worker_pairs(J, k, &p);
      const int64_t lo = J.pair_lo(k), cnt = J.pair_lo(k + 1) - lo;
      if (w.rc == ICIKT_OK && cnt > 0)
        w.rc = icikt_plan_download(p, raw + lo, pvalue ? pvalue + lo : nullptr, taumax ? taumax + lo : nullptr,
                                   completeness ? completeness + lo : nullptr, status ? status + lo : nullptr,
                                   counts ? counts + lo * ICIKT_NCOUNTS : nullptr, &w.mx);
      if (w.rc == ICIKT_OK && cnt > 0) w.rc = icikt_plan_timings(p, &w.tm);
      if (w.rc != ICIKT_OK) {
        w.err = g_err;  // thread-local message of this worker
        if (g_cached_multi[k]) { icikt_plan_destroy(g_cached_multi[k]); g_cached_multi[k] = nullptr; }
      }
    });
  }
  for (auto& t : threads) t.join();
  return multi_collect(work, max_taumax, timings);
}

int icikt_matrices_multi(const double* data, int64_t n, int64_t C, int64_t ld, const double* global_na,
                         int32_t n_global_na, const icikt_opts* opts, const int32_t* devices, int32_t n_devices,
                         int32_t scale_max, int32_t diag_good, const int32_t* n_good, double* cor, double* raw,
                         double* pvalue, double* taumax, double* completeness, int64_t* status_counts,
                         double* max_taumax, icikt_timings* timings) {
  if (!data) return fail(ICIKT_ERR_BAD_ARG, "data must not be NULL");
  if (n_devices < 1 || n_devices > kMaxMatrixDevices) return fail(ICIKT_ERR_BAD_ARG, "n_devices must be 1..16");
  if (n < 1 || C < 1) return fail(ICIKT_ERR_BAD_ARG, "n and C must be >= 1");
  icikt_opts base;
  if (opts) base = *opts; else icikt_default_opts(&base);
  base.include_diag = diag_good ? 0 : 1;  // setup_comparisons, R/kendalltau.R:191-194
  base.pair_lo = base.pair_hi = 0;
  const int64_t ptot = tri_pairs(C) + (base.include_diag ? C : 0);
  MultiJob J(n_devices);
  J.data = data; J.n = n; J.C = C; J.ld = ld; J.global_na = global_na; J.n_global_na = n_global_na;
  J.devices = devices; J.base = base; J.base.want_counts = 0; J.ptot = ptot;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    multi_probe_peers(J);
  }
  if (n_devices == 1 || !J.peers_ok || ptot < n_devices || C < n_devices) {
    // one device, or devices that cannot read each other's results: the single-device path
    base.device = J.device_of(0);
    return icikt_matrices(data, n, C, ld, global_na, n_global_na, nullptr, nullptr, 0, &base, scale_max, diag_good, n_good,
                          cor, raw, pvalue, taumax, completeness, status_counts, max_taumax, timings);
  }
  std::lock_guard<std::mutex> lock(g_cache_mu);  // the cached plans are not re-entrant
  double* outs[5] = {cor, raw, pvalue, taumax, completeness};
  std::vector<MultiWork> work((size_t)n_devices);
  std::vector<int32_t> good((size_t)C, 0);
  if (n_good) std::copy(n_good, n_good + C, good.begin());
  std::vector<icikt_plan*> plans((size_t)n_devices, nullptr);
  std::vector<std::thread> threads;
  for (int k = 0; k < n_devices; ++k) {
    threads.emplace_back([&, k]() {
      MultiWork& w = work[(size_t)k];
      icikt_plan* p = nullptr;
      w.rc = multi_worker_pairs(J, k, &p);
      unsigned long long bits = 0;
      if (w.rc == ICIKT_OK) {
        if (cudaMemcpyAsync(&bits, p->d_maxbits, sizeof(bits), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess ||
            cudaStreamSynchronize(p->stream) != cudaSuccess)
          w.rc = cuda_fail(cudaGetLastError(), "pair kernel");
      }
      if (w.rc == ICIKT_OK && bits) std::memcpy(&w.mx, &bits, sizeof(bits));
      if (w.rc == ICIKT_OK && k == 0 && diag_good && !n_good) {  // n_good = n - missing rows (R/kendalltau.R:165)
        w.rc = icikt_plan_column_info(p, good.data());
        for (int64_t c = 0; c < C && w.rc == ICIKT_OK; ++c) good[(size_t)c] = (int32_t)n - good[(size_t)c];
      }
      plans[(size_t)k] = (w.rc == ICIKT_OK) ? p : nullptr;
      J.barrier.wait();  // every slice of the pair results is complete on its device
      bool all_ok = true;
      double mx = std::nan("");
      for (int j = 0; j < n_devices; ++j) {
        all_ok = all_ok && plans[(size_t)j] != nullptr;
        if (work[(size_t)j].mx == work[(size_t)j].mx && !(mx >= work[(size_t)j].mx)) mx = work[(size_t)j].mx;
      }
      if (w.rc == ICIKT_OK && !all_ok) w.rc = fail(ICIKT_ERR_CUDA, "a peer device failed to compute its pairs");
      if (w.rc == ICIKT_OK) w.rc = enable_peers(J, k);
      if (w.rc == ICIKT_OK) w.rc = [&]() -> int {
        const int64_t c_lo = J.col_lo(k), c_hi = J.col_lo(k + 1);
        const size_t blk = (size_t)C * (size_t)(c_hi - c_lo);
        CK(cudaSetDevice(p->device));
        if (p->mat_elems < 5 * blk) {
          cudaFree(p->d_mat);
          p->d_mat = nullptr;
          p->mat_elems = 0;
          if (dmalloc(&p->d_mat, 5 * blk) != cudaSuccess) {
            cudaGetLastError();
            return fail(ICIKT_ERR_ALLOC, "device allocation of the result matrix block failed");
          }
          p->mat_elems = 5 * blk;
        }
        if (!p->d_hist) CK(dmalloc(&p->d_hist, 16));
        if (!p->d_ngood) CK(dmalloc(&p->d_ngood, (size_t)C));
        {
          const int r3 = ensure_stage(p);
          if (r3 != ICIKT_OK) return r3;
        }
        CK(cudaEventRecord(p->ev[6], p->stream));
        CK(cudaMemsetAsync(p->d_hist, 0, 16 * sizeof(unsigned long long), p->stream));
        BlockFill f{};
        f.n_dev = n_devices;
        for (int j = 0; j < n_devices; ++j) {
          const icikt_plan* q = plans[(size_t)j];
          f.tau[j] = q->d_tau; f.pvalue[j] = q->d_p; f.taumax[j] = q->d_tm; f.completeness[j] = q->d_comp;
          f.status[j] = q->d_status;
          f.pair_lo[j] = J.pair_lo(j);
        }
        f.pair_lo[n_devices] = ptot;
        f.n = n; f.C = C; f.c_lo = c_lo; f.c_hi = c_hi;
        f.scale_max = scale_max != 0; f.diag_good = diag_good != 0;
        f.max_taumax = mx;
        f.n_good = p->d_ngood;
        int best = 1;
        if (diag_good) {
          CK(cudaMemcpyAsync(p->d_ngood, good.data(), sizeof(int32_t) * (size_t)C, cudaMemcpyHostToDevice, p->stream));
          best = 0;
          for (int64_t c = 0; c < C; ++c) best = std::max(best, good[(size_t)c]);
        }
        f.best_good = best;
        for (int q = 0; q < 5; ++q) f.m[q] = outs[q] ? p->d_mat + (size_t)q * blk : nullptr;
        f.hist = p->d_hist;
        if (launch_matrix_block_fill(f, p->stream) < 0) return launch_fail("matrix block fill kernel");
        for (int q = 0; q < 5; ++q)
          if (outs[q] && blk) {
            const int r2 = staged_copy_out(p, outs[q] + (size_t)c_lo * (size_t)C, f.m[q], sizeof(double) * blk, p->stream);
            if (r2 != ICIKT_OK) return r2;
          }
        CK(cudaMemcpyAsync(w.hist, p->d_hist, sizeof(w.hist), cudaMemcpyDeviceToHost, p->stream));
        CK(cudaEventRecord(p->ev[7], p->stream));
        CK(cudaStreamSynchronize(p->stream));
        return icikt_plan_timings(p, &w.tm);
      }();
      J.barrier.wait();  // nobody's results are freed before every block has been filled
      if (w.rc != ICIKT_OK) {
        w.err = g_err;
        if (g_cached_multi[k]) { icikt_plan_destroy(g_cached_multi[k]); g_cached_multi[k] = nullptr; }
      }
    });
  }
  for (auto& t : threads) t.join();
  const int rc = multi_collect(work, max_taumax, timings);
  if (rc != ICIKT_OK) return rc;
  if (status_counts) {
    int64_t bad = 0;
    for (int q = 1; q < ICIKT_NSTATUS; ++q) {
      status_counts[q] = 0;
      for (const MultiWork& w : work) status_counts[q] += (int64_t)w.hist[q];
      bad += status_counts[q];
    }
    status_counts[0] = ptot - bad;
  }
  return ICIKT_OK;
}

int icikt_pair_list(const double* data, int64_t n, int64_t C, int64_t ld, const double* global_na,
                    int32_t n_global_na, const int32_t* pi, const int32_t* pj, int64_t P,
                    const icikt_opts* opts, double* raw, double* pvalue, double* taumax,
                    double* completeness, int32_t* status, int64_t* counts, double* max_taumax,
                    icikt_timings* timings) {
  if (!pi || !pj) return fail(ICIKT_ERR_BAD_ARG, "pi and pj must not be NULL");
  return one_shot(data, n, C, ld, global_na, n_global_na, pi, pj, P, opts, raw, pvalue, taumax,
                  completeness, status, counts, max_taumax, timings);
}

int icikt_pnorm_device(const double* z, int64_t n, int32_t lower_tail, double* out, int32_t device) {
  if (!z || !out || n < 0) return fail(ICIKT_ERR_BAD_ARG, "bad pnorm arguments");
  int rc = select_device(device);
  if (rc != ICIKT_OK) return rc;
  if (n == 0) return ICIKT_OK;
  double *dz = nullptr, *dout = nullptr;
  CK(dmalloc(&dz, (size_t)n));
  cudaError_t e = dmalloc(&dout, (size_t)n);
  if (e != cudaSuccess) { cudaFree(dz); return cuda_fail(e, "cudaMalloc"); }
  e = cudaMemcpy(dz, z, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = (launch_pnorm(dz, n, lower_tail, dout, 0) < 0) ? cudaGetLastError() : cudaSuccess;
  if (e == cudaSuccess) e = cudaMemcpy(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaFree(dz);
  cudaFree(dout);
  if (e != cudaSuccess) return cuda_fail(e, "pnorm");
  return ICIKT_OK;
}

}  // extern "C"
