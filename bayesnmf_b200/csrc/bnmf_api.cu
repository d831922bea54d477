// C ABI of bayesnmf_b200 (include/bnmf.h): handle management, state I/O and the
// per-iteration launch sequence.  No torch, no Rcpp; built by nvcc for sm_100a.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bnmf.h"
#include "bnmf_init.cuh"
#include "bnmf_mh.cuh"
#include "bnmf_poisson.cuh"
#include "bnmf_tc.cuh"
#include "bnmf_state.h"

using namespace bnmf;

static thread_local char g_err[1024] = "";
static int fail(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return 1;
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail("%s:%d CUDA error %s (%s)", __FILE__, __LINE__, cudaGetErrorName(e_), cudaGetErrorString(e_)); } while (0)

// ---------------------------------------------------------------------------------
// NCCL through dlopen: the library itself has no link-time dependency on it.
// ---------------------------------------------------------------------------------
struct Id128 { char b[128]; };
namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value (128 bytes)*/ Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi g_nccl;
static int load_nccl() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) return fail("NCCL not found (dlopen libnccl.so.2): %s", dlerror());
#define LD(field, sym) *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, sym); if (!g_nccl.field) return fail("NCCL symbol %s missing", sym)
  LD(GetUniqueId, "ncclGetUniqueId");
  LD(CommInitRank, "ncclCommInitRank");
  LD(AllReduce, "ncclAllReduce");
  LD(AllGather, "ncclAllGather");
  LD(GroupStart, "ncclGroupStart");
  LD(GroupEnd, "ncclGroupEnd");
  LD(CommDestroy, "ncclCommDestroy");
  LD(GetErrorString, "ncclGetErrorString");
#undef LD
  return 0;
}
// a communicator shared by the handles of one process (chains / ranks of a BIC fan-out reuse it)
struct CommBox {
  void* comm = nullptr;
  // one-shot all-reduce over NVLink peer memory (k_reduce_partials): this rank's exchange buffer, the peers'
  // buffers mapped through CUDA IPC, the number of the last exchange
  bool xchg_ok = false; int device = 0;
  unsigned long long* xlocal = nullptr; unsigned long long* xbuf[XCHG_MAX_WORLD] = {}; void* xopened[XCHG_MAX_WORLD] = {};
  int slot_words = 0; unsigned long long seq = 0; int* xerr = nullptr;
  ~CommBox() {
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(device);
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (xopened[r]) cudaIpcCloseMemHandle(xopened[r]);
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);       // (a collective: every peer has stopped pulling by now)
    if (xlocal) cudaFree(xlocal);
    if (xerr) cudaFree(xerr);
    cudaSetDevice(prev);
  }
};
enum { NCCL_INT64 = 4, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3, NCCL_INT32 = 2, NCCL_INT8 = 0 };

// ---------------------------------------------------------------------------------
struct bnmf_handle {
  virtual ~bnmf_handle() {}
  virtual int set_hyper(const char* name, const double* v, int64_t rows, int64_t cols) = 0;
  virtual int set_state(const char* name, const double* v, int64_t len) = 0;
  virtual int get_state(const char* name, double* out, int64_t len) = 0;
  virtual int set_temps(const double* t, int64_t n) = 0;
  virtual int init_from_prior(uint32_t have, uint32_t have_prior, double* row) = 0;
  virtual int step(int n_iters, int converged, double* metrics, double* P_out, double* A_out) = 0;
  virtual int run(const bnmf_convergence_control* cc, int post_warmup, double* metrics_out, int64_t rows_cap,
                  double* map_out, int64_t checks_cap, bnmf_run_result* res) = 0;
  virtual int ring_count(int* c) = 0;
  virtual int get_sample(const char* name, int ago, double* out, int64_t len) = 0;
  virtual int get_map(int n_samples, double* P, double* E, double* A, int* n_match) = 0;
  virtual int assign(int n_samples, const double* ref, int n_ref, double ci, int* n_keep, int* keep, double* votes, int* asg,
                     double* mapc, double* lo, double* hi, int* n_match) = 0;
  virtual int get_ci(int n_samples, double plo, double phi, double* P_lo, double* P_hi, double* E_lo, double* E_hi, int* n_match) = 0;
  virtual int comm_init(const char* id, int rank, int world) = 0;
  virtual int comm_share(bnmf_handle* src) = 0;
  std::shared_ptr<CommBox> commbox; int world = 1, rank = 0;
  virtual int timing(double* total, double* iter, double* z, int64_t* launches) = 0;
  virtual int set_l2_flush(size_t bytes) = 0;
  virtual int sample_z(int iter, double* ms) = 0;
  virtual int profile(int converged, char* names, double* ms, int32_t* counts, int cap, int32_t* n_out) = 0;
};

template <typename T> static __global__ void k_cvt_in(const double* src, T* dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (T)src[i];
}
template <typename S> static __global__ void k_cvt_out(const S* src, double* dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)src[i];
}
static __global__ void k_cvt_in_i32(const double* src, int32_t* dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (int32_t)llrint(src[i]);
}
static __global__ void k_cvt_in_u64(const double* src, unsigned long long* dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (unsigned long long)llrint(src[i]);
}
// data constants of the Poisson likelihood / padded KL (R/utils.R:103, :467-471)
// genome-major copy of the counts (32 x 32 tiles through shared memory, both sides coalesced)
static __global__ void k_transpose_counts(const int32_t* __restrict__ in, int32_t* __restrict__ out, int K, long long G) {
  __shared__ int32_t tile[32][33];
  const long long g0 = (long long)blockIdx.x * 32;
  const int k0 = (int)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int k = k0 + (int)threadIdx.x; const long long g = g0 + j;
    tile[j][threadIdx.x] = (k < K && g < G) ? in[k + (long long)K * g] : 0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int k = k0 + j; const long long g = g0 + threadIdx.x;
    if (k < K && g < G) out[g + G * (long long)k] = tile[threadIdx.x][j];
  }
}
static __global__ void k_data_consts(const int32_t* Mi, long long n, double* out /*2*/) {
  __shared__ double sc[8];
  double ll = 0.0, kl = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double m = (double)Mi[i];
    ll -= lgamma(m + 1.0);
    double mp = m > 1e-6 ? m : 1e-6;
    kl += mp * log(mp);
  }
  double a = block_sum<256>(ll, sc);
  double b = block_sum<256>(kl, sc);
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = a; out[2 * blockIdx.x + 1] = b; }
}

// padded-KL constant of real-valued data: sum M' log M', M' = max(M, 1e-6)   (R/utils.R:467-471)
template <typename T> static __global__ void k_data_consts_real(const T* Mr, long long n, double* out) {
  __shared__ double sc[8];
  double kl = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double m = (double)Mr[i];
    const double mp = m > 1e-6 ? m : 1e-6;
    kl += mp * log(mp);
  }
  const double b = block_sum<256>(kl, sc);
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = 0.0; out[2 * blockIdx.x + 1] = b; }
}

// samples$P / samples$A of the iteration into row ctrl->row of the step's history buffers
template <typename T> static __global__ void k_hist_copy(Dev<T> d, T* P_hist, int32_t* A_hist) {
  const long long KN = (long long)d.K * d.N;
  const long long row = d.ctrl->row;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (P_hist && i < KN) P_hist[row * KN + i] = d.P[i];
  if (A_hist && i < d.N) A_hist[row * d.N + i] = d.A[i];
}

constexpr int ET = 512;   // threads per block of k_eside
enum StType { ST_T, ST_I32, ST_U64 };
struct StEntry { void* p; long long len; StType ty; };

// Process-wide pool of the small pinned buffers handles fetch their metric rows into: cudaMallocHost costs
// 0.5 - 2 ms apiece, more than the rest of bnmf_create for a shard; a closed handle leaves its buffer here.
static std::mutex g_hpin_mutex;
static std::vector<std::pair<void*, size_t>> g_hpin_pool;
static cudaError_t pinned_get(void** p, size_t bytes) {
  {
    std::lock_guard<std::mutex> lk(g_hpin_mutex);
    for (size_t i = 0; i < g_hpin_pool.size(); ++i)
      if (g_hpin_pool[i].second == bytes) { *p = g_hpin_pool[i].first; g_hpin_pool.erase(g_hpin_pool.begin() + i); return cudaSuccess; }
  }
  return cudaMallocHost(p, bytes);
}
static void pinned_put(void* p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_hpin_mutex);
  if (g_hpin_pool.size() < 64) g_hpin_pool.push_back({p, bytes}); else cudaFreeHost(p);
}
static void pinned_release_all() {
  std::lock_guard<std::mutex> lk(g_hpin_mutex);
  for (auto& b : g_hpin_pool) cudaFreeHost(b.first);
  g_hpin_pool.clear();
}

// process-wide pinned staging buffer for the upload of count matrices (bnmf_create)
static std::mutex g_pin_mutex;
static void* g_pin = nullptr;
static size_t g_pin_bytes = 0;

// Process-wide cache of device blocks handed back by destroyed handles.  bayesNMF() builds one sampler
// after another (a rank at a time), every one with the same few block sizes; cudaMalloc / cudaFree cost
// 1-20 ms apiece on a device that is in use, more than the upload of the counts.  Blocks are matched by
// (device, exact size); the cache holds at most BNMF_CACHE_MB (default 4096) and is emptied by
// bnmf_release_cached_memory() or when an allocation fails.
struct DevBlock { void* p; size_t bytes; int device; };
static std::mutex g_dev_mutex;
static std::vector<DevBlock> g_dev_cache;
static size_t g_dev_cached = 0;
static size_t dev_cache_limit() {
  static const size_t v = (size_t)(getenv("BNMF_CACHE_MB") ? std::max(0LL, atoll(getenv("BNMF_CACHE_MB"))) : 4096LL) << 20;
  return v;
}
static int release_cached_blocks(int device) {       // device < 0: all devices
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  int prev = 0; cudaGetDevice(&prev);
  size_t keep = 0;
  for (size_t i = 0; i < g_dev_cache.size(); ++i) {
    DevBlock& b = g_dev_cache[i];
    if (device >= 0 && b.device != device) { g_dev_cache[keep++] = b; continue; }
    cudaSetDevice(b.device); cudaFree(b.p); g_dev_cached -= b.bytes;
  }
  g_dev_cache.resize(keep);
  cudaSetDevice(prev);
  return 0;
}
static cudaError_t cached_malloc(void** p, size_t bytes, int device) {
  {
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    for (size_t i = 0; i < g_dev_cache.size(); ++i)
      if (g_dev_cache[i].device == device && g_dev_cache[i].bytes == bytes) {
        *p = g_dev_cache[i].p; g_dev_cached -= bytes;
        g_dev_cache.erase(g_dev_cache.begin() + i);          // (keeps the list oldest-first)
        return cudaSuccess;
      }
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) {      // make room: give the cached blocks back to the driver, try once more
    cudaGetLastError();
    release_cached_blocks(device);
    e = cudaMalloc(p, bytes);
  }
  return e;
}
// the caller has synchronised every stream that touched the block
static void cached_free(void* p, size_t bytes, int device) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    if (bytes <= dev_cache_limit()) {
      // make room by dropping the oldest blocks (sizes nobody asked for again)
      int prev = 0; cudaGetDevice(&prev);
      size_t drop = 0;
      while (g_dev_cached + bytes > dev_cache_limit() && drop < g_dev_cache.size()) {
        DevBlock& b = g_dev_cache[drop++];
        cudaSetDevice(b.device); cudaFree(b.p); g_dev_cached -= b.bytes;
      }
      if (drop) { g_dev_cache.erase(g_dev_cache.begin(), g_dev_cache.begin() + drop); cudaSetDevice(prev); }
      g_dev_cache.push_back(DevBlock{p, bytes, device}); g_dev_cached += bytes;
      return;
    }
  }
  cudaFree(p);
}
// Device scratch of one call (MAP, credible intervals, assignment): blocks in power-of-two sizes from the
// cache, handed back -- after the stream has drained -- when the call returns, error paths included.
struct Scratch {
  cudaStream_t stream; int device;
  std::vector<std::pair<void*, size_t>> v;
  Scratch(cudaStream_t s, int dev) : stream(s), device(dev) {}
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  template <typename X> cudaError_t get(X** p, size_t bytes) {
    size_t r = 4096; while (r < bytes) r <<= 1;
    void* q = nullptr;
    const cudaError_t e = cached_malloc(&q, r, device);
    if (e == cudaSuccess) { v.push_back({q, r}); *p = static_cast<X*>(q); }
    return e;
  }
  ~Scratch() {
    if (v.empty()) return;
    cudaStreamSynchronize(stream);
    for (auto& a : v) cached_free(a.first, a.second, device);
  }
};

template <typename T>
struct Sampler : bnmf_handle {
  bnmf_config cfg;
  Dev<T> d;
  cudaStream_t stream = nullptr;
  std::vector<std::pair<void*, size_t>> allocs;      // device blocks of this handle (returned to the process-wide cache)
  // device arrays below 64 MiB are carved out of 64 MiB slabs (zero-filled once): a handle owns
  // ~40 arrays, and cudaMalloc costs up to a millisecond apiece when the device is in use
  struct Slab { char* base; size_t cap, used; };
  std::vector<Slab> slabs;
  static constexpr size_t SLAB_BYTES = (size_t)64 << 20;
  std::map<std::string, StEntry> st;
  std::map<std::string, Hyper<T>*> hy;
  std::map<std::string, long long> hy_len;
  double* stage = nullptr; long long stage_len = 0;      // device double staging
  int* work_ctr = nullptr; int n_ktiles = 1; int KT = 96; int ZR = 32, ZR_B = 2, z_ctB = 0; bool z_sparse = false, z_split = false;
  double* red_slices = nullptr; unsigned* red_ticket = nullptr;
  int* nanflags = nullptr;
  T* P_hist = nullptr; int32_t* A_hist = nullptr;
  double* h_metrics = nullptr;                            // pinned
  int NP = 0; size_t z_smem = 0;
  void* comm = nullptr;
  long long* red_i64 = nullptr;                           // [K*N + N] packed int64 reduction buffer
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> zev, iev;
  void* flush_buf = nullptr; size_t flush_bytes = 0;
  double last_iter_ms = 0;
  bool time_z = true;
  double last_total_ms = 0, last_z_ms = 0; int64_t last_launches = 0;
  bool have_temps = false; int64_t temps_cap = 0;
  double h_data_sum = 0.0;
  std::vector<double> h_temps;                             // host copy of the temperature schedule (bnmf_run)
  std::vector<double> h_rows;                              // every sample_metrics row so far, MC_COLS each (bnmf_run's windows)
  struct RunState {                                        // check_convergence_'s part of self$state (bnmf_run)
    bool have_prev = false; double prev = 0.0, best = 0.0;
    int inarow_no_change = 0, inarow_no_best = 0, inarow_na = 0, best_iter = 0;
    bool converged = false; int why = 0, converged_iter = 0, post_done = 0;
  } rs;

  ~Sampler() override {
    cudaSetDevice(cfg.device);
    commbox.reset();
    if (stream) cudaStreamSynchronize(stream);
    if (side) cudaStreamSynchronize(side);
    for (auto& a : allocs) cached_free(a.first, a.second, cfg.device);
    if (h_metrics) pinned_put(h_metrics, (size_t)d.metrics_cap * MC_COLS * sizeof(double));
    for (auto e : zev) cudaEventDestroy(e);
    for (auto e : iev) cudaEventDestroy(e);
    if (flush_buf) cudaFree(flush_buf);
    if (gexec) cudaGraphExecDestroy(gexec);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (side) { cudaStreamSynchronize(side); cudaStreamDestroy(side); }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_join_p) cudaEventDestroy(ev_join_p);
    if (stream) cudaStreamDestroy(stream);
  }

  template <typename X> int dalloc(X** p, long long n) {
    if (n < 1) n = 1;
    const size_t bytes = (((size_t)n * sizeof(X)) + 255) & ~(size_t)255;
    if (bytes >= SLAB_BYTES) {
      const size_t rounded = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
      CK(cached_malloc((void**)p, rounded, cfg.device));
      CK(cudaMemsetAsync(*p, 0, bytes, stream));
      allocs.push_back({*p, rounded});
      return 0;
    }
    if (slabs.empty() || slabs.back().used + bytes > slabs.back().cap) {
      char* base;
      CK(cached_malloc((void**)&base, SLAB_BYTES, cfg.device));
      CK(cudaMemsetAsync(base, 0, SLAB_BYTES, stream));
      allocs.push_back({base, SLAB_BYTES});
      slabs.push_back(Slab{base, SLAB_BYTES, 0});
    }
    *p = reinterpret_cast<X*>(slabs.back().base + slabs.back().used);
    slabs.back().used += bytes;
    return 0;
  }
  int reg(const char* name, T** p, long long n) {
    if (dalloc(p, n)) return 1;
    st[name] = StEntry{*p, n, ST_T};
    return 0;
  }
  int ensure_stage(long long n) {
    if (n <= stage_len) return 0;
    const size_t rounded = ((size_t)n * sizeof(double) + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    if (stage) {
      CK(cudaStreamSynchronize(stream));
      for (auto& a : allocs) if (a.first == stage) { cached_free(a.first, a.second, cfg.device); a.first = nullptr; }
    }
    CK(cached_malloc((void**)&stage, rounded, cfg.device));
    allocs.push_back({stage, rounded});
    stage_len = (long long)(rounded / sizeof(double));
    return 0;
  }
  static int blocks(long long n, int t) { return (int)((n + t - 1) / t); }

  int create(const bnmf_config* c, const double* data) {
    const bool trace = getenv("BNMF_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!trace) return;
      auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[bnmf_create] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
    };
    cfg = *c;
    if (cfg.K < 1 || cfg.N < 1 || cfg.G < 1) return fail("bnmf_create: K, N, G must be >= 1");
    if (cfg.N > 64) return fail("bnmf_create: N = %d > 64 is not supported by this build", cfg.N);
    // cells are addressed with 64-bit indices, but the column count of a handle and the item / block counts
    // derived from it are ints: one handle takes at most 2^31 - 64 columns (shard larger data sets by genome)
    if (cfg.G > 2147483583LL) return fail("bnmf_create: G = %lld columns on one handle exceeds the supported 2^31 - 65", (long long)cfg.G);
    CK(cudaSetDevice(cfg.device));
    {
      int least = 0, greatest = 0;
      CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      CK(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, greatest));
      CK(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, least));
      CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ev_join_p, cudaEventDisableTiming));
    }
    CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
    lap("stream + events");
    memset(&d, 0, sizeof(d));
    const int K = cfg.K, N = cfg.N; const long long G = cfg.G;
    d.K = K; d.N = N; d.G = (int)G; d.G_total = cfg.G_total; d.g0 = cfg.g0;
    d.likelihood = cfg.likelihood; d.prior = cfg.prior; d.MH = cfg.MH;
    d.learning_rank = cfg.learning_rank; d.rank_method = cfg.rank_method; d.seed = cfg.seed;
    const long long KN = (long long)K * N, NG = (long long)N * G, KG = (long long)K * G;

    // parameters
    if (reg("P", &d.P, KN) || reg("E", &d.E, NG) || reg("sigmasq", &d.sigmasq, G)) return 1;
    if (dalloc(&d.A, N) || dalloc(&d.R, 1)) return 1;
    st["A"] = StEntry{d.A, N, ST_I32}; st["R"] = StEntry{d.R, 1, ST_I32};
    // prior parameters (all allocated: cheap relative to E, keeps the name table uniform)
    if (reg("Mu_p", &d.Mu_p, KN) || reg("Sigmasq_p", &d.Sigmasq_p, KN) || reg("Lambda_p", &d.Lambda_p, KN) ||
        reg("Alpha_p", &d.Alpha_p, KN) || reg("Beta_p", &d.Beta_p, KN)) return 1;
    if (cfg.prior == BNMF_TRUNCNORMAL) { if (reg("Mu_e", &d.Mu_e, NG) || reg("Sigmasq_e", &d.Sigmasq_e, NG)) return 1; }
    if (cfg.prior == BNMF_EXPONENTIAL) { if (reg("Lambda_e", &d.Lambda_e, NG)) return 1; }
    if (cfg.prior == BNMF_GAMMA) { if (reg("Alpha_e", &d.Alpha_e, NG) || reg("Beta_e", &d.Beta_e, NG)) return 1; }
    if (cfg.likelihood == BNMF_NORMAL) { if (reg("Alpha", &d.Alpha_g, G) || reg("Beta", &d.Beta_g, G)) return 1; }
    // hyperparameters: start as scalars (1 element), may be replaced by matrices
    struct HN { const char* n; Hyper<T>* h; long long len; };
    HN hn[] = {{"A_p", &d.A_p, KN}, {"B_p", &d.B_p, KN}, {"C_p", &d.C_p, KN}, {"D_p", &d.D_p, KN},
               {"M_p", &d.M_p, KN}, {"S_p", &d.S_p, KN}, {"A_e", &d.A_e, NG}, {"B_e", &d.B_e, NG},
               {"C_e", &d.C_e, NG}, {"D_e", &d.D_e, NG}, {"M_e", &d.M_e, NG}, {"S_e", &d.S_e, NG}};
    for (auto& h : hn) {
      T* p; if (dalloc(&p, 1)) return 1;
      h.h->p = p; h.h->is_matrix = 0; hy[h.n] = h.h; hy_len[h.n] = h.len;
    }
    // statistics / reductions
    if (dalloc(&red_i64, KN + N)) return 1;
    d.SP = (unsigned long long*)red_i64; d.rowsumE_fx = red_i64 + KN;
    st["SP"] = StEntry{d.SP, KN, ST_U64};
    if (dalloc(&d.SE, NG)) return 1;
    st["SE"] = StEntry{d.SE, NG, ST_I32};
    if (dalloc(&d.colsumP, N) || dalloc(&d.lp_P, 1) || dalloc(&d.pacc_sum, 1)) return 1;
    if (cfg.MH) { if (reg("P_acceptance_rate", &d.P_acc, KN) || reg("E_acceptance_rate", &d.E_acc, NG)) return 1; }
    if (cfg.MH || cfg.likelihood == BNMF_NORMAL || cfg.learning_rank) { if (reg("Mhat", &d.Mhat, KG)) return 1; }

    // work decomposition of the latent-count kernel: every warp owns its tables (nothing block-wide)
    NP = ((N + 3) / 4) * 4; if (NP > 32) NP = ((N + 7) / 8) * 8;
    const int zw = NP <= 32 ? 8 : 4;
    z_smem = zstat_smem_bytes<T>(NP, zw);
    KT = K; n_ktiles = 1;
    const int cts = (int)((G + 31) / 32);
    // rows per work item: an item costs ~12k cycles of fixed latency (E tile in, SE out), rows differ
    // widely in their counts: as coarse as leaves every resident warp (148 SMs x 16) about eight items
    // to balance with (measured at 12.5k .. 100k genomes)
    // (an item also pays ~3 us of a warp's time for its E tile and its SE flush, a row ~10 us at WGS-like counts.)
    // The last z_ctB column tiles are cut finer (ZR_B rows): the tail of a launch is one small item.
    ZR_B = 2; if (ZR_B > K) ZR_B = K;
    z_ctB = (int)std::min<long long>(cts / 2, (4LL * 148 * 16 + (K + ZR_B - 1) / ZR_B - 1) / ((K + ZR_B - 1) / ZR_B));
    ZR = 16;
    // (a 12,500-genome shard: 8 rows 0.193 ms, 4 rows 0.197, 16 rows 0.211; 25,000 genomes: 8 rows 0.344, 16 rows 0.350, 4 rows 0.363)
    {
      auto coarse_items = [&](int zr) { return (long long)(cts - z_ctB) * ((K + zr - 1) / zr); };
      if (coarse_items(16) < 4000LL) ZR = 8;
      while (ZR > 1 && coarse_items(ZR) < 2300LL) ZR >>= 1;
    }
    if (const char* e = getenv("BNMF_ZR")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) ZR = v; }   // tuning knobs
    if (const char* e = getenv("BNMF_ZR_B")) { const int v = atoi(e); if (v >= 1 && v <= 32) ZR_B = v; }
    if (const char* e = getenv("BNMF_Z_CTB")) { const int v = atoi(e); if (v >= 0 && v <= cts) z_ctB = v; }
    if (ZR > K) ZR = K;
    if (ZR_B > ZR) ZR_B = ZR;
    {
      const long long nz = (long long)(cts - z_ctB) * ((K + ZR - 1) / ZR) + (long long)z_ctB * ((K + ZR_B - 1) / ZR_B);
      if (nz > 2147483647LL - (1 << 20)) return fail("bnmf_create: %lld work items exceed the supported 2^31 - 1 (K = %d, G = %lld)", nz, K, (long long)G);
      d.n_zitems = (int)nz;
    }
    d.n_eblocks = (cfg.MH || cfg.likelihood == BNMF_NORMAL) ? (int)((G + 7) / 8) : blocks(NG, ET);
    if (dalloc(&d.zpart, (long long)d.n_zitems * PC_COLS + 2 * N) || dalloc(&d.epart, (long long)d.n_eblocks * PC_COLS) ||
        dalloc(&d.red, PC_COLS) || dalloc(&work_ctr, n_ktiles + 8) || dalloc(&nanflags, 5 * N)) return 1;
    if (dalloc(&d.ctrl, 1) || dalloc(&red_slices, RED_BLOCKS * PC_COLS) || dalloc(&red_ticket, 3)) return 1;
    d.metrics_cap = 256;
    if (dalloc(&d.metrics, (long long)d.metrics_cap * MC_COLS)) return 1;
    CK(pinned_get((void**)&h_metrics, (size_t)d.metrics_cap * MC_COLS * sizeof(double)));
    if (dalloc(&P_hist, (long long)d.metrics_cap * KN) || dalloc(&A_hist, (long long)d.metrics_cap * N)) return 1;
    { double* t; if (dalloc(&t, 1)) return 1; d.temps = t; d.n_temps = 0; }
    // ring
    d.ring_cap = cfg.ring_cap > 0 ? cfg.ring_cap : 0;
    if (d.ring_cap > 0) {
      if (dalloc(&d.ring_P, (long long)d.ring_cap * KN) || dalloc(&d.ring_E, (long long)d.ring_cap * NG) ||
          dalloc(&d.ring_A, (long long)d.ring_cap * N)) return 1;
    }
    lap("allocations");
    // data: one threaded pass over the caller's matrix (validation, column totals, the sum for
    // mean(data), conversion of counts to int32) in fixed chunks of columns, so that the result
    // does not depend on the number of threads
    const long long CH = 1024;                                   // columns per chunk
    const long long n_ch = (G + CH - 1) / CH;
    std::vector<long double> ch_sum((size_t)n_ch, 0.0L);
    std::vector<double> ch_row;                                   // [chunk][K] row totals (the order k_zstat visits the mutation types in)
    std::vector<long long> ch_bad((size_t)n_ch, -1);             // first offending cell of a chunk
    std::vector<int> ch_why((size_t)n_ch, 0);
    // counts go through a process-wide pinned staging buffer (kept between handles: a fresh 38 MB
    // vector costs more in page faults than the whole pass over the data)
    const bool pois = cfg.likelihood == BNMF_POISSON;
    if (pois) ch_row.assign((size_t)n_ch * K, 0.0);
    std::unique_lock<std::mutex> pin_lock(g_pin_mutex, std::defer_lock);
    int32_t* h_mi = nullptr;
    if (pois) {
      pin_lock.lock();
      const size_t need = (size_t)KG * sizeof(int32_t);
      if (g_pin_bytes < need) {
        if (g_pin) cudaFreeHost(g_pin);
        g_pin = nullptr; g_pin_bytes = 0;
        CK(cudaMallocHost(&g_pin, need));
        g_pin_bytes = need;
      }
      h_mi = static_cast<int32_t*>(g_pin);
    }
    // the device copy of the counts exists before the pass: every chunk is uploaded as soon as it is converted
    int32_t* mi = nullptr;
    if (pois) { if (dalloc(&mi, KG)) return 1; }
    std::vector<int> ch_cuda((size_t)n_ch, 0);
    {
      auto work = [&](long long c) {
        const long long g_lo = c * CH, g_hi = std::min(G, g_lo + CH);
        long double sum = 0.0L;
        for (long long g = g_lo; g < g_hi; ++g) {
          double colsum = 0.0;
          for (int k = 0; k < K; ++k) {
            const long long i = k + (long long)K * g;
            const double v = data[i];
            if (pois) {
              // counts: non-negative integers; a cell up to 2^24 and a genome (column) total below
              // 2^31 keep the quad offsets of k_zstat and the int32 margins SE exact
              if (!(v >= 0.0) || v != std::floor(v)) { if (ch_bad[c] < 0) { ch_bad[c] = i; ch_why[c] = 1; } continue; }
              if (v > 16777216.0) { if (ch_bad[c] < 0) { ch_bad[c] = i; ch_why[c] = 2; } continue; }
              h_mi[(size_t)i] = (int32_t)v;
              ch_row[(size_t)c * K + k] += v;
            }
            colsum += v;
          }
          if (pois && colsum >= 2147483648.0 && ch_bad[c] < 0) { ch_bad[c] = g; ch_why[c] = 3; }
          sum += (long double)colsum;
        }
        ch_sum[(size_t)c] = sum;
        if (pois && ch_bad[c] < 0) {
          const size_t off = (size_t)g_lo * K, cnt = (size_t)(g_hi - g_lo) * K;
          ch_cuda[(size_t)c] = (int)cudaMemcpyAsync(mi + off, h_mi + off, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
        }
      };
      unsigned nt = std::thread::hardware_concurrency();
      if (nt < 1) nt = 1;
      if (nt > 16) nt = 16;
      if ((long long)nt > n_ch) nt = (unsigned)n_ch;
      if (KG < (1 << 20)) nt = 1;
      std::atomic<long long> next(0);
      const int dev = cfg.device;
      auto loop = [&]() { cudaSetDevice(dev); for (long long c; (c = next.fetch_add(1)) < n_ch;) work(c); };
      std::vector<std::thread> th;
      for (unsigned t = 1; t < nt; ++t) th.emplace_back(loop);
      loop();
      for (auto& t : th) t.join();
    }
    for (long long c = 0; c < n_ch; ++c) {
      if (ch_bad[c] < 0) continue;
      const long long i = ch_bad[c];
      if (ch_why[c] == 1) return fail("bnmf_create: Poisson likelihood needs non-negative integer counts (data[%lld] = %g)", i, data[i]);
      if (ch_why[c] == 2) return fail("bnmf_create: count %g at data[%lld] exceeds the supported 2^24 per cell", data[i], i);
      return fail("bnmf_create: column %lld sums to 2^31 counts or more, the supported maximum is 2^31 - 1", i);
    }
    long double data_sum = 0.0L;
    for (long long c = 0; c < n_ch; ++c) data_sum += ch_sum[(size_t)c];
    h_data_sum = (double)data_sum;
    lap("host pass over data");
    if (pois) {
      for (long long c = 0; c < n_ch; ++c) if (ch_cuda[(size_t)c]) return fail("bnmf_create: upload of the counts: %s", cudaGetErrorString((cudaError_t)ch_cuda[(size_t)c]));
      d.Mi = mi;
      if (!cfg.MH) {   // k_zstat reads the counts genome-major
        int32_t* mt; if (dalloc(&mt, KG)) return 1;
        k_transpose_counts<<<dim3((unsigned)((G + 31) / 32), (unsigned)((K + 31) / 32)), dim3(32, 8), 0, stream>>>(mi, mt, K, (long long)G);
        d.Mt = mt;
      }
      {   // mutation types by descending total count (ties: by index): heavy rows first, the tail of a launch is light
        std::vector<double> rs((size_t)K, 0.0);
        for (long long c = 0; c < n_ch; ++c) for (int k = 0; k < K; ++k) rs[k] += ch_row[(size_t)c * K + k];
        std::vector<int32_t> ord((size_t)K);
        for (int k = 0; k < K; ++k) ord[k] = k;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return rs[a] > rs[b]; });
        int32_t* ko; if (dalloc(&ko, K)) return 1;
        CK(cudaMemcpyAsync(ko, ord.data(), sizeof(int32_t) * K, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));     // (`ord` is a local)
        d.korder = ko;
      }
      const int nb = 296;
      double* cpart; if (dalloc(&cpart, 2 * nb)) return 1;
      k_data_consts<<<nb, 256, 0, stream>>>(mi, KG, cpart);
      std::vector<double> hc(2 * nb);
      CK(cudaMemcpyAsync(hc.data(), cpart, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      double a = 0, b = 0;
      for (int i = 0; i < nb; ++i) { a += hc[2 * i]; b += hc[2 * i + 1]; }
      d.ll_const = a; d.ll_const_all = a; d.kl_const = b;
      pin_lock.unlock();                       // (the stream was synchronised above: the upload is done)
    } else {
      if (ensure_stage(KG)) return 1;
      CK(cudaMemcpyAsync(stage, data, (size_t)KG * sizeof(double), cudaMemcpyHostToDevice, stream));
      T* mr; if (dalloc(&mr, KG)) return 1;
      k_cvt_in<T><<<blocks(KG, 256), 256, 0, stream>>>(stage, mr, KG);
      d.Mr = mr;
      st["data"] = StEntry{mr, KG, ST_T};
      const int nb = 296;
      double* cpart; if (dalloc(&cpart, 2 * nb)) return 1;
      k_data_consts_real<T><<<nb, 256, 0, stream>>>(mr, KG, cpart);
      std::vector<double> hc(2 * nb);
      CK(cudaMemcpyAsync(hc.data(), cpart, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      double b = 0;
      for (int i = 0; i < nb; ++i) b += hc[2 * i + 1];
      d.ll_const = 0.0; d.ll_const_all = 0.0; d.kl_const = b;
    }
    lap("upload + data constants");
    // fewer than four counts per cell on average: the kernel variant whose sparse rows skip the share machinery
    z_sparse = (double)(data_sum / (long double)KG) < 4.0;
    if (const char* e = getenv("BNMF_Z_SPARSE")) z_sparse = atoi(e) != 0;
    // fewer work items than half the resident warps: the variant in which 2, 4 or 8 warps share an item
    z_split = !z_sparse && 2LL * d.n_zitems <= 148LL * 16;
    if (const char* e = getenv("BNMF_Z_SPLIT")) z_split = !z_sparse && atoi(e) > 1;
    if (cfg.likelihood == BNMF_POISSON && !cfg.MH) { if (z_config()) return 1; }
    if (mh_setup()) return 1;
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    lap("kernel attributes, sweep setup");
    // default hyperprior parameters (get_default_*_hyperprior_params_, R/setup.R:123-181);
    // bnmf_set_hyper overrides them.  mean(data) is over this handle's columns: sharded runs
    // pass the global defaults explicitly.
    {
      const double mean = (double)(data_sum / (long double)KG), dN = (double)N;
      struct DV { const char* n; double v; };
      std::vector<DV> dv;
      if (cfg.prior == BNMF_TRUNCNORMAL) dv = {{"M", 0.0}, {"S", std::sqrt(mean / dN)}, {"A", dN + 1.0}, {"B", std::sqrt(dN)}};
      else if (cfg.prior == BNMF_EXPONENTIAL) dv = {{"A", 10.0 * std::sqrt(dN)}, {"B", 10.0 * std::sqrt(mean)}};
      else dv = {{"A", 10.0 * std::sqrt(dN)}, {"B", 10.0}, {"C", 10.0 * std::sqrt(mean)}, {"D", 10.0}};
      for (auto& e : dv)
        for (const char* side : {"_p", "_e"}) {
          const std::string nm = std::string(e.n) + side;
          if (set_hyper(nm.c_str(), &e.v, 1, 1)) return 1;
        }
    }
    return 0;
  }

  // ---- k_zstat dispatch over the compile-time signature count -------------------
  template <int NPV> int z_launch_t(bool configure) {
    auto kern = z_sparse ? k_zstat<T, NPV, 1> : z_split ? k_zstat<T, NPV, 2> : k_zstat<T, NPV, 0>;
    if (configure) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)z_smem));
      return 0;
    }
    int dev_sms = 148;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, cfg.device);
    int per_sm = NPV <= 32 ? 2 : 1;
    if (const char* e = getenv("BNMF_Z_PER_SM")) { const int v = atoi(e); if (v >= 1 && v <= per_sm) per_sm = v; }   // occupancy experiments
    int bx = dev_sms * per_sm;
    const long long items = d.n_zitems;
    constexpr int ZW = ZWarps<NPV>::value;
    // fewer items than resident warps: 2, 4 or 8 warps share an item (the dense kernel only)
    int S = 1;
    if (z_split) {
      while (S < 8 && items * S * 2 <= (long long)bx * ZW) S *= 2;
      if (const char* e = getenv("BNMF_Z_SPLIT")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8) S = v; }
    }
    long long need = (items * S + ZW - 1) / ZW;
    if (bx > need) bx = (int)need;
    kern<<<bx, 32 * ZW, z_smem, stream>>>(d, make_zkeys(d.seed), ZR, ZR_B, (int)((d.G + 31) / 32) - z_ctB, (S == 8 ? 3 : S == 4 ? 2 : S == 2 ? 1 : 0), work_ctr); mark("k_zstat");
    return 0;
  }
  int z_dispatch(bool configure) {
    switch (NP) {
      case 4: return z_launch_t<4>(configure);
      case 8: return z_launch_t<8>(configure);
      case 12: return z_launch_t<12>(configure);
      case 16: return z_launch_t<16>(configure);
      case 20: return z_launch_t<20>(configure);
      case 24: return z_launch_t<24>(configure);
      case 28: return z_launch_t<28>(configure);
      case 32: return z_launch_t<32>(configure);
      case 40: return z_launch_t<40>(configure);
      case 48: return z_launch_t<48>(configure);
      case 56: return z_launch_t<56>(configure);
      case 64: return z_launch_t<64>(configure);
    }
    return fail("k_zstat: unsupported padded signature count %d", NP);
  }
  int z_config() {
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, cfg.device);
    if (z_smem > 227 * 1024) return fail("k_zstat: %zu bytes of shared memory needed (N %d)", z_smem, cfg.N);
    return z_dispatch(true);
  }

  // ---- state I/O -------------------------------------------------------------------
  int set_hyper(const char* name, const double* v, int64_t rows, int64_t cols) override {
    CK(cudaSetDevice(cfg.device));
    if (!strcmp(name, "alpha") || !strcmp(name, "beta")) {   // sigmasq prior scalars (R/bayesNMF_sampler.R:222-230)
      if (rows * cols != 1) return fail("bnmf_set_hyper: '%s' is a scalar", name);
      (name[0] == 'a' ? sig_alpha : sig_beta) = v[0];
      return 0;
    }
    auto it = hy.find(name);
    if (it == hy.end()) return fail("bnmf_set_hyper: unknown hyperparameter '%s'", name);
    long long n = (long long)rows * cols;
    long long full = hy_len[name];
    if (n != 1 && n != full) return fail("bnmf_set_hyper: '%s' must be scalar or have %lld elements (got %lld)", name, full, n);
    // a value of the same shape as the current one overwrites it in place (repeated updates do not grow the
    // handle; the captured graph keeps its pointer); a change of shape takes a new array from the slab
    const bool same_shape = (it->second->is_matrix != 0) == (n != 1);
    T* p = const_cast<T*>(it->second->p);
    if (n == 1 && same_shape) {
      // a scalar over a scalar (the defaults, the usual overrides): one stream-ordered 8-byte copy, no staging, no
      // synchronisation (a pageable source is staged by the driver before the call returns; kernels launched later
      // on this stream -- and, through the fork event, on the side stream -- see the new value)
      const T hv = (T)v[0];
      CK(cudaMemcpyAsync(p, &hv, sizeof(T), cudaMemcpyHostToDevice, stream));
      return 0;
    }
    if (!same_shape) { if (dalloc(&p, n)) return 1; }
    if (ensure_stage(n)) return 1;
    CK(cudaStreamSynchronize(stream));
    if (side) CK(cudaStreamSynchronize(side));
    CK(cudaMemcpyAsync(stage, v, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
    k_cvt_in<T><<<blocks(n, 256), 256, 0, stream>>>(stage, p, n);
    CK(cudaStreamSynchronize(stream));
    if (!same_shape) {
      it->second->p = p; it->second->is_matrix = n != 1;
      drop_graph();      // the captured kernels hold the old pointer
    }
    return 0;
  }
  int set_state(const char* name, const double* v, int64_t len) override {
    CK(cudaSetDevice(cfg.device));
    if (!strcmp(name, "rowsumE")) {          // the fixed-point row sums of E, e.g. summed over shards by the host
      if (len != cfg.N) return fail("bnmf_set_state: 'rowsumE' has %d elements", cfg.N);
      std::vector<long long> hfx(cfg.N);
      for (int i = 0; i < cfg.N; ++i) hfx[i] = llrint(v[i] * RS_FX);
      CK(cudaMemcpyAsync(d.rowsumE_fx, hfx.data(), sizeof(long long) * cfg.N, cudaMemcpyHostToDevice, stream));
      CK(cudaStreamSynchronize(stream));
      return 0;
    }
    auto it = st.find(name);
    if (it == st.end()) return fail("bnmf_set_state: unknown or unallocated state '%s' for this model", name);
    if (len != it->second.len) return fail("bnmf_set_state: '%s' has %lld elements, got %lld", name, it->second.len, (long long)len);
    if (ensure_stage(len)) return 1;
    CK(cudaMemcpyAsync(stage, v, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, stream));
    const int b = blocks(len, 256);
    if (it->second.ty == ST_T) k_cvt_in<T><<<b, 256, 0, stream>>>(stage, (T*)it->second.p, len);
    else if (it->second.ty == ST_I32) k_cvt_in_i32<<<b, 256, 0, stream>>>(stage, (int32_t*)it->second.p, len);
    else k_cvt_in_u64<<<b, 256, 0, stream>>>(stage, (unsigned long long*)it->second.p, len);
    CK(cudaStreamSynchronize(stream));
    if (!strcmp(name, "E")) { if (refresh_rowsumE()) return 1; }
    if (!strcmp(name, "P")) { if (refresh_colsumP()) return 1; }
    return 0;
  }
  int dev_to_host(const void* p, StType ty, long long len, double* out) {
    if (ensure_stage(len)) return 1;
    const int b = blocks(len, 256);
    if (ty == ST_T) k_cvt_out<T><<<b, 256, 0, stream>>>((const T*)p, stage, len);
    else if (ty == ST_I32) k_cvt_out<int32_t><<<b, 256, 0, stream>>>((const int32_t*)p, stage, len);
    else k_cvt_out<unsigned long long><<<b, 256, 0, stream>>>((const unsigned long long*)p, stage, len);
    const size_t bytes = (size_t)len * sizeof(double);
    if (bytes < ((size_t)1 << 20)) {
      CK(cudaMemcpyAsync(out, stage, bytes, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      return 0;
    }
    // large results (E, Mhat): DMA into the process-wide pinned buffer, then a threaded copy into the
    // caller's pageable array (its first-touch page faults spread over the threads) -- the driver's
    // own pageable path does both serially
    std::lock_guard<std::mutex> pin_lock(g_pin_mutex);
    if (g_pin_bytes < bytes) {
      if (g_pin) cudaFreeHost(g_pin);
      g_pin = nullptr; g_pin_bytes = 0;
      CK(cudaMallocHost(&g_pin, bytes));
      g_pin_bytes = bytes;
    }
    CK(cudaMemcpyAsync(g_pin, stage, bytes, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 8) nt = 8;
    const size_t per = ((bytes / nt + 4095) / 4096) * 4096;
    auto part = [&](unsigned t) {
      const size_t lo = (size_t)t * per, hi = std::min(bytes, lo + per);
      if (lo < hi) memcpy((char*)out + lo, (const char*)g_pin + lo, hi - lo);
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(part, t);
    part(0);
    for (auto& t : th) t.join();
    return 0;
  }
  int get_state(const char* name, double* out, int64_t len) override {
    CK(cudaSetDevice(cfg.device));
    if (!strcmp(name, "data_sum")) {          // sum of this handle's data (what mean(data) of a sharded matrix is made of)
      if (len != 1) return fail("bnmf_get_state: 'data_sum' is a scalar");
      out[0] = h_data_sum;
      return 0;
    }
    if (!strcmp(name, "rowsumE")) {
      if (len != cfg.N) return fail("bnmf_get_state: 'rowsumE' has %d elements", cfg.N);
      std::vector<long long> h(cfg.N);
      CK(cudaMemcpyAsync(h.data(), d.rowsumE_fx, sizeof(long long) * cfg.N, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      for (int i = 0; i < cfg.N; ++i) out[i] = (double)h[i] / RS_FX;
      return 0;
    }
    auto it = st.find(name);
    if (it == st.end()) return fail("bnmf_get_state: unknown or unallocated state '%s' for this model", name);
    if (len != it->second.len) return fail("bnmf_get_state: '%s' has %lld elements, got %lld", name, it->second.len, (long long)len);
    return dev_to_host(it->second.p, it->second.ty, len, out);
  }
  int set_temps(const double* t, int64_t n) override {
    CK(cudaSetDevice(cfg.device));
    if (n < 1) return fail("bnmf_set_temperature_schedule: empty schedule");
    const bool reuse = have_temps && n <= temps_cap;        // a schedule that fits the current array overwrites it
    double* p = const_cast<double*>(d.temps);
    if (!reuse) { if (dalloc(&p, n)) return 1; temps_cap = n; }
    CK(cudaStreamSynchronize(stream));
    CK(cudaMemcpyAsync(p, t, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    d.temps = p; d.n_temps = (int)n; have_temps = true;
    h_temps.assign(t, t + n);
    drop_graph();
    return 0;
  }

  // rowSums(E) / colSums(P) after a user-supplied state (fixed point for E)
  int refresh_rowsumE();
  int refresh_colsumP();

  // ---- cross-shard sum -----------------------------------------------------------
  int comm_init(const char* id, int rank_, int world_) override {
    CK(cudaSetDevice(cfg.device));
    if (sweep_model)
      return fail("bnmf_comm_init: genome sharding is built for the Poisson latent-count models; run Normal / MH models as independent chains, one per GPU");
    if (load_nccl()) return 1;
    Id128 uid; memcpy(uid.b, id, 128);
    auto box = std::make_shared<CommBox>();
    int r = g_nccl.CommInitRank(&box->comm, world_, uid, rank_);
    if (r) return fail("ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    commbox = box; comm = box->comm; world = world_; rank = rank_;
    box->device = cfg.device;
    if (setup_xchg()) return 1;
    return sync_data_consts();
  }
  // Exchange buffers of the one-shot all-reduce: allocate this rank's, hand its IPC handle round with NCCL, map
  // the peers'.  Every rank must end up with the same answer to "is the peer path usable?" (a rank that cannot map
  // a peer -- no NVLink / P2P between the two devices, both ranks in one process -- sends everyone back to NCCL).
  int setup_xchg() {
    CommBox& b = *commbox;
    static const bool off = getenv("BNMF_XCHG") && !strcmp(getenv("BNMF_XCHG"), "0");
    if (off || world < 2 || world > XCHG_MAX_WORLD) return 0;
    const long long nw = (long long)cfg.K * cfg.N + cfg.N + PC_COLS;
    b.slot_words = (int)std::max<long long>(4096, 2 * nw);
    const size_t bytes = ((size_t)XCHG_SLOTS * world * b.slot_words + (size_t)XCHG_SLOTS * world + 64) * sizeof(unsigned long long);
    int ok = 1;
    if (cudaMalloc((void**)&b.xlocal, bytes) != cudaSuccess || cudaMalloc((void**)&b.xerr, sizeof(int)) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    cudaIpcMemHandle_t mine; memset(&mine, 0, sizeof(mine));
    if (ok) {
      CK(cudaMemsetAsync(b.xlocal, 0, bytes, stream));
      CK(cudaMemsetAsync(b.xerr, 0, sizeof(int), stream));
      if (cudaIpcGetMemHandle(&mine, b.xlocal) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    char* dsend; char* drecv; int* dflag;
    if (dalloc(&dsend, 64) || dalloc(&drecv, 64 * world) || dalloc(&dflag, 1)) return 1;
    CK(cudaMemcpyAsync(dsend, &mine, 64, cudaMemcpyHostToDevice, stream));
    int r = g_nccl.AllGather(dsend, drecv, 64, NCCL_INT8, comm, stream);
    if (r) return fail("ncclAllGather(ipc handles): %s", g_nccl.GetErrorString(r));
    std::vector<cudaIpcMemHandle_t> all(world);
    CK(cudaMemcpyAsync(all.data(), drecv, 64 * (size_t)world, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int p = 0; p < world && ok; ++p) {
      if (p == rank) { b.xbuf[p] = b.xlocal; continue; }
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      b.xopened[p] = ptr; b.xbuf[p] = static_cast<unsigned long long*>(ptr);
    }
    CK(cudaMemcpyAsync(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice, stream));
    if (allreduce_buf(dflag, 1, NCCL_INT32, NCCL_MIN)) return 1;
    int agreed = 0;
    CK(cudaMemcpyAsync(&agreed, dflag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    b.xchg_ok = agreed == 1;
    if (getenv("BNMF_TRACE")) fprintf(stderr, "[bnmf_comm_init] rank %d of %d: one-shot peer-memory all-reduce %s\n", rank, world, b.xchg_ok ? "on" : "off (NCCL)");
    return 0;
  }
  bool xchg_usable() const {
    return world > 1 && commbox && commbox->xchg_ok && (long long)cfg.K * cfg.N + cfg.N + PC_COLS <= commbox->slot_words;
  }
  // -sum lgamma(M+1) over every shard, the same bits on every rank (the rank learner compares a
  // replicated uniform with a probability formed from it: R/sample_params.R:115-165)
  int sync_data_consts() {
    if (world <= 1) return 0;
    double* buf; if (dalloc(&buf, 1)) return 1;
    CK(cudaMemcpyAsync(buf, &d.ll_const, sizeof(double), cudaMemcpyHostToDevice, stream));
    if (allreduce_buf(buf, 1, NCCL_FLOAT64, NCCL_SUM)) return 1;
    double v = 0.0;
    CK(cudaMemcpyAsync(&v, buf, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    d.ll_const_all = v;
    drop_graph();
    return 0;
  }
  int comm_share(bnmf_handle* src) override {
    if (sweep_model)
      return fail("bnmf_comm_share: genome sharding is built for the Poisson latent-count models");
    if (!src->commbox) return fail("bnmf_comm_share: the source handle has no communicator");
    commbox = src->commbox; comm = commbox->comm; world = src->world; rank = src->rank;
    return sync_data_consts();
  }
  int allreduce_stats() {   // SP + rowsumE_fx (exact int64 sums)
    if (world <= 1) return 0;
    int r = g_nccl.AllReduce(red_i64, red_i64, (size_t)cfg.K * cfg.N + cfg.N, NCCL_INT64, NCCL_SUM, comm, stream);
    if (r) return fail("ncclAllReduce(stats): %s", g_nccl.GetErrorString(r));
    return 0;
  }
  int allreduce_red() {     // metric partials
    if (world <= 1) return 0;
    int r = g_nccl.AllReduce(d.red, d.red, PC_COLS, NCCL_FLOAT64, NCCL_SUM, comm, stream);
    if (r) return fail("ncclAllReduce(metrics): %s", g_nccl.GetErrorString(r));
    return 0;
  }
  int allreduce_buf(void* p, size_t n, int dtype, int op) {
    if (world <= 1) return 0;
    int r = g_nccl.AllReduce(p, p, n, dtype, op, comm, stream);
    if (r) return fail("ncclAllReduce: %s", g_nccl.GetErrorString(r));
    return 0;
  }

  // ---- per-kernel timing of one iteration (bnmf_profile_iteration) ---------------------
  // mark(name) follows every kernel launch of an iteration: with profiling on it records an event on the
  // launching stream, so that the span since the previous mark is that kernel's device time (the stream is
  // in order and the host runs ahead of it)
  bool prof_on = false;
  std::vector<std::pair<const char*, cudaEvent_t>> prof_ev;
  std::vector<cudaEvent_t> prof_pool;
  void mark(const char* name) {
    if (!prof_on) return;
    cudaEvent_t e;
    if (!prof_pool.empty()) { e = prof_pool.back(); prof_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, stream);
    prof_ev.push_back({name, e});
  }
  int profile(int converged, char* names, double* ms, int32_t* counts, int cap, int32_t* n_out) override {
    CK(cudaSetDevice(cfg.device));
    h_converged = converged ? 1 : 0;
    int two[2] = {converged, -1};
    CK(cudaMemcpyAsync(&d.ctrl->converged, two, sizeof(two), cudaMemcpyHostToDevice, stream));
    if (overlap_allowed()) {
      Ctrl hc; CK(cudaMemcpyAsync(&hc, d.ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      h_iter = hc.iter;
    }
    hyper_ready = false; spec_next = false; ++h_iter;
    if (flush_bytes) CK(cudaMemsetAsync(flush_buf, 0, flush_bytes, stream));
    CK(cudaStreamSynchronize(stream));
    prof_on = true; prof_ev.clear();
    mark("(start)");
    const int rc = launch_iteration(false, false, nullptr, nullptr);
    prof_on = false;
    if (rc) return 1;
    CK(cudaMemcpyAsync(h_metrics, d.metrics, sizeof(double) * MC_COLS, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    h_rows.insert(h_rows.end(), h_metrics, h_metrics + MC_COLS);   // the chain has advanced by one iteration
    std::vector<std::string> nm; std::vector<double> tt; std::vector<int> cc;
    for (size_t i = 1; i < prof_ev.size(); ++i) {
      float t = 0; CK(cudaEventElapsedTime(&t, prof_ev[i - 1].second, prof_ev[i].second));
      size_t j = 0; while (j < nm.size() && nm[j] != prof_ev[i].first) ++j;
      if (j == nm.size()) { nm.push_back(prof_ev[i].first); tt.push_back(0.0); cc.push_back(0); }
      tt[j] += t; cc[j] += 1;
    }
    for (auto& pe : prof_ev) prof_pool.push_back(pe.second);
    prof_ev.clear();
    const int n = (int)std::min<size_t>(nm.size(), (size_t)(cap < 0 ? 0 : cap));
    for (int j = 0; j < n; ++j) {
      if (names) { snprintf(names + 32 * (size_t)j, 32, "%s", nm[j].c_str()); }
      if (ms) ms[j] = tt[j];
      if (counts) counts[j] = cc[j];
    }
    if (n_out) *n_out = (int32_t)nm.size();
    return 0;
  }

  // ---- the iteration ---------------------------------------------------------------
  int launches = 0;
  // Overlap of the E side's hyper-draws of iteration t+1 with k_zstat of iteration t (k_eside_hyper):
  // a low-priority side stream, fork / join events, the host's copy of the iteration counter.
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join_p = nullptr;
  double* alpha_retry = nullptr; int* alpha_n_retry = nullptr; int alpha_cap = 0;   // parked Alpha_e envelopes (k_eside_hyper -> k_alpha_retry)
  bool hyper_ready = false, spec_next = false;
  int h_iter = 0;
  unsigned long long* sides_pub = nullptr; unsigned long long sides_seq = 0;     // k_sides: "colSums(P) are out" counter
  // k_sides on small shards -- at most two waves of its blocks (512 threads at 64 registers: two per SM): measured on a 12,500-genome
  // shard 261.7 -> 257.2 us per iteration, at 25,000 genomes 448.3 -> 447.7, at 100,000 (seven waves; the P-side
  // blocks share their SMs with E-side blocks and the first wave waits for them) 1,580 -> 1,592: not used there.
  // BNMF_SIDES = "0": never, "1": always.
  bool sides_ok() const {
    const char* e = getenv("BNMF_SIDES");
    if (e && !strcmp(e, "0")) return false;
    if (e && !strcmp(e, "1")) return true;
    return (long long)cfg.N + d.n_eblocks <= 4LL * n_sms;
  }
  int n_sms = 148;
  bool overlap_allowed() const {
    static const bool off = getenv("BNMF_OVERLAP") && !strcmp(getenv("BNMF_OVERLAP"), "0");
    return !off && side != nullptr && cfg.prior == BNMF_GAMMA && cfg.likelihood == BNMF_POISSON && !cfg.MH && !graphs_allowed();
  }
  // within-iteration fork of the two sides' hyper-draws (graph-replayed problems, gamma prior); the list of parked
  // Alpha_e cells is allocated by ensure_graph() before the capture starts
  bool fork_hyper() const {
    static const bool off = getenv("BNMF_FORK_HYPER") && !strcmp(getenv("BNMF_FORK_HYPER"), "0");
    return !off && side != nullptr && alpha_retry != nullptr && !prof_on && graphs_allowed() &&
           cfg.prior == BNMF_GAMMA && cfg.likelihood == BNMF_POISSON && !cfg.MH;
  }
  int mh_setup();
  int mh_iteration(int from_prior, uint32_t have);
  int poisson_iteration(int from_prior, uint32_t have, cudaEvent_t z0, cudaEvent_t z1, bool fold_begin = false) {
    // fold_begin: k_pside is also k_begin_iter (see the kernel)
    int* const bc = fold_begin ? work_ctr : nullptr; const int bn = n_ktiles; unsigned* const bt = red_ticket + 2;
    const int keepP = (have & BNMF_HAVE_P) ? 1 : 0, keepE = (have & BNMF_HAVE_E) ? 1 : 0;
    // instantiated per (prior, prior draw or not): halves the code each launch has to fetch
    const int var = (cfg.prior == BNMF_GAMMA ? 2 : 0) | (from_prior ? 1 : 0);
    switch (var) {
      case 0: k_pside<T, 128, PRIOR_EXPONENTIAL, 0><<<cfg.N, 128, 0, stream>>>(d, keepP, bc, bn, bt); mark("k_pside");
              k_eside<T, ET, PRIOR_EXPONENTIAL, 0><<<d.n_eblocks, ET, 0, stream>>>(d, keepE); mark("k_eside"); break;
      case 1: k_pside<T, 128, PRIOR_EXPONENTIAL, 1><<<cfg.N, 128, 0, stream>>>(d, keepP); mark("k_pside");
              k_eside<T, ET, PRIOR_EXPONENTIAL, 1><<<d.n_eblocks, ET, 0, stream>>>(d, keepE); mark("k_eside"); break;
      case 2: if (!hyper_ready && fork_hyper()) {
                // Small problems (the iteration replayed as a CUDA graph, no overlap across iterations): the hyper-draws of
                // the two sides read nothing of each other -- the E side's run on the side stream (a parallel branch of
                // the graph) while the P side's and the draw of P run here; the draw of E joins them.  The kernels take
                // the iteration number from the device's counter (a graph freezes kernel arguments).
                CK(cudaEventRecord(ev_fork, stream));
                CK(cudaStreamWaitEvent(side, ev_fork, 0));
                const long long ncell = (long long)cfg.N * cfg.G;
                CK(cudaMemsetAsync(alpha_n_retry, 0, sizeof(int), side));
                k_eside_hyper<T, 256><<<blocks(ncell, 256), 256, 0, side>>>(d, -1, alpha_retry, alpha_n_retry, alpha_cap);
                k_alpha_retry<T><<<blocks((alpha_cap + ALPHA_RETRY_PER_WARP - 1) / ALPHA_RETRY_PER_WARP, 8), 256, 0, side>>>(d, -1, alpha_retry, alpha_n_retry, alpha_cap);
                CK(cudaEventRecord(ev_join, side));
                k_pside_hyper<T, 128><<<cfg.N, 128, 0, stream>>>(d, -1); mark("k_pside_hyper");
                k_pside<T, 128, PRIOR_GAMMA, 0, 1><<<cfg.N, 128, 0, stream>>>(d, keepP); mark("k_pside");
                CK(cudaStreamWaitEvent(stream, ev_join, 0));
                k_eside<T, ET, PRIOR_GAMMA, 0, 1><<<d.n_eblocks, ET, 0, stream>>>(d, keepE); mark("k_eside");
                launches += 3;
                break;
              }
              if (hyper_ready && fold_begin && sides_ok() && !keepP && !keepE) {
                // steady state: both sides' hyper-draws were made under the previous k_zstat -- the P side, the E side
                // and k_begin_iter are one launch (k_sides)
                CK(cudaStreamWaitEvent(stream, ev_join_p, 0));
                CK(cudaStreamWaitEvent(stream, ev_join, 0));
                if (!sides_pub) { if (dalloc(&sides_pub, 1)) return 1; sides_seq = 0; }
                ++sides_seq;
                k_sides<T, ET><<<cfg.N + d.n_eblocks, ET, 0, stream>>>(d, h_iter, bc, bn, bt, sides_pub, (unsigned long long)cfg.N * sides_seq); mark("k_sides");
                hyper_ready = false;
                --launches;               // (one launch where the count below expects two)
              } else {
              if (hyper_ready) {          // Beta_p / Alpha_p of this iteration were drawn under the previous k_zstat
                CK(cudaStreamWaitEvent(stream, ev_join_p, 0));
                k_pside<T, 128, PRIOR_GAMMA, 0, 1><<<cfg.N, 128, 0, stream>>>(d, keepP, bc, bn, bt); mark("k_pside");
              } else k_pside<T, 128, PRIOR_GAMMA, 0><<<cfg.N, 128, 0, stream>>>(d, keepP, bc, bn, bt); mark("k_pside");
              if (hyper_ready) {          // ... and so were Beta_e / Alpha_e
                CK(cudaStreamWaitEvent(stream, ev_join, 0));
                k_eside<T, ET, PRIOR_GAMMA, 0, 1><<<d.n_eblocks, ET, 0, stream>>>(d, keepE); mark("k_eside");
                hyper_ready = false;
              } else k_eside<T, ET, PRIOR_GAMMA, 0><<<d.n_eblocks, ET, 0, stream>>>(d, keepE); mark("k_eside");
              }
              if (spec_next && overlap_allowed()) {
                if (!alpha_retry) {      // list of parked Alpha_e cells, a quarter of the cells long: 16 % are rejected once, an
                                         // overflowing cell finishes in place (allocated -- zero-filled on `stream` -- before the fork)
                  const long long ncell = (long long)cfg.N * cfg.G;
                  alpha_cap = (int)std::max<long long>(256, std::min<long long>(ncell / 4 + 1024, 1LL << 30));
                  if (const char* e = getenv("BNMF_ALPHA_CAP")) { const int v = atoi(e); if (v >= 1) alpha_cap = v; }   // test knob (overflow path)
                  if (dalloc(&alpha_retry, (long long)BNMF_ALPHA_ENV_COLS * alpha_cap) || dalloc(&alpha_n_retry, 1)) return 1;
                }
                CK(cudaEventRecord(ev_fork, stream));
                CK(cudaStreamWaitEvent(side, ev_fork, 0));
                k_pside_hyper<T, 128><<<cfg.N, 128, 0, side>>>(d, h_iter + 1);
                CK(cudaEventRecord(ev_join_p, side));
                {
                  static const int ht = getenv("BNMF_HYPER_THREADS") ? atoi(getenv("BNMF_HYPER_THREADS")) : 256;
                  const long long ncell = (long long)cfg.N * cfg.G;
                  CK(cudaMemsetAsync(alpha_n_retry, 0, sizeof(int), side));
                  if (ht == 128) k_eside_hyper<T, 128><<<blocks(ncell, 128), 128, 0, side>>>(d, h_iter + 1, alpha_retry, alpha_n_retry, alpha_cap);
                  else if (ht == 512) k_eside_hyper<T, 512><<<blocks(ncell, 512), 512, 0, side>>>(d, h_iter + 1, alpha_retry, alpha_n_retry, alpha_cap);
                  else k_eside_hyper<T, 256><<<blocks(ncell, 256), 256, 0, side>>>(d, h_iter + 1, alpha_retry, alpha_n_retry, alpha_cap);
                  k_alpha_retry<T><<<blocks((alpha_cap + ALPHA_RETRY_PER_WARP - 1) / ALPHA_RETRY_PER_WARP, 8), 256, 0, side>>>(d, h_iter + 1, alpha_retry, alpha_n_retry, alpha_cap);
                  ++launches;
                }
                CK(cudaEventRecord(ev_join, side));
                hyper_ready = true; launches += 2;
              }
              break;
      default: k_pside<T, 128, PRIOR_GAMMA, 1><<<cfg.N, 128, 0, stream>>>(d, keepP); mark("k_pside");
               k_eside<T, ET, PRIOR_GAMMA, 1><<<d.n_eblocks, ET, 0, stream>>>(d, keepE); mark("k_eside"); break;
    }
    launches += 2;
    if (from_prior) { k_init_rank<T><<<1, 32, 0, stream>>>(d, (have & BNMF_HAVE_A) ? 1 : 0); mark("k_init_rank"); ++launches; }
    else if (cfg.learning_rank) { if (rank_sweep()) return 1; }
    if (!(from_prior && (have & BNMF_HAVE_Z))) {
      if (z0) CK(cudaEventRecord(z0, stream));
      if (z_dispatch(false)) return 1; ++launches;
      if (z1) CK(cudaEventRecord(z1, stream));
    } else {
      if (refresh_metrics_only()) return 1;
    }
    return 0;
  }
  int rank_sweep();
  int rank_sweep_kernels(int* pending);
  int refresh_metrics_only();
  bool sweep_model = false; int h_converged = 0;
  int p_kx = 32, p_gy = 8, p_ktiles = 1, col_blocks = 1, e_wpb = 8, e_stage = 0, e_lpg = 32, e_slots = 8; size_t e_smem = 0;
  // row-resident P sweep (k_p_rows): cluster size, genomes per block, threads, shared memory; 0 = not used
  int pr_cs = 0, pr_gslice = 0, pr_threads = 0; size_t pr_smem = 0; long long pr_Gp = 0; T* Et = nullptr;
  int p_rows_launch();
  double* gram_part = nullptr; double* gram_buf = nullptr; int gram_chunks = 0;   // Normal likelihood: Gram-matrix P sweep
  int p_gram_launch();
  int mhat_rebuild(); size_t eg_smem = 0; int eg_gb = 0; bool a_coop = false;
  unsigned char* tc_Pd = nullptr; int* tc_eP = nullptr; size_t tc_smem = 0;     // tensor-core Mhat (Normal likelihood): digit planes of P
  double* asum = nullptr; double sig_alpha = 3.0, sig_beta = 3.0;

  // End of an iteration: one kernel folds the partials, sums over the shards and composes the metrics row.
  // Sharded runs sum everything the next iteration (SP, rowSums(E): read by k_pside) and this iteration's
  // metrics row need in ONE exchange of ~16 KB inside that kernel (one-shot all-reduce over NVLink peer memory,
  // bnmf_poisson.cuh); where the peers' memory cannot be mapped, one grouped NCCL operation + k_metrics.
  int finish_iteration() {
    Xchg x; memset(&x, 0, sizeof(x));
    x.world = 1;
    if (world > 1) {
      if (xchg_usable()) {
        CommBox& b = *commbox;
        x.world = world; x.rank = rank; x.seq = ++b.seq; x.slot_words = b.slot_words; x.err = b.xerr;
        for (int r = 0; r < world; ++r) x.buf[r] = b.xbuf[r];
      } else x.world = -1;                       // NCCL sums, then k_metrics
    }
    k_reduce_partials<T, 256><<<RED_BLOCKS, 256, 0, stream>>>(d, red_slices, red_ticket, x); mark("k_reduce_partials"); ++launches;
    if (x.world == -1) {
      const bool stats = cfg.likelihood == BNMF_POISSON && !cfg.MH;
      if (stats) g_nccl.GroupStart();
      int rc = stats ? allreduce_stats() : 0;
      if (!rc) rc = allreduce_red();
      if (stats) { const int r = g_nccl.GroupEnd(); if (!rc && r) rc = fail("ncclGroupEnd: %s", g_nccl.GetErrorString(r)); }
      if (rc) return 1;
      k_metrics<T><<<1, 32, 0, stream>>>(d); mark("k_metrics"); ++launches;
    }
    return 0;
  }
  // a peer that never published (process gone): the exchange kernel gives up after ~4 s and says so
  int check_xchg() {
    if (!(world > 1 && commbox && commbox->xchg_ok)) return 0;
    int e = 0;
    CK(cudaMemcpyAsync(&e, commbox->xerr, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (e) return fail("cross-GPU exchange: rank %d never published (peer process gone?)", e - 1);
    return 0;
  }

  int init_from_prior(uint32_t have, uint32_t have_prior, double* row) override {
    CK(cudaSetDevice(cfg.device));
    const int K = cfg.K, N = cfg.N; const long long KN = (long long)K * N, NG = (long long)N * cfg.G;
    // which prior-parameter columns to draw: all of them unless supplied without NaN
    std::vector<int> fl(5 * N, 1);
    CK(cudaMemcpyAsync(nanflags, fl.data(), sizeof(int) * 5 * N, cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    (void)have_prior;
    for (int side = 0; side < 2; ++side) {
      const char* names_p[5] = {"Mu_p", "Sigmasq_p", "Lambda_p", "Alpha_p", "Beta_p"};
      const char* names_e[5] = {"Mu_e", "Sigmasq_e", "Lambda_e", "Alpha_e", "Beta_e"};
      // (supplied matrices: bit i of have_prior for side p, bit 8+i for side e)
      for (int w = 0; w < 5; ++w) {
        const bool supplied = have_prior & (1u << (w + 8 * side));
        if (!supplied) continue;
        auto it = st.find(side == 0 ? names_p[w] : names_e[w]);
        if (it == st.end()) continue;
        CK(cudaMemsetAsync(nanflags + w * N, 0, sizeof(int) * N, stream));
        const long long len = side == 0 ? KN : NG;
        k_nan_cols<T><<<blocks(len, 256), 256, 0, stream>>>((const T*)it->second.p, len, K, N, side, nanflags + w * N);
        if (allreduce_buf(nanflags + w * N, N, NCCL_INT32, NCCL_MAX)) return 1;
      }
      const long long cells = side == 0 ? KN : NG;
      k_init_prior<T><<<blocks(cells, 128), 128, 0, stream>>>(d, side, nanflags);
      // reset flags for the next side
      CK(cudaMemcpyAsync(nanflags, fl.data(), sizeof(int) * 5 * N, cudaMemcpyHostToDevice, stream));
      CK(cudaStreamSynchronize(stream));
    }
    if (cfg.likelihood == BNMF_NORMAL) { if (init_sigmasq_prior()) return 1; }
    Ctrl c; c.iter = 1; c.converged = 0; c.row = 0; c.ring_pos = 0; c.ring_count = 0;
    CK(cudaMemcpyAsync(d.ctrl, &c, sizeof(c), cudaMemcpyHostToDevice, stream));
    CK(cudaMemsetAsync(red_i64, 0, sizeof(long long) * (KN + N), stream));
    CK(cudaMemsetAsync(d.SE, 0, sizeof(int32_t) * NG, stream));
    CK(cudaMemsetAsync(work_ctr, 0, sizeof(int) * (n_ktiles + 8), stream));
    launches = 0;
    if (cfg.likelihood == BNMF_POISSON && !cfg.MH) { if (poisson_iteration(1, have, nullptr, nullptr)) return 1; }
    else { if (mh_iteration(1, have)) return 1; }
    if (finish_iteration()) return 1;
    CK(cudaMemcpyAsync(h_metrics, d.metrics, sizeof(double) * MC_COLS, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    if (row) memcpy(row, h_metrics, sizeof(double) * MC_COLS);
    h_rows.assign(h_metrics, h_metrics + MC_COLS);
    rs = RunState();
    return 0;
  }
  int init_sigmasq_prior();

  // ---- one iteration as a CUDA graph -------------------------------------------------
  // Small and sweep-based problems are launch-bound (60-170 tiny kernels per iteration at
  // 96 x 500): the fixed launch sequence of an iteration is captured once per
  // (converged, want P, want A) and replayed.  Everything that changes from one iteration to
  // the next (iteration number, metrics row, ring slot) lives in device memory (Ctrl).
  cudaGraphExec_t gexec = nullptr; int gkey = -1; int glaunches = 0;
  void drop_graph() { if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; } gkey = -1; }
  bool graphs_allowed() const {
    static const bool off = getenv("BNMF_GRAPH") && !strcmp(getenv("BNMF_GRAPH"), "0");
    if (off || world > 1) return false;
    return sweep_model || (long long)cfg.K * cfg.G <= 2000000LL;
  }
  int launch_iteration(bool wantP, bool wantA, cudaEvent_t z0, cudaEvent_t z1) {
    const long long KN = (long long)cfg.K * cfg.N;
    // (iterations replayed from a graph keep k_begin_iter: their side-branch kernels read the device's counter)
    const char* nf = getenv("BNMF_FOLD_BEGIN");      // "0": k_begin_iter as a launch of its own (tests compare the two)
    const bool fold_begin = !(nf && !strcmp(nf, "0")) && cfg.likelihood == BNMF_POISSON && !cfg.MH && !graphs_allowed();
    if (!fold_begin) { k_begin_iter<T><<<1, 64, 0, stream>>>(d, work_ctr, n_ktiles); mark("k_begin_iter"); ++launches; }
    if (cfg.likelihood == BNMF_POISSON && !cfg.MH) { if (poisson_iteration(0, 0, z0, z1, fold_begin)) return 1; }
    else { if (mh_iteration(0, 0)) return 1; }
    if (finish_iteration()) return 1;
    if (wantP || wantA) {
      k_hist_copy<T><<<blocks(KN, 256), 256, 0, stream>>>(d, wantP ? P_hist : nullptr, wantA ? A_hist : nullptr); mark("k_hist_copy"); ++launches;
    }
    return 0;
  }
  int ensure_graph(bool wantP, bool wantA) {
    const int key = (h_converged ? 1 : 0) | (wantP ? 2 : 0) | (wantA ? 4 : 0);
    if (gexec && gkey == key) return 0;
    if (!alpha_retry && cfg.prior == BNMF_GAMMA && cfg.likelihood == BNMF_POISSON && !cfg.MH) {     // (fork_hyper)
      const long long ncell = (long long)cfg.N * cfg.G;
      alpha_cap = (int)std::max<long long>(256, std::min<long long>(ncell / 4 + 1024, 1LL << 30));
      if (const char* e = getenv("BNMF_ALPHA_CAP")) { const int v = atoi(e); if (v >= 1) alpha_cap = v; }
      if (dalloc(&alpha_retry, (long long)BNMF_ALPHA_ENV_COLS * alpha_cap) || dalloc(&alpha_n_retry, 1)) return 1;
      CK(cudaStreamSynchronize(stream));
    }
    if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
    cudaGraph_t g = nullptr;
    const int before = launches;
    CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    const int rc = launch_iteration(wantP, wantA, nullptr, nullptr);
    const cudaError_t e = cudaStreamEndCapture(stream, &g);
    glaunches = launches - before; launches = before;
    if (rc) { if (g) cudaGraphDestroy(g); return 1; }
    if (e != cudaSuccess) return fail("graph capture of the iteration failed: %s", cudaGetErrorString(e));
    const cudaError_t e2 = cudaGraphInstantiate(&gexec, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) { gexec = nullptr; return fail("cudaGraphInstantiate: %s", cudaGetErrorString(e2)); }
    gkey = key;
    return 0;
  }

  int step(int n_iters, int converged, double* metrics, double* P_out, double* A_out) override {
    CK(cudaSetDevice(cfg.device));
    if (n_iters < 0) return fail("bnmf_step: n_iters < 0");
    h_converged = converged ? 1 : 0;
    const int K = cfg.K, N = cfg.N; const long long KN = (long long)K * N;
    launches = 0;
    last_z_ms = 0; last_iter_ms = 0;
    const bool use_graph = graphs_allowed();
    if (use_graph) { if (ensure_graph(P_out != nullptr, A_out != nullptr)) return 1; }
    if (overlap_allowed()) {     // the host's copy of state$iter (the side stream's draws are keyed by it)
      Ctrl hc; CK(cudaMemcpyAsync(&hc, d.ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      h_iter = hc.iter;
    }
    hyper_ready = false;
    CK(cudaEventRecord(ev0, stream));
    int done = 0;
    while (done < n_iters) {
      const int chunk = std::min(n_iters - done, d.metrics_cap);
      // ctrl.converged / ctrl.row for this chunk (iter and ring position live on the device)
      int two[2] = {converged, -1};
      CK(cudaMemcpyAsync(&d.ctrl->converged, two, sizeof(two), cudaMemcpyHostToDevice, stream));
      // bnmf_timing's iter_ms: two event records per iteration, at its boundaries; zstat_ms: two more INSIDE the
      // chain of an iteration, on either side of k_zstat (together 2 us of a 257 us shard iteration, 10 of 1,580 at
      // C3) -- recorded with the L2 flush (benchmark mode) or when asked for.  BNMF_TIMING = "0": no events,
      // "iter": the iteration's only, "z": both
      const char* tev = getenv("BNMF_TIMING");
      const bool timei = !(tev && !strcmp(tev, "0"));
      const bool wantz = tev ? !strcmp(tev, "z") : flush_bytes > 0;
      const bool timez = timei && wantz && !use_graph && time_z && cfg.likelihood == BNMF_POISSON && !cfg.MH;
      if (timez) while ((int)zev.size() < 2 * chunk) { cudaEvent_t e; CK(cudaEventCreate(&e)); zev.push_back(e); }
      if (timei) while ((int)iev.size() < 2 * chunk) { cudaEvent_t e; CK(cudaEventCreate(&e)); iev.push_back(e); }
      for (int i = 0; i < chunk; ++i) {
        if (flush_bytes) CK(cudaMemsetAsync(flush_buf, i & 0xff, flush_bytes, stream));
        if (timei) CK(cudaEventRecord(iev[2 * i], stream));
        spec_next = done + i + 1 < n_iters;     // never past the end of this call: the state handed back is iteration n's
        ++h_iter;
        if (use_graph) { CK(cudaGraphLaunch(gexec, stream)); launches += glaunches; }
        else if (launch_iteration(P_out != nullptr, A_out != nullptr, timez ? zev[2 * i] : nullptr, timez ? zev[2 * i + 1] : nullptr)) return 1;
        if (timei) CK(cudaEventRecord(iev[2 * i + 1], stream));
      }
      CK(cudaMemcpyAsync(h_metrics, d.metrics, sizeof(double) * MC_COLS * chunk, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      CK(cudaGetLastError());
      if (check_xchg()) return 1;
      if (metrics) memcpy(metrics + (long long)done * MC_COLS, h_metrics, sizeof(double) * MC_COLS * chunk);
      if (h_rows.size() > (size_t)MC_COLS * 2000000) h_rows.erase(h_rows.begin(), h_rows.begin() + (long long)MC_COLS * 1000000);
      h_rows.insert(h_rows.end(), h_metrics, h_metrics + (size_t)MC_COLS * chunk);
      if (P_out) { if (dev_to_host(P_hist, ST_T, (long long)chunk * KN, P_out + (long long)done * KN)) return 1; }
      if (A_out) { if (dev_to_host(A_hist, ST_I32, (long long)chunk * N, A_out + (long long)done * N)) return 1; }
      if (timez) for (int i = 0; i < chunk; ++i) { float ms = 0; CK(cudaEventElapsedTime(&ms, zev[2 * i], zev[2 * i + 1])); last_z_ms += ms; }
      if (timei) for (int i = 0; i < chunk; ++i) { float ms = 0; CK(cudaEventElapsedTime(&ms, iev[2 * i], iev[2 * i + 1])); last_iter_ms += ms; }
      done += chunk;
    }
    CK(cudaEventRecord(ev1, stream));
    CK(cudaEventSynchronize(ev1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, ev0, ev1));
    last_total_ms = ms; last_launches = launches;
    return 0;
  }

  int ring_count(int* c) override {
    CK(cudaSetDevice(cfg.device));
    Ctrl h; CK(cudaMemcpyAsync(&h, d.ctrl, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    *c = h.ring_count;
    return 0;
  }
  int get_sample(const char* name, int ago, double* out, int64_t len) override {
    CK(cudaSetDevice(cfg.device));
    if (d.ring_cap <= 0) return fail("bnmf_get_sample: the handle was created with ring_cap = 0");
    Ctrl h; CK(cudaMemcpyAsync(&h, d.ctrl, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (ago < 0 || ago >= h.ring_count) return fail("bnmf_get_sample: ago = %d outside the %d samples held", ago, h.ring_count);
    const int slot = ((h.ring_pos - 1 - ago) % d.ring_cap + d.ring_cap) % d.ring_cap;
    const long long KN = (long long)cfg.K * cfg.N, NG = (long long)cfg.N * cfg.G;
    if (!strcmp(name, "P")) { if (len != KN) return fail("bnmf_get_sample: P has %lld elements", KN); return dev_to_host(d.ring_P + slot * KN, ST_T, KN, out); }
    if (!strcmp(name, "E")) { if (len != NG) return fail("bnmf_get_sample: E has %lld elements", NG); return dev_to_host(d.ring_E + slot * NG, ST_T, NG, out); }
    if (!strcmp(name, "A")) { if (len != cfg.N) return fail("bnmf_get_sample: A has %d elements", cfg.N); return dev_to_host(d.ring_A + (long long)slot * cfg.N, ST_I32, cfg.N, out); }
    return fail("bnmf_get_sample: the ring holds P, E and A (got '%s')", name);
  }
  int run(const bnmf_convergence_control* cc, int post_warmup, double* metrics_out, int64_t rows_cap,
          double* map_out, int64_t checks_cap, bnmf_run_result* res) override;
  int get_map(int n_samples, double* P, double* E, double* A, int* n_match) override;
  int get_ci(int n_samples, double plo, double phi, double* P_lo, double* P_hi, double* E_lo, double* E_hi, int* n_match) override;
  int map_slots(int n_samples, std::vector<int>& match, std::string& mode);
  int assign(int n_samples, const double* ref, int n_ref, double ci, int* n_keep, int* keep, double* votes, int* asg,
             double* mapc, double* lo, double* hi, int* n_match) override;

  int set_l2_flush(size_t bytes) override {
    CK(cudaSetDevice(cfg.device));
    CK(cudaStreamSynchronize(stream));
    if (flush_buf) { CK(cudaFree(flush_buf)); flush_buf = nullptr; }
    flush_bytes = 0;
    if (bytes) { CK(cudaMalloc(&flush_buf, bytes)); flush_bytes = bytes; }
    return 0;
  }
  int timing(double* total, double* iter, double* z, int64_t* l) override {
    if (total) *total = last_total_ms;
    if (iter) *iter = last_iter_ms;
    if (z) *z = last_z_ms;
    if (l) *l = last_launches;
    return 0;
  }
  int sample_z(int iter, double* ms) override {
    CK(cudaSetDevice(cfg.device));
    if (!(cfg.likelihood == BNMF_POISSON && !cfg.MH)) return fail("bnmf_sample_z: only the Poisson non-MH model has latent counts");
    const long long KN = (long long)cfg.K * cfg.N, NG = (long long)cfg.N * cfg.G;
    CK(cudaMemcpyAsync(&d.ctrl->iter, &iter, sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemsetAsync(d.SP, 0, sizeof(long long) * KN, stream));
    CK(cudaMemsetAsync(d.SE, 0, sizeof(int32_t) * NG, stream));
    CK(cudaMemsetAsync(work_ctr, 0, sizeof(int) * (n_ktiles + 8), stream));
    CK(cudaEventRecord(ev0, stream));
    if (z_dispatch(false)) return 1;
    CK(cudaEventRecord(ev1, stream));
    CK(cudaEventSynchronize(ev1));
    CK(cudaGetLastError());
    float t = 0; CK(cudaEventElapsedTime(&t, ev0, ev1));
    if (ms) *ms = t;
    if (allreduce_buf(d.SP, (size_t)KN, NCCL_UINT64, NCCL_SUM)) return 1;
    CK(cudaStreamSynchronize(stream));
    return 0;
  }
};

// ---- small reductions after a user-supplied P / E -----------------------------------
template <typename T> static __global__ void k_rowsumE(Dev<T> d) {
  // one block per n; fixed-point sum is order independent
  const int n = blockIdx.x;
  long long s = 0;
  for (long long g = threadIdx.x; g < d.G; g += blockDim.x) s += llrint((double)d.E[n + (long long)d.N * g] * RS_FX);
  atomicAdd((unsigned long long*)&d.rowsumE_fx[n], (unsigned long long)s);
}
template <typename T> static __global__ void k_colsumP(Dev<T> d) {
  __shared__ double sc[4];
  const int n = blockIdx.x;
  double s = 0;
  for (int k = threadIdx.x; k < d.K; k += 128) s += (double)d.P[k + (long long)d.K * n];
  double r = block_sum<128>(s, sc);
  if (threadIdx.x == 0) d.colsumP[n] = (T)r;
}
template <typename T> int Sampler<T>::refresh_rowsumE() {
  CK(cudaMemsetAsync(d.rowsumE_fx, 0, sizeof(long long) * cfg.N, stream));
  k_rowsumE<T><<<cfg.N, 256, 0, stream>>>(d);
  if (allreduce_buf(d.rowsumE_fx, cfg.N, NCCL_INT64, NCCL_SUM)) return 1;
  CK(cudaStreamSynchronize(stream));
  return 0;
}
template <typename T> int Sampler<T>::refresh_colsumP() {
  k_colsumP<T><<<cfg.N, 128, 0, stream>>>(d);
  CK(cudaStreamSynchronize(stream));
  return 0;
}

#include "bnmf_api_mh.inl"
#include "bnmf_api_map.inl"
#include "bnmf_api_run.inl"

// ---------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------
extern "C" {

const char* bnmf_last_error(void) { return g_err; }

int bnmf_check_model(int likelihood, int prior, int MH, char* msg, size_t msg_len) {
  const char* m = nullptr;
  if (likelihood != BNMF_NORMAL && likelihood != BNMF_POISSON) m = "likelihood must be one of normal, poisson";
  else if (likelihood == BNMF_NORMAL) {
    if (!(prior == BNMF_TRUNCNORMAL || prior == BNMF_EXPONENTIAL)) m = "prior must be one of c('truncnormal','exponential') with `likelihood = 'normal'`";
  } else {
    if (!(prior == BNMF_GAMMA || prior == BNMF_EXPONENTIAL || prior == BNMF_TRUNCNORMAL)) m = "prior must be one of c('gamma','exponential','truncnormal') with `likelihood = 'poisson'`";
    else if (prior == BNMF_GAMMA && MH) m = "gamma prior cannot be used in a MH-within-gibbs sampler";
    else if (prior == BNMF_TRUNCNORMAL && !MH) m = "truncnormal prior can only be used in a MH-within-gibbs sampler";
  }
  if (!m) { if (msg && msg_len) msg[0] = 0; return 0; }
  if (msg && msg_len) snprintf(msg, msg_len, "%s", m);
  fail("%s", m);
  return 1;
}

int bnmf_create(const bnmf_config* cfg, const double* data, bnmf_handle** out) {
  if (!cfg || !data || !out) return fail("bnmf_create: null argument");
  *out = nullptr;
  char msg[256];
  if (bnmf_check_model(cfg->likelihood, cfg->prior, cfg->MH, msg, sizeof(msg))) return 1;
  if (cfg->likelihood == BNMF_NORMAL && cfg->MH) return fail("MH applies to the Poisson likelihood only");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1)
    return fail("bnmf_create: no CUDA device available (%s); this library has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail("bnmf_create: device %d out of range (%d visible)", cfg->device, ndev);
  if (cfg->precision == BNMF_F64) {
    auto* s = new Sampler<double>();
    if (s->create(cfg, data)) { delete s; return 1; }
    *out = s;
  } else if (cfg->precision == BNMF_F32) {
    auto* s = new Sampler<float>();
    if (s->create(cfg, data)) { delete s; return 1; }
    *out = s;
  } else return fail("bnmf_create: precision must be BNMF_F64 or BNMF_F32");
  return 0;
}
void bnmf_destroy(bnmf_handle* h) { delete h; }

#define NEED(h) if (!(h)) return fail("null handle")
int bnmf_set_hyper(bnmf_handle* h, const char* name, const double* v, int64_t rows, int64_t cols) { NEED(h); return h->set_hyper(name, v, rows, cols); }
int bnmf_set_state(bnmf_handle* h, const char* name, const double* v, int64_t len) { NEED(h); return h->set_state(name, v, len); }
int bnmf_get_state(bnmf_handle* h, const char* name, double* out, int64_t len) { NEED(h); return h->get_state(name, out, len); }
int bnmf_set_temperature_schedule(bnmf_handle* h, const double* t, int64_t n) { NEED(h); return h->set_temps(t, n); }
int bnmf_init_from_prior(bnmf_handle* h, uint32_t have, uint32_t have_prior, double* row) { NEED(h); return h->init_from_prior(have, have_prior, row); }
int bnmf_step(bnmf_handle* h, int32_t n, int32_t conv, double* m, double* P, double* A) { NEED(h); return h->step(n, conv, m, P, A); }
int bnmf_run(bnmf_handle* h, const bnmf_convergence_control* cc, int32_t pw, double* m, int64_t rc, double* mm, int64_t cc_cap, bnmf_run_result* res) {
  NEED(h); return h->run(cc, pw, m, rc, mm, cc_cap, res);
}
int bnmf_ring_count(bnmf_handle* h, int32_t* c) { NEED(h); return h->ring_count(c); }
int bnmf_get_sample(bnmf_handle* h, const char* name, int32_t ago, double* out, int64_t len) { NEED(h); return h->get_sample(name, ago, out, len); }
int bnmf_get_map(bnmf_handle* h, int32_t n, double* P, double* E, double* A, int32_t* nm) { NEED(h); return h->get_map(n, P, E, A, nm); }
int bnmf_assign_signatures(bnmf_handle* h, int32_t n, const double* ref, int32_t n_ref, double ci, int32_t* n_keep, int32_t* keep, double* votes,
                           int32_t* asg, double* mapc, double* lo, double* hi, int32_t* nm) {
  NEED(h); return h->assign(n, ref, n_ref, ci, n_keep, keep, votes, asg, mapc, lo, hi, nm);
}
int bnmf_get_credible_intervals(bnmf_handle* h, int32_t n, double plo, double phi, double* P_lo, double* P_hi, double* E_lo, double* E_hi, int32_t* nm) {
  NEED(h); return h->get_ci(n, plo, phi, P_lo, P_hi, E_lo, E_hi, nm);
}
int bnmf_comm_unique_id(char* id128) {
  if (load_nccl()) return 1;
  int r = g_nccl.GetUniqueId(id128);
  if (r) return fail("ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  return 0;
}
int bnmf_comm_init(bnmf_handle* h, const char* id, int32_t rank, int32_t world) { NEED(h); return h->comm_init(id, rank, world); }
int bnmf_comm_share(bnmf_handle* h, bnmf_handle* src) { NEED(h); NEED(src); return h->comm_share(src); }
int bnmf_timing(bnmf_handle* h, double* t, double* it, double* z, int64_t* l) { NEED(h); return h->timing(t, it, z, l); }
int bnmf_set_l2_flush(bnmf_handle* h, size_t bytes) { NEED(h); return h->set_l2_flush(bytes); }
int bnmf_sample_z(bnmf_handle* h, int32_t iter, double* ms) { NEED(h); return h->sample_z(iter, ms); }
int bnmf_profile_iteration(bnmf_handle* h, int32_t conv, char* names, double* ms, int32_t* counts, int32_t cap, int32_t* n) { NEED(h); return h->profile(conv, names, ms, counts, cap, n); }
int bnmf_release_cached_memory(void) {
  {   // the process-wide pinned staging buffer too (no handle is inside create / get_state: the lock is free)
    std::lock_guard<std::mutex> lk(g_pin_mutex);
    if (g_pin) { cudaFreeHost(g_pin); g_pin = nullptr; g_pin_bytes = 0; }
  }
  pinned_release_all();
  return release_cached_blocks(-1);
}

}  // extern "C"
