// lmpc_capi.cu -- implementation of the C ABI declared in include/lmpc_b200.h.
//
// Host side of the drop-in boundary: owns the device workspace, the device-resident safe-set slab
// (host-side ingestion mirrors SafeSetManager::add_lap / SafeSetRecorder::load, reference
// safe_set.cpp:116-151,260-276) and launches the three kernels of lmpc_kernels.cuh on the handle's
// stream.  No CPU compute path exists: every entry point needs the CUDA device.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lmpc_b200.h"
#include "lmpc_host_params.h"
#include "lmpc_kernels.cuh"
#include "lmpc_qp_launch.h"
#include "lmpc_loop.cuh"

namespace {

struct HostLap {
  int n;
  double L;   // total_length the lap was added with
  std::vector<double> ps, pe, xr, J;
  std::vector<int> canon;
  std::vector<double> x, u, k, t;   // the lap as given (regression points; u/k/t empty if the caller passed none)
};

// SafeSetRecorder state (safe_set.hpp:127-150)
struct Recorder {
  bool last_x_valid = false, initialized = false, to_file = false;
  std::string prefix;
  int lap_count = 0;
  double last_px = 0.0;
  std::vector<double> x, u, k, t;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct lmpc_handle {
  lmpc_mpc_config cfg;
  lmpc_vehicle_params veh;
  LmpcQpParams P;
  LmpcModel M;
  int device = 0;
  int max_batch = 0;
  cudaStream_t stream = nullptr;
  // the safe-set query of a tick does not depend on the linearisation: it runs on a side stream, forked from and joined to
  // the caller's stream with two events
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  int64_t launches = 0;
  // safe set: circular buffer, oldest first (boost::circular_buffer semantics, safe_set.cpp:139-151)
  std::vector<HostLap> laps;
  DevBuf slab;
  std::vector<LmpcLapView> dev_laps;   // newest first, device pointers into the slab
  // device workspace
  DevBuf ws_abg, ws_cen, ws_ssx, ws_ssj, ws_sqp, ws_qpscr;
  // error-dynamics regression: points of all laps (built lazily after a safe-set change), optional in-tick plan
  DevBuf reg_slab, ws_reg;
  LmpcRegView reg_view{nullptr, nullptr, 0, 0, -1};
  bool reg_dirty = true;
  bool reg_in_tick = false;
  LmpcRegPlan reg_plan{};
  Recorder rec;
  // track interpolants (device copy of LmpcTrackHost) and closed-loop workspace
  LmpcTrackHost track_host;
  DevBuf track_dev, ws_loop, st_loop;
  LmpcTrack track{};   // device view; m == 0 until lmpc_track_set
  // device staging for host-memory callers
  DevBuf st_in, st_out;
  size_t qp_smem = 0;
  // multi-GPU result exchange over peer memory (lmpc_gather_*): one allocation
  //   [sets][world][slab] | flags: u64 [world] | done counter | error word
  struct Gather {
    int world = 0, rank = 0, B = 0, sets = 0;
    size_t slab_bytes = 0, set_bytes = 0, flags_off = 0, total_bytes = 0;
    void* base = nullptr;
    void* peer[LMPC_MAX_PEERS] = {nullptr};   // IPC mappings of the peers' allocations (null for self)
    bool connected = false;
    unsigned long long seq = 0;
    int active_set = -1;                      // >= 0 while lmpc_solve_gather_batch routes the trajectory outputs
  } gat;
  // per-agent safe sets + device-side lap recording (lmpc_agents_*; csrc/lmpc_agents.cuh)
  LmpcAgentSets ag{};
  DevBuf ag_buf, ag_cnt;       // one allocation for the agents' arrays; per-instance column counts of the tick
  bool ag_on = false;
  // optional per-kernel timing: events recorded on the stream around the three kernels of each solve
  bool timing = false;
  std::vector<cudaEvent_t> tev;   // 4 events per recorded solve (ring)
  int tcount = 0;
};

static const int kTimingRing = 1024;

// compile-time-layout kernels exist for N in {20, 40} with 16 row slots per stage (LMPC_FIXED_LAYOUT=0 disables)
static int qp_fixed_n(const lmpc_handle* h) {
  const char* e = getenv("LMPC_FIXED_LAYOUT");
  if (e && atoi(e) == 0) return 0;
  if (h->P.NW == 1 && h->P.RS == 16 && (h->P.N == 20 || h->P.N == 40)) return h->P.N;
  // the long shipped horizons (lmpc_qp_tu4.cu): N = 60 with K = 0 or 65..96, N = 80 with K = 0
  const int kpl_ = std::max(1, (h->P.K + 31) / 32);
  if (h->P.NW == 1 && h->P.RS == 16 && ((h->P.N == 60 && (kpl_ == 1 || kpl_ == 3)) || (h->P.N == 80 && kpl_ == 1))) return h->P.N;
  if (h->P.NW == 2 && h->P.RS == 16 && h->P.N == 20) return h->P.N;
  return 0;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
      return LMPC_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)

static int dev_reserve(lmpc_handle* h, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes) return LMPC_OK;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
  if (cudaMalloc(&b.p, bytes) != cudaSuccess) { h->err = "cudaMalloc failed"; return LMPC_ERR_ALLOC; }
  b.bytes = bytes;
  return LMPC_OK;
}

extern "C" int lmpc_version(void) { return 100; }

extern "C" const char* lmpc_status_string(int s) {
  switch (s) {
    case LMPC_OK: return "ok";
    case LMPC_ERR_INVALID: return "invalid argument or unsupported configuration";
    case LMPC_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
    case LMPC_ERR_CUDA: return "CUDA runtime error";
    case LMPC_ERR_ALLOC: return "allocation failed";
    case LMPC_ERR_IO: return "safe-set file missing or malformed";
    case LMPC_ERR_CAPACITY: return "batch exceeds max_batch";
    default: return "unknown";
  }
}

extern "C" const char* lmpc_last_error(const lmpc_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int64_t lmpc_launch_count(const lmpc_handle* h) { return h ? h->launches : 0; }

static int qp_kpl(const lmpc_handle* h) { return std::max(1, (h->P.K + 32 * h->P.NW - 1) / (32 * h->P.NW)); }

static int set_qp_attr(lmpc_handle* h) {
  cudaError_t e = cudaSuccess;
  if (!lmpc_qp_set_smem(h->P.NW, qp_kpl(h), qp_fixed_n(h), h->qp_smem, &e)) { h->err = "no QP kernel instantiation for this configuration"; return LMPC_ERR_INVALID; }
  if (e != cudaSuccess) { h->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return LMPC_ERR_CUDA; }
  return LMPC_OK;
}

extern "C" int lmpc_create(const lmpc_mpc_config* config, const lmpc_vehicle_params* vehicle, int device_ordinal,
                           int max_batch, lmpc_handle** out) {
  if (!config || !vehicle || !out || max_batch < 1) return LMPC_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return LMPC_ERR_NO_DEVICE;
  if (device_ordinal < 0 || device_ordinal >= ndev) return LMPC_ERR_INVALID;
  if (vehicle->integrator != 0 && vehicle->integrator != 1) return LMPC_ERR_INVALID;
  lmpc_handle* h = new lmpc_handle();
  h->cfg = *config; h->veh = *vehicle; h->device = device_ordinal; h->max_batch = max_batch;
  // warps per MPC instance: 1, 2 or 4 (LMPC_WARPS_PER_INSTANCE overrides; 4 needs K <= 128, 2 needs K <= 256)
  int nw = 1;   // measured fastest on B200 at B = 1024 and 8192 (profiles/); 2 and 4 are kept selectable
  if (const char* e = getenv("LMPC_WARPS_PER_INSTANCE")) nw = atoi(e);
  if (nw != 1 && nw != 2 && nw != 4) nw = 1;
  if (config->learning && nw == 4 && config->num_ss_pts > 128) nw = 2;
  int rc = lmpc_make_qp_params(*config, *vehicle, &h->P, nw);
  if (rc != LMPC_OK) { delete h; return rc; }
  h->M = lmpc_make_model(*vehicle);
  if (cudaSetDevice(device_ordinal) != cudaSuccess) { delete h; return LMPC_ERR_NO_DEVICE; }
  h->qp_smem = sizeof(double) * (size_t)h->P.lay.total;
  if (const char* e = getenv("LMPC_QP_EXTRA_SMEM")) h->qp_smem += (size_t)atoi(e);   // occupancy experiments (profiles/README.md)
  rc = set_qp_attr(h);
  if (rc != LMPC_OK) { fprintf(stderr, "lmpc_create: %s\n", h->err.c_str()); delete h; return rc; }
  const size_t B = (size_t)max_batch, N = (size_t)h->P.N, NS = (size_t)h->P.NS, K = (size_t)std::max(h->P.K, 1);
  rc = dev_reserve(h, h->ws_abg, sizeof(double) * 54 * NS * B);
  if (rc == LMPC_OK) rc = dev_reserve(h, h->ws_cen, sizeof(double) * 6 * B);
  if (rc == LMPC_OK) rc = dev_reserve(h, h->ws_ssx, sizeof(double) * 6 * K * B);
  if (rc == LMPC_OK) rc = dev_reserve(h, h->ws_ssj, sizeof(double) * K * B);
  if (rc == LMPC_OK) rc = dev_reserve(h, h->ws_qpscr, sizeof(double) * (size_t)LMPC_QP_SCRATCH(N, K) * B);
  (void)N;
  if (rc == LMPC_OK && (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
                        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess))
    rc = LMPC_ERR_CUDA;
  if (rc != LMPC_OK) { lmpc_destroy(h); return rc; }
  *out = h;
  return LMPC_OK;
}

extern "C" int lmpc_destroy(lmpc_handle* h) {
  if (!h) return LMPC_ERR_INVALID;
  cudaSetDevice(h->device);
  for (auto& e : h->tev) cudaEventDestroy(e);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  for (int p = 0; p < LMPC_MAX_PEERS; p++) if (h->gat.peer[p]) cudaIpcCloseMemHandle(h->gat.peer[p]);
  if (h->gat.base) cudaFree(h->gat.base);
  for (DevBuf* b : {&h->slab, &h->ws_abg, &h->ws_cen, &h->ws_ssx, &h->ws_ssj, &h->ws_sqp, &h->ws_qpscr, &h->reg_slab, &h->ws_reg, &h->track_dev, &h->ws_loop, &h->st_loop, &h->st_in, &h->st_out, &h->ag_buf, &h->ag_cnt})
    if (b->p) cudaFree(b->p);
  delete h;
  return LMPC_OK;
}

extern "C" int lmpc_set_stream(lmpc_handle* h, void* s) {
  if (!h) return LMPC_ERR_INVALID;
  h->stream = (cudaStream_t)s;
  return LMPC_OK;
}

extern "C" int lmpc_set_timing(lmpc_handle* h, int enable) {
  if (!h) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  if (enable && h->tev.empty()) {
    h->tev.resize(4 * (size_t)kTimingRing);
    for (auto& e : h->tev) CK(cudaEventCreate(&e));
  }
  h->timing = enable != 0;
  h->tcount = 0;
  return LMPC_OK;
}

extern "C" int lmpc_get_kernel_ms(lmpc_handle* h, double* ms3, int* nsolves) {
  if (!h || !ms3) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  ms3[0] = ms3[1] = ms3[2] = 0.0;
  const int n = std::min(h->tcount, kTimingRing);
  for (int k = 0; k < n; k++)
    for (int j = 0; j < 3; j++) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, h->tev[4 * (size_t)k + j], h->tev[4 * (size_t)k + j + 1]));
      ms3[j] += ms;
    }
  if (nsolves) *nsolves = n;
  return LMPC_OK;
}

// fp64 FMA throughput probe: 8 independent accumulator chains per thread, no memory traffic in the loop
__global__ void __launch_bounds__(256) lmpc_dfma_probe_kernel(double* sink, double a, double b, int iters) {
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = (double)(threadIdx.x + k);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = fma(acc[k], a, b);
  }
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) t += acc[k];
  if (t == 123.456) sink[0] = t;   // never true: keeps the chains alive
}

extern "C" int lmpc_measure_fp64_peak(lmpc_handle* h, double* tflops) {
  if (!h || !tflops) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  DevBuf sink;
  const int rc = dev_reserve(h, sink, sizeof(double));
  if (rc != LMPC_OK) return rc;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  lmpc_dfma_probe_kernel<<<blocks, threads, 0, h->stream>>>((double*)sink.p, 0.999999, 1e-9, 256);   // warm-up
  CK(cudaEventRecord(e0, h->stream));
  lmpc_dfma_probe_kernel<<<blocks, threads, 0, h->stream>>>((double*)sink.p, 0.999999, 1e-9, iters);
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink.p);
  *tflops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads / ((double)ms * 1e-3) * 1e-12;
  return LMPC_OK;
}

extern "C" int lmpc_synchronize(lmpc_handle* h) {
  if (!h) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return LMPC_OK;
}

// ------------------------------------------------------------------------------------------ safe set
static int upload_safe_set(lmpc_handle* h) {
  CK(cudaSetDevice(h->device));
  h->dev_laps.clear();
  size_t total_pts = 0;
  for (const HostLap& l : h->laps) total_pts += 3 * (size_t)l.n;
  if (total_pts == 0) return LMPC_OK;
  // slab layout per point set: ps | pe | J | xr[6] (doubles) then canon (ints); one allocation
  const size_t dbl = total_pts * 9, ints = total_pts;
  const size_t bytes = dbl * sizeof(double) + ints * sizeof(int);
  // a new slab every time: kernels already enqueued on the stream may still read the old one
  CK(cudaStreamSynchronize(h->stream));
  int rc = dev_reserve(h, h->slab, bytes);
  if (rc != LMPC_OK) return rc;
  std::vector<double> hd(dbl);
  std::vector<int> hi(ints);
  double* dbase = (double*)h->slab.p;
  int* ibase = (int*)((char*)h->slab.p + dbl * sizeof(double));
  size_t od = 0, oi = 0;
  std::vector<LmpcLapView> views(h->laps.size());
  for (size_t li = 0; li < h->laps.size(); li++) {
    const HostLap& l = h->laps[li];
    const size_t m = 3 * (size_t)l.n;
    LmpcLapView v;
    std::memcpy(&hd[od], l.ps.data(), m * sizeof(double)); v.ps = dbase + od; od += m;
    std::memcpy(&hd[od], l.pe.data(), m * sizeof(double)); v.pe = dbase + od; od += m;
    std::memcpy(&hd[od], l.J.data(), m * sizeof(double)); v.J = dbase + od; od += m;
    std::memcpy(&hd[od], l.xr.data(), 6 * m * sizeof(double)); v.xr = dbase + od; od += 6 * m;
    std::memcpy(&hi[oi], l.canon.data(), m * sizeof(int)); v.canon = ibase + oi; oi += m;
    v.m = (int)m; v.take = 0; v.out_off = 0;
    views[li] = v;
  }
  CK(cudaMemcpyAsync(dbase, hd.data(), dbl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(ibase, hi.data(), ints * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));   // hd / hi are stack-scoped
  for (size_t li = h->laps.size(); li-- > 0;) h->dev_laps.push_back(views[li]);   // newest first
  return LMPC_OK;
}

extern "C" int lmpc_safe_set_add_lap(lmpc_handle* h, int n, const double* x, const double* u, const double* k,
                                     const double* t, double L) {
  if (!h || n < 1 || !x) return LMPC_ERR_INVALID;
  HostLap lap;
  lap.n = n; lap.L = L;
  lap.x.assign(x, x + 6 * (size_t)n);   // u, k, t are read by the error-dynamics regression only
  if (u && k && t) { lap.u.assign(u, u + 2 * (size_t)n); lap.k.assign(k, k + n); lap.t.assign(t, t + n); }
  const size_t m = 3 * (size_t)n;
  lap.ps.resize(m); lap.pe.resize(m); lap.J.resize(m); lap.xr.resize(6 * m); lap.canon.resize(m);
  // SSTrajectory::process_lap_data (safe_set.cpp:116-137): x_repeat = [x - L e0, x, x + L e0],
  // J = [J + n - 1, J, J - n + 1] with J_j = n - 1 - j
  for (int rep = 0; rep < 3; rep++)
    for (int j = 0; j < n; j++) {
      const size_t q = (size_t)rep * n + j;
      for (int c = 0; c < 6; c++) lap.xr[6 * q + c] = x[6 * (size_t)j + c];
      lap.xr[6 * q] += (rep - 1) * L;
      lap.ps[q] = lap.xr[6 * q]; lap.pe[q] = lap.xr[6 * q + 1];
      lap.J[q] = (double)(n - 1 - j) + (1 - rep) * (double)(n - 1);
    }
  // exact-duplicate keys resolve to the first inserted index (trajectory_kd_tree.cpp:38)
  std::vector<int> order(m);
  for (size_t i = 0; i < m; i++) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    if (lap.ps[a] != lap.ps[b]) return lap.ps[a] < lap.ps[b];
    if (lap.pe[a] != lap.pe[b]) return lap.pe[a] < lap.pe[b];
    return a < b;
  });
  for (size_t i = 0; i < m;) {
    size_t j = i;
    while (j < m && lap.ps[order[j]] == lap.ps[order[i]] && lap.pe[order[j]] == lap.pe[order[i]]) { lap.canon[order[j]] = order[i]; j++; }
    i = j;
  }
  const size_t cap = (size_t)std::max(h->cfg.max_lap_stored, 1);
  if (h->laps.size() == cap) h->laps.erase(h->laps.begin());   // circular_buffer::push_back overwrites the oldest
  h->laps.push_back(std::move(lap));
  h->reg_dirty = true;
  return upload_safe_set(h);
}

static bool read_matrix(const std::string& path, int cols, std::vector<double>& out, int& rows) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return false;
  out.clear();
  double v;
  while (fscanf(f, "%lf", &v) == 1) out.push_back(v);
  fclose(f);
  if (out.empty() || out.size() % (size_t)cols) return false;
  rows = (int)(out.size() / (size_t)cols);
  return true;
}

extern "C" int lmpc_safe_set_load(lmpc_handle* h, const char* prefix, double L) {
  if (!h || !prefix) return LMPC_ERR_INVALID;
  std::vector<double> x, u, k, t;
  int nx = 0, nu = 0, nk = 0, nt = 0;
  const std::string p(prefix);
  if (!read_matrix(p + "_x.txt", 6, x, nx) || !read_matrix(p + "_u.txt", 2, u, nu) ||
      !read_matrix(p + "_k.txt", 1, k, nk) || !read_matrix(p + "_t.txt", 1, t, nt) || nx != nu || nx != nk || nx != nt) {
    h->err = "cannot load lap " + p;
    return LMPC_ERR_IO;
  }
  const int rc = lmpc_safe_set_add_lap(h, nx, x.data(), u.data(), k.data(), t.data(), L);
  if (rc == LMPC_OK) h->rec.lap_count++;   // SafeSetRecorder::load, safe_set.cpp:269: recorded laps are numbered after the loaded ones
  return rc;
}

extern "C" int lmpc_safe_set_clear(lmpc_handle* h) {
  if (!h) return LMPC_ERR_INVALID;
  h->laps.clear();
  h->dev_laps.clear();
  h->reg_dirty = true;
  return LMPC_OK;
}

extern "C" int lmpc_safe_set_num_laps(const lmpc_handle* h) { return h ? (int)h->laps.size() : 0; }

// ------------------------------------------------------------------------------------------ lap recorder
extern "C" int lmpc_recorder_config(lmpc_handle* h, int to_file, const char* file_prefix) {
  if (!h || (to_file && !file_prefix)) return LMPC_ERR_INVALID;
  h->rec.to_file = to_file != 0;
  h->rec.prefix = file_prefix ? file_prefix : "";
  return LMPC_OK;
}

extern "C" int lmpc_recorder_lap_count(const lmpc_handle* h) { return h ? h->rec.lap_count : 0; }

static bool write_matrix(const std::string& path, const std::vector<double>& v, int cols) {
  FILE* f = fopen(path.c_str(), "w");
  if (!f) return false;
  for (size_t i = 0; i < v.size(); i++) fprintf(f, "%.16e%c", v[i], ((i + 1) % (size_t)cols) ? ' ' : '\n');
  return fclose(f) == 0;
}

// SafeSetRecorder::step (safe_set.cpp:278-322)
extern "C" int lmpc_recorder_step(lmpc_handle* h, const double* x, const double* u, double k, double t, double L, int32_t* lap_added) {
  if (!h || !x || !u) return LMPC_ERR_INVALID;
  Recorder& r = h->rec;
  if (lap_added) *lap_added = 0;
  if (!r.last_x_valid) {   // :282-286 -- the very first sample only arms the wrap test
    r.last_px = x[0]; r.last_x_valid = true;
    return LMPC_OK;
  }
  int rc = LMPC_OK;
  if (r.last_px - x[0] > 0.5 * L) {   // :290 new lap
    if (r.initialized) {
      const int n = (int)r.k.size();
      rc = lmpc_safe_set_add_lap(h, n, r.x.data(), r.u.data(), r.k.data(), r.t.data(), L);
      if (rc == LMPC_OK && lap_added) *lap_added = 1;
      if (rc == LMPC_OK && r.to_file) {
        const std::string fn = r.prefix + "lap_" + std::to_string(r.lap_count);
        if (!write_matrix(fn + "_x.txt", r.x, 6) || !write_matrix(fn + "_u.txt", r.u, 2) || !write_matrix(fn + "_t.txt", r.t, 1) ||
            !write_matrix(fn + "_k.txt", r.k, 1)) { h->err = "cannot write lap " + fn; rc = LMPC_ERR_IO; }
      }
    } else r.initialized = true;
    r.lap_count++;
    r.x.clear(); r.u.clear(); r.k.clear(); r.t.clear();
  }
  // samples recorded before the first wrap are never used (the reference keeps appending to them and drops them at
  // the first wrap, :306-317); they are not stored here
  if (r.initialized) {
    r.x.insert(r.x.end(), x, x + 6); r.u.insert(r.u.end(), u, u + 2); r.k.push_back(k); r.t.push_back(t);
  }
  r.last_px = x[0];
  return rc;
}

// ------------------------------------------------------------------------------------------ error-dynamics regression
static void launch_regress(lmpc_handle* h, const LmpcRegPlan& plan, const LmpcRegItems& ri, int blocks) {
  // two resident blocks per SM (120 registers, no spills) measured level with three (80 registers, spills): profiles/README.md
  if (lmpc_reg_size_class(plan) == 5) {
    lmpc_regress_tiled_kernel<5, 2><<<blocks, 32 * LMPC_REG_WARPS, 0, h->stream>>>(plan, h->reg_view, ri);
  } else lmpc_regress_tiled_kernel<LMPC_REG_D><<<blocks, 32 * LMPC_REG_WARPS, 0, h->stream>>>(plan, h->reg_view, ri);
}

static int make_reg_plan(lmpc_handle* h, const lmpc_reg_spec* sp, LmpcRegPlan* plan) {
  if (lmpc_make_reg_plan(sp, plan)) return LMPC_OK;
  h->err = "regression spec rejected (1..6 regressions, indices in range, dist_max > 0, ridge > 0, sign = +-1)";
  return LMPC_ERR_INVALID;
}

// points of all stored laps, oldest lap first (the order SafeSetManager::query concatenates them, safe_set.cpp:196-203)
static int ensure_reg_slab(lmpc_handle* h) {
  if (!h->reg_dirty) return LMPC_OK;
  size_t M = 0;
  for (const HostLap& l : h->laps) {
    if (l.u.empty()) { h->err = "a stored lap has no u / k / t: the regression needs them"; return LMPC_ERR_INVALID; }
    M += (size_t)(l.n - 1);
  }
  h->reg_view = LmpcRegView{nullptr, nullptr, 0, 0, -1};
  if (M == 0) { h->reg_dirty = false; return LMPC_OK; }
  // slab, by column with stride ld (M rounded up to 32): Z [8][ld] | E [6][ld] | Xn [6][ld] | kappa [ld] | dt [ld]
  // (the last three only feed the prepare kernel)
  const size_t ld = (M + 31) & ~(size_t)31;
  std::vector<double> hd(22 * ld, 0.0);
  double* Z = hd.data(); double* Xn = Z + 14 * ld; double* kp = Xn + 6 * ld; double* dt = kp + ld;
  // The samples are stored SORTED by one component of (x, u) -- the one with the widest spread among e_psi, v_x, v_y,
  // omega and the controls (the abscissa and e_y are not dynamics inputs) -- so that a query's scan can be cut to the
  // window |z - z_q| < dist_max of that component.  The order of the sums changes, nothing else.
  struct Src { const HostLap* l; int j; };
  std::vector<Src> src; src.reserve(M);
  for (const HostLap& l : h->laps) for (int j = 0; j + 1 < l.n; j++) src.push_back(Src{&l, j});
  auto comp = [](const Src& s_, int c) { return c < 6 ? s_.l->x[6 * (size_t)s_.j + c] : s_.l->u[2 * (size_t)s_.j + (c - 6)]; };
  int sort_dim = 2; double best = -1.0;
  for (int c = 2; c < 8; c++) {
    double lo = 1e300, hi = -1e300;
    for (const Src& s_ : src) { const double v = comp(s_, c); lo = std::min(lo, v); hi = std::max(hi, v); }
    if (hi - lo > best) { best = hi - lo; sort_dim = c; }
  }
  std::stable_sort(src.begin(), src.end(), [&](const Src& a, const Src& b) { return comp(a, sort_dim) < comp(b, sort_dim); });
  size_t p = 0;
  for (const Src& s_ : src) {
    const HostLap& l = *s_.l; const int j = s_.j;
    for (int c = 0; c < 6; c++) { Z[c * ld + p] = l.x[6 * (size_t)j + c]; Xn[c * ld + p] = l.x[6 * (size_t)(j + 1) + c]; }
    Z[6 * ld + p] = l.u[2 * (size_t)j]; Z[7 * ld + p] = l.u[2 * (size_t)j + 1];
    kp[p] = l.k[j]; dt[p] = l.t[j + 1] - l.t[j];
    p++;
  }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));   // kernels in flight may still read the old slab
  int rc = dev_reserve(h, h->reg_slab, sizeof(double) * 22 * ld);
  if (rc != LMPC_OK) return rc;
  double* d = (double*)h->reg_slab.p;
  CK(cudaMemcpyAsync(d, hd.data(), sizeof(double) * 22 * ld, cudaMemcpyHostToDevice, h->stream));
  const int threads = 128, blocks = (int)((M + threads - 1) / threads);
  lmpc_reg_prepare_kernel<<<blocks, threads, 0, h->stream>>>(h->M, (int)M, (int)ld, d, d + 14 * ld, d + 20 * ld, d + 21 * ld, d + 8 * ld);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));   // hd is stack-scoped
  h->reg_view = LmpcRegView{d, d + 8 * ld, (int)M, (int)ld, sort_dim};
  h->reg_dirty = false;
  return LMPC_OK;
}

extern "C" int lmpc_set_error_dynamics(lmpc_handle* h, const lmpc_reg_spec* spec) {
  if (!h) return LMPC_ERR_INVALID;
  if (!spec) { h->reg_in_tick = false; return LMPC_OK; }
  int rc = make_reg_plan(h, spec, &h->reg_plan);
  if (rc != LMPC_OK) return rc;
  h->reg_in_tick = true;
  return LMPC_OK;
}

extern "C" int lmpc_safe_set_regress_batch(lmpc_handle* h, int n, const lmpc_reg_spec* spec, const double* xq, const double* uq,
                                           double* A, double* Bm, double* C, int32_t* npts, int memspace) {
  if (!h || n < 1 || !xq || !uq || !A || !Bm || !C) return LMPC_ERR_INVALID;
  LmpcRegPlan plan;
  int rc = make_reg_plan(h, spec, &plan);
  if (rc == LMPC_OK) rc = ensure_reg_slab(h);
  if (rc != LMPC_OK) return rc;
  CK(cudaSetDevice(h->device));
  const size_t nz = (size_t)n;
  const double *dxq = xq, *duq = uq; double *dA = A, *dB = Bm, *dC = C; int32_t* dn = npts;
  if (memspace == LMPC_MEM_HOST) {
    rc = dev_reserve(h, h->ws_reg, sizeof(double) * 62 * nz + sizeof(int32_t) * LMPC_REG_MAX_OUT * nz);
    if (rc != LMPC_OK) return rc;
    double* d = (double*)h->ws_reg.p;
    double* q6 = d; double* q2 = d + 6 * nz; dA = d + 8 * nz; dB = dA + 36 * nz; dC = dB + 12 * nz; dn = npts ? (int32_t*)(dC + 6 * nz) : nullptr;
    CK(cudaMemcpyAsync(q6, xq, sizeof(double) * 6 * nz, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(q2, uq, sizeof(double) * 2 * nz, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dA, A, sizeof(double) * 36 * nz, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dB, Bm, sizeof(double) * 12 * nz, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dC, C, sizeof(double) * 6 * nz, cudaMemcpyHostToDevice, h->stream));
    dxq = q6; duq = q2;
  }
  if (h->reg_view.M > 0) {
    LmpcRegItems ri{};
    ri.n = n; ri.tick = 0; ri.xq = dxq; ri.uq = duq; ri.A = dA; ri.Bm = dB; ri.C = dC; ri.npts = dn;
    const int blocks = (n + LMPC_REG_WARPS - 1) / LMPC_REG_WARPS;
    launch_regress(h, plan, ri, blocks);
    h->launches++;
    CK(cudaGetLastError());
  } else if (dn) CK(cudaMemsetAsync(dn, 0, sizeof(int32_t) * (size_t)plan.n_out * nz, h->stream));
  if (memspace == LMPC_MEM_HOST) {
    CK(cudaMemcpyAsync(A, dA, sizeof(double) * 36 * nz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(Bm, dB, sizeof(double) * 12 * nz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(C, dC, sizeof(double) * 6 * nz, cudaMemcpyDeviceToHost, h->stream));
    if (npts) CK(cudaMemcpyAsync(npts, dn, sizeof(int32_t) * (size_t)plan.n_out * nz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LMPC_OK;
}

// newest -> oldest while num_total < max_total (safe_set.cpp:164); the columns a lap contributes and
// where they land do not depend on the query
static int make_lap_table(lmpc_handle* h, int max_total, int per_lap, LmpcLapTable* tab) {
  tab->n_used = 0; tab->count = 0;
  int total = 0;
  for (size_t j = 0; j < h->dev_laps.size() && total < max_total; j++) {
    if (tab->n_used >= LMPC_MAX_LAPS_USED) {   // never truncate silently: the result would differ from the reference's
      h->err = "safe-set query needs more than " + std::to_string(LMPC_MAX_LAPS_USED) + " laps (num_ss_pts / num_ss_pts_per_lap too large)";
      return LMPC_ERR_INVALID;
    }
    LmpcLapView v = h->dev_laps[j];
    v.take = std::min(per_lap, v.m);
    v.out_off = total;
    total += v.take;
    tab->lap[tab->n_used++] = v;
  }
  tab->count = std::min(total, max_total);
  return LMPC_OK;
}

static int launch_ss_query(lmpc_handle* h, const LmpcLapTable& tab, int B, const double* d_query, int qstride,
                           int max_total, int pad_to, double* d_ssx, double* d_ssj) {
  if (tab.n_used == 0) return LMPC_OK;
  const int warps = B * tab.n_used, threads = 128;
  const int blocks = (warps * 32 + threads - 1) / threads;
  lmpc_ss_query_kernel<<<blocks, threads, 0, h->stream>>>(tab, B, d_query, qstride, max_total, pad_to, d_ssx, d_ssj);
  h->launches++;
  CK(cudaGetLastError());
  return LMPC_OK;
}

extern "C" int lmpc_safe_set_query_batch(lmpc_handle* h, int B, const double* query, int max_total, int max_per_lap,
                                         double* ss_x, double* ss_j, int32_t* count, int memspace) {
  if (!h || B < 1 || !query || !ss_x || !ss_j || max_total < 1 || max_per_lap < 1) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  LmpcLapTable tab;
  { const int rc = make_lap_table(h, max_total, max_per_lap, &tab); if (rc != LMPC_OK) return rc; }
  const size_t nq = 2 * (size_t)B, nx = 6 * (size_t)max_total * B, nj = (size_t)max_total * B;
  const double* dq = query; double* dx = ss_x; double* dj = ss_j;
  if (memspace == LMPC_MEM_HOST) {
    int rc = dev_reserve(h, h->st_in, sizeof(double) * nq);
    if (rc == LMPC_OK) rc = dev_reserve(h, h->st_out, sizeof(double) * (nx + nj));
    if (rc != LMPC_OK) return rc;
    CK(cudaMemcpyAsync(h->st_in.p, query, sizeof(double) * nq, cudaMemcpyHostToDevice, h->stream));
    dq = (const double*)h->st_in.p; dx = (double*)h->st_out.p; dj = dx + nx;
  }
  // pad_to == count: no padding in the raw query (SafeSetManager::query returns `count` columns)
  int rc = launch_ss_query(h, tab, B, dq, 2, max_total, max_total, dx, dj);
  if (rc != LMPC_OK) return rc;
  if (memspace == LMPC_MEM_HOST) {
    // only the `count` columns the query found travel back: the caller's columns beyond them stay untouched (header contract)
    if (tab.count > 0) {
      CK(cudaMemcpy2DAsync(ss_x, sizeof(double) * 6 * (size_t)max_total, dx, sizeof(double) * 6 * (size_t)max_total,
                           sizeof(double) * 6 * (size_t)tab.count, (size_t)B, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpy2DAsync(ss_j, sizeof(double) * (size_t)max_total, dj, sizeof(double) * (size_t)max_total,
                           sizeof(double) * (size_t)tab.count, (size_t)B, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    if (count) for (int b = 0; b < B; b++) count[b] = tab.count;
  } else if (count) {
    std::vector<int32_t> hc((size_t)B, tab.count);
    CK(cudaMemcpyAsync(count, hc.data(), sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LMPC_OK;
}

// ------------------------------------------------------------------------------------------ model
static int model_batch(lmpc_handle* h, int n, const double* x, const double* u, const double* kappa, const double* dt,
                       double* A, double* Bm, double* g, double* xnext, int memspace, bool jac) {
  if (!h || n < 1 || !x || !u || !kappa || !dt) return LMPC_ERR_INVALID;
  if (jac && (!A || !Bm || !g)) return LMPC_ERR_INVALID;
  if (!jac && !xnext) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  const size_t nn = (size_t)n;
  const double *dx = x, *du = u, *dk = kappa, *dd = dt;
  double *dA = A, *dB = Bm, *dg = g, *dxn = xnext;
  if (memspace == LMPC_MEM_HOST) {
    int rc = dev_reserve(h, h->st_in, sizeof(double) * nn * 10);
    if (rc == LMPC_OK) rc = dev_reserve(h, h->st_out, sizeof(double) * nn * 60);
    if (rc != LMPC_OK) return rc;
    double* si = (double*)h->st_in.p;
    CK(cudaMemcpyAsync(si, x, sizeof(double) * 6 * nn, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(si + 6 * nn, u, sizeof(double) * 2 * nn, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(si + 8 * nn, kappa, sizeof(double) * nn, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(si + 9 * nn, dt, sizeof(double) * nn, cudaMemcpyHostToDevice, h->stream));
    dx = si; du = si + 6 * nn; dk = si + 8 * nn; dd = si + 9 * nn;
    double* so = (double*)h->st_out.p;
    dA = so; dB = so + 36 * nn; dg = so + 48 * nn; dxn = so + 54 * nn;
  }
  const int threads = 64, blocks = (n + threads - 1) / threads;
  if (jac) lmpc_linearise_items_kernel<<<blocks, threads, 0, h->stream>>>(h->M, n, dx, du, dk, dd, dA, dB, dg, (xnext || memspace == LMPC_MEM_HOST) ? dxn : nullptr);
  else lmpc_step_items_kernel<<<blocks, threads, 0, h->stream>>>(h->M, n, dx, du, dk, dd, dxn);
  h->launches++;
  CK(cudaGetLastError());
  if (memspace == LMPC_MEM_HOST) {
    if (jac) {
      CK(cudaMemcpyAsync(A, dA, sizeof(double) * 36 * nn, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(Bm, dB, sizeof(double) * 12 * nn, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(g, dg, sizeof(double) * 6 * nn, cudaMemcpyDeviceToHost, h->stream));
    }
    if (xnext) CK(cudaMemcpyAsync(xnext, dxn, sizeof(double) * 6 * nn, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LMPC_OK;
}

// BaseVehicleModel::to_base_control / from_base_control (single_track_planar_model.cpp:390-417)
static int control_map(lmpc_handle* h, int n, int dir, const double* in, double* out, int memspace) {
  if (!h || n < 1 || !in || !out) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  const size_t wi = dir ? 3 : 2, wo = dir ? 2 : 3;
  const double* din = in; double* dout = out;
  if (memspace == LMPC_MEM_HOST) {
    int rc = dev_reserve(h, h->st_in, sizeof(double) * wi * (size_t)n);
    if (rc == LMPC_OK) rc = dev_reserve(h, h->st_out, sizeof(double) * wo * (size_t)n);
    if (rc != LMPC_OK) return rc;
    CK(cudaMemcpyAsync(h->st_in.p, in, sizeof(double) * wi * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    din = (const double*)h->st_in.p; dout = (double*)h->st_out.p;
  }
  lmpc_control_map_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(n, dir, din, dout);
  h->launches++;
  CK(cudaGetLastError());
  if (memspace == LMPC_MEM_HOST) {
    CK(cudaMemcpyAsync(out, dout, sizeof(double) * wo * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LMPC_OK;
}
extern "C" int lmpc_to_base_control_batch(lmpc_handle* h, int n, const double* u, double* u_base, int memspace) { return control_map(h, n, 0, u, u_base, memspace); }
extern "C" int lmpc_from_base_control_batch(lmpc_handle* h, int n, const double* u_base, double* u, int memspace) { return control_map(h, n, 1, u_base, u, memspace); }

// A handle for the model functions alone (the vehicle_model_factory product: no MPC configuration exists yet when the node
// builds the model, racing_mpc_node.cpp:43-47): a minimal tracking configuration stands in.
extern "C" int lmpc_model_create(const lmpc_vehicle_params* vehicle, int device_ordinal, lmpc_handle** out) {
  if (!vehicle || !out) return LMPC_ERR_INVALID;
  lmpc_mpc_config c;
  memset(&c, 0, sizeof c);
  c.N = 3; c.learning = 0; c.q_boundary = 1.0;
  c.R[0] = c.R[3] = c.R_d[0] = c.R_d[3] = 1.0;
  for (int k = 0; k < 6; k++) { c.x_max[k] = INFINITY; c.x_min[k] = -INFINITY; }
  for (int k = 0; k < 2; k++) { c.u_max[k] = INFINITY; c.u_min[k] = -INFINITY; }
  c.num_ss_pts = 0; c.num_ss_pts_per_lap = 1; c.max_lap_stored = 1;
  return lmpc_create(&c, vehicle, device_ordinal, 1, out);
}

// columns the tick's safe-set query finds before padding (racing_mpc.cpp:249-262): min(num_ss_pts, sum over the newest laps
// of min(num_ss_pts_per_lap, points of the lap)); what out["ss_x"].size2() is in the reference
extern "C" int lmpc_safe_set_tick_count(lmpc_handle* h, int32_t* count) {
  if (!h || !count) return LMPC_ERR_INVALID;
  *count = 0;
  const int K = h->cfg.num_ss_pts;
  if (K < 1 || h->dev_laps.empty()) return LMPC_OK;
  LmpcLapTable tab;
  const int rc = make_lap_table(h, K, std::max(1, (int)h->cfg.num_ss_pts_per_lap), &tab);
  if (rc != LMPC_OK) return rc;
  *count = tab.count;
  return LMPC_OK;
}

extern "C" int lmpc_discrete_dynamics_batch(lmpc_handle* h, int n, const double* x, const double* u, const double* kappa,
                                            const double* dt, double* x_next, int memspace) {
  return model_batch(h, n, x, u, kappa, dt, nullptr, nullptr, nullptr, x_next, memspace, false);
}

extern "C" int lmpc_linearise_batch(lmpc_handle* h, int n, const double* x, const double* u, const double* kappa,
                                    const double* dt, double* A, double* Bm, double* g, double* x_next, int memspace) {
  return model_batch(h, n, x, u, kappa, dt, A, Bm, g, x_next, memspace, true);
}

// ------------------------------------------------------------------------------------------ solve
static void launch_qp(lmpc_handle* h, const LmpcQpBatch& a) {
  lmpc_qp_launch(h->P.NW, qp_kpl(h), qp_fixed_n(h), a.B, h->qp_smem, h->stream, h->P, a);
}

// Device views of one batch: caller's buffers (LMPC_MEM_DEVICE) or the handle's staging area (LMPC_MEM_HOST).
struct DevIO {
  const double* din[11];
  double* dout[7];
  int32_t *d_status, *d_iters;
  double *ssx, *ssj;
  size_t nin[11], nout[7];
  // host callers: where the safe-set columns go.  They are final when K2 ends, so their D2H copies are issued on the
  // side stream and travel while the QP kernel runs (5.5 of the 7.9 MB a 1024-instance tick returns).
  double *h_ssx = nullptr, *h_ssj = nullptr;
  mutable bool ss_copied = false;
  // host callers whose output buffers are one pinned, device-mapped arena laid out like the staging area: the QP kernel
  // stores every instance's result straight into it (byte offset host - staging); no D2H copy of those arrays follows
  long long host_mirror_off = 0;
  bool host_mirror = false;
};

static int check_batch_args(lmpc_handle* h, int B, const lmpc_batch_in* in, const lmpc_batch_out* out) {
  if (!h || !in || !out || B < 1) return LMPC_ERR_INVALID;
  if (B > h->max_batch) return LMPC_ERR_CAPACITY;
  if (!in->x_ic || !in->u_ic || !in->X_ref || !in->U_ref || !in->T_ref || !in->bound_left || !in->bound_right ||
      !in->curvatures || !in->vel_ref || !in->total_length)
    return LMPC_ERR_INVALID;
  if (!out->X_optm || !out->U_optm || !out->dU_optm || !out->status || !out->iters) return LMPC_ERR_INVALID;
  return LMPC_OK;
}

// H2D of the 11 inputs (host callers) / pointer pass-through (device callers)
static int stage_in(lmpc_handle* h, int B, const lmpc_batch_in* in, const lmpc_batch_out* out, int memspace, DevIO& io) {
  const size_t Bz = (size_t)B, N = (size_t)h->P.N, NS = (size_t)h->P.NS, K = (size_t)h->P.K;
  // element counts of the 11 inputs and 7 double outputs, in struct order
  const size_t nin[11] = {6 * Bz, 2 * Bz, 6 * N * Bz, 2 * NS * Bz, NS * Bz, N * Bz, N * Bz, N * Bz, N * Bz, Bz, 2 * NS * Bz};
  const size_t nout[7] = {6 * N * Bz, 2 * NS * Bz, 2 * NS * Bz, K * Bz, 6 * K * Bz, K * Bz, Bz};
  const double* hin[11] = {in->x_ic, in->u_ic, in->X_ref, in->U_ref, in->T_ref, in->bound_left, in->bound_right,
                           in->curvatures, in->vel_ref, in->total_length, in->U_warm};
  double* hout[7] = {out->X_optm, out->U_optm, out->dU_optm, out->convex_combi_optm, out->ss_x, out->ss_j, out->cost};
  for (int k = 0; k < 11; k++) io.nin[k] = nin[k];
  for (int k = 0; k < 7; k++) io.nout[k] = nout[k];
  io.d_status = out->status; io.d_iters = out->iters;
  if (memspace == LMPC_MEM_HOST) {
    size_t tin = 0, tout = 0;
    for (int k = 0; k < 11; k++) tin += nin[k];
    for (int k = 0; k < 7; k++) tout += nout[k];
    int rc = dev_reserve(h, h->st_in, sizeof(double) * tin);
    if (rc == LMPC_OK) rc = dev_reserve(h, h->st_out, sizeof(double) * tout + 2 * sizeof(int32_t) * Bz);
    if (rc != LMPC_OK) return rc;
    // the staging area holds the inputs back to back in struct order; host buffers that are adjacent in the same order
    // (one arena, e.g. BatchedRacingMPC.alloc_host_inputs) travel as one copy
    double* p = (double*)h->st_in.p;
    const double* run_src = nullptr; double* run_dst = nullptr; size_t run_n = 0;
    for (int k = 0; k < 11; k++) {
      if (hin[k]) {
        io.din[k] = p;
        if (run_n && hin[k] == run_src + run_n && p == run_dst + run_n) run_n += nin[k];
        else {
          if (run_n) CK(cudaMemcpyAsync(run_dst, run_src, sizeof(double) * run_n, cudaMemcpyHostToDevice, h->stream));
          run_src = hin[k]; run_dst = p; run_n = nin[k];
        }
      } else io.din[k] = nullptr;
      p += nin[k];
    }
    if (run_n) CK(cudaMemcpyAsync(run_dst, run_src, sizeof(double) * run_n, cudaMemcpyHostToDevice, h->stream));
    double* q = (double*)h->st_out.p;
    for (int k = 0; k < 7; k++) { io.dout[k] = q; q += nout[k]; }
    io.d_status = (int32_t*)q; io.d_iters = io.d_status + Bz;
    if (!out->convex_combi_optm) io.dout[3] = nullptr;
    if (!out->cost) io.dout[6] = nullptr;
    io.ssx = (double*)h->st_out.p + nout[0] + nout[1] + nout[2] + nout[3]; io.ssj = io.ssx + nout[4];
    io.h_ssx = out->ss_x; io.h_ssj = out->ss_j;
    // zero-copy results: one offset must map the staging X | U | dU | lambda | cost | status | iters onto the caller's
    // buffers (true for BatchedRacingMPC.alloc_host_outputs' arena) and the memory must be pinned and device-mapped
    {
      const char* e = getenv("LMPC_HOST_MIRROR");
      cudaPointerAttributes at;
      if (!(e && atoi(e) == 0) && out->X_optm && cudaPointerGetAttributes(&at, out->X_optm) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
        const long long off = (const char*)at.devicePointer - (const char*)io.dout[0];
        auto same = [&](const void* hp, const void* dp) { return hp && dp && ((const char*)hp - (const char*)out->X_optm) == ((const char*)dp - (const char*)io.dout[0]); };
        const bool lam_ok = (!h->P.learning) || !out->convex_combi_optm || same(out->convex_combi_optm, io.dout[3]);
        if (same(out->U_optm, io.dout[1]) && same(out->dU_optm, io.dout[2]) && lam_ok && (!out->cost || same(out->cost, io.dout[6])) &&
            same(out->status, io.d_status) && same(out->iters, io.d_iters)) {
          io.host_mirror = true; io.host_mirror_off = off;
        }
      } else cudaGetLastError();   // a pageable pointer makes cudaPointerGetAttributes fail on some drivers: not an error here
    }
  } else {
    for (int k = 0; k < 11; k++) io.din[k] = hin[k];
    for (int k = 0; k < 7; k++) io.dout[k] = hout[k];
    // the safe-set columns go straight to the caller's ss_x / ss_j when given, else to the workspace
    io.ssx = out->ss_x ? out->ss_x : (double*)h->ws_ssx.p;
    io.ssj = out->ss_j ? out->ss_j : (double*)h->ws_ssj.p;
  }
  return LMPC_OK;
}

// D2H of the outputs + stream synchronise (host callers only)
static int stage_out(lmpc_handle* h, int B, const lmpc_batch_out* out, int memspace, const DevIO& io) {
  if (memspace != LMPC_MEM_HOST) return LMPC_OK;
  const bool learn = h->P.learning != 0;
  double* hout[7] = {out->X_optm, out->U_optm, out->dU_optm, out->convex_combi_optm, out->ss_x, out->ss_j, out->cost};
  const double* dsrc[7] = {io.dout[0], io.dout[1], io.dout[2], io.dout[3], io.ssx, io.ssj, io.dout[6]};
  // adjacent (host, device) pairs are merged into one copy, as on the way in
  char* run_dst = nullptr; const char* run_src = nullptr; size_t run_b = 0;
  auto push = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    if (run_b && (char*)dst == run_dst + run_b && (const char*)src == run_src + run_b) { run_b += bytes; return cudaSuccess; }
    cudaError_t e = run_b ? cudaMemcpyAsync(run_dst, run_src, run_b, cudaMemcpyDeviceToHost, h->stream) : cudaSuccess;
    run_dst = (char*)dst; run_src = (const char*)src; run_b = bytes;
    return e;
  };
  for (int k = 0; k < 7; k++) {
    if (!hout[k] || !dsrc[k]) continue;
    if ((k == 3 || k == 4 || k == 5) && !learn) continue;
    if ((k == 4 || k == 5) && io.ss_copied) continue;   // already on their way (side stream)
    if (io.host_mirror && k != 4 && k != 5) continue;    // stored by the QP kernel itself (zero-copy), complete at the synchronise below
    CK(push(hout[k], dsrc[k], sizeof(double) * io.nout[k]));
  }
  if (!io.host_mirror) {
    CK(push(out->status, io.d_status, sizeof(int32_t) * (size_t)B));
    CK(push(out->iters, io.d_iters, sizeof(int32_t) * (size_t)B));
  }
  if (run_b) CK(cudaMemcpyAsync(run_dst, run_src, run_b, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (io.ss_copied) CK(cudaStreamSynchronize(h->side));
  return LMPC_OK;
}

// The three kernels of one tick.  X_lin / U_lin: linearisation point; U0: initial controls of the interior point;
// first: also write the safe-set query point and run the k-NN (later SQP iterations keep the columns of the first);
// skip: optional per-instance mask (SQP instances that have converged).
static int run_tick_kernels(lmpc_handle* h, int B, const DevIO& io, const double* X_lin, const double* U_lin, const double* U0,
                            bool first, const int* skip, int* ss_count_io) {
  const size_t N = (size_t)h->P.N, NS = (size_t)h->P.NS, K = (size_t)h->P.K;
  const bool learn = h->P.learning != 0;
  double* abg = (double*)h->ws_abg.p; double* cen = (double*)h->ws_cen.p;
  cudaEvent_t* tev = (h->timing && !h->tev.empty()) ? &h->tev[4 * (size_t)(h->tcount % kTimingRing)] : nullptr;
  if (tev) CK(cudaEventRecord(tev[0], h->stream));
  // K2: safe-set query at X_ref[:, N-1] (racing_mpc.cpp:249-255), padded to K columns (:263-272).  Independent of K1:
  // forked onto the side stream here, joined before K3.
  const bool fork_ss = learn && first;
  const bool per_agent = h->ag_on && B == h->ag.B;
  if (fork_ss && per_agent) {
    // every agent queries its OWN laps (lmpc_agents_query_kernel); the column count is per instance
    const int per_lap = std::max(1, (int)h->cfg.num_ss_pts_per_lap);
    const int max_used = std::min(h->ag.slots, ((int)K + per_lap - 1) / per_lap);
    CK(cudaEventRecord(h->ev_fork, h->stream));
    CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    const int warps = B * max_used, threads = 128, blocks = (warps * 32 + threads - 1) / threads;
    lmpc_agents_query_kernel<<<blocks, threads, 0, h->side>>>(h->ag, max_used, per_lap, (int)N, io.din[0], X_lin, io.din[9], nullptr, (int)K, (int)K,
                                                               io.ssx, io.ssj, (int*)h->ag_cnt.p);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev_join, h->side));
    *ss_count_io = (int)K;
  } else if (fork_ss) {
    LmpcLapTable tab;
    int rc = make_lap_table(h, (int)K, h->cfg.num_ss_pts_per_lap, &tab);
    if (rc != LMPC_OK) return rc;
    *ss_count_io = tab.count;
    if (tab.n_used > 0) {
      CK(cudaEventRecord(h->ev_fork, h->stream));
      CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
      const int warps = B * tab.n_used, threads = 128, blocks = (warps * 32 + threads - 1) / threads;
      lmpc_ss_query_tick_kernel<<<blocks, threads, 0, h->side>>>(tab, B, (int)N, io.din[0], X_lin, io.din[9], (int)K, (int)K, io.ssx, io.ssj);
      h->launches++;
      CK(cudaGetLastError());
      CK(cudaEventRecord(h->ev_join, h->side));
    }
  }
  // K1: linearise
  {
    const int n = B * (int)NS, threads = LMPC_K1_THREADS, blocks = (n + threads - 1) / threads;
    lmpc_linearise_kernel<<<blocks, threads, 0, h->stream>>>(h->M, B, (int)N, io.din[0], X_lin, U_lin, io.din[4], io.din[7], io.din[9], abg,
                                                             first ? cen : nullptr, skip);
    h->launches++;
    CK(cudaGetLastError());
  }
  // KR: error-dynamics regression on every stage's (A, B, g) (optional, lmpc_set_error_dynamics)
  if (h->reg_in_tick) {
    int rc = ensure_reg_slab(h);
    if (rc != LMPC_OK) return rc;
    if (h->reg_view.M > 0) {
      LmpcRegItems ri{};
      ri.n = B * (int)NS; ri.tick = 1; ri.B = B; ri.N = (int)N; ri.x_ic = io.din[0]; ri.X_ref = X_lin; ri.U_ref = U_lin;
      ri.total_length = io.din[9]; ri.ABg = abg; ri.skip = skip;
      const int blocks = (ri.n + LMPC_REG_WARPS - 1) / LMPC_REG_WARPS;
      launch_regress(h, h->reg_plan, ri, blocks);
      h->launches++;
      CK(cudaGetLastError());
    }
  }
  if (tev) CK(cudaEventRecord(tev[1], h->stream));
  if (fork_ss) CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));   // a never-recorded event is complete: no-op when nothing was forked
  if (tev) CK(cudaEventRecord(tev[2], h->stream));
  // K3: QP
  LmpcQpBatch a;
  a.x_ic = io.din[0]; a.u_ic = io.din[1]; a.U0 = U0; a.T_ref = io.din[4];
  a.bl = io.din[5]; a.br = io.din[6]; a.vref = io.din[8]; a.ABg = abg; a.ssx = io.ssx; a.ssj = io.ssj; a.cen = cen;
  a.X = io.dout[0]; a.U = io.dout[1]; a.dU = io.dout[2]; a.lam = io.dout[3]; a.cost = io.dout[6];
  a.status = io.d_status; a.iters = io.d_iters; a.ss_count = *ss_count_io; a.B = B; a.skip = skip;
  a.ss_count_v = (learn && per_agent) ? (const int*)h->ag_cnt.p : nullptr;
  a.n_mirror = 0; a.done = nullptr; a.seq = 0; a.mirror_all = 0;
  if (io.host_mirror && first && !skip && h->gat.active_set < 0) {
    a.n_mirror = 1; a.mirror_off[0] = io.host_mirror_off; a.peer_flag[0] = nullptr; a.mirror_all = 1;
  }
  if (h->gat.active_set >= 0) {   // lmpc_solve_gather_batch: X, U, dU, cost, status live in this rank's block of the gather buffer
    const lmpc_handle::Gather& G = h->gat;
    for (int p = 0; p < G.world; p++) {
      if (p == G.rank) continue;
      a.mirror_off[a.n_mirror] = (long long)((char*)G.peer[p] - (char*)G.base);
      a.peer_flag[a.n_mirror] = (unsigned long long*)((char*)G.peer[p] + G.flags_off) + G.rank;
      a.n_mirror++;
    }
    a.done = (unsigned int*)((char*)G.base + G.flags_off + sizeof(unsigned long long) * LMPC_MAX_PEERS);
    a.seq = G.seq;
  }
  a.scratch = (double*)h->ws_qpscr.p;
  launch_qp(h, a);
  h->launches++;
  CK(cudaGetLastError());
  if (tev) { CK(cudaEventRecord(tev[3], h->stream)); h->tcount++; }
  // host callers: the safe-set columns leave on the side stream now (it is idle once K2 has ended), behind the QP kernel.
  // Issued after the QP launch so that a pageable destination, which makes the copy call block, cannot delay the launches.
  if (fork_ss && (io.h_ssx || io.h_ssj)) {
    if (io.h_ssx) CK(cudaMemcpyAsync(io.h_ssx, io.ssx, sizeof(double) * io.nout[4], cudaMemcpyDeviceToHost, h->side));
    if (io.h_ssj) CK(cudaMemcpyAsync(io.h_ssj, io.ssj, sizeof(double) * io.nout[5], cudaMemcpyDeviceToHost, h->side));
    io.ss_copied = true;
  }
  return LMPC_OK;
}

extern "C" int lmpc_solve_batch(lmpc_handle* h, int B, const lmpc_batch_in* in, const lmpc_batch_out* out, int memspace) {
  int rc = check_batch_args(h, B, in, out);
  if (rc != LMPC_OK) return rc;
  CK(cudaSetDevice(h->device));
  DevIO io;
  rc = stage_in(h, B, in, out, memspace, io);
  if (rc != LMPC_OK) return rc;
  int ss_count = 0;
  rc = run_tick_kernels(h, B, io, io.din[2], io.din[3], io.din[10] ? io.din[10] : io.din[3], true, nullptr, &ss_count);
  if (rc != LMPC_OK) return rc;
  return stage_out(h, B, out, memspace, io);
}

// ------------------------------------------------------------------------------------------ multi-GPU result exchange
// One process per GPU; instances are sharded and independent, the only exchange is "every rank receives every rank's
// converged trajectories" (SURVEY.md 8e).  Instead of a collective AFTER the solve, the QP kernel's epilogue stores each
// instance's result into every peer's gather buffer through NVLink peer mappings (lmpc_qp_kernel.cuh) and the last CTA
// publishes a sequence number in the peers' flag arrays; a one-warp wait kernel on the receiving side spins on LOCAL
// memory.  No SM-resident communication kernel runs beside the solve, no packing kernel, no extra launch on the sender.
// Waits (on the device, one lane per peer) until every peer has published sequence number >= seq in this rank's LOCAL
// flag array.  A peer that never arrives would hang the stream: after `timeout_cycles` the lane gives up and records
// 1 + peer in *err (checked by the host at the next synchronisation).
__global__ void lmpc_gather_wait_kernel(const unsigned long long* __restrict__ flags, int world, int rank, unsigned long long seq,
                                        long long timeout_cycles, int* err) {
  const int p = (int)threadIdx.x;
  if (p >= world || p == rank) return;
  const long long t0 = clock64();
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + p) : "memory");
    if (v >= seq) break;
    if (clock64() - t0 > timeout_cycles) { atomicExch(err, 1 + p); break; }
    __nanosleep(100);
  }
}


static size_t gather_slab_bytes(const lmpc_handle* h, int B) {
  const size_t N = (size_t)h->P.N, NS = (size_t)h->P.NS, Bz = (size_t)B;
  return sizeof(double) * (Bz * (6 * N + 4 * NS + 1) + (Bz + 1) / 2);   // X | U | dU | cost | status (int32, padded): solver.alloc_device_outputs' slab
}

extern "C" int lmpc_gather_init(lmpc_handle* h, int world, int rank, int B, int sets, void* ipc_handle_out, size_t* slab_bytes) {
  if (!h || world < 1 || world > LMPC_MAX_PEERS || rank < 0 || rank >= world || B < 1 || B > h->max_batch || sets < 1 || sets > 4 || !ipc_handle_out)
    return LMPC_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == LMPC_IPC_HANDLE_BYTES, "IPC handle size");
  CK(cudaSetDevice(h->device));
  lmpc_handle::Gather& G = h->gat;
  if (G.base) { h->err = "gather already initialised on this handle"; return LMPC_ERR_INVALID; }
  G.world = world; G.rank = rank; G.B = B; G.sets = sets;
  G.slab_bytes = (gather_slab_bytes(h, B) + 255) & ~(size_t)255;
  G.set_bytes = G.slab_bytes * (size_t)world;
  G.flags_off = G.set_bytes * (size_t)sets;
  G.total_bytes = G.flags_off + sizeof(unsigned long long) * LMPC_MAX_PEERS + 256;
  // cudaMalloc (not the stream-ordered allocator): legacy IPC handles exist for these allocations only
  if (cudaMalloc(&G.base, G.total_bytes) != cudaSuccess) { G.base = nullptr; h->err = "cudaMalloc (gather buffer) failed"; return LMPC_ERR_ALLOC; }
  CK(cudaMemset(G.base, 0, G.total_bytes));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t ih;
  if (world > 1) CK(cudaIpcGetMemHandle(&ih, G.base)); else memset(&ih, 0, sizeof ih);
  memcpy(ipc_handle_out, &ih, sizeof ih);
  if (slab_bytes) *slab_bytes = G.slab_bytes;
  G.connected = world == 1;
  return LMPC_OK;
}

extern "C" int lmpc_gather_connect(lmpc_handle* h, const void* ipc_handles_all) {
  if (!h || !h->gat.base || !ipc_handles_all) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  lmpc_handle::Gather& G = h->gat;
  for (int p = 0; p < G.world; p++) {
    if (p == G.rank || G.peer[p]) continue;
    cudaIpcMemHandle_t ih;
    memcpy(&ih, (const char*)ipc_handles_all + (size_t)p * LMPC_IPC_HANDLE_BYTES, sizeof ih);
    CK(cudaIpcOpenMemHandle(&G.peer[p], ih, cudaIpcMemLazyEnablePeerAccess));
  }
  G.connected = true;
  return LMPC_OK;
}

extern "C" int lmpc_gather_buffer(lmpc_handle* h, int set, double** device_ptr, size_t* doubles_per_rank) {
  if (!h || !h->gat.base || set < 0 || set >= h->gat.sets || !device_ptr) return LMPC_ERR_INVALID;
  *device_ptr = (double*)((char*)h->gat.base + h->gat.set_bytes * (size_t)set);
  if (doubles_per_rank) *doubles_per_rank = h->gat.slab_bytes / sizeof(double);
  return LMPC_OK;
}

extern "C" int lmpc_gather_wait(lmpc_handle* h, uint64_t seq) {
  if (!h || !h->gat.base) return LMPC_ERR_INVALID;
  lmpc_handle::Gather& G = h->gat;
  if (G.world == 1) return LMPC_OK;
  CK(cudaSetDevice(h->device));
  const unsigned long long* flags = (const unsigned long long*)((char*)G.base + G.flags_off);
  int* err = (int*)((char*)G.base + G.flags_off + sizeof(unsigned long long) * LMPC_MAX_PEERS + 8);
  lmpc_gather_wait_kernel<<<1, 32, 0, h->stream>>>(flags, G.world, G.rank, (unsigned long long)seq, 4000000000ll /* ~2 s */, err);
  h->launches++;
  CK(cudaGetLastError());
  return LMPC_OK;
}

extern "C" int lmpc_gather_error(lmpc_handle* h, int32_t* peer_timed_out) {
  if (!h || !h->gat.base || !peer_timed_out) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  int e = 0;
  CK(cudaMemcpy(&e, (char*)h->gat.base + h->gat.flags_off + sizeof(unsigned long long) * LMPC_MAX_PEERS + 8, sizeof e, cudaMemcpyDeviceToHost));
  *peer_timed_out = e;
  return LMPC_OK;
}

extern "C" int lmpc_solve_gather_batch(lmpc_handle* h, int B, const lmpc_batch_in* in, const lmpc_batch_out* out, int set,
                                       int wait, double* gathered_host, uint64_t* seq_out, int memspace) {
  if (!h || !in || !out || B < 1) return LMPC_ERR_INVALID;
  lmpc_handle::Gather& G = h->gat;
  if (!G.base || !G.connected || set < 0 || set >= G.sets || B != G.B) { if (h) h->err = "gather not initialised / connected, or batch size differs from lmpc_gather_init's"; return LMPC_ERR_INVALID; }
  CK(cudaSetDevice(h->device));
  // the trajectory outputs are routed into this rank's block of the chosen set
  const size_t Bz = (size_t)B, N = (size_t)h->P.N, NS = (size_t)h->P.NS;
  double* blk = (double*)((char*)G.base + G.set_bytes * (size_t)set + G.slab_bytes * (size_t)G.rank);
  lmpc_batch_out o2 = *out;
  lmpc_batch_out dev = *out;
  dev.X_optm = blk; dev.U_optm = blk + 6 * N * Bz; dev.dU_optm = dev.U_optm + 2 * NS * Bz; dev.cost = dev.dU_optm + 2 * NS * Bz;
  dev.status = (int32_t*)(dev.cost + Bz);
  // placeholders so that the argument check passes; the host path's own X/U/dU/cost/status copies are skipped below
  if (!o2.X_optm) o2.X_optm = dev.X_optm;
  if (!o2.U_optm) o2.U_optm = dev.U_optm;
  if (!o2.dU_optm) o2.dU_optm = dev.dU_optm;
  if (!o2.status) o2.status = dev.status;
  int rc = check_batch_args(h, B, in, &o2);
  if (rc != LMPC_OK) return rc;
  DevIO io;
  rc = stage_in(h, B, in, &o2, memspace, io);
  if (rc != LMPC_OK) return rc;
  io.dout[0] = dev.X_optm; io.dout[1] = dev.U_optm; io.dout[2] = dev.dU_optm; io.dout[6] = dev.cost; io.d_status = dev.status;
  G.seq++;
  G.active_set = set;
  int ss_count = 0;
  rc = run_tick_kernels(h, B, io, io.din[2], io.din[3], io.din[10] ? io.din[10] : io.din[3], true, nullptr, &ss_count);
  G.active_set = -1;
  if (rc != LMPC_OK) return rc;
  if (seq_out) *seq_out = G.seq;
  if (wait) { rc = lmpc_gather_wait(h, G.seq); if (rc != LMPC_OK) return rc; }
  if (memspace == LMPC_MEM_HOST) {
    // host callers: lambda, iterations and (behind the QP kernel, side stream) the safe-set columns as in lmpc_solve_batch;
    // the trajectories of ALL ranks as one copy of the gathered set when asked for, else this rank's block
    const bool learn = h->P.learning != 0;
    if (out->convex_combi_optm && io.dout[3] && learn) CK(cudaMemcpyAsync(out->convex_combi_optm, io.dout[3], sizeof(double) * io.nout[3], cudaMemcpyDeviceToHost, h->stream));
    if (out->iters) CK(cudaMemcpyAsync(out->iters, io.d_iters, sizeof(int32_t) * Bz, cudaMemcpyDeviceToHost, h->stream));
    if (!io.ss_copied && learn) {
      if (out->ss_x) CK(cudaMemcpyAsync(out->ss_x, io.ssx, sizeof(double) * io.nout[4], cudaMemcpyDeviceToHost, h->stream));
      if (out->ss_j) CK(cudaMemcpyAsync(out->ss_j, io.ssj, sizeof(double) * io.nout[5], cudaMemcpyDeviceToHost, h->stream));
    }
    if (gathered_host && wait) CK(cudaMemcpyAsync(gathered_host, (char*)G.base + G.set_bytes * (size_t)set, G.set_bytes, cudaMemcpyDeviceToHost, h->stream));
    else {
      if (out->X_optm && out->U_optm == out->X_optm + 6 * N * Bz && out->dU_optm == out->U_optm + 2 * NS * Bz) {
        // one arena (alloc_host_outputs): the three trajectory arrays are adjacent on both sides -> one copy
        CK(cudaMemcpyAsync(out->X_optm, dev.X_optm, sizeof(double) * (6 * N + 4 * NS) * Bz, cudaMemcpyDeviceToHost, h->stream));
      } else {
        if (out->X_optm) CK(cudaMemcpyAsync(out->X_optm, dev.X_optm, sizeof(double) * 6 * N * Bz, cudaMemcpyDeviceToHost, h->stream));
        if (out->U_optm) CK(cudaMemcpyAsync(out->U_optm, dev.U_optm, sizeof(double) * 2 * NS * Bz, cudaMemcpyDeviceToHost, h->stream));
        if (out->dU_optm) CK(cudaMemcpyAsync(out->dU_optm, dev.dU_optm, sizeof(double) * 2 * NS * Bz, cudaMemcpyDeviceToHost, h->stream));
      }
      if (out->cost) CK(cudaMemcpyAsync(out->cost, dev.cost, sizeof(double) * Bz, cudaMemcpyDeviceToHost, h->stream));
      if (out->status) CK(cudaMemcpyAsync(out->status, dev.status, sizeof(int32_t) * Bz, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    if (io.ss_copied) CK(cudaStreamSynchronize(h->side));
  }
  return LMPC_OK;
}

// ------------------------------------------------------------------------------------------ SQP to convergence
extern "C" int lmpc_solve_sqp_batch(lmpc_handle* h, int B, const lmpc_batch_in* in, const lmpc_batch_out* out, int max_sqp_iter,
                                    double sqp_tol, int32_t* sqp_iters, double* defect, int memspace) {
  int rc = check_batch_args(h, B, in, out);
  if (rc != LMPC_OK) return rc;
  if (max_sqp_iter < 1 || !(sqp_tol > 0.0)) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  DevIO io;
  rc = stage_in(h, B, in, out, memspace, io);
  if (rc != LMPC_OK) return rc;
  const size_t Bz = (size_t)B, N = (size_t)h->P.N, NS = (size_t)h->P.NS;
  // workspace: current linearisation point, per-instance flags, defect, active counter
  rc = dev_reserve(h, h->ws_sqp, sizeof(double) * (2 * (6 * N + 2 * NS) + 2) * Bz + sizeof(int32_t) * (2 * Bz + 2));
  if (rc != LMPC_OK) return rc;
  double* Xk = (double*)h->ws_sqp.p; double* Uk = Xk + 6 * N * Bz; double* Dprev = Uk + 2 * NS * Bz;
  double* alpha = Dprev + (6 * N + 2 * NS) * Bz; double* dfc = alpha + Bz;
  int32_t* done = (int32_t*)(dfc + Bz); int32_t* its = done + Bz; int32_t* n_active = its + Bz;
  const int threads = 128, blocks = (B + threads - 1) / threads;
  lmpc_sqp_init_kernel<<<blocks, threads, 0, h->stream>>>(B, (int)N, io.din[0], io.din[2], io.din[3], io.din[9], Xk, Uk, Dprev, alpha, done, its);
  h->launches++;
  int ss_count = 0;
  for (int k = 0; k < max_sqp_iter; k++) {
    // first pass: interior point started from U_warm when given; later passes from the previous solution
    const double* U0 = (k == 0 && io.din[10]) ? io.din[10] : Uk;
    rc = run_tick_kernels(h, B, io, Xk, Uk, U0, k == 0, done, &ss_count);
    if (rc != LMPC_OK) return rc;
    CK(cudaMemsetAsync(n_active, 0, sizeof(int32_t), h->stream));
    lmpc_sqp_update_kernel<<<blocks, threads, 0, h->stream>>>(B, (int)N, sqp_tol, io.dout[0], io.dout[1], io.d_status, Xk, Uk, Dprev, alpha, done, its, n_active);
    h->launches++;
    CK(cudaGetLastError());
    int32_t na = 0;
    CK(cudaMemcpyAsync(&na, n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (na == 0) break;
  }
  // instances whose step test never passed within max_sqp_iter: LMPC_SQP_MAX_ITER instead of the last QP's SOLVED
  lmpc_sqp_finalize_kernel<<<blocks, threads, 0, h->stream>>>(B, done, io.d_status);
  h->launches++;
  if (defect) {
    lmpc_sqp_defect_kernel<<<blocks, threads, 0, h->stream>>>(h->M, B, (int)N, io.dout[0], io.dout[1], io.din[4], io.din[7], dfc);
    h->launches++;
    CK(cudaGetLastError());
  }
  if (memspace == LMPC_MEM_HOST) {
    if (sqp_iters) CK(cudaMemcpyAsync(sqp_iters, its, sizeof(int32_t) * Bz, cudaMemcpyDeviceToHost, h->stream));
    if (defect) CK(cudaMemcpyAsync(defect, dfc, sizeof(double) * Bz, cudaMemcpyDeviceToHost, h->stream));
  } else {
    if (sqp_iters) CK(cudaMemcpyAsync(sqp_iters, its, sizeof(int32_t) * Bz, cudaMemcpyDeviceToDevice, h->stream));
    if (defect) CK(cudaMemcpyAsync(defect, dfc, sizeof(double) * Bz, cudaMemcpyDeviceToDevice, h->stream));
  }
  return stage_out(h, B, out, memspace, io);
}

// ------------------------------------------------------------------------------------------ track
extern "C" int lmpc_track_set(lmpc_handle* h, int n_rows, int n_cols, const double* table) {
  if (!h || !table) return LMPC_ERR_INVALID;
  LmpcTrackHost th;
  if (!lmpc_track_build(n_rows, n_cols, table, &th)) { h->err = "track table rejected (needs >= 8 rows, >= 13 columns, increasing abscissa, total length > 0)"; return LMPC_ERR_INVALID; }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  const size_t nb = th.brk.size(), nc = th.coef.size(), nw = th.way.size();
  int rc = dev_reserve(h, h->track_dev, sizeof(double) * (nb + nc + nw));
  if (rc != LMPC_OK) return rc;
  double* d = (double*)h->track_dev.p;
  CK(cudaMemcpyAsync(d, th.brk.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d + nb, th.coef.data(), sizeof(double) * nc, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d + nb + nc, th.way.data(), sizeof(double) * nw, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->track_host = std::move(th);
  h->track.m = h->track_host.m; h->track.n_way = h->track_host.n_way; h->track.L = h->track_host.L;
  h->track.brk = d; h->track.coef = d + nb; h->track.way = d + nb + nc;
  return LMPC_OK;
}

extern "C" int lmpc_track_load(lmpc_handle* h, const char* file) {
  if (!h || !file) return LMPC_ERR_INVALID;
  // casadi::DM::from_file(file_name).T() (racing_trajectory.cpp:188-191): whitespace-separated rows of equal length
  FILE* f = fopen(file, "r");
  if (!f) { h->err = std::string("cannot open ") + file; return LMPC_ERR_IO; }
  std::vector<double> v; int rows = 0, cols = -1;
  char* line = nullptr; size_t cap = 0;
  bool bad = false;
  while (getline(&line, &cap, f) > 0) {
    int c = 0; char* p = line; char* e = nullptr;
    for (;;) { const double x = strtod(p, &e); if (e == p) break; v.push_back(x); c++; p = e; while (*p == ',' ) p++; }
    if (c == 0) continue;
    if (cols < 0) cols = c; else if (c != cols) bad = true;
    rows++;
  }
  free(line); fclose(f);
  if (bad || rows == 0) { h->err = std::string("malformed track file ") + file; return LMPC_ERR_IO; }
  return lmpc_track_set(h, rows, cols, v.data());
}

extern "C" int lmpc_track_total_length(const lmpc_handle* h, double* L) {
  if (!h || !L || h->track.m == 0) return LMPC_ERR_INVALID;
  *L = h->track.L;
  return LMPC_OK;
}

// n items of `in_w` doubles in, `out_w` doubles out, through one of the three track kernels
static int track_batch(lmpc_handle* h, int which, int n, const double* in, int in_w, double* out, int out_w, int memspace) {
  if (!h || n < 1 || !in || !out) return LMPC_ERR_INVALID;
  if (h->track.m == 0) { h->err = "no track set (lmpc_track_set / lmpc_track_load)"; return LMPC_ERR_INVALID; }
  CK(cudaSetDevice(h->device));
  const double* din = in; double* dout = out;
  if (memspace == LMPC_MEM_HOST) {
    int rc = dev_reserve(h, h->st_in, sizeof(double) * (size_t)n * in_w);
    if (rc == LMPC_OK) rc = dev_reserve(h, h->st_out, sizeof(double) * (size_t)n * out_w);
    if (rc != LMPC_OK) return rc;
    CK(cudaMemcpyAsync(h->st_in.p, in, sizeof(double) * (size_t)n * in_w, cudaMemcpyHostToDevice, h->stream));
    din = (const double*)h->st_in.p; dout = (double*)h->st_out.p;
  }
  const int threads = 128, blocks = (n + threads - 1) / threads;
  if (which == 0) lmpc_track_eval_kernel<<<blocks, threads, 0, h->stream>>>(h->track, n, din, dout);
  else if (which == 1) lmpc_frenet_to_global_kernel<<<blocks, threads, 0, h->stream>>>(h->track, n, din, dout);
  else lmpc_global_to_frenet_kernel<<<blocks, threads, 0, h->stream>>>(h->track, n, din, dout);
  h->launches++;
  CK(cudaGetLastError());
  if (memspace == LMPC_MEM_HOST) {
    CK(cudaMemcpyAsync(out, dout, sizeof(double) * (size_t)n * out_w, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LMPC_OK;
}

extern "C" int lmpc_track_eval_batch(lmpc_handle* h, int n, const double* s, double* out, int memspace) { return track_batch(h, 0, n, s, 1, out, 7, memspace); }
extern "C" int lmpc_frenet_to_global_batch(lmpc_handle* h, int n, const double* frenet, double* global, int memspace) { return track_batch(h, 1, n, frenet, 3, global, 3, memspace); }
extern "C" int lmpc_global_to_frenet_batch(lmpc_handle* h, int n, const double* global, double* frenet, int memspace) { return track_batch(h, 2, n, global, 3, frenet, 3, memspace); }

// ------------------------------------------------------------------------------------------ per-agent safe sets
extern "C" int lmpc_agents_destroy(lmpc_handle* h) {
  if (!h) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->ag_buf.p) { cudaFree(h->ag_buf.p); h->ag_buf = DevBuf(); }
  h->ag_on = false; h->ag = LmpcAgentSets{};
  return LMPC_OK;
}

extern "C" int lmpc_agents_create(lmpc_handle* h, int B, int max_lap_samples) {
  if (!h || B < 1 || B > h->max_batch || max_lap_samples < 8) return LMPC_ERR_INVALID;
  if (!h->P.learning) { h->err = "per-agent safe sets need a learning configuration"; return LMPC_ERR_INVALID; }
  if (h->reg_in_tick) { h->err = "the in-tick error-dynamics regression reads the shared laps: disable it for per-agent safe sets"; return LMPC_ERR_INVALID; }
  lmpc_agents_destroy(h);
  CK(cudaSetDevice(h->device));
  LmpcAgentSets A{};
  A.B = B; A.slots = std::max(1, (int)h->cfg.max_lap_stored); A.cap = max_lap_samples;
  for (const HostLap& l : h->laps) if (l.n > A.cap) { h->err = "a stored lap is longer than max_lap_samples"; return LMPC_ERR_INVALID; }
  const size_t Bz = (size_t)B, pts = Bz * A.slots * 3 * (size_t)A.cap, rec = Bz * (size_t)A.cap;
  const size_t n_dbl = pts * 9 + rec * 10 + Bz * 10 + Bz, n_int = pts + Bz * A.slots + 5 * Bz;
  int rc = dev_reserve(h, h->ag_buf, sizeof(double) * n_dbl + sizeof(int) * n_int);
  if (rc == LMPC_OK) rc = dev_reserve(h, h->ag_cnt, sizeof(int) * (size_t)h->max_batch);
  if (rc != LMPC_OK) return rc;
  double* d = (double*)h->ag_buf.p;
  A.ps = d; d += pts; A.pe = d; d += pts; A.J = d; d += pts; A.xr = d; d += 6 * pts;
  A.rec_x = d; d += 6 * rec; A.rec_u = d; d += 2 * rec; A.rec_k = d; d += rec; A.rec_t = d; d += rec;
  A.carry = d; d += 10 * Bz; A.last_px = d; d += Bz;
  int* q = (int*)d;
  A.canon = q; q += pts; A.n = q; q += Bz * A.slots; A.head = q; q += Bz; A.count = q; q += Bz; A.rec_n = q; q += Bz; A.flags = q; q += Bz;
  A.lap_count = q; q += Bz;
  CK(cudaMemsetAsync(A.n, 0, sizeof(int) * (Bz * A.slots + 5 * Bz), h->stream));
  h->ag = A; h->ag_on = true;
  // every agent starts from the handle's laps (oldest first), as if each had called SafeSetRecorder::load on them; the
  // recorder's lap counter starts after them (safe_set.cpp:269)
  DevBuf tmp;
  for (const HostLap& l : h->laps) {
    rc = dev_reserve(h, tmp, sizeof(double) * 6 * (size_t)l.n);
    if (rc != LMPC_OK) return rc;
    CK(cudaMemcpyAsync(tmp.p, l.x.data(), sizeof(double) * 6 * (size_t)l.n, cudaMemcpyHostToDevice, h->stream));
    lmpc_agents_add_lap_kernel<<<(B * 32 + 127) / 128, 128, 0, h->stream>>>(h->ag, nullptr, (const double*)tmp.p, l.n, l.L);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
  }
  if (tmp.p) cudaFree(tmp.p);
  if (!h->laps.empty()) {
    std::vector<int> lc((size_t)B, (int)h->laps.size());
    CK(cudaMemcpyAsync(A.lap_count, lc.data(), sizeof(int) * Bz, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LMPC_OK;
}

// read-back for tests and for saving learned laps: the agent's `which`-th newest stored lap (0 = newest); returns the
// un-tripled samples x [n][6] (capacity max_n) and the lap's cost-to-go is implied (J_j = n - 1 - j)
extern "C" int lmpc_agents_get_lap(lmpc_handle* h, int agent, int which, int max_n, int32_t* n_out, double* x) {
  if (!h || !h->ag_on || agent < 0 || agent >= h->ag.B || which < 0 || !n_out) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  const LmpcAgentSets& A = h->ag;
  int cnt = 0, hd = 0;
  CK(cudaMemcpy(&cnt, A.count + agent, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&hd, A.head + agent, sizeof(int), cudaMemcpyDeviceToHost));
  if (which >= cnt) { *n_out = 0; return LMPC_OK; }
  const int slot = (hd + cnt - 1 - which) % A.slots;
  int n = 0;
  CK(cudaMemcpy(&n, A.n + (size_t)agent * A.slots + slot, sizeof(int), cudaMemcpyDeviceToHost));
  *n_out = n;
  if (x) {
    if (n > max_n) return LMPC_ERR_CAPACITY;
    // the middle copy of the tripled set is the lap itself (x_repeat = [x - L e0, x, x + L e0])
    CK(cudaMemcpy(x, A.xr + 6 * (lmpc_ag_slot(A, agent, slot) + (size_t)n), sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost));
  }
  return LMPC_OK;
}

// recorder state per agent: lap_count [B] (SafeSetRecorder::lap_count_), stored [B] laps in the safe set, flags [B]
// (bit 3: a lap overflowed max_lap_samples and was truncated).  Host buffers.
extern "C" int lmpc_agents_status(lmpc_handle* h, int32_t* lap_count, int32_t* stored, int32_t* flags) {
  if (!h || !h->ag_on) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  const size_t nb = sizeof(int) * (size_t)h->ag.B;
  if (lap_count) CK(cudaMemcpy(lap_count, h->ag.lap_count, nb, cudaMemcpyDeviceToHost));
  if (stored) CK(cudaMemcpy(stored, h->ag.count, nb, cudaMemcpyDeviceToHost));
  if (flags) CK(cudaMemcpy(flags, h->ag.flags, nb, cudaMemcpyDeviceToHost));
  return LMPC_OK;
}

// SafeSetManager::query for every agent on its own laps: query [B][2] (s, e_y) -> ss_x [B][max_total][6], ss_j [B][max_total]
// (raw J), count [B].  Host buffers.  Columns beyond count[b] are left untouched.
extern "C" int lmpc_agents_query_batch(lmpc_handle* h, const double* query, int max_total, int max_per_lap, double* ss_x, double* ss_j, int32_t* count) {
  if (!h || !h->ag_on || !query || !ss_x || !ss_j || !count || max_total < 1 || max_per_lap < 1) return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  const int B = h->ag.B;
  const size_t nq = 2 * (size_t)B, nx = 6 * (size_t)max_total * B, nj = (size_t)max_total * B;
  int rc = dev_reserve(h, h->st_in, sizeof(double) * nq);
  if (rc == LMPC_OK) rc = dev_reserve(h, h->st_out, sizeof(double) * (nx + nj));
  if (rc != LMPC_OK) return rc;
  CK(cudaMemcpyAsync(h->st_in.p, query, sizeof(double) * nq, cudaMemcpyHostToDevice, h->stream));
  double* dx = (double*)h->st_out.p; double* dj = dx + nx;
  const int max_used = std::min(h->ag.slots, (max_total + max_per_lap - 1) / max_per_lap);
  lmpc_agents_query_kernel<<<(B * max_used * 32 + 127) / 128, 128, 0, h->stream>>>(h->ag, max_used, max_per_lap, 0, nullptr, nullptr, nullptr,
                                                                                   (const double*)h->st_in.p, max_total, max_total, dx, dj, (int*)h->ag_cnt.p);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(count, h->ag_cnt.p, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int b = 0; b < B; b++) {
    if (count[b] < 1) continue;
    CK(cudaMemcpyAsync(ss_x + 6 * (size_t)max_total * b, dx + 6 * (size_t)max_total * b, sizeof(double) * 6 * (size_t)count[b], cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(ss_j + (size_t)max_total * b, dj + (size_t)max_total * b, sizeof(double) * (size_t)count[b], cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return LMPC_OK;
}

// ------------------------------------------------------------------------------------------ closed loop
static LmpcLoopParams loop_params(const lmpc_loop_options* o) {
  LmpcLoopParams P;
  P.step_mode = o->step_mode; P.delay_step = o->delay_step; P.plant_substeps = o->plant_substeps;
  P.dt = o->dt; P.plant_dt = o->plant_dt; P.speed_limit = o->speed_limit; P.speed_scale = o->speed_scale;
  P.max_vel_ref_diff = o->max_vel_ref_diff;
  return P;
}

static int closed_loop_impl(lmpc_handle* h, int B, int ticks, const lmpc_loop_options* opt, double* x, double* u_prev,
                            double* X_last, double* U_last, int32_t* lap_count, int32_t* fail_count, double* log_x,
                            double* log_u, double* log_rec, double t0, int memspace);

extern "C" int lmpc_closed_loop_run(lmpc_handle* h, int B, int ticks, const lmpc_loop_options* opt, double* x, double* u_prev,
                                    double* X_last, double* U_last, int32_t* lap_count, int32_t* fail_count, double* log_x,
                                    double* log_u, int memspace) {
  return closed_loop_impl(h, B, ticks, opt, x, u_prev, X_last, U_last, lap_count, fail_count, log_x, log_u, nullptr, 0.0, memspace);
}

extern "C" int lmpc_closed_loop_run_agents(lmpc_handle* h, int B, int ticks, const lmpc_loop_options* opt, double* x, double* u_prev,
                                           double* X_last, double* U_last, int32_t* lap_count, int32_t* fail_count, double* log_x,
                                           double* log_u, double* log_rec, double t0, int memspace) {
  if (!h || !h->ag_on || B != h->ag.B) { if (h) h->err = "lmpc_agents_create(B) first"; return LMPC_ERR_INVALID; }
  return closed_loop_impl(h, B, ticks, opt, x, u_prev, X_last, U_last, lap_count, fail_count, log_x, log_u, log_rec, t0, memspace);
}

static int closed_loop_impl(lmpc_handle* h, int B, int ticks, const lmpc_loop_options* opt, double* x, double* u_prev,
                            double* X_last, double* U_last, int32_t* lap_count, int32_t* fail_count, double* log_x,
                            double* log_u, double* log_rec, double t0, int memspace) {
  if (!h || !opt || B < 1 || ticks < 1 || !x || !u_prev || !X_last || !U_last) return LMPC_ERR_INVALID;
  if (B > h->max_batch) return LMPC_ERR_CAPACITY;
  if (h->track.m == 0) { h->err = "no track set (lmpc_track_set / lmpc_track_load)"; return LMPC_ERR_INVALID; }
  if (opt->step_mode < 0 || opt->step_mode > 1 || opt->delay_step < 0 || opt->delay_step >= h->P.NS || opt->plant_substeps < 1 ||
      !(opt->dt > 0.0) || !(opt->plant_dt > 0.0))
    return LMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  const size_t Bz = (size_t)B, N = (size_t)h->P.N, NS = (size_t)h->P.NS, K = (size_t)std::max(h->P.K, 1);
  const LmpcLoopParams O = loop_params(opt);
  // tick workspace: the solve's inputs and outputs, counters
  const size_t n_in = 6 * Bz + 2 * Bz + 6 * N * Bz + 2 * NS * Bz + NS * Bz + 4 * N * Bz + Bz;
  const size_t n_out = 6 * N * Bz + 4 * NS * Bz + K * Bz + Bz;
  int rc = dev_reserve(h, h->ws_loop, sizeof(double) * (n_in + n_out) + sizeof(int32_t) * (4 * Bz + 4));
  if (rc != LMPC_OK) return rc;
  double* w = (double*)h->ws_loop.p;
  LmpcTickIn I;
  I.x_ic = w; w += 6 * Bz; I.u_ic = w; w += 2 * Bz; I.X_ref = w; w += 6 * N * Bz; I.U_ref = w; w += 2 * NS * Bz;
  I.T_ref = w; w += NS * Bz; I.bl = w; w += N * Bz; I.br = w; w += N * Bz; I.kap = w; w += N * Bz; I.vref = w; w += N * Bz;
  I.L = w; w += Bz;
  double* oX = w; w += 6 * N * Bz; double* oU = w; w += 2 * NS * Bz; double* odU = w; w += 2 * NS * Bz;
  double* olam = w; w += K * Bz; double* ocost = w; w += Bz;
  int32_t* ostatus = (int32_t*)w; int32_t* oiters = ostatus + Bz; int32_t* d_lap = oiters + Bz; int32_t* d_fail = d_lap + Bz;
  int32_t* d_tick = d_fail + Bz;
  // agent state: caller's device buffers, or a staged copy of the caller's host buffers
  LmpcLoopState S;
  double *dlogx = log_x, *dlogu = log_u, *dlogr = log_rec;
  const bool agents = h->ag_on && B == h->ag.B;
  const size_t n_state = 6 * Bz + 2 * Bz + 6 * N * Bz + 2 * NS * Bz;
  if (memspace == LMPC_MEM_HOST) {
    const size_t n_log = (log_x ? 6 * Bz * (size_t)ticks : 0) + (log_u ? 2 * Bz * (size_t)ticks : 0) + (log_rec ? 10 * Bz * (size_t)ticks : 0);
    rc = dev_reserve(h, h->st_loop, sizeof(double) * (n_state + n_log));
    if (rc != LMPC_OK) return rc;
    double* q = (double*)h->st_loop.p;
    S.x = q; q += 6 * Bz; S.u_prev = q; q += 2 * Bz; S.X_last = q; q += 6 * N * Bz; S.U_last = q; q += 2 * NS * Bz;
    if (log_x) { dlogx = q; q += 6 * Bz * (size_t)ticks; }
    if (log_u) { dlogu = q; q += 2 * Bz * (size_t)ticks; }
    if (log_rec) { dlogr = q; q += 10 * Bz * (size_t)ticks; }
    CK(cudaMemcpyAsync(S.x, x, sizeof(double) * 6 * Bz, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(S.u_prev, u_prev, sizeof(double) * 2 * Bz, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(S.X_last, X_last, sizeof(double) * 6 * N * Bz, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(S.U_last, U_last, sizeof(double) * 2 * NS * Bz, cudaMemcpyHostToDevice, h->stream));
    if (lap_count) CK(cudaMemcpyAsync(d_lap, lap_count, sizeof(int32_t) * Bz, cudaMemcpyHostToDevice, h->stream));
    else CK(cudaMemsetAsync(d_lap, 0, sizeof(int32_t) * Bz, h->stream));
  } else {
    S.x = x; S.u_prev = u_prev; S.X_last = X_last; S.U_last = U_last;
    if (lap_count) CK(cudaMemcpyAsync(d_lap, lap_count, sizeof(int32_t) * Bz, cudaMemcpyDeviceToDevice, h->stream));
    else CK(cudaMemsetAsync(d_lap, 0, sizeof(int32_t) * Bz, h->stream));
  }
  S.lap_count = d_lap; S.fail_count = d_fail; S.tick = d_tick;
  CK(cudaMemsetAsync(d_fail, 0, sizeof(int32_t) * (Bz + 1), h->stream));   // fail counters and the tick counter

  DevIO io;
  const double* din[11] = {I.x_ic, I.u_ic, I.X_ref, I.U_ref, I.T_ref, I.bl, I.br, I.kap, I.vref, I.L, nullptr};
  for (int k = 0; k < 11; k++) io.din[k] = din[k];
  io.dout[0] = oX; io.dout[1] = oU; io.dout[2] = odU; io.dout[3] = olam; io.dout[4] = nullptr; io.dout[5] = nullptr; io.dout[6] = ocost;
  io.d_status = ostatus; io.d_iters = oiters; io.ssx = (double*)h->ws_ssx.p; io.ssj = (double*)h->ws_ssj.p;
  const int pthreads = 128, pblocks = (B * (int)N + pthreads - 1) / pthreads, ablocks = (B + pthreads - 1) / pthreads;
  for (int t = 0; t < ticks; t++) {
    lmpc_prepare_kernel<<<pblocks, pthreads, 0, h->stream>>>(h->M, h->track, O, B, (int)N, S, I);
    h->launches++;
    CK(cudaGetLastError());
    if (agents) {
      // RacingMPC::solve feeds its recorder before the query (racing_mpc.cpp:245-255): a lap completed by this tick's
      // sample is in the agent's safe set for this tick's solve
      lmpc_agents_record_kernel<<<ablocks, pthreads, 0, h->stream>>>(h->ag, I.x_ic, I.u_ic, I.kap, (int)N, I.L, t0, opt->dt, d_tick, dlogr, ticks);
      lmpc_agents_add_lap_kernel<<<(B * 32 + pthreads - 1) / pthreads, pthreads, 0, h->stream>>>(h->ag, I.L, nullptr, 0, 0.0);
      h->launches += 2;
      CK(cudaGetLastError());
    }
    int ss_count = 0;
    rc = run_tick_kernels(h, B, io, I.X_ref, I.U_ref, I.U_ref, true, nullptr, &ss_count);
    if (rc != LMPC_OK) return rc;
    lmpc_plant_kernel<<<ablocks, pthreads, 0, h->stream>>>(h->M, h->track, O, B, (int)N, S, I, oX, oU, ostatus, dlogx, dlogu, ticks);
    lmpc_tick_advance_kernel<<<1, 1, 0, h->stream>>>(d_tick);
    h->launches += 2;
    CK(cudaGetLastError());
  }
  if (memspace == LMPC_MEM_HOST) {
    CK(cudaMemcpyAsync(x, S.x, sizeof(double) * 6 * Bz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(u_prev, S.u_prev, sizeof(double) * 2 * Bz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(X_last, S.X_last, sizeof(double) * 6 * N * Bz, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(U_last, S.U_last, sizeof(double) * 2 * NS * Bz, cudaMemcpyDeviceToHost, h->stream));
    if (lap_count) CK(cudaMemcpyAsync(lap_count, d_lap, sizeof(int32_t) * Bz, cudaMemcpyDeviceToHost, h->stream));
    if (fail_count) CK(cudaMemcpyAsync(fail_count, d_fail, sizeof(int32_t) * Bz, cudaMemcpyDeviceToHost, h->stream));
    if (log_x) CK(cudaMemcpyAsync(log_x, dlogx, sizeof(double) * 6 * Bz * (size_t)ticks, cudaMemcpyDeviceToHost, h->stream));
    if (log_u) CK(cudaMemcpyAsync(log_u, dlogu, sizeof(double) * 2 * Bz * (size_t)ticks, cudaMemcpyDeviceToHost, h->stream));
    if (log_rec) CK(cudaMemcpyAsync(log_rec, dlogr, sizeof(double) * 10 * Bz * (size_t)ticks, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  } else {
    if (lap_count) CK(cudaMemcpyAsync(lap_count, d_lap, sizeof(int32_t) * Bz, cudaMemcpyDeviceToDevice, h->stream));
    if (fail_count) CK(cudaMemcpyAsync(fail_count, d_fail, sizeof(int32_t) * Bz, cudaMemcpyDeviceToDevice, h->stream));
  }
  return LMPC_OK;
}

// one preparation step on its own (parity tests; device buffers only): fills the solve's input keys from the agent state
extern "C" int lmpc_prepare_batch(lmpc_handle* h, int B, const lmpc_loop_options* opt, const double* x, const double* u_prev,
                                  const double* X_last, const double* U_last, double* x_ic, double* u_ic, double* X_ref,
                                  double* U_ref, double* T_ref, double* bound_left, double* bound_right, double* curvatures,
                                  double* vel_ref, double* total_length) {
  if (!h || !opt || B < 1 || !x || !u_prev || !X_last || !U_last || !x_ic || !u_ic || !X_ref || !U_ref || !T_ref || !bound_left ||
      !bound_right || !curvatures || !vel_ref || !total_length)
    return LMPC_ERR_INVALID;
  if (h->track.m == 0) { h->err = "no track set (lmpc_track_set / lmpc_track_load)"; return LMPC_ERR_INVALID; }
  CK(cudaSetDevice(h->device));
  LmpcLoopState S{};
  S.x = const_cast<double*>(x); S.u_prev = const_cast<double*>(u_prev); S.X_last = const_cast<double*>(X_last); S.U_last = const_cast<double*>(U_last);
  LmpcTickIn I{x_ic, u_ic, X_ref, U_ref, T_ref, bound_left, bound_right, curvatures, vel_ref, total_length};
  const int N = h->P.N, threads = 128, blocks = (B * N + threads - 1) / threads;
  lmpc_prepare_kernel<<<blocks, threads, 0, h->stream>>>(h->M, h->track, loop_params(opt), B, N, S, I);
  h->launches++;
  CK(cudaGetLastError());
  return LMPC_OK;
}
