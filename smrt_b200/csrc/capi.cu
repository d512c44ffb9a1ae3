// capi.cu — extern "C" entry points of libsmrt_dort_b200.so (declared in include/smrt_dort_b200.h).
//
// A plan owns: the Gauss-Legendre node table, two "slots" (CUDA stream + per-chunk workspace: layer eigen records,
// per-layer aux values, work counters, optional per-CTA global scratch) so that the eigen kernel of chunk c+1 overlaps
// the boundary kernel of chunk c, pinned host staging buffers for the *_host entry point, and CUDA events for timing.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges around pack / H2D / kernels / D2H

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "dort_host.h"

static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                                 \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) return fail(-2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

namespace {

constexpr int kSlots = 2;
constexpr size_t kMaxSmemOptin = 227 * 1024 - 256;  // opt-in limit minus the kernels' static shared memory

struct Slot {
  cudaStream_t stream = nullptr;
  double* aux = nullptr;
  double* eig = nullptr;
  double* kmin = nullptr;
  int* scat_flag = nullptr;
  int* counters = nullptr;
  double* scratch = nullptr;
  cudaEvent_t done = nullptr;
  bool used = false;
};

}  // namespace

struct smrtb200_plan {
  smrtb200_options opt;
  smrt_host::Layout layout;
  int device = 0;
  int sm_count = 0;
  int chunk = 0;
  bool use_global_scratch = false;  // the BOUNDARY kernel keeps its matrices in the per-CTA global scratch
  int eigen_variant = 0;            // 0: shared memory (h <= 64), 1: global scratch, 2: shared memory, 64 < h <= 128
  bool boundary_mid = false;        // boundary kernel for 64 < h <= 128: one resident matrix + L2-resident scratch
  int gj_single = 1;                // one blocked Gauss-Jordan instantiation for every block of a plan (h <= 64 kernels)
  int eigen_grid = 0, boundary_grid = 0;
  int boundary_threads = SMRT_NT_B;
  int eigen_threads = SMRT_NT;
  void (*eigen_fn)(KArgs) = nullptr;
  void (*boundary_fn)(KArgs) = nullptr;
  void (*boundary_fn_rough)(KArgs) = nullptr;  // the same instantiation with the rough-interface code (kRough)
  size_t eigen_smem = 0, boundary_smem = 0;
  long long scratch_stride = 0;
  double* gl_mu = nullptr;
  unsigned long long* prof = nullptr;  // SMRT_B200_PROFILE: cycle counters of the boundary kernel phases
  Slot slots[kSlots];
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  unsigned long long workspace_bytes = 0;
  unsigned long long launches = 0;
  float last_total_ms = 0.f, last_eigen_ms = 0.f, last_boundary_ms = 0.f;  // sums over the chunks of the last solve
  int last_chunks = 0;
  std::vector<cudaEvent_t> chunk_events;  // 3 per chunk: before eigen, after eigen, after boundary
  cudaEvent_t ev_dep = nullptr;  // orders the plan's two slot streams after the caller's stream
  // *_host entry point: a ring of kRing fixed-size staging slots (pinned host block + device block + three events per
  // slot); H2D of host chunk k+1 and D2H of host chunk k-1 run under the kernels of host chunk k
  struct RingSlot {
    std::vector<void*> dev, pinned;  // one buffer per field of field_table()
    smrtb200_batch dev_batch;
    cudaEvent_t h2d_done = nullptr, solve_done = nullptr, d2h_done = nullptr;
    int b0 = 0, nb = 0;  // host chunk in flight (nb == 0: none)
  };
  std::vector<RingSlot> ring;
  std::vector<void*> shared_dev;  // theta / theta_inc (shared by every problem)
  int host_chunk = 0;
  cudaStream_t host_stream = nullptr, h2d_stream = nullptr, d2h_stream = nullptr;
  bool staging_ready = false;
};

extern "C" int smrtb200_abi_version(void) { return SMRTB200_ABI_VERSION; }
extern "C" const char* smrtb200_last_error(void) { return g_last_error.c_str(); }

extern "C" int smrtb200_device_count(int* count) {
  if (!count) return fail(-1, "count is NULL");
  CUDA_TRY(cudaGetDeviceCount(count));
  return 0;
}

template <typename T>
static int dev_alloc(smrtb200_plan* p, T** ptr, size_t count) {
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CUDA_TRY(cudaMalloc((void**)ptr, bytes));
  p->workspace_bytes += bytes;
  return 0;
}

extern "C" int smrtb200_plan_destroy(smrtb200_plan* p) {
  if (!p) return 0;
  cudaSetDevice(p->device);
  for (auto& s : p->slots) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.aux);
    cudaFree(s.eig);
    cudaFree(s.kmin);
    cudaFree(s.scat_flag);
    cudaFree(s.counters);
    cudaFree(s.scratch);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  for (cudaEvent_t e : p->chunk_events) cudaEventDestroy(e);
  for (auto& r : p->ring) {
    for (void* d : r.dev) cudaFree(d);
    for (void* h : r.pinned) cudaFreeHost(h);
    if (r.h2d_done) cudaEventDestroy(r.h2d_done);
    if (r.solve_done) cudaEventDestroy(r.solve_done);
    if (r.d2h_done) cudaEventDestroy(r.d2h_done);
  }
  for (void* d : p->shared_dev) cudaFree(d);
  if (p->host_stream) cudaStreamDestroy(p->host_stream);
  if (p->h2d_stream) cudaStreamDestroy(p->h2d_stream);
  if (p->d2h_stream) cudaStreamDestroy(p->d2h_stream);
  if (p->ev_dep) cudaEventDestroy(p->ev_dep);
  if (p->prof) {
    unsigned long long c[16];
    if (cudaMemcpy(c, p->prof, sizeof(c), cudaMemcpyDeviceToHost) == cudaSuccess) {
      static const char* names[8] = {"between problems", "problem setup", "layer head", "rhs + formation", "elimination 1",
                                     "extraction + products", "elimination 2", "R + source"};
      double tot = 0;
      for (int i = 0; i < 8; ++i) tot += (double)c[i];
      for (int i = 0; i < 8; ++i)
        std::fprintf(stderr, "[smrtb200 profile] %-24s %6.2f %%  %.3e cycles\n", names[i], 100.0 * c[i] / (tot > 0 ? tot : 1), (double)c[i]);
    }
    cudaFree(p->prof);
  }
  cudaFree(p->gl_mu);
  if (p->ev_begin) cudaEventDestroy(p->ev_begin);
  if (p->ev_end) cudaEventDestroy(p->ev_end);
  delete p;
  return 0;
}

extern "C" int smrtb200_plan_create(const smrtb200_options* options, smrtb200_plan** out) {
  if (!options || !out) return fail(-1, "options / plan pointer is NULL");
  const char* err = smrt_host::validate_options(*options);
  if (err) return fail(-1, "invalid options: %s", err);
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (options->device < 0 || options->device >= ndev) return fail(-1, "device %d out of range (%d devices)", options->device, ndev);
  CUDA_TRY(cudaSetDevice(options->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, options->device));
  if (prop.major < 10) return fail(-3, "device %d is sm_%d%d: this library contains sm_100a code only", options->device, prop.major, prop.minor);

  smrtb200_plan* p = new (std::nothrow) smrtb200_plan();
  if (!p) return fail(-4, "out of host memory");
  p->opt = *options;
  p->device = options->device;
  p->sm_count = prop.multiProcessorCount;
  p->layout = smrt_host::make_layout(*options);
  // default on (measured: boundary kernel 8.50 -> 8.31 ms per cfg-2 launch); SMRT_B200_GJ_SINGLE=0 picks by block size
  p->gj_single = std::getenv("SMRT_B200_GJ_SINGLE") ? std::atoi(std::getenv("SMRT_B200_GJ_SINGLE")) : 1;
  const smrt_host::Layout& L = p->layout;

  // shared-memory path if both kernels fit into the opt-in limit, else matrices in an L2-resident global scratch
  p->use_global_scratch = L.boundary_smem_bytes > kMaxSmemOptin;
  if (L.hmax <= 64 && L.eigen_smem_bytes <= kMaxSmemOptin)
    p->eigen_variant = std::getenv("SMRT_B200_EIGEN4") ? 3 : 0;  // experiment: four CTAs per SM (packed C, 128 registers)
  else if (L.eigen_mid_smem_bytes > 0 && L.eigen_mid_smem_bytes <= kMaxSmemOptin && !std::getenv("SMRT_B200_NO_MID"))
    p->eigen_variant = 2;
  else
    p->eigen_variant = 1;
  p->eigen_smem = p->eigen_variant == 0   ? L.eigen_smem_bytes
                  : p->eigen_variant == 2 ? L.eigen_mid_smem_bytes
                  : p->eigen_variant == 3 ? L.eigen_small_packed_smem_bytes
                                          : L.eigen_vec_bytes;
  p->boundary_smem = p->use_global_scratch ? L.boundary_vec_bytes : L.boundary_smem_bytes;
  if (p->use_global_scratch && L.boundary_mid_smem_bytes > 0 && L.boundary_mid_smem_bytes <= kMaxSmemOptin &&
      !std::getenv("SMRT_B200_NO_MID")) {
    p->boundary_mid = true;
    p->use_global_scratch = false;
    p->boundary_smem = L.boundary_mid_smem_bytes;
  }
  int rc = 0;
  auto bail = [&](int code) {
    smrtb200_plan_destroy(p);
    return code;
  };
#define PLAN_TRY(expr)                 \
  do {                                 \
    rc = (expr);                       \
    if (rc != 0) return bail(rc);      \
  } while (0)
#define PLAN_CUDA(expr)                                                                               \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return bail(fail(-2, "%s failed: %s", #expr, cudaGetErrorString(_e)));     \
  } while (0)

  p->eigen_fn = p->eigen_variant == 0   ? eigen_kernel<0>
                : p->eigen_variant == 2 ? eigen_kernel<2>
                : p->eigen_variant == 3 ? eigen_kernel<3>
                                        : eigen_kernel<1>;
  // boundary kernel: 512 threads (128 registers) or 256 threads (255 registers: no spills in the register-tiled
  // eliminations); the block size is a plan parameter
  if (!p->use_global_scratch) p->boundary_threads = 256;  // measured: 14.9 ms per launch vs 17.9 ms with 512 (cfg 2)
  if (p->boundary_mid) p->boundary_threads = 512;         // the tile maps of the mid instantiation are written for 512
  if (const char* e = std::getenv("SMRT_B200_BOUNDARY_THREADS")) {
    int v = std::atoi(e);
    if (v >= 64 && v <= SMRT_NT_B && (v % 64) == 0) p->boundary_threads = v;
  }
  // two problems per SM: either the resident layout already allows it (small stream counts: 128 threads so that the
  // registers allow it too), or the instantiation that stages F and G into [T | R] does (h <= 64, <= 113 KB per CTA)
  const size_t two_per_sm = (size_t)(228 * 1024) / 2 - 1024 - 256;
  bool stream_fg = false;
  if (!p->use_global_scratch && !p->boundary_mid && !std::getenv("SMRT_B200_BOUNDARY_THREADS")) {
    if (L.boundary_smem_bytes <= two_per_sm) {
      p->boundary_threads = 128;
    } else if (L.boundary_stream_smem_bytes > 0 && L.boundary_stream_smem_bytes <= two_per_sm) {
      p->boundary_threads = 128;
      stream_fg = true;
    }
  }
  if (const char* e = std::getenv("SMRT_B200_STREAM_FG")) {  // experiments: force the choice
    const bool want = std::atoi(e) != 0;
    if (!want) {
      if (stream_fg) p->boundary_threads = 256;
      stream_fg = false;
    } else if (!p->use_global_scratch && !p->boundary_mid && L.boundary_stream_smem_bytes > 0) {
      stream_fg = true;
      p->boundary_threads = 128;
    }
  }
  if (stream_fg) p->boundary_smem = L.boundary_stream_smem_bytes;
  if (p->boundary_mid) {
    p->boundary_fn = boundary_kernel<false, 512, false, true>;
    p->boundary_fn_rough = boundary_kernel<false, 512, false, true, true>;
  } else if (p->use_global_scratch) {
    p->boundary_fn = boundary_kernel<true, SMRT_NT_B>;
    p->boundary_fn_rough = boundary_kernel<true, SMRT_NT_B, false, false, true>;
  } else if (stream_fg) {
    p->boundary_fn = boundary_kernel<false, 128, true>;
    p->boundary_fn_rough = boundary_kernel<false, 128, true, false, true>;
  } else {
    p->boundary_fn = (p->boundary_threads <= 128)   ? boundary_kernel<false, 128>
                     : (p->boundary_threads <= 256) ? boundary_kernel<false, 256>
                                                    : boundary_kernel<false, SMRT_NT_B>;
    p->boundary_fn_rough = (p->boundary_threads <= 128)   ? boundary_kernel<false, 128, false, false, true>
                           : (p->boundary_threads <= 256) ? boundary_kernel<false, 256, false, false, true>
                                                          : boundary_kernel<false, SMRT_NT_B, false, false, true>;
  }
  // the attribute belongs to the FUNCTION, not to the plan: always raise it to the opt-in maximum so that plans with
  // different shared-memory footprints can coexist
  {
    int optin = 0;
    PLAN_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, options->device));
    cudaFuncAttributes fa;
    PLAN_CUDA(cudaFuncGetAttributes(&fa, p->eigen_fn));
    PLAN_CUDA(cudaFuncSetAttribute(p->eigen_fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   optin - (int)fa.sharedSizeBytes));
    PLAN_CUDA(cudaFuncGetAttributes(&fa, p->boundary_fn));
    PLAN_CUDA(cudaFuncSetAttribute(p->boundary_fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   optin - (int)fa.sharedSizeBytes));
    PLAN_CUDA(cudaFuncGetAttributes(&fa, p->boundary_fn_rough));
    PLAN_CUDA(cudaFuncSetAttribute(p->boundary_fn_rough, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   optin - (int)fa.sharedSizeBytes));
  }
  int occ_e = 0, occ_b = 0;
  p->eigen_threads = (p->eigen_variant == 0 || p->eigen_variant == 3) ? SMRT_NT_SMEM
                     : p->eigen_variant == 2                          ? SMRT_NT_MID
                                                                      : SMRT_NT;
  PLAN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, p->eigen_fn, p->eigen_threads, p->eigen_smem));
  PLAN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, p->boundary_fn, p->boundary_threads, p->boundary_smem));
  if (occ_e < 1 || occ_b < 1) return bail(fail(-2, "kernels do not fit on an SM (occupancy %d / %d)", occ_e, occ_b));
  // keep the scratch L2-resident: at most 2 CTAs per SM
  if (p->eigen_variant == 1) occ_e = std::min(occ_e, 2);
  if (p->use_global_scratch) occ_b = std::min(occ_b, 2);
  p->eigen_grid = p->sm_count * occ_e;
  p->boundary_grid = p->sm_count * occ_b;
  // stride in doubles, a multiple of 16 (128 B): the kernels use 16-byte vector accesses on the per-CTA scratch
  const bool any_scratch = p->use_global_scratch || p->eigen_variant == 1 || p->boundary_mid;
  long long need = 0;
  if (p->eigen_variant == 1) need = std::max(need, L.eigen_scratch_doubles);
  if (p->use_global_scratch) need = std::max(need, L.boundary_scratch_doubles);
  if (p->boundary_mid) need = std::max(need, L.boundary_mid_scratch_doubles);
  p->scratch_stride = any_scratch ? ((need + 31) & ~15LL) : 0;

  // chunk: enough problems to fill the machine several times, bounded by a workspace budget of ~6 GB per slot
  {
    const double bytes_per_problem = (double)options->max_layers * (double)(L.eig_stride + SMRT_AUX_STRIDE + SMRT_MAX_MODES) * 8.0;
    long long by_mem = (long long)(6.0e9 / std::max(bytes_per_problem, 1.0));
    long long want = options->chunk > 0 ? options->chunk : std::max<long long>(4LL * p->boundary_grid, 1024);
    if (const char* e = std::getenv("SMRT_B200_CHUNK")) {
      long long v = std::atoll(e);
      if (v > 0) want = v;
    }
    long long c = std::min<long long>(std::min<long long>(want, std::max<long long>(by_mem, 1)), options->max_batch);
    // whole waves of the boundary kernel (one problem per CTA and wave): 456 problems on 148 CTAs would make 12 CTAs run
    // a fourth problem while the others idle (cfg 4: 63.6 vs 53.6 ms per launch)
    if (c > p->boundary_grid && c < options->max_batch) c -= c % p->boundary_grid;
    p->chunk = (int)std::max<long long>(c, 1);
  }

  std::vector<double> gl(L.n);
  smrt_host::gauss_legendre_positive_nodes(L.n, gl.data());
  PLAN_TRY(dev_alloc(p, &p->gl_mu, (size_t)L.n));
  PLAN_CUDA(cudaMemcpy(p->gl_mu, gl.data(), sizeof(double) * L.n, cudaMemcpyHostToDevice));

  if (std::getenv("SMRT_B200_PROFILE")) {
    PLAN_TRY(dev_alloc(p, &p->prof, (size_t)16));
    PLAN_CUDA(cudaMemset(p->prof, 0, 16 * sizeof(unsigned long long)));
  }
  const size_t CL = (size_t)p->chunk * options->max_layers;
  for (auto& s : p->slots) {
    PLAN_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    PLAN_TRY(dev_alloc(p, &s.aux, CL * SMRT_AUX_STRIDE));
    PLAN_TRY(dev_alloc(p, &s.eig, CL * (size_t)L.eig_stride));
    PLAN_TRY(dev_alloc(p, &s.kmin, CL * SMRT_MAX_MODES));
    PLAN_TRY(dev_alloc(p, &s.scat_flag, CL));
    PLAN_TRY(dev_alloc(p, &s.counters, 2));
    if (any_scratch)
      PLAN_TRY(dev_alloc(p, &s.scratch, (size_t)std::max(p->eigen_grid, p->boundary_grid) * (size_t)p->scratch_stride));
    PLAN_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  }
  PLAN_CUDA(cudaEventCreate(&p->ev_begin));
  PLAN_CUDA(cudaEventCreate(&p->ev_end));
  PLAN_CUDA(cudaEventCreateWithFlags(&p->ev_dep, cudaEventDisableTiming));
  *out = p;
  return 0;
}

extern "C" int smrtb200_plan_workspace_bytes(const smrtb200_plan* p, unsigned long long* bytes) {
  if (!p || !bytes) return fail(-1, "NULL argument");
  *bytes = p->workspace_bytes;
  return 0;
}

extern "C" int smrtb200_plan_launch_count(const smrtb200_plan* p, unsigned long long* launches) {
  if (!p || !launches) return fail(-1, "NULL argument");
  *launches = p->launches;
  return 0;
}

extern "C" int smrtb200_plan_last_timing(const smrtb200_plan* p, float* total_ms, float* eigen_ms, float* boundary_ms,
                                         int* n_chunks) {
  if (!p) return fail(-1, "NULL plan");
  if (n_chunks) *n_chunks = p->last_chunks;
  if (total_ms) *total_ms = p->last_total_ms;
  if (eigen_ms) *eigen_ms = p->last_eigen_ms;
  if (boundary_ms) *boundary_ms = p->last_boundary_ms;
  return 0;
}

static int check_batch(const smrtb200_plan* p, const smrtb200_batch* b) {
  if (!p || !b) return fail(-1, "NULL plan / batch");
  if (b->B < 0 || b->B > p->opt.max_batch) return fail(-1, "batch size %d exceeds plan max_batch %d", b->B, p->opt.max_batch);
  const void* req[] = {b->frequency, b->nlayer, b->thickness, b->temperature, b->frac_volume, b->eps_bg, b->eps_sc,
                       b->emmodel, b->ms_kind, b->ms_p0, b->ms_p1, b->interface_kind, b->dense_snow_correction,
                       b->substrate_kind, b->substrate_eps, b->substrate_temperature, b->values, b->ks, b->ka,
                       b->eps_eff, b->n_streams_out, b->stream_angles, b->optical_depth, b->status};
  for (const void* q : req)
    if (!q) return fail(-1, "a required batch pointer is NULL");
  if (p->opt.mode == SMRTB200_MODE_PASSIVE && !b->theta) return fail(-1, "theta is NULL");
  if (p->opt.mode == SMRTB200_MODE_ACTIVE && !b->theta_inc) return fail(-1, "theta_inc is NULL");
  return 0;
}

// Queue the kernels of one batch (device pointers).  The plan's two slot streams are ordered after `user` on entry and
// `user` after them on exit.  ev_base: index of the first per-chunk event triple to use (the host entry point queues
// several batches per call and keeps all their events); returns the number of chunks in *nchunks_out.
static int solve_device_impl(smrtb200_plan* p, const smrtb200_batch* batch, cudaStream_t user, int ev_base,
                             int* nchunks_out) {
  const smrt_host::Layout& L = p->layout;
  const int B = batch->B;
  CUDA_TRY(cudaMemsetAsync(batch->status, 0, sizeof(int) * (size_t)B, user));
  CUDA_TRY(cudaEventRecord(p->ev_dep, user));
  for (auto& s : p->slots) {
    CUDA_TRY(cudaStreamWaitEvent(s.stream, p->ev_dep, 0));
    s.used = false;
  }
  int nchunks = 0;
  const int nslots = (p->opt.reserved & 1) ? 1 : kSlots;  // bit 0 of `reserved`: serialise the chunks on one stream
  for (int b0 = 0; b0 < B; b0 += p->chunk, ++nchunks) {
    const int nb = std::min(p->chunk, B - b0);
    Slot& s = p->slots[nchunks % nslots];
    KArgs A = smrt_host::make_kargs(p->opt, L, *batch, b0, nb);
    A.gl_mu = p->gl_mu;
    A.prof = p->prof;
    A.aux = s.aux;
    A.eig = s.eig;
    A.kmin = s.kmin;
    A.scat_flag = s.scat_flag;
    A.counters = s.counters;
    A.scratch = s.scratch;
    A.scratch_stride = p->scratch_stride;
    A.use_global_scratch = p->use_global_scratch ? 1 : 0;
    A.gj_single = p->gj_single;
    CUDA_TRY(cudaMemsetAsync(s.counters, 0, 2 * sizeof(int), s.stream));
    const int nthreads_opt = 128;
    const int items = nb * p->opt.max_layers;
    while (p->chunk_events.size() < 3 * (size_t)(ev_base + nchunks + 1)) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreate(&e));
      p->chunk_events.push_back(e);
    }
    cudaEvent_t* ev = &p->chunk_events[3 * (size_t)(ev_base + nchunks)];
    optics_kernel<<<(items + nthreads_opt - 1) / nthreads_opt, nthreads_opt, 0, s.stream>>>(A);
    CUDA_TRY(cudaEventRecord(ev[0], s.stream));
    p->eigen_fn<<<std::min(p->eigen_grid, items), p->eigen_threads, p->eigen_smem, s.stream>>>(A);
    CUDA_TRY(cudaEventRecord(ev[1], s.stream));
    // batches with rough interfaces run the instantiation that holds their code; the others the plain one
    (A.interface_params ? p->boundary_fn_rough : p->boundary_fn)<<<std::min(p->boundary_grid, nb), p->boundary_threads,
                                                                   p->boundary_smem, s.stream>>>(A);
    CUDA_TRY(cudaEventRecord(ev[2], s.stream));
    CUDA_TRY(cudaGetLastError());
    p->launches += 3;
    s.used = true;
  }
  for (auto& s : p->slots) {
    if (!s.used) continue;
    CUDA_TRY(cudaEventRecord(s.done, s.stream));
    CUDA_TRY(cudaStreamWaitEvent(user, s.done, 0));
  }
  if (nchunks_out) *nchunks_out = nchunks;
  return 0;
}

extern "C" int smrtb200_solve_batch_device(smrtb200_plan* p, const smrtb200_batch* batch, void* cuda_stream) {
  int rc = check_batch(p, batch);
  if (rc) return rc;
  if (batch->B == 0) return 0;
  CUDA_TRY(cudaSetDevice(p->device));
  cudaStream_t user = (cudaStream_t)cuda_stream;
  nvtxRangePushA("smrtb200 solve_batch_device (queue optics / eigen / boundary kernels)");
  CUDA_TRY(cudaEventRecord(p->ev_begin, user));
  int nchunks = 0;
  rc = solve_device_impl(p, batch, user, 0, &nchunks);
  nvtxRangePop();
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(p->ev_end, user));
  p->last_chunks = nchunks;
  return 0;
}

// timing of the last solve; call after the user stream has been synchronised
static void collect_timing(smrtb200_plan* p) {
  float t = 0.f;
  if (cudaEventElapsedTime(&t, p->ev_begin, p->ev_end) == cudaSuccess) p->last_total_ms = t;
  float e = 0.f, b = 0.f;
  for (int c = 0; c < p->last_chunks; ++c) {
    cudaEvent_t* ev = &p->chunk_events[3 * (size_t)c];
    if (cudaEventElapsedTime(&t, ev[0], ev[1]) == cudaSuccess) e += t;
    if (cudaEventElapsedTime(&t, ev[1], ev[2]) == cudaSuccess) b += t;
  }
  p->last_eigen_ms = e;
  p->last_boundary_ms = b;
}

extern "C" int smrtb200_plan_sync_timing(smrtb200_plan* p, void* cuda_stream) {
  if (!p) return fail(-1, "NULL plan");
  CUDA_TRY(cudaSetDevice(p->device));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  collect_timing(p);
  return 0;
}

namespace {
struct FieldDesc {
  size_t offset;     // offset of the pointer inside smrtb200_batch
  size_t elem;       // bytes per problem (per-problem arrays) or total bytes (shared arrays, per_problem = false)
  bool per_problem;
  bool is_output;
};
}  // namespace

#define FOFF(f) offsetof(smrtb200_batch, f)

static std::vector<FieldDesc> field_table(const smrtb200_plan* p) {
  const size_t Ls = (size_t)p->opt.max_layers, d = sizeof(double), i = sizeof(int);
  const size_t nout = (p->opt.mode == SMRTB200_MODE_PASSIVE) ? 2 * (size_t)p->opt.n_theta : 9 * (size_t)p->opt.n_inc;
  return {
      {FOFF(frequency), d, true, false},        {FOFF(nlayer), i, true, false},
      {FOFF(thickness), Ls * d, true, false},   {FOFF(temperature), Ls * d, true, false},
      {FOFF(frac_volume), Ls * d, true, false}, {FOFF(eps_bg), 2 * Ls * d, true, false},
      {FOFF(eps_sc), 2 * Ls * d, true, false},  {FOFF(emmodel), Ls * i, true, false},
      {FOFF(ms_kind), Ls * i, true, false},     {FOFF(ms_p0), Ls * d, true, false},
      {FOFF(ms_p1), Ls * d, true, false},       {FOFF(interface_kind), Ls * i, true, false},
      {FOFF(dense_snow_correction), Ls * i, true, false},
      {FOFF(substrate_kind), i, true, false},   {FOFF(substrate_eps), 2 * d, true, false},
      {FOFF(substrate_temperature), d, true, false},
      {FOFF(substrate_params), 4 * d, true, false}, {FOFF(atmosphere), 3 * d, true, false},
      {FOFF(inclusion), 5 * Ls * d, true, false},
      {FOFF(interface_params), 4 * Ls * d, true, false},
      {FOFF(theta), std::max<size_t>(p->opt.n_theta, 1) * d, false, false},
      {FOFF(theta_inc), std::max<size_t>(p->opt.n_inc, 1) * d, false, false},
      {FOFF(values), nout * d, true, true},     {FOFF(ks), Ls * d, true, true},
      {FOFF(ka), Ls * d, true, true},           {FOFF(eps_eff), 2 * Ls * d, true, true},
      {FOFF(n_streams_out), i, true, true},     {FOFF(stream_angles), (size_t)p->opt.n_max_stream * d, true, true},
      {FOFF(optical_depth), d, true, true},     {FOFF(status), i, true, true},
  };
}

static void*& field_ptr(smrtb200_batch* b, size_t off) { return *reinterpret_cast<void**>(reinterpret_cast<char*>(b) + off); }
static const void* field_cptr(const smrtb200_batch* b, size_t off) {
  return *reinterpret_cast<void* const*>(reinterpret_cast<const char*>(b) + off);
}

constexpr int kRing = 3;

static int ensure_staging(smrtb200_plan* p) {
  if (p->staging_ready) return 0;
  auto tab = field_table(p);
  // host chunk: several kernel waves, so that the two-slot kernel pipeline of a chunk reaches its steady state, yet
  // small enough for the pinned ring to stay at tens of MB whatever the batch size
  long long hc = 8LL * p->chunk;
  if (const char* e = std::getenv("SMRT_B200_HOST_CHUNK")) {
    long long v = std::atoll(e);
    if (v > 0) hc = v;
  }
  p->host_chunk = (int)std::max<long long>(1, std::min<long long>(hc, p->opt.max_batch));
  CUDA_TRY(cudaStreamCreateWithFlags(&p->host_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&p->h2d_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking));
  p->ring.resize(kRing);
  p->shared_dev.assign(tab.size(), nullptr);
  for (size_t k = 0; k < tab.size(); ++k) {
    if (tab[k].per_problem) continue;
    CUDA_TRY(cudaMalloc(&p->shared_dev[k], std::max<size_t>(tab[k].elem, 16)));
    p->workspace_bytes += tab[k].elem;
  }
  for (auto& r : p->ring) {
    std::memset(&r.dev_batch, 0, sizeof(r.dev_batch));
    CUDA_TRY(cudaEventCreateWithFlags(&r.h2d_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&r.solve_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&r.d2h_done, cudaEventDisableTiming));
    r.dev.assign(tab.size(), nullptr);
    r.pinned.assign(tab.size(), nullptr);
    for (size_t k = 0; k < tab.size(); ++k) {
      const auto& f = tab[k];
      if (!f.per_problem) {
        field_ptr(&r.dev_batch, f.offset) = p->shared_dev[k];
        continue;
      }
      const size_t bytes = std::max<size_t>(f.elem * (size_t)p->host_chunk, 16);
      CUDA_TRY(cudaMalloc(&r.dev[k], bytes));
      CUDA_TRY(cudaMallocHost(&r.pinned[k], bytes));
      p->workspace_bytes += bytes;
      field_ptr(&r.dev_batch, f.offset) = r.dev[k];
    }
  }
  p->staging_ready = true;
  return 0;
}

// results of the host chunk in flight in ring slot `r` -> the caller's arrays
static int ring_drain(smrtb200_plan* p, smrtb200_plan::RingSlot& r, const smrtb200_batch* batch,
                      const std::vector<FieldDesc>& tab) {
  if (r.nb == 0) return 0;
  CUDA_TRY(cudaEventSynchronize(r.d2h_done));
  nvtxRangePushA("smrtb200 unpack results (pinned -> caller)");
  for (size_t k = 0; k < tab.size(); ++k) {
    const auto& f = tab[k];
    if (!f.is_output) continue;
    char* dst = static_cast<char*>(const_cast<void*>(field_cptr(batch, f.offset)));
    std::memcpy(dst + f.elem * (size_t)r.b0, r.pinned[k], f.elem * (size_t)r.nb);
  }
  nvtxRangePop();
  r.nb = 0;
  return 0;
}

extern "C" int smrtb200_solve_batch_host(smrtb200_plan* p, const smrtb200_batch* batch) {
  int rc = check_batch(p, batch);
  if (rc) return rc;
  if (batch->B == 0) return 0;
  CUDA_TRY(cudaSetDevice(p->device));
  rc = ensure_staging(p);
  if (rc) return rc;
  auto tab = field_table(p);
  const int B = batch->B;
  // arrays shared by every problem (viewing / incidence angles): pageable copy, ordered before the first kernels
  for (size_t k = 0; k < tab.size(); ++k) {
    const auto& f = tab[k];
    if (f.per_problem) continue;
    const void* src = field_cptr(batch, f.offset);
    if (src) CUDA_TRY(cudaMemcpyAsync(p->shared_dev[k], src, f.elem, cudaMemcpyHostToDevice, p->host_stream));
  }
  CUDA_TRY(cudaEventRecord(p->ev_begin, p->host_stream));
  int ev_base = 0, chunk_no = 0;
  for (int b0 = 0; b0 < B; b0 += p->host_chunk, ++chunk_no) {
    const int nb = std::min(p->host_chunk, B - b0);
    auto& r = p->ring[chunk_no % kRing];
    rc = ring_drain(p, r, batch, tab);  // the slot's previous chunk: wait for its D2H, hand the results over
    if (rc) return rc;
    nvtxRangePushA("smrtb200 pack chunk (caller -> pinned) + H2D");
    for (size_t k = 0; k < tab.size(); ++k) {
      const auto& f = tab[k];
      if (f.is_output || !f.per_problem) continue;
      const char* src = static_cast<const char*>(field_cptr(batch, f.offset));
      if (!src) continue;  // optional input not given
      const size_t bytes = f.elem * (size_t)nb;
      std::memcpy(r.pinned[k], src + f.elem * (size_t)b0, bytes);
      CUDA_TRY(cudaMemcpyAsync(r.dev[k], r.pinned[k], bytes, cudaMemcpyHostToDevice, p->h2d_stream));
    }
    CUDA_TRY(cudaEventRecord(r.h2d_done, p->h2d_stream));
    nvtxRangePop();
    smrtb200_batch db = r.dev_batch;  // optional inputs the caller did not give stay NULL on the device side
    db.B = nb;
    db.phi = batch->phi;
    if (!batch->substrate_params) db.substrate_params = nullptr;
    if (!batch->atmosphere) db.atmosphere = nullptr;
    if (!batch->inclusion) db.inclusion = nullptr;
    if (!batch->interface_params) db.interface_params = nullptr;
    CUDA_TRY(cudaStreamWaitEvent(p->host_stream, r.h2d_done, 0));
    int nchunks = 0;
    nvtxRangePushA("smrtb200 queue kernels of a host chunk");
    rc = solve_device_impl(p, &db, p->host_stream, ev_base, &nchunks);
    nvtxRangePop();
    if (rc) return rc;
    ev_base += nchunks;
    CUDA_TRY(cudaEventRecord(r.solve_done, p->host_stream));
    CUDA_TRY(cudaStreamWaitEvent(p->d2h_stream, r.solve_done, 0));
    for (size_t k = 0; k < tab.size(); ++k) {
      const auto& f = tab[k];
      if (!f.is_output) continue;
      CUDA_TRY(cudaMemcpyAsync(r.pinned[k], r.dev[k], f.elem * (size_t)nb, cudaMemcpyDeviceToHost, p->d2h_stream));
    }
    CUDA_TRY(cudaEventRecord(r.d2h_done, p->d2h_stream));
    r.b0 = b0;
    r.nb = nb;
  }
  CUDA_TRY(cudaEventRecord(p->ev_end, p->host_stream));
  for (int k = 0; k < kRing; ++k) {  // oldest first
    rc = ring_drain(p, p->ring[(chunk_no + k) % kRing], batch, tab);
    if (rc) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(p->host_stream));
  p->last_chunks = ev_base;
  collect_timing(p);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// FP64 roofline denominator: dependent-chain-free DFMA loop, 8 independent accumulators per thread
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3;
  double x4 = x0 + 4e-3, x5 = x0 + 5e-3, x6 = x0 + 6e-3, x7 = x0 + 7e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;  // never true; keeps the loop alive
}

extern "C" int smrtb200_measure_fp64_peak(int device, float ms, double* tflops) {
  if (!tflops) return fail(-1, "tflops is NULL");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  double* out = nullptr;
  CUDA_TRY(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int grid = prop.multiProcessorCount * 8, block = 256;
  int iters = 2000;
  double best = 0.0;
  float elapsed_total = 0.f;
  for (int rep = 0; rep < 50 && elapsed_total < ms; ++rep) {
    CUDA_TRY(cudaEventRecord(e0));
    fp64_peak_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-7);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, e0, e1));
    elapsed_total += t;
    double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)grid * (double)block;
    if (rep > 0) best = std::max(best, flops / (t * 1e-3) / 1e12);
    if (t < 5.f) iters *= 2;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return 0;
}
