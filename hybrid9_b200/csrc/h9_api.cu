/*
 * h9_api.cu -- the C-ABI of libh9gpu (include/h9gpu.h): context, land
 * compaction, state upload/download, the forcing ingest pipeline and the
 * launches of the time-stepping kernels.  No CPU implementation of the physics
 * exists in this library: every compute entry launches CUDA kernels and fails
 * with H9_ERR_CUDA when no device is usable.
 *
 * Replaces the loop nest HYBRID9.f90:120-295 of the reference's driver; see
 * include/h9gpu.h for the per-entry citations.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> /* types and prototypes only: libnccl is dlopen'ed by h9_comm_init */

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/h9gpu.h"
#include "h9_device.h"

using namespace h9;

namespace h9 {
int launch_budget(void* stream, const DevArrays& a, int iyr, double* d_budget);
}

struct h9_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr, comm_stream = nullptr;
  std::string err;
  bool configured = false, have_soil = false;
  int lon_c = 0, lat_c = 0, nisurf = 48, nyr = 1;
  size_t ngrid = 0;
  int math_mode = H9_MATH_FAST;
  int tile_days = 8, block = 1064; /* 64 threads/block, 128-register variant */
  int pair_max_cells = -1;         /* shards up to this size step with two lanes per cell; -1: auto */
  /* h9_run_days ingest (env H9_INGEST): 3 hostread (default; page-locked arrays are read in
   * place by the stepping kernel, pageable ones fall back to gather), 0 gather, 1 pack,
   * 2 zerocopy pack */
  int ingest = 3;
  bool zero_copy = false;
  int real_evap = 0; /* H9_OPT_REAL_EVAP */
  Geo geo;
  std::vector<int32_t> cell_xy;
  DevArrays dev;
  int32_t* d_cell_xy = nullptr;
  std::vector<void*> allocs; /* everything sized by the soil/land set */
  /* year bookkeeping (all cells step in lockstep, so these are scalars) */
  int cur_year = 0, nt = 0, max_year_seen = 0;
  unsigned long long step_counter = 0;
  /* forcing pipeline */
  float* d_stage[2] = {nullptr, nullptr};
  float* d_packed[2] = {nullptr, nullptr};
  float* h_stage[2] = {nullptr, nullptr};
  size_t stage_days = 0;
  cudaEvent_t ev_packed[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr},
              ev_h2d[2] = {nullptr, nullptr};
  int32_t* d_year_index = nullptr;
  size_t year_index_cap = 0;
  float* d_pack_full = nullptr;
  size_t pack_full_cap = 0;
  float* d_step_forcing = nullptr; /* [7][ncs] for the fine-grained entries */
  float* d_diag = nullptr;         /* [12][ncs] + jwt */
  double* d_budget = nullptr;
  /* multi-GPU (h9_comm_init): one ctx per GPU, NCCL on the ctx's own stream */
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  std::vector<int64_t> land_of_rank, stride_of_rank; /* nc and ncs of every rank */
  float* d_gather = nullptr;      /* [sum_r 13*ncs_r]: every rank's annual planes of one year */
  double* d_budget_all = nullptr; /* [nyr][8]: all-reduced budget of every year slot */
  size_t gather_elems = 0;
  int gathered_year = 0;
  /* the NCCL part of h9_annual_collective runs on comm_stream behind ev_coll_ready (budget
   * kernel done on the stepping stream) so that the next year's stepping kernel does not queue
   * behind it; ev_coll_done guards the year slot that collective still reads */
  cudaEvent_t ev_coll_ready = nullptr, ev_coll_done = nullptr;
  bool coll_inflight = false, coll_guard = false;
  /* counters */
  int64_t launches = 0, h2d = 0, d2h = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing;
  std::vector<cudaEvent_t> event_pool;
  double step_ms = 0.0;
};

namespace {

int fail(h9_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      return fail(ctx, H9_ERR_CUDA,                                                   \
                  std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ \
                                                                         ")");        \
    }                                                                                 \
  } while (0)

#define CKL(expr)                                                                         \
  do {                                                                                    \
    int e_ = (expr);                                                                      \
    if (e_ != 0)                                                                          \
      return fail(ctx, H9_ERR_CUDA,                                                       \
                  std::string(#expr) + ": " + cudaGetErrorString((cudaError_t)e_));      \
    ctx->launches += 1;                                                                   \
  } while (0)

/* `written_first`: the array is completely written (upload or kernel) before anything reads
 * it.  With env H9_POISON=1 such arrays start as 0xFF bytes (NaN floats) instead of zeros: a
 * run that still gives the bits of the normal run reads no uninitialised device memory
 * (tests/test_gpu_poison.py; the stand-in for compute-sanitizer's initcheck on this pool). */
template <class T>
int dev_alloc(h9_ctx* ctx, T** p, size_t n, bool track = true, bool written_first = false) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) return fail(ctx, H9_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  /* on the ctx's own stream: it is non-blocking, so a memset on the legacy stream would not be
   * ordered before the uploads and kernels that follow */
  const char* poison = getenv("H9_POISON");
  const int fill = (written_first && poison && atoi(poison) != 0) ? 0xFF : 0;
  e = cudaMemsetAsync(q, fill, std::max<size_t>(n, 1) * sizeof(T), ctx->stream);
  if (e != cudaSuccess) return fail(ctx, H9_ERR_CUDA, std::string("cudaMemsetAsync: ") + cudaGetErrorString(e));
  *p = (T*)q;
  if (track) ctx->allocs.push_back(q);
  return 0;
}

void free_land(h9_ctx* ctx) {
  for (void* q : ctx->allocs) cudaFree(q);
  ctx->allocs.clear();
  for (int b = 0; b < 2; ++b) {
    if (ctx->d_stage[b]) cudaFree(ctx->d_stage[b]);
    if (ctx->d_packed[b]) cudaFree(ctx->d_packed[b]);
    if (ctx->h_stage[b]) cudaFreeHost(ctx->h_stage[b]);
    ctx->d_stage[b] = ctx->d_packed[b] = ctx->h_stage[b] = nullptr;
  }
  ctx->stage_days = 0;
  if (ctx->d_pack_full) cudaFree(ctx->d_pack_full);
  ctx->d_pack_full = nullptr;
  ctx->pack_full_cap = 0;
  ctx->d_cell_xy = nullptr;
  ctx->d_step_forcing = nullptr;
  ctx->d_diag = nullptr;
  ctx->d_budget = nullptr;
  memset(&ctx->dev, 0, sizeof(ctx->dev));
  ctx->dev.real_evap = ctx->real_evap;
  ctx->have_soil = false;
}

/* fast mode: two lanes per cell (h9_kernels_pair.cu) when forced (block 4000) or when the
 * shard leaves at most one 16-cell warp per scheduler */
bool use_pair(const h9_ctx* ctx) {
  if (ctx->math_mode != H9_MATH_FAST) return false;
  if (ctx->block >= 4000) return true;
  if (ctx->block != 1064) return false;
  int limit = ctx->pair_max_cells;
  if (limit < 0) {
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    limit = nsm * 4 * 16;
  }
  return ctx->dev.nc <= limit;
}

int set_device(h9_ctx* ctx) {
  CK(cudaSetDevice(ctx->device));
  return 0;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

cudaEvent_t get_event(h9_ctx* ctx) {
  if (!ctx->event_pool.empty()) {
    cudaEvent_t e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) {
    cudaGetLastError();
    return nullptr; /* the launch goes ahead untimed */
  }
  return e;
}

/* fold finished timing pairs into step_ms and recycle their events; called where the stream
 * has just been synchronised, so a host loop that never asks for the time
 * keeps a bounded number of live events */
void fold_timing(h9_ctx* ctx) {
  for (auto& pr : ctx->timing) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) ctx->step_ms += ms;
    ctx->event_pool.push_back(pr.first);
    ctx->event_pool.push_back(pr.second);
  }
  ctx->timing.clear();
}

ForcingView strided_view(const float* base, size_t day_stride, size_t plane_stride, size_t limit) {
  ForcingView fv;
  for (int p = 0; p < kForcingPlanes; ++p) fv.plane[p] = base + p * plane_stride;
  fv.day_stride = day_stride;
  fv.limit = limit;
  return fv;
}

int launch_days(h9_ctx* ctx, int ndays, const int32_t* d_yi, const ForcingView& fv,
                const int32_t* d_cell_index = nullptr) {
  if (ctx->coll_inflight && ctx->coll_guard) {
    /* this batch writes the annual planes of the year an unfinished collective still reads */
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_coll_done, 0));
    ctx->coll_inflight = false;
  }
  cudaEvent_t e0 = get_event(ctx), e1 = get_event(ctx);
  const bool timed = e0 && e1;
  if (timed) CK(cudaEventRecord(e0, ctx->stream));
  if (ctx->math_mode == H9_MATH_EXACT) {
    CKL(launch_days_exact(ctx->stream, ctx->dev, ctx->geo, ndays, d_yi, fv, ctx->cur_year, ctx->nt,
                          ctx->step_counter, ctx->block, d_cell_index));
  } else if (use_pair(ctx)) {
    CKL(launch_days_pair(ctx->stream, ctx->dev, ctx->geo, ndays, d_yi, fv, ctx->cur_year, ctx->nt,
                         ctx->step_counter, d_cell_index));
  } else {
    CKL(launch_days_fast(ctx->stream, ctx->dev, ctx->geo, ndays, d_yi, fv, ctx->cur_year, ctx->nt,
                         ctx->step_counter, ctx->block, d_cell_index));
  }
  if (timed) {
    CK(cudaEventRecord(e1, ctx->stream));
    ctx->timing.emplace_back(e0, e1);
  } else {
    if (e0) ctx->event_pool.push_back(e0);
    if (e1) ctx->event_pool.push_back(e1);
  }
  return 0;
}

/* host mirror of the kernel's year bookkeeping */
void advance_years(h9_ctx* ctx, int ndays, const int32_t* yi) {
  for (int d = 0; d < ndays; ++d) {
    if (yi[d] != ctx->cur_year) {
      ctx->cur_year = yi[d];
      ctx->nt = 0;
    }
    ctx->nt += 1;
    ctx->max_year_seen = std::max(ctx->max_year_seen, (int)yi[d]);
  }
  ctx->step_counter += (unsigned long long)ndays * ctx->nisurf;
}

int upload_year_index(h9_ctx* ctx, int ndays, const int32_t* yi) {
  ctx->coll_guard = false;
  if (ctx->coll_inflight)
    for (int d = 0; d < ndays; ++d)
      if (yi[d] == ctx->gathered_year) ctx->coll_guard = true;
  if ((size_t)ndays > ctx->year_index_cap) {
    if (ctx->d_year_index) cudaFree(ctx->d_year_index);
    ctx->d_year_index = nullptr;
    CK(cudaMalloc((void**)&ctx->d_year_index, (size_t)ndays * sizeof(int32_t)));
    ctx->year_index_cap = (size_t)ndays;
  }
  CK(cudaMemcpyAsync(ctx->d_year_index, yi, (size_t)ndays * sizeof(int32_t),
                     cudaMemcpyHostToDevice, ctx->stream));
  ctx->h2d += (int64_t)ndays * 4;
  return 0;
}

int read_any_fault(h9_ctx* ctx, uint32_t* any) {
  CK(cudaMemcpyAsync(any, ctx->dev.any_fault, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->d2h += 4;
  if (ctx->timing.size() > 32) fold_timing(ctx);
  return 0;
}

int ensure_stage(h9_ctx* ctx, bool need_host_stage) {
  const size_t td = (size_t)ctx->tile_days;
  if (ctx->stage_days != td) {
    for (int b = 0; b < 2; ++b) {
      if (ctx->d_stage[b]) cudaFree(ctx->d_stage[b]);
      if (ctx->d_packed[b]) cudaFree(ctx->d_packed[b]);
      if (ctx->h_stage[b]) cudaFreeHost(ctx->h_stage[b]);
      ctx->d_stage[b] = ctx->d_packed[b] = ctx->h_stage[b] = nullptr;
      CK(cudaMalloc((void**)&ctx->d_stage[b], td * kForcingPlanes * ctx->ngrid * sizeof(float)));
      CK(cudaMalloc((void**)&ctx->d_packed[b], td * kForcingPlanes * ctx->dev.ncs * sizeof(float)));
    }
    ctx->stage_days = td;
  }
  if (need_host_stage && !ctx->h_stage[0]) {
    for (int b = 0; b < 2; ++b)
      CK(cudaMallocHost((void**)&ctx->h_stage[b], td * kForcingPlanes * ctx->ngrid * sizeof(float)));
  }
  for (int b = 0; b < 2; ++b) {
    if (!ctx->ev_packed[b]) CK(cudaEventCreateWithFlags(&ctx->ev_packed[b], cudaEventDisableTiming));
    if (!ctx->ev_done[b]) CK(cudaEventCreateWithFlags(&ctx->ev_done[b], cudaEventDisableTiming));
    if (!ctx->ev_h2d[b]) CK(cudaEventCreateWithFlags(&ctx->ev_h2d[b], cudaEventDisableTiming));
  }
  return 0;
}

/* days [d0,d0+nd) of the seven (lon_c,lat_c,ndays) host arrays -> `d_out` ([nd][7][ncs]) on
 * the copy stream.  Page-locked arrays are gathered in place by the pack kernel (zero-copy:
 * only land cells cross PCIe); pageable ones are bounced through the pinned ring and copied
 * to stage buffer b first. */
int stage_and_pack(h9_ctx* ctx, int b, int d0, int nd, const float* const src[7], bool pinned,
                   float* d_out) {
  const size_t plane = (size_t)nd * ctx->ngrid;
  const float* planes[kForcingPlanes];
  if (pinned && ctx->zero_copy) {
    for (int p = 0; p < kForcingPlanes; ++p) {
      void* dp = nullptr;
      CK(cudaHostGetDevicePointer(&dp, (void*)(src[p] + (size_t)d0 * ctx->ngrid), 0));
      planes[p] = (const float*)dp;
    }
    /* bytes that cross the bus: the land cells' values (32-byte sector granularity ignored) */
    ctx->h2d += (int64_t)nd * kForcingPlanes * ctx->dev.nc * (int64_t)sizeof(float);
  } else {
    for (int p = 0; p < kForcingPlanes; ++p) {
      const float* hp = src[p] + (size_t)d0 * ctx->ngrid;
      if (!pinned) { /* pageable caller memory: bounce through our pinned ring */
        memcpy(ctx->h_stage[b] + p * plane, hp, plane * sizeof(float));
        hp = ctx->h_stage[b] + p * plane;
      }
      CK(cudaMemcpyAsync(ctx->d_stage[b] + p * plane, hp, plane * sizeof(float),
                         cudaMemcpyHostToDevice, ctx->copy_stream));
      ctx->h2d += (int64_t)(plane * sizeof(float));
      planes[p] = ctx->d_stage[b] + p * plane;
    }
  }
  CK(cudaEventRecord(ctx->ev_h2d[b], ctx->copy_stream));
  CKL(launch_pack_forcing(ctx->copy_stream, planes, nd, ctx->ngrid, ctx->d_cell_xy, ctx->dev.nc,
                          ctx->dev.ncs, d_out));
  return 0;
}

template <class T>
void gather(const h9_ctx* ctx, const T* src, int per, int take, std::vector<T>& out) {
  /* Fortran (per,lon_c,lat_c) -> compact [ncs][take] */
  out.assign((size_t)ctx->dev.ncs * take, T(0));
  for (int k = 0; k < ctx->dev.nc; ++k) {
    const size_t ci = (size_t)ctx->cell_xy[k];
    for (int i = 0; i < take; ++i) out[(size_t)k * take + i] = src[ci * per + i];
  }
}

template <class T>
void scatter(const h9_ctx* ctx, const T* compact, int take, int per, T* dst) {
  for (int k = 0; k < ctx->dev.nc; ++k) {
    const size_t ci = (size_t)ctx->cell_xy[k];
    for (int i = 0; i < take; ++i) dst[ci * per + i] = compact[(size_t)k * take + i];
  }
}

template <class T>
int upload(h9_ctx* ctx, T* d, const std::vector<T>& h) {
  CK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->h2d += (int64_t)(h.size() * sizeof(T));
  return 0;
}

template <class T>
int download(h9_ctx* ctx, std::vector<T>& h, const T* d, size_t n) {
  h.resize(n);
  CK(cudaMemcpyAsync(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->d2h += (int64_t)(n * sizeof(T));
  return 0;
}

} /* namespace */

extern "C" {

int h9_create(h9_ctx** out, int device_id) {
  if (!out) return H9_ERR_ARG;
  *out = nullptr;
  h9_ctx* ctx = new h9_ctx();
  memset(&ctx->dev, 0, sizeof(ctx->dev));
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    /* no CPU path exists: creation fails loudly */
    fprintf(stderr, "libh9gpu: no usable CUDA device (%s)\n",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    delete ctx;
    return H9_ERR_CUDA;
  }
  if (device_id < 0) {
    if (cudaGetDevice(&ctx->device) != cudaSuccess) ctx->device = 0;
  } else {
    ctx->device = device_id;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    fprintf(stderr, "libh9gpu: cannot initialise device %d: %s\n", ctx->device,
            cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return H9_ERR_CUDA;
  }
  const char* td = getenv("H9_TILE_DAYS");
  if (td && atoi(td) > 0) ctx->tile_days = atoi(td);
  const char* ing = getenv("H9_INGEST");
  if (ing && !strcmp(ing, "gather")) ctx->ingest = 0;
  if (ing && !strcmp(ing, "pack")) ctx->ingest = 1;
  if (ing && !strcmp(ing, "hostread")) ctx->ingest = 3;
  if (ing && !strcmp(ing, "zerocopy")) {
    ctx->ingest = 2;
    ctx->zero_copy = true;
  }
  const char* bl = getenv("H9_BLOCK");
  if (bl && (atoi(bl) == 2000 || atoi(bl) == 4000 || atoi(bl) % 1000 == 32 || atoi(bl) % 1000 == 64 || atoi(bl) % 1000 == 128)) ctx->block = atoi(bl);
  const char* pm = getenv("H9_PAIR_MAX_CELLS");
  if (pm) ctx->pair_max_cells = atoi(pm);
  *out = ctx;
  return H9_OK;
}

int h9_destroy(h9_ctx* ctx) {
  if (!ctx) return H9_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  h9_comm_destroy(ctx);
  free_land(ctx);
  if (ctx->d_year_index) cudaFree(ctx->d_year_index);
  for (auto& pr : ctx->timing) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  for (int b = 0; b < 2; ++b) {
    if (ctx->ev_packed[b]) cudaEventDestroy(ctx->ev_packed[b]);
    if (ctx->ev_done[b]) cudaEventDestroy(ctx->ev_done[b]);
    if (ctx->ev_h2d[b]) cudaEventDestroy(ctx->ev_h2d[b]);
  }
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return H9_OK;
}

const char* h9_last_error(const h9_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int h9_configure(h9_ctx* ctx, int lon_c, int lat_c, int nisurf, const float zi[10], int nyr) {
  if (!ctx) return H9_ERR_ARG;
  if (lon_c < 1 || lat_c < 1 || nisurf < 1 || !zi || nyr < 1)
    return fail(ctx, H9_ERR_ARG, "h9_configure: bad argument");
  if (set_device(ctx)) return H9_ERR_CUDA;
  free_land(ctx);
  ctx->lon_c = lon_c;
  ctx->lat_c = lat_c;
  ctx->nisurf = nisurf;
  ctx->nyr = nyr;
  ctx->ngrid = (size_t)lon_c * lat_c;
  geo_init(ctx->geo, zi, nisurf);
  ctx->cur_year = 0;
  ctx->nt = 0;
  ctx->max_year_seen = 0;
  ctx->step_counter = 0;
  ctx->configured = true;
  return H9_OK;
}

int h9_set_math(h9_ctx* ctx, int mode) {
  if (!ctx) return H9_ERR_ARG;
  if (mode != H9_MATH_EXACT && mode != H9_MATH_FAST) return fail(ctx, H9_ERR_ARG, "h9_set_math: bad mode");
  ctx->math_mode = mode;
  return H9_OK;
}

int h9_set_option(h9_ctx* ctx, int option, int value) {
  if (!ctx) return H9_ERR_ARG;
  if (option != H9_OPT_REAL_EVAP) return fail(ctx, H9_ERR_ARG, "h9_set_option: unknown option");
  ctx->real_evap = value ? 1 : 0;
  ctx->dev.real_evap = ctx->real_evap;
  return H9_OK;
}

int h9_set_tuning(h9_ctx* ctx, int tile_days, int block) {
  if (!ctx) return H9_ERR_ARG;
  if (tile_days > 0) ctx->tile_days = tile_days;
  if (block == 2000 || block == 4000 || block % 1000 == 32 || block % 1000 == 64 || block % 1000 == 128) ctx->block = block;
  return H9_OK;
}

int h9_set_soil(h9_ctx* ctx, const int32_t* soil_tex, const float* theta_s, const float* hksat,
                const float* bsw, const float* psi_s, const float* fmax) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->configured) return fail(ctx, H9_ERR_STATE, "h9_set_soil: call h9_configure first");
  if (!soil_tex || !theta_s || !hksat || !bsw || !psi_s || !fmax)
    return fail(ctx, H9_ERR_ARG, "h9_set_soil: null pointer");
  if (set_device(ctx)) return H9_ERR_CUDA;
  free_land(ctx);
  /* land predicate and iteration order of HYBRID9.f90:120-123: y outer, x inner,
   * SUM(theta_s(:,x,y)) in index order, in float */
  ctx->cell_xy.clear();
  for (int y = 0; y < ctx->lat_c; ++y)
    for (int x = 0; x < ctx->lon_c; ++x) {
      const size_t ci = (size_t)y * ctx->lon_c + x;
      volatile float sum = 0.0f;
      for (int i = 0; i < NL; ++i) sum = sum + theta_s[ci * NL + i];
      if (soil_tex[ci] > 0 && soil_tex[ci] != 13 && sum > 1.0E-8f) ctx->cell_xy.push_back((int32_t)ci);
    }
  DevArrays& d = ctx->dev;
  d.nc = (int)ctx->cell_xy.size();
  d.ncs = std::max(kCellPad, (d.nc + kCellPad - 1) / kCellPad * kCellPad);
  d.nyr = ctx->nyr;
  d.real_evap = ctx->real_evap;
  {
    const char* bc = getenv("H9_BUDGET_CELL"); /* read by -DH9_CYCLE_BUDGET builds only */
    d.budget_cell = bc ? atoi(bc) : 0;
  }
  const size_t n = (size_t)d.ncs;
  int rc = 0;
  float *ths = nullptr, *hks = nullptr, *bs = nullptr, *ps = nullptr, *fm = nullptr;
  rc |= dev_alloc(ctx, &d.h2o, n * NL, true, true);
  rc |= dev_alloc(ctx, &d.smp, n * NL, true, true);
  rc |= dev_alloc(ctx, &d.rootr, n * NL, true, true);
  rc |= dev_alloc(ctx, &ths, n * NL, true, true);
  rc |= dev_alloc(ctx, &hks, n * NL, true, true);
  rc |= dev_alloc(ctx, &bs, n * NL, true, true);
  rc |= dev_alloc(ctx, &ps, n * NL, true, true);
  rc |= dev_alloc(ctx, &fm, n, true, true);
  rc |= dev_alloc(ctx, &d.zwt, n, true, true);
  rc |= dev_alloc(ctx, &d.wa, n, true, true);
  rc |= dev_alloc(ctx, &d.lai, n, true, true);
  rc |= dev_alloc(ctx, &d.lai_litter, n, true, true);
  rc |= dev_alloc(ctx, &d.plant_mass, n, true, true);
  rc |= dev_alloc(ctx, &d.plant_foliage_mass, n, true, true);
  rc |= dev_alloc(ctx, &d.plant_length, n, true, true);
  rc |= dev_alloc(ctx, &d.rdepth, n, true, true);
  rc |= dev_alloc(ctx, &d.nplants, n, true, true);
  rc |= dev_alloc(ctx, &d.rnf_sum, n);
  rc |= dev_alloc(ctx, &d.npp_sum, n);
  rc |= dev_alloc(ctx, &d.plant_mass_sum, n);
  rc |= dev_alloc(ctx, &d.h2osoi_sum_total, n);
  rc |= dev_alloc(ctx, &d.evap_sum, n);
  rc |= dev_alloc(ctx, &d.theta_sum, n * NL);
  rc |= dev_alloc(ctx, &d.annual, (size_t)d.nyr * kAnnualPlanes * n);
  rc |= dev_alloc(ctx, &d.fault, n);
  rc |= dev_alloc(ctx, &d.first_code, n);
  rc |= dev_alloc(ctx, &d.first_step, n);
  rc |= dev_alloc(ctx, &d.first_imb, n);
  rc |= dev_alloc(ctx, &d.first_key, 1);
  rc |= dev_alloc(ctx, &d.any_fault, 1);
  rc |= dev_alloc(ctx, &ctx->d_cell_xy, n, true, true);
  rc |= dev_alloc(ctx, &ctx->d_step_forcing, n * kForcingPlanes, true, true);
  rc |= dev_alloc(ctx, &ctx->d_diag, n * 13, true, true);
  rc |= dev_alloc(ctx, &ctx->d_budget, 8);
  if (rc) {
    free_land(ctx);
    return rc < 0 ? (rc == H9_ERR_NOMEM ? H9_ERR_NOMEM : H9_ERR_CUDA) : H9_ERR_CUDA;
  }
  d.theta_s = ths;
  d.hksat = hks;
  d.bsw = bs;
  d.psi_s = ps;
  d.fmax = fm;
  CK(cudaMemsetAsync(d.first_key, 0xFF, sizeof(unsigned long long), ctx->stream));
  std::vector<float> h;
  /* padding cells get a benign copy of the last land cell's parameters */
  auto up8 = [&](const float* src, float* dst) -> int {
    gather(ctx, src, NL, NL, h);
    for (int k = d.nc; k < d.ncs && d.nc > 0; ++k)
      for (int i = 0; i < NL; ++i) h[(size_t)k * NL + i] = h[(size_t)(d.nc - 1) * NL + i];
    return upload(ctx, dst, h);
  };
  if (up8(theta_s, ths) || up8(hksat, hks) || up8(bsw, bs) || up8(psi_s, ps)) return H9_ERR_CUDA;
  gather(ctx, fmax, 1, 1, h);
  if (upload(ctx, fm, h)) return H9_ERR_CUDA;
  std::vector<int32_t> xy((size_t)d.ncs, d.nc > 0 ? ctx->cell_xy[d.nc - 1] : 0);
  std::copy(ctx->cell_xy.begin(), ctx->cell_xy.end(), xy.begin());
  if (upload(ctx, ctx->d_cell_xy, xy)) return H9_ERR_CUDA;
  ctx->have_soil = true;
  return H9_OK;
}

int64_t h9_num_land(const h9_ctx* ctx) { return (ctx && ctx->have_soil) ? ctx->dev.nc : -1; }

int h9_get_land_index(const h9_ctx* ctx, int32_t* cell_xy) {
  if (!ctx || !ctx->have_soil || !cell_xy) return H9_ERR_ARG;
  std::copy(ctx->cell_xy.begin(), ctx->cell_xy.end(), cell_xy);
  return H9_OK;
}

int h9_set_state(h9_ctx* ctx, const float* h2osoi_liq, const float* zwt, const float* wa,
                 const float* lai, const float* lai_litter, const float* plant_mass,
                 const float* plant_foliage_mass, const float* plant_length, const float* rdepth,
                 const float* rootr_col, const int32_t* nplants, const float* smp) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_set_state: call h9_set_soil first");
  if (!h2osoi_liq || !zwt || !wa || !lai || !lai_litter || !plant_mass || !plant_foliage_mass ||
      !plant_length || !rdepth || !rootr_col || !nplants)
    return fail(ctx, H9_ERR_ARG, "h9_set_state: null pointer");
  if (set_device(ctx)) return H9_ERR_CUDA;
  DevArrays& d = ctx->dev;
  std::vector<float> h;
  gather(ctx, h2osoi_liq, NL, NL, h);
  if (upload(ctx, d.h2o, h)) return H9_ERR_CUDA;
  gather(ctx, rootr_col, 9, NL, h); /* rootr_col(9,...) : element 9 is always 0 (GROW.f90:76) */
  if (upload(ctx, d.rootr, h)) return H9_ERR_CUDA;
  if (smp) {
    gather(ctx, smp, NL, NL, h);
  } else {
    h.assign((size_t)d.ncs * NL, 0.0f); /* INIT.f90:109 leaves smp undefined; we define it as 0 */
  }
  if (upload(ctx, d.smp, h)) return H9_ERR_CUDA;
  struct {
    const float* src;
    float* dst;
  } sc[] = {{zwt, d.zwt},
            {wa, d.wa},
            {lai, d.lai},
            {lai_litter, d.lai_litter},
            {plant_mass, d.plant_mass},
            {plant_foliage_mass, d.plant_foliage_mass},
            {plant_length, d.plant_length},
            {rdepth, d.rdepth}};
  for (auto& e : sc) {
    gather(ctx, e.src, 1, 1, h);
    if (upload(ctx, e.dst, h)) return H9_ERR_CUDA;
  }
  std::vector<int32_t> hi;
  gather(ctx, nplants, 1, 1, hi);
  if (upload(ctx, d.nplants, hi)) return H9_ERR_CUDA;
  return H9_OK;
}

int h9_get_state(h9_ctx* ctx, float* h2osoi_liq, float* zwt, float* wa, float* lai,
                 float* lai_litter, float* plant_mass, float* plant_foliage_mass,
                 float* plant_length, float* rdepth, float* rootr_col, int32_t* nplants, float* smp) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_get_state: no state");
  if (set_device(ctx)) return H9_ERR_CUDA;
  DevArrays& d = ctx->dev;
  const size_t n = (size_t)d.ncs;
  std::vector<float> h;
  if (h2osoi_liq) {
    if (download(ctx, h, d.h2o, n * NL)) return H9_ERR_CUDA;
    scatter(ctx, h.data(), NL, NL, h2osoi_liq);
  }
  if (smp) {
    if (download(ctx, h, d.smp, n * NL)) return H9_ERR_CUDA;
    scatter(ctx, h.data(), NL, NL, smp);
  }
  if (rootr_col) {
    if (download(ctx, h, d.rootr, n * NL)) return H9_ERR_CUDA;
    scatter(ctx, h.data(), NL, 9, rootr_col);
    for (int k = 0; k < d.nc; ++k) rootr_col[(size_t)ctx->cell_xy[k] * 9 + 8] = 0.0f;
  }
  struct {
    float* dst;
    const float* src;
  } sc[] = {{zwt, d.zwt},
            {wa, d.wa},
            {lai, d.lai},
            {lai_litter, d.lai_litter},
            {plant_mass, d.plant_mass},
            {plant_foliage_mass, d.plant_foliage_mass},
            {plant_length, d.plant_length},
            {rdepth, d.rdepth}};
  for (auto& e : sc) {
    if (!e.dst) continue;
    if (download(ctx, h, e.src, n)) return H9_ERR_CUDA;
    scatter(ctx, h.data(), 1, 1, e.dst);
  }
  if (nplants) {
    std::vector<int32_t> hi;
    if (download(ctx, hi, (const int32_t*)d.nplants, n)) return H9_ERR_CUDA;
    scatter(ctx, hi.data(), 1, 1, nplants);
  }
  return H9_OK;
}

int h9_run_days_device(h9_ctx* ctx, int ndays, const int32_t* year_index_of_day,
                       const float* d_forcing, size_t day_stride, size_t plane_stride) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_run_days_device: no soil/state");
  if (ndays < 0 || !year_index_of_day || !d_forcing)
    return fail(ctx, H9_ERR_ARG, "h9_run_days_device: bad argument");
  if (ndays == 0) return H9_OK;
  if (set_device(ctx)) return H9_ERR_CUDA;
  if (upload_year_index(ctx, ndays, year_index_of_day)) return H9_ERR_CUDA;
  if (launch_days(ctx, ndays, ctx->d_year_index,
                  strided_view(d_forcing, day_stride, plane_stride,
                               (size_t)(ndays - 1) * day_stride + (size_t)ctx->dev.ncs)))
    return H9_ERR_CUDA;
  advance_years(ctx, ndays, year_index_of_day);
  uint32_t any = 0;
  if (read_any_fault(ctx, &any)) return H9_ERR_CUDA;
  return (int)any;
}

int h9_run_days(h9_ctx* ctx, int ndays, const int32_t* year_index_of_day, const float* tas,
                const float* rlds, const float* rsds, const float* huss, const float* ps,
                const float* pr, const float* rhs) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_run_days: no soil/state");
  if (ndays < 0 || !year_index_of_day || !tas || !rlds || !rsds || !huss || !ps || !pr || !rhs)
    return fail(ctx, H9_ERR_ARG, "h9_run_days: bad argument");
  if (ndays == 0) return H9_OK;
  if (set_device(ctx)) return H9_ERR_CUDA;
  const float* src[7] = {tas, rlds, rsds, huss, ps, pr, rhs};
  bool pinned = true;
  for (int p = 0; p < 7; ++p) pinned = pinned && is_pinned(src[p]);
  if (upload_year_index(ctx, ndays, year_index_of_day)) return H9_ERR_CUDA;
  const size_t day_stride = (size_t)kForcingPlanes * ctx->dev.ncs;
  const int td = ctx->tile_days;
  const int ntiles = (ndays + td - 1) / td;
  /* Pageable host arrays (or H9_INGEST=gather): two-deep pipeline, tile k+1 is copied on the
   * copy stream (copy engine, no SM) while tile k is stepped on the compute stream; the
   * stepping kernel reads the grid tile exactly as the host holds it, through the land
   * index.  The alternatives with a separate compaction pass (H9_INGEST=pack, =zerocopy) were
   * measured slower: the stepping kernel owns every register of every SM, so any second
   * kernel queues behind it.  Measured on a B200, 0.5 deg year: hostread 65.8 ms, gather
   * 69.1 ms (PCIe-bound: the whole grid, 2.65 GB, crosses the bus), pack 68.4, zerocopy 71.6. */
  if (ctx->ingest == 3 && pinned) {
    /* hostread: no copies at all.  The stepping kernel loads each day's seven values for
     * its cell straight from the host's page-locked arrays (mapped into the device address
     * space), one day ahead of use; only the land cells' values cross PCIe. */
    ForcingView fv;
    for (int p = 0; p < kForcingPlanes; ++p) {
      void* dp = nullptr;
      CK(cudaHostGetDevicePointer(&dp, (void*)src[p], 0));
      fv.plane[p] = (const float*)dp;
    }
    fv.day_stride = ctx->ngrid;
    fv.limit = (size_t)ndays * ctx->ngrid;
    if (launch_days(ctx, ndays, ctx->d_year_index, fv, ctx->d_cell_xy)) return H9_ERR_CUDA;
    ctx->h2d += (int64_t)ndays * kForcingPlanes * ctx->dev.nc * (int64_t)sizeof(float);
    advance_years(ctx, ndays, year_index_of_day);
    uint32_t any_hr = 0;
    if (read_any_fault(ctx, &any_hr)) return H9_ERR_CUDA;
    return (int)any_hr;
  }
  if (ensure_stage(ctx, !pinned)) return H9_ERR_CUDA; /* staging buffers only on the copy paths */
  for (int k = 0; k < ntiles; ++k) {
    const int b = k & 1;
    const int d0 = k * td, nd = std::min(td, ndays - d0);
    if (k >= 2) {
      CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
      if (!pinned) CK(cudaEventSynchronize(ctx->ev_h2d[b])); /* h_stage[b] is free again */
    }
    if (ctx->ingest == 0 || ctx->ingest == 3) { /* gather: H2D only */
      const size_t plane = (size_t)nd * ctx->ngrid;
      for (int p = 0; p < kForcingPlanes; ++p) {
        const float* hp = src[p] + (size_t)d0 * ctx->ngrid;
        if (!pinned) {
          memcpy(ctx->h_stage[b] + p * plane, hp, plane * sizeof(float));
          hp = ctx->h_stage[b] + p * plane;
        }
        CK(cudaMemcpyAsync(ctx->d_stage[b] + p * plane, hp, plane * sizeof(float),
                           cudaMemcpyHostToDevice, ctx->copy_stream));
        ctx->h2d += (int64_t)(plane * sizeof(float));
      }
      CK(cudaEventRecord(ctx->ev_h2d[b], ctx->copy_stream));
      CK(cudaEventRecord(ctx->ev_packed[b], ctx->copy_stream));
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_packed[b], 0));
      if (launch_days(ctx, nd, ctx->d_year_index + d0,
                      strided_view(ctx->d_stage[b], ctx->ngrid, plane, plane), ctx->d_cell_xy))
        return H9_ERR_CUDA;
    } else {
      if (stage_and_pack(ctx, b, d0, nd, src, pinned, ctx->d_packed[b])) return H9_ERR_CUDA;
      CK(cudaEventRecord(ctx->ev_packed[b], ctx->copy_stream));
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_packed[b], 0));
      if (launch_days(ctx, nd, ctx->d_year_index + d0,
                      strided_view(ctx->d_packed[b], day_stride, (size_t)ctx->dev.ncs,
                                   (size_t)(nd - 1) * day_stride + (size_t)ctx->dev.ncs)))
        return H9_ERR_CUDA;
    }
    CK(cudaEventRecord(ctx->ev_done[b], ctx->stream));
    advance_years(ctx, nd, year_index_of_day + d0);
  }
  uint32_t any = 0;
  if (read_any_fault(ctx, &any)) return H9_ERR_CUDA;
  CK(cudaStreamSynchronize(ctx->copy_stream));
  return (int)any;
}

int h9_pack_forcing(h9_ctx* ctx, int ndays, const float* tas, const float* rlds,
                    const float* rsds, const float* huss, const float* ps, const float* pr,
                    const float* rhs, const float** d_forcing, size_t* day_stride,
                    size_t* plane_stride) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_pack_forcing: no soil");
  if (ndays < 1 || !tas || !rlds || !rsds || !huss || !ps || !pr || !rhs || !d_forcing)
    return fail(ctx, H9_ERR_ARG, "h9_pack_forcing: bad argument");
  if (set_device(ctx)) return H9_ERR_CUDA;
  const float* src[7] = {tas, rlds, rsds, huss, ps, pr, rhs};
  bool pinned = true;
  for (int p = 0; p < 7; ++p) pinned = pinned && is_pinned(src[p]);
  if (ensure_stage(ctx, !pinned)) return H9_ERR_CUDA;
  const size_t dstride = (size_t)kForcingPlanes * ctx->dev.ncs;
  const size_t need = (size_t)ndays * dstride;
  if (need > ctx->pack_full_cap) {
    if (ctx->d_pack_full) cudaFree(ctx->d_pack_full);
    ctx->d_pack_full = nullptr;
    ctx->pack_full_cap = 0;
    cudaError_t e = cudaMalloc((void**)&ctx->d_pack_full, need * sizeof(float));
    if (e != cudaSuccess) return fail(ctx, H9_ERR_NOMEM, "h9_pack_forcing: out of device memory");
    ctx->pack_full_cap = need;
  }
  const int td = ctx->tile_days;
  for (int d0 = 0, k = 0; d0 < ndays; d0 += td, ++k) {
    const int b = k & 1;
    const int nd = std::min(td, ndays - d0);
    if (k >= 2) CK(cudaEventSynchronize(ctx->ev_packed[b])); /* stage b consumed */
    if (stage_and_pack(ctx, b, d0, nd, src, pinned, ctx->d_pack_full + (size_t)d0 * dstride))
      return H9_ERR_CUDA;
    CK(cudaEventRecord(ctx->ev_packed[b], ctx->copy_stream));
  }
  CK(cudaStreamSynchronize(ctx->copy_stream));
  *d_forcing = ctx->d_pack_full;
  if (day_stride) *day_stride = dstride;
  if (plane_stride) *plane_stride = (size_t)ctx->dev.ncs;
  return H9_OK;
}

int h9_get_annual(h9_ctx* ctx, int iyr, float* axy_npp, float* axy_plant_mass, float* axy_rnf,
                  float* axy_evap, float* axy_theta_total, float* axy_theta) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_get_annual: no state");
  if (iyr < 1 || iyr > ctx->nyr) return fail(ctx, H9_ERR_ARG, "h9_get_annual: iyr out of range");
  if (iyr > ctx->max_year_seen) return H9_OK; /* year not reached: leave the caller's fill values */
  if (set_device(ctx)) return H9_ERR_CUDA;
  const DevArrays& d = ctx->dev;
  const size_t n = (size_t)d.ncs;
  std::vector<float> h;
  if (download(ctx, h, (const float*)d.annual + (size_t)(iyr - 1) * kAnnualPlanes * n, kAnnualPlanes * n))
    return H9_ERR_CUDA;
  float* dst[5] = {axy_npp, axy_plant_mass, axy_rnf, axy_evap, axy_theta_total};
  for (int k = 0; k < d.nc; ++k) {
    const size_t ci = (size_t)ctx->cell_xy[k];
    for (int p = 0; p < 5; ++p)
      if (dst[p]) dst[p][ci] = h[(size_t)p * n + k];
    if (axy_theta)
      for (int i = 0; i < NL; ++i) axy_theta[ci * NL + i] = h[(size_t)(5 + i) * n + k];
  }
  return H9_OK;
}

int h9_annual_device(h9_ctx* ctx, int iyr, const float** d_means, size_t* plane_stride,
                     const double** d_budget) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_annual_device: no state");
  if (iyr < 1 || iyr > ctx->nyr) return fail(ctx, H9_ERR_ARG, "h9_annual_device: iyr out of range");
  if (set_device(ctx)) return H9_ERR_CUDA;
  const size_t n = (size_t)ctx->dev.ncs;
  if (d_means) *d_means = ctx->dev.annual + (size_t)(iyr - 1) * kAnnualPlanes * n;
  if (plane_stride) *plane_stride = n;
  if (d_budget) {
    CKL(launch_budget(ctx->stream, ctx->dev, iyr, ctx->d_budget));
    *d_budget = ctx->d_budget;
  }
  return H9_OK;
}

int h9_get_fault(h9_ctx* ctx, h9_fault* out) {
  if (!ctx || !out) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_get_fault: no state");
  if (set_device(ctx)) return H9_ERR_CUDA;
  memset(out, 0, sizeof(*out));
  const DevArrays& d = ctx->dev;
  std::vector<uint32_t> f;
  if (download(ctx, f, (const uint32_t*)d.fault, (size_t)d.nc)) return H9_ERR_CUDA;
  for (int k = 0; k < d.nc; ++k)
    if (f[k]) {
      out->any |= f[k];
      out->n_faulted += 1;
    }
  if (!out->any) return H9_OK;
  unsigned long long key = 0;
  CK(cudaMemcpy(&key, d.first_key, sizeof(key), cudaMemcpyDeviceToHost));
  const unsigned cell = (unsigned)(key & 0xFFFFFFFFull);
  if (key == ~0ull || (int)cell >= d.nc) return H9_OK;
  unsigned long long step = 0;
  CK(cudaMemcpy(&out->code, d.first_code + cell, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&step, d.first_step + cell, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&out->imbalance, d.first_imb + cell, 4, cudaMemcpyDeviceToHost));
  ctx->d2h += 24;
  const int32_t ci = ctx->cell_xy[cell];
  out->x = ci % ctx->lon_c + 1;
  out->y = ci / ctx->lon_c + 1;
  out->day = (int32_t)(step / (unsigned long long)ctx->nisurf) + 1;
  out->substep = (int32_t)(step % (unsigned long long)ctx->nisurf) + 1;
  return H9_OK;
}

int h9_clear_fault(h9_ctx* ctx) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return H9_OK;
  if (set_device(ctx)) return H9_ERR_CUDA;
  const DevArrays& d = ctx->dev;
  CK(cudaMemsetAsync(d.fault, 0, (size_t)d.ncs * 4, ctx->stream));
  CK(cudaMemsetAsync(d.any_fault, 0, 4, ctx->stream));
  CK(cudaMemsetAsync(d.first_key, 0xFF, 8, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return H9_OK;
}

void* h9_stream(h9_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

const char* h9_kernel_variant(h9_ctx* ctx) {
  if (!ctx) return "";
  if (ctx->math_mode == H9_MATH_EXACT) return days_variant_exact(ctx->dev.nc, ctx->block);
  if (use_pair(ctx)) return "h9::days_kernel_pair<128>";
  return days_variant_fast(ctx->dev.nc, ctx->block);
}

int h9_synchronize(h9_ctx* ctx) {
  if (!ctx) return H9_ERR_ARG;
  if (set_device(ctx)) return H9_ERR_CUDA;
  CK(cudaStreamSynchronize(ctx->copy_stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->comm_stream) CK(cudaStreamSynchronize(ctx->comm_stream));
  return H9_OK;
}

void* h9_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}

void h9_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int64_t h9_launch_count(const h9_ctx* ctx) { return ctx ? ctx->launches : -1; }
int64_t h9_h2d_bytes(const h9_ctx* ctx) { return ctx ? ctx->h2d : -1; }
int64_t h9_d2h_bytes(const h9_ctx* ctx) { return ctx ? ctx->d2h : -1; }

double h9_step_kernel_ms(h9_ctx* ctx) {
  if (!ctx) return -1.0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  fold_timing(ctx);
  return ctx->step_ms;
}

int h9_reset_counters(h9_ctx* ctx) {
  if (!ctx) return H9_ERR_ARG;
  h9_step_kernel_ms(ctx);
  ctx->step_ms = 0.0;
  ctx->launches = ctx->h2d = ctx->d2h = 0;
  return H9_OK;
}

int h9_hydrology_step(h9_ctx* ctx, const float* tas, const float* rlds, const float* rsds,
                      const float* huss, const float* ps, const float* pr, const float* rhs,
                      float* theta, float* qflx_tran_veg_col, float* qflx_evap_grnd,
                      float* rnf_inc, float* w_imbalance, int32_t* jwt) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_hydrology_step: no soil/state");
  if (!tas || !rlds || !rsds || !huss || !ps || !pr || !rhs)
    return fail(ctx, H9_ERR_ARG, "h9_hydrology_step: null forcing");
  if (set_device(ctx)) return H9_ERR_CUDA;
  const DevArrays& d = ctx->dev;
  const size_t n = (size_t)d.ncs;
  const float* src[7] = {tas, rlds, rsds, huss, ps, pr, rhs};
  std::vector<float> h(n * kForcingPlanes, 0.0f);
  for (int p = 0; p < 7; ++p)
    for (int k = 0; k < d.nc; ++k) h[(size_t)p * n + k] = src[p][ctx->cell_xy[k]];
  if (upload(ctx, ctx->d_step_forcing, h)) return H9_ERR_CUDA;
  StepDiagArrays diag;
  diag.theta = ctx->d_diag;
  diag.qflx_tran_veg_col = ctx->d_diag + 8 * n;
  diag.qflx_evap_grnd = ctx->d_diag + 9 * n;
  diag.rnf_inc = ctx->d_diag + 10 * n;
  diag.w_imbalance = ctx->d_diag + 11 * n;
  diag.jwt = (int32_t*)(ctx->d_diag + 12 * n);
  if (ctx->math_mode == H9_MATH_EXACT) {
    CKL(launch_hydrology_step_exact(ctx->stream, d, ctx->geo, strided_view(ctx->d_step_forcing, 0, n, n),
                                    ctx->step_counter, diag));
  } else if (use_pair(ctx)) {
    CKL(launch_hydrology_step_pair(ctx->stream, d, ctx->geo, strided_view(ctx->d_step_forcing, 0, n, n),
                                   ctx->step_counter, diag));
  } else {
    CKL(launch_hydrology_step_fast_variant(ctx->stream, d, ctx->geo, strided_view(ctx->d_step_forcing, 0, n, n),
                                           ctx->step_counter, diag, ctx->block));
  }
  ctx->step_counter += 1;
  std::vector<float> o;
  if (download(ctx, o, (const float*)ctx->d_diag, n * 13)) return H9_ERR_CUDA;
  if (theta) scatter(ctx, o.data(), NL, NL, theta);
  float* dst[4] = {qflx_tran_veg_col, qflx_evap_grnd, rnf_inc, w_imbalance};
  for (int q = 0; q < 4; ++q)
    if (dst[q]) scatter(ctx, o.data() + (8 + q) * n, 1, 1, dst[q]);
  if (jwt) scatter(ctx, (const int32_t*)(o.data() + 12 * n), 1, 1, jwt);
  uint32_t any = 0;
  if (read_any_fault(ctx, &any)) return H9_ERR_CUDA;
  return (int)any;
}

int h9_grow_day(h9_ctx* ctx, const float* tas, float* npp, float* w_i, float* fT) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_grow_day: no soil/state");
  if (!tas) return fail(ctx, H9_ERR_ARG, "h9_grow_day: null tas");
  if (set_device(ctx)) return H9_ERR_CUDA;
  const DevArrays& d = ctx->dev;
  const size_t n = (size_t)d.ncs;
  std::vector<float> h(n, 0.0f);
  for (int k = 0; k < d.nc; ++k) h[k] = tas[ctx->cell_xy[k]];
  if (upload(ctx, ctx->d_step_forcing, h)) return H9_ERR_CUDA;
  GrowDiagArrays diag;
  diag.npp = ctx->d_diag;
  diag.w_i = ctx->d_diag + n;
  diag.fT = ctx->d_diag + 2 * n;
  if (ctx->math_mode == H9_MATH_EXACT) {
    CKL(launch_grow_day_exact(ctx->stream, d, ctx->geo, ctx->d_step_forcing, diag));
  } else {
    CKL(launch_grow_day_fast(ctx->stream, d, ctx->geo, ctx->d_step_forcing, diag));
  }
  std::vector<float> o;
  if (download(ctx, o, (const float*)ctx->d_diag, n * 3)) return H9_ERR_CUDA;
  if (npp) scatter(ctx, o.data(), 1, 1, npp);
  if (w_i) scatter(ctx, o.data() + n, 1, 1, w_i);
  if (fT) scatter(ctx, o.data() + 2 * n, 1, 1, fT);
  return H9_OK;
}

int h9_regrid_soil_layer(h9_ctx* ctx, int lon_c, int lat_c, int layer, const float* theta_s_in,
                         const float* k_s_in, const float* lambda_in, const float* psi_s_in,
                         float* theta_s, float* hksat, float* bsw, float* psi_s) {
  if (!ctx) return H9_ERR_ARG;
  if (lon_c < 1 || lat_c < 1 || layer < 1 || layer > NL || !theta_s_in || !k_s_in || !lambda_in ||
      !psi_s_in || !theta_s || !hksat || !bsw || !psi_s)
    return fail(ctx, H9_ERR_ARG, "h9_regrid_soil_layer: bad argument");
  if (set_device(ctx)) return H9_ERR_CUDA;
  const float* src[4] = {theta_s_in, k_s_in, lambda_in, psi_s_in};
  float* dst[4] = {theta_s, hksat, bsw, psi_s};
  const size_t fw = (size_t)lon_c * 60;
  /* latitude bands of <= 64 MB per field on the device */
  int band = (int)std::max<size_t>(1, (64u << 20) / (fw * 60 * sizeof(float)));
  band = std::min(band, lat_c);
  float *d_in = nullptr, *d_out = nullptr;
  CK(cudaMalloc((void**)&d_in, 4 * (size_t)band * 60 * fw * sizeof(float)));
  cudaError_t e = cudaMalloc((void**)&d_out, 4 * (size_t)band * lon_c * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(d_in);
    return fail(ctx, H9_ERR_NOMEM, "h9_regrid_soil_layer: out of device memory");
  }
  std::vector<float> h((size_t)4 * band * lon_c);
  int rc = H9_OK;
  for (int y0 = 0; y0 < lat_c && rc == H9_OK; y0 += band) {
    const int rows = std::min(band, lat_c - y0);
    const size_t n = (size_t)rows * 60 * fw;
    for (int f = 0; f < 4 && rc == H9_OK; ++f) {
      if (cudaMemcpyAsync(d_in + f * n, src[f] + (size_t)y0 * 60 * fw, n * sizeof(float),
                          cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        rc = fail(ctx, H9_ERR_CUDA, "h9_regrid_soil_layer: H2D copy failed");
      ctx->h2d += (int64_t)(n * sizeof(float));
    }
    if (rc != H9_OK) break;
    if (launch_regrid_soil(ctx->stream, d_in, lon_c, rows, d_out) != 0) {
      rc = fail(ctx, H9_ERR_CUDA, "h9_regrid_soil_layer: launch failed");
      break;
    }
    ctx->launches += 1;
    if (cudaMemcpyAsync(h.data(), d_out, (size_t)4 * rows * lon_c * sizeof(float), cudaMemcpyDeviceToHost,
                        ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      rc = fail(ctx, H9_ERR_CUDA, std::string("h9_regrid_soil_layer: ") + cudaGetErrorString(cudaGetLastError()));
      break;
    }
    ctx->d2h += (int64_t)((size_t)4 * rows * lon_c * sizeof(float));
    for (int f = 0; f < 4; ++f)
      for (int y = 0; y < rows; ++y)
        for (int x = 0; x < lon_c; ++x)
          dst[f][((size_t)(y0 + y) * lon_c + x) * NL + (layer - 1)] = h[((size_t)f * rows + y) * lon_c + x];
  }
  cudaFree(d_in);
  cudaFree(d_out);
  return rc;
}

int h9_partition_lat_bands(int lon_c, int lat_c, const int32_t* soil_tex, const float* theta_s,
                           int nranks, int32_t* lat_s, int32_t* lat_count, int64_t* n_land) {
  if (lon_c < 1 || lat_c < 1 || !soil_tex || !theta_s || nranks < 1 || !lat_s || !lat_count)
    return H9_ERR_ARG;
  std::vector<int64_t> row(lat_c, 0);
  int64_t total = 0;
  for (int y = 0; y < lat_c; ++y) {
    for (int x = 0; x < lon_c; ++x) {
      const size_t ci = (size_t)y * lon_c + x;
      volatile float sum = 0.0f;
      for (int i = 0; i < NL; ++i) sum = sum + theta_s[ci * NL + i];
      if (soil_tex[ci] > 0 && soil_tex[ci] != 13 && sum > 1.0E-8f) row[y] += 1;
    }
    total += row[y];
  }
  /* greedy prefix split: band r ends at the first row where the running count
   * reaches (r+1)/nranks of the land cells; every band keeps >= 0 rows and the
   * bands tile [1, lat_c] in order (north to south, INIT.f90:145) */
  int y = 0;
  int64_t run = 0;
  for (int r = 0; r < nranks; ++r) {
    const int y0 = y;
    int64_t mine = 0;
    const int64_t target = (total * (r + 1) + nranks - 1) / nranks;
    if (r == nranks - 1) {
      while (y < lat_c) mine += row[y++];
    } else {
      while (y < lat_c && run + mine < target) {
        /* stop before a row that would overshoot more than it undershoots */
        const int64_t after = run + mine + row[y];
        if (after > target && (after - target) > (target - (run + mine)) && mine > 0) break;
        mine += row[y++];
      }
    }
    run += mine;
    lat_s[r] = y0 + 1;
    lat_count[r] = y - y0;
    if (n_land) n_land[r] = mine;
  }
  return H9_OK;
}

} /* extern "C" */

/* ---- multi-GPU: NCCL behind the C ABI ------------------------------------------------- */

namespace {

/* libnccl is loaded on first use so that libh9gpu.so itself has no NCCL dependency (a
 * single-GPU Fortran host needs none).  A process that already holds a libnccl.so.2 (MPI
 * launcher, torch) gets that copy: dlopen matches the loaded SONAME. */
struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib) return &api;
  const char* names[] = {getenv("H9_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
    api.err = dlerror();
  }
  if (!api.lib) return &api;
#define H9_NCCL_SYM(field, sym)                                 \
  api.field = (decltype(api.field))dlsym(api.lib, #sym);        \
  if (!api.field) {                                             \
    api.err = "libnccl: missing symbol " #sym;                  \
    dlclose(api.lib);                                           \
    api.lib = nullptr;                                          \
    return &api;                                                \
  }
  H9_NCCL_SYM(GetUniqueId, ncclGetUniqueId)
  H9_NCCL_SYM(CommInitRank, ncclCommInitRank)
  H9_NCCL_SYM(CommDestroy, ncclCommDestroy)
  H9_NCCL_SYM(AllReduce, ncclAllReduce)
  H9_NCCL_SYM(AllGather, ncclAllGather)
  H9_NCCL_SYM(Broadcast, ncclBroadcast)
  H9_NCCL_SYM(GroupStart, ncclGroupStart)
  H9_NCCL_SYM(GroupEnd, ncclGroupEnd)
  H9_NCCL_SYM(GetErrorString, ncclGetErrorString)
#undef H9_NCCL_SYM
  return &api;
}

#define CKN(call)                                                                          \
  do {                                                                                     \
    ncclResult_t r_ = (call);                                                              \
    if (r_ != ncclSuccess)                                                                 \
      return fail(ctx, H9_ERR_NCCL, std::string(#call) + ": " + nccl_api()->GetErrorString(r_)); \
  } while (0)

/* first collective after h9_set_soil: every rank learns every rank's cell count and stride
 * (one small all-gather and one host sync, once), and the persistent buffers are sized */
int comm_prepare(h9_ctx* ctx) {
  NcclApi* n = nccl_api();
  if (!ctx->land_of_rank.empty()) return 0;
  const int R = ctx->nranks;
  long long* d_tmp = nullptr;
  CK(cudaMalloc((void**)&d_tmp, sizeof(long long) * 2 * (size_t)(R + 1)));
  const long long mine[2] = {(long long)ctx->dev.nc, (long long)ctx->dev.ncs};
  CK(cudaMemcpyAsync(d_tmp + 2 * R, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
  CKN(n->AllGather(d_tmp + 2 * R, d_tmp, 2, ncclInt64, ctx->comm, ctx->stream));
  std::vector<long long> all(2 * (size_t)R);
  CK(cudaMemcpyAsync(all.data(), d_tmp, sizeof(long long) * 2 * R, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_tmp);
  ctx->land_of_rank.resize(R);
  ctx->stride_of_rank.resize(R);
  size_t elems = 0;
  for (int r = 0; r < R; ++r) {
    ctx->land_of_rank[r] = all[2 * r];
    ctx->stride_of_rank[r] = all[2 * r + 1];
    elems += (size_t)kAnnualPlanes * (size_t)all[2 * r + 1];
  }
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  if (ctx->d_budget_all) cudaFree(ctx->d_budget_all);
  ctx->d_gather = nullptr;
  ctx->d_budget_all = nullptr;
  CK(cudaMalloc((void**)&ctx->d_gather, elems * sizeof(float)));
  CK(cudaMalloc((void**)&ctx->d_budget_all, (size_t)ctx->nyr * 8 * sizeof(double)));
  CK(cudaMemsetAsync(ctx->d_budget_all, 0, (size_t)ctx->nyr * 8 * sizeof(double), ctx->stream));
  ctx->gather_elems = elems;
  return 0;
}

} /* namespace */

extern "C" {

int h9_comm_unique_id(void* id) {
  if (!id) return H9_ERR_ARG;
  NcclApi* n = nccl_api();
  if (!n->lib) {
    fprintf(stderr, "libh9gpu: cannot load libnccl (%s); set H9_NCCL_LIB\n", n->err.c_str());
    return H9_ERR_NCCL;
  }
  static_assert(sizeof(ncclUniqueId) == H9_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  if (n->GetUniqueId(&u) != ncclSuccess) return H9_ERR_NCCL;
  memcpy(id, &u, sizeof(u));
  return H9_OK;
}

int h9_comm_init(h9_ctx* ctx, int nranks, int rank, const void* id) {
  if (!ctx) return H9_ERR_ARG;
  if (nranks < 1 || rank < 0 || rank >= nranks || !id) return fail(ctx, H9_ERR_ARG, "h9_comm_init: bad argument");
  if (ctx->comm) return fail(ctx, H9_ERR_STATE, "h9_comm_init: communicator already initialised");
  NcclApi* n = nccl_api();
  if (!n->lib) return fail(ctx, H9_ERR_NCCL, "h9_comm_init: cannot load libnccl (" + n->err + "); set H9_NCCL_LIB");
  if (set_device(ctx)) return H9_ERR_CUDA;
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  CKN(n->CommInitRank(&ctx->comm, nranks, u, rank));
  if (!ctx->comm_stream) CK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  if (!ctx->ev_coll_ready) CK(cudaEventCreateWithFlags(&ctx->ev_coll_ready, cudaEventDisableTiming));
  if (!ctx->ev_coll_done) CK(cudaEventCreateWithFlags(&ctx->ev_coll_done, cudaEventDisableTiming));
  ctx->nranks = nranks;
  ctx->rank = rank;
  ctx->land_of_rank.clear();
  ctx->stride_of_rank.clear();
  return H9_OK;
}

int h9_comm_destroy(h9_ctx* ctx) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->comm) return H9_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
  nccl_api()->CommDestroy(ctx->comm);
  ctx->comm = nullptr;
  ctx->coll_inflight = false;
  if (ctx->ev_coll_ready) cudaEventDestroy(ctx->ev_coll_ready);
  if (ctx->ev_coll_done) cudaEventDestroy(ctx->ev_coll_done);
  ctx->ev_coll_ready = ctx->ev_coll_done = nullptr;
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  ctx->comm_stream = nullptr;
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  if (ctx->d_budget_all) cudaFree(ctx->d_budget_all);
  ctx->d_gather = nullptr;
  ctx->d_budget_all = nullptr;
  ctx->land_of_rank.clear();
  ctx->stride_of_rank.clear();
  ctx->nranks = 1;
  ctx->rank = 0;
  return H9_OK;
}

int h9_comm_land_counts(h9_ctx* ctx, int64_t* n_land) {
  if (!ctx || !n_land) return H9_ERR_ARG;
  if (!ctx->comm) return fail(ctx, H9_ERR_STATE, "h9_comm_land_counts: call h9_comm_init first");
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_comm_land_counts: call h9_set_soil first");
  if (set_device(ctx)) return H9_ERR_CUDA;
  if (comm_prepare(ctx)) return ctx->err.find("nccl") != std::string::npos ? H9_ERR_NCCL : H9_ERR_CUDA;
  for (int r = 0; r < ctx->nranks; ++r) n_land[r] = ctx->land_of_rank[r];
  return H9_OK;
}

int h9_annual_collective(h9_ctx* ctx, int iyr) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->comm) return fail(ctx, H9_ERR_STATE, "h9_annual_collective: call h9_comm_init first");
  if (!ctx->have_soil) return fail(ctx, H9_ERR_STATE, "h9_annual_collective: no state");
  if (iyr < 1 || iyr > ctx->nyr) return fail(ctx, H9_ERR_ARG, "h9_annual_collective: iyr out of range");
  if (set_device(ctx)) return H9_ERR_CUDA;
  if (comm_prepare(ctx)) return H9_ERR_CUDA;
  NcclApi* n = nccl_api();
  /* K5 is enqueued on the stepping stream right behind the stepping kernel and writes the
   * year's own budget slot; the two collectives follow behind an event: no host
   * synchronisation, persistent buffers */
  double* slot = ctx->d_budget_all + (size_t)(iyr - 1) * 8;
  if (ctx->coll_inflight) { /* the previous collective (a year ago) before its buffers are reused */
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_coll_done, 0));
    ctx->coll_inflight = false;
  }
  CKL(launch_budget(ctx->stream, ctx->dev, iyr, slot));
  /* Default: the collectives stay on the stepping stream, between two years.  Running them on
   * the communication stream (env H9_COMM_OVERLAP=1) lets the next year's stepping kernel start
   * at once, but the NCCL kernels then share the SMs with it, and a rank that finished early
   * spins in them until the slowest rank arrives: measured per simulated year, one 0.5 deg grid
   * on 2 GPUs 31.2 -> 31.1 ms, on 8 GPUs (bands of unequal cost, one warp per scheduler) 20.0 ->
   * 21.4 ms, replicated grids 58.5 -> 69 ms (profiles/r02/README.md). */
  static const bool overlap = [] { const char* e = getenv("H9_COMM_OVERLAP"); return e && atoi(e) != 0; }();
  cudaStream_t cs = overlap ? ctx->comm_stream : ctx->stream;
  CK(cudaEventRecord(ctx->ev_coll_ready, ctx->stream));
  CK(cudaStreamWaitEvent(cs, ctx->ev_coll_ready, 0));
  CKN(n->AllReduce(slot, slot, 8, ncclDouble, ncclSum, ctx->comm, cs));
  /* ragged all-gather of the 13 annual-mean planes: rank r's [13][ncs_r] block lands at its
   * offset in every rank's buffer (grouped broadcasts: shards differ in size, no padding) */
  const float* mine = ctx->dev.annual + (size_t)(iyr - 1) * kAnnualPlanes * ctx->dev.ncs;
  CKN(n->GroupStart());
  size_t off = 0;
  for (int r = 0; r < ctx->nranks; ++r) {
    const size_t cnt = (size_t)kAnnualPlanes * (size_t)ctx->stride_of_rank[r];
    ncclResult_t rr = n->Broadcast(r == ctx->rank ? (const void*)mine : (const void*)(ctx->d_gather + off),
                                   ctx->d_gather + off, cnt, ncclFloat, r, ctx->comm, cs);
    if (rr != ncclSuccess) {
      n->GroupEnd();
      return fail(ctx, H9_ERR_NCCL, std::string("ncclBroadcast: ") + n->GetErrorString(rr));
    }
    off += cnt;
  }
  CKN(n->GroupEnd());
  CK(cudaEventRecord(ctx->ev_coll_done, cs));
  ctx->coll_inflight = true;
  ctx->gathered_year = iyr;
  return H9_OK;
}

int h9_get_gathered_annual(h9_ctx* ctx, int r, float* planes) {
  if (!ctx || !planes) return H9_ERR_ARG;
  if (!ctx->comm || ctx->gathered_year < 1) return fail(ctx, H9_ERR_STATE, "h9_get_gathered_annual: no collective has run");
  if (r < 0 || r >= ctx->nranks) return fail(ctx, H9_ERR_ARG, "h9_get_gathered_annual: bad rank");
  if (set_device(ctx)) return H9_ERR_CUDA;
  size_t off = 0;
  for (int q = 0; q < r; ++q) off += (size_t)kAnnualPlanes * (size_t)ctx->stride_of_rank[q];
  const size_t nl = (size_t)ctx->land_of_rank[r], st = (size_t)ctx->stride_of_rank[r];
  if (nl == 0) return H9_OK;
  CK(cudaEventSynchronize(ctx->ev_coll_done)); /* whichever stream the collective ran on */
  CK(cudaMemcpy2DAsync(planes, nl * sizeof(float), ctx->d_gather + off, st * sizeof(float),
                       nl * sizeof(float), kAnnualPlanes, cudaMemcpyDeviceToHost, ctx->comm_stream));
  CK(cudaStreamSynchronize(ctx->comm_stream));
  ctx->d2h += (int64_t)(nl * kAnnualPlanes * sizeof(float));
  return H9_OK;
}

int h9_get_budget(h9_ctx* ctx, int iyr, double* budget) {
  if (!ctx || !budget) return H9_ERR_ARG;
  if (!ctx->comm || !ctx->d_budget_all) return fail(ctx, H9_ERR_STATE, "h9_get_budget: no collective has run");
  if (iyr < 1 || iyr > ctx->nyr) return fail(ctx, H9_ERR_ARG, "h9_get_budget: iyr out of range");
  if (set_device(ctx)) return H9_ERR_CUDA;
  CK(cudaEventSynchronize(ctx->ev_coll_done));
  CK(cudaMemcpyAsync(budget, ctx->d_budget_all + (size_t)(iyr - 1) * 8, 8 * sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->comm_stream));
  CK(cudaStreamSynchronize(ctx->comm_stream));
  ctx->d2h += 64;
  return H9_OK;
}

int h9_gathered_device(h9_ctx* ctx, const float** d_planes, const double** d_budget_years) {
  if (!ctx) return H9_ERR_ARG;
  if (!ctx->comm || !ctx->d_gather) return fail(ctx, H9_ERR_STATE, "h9_gathered_device: no collective has run");
  if (set_device(ctx)) return H9_ERR_CUDA;
  if (ctx->coll_inflight) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_coll_done, 0));
  if (d_planes) *d_planes = ctx->d_gather;
  if (d_budget_years) *d_budget_years = ctx->d_budget_all;
  return H9_OK;
}

} /* extern "C" */
