// capi.cpp — the C ABI of libqcoh.so (include/qcoh.h): handles, ownership, error channel.
// Everything numerical is a kernel launch (kernels.cu); there is no CPU compute path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/qcoh.h"
#include "forest.hpp"
#include "kernels.hpp"

using namespace qcoh;

namespace {

thread_local std::string g_err;

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CU(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      throw Error(std::string("CUDA error in " #call ": ") + cudaGetErrorName(e__) + " — " +          \
                  cudaGetErrorString(e__));                                                           \
  } while (0)

#define API_BEGIN try {
#define API_END                      \
  }                                  \
  catch (const std::exception &e) {  \
    g_err = e.what();                \
    return -1;                       \
  }                                  \
  catch (...) {                      \
    g_err = "unknown error";         \
    return -1;                       \
  }                                  \
  return 0;

// ---- device context ---------------------------------------------------------------
struct Ctx {
  bool ready = false;
  int device = -1;
  cudaStream_t stream = nullptr;       // compute (and everything ordered with it)
  cudaStream_t copy_stream = nullptr;  // H2D of matrix chunks
  cudaStream_t d2h_stream = nullptr;   // D2H of result chunks
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> chunk_events;
  int speculate = 1;                   // pipeline prediction into XGDMatrixCreateFromMat
  uint64_t chunk_rows = 1ull << 21;
  void *flush = nullptr;
  size_t flush_bytes = 0;
  Tunables tun;
} g;

int requested_device = -1;

void ensure_device() {
  if (g.ready) return;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    throw Error(std::string("libqcoh: no CUDA device is visible (") + (e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)) +
                "); the OH path runs on the GPU only, there is no CPU fallback");
  }
  int dev = requested_device;
  if (dev < 0) {
    const char *lr = getenv("LOCAL_RANK");
    dev = lr ? atoi(lr) % n : 0;
  }
  if (dev >= n) throw Error("libqcoh: device " + std::to_string(dev) + " requested but only " + std::to_string(n) + " visible");
  CU(cudaSetDevice(dev));
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, dev));
  if (p.major < 10)
    throw Error(std::string("libqcoh is built for sm_100a (B200); device '") + p.name + "' is sm_" + std::to_string(p.major) +
                std::to_string(p.minor));
  CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&g.d2h_stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&g.ev0));
  CU(cudaEventCreate(&g.ev1));
  g.device = dev;
  g.ready = true;
}

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  T *need(size_t n) {
    if (n > cap) {
      if (p) cudaFree(p);
      p = nullptr, cap = 0;
      CU(cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T)));
      cap = n;
    }
    return p;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr, cap = 0;
  }
  void swap(DevBuf &o) {
    std::swap(p, o.p);
    std::swap(cap, o.cap);
  }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
};

template <class T>
struct PinBuf {
  T *p = nullptr;
  size_t cap = 0;
  T *need(size_t n) {
    if (n > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr, cap = 0;
      CU(cudaHostAlloc((void **)&p, (n ? n : 1) * sizeof(T), cudaHostAllocDefault));
      cap = n;
    }
    return p;
  }
  void swap(PinBuf &o) {
    std::swap(p, o.p);
    std::swap(cap, o.cap);
  }
  PinBuf() = default;
  PinBuf(const PinBuf &) = delete;
  PinBuf &operator=(const PinBuf &) = delete;
  ~PinBuf() {
    if (p) cudaFreeHost(p);
  }
};

bool is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- handles ------------------------------------------------------------------------
constexpr uint32_t kBoosterMagic = 0x51434253;  // 'QCBS'
constexpr uint32_t kDMatrixMagic = 0x5143444d;  // 'QCDM'
constexpr uint32_t kOhMagic = 0x51434f48;       // 'QCOH'

struct Booster {
  uint32_t magic = kBoosterMagic;
  uint64_t version = 0;  // changes with every (re)load
  bool loaded = false, uploaded = false;
  HostForest host;
  FlatForest flat;
  DeviceForest dev;
  DevBuf<uint2> d_nodes;
  DevBuf<uint32_t> d_off;
  DevBuf<int32_t> d_depth, d_orig;
  DevBuf<float> d_result;
  PinBuf<float> h_result;
};

struct DMatrix {
  uint32_t magic = kDMatrixMagic;
  uint64_t nrow = 0, ncol = 0;
  float missing = NAN;
  DevBuf<float> X;
  DevBuf<int> flags;
  int hflags = 1;  // bit0 has-missing, bit1 has-inf; conservative until sealed
  bool sealed = false;
  // prediction pipelined into XGDMatrixCreateFromMat (see create_pipelined)
  const void *spec_booster = nullptr;
  uint64_t spec_version = 0;
  bool spec_ready = false;
  DevBuf<float> spec_dev;
  PinBuf<float> spec_host;
};

// XGDMatrixFree keeps the largest freed matrix buffer for the next XGDMatrixCreateFromMat: the
// reference creates and frees a same-sized DMatrix on every call (OH_GridCompMod.F90:347,377) and
// cudaMalloc / cudaFree of multi-GB buffers would otherwise dominate the step.
DevBuf<float> g_spare_X;
PinBuf<float> g_spare_pin;
DevBuf<float> g_spare_spec;
DevBuf<int> g_chunk_flags;
PinBuf<int> g_h_chunk_flags;
struct Booster;
Booster *g_last_booster = nullptr;  // the process's booster (the reference keeps exactly one, SAVE :182)
uint64_t g_version_counter = 0;

Booster *B(BoosterHandle h) {
  Booster *b = (Booster *)h;
  if (!b || b->magic != kBoosterMagic) throw Error("Invalid booster handle");
  return b;
}
DMatrix *D(DMatrixHandle h) {
  DMatrix *d = (DMatrix *)h;
  if (!d || d->magic != kDMatrixMagic) throw Error("Invalid DMatrix handle");
  return d;
}

void upload(Booster *b) {
  if (b->uploaded) return;
  if (!b->loaded) throw Error("Booster has no model: call XGBoosterLoadModel first");
  ensure_device();
  const FlatForest &f = b->flat;
  const size_t nn = (size_t)f.num_nodes();
  {
    // device copy: an internal node's x word becomes -key(threshold) mod 2^32 (kernels.cu "Order-
    // preserving integer keys"); leaves keep the float bits of their value
    std::vector<uint32_t> dev_nodes(f.nodes_xy);
    for (size_t i = 0; i < nn; ++i) {
      if ((dev_nodes[2 * i + 1] & kMetaRelMask) == 0) continue;
      float thr;
      memcpy(&thr, &dev_nodes[2 * i], 4);
      if (std::isnan(thr)) throw Error("split threshold is NaN (node " + std::to_string(i) + ")");
      thr += 0.0f;  // -0.0 -> +0.0
      uint32_t bits;
      memcpy(&bits, &thr, 4);
      const uint32_t key = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
      dev_nodes[2 * i] = 0u - key;  // key >= 0x007FFFFF (-inf), never 0
    }
    CU(cudaMemcpy(b->d_nodes.need(nn), dev_nodes.data(), nn * 8, cudaMemcpyHostToDevice));
  }
  CU(cudaMemcpy(b->d_off.need(f.tree_offset.size()), f.tree_offset.data(), f.tree_offset.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(b->d_depth.need(f.tree_depth.size()), f.tree_depth.data(), f.tree_depth.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(b->d_orig.need(nn), f.orig_id.data(), nn * 4, cudaMemcpyHostToDevice));
  b->dev.nodes = b->d_nodes.p, b->dev.tree_offset = b->d_off.p, b->dev.tree_depth = b->d_depth.p, b->dev.orig_id = b->d_orig.p;
  b->dev.ntree = (int32_t)b->host.trees.size();
  b->dev.nfeat = (int32_t)b->host.num_feature;
  b->dev.max_depth = f.max_depth;
  b->dev.num_nodes = f.num_nodes();
  b->dev.base_score = b->host.base_score;
  if (b->dev.tex) cudaDestroyTextureObject(b->dev.tex);
  b->dev.tex = 0;
  if (nn > 0 && nn < ((size_t)1 << 27)) {
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = b->d_nodes.p;
    rd.res.linear.desc = cudaCreateChannelDesc<uint2>();
    rd.res.linear.sizeInBytes = nn * 8;
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.readMode = cudaReadModeElementType;
    CU(cudaCreateTextureObject(&b->dev.tex, &rd, &td, nullptr));
  }
  b->uploaded = true;
}

void seal(DMatrix *d) {
  ensure_device();
  int *fl = d->flags.need(1);
  CU(cudaMemsetAsync(fl, 0, sizeof(int), g.stream));
  CU(launch_scan_matrix(d->X.p, d->nrow * d->ncol, d->missing, fl, g.stream));
  CU(cudaMemcpyAsync(&d->hflags, fl, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  d->sealed = true;
  // xgboost src/data/data.cc SparsePage::Push: a finite `missing` with inf data is an error
  if (d->hflags & 2) throw Error("Check failed: valid: Input data contains `inf` or `nan`");
}

unsigned trees_used(const Booster *b, unsigned ntree_limit) {
  const unsigned nt = (unsigned)b->host.trees.size();
  return (ntree_limit == 0 || ntree_limit > nt) ? nt : ntree_limit;
}

void predict_into(Booster *b, DMatrix *d, int option_mask, unsigned ntree_limit, const qcoh_epilogue *epi, float *out_dev) {
  upload(b);
  if (!d->sealed) seal(d);
  if (option_mask & ~3) throw Error("option_mask " + std::to_string(option_mask) + ": only 0 (value), 1 (margin) and 2 (leaf index) are supported");
  if (d->ncol > b->host.num_feature)
    throw Error("Check failed: Number of columns does not match number of features in booster. Columns: " +
                std::to_string(d->ncol) + " Features: " + std::to_string(b->host.num_feature));
  PredictArgs a;
  a.X = d->X.p, a.nrow = d->nrow, a.ncol = (int32_t)d->ncol, a.missing = d->missing;
  a.has_missing = ((d->hflags & 1) || d->ncol < b->host.num_feature) ? 1 : 0;
  a.pred_leaf = (option_mask & 2) ? 1 : 0;
  a.ntree_used = (int32_t)trees_used(b, ntree_limit);
  a.exp10 = epi ? epi->exp10 : 0;
  a.scale = epi ? epi->scale : 1.f;
  a.out = out_dev;
  CU(launch_predict(b->dev, a, g.tun, g.stream));
}

cudaEvent_t chunk_event(size_t i) {
  while (g.chunk_events.size() <= i) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g.chunk_events.push_back(e);
  }
  return g.chunk_events[i];
}

void drain() {
  cudaStreamSynchronize(g.copy_stream);
  cudaStreamSynchronize(g.stream);
  cudaStreamSynchronize(g.d2h_stream);
}

bool is_pinned_host(const void *p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// multi-threaded memcpy (pageable host -> pinned staging)
void parallel_memcpy(void *dst, const void *src, size_t bytes) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt > 16) nt = 16;
  if (bytes < ((size_t)8 << 20)) nt = 1;
  if (nt == 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = (bytes / nt + 4095) / 4096 * 4096;
  for (unsigned t = 0; t < nt; ++t) {
    const size_t o = (size_t)t * per;
    if (o >= bytes) break;
    const size_t n = std::min(per, bytes - o);
    th.emplace_back([=] { memcpy((char *)dst + o, (const char *)src + o, n); });
  }
  for (auto &t : th) t.join();
}

constexpr int kStageSlots = 3;
PinBuf<float> g_stage[kStageSlots];
cudaEvent_t g_stage_free[kStageSlots] = {nullptr, nullptr, nullptr};

// XGDMatrixCreateFromMat from HOST memory, pipelined: the matrix is cut into row chunks; while chunk
// c+1 crosses PCIe, chunk c is scanned (missing / inf) and — since the reference keeps exactly one
// booster per process and always predicts with option_mask = 0, ntree_limit = 0 right after creating
// the matrix (OH_GridCompMod.F90:347-356) — already predicted with that booster, and its results
// stream back into a pinned buffer.  XGBoosterPredict then finds the answer ready if it is called
// with that booster and those options; any other call takes the ordinary path.  Pageable host
// memory (what a Fortran ALLOCATE gives) is staged through a ring of pinned buffers by a threaded
// memcpy so that the DMA engine never waits on the driver's own bounce buffer.  The caller's buffer
// is fully consumed before this returns.
//   copy_stream : H2D(c) -> scan(c) -> flag D2H(c)      g.stream : predict(c)      d2h_stream : result D2H(c)
void create_pipelined(DMatrix *d, const float *data, Booster *b) {
  const uint64_t nrow = d->nrow, ncol = d->ncol;
  const bool pinned = is_pinned_host(data);
  uint64_t cr = g.chunk_rows;
  if (!pinned && cr > (1ull << 19)) cr = 1ull << 19;
  const size_t nchunk = (size_t)((nrow + cr - 1) / cr);
  float *X = d->X.p;
  const bool spec = b != nullptr;
  float *sdev = nullptr, *shost = nullptr;
  if (spec) {
    upload(b);
    if (g_spare_spec.cap >= nrow && g_spare_spec.p) d->spec_dev.swap(g_spare_spec);
    sdev = d->spec_dev.need(nrow);
    if (g_spare_pin.cap >= nrow && g_spare_pin.p) d->spec_host.swap(g_spare_pin);
    shost = d->spec_host.need(nrow);
  }
  int *fl = g_chunk_flags.need(nchunk);
  int *hfl = g_h_chunk_flags.need(nchunk);
  CU(cudaMemsetAsync(fl, 0, nchunk * sizeof(int), g.copy_stream));
  if (!pinned)
    for (int s = 0; s < kStageSlots; ++s)
      if (!g_stage_free[s]) CU(cudaEventCreateWithFlags(&g_stage_free[s], cudaEventDisableTiming));
  int flags = 0;
  auto issue = [&](size_t c) {
    const uint64_t r0 = c * cr, nr = std::min(cr, nrow - r0);
    const size_t bytes = nr * ncol * sizeof(float);
    const float *src = data + r0 * ncol;
    if (!pinned) {
      const int s = (int)(c % kStageSlots);
      float *st = g_stage[s].need(cr * ncol);
      if (c >= (size_t)kStageSlots) CU(cudaEventSynchronize(g_stage_free[s]));  // its previous H2D has drained
      parallel_memcpy(st, src, bytes);
      src = st;
      CU(cudaMemcpyAsync(X + r0 * ncol, src, bytes, cudaMemcpyHostToDevice, g.copy_stream));
      CU(cudaEventRecord(g_stage_free[s], g.copy_stream));
    } else {
      CU(cudaMemcpyAsync(X + r0 * ncol, src, bytes, cudaMemcpyHostToDevice, g.copy_stream));
    }
    CU(launch_scan_matrix(X + r0 * ncol, nr * ncol, d->missing, fl + c, g.copy_stream));
    CU(cudaMemcpyAsync(hfl + c, fl + c, sizeof(int), cudaMemcpyDeviceToHost, g.copy_stream));
    CU(cudaEventRecord(chunk_event(2 * c), g.copy_stream));
  };
  auto process = [&](size_t c) {
    const uint64_t r0 = c * cr, nr = std::min(cr, nrow - r0);
    CU(cudaEventSynchronize(chunk_event(2 * c)));
    flags |= hfl[c];
    if (hfl[c] & 2) {
      drain();
      throw Error("Check failed: valid: Input data contains `inf` or `nan`");
    }
    if (!spec) return;
    PredictArgs a;
    a.X = X + r0 * ncol, a.nrow = nr, a.ncol = (int32_t)ncol, a.missing = d->missing;
    a.has_missing = ((hfl[c] & 1) || ncol < b->host.num_feature) ? 1 : 0;
    a.ntree_used = (int32_t)b->host.trees.size();
    a.out = sdev + r0;
    CU(launch_predict(b->dev, a, g.tun, g.stream));
    CU(cudaEventRecord(chunk_event(2 * c + 1), g.stream));
    CU(cudaStreamWaitEvent(g.d2h_stream, chunk_event(2 * c + 1), 0));
    CU(cudaMemcpyAsync(shost + r0, sdev + r0, nr * sizeof(float), cudaMemcpyDeviceToHost, g.d2h_stream));
  };
  // pinned source: every copy can be queued up front; pageable: stage one chunk ahead of the GPU
  const size_t lookahead = pinned ? nchunk : 1;
  for (size_t c = 0; c < nchunk + lookahead; ++c) {
    if (c < nchunk) issue(c);
    if (c >= lookahead && c - lookahead < nchunk) process(c - lookahead);
  }
  CU(cudaStreamSynchronize(g.copy_stream));  // the borrowed host buffer has been read completely
  d->hflags = flags;
  d->sealed = true;
  if (spec) d->spec_booster = b, d->spec_version = b->version, d->spec_ready = true;
}

}  // namespace

// =====================================================================================
// (1) xgb_fortran_api boundary
// =====================================================================================
extern "C" {

const char *XGBGetLastError(void) { return g_err.c_str(); }

int XGBoosterCreate(const DMatrixHandle dmats[], bst_ulong len, BoosterHandle *out) {
  API_BEGIN
  // The reference passes a DMatrix handle by value here with len = 0 (OH_GridCompMod.F90:255-256):
  // dmats must not be dereferenced.  Cached matrices are a training concept; ignored for len > 0.
  (void)dmats, (void)len;
  if (!out) throw Error("XGBoosterCreate: out is NULL");
  *out = new Booster();
  API_END
}

int XGBoosterFree(BoosterHandle handle) {
  API_BEGIN
  Booster *b = B(handle);
  if (g_last_booster == b) g_last_booster = nullptr;
  if (g.ready) CU(cudaStreamSynchronize(g.stream));
  b->magic = 0;
  delete b;
  API_END
}

int qcoh_booster_parse(BoosterHandle handle, const char *fname) {
  API_BEGIN
  Booster *b = B(handle);
  if (!fname) throw Error("model file name is NULL");
  HostForest hf = load_model_file(fname);
  FlatForest ff = flatten(hf);
  b->host = std::move(hf), b->flat = std::move(ff);
  b->loaded = true, b->uploaded = false;
  b->version = ++g_version_counter;
  API_END
}

int XGBoosterLoadModel(BoosterHandle handle, const char *fname) {
  if (qcoh_booster_parse(handle, fname) != 0) return -1;
  API_BEGIN
  upload(B(handle));
  g_last_booster = B(handle);
  API_END
}

int XGBoosterSaveModel(BoosterHandle handle, const char *fname) {
  API_BEGIN
  Booster *b = B(handle);
  if (!b->loaded) throw Error("Booster has no model to save");
  save_model_file(b->host, fname);
  API_END
}

int XGDMatrixCreateFromMat(const float *data, bst_ulong nrow, bst_ulong ncol, float missing, DMatrixHandle *out) {
  API_BEGIN
  if (!out) throw Error("XGDMatrixCreateFromMat: out is NULL");
  if (!data && nrow * ncol) throw Error("XGDMatrixCreateFromMat: data is NULL");
  ensure_device();
  std::unique_ptr<DMatrix> d(new DMatrix());
  d->nrow = nrow, d->ncol = ncol, d->missing = missing;
  const size_t n = (size_t)nrow * ncol;
  if (g_spare_X.cap >= n && g_spare_X.p) d->X.swap(g_spare_X);
  d->X.need(n);
  // borrowed for this call only (the reference deallocates xx_carr right after predict,
  // OH_GridCompMod.F90:383): the data is in HBM when this returns.
  if (n && !is_device_ptr(data)) {
    Booster *b = g.speculate ? g_last_booster : nullptr;
    if (b && !(b->loaded && ncol <= b->host.num_feature)) b = nullptr;
    create_pipelined(d.get(), data, b);
  } else {
    if (n) {
      CU(cudaMemcpyAsync(d->X.p, data, n * sizeof(float), cudaMemcpyDeviceToDevice, g.stream));
      CU(cudaStreamSynchronize(g.stream));
    }
    seal(d.get());
  }
  *out = d.release();
  API_END
}

int XGDMatrixFree(DMatrixHandle handle) {
  API_BEGIN
  DMatrix *d = D(handle);
  if (d->spec_ready) drain();  // pipelined work may still reference the buffers
  d->magic = 0;
  if (d->X.cap > g_spare_X.cap) d->X.swap(g_spare_X);
  if (d->spec_host.cap > g_spare_pin.cap) d->spec_host.swap(g_spare_pin);
  if (d->spec_dev.cap > g_spare_spec.cap) d->spec_dev.swap(g_spare_spec);
  delete d;
  API_END
}

int XGDMatrixNumRow(DMatrixHandle handle, bst_ulong *out) {
  API_BEGIN
  *out = D(handle)->nrow;
  API_END
}

int XGDMatrixNumCol(DMatrixHandle handle, bst_ulong *out) {
  API_BEGIN
  *out = D(handle)->ncol;
  API_END
}

int XGBoosterPredict(BoosterHandle handle, DMatrixHandle dmat, int option_mask, unsigned ntree_limit, int training,
                     bst_ulong *out_len, const float **out_result) {
  API_BEGIN
  (void)training;  // inference only; the reference passes 0 (OH_GridCompMod.F90:235)
  Booster *b = B(handle);
  DMatrix *d = D(dmat);
  if (!out_len || !out_result) throw Error("XGBoosterPredict: NULL output argument");
  if (!b->loaded) throw Error("Booster has no model: call XGBoosterLoadModel first");
  const size_t n = (size_t)d->nrow * ((option_mask & 2) ? trees_used(b, ntree_limit) : 1);
  ensure_device();
  g_last_booster = b;
  if (d->spec_ready && d->spec_booster == b && d->spec_version == b->version && (option_mask & ~1) == 0 &&
      trees_used(b, ntree_limit) == b->host.trees.size()) {
    // already predicted while the matrix was crossing PCIe: adopt the pinned result buffer
    CU(cudaStreamSynchronize(g.d2h_stream));
    b->h_result.swap(d->spec_host);
    d->spec_ready = false;
    *out_len = n;
    *out_result = b->h_result.p;
    return 0;
  }
  float *dv = b->d_result.need(n);
  float *hv = b->h_result.need(n);
  predict_into(b, d, option_mask, ntree_limit, nullptr, dv);
  CU(cudaMemcpyAsync(hv, dv, n * sizeof(float), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  *out_len = n;
  *out_result = hv;
  API_END
}

// Dense container written by XGDMatrixSaveBinary: "QCDM" u32 version, u64 nrow, u64 ncol, f32 missing, data.
int XGDMatrixSaveBinary(DMatrixHandle handle, const char *fname, int silent) {
  API_BEGIN
  (void)silent;
  DMatrix *d = D(handle);
  const size_t n = (size_t)d->nrow * d->ncol;
  std::vector<float> h(n);
  if (n) {
    ensure_device();
    CU(cudaMemcpy(h.data(), d->X.p, n * 4, cudaMemcpyDeviceToHost));
  }
  FILE *fp = fopen(fname, "wb");
  if (!fp) throw Error(std::string("Opening ") + fname + " for writing failed");
  uint32_t ver = 1;
  uint64_t nr = d->nrow, nc = d->ncol;
  bool ok = fwrite("QCDM", 1, 4, fp) == 4 && fwrite(&ver, 4, 1, fp) == 1 && fwrite(&nr, 8, 1, fp) == 1 &&
            fwrite(&nc, 8, 1, fp) == 1 && fwrite(&d->missing, 4, 1, fp) == 1 && fwrite(h.data(), 4, n, fp) == n;
  fclose(fp);
  if (!ok) throw Error(std::string("Short write on ") + fname);
  API_END
}

int XGDMatrixCreateFromFile(const char *fname, int silent, DMatrixHandle *out) {
  (void)silent;
  std::vector<float> h;
  uint64_t nr = 0, nc = 0;
  float missing = NAN;
  try {
    FILE *fp = fopen(fname, "rb");
    if (!fp) throw Error(std::string("Opening ") + fname + " failed");
    char magic[4];
    uint32_t ver = 0;
    bool ok = fread(magic, 1, 4, fp) == 4 && !memcmp(magic, "QCDM", 4) && fread(&ver, 4, 1, fp) == 1 && ver == 1 &&
              fread(&nr, 8, 1, fp) == 1 && fread(&nc, 8, 1, fp) == 1 && fread(&missing, 4, 1, fp) == 1;
    if (ok) {
      h.resize((size_t)nr * nc);
      ok = fread(h.data(), 4, h.size(), fp) == h.size();
    }
    fclose(fp);
    if (!ok) throw Error(std::string(fname) + ": not a libqcoh dense matrix file (XGBoost's own binary DMatrix, libsvm and csv inputs are not supported)");
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
  return XGDMatrixCreateFromMat(h.data(), nr, nc, missing, out);
}

// =====================================================================================
// (2) qcoh_* extension
// =====================================================================================
const char *qcoh_version(void) { return "libqcoh 0.1 (sm_100a)"; }

int qcoh_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

int qcoh_set_device(int device) {
  API_BEGIN
  if (g.ready && g.device != device) throw Error("qcoh_set_device: the library is already bound to device " + std::to_string(g.device));
  requested_device = device;
  ensure_device();
  API_END
}

int qcoh_host_alloc(size_t bytes, void **out) {
  API_BEGIN
  ensure_device();
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  API_END
}
int qcoh_host_free(void *p) {
  API_BEGIN
  CU(cudaFreeHost(p));
  API_END
}
int qcoh_device_alloc(size_t bytes, void **out) {
  API_BEGIN
  ensure_device();
  CU(cudaMalloc(out, bytes ? bytes : 1));
  API_END
}
int qcoh_device_free(void *p) {
  API_BEGIN
  CU(cudaFree(p));
  API_END
}
int qcoh_memcpy_h2d(void *dst, const void *src, size_t bytes) {
  API_BEGIN
  ensure_device();
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  API_END
}
int qcoh_memcpy_d2h(void *dst, const void *src, size_t bytes) {
  API_BEGIN
  ensure_device();
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  API_END
}
int qcoh_device_synchronize(void) {
  API_BEGIN
  ensure_device();
  CU(cudaStreamSynchronize(g.stream));
  API_END
}
int qcoh_timer_start(void) {
  API_BEGIN
  ensure_device();
  CU(cudaEventRecord(g.ev0, g.stream));
  API_END
}
int qcoh_timer_stop(float *ms) {
  API_BEGIN
  ensure_device();
  CU(cudaEventRecord(g.ev1, g.stream));
  CU(cudaEventSynchronize(g.ev1));
  CU(cudaEventElapsedTime(ms, g.ev0, g.ev1));
  API_END
}
int qcoh_flush_l2(void) {
  API_BEGIN
  ensure_device();
  if (!g.flush) {
    g.flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    CU(cudaMalloc(&g.flush, g.flush_bytes));
  }
  CU(cudaMemsetAsync(g.flush, 0, g.flush_bytes, g.stream));
  API_END
}

int qcoh_set_param(const char *name, const char *value) {
  API_BEGIN
  if (!name || !value) throw Error("qcoh_set_param: NULL argument");
  const int v = atoi(value);
  std::string n(name);
  if (n == "variant") g.tun.variant = v;
  else if (n == "ilp") g.tun.ilp = v;
  else if (n == "block") g.tun.block = v;
  else if (n == "top_levels") g.tun.top_levels = v;
  else if (n == "park") g.tun.park = v;
  else if (n == "minb") g.tun.minb = v;
  else if (n == "speculate") g.speculate = v;
  else if (n == "chunk_rows") g.chunk_rows = v > 0 ? ((uint64_t)v + 255) / 256 * 256 : (1ull << 21);
  else throw Error("qcoh_set_param: unknown parameter '" + n + "'");
  API_END
}

uint64_t qcoh_launch_count(void) { return launch_count(); }

int qcoh_booster_get_info(BoosterHandle handle, qcoh_booster_info *out) {
  API_BEGIN
  Booster *b = B(handle);
  if (!b->loaded) throw Error("Booster has no model");
  out->num_trees = (int32_t)b->host.trees.size();
  out->num_feature = (int32_t)b->host.num_feature;
  out->max_depth = b->flat.max_depth;
  out->num_nodes = b->flat.num_nodes();
  out->base_score = b->host.base_score;
  out->format = (int32_t)b->host.format;
  for (int i = 0; i < 3; ++i) out->version[i] = b->host.version[i];
  API_END
}

int qcoh_booster_get_flat(BoosterHandle handle, const uint32_t **nodes_xy, const uint32_t **tree_offset,
                          const int32_t **tree_depth, const int32_t **orig_id) {
  API_BEGIN
  Booster *b = B(handle);
  if (!b->loaded) throw Error("Booster has no model");
  if (nodes_xy) *nodes_xy = b->flat.nodes_xy.data();
  if (tree_offset) *tree_offset = b->flat.tree_offset.data();
  if (tree_depth) *tree_depth = b->flat.tree_depth.data();
  if (orig_id) *orig_id = b->flat.orig_id.data();
  API_END
}

int qcoh_dmatrix_create_device(bst_ulong nrow, bst_ulong ncol, float missing, DMatrixHandle *out) {
  API_BEGIN
  ensure_device();
  std::unique_ptr<DMatrix> d(new DMatrix());
  d->nrow = nrow, d->ncol = ncol, d->missing = missing;
  d->X.need((size_t)nrow * ncol);
  *out = d.release();
  API_END
}
int qcoh_dmatrix_device_ptr(DMatrixHandle handle, float **out_dev) {
  API_BEGIN
  *out_dev = D(handle)->X.p;
  API_END
}
int qcoh_dmatrix_upload(DMatrixHandle handle, const float *host_rows, bst_ulong row0, bst_ulong nrows) {
  API_BEGIN
  DMatrix *d = D(handle);
  if (row0 + nrows > d->nrow) throw Error("qcoh_dmatrix_upload: row range out of bounds");
  CU(cudaMemcpyAsync(d->X.p + (size_t)row0 * d->ncol, host_rows, (size_t)nrows * d->ncol * 4,
                     is_device_ptr(host_rows) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  d->sealed = false;
  API_END
}
int qcoh_dmatrix_seal(DMatrixHandle handle) {
  API_BEGIN
  seal(D(handle));
  API_END
}

int qcoh_booster_predict_device(BoosterHandle handle, DMatrixHandle dmat, int option_mask, unsigned ntree_limit,
                                const qcoh_epilogue *epi, float *out_dev) {
  API_BEGIN
  predict_into(B(handle), D(dmat), option_mask, ntree_limit, epi, out_dev);
  API_END
}

int qcoh_partition_columns(int64_t ncol_global, int nranks, int rank, int64_t *col0, int64_t *ncol_local) {
  API_BEGIN
  if (nranks <= 0 || rank < 0 || rank >= nranks || ncol_global < 0) throw Error("qcoh_partition_columns: bad arguments");
  const int64_t q = ncol_global / nranks, r = ncol_global % nranks;
  // the first r ranks own one extra column; contiguous ranges, no halo
  *col0 = q * rank + (rank < r ? rank : r);
  *ncol_local = q + (rank < r ? 1 : 0);
  API_END
}

}  // extern "C"

// =====================================================================================
// fused Run1
// =====================================================================================
namespace {

// JulianDay / leap_year — OH_GridCompMod.F90:1905-1971
bool is_leap(int ny) { return ny >= 0 && ((ny % 100 == 0 && ny % 400 == 0) || (ny % 4 == 0 && ny % 100 != 0)); }
int julian_day(int nymd) {
  static const int days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  const int ny = nymd / 10000, mm = (nymd % 10000) / 100;
  int ds = nymd % 100;
  for (int m = 1; m < mm; ++m) ds += (m == 2 && is_leap(ny)) ? 29 : days[m - 1];
  return ds;
}

// computeSolarZenithAngle_LocalNoon — OH_GridCompMod.F90:401-466.  Evaluated on the HOST with
// the C library's float32 sin/asin/cos/acos: that is what the compiled Fortran calls, and the
// GPU's libdevice versions differ from it in the last bit (a 1-ulp SZA change can flip a leaf).
// It is a 2-D field that only changes with the day of year, so it is cached per (jday, grid).
void noon_sza_range(int jday, const float *lat, const float *lon, int i0, int i1, float r2d, float d2r, float *out) {
  const float sindec = 0.3978f * sinf(0.9863f * ((float)jday - 80.0f) * d2r);
  const float cosdec = cosf(asinf(sindec));
  for (int i = i0; i < i1; ++i) {
    const float sinlat = sinf(lat[i]);
    const float coslat = cosf(asinf(sinlat));
    float mylon = lon[i] * r2d;
    if (mylon > 180.0f) mylon = mylon - 360.0f;
    if (mylon < -180.0f) mylon = mylon + 360.0f;
    const float tau = 12.0f + (mylon / -180.0f) * 12.0f;
    const float loct = ((tau * 15.0f) - 180.0f) * d2r + lon[i];
    float cosz = cosdec * coslat * cosf(loct) + sindec * sinlat;
    cosz = fminf(1.0f, cosz);
    cosz = fmaxf(-1.0f, cosz);
    out[i] = acosf(cosz) * r2d;
  }
}

void noon_sza(int jday, const float *lat, const float *lon, int n, float r2d, float d2r, float *out) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt > 64) nt = 64;
  if ((unsigned)n < nt * 1024) nt = 1;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    const int i0 = (int)((int64_t)n * t / nt), i1 = (int)((int64_t)n * (t + 1) / nt);
    th.emplace_back(noon_sza_range, jday, lat, lon, i0, i1, r2d, d2r, out);
  }
  for (auto &t : th) t.join();
}

struct Oh {
  uint32_t magic = kOhMagic;
  Booster *booster = nullptr;
  qcoh_oh_config cfg;
  // resident copies of host-provided inputs, one slot per input field
  static constexpr int kNumIn = 13 + 7 + 11 + 5 + 1 + 1;
  DevBuf<float> in[kNumIn];
  DevBuf<float> PL_MOD, NDWET, sums[6], OH_ML, OH, OH_boost, X, pred, sza, lat_deg, so3;
  DevBuf<int> ctl;
  DevBuf<double> diag;
  PinBuf<float> h_sza, h_lat, h_lon;
  int sza_jday = -1;
  const float *sza_lat_key = nullptr, *sza_lon_key = nullptr;
  bool oh_ml_valid = false;
};

Oh *O(qcoh_oh_handle h) {
  Oh *o = (Oh *)h;
  if (!o || o->magic != kOhMagic) throw Error("Invalid OH handle");
  return o;
}

// host pointer -> resident device copy; device pointer -> used in place
const float *resident(Oh *o, int slot, const float *p, size_t n, const char *name) {
  if (!p) throw Error(std::string("qcoh_oh_run1: input field ") + name + " is NULL");
  if (is_device_ptr(p)) return p;
  float *d = o->in[slot].need(n);
  CU(cudaMemcpyAsync(d, p, n * 4, cudaMemcpyHostToDevice, g.stream));
  return d;
}

void deliver(float *user, const float *dev, size_t n) {
  if (!user) return;
  CU(cudaMemcpyAsync(user, dev, n * 4, is_device_ptr(user) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, g.stream));
}

}  // namespace

extern "C" {

int qcoh_oh_create(BoosterHandle booster, const qcoh_oh_config *cfg, qcoh_oh_handle *out) {
  API_BEGIN
  Booster *b = B(booster);
  if (!cfg || !out) throw Error("qcoh_oh_create: NULL argument");
  if (cfg->ncol <= 0 || cfg->km <= 0) throw Error("qcoh_oh_create: ncol and km must be positive");
  if (!b->loaded) throw Error("Booster has no model: call XGBoosterLoadModel first");
  if (b->host.num_feature != 27)
    throw Error("OH_GridComp packs exactly 27 features (OH_GridCompMod.F90:228); the booster has " + std::to_string(b->host.num_feature));
  ensure_device();
  upload(b);
  std::unique_ptr<Oh> o(new Oh());
  o->booster = b, o->cfg = *cfg;
  const size_t n3 = (size_t)cfg->ncol * cfg->km;
  o->OH_ML.need(n3);
  CU(cudaMemsetAsync(o->OH_ML.p, 0, n3 * 4, g.stream));
  *out = o.release();
  API_END
}

int qcoh_oh_get_diag(qcoh_oh_handle h, const char *name, float *out) {
  API_BEGIN
  Oh *o = O(h);
  if (!name || !out) throw Error("qcoh_oh_get_diag: NULL argument");
  if (!o->oh_ml_valid) throw Error("qcoh_oh_get_diag: no boost step has run yet");
  const size_t n2 = (size_t)o->cfg.ncol, n3 = n2 * o->cfg.km;
  const std::string s(name);
  const float *src = nullptr;
  size_t n = n3;
  static const char *sum_names[6] = {"TAUCLWDN", "TAUCLIDN", "TAUCLIUP", "TAUCLWUP", "AODUP", "AODDN"};
  for (int i = 0; i < 6; ++i)
    if (s == sum_names[i]) src = o->sums[i].p;
  if (s == "PL") src = o->PL_MOD.p;
  if (s == "NDWET") src = o->NDWET.p;
  if (s == "OH_boost") src = o->OH_ML.p;
  if (s == "LAT") src = o->lat_deg.p, n = n2;
  if (s == "SZA") src = o->sza.p, n = n2;
  if (s == "stratO3") src = o->so3.p, n = n2;
  if (!src) throw Error("qcoh_oh_get_diag: unknown or unavailable field '" + s + "'");
  deliver(out, src, n);
  CU(cudaStreamSynchronize(g.stream));
  API_END
}

int qcoh_oh_free(qcoh_oh_handle h) {
  API_BEGIN
  Oh *o = O(h);
  o->magic = 0;
  delete o;
  API_END
}

int qcoh_oh_run1(qcoh_oh_handle h, const qcoh_run1_in *in, qcoh_run1_out *out) {
  API_BEGIN
  Oh *o = O(h);
  if (!in || !out) throw Error("qcoh_oh_run1: NULL argument");
  if (!out->OH) throw Error("qcoh_oh_run1: out->OH is required");
  ensure_device();
  const qcoh_oh_config &c = o->cfg;
  const int nc = c.ncol, km = c.km;
  const size_t n2 = (size_t)nc, n3 = (size_t)nc * km, ne = (size_t)nc * (km + 1);
  Run1Dev r;
  memset(&r, 0, sizeof r);
  r.ncol = nc, r.km = km;
  r.eps = c.mapl_epsilon, r.avogad = c.mapl_avogad, r.runiv = c.mapl_runiv, r.r2d = c.mapl_radians_to_degrees;
  r.ohscale = c.ohscale, r.tropp_min = c.tropp_min, r.missing = c.missing;
  r.dynamic_k = c.compute_once_per_day ? 0 : 1;  // :1561
  int s = 0;
  r.T_MOD = resident(o, s++, in->T_MOD, n3, "T_MOD");
  r.Q_MOD = resident(o, s++, in->Q_MOD, n3, "Q_MOD");
  r.PLE_MOD = resident(o, s++, in->PLE_MOD, ne, "PLE_MOD");
  r.TROPP = resident(o, s++, in->TROPP, n2, "TROPP");
  r.OH_CLIM = resident(o, s++, in->OH_CLIM, n3, "OH_CLIM");
  const bool boost = in->need_to_call_boost != 0;
  if (!boost && !o->oh_ml_valid) throw Error("qcoh_oh_run1: need_to_call_boost = 0 before any boost call (OH_ML is undefined)");
  const bool want_diag = in->AREA != nullptr;
  if (boost || want_diag) {
    r.ZLE_BST = resident(o, s++, in->ZLE_BST, ne, "ZLE_BST");
    r.CH4 = resident(o, s++, in->CH4, n3, "CH4");
  } else {
    s += 2;
  }
  if (boost) {
    // aliasing (ONLINE_INST hands the same arrays as model state and boost input) is preserved:
    // identical host pointers are uploaded once
    auto same = [&](const float *p, const float *q, const float *dq) { return p == q ? dq : nullptr; };
    const float *d;
    r.T_BST = (d = same(in->T_BST, in->T_MOD, r.T_MOD)) ? d : resident(o, s, in->T_BST, n3, "T_BST");
    ++s;
    r.Q_BST = (d = same(in->Q_BST, in->Q_MOD, r.Q_MOD)) ? d : resident(o, s, in->Q_BST, n3, "Q_BST");
    ++s;
    r.PLE_BST = (d = same(in->PLE_BST, in->PLE_MOD, r.PLE_MOD)) ? d : resident(o, s, in->PLE_BST, ne, "PLE_BST");
    ++s;
    r.TAUCLW = resident(o, s++, in->TAUCLW, n3, "TAUCLW");
    r.TAUCLI = resident(o, s++, in->TAUCLI, n3, "TAUCLI");
    r.FCLD = resident(o, s++, in->FCLD, n3, "FCLD");
    r.CO = resident(o, s++, in->CO, n3, "CO");
    for (int i = 0; i < 7; ++i) r.SCA[i] = resident(o, s++, in->SCA[i], n3, "SCACOEF");
    const float *const gases[11] = {in->NO2, in->O3, in->ISOP, in->ACET, in->C2H6, in->C3H8, in->PRPE, in->ALK4, in->MP, in->H2O2, in->CH2O};
    const float **dst[11] = {&r.NO2, &r.O3, &r.ISOP, &r.ACET, &r.C2H6, &r.C3H8, &r.PRPE, &r.ALK4, &r.MP, &r.H2O2, &r.CH2O};
    for (int i = 0; i < 11; ++i) *dst[i] = resident(o, s++, gases[i], n3, "climatological gas");
    r.GMITO3 = resident(o, s++, in->GMITO3, n2, "GMITO3");
    r.GMITTO3 = resident(o, s++, in->GMITTO3, n2, "GMITTO3");
    r.ALBUV = resident(o, s++, in->ALBUV, n2, "ALBUV");
    r.LATS = resident(o, s++, in->LATS, n2, "LATS");
    if (!in->LONS) throw Error("qcoh_oh_run1: input field LONS is NULL");
    // noon SZA (host libm, cached per day and grid)
    const int jday = julian_day(in->nymd);
    if (jday != o->sza_jday || in->LATS != o->sza_lat_key || in->LONS != o->sza_lon_key) {
      float *hl = o->h_lat.need(n2), *hn = o->h_lon.need(n2), *hs = o->h_sza.need(n2);
      CU(cudaMemcpyAsync(hl, in->LATS, n2 * 4, cudaMemcpyDefault, g.stream));
      CU(cudaMemcpyAsync(hn, in->LONS, n2 * 4, cudaMemcpyDefault, g.stream));
      CU(cudaStreamSynchronize(g.stream));
      noon_sza(jday, hl, hn, nc, c.mapl_radians_to_degrees, c.mapl_degrees_to_radians, hs);
      CU(cudaMemcpyAsync(o->sza.need(n2), hs, n2 * 4, cudaMemcpyHostToDevice, g.stream));
      o->sza_jday = jday, o->sza_lat_key = in->LATS, o->sza_lon_key = in->LONS;
    }
    r.SZA = o->sza.p;
  }
  if (want_diag) r.AREA = resident(o, Oh::kNumIn - 1, in->AREA, n2, "AREA");
  r.PL_MOD = o->PL_MOD.need(n3), r.NDWET = o->NDWET.need(n3);
  r.OH_ML = o->OH_ML.p, r.OH = o->OH.need(n3), r.OH_boost = o->OH_boost.need(n3);
  r.ctl = o->ctl.need(4);
  r.diag = o->diag.need(4);
  CU(cudaMemsetAsync(r.ctl, 0, 4 * sizeof(int), g.stream));
  CU(launch_oh_state(r, g.stream));
  out->k1 = 0;
  bool check_inf_after = false;
  if (boost) {
    int ctl[4];
    CU(cudaMemcpyAsync(ctl, r.ctl, sizeof ctl, cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    if (!r.dynamic_k && ctl[1] != 0) throw Error("OH Prediction: Minimum tropopause pressure is not low enough!");  // :288
    const int ksub = ctl[0];
    const int k1 = km - ksub + 1;  // :300
    out->k1 = k1;
    const uint64_t npred = (uint64_t)nc * ksub;
    for (int i = 0; i < 6; ++i) r.sums[i] = o->sums[i].need(n3);
    r.lat_deg = o->lat_deg.need(n2), r.so3 = o->so3.need(n2);
    CU(launch_oh_sums(r, g.stream));
    CU(cudaMemsetAsync(r.OH_ML, 0, n3 * 4, g.stream));  // self%OH_ML = 0.0 (:1559)
    if (npred && !out->X) {
      // fused: pack (:303-345) + create (:347) + predict (:356) + 10**x (:369) * OHscale (:1569) in one
      // kernel reading the SoA fields; the [N x 27] matrix is never formed
      SoaArgs a;
      const float *s3[27] = {nullptr, nullptr, r.T_BST, r.NO2, r.O3, r.CH4, r.CO, r.ISOP, r.ACET, r.C2H6, r.C3H8, r.PRPE,
                             r.ALK4, r.MP, r.H2O2, r.sums[0], r.sums[1], r.sums[2], r.sums[3], r.FCLD, r.Q_BST, nullptr,
                             nullptr, r.sums[4], r.sums[5], r.CH2O, nullptr};
      const float *s2[27] = {nullptr};
      s2[0] = r.lat_deg, s2[21] = r.so3, s2[22] = r.ALBUV, s2[26] = r.SZA;
      for (int f = 0; f < 27; ++f) a.src3[f] = s3[f], a.src2[f] = s2[f];
      a.ple = r.PLE_BST, a.ncol = nc, a.e0 = (uint64_t)(k1 - 1) * nc, a.nrow = npred, a.missing = c.missing;
      a.ntree_used = o->booster->dev.ntree, a.exp10 = 1, a.scale = c.ohscale;
      a.out = r.OH_ML + (size_t)(k1 - 1) * nc;
      a.pred = out->pred ? o->pred.need(npred) : nullptr;
      a.flags = r.ctl + 2;
      CU(launch_predict_soa(o->booster->dev, a, g.tun, g.stream));
      if (out->pred) deliver(out->pred, a.pred, npred);
      check_inf_after = true;
    } else if (npred) {
      // debug / parity path: materialise xx_carr so that it can be handed back (out->X)
      float *X = o->X.need(npred * 27);
      CU(launch_oh_pack(r, k1, X, g.stream));
      int flags = 0;
      CU(cudaMemcpyAsync(&flags, r.ctl + 2, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
      CU(cudaStreamSynchronize(g.stream));
      if ((flags & 2) && !std::isinf(c.missing)) throw Error("Check failed: valid: Input data contains `inf` or `nan`");
      PredictArgs a;
      a.X = X, a.nrow = npred, a.ncol = 27, a.missing = c.missing, a.has_missing = flags & 1, a.pred_leaf = 0;
      a.ntree_used = o->booster->dev.ntree, a.exp10 = 1, a.scale = c.ohscale;
      a.out = r.OH_ML + (size_t)(k1 - 1) * nc;
      CU(launch_predict(o->booster->dev, a, g.tun, g.stream));
      if (out->pred) {
        a.exp10 = 0, a.scale = 1.f, a.out = o->pred.need(npred);
        CU(launch_predict(o->booster->dev, a, g.tun, g.stream));
        deliver(out->pred, a.out, npred);
      }
      deliver(out->X, X, npred * 27);
    }
    o->oh_ml_valid = true;
  }
  CU(launch_oh_finalize(r, g.stream));
  if (want_diag) {
    CU(cudaMemsetAsync(r.diag, 0, 4 * sizeof(double), g.stream));
    CU(launch_oh_diag(r, g.stream));
    CU(cudaMemcpyAsync(out->diag, r.diag, 4 * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  }
  deliver(out->OH, r.OH, n3);
  deliver(out->OH_boost, r.OH_boost, n3);
  deliver(out->NDWET, r.NDWET, n3);
  int inf_flags = 0;
  if (check_inf_after) CU(cudaMemcpyAsync(&inf_flags, r.ctl + 2, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  if ((inf_flags & 2) && !std::isinf(c.missing)) {
    o->oh_ml_valid = false;
    throw Error("Check failed: valid: Input data contains `inf` or `nan`");
  }
  API_END
}

}  // extern "C"
