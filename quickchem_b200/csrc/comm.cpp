// comm.cpp — NCCL plumbing for the one collective of the path: the all-reduce of the build-defined
// diagnostic's float64 partial sums (mass-weighted mean OH, CH4 lifetime).  The data
// path itself has no exchange step.  libnccl is dlopen'ed on first use, so libqcoh.so has no link-time
// dependency on it (a single-GPU host, or a Fortran host that reduces with MPI, never loads it).
// One process per GPU: rank r calls qcoh_comm_init(nranks, r, id) with the 128-byte id that rank 0 got from
// qcoh_comm_get_unique_id and handed to the others (MPI_Bcast in a MAPL host; torch.distributed in bench.py).
#include <dlfcn.h>

#include "context.hpp"

using namespace qcoh;

namespace {

// the part of nccl.h this file needs (NCCL 2.x ABI)
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;     // ncclSuccess == 0
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
// the dlopen'ed entry points are process-wide; the communicator belongs to the calling thread's rank
struct Nccl : NcclApi {
  ncclComm_t comm = nullptr;
  int nranks = 0, rank = -1;
  DevBuf<double> buf;
};
thread_local Nccl nccl;

void load_nccl() {
  if (nccl.lib) return;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names)
    if ((nccl.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
  if (!nccl.lib) throw Error(std::string("libnccl.so.2 not found (") + dlerror() + "); needed only for qcoh_comm_*");
  auto sym = [&](const char *s) {
    void *p = dlsym(nccl.lib, s);
    if (!p) throw Error(std::string("libnccl: missing symbol ") + s);
    return p;
  };
  nccl.GetUniqueId = (decltype(nccl.GetUniqueId))sym("ncclGetUniqueId");
  nccl.CommInitRank = (decltype(nccl.CommInitRank))sym("ncclCommInitRank");
  nccl.AllReduce = (decltype(nccl.AllReduce))sym("ncclAllReduce");
  nccl.CommDestroy = (decltype(nccl.CommDestroy))sym("ncclCommDestroy");
  nccl.GetErrorString = (decltype(nccl.GetErrorString))sym("ncclGetErrorString");
}

void nccl_check(ncclResult_t r, const char *what) {
  if (r != 0) throw Error(std::string("NCCL error in ") + what + ": " + (nccl.GetErrorString ? nccl.GetErrorString(r) : "?"));
}

}  // namespace

extern "C" {

int qcoh_comm_get_unique_id(char id[128]) {
  API_BEGIN
  if (!id) throw Error("qcoh_comm_get_unique_id: id is NULL");
  load_nccl();
  ncclUniqueId u;
  nccl_check(nccl.GetUniqueId(&u), "ncclGetUniqueId");
  memcpy(id, u.internal, 128);
  API_END
}

int qcoh_comm_init(int nranks, int rank, const char id[128]) {
  API_BEGIN
  if (!id || nranks < 1 || rank < 0 || rank >= nranks) throw Error("qcoh_comm_init: bad arguments");
  if (nccl.comm) throw Error("qcoh_comm_init: a communicator already exists (qcoh_comm_destroy first)");
  ensure_device();
  load_nccl();
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  nccl_check(nccl.CommInitRank(&nccl.comm, nranks, u, rank), "ncclCommInitRank");
  nccl.nranks = nranks, nccl.rank = rank;
  API_END
}

// In-place sum over all ranks of n float64 values held in HOST memory (n is 4 for the diagnostic): staged
// through a device buffer, reduced by ncclAllReduce on the library stream over NVLink / NVSwitch.
int qcoh_comm_allreduce_sum_f64(double *values, int n) {
  API_BEGIN
  if (!values || n <= 0) throw Error("qcoh_comm_allreduce_sum_f64: bad arguments");
  if (!nccl.comm) throw Error("qcoh_comm_allreduce_sum_f64: call qcoh_comm_init first");
  double *d = nccl.buf.need((size_t)n);
  CU(cudaMemcpyAsync(d, values, sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
  nccl_check(nccl.AllReduce(d, d, (size_t)n, kNcclFloat64, kNcclSum, nccl.comm, g.stream), "ncclAllReduce");
  CU(cudaMemcpyAsync(values, d, sizeof(double) * n, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  API_END
}

int qcoh_comm_destroy(void) {
  API_BEGIN
  if (nccl.comm) {
    nccl_check(nccl.CommDestroy(nccl.comm), "ncclCommDestroy");
    nccl.comm = nullptr;
  }
  API_END
}

}  // extern "C"
