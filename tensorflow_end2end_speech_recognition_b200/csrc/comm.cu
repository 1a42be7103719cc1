// Gradient exchange of the data-parallel step behind the C ABI: NCCL all-reduce (mean) over NVLink 5 /
// NVSwitch, one communicator rank per GPU.
//
// Replaces average_gradients (utils/training/multi_gpu.py:13-48), which concatenates every tower's gradient
// on /cpu:0 and takes the mean, as called at examples/librispeech/training/train_ctc.py:143.  Here every rank
// holds one tower; the buckets (one per BLSTM layer, issued in BPTT completion order by the host mirror) are
// reduced in place with ncclAvg inside one NCCL group call.
//
// libnccl is bound at run time (dlopen of the soname torch already has in the process, or the system copy),
// so libb2asr.so itself loads on a box without NCCL; the entry points then fail with B2_ERR_UNSUPPORTED.
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>

namespace b2 {

// the slice of nccl.h this file needs (NCCL >= 2.10: ncclAvg)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclFloat = 7, kNcclAvg = 4 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommCount)(ncclComm_t, int*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return nullptr;
#define B2_SYM(field, sym)                                    \
  *(void**)(&api.field) = dlsym(h, sym);                      \
  if (!api.field) { dlclose(h); return nullptr; }
  B2_SYM(GetUniqueId, "ncclGetUniqueId")
  B2_SYM(CommInitRank, "ncclCommInitRank")
  B2_SYM(CommInitAll, "ncclCommInitAll")
  B2_SYM(CommDestroy, "ncclCommDestroy")
  B2_SYM(CommCount, "ncclCommCount")
  B2_SYM(AllReduce, "ncclAllReduce")
  B2_SYM(GroupStart, "ncclGroupStart")
  B2_SYM(GroupEnd, "ncclGroupEnd")
  B2_SYM(GetErrorString, "ncclGetErrorString")
  B2_SYM(GetVersion, "ncclGetVersion")
#undef B2_SYM
  api.handle = h;
  return &api;
}

#define B2_NCCL(api, call)                                                                       \
  do {                                                                                           \
    int r__ = (call);                                                                            \
    if (r__ != kNcclSuccess) {                                                                   \
      b2::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, (api)->GetErrorString(r__));   \
      return B2_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

}  // namespace b2

using namespace b2;

extern "C" int b2_comm_available(void) {
  NcclApi* a = nccl_api();
  if (!a) return 0;
  int v = 0;
  return a->GetVersion(&v) == kNcclSuccess ? v : 0;
}

extern "C" int b2_comm_get_unique_id(void* id_out) {
  B2_CHECK_ARG(id_out, "b2_comm_get_unique_id: null pointer");
  NcclApi* a = nccl_api();
  if (!a) { set_error("b2_comm: libnccl.so.2 not found"); return B2_ERR_UNSUPPORTED; }
  ncclUniqueId id;
  B2_NCCL(a, a->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return B2_OK;
}

extern "C" int b2_comm_init_rank(b2_comm_t* comm_out, int nranks, const void* id, int rank) {
  B2_CHECK_ARG(comm_out && id && nranks >= 1 && rank >= 0 && rank < nranks,
               "b2_comm_init_rank: bad argument (nranks %d rank %d)", nranks, rank);
  NcclApi* a = nccl_api();
  if (!a) { set_error("b2_comm: libnccl.so.2 not found"); return B2_ERR_UNSUPPORTED; }
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclComm_t c = nullptr;
  B2_NCCL(a, a->CommInitRank(&c, nranks, uid, rank));
  *comm_out = (b2_comm_t)c;
  return B2_OK;
}

extern "C" int b2_comm_init_all(b2_comm_t* comms_out, int ndev, const int* devices) {
  B2_CHECK_ARG(comms_out && ndev >= 1, "b2_comm_init_all: bad argument");
  NcclApi* a = nccl_api();
  if (!a) { set_error("b2_comm: libnccl.so.2 not found"); return B2_ERR_UNSUPPORTED; }
  B2_NCCL(a, a->CommInitAll((ncclComm_t*)comms_out, ndev, devices));
  return B2_OK;
}

extern "C" int b2_comm_size(b2_comm_t comm) {
  NcclApi* a = nccl_api();
  int n = 0;
  if (!a || !comm || a->CommCount((ncclComm_t)comm, &n) != kNcclSuccess) return -1;
  return n;
}

extern "C" int b2_comm_destroy(b2_comm_t comm) {
  NcclApi* a = nccl_api();
  if (!a) { set_error("b2_comm: libnccl.so.2 not found"); return B2_ERR_UNSUPPORTED; }
  if (comm) B2_NCCL(a, a->CommDestroy((ncclComm_t)comm));
  return B2_OK;
}

extern "C" int b2_allreduce_mean(b2_comm_t comm, float* const* buckets, const int64_t* sizes, int n_buckets,
                                 b2_stream_t stream_) {
  B2_CHECK_ARG(comm && buckets && sizes && n_buckets >= 1, "b2_allreduce_mean: bad argument");
  NcclApi* a = nccl_api();
  if (!a) { set_error("b2_comm: libnccl.so.2 not found"); return B2_ERR_UNSUPPORTED; }
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_NCCL(a, a->GroupStart());
  for (int k = 0; k < n_buckets; ++k) {
    if (sizes[k] <= 0) continue;
    const int r = a->AllReduce(buckets[k], buckets[k], (size_t)sizes[k], kNcclFloat, kNcclAvg, (ncclComm_t)comm,
                               stream);
    if (r != kNcclSuccess) {
      a->GroupEnd();
      set_error("b2_allreduce_mean: ncclAllReduce(bucket %d) -> %s", k, a->GetErrorString(r));
      return B2_ERR_CUDA;
    }
  }
  B2_NCCL(a, a->GroupEnd());
  return B2_OK;
}

// all local towers of ONE process (the reference's in-graph multi-GPU flow, train_ctc.py:82-147): comm k reduces
// bucket k, each on its own device/stream, inside one group call
extern "C" int b2_allreduce_mean_local(const b2_comm_t* comms, float* const* buffers, int64_t n, int ndev,
                                       const b2_stream_t* streams) {
  B2_CHECK_ARG(comms && buffers && streams && n > 0 && ndev >= 1, "b2_allreduce_mean_local: bad argument");
  NcclApi* a = nccl_api();
  if (!a) { set_error("b2_comm: libnccl.so.2 not found"); return B2_ERR_UNSUPPORTED; }
  B2_NCCL(a, a->GroupStart());
  for (int k = 0; k < ndev; ++k) {
    const int r = a->AllReduce(buffers[k], buffers[k], (size_t)n, kNcclFloat, kNcclAvg, (ncclComm_t)comms[k],
                               (cudaStream_t)streams[k]);
    if (r != kNcclSuccess) {
      a->GroupEnd();
      set_error("b2_allreduce_mean_local: ncclAllReduce(device %d) -> %s", k, a->GetErrorString(r));
      return B2_ERR_CUDA;
    }
  }
  B2_NCCL(a, a->GroupEnd());
  return B2_OK;
}
