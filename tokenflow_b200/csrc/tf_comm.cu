// tf_comm — the one collective of the path: all-gather of keyframe tensors along the pivotal-sample axis
// (keyframe K/V, pivot unit rows and extended-attention outputs, so that every GPU holds all
// K keyframes), as a C-ABI call on the caller's CUDA stream.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has loaded — PyTorch bundles
// 2.28.9 — or of the system one), so the library keeps loading on a box without NCCL or a GPU; the few
// prototypes used are declared here.  The rendezvous (distributing the 128-byte unique id) is the caller's
// business: any channel works (torch.distributed broadcast, a file, MPI).
// ncclAllGather over NVLink 5 / NVSwitch; capturable into a CUDA graph like any stream-ordered NCCL call.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/tokenflow_b200.h"
#include "tf_common.cuh"

namespace tf {
namespace {

typedef struct ncclComm* NcclComm;
struct NcclUniqueId { char internal[128]; };
typedef int NcclResult;      // 0 = ncclSuccess
constexpr int kNcclInt8 = 0; // ncclInt8 / ncclChar

struct NcclApi {
  NcclResult (*GetUniqueId)(NcclUniqueId*);
  NcclResult (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
  NcclResult (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
  NcclResult (*CommDestroy)(NcclComm);
  const char* (*GetErrorString)(NcclResult);
  NcclResult (*GetVersion)(int*);
  bool ok = false;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  const char* names[] = {getenv("TF_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // the copy the process already uses, if any
    if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return;
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(h, "ncclAllGather"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  g_nccl.GetVersion = reinterpret_cast<decltype(g_nccl.GetVersion)>(dlsym(h, "ncclGetVersion"));
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.CommDestroy;
}

int need_nccl(const char* who) {
  std::call_once(g_nccl_once, load_nccl);
  if (g_nccl.ok) return TF_OK;
  set_last_error("%s: NCCL (libnccl.so.2) could not be loaded; set TF_NCCL_LIB to its path", who);
  return TF_ERR_UNSUPPORTED;
}

int check_nccl(NcclResult r, const char* what) {
  if (r == 0) return TF_OK;
  set_last_error("%s: NCCL error %d (%s)", what, r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return TF_ERR_DRIVER;
}

}  // namespace
}  // namespace tf

using namespace tf;

extern "C" {

int tf_comm_nccl_version(void) {
  if (need_nccl("tf_comm_nccl_version")) return 0;
  int v = 0;
  if (!g_nccl.GetVersion || g_nccl.GetVersion(&v) != 0) return 0;
  return v;
}

int tf_comm_unique_id(void* id_out) {
  if (!id_out) { set_last_error("tf_comm_unique_id: NULL output"); return TF_ERR_INVALID_ARGUMENT; }
  if (int e = need_nccl("tf_comm_unique_id")) return e;
  NcclUniqueId id;
  if (int e = check_nccl(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return e;
  memcpy(id_out, id.internal, TF_COMM_ID_BYTES);
  return TF_OK;
}

int tf_comm_init(const void* id, int nranks, int rank, tf_comm_t* comm_out) {
  if (!id || !comm_out || nranks <= 0 || rank < 0 || rank >= nranks) {
    set_last_error("tf_comm_init: bad arguments (nranks=%d rank=%d)", nranks, rank);
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (int e = need_nccl("tf_comm_init")) return e;
  NcclUniqueId uid;
  memcpy(uid.internal, id, TF_COMM_ID_BYTES);
  NcclComm comm = nullptr;
  if (int e = check_nccl(g_nccl.CommInitRank(&comm, nranks, uid, rank), "ncclCommInitRank")) return e;
  *comm_out = comm;
  return TF_OK;
}

int tf_allgather(tf_comm_t comm, const void* send, void* recv, int64_t bytes_per_rank, tf_stream_t stream) {
  if (!comm || bytes_per_rank < 0 || (bytes_per_rank > 0 && (!send || !recv))) {
    set_last_error("tf_allgather: bad arguments");
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (bytes_per_rank == 0) return TF_OK;
  if (int e = need_nccl("tf_allgather")) return e;
  return check_nccl(g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, kNcclInt8, static_cast<NcclComm>(comm),
                                     static_cast<cudaStream_t>(stream)),
                    "ncclAllGather");
}

int tf_comm_destroy(tf_comm_t comm) {
  if (!comm) return TF_OK;
  if (int e = need_nccl("tf_comm_destroy")) return e;
  return check_nccl(g_nccl.CommDestroy(static_cast<NcclComm>(comm)), "ncclCommDestroy");
}

}  // extern "C"
