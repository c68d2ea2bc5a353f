// libb2f_comm.so: the gradient all-reduce of the training path over NCCL (NVLink 5 / NVSwitch), see include/b2f_comm.h.
// NCCL is bound at run time with dlopen so that neither this library nor libb2f_cuda.so carries a DT_NEEDED on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/b2f_comm.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  const char* (*GetLastError)(ncclComm_t) = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.handle) break;
    }
    if (!n.handle) return;
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(n.handle, "ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(n.handle, "ncclCommInitRank"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.handle, "ncclAllReduce"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.handle, "ncclCommDestroy"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(n.handle, "ncclGetErrorString"));
    n.GetLastError = reinterpret_cast<decltype(n.GetLastError)>(dlsym(n.handle, "ncclGetLastError"));
    n.ok = n.GetUniqueId && n.CommInitRank && n.AllReduce && n.CommDestroy && n.GetErrorString;
  });
  return n;
}

int need_nccl() {
  if (!nccl().ok) return fail(-4, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbol");
  return 0;
}

int nccl_fail(ncclResult_t r, const char* what) {
  const char* detail = nccl().GetLastError ? nccl().GetLastError(nullptr) : "";
  return fail(1000 + (int)r, "%s: %s %s", what, nccl().GetErrorString(r), detail ? detail : "");
}

}  // namespace

struct b2f_comm {
  ncclComm_t comm;
  int world, rank;
};

static_assert(sizeof(ncclUniqueId) == B2F_COMM_ID_BYTES, "NCCL unique id size");

extern "C" {

int b2f_comm_abi_version(void) { return 1; }
const char* b2f_comm_last_error(void) { return g_err; }

int b2f_comm_unique_id(void* id_out) {
  if (!id_out) return fail(-1, "comm_unique_id: NULL");
  if (int rc = need_nccl()) return rc;
  ncclUniqueId id;
  ncclResult_t r = nccl().GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

int b2f_comm_init(b2f_comm_t* comm, const void* id, int world, int rank) {
  if (!comm || !id) return fail(-1, "comm_init: NULL argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(-1, "comm_init: bad world / rank %d / %d", world, rank);
  if (int rc = need_nccl()) return rc;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclComm_t c;
  ncclResult_t r = nccl().CommInitRank(&c, world, uid, rank);
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
  *comm = new b2f_comm{c, world, rank};
  return 0;
}

int b2f_comm_world(b2f_comm_t comm, int* world, int* rank) {
  if (!comm) return fail(-1, "comm_world: NULL communicator");
  if (world) *world = comm->world;
  if (rank) *rank = comm->rank;
  return 0;
}

static int allreduce(b2f_comm_t comm, void* buf, size_t count, ncclDataType_t dt, b2f_comm_stream_t stream) {
  if (!comm) return fail(-1, "comm_allreduce: NULL communicator");
  if (!count) return 0;
  if (!buf) return fail(-1, "comm_allreduce: NULL buffer");
  if (comm->world == 1) return 0;
  ncclResult_t r = nccl().AllReduce(buf, buf, count, dt, ncclSum, comm->comm, reinterpret_cast<cudaStream_t>(stream));
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
  return 0;
}

int b2f_comm_allreduce_sum_f32(b2f_comm_t comm, float* buf, size_t count, b2f_comm_stream_t stream) {
  return allreduce(comm, buf, count, ncclFloat32, stream);
}
int b2f_comm_allreduce_sum_f64(b2f_comm_t comm, double* buf, size_t count, b2f_comm_stream_t stream) {
  return allreduce(comm, buf, count, ncclFloat64, stream);
}

int b2f_comm_destroy(b2f_comm_t comm) {
  if (!comm) return 0;
  ncclResult_t r = nccl().CommDestroy(comm->comm);
  delete comm;
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommDestroy");
  return 0;
}

}  // extern "C"
