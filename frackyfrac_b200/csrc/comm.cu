// NCCL plumbing of libfrcfrc_cuda (one process per GPU).
//
// The only exchange step of the path is the all-gather of the sample-sharded
// compact embedding (presence bit columns + row sums; weighted: fp32 operand panels +
// denominators) over NVLink.  NCCL is bound at run time with dlopen so that the single-GPU
// library has no link-time dependency on it: inside a Python process torch has already loaded
// libnccl.so.2 and the same copy is picked up; elsewhere FRC_NCCL_LIB names the library.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "frc_internal.h"

namespace frc {

namespace {

struct NcclId { char internal[128]; };
typedef void* NcclComm;
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(NcclComm*, int, NcclId, int);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_destroy)(NcclComm);
typedef int (*fn_group)(void);
typedef const char* (*fn_errstr)(int);

struct Api {
  void* handle = nullptr;
  fn_get_id get_id = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  fn_errstr errstr = nullptr;
  std::string error;
};

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("FRC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) { a.error = "libnccl.so.2 not found (set FRC_NCCL_LIB)"; return; }
    a.get_id = reinterpret_cast<fn_get_id>(dlsym(a.handle, "ncclGetUniqueId"));
    a.init_rank = reinterpret_cast<fn_init_rank>(dlsym(a.handle, "ncclCommInitRank"));
    a.all_gather = reinterpret_cast<fn_all_gather>(dlsym(a.handle, "ncclAllGather"));
    a.all_reduce = reinterpret_cast<fn_all_reduce>(dlsym(a.handle, "ncclAllReduce"));
    a.destroy = reinterpret_cast<fn_destroy>(dlsym(a.handle, "ncclCommDestroy"));
    a.group_start = reinterpret_cast<fn_group>(dlsym(a.handle, "ncclGroupStart"));
    a.group_end = reinterpret_cast<fn_group>(dlsym(a.handle, "ncclGroupEnd"));
    a.errstr = reinterpret_cast<fn_errstr>(dlsym(a.handle, "ncclGetErrorString"));
    if (!a.get_id || !a.init_rank || !a.all_gather || !a.all_reduce || !a.destroy || !a.group_start || !a.group_end)
      a.error = "libnccl lacks a required symbol";
  });
  return a;
}

std::string nccl_error(const char* what, int rc) {
  Api& a = api();
  return std::string(what) + ": " + (a.errstr ? a.errstr(rc) : "NCCL error") + " (" + std::to_string(rc) + ")";
}

constexpr int kNcclUint8 = 1;

}  // namespace

struct Comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
  char* scratch = nullptr;  // device: small host<->host exchanges and the barrier
};
constexpr size_t kScratchBytes = 64 * 1024;

bool comm_unique_id(char* id128, std::string* err) {
  Api& a = api();
  if (!a.error.empty()) { *err = a.error; return false; }
  NcclId id;
  int rc = a.get_id(&id);
  if (rc) { *err = nccl_error("ncclGetUniqueId", rc); return false; }
  memcpy(id128, id.internal, 128);
  return true;
}

Comm* comm_create(const char* id128, int rank, int world, std::string* err) {
  Api& a = api();
  if (!a.error.empty()) { *err = a.error; return nullptr; }
  NcclId id;
  memcpy(id.internal, id128, 128);
  Comm* c = new Comm();
  c->rank = rank; c->world = world;
  int rc = a.init_rank(&c->comm, world, id, rank);
  if (rc) { *err = nccl_error("ncclCommInitRank", rc); delete c; return nullptr; }
  if (cudaMalloc(reinterpret_cast<void**>(&c->scratch), kScratchBytes) != cudaSuccess) {
    cudaGetLastError();
    *err = "cudaMalloc of the communicator scratch failed";
    a.destroy(c->comm);
    delete c;
    return nullptr;
  }
  return c;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->comm) api().destroy(c->comm);
  if (c->scratch) cudaFree(c->scratch);
  delete c;
}

int comm_rank(const Comm* c) { return c->rank; }
int comm_world(const Comm* c) { return c->world; }

bool comm_all_gather_inplace(Comm* c, void* const* bufs, const size_t* bytes_per_rank, int n, cudaStream_t s,
                             std::string* err) {
  Api& a = api();
  int rc = a.group_start();
  for (int k = 0; k < n && !rc; ++k) {
    char* base = static_cast<char*>(bufs[k]);
    rc = a.all_gather(base + static_cast<size_t>(c->rank) * bytes_per_rank[k], base, bytes_per_rank[k], kNcclUint8,
                      c->comm, s);
  }
  int rc2 = a.group_end();
  if (rc || rc2) { *err = nccl_error("ncclAllGather", rc ? rc : rc2); return false; }
  return true;
}

// Host-side all-gather of a few bytes per rank (IPC handles): through the device scratch, blocking.
bool comm_all_gather_host(Comm* c, const void* mine, size_t bytes, void* all, cudaStream_t s, std::string* err) {
  Api& a = api();
  if (bytes * c->world > kScratchBytes) { *err = "comm_all_gather_host: message too large"; return false; }
  if (cudaMemcpyAsync(c->scratch + static_cast<size_t>(c->rank) * bytes, mine, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) {
    *err = "comm_all_gather_host: H2D failed"; cudaGetLastError(); return false;
  }
  int rc = a.all_gather(c->scratch + static_cast<size_t>(c->rank) * bytes, c->scratch, bytes, kNcclUint8, c->comm, s);
  if (rc) { *err = nccl_error("ncclAllGather", rc); return false; }
  if (cudaMemcpyAsync(all, c->scratch, bytes * c->world, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    *err = "comm_all_gather_host: D2H failed"; cudaGetLastError(); return false;
  }
  return true;
}

// Stream-ordered barrier over the ranks (a 4-byte all-reduce).
bool comm_barrier(Comm* c, cudaStream_t s, std::string* err) {
  Api& a = api();
  constexpr int kNcclInt32 = 2, kNcclSum = 0;
  int rc = a.all_reduce(c->scratch + kScratchBytes - 256, c->scratch + kScratchBytes - 128, 1, kNcclInt32, kNcclSum, c->comm, s);
  if (rc) { *err = nccl_error("ncclAllReduce", rc); return false; }
  return true;
}

}  // namespace frc
