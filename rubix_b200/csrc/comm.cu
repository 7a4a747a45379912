// The one exchange step of the path behind the C ABI: the per-rank partial cubes are summed over
// NVLink with NCCL -- rubix's `jnp.sum(ifu_cubes, axis=0)` over its device axis (rubix/core/ifu.py:324-333).
//   rbx_reduce_cube          MUSE-size cubes (9.3 MB): one ncclReduce onto the rank that applies PSF + LSF
//   rbx_reduce_scatter_cube  large-FOV cubes: one ncclReduceScatter of the slab-major partial cube
//                            (rbx_assign_build_cube_slabs), every rank keeps its summed wavelength slab + halo
//   rbx_allreduce_cube       every rank gets the whole cube
// NCCL is bound at run time (dlopen "libnccl.so.2"): a host process that already loaded NCCL (torch, jaxlib)
// shares that copy, and the library itself has no link-time NCCL dependency.  All calls are asynchronous on
// the caller's stream.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

using namespace rbx;

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  std::string error;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  NcclApi &a = g_nccl;
  // the copy the process already holds (torch / jaxlib bundle their own libnccl.so.2), else the system one
  a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!a.handle) {
    a.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
    return;
  }
  bool ok = true;
  auto sym = [&](const char *name) {
    void *p = dlsym(a.handle, name);
    if (!p) { ok = false; a.error = std::string("libnccl.so.2 lacks ") + name; }
    return p;
  };
  a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
  a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
  a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
  a.Reduce = (decltype(a.Reduce))sym("ncclReduce");
  a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
  a.ReduceScatter = (decltype(a.ReduceScatter))sym("ncclReduceScatter");
  a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
  if (!ok) a.handle = nullptr;
}

int nccl_api(NcclApi **out) {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.handle) {
    set_error("rbx_comm: " + g_nccl.error);
    return RBX_ERR_NCCL;
  }
  *out = &g_nccl;
  return RBX_OK;
}

}  // namespace

struct rbx_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

#define RBX_NCCL_OK(api, expr)                                                              \
  do {                                                                                      \
    ncclResult_t _r = (expr);                                                               \
    if (_r != ncclSuccess) {                                                                \
      rbx::set_error(std::string(#expr) + ": " + (api)->GetErrorString(_r));                \
      return RBX_ERR_NCCL;                                                                  \
    }                                                                                       \
  } while (0)

extern "C" int rbx_comm_unique_id(void *id) {
  RBX_REQUIRE(id, "rbx_comm_unique_id: null id");
  static_assert(sizeof(ncclUniqueId) == RBX_COMM_ID_BYTES, "ncclUniqueId size");
  NcclApi *a;
  int rc = nccl_api(&a);
  if (rc != RBX_OK) return rc;
  RBX_NCCL_OK(a, a->GetUniqueId(reinterpret_cast<ncclUniqueId *>(id)));
  return RBX_OK;
}

extern "C" int rbx_comm_init(rbx_comm **comm, const void *id, int rank, int world) {
  RBX_REQUIRE(comm && id && world >= 1 && rank >= 0 && rank < world, "rbx_comm_init: bad argument");
  NcclApi *a;
  int rc = nccl_api(&a);
  if (rc != RBX_OK) return rc;
  rbx_comm *c = new rbx_comm;
  c->rank = rank;
  c->world = world;
  RBX_CUDA_OK(cudaGetDevice(&c->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = a->CommInitRank(&c->comm, world, uid, rank);
  if (r != ncclSuccess) {
    set_error(std::string("ncclCommInitRank: ") + a->GetErrorString(r));
    delete c;
    return RBX_ERR_NCCL;
  }
  *comm = c;
  return RBX_OK;
}

extern "C" int rbx_comm_destroy(rbx_comm *comm) {
  if (!comm) return RBX_OK;
  NcclApi *a;
  int rc = nccl_api(&a);
  if (rc != RBX_OK) return rc;
  if (comm->comm) a->CommDestroy(comm->comm);
  delete comm;
  return RBX_OK;
}

extern "C" int rbx_comm_info(const rbx_comm *comm, int *rank, int *world, int *nccl_version) {
  RBX_REQUIRE(comm, "rbx_comm_info: null comm");
  if (rank) *rank = comm->rank;
  if (world) *world = comm->world;
  if (nccl_version) {
    NcclApi *a;
    int rc = nccl_api(&a);
    if (rc != RBX_OK) return rc;
    RBX_NCCL_OK(a, a->GetVersion(nccl_version));
  }
  return RBX_OK;
}

extern "C" int rbx_reduce_cube(rbx_comm *comm, const float *d_send, float *d_recv, int64_t count, int root, void *stream) {
  RBX_REQUIRE(comm && d_send && count >= 0 && root >= 0 && root < comm->world, "rbx_reduce_cube: bad argument");
  RBX_REQUIRE(d_recv || comm->rank != root, "rbx_reduce_cube: the root needs a receive buffer");
  NcclApi *a;
  int rc = nccl_api(&a);
  if (rc != RBX_OK) return rc;
  RBX_NCCL_OK(a, a->Reduce(d_send, d_recv, (size_t)count, ncclFloat32, ncclSum, root, comm->comm, (cudaStream_t)stream));
  return RBX_OK;
}

extern "C" int rbx_allreduce_cube(rbx_comm *comm, const float *d_send, float *d_recv, int64_t count, void *stream) {
  RBX_REQUIRE(comm && d_send && d_recv && count >= 0, "rbx_allreduce_cube: bad argument");
  NcclApi *a;
  int rc = nccl_api(&a);
  if (rc != RBX_OK) return rc;
  RBX_NCCL_OK(a, a->AllReduce(d_send, d_recv, (size_t)count, ncclFloat32, ncclSum, comm->comm, (cudaStream_t)stream));
  return RBX_OK;
}

extern "C" int rbx_reduce_scatter_cube(rbx_comm *comm, const float *d_send, float *d_recv, int64_t recv_count,
                                       void *stream) {
  RBX_REQUIRE(comm && d_send && d_recv && recv_count >= 0, "rbx_reduce_scatter_cube: bad argument");
  NcclApi *a;
  int rc = nccl_api(&a);
  if (rc != RBX_OK) return rc;
  RBX_NCCL_OK(a, a->ReduceScatter(d_send, d_recv, (size_t)recv_count, ncclFloat32, ncclSum, comm->comm,
                                  (cudaStream_t)stream));
  return RBX_OK;
}

extern "C" int rbx_allgather_cube(rbx_comm *comm, const float *d_send, float *d_recv, int64_t send_count, void *stream) {
  RBX_REQUIRE(comm && d_send && d_recv && send_count >= 0, "rbx_allgather_cube: bad argument");
  NcclApi *a;
  int rc = nccl_api(&a);
  if (rc != RBX_OK) return rc;
  RBX_NCCL_OK(a, a->AllGather(d_send, d_recv, (size_t)send_count, ncclFloat32, comm->comm, (cudaStream_t)stream));
  return RBX_OK;
}

// float64 sums (the inertia moments of a sharded galaxy, rbx_rotate_moments)
extern "C" int rbx_allreduce_f64(rbx_comm *comm, const double *d_send, double *d_recv, int64_t count, void *stream) {
  RBX_REQUIRE(comm && d_send && d_recv && count >= 0, "rbx_allreduce_f64: bad argument");
  NcclApi *a;
  int rc = nccl_api(&a);
  if (rc != RBX_OK) return rc;
  RBX_NCCL_OK(a, a->AllReduce(d_send, d_recv, (size_t)count, ncclFloat64, ncclSum, comm->comm, (cudaStream_t)stream));
  return RBX_OK;
}
