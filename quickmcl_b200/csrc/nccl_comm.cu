// nccl_comm.cu — a built-in qmcl_comm (include/qmcl.h) on NCCL over NVLink.
//
// The sharded filter's exchanges are small (8 B to a few MB) and there are
// about fifteen of them per update, so their cost is launch latency, not
// bandwidth.  Calling NCCL from inside the library on the filter's own stream
// keeps each one at an enqueue (no interpreter, no stream hand-over).  NCCL is
// resolved with dlopen at first use: the library still loads on a machine
// without NCCL, and in a torch process the already loaded libnccl.so.2 is
// reused.  Hosts that prefer their own transport fill qmcl_comm themselves
// (quickmcl_b200/sharded.py: TorchComm).

#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

namespace qmcl {
namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

template <typename T> bool sym(void *h, const char *name, T *out) {
  *out = reinterpret_cast<T>(dlsym(h, name));
  return *out != nullptr;
}

const NcclApi &nccl() {
  std::call_once(g_nccl_once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
    void *h = g_nccl.handle;
    g_nccl.ok = sym(h, "ncclGetUniqueId", &g_nccl.GetUniqueId) &&
                sym(h, "ncclCommInitRank", &g_nccl.CommInitRank) &&
                sym(h, "ncclCommDestroy", &g_nccl.CommDestroy) &&
                sym(h, "ncclAllGather", &g_nccl.AllGather) &&
                sym(h, "ncclAllReduce", &g_nccl.AllReduce) && sym(h, "ncclSend", &g_nccl.Send) &&
                sym(h, "ncclRecv", &g_nccl.Recv) && sym(h, "ncclGroupStart", &g_nccl.GroupStart) &&
                sym(h, "ncclGroupEnd", &g_nccl.GroupEnd) &&
                sym(h, "ncclGetErrorString", &g_nccl.GetErrorString);
  });
  return g_nccl;
}

struct NcclCtx {
  ncclComm_t comm;
  int rank, size;
};

int fail(ncclResult_t r, const char *what) {
  set_error("%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
  return 1;
}

int cb_allgather(void *ctx, void *dev_buf, size_t bytes_per_rank, void *stream) {
  NcclCtx *c = static_cast<NcclCtx *>(ctx);
  // in place: this rank's part already sits at rank * bytes_per_rank
  const char *mine = static_cast<const char *>(dev_buf) + static_cast<size_t>(c->rank) * bytes_per_rank;
  const ncclResult_t r = nccl().AllGather(mine, dev_buf, bytes_per_rank, ncclInt8, c->comm,
                                          static_cast<cudaStream_t>(stream));
  return r == ncclSuccess ? 0 : fail(r, "ncclAllGather");
}

int cb_allreduce(void *ctx, void *dev_buf, size_t count, int dtype, int op, void *stream) {
  NcclCtx *c = static_cast<NcclCtx *>(ctx);
  const ncclDataType_t dt = dtype == QMCL_DT_I32 ? ncclInt32 : dtype == QMCL_DT_I64 ? ncclInt64 : ncclFloat64;
  const ncclRedOp_t ro = op == QMCL_OP_SUM ? ncclSum : op == QMCL_OP_MIN ? ncclMin : ncclMax;
  const ncclResult_t r =
      nccl().AllReduce(dev_buf, dev_buf, count, dt, ro, c->comm, static_cast<cudaStream_t>(stream));
  return r == ncclSuccess ? 0 : fail(r, "ncclAllReduce");
}

int cb_alltoallv(void *ctx, const void *dev_send, const uint64_t *send_bytes,
                 const uint64_t *send_offsets, void *dev_recv, const uint64_t *recv_bytes,
                 const uint64_t *recv_offsets, void *stream) {
  NcclCtx *c = static_cast<NcclCtx *>(ctx);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // What stays on this rank (the bulk of a rebalancing exchange) is a plain device copy,
  // not a send to oneself through NCCL; peers with nothing to exchange are skipped, and a
  // call without any remote traffic issues no NCCL operation at all.
  const int me = c->rank;
  if (send_bytes[me] && recv_bytes[me]) {
    if (send_bytes[me] != recv_bytes[me]) return fail(ncclInvalidArgument, "alltoallv: self piece");
    if (cudaMemcpyAsync(static_cast<char *>(dev_recv) + recv_offsets[me],
                        static_cast<const char *>(dev_send) + send_offsets[me], send_bytes[me],
                        cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return fail(ncclUnhandledCudaError, "alltoallv: self copy");
  }
  bool remote = false;
  for (int p = 0; p < c->size; ++p) remote |= p != me && (send_bytes[p] || recv_bytes[p]);
  if (!remote) return 0;
  ncclResult_t r = nccl().GroupStart();
  if (r != ncclSuccess) return fail(r, "ncclGroupStart");
  for (int p = 0; p < c->size; ++p) {
    if (p == me) continue;
    if (send_bytes[p]) {
      r = nccl().Send(static_cast<const char *>(dev_send) + send_offsets[p], send_bytes[p], ncclInt8, p,
                      c->comm, s);
      if (r != ncclSuccess) break;
    }
    if (recv_bytes[p]) {
      r = nccl().Recv(static_cast<char *>(dev_recv) + recv_offsets[p], recv_bytes[p], ncclInt8, p,
                      c->comm, s);
      if (r != ncclSuccess) break;
    }
  }
  const ncclResult_t e = nccl().GroupEnd();
  if (r != ncclSuccess) return fail(r, "ncclSend/ncclRecv");
  return e == ncclSuccess ? 0 : fail(e, "ncclGroupEnd");
}

}  // namespace
}  // namespace qmcl

using namespace qmcl;

extern "C" {

qmcl_status qmcl_nccl_unique_id(uint8_t out[128]) {
  if (!out) return QMCL_ERR_INVALID_ARG;
  if (!nccl().ok) {
    set_error("NCCL (libnccl.so.2) could not be loaded");
    return QMCL_ERR_COMM;
  }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  const ncclResult_t r = nccl().GetUniqueId(&id);
  if (r != ncclSuccess) {
    fail(r, "ncclGetUniqueId");
    return QMCL_ERR_COMM;
  }
  std::memcpy(out, &id, sizeof(id));
  return QMCL_OK;
}

qmcl_status qmcl_nccl_comm_create(const uint8_t id_bytes[128], int rank, int size, int device,
                                  qmcl_comm *out) {
  if (!id_bytes || !out || size < 1 || rank < 0 || rank >= size) return QMCL_ERR_INVALID_ARG;
  if (!nccl().ok) {
    set_error("NCCL (libnccl.so.2) could not be loaded");
    return QMCL_ERR_COMM;
  }
  if (device >= 0) QMCL_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, sizeof(id));
  NcclCtx *c = new NcclCtx{nullptr, rank, size};
  const ncclResult_t r = nccl().CommInitRank(&c->comm, size, id, rank);
  if (r != ncclSuccess) {
    delete c;
    fail(r, "ncclCommInitRank");
    return QMCL_ERR_COMM;
  }
  out->rank = rank;
  out->size = size;
  out->ctx = c;
  out->allgather = cb_allgather;
  out->allreduce = cb_allreduce;
  out->alltoallv = cb_alltoallv;
  return QMCL_OK;
}

qmcl_status qmcl_nccl_comm_destroy(qmcl_comm *comm) {
  if (!comm || !comm->ctx) return QMCL_ERR_INVALID_ARG;
  NcclCtx *c = static_cast<NcclCtx *>(comm->ctx);
  if (nccl().ok && c->comm) nccl().CommDestroy(c->comm);
  delete c;
  comm->ctx = nullptr;
  return QMCL_OK;
}

}  // extern "C"
