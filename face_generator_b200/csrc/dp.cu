// Data parallelism: one process per GPU, one NCCL communicator, one in-place sum-allreduce of the
// flat gradient (+8 tail scalars) per optimizer step.  The reference has no
// multi-GPU path at all; this is new functionality behind the same train-step call.
#include <nccl.h>

#include <cstring>

#include "fg_internal.h"

#define FG_NCCL(call)                                                                          \
  do {                                                                                         \
    ncclResult_t r__ = (call);                                                                 \
    if (r__ != ncclSuccess) {                                                                  \
      fg_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__));     \
      return FG_ERR_NCCL;                                                                      \
    }                                                                                          \
  } while (0)

int net_allreduce(fg_ctx* c, float* buf, int64_t n) {
  if (c->world <= 1) return FG_OK;
  ScopedTimer t(c, "nccl.allreduce");
  FG_NCCL(ncclAllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
  return FG_OK;
}

extern "C" {
int fg_dp_unique_id(void* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  if (!out128) return FG_ERR_INVALID;
  ncclUniqueId id;
  FG_NCCL(ncclGetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return FG_OK;
}
int fg_dp_init(fg_ctx* c, const void* id128, int nranks, int rank) {
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) {
    fg_set_error("fg_dp_init: bad arguments");
    return FG_ERR_INVALID;
  }
  FG_CUDA(cudaSetDevice(c->device));
  if (c->nccl_comm) {
    ncclCommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
  }
  c->world = 1;
  c->rank = 0;
  if (nranks == 1) return FG_OK;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  FG_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
  c->nccl_comm = comm;
  c->world = nranks;
  c->rank = rank;
  return FG_OK;
}
int fg_dp_broadcast_params(fg_ctx* c) {
  if (!c) return FG_ERR_INVALID;
  if (c->world <= 1) return FG_OK;
  FG_CUDA(cudaSetDevice(c->device));
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  FG_NCCL(ncclGroupStart());
  FG_NCCL(ncclBroadcast(c->PG, c->PG, c->gl.total, ncclFloat, 0, comm, c->stream));
  FG_NCCL(ncclBroadcast(c->PD, c->PD, c->dl.total, ncclFloat, 0, comm, c->stream));
  FG_NCCL(ncclBroadcast(c->mG, c->mG, c->gl.total, ncclFloat, 0, comm, c->stream));
  FG_NCCL(ncclBroadcast(c->vG, c->vG, c->gl.total, ncclFloat, 0, comm, c->stream));
  FG_NCCL(ncclBroadcast(c->mD, c->mD, c->dl.total, ncclFloat, 0, comm, c->stream));
  FG_NCCL(ncclBroadcast(c->vD, c->vD, c->dl.total, ncclFloat, 0, comm, c->stream));
  FG_NCCL(ncclBroadcast(c->bnG, c->bnG, 768, ncclFloat, 0, comm, c->stream));
  FG_NCCL(ncclGroupEnd());
  FG_CUDA(cudaStreamSynchronize(c->stream));
  c->G_packed = c->D_packed = false;
  return FG_OK;
}
int fg_dp_world(fg_ctx* c) { return c ? c->world : 0; }
}
