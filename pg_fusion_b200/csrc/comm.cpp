// Multi-GPU plumbing: one NCCL communicator per context (one process per GPU), loaded lazily so that a
// single-GPU host never needs libnccl.  Collectives are enqueued on the context's compute stream.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the functions are resolved with dlsym

#include <cstring>

#include "context.hpp"

namespace pgf {
namespace {

struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    // a process that already carries an NCCL (e.g. the one PyTorch bundles) keeps using that one: same soname
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return;
    auto sym = [&](const char* n) { return dlsym(h, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Send && api.Recv && api.GroupStart &&
             api.GroupEnd && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}

pgf_status nccl_fail(pgf_ctx* ctx, ncclResult_t r, const char* what) {
  NcclApi* n = nccl();
  return ctx->fail(PGF_ERR_COMM, "NCCL error %d (%s) in %s", int(r), n ? n->GetErrorString(r) : "?", what);
}

}  // namespace

pgf_status comm_all_gather(pgf_ctx* ctx, const void* send, void* recv, uint64_t bytes) {
  NcclApi* n = nccl();
  if (!n || !ctx->nccl_comm) return ctx->fail(PGF_ERR_STATE, "the context has no communicator (pgf_comm_init)");
  const ncclResult_t r = n->AllGather(send, recv, bytes, ncclChar, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->compute_stream);
  return r == ncclSuccess ? PGF_OK : nccl_fail(ctx, r, "ncclAllGather");
}

// all-to-all with per-peer byte counts: grouped ncclSend / ncclRecv (NVSwitch gives every pair full bandwidth)
pgf_status comm_all_to_all_v(pgf_ctx* ctx, const void* send, const uint64_t* send_off, const uint64_t* send_bytes, void* recv,
                             const uint64_t* recv_off, const uint64_t* recv_bytes) {
  NcclApi* n = nccl();
  if (!n || !ctx->nccl_comm) return ctx->fail(PGF_ERR_STATE, "the context has no communicator (pgf_comm_init)");
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
  ncclResult_t r = n->GroupStart();
  for (int p = 0; p < ctx->comm_world && r == ncclSuccess; ++p) {
    if (send_bytes[p]) r = n->Send(static_cast<const uint8_t*>(send) + send_off[p], send_bytes[p], ncclChar, p, comm, ctx->compute_stream);
    if (r == ncclSuccess && recv_bytes[p]) r = n->Recv(static_cast<uint8_t*>(recv) + recv_off[p], recv_bytes[p], ncclChar, p, comm, ctx->compute_stream);
  }
  const ncclResult_t e = n->GroupEnd();
  if (r == ncclSuccess) r = e;
  return r == ncclSuccess ? PGF_OK : nccl_fail(ctx, r, "grouped ncclSend / ncclRecv");
}

void comm_release(pgf_ctx* ctx) {
  if (ctx->nccl_comm) {
    if (NcclApi* n = nccl()) n->CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
  }
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
}

}  // namespace pgf

using namespace pgf;

extern "C" {

pgf_status pgf_comm_unique_id(uint8_t id_out[PGF_COMM_ID_BYTES]) {
  static_assert(PGF_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
  if (!id_out) return PGF_ERR_INVALID_ARGUMENT;
  NcclApi* n = nccl();
  if (!n) return PGF_ERR_NO_DEVICE;
  ncclUniqueId id;
  if (n->GetUniqueId(&id) != ncclSuccess) return PGF_ERR_CUDA;
  std::memcpy(id_out, id.internal, PGF_COMM_ID_BYTES);
  return PGF_OK;
}

pgf_status pgf_comm_init(pgf_ctx* ctx, const uint8_t id[PGF_COMM_ID_BYTES], int32_t rank, int32_t world) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (ctx->nccl_comm) return ctx->fail(PGF_ERR_STATE, "the context already has a communicator");
  NcclApi* n = nccl();
  if (!n) return ctx->fail(PGF_ERR_NO_DEVICE, "libnccl.so.2 could not be loaded: %s", dlerror());
  CU(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  std::memcpy(uid.internal, id, PGF_COMM_ID_BYTES);
  ncclComm_t comm = nullptr;
  const ncclResult_t r = n->CommInitRank(&comm, world, uid, rank);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclCommInitRank");
  ctx->nccl_comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return PGF_OK;
}

pgf_status pgf_comm_destroy(pgf_ctx* ctx) {
  if (!ctx) return PGF_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->compute_stream);
  comm_release(ctx);
  return PGF_OK;
}

pgf_status pgf_comm_info(pgf_ctx* ctx, int32_t* rank_out, int32_t* world_out) {
  if (!ctx || !rank_out || !world_out) return PGF_ERR_INVALID_ARGUMENT;
  *rank_out = ctx->comm_rank;
  *world_out = ctx->comm_world;
  return PGF_OK;
}

pgf_status pgf_comm_all_gather(pgf_ctx* ctx, const void* dev_send, void* dev_recv, uint64_t bytes) {
  if (!ctx || !dev_send || !dev_recv) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  std::lock_guard<std::mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  if (ctx->comm_world == 1) {
    CU(ctx, cudaMemcpyAsync(dev_recv, dev_send, bytes, cudaMemcpyDeviceToDevice, ctx->compute_stream));
    return PGF_OK;
  }
  return comm_all_gather(ctx, dev_send, dev_recv, bytes);
}

pgf_status pgf_comm_all_gather_host(pgf_ctx* ctx, const void* host_send, void* host_recv, uint64_t bytes) {
  if (!ctx || !host_send || !host_recv) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  std::lock_guard<std::mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  const uint64_t world = uint64_t(ctx->comm_world);
  if (world == 1) {
    std::memcpy(host_recv, host_send, bytes);
    return PGF_OK;
  }
  const size_t need = bytes * (world + 1);
  if (ctx->d_xchg_cap < need) {
    CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
    if (ctx->d_xchg) cudaFree(ctx->d_xchg);
    ctx->d_xchg = nullptr;
    ctx->d_xchg_cap = 0;
    void* p = nullptr;
    if (cudaMalloc(&p, need) != cudaSuccess) { cudaGetLastError(); return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "exchange scratch"); }
    ctx->d_xchg = static_cast<uint8_t*>(p);
    ctx->d_xchg_cap = need;
  }
  CU(ctx, cudaMemcpyAsync(ctx->d_xchg, host_send, bytes, cudaMemcpyHostToDevice, ctx->compute_stream));
  PGF_TRY(comm_all_gather(ctx, ctx->d_xchg, ctx->d_xchg + bytes, bytes));
  CU(ctx, cudaMemcpyAsync(host_recv, ctx->d_xchg + bytes, bytes * world, cudaMemcpyDeviceToHost, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  return PGF_OK;
}

pgf_status pgf_pipeline_run_sharded(pgf_ctx* ctx, const pgf_pipeline* plan, uint64_t max_groups, pgf_result** result_out) {
  if (!ctx || !plan || !result_out) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  return pipeline_run_sharded(ctx, plan, max_groups, result_out);
}

pgf_status pgf_join_table_exchange(pgf_ctx* ctx, uint64_t table_or_rows, uint32_t mode, uint64_t* out_handle, uint64_t* nvlink_bytes_out) {
  if (!ctx || !out_handle) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  return join_exchange(ctx, table_or_rows, mode, out_handle, nvlink_bytes_out);
}

}  // extern "C"
