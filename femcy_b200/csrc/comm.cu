// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
// The reference is single-device (no collectives anywhere), so everything here is new:
//   * halo exchange of the ghost entries of the CG direction vector before each SpMV
//     (pack kernel -> grouped ncclSend/ncclRecv -> unpack kernel),
//   * CG reductions as ncclAllGather of per-rank partials which every rank then folds in rank
//     order (bitwise identical scalars on all ranks, no divergence of alpha/beta).
// NCCL is resolved at run time with dlopen from the path the host passes in (the library torch
// already loaded), so the .so has no link-time NCCL dependency and single-GPU use needs no NCCL.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "ctx.cuh"

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int ncclResult_t;
enum { NCCL_DOUBLE = 8 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static int load_nccl(NcclApi& api, const char* path, std::string& err) {
  const char* p = (path && path[0]) ? path : "libnccl.so.2";
  api.handle = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
  if (!api.handle) { err = std::string("dlopen NCCL failed: ") + dlerror(); return 1; }
#define LOADSYM(field, name)                                                   \
  *(void**)(&api.field) = dlsym(api.handle, name);                             \
  if (!api.field) { err = std::string("missing NCCL symbol ") + name; return 1; }
  LOADSYM(GetUniqueId, "ncclGetUniqueId");
  LOADSYM(CommInitRank, "ncclCommInitRank");
  LOADSYM(CommDestroy, "ncclCommDestroy");
  LOADSYM(AllGather, "ncclAllGather");
  LOADSYM(Send, "ncclSend");
  LOADSYM(Recv, "ncclRecv");
  LOADSYM(GroupStart, "ncclGroupStart");
  LOADSYM(GroupEnd, "ncclGroupEnd");
  LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
  return 0;
}

struct CommState {
  NcclApi api;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // halo plan
  int npeers = 0;
  std::vector<int> peers;
  std::vector<int64_t> send_ptr, recv_ptr;  // in nodes
  int32_t* d_send_nodes = nullptr;
  int32_t* d_recv_nodes = nullptr;
  double* sendbuf = nullptr;
  double* recvbuf = nullptr;
  int64_t n_send = 0, n_recv = 0;
  // peer-memory path
  bool p2p = false;
  P2PView pv;
  unsigned long long* window = nullptr;      // my window (cudaMalloc, exported with cudaIpc)
  void* opened[2 * FEMCY_MAX_RANKS] = {nullptr};
  unsigned char* bflag = nullptr;            // [nn_own] 1 = node has copies on other ranks
  int32_t* push_ptr = nullptr;               // [nn_own+1] CSR over owned nodes -> push entries
  int32_t* push_peer = nullptr;              // [n_push] destination rank
  int32_t* push_ridx = nullptr;              // [n_push] node index in the destination's numbering
  int32_t* bnodes = nullptr;                 // [n_bnodes] the owned nodes with bflag = 1 (compact list)
  int64_t n_bnodes = 0;
  int4* bpush = nullptr;                     // [n_bnodes] {node, first peer, first remote index, further push entries}: one load per boundary node
  int32_t* slice_order = nullptr;            // [nslice] SpMV slice order: slices without ghost columns first
  unsigned char* slice_ghost = nullptr;      // [nslice] 1 = the slice reads a ghost column
};

#define NCK(call)                                                                         \
  do {                                                                                    \
    ncclResult_t _r = (call);                                                             \
    if (_r != 0) return femcy_fail_msg(ctx, std::string(#call) + ": " + cs->api.GetErrorString(_r)); \
  } while (0)

int femcy_comm_size(femcy_ctx* ctx) { return ctx->comm ? ctx->comm->nranks : 1; }
int femcy_comm_rank(femcy_ctx* ctx) { return ctx->comm ? ctx->comm->rank : 0; }

void femcy_comm_free(femcy_ctx* ctx) {
  CommState* cs = ctx->comm;
  if (!cs) return;
  femcy_drop_graph(ctx);
  femcy_free(&cs->d_send_nodes); femcy_free(&cs->d_recv_nodes); femcy_free(&cs->sendbuf); femcy_free(&cs->recvbuf);
  for (int i = 0; i < 2 * FEMCY_MAX_RANKS; ++i)
    if (cs->opened[i]) cudaIpcCloseMemHandle(cs->opened[i]);
  femcy_free(&cs->window); femcy_free(&cs->bflag); femcy_free(&cs->push_ptr); femcy_free(&cs->push_peer); femcy_free(&cs->push_ridx); femcy_free(&cs->bnodes); femcy_free(&cs->bpush);
  femcy_free(&cs->slice_order); femcy_free(&cs->slice_ghost);
  if (cs->comm && cs->api.CommDestroy) cs->api.CommDestroy(cs->comm);
  delete cs;
  ctx->comm = nullptr;
}

extern "C" int femcy_comm_unique_id(const char* nccl_library_path, void* id_out) {
  NcclApi api;
  std::string err;
  if (load_nccl(api, nccl_library_path, err)) return 1;
  ncclUniqueId_t id;
  if (api.GetUniqueId(&id) != 0) return 2;
  memcpy(id_out, &id, sizeof id);
  return 0;
}

extern "C" int femcy_comm_init(femcy_ctx* ctx, int rank, int nranks, const void* unique_id, const char* nccl_library_path) {
  cudaSetDevice(ctx->device);
  femcy_comm_free(ctx);
  if (nranks > 8) return femcy_fail_msg(ctx, "at most 8 ranks (one NVSwitch box)");
  CommState* cs = new CommState();
  ctx->comm = cs;
  cs->rank = rank; cs->nranks = nranks;
  if (nranks == 1) return 0;
  std::string err;
  if (load_nccl(cs->api, nccl_library_path, err)) return femcy_fail_msg(ctx, err);
  ncclUniqueId_t id;
  memcpy(&id, unique_id, sizeof id);
  NCK(cs->api.CommInitRank(&cs->comm, nranks, id, rank));
  return 0;
}

extern "C" int femcy_set_halo(femcy_ctx* ctx, int npeers, const int32_t* peer_ranks, const int64_t* send_ptr,
                              const int32_t* send_nodes, const int64_t* recv_ptr, const int32_t* recv_nodes) {
  cudaSetDevice(ctx->device);
  CommState* cs = ctx->comm;
  if (!cs) return femcy_fail_msg(ctx, "comm_init first");
  cs->npeers = npeers;
  cs->peers.assign(peer_ranks, peer_ranks + npeers);
  cs->send_ptr.assign(send_ptr, send_ptr + npeers + 1);
  cs->recv_ptr.assign(recv_ptr, recv_ptr + npeers + 1);
  cs->n_send = send_ptr[npeers];
  cs->n_recv = recv_ptr[npeers];
  int dm = ctx->dm;
  if (femcy_alloc(ctx, &cs->d_send_nodes, cs->n_send) || femcy_alloc(ctx, &cs->d_recv_nodes, cs->n_recv) ||
      femcy_alloc(ctx, &cs->sendbuf, cs->n_send * dm) || femcy_alloc(ctx, &cs->recvbuf, cs->n_recv * dm))
    return 1;
  if (cs->n_send) CK(cudaMemcpyAsync(cs->d_send_nodes, send_nodes, (size_t)cs->n_send * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (cs->n_recv) CK(cudaMemcpyAsync(cs->d_recv_nodes, recv_nodes, (size_t)cs->n_recv * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

__global__ void k_pack(const double* __restrict__ v, const int32_t* __restrict__ idx, int64_t n, int dm, double* __restrict__ buf) {
  int64_t tot = n * dm;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = t / dm; int c = (int)(t - k * dm);
    buf[t] = v[(int64_t)idx[k] * dm + c];
  }
}
__global__ void k_unpack(double* __restrict__ v, const int32_t* __restrict__ idx, int64_t n, int dm, const double* __restrict__ buf) {
  int64_t tot = n * dm;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = t / dm; int c = (int)(t - k * dm);
    v[(int64_t)idx[k] * dm + c] = buf[t];
  }
}

int femcy_comm_halo(femcy_ctx* ctx, double* v) {
  CommState* cs = ctx->comm;
  if (!cs || cs->nranks == 1) return 0;
  int dm = ctx->dm;
  cudaStream_t st = ctx->stream;
  if (cs->n_send) {
    int g = (int)ceil_div64(cs->n_send * dm, 256); if (g > 1184) g = 1184;
    k_pack<<<g, 256, 0, st>>>(v, cs->d_send_nodes, cs->n_send, dm, cs->sendbuf);
    CK_LAUNCH();
  }
  NCK(cs->api.GroupStart());
  for (int p = 0; p < cs->npeers; ++p) {
    int64_t ns = cs->send_ptr[p + 1] - cs->send_ptr[p], nr = cs->recv_ptr[p + 1] - cs->recv_ptr[p];
    if (ns) NCK(cs->api.Send(cs->sendbuf + cs->send_ptr[p] * dm, (size_t)ns * dm, NCCL_DOUBLE, cs->peers[p], cs->comm, st));
    if (nr) NCK(cs->api.Recv(cs->recvbuf + cs->recv_ptr[p] * dm, (size_t)nr * dm, NCCL_DOUBLE, cs->peers[p], cs->comm, st));
  }
  NCK(cs->api.GroupEnd());
  ctx->launches++;
  if (cs->n_recv) {
    int g = (int)ceil_div64(cs->n_recv * dm, 256); if (g > 1184) g = 1184;
    k_unpack<<<g, 256, 0, st>>>(v, cs->d_recv_nodes, cs->n_recv, dm, cs->recvbuf);
    CK_LAUNCH();
  }
  return 0;
}

extern "C" int femcy_halo_exchange(femcy_ctx* ctx, int which) {
  cudaSetDevice(ctx->device);
  if (which < 0 || which >= FEMCY_VEC_COUNT || !ctx->vec[which]) return femcy_fail_msg(ctx, "bad vector selector");
  return femcy_comm_halo(ctx, ctx->vec[which]);
}

// all-gather `nvals` doubles per rank from scal[16..] into scal[24 + rank*nvals ..]
int femcy_cg_comm_allgather(femcy_ctx* ctx, int nvals) {
  CommState* cs = ctx->comm;
  if (!cs || cs->nranks == 1) return 0;
  NCK(cs->api.AllGather(ctx->scal + 16, ctx->scal + 24, (size_t)nvals, NCCL_DOUBLE, cs->comm, ctx->stream));
  ctx->launches++;
  return 0;
}

// ---- peer-memory (NVLink P2P) setup -------------------------------------------------------------
// export: allocate my window, return the cudaIpc handles of {window, vec[D]} (2 x 64 bytes)
extern "C" int femcy_p2p_export(femcy_ctx* ctx, void* handles_out) {
  cudaSetDevice(ctx->device);
  CommState* cs = ctx->comm;
  if (!cs) return femcy_fail_msg(ctx, "comm_init first");
  if (!ctx->vec[FEMCY_VEC_D]) return femcy_fail_msg(ctx, "state not allocated");
  if (!cs->window) {
    if (femcy_alloc(ctx, &cs->window, P2P_WINDOW_WORDS)) return 1;
    CK(cudaMemset(cs->window, 0, P2P_WINDOW_WORDS * sizeof(unsigned long long)));
  }
  cudaIpcMemHandle_t h[2];
  CK(cudaIpcGetMemHandle(&h[0], cs->window));
  CK(cudaIpcGetMemHandle(&h[1], ctx->vec[FEMCY_VEC_D]));
  memcpy(handles_out, h, sizeof h);
  return 0;
}

// import: open every peer's {window, d}; install the push plan.  remote_start[k] = index, in the
// numbering of peer k (same order as femcy_set_halo's peer list), of the first ghost node it holds
// for me; my send list to that peer maps to consecutive nodes from there.
extern "C" int femcy_p2p_import(femcy_ctx* ctx, const void* all_handles /*[nranks][2][64 B]*/, const int64_t* remote_start) {
  cudaSetDevice(ctx->device);
  CommState* cs = ctx->comm;
  if (!cs || !cs->window) return femcy_fail_msg(ctx, "p2p_export first");
  const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
  P2PView& pv = cs->pv;
  pv.nranks = cs->nranks; pv.rank = cs->rank;
  for (int r = 0; r < cs->nranks; ++r) {
    if (r == cs->rank) {
      pv.win_of[r] = cs->window;
      pv.d_of[r] = ctx->vec[FEMCY_VEC_D];
      continue;
    }
    void *pw = nullptr, *pd = nullptr;
    CK(cudaIpcOpenMemHandle(&pw, h[2 * r + 0], cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle(&pd, h[2 * r + 1], cudaIpcMemLazyEnablePeerAccess));
    cs->opened[2 * r] = pw; cs->opened[2 * r + 1] = pd;
    pv.win_of[r] = (unsigned long long*)pw;
    pv.d_of[r] = (double*)pd;
  }
  // push plan from the halo plan: owned node -> (peer rank, remote node index)
  int64_t nown = ctx->nn_own;
  std::vector<int32_t> cnt(nown + 1, 0);
  std::vector<int32_t> send_nodes(cs->n_send);
  if (cs->n_send) CK(cudaMemcpy(send_nodes.data(), cs->d_send_nodes, (size_t)cs->n_send * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int64_t t = 0; t < cs->n_send; ++t) {
    if (send_nodes[t] < 0 || send_nodes[t] >= nown) return femcy_fail_msg(ctx, "send list holds a node this rank does not own");
    cnt[send_nodes[t] + 1]++;
  }
  for (int64_t i = 0; i < nown; ++i) cnt[i + 1] += cnt[i];
  std::vector<int32_t> ppeer(cs->n_send), pridx(cs->n_send), fill(cnt.begin(), cnt.end() - 1);
  std::vector<unsigned char> bf(nown, 0);
  for (int p = 0; p < cs->npeers; ++p)
    for (int64_t t = cs->send_ptr[p]; t < cs->send_ptr[p + 1]; ++t) {
      int32_t nd = send_nodes[t];
      int32_t o = fill[nd]++;
      ppeer[o] = cs->peers[p];
      pridx[o] = (int32_t)(remote_start[p] + (t - cs->send_ptr[p]));
      bf[nd] = 1;
    }
  if (femcy_alloc(ctx, &cs->bflag, nown) || femcy_alloc(ctx, &cs->push_ptr, nown + 1) ||
      femcy_alloc(ctx, &cs->push_peer, cs->n_send) || femcy_alloc(ctx, &cs->push_ridx, cs->n_send))
    return 1;
  CK(cudaMemcpy(cs->bflag, bf.data(), (size_t)nown, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(cs->push_ptr, cnt.data(), (size_t)(nown + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (cs->n_send) {
    CK(cudaMemcpy(cs->push_peer, ppeer.data(), (size_t)cs->n_send * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(cs->push_ridx, pridx.data(), (size_t)cs->n_send * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  std::vector<int32_t> bn;
  for (int64_t i = 0; i < nown; ++i)
    if (bf[i]) bn.push_back((int32_t)i);
  cs->n_bnodes = (int64_t)bn.size();
  if (femcy_alloc(ctx, &cs->bnodes, cs->n_bnodes)) return 1;
  if (cs->n_bnodes) CK(cudaMemcpy(cs->bnodes, bn.data(), (size_t)cs->n_bnodes * sizeof(int32_t), cudaMemcpyHostToDevice));
  {
    std::vector<int4> bp(bn.size());
    for (size_t k = 0; k < bn.size(); ++k) {
      const int32_t nd = bn[k], o = cnt[nd];
      bp[k] = make_int4(nd, ppeer[o], pridx[o], cnt[nd + 1] - o - 1);
    }
    if (femcy_alloc(ctx, &cs->bpush, cs->n_bnodes)) return 1;
    if (cs->n_bnodes) CK(cudaMemcpy(cs->bpush, bp.data(), bp.size() * sizeof(int4), cudaMemcpyHostToDevice));
  }
  // SpMV slice order: slices whose rows reference no ghost column first
  {
    BsellPattern& P = ctx->P;
    if (!P.colidx) return femcy_fail_msg(ctx, "build_pattern before p2p_import");
    std::vector<int32_t> sp(P.nslice + 1), ci(P.nslots);
    CK(cudaMemcpy(sp.data(), P.slice_ptr, (size_t)(P.nslice + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ci.data(), P.colidx, (size_t)P.nslots * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<unsigned char> gh(P.nslice, 0);
    for (int64_t sl = 0; sl < P.nslice; ++sl)
      for (int32_t t = sp[sl]; t < sp[sl + 1]; ++t)
        if (ci[t] >= nown) { gh[sl] = 1; break; }
    std::vector<int32_t> order;
    order.reserve(P.nslice);
    for (int64_t sl = 0; sl < P.nslice; ++sl) if (!gh[sl]) order.push_back((int32_t)sl);
    for (int64_t sl = 0; sl < P.nslice; ++sl) if (gh[sl]) order.push_back((int32_t)sl);
    if (femcy_alloc(ctx, &cs->slice_order, P.nslice) || femcy_alloc(ctx, &cs->slice_ghost, P.nslice)) return 1;
    CK(cudaMemcpy(cs->slice_order, order.data(), (size_t)P.nslice * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(cs->slice_ghost, gh.data(), (size_t)P.nslice, cudaMemcpyHostToDevice));
  }
  cs->p2p = true;
  femcy_drop_graph(ctx);
  return 0;
}

bool femcy_p2p_view(femcy_ctx* ctx, P2PView* pv, const unsigned char** bflag, const int32_t** push_ptr,
                    const int32_t** push_peer, const int32_t** push_ridx, const int32_t** bnodes, int64_t* n_bnodes,
                    const int32_t** slice_order, const unsigned char** slice_ghost, const int4** bpush) {
  CommState* cs = ctx->comm;
  if (!cs || !cs->p2p || cs->nranks == 1 || ctx->opt.no_p2p) return false;
  *pv = cs->pv; *bflag = cs->bflag; *push_ptr = cs->push_ptr; *push_peer = cs->push_peer; *push_ridx = cs->push_ridx;
  *bnodes = cs->bnodes; *n_bnodes = cs->n_bnodes;
  if (bpush) *bpush = cs->bpush;
  *slice_order = cs->slice_order; *slice_ghost = cs->slice_ghost;
  return true;
}
