// Hot path rows a6 + a7: Jacobi-preconditioned conjugate gradients on the node-block SELL-32 matrix.
//
// Reference: ConjugateGradientSolver_rowMajor       /root/reference/conjugateGradientSolver.py:8-127
//   M_init :48   compute_Ad :53   r_d_init :60   rmax :67   compute_rMr :74
//   update_x/r/d :81/:86/:91   dot_product :96   solve :103-127
// The reference launches 8 kernels and reads 4 scalars back to the host per iteration.  Kernels here (cg_kernels.cuh):
//   k_cg_stream      (default: one GPU and the NVLink peer-memory path) ONE persistent cooperative kernel runs `check_every`
//                    whole iterations; the matrix stream is staged through shared memory by the TMA engine
//                    (cp.async.bulk on mbarriers), grid barriers replace kernel boundaries, partial sums and the halo
//                    travel through peer memory from inside the kernel.  FEMCY_OPT_CG_SYM: upper-half matrix stream.
//   k_cg_persistent  (option cg_kernel = 2) the same iteration with plain loads, 6 blocks per SM -- the round-1 kernel,
//                    kept as the A/B partner of the streaming kernel.
//   k_spmv_dot + k_update_xr + k_update_d   (NCCL exchange path, option cg_kernel = 1, per-kernel profile) three kernels
//                    per iteration in a CUDA graph; the last block of each kernel folds the partials.
// All reductions are two-stage with a fixed fold order => bit-reproducible (except cg_sym: fp64 atomics).  Once the
// device-side stop flag is set every later iteration is a no-op, so polling the flag from the host only every
// `check_every` iterations still stops at exactly the reference's iteration (conjugateGradientSolver.py:124).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ctx.cuh"
#include "cg_kernels.cuh"

static inline int vec_grid(int64_t n) {
  int64_t g = ceil_div64(n, 256 * 4);
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

static P2PView g_empty_view;

// persistent kernels by block size
static const void* persistent_kernel(int dm, bool sym) {
  if (sym) return dm == 1 ? (const void*)k_cg_persistent<1, 4, true> : dm == 2 ? (const void*)k_cg_persistent<2, 4, true> : (const void*)k_cg_persistent<3, 4, true>;
  return dm == 1 ? (const void*)k_cg_persistent<1, 6> : dm == 2 ? (const void*)k_cg_persistent<2, 6> : (const void*)k_cg_persistent<3, 6>;
}

// streaming kernels: (warps per block, block columns per stage, stages) -- cfg 0 is the default; 1..3 are A/B partners for dm = 3
struct StreamKernel { const void* fn; int threads; int smem; };
template <int DM, int NW, int KC, int NS>
static StreamKernel stream_kernel_of(bool sym) {
  StreamKernel k;
  k.fn = sym ? (const void*)k_cg_stream<DM, NW, KC, NS, true> : (const void*)k_cg_stream<DM, NW, KC, NS, false>;
  k.threads = NW * 32;
  k.smem = CGStreamCfg<DM, NW, KC, NS>::SMEM_BYTES;
  return k;
}
static StreamKernel stream_kernel(int dm, bool sym, int cfg) {
  if (dm == 1) return stream_kernel_of<1, 16, 8, 2>(sym);
  if (dm == 2) return stream_kernel_of<2, 16, 4, 2>(sym);
  switch (cfg) {
    case 1: return stream_kernel_of<3, 8, 4, 2>(sym);
    case 2: return stream_kernel_of<3, 8, 2, 4>(sym);
    case 3: return stream_kernel_of<3, 16, 1, 4>(sym);
    default: return stream_kernel_of<3, 16, 2, 2>(sym);
  }
}

template <int DM>
static int spmv_launch(femcy_ctx* ctx, const double* x, double* y, int cg_mode, int multi, const P2PView& pv,
                       const int32_t* slice_order, const unsigned char* slice_ghost) {
  BsellPattern& P = ctx->P;
  int grid = (int)ceil_div64(P.nslice, 8);
  if (grid < 1) grid = 1;
  if (femcy_ensure_reduction_scratch(ctx, grid)) return 1;
  if (multi == 2 && cg_mode)
    k_spmv_dot<DM, true><<<grid, 256, 0, ctx->stream>>>(P.slice_ptr, P.colidx, P.val, x, y, P.nn_own, P.nslice,
                                                        ctx->red_partials, ctx->red_ticket, ctx->scal, cg_mode, multi, pv,
                                                        slice_order, slice_ghost, P.rowof);
  else
    k_spmv_dot<DM, false><<<grid, 256, 0, ctx->stream>>>(P.slice_ptr, P.colidx, P.val, x, y, P.nn_own, P.nslice,
                                                         ctx->red_partials, ctx->red_ticket, ctx->scal, cg_mode, multi, pv,
                                                         slice_order, slice_ghost, P.rowof);
  CK_LAUNCH();
  return 0;
}

static int spmv_dispatch(femcy_ctx* ctx, const double* x, double* y, int cg_mode, int multi,
                         const P2PView& pv = g_empty_view, const int32_t* slice_order = nullptr,
                         const unsigned char* slice_ghost = nullptr) {
  switch (ctx->P.dm) {
    case 1: return spmv_launch<1>(ctx, x, y, cg_mode, multi, pv, slice_order, slice_ghost);
    case 2: return spmv_launch<2>(ctx, x, y, cg_mode, multi, pv, slice_order, slice_ghost);
    case 3: return spmv_launch<3>(ctx, x, y, cg_mode, multi, pv, slice_order, slice_ghost);
  }
  return femcy_fail_msg(ctx, "bad block size");
}

extern "C" int femcy_spmv(femcy_ctx* ctx, int x_sel, int y_sel) {
  cudaSetDevice(ctx->device);
  if (!ctx->P.val) return femcy_fail_msg(ctx, "no matrix");
  if (x_sel < 0 || x_sel >= FEMCY_VEC_COUNT || y_sel < 0 || y_sel >= FEMCY_VEC_COUNT || x_sel == y_sel)
    return femcy_fail_msg(ctx, "bad vector selector");
  if (femcy_comm_size(ctx) > 1 && femcy_comm_halo(ctx, ctx->vec[x_sel])) return 1;
  return spmv_dispatch(ctx, ctx->vec[x_sel], ctx->vec[y_sel], 0, 0);
}

// plain SpMV / SpMV with the fused d.Ad reduction (alpha = S_RMR / d.Ad) on raw device vectors: used by precond.cu
int femcy_spmv_plain(femcy_ctx* ctx, const double* x, double* y) { return spmv_dispatch(ctx, x, y, 0, 0); }
int femcy_spmv_cg(femcy_ctx* ctx, const double* x, double* y) { return spmv_dispatch(ctx, x, y, 1, 0); }
int femcy_cg_set_scalars(femcy_ctx* ctx, double eps, int fixed_iters) {
  k_set_scalars<<<1, 1, 0, ctx->stream>>>(ctx->scal, eps, fixed_iters ? 1.0 : 0.0);
  CK_LAUNCH();
  return 0;
}
int femcy_cg_solve_two_level(femcy_ctx* ctx, int b_sel, double eps, int64_t max_iter, int check_every, int fixed_iters,
                             int64_t* iters_out, double* rmax0_out, double* rmax_out);   // precond.cu

template <int DM>
static int cg_init_launch(femcy_ctx* ctx, const double* b, int multi) {
  BsellPattern& P = ctx->P;
  int64_t n = P.nn_own * DM;
  int grid = vec_grid(n);
  if (femcy_ensure_reduction_scratch(ctx, grid)) return 1;
  k_cg_init<DM><<<grid, 256, 0, ctx->stream>>>(P.diag_slot, P.val, b, ctx->vec[FEMCY_VEC_X], ctx->vec[FEMCY_VEC_R],
                                               ctx->vec[FEMCY_VEC_D], ctx->vec[FEMCY_VEC_M], ctx->vec[FEMCY_VEC_AD],
                                               P.nn_own, ctx->red_partials, ctx->red_ticket, ctx->scal, multi);
  CK_LAUNCH();
  return 0;
}

extern "C" int femcy_cg_solve(femcy_ctx* ctx, int b_sel, double eps, int64_t max_iter, int check_every, int fixed_iters,
                              int64_t* iters_out, double* rmax0_out, double* rmax_out) {
  cudaSetDevice(ctx->device);
  BsellPattern& P = ctx->P;
  if (!P.val) return femcy_fail_msg(ctx, "no matrix: build_pattern / assemble first");
  if (b_sel < 0 || b_sel >= FEMCY_VEC_COUNT || b_sel >= FEMCY_VEC_X) return femcy_fail_msg(ctx, "b must be one of dof/rhs/residual/nodal_force/du");
  if (check_every < 1) check_every = 1;
  if (ctx->opt.cg_precond == 1)      // opt-in two-level preconditioner (row f2): its own loop, same contract
    return femcy_cg_solve_two_level(ctx, b_sel, eps, max_iter, check_every, fixed_iters, iters_out, rmax0_out, rmax_out);
  cudaStream_t st = ctx->stream;
  int nranks = femcy_comm_size(ctx);
  int multi = nranks > 1 ? 1 : 0;
  P2PView pv;
  const unsigned char* bflag = nullptr;
  const int32_t *push_ptr = nullptr, *push_peer = nullptr, *push_ridx = nullptr, *bnodes = nullptr;
  int64_t n_bnodes = 0;
  const int32_t* slice_order = nullptr;
  const unsigned char* slice_ghost = nullptr;
  const int4* bpush = nullptr;
  if (multi && femcy_p2p_view(ctx, &pv, &bflag, &push_ptr, &push_peer, &push_ridx, &bnodes, &n_bnodes, &slice_order, &slice_ghost, &bpush))
    multi = 2;   // peer-memory path
  int64_t n = P.nn_own * P.dm;
  const double* b = ctx->vec[b_sel];
  double *x = ctx->vec[FEMCY_VEC_X], *r = ctx->vec[FEMCY_VEC_R], *d = ctx->vec[FEMCY_VEC_D], *M = ctx->vec[FEMCY_VEC_M],
         *Ad = ctx->vec[FEMCY_VEC_AD];
  // NCCL path: ghost part of the work vectors must not hold garbage.  (Peer-memory path: the ghost part
  // of d is written by the owners' pushes, possibly before this rank gets here -- do not touch it.)
  if (multi == 1) {
    for (int v : {FEMCY_VEC_X, FEMCY_VEC_R, FEMCY_VEC_D, FEMCY_VEC_M, FEMCY_VEC_AD})
      CK(cudaMemsetAsync(ctx->vec[v], 0, (size_t)ctx->nn * ctx->dm * sizeof(double), st));
  }
  CK(cudaMemsetAsync(ctx->red_ticket, 0, 8 * sizeof(unsigned int), st));
  k_set_scalars<<<1, 1, 0, st>>>(ctx->scal, eps, fixed_iters ? 1.0 : 0.0);
  CK_LAUNCH();
  int rc = 0;
  switch (P.dm) {
    case 1: rc = cg_init_launch<1>(ctx, b, multi); break;
    case 2: rc = cg_init_launch<2>(ctx, b, multi); break;
    case 3: rc = cg_init_launch<3>(ctx, b, multi); break;
    default: return femcy_fail_msg(ctx, "bad block size");
  }
  if (rc) return rc;
  if (multi) {
    if (femcy_cg_comm_allgather(ctx, 2)) return 1;
    k_finish_init<<<1, 1, 0, st>>>(ctx->scal, nranks);
    CK_LAUNCH();
  }
  int vg = vec_grid(n);
  {
    // all reduction scratch must exist BEFORE any stream capture: cudaMalloc/cudaFree are not permitted
    // while a stream is capturing (a realloc inside the capture invalidated it and left a null scratch)
    int64_t spmv_grid = ceil_div64(P.nslice, 8);
    if (femcy_ensure_reduction_scratch(ctx, spmv_grid > vg ? spmv_grid : vg)) return 1;
  }

  auto update_d_launch = [&]() -> int {
    if (multi == 2) {
      unsigned int* tk = ctx->red_ticket + 4;
      switch (P.dm) {
        case 1: k_update_d_p2p<1><<<vg, 256, 0, st>>>(d, r, M, (int)n, ctx->scal, pv, bflag, push_ptr, push_peer, push_ridx, bnodes, (int)n_bnodes, tk); break;
        case 2: k_update_d_p2p<2><<<vg, 256, 0, st>>>(d, r, M, (int)n, ctx->scal, pv, bflag, push_ptr, push_peer, push_ridx, bnodes, (int)n_bnodes, tk); break;
        default: k_update_d_p2p<3><<<vg, 256, 0, st>>>(d, r, M, (int)n, ctx->scal, pv, bflag, push_ptr, push_peer, push_ridx, bnodes, (int)n_bnodes, tk); break;
      }
    } else
      k_update_d<<<vg, 256, 0, st>>>(d, r, M, n, ctx->scal);
    CK_LAUNCH();
    return 0;
  };
  // peer-memory path: first push of d0 = M r0 (beta = 0) so that every rank's ghosts are filled
  if (multi == 2 && update_d_launch()) return 1;

  // option cg_profile: plain launches of the three-kernel path with a CUDA event after every kernel of the first
  // iterations; the per-kernel averages are returned by femcy_last_time_ms(kind 4/5/6 = spmv/update_xr/update_d)
  const bool profile = ctx->opt.cg_profile != 0;
  const int PROF_MAX = 64;
  std::vector<cudaEvent_t> pev;
  int prof_iters = 0;
  if (profile) {
    pev.resize(PROF_MAX * 3 + 1);
    for (auto& e : pev) cudaEventCreate(&e);
  }
  auto cleanup = [&]() { for (auto& e : pev) cudaEventDestroy(e); pev.clear(); };
#define CG_FAIL(expr) do { if (expr) { cleanup(); return 1; } } while (0)
  auto mark = [&](int slot) { if (profile && prof_iters < PROF_MAX) cudaEventRecord(pev[prof_iters * 3 + slot], st); };

  // one CG iteration of the three-kernel path = the launches below, always in this order (plain launches or graph capture)
  auto enqueue_iteration = [&]() -> int {
    if (profile && prof_iters == 0) cudaEventRecord(pev[0], st);
    if (multi == 1 && femcy_comm_halo(ctx, d)) return 1;
    if (spmv_dispatch(ctx, d, Ad, 1, multi, pv, slice_order, slice_ghost)) return 1;
    if (multi == 1) {
      if (femcy_cg_comm_allgather(ctx, 1)) return 1;
      k_finish_alpha<<<1, 1, 0, st>>>(ctx->scal, nranks);
      CK_LAUNCH();
    }
    mark(1);
    k_update_xr<<<vg, 256, 0, st>>>(x, r, d, Ad, M, n, ctx->red_partials, ctx->red_ticket, ctx->scal, multi, pv);
    CK_LAUNCH();
    if (multi == 1) {
      if (femcy_cg_comm_allgather(ctx, 2)) return 1;
      k_finish_beta<<<1, 1, 0, st>>>(ctx->scal, nranks);
      CK_LAUNCH();
    }
    mark(2);
    int rc2 = update_d_launch();
    mark(3);
    if (profile && prof_iters < PROF_MAX) ++prof_iters;
    return rc2;
  };

  // which kernel: 0 = auto, 1 = three-kernel graph, 2 = persistent with plain loads, 3 = streaming persistent
  int kernel = ctx->opt.cg_kernel;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
  if (profile || multi == 1) kernel = 1;             // the per-kernel profile and the NCCL exchange exist on this path only
  else if (kernel == 0) kernel = coop ? 3 : 1;
  if ((kernel == 2 || kernel == 3) && !coop) { cleanup(); return femcy_fail_msg(ctx, "persistent PCG kernels need cooperative launch"); }
  // cg_sym: the SpMV streams only the upper half of the matrix (SymPattern: built once per pattern, values copied from
  // the eliminated K at the start of every solve) and scatters the transposed products with fp64 atomics.  K is symmetric
  // after the reference's symmetric Dirichlet elimination (stiffnessMtrx.py:279-307); the iterates differ from the
  // full-matrix path by rounding only, and from run to run in the last bits (atomic order).
  const bool sym = ctx->opt.cg_sym != 0 && kernel != 1;

  // CUDA graph of `check_every` iterations of the three-kernel path (launch-bound at small per-GPU sizes / with NCCL
  // nodes): captured once per (matrix, chunk) and replayed
  bool use_graph = kernel == 1 && ctx->opt.no_graph == 0 && !profile && check_every > 1 && max_iter >= check_every;
  if (use_graph && (ctx->cg_graph_exec == nullptr || ctx->cg_graph_chunk != check_every || ctx->cg_graph_mode != multi)) {
    if (ctx->cg_graph_exec) { cudaGraphExecDestroy(ctx->cg_graph_exec); ctx->cg_graph_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    int64_t launches_before = ctx->launches;
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    int erc = 0;
    if (ce == cudaSuccess) {
      for (int c = 0; c < check_every && !erc; ++c) erc = enqueue_iteration();
      ce = cudaStreamEndCapture(st, &graph);
    }
    ctx->launches = launches_before;
    if (ce != cudaSuccess || erc || graph == nullptr ||
        cudaGraphInstantiate(&ctx->cg_graph_exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      ctx->cg_graph_exec = nullptr;
      use_graph = false;          // capture not possible (e.g. an old NCCL): plain launches
    } else {
      ctx->cg_graph_chunk = check_every;
      ctx->cg_graph_mode = multi;
      ctx->cg_graph_launches = (multi == 1 ? 11 : 3) * (int64_t)check_every;
    }
    if (graph) cudaGraphDestroy(graph);
  }

  CGPersistArgs pa;
  int pgrid = 0;
  StreamKernel sk = {nullptr, 0, 0};
  if (kernel == 2 || kernel == 3) {
    int nbsm = 0, nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    int64_t need_blocks;
    if (kernel == 3) {
      sk = stream_kernel(P.dm, sym, ctx->opt.cg_stream_cfg);
      cudaError_t ae = cudaFuncSetAttribute(sk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, sk.smem);
      cudaError_t oe = ae == cudaSuccess ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbsm, sk.fn, sk.threads, sk.smem) : ae;
      if (oe != cudaSuccess || nbsm < 1) { cleanup(); return femcy_fail(ctx, "streaming PCG kernel does not fit this device", oe, __FILE__, __LINE__); }
      nbsm = 1;                                                   // one block per SM: the ring takes the shared memory
      need_blocks = ceil_div64(P.nslice, sk.threads / 32);
    } else {
      cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbsm, persistent_kernel(P.dm, sym), 256, 0);
      if (oe != cudaSuccess || nbsm < 1) { cleanup(); return femcy_fail(ctx, "persistent PCG kernel does not fit this device", oe, __FILE__, __LINE__); }
      need_blocks = ceil_div64(P.nslice, 8);
    }
    pgrid = nbsm * nsm;
    if (need_blocks < pgrid) pgrid = (int)(need_blocks < 1 ? 1 : need_blocks);   // no point in more blocks than slice groups
    CG_FAIL(femcy_ensure_reduction_scratch(ctx, pgrid));        // capacity >= 4 doubles per block
    pa.slice_ptr = P.slice_ptr; pa.colidx = P.colidx; pa.val = P.val; pa.nrows = P.nn_own; pa.nslice = P.nslice;
    pa.x = x; pa.r = r; pa.d = d; pa.Ad = Ad; pa.M = M; pa.n = n;
    pa.part1 = ctx->red_partials; pa.part2 = ctx->red_partials + pgrid;
    pa.scal = ctx->scal; pa.p2p = (multi == 2) ? 1 : 0;
    pa.pv = pv; pa.bflag = bflag; pa.push_ptr = push_ptr; pa.push_peer = push_peer; pa.push_ridx = push_ridx;
    pa.bnodes = bnodes; pa.n_bnodes = (int)n_bnodes; pa.slice_order = slice_order; pa.slice_ghost = slice_ghost;
    pa.ticket = ctx->red_ticket + 6;
    pa.bpush = bpush;
    pa.rowof = P.rowof;
    if (sym) {
      CG_FAIL(femcy_build_sym_pattern(ctx) || femcy_sym_extract(ctx));
      if (cudaMemsetAsync(Ad, 0, (size_t)n * sizeof(double), st) != cudaSuccess) { cleanup(); return femcy_fail_msg(ctx, "memset(Ad)"); }
      pa.sym = 1; pa.u_slice_ptr = ctx->U.slice_ptr; pa.u_colidx = ctx->U.colidx; pa.u_val = ctx->U.val;
    }
  }
  auto launch_persistent = [&](int iters) -> int {
    pa.iters = iters;
    void* kargs[] = {(void*)&pa};
    cudaError_t le = kernel == 3
        ? cudaLaunchCooperativeKernel(sk.fn, dim3(pgrid), dim3(sk.threads), kargs, (size_t)sk.smem, st)
        : cudaLaunchCooperativeKernel(persistent_kernel(P.dm, sym), dim3(pgrid), dim3(256), kargs, 0, st);
    if (le != cudaSuccess) return femcy_fail(ctx, "cooperative launch", le, __FILE__, __LINE__);
    ctx->launches++;
    return 0;
  };

  if (cudaEventRecord(ctx->ev0, st) != cudaSuccess) { cleanup(); return femcy_fail_msg(ctx, "event record"); }
  int64_t it = 0;
  bool done = false;
  while (it < max_iter && !done) {
    int64_t chunk = check_every;
    if (it + chunk > max_iter) chunk = max_iter - it;
    if (kernel != 1) {
      CG_FAIL(launch_persistent((int)chunk));
    } else if (use_graph && chunk == check_every) {
      if (cudaGraphLaunch(ctx->cg_graph_exec, st) != cudaSuccess) { cleanup(); return femcy_fail_msg(ctx, "graph launch"); }
      ctx->launches += ctx->cg_graph_launches;
    } else {
      for (int64_t c = 0; c < chunk; ++c) CG_FAIL(enqueue_iteration());
    }
    it += chunk;
    if (!fixed_iters || it >= max_iter) {
      cudaError_t me = cudaMemcpyAsync(ctx->h_scal, ctx->scal, 16 * sizeof(double), cudaMemcpyDeviceToHost, st);
      if (me == cudaSuccess) me = cudaStreamSynchronize(st);
      if (me != cudaSuccess) { cleanup(); return femcy_fail(ctx, "PCG: reading the stop flag", me, __FILE__, __LINE__); }
      if (ctx->h_scal[S_DONE] != 0.0) done = true;
    }
  }
  {
    cudaError_t me = cudaEventRecord(ctx->ev1, st);
    if (me == cudaSuccess) me = cudaMemcpyAsync(ctx->h_scal, ctx->scal, 16 * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (me == cudaSuccess) me = cudaMemcpyAsync(ctx->h_scal + S_PHASE, ctx->scal + S_PHASE, S_PHASE_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (me == cudaSuccess) me = cudaStreamSynchronize(st);
    if (me != cudaSuccess) { cleanup(); return femcy_fail(ctx, "PCG: reading the result scalars", me, __FILE__, __LINE__); }
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->last_ms[1] = ms;
  if (profile) {
    double acc3[3] = {0, 0, 0};
    for (int i = 0; i < prof_iters; ++i)
      for (int k = 0; k < 3; ++k) {
        float t = 0;
        cudaEventElapsedTime(&t, pev[i * 3 + k], pev[i * 3 + k + 1]);
        acc3[k] += t;
      }
    for (int k = 0; k < 3; ++k) ctx->prof_ms[k] = prof_iters ? acc3[k] / prof_iters : 0.0;
  }
  cleanup();
#undef CG_FAIL
  for (int q = 0; q < S_PHASE_COUNT; ++q) ctx->cg_phase_ns[q] = ctx->h_scal[S_PHASE + q];
  if (iters_out) *iters_out = (int64_t)ctx->h_scal[S_ITER];
  if (rmax0_out) *rmax0_out = ctx->h_scal[S_R0];
  if (rmax_out) *rmax_out = ctx->h_scal[S_RMAX];
  // S_DONE == 3: a bounded wait on a peer (halo flag / partial-sum window) ran out inside the loop -- the iterate is
  // not a PCG iterate any more; fail loudly instead of returning numbers
  if (ctx->h_scal[S_DONE] == 3.0)
    return femcy_fail_msg(ctx, "PCG: timed out waiting for a peer GPU (NVLink peer-memory exchange); solution invalid");
  // S_DONE == 2: NaN / inf in the residual (a singular or indefinite system, a zero diagonal): distinct return code, so
  // that callers other than the Newton driver (which tests the residual norm itself) can detect the breakdown
  ctx->cg_breakdown = ctx->h_scal[S_DONE] == 2.0;
  return 0;
}

// Phase clock of the last femcy_cg_solve that ran the persistent kernel: nanoseconds (device globaltimer, block 0) summed
// over the iterations -- SpMV loop | barrier + fold | cross-rank exchange | x/r update | barrier + fold | exchange |
// d update + halo push + barrier.  All zero when another path ran.
extern "C" int femcy_cg_phase_ns(femcy_ctx* ctx, double* out7) {
  for (int q = 0; q < S_PHASE_COUNT; ++q) out7[q] = ctx->cg_phase_ns[q];
  return 0;
}
