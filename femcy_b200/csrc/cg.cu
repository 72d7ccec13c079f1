// Hot path rows a6 + a7: Jacobi-preconditioned conjugate gradients on the node-block SELL-32 matrix.
//
// Reference: ConjugateGradientSolver_rowMajor       /root/reference/conjugateGradientSolver.py:8-127
//   M_init :48   compute_Ad :53   r_d_init :60   rmax :67   compute_rMr :74
//   update_x/r/d :81/:86/:91   dot_product :96   solve :103-127
// The reference launches 8 kernels and reads 4 scalars back to the host per iteration.  Here one
// iteration is 3 kernels and no host round trip:
//   k_spmv_dot   : Ad = A d (warp per 32-row slice, coalesced plane loads), fused d.Ad reduction;
//                  the last block folds the per-block partials in index order and writes
//                  alpha = rMr / dAd                                     (:112-113)
//   k_update_xr  : x += alpha d ; r -= alpha Ad ; fused r.M.r and max|r| reductions; last block
//                  writes beta = rMr'/rMr, carries rMr' and evaluates the stopping rule
//                  max|r| < eps*max|r0| on the device                     (:114-124)
//   k_update_d   : d = M r + beta d                                       (:117)
// All reductions are two-stage with a fixed fold order (grid_reduce in elem_math.cuh) => bit-reproducible.
// Once the device-side stop flag is set every later kernel is a no-op, so polling the flag
// from the host only every `check_every` iterations still stops at exactly the reference's
// iteration.
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

// persistent kernels by (block size dm, blocks/SM they are compiled for)
static const void* persistent_kernel(int dm, bool single_red, int minb, bool sym = false) {
  if (sym && single_red) return dm == 1 ? (const void*)k_cg_persistent_sr<1, 4, true> : dm == 2 ? (const void*)k_cg_persistent_sr<2, 4, true> : (const void*)k_cg_persistent_sr<3, 4, true>;
  if (sym) return dm == 1 ? (const void*)k_cg_persistent<1, 4, true> : dm == 2 ? (const void*)k_cg_persistent<2, 4, true> : (const void*)k_cg_persistent<3, 4, true>;
  if (single_red) {
    if (minb == 5) return dm == 1 ? (const void*)k_cg_persistent_sr<1, 5> : dm == 2 ? (const void*)k_cg_persistent_sr<2, 5> : (const void*)k_cg_persistent_sr<3, 5>;
    return dm == 1 ? (const void*)k_cg_persistent_sr<1, 6> : dm == 2 ? (const void*)k_cg_persistent_sr<2, 6> : (const void*)k_cg_persistent_sr<3, 6>;
  }
  if (minb == 5) return dm == 1 ? (const void*)k_cg_persistent<1, 5> : dm == 2 ? (const void*)k_cg_persistent<2, 5> : (const void*)k_cg_persistent<3, 5>;
  return dm == 1 ? (const void*)k_cg_persistent<1, 6> : dm == 2 ? (const void*)k_cg_persistent<2, 6> : (const void*)k_cg_persistent<3, 6>;
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
  cudaStream_t st = ctx->stream;
  int nranks = femcy_comm_size(ctx);
  int multi = nranks > 1 ? 1 : 0;
  P2PView pv;
  const unsigned char* bflag = nullptr;
  const int32_t *push_ptr = nullptr, *push_peer = nullptr, *push_ridx = nullptr, *bnodes = nullptr;
  int64_t n_bnodes = 0;
  const int32_t* slice_order = nullptr;
  const unsigned char* slice_ghost = nullptr;
  if (multi && femcy_p2p_view(ctx, &pv, &bflag, &push_ptr, &push_peer, &push_ridx, &bnodes, &n_bnodes, &slice_order, &slice_ghost))
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

  // FEMCY_CG_PROFILE=1: plain launches with a CUDA event after every kernel of the first iterations;
  // the per-kernel averages are returned by femcy_last_time_ms(kind 4/5/6 = spmv/update_xr/update_d)
  const bool profile = getenv("FEMCY_CG_PROFILE") != nullptr;
  const int PROF_MAX = 64;
  std::vector<cudaEvent_t> pev;
  int prof_iters = 0;
  if (profile) {
    pev.resize(PROF_MAX * 3 + 1);
    for (auto& e : pev) cudaEventCreate(&e);
  }
  auto mark = [&](int slot) { if (profile && prof_iters < PROF_MAX) cudaEventRecord(pev[prof_iters * 3 + slot], st); };

  // one CG iteration = the launches below, always in this order (plain launches or graph capture)
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

  // CUDA graph of `check_every` iterations (launch-bound at small per-GPU sizes / with NCCL nodes):
  // captured once per (matrix, chunk) and replayed; FEMCY_NO_GRAPH=1 falls back to plain launches.
  bool use_graph = (getenv("FEMCY_NO_GRAPH") == nullptr) && !profile && check_every > 1 && max_iter >= check_every;
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

  // persistent cooperative kernel (single GPU and peer-memory path): one launch per `check_every` iterations
  // default (measured, profiles/r1_notes.md): on for a single GPU (3 % faster than the graph of three kernels at
  // 10 M elements, several times faster on launch-bound small systems) and on the peer-memory path from 4 ranks up
  // (N=4: 0.129 vs 0.135 ms/iteration, N=8: 0.0843 vs 0.0855); at N=2 the three-kernel graph is 4 % faster
  // (0.220 vs 0.229).  FEMCY_CG_PERSISTENT=1 / FEMCY_CG_MULTIKERNEL=1 force either path.
  const int cg_minb = (getenv("FEMCY_CG_MINB") != nullptr && atoi(getenv("FEMCY_CG_MINB")) == 5) ? 5 : 6;
  bool persistent = (multi != 1) && !profile && getenv("FEMCY_CG_MULTIKERNEL") == nullptr &&
                    (multi == 0 || nranks >= 4 || getenv("FEMCY_CG_PERSISTENT") != nullptr ||
                     (getenv("FEMCY_CG_VARIANT") != nullptr && strcmp(getenv("FEMCY_CG_VARIANT"), "sr") == 0) ||
                     (getenv("FEMCY_CG_SYM") != nullptr && atoi(getenv("FEMCY_CG_SYM")) != 0));   // (!profile is part of the product)
  // (FEMCY_CG_PROFILE, the per-kernel timing hook of the three-kernel path, always measures that path: the switch is
  //  ignored there instead of failing the call -- bench.py runs one profiled solve whatever the A/B environment is)
  const bool sym_req = !profile && getenv("FEMCY_CG_SYM") != nullptr && atoi(getenv("FEMCY_CG_SYM")) != 0;
  CGPersistArgs pa;
  int pgrid = 0;
  if (persistent) {
    int nbsm = 0, nsm = 0;
    cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbsm, persistent_kernel(P.dm, false, cg_minb, sym_req), 256, 0);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    if (oe != cudaSuccess || nbsm < 1 || !coop) {
      cudaGetLastError();
      persistent = false;
    } else {
      if (getenv("FEMCY_CG_BLOCKS_PER_SM") != nullptr) {         // A/B: fewer, fatter-loaded blocks make grid.sync / folds cheaper
        int want = atoi(getenv("FEMCY_CG_BLOCKS_PER_SM"));
        if (want >= 1 && want < nbsm) nbsm = want;
      }
      pgrid = nbsm * nsm;
      int64_t need_blocks = ceil_div64(P.nslice, 8);            // no point in more blocks than slice groups
      if (need_blocks < pgrid) pgrid = (int)(need_blocks < 1 ? 1 : need_blocks);
      if (femcy_ensure_reduction_scratch(ctx, pgrid)) return 1;   // capacity >= 4 doubles per block
      pa.slice_ptr = P.slice_ptr; pa.colidx = P.colidx; pa.val = P.val; pa.nrows = P.nn_own; pa.nslice = P.nslice;
      pa.x = x; pa.r = r; pa.d = d; pa.Ad = Ad; pa.M = M; pa.n = n;
      pa.part1 = ctx->red_partials; pa.part2 = ctx->red_partials + pgrid;
      pa.scal = ctx->scal; pa.p2p = (multi == 2) ? 1 : 0;
      pa.pv = pv; pa.bflag = bflag; pa.push_ptr = push_ptr; pa.push_peer = push_peer; pa.push_ridx = push_ridx;
      pa.bnodes = bnodes; pa.n_bnodes = (int)n_bnodes; pa.slice_order = slice_order; pa.slice_ghost = slice_ghost;
      pa.ticket = ctx->red_ticket + 6;
      pa.rowof = P.rowof;
      pa.fold_bar = (getenv("FEMCY_CG_FOLD_BARRIER") != nullptr && atoi(getenv("FEMCY_CG_FOLD_BARRIER")) != 0) ? 1 : 0;
      pa.bar_counter = ctx->red_ticket + 3; pa.bar_gen = ctx->red_ticket + 7; pa.bar_tot = ctx->scal + 48;
      pa.late_fence = (getenv("FEMCY_CG_LATE_FENCE") != nullptr && atoi(getenv("FEMCY_CG_LATE_FENCE")) != 0) ? 1 : 0;
      use_graph = false;
    }
  }
  // FEMCY_CG_SYM=1 (opt-in, unmeasured): the SpMV of the persistent kernel streams only the upper half of the matrix
  // (SymPattern: built once per pattern, values copied from the eliminated K at the start of every solve) and
  // scatters the transposed products with fp64 atomics.  K is symmetric after the reference's symmetric Dirichlet
  // elimination (stiffnessMtrx.py:279-307); the iterates differ from the default path by rounding only.
  if (sym_req && !persistent)
    return femcy_fail_msg(ctx, "FEMCY_CG_SYM needs a persistent kernel (not the NCCL path or FEMCY_CG_MULTIKERNEL)");
  if (sym_req) {
    if (femcy_build_sym_pattern(ctx) || femcy_sym_extract(ctx)) return 1;
    CK(cudaMemsetAsync(Ad, 0, (size_t)n * sizeof(double), st));
    pa.sym = 1; pa.u_slice_ptr = ctx->U.slice_ptr; pa.u_colidx = ctx->U.colidx; pa.u_val = ctx->U.val;
  }
  // opt-in single-reduction variant (k_cg_persistent_sr): FEMCY_CG_VARIANT=sr, cooperative launch required
  const char* cg_variant = getenv("FEMCY_CG_VARIANT");
  const bool single_red = persistent && cg_variant != nullptr && strcmp(cg_variant, "sr") == 0;
  CGSingleRedArgs sa;
  bool sr_first = true;
  if (single_red) {
    int64_t Nfull = ctx->nn * ctx->dm;
    if (ctx->cg_ps_len != Nfull) {
      if (femcy_alloc(ctx, &ctx->cg_p, Nfull) || femcy_alloc(ctx, &ctx->cg_s, Nfull)) return 1;
      ctx->cg_ps_len = Nfull;
    }
    CK(cudaMemsetAsync(ctx->cg_p, 0, (size_t)Nfull * sizeof(double), st));
    CK(cudaMemsetAsync(ctx->cg_s, 0, (size_t)Nfull * sizeof(double), st));
    int nbsm = 0, nsm = 0;
    cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbsm, persistent_kernel(P.dm, true, cg_minb, sym_req), 256, 0);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    if (oe != cudaSuccess || nbsm < 1) return femcy_fail_msg(ctx, "FEMCY_CG_VARIANT=sr: occupancy query failed");
    if (getenv("FEMCY_CG_BLOCKS_PER_SM") != nullptr) {
      int want = atoi(getenv("FEMCY_CG_BLOCKS_PER_SM"));
      if (want >= 1 && want < nbsm) nbsm = want;
    }
    int sgrid = nbsm * nsm;
    int64_t need_blocks = ceil_div64(P.nslice, 8);
    if (need_blocks < sgrid) sgrid = (int)(need_blocks < 1 ? 1 : need_blocks);
    pgrid = sgrid;
    if (femcy_ensure_reduction_scratch(ctx, 2 * (int64_t)pgrid)) return 1;   // capacity >= 4 doubles per block: 2 x [grid*3] fits
    sa.slice_ptr = P.slice_ptr; sa.colidx = P.colidx; sa.val = P.val; sa.nrows = P.nn_own; sa.nslice = P.nslice;
    sa.x = x; sa.r = r; sa.u = d; sa.w = Ad; sa.p = ctx->cg_p; sa.s = ctx->cg_s; sa.M = M; sa.n = n;
    sa.part = ctx->red_partials; sa.scal = ctx->scal; sa.p2p = (multi == 2) ? 1 : 0;
    sa.pv = pv; sa.bflag = bflag; sa.push_ptr = push_ptr; sa.push_peer = push_peer; sa.push_ridx = push_ridx;
    sa.bnodes = bnodes; sa.n_bnodes = (int)n_bnodes; sa.slice_order = slice_order; sa.slice_ghost = slice_ghost;
    sa.ticket = ctx->red_ticket + 6;
    sa.rowof = P.rowof;
    sa.fold_bar = (getenv("FEMCY_CG_FOLD_BARRIER") != nullptr && atoi(getenv("FEMCY_CG_FOLD_BARRIER")) != 0) ? 1 : 0;
    sa.bar_counter = ctx->red_ticket + 3; sa.bar_gen = ctx->red_ticket + 7; sa.bar_tot = ctx->scal + 48;
    sa.late_fence = (getenv("FEMCY_CG_LATE_FENCE") != nullptr && atoi(getenv("FEMCY_CG_LATE_FENCE")) != 0) ? 1 : 0;
    if (sym_req) { sa.sym = 1; sa.u_slice_ptr = ctx->U.slice_ptr; sa.u_colidx = ctx->U.colidx; sa.u_val = ctx->U.val; }   // sa.w = Ad is zero (memset above)
  }
  auto launch_persistent = [&](int iters) -> int {
    if (single_red) {
      sa.iters = iters;
      sa.first = sr_first ? 1 : 0;
      sr_first = false;
      void* kargs[] = {(void*)&sa};
      cudaError_t le = cudaLaunchCooperativeKernel(persistent_kernel(P.dm, true, cg_minb, sa.sym != 0), dim3(pgrid), dim3(256), kargs, 0, st);
      if (le != cudaSuccess) return femcy_fail(ctx, "cooperative launch (single-reduction PCG)", le, __FILE__, __LINE__);
      ctx->launches++;
      return 0;
    }
    pa.iters = iters;
    void* kargs[] = {(void*)&pa};
    cudaError_t le = cudaLaunchCooperativeKernel(persistent_kernel(P.dm, false, cg_minb, pa.sym != 0), dim3(pgrid), dim3(256), kargs, 0, st);
    if (le != cudaSuccess) return femcy_fail(ctx, "cooperative launch", le, __FILE__, __LINE__);
    ctx->launches++;
    return 0;
  };

  // FEMCY_CG_L2_PERSIST (opt-in, unmeasured): an L2 access-policy window for the duration of the solve, reset at its end.
  //   1: on the direction vector -- the SpMV's gather target, read ~15 times per iteration, then once by each vector pass;
  //   2: on the matrix values the SpMV streams (upper half with FEMCY_CG_SYM, whose loads then drop the evict-first hint):
  //      hitRatio = persisting capacity / window, so that fraction of the matrix stays in L2 from one iteration to the
  //      next.  Meant for the multi-GPU case: at 8 ranks the upper half of cfg 4 is 131 MB per rank against 126 MB of L2.
  bool l2_window = false;
  const int l2_mode = getenv("FEMCY_CG_L2_PERSIST") != nullptr ? atoi(getenv("FEMCY_CG_L2_PERSIST")) : 0;
  if (l2_mode == 1 || l2_mode == 2) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    const void* base = (const void*)d;
    size_t want = (size_t)(ctx->nn * ctx->dm) * sizeof(double);       // owned + ghost entries of d
    if (l2_mode == 2) {
      base = sym_req ? (const void*)ctx->U.val : (const void*)P.val;
      want = (size_t)((sym_req ? ctx->U.nslots : P.nslots) * P.dm * P.dm) * sizeof(double);
    }
    if (max_persist > 0 && max_window > 0 && want > 0) {
      size_t bytes = want < (size_t)max_window ? want : (size_t)max_window;
      size_t carve = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
      cudaStreamAttrValue av;
      memset(&av, 0, sizeof(av));
      av.accessPolicyWindow.base_ptr = const_cast<void*>(base);
      av.accessPolicyWindow.num_bytes = bytes;
      av.accessPolicyWindow.hitRatio = (float)((double)carve / (double)bytes);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = (l2_mode == 2) ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess &&
          cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess) {
        l2_window = true;
        if (l2_mode == 2) { pa.mat_plain = 1; sa.mat_plain = 1; }
      } else {
        cudaGetLastError();
      }
    }
  }
  auto drop_l2_window = [&]() {
    if (!l2_window) return;
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof(av));
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
    cudaCtxResetPersistingL2Cache();
    l2_window = false;
  };
  CK(cudaEventRecord(ctx->ev0, st));
  int64_t it = 0;
  bool done = false;
  while (it < max_iter && !done) {
    int64_t chunk = check_every;
    if (it + chunk > max_iter) chunk = max_iter - it;
    if (persistent) {
      if (launch_persistent((int)chunk)) return 1;
    } else if (use_graph && chunk == check_every) {
      CK(cudaGraphLaunch(ctx->cg_graph_exec, st));
      ctx->launches += ctx->cg_graph_launches;
    } else {
      for (int64_t c = 0; c < chunk; ++c)
        if (enqueue_iteration()) return 1;
    }
    it += chunk;
    if (!fixed_iters || it >= max_iter) {
      CK(cudaMemcpyAsync(ctx->h_scal, ctx->scal, 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (ctx->h_scal[S_DONE] != 0.0) done = true;
    }
  }
  CK(cudaEventRecord(ctx->ev1, st));
  CK(cudaMemcpyAsync(ctx->h_scal, ctx->scal, 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  drop_l2_window();
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
    for (auto& e : pev) cudaEventDestroy(e);
  }
  if (iters_out) *iters_out = (int64_t)ctx->h_scal[S_ITER];
  if (rmax0_out) *rmax0_out = ctx->h_scal[S_R0];
  if (rmax_out) *rmax_out = ctx->h_scal[S_RMAX];
  // S_DONE == 3: a bounded wait on a peer (halo flag / partial-sum window) ran out inside the loop -- the iterate is
  // not a PCG iterate any more; fail loudly instead of returning numbers
  if (ctx->h_scal[S_DONE] == 3.0)
    return femcy_fail_msg(ctx, "PCG: timed out waiting for a peer GPU (NVLink peer-memory exchange); solution invalid");
  return 0;
}
