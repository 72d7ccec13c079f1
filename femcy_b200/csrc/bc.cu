// Row a4: Dirichlet boundary conditions on the node-block SELL-32 matrix.
//
// Reference kernels replaced:
//   dirichletBC_linearEquations          /root/reference/stiffnessMtrx.py:279-307
//   dirichletBC_forNewtonMethod_kernel   /root/reference/stiffnessMtrx.py:317-341
//   dirichletBC_val                      /root/reference/stiffnessMtrx.py:357-366
//
// The reference runs one thread per constrained node and searches rows (and has a benign rhs
// race).  Here the constrained dofs are first marked in a flag vector, then one warp
// per 32-row slice walks its column indices; values are touched only where a row or column dof
// is flagged.  Each row is handled by exactly one lane => no atomics, deterministic, and equal
// to the reference's *sequential* semantics whenever no dof is listed twice in one call:
//   free row r:        rhs[r] -= sum_{constrained c} val_c * K[r][c] ; K[r][c] = 0
//   constrained row i: rhs[i] = val_i ; K[i][:] = 0 ; K[i][i] = 1
#include "ctx.cuh"
#include "bc_kernels.cuh"

static int upload_bc(femcy_ctx* ctx, const int32_t* nodes, const int32_t* comps, const double* vals, int64_t n) {
  if (n > ctx->bc_cap) {
    int64_t cap = n + 1024;
    if (femcy_alloc(ctx, &ctx->bc_nodes, cap) || femcy_alloc(ctx, &ctx->bc_comps, cap) || femcy_alloc(ctx, &ctx->bc_vals, cap)) return 1;
    ctx->bc_cap = cap;
  }
  for (int64_t t = 0; t < n; ++t) {
    if (nodes[t] < 0 || nodes[t] >= ctx->nn || comps[t] < 0 || comps[t] >= ctx->dm)
      return femcy_fail_msg(ctx, "Dirichlet node/component out of range");
  }
  CK(cudaMemcpyAsync(ctx->bc_nodes, nodes, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->bc_comps, comps, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (vals) CK(cudaMemcpyAsync(ctx->bc_vals, vals, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // host buffers are borrowed only for the call
  return 0;
}

static inline int gridb(int64_t n) {
  int64_t g = ceil_div64(n, 256);
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

static int bc_apply(femcy_ctx* ctx, int64_t n, bool have_vals, int mode, int target_vec) {
  BsellPattern& P = ctx->P;
  cudaStream_t st = ctx->stream;
  if (n == 0) return 0;
  k_bc_mark<<<gridb(n), 256, 0, st>>>(ctx->bc_nodes, ctx->bc_comps, have_vals ? ctx->bc_vals : nullptr, n, ctx->dm,
                                      ctx->bc_flag, ctx->bc_val_full, 1);
  CK_LAUNCH();
  int grid = (int)ceil_div64(P.nslice, 8);
  if (grid < 1) grid = 1;
  switch (P.dm) {
    case 1: k_bc_apply<1><<<grid, 256, 0, st>>>(P.slice_ptr, P.colidx, P.val, P.nn_own, P.nslice, ctx->bc_flag, ctx->bc_val_full, ctx->vec[target_vec], mode, P.rowof); break;
    case 2: k_bc_apply<2><<<grid, 256, 0, st>>>(P.slice_ptr, P.colidx, P.val, P.nn_own, P.nslice, ctx->bc_flag, ctx->bc_val_full, ctx->vec[target_vec], mode, P.rowof); break;
    case 3: k_bc_apply<3><<<grid, 256, 0, st>>>(P.slice_ptr, P.colidx, P.val, P.nn_own, P.nslice, ctx->bc_flag, ctx->bc_val_full, ctx->vec[target_vec], mode, P.rowof); break;
    default: return femcy_fail_msg(ctx, "bad block size");
  }
  CK_LAUNCH();
  k_bc_mark<<<gridb(n), 256, 0, st>>>(ctx->bc_nodes, ctx->bc_comps, nullptr, n, ctx->dm, ctx->bc_flag, ctx->bc_val_full, 0);
  CK_LAUNCH();
  return 0;
}

extern "C" int femcy_dirichlet_linear(femcy_ctx* ctx, const int32_t* nodes, const int32_t* comps, const double* vals, int64_t n) {
  cudaSetDevice(ctx->device);
  if (!ctx->P.val) return femcy_fail_msg(ctx, "no matrix");
  if (upload_bc(ctx, nodes, comps, vals, n)) return 1;
  return bc_apply(ctx, n, true, 0, FEMCY_VEC_RHS);
}

extern "C" int femcy_dirichlet_newton(femcy_ctx* ctx, const int32_t* nodes, const int32_t* comps, int64_t n) {
  cudaSetDevice(ctx->device);
  if (!ctx->P.val) return femcy_fail_msg(ctx, "no matrix");
  if (upload_bc(ctx, nodes, comps, nullptr, n)) return 1;
  return bc_apply(ctx, n, false, 1, FEMCY_VEC_RESIDUAL);
}

extern "C" int femcy_dirichlet_val(femcy_ctx* ctx, const int32_t* nodes, const int32_t* comps, const double* vals, int64_t n) {
  cudaSetDevice(ctx->device);
  if (!ctx->vec[FEMCY_VEC_DOF]) return femcy_fail_msg(ctx, "state not allocated");
  if (n == 0) return 0;
  if (upload_bc(ctx, nodes, comps, vals, n)) return 1;
  k_bc_val<<<gridb(n), 256, 0, ctx->stream>>>(ctx->bc_nodes, ctx->bc_comps, ctx->bc_vals, n, ctx->dm, ctx->vec[FEMCY_VEC_DOF]);
  CK_LAUNCH();
  return 0;
}
