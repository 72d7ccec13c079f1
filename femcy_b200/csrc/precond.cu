// Row f2: a stronger preconditioner behind the same CG surface -- OPT-IN (option cg_precond = 1), because it
// changes the iteration path: the reference's solver is plain Jacobi-PCG (conjugateGradientSolver.py:48-51, :103-127) and
// that stays the default, bit-unchanged.
//
// Two-level additive preconditioner   z = Cheb_2(D^-1 A) D^-1 r  +  P (P^T A P)^-1 P^T r
//   * smoother: 2 steps of the Chebyshev iteration on the Jacobi-scaled operator over [lmax/30, lmax], lmax from a few power
//     iterations (one extra SpMV per application);
//   * coarse space: aggregates of nodes (host: geometric bins, femcy_set_aggregates) x the rigid-body modes of each aggregate
//     (3 translations + 3 rotations in 3-D, 2 + 1 in 2-D) -- the near-null space of elasticity that Jacobi cannot see;
//     A_c = P^T A P is assembled from the block matrix on the device, inverted densely once per solve (cuSOLVER potrf / potri
//     through dlopen -- a library call on the coarse level, not on the hot path) and applied as a dense mat-vec.
// Measured: 15-50x fewer iterations at 2 SpMV per iteration on the meshes of BASELINE.json.
// Single GPU only: under a partition femcy_cg_solve refuses the option instead of ignoring it.
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ctx.cuh"
#include "device_compat.cuh"
#include "elem_math.cuh"

// rigid-body modes per aggregate
template <int DM> struct Rbm { static constexpr int NR = (DM == 2) ? 3 : 6; };

// T_i (DM x NR): displacement of node i (position rel to its aggregate's centroid) under the unit rigid-body modes
template <int DM>
__device__ __forceinline__ void rbm_rows(const double (&rel)[DM], double (&T)[DM][Rbm<DM>::NR]) {
  constexpr int NR = Rbm<DM>::NR;
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int m = 0; m < NR; ++m) T[i][m] = (m == i) ? 1.0 : 0.0;
  if constexpr (DM == 2) {
    T[0][2] = -rel[1]; T[1][2] = rel[0];                   // rotation about z: u = (-y, x)
  } else {
    T[0][4] = rel[2]; T[0][5] = -rel[1];                   // u = e_m x rel
    T[1][3] = -rel[2]; T[1][5] = rel[0];
    T[2][3] = rel[1]; T[2][4] = -rel[0];
  }
}

struct Precond2 {
  int64_t nagg = 0, nn = 0;
  int dm = 0;
  int32_t* agg = nullptr;        // [nn] aggregate of every (owned) node
  int32_t* agg_ptr = nullptr;    // [nagg+1] nodes grouped by aggregate
  int32_t* agg_nodes = nullptr;  // [nn]
  double* cen = nullptr;         // [nagg*dm] centroids
  double* Ac = nullptr;          // [nc*nc] coarse matrix -> its inverse
  double* rc = nullptr;          // [nc]
  double* yc = nullptr;          // [nc]
  double *z = nullptr, *dv = nullptr, *tmp = nullptr;   // [nn*dm]
  int64_t vec_len = 0, nc_alloc = 0;
  void* solver_lib = nullptr; void* solver_handle = nullptr; void* work = nullptr; int64_t work_len = 0; int* info = nullptr;
};

static Precond2* pc_of(femcy_ctx* ctx) {
  if (!ctx->precond2) ctx->precond2 = new Precond2();
  return static_cast<Precond2*>(ctx->precond2);
}

void femcy_precond_free(femcy_ctx* ctx) {
  Precond2* p = static_cast<Precond2*>(ctx->precond2);
  if (!p) return;
  femcy_free(&p->agg); femcy_free(&p->agg_ptr); femcy_free(&p->agg_nodes); femcy_free(&p->cen); femcy_free(&p->Ac);
  femcy_free(&p->rc); femcy_free(&p->yc); femcy_free(&p->z); femcy_free(&p->dv); femcy_free(&p->tmp);
  if (p->work) cudaFree(p->work);
  if (p->info) cudaFree(p->info);
  if (p->solver_handle && p->solver_lib) {
    typedef int (*Destroy)(void*);
    Destroy d = (Destroy)dlsym(p->solver_lib, "cusolverDnDestroy");
    if (d) d(p->solver_handle);
  }
  if (p->solver_lib) dlclose(p->solver_lib);
  delete p;
  ctx->precond2 = nullptr;
}

// ---- setup kernels --------------------------------------------------------------------------------------------------
template <int DM>
__global__ void k_agg_centroid(const double* __restrict__ nodes, const int32_t* __restrict__ agg_ptr,
                               const int32_t* __restrict__ agg_nodes, int64_t nagg, double* __restrict__ cen) {
  // one warp per aggregate, fixed order => reproducible
  const int lane = threadIdx.x & 31;
  const int64_t a = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (a >= nagg) return;
  double s[DM];
#pragma unroll
  for (int c = 0; c < DM; ++c) s[c] = 0.0;
  const int b = agg_ptr[a], e = agg_ptr[a + 1];
  for (int t = b + lane; t < e; t += 32) {
    const int64_t n = agg_nodes[t];
#pragma unroll
    for (int c = 0; c < DM; ++c) s[c] += nodes[n * DM + c];
  }
#pragma unroll
  for (int c = 0; c < DM; ++c) s[c] = warp_sum(s[c]);
  if (lane == 0) {
    const double inv = (e > b) ? 1.0 / (double)(e - b) : 0.0;
#pragma unroll
    for (int c = 0; c < DM; ++c) cen[a * DM + c] = s[c] * inv;
  }
}

// A_c += T_i^T K_ij T_j for every stored block (fp64 atomics: once per solve, off the iteration path)
template <int DM>
__global__ void __launch_bounds__(256)
k_coarse_matrix(const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ colidx, const double* __restrict__ val,
                int64_t nrows, int64_t nslice, const int32_t* __restrict__ rowof, const double* __restrict__ nodes,
                const int32_t* __restrict__ agg, const double* __restrict__ cen, double* __restrict__ Ac, int64_t nc) {
  constexpr int DM2 = DM * DM, NR = Rbm<DM>::NR;
  const int lane = threadIdx.x & 31;
  const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (s >= nslice) return;
  int64_t i = s * 32 + lane;
  bool ok = i < nrows;
  if (rowof) { i = ok ? rowof[i] : -1; ok = i >= 0; }
  if (!ok) return;
  const int ai = agg[i];
  double rel[DM], Ti[DM][NR];
#pragma unroll
  for (int c = 0; c < DM; ++c) rel[c] = nodes[i * DM + c] - cen[(int64_t)ai * DM + c];
  rbm_rows<DM>(rel, Ti);
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) >> 5;
  for (int k = 0; k < w; ++k) {
    const int c = colidx[base + (k << 5) + lane];
    if (c < 0 || c >= nrows) continue;            // padding / ghost column (single GPU: none)
    double a[DM2];
#pragma unroll
    for (int q = 0; q < DM2; ++q) a[q] = val[((((int64_t)(base >> 5) + k) * DM2 + q) << 5) + lane];
    const int aj = agg[c];
    double relj[DM], Tj[DM][NR];
#pragma unroll
    for (int d = 0; d < DM; ++d) relj[d] = nodes[(int64_t)c * DM + d] - cen[(int64_t)aj * DM + d];
    rbm_rows<DM>(relj, Tj);
    // KT = K_ij T_j (DM x NR), then T_i^T KT (NR x NR)
    double KT[DM][NR];
#pragma unroll
    for (int r = 0; r < DM; ++r)
#pragma unroll
      for (int m = 0; m < NR; ++m) {
        double t = 0.0;
#pragma unroll
        for (int d = 0; d < DM; ++d) t += a[r * DM + d] * Tj[d][m];
        KT[r][m] = t;
      }
    double* dst = Ac + ((int64_t)ai * NR) * nc + (int64_t)aj * NR;
#pragma unroll
    for (int p = 0; p < NR; ++p)
#pragma unroll
      for (int m = 0; m < NR; ++m) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < DM; ++r) t += Ti[r][p] * KT[r][m];
        if (t != 0.0) femcy_red_add_f64(dst + (int64_t)p * nc + m, t);
      }
  }
}

// symmetrise (average) + shift the diagonal; empty modes (zero diagonal) become identity rows
__global__ void k_coarse_fix(double* __restrict__ Ac, int64_t nc, double shift_rel, const double* __restrict__ dmax) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nc * nc) return;
  const int64_t r = t / nc, c = t - r * nc;
  if (r > c) return;
  if (r == c) {
    double d = Ac[t];
    Ac[t] = (d > 0.0) ? d + shift_rel * dmax[0] : 1.0;
  } else {
    double v = 0.5 * (Ac[t] + Ac[c * nc + r]);
    Ac[t] = v; Ac[c * nc + r] = v;
  }
}
__global__ void k_diag_max(const double* __restrict__ Ac, int64_t nc, double* __restrict__ out) {
  // single block
  __shared__ double sh[256];
  double m = 0.0;
  for (int64_t i = threadIdx.x; i < nc; i += blockDim.x) m = fmax(m, Ac[i * nc + i]);
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]); __syncthreads(); }
  if (threadIdx.x == 0) out[0] = sh[0];
}
__global__ void k_mirror_lower(double* __restrict__ A, int64_t nc) {   // potri fills one triangle: copy it to the other
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nc * nc) return;
  const int64_t r = t / nc, c = t - r * nc;
  // cuSOLVER's "lower" triangle of the column-major view (i >= j at A[i + j*n]) is the UPPER triangle (c >= r) of this
  // row-major view: fill the other half from it
  if (r > c) A[t] = A[c * nc + r];
}

// ---- per-iteration kernels ------------------------------------------------------------------------------------------
// rc = P^T r : one block per aggregate (fixed order => reproducible)
template <int DM>
__global__ void __launch_bounds__(128)
k_restrict(const double* __restrict__ r, const double* __restrict__ nodes, const int32_t* __restrict__ agg_ptr,
           const int32_t* __restrict__ agg_nodes, const double* __restrict__ cen, double* __restrict__ rc) {
  constexpr int NR = Rbm<DM>::NR;
  __shared__ double sh[NR][4];
  const int a = blockIdx.x;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double acc[NR];
#pragma unroll
  for (int m = 0; m < NR; ++m) acc[m] = 0.0;
  const int b = agg_ptr[a], e = agg_ptr[a + 1];
  for (int t = b + threadIdx.x; t < e; t += blockDim.x) {
    const int64_t n = agg_nodes[t];
    double rel[DM], T[DM][NR];
#pragma unroll
    for (int c = 0; c < DM; ++c) rel[c] = nodes[n * DM + c] - cen[(int64_t)a * DM + c];
    rbm_rows<DM>(rel, T);
#pragma unroll
    for (int c = 0; c < DM; ++c) {
      const double rv = r[n * DM + c];
#pragma unroll
      for (int m = 0; m < NR; ++m) acc[m] += T[c][m] * rv;
    }
  }
#pragma unroll
  for (int m = 0; m < NR; ++m) { acc[m] = warp_sum(acc[m]); if (lane == 0) sh[m][wib] = acc[m]; }
  __syncthreads();
  if (threadIdx.x < NR) rc[(int64_t)a * NR + threadIdx.x] = (sh[threadIdx.x][0] + sh[threadIdx.x][1]) + (sh[threadIdx.x][2] + sh[threadIdx.x][3]);
}

// yc = Ainv rc : one warp per row
__global__ void __launch_bounds__(256)
k_dense_matvec(const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  double s = 0.0;
  for (int64_t c = lane; c < n; c += 32) s += A[row * n + c] * x[c];
  s = warp_sum(s);
  if (lane == 0) y[row] = s;
}

// dv = (D^-1 r) / theta
__global__ void k_cheb_first(const double* __restrict__ r, const double* __restrict__ Dinv, double* __restrict__ dv, int64_t n,
                             double inv_theta) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dv[i] = Dinv[i] * r[i] * inv_theta;
}

// z = dv0 + [c1 dv0 + c2 D^-1 (r - A dv0)] + P yc ; partial r.z ; the last block: rz -> beta = rz / rz_old (first: beta = 0)
template <int DM>
__global__ void __launch_bounds__(256)
k_precond_finish(const double* __restrict__ r, const double* __restrict__ Dinv, const double* __restrict__ dv,
                 const double* __restrict__ Adv, const double* __restrict__ nodes, const int32_t* __restrict__ agg,
                 const double* __restrict__ cen, const double* __restrict__ yc, double* __restrict__ z, int64_t nn_own,
                 double c1, double c2, double* partials, unsigned int* ticket, double* scal, int first) {
  constexpr int NR = Rbm<DM>::NR;
  double rz = 0.0;
  for (int64_t nd = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; nd < nn_own; nd += (int64_t)gridDim.x * blockDim.x) {
    const int a = agg[nd];
    double rel[DM], T[DM][NR], y[NR];
#pragma unroll
    for (int c = 0; c < DM; ++c) rel[c] = nodes[nd * DM + c] - cen[(int64_t)a * DM + c];
    rbm_rows<DM>(rel, T);
#pragma unroll
    for (int m = 0; m < NR; ++m) y[m] = yc[(int64_t)a * NR + m];
#pragma unroll
    for (int c = 0; c < DM; ++c) {
      const int64_t i = nd * DM + c;
      const double d0 = dv[i], rv = r[i];
      double zi = d0 + (c1 * d0 + c2 * (Dinv[i] * (rv - Adv[i])));
#pragma unroll
      for (int m = 0; m < NR; ++m) zi += T[c][m] * y[m];
      z[i] = zi;
      rz += rv * zi;
    }
  }
  double mine[1] = {rz}, tot[1];
  const bool is_max[1] = {false};
  if (grid_reduce<1>(mine, partials, ticket, tot, is_max)) {
    scal[S_BETA] = first ? 0.0 : tot[0] / scal[S_RMR];
    scal[S_RMR] = tot[0];
  }
}

// x += alpha d ; r -= alpha Ad ; max|r| ; the last block: iteration count + the reference's stop rule (:124)
__global__ void __launch_bounds__(256)
k2_update_xr(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ d, const double* __restrict__ Ad,
             int64_t n, double* partials, unsigned int* ticket, double* scal) {
  if (scal[S_DONE] != 0.0) return;
  const double alpha = scal[S_ALPHA];
  double rmax = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = x[i] + alpha * d[i];
    const double rn = r[i] - alpha * Ad[i];
    r[i] = rn;
    rmax = fmax(rmax, fabs(rn));
    if (rn != rn) rmax = 1.0 / 0.0;
  }
  double mine[1] = {rmax}, tot[1];
  const bool is_max[1] = {true};
  if (grid_reduce<1>(mine, partials, ticket, tot, is_max)) {
    scal[S_RMAX] = tot[0];
    scal[S_ITER] = scal[S_ITER] + 1.0;
    if (scal[S_FIXED] == 0.0 && tot[0] < scal[S_EPS] * scal[S_R0]) scal[S_DONE] = 1.0;
    if (!(tot[0] < 1.0e300)) scal[S_DONE] = 2.0;
  }
}

// d = z + beta d  (first: d = z)
__global__ void k2_update_d(double* __restrict__ d, const double* __restrict__ z, int64_t n, const double* __restrict__ scal) {
  if (scal[S_DONE] != 0.0) return;
  const double beta = scal[S_BETA];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = z[i] + beta * d[i];
}

__global__ void k2_init(const int32_t* __restrict__ diag_slot, const double* __restrict__ val, const double* __restrict__ b,
                        double* __restrict__ x, double* __restrict__ r, double* __restrict__ Dinv, int64_t nrows, int dm,
                        double* partials, unsigned int* ticket, double* scal) {
  const int dm2 = dm * dm;
  double rmax = 0.0;
  const int64_t n = nrows * dm;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / dm;
    const int c = (int)(t - i * dm);
    const int slot = diag_slot[i];
    const double diag = (slot >= 0) ? val[(((int64_t)(slot >> 5) * dm2 + (c * dm + c)) << 5) + (slot & 31)] : 0.0;
    Dinv[t] = 1.0 / diag;
    const double bi = b[t];
    r[t] = bi;
    x[t] = 0.0;
    rmax = fmax(rmax, fabs(bi));
  }
  double mine[1] = {rmax}, tot[1];
  const bool is_max[1] = {true};
  if (grid_reduce<1>(mine, partials, ticket, tot, is_max)) {
    scal[S_R0] = tot[0]; scal[S_RMAX] = tot[0];
    if (tot[0] == 0.0) scal[S_DONE] = 1.0;
  }
}

// max|v| -> out3[1] (deterministic grid reduction)
__global__ void __launch_bounds__(256)
k2_absmax(const double* __restrict__ v, int64_t n, double* partials, unsigned int* ticket, double* __restrict__ out3) {
  double m = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmax(m, fabs(v[i]));
  double mine[1] = {m}, tot[1];
  const bool is_max[1] = {true};
  if (grid_reduce<1>(mine, partials, ticket, tot, is_max)) out3[1] = tot[0];
}

// power iteration helpers: v <- D^-1 (A v) / max|.|
__global__ void k_scale_by(double* __restrict__ v, const double* __restrict__ Av, const double* __restrict__ Dinv, int64_t n,
                           const double* __restrict__ norm3) {
  const double inv = norm3[1] > 0.0 ? 1.0 / norm3[1] : 1.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = Dinv[i] * Av[i] * inv;
}
__global__ void k_mul(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = a[i] * b[i];
}
__global__ void k_seed(double* __restrict__ v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    v[i] = 0.5 + (double)(h & 0xffffff) / 16777216.0;
  }
}

// ---- host -------------------------------------------------------------------------------------------------------------
extern "C" int femcy_set_aggregates(femcy_ctx* ctx, int64_t nagg, const int32_t* agg_of_node) {
  cudaSetDevice(ctx->device);
  if (nagg < 1 || !agg_of_node) return femcy_fail_msg(ctx, "femcy_set_aggregates: need at least one aggregate");
  if (ctx->nn_own != ctx->nn) return femcy_fail_msg(ctx, "the two-level preconditioner is single-GPU only");
  Precond2* p = pc_of(ctx);
  const int64_t nn = ctx->nn;
  std::vector<int32_t> cnt(nagg + 1, 0), nodes_sorted(nn);
  for (int64_t i = 0; i < nn; ++i) {
    if (agg_of_node[i] < 0 || agg_of_node[i] >= nagg) return femcy_fail_msg(ctx, "femcy_set_aggregates: aggregate id out of range");
    cnt[agg_of_node[i] + 1]++;
  }
  for (int64_t a = 0; a < nagg; ++a) cnt[a + 1] += cnt[a];
  std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
  for (int64_t i = 0; i < nn; ++i) nodes_sorted[fill[agg_of_node[i]]++] = (int32_t)i;
  p->nagg = nagg; p->nn = nn; p->dm = ctx->dm;
  if (femcy_alloc(ctx, &p->agg, nn) || femcy_alloc(ctx, &p->agg_ptr, nagg + 1) || femcy_alloc(ctx, &p->agg_nodes, nn) ||
      femcy_alloc(ctx, &p->cen, nagg * ctx->dm))
    return 1;
  CK(cudaMemcpy(p->agg, agg_of_node, (size_t)nn * sizeof(int32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->agg_ptr, cnt.data(), (size_t)(nagg + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->agg_nodes, nodes_sorted.data(), (size_t)nn * sizeof(int32_t), cudaMemcpyHostToDevice));
  const int g = (int)ceil_div64(nagg * 32, 256);
  if (ctx->dm == 2) k_agg_centroid<2><<<g, 256, 0, ctx->stream>>>(ctx->nodes, p->agg_ptr, p->agg_nodes, nagg, p->cen);
  else k_agg_centroid<3><<<g, 256, 0, ctx->stream>>>(ctx->nodes, p->agg_ptr, p->agg_nodes, nagg, p->cen);
  CK_LAUNCH();
  return 0;
}

// dense SPD inverse on the device: cuSOLVER potrf + potri (dlopen, like NCCL: the library is needed only when the option is used)
static int dense_spd_inverse(femcy_ctx* ctx, Precond2* p, double* A, int n) {
  typedef int (*Create)(void**);
  typedef int (*SetStream)(void*, cudaStream_t);
  typedef int (*PotrfBuf)(void*, int, int, double*, int, int*);
  typedef int (*Potrf)(void*, int, int, double*, int, double*, int, int*);
  typedef int (*PotriBuf)(void*, int, int, double*, int, int*);
  typedef int (*Potri)(void*, int, int, double*, int, double*, int, int*);
  if (!p->solver_lib) {
    p->solver_lib = dlopen("libcusolver.so", RTLD_NOW | RTLD_GLOBAL);
    if (!p->solver_lib) p->solver_lib = dlopen("libcusolver.so.11", RTLD_NOW | RTLD_GLOBAL);
    if (!p->solver_lib) p->solver_lib = dlopen("/usr/local/cuda/lib64/libcusolver.so", RTLD_NOW | RTLD_GLOBAL);
    if (!p->solver_lib) return femcy_fail_msg(ctx, "cg_precond: libcusolver.so not found (needed for the coarse-level factorisation)");
  }
  Create create = (Create)dlsym(p->solver_lib, "cusolverDnCreate");
  SetStream set_stream = (SetStream)dlsym(p->solver_lib, "cusolverDnSetStream");
  PotrfBuf potrf_buf = (PotrfBuf)dlsym(p->solver_lib, "cusolverDnDpotrf_bufferSize");
  Potrf potrf = (Potrf)dlsym(p->solver_lib, "cusolverDnDpotrf");
  PotriBuf potri_buf = (PotriBuf)dlsym(p->solver_lib, "cusolverDnDpotri_bufferSize");
  Potri potri = (Potri)dlsym(p->solver_lib, "cusolverDnDpotri");
  if (!create || !set_stream || !potrf_buf || !potrf || !potri_buf || !potri) return femcy_fail_msg(ctx, "cg_precond: cuSOLVER symbols missing");
  if (!p->solver_handle && create(&p->solver_handle) != 0) return femcy_fail_msg(ctx, "cusolverDnCreate failed");
  if (set_stream(p->solver_handle, ctx->stream) != 0) return femcy_fail_msg(ctx, "cusolverDnSetStream failed");
  const int LOWER = 0;   // CUBLAS_FILL_MODE_LOWER (column-major view of the symmetric matrix)
  int l1 = 0, l2 = 0;
  if (potrf_buf(p->solver_handle, LOWER, n, A, n, &l1) != 0 || potri_buf(p->solver_handle, LOWER, n, A, n, &l2) != 0)
    return femcy_fail_msg(ctx, "cuSOLVER buffer query failed");
  const int64_t need = l1 > l2 ? l1 : l2;
  if (need > p->work_len) {
    if (p->work) cudaFree(p->work);
    CK(cudaMalloc(&p->work, (size_t)(need + 16) * sizeof(double)));
    p->work_len = need;
  }
  if (!p->info) CK(cudaMalloc((void**)&p->info, sizeof(int)));
  int st1 = potrf(p->solver_handle, LOWER, n, A, n, (double*)p->work, (int)p->work_len, p->info);
  int h_info = 0;
  CK(cudaMemcpyAsync(&h_info, p->info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (st1 != 0 || h_info != 0) return 2;        // not positive definite (NaN / indefinite K): the caller reports a breakdown
  int st2 = potri(p->solver_handle, LOWER, n, A, n, (double*)p->work, (int)p->work_len, p->info);
  CK(cudaMemcpyAsync(&h_info, p->info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (st2 != 0 || h_info != 0) return femcy_fail_msg(ctx, "cg_precond: potri failed");
  k_mirror_lower<<<(unsigned)ceil_div64((int64_t)n * n, 256), 256, 0, ctx->stream>>>(A, n);
  CK_LAUNCH();
  return 0;
}

static inline int vgrid(int64_t n) {
  int64_t g = ceil_div64(n, 256 * 4);
  if (g > 148 * 8) g = 148 * 8;
  return (int)(g < 1 ? 1 : g);
}

int femcy_spmv_plain(femcy_ctx* ctx, const double* x, double* y);   // cg.cu
int femcy_spmv_cg(femcy_ctx* ctx, const double* x, double* y);      // cg.cu: SpMV + d.Ad + alpha = S_RMR / d.Ad
int femcy_cg_set_scalars(femcy_ctx* ctx, double eps, int fixed_iters);

// PCG with the two-level preconditioner; same contract as femcy_cg_solve (called from it when the option is set)
int femcy_cg_solve_two_level(femcy_ctx* ctx, int b_sel, double eps, int64_t max_iter, int check_every, int fixed_iters,
                             int64_t* iters_out, double* rmax0_out, double* rmax_out) {
  BsellPattern& P = ctx->P;
  Precond2* p = static_cast<Precond2*>(ctx->precond2);
  if (!p || !p->agg || p->nn != ctx->nn || p->dm != ctx->dm)
    return femcy_fail_msg(ctx, "cg_precond = 1 needs femcy_set_aggregates for this mesh");
  if (femcy_comm_size(ctx) > 1) return femcy_fail_msg(ctx, "cg_precond = 1 (two-level preconditioner) is single-GPU only");
  cudaStream_t st = ctx->stream;
  const int dm = P.dm;
  const int NR = dm == 2 ? 3 : 6;
  const int64_t n = P.nn_own * dm, nc = p->nagg * NR;
  if (nc > 16384) return femcy_fail_msg(ctx, "cg_precond: at most 16384 coarse unknowns (dense coarse level)");
  if (p->vec_len != n) {
    if (femcy_alloc(ctx, &p->z, n) || femcy_alloc(ctx, &p->dv, n) || femcy_alloc(ctx, &p->tmp, n)) return 1;
    p->vec_len = n;
  }
  if (p->nc_alloc != nc) {
    if (femcy_alloc(ctx, &p->Ac, nc * nc) || femcy_alloc(ctx, &p->rc, nc) || femcy_alloc(ctx, &p->yc, nc)) return 1;
    p->nc_alloc = nc;
  }
  const double* b = ctx->vec[b_sel];
  double *x = ctx->vec[FEMCY_VEC_X], *r = ctx->vec[FEMCY_VEC_R], *d = ctx->vec[FEMCY_VEC_D], *Dinv = ctx->vec[FEMCY_VEC_M],
         *Ad = ctx->vec[FEMCY_VEC_AD];
  const int vg = vgrid(n);
  {
    int64_t spmv_grid = ceil_div64(P.nslice, 8);
    if (femcy_ensure_reduction_scratch(ctx, spmv_grid > vg ? spmv_grid : vg)) return 1;
  }
  CK(cudaMemsetAsync(ctx->red_ticket, 0, 8 * sizeof(unsigned int), st));
  if (femcy_cg_set_scalars(ctx, eps, fixed_iters)) return 1;
  k2_init<<<vg, 256, 0, st>>>(P.diag_slot, P.val, b, x, r, Dinv, P.nn_own, dm, ctx->red_partials, ctx->red_ticket, ctx->scal);
  CK_LAUNCH();
  CK(cudaEventRecord(ctx->ev0, st));

  // ---- setup 1: lambda_max(D^-1 A) by power iteration (8 steps, +10 %) -------------------------------------------------
  double lmax = 0.0;
  {
    k_seed<<<vg, 256, 0, st>>>(p->dv, n);
    CK_LAUNCH();
    double norms[3];
    for (int itp = 0; itp < 8; ++itp) {
      if (femcy_spmv_plain(ctx, p->dv, p->tmp)) return 1;
      k_mul<<<vg, 256, 0, st>>>(p->z, p->tmp, Dinv, n);          // z = D^-1 A v
      CK_LAUNCH();
      // max|z| / max|v| with v normalised to max 1 after the first step
      k2_absmax<<<vg, 256, 0, st>>>(p->z, n, ctx->red_partials, ctx->red_ticket + 1, ctx->scal + 40);
      CK_LAUNCH();
      k_scale_by<<<vg, 256, 0, st>>>(p->dv, p->tmp, Dinv, n, ctx->scal + 40);
      CK_LAUNCH();
      if (itp == 7) {
        CK(cudaMemcpyAsync(norms, ctx->scal + 40, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        lmax = norms[1];
      }
    }
    if (!(lmax > 0.0) || !(lmax < 1e300)) return femcy_fail_msg(ctx, "cg_precond: eigenvalue estimate failed (singular diagonal?)");
    lmax *= 1.1;
  }
  const double lmin = lmax / 30.0;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  const double rho0 = 1.0 / sigma, rho1 = 1.0 / (2.0 * sigma - rho0);
  const double c1 = rho1 * rho0, c2 = 2.0 * rho1 / delta;

  // ---- setup 2: A_c = P^T A P, inverted densely -----------------------------------------------------------------------
  CK(cudaMemsetAsync(p->Ac, 0, (size_t)(nc * nc) * sizeof(double), st));
  {
    const unsigned g = (unsigned)ceil_div64(P.nslice * 32, 256);
    if (dm == 2) k_coarse_matrix<2><<<g, 256, 0, st>>>(P.slice_ptr, P.colidx, P.val, P.nn_own, P.nslice, P.rowof, ctx->nodes, p->agg, p->cen, p->Ac, nc);
    else k_coarse_matrix<3><<<g, 256, 0, st>>>(P.slice_ptr, P.colidx, P.val, P.nn_own, P.nslice, P.rowof, ctx->nodes, p->agg, p->cen, p->Ac, nc);
    CK_LAUNCH();
    k_diag_max<<<1, 256, 0, st>>>(p->Ac, nc, ctx->scal + 46);
    CK_LAUNCH();
    k_coarse_fix<<<(unsigned)ceil_div64(nc * nc, 256), 256, 0, st>>>(p->Ac, nc, 1e-10, ctx->scal + 46);
    CK_LAUNCH();
    const int irc = dense_spd_inverse(ctx, p, p->Ac, (int)nc);
    if (irc == 2) {
      // K is not positive definite (typically NaN after a diverged Newton step): behave like the Jacobi path, whose
      // recurrence runs into NaN -- x = NaN, breakdown flag set, the Newton driver cuts the step (stiffnessMtrx.py:790-793)
      CK(cudaMemsetAsync(x, 0xFF, (size_t)n * sizeof(double), st));
      CK(cudaEventRecord(ctx->ev1, st));
      CK(cudaStreamSynchronize(st));
      ctx->cg_breakdown = true;
      if (iters_out) *iters_out = 0;
      if (rmax0_out) *rmax0_out = 0.0;
      if (rmax_out) *rmax_out = 0.0 / 0.0;
      return 0;
    }
    if (irc) return 1;
  }

  // ---- z = M^-1 r ------------------------------------------------------------------------------------------------------
  auto apply_precond = [&](int first) -> int {
    if (dm == 2) k_restrict<2><<<(unsigned)p->nagg, 128, 0, st>>>(r, ctx->nodes, p->agg_ptr, p->agg_nodes, p->cen, p->rc);
    else k_restrict<3><<<(unsigned)p->nagg, 128, 0, st>>>(r, ctx->nodes, p->agg_ptr, p->agg_nodes, p->cen, p->rc);
    CK_LAUNCH();
    k_dense_matvec<<<(unsigned)ceil_div64(nc * 32, 256), 256, 0, st>>>(p->Ac, p->rc, p->yc, nc);
    CK_LAUNCH();
    k_cheb_first<<<vg, 256, 0, st>>>(r, Dinv, p->dv, n, 1.0 / theta);
    CK_LAUNCH();
    if (femcy_spmv_plain(ctx, p->dv, p->tmp)) return 1;
    if (dm == 2) k_precond_finish<2><<<vg, 256, 0, st>>>(r, Dinv, p->dv, p->tmp, ctx->nodes, p->agg, p->cen, p->yc, p->z, P.nn_own, c1, c2, ctx->red_partials, ctx->red_ticket, ctx->scal, first);
    else k_precond_finish<3><<<vg, 256, 0, st>>>(r, Dinv, p->dv, p->tmp, ctx->nodes, p->agg, p->cen, p->yc, p->z, P.nn_own, c1, c2, ctx->red_partials, ctx->red_ticket, ctx->scal, first);
    CK_LAUNCH();
    k2_update_d<<<vg, 256, 0, st>>>(d, p->z, n, ctx->scal);
    CK_LAUNCH();
    return 0;
  };
  if (apply_precond(1)) return 1;

  auto iteration = [&]() -> int {
    if (femcy_spmv_cg(ctx, d, Ad)) return 1;
    k2_update_xr<<<vg, 256, 0, st>>>(x, r, d, Ad, n, ctx->red_partials, ctx->red_ticket, ctx->scal);
    CK_LAUNCH();
    return apply_precond(0);
  };

  // CUDA graph of `check_every` iterations (9 kernels each)
  cudaGraphExec_t gexec = nullptr;
  if (check_every > 1 && max_iter >= check_every) {
    cudaGraph_t graph = nullptr;
    int64_t l0 = ctx->launches;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int erc = 0;
      for (int c = 0; c < check_every && !erc; ++c) erc = iteration();
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (erc || ce != cudaSuccess || !graph || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) { cudaGetLastError(); gexec = nullptr; }
      if (graph) cudaGraphDestroy(graph);
    } else cudaGetLastError();
    ctx->launches = l0;
  }
  int64_t it = 0;
  bool done = false;
  int rc = 0;
  while (it < max_iter && !done && !rc) {
    int64_t chunk = check_every;
    if (it + chunk > max_iter) chunk = max_iter - it;
    if (gexec && chunk == check_every) {
      if (cudaGraphLaunch(gexec, st) != cudaSuccess) { rc = femcy_fail_msg(ctx, "graph launch"); break; }
      ctx->launches += 9 * chunk;
    } else {
      for (int64_t c = 0; c < chunk && !rc; ++c) rc = iteration();
    }
    it += chunk;
    if (!fixed_iters || it >= max_iter) {
      cudaError_t me = cudaMemcpyAsync(ctx->h_scal, ctx->scal, 16 * sizeof(double), cudaMemcpyDeviceToHost, st);
      if (me == cudaSuccess) me = cudaStreamSynchronize(st);
      if (me != cudaSuccess) { rc = femcy_fail(ctx, "PCG: reading the stop flag", me, __FILE__, __LINE__); break; }
      if (ctx->h_scal[S_DONE] != 0.0) done = true;
    }
  }
  if (gexec) cudaGraphExecDestroy(gexec);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev1, st));
  CK(cudaMemcpyAsync(ctx->h_scal, ctx->scal, 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->last_ms[1] = ms;
  for (int q = 0; q < S_PHASE_COUNT; ++q) ctx->cg_phase_ns[q] = 0.0;
  if (iters_out) *iters_out = (int64_t)ctx->h_scal[S_ITER];
  if (rmax0_out) *rmax0_out = ctx->h_scal[S_R0];
  if (rmax_out) *rmax_out = ctx->h_scal[S_RMAX];
  ctx->cg_breakdown = ctx->h_scal[S_DONE] == 2.0;
  return 0;
}
