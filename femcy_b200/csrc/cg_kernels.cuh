// Device code of the PCG path (rows a6 + a7).  Kept in a header so that the host launch code (cg.cu) and the
// CPU SIMT emulation used by the not-gpu kernel-logic tests (tests/simt, test infrastructure only) compile the
// same source.  Reference: ConjugateGradientSolver_rowMajor  /root/reference/conjugateGradientSolver.py:8-127
#pragma once
#include "device_compat.cuh"
#include "kernel_types.cuh"
#include "elem_math.cuh"

template <int DM>
__device__ __forceinline__ void bsell_row(const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ colidx,
                                          const double* __restrict__ val, const double* __restrict__ x, int64_t s,
                                          int lane, double (&acc)[DM], int ghost_from = 0x7fffffff) {
  constexpr int DM2 = DM * DM;
  int base = slice_ptr[s];
  int w = (slice_ptr[s + 1] - base) >> 5;
#pragma unroll
  for (int r = 0; r < DM; ++r) acc[r] = 0.0;
  const int32_t* ci = colidx + base + lane;
  const double* v = val + (((int64_t)(base >> 5) * DM2) << 5) + lane;
#pragma unroll 2
  for (int k = 0; k < w; ++k) {
    // matrix stream: read once per SpMV -> evict-first (ld.global.cs) so the 1.9 GB of values do not
    // push the vectors (d, Ad, r, M: reused by the next kernels) out of L2
    int c = __ldcs(ci + (k << 5));
    double a[DM2];
#pragma unroll
    for (int q = 0; q < DM2; ++q) a[q] = __ldcs(v + (((int64_t)k * DM2 + q) << 5));
    if (c >= 0) {
      double xv[DM];
      if (c >= ghost_from) {
        // ghost column (peer-memory path): written by another GPU during this kernel -> read at L2
        // (ld.global.cg); an L1 line brought in earlier by a neighbouring owned column could be stale
#pragma unroll
        for (int j = 0; j < DM; ++j) xv[j] = __ldcg(x + (int64_t)c * DM + j);
      } else {
#pragma unroll
        for (int j = 0; j < DM; ++j) xv[j] = x[(int64_t)c * DM + j];
      }
#pragma unroll
      for (int r = 0; r < DM; ++r)
#pragma unroll
        for (int j = 0; j < DM; ++j) acc[r] += a[r * DM + j] * xv[j];
    }
  }
}

// Row of the upper-half matrix (SymPattern; option cg_sym): lane = row i holds the blocks K_ij with j >= i.
// Each block is used twice: y_i += K_ij x_j and, for an owned off-diagonal column, y_j += K_ij^T x_i (fp64 atomics;
// y must be zero when the SpMV starts).  Half the matrix stream of bsell_row -- the SpMV is bound by it -- for
// 3 atomics per off-diagonal block; the summation order, hence the last bits of y, varies from run to run.
// Ghost columns (j >= n_own: rows of another rank, which stores block (j,i) itself) are not scattered to.
// Returns this row's share of x.y: x_i.(K_ii x_i + 2 sum_{owned j>i} K_ij x_j + sum_{ghost j} K_ij x_j).
template <int DM>
__device__ __forceinline__ double bsell_row_sym(const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ colidx,
                                                const double* __restrict__ val, const double* __restrict__ x,
                                                double* __restrict__ y, int64_t s, int lane, int n_own, bool ghost_l2,
                                                const int32_t* __restrict__ rowof = nullptr) {
  constexpr int DM2 = DM * DM;
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) >> 5;
  int i = (int)(s * 32 + lane);                     // position in the (sigma-sorted) row order -> row node
  if (rowof) { i = rowof[i]; if (i < 0) i = 0x7fffffff; }
  double xi[DM], acc[DM];
#pragma unroll
  for (int r = 0; r < DM; ++r) { xi[r] = (i < n_own) ? x[(int64_t)i * DM + r] : 0.0; acc[r] = 0.0; }
  double dot = 0.0;
  const int32_t* ci = colidx + base + lane;
  const double* v = val + (((int64_t)(base >> 5) * DM2) << 5) + lane;
#pragma unroll 2
  for (int k = 0; k < w; ++k) {
    int c = __ldcs(ci + (k << 5));
    double a[DM2];
#pragma unroll
    for (int q = 0; q < DM2; ++q) a[q] = __ldcs(v + (((int64_t)k * DM2 + q) << 5));
    if (c >= 0) {
      double xv[DM];
      if (ghost_l2 && c >= n_own) {
#pragma unroll
        for (int j = 0; j < DM; ++j) xv[j] = __ldcg(x + (int64_t)c * DM + j);
      } else {
#pragma unroll
        for (int j = 0; j < DM; ++j) xv[j] = x[(int64_t)c * DM + j];
      }
      double t[DM], xt = 0.0;
#pragma unroll
      for (int r = 0; r < DM; ++r) {
        t[r] = 0.0;
#pragma unroll
        for (int j = 0; j < DM; ++j) t[r] += a[r * DM + j] * xv[j];
        acc[r] += t[r];
        xt += xi[r] * t[r];
      }
      if (c != i && c < n_own) {
#pragma unroll
        for (int j = 0; j < DM; ++j) {
          double u = 0.0;
#pragma unroll
          for (int r = 0; r < DM; ++r) u += a[r * DM + j] * xi[r];
          femcy_red_add_f64(y + (int64_t)c * DM + j, u);
        }
        dot += 2.0 * xt;
      } else {
        dot += xt;
      }
    }
  }
  if (i < n_own) {
#pragma unroll
    for (int r = 0; r < DM; ++r) femcy_red_add_f64(y + (int64_t)i * DM + r, acc[r]);
  }
  return dot;
}

// ---- peer-memory exchange (multi == 2) ------------------------------------------------------------
// Called by ONE thread (thread 0 of the last block of a kernel).  Every double is sent to every rank's
// window as two 8-byte words {half of the value | 32-bit tag of this exchange} with plain relaxed
// system-scope stores (NVLink peer stores for the other ranks); the reader polls its own window until both
// halves of every rank's value carry the tag.  No fence, no separate flag: one NVLink one-way latency.
// Contributions are returned in rank order.  The spin is bounded (a lost peer must not hang the GPU).
// (st_sys_u64 / ld_sys_u64 / ld_acquire_sys_u64: device_compat.cuh)

template <int NV_>
__device__ __forceinline__ bool p2p_allgather(const P2PView& pv, int which, const double (&mine)[NV_],
                                              unsigned long long seq1, double (&all)[FEMCY_MAX_RANKS][NV_]) {
  const unsigned long long tag = seq1 & 0xffffffffull;
  const int base = (which == 0) ? P2P_SLOT_A(pv.rank) : P2P_SLOT_B(pv.rank);
  unsigned long long w[NV_][2];
#pragma unroll
  for (int i = 0; i < NV_; ++i) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(mine[i]);
    w[i][0] = (bits << 32) | tag;                          // low half of the value | tag
    w[i][1] = (bits & 0xffffffff00000000ull) | tag;        // high half of the value | tag
  }
  for (int r = 0; r < pv.nranks; ++r)
#pragma unroll
    for (int i = 0; i < NV_; ++i) {
      st_sys_u64(pv.win_of[r] + base + 2 * i, w[i][0]);
      st_sys_u64(pv.win_of[r] + base + 2 * i + 1, w[i][1]);
    }
  bool ok = true;
  for (int r = 0; r < pv.nranks; ++r) {
    const unsigned long long* src = pv.win_of[pv.rank] + ((which == 0) ? P2P_SLOT_A(r) : P2P_SLOT_B(r));
#pragma unroll
    for (int i = 0; i < NV_; ++i) {
      unsigned long long a = 0, b = 0;
      long long spins = 0;
      for (;;) {
        a = ld_sys_u64(src + 2 * i);
        b = ld_sys_u64(src + 2 * i + 1);
        if ((a & 0xffffffffull) == tag && (b & 0xffffffffull) == tag) break;
        if (++spins > (1ll << 24)) { ok = false; break; }
        FEMCY_SPIN_PAUSE();
      }
      all[r][i] = __longlong_as_double((long long)((b & 0xffffffff00000000ull) | (a >> 32)));
    }
  }
  return ok;
}

// y = A x ; optional fused dot(x_own, y).  One warp per slice.
template <int DM, bool P2P>
__global__ void __launch_bounds__(256)
k_spmv_dot(const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ colidx, const double* __restrict__ val,
           const double* __restrict__ x, double* __restrict__ y, int64_t nrows, int64_t nslice, double* partials,
           unsigned int* ticket, double* scal, int cg_mode, int multi, const __grid_constant__ P2PView pv,
           const int32_t* __restrict__ slice_order, const unsigned char* __restrict__ slice_ghost,
           const int32_t* __restrict__ rowof) {
  if (cg_mode && scal[S_DONE] != 0.0) return;
  int lane = threadIdx.x & 31;
  int64_t s = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  double dot = 0.0;
  if (s < nslice) {
    if (P2P) {
      // peer-memory path: slices that read no ghost column come first in launch order; a slice that does
      // waits (in-kernel) until every rank has published the halo push of the previous update_d, so the
      // NVLink flight time of the boundary values hides behind the interior rows.
      s = slice_order[s];
      if (slice_ghost[s]) {
        unsigned long long want = (unsigned long long)scal[S_SEQ];
        const unsigned long long* myflags = pv.win_of[pv.rank] + P2P_FLAG_D(0);
        if (lane < pv.nranks) {
          long long spins = 0;
          while (ld_acquire_sys_u64(myflags + lane) < want) {   // acquire: the ghost loads below stay behind it
            if (++spins > (1ll << 24)) { scal[S_DONE] = 3.0; break; }
            FEMCY_SPIN_PAUSE();
          }
        }
        __syncwarp();
      }
    }
    double acc[DM];
    bsell_row<DM>(slice_ptr, colidx, val, x, s, lane, acc, P2P ? (int)nrows : 0x7fffffff);
    int64_t i = s * 32 + lane;      // position in the (sigma-sorted) row order; rowof == nullptr: position == row
    if (i < nrows) {
      if (rowof) i = rowof[i];
#pragma unroll
      for (int r = 0; r < DM; ++r) {
        y[i * DM + r] = acc[r];
        dot += acc[r] * x[i * DM + r];
      }
    }
  }
  if (!cg_mode) return;
  double mine[1] = {dot}, tot[1];
  const bool is_max[1] = {false};
  if (grid_reduce<1>(mine, partials, ticket, tot, is_max)) {
    if (P2P) {
      double all[FEMCY_MAX_RANKS][1];
      bool ok = p2p_allgather<1>(pv, 0, tot, (unsigned long long)scal[S_SEQ] + 1ull, all);
      double t = 0.0;
      for (int r = 0; r < pv.nranks; ++r) t += all[r][0];       // rank order: identical on every rank
      scal[S_DAD] = t;
      scal[S_ALPHA] = scal[S_RMR] / t;
      if (!(t > 0.0)) scal[S_DONE] = 2.0;          // d.Ad <= 0 (or NaN): K is not positive definite -- CG breaks down
      if (!ok) scal[S_DONE] = 3.0;
    } else if (multi) {
      scal[S_SEND] = tot[0];
    } else {
      scal[S_DAD] = tot[0];
      scal[S_ALPHA] = scal[S_RMR] / tot[0];
      if (!(tot[0] > 0.0)) scal[S_DONE] = 2.0;     // d.Ad <= 0 (or NaN): K is not positive definite -- CG breaks down
    }
  }
}

// multi-GPU: fold the all-gathered partial d.Ad in rank order
__global__ void k_finish_alpha(double* scal, int nranks) {
  if (scal[S_DONE] != 0.0) return;
  double t = 0.0;
  for (int r = 0; r < nranks; ++r) t += scal[S_GATHER + r];
  scal[S_DAD] = t;
  scal[S_ALPHA] = scal[S_RMR] / t;
  if (!(t > 0.0)) scal[S_DONE] = 2.0;              // d.Ad <= 0 (or NaN): K is not positive definite -- CG breaks down
}

__device__ __forceinline__ void finish_beta(double* scal, double rmr_new, double rmax) {
  double rmr = scal[S_RMR];
  scal[S_BETA] = rmr_new / rmr;
  scal[S_RMR] = rmr_new;
  scal[S_RMAX] = rmax;
  double it = scal[S_ITER] + 1.0;
  scal[S_ITER] = it;
  if (scal[S_FIXED] == 0.0 && (rmax < scal[S_EPS] * scal[S_R0])) scal[S_DONE] = 1.0;  // :124
  if (!(rmax < 1.0e300) || rmr_new != rmr_new) scal[S_DONE] = 2.0;                   // NaN/inf: stop
}

__global__ void __launch_bounds__(256)
k_update_xr(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ d, const double* __restrict__ Ad,
            const double* __restrict__ M, int64_t n, double* partials, unsigned int* ticket, double* scal, int multi,
            const __grid_constant__ P2PView pv) {
  if (scal[S_DONE] != 0.0) return;
  double alpha = scal[S_ALPHA];
  double rmr = 0.0, rmax = 0.0;
  // 16-byte vector accesses (the vectors are 256 B aligned), x streamed with evict-first hints: it is
  // touched once per iteration, while r, M, d are re-read by update_d right after
  const int64_t n2 = n >> 1;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  double2* x2 = reinterpret_cast<double2*>(x);
  double2* r2 = reinterpret_cast<double2*>(r);
  const double2* d2 = reinterpret_cast<const double2*>(d);
  const double2* A2 = reinterpret_cast<const double2*>(Ad);
  const double2* M2 = reinterpret_cast<const double2*>(M);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += gs) {
    double2 xv = __ldcs(x2 + i), dv = d2[i], rv = r2[i], av = __ldcs(A2 + i), mv = M2[i];
    xv.x = xv.x + alpha * dv.x; xv.y = xv.y + alpha * dv.y;
    rv.x = rv.x - alpha * av.x; rv.y = rv.y - alpha * av.y;
    __stcs(x2 + i, xv);
    r2[i] = rv;
    rmr += rv.x * mv.x * rv.x;
    rmr += rv.y * mv.y * rv.y;
    rmax = fmax(rmax, fmax(fabs(rv.x), fabs(rv.y)));
    if (rv.x != rv.x || rv.y != rv.y) rmax = 1.0 / 0.0;  // NaN in r: force the stop flag through an inf max
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t i = n - 1;
    x[i] = x[i] + alpha * d[i];
    double rn = r[i] - alpha * Ad[i];
    r[i] = rn;
    rmr += rn * M[i] * rn;
    rmax = fmax(rmax, fabs(rn));
    if (rn != rn) rmax = 1.0 / 0.0;
  }
  double mine[2] = {rmr, rmax}, tot[2];
  const bool is_max[2] = {false, true};
  if (grid_reduce<2>(mine, partials, ticket, tot, is_max)) {
    if (multi == 2) {
      double all[FEMCY_MAX_RANKS][2];
      bool ok = p2p_allgather<2>(pv, 1, tot, (unsigned long long)scal[S_SEQ] + 1ull, all);
      double t = 0.0, m = 0.0;
      for (int r = 0; r < pv.nranks; ++r) { t += all[r][0]; m = fmax(m, all[r][1]); }
      finish_beta(scal, t, m);
      if (!ok) scal[S_DONE] = 3.0;
    } else if (multi) { scal[S_SEND] = tot[0]; scal[S_SEND + 1] = tot[1]; }
    else finish_beta(scal, tot[0], tot[1]);
  }
}

__global__ void k_finish_beta(double* scal, int nranks) {
  if (scal[S_DONE] != 0.0) return;
  double t = 0.0, m = 0.0;
  for (int r = 0; r < nranks; ++r) { t += scal[S_GATHER + 2 * r]; m = fmax(m, scal[S_GATHER + 2 * r + 1]); }
  finish_beta(scal, t, m);
}

__global__ void __launch_bounds__(256)
k_update_d(double* __restrict__ d, const double* __restrict__ r, const double* __restrict__ M, int64_t n,
           const double* __restrict__ scal) {
  if (scal[S_DONE] != 0.0) return;
  double beta = scal[S_BETA];
  const int64_t n2 = n >> 1;
  double2* d2 = reinterpret_cast<double2*>(d);
  const double2* r2 = reinterpret_cast<const double2*>(r);
  const double2* M2 = reinterpret_cast<const double2*>(M);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    double2 dv = d2[i], rv = r2[i], mv = M2[i];
    dv.x = mv.x * rv.x + beta * dv.x;
    dv.y = mv.y * rv.y + beta * dv.y;
    d2[i] = dv;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) d[n - 1] = M[n - 1] * r[n - 1] + beta * d[n - 1];
}

// update_d fused with the halo push (multi == 2).  d = M r + beta d   (conjugateGradientSolver.py:91-94)
//   phase 1  the boundary entries (compact list push_dof) are updated FIRST and stored into the ghost
//            slots of every rank holding a copy (NVLink peer stores); the last block through ticket 1
//            publishes flag D, so the values travel while phase 2 runs;
//   phase 2  all other entries (boundary nodes skipped via bflag);
//   tail     the last block through ticket 2 advances the exchange counter; nobody waits here -- the next
//            SpMV's boundary slices poll flag D themselves (k_spmv_dot).
template <int DM>
__global__ void __launch_bounds__(256)
k_update_d_p2p(double* __restrict__ d, const double* __restrict__ r, const double* __restrict__ M, int n,
               double* scal, const __grid_constant__ P2PView pv, const unsigned char* __restrict__ bflag,
               const int32_t* __restrict__ push_ptr, const int32_t* __restrict__ push_peer,
               const int32_t* __restrict__ push_ridx, const int32_t* __restrict__ bnodes, int n_bnodes,
               unsigned int* tickets) {
  if (scal[S_DONE] != 0.0) return;
  double beta = scal[S_BETA];
  __shared__ bool last1, last2;
  bool pushed = false;
  const int stride = gridDim.x * blockDim.x;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_bnodes * DM; t += stride) {
    int k = t / DM;
    int c = t - k * DM;
    int node = bnodes[k];
    int i = node * DM + c;
    double dn = M[i] * r[i] + beta * d[i];
    d[i] = dn;
    for (int e = push_ptr[node]; e < push_ptr[node + 1]; ++e)
      pv.d_of[push_peer[e]][(int64_t)push_ridx[e] * DM + c] = dn;
    pushed = true;
  }
  if (pushed) __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last1 = (atomicAdd(&tickets[0], 1u) == gridDim.x - 1);
  __syncthreads();
  unsigned long long seq1 = (unsigned long long)scal[S_SEQ] + 1ull;
  if (last1 && threadIdx.x == 0) {
    __threadfence_system();        // (release pattern at system scope, see k_cg_stream)
    for (int rk = 0; rk < pv.nranks; ++rk) st_sys_u64(pv.win_of[rk] + P2P_FLAG_D(pv.rank), seq1);
    tickets[0] = 0;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (bflag[i / DM]) continue;
    d[i] = M[i] * r[i] + beta * d[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) last2 = (atomicAdd(&tickets[1], 1u) == gridDim.x - 1);
  __syncthreads();
  if (last2 && threadIdx.x == 0) {
    scal[S_SEQ] = (double)seq1;     // the next SpMV's boundary slices wait for flag D >= this value
    tickets[1] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent cooperative CG: ONE kernel runs `iters` whole iterations (SpMV + dot, x/r update + rMr/max|r|,
// d update + halo push) with grid-wide barriers instead of kernel boundaries, and -- on several GPUs --
// exchanges the partial sums and the halo through NVLink peer memory from inside the same kernel.  At 8 GPUs
// an iteration is ~55 us of work: three kernel launches + three reduction tails cost almost as much as the
// work (profiles/r1h_scaling.md); here the per-iteration overhead is three grid.sync() + two window polls.
// Arithmetic and operation order per entry are those of the three-kernel path; the block-partial folds are
// done redundantly by every block in one fixed order, so all blocks (and all ranks) hold identical scalars.
struct CGPersistArgs {
  const int32_t* slice_ptr; const int32_t* colidx; const double* val;
  int64_t nrows, nslice;
  double *x, *r, *d, *Ad; const double* M; int64_t n;
  double* part1;   // [grid]     d.Ad block partials
  double* part2;   // [grid*2]   rMr / max|r| block partials
  double* scal;
  int iters, p2p;
  P2PView pv;
  const unsigned char* bflag; const int32_t *push_ptr, *push_peer, *push_ridx, *bnodes; int n_bnodes;
  const int32_t* slice_order; const unsigned char* slice_ghost;
  unsigned int* ticket;
  const int4* bpush = nullptr;      // [n_bnodes] {node, first peer, first remote index, further push entries} (k_cg_stream)
  const int32_t* rowof = nullptr;   // sigma-sorted SELL: position -> row node (nullptr: identity)
  // option cg_sym: SpMV over the upper half of the matrix (bsell_row_sym); Ad is zero on entry and re-zeroed in P2
  int sym = 0; const int32_t* u_slice_ptr = nullptr; const int32_t* u_colidx = nullptr; const double* u_val = nullptr;
};

// every block calls this after a grid.sync(): fixed-order fold of `nb` block partials (stride NVs) with all
// 256 threads -- thread t sums partials t, t+256, ... (loads issued together), then a fixed shared-memory
// tree.  Same operations in the same order in every block => identical result everywhere.
// (A 32-lane fold walks ~28 dependent L2 round trips per lane at nb = 888: ~8 us, twice per iteration.)
template <int NVs>
__device__ __forceinline__ void fold_partials(const double* part, int nb, double (&out)[NVs], const bool (&is_max)[NVs],
                                              double (*sh)[256] /*[NVs][256]*/) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NVs; ++i) {
    double p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int b = t + u * 256;
      p[u] = (b < nb) ? __ldcg(part + (int64_t)b * NVs + i) : 0.0;
    }
    double acc = is_max[i] ? fmax(fmax(p[0], p[1]), fmax(p[2], p[3])) : ((p[0] + p[1]) + (p[2] + p[3]));
    for (int b = t + 1024; b < nb; b += 256) {
      double q = __ldcg(part + (int64_t)b * NVs + i);
      acc = is_max[i] ? fmax(acc, q) : acc + q;
    }
    sh[i][t] = acc;
  }
  __syncthreads();
  for (int s2 = 128; s2 > 0; s2 >>= 1) {
    if (t < s2) {
#pragma unroll
      for (int i = 0; i < NVs; ++i) sh[i][t] = is_max[i] ? fmax(sh[i][t], sh[i][t + s2]) : sh[i][t] + sh[i][t + s2];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < NVs; ++i) out[i] = sh[i][0];
  __syncthreads();
}

// the cross-rank part, called by every thread of every block: block 0 publishes this rank's values, every block polls
// its own window (local L2 reads) and folds the ranks in order.  One LANE per (rank, value) pair publishes / polls, so
// an exchange costs one L2 round trip after the data has arrived instead of nranks*NV_ dependent ones (a single
// thread walking 8 ranks x 3 values spent ~5 us per iteration at 8 GPUs); thread 0 then folds in rank order, so the
// totals are bitwise those of the sequential walk and identical on every block and rank.
// which: 0 = slot A (1 value), 1 = slot B (<= 2 values), 2 = A + B (3 values: value 0 in A, values 1, 2 in B).
// Returns false on a spin timeout.
__device__ __forceinline__ int p2p_word(int which, int rank, int i) {
  if (which == 0) return P2P_SLOT_A(rank) + 2 * i;
  if (which == 1) return P2P_SLOT_B(rank) + 2 * i;
  return i == 0 ? P2P_SLOT_A(rank) : P2P_SLOT_B(rank) + 2 * (i - 1);
}

template <int NV_>
__device__ __forceinline__ bool p2p_exchange_all_blocks(const P2PView& pv, int which, const double (&mine)[NV_],
                                                        unsigned long long seq1, double (&tot)[NV_],
                                                        const bool (&is_max)[NV_], const double* err_flag) {
  __shared__ double xv_s[FEMCY_MAX_RANKS * NV_];
  __shared__ int xok_s[FEMCY_MAX_RANKS * NV_];
  __shared__ double xt_s[NV_];
  __shared__ int ok_s;
  const int t = threadIdx.x;
  const unsigned long long tag = seq1 & 0xffffffffull;
  if (t < pv.nranks * NV_) {
    const int rk = t / NV_, i = t - rk * NV_;
    if (blockIdx.x == 0) {
      // NV_ is tiny: select mine[i] without dynamic register indexing
      double mv = mine[0];
#pragma unroll
      for (int u = 1; u < NV_; ++u) mv = (i == u) ? mine[u] : mv;
      unsigned long long bits = (unsigned long long)__double_as_longlong(mv);
      unsigned long long* dst = pv.win_of[rk] + p2p_word(which, pv.rank, i);
      st_sys_u64(dst, (bits << 32) | tag);
      st_sys_u64(dst + 1, (bits & 0xffffffff00000000ull) | tag);
    }
    const unsigned long long* src = pv.win_of[pv.rank] + p2p_word(which, rk, i);
    unsigned long long a = 0, b = 0;
    long long spins = 0;
    int ok = 1;
    // after one timeout (S_ERR set by the caller) every later exchange of the launch gives up after a short wait, so
    // a lost peer costs one bounded spin, not one per remaining iteration
    const long long limit = (*reinterpret_cast<const volatile double*>(err_flag) != 0.0) ? (1ll << 10) : (1ll << 24);
    for (;;) {
      a = ld_sys_u64(src);
      b = ld_sys_u64(src + 1);
      if ((a & 0xffffffffull) == tag && (b & 0xffffffffull) == tag) break;
      if (++spins > limit) { ok = 0; break; }
      FEMCY_SPIN_PAUSE();
    }
    xv_s[t] = __longlong_as_double((long long)((b & 0xffffffff00000000ull) | (a >> 32)));
    xok_s[t] = ok;
  }
  __syncthreads();
  if (t == 0) {
    int ok = 1;
#pragma unroll
    for (int i = 0; i < NV_; ++i) {
      double acc = 0.0;
      for (int rk = 0; rk < pv.nranks; ++rk) {
        double v = xv_s[rk * NV_ + i];
        acc = is_max[i] ? fmax(acc, v) : acc + v;
        ok &= xok_s[rk * NV_ + i];
      }
      xt_s[i] = acc;
    }
    ok_s = ok;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV_; ++i) tot[i] = xt_s[i];
  bool ok = ok_s != 0;
  __syncthreads();
  return ok;
}

// MINB = blocks per SM the kernel is compiled for: 6 -> 40 registers (64-180 B of spills), 5 -> 48 registers, no spills
// (measured on 1 and 8 GPUs, profiles/r2a + r2d: 6 blocks/SM wins)
// SYM = the cg_sym variant (upper-half SpMV with transposed scatter): its own instantiation, so that the default
// kernel's register allocation is untouched
template <int DM, int MINB = 6, bool SYM = false>
__global__ void __launch_bounds__(256, MINB)
k_cg_persistent(const __grid_constant__ CGPersistArgs a) {
  namespace cgx = cooperative_groups;
  cgx::grid_group grid = cgx::this_grid();
  __shared__ double shf[2][256];
  __shared__ double shw[2][8];
  double* scal = a.scal;
  if (scal[S_DONE] != 0.0) return;                 // stable during this launch: set only by earlier launches
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nb = gridDim.x;
  const int64_t gw = (int64_t)blockIdx.x * 8 + wib, nwarps = (int64_t)nb * 8;
  const int64_t gs = (int64_t)nb * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double rmr = scal[S_RMR];
  const double eps = scal[S_EPS], r0 = scal[S_R0];
  const bool fixed = scal[S_FIXED] != 0.0;
  unsigned long long seq = (unsigned long long)scal[S_SEQ];
  double it_count = scal[S_ITER];
  int done = 0;
  double alpha = 0.0, beta = 0.0, dAd = 0.0, rmax_g = 0.0;
  const bool clk = (blockIdx.x == 0 && threadIdx.x == 0);
  __shared__ unsigned long long ph[S_PHASE_COUNT + 1];     // phase sums + the last stamp: shared memory, not registers
  if (clk) {
    for (int q = 0; q < S_PHASE_COUNT; ++q) ph[q] = 0ull;
    ph[S_PHASE_COUNT] = femcy_globaltimer();
  }
  auto stamp = [&](int slot) {
    if (clk) { unsigned long long t = femcy_globaltimer(); ph[slot] += t - ph[S_PHASE_COUNT]; ph[S_PHASE_COUNT] = t; }
  };

  // stop decisions: taken by block 0, read by every block after the next grid barrier (see k_cg_stream)
  double* stop_word = a.part1 + 3 * (int64_t)nb;
  if (blockIdx.x == 0 && threadIdx.x == 0) *stop_word = 0.0;
  for (int it = 0; it < a.iters; ++it) {
    // ---- P1: Ad = A d, partial d.Ad --------------------------------------------------------------
    double dot = 0.0;
    for (int64_t sidx = gw; sidx < a.nslice; sidx += nwarps) {
      int64_t s = a.p2p ? a.slice_order[sidx] : sidx;
      if (a.p2p && a.slice_ghost[s]) {
        const unsigned long long* myflags = a.pv.win_of[a.pv.rank] + P2P_FLAG_D(0);
        if (lane < a.pv.nranks) {
          long long spins = 0;
          const long long limit = (*reinterpret_cast<const volatile double*>(scal + S_ERR) != 0.0) ? (1ll << 10) : (1ll << 24);
          while (ld_acquire_sys_u64(myflags + lane) < seq) {
            if (++spins > limit) { scal[S_ERR] = 3.0; break; }   // never changes control flow (grid.sync!)
            FEMCY_SPIN_PAUSE();
          }
        }
        __syncwarp();
      }
      if constexpr (SYM) {
        dot += bsell_row_sym<DM>(a.u_slice_ptr, a.u_colidx, a.u_val, a.d, a.Ad, s, lane, (int)a.nrows, a.p2p != 0, a.rowof);
        continue;
      }
      double acc[DM];
      bsell_row<DM>(a.slice_ptr, a.colidx, a.val, a.d, s, lane, acc, a.p2p ? (int)a.nrows : 0x7fffffff);
      int64_t i = s * 32 + lane;
      if (i < a.nrows) {
        if (a.rowof) i = a.rowof[i];
#pragma unroll
        for (int rr = 0; rr < DM; ++rr) {
          a.Ad[i * DM + rr] = acc[rr];
          dot += acc[rr] * a.d[i * DM + rr];
        }
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) shw[0][wib] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double b = 0.0;
      for (int j = 0; j < 8; ++j) b += shw[0][j];
      a.part1[blockIdx.x] = b;
    }
    stamp(0);
    {
      double loc[1], tot[1];
      const bool im[1] = {false};
      grid.sync();
      const double stop_now = __ldcg(stop_word);
      fold_partials<1>(a.part1, nb, loc, im, shf);
      stamp(1);
      if (stop_now != 0.0) { done = (int)stop_now; break; }
      if (a.p2p) { if (!p2p_exchange_all_blocks<1>(a.pv, 0, loc, seq + 1ull, tot, im, scal + S_ERR)) scal[S_ERR] = 3.0; }
      else tot[0] = loc[0];
      stamp(2);
      dAd = tot[0];
      alpha = rmr / dAd;
    }
    if (!(dAd > 0.0)) done = 2;      // K not positive definite: CG breaks down (acted upon through the stop word)
    // ---- P2: x += alpha d ; r -= alpha Ad ; partial r.M.r, max|r| ---------------------------------
    double prmr = 0.0, prmax = 0.0;
    {
      const int64_t n2 = a.n >> 1;
      double2* x2 = reinterpret_cast<double2*>(a.x);
      double2* r2 = reinterpret_cast<double2*>(a.r);
      const double2* d2 = reinterpret_cast<const double2*>(a.d);
      const double2* A2 = reinterpret_cast<const double2*>(a.Ad);
      const double2* M2 = reinterpret_cast<const double2*>(a.M);
      for (int64_t i = tid; i < n2; i += gs) {
        double2 xv = x2[i], dv = d2[i], rv = r2[i], av = A2[i], mv = M2[i];
        xv.x = xv.x + alpha * dv.x; xv.y = xv.y + alpha * dv.y;
        rv.x = rv.x - alpha * av.x; rv.y = rv.y - alpha * av.y;
        x2[i] = xv;
        r2[i] = rv;
        if constexpr (SYM) reinterpret_cast<double2*>(a.Ad)[i] = make_double2(0.0, 0.0);   // the next SpMV accumulates into Ad
        prmr += rv.x * mv.x * rv.x;
        prmr += rv.y * mv.y * rv.y;
        prmax = fmax(prmax, fmax(fabs(rv.x), fabs(rv.y)));
        if (rv.x != rv.x || rv.y != rv.y) prmax = 1.0 / 0.0;
      }
      if ((a.n & 1) && tid == 0) {
        int64_t i = a.n - 1;
        a.x[i] = a.x[i] + alpha * a.d[i];
        double rn = a.r[i] - alpha * a.Ad[i];
        a.r[i] = rn;
        if constexpr (SYM) a.Ad[i] = 0.0;
        prmr += rn * a.M[i] * rn;
        prmax = fmax(prmax, fabs(rn));
        if (rn != rn) prmax = 1.0 / 0.0;
      }
    }
    prmr = warp_sum(prmr);
    prmax = warp_max(prmax);
    if (lane == 0) { shw[0][wib] = prmr; shw[1][wib] = prmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double b0 = 0.0, b1 = 0.0;
      for (int j = 0; j < 8; ++j) { b0 += shw[0][j]; b1 = fmax(b1, shw[1][j]); }
      a.part2[blockIdx.x * 2] = b0;
      a.part2[blockIdx.x * 2 + 1] = b1;
    }
    stamp(3);
    {
      double loc[2], tot[2];
      const bool im[2] = {false, true};
      grid.sync();
      fold_partials<2>(a.part2, nb, loc, im, shf);
      stamp(4);
      if (a.p2p) { if (!p2p_exchange_all_blocks<2>(a.pv, 1, loc, seq + 1ull, tot, im, scal + S_ERR)) scal[S_ERR] = 3.0; }
      else { tot[0] = loc[0]; tot[1] = loc[1]; }
      stamp(5);
      beta = tot[0] / rmr;
      rmr = tot[0];
      rmax_g = tot[1];
      it_count += 1.0;
      if (!fixed && rmax_g < eps * r0) done = done ? done : 1;                 // conjugateGradientSolver.py:124
      if (!(rmax_g < 1.0e300) || rmr != rmr) done = 2;
    }
    if (done && blockIdx.x == 0 && threadIdx.x == 0) *stop_word = (double)done;
    // ---- P3: d = M r + beta d (boundary entries first, pushed to the neighbours' ghost slots) -----
    bool pushed = false;
    if (a.p2p) {
      for (int64_t t = tid; t < (int64_t)a.n_bnodes * DM; t += gs) {
        int k = (int)(t / DM);
        int c = (int)(t - (int64_t)k * DM);
        int node = a.bnodes[k];
        int64_t i = (int64_t)node * DM + c;
        double dn = a.M[i] * a.r[i] + beta * a.d[i];
        a.d[i] = dn;
        for (int e = a.push_ptr[node]; e < a.push_ptr[node + 1]; ++e)
          a.pv.d_of[a.push_peer[e]][(int64_t)a.push_ridx[e] * DM + c] = dn;
        pushed = true;
      }
    }
    if (a.p2p) {
      // publish the halo flag as soon as every block's pushes are fenced (ticket), before the interior
      // entries: the values travel while the rest of update_d runs
      if (pushed) __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0 && atomicAdd(a.ticket, 1u) == (unsigned)nb - 1u) {
        __threadfence_system();      // (release pattern at system scope, see k_cg_stream)
        for (int rk = 0; rk < a.pv.nranks; ++rk) st_sys_u64(a.pv.win_of[rk] + P2P_FLAG_D(a.pv.rank), seq + 1ull);
        *a.ticket = 0;
      }
    }
    for (int64_t i = tid; i < a.n; i += gs) {
      if (a.p2p && a.bflag[i / DM]) continue;
      a.d[i] = a.M[i] * a.r[i] + beta * a.d[i];
    }
    grid.sync();
    stamp(6);
    seq += 1ull;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int q = 0; q < S_PHASE_COUNT; ++q) scal[S_PHASE + q] += (double)ph[q];
    scal[S_RMR] = rmr; scal[S_ALPHA] = alpha; scal[S_BETA] = beta; scal[S_DAD] = dAd; scal[S_RMAX] = rmax_g;
    scal[S_ITER] = it_count; scal[S_SEQ] = (double)seq;
    if (done) scal[S_DONE] = (double)done;
    if (scal[S_ERR] != 0.0) scal[S_DONE] = 3.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming persistent PCG (the default solver kernel): the iteration of k_cg_persistent -- same recurrence, same
// per-entry operations -- with the MATRIX STREAM STAGED THROUGH SHARED MEMORY BY THE TMA ENGINE.
//
// Why: the SpMV is a pure stream of the matrix (1.95 GB per iteration on cfg 4, read once, no reuse).  With plain loads the
// bytes in flight are bounded by registers (a lane can keep ~one block column = 76 B outstanding at 40 registers), so the
// kernel needs 48 resident warps per SM to cover the HBM latency and, on a partitioned system where a warp owns a single
// 32-row slice, the whole SpMV becomes a chain of ~10 dependent DRAM round trips (phase clock, profiles/r2f: 58 us for
// 252 MB on one GPU = 4.3 TB/s).  Here every warp owns a ring of NS shared-memory stages; one lane issues
// `cp.async.bulk.shared::cta.global` copies (UBLKCP) of KC block columns of its slice per stage -- values [k][dm*dm][32]
// and column indices [k][32] are contiguous in the SELL-32 layout -- completing on the stage's mbarrier, with an L2
// evict-first hint.  The ring runs AHEAD OF THE GRID BARRIERS: the matrix does not depend on the vectors, so while the
// vector phases and the barriers of iteration i run, the first stages of iteration i+1 are already in flight.
// One block of NW warps per SM (148 partials to fold instead of 888, cheaper grid barriers).
//
// The per-warp chunk sequence is cyclic: slices sidx = gw, gw + nwarps, ... in KC-column chunks, then again from gw.
template <int DM, int NW, int KC, int NS>
struct CGStreamCfg {
  static constexpr int DM2 = DM * DM;
  static constexpr int STAGE_VALS = KC * DM2 * 32;                       // doubles
  static constexpr int STAGE_BYTES = STAGE_VALS * 8 + KC * 32 * 4;       // + column indices
  static constexpr int SMEM_BYTES = NW * NS * STAGE_BYTES;
};

// fold of <= a few hundred block partials by warp 0 (fixed order: lane l sums partials l, l+32, ..., then a butterfly):
// identical result in every block
template <int NVs>
__device__ __forceinline__ void fold_small(const double* part, int nb, double (&out)[NVs], const bool (&is_max)[NVs], double* sh) {
  if (threadIdx.x < 32) {
#pragma unroll
    for (int i = 0; i < NVs; ++i) {
      double acc = 0.0;
      for (int b = threadIdx.x; b < nb; b += 32) {
        double q = __ldcg(part + (int64_t)b * NVs + i);
        acc = is_max[i] ? fmax(acc, q) : acc + q;
      }
      acc = is_max[i] ? warp_max(acc) : warp_sum(acc);
      if (threadIdx.x == 0) sh[i] = acc;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NVs; ++i) out[i] = sh[i];
  __syncthreads();
}

template <int DM, int NW, int KC, int NS, bool SYM>
__global__ void __launch_bounds__(NW * 32, 1)
k_cg_stream(const __grid_constant__ CGPersistArgs a) {
  namespace cgx = cooperative_groups;
  using Cfg = CGStreamCfg<DM, NW, KC, NS>;
  constexpr int DM2 = DM * DM;
  cgx::grid_group grid = cgx::this_grid();
#ifdef FEMCY_SIMT_EMU
  unsigned char* ring_all = static_cast<unsigned char*>(simt::dyn_smem());
#else
  extern __shared__ __align__(128) unsigned char ring_all[];
#endif
  __shared__ unsigned long long full_bar[NW * NS];
  __shared__ double shw[2][NW];
  __shared__ double shf[4];
  __shared__ unsigned long long ph[S_PHASE_COUNT + 1];
  double* scal = a.scal;
  if (scal[S_DONE] != 0.0) return;                 // stable during this launch: set only by earlier launches
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nb = gridDim.x;
  // peer-memory path: warp 0 of block 0 is the COMMUNICATION warp -- it publishes the halo flag behind a system-scope fence
  // (~3 us) at the start of every SpMV phase and takes no slices, so the fence delays nobody (its share of the slices is
  // spread over the other nb*NW - 1 warps).  Slices are dealt round-robin to the remaining warps.
  const bool comm_warp = a.p2p && blockIdx.x == 0 && wib == 0;
  const int64_t nwarps = (int64_t)nb * NW - (a.p2p ? 1 : 0);
  const int64_t gw = comm_warp ? a.nslice : (int64_t)blockIdx.x * NW + wib - (a.p2p ? 1 : 0);
  const int64_t gs = (int64_t)nb * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double rmr = scal[S_RMR];
  const double eps = scal[S_EPS], r0 = scal[S_R0];
  const bool fixed = scal[S_FIXED] != 0.0;
  unsigned long long seq = (unsigned long long)scal[S_SEQ];
  double it_count = scal[S_ITER];
  int done = 0;
  double alpha = 0.0, beta = 0.0, dAd = 0.0, rmax_g = 0.0;
  const bool clk = (blockIdx.x == 0 && threadIdx.x == 0);
  if (clk) {
    for (int q = 0; q < S_PHASE_COUNT; ++q) ph[q] = 0ull;
    ph[S_PHASE_COUNT] = femcy_globaltimer();
  }
  auto stamp = [&](int slot) {
    if (clk) { unsigned long long t = femcy_globaltimer(); ph[slot] += t - ph[S_PHASE_COUNT]; ph[S_PHASE_COUNT] = t; }
  };

  // the matrix this kernel streams: all blocks, or the upper half (SYM)
  const int32_t* m_slice_ptr = SYM ? a.u_slice_ptr : a.slice_ptr;
  const int32_t* m_colidx = SYM ? a.u_colidx : a.colidx;
  const double* m_val = SYM ? a.u_val : a.val;

  unsigned char* ring = ring_all + (size_t)wib * NS * Cfg::STAGE_BYTES;
  unsigned long long* bars = full_bar + wib * NS;
  if (lane == 0) {
    for (int st = 0; st < NS; ++st) femcy_mbar_init(bars + st, 1);
  }
  __syncthreads();

  // ---- producer cursor (used by lane 0; kept uniform in the warp): next chunk of the cyclic sequence ----
  const bool has_work = gw < a.nslice;
  int64_t p_sidx = gw;
  int p_base = 0, p_w = 0, p_k = 0;
  auto p_open = [&]() {                              // read the producer's current slice
    int64_t s = a.p2p ? a.slice_order[p_sidx] : p_sidx;
    p_base = m_slice_ptr[s];
    p_w = (m_slice_ptr[s + 1] - p_base) >> 5;
    p_k = 0;
  };
  auto p_skip_empty = [&]() {                        // a slice without blocks has no chunk (the consumer writes zeros)
    int guard = 0;
    while (p_w == 0 && guard < 4) {
      p_sidx += nwarps;
      if (p_sidx >= a.nslice) { p_sidx = gw; ++guard; }
      p_open();
    }
  };
  auto p_issue = [&](int st) {                       // issue the producer's chunk into stage st, advance the cursor
    if (p_w == 0) return;                            // (all slices of this warp are empty)
    const int nk = (p_w - p_k) < KC ? (p_w - p_k) : KC;
    if (lane == 0) {
      unsigned char* dst = ring + (size_t)st * Cfg::STAGE_BYTES;
      const unsigned bytes_v = (unsigned)(nk * DM2 * 32 * 8), bytes_c = (unsigned)(nk * 32 * 4);
      femcy_mbar_arrive_expect_tx(bars + st, bytes_v + bytes_c);
      femcy_bulk_load_stream(dst, m_val + (((int64_t)(p_base >> 5) + p_k) * DM2 << 5), bytes_v, bars + st);
      femcy_bulk_load_stream(dst + Cfg::STAGE_VALS * 8, m_colidx + p_base + (p_k << 5), bytes_c, bars + st);
    }
    p_k += nk;
    if (p_k >= p_w) {
      p_sidx += nwarps;
      if (p_sidx >= a.nslice) p_sidx = gw;
      p_open();
      p_skip_empty();
    }
  };
  unsigned n_cons = 0, n_iss = 0;                    // chunks consumed / issued by this warp since kernel start
  if (has_work) {
    p_open();
    p_skip_empty();
    if (p_w > 0)
      for (int st = 0; st < NS; ++st) { p_issue(st); ++n_iss; }
  }

  // Stop decisions are taken by BLOCK 0 ALONE and travel to the other blocks with the next reduction (a word behind the block
  // partials, read by every block after the grid barrier that follows the next SpMV): all blocks of a rank leave the loop at
  // the same point even if a bounded peer wait ran out in some of them and their totals differ.  Price: one more d update +
  // SpMV after the deciding iteration (x and r are final by then).
  double* stop_word = a.part1 + 3 * (int64_t)nb;
  if (blockIdx.x == 0 && threadIdx.x == 0) *stop_word = 0.0;
  for (int it = 0; it < a.iters; ++it) {
    // ---- P1: Ad = A d, partial d.Ad ---------------------------------------------------------------
    double dot = 0.0;
    for (int64_t sidx = gw; sidx < a.nslice; sidx += nwarps) {
      const int64_t s = a.p2p ? a.slice_order[sidx] : sidx;
      if (a.p2p && a.slice_ghost[s]) {
        const unsigned long long* myflags = a.pv.win_of[a.pv.rank] + P2P_FLAG_D(0);
        if (lane < a.pv.nranks) {
          long long spins = 0;
          const long long limit = (*reinterpret_cast<const volatile double*>(scal + S_ERR) != 0.0) ? (1ll << 10) : (1ll << 24);
          while (ld_acquire_sys_u64(myflags + lane) < seq) {
            if (++spins > limit) { scal[S_ERR] = 3.0; break; }   // never changes control flow (grid.sync!)
            FEMCY_SPIN_PAUSE();
          }
        }
        __syncwarp();
      }
      const int base = m_slice_ptr[s];
      const int w = (m_slice_ptr[s + 1] - base) >> 5;
      int64_t i = s * 32 + lane;                      // position in the (sigma-sorted) row order -> row node
      bool row_ok = i < a.nrows;
      if (a.rowof) { i = row_ok ? a.rowof[i] : -1; row_ok = i >= 0; }
      double acc[DM], xi[DM];
#pragma unroll
      for (int r = 0; r < DM; ++r) { acc[r] = 0.0; xi[r] = (SYM && row_ok) ? a.d[i * DM + r] : 0.0; }
      for (int k0 = 0; k0 < w; k0 += KC) {
        const int st = (int)(n_cons % NS);
        femcy_mbar_wait(bars + st, (n_cons / NS) & 1u);
        const double* vs = reinterpret_cast<const double*>(ring + (size_t)st * Cfg::STAGE_BYTES);
        const int32_t* cs = reinterpret_cast<const int32_t*>(ring + (size_t)st * Cfg::STAGE_BYTES + Cfg::STAGE_VALS * 8);
        const int nk = (w - k0) < KC ? (w - k0) : KC;
        int c[KC];
        double xv[KC][DM];
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
          c[kk] = (kk < nk) ? cs[(kk << 5) + lane] : -1;
          if (c[kk] >= 0) {
            if (a.p2p && c[kk] >= (int)a.nrows) {
              // ghost column: written by another GPU during this kernel -> read at L2 (an L1 line brought in earlier by
              // a neighbouring owned column could be stale)
#pragma unroll
              for (int j = 0; j < DM; ++j) xv[kk][j] = __ldcg(a.d + (int64_t)c[kk] * DM + j);
            } else {
#pragma unroll
              for (int j = 0; j < DM; ++j) xv[kk][j] = a.d[(int64_t)c[kk] * DM + j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < DM; ++j) xv[kk][j] = 0.0;
          }
        }
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
          if (c[kk] >= 0) {
            double av[DM2];
#pragma unroll
            for (int q = 0; q < DM2; ++q) av[q] = vs[((kk * DM2 + q) << 5) + lane];
            if constexpr (!SYM) {
#pragma unroll
              for (int r = 0; r < DM; ++r)
#pragma unroll
                for (int j = 0; j < DM; ++j) acc[r] += av[r * DM + j] * xv[kk][j];
            } else {
              // upper half: y_i += K_ij x_j and, for an owned off-diagonal column, y_j += K_ij^T x_i (see bsell_row_sym)
              double xt = 0.0;
#pragma unroll
              for (int r = 0; r < DM; ++r) {
                double t = 0.0;
#pragma unroll
                for (int j = 0; j < DM; ++j) t += av[r * DM + j] * xv[kk][j];
                acc[r] += t;
                xt += xi[r] * t;
              }
              if (c[kk] != (int)i && c[kk] < (int)a.nrows) {
#pragma unroll
                for (int j = 0; j < DM; ++j) {
                  double u = 0.0;
#pragma unroll
                  for (int r = 0; r < DM; ++r) u += av[r * DM + j] * xi[r];
                  femcy_red_add_f64(a.Ad + (int64_t)c[kk] * DM + j, u);
                }
                dot += 2.0 * xt;
              } else {
                dot += xt;
              }
            }
          }
        }
        __syncwarp();                                 // every lane is done with stage st
        ++n_cons;
        p_issue(st);                                  // refill it with the next chunk of the cyclic sequence
        ++n_iss;
      }
      if (row_ok) {
        if constexpr (SYM) {
#pragma unroll
          for (int r = 0; r < DM; ++r) femcy_red_add_f64(a.Ad + i * DM + r, acc[r]);
        } else {
#pragma unroll
          for (int rr = 0; rr < DM; ++rr) {
            a.Ad[i * DM + rr] = acc[rr];
            dot += acc[rr] * a.d[i * DM + rr];
          }
        }
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) shw[0][wib] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double b = 0.0;
      for (int j = 0; j < NW; ++j) b += shw[0][j];
      a.part1[blockIdx.x] = b;
    }
    stamp(0);
    {
      double loc[1], tot[1];
      const bool im[1] = {false};
      grid.sync();
      const double stop_now = __ldcg(stop_word);        // block 0's decision of the previous iteration (uniform)
      fold_small<1>(a.part1, nb, loc, im, shf);
      stamp(1);
      if (stop_now != 0.0) { done = (int)stop_now; break; }
      if (a.p2p) { if (!p2p_exchange_all_blocks<1>(a.pv, 0, loc, seq + 1ull, tot, im, scal + S_ERR)) scal[S_ERR] = 3.0; }
      else tot[0] = loc[0];
      stamp(2);
      dAd = tot[0];
      alpha = rmr / dAd;
    }
    // d.Ad <= 0 (or NaN): K is not positive definite (a diverged Newton step) -- CG breaks down.  (Without this the recurrence
    // wanders until the iteration bound.)  Recorded here, acted upon through the stop word like every other decision.
    if (!(dAd > 0.0)) done = 2;
    // ---- P2: x += alpha d ; r -= alpha Ad ; partial r.M.r, max|r| ---------------------------------
    double prmr = 0.0, prmax = 0.0;
    {
      const int64_t n2 = a.n >> 1;
      double2* x2 = reinterpret_cast<double2*>(a.x);
      double2* r2 = reinterpret_cast<double2*>(a.r);
      const double2* d2 = reinterpret_cast<const double2*>(a.d);
      double2* A2 = reinterpret_cast<double2*>(a.Ad);
      const double2* M2 = reinterpret_cast<const double2*>(a.M);
      for (int64_t i = tid; i < n2; i += gs) {
        double2 xv = x2[i], dv = d2[i], rv = r2[i], av = A2[i], mv = M2[i];
        xv.x = xv.x + alpha * dv.x; xv.y = xv.y + alpha * dv.y;
        rv.x = rv.x - alpha * av.x; rv.y = rv.y - alpha * av.y;
        x2[i] = xv;
        r2[i] = rv;
        if constexpr (SYM) A2[i] = make_double2(0.0, 0.0);     // the next SpMV accumulates into Ad
        prmr += rv.x * mv.x * rv.x;
        prmr += rv.y * mv.y * rv.y;
        prmax = fmax(prmax, fmax(fabs(rv.x), fabs(rv.y)));
        if (rv.x != rv.x || rv.y != rv.y) prmax = 1.0 / 0.0;
      }
      if ((a.n & 1) && tid == 0) {
        int64_t i = a.n - 1;
        a.x[i] = a.x[i] + alpha * a.d[i];
        double rn = a.r[i] - alpha * a.Ad[i];
        a.r[i] = rn;
        if constexpr (SYM) a.Ad[i] = 0.0;
        prmr += rn * a.M[i] * rn;
        prmax = fmax(prmax, fabs(rn));
        if (rn != rn) prmax = 1.0 / 0.0;
      }
    }
    prmr = warp_sum(prmr);
    prmax = warp_max(prmax);
    if (lane == 0) { shw[0][wib] = prmr; shw[1][wib] = prmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double b0 = 0.0, b1 = 0.0;
      for (int j = 0; j < NW; ++j) { b0 += shw[0][j]; b1 = fmax(b1, shw[1][j]); }
      a.part2[blockIdx.x * 2] = b0;
      a.part2[blockIdx.x * 2 + 1] = b1;
    }
    stamp(3);
    {
      double loc[2], tot[2];
      const bool im[2] = {false, true};
      grid.sync();
      fold_small<2>(a.part2, nb, loc, im, shf);
      stamp(4);
      if (a.p2p) { if (!p2p_exchange_all_blocks<2>(a.pv, 1, loc, seq + 1ull, tot, im, scal + S_ERR)) scal[S_ERR] = 3.0; }
      else { tot[0] = loc[0]; tot[1] = loc[1]; }
      stamp(5);
      beta = tot[0] / rmr;
      rmr = tot[0];
      rmax_g = tot[1];
      it_count += 1.0;
      if (!fixed && rmax_g < eps * r0) done = done ? done : 1;                 // conjugateGradientSolver.py:124
      if (!(rmax_g < 1.0e300) || rmr != rmr) done = 2;
    }
    if (done && blockIdx.x == 0 && threadIdx.x == 0) *stop_word = (double)done;   // read by all blocks after the next SpMV
    // ---- P3: d = M r + beta d (boundary entries first, pushed to the neighbours' ghost slots) -----
    if (a.p2p) {
      // boundary entries: one 16-byte record per boundary node holds the node, its (first) destination rank and remote
      // index -- update, store locally and into the neighbour's ghost slot (NVLink peer store)
      bool pushed = false;
      for (int64_t t = tid; t < (int64_t)a.n_bnodes * DM; t += gs) {
        const int k = (int)(t / DM);
        const int c = (int)(t - (int64_t)k * DM);
        const int4 bp = a.bpush[k];
        const int64_t i = (int64_t)bp.x * DM + c;
        const double dn = a.M[i] * a.r[i] + beta * a.d[i];
        a.d[i] = dn;
        a.pv.d_of[bp.y][(int64_t)bp.z * DM + c] = dn;
        if (bp.w > 0) {
          for (int e = a.push_ptr[bp.x] + 1; e < a.push_ptr[bp.x + 1]; ++e)
            a.pv.d_of[a.push_peer[e]][(int64_t)a.push_ridx[e] * DM + c] = dn;
        }
        pushed = true;
      }
      // interior entries, two at a time (16-byte accesses); an entry of a boundary node was updated above
      {
        const int64_t n2 = a.n >> 1;
        double2* d2 = reinterpret_cast<double2*>(a.d);
        const double2* r2 = reinterpret_cast<const double2*>(a.r);
        const double2* M2 = reinterpret_cast<const double2*>(a.M);
        for (int64_t i = tid; i < n2; i += gs) {
          const int64_t e0 = 2 * i;
          // (all five loads are independent: the flags do not gate the data loads)
          double2 dv = d2[i], rv = r2[i], mv = M2[i];
          const bool f0 = a.bflag[e0 / DM] != 0, f1 = a.bflag[(e0 + 1) / DM] != 0;
          if (f0 && f1) continue;
          if (!f0) dv.x = mv.x * rv.x + beta * dv.x;
          if (!f1) dv.y = mv.y * rv.y + beta * dv.y;
          if (!f0 && !f1) d2[i] = dv;
          else if (!f0) a.d[e0] = dv.x;
          else a.d[e0 + 1] = dv.y;
        }
        if ((a.n & 1) && tid == 0 && !a.bflag[(a.n - 1) / DM]) a.d[a.n - 1] = a.M[a.n - 1] * a.r[a.n - 1] + beta * a.d[a.n - 1];
      }
      // the peer stores were issued before the interior entries and are acknowledged while those run; the system-scope
      // fence of every pushing thread precedes the grid barrier, the halo flag is published right after it: a neighbour
      // asks for the flag only when its SpMV reaches the slices with ghost columns (last in its slice order)
      if (pushed) __threadfence_system();
    } else {
      const int64_t n2 = a.n >> 1;
      double2* d2 = reinterpret_cast<double2*>(a.d);
      const double2* r2 = reinterpret_cast<const double2*>(a.r);
      const double2* M2 = reinterpret_cast<const double2*>(a.M);
      for (int64_t i = tid; i < n2; i += gs) {
        double2 dv = d2[i], rv = r2[i], mv = M2[i];
        dv.x = mv.x * rv.x + beta * dv.x;
        dv.y = mv.y * rv.y + beta * dv.y;
        d2[i] = dv;
      }
      if ((a.n & 1) && tid == 0) a.d[a.n - 1] = a.M[a.n - 1] * a.r[a.n - 1] + beta * a.d[a.n - 1];
    }
    grid.sync();
    if (a.p2p && blockIdx.x == 0 && threadIdx.x < a.pv.nranks) {
      // release pattern at system scope: this thread has observed (through the grid barrier) the fenced peer stores of all
      // pushing threads; its own fence makes them precede the flag for an observer on another GPU
      __threadfence_system();
      st_sys_u64(a.pv.win_of[threadIdx.x] + P2P_FLAG_D(a.pv.rank), seq + 1ull);
    }
    stamp(6);
    seq += 1ull;
  }
  // drain the ring: the copies issued ahead must land before the shared memory is released
  while (n_cons < n_iss) {
    if (p_w > 0) femcy_mbar_wait(bars + (n_cons % NS), (n_cons / NS) & 1u);
    ++n_cons;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int q = 0; q < S_PHASE_COUNT; ++q) scal[S_PHASE + q] += (double)ph[q];
    scal[S_RMR] = rmr; scal[S_ALPHA] = alpha; scal[S_BETA] = beta; scal[S_DAD] = dAd; scal[S_RMAX] = rmax_g;
    scal[S_ITER] = it_count; scal[S_SEQ] = (double)seq;
    if (done) scal[S_DONE] = (double)done;
    if (scal[S_ERR] != 0.0) scal[S_DONE] = 3.0;
  }
}

// M = 1/diag(A) (M_init :48-51) ; r = b ; d = M r (r_d_init :60-65) ; x = 0 ; partials: rMr, max|r|
template <int DM>
__global__ void __launch_bounds__(256)
k_cg_init(const int32_t* __restrict__ diag_slot, const double* __restrict__ val, const double* __restrict__ b,
          double* __restrict__ x, double* __restrict__ r, double* __restrict__ d, double* __restrict__ M,
          double* __restrict__ Ad, int64_t nrows, double* partials, unsigned int* ticket, double* scal, int multi) {
  constexpr int DM2 = DM * DM;
  double rmr = 0.0, rmax = 0.0;
  int64_t n = nrows * DM;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / DM;
    int c = (int)(t - i * DM);
    int slot = diag_slot[i];
    // A_get returns A[i][0] when the diagonal is absent (:40-46); a row without a diagonal block
    // cannot come out of an FE assembly, treat it as 1/0 like the reference would in effect.
    double diag = (slot >= 0) ? val[(((int64_t)(slot >> 5) * DM2 + (c * DM + c)) << 5) + (slot & 31)] : 0.0;
    double m = 1.0 / diag;
    double bi = b[t];
    M[t] = m;
    r[t] = bi;
    d[t] = m * bi;
    x[t] = 0.0;
    Ad[t] = 0.0;
    rmr += bi * m * bi;
    rmax = fmax(rmax, fabs(bi));
  }
  double mine[2] = {rmr, rmax}, tot[2];
  const bool is_max[2] = {false, true};
  if (grid_reduce<2>(mine, partials, ticket, tot, is_max)) {
    if (multi) { scal[S_SEND] = tot[0]; scal[S_SEND + 1] = tot[1]; }
    else {
      scal[S_RMR] = tot[0]; scal[S_R0] = tot[1]; scal[S_RMAX] = tot[1];
      if (tot[1] == 0.0) scal[S_DONE] = 1.0;      // b = 0: x = 0 is the solution; without this alpha = 0/0 poisons x
    }
  }
}

__global__ void k_finish_init(double* scal, int nranks) {
  double t = 0.0, m = 0.0;
  for (int r = 0; r < nranks; ++r) { t += scal[S_GATHER + 2 * r]; m = fmax(m, scal[S_GATHER + 2 * r + 1]); }
  scal[S_RMR] = t; scal[S_R0] = m; scal[S_RMAX] = m;
  if (m == 0.0) scal[S_DONE] = 1.0;               // b = 0 on every rank: x = 0 is the solution
}

__global__ void k_set_scalars(double* scal, double eps, double fixed) {
  scal[S_EPS] = eps; scal[S_DONE] = 0.0; scal[S_ITER] = 0.0; scal[S_FIXED] = fixed;
  scal[S_ALPHA] = 0.0; scal[S_BETA] = 0.0; scal[S_DAD] = 0.0; scal[S_ERR] = 0.0;
  for (int q = 0; q < S_PHASE_COUNT; ++q) scal[S_PHASE + q] = 0.0;
}
