// Device code of the Dirichlet path (row a4); header for the same reason as assembly_kernels.cuh / cg_kernels.cuh
// (host launch code in bc.cu, CPU SIMT emulation in tests/simt).
//   dirichletBC_linearEquations          /root/reference/stiffnessMtrx.py:279-307
//   dirichletBC_forNewtonMethod_kernel   /root/reference/stiffnessMtrx.py:317-341
//   dirichletBC_val                      /root/reference/stiffnessMtrx.py:357-366
#pragma once
#include "device_compat.cuh"
#include "kernel_types.cuh"

__global__ void k_bc_mark(const int32_t* __restrict__ nodes, const int32_t* __restrict__ comps,
                          const double* __restrict__ vals, int64_t n, int dm, unsigned char* __restrict__ flag,
                          double* __restrict__ valfull, unsigned char f) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t idx = (int64_t)nodes[t] * dm + comps[t];
    flag[idx] = f;
    if (vals) valfull[idx] = vals[t];
  }
}

__global__ void k_bc_val(const int32_t* __restrict__ nodes, const int32_t* __restrict__ comps,
                         const double* __restrict__ vals, int64_t n, int dm, double* __restrict__ dof) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    dof[(int64_t)nodes[t] * dm + comps[t]] = vals[t];
}

// mode 0: linear equations (rhs corrected, rhs[i]=val) ; mode 1: Newton (target[i] = 0)
template <int DM>
__global__ void __launch_bounds__(256)
k_bc_apply(const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ colidx, double* __restrict__ val,
           int64_t nrows, int64_t nslice, const unsigned char* __restrict__ flag, const double* __restrict__ valfull,
           double* __restrict__ target, int mode, const int32_t* __restrict__ rowof) {
  constexpr int DM2 = DM * DM;
  int lane = threadIdx.x & 31;
  int64_t s = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= nslice) return;
  int64_t i = s * 32 + lane;         // position in the (sigma-sorted) row order
  if (i >= nrows) return;
  if (rowof) i = rowof[i];           // -> row node
  int base = slice_ptr[s];
  int w = (slice_ptr[s + 1] - base) >> 5;
  bool fi[DM];
  bool any_row = false;
#pragma unroll
  for (int r = 0; r < DM; ++r) { fi[r] = flag[i * DM + r] != 0; any_row |= fi[r]; }
  double corr[DM];
#pragma unroll
  for (int r = 0; r < DM; ++r) corr[r] = 0.0;
  for (int k = 0; k < w; ++k) {
    int c = colidx[base + (k << 5) + lane];
    if (c < 0) continue;
    bool fc[DM];
    bool any_col = false;
#pragma unroll
    for (int j = 0; j < DM; ++j) { fc[j] = flag[(int64_t)c * DM + j] != 0; any_col |= fc[j]; }
    if (!any_row && !any_col) continue;
    double* v = val + (((int64_t)((base >> 5) + k) * DM2) << 5) + lane;
#pragma unroll
    for (int r = 0; r < DM; ++r)
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        if (fi[r] || fc[j]) {
          double* p = v + ((r * DM + j) << 5);
          if (mode == 0 && fc[j] && !fi[r]) corr[r] += valfull[(int64_t)c * DM + j] * (*p);
          *p = (c == i && r == j && fi[r]) ? 1.0 : 0.0;
        }
      }
  }
#pragma unroll
  for (int r = 0; r < DM; ++r) {
    if (fi[r]) target[i * DM + r] = (mode == 0) ? valfull[i * DM + r] : 0.0;
    else if (mode == 0 && corr[r] != 0.0) target[i * DM + r] -= corr[r];
  }
}

