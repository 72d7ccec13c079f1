// Element-local fp64 math shared by the assembly / post-processing kernels.
//
// Conventions follow the reference exactly:
//   J      = x^T . dN/dxi          (stiffnessMtrx.py:148  `localNodes.transpose() @ dsdn`)
//   grad N = dN/dxi . J^-1         (stiffnessMtrx.py:149)
//   vol    = det(J) * w_gp         (stiffnessMtrx.py:150)
//   B      = strainMtrx(grad N): Voigt rows 2-D [xx,yy,xy], 3-D [xx,yy,zz,xy,zx,yz],
//            columns node-major / component-minor (element_linear_tetrahedral.py:137-177 etc.)
#pragma once
#include "device_compat.cuh"
#include "kernel_types.cuh"

template <int DM>
struct Voigt { static constexpr int NV = (DM == 2) ? 3 : 6; };

// 2x2 / 3x3 inverse by adjugate; returns det.
__device__ __forceinline__ double inv2(const double (&J)[2][2], double (&Ji)[2][2]) {
  double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  double id = 1.0 / det;
  Ji[0][0] = J[1][1] * id;  Ji[0][1] = -J[0][1] * id;
  Ji[1][0] = -J[1][0] * id; Ji[1][1] = J[0][0] * id;
  return det;
}
__device__ __forceinline__ double inv3(const double (&J)[3][3], double (&Ji)[3][3]) {
  double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  double id = 1.0 / det;
  Ji[0][0] = c00 * id;
  Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
  Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
  Ji[1][0] = c01 * id;
  Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
  Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
  Ji[2][0] = c02 * id;
  Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
  Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
  return det;
}
template <int DM>
__device__ __forceinline__ double inv_dm(const double (&J)[DM][DM], double (&Ji)[DM][DM]) {
  if constexpr (DM == 2) return inv2(J, Ji);
  else return inv3(J, Ji);
}
template <int DM>
__device__ __forceinline__ double det_dm(const double (&J)[DM][DM]) {
  if constexpr (DM == 2) return J[0][0] * J[1][1] - J[0][1] * J[1][0];
  else
    return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) +
           J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

// gradients of the NEN shape functions at one Gauss point; returns det(J).
// x[a][i]: nodal coordinates of the configuration in which the gradient is wanted.
// dN: [NEN][DM] natural derivatives at this Gauss point.
template <int DM, int NEN>
__device__ __forceinline__ double shape_gradients(const double (&x)[NEN][DM], const double* __restrict__ dN,
                                                  double (&g)[NEN][DM]) {
  double J[DM][DM], Ji[DM][DM];
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int k = 0; k < DM; ++k) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < NEN; ++a) s += x[a][i] * dN[a * DM + k];
      J[i][k] = s;
    }
  double det = inv_dm<DM>(J, Ji);
#pragma unroll
  for (int a = 0; a < NEN; ++a)
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < DM; ++k) s += dN[a * DM + k] * Ji[k][j];
      g[a][j] = s;
    }
  return det;
}

// T = C . B_b  (NV x DM) for one node's strain columns, using the sparsity of B.
template <int DM>
__device__ __forceinline__ void C_times_B(const double* __restrict__ C, const double (&gb)[DM],
                                          double (&T)[Voigt<DM>::NV][DM]) {
  constexpr int NV = Voigt<DM>::NV;
  if constexpr (DM == 2) {
    // B_b = [[gx,0],[0,gy],[gy,gx]]
#pragma unroll
    for (int p = 0; p < NV; ++p) {
      T[p][0] = C[p * NV + 0] * gb[0] + C[p * NV + 2] * gb[1];
      T[p][1] = C[p * NV + 1] * gb[1] + C[p * NV + 2] * gb[0];
    }
  } else {
    // B_b = [[gx,0,0],[0,gy,0],[0,0,gz],[gy,gx,0],[gz,0,gx],[0,gz,gy]]
#pragma unroll
    for (int p = 0; p < NV; ++p) {
      T[p][0] = C[p * NV + 0] * gb[0] + C[p * NV + 3] * gb[1] + C[p * NV + 4] * gb[2];
      T[p][1] = C[p * NV + 1] * gb[1] + C[p * NV + 3] * gb[0] + C[p * NV + 5] * gb[2];
      T[p][2] = C[p * NV + 2] * gb[2] + C[p * NV + 4] * gb[0] + C[p * NV + 5] * gb[1];
    }
  }
}

// acc[i][j] += s * (B_a^T . T)[i][j]
template <int DM>
__device__ __forceinline__ void Bt_times_T_acc(const double (&ga)[DM], const double (&T)[Voigt<DM>::NV][DM],
                                               double s, double (&acc)[DM][DM]) {
  if constexpr (DM == 2) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      acc[0][j] += s * (ga[0] * T[0][j] + ga[1] * T[2][j]);
      acc[1][j] += s * (ga[1] * T[1][j] + ga[0] * T[2][j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      acc[0][j] += s * (ga[0] * T[0][j] + ga[1] * T[3][j] + ga[2] * T[4][j]);
      acc[1][j] += s * (ga[1] * T[1][j] + ga[0] * T[3][j] + ga[2] * T[5][j]);
      acc[2][j] += s * (ga[2] * T[2][j] + ga[0] * T[4][j] + ga[1] * T[5][j]);
    }
  }
}

// acc += s * (B_a^T C B_b) for a tangent of cubic form -- C_ii = p, C_ij = q (i != j, normal block), shear diagonal r,
// everything else 0 -- which covers every tangent the reference ships (isotropic 3-D, plane strain, plane stress,
// and the fixed neo-Hookean `C` of neo_hookean.py:22-42).  Worked out from sigma_ii = p e_ii + q sum_{k != i} e_kk,
// sigma_ij = r gamma_ij:   K_ij = q ga_i gb_j + r ga_j gb_i (i != j),   K_ii = p ga_i gb_i + r (ga.gb - ga_i gb_i).
// ~27 FP64 instructions instead of the 99 of C_times_B + Bt_times_T_acc; rounding differs from the general path in
// the last bits only.
template <int DM>
__device__ __forceinline__ void block_cubic_acc(double p, double q, double r, const double (&ga)[DM],
                                                const double (&gb)[DM], double s, double (&acc)[DM][DM]) {
  double sa[DM];
  double dot = 0.0;
#pragma unroll
  for (int i = 0; i < DM; ++i) { sa[i] = s * ga[i]; dot += sa[i] * gb[i]; }
  const double rd = r * dot, pr = p - r;
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      if (i == j) acc[i][i] += pr * (sa[i] * gb[i]) + rd;
      else acc[i][j] += q * (sa[i] * gb[j]) + r * (sa[j] * gb[i]);
    }
}

// host + device: does the row-major [NV][NV] tangent have the cubic form above?
static inline bool tangent_is_cubic(const double* C, int dm) {
  const int nv = (dm == 2) ? 3 : 6;
  const double p = C[0], q = C[1], r = C[nv * nv - 1];
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) {
      double want = 0.0;
      if (i < dm && j < dm) want = (i == j) ? p : q;
      else if (i == j) want = r;
      if (C[i * nv + j] != want) return false;
    }
  return true;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Deterministic grid reduction: every block writes its partial(s), the last block to finish (ticket)
// folds them with ALL its threads -- thread t sums partials t, t+B, t+2B, ... in that order, then a
// fixed-shape shared-memory tree combines the B per-thread sums.  The fold order depends only on
// (gridDim, blockDim), never on scheduling => bit-reproducible.  Returns true on thread 0 of the
// last block with the totals in tot[].  NV_ values per block; is_max[i] selects max instead of sum.
// (A single-thread fold of ~7000 partials costs ~250 us of serial L2 latency per kernel: measured,
//  profiles/r1_notes.md.)
template <int NV_>
__device__ __forceinline__ bool grid_reduce(const double (&mine)[NV_], double* partials, unsigned int* ticket,
                                            double (&tot)[NV_], const bool (&is_max)[NV_]) {
  __shared__ double sh[NV_][256];
  __shared__ bool last;
  int t = threadIdx.x, B = blockDim.x;   // B <= 256, multiple of 32
  int w = t >> 5, l = t & 31, nw = B >> 5;
#pragma unroll
  for (int i = 0; i < NV_; ++i) {
    double v = is_max[i] ? warp_max(mine[i]) : warp_sum(mine[i]);
    if (l == 0) sh[i][w] = v;
  }
  __syncthreads();
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < NV_; ++i) {
      double b = 0.0;
      for (int j = 0; j < nw; ++j) b = is_max[i] ? fmax(b, sh[i][j]) : b + sh[i][j];
      partials[(int64_t)blockIdx.x * NV_ + i] = b;
    }
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  // L2-coherent (ld.global.cg) loads, 4 in flight per thread: a volatile loop serialises one L2
  // round trip per partial (~27 per thread for the SpMV grid => ~8 us of tail per kernel).
#pragma unroll
  for (int i = 0; i < NV_; ++i) {
    double acc = 0.0;
    unsigned int nb = gridDim.x;
    unsigned int b = t;
    for (; b + 3u * B < nb; b += 4u * B) {
      double p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) p[u] = __ldcg(partials + (int64_t)(b + u * B) * NV_ + i);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = is_max[i] ? fmax(acc, p[u]) : acc + p[u];
    }
    for (; b < nb; b += B) {
      double p = __ldcg(partials + (int64_t)b * NV_ + i);
      acc = is_max[i] ? fmax(acc, p) : acc + p;
    }
    sh[i][t] = acc;
  }
  __syncthreads();
  for (int s = B >> 1; s > 0; s >>= 1) {
    if (t < s) {
#pragma unroll
      for (int i = 0; i < NV_; ++i) sh[i][t] = is_max[i] ? fmax(sh[i][t], sh[i][t + s]) : sh[i][t] + sh[i][t + s];
    }
    __syncthreads();
  }
  if (t != 0) return false;
#pragma unroll
  for (int i = 0; i < NV_; ++i) tot[i] = sh[i][0];
  *ticket = 0;
  return true;
}
