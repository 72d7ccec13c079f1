// Device code of the assembly path (rows a2 + a3): shape-function gradients / volumes and the stiffness
// assembly kernels.  Kept in a header so that the host launch code (assembly.cu) and the CPU SIMT emulation
// used by the not-gpu kernel-logic tests (tests/simt, test infrastructure only) compile the same source.
//
// Reference kernels replaced:
//   System_of_equations.get_dsdx_and_vol        /root/reference/stiffnessMtrx.py:132-150
//   System_of_equations.assemble_stiffnessMtrx  /root/reference/stiffnessMtrx.py:161-186
//   System_of_equations.sparseMatrix_get_j      /root/reference/stiffnessMtrx.py:414-420  (row scan -> precomputed slot)
#pragma once
#include "device_compat.cuh"
#include "kernel_types.cuh"
#include "elem_math.cuh"
#include "constitutive.cuh"     // sigma_of_F: the constitutive laws (the consistent tangent differentiates them)

// ---------------------------------------------------------------------------------------------
// geometry at all Gauss points of one element on the configuration X + u
template <int DM, int NEN>
__device__ __forceinline__ void load_current_coords(const double* __restrict__ nodes, const double* __restrict__ dof,
                                                    const int32_t* __restrict__ conn, double (&x)[NEN][DM]) {
#pragma unroll
  for (int a = 0; a < NEN; ++a) {
    int64_t n = conn[a];
#pragma unroll
    for (int i = 0; i < DM; ++i) x[a][i] = nodes[n * DM + i] + dof[n * DM + i];
  }
}

template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(128)
k_dsdx_vol(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes, const double* __restrict__ dof,
           const int32_t* __restrict__ elems, int64_t ne, double* __restrict__ dsdx, double* __restrict__ vol) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int32_t conn[NEN];
#pragma unroll
  for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
  double x[NEN][DM];
  load_current_coords<DM, NEN>(nodes, dof, conn, x);
#pragma unroll 1
  for (int gp = 0; gp < NGP; ++gp) {
    double g[NEN][DM];
    double det = shape_gradients<DM, NEN>(x, &tab.dN[gp * NEN * DM], g);
    vol[e * NGP + gp] = det * tab.w[gp];
    if (dsdx) {
      double* o = dsdx + (e * NGP + gp) * (NEN * DM);
#pragma unroll
      for (int a = 0; a < NEN; ++a)
#pragma unroll
        for (int j = 0; j < DM; ++j) o[a * DM + j] = g[a][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// scatter assembly: thread per element
template <int DM, int NEN, int NGP, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_assemble_scatter(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                   const double* __restrict__ dof, const int32_t* __restrict__ elems,
                   const int32_t* __restrict__ elem_slot, int64_t ne, double* __restrict__ val) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int32_t conn[NEN];
#pragma unroll
  for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
  double x[NEN][DM];
  load_current_coords<DM, NEN>(nodes, dof, conn, x);

  double g[NGP][NEN][DM];
  double vol[NGP];
#pragma unroll
  for (int gp = 0; gp < NGP; ++gp) vol[gp] = shape_gradients<DM, NEN>(x, &tab.dN[gp * NEN * DM], g[gp]) * tab.w[gp];

  const int32_t* slots = elem_slot + e * (NEN * NEN);
  // big elements keep the pair loops rolled (g is then indexed dynamically -> local memory)
  constexpr bool ROLL = (NEN * NGP > 16);
#pragma unroll(ROLL ? 1 : NEN)
  for (int b = 0; b < NEN; ++b) {
    double T[NGP][NV][DM];
#pragma unroll
    for (int gp = 0; gp < NGP; ++gp) C_times_B<DM>(tab.C, g[gp][b], T[gp]);
#pragma unroll(ROLL ? 1 : NEN)
    for (int a = 0; a < NEN; ++a) {
      int32_t slot = slots[a * NEN + b];
      if (slot < 0) continue;  // row node owned by another rank
      double acc[DM][DM];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
#pragma unroll
      for (int gp = 0; gp < NGP; ++gp) Bt_times_T_acc<DM>(g[gp][a], T[gp], vol[gp], acc);
      double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + (slot & 31);
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) atomicAdd(dst + ((i * DM + j) << 5), acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row f2 (opt-in, option consistent_tangent): the CONSISTENT tangent of the internal force
//     f_a,i = sum_gp  sigma_im(F) grad_x N_a,m  vol            (assemble_nodal_force_GN, stiffnessMtrx.py:609-644)
// instead of the reference's constant-C stiffness (ddsdde is never updated: neo_hookean.py:62-64 is commented out, so the
// reference's Newton loop is a modified Newton iteration).  Exact linearisation (material + geometric part in one tensor):
//     K_ab,ij = sum_gp  grad N_a,m  A_imjn  grad N_b,n  vol ,     A_imjn = (1/J) d tau_im / dh [(I + h e_j e_n^T) F]  -  sigma_in delta_mj
// with tau = det(F) sigma(F).  The derivative is a central difference of the constitutive law ITSELF (h = 1e-6: truncation and
// round-off ~1e-10), so every material of material_zoo is covered by construction and the tangent is consistent with exactly
// the stress the residual uses; A is symmetrised over (im) <-> (jn), which is exact for a hyperelastic law.
template <int DM>
__device__ __forceinline__ void kirchhoff_of_F(const ElemTables& tab, int kind, const double (&F)[DM][DM], double (&T)[DM][DM]) {
  sigma_of_F<DM>(tab, kind, 1, F, T);
  const double J = det_dm<DM>(F);
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int m = 0; m < DM; ++m) T[i][m] *= J;
}

template <int DM>
__device__ void spatial_tangent(const ElemTables& tab, int kind, const double (&F)[DM][DM], double (&A)[DM][DM][DM][DM]) {
  const double h = 1.0e-6;
  const double Jinv = 1.0 / det_dm<DM>(F);
  double S[DM][DM];
  sigma_of_F<DM>(tab, kind, 1, F, S);
  for (int j = 0; j < DM; ++j)
    for (int n = 0; n < DM; ++n) {
      // (I +- h e_j e_n^T) F : row j of F gets +- h times row n
      double Fp[DM][DM], Fm[DM][DM], Tp[DM][DM], Tm[DM][DM];
#pragma unroll
      for (int r = 0; r < DM; ++r)
#pragma unroll
        for (int c = 0; c < DM; ++c) {
          const double d = (r == j) ? h * F[n][c] : 0.0;
          Fp[r][c] = F[r][c] + d;
          Fm[r][c] = F[r][c] - d;
        }
      kirchhoff_of_F<DM>(tab, kind, Fp, Tp);
      kirchhoff_of_F<DM>(tab, kind, Fm, Tm);
      for (int i = 0; i < DM; ++i)
        for (int m = 0; m < DM; ++m)
          A[i][m][j][n] = (Tp[i][m] - Tm[i][m]) * (0.5 / h) * Jinv - ((m == j) ? S[i][n] : 0.0);
    }
  for (int i = 0; i < DM; ++i)
    for (int m = 0; m < DM; ++m)
      for (int j = 0; j < DM; ++j)
        for (int n = 0; n < DM; ++n)
          if (i * DM + m < j * DM + n) {
            const double v = 0.5 * (A[i][m][j][n] + A[j][n][i][m]);
            A[i][m][j][n] = v;
            A[j][n][i][m] = v;
          }
}

// thread per element, all element kinds; per Gauss point: F, the current-configuration gradients, A, then dm*dm atomic adds per
// node pair (K is zero-filled by the caller).  Not a benchmark path: it trades the constant-C assembly's speed for Newton steps.
template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(128)
k_assemble_scatter_ct(const __grid_constant__ ElemTables tab, int kind, const double* __restrict__ nodes,
                      const double* __restrict__ dof, const int32_t* __restrict__ elems,
                      const int32_t* __restrict__ elem_slot, int64_t ne, double* __restrict__ val) {
  constexpr int DM2 = DM * DM;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  double X[NEN][DM], x[NEN][DM];
#pragma unroll
  for (int a = 0; a < NEN; ++a) {
    const int64_t nd = elems[e * NEN + a];
#pragma unroll
    for (int i = 0; i < DM; ++i) { X[a][i] = nodes[nd * DM + i]; x[a][i] = X[a][i] + dof[nd * DM + i]; }
  }
  const int32_t* slots = elem_slot + e * (NEN * NEN);
#pragma unroll 1
  for (int gp = 0; gp < NGP; ++gp) {
    const double* dN = &tab.dN[gp * NEN * DM];
    double JX[DM][DM], Jx[DM][DM], JXi[DM][DM], Jxi[DM][DM];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int k = 0; k < DM; ++k) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int a = 0; a < NEN; ++a) { s0 += X[a][i] * dN[a * DM + k]; s1 += x[a][i] * dN[a * DM + k]; }
        JX[i][k] = s0;
        Jx[i][k] = s1;
      }
    inv_dm<DM>(JX, JXi);
    const double vol = inv_dm<DM>(Jx, Jxi) * tab.w[gp];
    double F[DM][DM];                                   // dx/dX = (dx/dxi)(dX/dxi)^-1
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int M = 0; M < DM; ++M) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DM; ++k) s += Jx[i][k] * JXi[k][M];
        F[i][M] = s;
      }
    double g[NEN][DM];                                  // grad_x N
#pragma unroll
    for (int a = 0; a < NEN; ++a)
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DM; ++k) s += dN[a * DM + k] * Jxi[k][j];
        g[a][j] = s;
      }
    double A[DM][DM][DM][DM];
    spatial_tangent<DM>(tab, kind, F, A);
#pragma unroll 1
    for (int b = 0; b < NEN; ++b) {
      double T[DM][DM][DM];                             // T[i][m][j] = sum_n A_imjn grad N_b,n
      for (int i = 0; i < DM; ++i)
        for (int m = 0; m < DM; ++m)
          for (int j = 0; j < DM; ++j) {
            double s = 0.0;
#pragma unroll
            for (int n = 0; n < DM; ++n) s += A[i][m][j][n] * g[b][n];
            T[i][m][j] = s;
          }
#pragma unroll 1
      for (int a = 0; a < NEN; ++a) {
        const int32_t slot = slots[a * NEN + b];
        if (slot < 0) continue;                         // row node owned by another rank
        double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + (slot & 31);
        for (int i = 0; i < DM; ++i)
          for (int j = 0; j < DM; ++j) {
            double s = 0.0;
#pragma unroll
            for (int m = 0; m < DM; ++m) s += g[a][m] * T[i][m][j];
            atomicAdd(dst + ((i * DM + j) << 5), s * vol);
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// scatter assembly for big elements (C3D10, CPS8/CPE8): one WARP per element.
//   stage 1  lanes < NEN load the element's nodes (X + u) into shared memory
//   stage 2  lanes < NGP invert the Jacobian of their Gauss point; then the NGP*NEN (gp, node) pairs
//            are spread over the lanes to form grad N -> shared memory
//   stage 3  register tiling of K_e: lane = (column node b, row group a0); per Gauss point the lane
//            forms T = C.B_b once in registers and applies it to its <= APL row nodes a = a0, a0+G, ...
//            (3 + 3*APL shared loads per Gauss point instead of 21 per node pair), then DM*DM atomics per
//            pair into the precomputed slot.
// (The thread-per-element kernel needs NGP*NEN*DM gradients live per thread: 120 doubles for C3D10.)
template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(128, 4)
k_assemble_scatter_warp(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                        const double* __restrict__ dof, const int32_t* __restrict__ elems,
                        const int32_t* __restrict__ elem_slot, int64_t ne, double* __restrict__ val) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int WPB = 4;                   // warps per block
  constexpr int G = 32 / NEN;              // row groups per warp (3 for NEN=10, 4 for NEN=8)
  constexpr int APL = (NEN + G - 1) / G;   // row nodes per lane
  __shared__ double xs[WPB][NEN][DM];
  __shared__ double Ji_s[WPB][NGP][DM][DM];
  __shared__ double vol_s[WPB][NGP];
  __shared__ double g_s[WPB][NGP][NEN][DM];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t nwarps = (int64_t)gridDim.x * WPB;
  const int b = lane % NEN, a0 = lane / NEN;
  const bool active = lane < G * NEN;
  // grid-stride: a contiguous element range per warp was measured 2.4x slower (21.6 vs 8.9 ms on cfg 5, r1z)
  for (int64_t e = blockIdx.x * (int64_t)WPB + w; e < ne; e += nwarps) {
    if (lane < NEN) {
      int64_t n = elems[e * NEN + lane];
#pragma unroll
      for (int i = 0; i < DM; ++i) xs[w][lane][i] = nodes[n * DM + i] + dof[n * DM + i];
    }
    __syncwarp();
    if (lane < NGP) {
      const double* dN = &tab.dN[lane * NEN * DM];
      double J[DM][DM], Ji[DM][DM];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int k = 0; k < DM; ++k) {
          double sacc = 0.0;
#pragma unroll
          for (int a = 0; a < NEN; ++a) sacc += xs[w][a][i] * dN[a * DM + k];
          J[i][k] = sacc;
        }
      double det = inv_dm<DM>(J, Ji);
      vol_s[w][lane] = det * tab.w[lane];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int k = 0; k < DM; ++k) Ji_s[w][lane][i][k] = Ji[i][k];
    }
    __syncwarp();
    for (int p = lane; p < NGP * NEN; p += 32) {
      int gp = p / NEN, a = p - gp * NEN;
      const double* dN = &tab.dN[(gp * NEN + a) * DM];
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < DM; ++k) sacc += dN[k] * Ji_s[w][gp][k][j];
        g_s[w][gp][a][j] = sacc;
      }
    }
    __syncwarp();
    if (active) {
      double acc[APL][DM][DM];
#pragma unroll
      for (int m = 0; m < APL; ++m)
#pragma unroll
        for (int i = 0; i < DM; ++i)
#pragma unroll
          for (int j = 0; j < DM; ++j) acc[m][i][j] = 0.0;
#pragma unroll 1
      for (int gp = 0; gp < NGP; ++gp) {
        double gb[DM], T[NV][DM];
#pragma unroll
        for (int j = 0; j < DM; ++j) gb[j] = g_s[w][gp][b][j];
        C_times_B<DM>(tab.C, gb, T);
        double v = vol_s[w][gp];
#pragma unroll
        for (int m = 0; m < APL; ++m) {
          int a = a0 + m * G;
          if (a < NEN) {
            double ga[DM];
#pragma unroll
            for (int j = 0; j < DM; ++j) ga[j] = g_s[w][gp][a][j];
            Bt_times_T_acc<DM>(ga, T, v, acc[m]);
          }
        }
      }
      const int32_t* slots = elem_slot + e * (NEN * NEN);
#pragma unroll
      for (int m = 0; m < APL; ++m) {
        int a = a0 + m * G;
        if (a < NEN) {
          int32_t slot = slots[a * NEN + b];
          if (slot >= 0) {
            double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + (slot & 31);
#pragma unroll
            for (int i = 0; i < DM; ++i)
#pragma unroll
              for (int j = 0; j < DM; ++j) atomicAdd(dst + ((i * DM + j) << 5), acc[m][i][j]);
          }
        }
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Gather assembly (default): atomic-free, no zero-fill, bit-reproducible.
//   pass 1  one 32-byte sector per (element, node, Gauss point):  rec[e][a][gp] = (dN_a/dx, dN_a/dy, dN_a/dz | 0, vol_gp),
//           staged per block in shared memory and written as whole records (k_elem_geometry4s) or, for C3D4 whose record
//           is exactly one 128-byte row, handed to the TMA unit as ONE tensor store per block (k_elem_geometry4t);
//   pass 2  k_assemble_gather_p: one thread per stored block sums its (element, a, b) list with one 256-bit load per
//           record (LDG.E.256) and accumulates the gradient products first -- the tangent enters once per block.
// Measured on B200 (profiles/r2a, r2b, r2i): 2.48 ms for 10.1 M C3D4 (scatter 3.27), 4.16 ms for 1.0 M C3D10 (scatter
// 8.9); pass 2 is bound by the L1 data pipe (one wavefront per scattered 32-byte sector: 95 % busy, ncu r2i).
template <int NEN, int NGP>
struct Geo4Cfg {
  static constexpr int CH = NEN * NGP * 2;                 // 16-byte chunks per element record
  static constexpr int TPB = (CH <= 8) ? 128 : 32;         // elements (= threads) per block
};

template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(Geo4Cfg<NEN, NGP>::TPB)
k_elem_geometry4s(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                  const double* __restrict__ dof, const int32_t* __restrict__ elems, int64_t ne,
                  double* __restrict__ rec, double* __restrict__ vol_out) {
  constexpr int CH = Geo4Cfg<NEN, NGP>::CH, TPB = Geo4Cfg<NEN, NGP>::TPB;
  __shared__ double2 tile[TPB * (CH + 1)];
  const int t = threadIdx.x;
  const int64_t e0 = blockIdx.x * (int64_t)TPB;
  const int64_t e = e0 + t;
  if (e < ne) {
    int32_t conn[NEN];
#pragma unroll
    for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
    double x[NEN][DM];
    load_current_coords<DM, NEN>(nodes, dof, conn, x);
    double2* o = tile + t * (CH + 1);
#pragma unroll 1
    for (int gp = 0; gp < NGP; ++gp) {
      double g[NEN][DM];
      double v = shape_gradients<DM, NEN>(x, &tab.dN[gp * NEN * DM], g) * tab.w[gp];
#pragma unroll
      for (int a = 0; a < NEN; ++a) {
        double2 lo, hi;
        lo.x = g[a][0]; lo.y = g[a][1];
        hi.x = (DM == 3) ? g[a][DM - 1] : 0.0; hi.y = v;
        o[(a * NGP + gp) * 2] = lo;
        o[(a * NGP + gp) * 2 + 1] = hi;
      }
      vol_out[e * NGP + gp] = v;
    }
  }
  __syncthreads();
  const int64_t rem = ne - e0;
  const int nel = rem < TPB ? (int)rem : TPB;
  double2* out = reinterpret_cast<double2*>(rec + e0 * (NEN * NGP * 4));
  for (int gi = t; gi < nel * CH; gi += TPB) {
    int el = gi / CH, c = gi - el * CH;
    out[gi] = tile[el * (CH + 1) + c];
  }
}

// pass 1 with a TMA tensor store (C3D4: the record is exactly one 128-byte row).  The block's 128 records
// are written into a 16 KB shared-memory tile in the 128-byte-swizzled layout (16-byte chunk c of row t at chunk
// c ^ (t & 7): conflict-free for the per-thread row writes) and ONE thread hands the whole tile to the TMA unit
// (cp.async.bulk.tensor.2d, CU_TENSOR_MAP_SWIZZLE_128B un-swizzles on the way out; rows past `ne` are clipped by the
// tensor map) -- no copy-out loop, no per-thread global stores.  Under the CPU emulation the store is a plain loop.
template <int DM, int NEN>
__global__ void __launch_bounds__(128)
k_elem_geometry4t(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                  const double* __restrict__ dof, const int32_t* __restrict__ elems, int64_t ne,
                  const __grid_constant__ FemcyTmap tm, double* __restrict__ vol_out) {
  static_assert(NEN * 2 == 8, "k_elem_geometry4t: records of exactly 128 bytes (4 nodes, one Gauss point)");
  constexpr int TPB = 128, CH = 8;
#ifdef FEMCY_SIMT_EMU
  __shared__ double2 tile[TPB * CH];
#else
  __shared__ __align__(1024) double2 tile[TPB * CH];
#endif
  const int t = threadIdx.x;
  const int64_t e0 = blockIdx.x * (int64_t)TPB;
  const int64_t e = e0 + t;
  if (e < ne) {
    int32_t conn[NEN];
#pragma unroll
    for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
    double x[NEN][DM], g[NEN][DM];
    load_current_coords<DM, NEN>(nodes, dof, conn, x);
    double v = shape_gradients<DM, NEN>(x, tab.dN, g) * tab.w[0];
    double2* o = tile + t * CH;
    const int sw = t & 7;
#pragma unroll
    for (int a = 0; a < NEN; ++a) {
      double2 lo, hi;
      lo.x = g[a][0]; lo.y = g[a][1];
      hi.x = (DM == 3) ? g[a][DM - 1] : 0.0; hi.y = v;
      o[(2 * a) ^ sw] = lo;
      o[(2 * a + 1) ^ sw] = hi;
    }
    vol_out[e] = v;
  }
#ifdef FEMCY_SIMT_EMU
  __syncthreads();
  const int64_t rem = tm.rows - e0;
  const int nel = rem < TPB ? (int)rem : TPB;
  double2* out = reinterpret_cast<double2*>(tm.base + e0 * (NEN * 4));
  for (int gi = t; gi < nel * CH; gi += TPB) {
    int el = gi / CH, c = gi - el * CH;
    out[gi] = tile[el * CH + (c ^ (el & 7))];
  }
#else
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of the tile -> visible to the TMA unit
  __syncthreads();
  if (t == 0) {
    const uint64_t tmap = reinterpret_cast<uint64_t>(&tm);
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
    const int c0 = 0, c1 = (int)e0;
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(src) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile must outlive the read
  }
#endif
}

// Block from the accumulated gradient products  P = sum_{contributions, Gauss points} vol * (grad N_a) (grad N_b)^T :
//   K_ab[i][j] = sum_{k,l} C[voigt(i,k)][voigt(j,l)] * P[k][l]          (B_a^T C B_b is bilinear in grad N_a, grad N_b)
// valid because the tangent C is one constant matrix per run (stiffnessMtrx.py:124-129).  Cubic-form tangents (every
// tangent the reference ships, see tangent_is_cubic):  K_ij = q P_ij + r P_ji (i != j),  K_ii = p P_ii + r (tr P - P_ii).
template <int DM, bool CUBIC>
__device__ __forceinline__ void block_from_products(const double* __restrict__ C, const double (&P)[DM][DM],
                                                    double (&K)[DM][DM]) {
  constexpr int NV = Voigt<DM>::NV;
  if constexpr (CUBIC) {
    const double p = C[0], q = C[1], r = C[NV * NV - 1];
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < DM; ++i) tr += P[i][i];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) K[i][j] = (i == j) ? (p - r) * P[i][i] + r * tr : q * P[i][j] + r * P[j][i];
  } else {
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DM; ++k)
#pragma unroll
          for (int l = 0; l < DM; ++l) {
            // Voigt row of the symmetric pair: 2-D [xx, yy, xy]; 3-D [xx, yy, zz, xy, zx, yz]
            const int vik = (i == k) ? i : (DM == 2 ? 2 : (i + k == 1 ? 3 : (i + k == 2 ? 4 : 5)));
            const int vjl = (j == l) ? j : (DM == 2 ? 2 : (j + l == 1 ? 3 : (j + l == 2 ? 4 : 5)));
            s += C[vik * NV + vjl] * P[k][l];
          }
        K[i][j] = s;
      }
  }
}

// Default assembly, pass 2: per-block gather over the node-sector records with the gradient products accumulated first
// (12 FP64 instructions per contribution and Gauss point instead of 43 / 99; the tangent enters once per stored block).
// One thread per stored block; a block of 8 warps walks the block columns k = ty, ty+8, ... of ONE 32-row slice, so
// all loads of a slice's element records come from the same SM (L1 reuse) and no warp exits idle.
template <int DM, int NEN, int NGP, bool CUBIC>
__global__ void __launch_bounds__(256)
k_assemble_gather_p(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr,
                    const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                    const uint32_t* __restrict__ ent_list, const double* __restrict__ rec, double* __restrict__ val,
                    int64_t nslice) {
  constexpr int DM2 = DM * DM;
  constexpr int P = NEN * NEN;
  const int lane = threadIdx.x;
  for (int64_t s = blockIdx.x; s < nslice; s += gridDim.x) {
    const int base = slice_ptr[s];
    const int w = (slice_ptr[s + 1] - base) >> 5;
    for (int k = threadIdx.y; k < w; k += blockDim.y) {
      const int slot = base + (k << 5) + lane;
      const int beg = slot_beg[slot], end = slot_end[slot];
      double acc[DM][DM];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
      uint32_t id_next = (beg < end) ? ent_list[beg] : 0u;
      for (int t = beg; t < end; ++t) {
        const uint32_t id = id_next;
        if (t + 1 < end) id_next = ent_list[t + 1];
        const uint32_t e = id / P;
        const int p = (int)(id - e * P);
        const int a = p / NEN, b = p - a * NEN;
        const double* r1 = rec + (int64_t)e * (NEN * NGP * 4);
#pragma unroll
        for (int gp = 0; gp < NGP; ++gp) {
          const femcy_d4 ra = femcy_ld256_nc(r1 + (a * NGP + gp) * 4), rb = femcy_ld256_nc(r1 + (b * NGP + gp) * 4);
          const double sa0 = ra.w * ra.x, sa1 = ra.w * ra.y;
          acc[0][0] += sa0 * rb.x; acc[0][1] += sa0 * rb.y;
          acc[1][0] += sa1 * rb.x; acc[1][1] += sa1 * rb.y;
          if constexpr (DM == 3) {
            const double sa2 = ra.w * ra.z;
            acc[0][2] += sa0 * rb.z; acc[1][2] += sa1 * rb.z;
            acc[2][0] += sa2 * rb.x; acc[2][1] += sa2 * rb.y; acc[2][2] += sa2 * rb.z;
          }
        }
      }
      double K[DM][DM];
      block_from_products<DM, CUBIC>(tab.C, acc, K);
      double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + lane;
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) dst[(i * DM + j) << 5] = K[i][j];
    }
  }
}

// Pass 2 for elements with FOUR Gauss points (C3D10, CPS4, CPS8): a QUAD of lanes per stored block, lane j of the quad = Gauss
// point j.  The record of (element, node) is one 128-byte line holding its 4 Gauss-point sectors, so the four lanes of a quad
// read one whole line per instruction: the L1 data pipe -- the bound of the thread-per-block gather, which spends one wavefront
// per scattered 32-byte sector (ncu r2i: 95.6 % busy) -- serves a contribution with 2 wavefronts instead of 8.  Every lane
// accumulates its Gauss point's gradient products over the block's whole element list; the quad folds them once per block
// (two butterfly steps, fixed order => bit-reproducible) and shares the 9 stores.  A warp covers 8 consecutive rows of a slice
// at one block column: the stores of a plane are 64 contiguous bytes (two full sectors).
template <int DM, int NEN, bool CUBIC>
__global__ void __launch_bounds__(256)
k_assemble_gather_q(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr,
                    const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                    const uint32_t* __restrict__ ent_list, const double* __restrict__ rec, double* __restrict__ val,
                    int64_t nslice) {
  constexpr int NGP = 4;
  constexpr int DM2 = DM * DM;
  constexpr int P = NEN * NEN;
  const int lane = threadIdx.x, gp = lane & 3, q = lane >> 2;
  const int64_t s = blockIdx.x;
  if (s >= nslice) return;
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) >> 5;
  for (int task = threadIdx.y; task < 4 * w; task += blockDim.y) {
    const int rg = task & 3, k = task >> 2;
    const int slot = base + (k << 5) + (rg << 3) + q;
    const int beg = slot_beg[slot], end = slot_end[slot];
    double acc[DM][DM];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
    uint32_t id_next = (beg < end) ? ent_list[beg] : 0u;
    for (int t = beg; t < end; ++t) {
      const uint32_t id = id_next;
      if (t + 1 < end) id_next = ent_list[t + 1];
      const uint32_t e = id / P;
      const int p = (int)(id - e * P);
      const int a = p / NEN, b = p - a * NEN;
      const double* r1 = rec + (int64_t)e * (NEN * NGP * 4);
      const femcy_d4 ra = femcy_ld256_nc(r1 + (a * NGP + gp) * 4), rb = femcy_ld256_nc(r1 + (b * NGP + gp) * 4);
      const double sa0 = ra.w * ra.x, sa1 = ra.w * ra.y;
      acc[0][0] += sa0 * rb.x; acc[0][1] += sa0 * rb.y;
      acc[1][0] += sa1 * rb.x; acc[1][1] += sa1 * rb.y;
      if constexpr (DM == 3) {
        const double sa2 = ra.w * ra.z;
        acc[0][2] += sa0 * rb.z; acc[1][2] += sa1 * rb.z;
        acc[2][0] += sa2 * rb.x; acc[2][1] += sa2 * rb.y; acc[2][2] += sa2 * rb.z;
      }
    }
    double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + (slot & 31);
    if constexpr (CUBIC && DM == 3) {
      // reduce-scatter over the quad (8 instead of 18 double shuffles -- shuffles share the L1 data pipe with the loads):
      // lane 0 ends with the sums of P00, P11, P22, lane 1 with P01, P10, lane 2 with P02, P20, lane 3 with P12, P21 --
      // exactly what K_ii = (p - r) P_ii + r tr P and K_ij = q P_ij + r P_ji need.  Order: (g + g^2) + (g^1 + g^3), fixed.
      const bool hi = (gp & 2) != 0, odd = (gp & 1) != 0;
      double s0 = hi ? acc[0][0] : acc[0][2], s1 = hi ? acc[1][1] : acc[2][0], s2 = hi ? acc[2][2] : acc[1][2],
             s3 = hi ? acc[0][1] : acc[2][1], s4 = hi ? acc[1][0] : 0.0;
      s0 = __shfl_xor_sync(0xffffffffu, s0, 2); s1 = __shfl_xor_sync(0xffffffffu, s1, 2); s2 = __shfl_xor_sync(0xffffffffu, s2, 2);
      s3 = __shfl_xor_sync(0xffffffffu, s3, 2); s4 = __shfl_xor_sync(0xffffffffu, s4, 2);
      if (!hi) { acc[0][0] += s0; acc[1][1] += s1; acc[2][2] += s2; acc[0][1] += s3; acc[1][0] += s4; }
      else { acc[0][2] += s0; acc[2][0] += s1; acc[1][2] += s2; acc[2][1] += s3; }
      double t0 = !hi ? (odd ? acc[0][0] : acc[0][1]) : (odd ? acc[0][2] : acc[1][2]);
      double t1 = !hi ? (odd ? acc[1][1] : acc[1][0]) : (odd ? acc[2][0] : acc[2][1]);
      double t2 = (!hi && odd) ? acc[2][2] : 0.0;
      t0 = __shfl_xor_sync(0xffffffffu, t0, 1); t1 = __shfl_xor_sync(0xffffffffu, t1, 1); t2 = __shfl_xor_sync(0xffffffffu, t2, 1);
      constexpr int NV = Voigt<DM>::NV;
      const double cp = tab.C[0], cq = tab.C[1], cr = tab.C[NV * NV - 1];
      if (gp == 0) {
        const double p00 = acc[0][0] + t0, p11 = acc[1][1] + t1, p22 = acc[2][2] + t2;
        const double rtr = cr * ((p00 + p11) + p22);
        dst[0 << 5] = (cp - cr) * p00 + rtr; dst[4 << 5] = (cp - cr) * p11 + rtr; dst[8 << 5] = (cp - cr) * p22 + rtr;
      } else if (gp == 1) {
        const double p01 = acc[0][1] + t0, p10 = acc[1][0] + t1;
        dst[1 << 5] = cq * p01 + cr * p10; dst[3 << 5] = cq * p10 + cr * p01;
      } else if (gp == 2) {
        const double p02 = acc[0][2] + t0, p20 = acc[2][0] + t1;
        dst[2 << 5] = cq * p02 + cr * p20; dst[6 << 5] = cq * p20 + cr * p02;
      } else {
        const double p12 = acc[1][2] + t0, p21 = acc[2][1] + t1;
        dst[5 << 5] = cq * p12 + cr * p21; dst[7 << 5] = cq * p21 + cr * p12;
      }
    } else {
      // fold the four Gauss points of the quad: (g0 + g1) + (g2 + g3) in every lane
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) {
          double v = acc[i][j];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          acc[i][j] = v;
        }
      double K[DM][DM];
      block_from_products<DM, CUBIC>(tab.C, acc, K);
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j)
          if (((i * DM + j) & 3) == gp) dst[(i * DM + j) << 5] = K[i][j];
    }
  }
}

// Pass 2 for elements with ONE Gauss point (C3D4, CPS3): a PAIR of lanes per stored block.  The record of an element is one
// line of NEN 32-byte sectors; the even lane reads the sector of the row node a, the odd lane that of the column node b -- one
// L1 wavefront per contribution instead of two (the data pipe is the bound, see k_assemble_gather_q) -- and hands its gradient
// to the even lane through register shuffles.  The even lane accumulates the gradient products and writes the block.  A warp
// covers 16 consecutive rows of a slice at one block column: the stores of a plane are 128 contiguous bytes.
template <int DM, int NEN, bool CUBIC>
__global__ void __launch_bounds__(256)
k_assemble_gather_h(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr,
                    const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                    const uint32_t* __restrict__ ent_list, const double* __restrict__ rec, double* __restrict__ val,
                    int64_t nslice) {
  constexpr int DM2 = DM * DM;
  constexpr int P = NEN * NEN;
  const int lane = threadIdx.x, odd = lane & 1, q = lane >> 1;
  const int64_t s = blockIdx.x;
  if (s >= nslice) return;
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) >> 5;
  for (int task = threadIdx.y; task < 2 * w; task += blockDim.y) {
    const int rg = task & 1, k = task >> 1;
    const int slot = base + (k << 5) + (rg << 4) + q;
    const int beg = slot_beg[slot], end = slot_end[slot];
    // the pairs of a warp walk lists of different lengths: every lane takes part in every shuffle, so the loop runs to the
    // longest list of the warp and short lists idle (a = b = 0 of element 0 is a valid, unused read)
    int len = end - beg;
    int maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    double acc[DM][DM];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
    uint32_t id_next = (len > 0) ? ent_list[beg] : 0u;
    for (int t = 0; t < maxlen; ++t) {
      const bool live = t < len;
      const uint32_t id = id_next;
      if (t + 1 < len) id_next = ent_list[beg + t + 1];
      const uint32_t e = live ? id / P : 0u;
      const int p = live ? (int)(id - e * P) : 0;
      const int a = p / NEN, b = p - a * NEN;
      const femcy_d4 r = femcy_ld256_nc(rec + ((int64_t)e * NEN + (odd ? b : a)) * 4);
      const double bx = __shfl_sync(0xffffffffu, r.x, lane | 1), by = __shfl_sync(0xffffffffu, r.y, lane | 1);
      if (live && !odd) {
        const double sa0 = r.w * r.x, sa1 = r.w * r.y;
        acc[0][0] += sa0 * bx; acc[0][1] += sa0 * by;
        acc[1][0] += sa1 * bx; acc[1][1] += sa1 * by;
      }
      if constexpr (DM == 3) {
        const double bz = __shfl_sync(0xffffffffu, r.z, lane | 1);
        if (live && !odd) {
          const double sa0 = r.w * r.x, sa1 = r.w * r.y, sa2 = r.w * r.z;
          acc[0][2] += sa0 * bz; acc[1][2] += sa1 * bz;
          acc[2][0] += sa2 * bx; acc[2][1] += sa2 * by; acc[2][2] += sa2 * bz;
        }
      }
    }
    if (!odd) {
      double K[DM][DM];
      block_from_products<DM, CUBIC>(tab.C, acc, K);
      double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + (slot & 31);
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) dst[(i * DM + j) << 5] = K[i][j];
    }
  }
}
