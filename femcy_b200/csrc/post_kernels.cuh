// Device code of rows a8 + a9 (deformation gradient, constitutive laws, strain, von-Mises stress, internal force,
// elastic energy); header for the same reason as assembly_kernels.cuh (host launch code in post.cu, CPU SIMT
// emulation in tests/simt).  Reference kernels: see post.cu.
#pragma once
#include "device_compat.cuh"
#include "kernel_types.cuh"
#include "elem_math.cuh"
#include "constitutive.cuh"

template <int DM>
__device__ __forceinline__ double mises_of(const ElemTables& tab, int kind, const double (&S)[DM][DM]) {
  double s[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int j = 0; j < DM; ++j) s[i][j] = S[i][j];
  if (DM == 2 && kind == MAT_PSTRAIN) s[2][2] = tab.mat[1] * (S[0][0] + S[1][1]);
  double tr = (s[0][0] + s[1][1] + s[2][2]) / 3.0;
  double sum = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double d = s[i][j] - ((i == j) ? tr : 0.0);
      sum += d * d;
    }
  return sqrt(3.0 / 2.0 * sum);
}

template <int DM>
__device__ __forceinline__ double energy_of(const ElemTables& tab, int kind, const double (&F)[DM][DM]) {
  if constexpr (DM == 3) {
    if (kind == MAT_NEOHOOKE) {
      double J = det_dm<3>(F);
      double trB = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) trB += F[i][j] * F[i][j];
      return tab.mat[0] * (trB - 3.0 - 2.0 * log(J)) + tab.mat[1] * (J - 1.0) * (J - 1.0);
    }
    double E[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        E[i][j] = (F[0][i] * F[0][j] + F[1][i] * F[1][j] + F[2][i] * F[2][j] - ((i == j) ? 1.0 : 0.0)) / 2.0;
    double ev[6] = {E[0][0], E[1][1], E[2][2], 2.0 * E[0][1], 2.0 * E[2][0], 2.0 * E[1][2]};
    double tot = 0.0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 6; ++q) t += tab.C[p * 6 + q] * ev[q];
      tot += ev[p] * t;
    }
    return tot / 2.0;
  } else {
    double Em = tab.mat[0], nu = tab.mat[1];
    double G = Em / 2.0 / (1.0 + nu);
    double c00, c01, F33;
    if (kind == MAT_PSTRAIN) {
      double t1 = Em / (1.0 + nu), t2 = nu / (fabs(1.0 - 2.0 * nu) + 1.e-30);
      c00 = t1 * (1.0 + t2); c01 = t1 * t2; F33 = 1.0;
    } else {
      c00 = Em / (1.0 - nu * nu); c01 = c00 * nu;
      F33 = -nu / (1.0 - nu) * (F[0][0] + F[1][1] - 2.0) + 1.0;
    }
    double E00 = (F[0][0] * F[0][0] + F[1][0] * F[1][0] - 1.0) / 2.0;
    double E11 = (F[0][1] * F[0][1] + F[1][1] * F[1][1] - 1.0) / 2.0;
    double E01 = (F[0][0] * F[0][1] + F[1][0] * F[1][1]) / 2.0;
    double E22 = (F33 * F33 - 1.0) / 2.0;
    // C_6x6 of the two plane classes: zz row/col carries c01 couplings only for plane strain,
    // C[2][2] = 0 in both (linear_isotropic_plane_strain.py:30-39, ..._plane_stress.py:22-31)
    double s0 = c00 * E00 + c01 * E11, s1 = c01 * E00 + c00 * E11, s2 = 0.0;
    if (kind == MAT_PSTRAIN) { s0 += c01 * E22; s1 += c01 * E22; s2 = c01 * (E00 + E11); }
    double g01 = 2.0 * E01;
    return (E00 * s0 + E11 * s1 + E22 * s2 + g01 * G * g01) / 2.0;
  }
}

// ---- kernels --------------------------------------------------------------------------------
template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(128)
k_defgrad(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes, const double* __restrict__ dof,
          const int32_t* __restrict__ elems, int64_t ne, double* __restrict__ Fout) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  double X[NEN][DM], u[NEN][DM];
#pragma unroll
  for (int a = 0; a < NEN; ++a) {
    int64_t n = elems[e * NEN + a];
#pragma unroll
    for (int i = 0; i < DM; ++i) { X[a][i] = nodes[n * DM + i]; u[a][i] = dof[n * DM + i]; }
  }
#pragma unroll 1
  for (int gp = 0; gp < NGP; ++gp) {
    double g[NEN][DM];
    shape_gradients<DM, NEN>(X, &tab.dN[gp * NEN * DM], g);
    double* o = Fout + (e * NGP + gp) * (DM * DM);
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < NEN; ++a) s += u[a][i] * g[a][j];
        o[i * DM + j] = s + ((i == j) ? 1.0 : 0.0);
      }
  }
}

// what: 0 constitutive -> cauchy ; 1 strain ; 2 mises (from cauchy) ; 3 energy density (from F)
template <int DM>
__global__ void __launch_bounds__(256)
k_per_gp(const __grid_constant__ ElemTables tab, int kind, int large, int what, const double* __restrict__ Fin,
         double* __restrict__ cauchy, double* __restrict__ out, int64_t ngp) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= ngp) return;
  constexpr int DD = DM * DM;
  double F[DM][DM];
  if (what != 2) {
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) F[i][j] = Fin[t * DD + i * DM + j];
  }
  if (what == 0) {
    double S[DM][DM];
    sigma_of_F<DM>(tab, kind, large, F, S);
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) cauchy[t * DD + i * DM + j] = S[i][j];
  } else if (what == 1) {
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        double v;
        if (!large) v = (F[i][j] + F[j][i]) / 2.0 - ((i == j) ? 1.0 : 0.0);
        else {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < DM; ++k) s += F[k][i] * F[k][j];
          v = (s - ((i == j) ? 1.0 : 0.0)) / 2.0;
        }
        out[t * DD + i * DM + j] = v;
      }
  } else if (what == 2) {
    double S[DM][DM];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) S[i][j] = cauchy[t * DD + i * DM + j];
    out[t] = mises_of<DM>(tab, kind, S);
  } else {
    out[t] = energy_of<DM>(tab, kind, F);
  }
}

// F -> sigma(large) -> grad N, vol on X+u -> nodal force scatter
template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(128)
k_internal_force(const __grid_constant__ ElemTables tab, int kind, const double* __restrict__ nodes,
                 const double* __restrict__ dof, const int32_t* __restrict__ elems, int64_t ne, int64_t nn_own,
                 double* __restrict__ Fout, double* __restrict__ cauchy, double* __restrict__ vol,
                 double* __restrict__ dsdx, double* __restrict__ force) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int32_t conn[NEN];
  double X[NEN][DM], u[NEN][DM];
#pragma unroll
  for (int a = 0; a < NEN; ++a) {
    conn[a] = elems[e * NEN + a];
    int64_t n = conn[a];
#pragma unroll
    for (int i = 0; i < DM; ++i) { X[a][i] = nodes[n * DM + i]; u[a][i] = dof[n * DM + i]; }
  }
  double f[NEN][DM];
#pragma unroll
  for (int a = 0; a < NEN; ++a)
#pragma unroll
    for (int i = 0; i < DM; ++i) f[a][i] = 0.0;
#pragma unroll 1
  for (int gp = 0; gp < NGP; ++gp) {
    double g[NEN][DM];
    shape_gradients<DM, NEN>(X, &tab.dN[gp * NEN * DM], g);
    double F[DM][DM], S[DM][DM];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < NEN; ++a) s += u[a][i] * g[a][j];
        F[i][j] = s + ((i == j) ? 1.0 : 0.0);
      }
    sigma_of_F<DM>(tab, kind, 1, F, S);
    int64_t o = (e * NGP + gp) * (DM * DM);
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) { Fout[o + i * DM + j] = F[i][j]; cauchy[o + i * DM + j] = S[i][j]; }
    // current configuration
    double x[NEN][DM];
#pragma unroll
    for (int a = 0; a < NEN; ++a)
#pragma unroll
      for (int i = 0; i < DM; ++i) x[a][i] = X[a][i] + u[a][i];
    double v = shape_gradients<DM, NEN>(x, &tab.dN[gp * NEN * DM], g) * tab.w[gp];
    vol[e * NGP + gp] = v;
    if (dsdx) {
      double* od = dsdx + (e * NGP + gp) * (NEN * DM);
#pragma unroll
      for (int a = 0; a < NEN; ++a)
#pragma unroll
        for (int j = 0; j < DM; ++j) od[a * DM + j] = g[a][j];
    }
#pragma unroll
    for (int a = 0; a < NEN; ++a)
#pragma unroll
      for (int i = 0; i < DM; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < DM; ++j) s += g[a][j] * S[j][i];
        f[a][i] += s * v;
      }
  }
#pragma unroll
  for (int a = 0; a < NEN; ++a) {
    if (conn[a] < nn_own) {
#pragma unroll
      for (int i = 0; i < DM; ++i) atomicAdd(&force[(int64_t)conn[a] * DM + i], f[a][i]);
    }
  }
}

__global__ void __launch_bounds__(256)
k_weighted_sum(const double* __restrict__ a, const double* __restrict__ w, int64_t n, double* partials,
               unsigned int* ticket, double* out) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += w ? a[i] * w[i] : a[i];                      // w == nullptr: plain sum (femcy_gp_sum)
  double mine[1] = {s}, tot[1];
  const bool is_max[1] = {false};
  if (grid_reduce<1>(mine, partials, ticket, tot, is_max)) *out = tot[0];
}

