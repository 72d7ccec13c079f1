// Constitutive laws on one deformation gradient (material_zoo of the reference: linear_isotropic.py:35-76,
// linear_isotropic_plane_strain.py:44-86, linear_isotropic_plane_stress.py:36-96, neo_hookean.py:44-77) -- shared by the
// stress-recovery / internal-force kernels (post_kernels.cuh) and by the consistent tangent of the assembly
// (assembly_kernels.cuh: spatial_tangent differentiates exactly these functions).
#pragma once
#include "device_compat.cuh"
#include "kernel_types.cuh"
#include "elem_math.cuh"

enum { MAT_ISO3D = 0, MAT_PSTRAIN = 1, MAT_PSTRESS = 2, MAT_NEOHOOKE = 3 };

// ---- constitutive laws on one F -------------------------------------------------------------
// 3-D kinds
__device__ __forceinline__ void sigma_3d(const ElemTables& tab, int kind, int large, const double (&F)[3][3],
                                         double (&S)[3][3]) {
  if (kind == MAT_NEOHOOKE) {
    double C1 = tab.mat[0], D1 = tab.mat[1];
    double J = det_dm<3>(F);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double b = F[i][0] * F[j][0] + F[i][1] * F[j][1] + F[i][2] * F[j][2];  // B = F F^T
        double eye = (i == j) ? 1.0 : 0.0;
        S[i][j] = 2.0 * C1 / J * (b - eye) + 2.0 * D1 * (J - 1.0) * eye;
      }
    return;
  }
  double E[3][3];
  if (!large) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) E[i][j] = (F[i][j] + F[j][i]) / 2.0 - ((i == j) ? 1.0 : 0.0);
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        E[i][j] = (F[0][i] * F[0][j] + F[1][i] * F[1][j] + F[2][i] * F[2][j] - ((i == j) ? 1.0 : 0.0)) / 2.0;
  }
  double ev[6] = {E[0][0], E[1][1], E[2][2], 2.0 * E[0][1], 2.0 * E[2][0], 2.0 * E[1][2]};
  double s[6];
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) t += tab.C[p * 6 + q] * ev[q];
    s[p] = t;
  }
  double P2[3][3] = {{s[0], s[3], s[4]}, {s[3], s[1], s[5]}, {s[4], s[5], s[2]}};
  if (!large) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) S[i][j] = P2[i][j];
    return;
  }
  double J = det_dm<3>(F);
  double FP[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) FP[i][j] = F[i][0] * P2[0][j] + F[i][1] * P2[1][j] + F[i][2] * P2[2][j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) S[i][j] = (FP[i][0] * F[j][0] + FP[i][1] * F[j][1] + FP[i][2] * F[j][2]) / J;
}

// 2-D kinds
__device__ __forceinline__ void sigma_2d(const ElemTables& tab, int kind, int large, const double (&F)[2][2],
                                         double (&S)[2][2]) {
  if (kind == MAT_PSTRAIN) {
    double E[2][2];
    if (!large) {
      E[0][0] = F[0][0] - 1.0; E[1][1] = F[1][1] - 1.0;
      E[0][1] = E[1][0] = (F[0][1] + F[1][0]) / 2.0;
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) E[i][j] = (F[0][i] * F[0][j] + F[1][i] * F[1][j] - ((i == j) ? 1.0 : 0.0)) / 2.0;
    }
    double ev[3] = {E[0][0], E[1][1], E[0][1] + E[1][0]};
    double s[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) s[p] = tab.C[p * 3 + 0] * ev[0] + tab.C[p * 3 + 1] * ev[1] + tab.C[p * 3 + 2] * ev[2];
    double P2[2][2] = {{s[0], s[2]}, {s[2], s[1]}};
    if (!large) { S[0][0] = P2[0][0]; S[0][1] = P2[0][1]; S[1][0] = P2[1][0]; S[1][1] = P2[1][1]; return; }
    double J = det_dm<2>(F);
    double FP[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) FP[i][j] = F[i][0] * P2[0][j] + F[i][1] * P2[1][j];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) S[i][j] = (FP[i][0] * F[j][0] + FP[i][1] * F[j][1]) / J;
    return;
  }
  // plane stress: embed in 3-D with F33 from nu; uses its own C_6x6 (not ddsdde), rows zz/zx/yz zero
  double Em = tab.mat[0], nu = tab.mat[1];
  double c00 = Em / (1.0 - nu * nu), c01 = c00 * nu, G = Em / 2.0 / (1.0 + nu);
  double F33 = -nu / (1.0 - nu) * (F[0][0] + F[1][1] - 2.0) + 1.0;
  double E00, E11, E01;
  if (!large) {
    E00 = F[0][0] - 1.0; E11 = F[1][1] - 1.0; E01 = (F[0][1] + F[1][0]) / 2.0;
  } else {
    E00 = (F[0][0] * F[0][0] + F[1][0] * F[1][0] - 1.0) / 2.0;
    E11 = (F[0][1] * F[0][1] + F[1][1] * F[1][1] - 1.0) / 2.0;
    E01 = (F[0][0] * F[0][1] + F[1][0] * F[1][1]) / 2.0;
  }
  double s0 = c00 * E00 + c01 * E11, s1 = c01 * E00 + c00 * E11, s3 = G * (2.0 * E01);
  if (!large) { S[0][0] = s0; S[0][1] = s3; S[1][0] = s3; S[1][1] = s1; return; }
  double P2[2][2] = {{s0, s3}, {s3, s1}};
  double J = det_dm<2>(F) * F33;
  double FP[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) FP[i][j] = F[i][0] * P2[0][j] + F[i][1] * P2[1][j];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) S[i][j] = (FP[i][0] * F[j][0] + FP[i][1] * F[j][1]) / J;
}

template <int DM>
__device__ __forceinline__ void sigma_of_F(const ElemTables& tab, int kind, int large, const double (&F)[DM][DM],
                                           double (&S)[DM][DM]) {
  if constexpr (DM == 2) sigma_2d(tab, kind, large, F, S);
  else sigma_3d(tab, kind, large, F, S);
}
