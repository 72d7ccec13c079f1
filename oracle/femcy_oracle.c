/* CPU restatement of the reference's hot-path algorithm in C + OpenMP (TEST INFRASTRUCTURE ONLY:
 * used by tests/ as a checker at sizes NumPy is slow at, and by bench.py as the `cpu_baseline` /
 * `--impl reference` arm -- Taichi is not installable here, SURVEY H7, so the reference's
 * arch=cpu path is restated kernel by kernel with the same data layout and the same work per
 * thread).  Nothing under femcy_b200/ links or loads this file.
 *
 * Data layout = the reference's: ELL rows  sparseIJ[N][W+1] (count first, -1 padding) and
 * sparseMtrx_rowMajor[N][W]                       /root/reference/stiffnessMtrx.py:78-95
 *
 *   oracle_dsdx_vol      get_dsdx_and_vol         /root/reference/stiffnessMtrx.py:132-150
 *   oracle_assemble_ell  assemble_stiffnessMtrx   /root/reference/stiffnessMtrx.py:161-186
 *                        (+ sparseMatrix_get_j :414-420: full row scan, no early exit)
 *   oracle_pcg_ell       ConjugateGradientSolver_rowMajor.solve
 *                                                 /root/reference/conjugateGradientSolver.py:48-127
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXEN 10
#define MAXDM 3
#define MAXV 6

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static double inv_small(int dm, const double* J, double* Ji) {
  if (dm == 2) {
    double det = J[0] * J[3] - J[1] * J[2];
    Ji[0] = J[3] / det; Ji[1] = -J[1] / det; Ji[2] = -J[2] / det; Ji[3] = J[0] / det;
    return det;
  }
  double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
  double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
  Ji[0] = c00 / det; Ji[1] = (J[2] * J[7] - J[1] * J[8]) / det; Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
  Ji[3] = c01 / det; Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det; Ji[5] = (J[2] * J[3] - J[0] * J[5]) / det;
  Ji[6] = c02 / det; Ji[7] = (J[1] * J[6] - J[0] * J[7]) / det; Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
  return det;
}

/* one thread per element, serial over Gauss points (stiffnessMtrx.py:137-150) */
void oracle_dsdx_vol(int dm, int n_en, int n_gp, int64_t ne, const double* nodes, const double* dof,
                     const int32_t* elems, const double* dN /*[n_gp][n_en][dm]*/, const double* w,
                     double* dsdx /*[ne][n_gp][n_en][dm]*/, double* vol /*[ne][n_gp]*/) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < ne; ++e) {
    double x[MAXEN][MAXDM];
    for (int a = 0; a < n_en; ++a) {
      int64_t n = elems[e * n_en + a];
      for (int i = 0; i < dm; ++i) x[a][i] = nodes[n * dm + i] + dof[n * dm + i];
    }
    for (int g = 0; g < n_gp; ++g) {
      const double* d = dN + (size_t)g * n_en * dm;
      double J[9], Ji[9];
      for (int i = 0; i < dm; ++i)
        for (int k = 0; k < dm; ++k) {
          double s = 0.0;
          for (int a = 0; a < n_en; ++a) s += x[a][i] * d[a * dm + k];
          J[i * dm + k] = s;
        }
      double det = inv_small(dm, J, Ji);
      double* o = dsdx + ((size_t)(e * n_gp + g)) * n_en * dm;
      for (int a = 0; a < n_en; ++a)
        for (int j = 0; j < dm; ++j) {
          double s = 0.0;
          for (int k = 0; k < dm; ++k) s += d[a * dm + k] * Ji[k * dm + j];
          o[a * dm + j] = s;
        }
      vol[e * n_gp + g] = det * w[g];
    }
  }
}

static void strain_matrix(int dm, int n_en, const double* g /*[n_en][dm]*/, double* B /*[n_v][n_edof]*/) {
  int n_edof = n_en * dm, n_v = dm == 2 ? 3 : 6;
  memset(B, 0, sizeof(double) * n_v * n_edof);
  for (int a = 0; a < n_en; ++a) {
    if (dm == 2) {
      B[0 * n_edof + a * 2 + 0] = g[a * 2 + 0];
      B[1 * n_edof + a * 2 + 1] = g[a * 2 + 1];
      B[2 * n_edof + a * 2 + 0] = g[a * 2 + 1];
      B[2 * n_edof + a * 2 + 1] = g[a * 2 + 0];
    } else {
      B[0 * n_edof + a * 3 + 0] = g[a * 3 + 0];
      B[1 * n_edof + a * 3 + 1] = g[a * 3 + 1];
      B[2 * n_edof + a * 3 + 2] = g[a * 3 + 2];
      B[3 * n_edof + a * 3 + 0] = g[a * 3 + 1]; B[3 * n_edof + a * 3 + 1] = g[a * 3 + 0];
      B[4 * n_edof + a * 3 + 0] = g[a * 3 + 2]; B[4 * n_edof + a * 3 + 2] = g[a * 3 + 0];
      B[5 * n_edof + a * 3 + 1] = g[a * 3 + 2]; B[5 * n_edof + a * 3 + 2] = g[a * 3 + 1];
    }
  }
}

/* one thread per (element, Gauss point): dense B^T (C B) vol, then n_edof^2 searched atomic adds */
void oracle_assemble_ell(int dm, int n_en, int n_gp, int64_t ne, int64_t N, int W, const int32_t* elems,
                         const double* dsdx, const double* vol, const double* C /*[n_v][n_v]*/,
                         const int32_t* ij /*[N][W+1]*/, double* spm /*[N][W]*/) {
  int n_edof = n_en * dm, n_v = dm == 2 ? 3 : 6;
  memset(spm, 0, sizeof(double) * (size_t)N * W);  /* K.fill(0), :168 */
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < ne * n_gp; ++t) {
    int64_t e = t / n_gp;
    double B[MAXV * MAXEN * MAXDM], CB[MAXV * MAXEN * MAXDM];
    strain_matrix(dm, n_en, dsdx + (size_t)t * n_en * dm, B);
    for (int p = 0; p < n_v; ++p)
      for (int c = 0; c < n_edof; ++c) {
        double s = 0.0;
        for (int q = 0; q < n_v; ++q) s += C[p * n_v + q] * B[q * n_edof + c];
        CB[p * n_edof + c] = s;
      }
    int32_t Js[MAXEN * MAXDM];
    for (int a = 0; a < n_en; ++a)
      for (int i = 0; i < dm; ++i) Js[a * dm + i] = elems[e * n_en + a] * dm + i;
    double v = vol[t];
    for (int r = 0; r < n_edof; ++r) {
      int32_t ig = Js[r];
      const int32_t* row = ij + (size_t)ig * (W + 1);
      int cnt = row[0];
      for (int c = 0; c < n_edof; ++c) {
        double bcb = 0.0;
        for (int p = 0; p < n_v; ++p) bcb += B[p * n_edof + r] * CB[p * n_edof + c];
        int32_t jg = Js[c];
        int jl = 0;
        for (int j = 0; j < cnt; ++j)  /* sparseMatrix_get_j: scans the whole row */
          if (row[j + 1] == jg) jl = j;
        double add = bcb * v;
#pragma omp atomic
        spm[(size_t)ig * W + jl] += add;
      }
    }
  }
}

/* Jacobi-PCG on the ELL rows, kernel by kernel as the reference launches them */
int64_t oracle_pcg_ell(int64_t N, int W, const double* A, const int32_t* ij, const double* b, double* x, double eps,
                       int64_t max_iter, int fixed_iters, double* rmax0_out, double* rmax_out) {
  double* r = (double*)malloc(sizeof(double) * N);
  double* d = (double*)malloc(sizeof(double) * N);
  double* M = (double*)malloc(sizeof(double) * N);
  double* Ad = (double*)malloc(sizeof(double) * N);
  double r0 = 0.0;
#pragma omp parallel for schedule(static) reduction(max : r0)
  for (int64_t i = 0; i < N; ++i) {
    const int32_t* row = ij + (size_t)i * (W + 1);
    int t = 0;
    for (int j = 0; j < row[0]; ++j)
      if (row[j + 1] == i) t = j;                 /* A_get :40-46 */
    M[i] = 1.0 / A[(size_t)i * W + t];            /* M_init :48-51 */
    x[i] = 0.0;
    r[i] = b[i];                                   /* r_d_init :60-65 */
    d[i] = M[i] * r[i];
    double a = fabs(r[i]);
    if (a > r0) r0 = a;
  }
  int64_t it = 0;
  double rmax = r0;
  /* one parallel region, one static row partition for every loop (each "kernel" of the reference
   * is an omp-for; the implicit barriers are the kernel boundaries) */
  double rMr = 0.0, dAd = 0.0, rMr2 = 0.0, rm = 0.0;
  int stop = 0;
#pragma omp parallel
  {
    for (int64_t k = 0; k < max_iter; ++k) {
#pragma omp single
      { rMr = 0.0; dAd = 0.0; rMr2 = 0.0; rm = 0.0; }
#pragma omp for schedule(static)
      for (int64_t i = 0; i < N; ++i) {             /* compute_Ad :53-58 */
        const int32_t* row = ij + (size_t)i * (W + 1);
        const double* a = A + (size_t)i * W;
        double s = 0.0;
        for (int j = 0; j < row[0]; ++j) s = s + a[j] * d[row[j + 1]];
        Ad[i] = s;
      }
#pragma omp for schedule(static) reduction(+ : rMr)
      for (int64_t i = 0; i < N; ++i) rMr += r[i] * M[i] * r[i];      /* compute_rMr :74-79 */
#pragma omp for schedule(static) reduction(+ : dAd)
      for (int64_t i = 0; i < N; ++i) dAd += d[i] * Ad[i];            /* dot_product :96-101 */
      double alpha = rMr / dAd;
#pragma omp for schedule(static) nowait
      for (int64_t i = 0; i < N; ++i) x[i] = x[i] + alpha * d[i];     /* update_x :81-84 */
#pragma omp for schedule(static) nowait
      for (int64_t i = 0; i < N; ++i) r[i] = r[i] - alpha * Ad[i];    /* update_r :86-89 */
#pragma omp for schedule(static) reduction(+ : rMr2)
      for (int64_t i = 0; i < N; ++i) rMr2 += r[i] * M[i] * r[i];
      double beta = rMr2 / rMr;
#pragma omp for schedule(static) nowait
      for (int64_t i = 0; i < N; ++i) d[i] = M[i] * r[i] + beta * d[i];  /* update_d :91-94 */
#pragma omp for schedule(static) reduction(max : rm)
      for (int64_t i = 0; i < N; ++i) {                                 /* rmax :67-72 */
        double a = fabs(r[i]);
        if (a > rm) rm = a;
      }
#pragma omp single
      {
        rmax = rm;
        it = k + 1;
        if (!fixed_iters && rmax < eps * r0) stop = 1;                  /* :124 */
      }
      if (stop) break;
    }
  }
  if (rmax0_out) *rmax0_out = r0;
  if (rmax_out) *rmax_out = rmax;
  free(r); free(d); free(M); free(Ad);
  return it;
}
