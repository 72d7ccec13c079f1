// Rows a8 + a9: deformation gradient, constitutive laws, strain, von-Mises stress, internal force,
// elastic energy.
//
// Reference kernels replaced:
//   get_deformation_gradient                    /root/reference/stiffnessMtrx.py:532-556
//   get_strain_{small,large}Deformation         /root/reference/stiffnessMtrx.py:559-589
//   get_mises_stress_{planeStress,planeStrain,3d}  /root/reference/stiffnessMtrx.py:457-501
//   assemble_nodal_force_GN(_kernel)            /root/reference/stiffnessMtrx.py:609-644
//   get_elasEng_kernel                          /root/reference/stiffnessMtrx.py:597-606
//   LinearIsotropic.constitutiveOf*             /root/reference/material_zoo/linear_isotropic.py:35-76
//   LinearIsotropicPlaneStrain.constitutiveOf*  /root/reference/material_zoo/linear_isotropic_plane_strain.py:44-86
//   LinearIsotropicPlaneStress.constitutiveOf*  /root/reference/material_zoo/linear_isotropic_plane_stress.py:36-96
//   NeoHookean.constitutiveOf*                  /root/reference/material_zoo/neo_hookean.py:44-77
//   elasticEnergyDensity of the four classes
#include "ctx.cuh"
#include "post_kernels.cuh"

// ---- host wrappers -----------------------------------------------------------------------------
#define POST_DISPATCH(FN, ...)                                                         \
  do {                                                                                 \
    int key = ctx->dm * 1000 + ctx->n_en * 10 + ctx->n_gp;                             \
    switch (key) {                                                                     \
      case 2031: return FN<2, 3, 1>(__VA_ARGS__);                                      \
      case 2063: return FN<2, 6, 3>(__VA_ARGS__);                                      \
      case 2044: return FN<2, 4, 4>(__VA_ARGS__);                                      \
      case 2084: return FN<2, 8, 4>(__VA_ARGS__);                                      \
      case 3041: return FN<3, 4, 1>(__VA_ARGS__);                                      \
      case 3104: return FN<3, 10, 4>(__VA_ARGS__);                                     \
      default: return femcy_fail_msg(ctx, "no kernel instantiation for this (dm, n_en, n_gp)"); \
    }                                                                                  \
  } while (0)

template <int DM, int NEN, int NGP>
static int launch_defgrad(femcy_ctx* ctx) {
  if (ctx->ne == 0) return 0;
  k_defgrad<DM, NEN, NGP><<<(int)ceil_div64(ctx->ne, 128), 128, 0, ctx->stream>>>(ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF],
                                                                                ctx->elems, ctx->ne, ctx->F);
  CK_LAUNCH();
  return 0;
}
extern "C" int femcy_deformation_gradient(femcy_ctx* ctx) {
  cudaSetDevice(ctx->device);
  return femcy_for_sections(ctx, [&]() -> int {        // (a mesh of several sections, row f4: every section)
    if (!ctx->have_elem) return femcy_fail_msg(ctx, "set_element first");
    POST_DISPATCH(launch_defgrad, ctx);
  });
}

static int per_gp(femcy_ctx* ctx, int what, int large, double* out) {
  int64_t ngp = ctx->ne * ctx->n_gp;
  if (ngp == 0) return 0;
  int grid = (int)ceil_div64(ngp, 256);
  if (ctx->dm == 2) k_per_gp<2><<<grid, 256, 0, ctx->stream>>>(ctx->tab, ctx->mat_kind, large, what, ctx->F, ctx->cauchy, out, ngp);
  else k_per_gp<3><<<grid, 256, 0, ctx->stream>>>(ctx->tab, ctx->mat_kind, large, what, ctx->F, ctx->cauchy, out, ngp);
  CK_LAUNCH();
  return 0;
}
extern "C" int femcy_constitutive(femcy_ctx* ctx, int large_deform) {
  cudaSetDevice(ctx->device);
  return femcy_for_sections(ctx, [&]() -> int {
    if (!ctx->have_mat || !ctx->have_elem) return femcy_fail_msg(ctx, "set_element and set_material first");
    return per_gp(ctx, 0, large_deform, nullptr);
  });
}
extern "C" int femcy_strain(femcy_ctx* ctx, int large_deform) {
  cudaSetDevice(ctx->device);
  return femcy_for_sections(ctx, [&]() -> int {
    if (!ctx->have_elem) return femcy_fail_msg(ctx, "set_element first");
    if (!ctx->strain && femcy_alloc(ctx, &ctx->strain, ctx->ne * ctx->n_gp * ctx->dm * ctx->dm)) return 1;
    return per_gp(ctx, 1, large_deform, ctx->strain);
  });
}
extern "C" int femcy_mises(femcy_ctx* ctx) {
  cudaSetDevice(ctx->device);
  return femcy_for_sections(ctx, [&]() -> int {
    if (!ctx->have_mat || !ctx->have_elem) return femcy_fail_msg(ctx, "set_element and set_material first");
    return per_gp(ctx, 2, 0, ctx->mises);
  });
}

template <int DM, int NEN, int NGP>
static int launch_force(femcy_ctx* ctx) {
  if (ctx->ne == 0) return 0;
  k_internal_force<DM, NEN, NGP><<<(int)ceil_div64(ctx->ne, 128), 128, 0, ctx->stream>>>(
      ctx->tab, ctx->mat_kind, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, ctx->nn_own, ctx->F, ctx->cauchy,
      ctx->vol, ctx->dsdx, ctx->vec[FEMCY_VEC_NODAL_FORCE]);
  CK_LAUNCH();
  return 0;
}
extern "C" int femcy_internal_force(femcy_ctx* ctx) {
  cudaSetDevice(ctx->device);
  if (!ctx->vec[FEMCY_VEC_NODAL_FORCE]) return femcy_fail_msg(ctx, "set_element first");
  // f_int = sum over the elements of every section: one zero-fill, then one scatter-add pass per section
  CK(cudaMemsetAsync(ctx->vec[FEMCY_VEC_NODAL_FORCE], 0, (size_t)ctx->nn * ctx->dm * sizeof(double), ctx->stream));
  return femcy_for_sections(ctx, [&]() -> int {
    if (!ctx->have_mat || !ctx->have_elem) return femcy_fail_msg(ctx, "set_element and set_material first");
    POST_DISPATCH(launch_force, ctx);
  });
}

extern "C" int femcy_elastic_energy(femcy_ctx* ctx, double* total_out) {
  cudaSetDevice(ctx->device);
  double total = 0.0;                                   // (several sections: the sections' energies summed in section order)
  int rc = femcy_for_sections(ctx, [&]() -> int {
    if (!ctx->have_mat || !ctx->have_elem) return femcy_fail_msg(ctx, "set_element and set_material first");
    if (per_gp(ctx, 3, 1, ctx->energy)) return 1;
    int64_t ngp = ctx->ne * ctx->n_gp;
    int64_t g64 = ceil_div64(ngp > 0 ? ngp : 1, 1024);
    int grid = (int)(g64 > 592 ? 592 : g64);
    if (femcy_ensure_reduction_scratch(ctx, grid)) return 1;
    k_weighted_sum<<<grid, 256, 0, ctx->stream>>>(ctx->energy, ctx->vol, ngp, ctx->red_partials, ctx->red_ticket + 2, ctx->scal + 44);
    CK_LAUNCH();
    CK(cudaMemcpyAsync(ctx->h_scal + 44, ctx->scal + 44, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    total += ctx->h_scal[44];
    return 0;
  });
  if (rc) return rc;
  if (total_out) *total_out = total;
  return 0;
}

// sum of a scalar per-Gauss-point array (vol -> current mesh volume, energy, mises): an 8-byte result of a step,
// folded in a fixed order (bit-reproducible)
extern "C" int femcy_gp_sum(femcy_ctx* ctx, int which, double* total_out) {
  cudaSetDevice(ctx->device);
  const double* a = which == FEMCY_GP_VOL ? ctx->vol : which == FEMCY_GP_MISES ? ctx->mises : which == FEMCY_GP_ENERGY ? ctx->energy : nullptr;
  if (!a) return femcy_fail_msg(ctx, "femcy_gp_sum: vol, mises or energy (allocated) only");
  int64_t ngp = ctx->ne * ctx->n_gp;
  int64_t g64 = ceil_div64(ngp > 0 ? ngp : 1, 1024);
  int grid = (int)(g64 > 592 ? 592 : g64);
  if (femcy_ensure_reduction_scratch(ctx, grid)) return 1;
  k_weighted_sum<<<grid, 256, 0, ctx->stream>>>(a, nullptr, ngp, ctx->red_partials, ctx->red_ticket + 2, ctx->scal + 45);
  CK_LAUNCH();
  CK(cudaMemcpyAsync(ctx->h_scal + 45, ctx->scal + 45, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (total_out) *total_out = ctx->h_scal[45];
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Row f3: Gauss-point field -> nodes on the device.
//   ELE.extrapolate            /root/reference/element_zoo/element_*.py:202-293  (six hand-written variants; all are
//                              nodal = E . gp_values with a constant [n_en x n_gp] matrix E per element kind)
//   nodal averaging            what the reference's renderer shows by overdrawing (README future work, README.md:130)
// One thread per element forms its n_en nodal values (kept per element for ELE.extrapolate's [ne, n_en] result) and adds them
// to the node sums; a second pass divides by the number of adjacent elements.  Post-processing, not hot: fp64 atomics.
__global__ void __launch_bounds__(256)
k_extrapolate(const double* __restrict__ gp, int64_t ne, int n_gp, int ncomp, int comp, const double* __restrict__ E, int n_en,
              const int32_t* __restrict__ elems, double* __restrict__ elem_nodal, double* __restrict__ node_sum,
              double* __restrict__ node_cnt) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    double g[8];
    for (int q = 0; q < n_gp; ++q) g[q] = gp[(e * n_gp + q) * ncomp + comp];
    for (int a = 0; a < n_en; ++a) {
      double v = 0.0;
      for (int q = 0; q < n_gp; ++q) v += E[a * n_gp + q] * g[q];
      elem_nodal[e * n_en + a] = v;
      const int64_t nd = elems[e * n_en + a];
      atomicAdd(node_sum + nd, v);
      atomicAdd(node_cnt + nd, 1.0);
    }
  }
}
__global__ void k_node_mean(double* __restrict__ node_sum, const double* __restrict__ node_cnt, int64_t nn) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x)
    node_sum[i] = node_cnt[i] > 0.0 ? node_sum[i] / node_cnt[i] : 0.0;
}

extern "C" int femcy_extrapolate(femcy_ctx* ctx, int which_gp, int comp, const double* E_host, double* elem_nodal_out,
                                 double* node_mean_out) {
  cudaSetDevice(ctx->device);
  if (!ctx->have_elem) return femcy_fail_msg(ctx, "set_element first");
  if (ctx->n_gp > 8) return femcy_fail_msg(ctx, "femcy_extrapolate: at most 8 Gauss points");
  const double* src = nullptr;
  int ncomp = 1;
  const int dd = ctx->dm * ctx->dm;
  switch (which_gp) {
    case FEMCY_GP_VOL: src = ctx->vol; break;
    case FEMCY_GP_MISES: src = ctx->mises; break;
    case FEMCY_GP_ENERGY: src = ctx->energy; break;
    case FEMCY_GP_CAUCHY: src = ctx->cauchy; ncomp = dd; break;
    case FEMCY_GP_F: src = ctx->F; ncomp = dd; break;
    case FEMCY_GP_STRAIN: src = ctx->strain; ncomp = dd; break;
    default: return femcy_fail_msg(ctx, "femcy_extrapolate: vol, mises, energy, cauchy, F or strain");
  }
  if (!src) return femcy_fail_msg(ctx, "femcy_extrapolate: the field is not materialised yet");
  if (comp < 0 || comp >= ncomp) return femcy_fail_msg(ctx, "femcy_extrapolate: component out of range");
  const int64_t ne = ctx->ne, nn = ctx->nn;
  const int n_en = ctx->n_en, n_gp = ctx->n_gp;
  double *E = nullptr, *en = nullptr, *ns = nullptr, *nc = nullptr;
  if (femcy_alloc(ctx, &E, n_en * n_gp) || femcy_alloc(ctx, &en, ne * n_en) || femcy_alloc(ctx, &ns, nn) || femcy_alloc(ctx, &nc, nn)) return 1;
  int rc = 0;
  auto done = [&]() { femcy_free(&E); femcy_free(&en); femcy_free(&ns); femcy_free(&nc); };
#define EX_CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { done(); return femcy_fail(ctx, #call, _e, __FILE__, __LINE__); } } while (0)
  EX_CK(cudaMemcpyAsync(E, E_host, (size_t)(n_en * n_gp) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  EX_CK(cudaMemsetAsync(ns, 0, (size_t)nn * sizeof(double), ctx->stream));
  EX_CK(cudaMemsetAsync(nc, 0, (size_t)nn * sizeof(double), ctx->stream));
  if (ne > 0) {
    int64_t g = ceil_div64(ne, 256);
    if (g > 148 * 16) g = 148 * 16;
    k_extrapolate<<<(int)g, 256, 0, ctx->stream>>>(src, ne, n_gp, ncomp, comp, E, n_en, ctx->elems, en, ns, nc);
    ctx->launches++;
    EX_CK(cudaGetLastError());
  }
  {
    int64_t g = ceil_div64(nn > 0 ? nn : 1, 256);
    if (g > 148 * 16) g = 148 * 16;
    k_node_mean<<<(int)g, 256, 0, ctx->stream>>>(ns, nc, nn);
    ctx->launches++;
    EX_CK(cudaGetLastError());
  }
  if (elem_nodal_out && ne > 0) EX_CK(cudaMemcpyAsync(elem_nodal_out, en, (size_t)(ne * n_en) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (node_mean_out && nn > 0) EX_CK(cudaMemcpyAsync(node_mean_out, ns, (size_t)nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  EX_CK(cudaStreamSynchronize(ctx->stream));
#undef EX_CK
  done();
  return rc;
}
