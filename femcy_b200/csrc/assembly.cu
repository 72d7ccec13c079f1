// Hot path rows a2 + a3: shape-function gradients / volumes and the stiffness assembly.
//
// Reference kernels replaced:
//   System_of_equations.get_dsdx_and_vol        /root/reference/stiffnessMtrx.py:132-150
//   System_of_equations.assemble_stiffnessMtrx  /root/reference/stiffnessMtrx.py:161-186
//   System_of_equations.sparseMatrix_get_j      /root/reference/stiffnessMtrx.py:414-420  (row scan -> precomputed slot)
//
// Two formulations over the same node-block SELL-32 matrix (device code: assembly_kernels.cuh; numbers of
// femcy_assemble_K in include/femcy_b200.h; measurements: DESIGN.md section 4):
//   1 scatter  thread (or warp, n_en >= 8) per element; K_e accumulated over the Gauss points in registers, then dm*dm
//              fp64 atomic adds (RED.ADD.F64) per node pair into the precomputed slot.
//   2 gather   (default) pass 1 writes one 32-byte sector (grad N_a, vol) per (element, node, Gauss point) -- through a TMA
//              tensor store for C3D4 --, pass 2 runs one thread per stored block over its element list: no atomics, no
//              zero-fill, bit-reproducible.
// The 20 other variants of round 1 were measured on a B200 at the start of round 2 (profiles/r2a_quick_ab.jsonl) and
// removed.
#include "ctx.cuh"
#include "assembly_kernels.cuh"

// ---------------------------------------------------------------------------------------------
template <int DM, int NEN, int NGP>
static int launch_dsdx(femcy_ctx* ctx, bool want_dsdx) {
  if (want_dsdx && !ctx->dsdx) {
    if (femcy_alloc(ctx, &ctx->dsdx, ctx->ne * NGP * NEN * DM)) return 1;
  }
  if (ctx->ne == 0) return 0;
  int grid = (int)ceil_div64(ctx->ne, 128);
  k_dsdx_vol<DM, NEN, NGP><<<grid, 128, 0, ctx->stream>>>(ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems,
                                                         ctx->ne, want_dsdx ? ctx->dsdx : nullptr, ctx->vol);
  CK_LAUNCH();
  return 0;
}

// tensor map of the C3D4 record array for the TMA store of pass 1 (encoded once per record buffer)
static int record_tensor_map(femcy_ctx* ctx, FemcyTmap* tm) {
  if (ctx->egeo4_tmap_for == ctx->egeo4 && ctx->egeo4_tmap_ne == ctx->ne) { *tm = ctx->egeo4_tmap; return 0; }
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
    return femcy_fail_msg(ctx, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {16, (cuuint64_t)ctx->ne};           // 16 doubles (one 128 B record) x ne records
  cuuint64_t gstr[1] = {128};                               // bytes between records
  cuuint32_t box[2] = {16, 128};                            // one block's tile: 128 records
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = ((EncodeTiled)fn)(&ctx->egeo4_tmap.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ctx->egeo4, gdim, gstr, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return femcy_fail_msg(ctx, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)cr) + ")");
  ctx->egeo4_tmap_for = ctx->egeo4;
  ctx->egeo4_tmap_ne = ctx->ne;
  *tm = ctx->egeo4_tmap;
  return 0;
}

// option consistent_tangent (row f2): exact linearisation of the internal force, scatter-add
template <int DM, int NEN, int NGP>
static int launch_assemble_ct(femcy_ctx* ctx, bool zero_fill) {
  BsellPattern& P = ctx->P;
  if (!ctx->elem_slot) return femcy_fail_msg(ctx, "build_pattern first");
  if (zero_fill) CK(cudaMemsetAsync(P.val, 0, (size_t)(P.nslots * DM * DM) * sizeof(double), ctx->stream));
  if (ctx->ne == 0) return 0;
  k_assemble_scatter_ct<DM, NEN, NGP><<<(int)ceil_div64(ctx->ne, 128), 128, 0, ctx->stream>>>(
      ctx->tab, ctx->mat_kind, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->elem_slot, ctx->ne, P.val);
  CK_LAUNCH();
  return 0;
}

template <int DM, int NEN, int NGP>
static int launch_assemble(femcy_ctx* ctx, int variant, bool zero_fill) {
  BsellPattern& P = ctx->P;
  constexpr int DM2 = DM * DM;
  if (ctx->opt.consistent_tangent && variant <= FEMCY_ASSEMBLY_SCATTER) return launch_assemble_ct<DM, NEN, NGP>(ctx, zero_fill);
  if (ctx->ne == 0) return 0;
  const bool gather_ok = ctx->ent_list != nullptr;
  // default: the atomic-free gather (measured on B200, profiles/r2a + r2i: 2.48 vs 3.27 ms on 10.1 M C3D4,
  // 4.16 vs 8.9 ms on 1.0 M C3D10); the scatter is kept as the atomic scatter-add formulation it is compared with
  if (variant == 0) variant = gather_ok ? FEMCY_ASSEMBLY_GATHER : FEMCY_ASSEMBLY_SCATTER;
  const bool variant_q = (variant != 3);      // 3 = the gather with one THREAD per block also for 4-Gauss-point elements (A/B)
  if (variant == 3) variant = FEMCY_ASSEMBLY_GATHER;
  if (variant == FEMCY_ASSEMBLY_GATHER) {
    if (!gather_ok) return femcy_fail_msg(ctx, "gather assembly needs the element lists of build_pattern");
    if (!ctx->egeo4) {
      if (femcy_alloc(ctx, &ctx->egeo4, ctx->ne * NEN * NGP * 4)) return 1;
    }
    // pass 1: node-sector records rec[e][a][gp] = (grad N_a, vol_gp) + the vol field
    bool tma_store = false;
    if constexpr (NEN == 4 && NGP == 1) {
      // C3D4 / CPS4-like 128-byte records: one TMA tensor store per block of 128 records
      FemcyTmap tm;
      if (record_tensor_map(ctx, &tm)) return 1;
      k_elem_geometry4t<DM, NEN><<<(int)ceil_div64(ctx->ne, 128), 128, 0, ctx->stream>>>(
          ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, tm, ctx->vol);
      tma_store = true;
    }
    if (!tma_store) {
      using G = Geo4Cfg<NEN, NGP>;
      k_elem_geometry4s<DM, NEN, NGP><<<(int)ceil_div64(ctx->ne, G::TPB), G::TPB, 0, ctx->stream>>>(
          ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, ctx->egeo4, ctx->vol);
    }
    CK_LAUNCH();
    // pass 2: one block of 8 warps per 32-row slice; four-Gauss-point elements: a quad of lanes per stored block
    if constexpr (NGP == 4) {
      if (variant_q) {
        if (tangent_is_cubic(ctx->tab.C, DM))
          k_assemble_gather_q<DM, NEN, true><<<(unsigned)P.nslice, dim3(32, 8), 0, ctx->stream>>>(
              ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, P.nslice);
        else
          k_assemble_gather_q<DM, NEN, false><<<(unsigned)P.nslice, dim3(32, 8), 0, ctx->stream>>>(
              ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, P.nslice);
        CK_LAUNCH();
        return 0;
      }
    }
    if constexpr (NGP == 1) {
      if (variant_q) {      // one-Gauss-point elements: a pair of lanes per stored block
        if (tangent_is_cubic(ctx->tab.C, DM))
          k_assemble_gather_h<DM, NEN, true><<<(unsigned)P.nslice, dim3(32, 8), 0, ctx->stream>>>(
              ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, P.nslice);
        else
          k_assemble_gather_h<DM, NEN, false><<<(unsigned)P.nslice, dim3(32, 8), 0, ctx->stream>>>(
              ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, P.nslice);
        CK_LAUNCH();
        return 0;
      }
    }
    if (tangent_is_cubic(ctx->tab.C, DM))
      k_assemble_gather_p<DM, NEN, NGP, true><<<(unsigned)P.nslice, dim3(32, 8), 0, ctx->stream>>>(
          ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, P.nslice);
    else
      k_assemble_gather_p<DM, NEN, NGP, false><<<(unsigned)P.nslice, dim3(32, 8), 0, ctx->stream>>>(
          ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, P.nslice);
    CK_LAUNCH();
    return 0;
  }
  if (variant != FEMCY_ASSEMBLY_SCATTER) return femcy_fail_msg(ctx, "unknown assembly variant (1 = scatter, 2 = gather)");
  if (zero_fill) CK(cudaMemsetAsync(P.val, 0, (size_t)(P.nslots * DM2) * sizeof(double), ctx->stream));  // K.fill(0), :168
  if constexpr (NEN >= 8) {
    // one warp per element (C3D10, CPS8/CPE8)
    int64_t blocks = ceil_div64(ctx->ne, 4);
    if (blocks > 148 * 64) blocks = 148 * 64;
    k_assemble_scatter_warp<DM, NEN, NGP><<<(int)blocks, 128, 0, ctx->stream>>>(
        ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->elem_slot, ctx->ne, P.val);
  } else {
    k_assemble_scatter<DM, NEN, NGP, 1><<<(int)ceil_div64(ctx->ne, 128), 128, 0, ctx->stream>>>(
        ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->elem_slot, ctx->ne, P.val);
  }
  CK_LAUNCH();
  return 0;
}

#define FEMCY_DISPATCH(FN, ...)                                                        \
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

extern "C" int femcy_get_dsdx_and_vol(femcy_ctx* ctx) {
  cudaSetDevice(ctx->device);
  return femcy_for_sections(ctx, [&]() -> int {
    if (!ctx->have_elem) return femcy_fail_msg(ctx, "set_element first");
    FEMCY_DISPATCH(launch_dsdx, ctx, true);
  });
}

extern "C" int femcy_assemble_K(femcy_ctx* ctx, int variant) {
  cudaSetDevice(ctx->device);
  if (ctx->sections.empty() && (!ctx->have_elem || !ctx->have_mat)) return femcy_fail_msg(ctx, "set_element and set_material first");
  if (!ctx->P.val) return femcy_fail_msg(ctx, "build_pattern first");
  CK(cudaEventRecord(ctx->evA0, ctx->stream));
  int rc;
  if (!ctx->sections.empty()) {
    // Row f4: K = sum over the sections; one zero-fill, then every section scatter-adds with its own element tables and
    // tangent into the slots of the union pattern
    if (variant != 0 && variant != FEMCY_ASSEMBLY_SCATTER)
      return femcy_fail_msg(ctx, "a mesh of several sections assembles by scatter-add (variant 0 or 1)");
    CK(cudaMemsetAsync(ctx->P.val, 0, (size_t)(ctx->P.nslots * ctx->dm * ctx->dm) * sizeof(double), ctx->stream));
    rc = femcy_for_sections(ctx, [&]() -> int {
      if (!ctx->have_elem || !ctx->have_mat) return femcy_fail_msg(ctx, "set_element and set_material for every section first");
      if (!ctx->elem_slot) return femcy_fail_msg(ctx, "build_pattern first");
      FEMCY_DISPATCH(launch_assemble, ctx, FEMCY_ASSEMBLY_SCATTER, false);
    });
  } else {
    auto run = [&]() -> int { FEMCY_DISPATCH(launch_assemble, ctx, variant, true); };
    rc = run();
  }
  if (rc) return rc;
  CK(cudaEventRecord(ctx->evA1, ctx->stream));
  return 0;
}
