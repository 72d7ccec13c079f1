// Hot path rows a2 + a3: shape-function gradients / volumes and the stiffness assembly.
//
// Reference kernels replaced:
//   System_of_equations.get_dsdx_and_vol        /root/reference/stiffnessMtrx.py:132-150
//   System_of_equations.assemble_stiffnessMtrx  /root/reference/stiffnessMtrx.py:161-186
//   System_of_equations.sparseMatrix_get_j      /root/reference/stiffnessMtrx.py:414-420  (row scan -> precomputed slot)
//
// Assembly kernel families over the same node-block SELL-32 matrix (device code: assembly_kernels.cuh; the variant
// numbers are those of femcy_assemble_K in include/femcy_b200.h; measurements: DESIGN.md section 4 / 4a):
//   scatter (1, 3, 4): thread (or warp, n_en >= 8) per element; K_e accumulated over the Gauss points in registers,
//            then dm*dm fp64 atomic adds (RED.ADD.F64) per node pair into the precomputed slot.
//   gather  (2, 5 = default for single-Gauss-point elements, 9-13): pass 1 writes grad N + vol per element, pass 2 runs
//            one thread per stored block and sums its element list -- no atomics, no zero-fill, bit-reproducible,
//            coalesced 256 B plane stores.
//   rows    (6-8, 16, 17): owner-computes; the rows of a slice accumulate in shared memory from node->element lists.
//   tile    (14, 15): the gather with the slice's element records staged in shared memory.
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

template <int DM, int NEN, int NGP>
static int launch_assemble(femcy_ctx* ctx, int variant) {
  BsellPattern& P = ctx->P;
  constexpr int DM2 = DM * DM;
  if (ctx->ne == 0) return 0;
  bool gather_ok = ctx->ent_list != nullptr;   // NGP > 1: experimental k_assemble_gather_mgp
  // default (measured on B200, cfg 4 / cfg 5, profiles/r1z_quick_ab.jsonl): single-Gauss-point elements use the
  // atomic-free gather in slice-major launch order (2.94 ms vs 3.25 ms for the scatter on 10.1 M C3D4; bit-reproducible);
  // multi-Gauss-point elements keep the atomic scatter (C3D10: 8.8 ms; rows 8.2 ms, gather 9.7 ms -- no clear winner yet)
  if (variant == 0) variant = (NGP == 1 && gather_ok) ? 5 : 1;
  if (variant == 2 && !gather_ok) return femcy_fail_msg(ctx, "gather assembly needs the element lists of build_pattern");
  if (variant == 19) {
    // experimental: software-pipelined symmetric warp scatter for big elements (k_assemble_scatter_pairs); a
    // non-symmetric tangent takes the plain warp scatter
    if constexpr (NEN >= 6) {
      if (tangent_is_symmetric(ctx->tab.C, DM)) {
        CK(cudaMemsetAsync(P.val, 0, (size_t)(P.nslots * DM2) * sizeof(double), ctx->stream));  // K.fill(0), :168
        const bool cubic = tangent_is_cubic(ctx->tab.C, DM);
        const void* kfn = cubic ? (const void*)k_assemble_scatter_pairs<DM, NEN, NGP, true>
                                : (const void*)k_assemble_scatter_pairs<DM, NEN, NGP, false>;
        int nbsm = 0, nsm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbsm, kfn, 128, 0));
        CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
        int64_t blocks = ceil_div64(ctx->ne, 4);                 // persistent warps: the prefetch needs a loop to run ahead in
        if (nbsm >= 1 && blocks > (int64_t)nbsm * nsm) blocks = (int64_t)nbsm * nsm;
        if (cubic)
          k_assemble_scatter_pairs<DM, NEN, NGP, true><<<(int)blocks, 128, 0, ctx->stream>>>(
              ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->elem_slot, ctx->ne, P.val);
        else
          k_assemble_scatter_pairs<DM, NEN, NGP, false><<<(int)blocks, 128, 0, ctx->stream>>>(
              ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->elem_slot, ctx->ne, P.val);
        CK_LAUNCH();
        return 0;
      }
      variant = 1;
    } else {
      return femcy_fail_msg(ctx, "assembly variant 19 (pair scatter) is for elements with 6 or more nodes");
    }
  }
  if (variant == 18) {
    // experimental: first pass through a TMA tensor store (C3D4), second pass = the cubic-tangent gather of variant 10
    if constexpr (NGP == 1 && NEN == 4) {
      if (!gather_ok) return femcy_fail_msg(ctx, "gather assembly needs the element lists of build_pattern");
      if (!ctx->egeo4) {
        if (femcy_alloc(ctx, &ctx->egeo4, ctx->ne * NEN * NGP * 4)) return 1;
      }
      FemcyTmap tm;
      {
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
        CUresult cr = ((EncodeTiled)fn)(&tm.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ctx->egeo4, gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return femcy_fail_msg(ctx, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)cr) + ")");
      }
      k_elem_geometry4t<DM, NEN><<<(int)ceil_div64(ctx->ne, 128), 128, 0, ctx->stream>>>(
          ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, tm, ctx->vol);
      CK_LAUNCH();
      const int KB = 8;
      int kgroups = (P.max_row_blocks + KB - 1) / KB;
      if (tangent_is_cubic(ctx->tab.C, DM))
        k_assemble_gather4<DM, NEN, NGP, true><<<(unsigned)(P.nslice * kgroups), dim3(32, KB), 0, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, kgroups);
      else
        k_assemble_gather4<DM, NEN, NGP, false><<<(unsigned)(P.nslice * kgroups), dim3(32, KB), 0, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, kgroups);
      CK_LAUNCH();
      return 0;
    } else {
      return femcy_fail_msg(ctx, "assembly variant 18 (TMA store) is for 4-node single-Gauss-point elements (C3D4)");
    }
  }
  if (variant == 15) {
    // experimental tile assembly for multi-Gauss-point / large elements: 8-row blocks, one Gauss point staged at a time
    if (!gather_ok) return femcy_fail_msg(ctx, "tile assembly needs the element lists of build_pattern");
    if (femcy_build_tiles(ctx, 3)) return 1;
    if (!ctx->egeo4) {
      if (femcy_alloc(ctx, &ctx->egeo4, ctx->ne * NEN * NGP * 4)) return 1;
    }
    using G = Geo4Cfg<NEN, NGP>;
    k_elem_geometry4s<DM, NEN, NGP><<<(int)ceil_div64(ctx->ne, G::TPB), G::TPB, 0, ctx->stream>>>(
        ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, ctx->egeo4, ctx->vol);
    CK_LAUNCH();
    size_t smem = (size_t)ctx->max_tile * NEN * 2 * 16;
    if (smem > 200 * 1024) return femcy_fail_msg(ctx, "tile assembly: a row block touches too many elements for shared memory");
    unsigned tgrid = (unsigned)(P.nslice * 4);
    if (tangent_is_cubic(ctx->tab.C, DM)) {
      if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_assemble_tile_mgp<DM, NEN, NGP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_assemble_tile_mgp<DM, NEN, NGP, true><<<tgrid, dim3(FEMCY_TILE_RB, FEMCY_TILE_KT), smem, ctx->stream>>>(
          ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_tile, ctx->tile_ptr, ctx->tile_elems, ctx->egeo4, P.val);
    } else {
      if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_assemble_tile_mgp<DM, NEN, NGP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_assemble_tile_mgp<DM, NEN, NGP, false><<<tgrid, dim3(FEMCY_TILE_RB, FEMCY_TILE_KT), smem, ctx->stream>>>(
          ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_tile, ctx->tile_ptr, ctx->tile_elems, ctx->egeo4, P.val);
    }
    CK_LAUNCH();
    return 0;
  }
  if (variant == 22) {
    // experimental: the tile assembly of variant 14 with its records staged by TMA-engine bulk copies on an mbarrier
    if constexpr (NGP == 1) {
      if (!gather_ok) return femcy_fail_msg(ctx, "tile assembly needs the element lists of build_pattern");
      if (femcy_build_tiles(ctx, 5)) return 1;
      if (!ctx->egeo4) {
        if (femcy_alloc(ctx, &ctx->egeo4, ctx->ne * NEN * NGP * 4)) return 1;
      }
      using G = Geo4Cfg<NEN, NGP>;
      k_elem_geometry4s<DM, NEN, NGP><<<(int)ceil_div64(ctx->ne, G::TPB), G::TPB, 0, ctx->stream>>>(
          ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, ctx->egeo4, ctx->vol);
      CK_LAUNCH();
      size_t smem = (size_t)ctx->max_tile * TileBCfg<NEN>::PB;
      if (smem > 200 * 1024) return femcy_fail_msg(ctx, "tile assembly: a slice touches too many elements for shared memory");
      if (tangent_is_cubic(ctx->tab.C, DM)) {
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_assemble_tile_b<DM, NEN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_assemble_tile_b<DM, NEN, true><<<(unsigned)P.nslice, dim3(32, 8), smem, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_tile, ctx->tile_ptr, ctx->tile_elems, ctx->egeo4, P.val);
      } else {
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_assemble_tile_b<DM, NEN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_assemble_tile_b<DM, NEN, false><<<(unsigned)P.nslice, dim3(32, 8), smem, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_tile, ctx->tile_ptr, ctx->tile_elems, ctx->egeo4, P.val);
      }
      CK_LAUNCH();
      return 0;
    } else {
      return femcy_fail_msg(ctx, "assembly variant 22 (tile, bulk loads) is for single-Gauss-point elements");
    }
  }
  if (variant == 14) {
    // experimental "tile" assembly: per-block gather out of shared memory (single-Gauss-point elements)
    if constexpr (NGP == 1) {
      if (!gather_ok) return femcy_fail_msg(ctx, "tile assembly needs the element lists of build_pattern");
      if (femcy_build_tiles(ctx, 5)) return 1;
      if (!ctx->egeo4) {
        if (femcy_alloc(ctx, &ctx->egeo4, ctx->ne * NEN * NGP * 4)) return 1;
      }
      using G = Geo4Cfg<NEN, NGP>;
      k_elem_geometry4s<DM, NEN, NGP><<<(int)ceil_div64(ctx->ne, G::TPB), G::TPB, 0, ctx->stream>>>(
          ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, ctx->egeo4, ctx->vol);
      CK_LAUNCH();
      size_t smem = (size_t)ctx->max_tile * NEN * 2 * 16;
      if (smem > 200 * 1024) return femcy_fail_msg(ctx, "tile assembly: a slice touches too many elements for shared memory");
      const bool cubic = tangent_is_cubic(ctx->tab.C, DM);
      if (cubic) {
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_assemble_tile<DM, NEN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_assemble_tile<DM, NEN, true><<<(unsigned)P.nslice, dim3(32, 8), smem, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_tile, ctx->tile_ptr, ctx->tile_elems, ctx->egeo4, P.val);
      } else {
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_assemble_tile<DM, NEN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_assemble_tile<DM, NEN, false><<<(unsigned)P.nslice, dim3(32, 8), smem, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_tile, ctx->tile_ptr, ctx->tile_elems, ctx->egeo4, P.val);
      }
      CK_LAUNCH();
      return 0;
    } else {
      return femcy_fail_msg(ctx, "assembly variant 14 (tile) is for single-Gauss-point elements");
    }
  }
  if (variant == 23 || variant == 24) {
    // gradient-product gather: 23 = one block per slice, 24 = persistent grid-stride over the slices
    if (!gather_ok) return femcy_fail_msg(ctx, "gather assembly needs the element lists of build_pattern");
    if (!ctx->egeo4) {
      if (femcy_alloc(ctx, &ctx->egeo4, ctx->ne * NEN * NGP * 4)) return 1;
    }
    using G = Geo4Cfg<NEN, NGP>;
    k_elem_geometry4s<DM, NEN, NGP><<<(int)ceil_div64(ctx->ne, G::TPB), G::TPB, 0, ctx->stream>>>(
        ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, ctx->egeo4, ctx->vol);
    CK_LAUNCH();
    unsigned g = (unsigned)P.nslice;
    if (variant == 24 && g > 148u * 8u) g = 148u * 8u;
    if (tangent_is_cubic(ctx->tab.C, DM))
      k_assemble_gather_p<DM, NEN, NGP, true><<<g, dim3(32, 8), 0, ctx->stream>>>(
          ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, P.nslice);
    else
      k_assemble_gather_p<DM, NEN, NGP, false><<<g, dim3(32, 8), 0, ctx->stream>>>(
          ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, P.nslice);
    CK_LAUNCH();
    return 0;
  }
  if ((variant >= 6 && variant <= 10) || variant == 12 || variant == 13 || variant == 16 || variant == 17 || variant == 20) {
    // experimental atomic-free variants over the node-sector records rec[e][a][gp] = (grad N_a, vol_gp):
    //   6 = rows assembly (owner-computes in shared memory), plain loop, thread-per-element pass 1 (as measured r1z)
    //   7 / 8 = rows assembly with L2 / L1 software prefetch, pass 1 with coalesced (staged) record stores
    //   9 = per-block gather over the same records, slice-major, staged pass 1;  10 = 9 + cubic-form tangent fast path
    if (!ctx->egeo4) {
      if (femcy_alloc(ctx, &ctx->egeo4, ctx->ne * NEN * NGP * 4)) return 1;
    }
    if (variant == 6) {
      int grid = (int)ceil_div64(ctx->ne, 128);
      k_elem_geometry4<DM, NEN, NGP><<<grid, 128, 0, ctx->stream>>>(ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems,
                                                                    ctx->ne, ctx->egeo4, ctx->vol);
    } else {
      using G = Geo4Cfg<NEN, NGP>;
      int grid = (int)ceil_div64(ctx->ne, G::TPB);
      k_elem_geometry4s<DM, NEN, NGP><<<grid, G::TPB, 0, ctx->stream>>>(ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF],
                                                                        ctx->elems, ctx->ne, ctx->egeo4, ctx->vol);
    }
    CK_LAUNCH();
    if (variant == 9 || variant == 10 || variant == 12 || variant == 13 || variant == 20) {
      if (!gather_ok) return femcy_fail_msg(ctx, "gather assembly needs the element lists of build_pattern");
      const int KB = 8;
      int kgroups = (P.max_row_blocks + KB - 1) / KB;
      // 10 = 9 with the cubic-form tangent fast path; a tangent of another form silently takes the general kernel;
      // 12 = 10 compiled for 6 blocks/SM (<= 42 registers: 75 % instead of 62 % occupancy)
      // 20 = 10 with one 256-bit load per record (a general tangent: the general kernel, also with 256-bit loads)
      if (variant == 20 && tangent_is_cubic(ctx->tab.C, DM))
        k_assemble_gather4<DM, NEN, NGP, true, 0, true><<<(unsigned)(P.nslice * kgroups), dim3(32, KB), 0, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, kgroups);
      else if (variant == 20)
        k_assemble_gather4<DM, NEN, NGP, false, 0, true><<<(unsigned)(P.nslice * kgroups), dim3(32, KB), 0, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, kgroups);
      else if (variant == 12 && tangent_is_cubic(ctx->tab.C, DM))
        k_assemble_gather4<DM, NEN, NGP, true, 6><<<(unsigned)(P.nslice * kgroups), dim3(32, KB), 0, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, kgroups);
      else if (variant == 13 && tangent_is_cubic(ctx->tab.C, DM))   // 13 = 10 with the register cap lifted (compiler trades occupancy for ILP)
        k_assemble_gather4<DM, NEN, NGP, true, 1><<<(unsigned)(P.nslice * kgroups), dim3(32, KB), 0, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, kgroups);
      else if (variant >= 10 && tangent_is_cubic(ctx->tab.C, DM))
        k_assemble_gather4<DM, NEN, NGP, true><<<(unsigned)(P.nslice * kgroups), dim3(32, KB), 0, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, kgroups);
      else
        k_assemble_gather4<DM, NEN, NGP, false><<<(unsigned)(P.nslice * kgroups), dim3(32, KB), 0, ctx->stream>>>(
            ctx->tab, P.slice_ptr, ctx->slot_ent_beg, ctx->slot_ent_end, ctx->ent_list, ctx->egeo4, P.val, kgroups);
      CK_LAUNCH();
      return 0;
    }
    if (femcy_build_incidence(ctx)) return 1;
    using Cfg = RowsCfg<NEN>;
    size_t smem = (size_t)P.max_row_blocks * DM2 * Cfg::PITCH * sizeof(double);
    if (smem > 200 * 1024) return femcy_fail_msg(ctx, "rows assembly: a row has too many blocks for the shared-memory accumulator");
    unsigned rgrid = (unsigned)(P.nslice * (32 / Cfg::R));
    if (variant == 16 || variant == 17) {
      // 16 / 17 = rows with register double-buffering (17: + cubic-form tangent fast path); single-Gauss-point elements
      if constexpr (NGP == 1) {
        const bool cubic = (variant == 17) && tangent_is_cubic(ctx->tab.C, DM);
        if (cubic) {
          if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_assemble_rows<DM, NEN, NGP, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_assemble_rows<DM, NEN, NGP, 3, true><<<rgrid, Cfg::NW * 32, smem, ctx->stream>>>(
              ctx->tab, P.slice_ptr, P.nn_own, ctx->inc_ptr, ctx->inc_list, ctx->elem_slot, ctx->egeo4, P.val, P.rowof);
        } else {
          if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_assemble_rows<DM, NEN, NGP, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_assemble_rows<DM, NEN, NGP, 3, false><<<rgrid, Cfg::NW * 32, smem, ctx->stream>>>(
              ctx->tab, P.slice_ptr, P.nn_own, ctx->inc_ptr, ctx->inc_list, ctx->elem_slot, ctx->egeo4, P.val, P.rowof);
        }
        CK_LAUNCH();
        return 0;
      } else {
        return femcy_fail_msg(ctx, "assembly variants 16 / 17 are for single-Gauss-point elements");
      }
    }
#define FEMCY_ROWS_LAUNCH(PF)                                                                                         \
    do {                                                                                                              \
      if (smem > 48 * 1024)                                                                                           \
        CK(cudaFuncSetAttribute(k_assemble_rows<DM, NEN, NGP, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_assemble_rows<DM, NEN, NGP, PF><<<rgrid, Cfg::NW * 32, smem, ctx->stream>>>(                                  \
          ctx->tab, P.slice_ptr, P.nn_own, ctx->inc_ptr, ctx->inc_list, ctx->elem_slot, ctx->egeo4, P.val, P.rowof);   \
    } while (0)
    if (variant == 6) FEMCY_ROWS_LAUNCH(0);
    else if (variant == 7) FEMCY_ROWS_LAUNCH(1);
    else FEMCY_ROWS_LAUNCH(2);
#undef FEMCY_ROWS_LAUNCH
    CK_LAUNCH();
    return 0;
  }
  if (variant == 5) {
    if constexpr (NGP != 1) return femcy_fail_msg(ctx, "assembly variant 5 (slice-major gather) is for single-Gauss-point elements");
    if (!gather_ok) return femcy_fail_msg(ctx, "gather assembly needs the element lists of build_pattern");
  }
  const bool staged_pass1 = (variant == 11);      // 11 = 5 with the record stores of pass 1 staged through shared memory
  const bool bulk_pass1 = (variant == 21);        // 21 = 5 with the records leaving shared memory as one bulk copy per block
  if (variant == 11 || variant == 21) {
    if constexpr (NGP != 1) return femcy_fail_msg(ctx, "assembly variants 11 / 21 are for single-Gauss-point elements");
    if (!gather_ok) return femcy_fail_msg(ctx, "gather assembly needs the element lists of build_pattern");
    variant = 5;
  }
  if (variant < 0 || variant > 5) return femcy_fail_msg(ctx, "unknown assembly variant");
  if (variant == 1 || variant == 3 || variant == 4) {
    CK(cudaMemsetAsync(P.val, 0, (size_t)(P.nslots * DM2) * sizeof(double), ctx->stream));  // K.fill(0), :168
    int grid = (int)ceil_div64(ctx->ne, 128);
    if constexpr (NEN >= 8) {
      // one warp per element (C3D10, CPS8/CPE8)
      int64_t blocks = ceil_div64(ctx->ne, 4);
      if (blocks > 148 * 64) blocks = 148 * 64;
      // variant 4 (experimental): every warp owns a contiguous element range instead of a grid stride
      int64_t chunk = (variant == 4) ? ceil_div64(ctx->ne, blocks * 4) : 0;
      k_assemble_scatter_warp<DM, NEN, NGP><<<(int)blocks, 128, 0, ctx->stream>>>(
          ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->elem_slot, ctx->ne, P.val, chunk);
    } else if (variant == 3)   // experiment: cap registers for 2x occupancy
      k_assemble_scatter<DM, NEN, NGP, 8><<<grid, 128, 0, ctx->stream>>>(ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF],
                                                                        ctx->elems, ctx->elem_slot, ctx->ne, P.val);
    else
      k_assemble_scatter<DM, NEN, NGP, 1><<<grid, 128, 0, ctx->stream>>>(ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF],
                                                                        ctx->elems, ctx->elem_slot, ctx->ne, P.val);
    CK_LAUNCH();
  } else {
    if constexpr (NGP > 1) {
      // experimental: dsdx/vol pre-pass (the reference's own two-pass structure) + atomic-free gather
      if (launch_dsdx<DM, NEN, NGP>(ctx, true)) return 1;
      const int KB = 8;
      dim3 blk(32, KB);
      dim3 grd((unsigned)P.nslice, (unsigned)((P.max_row_blocks + KB - 1) / KB));
      k_assemble_gather_mgp<DM, NEN, NGP><<<grd, blk, 0, ctx->stream>>>(ctx->tab, P.slice_ptr, P.nslice, ctx->slot_ent_beg,
                                                                       ctx->slot_ent_end, ctx->ent_list, ctx->dsdx, ctx->vol, P.val);
      CK_LAUNCH();
    }
    if constexpr (NGP == 1) {
      constexpr int REC = GeoRec<DM, NEN>::N;
      if (!ctx->egeo) {
        if (femcy_alloc(ctx, &ctx->egeo, ctx->ne * REC)) return 1;
      }
      if (staged_pass1) {
        k_elem_geometry_s<DM, NEN><<<(int)ceil_div64(ctx->ne, 128), 128, 0, ctx->stream>>>(
            ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, ctx->egeo, ctx->vol);
      } else if (bulk_pass1) {
        k_elem_geometry_b<DM, NEN><<<(int)ceil_div64(ctx->ne, 128), 128, 0, ctx->stream>>>(
            ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems, ctx->ne, ctx->egeo, ctx->vol);
      } else {
        int grid = (int)ceil_div64(ctx->ne, 256);
        k_elem_geometry<DM, NEN><<<grid, 256, 0, ctx->stream>>>(ctx->tab, ctx->nodes, ctx->vec[FEMCY_VEC_DOF], ctx->elems,
                                                               ctx->ne, ctx->egeo, ctx->vol);
      }
      CK_LAUNCH();
      const int KB = 8;
      dim3 blk(32, KB);
      int kgroups = (P.max_row_blocks + KB - 1) / KB;
      dim3 grd((unsigned)P.nslice, (unsigned)kgroups);
      if (variant == 5) grd = dim3((unsigned)(P.nslice * kgroups), 1);      // slice-major launch order
      k_assemble_gather<DM, NEN><<<grd, blk, 0, ctx->stream>>>(ctx->tab, P.slice_ptr, P.nslice, ctx->slot_ent_beg,
                                                              ctx->slot_ent_end, ctx->ent_list, ctx->egeo, P.val,
                                                              variant == 5 ? kgroups : 0);
      CK_LAUNCH();
    }
  }
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
  if (!ctx->have_elem) return femcy_fail_msg(ctx, "set_element first");
  FEMCY_DISPATCH(launch_dsdx, ctx, true);
}

extern "C" int femcy_assemble_K(femcy_ctx* ctx, int variant) {
  cudaSetDevice(ctx->device);
  if (!ctx->have_elem || !ctx->have_mat) return femcy_fail_msg(ctx, "set_element and set_material first");
  if (!ctx->P.val) return femcy_fail_msg(ctx, "build_pattern first");
  CK(cudaEventRecord(ctx->evA0, ctx->stream));
  int rc;
  {
    auto run = [&]() -> int { FEMCY_DISPATCH(launch_assemble, ctx, variant); };
    rc = run();
  }
  if (rc) return rc;
  CK(cudaEventRecord(ctx->evA1, ctx->stream));
  return 0;
}
