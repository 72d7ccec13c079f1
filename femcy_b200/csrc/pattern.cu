// Device build of the node-block SELL-32 sparsity pattern from element connectivity (row a1).
//
// Replaces the reference's Python-loop topology code:
//   Body.get_nodeEles / get_coElement_nodes            /root/reference/body.py:165-194
//   sparseIJ + rows/cols lists                          /root/reference/stiffnessMtrx.py:78-107
// Design: every (element, a, b) pair emits the 64-bit key  row_node*nn + col_node ; one radix
// sort (CUB) groups equal keys; run heads are the non-zero blocks (columns come out sorted,
// unlike the reference's Python-set order -- only the summation order differs).
// The same sorted order gives (1) elem_slot: slot of each element-local block, used by the
// atomic scatter assembly and (2) ent_list/slot_ent_*: element lists per block, used by the
// atomic-free gather assembly.
#include <cub/cub.cuh>
#include <stdlib.h>

#include <vector>

#include "ctx.cuh"
#include "pattern_kernels.cuh"

int femcy_pattern_free(femcy_ctx* ctx) {
  BsellPattern& P = ctx->P;
  femcy_free(&P.slice_ptr); femcy_free(&P.blkptr); femcy_free(&P.colidx); femcy_free(&P.diag_slot); femcy_free(&P.val);
  femcy_free(&P.rowof); femcy_free(&P.rowpos);
  femcy_free(&ctx->elem_slot); femcy_free(&ctx->ent_list); femcy_free(&ctx->slot_ent_beg); femcy_free(&ctx->slot_ent_end);
  for (int s = 0; s < (int)ctx->sections.size(); ++s) {          // parked sections; the selected one's copy is stale
    if (s != ctx->cur_section) femcy_free(&ctx->sections[s].elem_slot);
    ctx->sections[s].elem_slot = nullptr;
  }
  femcy_free(&ctx->egeo4); ctx->egeo4_tmap_for = nullptr;
  femcy_free(&ctx->U.slice_ptr); femcy_free(&ctx->U.colidx); femcy_free(&ctx->U.src); femcy_free(&ctx->U.val); ctx->U = SymPattern();
  P = BsellPattern();
  ctx->n_ent = 0;
  femcy_drop_graph(ctx);
  return 0;
}

static inline int gridp(int64_t n) {
  int64_t g = ceil_div64(n, 256);
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

// Shared tail: keys/ids (unsorted, device) -> pattern + entry_slot[n_total].
static int build_from_keys(femcy_ctx* ctx, uint64_t* keys, uint32_t* ids, int64_t n_total, int64_t nrows, int64_t ncols,
                           int dm, int32_t* entry_slot, bool keep_lists) {
  cudaStream_t st = ctx->stream;
  BsellPattern& P = ctx->P;
  P.dm = dm; P.nn = ncols; P.nn_own = nrows;
  uint64_t invalid = (uint64_t)nrows * (uint64_t)ncols;
  int end_bit = 1;
  while (end_bit < 64 && (invalid >> end_bit) != 0) ++end_bit;

  uint64_t* keys2 = nullptr; uint32_t* ids2 = nullptr;
  if (femcy_alloc(ctx, &keys2, n_total) || femcy_alloc(ctx, &ids2, n_total)) return 1;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ids, ids2, n_total, 0, end_bit, st);
  void* tmp = nullptr;
  CK(cudaMalloc(&tmp, tmp_bytes + 16));
  CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ids, ids2, n_total, 0, end_bit, st));
  ctx->launches += 8;
  cudaFree(tmp);
  // keys2/ids2 sorted.  Valid prefix length:
  int64_t* d_nvalid = nullptr;
  if (femcy_alloc(ctx, &d_nvalid, 1)) return 1;
  CK(cudaMemsetAsync(d_nvalid, 0, sizeof(int64_t), st));
  k_count_valid<<<gridp(n_total), 256, 0, st>>>(keys2, n_total, invalid, d_nvalid);
  CK_LAUNCH();
  int64_t n_ent = 0;
  CK(cudaMemcpyAsync(&n_ent, d_nvalid, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  femcy_free(&d_nvalid);
  if (n_ent >= ((int64_t)1 << 31)) return femcy_fail_msg(ctx, "too many element-block entries for int32 offsets");
  ctx->n_ent = n_ent;

  int32_t *head = nullptr, *blk_of = nullptr;
  if (femcy_alloc(ctx, &head, n_ent) || femcy_alloc(ctx, &blk_of, n_ent)) return 1;
  k_heads<<<gridp(n_ent), 256, 0, st>>>(keys2, n_ent, head);
  CK_LAUNCH();
  tmp_bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, head, blk_of, n_ent, st);
  CK(cudaMalloc(&tmp, tmp_bytes + 16));
  CK(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, head, blk_of, n_ent, st));
  ctx->launches += 2;
  int32_t nnzb32 = 0;
  if (n_ent > 0) CK(cudaMemcpyAsync(&nnzb32, blk_of + (n_ent - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(tmp);
  int64_t nnzb = nnzb32;
  P.nnzb = nnzb;

  int32_t *brow = nullptr, *bcol = nullptr, *bfirst = nullptr, *bslot = nullptr;
  if (femcy_alloc(ctx, &brow, nnzb) || femcy_alloc(ctx, &bcol, nnzb) || femcy_alloc(ctx, &bfirst, nnzb) ||
      femcy_alloc(ctx, &bslot, nnzb))
    return 1;
  k_block_info<<<gridp(n_ent), 256, 0, st>>>(keys2, head, blk_of, n_ent, ncols, brow, bcol, bfirst);
  CK_LAUNCH();
  femcy_free(&head);
  femcy_free(&keys2);

  if (femcy_alloc(ctx, &P.blkptr, nrows + 1)) return 1;
  k_blkptr<<<gridp(nrows + 1), 256, 0, st>>>(brow, nnzb, nrows, P.blkptr);
  CK_LAUNCH();

  P.nslice = ceil_div64(nrows, FEMCY_SLICE);
  // SELL-32-sigma row order (option sell_sigma): -1 = automatic -- sigma = 1024 when natural-order slices would carry more than
  // 15 % padding (quadratic elements: corner rows of 65 blocks next to mid-edge rows of 14-42), else natural order.
  // Measured on cfg 5 (profiles/r2q, r2i): gather assembly 3.95 -> 2.92 ms, PCG iteration 0.698 -> 0.620 ms; no effect on
  // meshes with uniform rows (cfg 4: 0.5 % padding).
  {
    int sigma = ctx->opt.sell_sigma;
    if (sigma == -1) {
      int32_t* nat = nullptr; int32_t* nmax = nullptr;
      if (femcy_alloc(ctx, &nat, P.nslice + 1) || femcy_alloc(ctx, &nmax, 1)) return 1;
      CK(cudaMemsetAsync(nmax, 0, sizeof(int32_t), st));
      k_slice_width<<<gridp(P.nslice), 256, 0, st>>>(P.blkptr, nrows, P.nslice, nat, nmax, nullptr);
      CK_LAUNCH();
      std::vector<int32_t> h(P.nslice > 0 ? P.nslice : 1);
      if (P.nslice > 0) CK(cudaMemcpy(h.data(), nat, (size_t)P.nslice * sizeof(int32_t), cudaMemcpyDeviceToHost));
      int64_t slots_nat = 0;
      for (int64_t q = 0; q < P.nslice; ++q) slots_nat += h[q];
      femcy_free(&nat); femcy_free(&nmax);
      sigma = (nrows >= 4096 && (double)slots_nat > 1.15 * (double)nnzb) ? 1024 : 0;
    }
    if (sigma < 0 || (sigma % FEMCY_SLICE) != 0) return femcy_fail_msg(ctx, "sell_sigma must be -1 (automatic) or a multiple of 32");
    P.sigma = sigma;
    if (sigma > 0 && nrows > 0) {
      if ((uint64_t)(nrows / sigma) >= ((uint64_t)1 << 24)) return femcy_fail_msg(ctx, "sell_sigma too small for this many rows");
      uint32_t *k1 = nullptr, *k2 = nullptr; int32_t* r1 = nullptr;
      if (femcy_alloc(ctx, &k1, nrows) || femcy_alloc(ctx, &k2, nrows) || femcy_alloc(ctx, &r1, nrows) ||
          femcy_alloc(ctx, &P.rowof, P.nslice * FEMCY_SLICE) || femcy_alloc(ctx, &P.rowpos, nrows))
        return 1;
      k_fill_i32<<<gridp(P.nslice * FEMCY_SLICE), 256, 0, st>>>(P.rowof, -1, P.nslice * FEMCY_SLICE);
      CK_LAUNCH();
      k_sigma_keys<<<gridp(nrows), 256, 0, st>>>(P.blkptr, nrows, sigma, k1, r1);
      CK_LAUNCH();
      int eb = 9;
      while (eb < 32 && (((uint64_t)(nrows / sigma) << 8) >> eb) != 0) ++eb;
      size_t sb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, sb, k1, k2, r1, P.rowof, nrows, 0, eb, st);
      void* stmp = nullptr;
      CK(cudaMalloc(&stmp, sb + 16));
      CK(cub::DeviceRadixSort::SortPairs(stmp, sb, k1, k2, r1, P.rowof, nrows, 0, eb, st));
      ctx->launches += 4;
      k_rowpos<<<gridp(nrows), 256, 0, st>>>(P.rowof, nrows, P.rowpos);
      CK_LAUNCH();
      CK(cudaStreamSynchronize(st));
      cudaFree(stmp);
      femcy_free(&k1); femcy_free(&k2); femcy_free(&r1);
    }
  }
  int32_t* sps = nullptr; int32_t* d_maxw = nullptr;
  if (femcy_alloc(ctx, &sps, P.nslice + 1) || femcy_alloc(ctx, &d_maxw, 1)) return 1;
  CK(cudaMemsetAsync(d_maxw, 0, sizeof(int32_t), st));
  CK(cudaMemsetAsync(sps, 0, (size_t)(P.nslice + 1) * sizeof(int32_t), st));
  k_slice_width<<<gridp(P.nslice), 256, 0, st>>>(P.blkptr, nrows, P.nslice, sps, d_maxw, P.rowof);
  CK_LAUNCH();
  if (femcy_alloc(ctx, &P.slice_ptr, P.nslice + 1)) return 1;
  tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, sps, P.slice_ptr, P.nslice + 1, st);
  CK(cudaMalloc(&tmp, tmp_bytes + 16));
  CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, sps, P.slice_ptr, P.nslice + 1, st));
  ctx->launches += 2;
  int32_t nslots32 = 0, maxw = 0;
  CK(cudaMemcpyAsync(&nslots32, P.slice_ptr + P.nslice, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&maxw, d_maxw, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(tmp);
  femcy_free(&sps); femcy_free(&d_maxw);
  P.nslots = nslots32;
  P.max_row_blocks = maxw;

  if (femcy_alloc(ctx, &P.colidx, P.nslots) || femcy_alloc(ctx, &P.diag_slot, nrows) ||
      femcy_alloc(ctx, &ctx->slot_ent_beg, P.nslots) || femcy_alloc(ctx, &ctx->slot_ent_end, P.nslots))
    return 1;
  k_fill_i32<<<gridp(P.nslots), 256, 0, st>>>(P.colidx, -1, P.nslots);
  CK_LAUNCH();
  k_fill_i32<<<gridp(nrows), 256, 0, st>>>(P.diag_slot, -1, nrows);
  CK_LAUNCH();
  CK(cudaMemsetAsync(ctx->slot_ent_beg, 0, (size_t)P.nslots * sizeof(int32_t), st));
  CK(cudaMemsetAsync(ctx->slot_ent_end, 0, (size_t)P.nslots * sizeof(int32_t), st));
  k_block_slots<<<gridp(nnzb), 256, 0, st>>>(brow, bcol, bfirst, P.blkptr, P.slice_ptr, nnzb, n_ent, P.colidx,
                                             P.diag_slot, bslot, ctx->slot_ent_beg, ctx->slot_ent_end, P.rowpos);
  CK_LAUNCH();
  k_fill_i32<<<gridp(n_total), 256, 0, st>>>(entry_slot, -1, n_total);
  CK_LAUNCH();
  k_entry_slots<<<gridp(n_ent), 256, 0, st>>>(ids2, blk_of, bslot, n_ent, entry_slot);
  CK_LAUNCH();

  int64_t dm2 = (int64_t)dm * dm;
  if (femcy_alloc(ctx, &P.val, P.nslots * dm2)) return 1;
  CK(cudaMemsetAsync(P.val, 0, (size_t)(P.nslots * dm2) * sizeof(double), st));
  CK(cudaStreamSynchronize(st));
  femcy_free(&brow); femcy_free(&bcol); femcy_free(&bfirst); femcy_free(&bslot); femcy_free(&blk_of);
  if (keep_lists) {
    // ids2[0..n_ent) is the entry list ordered by block; keep it (shrunk) for the gather assembly
    if (femcy_alloc(ctx, &ctx->ent_list, n_ent)) return 1;
    CK(cudaMemcpyAsync(ctx->ent_list, ids2, (size_t)n_ent * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
  } else {
    femcy_free(&ctx->slot_ent_beg); femcy_free(&ctx->slot_ent_end);
  }
  femcy_free(&ids2);
  return 0;
}

// Row f4: the pattern of a mesh of several sections is the union of their element couplings -- the keys of all sections
// go through ONE sort; every section receives the slots of its own element-local blocks (the scatter assembly's input).
// The per-block element lists of the gather assembly are not kept: entries of different element kinds do not share a
// record format, so a multi-section mesh assembles by scatter-add.
static int build_pattern_sections(femcy_ctx* ctx, int64_t* nnz_out) {
  femcy_section_park(ctx);
  const int nsec = (int)ctx->sections.size();
  std::vector<int64_t> off(nsec + 1, 0);
  for (int s = 0; s < nsec; ++s) {
    const FemcySection& S = ctx->sections[s];
    if (!S.elems) return femcy_fail_msg(ctx, "a section has no elements array");
    off[s + 1] = off[s] + S.ne * (int64_t)S.n_en * S.n_en;
  }
  const int64_t total = off[nsec];
  if (total >= ((int64_t)1 << 32)) return femcy_fail_msg(ctx, "sum of ne*n_en^2 over the sections exceeds uint32 entry ids");
  uint64_t* keys = nullptr; uint32_t* ids = nullptr; int32_t* entry_slot = nullptr;
  auto drop = [&]() { femcy_free(&keys); femcy_free(&ids); femcy_free(&entry_slot); };
  if (femcy_alloc(ctx, &keys, total) || femcy_alloc(ctx, &ids, total) || femcy_alloc(ctx, &entry_slot, total)) { drop(); return 1; }
  for (int s = 0; s < nsec; ++s) {
    const FemcySection& S = ctx->sections[s];
    const int64_t cnt = off[s + 1] - off[s];
    if (cnt == 0) continue;
    k_elem_keys<<<gridp(cnt), 256, 0, ctx->stream>>>(S.elems, S.ne, S.n_en, ctx->nn, ctx->nn_own, keys + off[s], ids + off[s],
                                                     (uint32_t)off[s]);
    ctx->launches++;
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) { drop(); return femcy_fail(ctx, "kernel launch", le, __FILE__, __LINE__); }
  }
  int rc = build_from_keys(ctx, keys, ids, total, ctx->nn_own, ctx->nn, ctx->dm, entry_slot, false);
  femcy_free(&keys); femcy_free(&ids);
  if (rc) { drop(); return rc; }
  for (int s = 0; s < nsec; ++s) {
    FemcySection& S = ctx->sections[s];
    const int64_t cnt = off[s + 1] - off[s];
    S.elem_slot = nullptr;
    if (femcy_alloc(ctx, &S.elem_slot, cnt)) { drop(); return 1; }
    if (cnt > 0) {
      cudaError_t ce = cudaMemcpyAsync(S.elem_slot, entry_slot + off[s], (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream);
      if (ce != cudaSuccess) { drop(); return femcy_fail(ctx, "copy of a section's slots", ce, __FILE__, __LINE__); }
    }
  }
  {
    cudaError_t se = cudaStreamSynchronize(ctx->stream);
    drop();
    if (se != cudaSuccess) return femcy_fail(ctx, "cudaStreamSynchronize", se, __FILE__, __LINE__);
  }
  femcy_section_load(ctx, ctx->cur_section);      // the selected section's new elem_slot -> ctx field
  if (nnz_out) *nnz_out = ctx->P.nnzb * ctx->dm * ctx->dm;
  return 0;
}

extern "C" int femcy_build_pattern(femcy_ctx* ctx, int64_t* nnz_out) {
  cudaSetDevice(ctx->device);
  if (ctx->dm == 0 || !ctx->elems) return femcy_fail_msg(ctx, "set_mesh first");
  femcy_pattern_free(ctx);
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  if (!ctx->sections.empty()) {
    if (build_pattern_sections(ctx, nnz_out)) return 1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[2] = ms;
    return 0;
  }
  int64_t Pn = (int64_t)ctx->n_en * ctx->n_en;
  int64_t total = ctx->ne * Pn;
  if (total >= ((int64_t)1 << 32)) return femcy_fail_msg(ctx, "ne*n_en^2 exceeds uint32 entry ids");
  uint64_t* keys = nullptr; uint32_t* ids = nullptr;
  if (femcy_alloc(ctx, &keys, total) || femcy_alloc(ctx, &ids, total)) return 1;
  k_elem_keys<<<gridp(total), 256, 0, ctx->stream>>>(ctx->elems, ctx->ne, ctx->n_en, ctx->nn, ctx->nn_own, keys, ids);
  CK_LAUNCH();
  if (femcy_alloc(ctx, &ctx->elem_slot, total)) return 1;
  int rc = build_from_keys(ctx, keys, ids, total, ctx->nn_own, ctx->nn, ctx->dm, ctx->elem_slot, true);
  femcy_free(&keys); femcy_free(&ids);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->last_ms[2] = ms;
  if (nnz_out) *nnz_out = ctx->P.nnzb * ctx->dm * ctx->dm;
  return 0;
}

// Upper-half pattern for the PCG SpMV (SymPattern, kernel_types.cuh; option cg_sym): built once per pattern.
int femcy_build_sym_pattern(femcy_ctx* ctx) {
  if (ctx->U.slice_ptr) return 0;
  BsellPattern& P = ctx->P;
  if (!P.slice_ptr || P.nslice <= 0) return femcy_fail_msg(ctx, "cg_sym: no pattern");
  cudaStream_t st = ctx->stream;
  SymPattern& U = ctx->U;
  int32_t *kstart = nullptr, *uslots = nullptr;
  if (femcy_alloc(ctx, &kstart, P.nslice * FEMCY_SLICE) || femcy_alloc(ctx, &uslots, P.nslice + 1) ||
      femcy_alloc(ctx, &U.slice_ptr, P.nslice + 1))
    return 1;
  CK(cudaMemsetAsync(uslots, 0, (size_t)(P.nslice + 1) * sizeof(int32_t), st));
  const int wgrid = (int)(ceil_div64(P.nslice, 8) > 148 * 16 ? 148 * 16 : ceil_div64(P.nslice, 8));
  k_sym_rows<<<wgrid, 256, 0, st>>>(P.slice_ptr, P.colidx, P.nslice, kstart, uslots, P.rowof);
  CK_LAUNCH();
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, uslots, U.slice_ptr, (int)(P.nslice + 1), st);
  void* tmp = nullptr;
  CK(cudaMalloc(&tmp, tb + 16));
  CK(cub::DeviceScan::ExclusiveSum(tmp, tb, uslots, U.slice_ptr, (int)(P.nslice + 1), st));
  ctx->launches += 2;
  int32_t total = 0;
  CK(cudaMemcpyAsync(&total, U.slice_ptr + P.nslice, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(tmp);
  U.nslots = total;
  if (femcy_alloc(ctx, &U.colidx, U.nslots) || femcy_alloc(ctx, &U.src, U.nslots) ||
      femcy_alloc(ctx, &U.val, U.nslots * P.dm * P.dm))
    return 1;
  k_sym_fill<<<wgrid, 256, 0, st>>>(P.slice_ptr, P.colidx, kstart, U.slice_ptr, P.nslice, U.colidx, U.src);
  CK_LAUNCH();
  CK(cudaStreamSynchronize(st));
  femcy_free(&kstart); femcy_free(&uslots);
  return 0;
}

// values of the upper-half copy from the current full matrix (start of every cg_sym solve)
int femcy_sym_extract(femcy_ctx* ctx) {
  SymPattern& U = ctx->U;
  BsellPattern& P = ctx->P;
  if (U.nslots == 0) return 0;
  const int g = gridp(U.nslots);
  switch (P.dm) {
    case 1: k_sym_extract<1><<<g, 256, 0, ctx->stream>>>(U.src, U.nslots, P.val, U.val); break;
    case 2: k_sym_extract<2><<<g, 256, 0, ctx->stream>>>(U.src, U.nslots, P.val, U.val); break;
    default: k_sym_extract<3><<<g, 256, 0, ctx->stream>>>(U.src, U.nslots, P.val, U.val); break;
  }
  CK_LAUNCH();
  return 0;
}

extern "C" int femcy_pattern_stats(femcy_ctx* ctx, int64_t* out4) {
  out4[0] = ctx->P.nnzb; out4[1] = ctx->P.nslots; out4[2] = ctx->P.nslice; out4[3] = ctx->P.max_row_blocks;
  return 0;
}

// ---- scalar CSR view (kernels: pattern_kernels.cuh) ----
extern "C" int femcy_get_csr_pattern(femcy_ctx* ctx, int32_t* rowptr, int32_t* colidx) {
  cudaSetDevice(ctx->device);
  BsellPattern& P = ctx->P;
  if (!P.blkptr) return femcy_fail_msg(ctx, "build_pattern first");
  int64_t N = P.nn_own * P.dm, nnz = P.nnzb * P.dm * P.dm;
  int32_t *d_rp = nullptr, *d_ci = nullptr;
  if (femcy_alloc(ctx, &d_rp, N + 1) || femcy_alloc(ctx, &d_ci, nnz)) return 1;
  k_csr_rowptr<<<gridp(N + 1), 256, 0, ctx->stream>>>(P.blkptr, P.nn_own, P.dm, d_rp);
  CK_LAUNCH();
  k_csr_xfer<<<gridp(P.nn_own), 256, 0, ctx->stream>>>(P, d_ci, nullptr, 0);
  CK_LAUNCH();
  CK(cudaMemcpyAsync(rowptr, d_rp, (size_t)(N + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(colidx, d_ci, (size_t)nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  femcy_free(&d_rp); femcy_free(&d_ci);
  return 0;
}
extern "C" int femcy_get_K_csr_values(femcy_ctx* ctx, double* vals) {
  cudaSetDevice(ctx->device);
  BsellPattern& P = ctx->P;
  if (!P.blkptr) return femcy_fail_msg(ctx, "build_pattern first");
  int64_t nnz = P.nnzb * P.dm * P.dm;
  double* d_v = nullptr;
  if (femcy_alloc(ctx, &d_v, nnz)) return 1;
  k_csr_xfer<<<gridp(P.nn_own), 256, 0, ctx->stream>>>(P, nullptr, d_v, 1);
  CK_LAUNCH();
  CK(cudaMemcpyAsync(vals, d_v, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  femcy_free(&d_v);
  return 0;
}
extern "C" int femcy_set_K_csr_values(femcy_ctx* ctx, const double* vals) {
  cudaSetDevice(ctx->device);
  BsellPattern& P = ctx->P;
  if (!P.blkptr) return femcy_fail_msg(ctx, "build_pattern first");
  int64_t nnz = P.nnzb * P.dm * P.dm;
  double* d_v = nullptr;
  if (femcy_alloc(ctx, &d_v, nnz)) return 1;
  CK(cudaMemcpyAsync(d_v, vals, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_csr_xfer<<<gridp(P.nn_own), 256, 0, ctx->stream>>>(P, nullptr, d_v, 2);
  CK_LAUNCH();
  CK(cudaStreamSynchronize(ctx->stream));
  femcy_free(&d_v);
  return 0;
}

// ---- drop-in constructor path: the reference's ELL arrays -> scalar SELL-32 ---------------------
__global__ void k_ell_values(const double* __restrict__ spm, const int32_t* __restrict__ entry_slot, int64_t total,
                             double* __restrict__ val) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int32_t s = entry_slot[t];
    if (s >= 0) val[s] = spm[t];  // dm == 1: slot index == value index
  }
}

extern "C" int femcy_cg_from_ell(femcy_ctx* ctx, int64_t N, int W, const double* spm, const int32_t* sparseIJ) {
  cudaSetDevice(ctx->device);
  if (N * (int64_t)W >= ((int64_t)1 << 32)) return femcy_fail_msg(ctx, "ELL too large");
  // a "mesh" of N one-dof nodes and no elements
  int rc = femcy_set_mesh(ctx, 1, N, N, nullptr, 0, 1, nullptr);
  if (rc) return rc;
  ctx->n_gp = 1;
  if (femcy_alloc_state(ctx)) return 1;
  int64_t total = N * W;
  double* d_spm = nullptr; int32_t* d_ij = nullptr; uint64_t* keys = nullptr; uint32_t* ids = nullptr; int32_t* eslot = nullptr;
  if (femcy_alloc(ctx, &d_spm, total) || femcy_alloc(ctx, &d_ij, N * (W + 1)) || femcy_alloc(ctx, &keys, total) ||
      femcy_alloc(ctx, &ids, total) || femcy_alloc(ctx, &eslot, total))
    return 1;
  CK(cudaMemcpyAsync(d_spm, spm, (size_t)total * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_ij, sparseIJ, (size_t)N * (W + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  k_ell_keys<<<gridp(total), 256, 0, ctx->stream>>>(d_ij, N, W, keys, ids);
  CK_LAUNCH();
  rc = build_from_keys(ctx, keys, ids, total, N, N, 1, eslot, false);
  if (!rc) {
    k_ell_values<<<gridp(total), 256, 0, ctx->stream>>>(d_spm, eslot, total, ctx->P.val);
    ctx->launches++;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = femcy_fail(ctx, "ell values", e, __FILE__, __LINE__);
  }
  femcy_free(&d_spm); femcy_free(&d_ij); femcy_free(&keys); femcy_free(&ids); femcy_free(&eslot);
  return rc;
}
