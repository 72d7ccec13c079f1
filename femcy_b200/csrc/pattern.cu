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

#include "ctx.cuh"

int femcy_pattern_free(femcy_ctx* ctx) {
  BsellPattern& P = ctx->P;
  femcy_free(&P.slice_ptr); femcy_free(&P.blkptr); femcy_free(&P.colidx); femcy_free(&P.diag_slot); femcy_free(&P.val);
  femcy_free(&P.rowof); femcy_free(&P.rowpos);
  femcy_free(&ctx->elem_slot); femcy_free(&ctx->ent_list); femcy_free(&ctx->slot_ent_beg); femcy_free(&ctx->slot_ent_end);
  femcy_free(&ctx->egeo); femcy_free(&ctx->egeo4); femcy_free(&ctx->inc_ptr); femcy_free(&ctx->inc_list);
  P = BsellPattern();
  ctx->n_ent = 0;
  femcy_drop_graph(ctx);
  return 0;
}

__global__ void k_elem_keys(const int32_t* __restrict__ elems, int64_t ne, int n_en, int64_t nn, int64_t nn_own,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ ids) {
  int64_t P = (int64_t)n_en * n_en;
  int64_t total = ne * P;
  uint64_t invalid = (uint64_t)nn_own * (uint64_t)nn;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = t / P;
    int p = (int)(t - e * P);
    int a = p / n_en, b = p - a * n_en;
    int64_t i = elems[e * n_en + a], j = elems[e * n_en + b];
    keys[t] = (i < nn_own) ? (uint64_t)i * (uint64_t)nn + (uint64_t)j : invalid;
    ids[t] = (uint32_t)t;
  }
}

__global__ void k_ell_keys(const int32_t* __restrict__ ij, int64_t N, int W, uint64_t* __restrict__ keys,
                           uint32_t* __restrict__ ids) {
  int64_t total = N * W;
  uint64_t invalid = (uint64_t)N * (uint64_t)N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / W;
    int j0 = (int)(t - i * W);
    int cnt = ij[i * (W + 1)];
    keys[t] = (j0 < cnt) ? (uint64_t)i * (uint64_t)N + (uint64_t)ij[i * (W + 1) + 1 + j0] : invalid;
    ids[t] = (uint32_t)t;
  }
}

__global__ void k_count_valid(const uint64_t* __restrict__ keys, int64_t n, uint64_t invalid, int64_t* __restrict__ n_valid) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    bool v = keys[t] < invalid;
    bool vn = (t + 1 < n) ? (keys[t + 1] < invalid) : false;
    if (v && !vn) *n_valid = t + 1;
  }
}

__global__ void k_heads(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ head) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    head[t] = (t == 0 || keys[t] != keys[t - 1]) ? 1 : 0;
}

// blk_of[t] is the inclusive scan of head (1-based block number). Writes block row/col and first entry.
__global__ void k_block_info(const uint64_t* __restrict__ keys, const int32_t* __restrict__ head,
                             const int32_t* __restrict__ blk_of, int64_t n, int64_t nn, int32_t* __restrict__ brow,
                             int32_t* __restrict__ bcol, int32_t* __restrict__ bfirst) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    if (head[t]) {
      int32_t b = blk_of[t] - 1;
      uint64_t k = keys[t];
      brow[b] = (int32_t)(k / (uint64_t)nn);
      bcol[b] = (int32_t)(k % (uint64_t)nn);
      bfirst[b] = (int32_t)t;
    }
  }
}

__global__ void k_blkptr(const int32_t* __restrict__ brow, int64_t nnzb, int64_t nrows, int32_t* __restrict__ blkptr) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= nrows; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = nnzb;  // first block with brow >= i
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (brow[mid] < i) lo = mid + 1; else hi = mid;
    }
    blkptr[i] = (int32_t)lo;
  }
}

__global__ void k_slice_width(const int32_t* __restrict__ blkptr, int64_t nrows, int64_t nslice,
                              int32_t* __restrict__ slots_per_slice, int32_t* __restrict__ maxw,
                              const int32_t* __restrict__ rowof) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nslice; s += (int64_t)gridDim.x * blockDim.x) {
    int w = 0;
    for (int l = 0; l < FEMCY_SLICE; ++l) {
      int64_t i = s * FEMCY_SLICE + l;
      if (i < nrows) {
        if (rowof) i = rowof[i];
        w = max(w, blkptr[i + 1] - blkptr[i]);
      }
    }
    slots_per_slice[s] = w * FEMCY_SLICE;
    atomicMax(maxw, w);
  }
}

// SELL-32-sigma: sort key of row i = (window i / sigma, descending block count); a stable sort keeps the natural
// order among rows of equal length
__global__ void k_sigma_keys(const int32_t* __restrict__ blkptr, int64_t nrows, int sigma, uint32_t* __restrict__ keys,
                             int32_t* __restrict__ rows) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nrows; i += (int64_t)gridDim.x * blockDim.x) {
    int len = blkptr[i + 1] - blkptr[i];
    if (len > 255) len = 255;
    keys[i] = ((uint32_t)(i / sigma) << 8) | (uint32_t)(255 - len);
    rows[i] = (int32_t)i;
  }
}
__global__ void k_rowpos(const int32_t* __restrict__ rowof, int64_t nrows, int32_t* __restrict__ rowpos) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nrows; p += (int64_t)gridDim.x * blockDim.x)
    rowpos[rowof[p]] = (int32_t)p;
}

__global__ void k_fill_i32(int32_t* __restrict__ p, int32_t v, int64_t n) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) p[t] = v;
}

__global__ void k_block_slots(const int32_t* __restrict__ brow, const int32_t* __restrict__ bcol,
                              const int32_t* __restrict__ bfirst, const int32_t* __restrict__ blkptr,
                              const int32_t* __restrict__ slice_ptr, int64_t nnzb, int64_t n_ent,
                              int32_t* __restrict__ colidx, int32_t* __restrict__ diag_slot,
                              int32_t* __restrict__ bslot, int32_t* __restrict__ slot_beg, int32_t* __restrict__ slot_end,
                              const int32_t* __restrict__ rowpos) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < nnzb; b += (int64_t)gridDim.x * blockDim.x) {
    int32_t i = brow[b];
    int32_t k = (int32_t)b - blkptr[i];
    int32_t pos = rowpos ? rowpos[i] : i;
    int32_t slot = slice_ptr[pos / FEMCY_SLICE] + k * FEMCY_SLICE + (pos % FEMCY_SLICE);
    colidx[slot] = bcol[b];
    bslot[b] = slot;
    if (bcol[b] == i) diag_slot[i] = slot;
    slot_beg[slot] = bfirst[b];
    slot_end[slot] = (b + 1 < nnzb) ? bfirst[b + 1] : (int32_t)n_ent;
  }
}

__global__ void k_entry_slots(const uint32_t* __restrict__ ids, const int32_t* __restrict__ blk_of,
                              const int32_t* __restrict__ bslot, int64_t n_ent, int32_t* __restrict__ entry_slot) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_ent; t += (int64_t)gridDim.x * blockDim.x)
    entry_slot[ids[t]] = bslot[blk_of[t] - 1];
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
  // optional SELL-32-sigma row order (FEMCY_SELL_SIGMA=<multiple of 32>; off by default until measured)
  {
    const char* sg = getenv("FEMCY_SELL_SIGMA");
    int sigma = sg ? atoi(sg) : 0;
    if (sigma < 0 || (sigma % FEMCY_SLICE) != 0) return femcy_fail_msg(ctx, "FEMCY_SELL_SIGMA must be a multiple of 32");
    P.sigma = sigma;
    if (sigma > 0 && nrows > 0) {
      if ((uint64_t)(nrows / sigma) >= ((uint64_t)1 << 24)) return femcy_fail_msg(ctx, "FEMCY_SELL_SIGMA too small for this many rows");
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

extern "C" int femcy_build_pattern(femcy_ctx* ctx, int64_t* nnz_out) {
  cudaSetDevice(ctx->device);
  if (ctx->dm == 0 || !ctx->elems) return femcy_fail_msg(ctx, "set_mesh first");
  femcy_pattern_free(ctx);
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
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

// ---- node -> element incidence lists (rows assembly) -------------------------------------------------
__global__ void k_inc_keys(const int32_t* __restrict__ elems, int64_t total, int64_t nn_own, uint32_t* __restrict__ keys,
                           uint32_t* __restrict__ ids) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int32_t nd = elems[t];
    keys[t] = (nd < nn_own) ? (uint32_t)nd : (uint32_t)nn_own;   // rows of other ranks sort behind the owned ones
    ids[t] = (uint32_t)t;
  }
}
__global__ void k_inc_ptr(const uint32_t* __restrict__ keys, int64_t total, int64_t nrows, int32_t* __restrict__ ptr) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= nrows; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = total;   // first entry with key >= i
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if ((int64_t)keys[mid] < i) lo = mid + 1; else hi = mid;
    }
    ptr[i] = (int32_t)lo;
  }
}

int femcy_build_incidence(femcy_ctx* ctx) {
  if (ctx->inc_ptr && ctx->inc_list) return 0;
  cudaStream_t st = ctx->stream;
  int64_t total = ctx->ne * ctx->n_en;
  if (total >= ((int64_t)1 << 31)) return femcy_fail_msg(ctx, "ne*n_en exceeds int32 incidence offsets");
  uint32_t *keys = nullptr, *ids = nullptr, *keys2 = nullptr;
  if (femcy_alloc(ctx, &keys, total) || femcy_alloc(ctx, &ids, total) || femcy_alloc(ctx, &keys2, total) ||
      femcy_alloc(ctx, &ctx->inc_list, total) || femcy_alloc(ctx, &ctx->inc_ptr, ctx->nn_own + 1))
    return 1;
  if (total > 0) {
    k_inc_keys<<<gridp(total), 256, 0, st>>>(ctx->elems, total, ctx->nn_own, keys, ids);
    CK_LAUNCH();
    int end_bit = 1;
    while (end_bit < 32 && ((uint64_t)ctx->nn_own >> end_bit) != 0) ++end_bit;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ids, ctx->inc_list, total, 0, end_bit, st);
    void* tmp = nullptr;
    CK(cudaMalloc(&tmp, tmp_bytes + 16));
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ids, ctx->inc_list, total, 0, end_bit, st));   // stable: ascending element id per node
    ctx->launches += 8;
    CK(cudaStreamSynchronize(st));
    cudaFree(tmp);
  }
  k_inc_ptr<<<gridp(ctx->nn_own + 1), 256, 0, st>>>(keys2, total, ctx->nn_own, ctx->inc_ptr);
  CK_LAUNCH();
  CK(cudaStreamSynchronize(st));
  femcy_free(&keys); femcy_free(&ids); femcy_free(&keys2);
  return 0;
}

extern "C" int femcy_pattern_stats(femcy_ctx* ctx, int64_t* out4) {
  out4[0] = ctx->P.nnzb; out4[1] = ctx->P.nslots; out4[2] = ctx->P.nslice; out4[3] = ctx->P.max_row_blocks;
  return 0;
}

// ---- scalar CSR view ---------------------------------------------------------------------------
__global__ void k_csr_rowptr(const int32_t* __restrict__ blkptr, int64_t nrows, int dm, int32_t* __restrict__ rowptr) {
  int64_t N = nrows * dm;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= N; r += (int64_t)gridDim.x * blockDim.x) {
    if (r == N) { rowptr[r] = blkptr[nrows] * dm * dm; continue; }
    int64_t i = r / dm; int c = (int)(r - i * dm);
    int nb = blkptr[i + 1] - blkptr[i];
    rowptr[r] = blkptr[i] * dm * dm + c * nb * dm;
  }
}
// mode 0: write colidx; 1: export values; 2: import values
__global__ void k_csr_xfer(BsellPattern P, int32_t* __restrict__ colidx, double* __restrict__ vals, int mode) {
  int dm = P.dm; int dm2 = dm * dm;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P.nn_own; i += (int64_t)gridDim.x * blockDim.x) {
    int nb = P.blkptr[i + 1] - P.blkptr[i];
    int64_t pos = P.rowpos ? P.rowpos[i] : i;
    int64_t base = P.slice_ptr[pos / FEMCY_SLICE];
    int lane = (int)(pos % FEMCY_SLICE);
    for (int r = 0; r < dm; ++r) {
      int64_t o = (int64_t)P.blkptr[i] * dm2 + (int64_t)r * nb * dm;
      for (int k = 0; k < nb; ++k) {
        int64_t slot = base + (int64_t)k * FEMCY_SLICE + lane;
        int32_t cn = P.colidx[slot];
        for (int c = 0; c < dm; ++c) {
          int64_t vi = bsell_val_index(slot, dm2, r * dm + c);
          if (mode == 0) colidx[o + k * dm + c] = cn * dm + c;
          else if (mode == 1) vals[o + k * dm + c] = P.val[vi];
          else P.val[vi] = vals[o + k * dm + c];
        }
      }
    }
  }
}

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
