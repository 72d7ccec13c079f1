// Device code of the pattern build (row a1) -- everything except the CUB sorts / scans that pattern.cu calls between
// these kernels.  Header for the same reason as assembly_kernels.cuh (CPU SIMT emulation in tests/simt).
#pragma once
#include "device_compat.cuh"
#include "kernel_types.cuh"

__global__ void k_elem_keys(const int32_t* __restrict__ elems, int64_t ne, int n_en, int64_t nn, int64_t nn_own,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ ids, uint32_t id_base = 0) {
  int64_t P = (int64_t)n_en * n_en;
  int64_t total = ne * P;
  uint64_t invalid = (uint64_t)nn_own * (uint64_t)nn;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = t / P;
    int p = (int)(t - e * P);
    int a = p / n_en, b = p - a * n_en;
    int64_t i = elems[e * n_en + a], j = elems[e * n_en + b];
    keys[t] = (i < nn_own) ? (uint64_t)i * (uint64_t)nn + (uint64_t)j : invalid;
    ids[t] = id_base + (uint32_t)t;      // id_base: offset of this section's entries in a mesh of several sections (row f4)
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

// ---- symmetric half storage for the PCG SpMV (SymPattern, kernel_types.cuh) -----------------------------------
// one warp per slice: kstart[i] = number of blocks of row i left of the diagonal, uslots[s] = 32 x the widest kept suffix
// (rowof: SELL-32-sigma position -> row node, nullptr = identity; a position behind the last row holds no blocks)
__global__ void k_sym_rows(const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ colidx, int64_t nslice,
                           int32_t* __restrict__ kstart, int32_t* __restrict__ uslots, const int32_t* __restrict__ rowof) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t s = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); s < nslice; s += nw) {
    const int base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5;
    const int64_t pos = s * 32 + lane;
    const int64_t i = rowof ? (int64_t)rowof[pos] : pos;
    int ks = 0, len = 0;
    for (int k = 0; k < w; ++k) {
      int c = colidx[base + (k << 5) + lane];
      if (c >= 0) { ++len; if (c < i) ++ks; }
    }
    kstart[pos] = ks;
    int ul = len - ks;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ul = max(ul, __shfl_xor_sync(0xffffffffu, ul, o));
    if (lane == 0) uslots[s] = ul * 32;
  }
}

__global__ void k_sym_fill(const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ colidx,
                           const int32_t* __restrict__ kstart, const int32_t* __restrict__ u_slice_ptr, int64_t nslice,
                           int32_t* __restrict__ u_colidx, int32_t* __restrict__ u_src) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t s = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); s < nslice; s += nw) {
    const int base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5;
    const int ubase = u_slice_ptr[s], uw = (u_slice_ptr[s + 1] - ubase) >> 5;
    const int ks = kstart[s * 32 + lane];
    for (int ku = 0; ku < uw; ++ku) {
      const int k = ks + ku;
      int src = (k < w) ? base + (k << 5) + lane : -1;
      int c = (src >= 0) ? colidx[src] : -1;
      if (c < 0) src = -1;
      u_colidx[ubase + (ku << 5) + lane] = c;
      u_src[ubase + (ku << 5) + lane] = src;
    }
  }
}

// values of the kept blocks, in the plane layout of the full matrix (run at the start of every solve)
template <int DM>
__global__ void k_sym_extract(const int32_t* __restrict__ u_src, int64_t nslots_u, const double* __restrict__ val,
                              double* __restrict__ u_val) {
  constexpr int DM2 = DM * DM;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nslots_u; t += (int64_t)gridDim.x * blockDim.x) {
    const int src = u_src[t];
    const int64_t g = t >> 5;
    const int lane = (int)(t & 31);
#pragma unroll
    for (int q = 0; q < DM2; ++q)
      u_val[((g * DM2 + q) << 5) + lane] = (src >= 0) ? val[((((int64_t)(src >> 5)) * DM2 + q) << 5) + (src & 31)] : 0.0;
  }
}

