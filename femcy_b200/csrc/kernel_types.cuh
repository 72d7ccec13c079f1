// Plain-data types and constants shared by host code and kernels (no CUDA runtime dependency).
#pragma once
#include <stdint.h>

#define FEMCY_MAX_GP 4
#define FEMCY_MAX_EN 10
#define FEMCY_SLICE 32  // rows (nodes) per SELL slice == warp size

// Element + material tables handed to kernels by value (lives in the constant bank).
struct ElemTables {
  double dN[FEMCY_MAX_GP * FEMCY_MAX_EN * 3];  // [gp][a][k], k < dm
  double w[FEMCY_MAX_GP];
  double C[36];  // [n_v][n_v] row-major (n_v = 3 in 2-D, 6 in 3-D)
  double mat[4]; // constitutive parameters (E,nu | C1,D1)
};

// Node-block SELL-32 matrix: slice s holds 32 consecutive node rows; block k of lane l of the
// slice is slot = slice_ptr[s] + k*32 + l (slice_ptr[s] is a multiple of 32); its dm*dm values
// live in "planes" of 32 lanes, one 32-lane group per (slice, k):
//   val[((slot >> 5)*dm2 + q)*32 + (slot & 31)],  q = r*dm + c
// so for fixed (slice, k, q) a warp reads 256 contiguous bytes and the dm2 planes of one
// (slice, k) group are one contiguous dm2*256 B chunk.
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline int64_t bsell_val_index(int64_t slot, int dm2, int q) {
  return (((slot >> 5) * dm2 + q) << 5) + (slot & 31);
}
// Upper half of the (symmetric) matrix for the PCG SpMV (opt-in, option cg_sym): row i keeps its blocks with
// column >= i -- columns are sorted, so a suffix of the row; ghost columns have the largest local indices and are all
// kept -- in the same SELL-32 layout.  src[slot] = slot of the block in the full pattern (-1: padding).
struct SymPattern {
  int32_t* slice_ptr = nullptr;  // [nslice+1]
  int32_t* colidx = nullptr;     // [nslots]
  int32_t* src = nullptr;        // [nslots]
  double* val = nullptr;         // [nslots*dm2], filled from the full matrix at the start of every solve
  int64_t nslots = 0;
};

struct BsellPattern {
  int dm = 0;
  int64_t nn = 0, nn_own = 0;  // columns / rows (nodes)
  int64_t nslice = 0;
  int64_t nslots = 0;   // stored block slots incl. padding
  int64_t nnzb = 0;     // real blocks
  int max_row_blocks = 0;
  int32_t* slice_ptr = nullptr;  // [nslice+1] in slots
  int32_t* blkptr = nullptr;     // [nn_own+1] CSR-style block row pointer (sorted columns)
  int32_t* colidx = nullptr;     // [nslots] column node or -1
  int32_t* diag_slot = nullptr;  // [nn_own]
  // SELL-32-sigma (option sell_sigma; automatic for quadratic-element meshes): rows are ordered by descending block count inside windows of sigma
  // consecutive nodes before being cut into slices, which removes the padding of meshes whose neighbouring rows
  // differ in length (quadratic elements: 37-40 % padding in natural order, 5 % at sigma = 256).
  // rowof[pos] = row node stored at position pos (slice pos/32, lane pos%32), rowpos = its inverse; nullptr = identity.
  int32_t* rowof = nullptr;      // [nslice*32], -1 behind the last row
  int32_t* rowpos = nullptr;     // [nn_own]
  int sigma = 0;
  double* val = nullptr;         // [nslots*dm2]
};

// Peer-memory view of the other ranks of the box (NVLink/NVSwitch, cudaIpc-mapped): the CG kernels
// store their boundary values / partial sums straight into the peers' memory and spin on flags in
// their own window -- no NCCL call and no extra kernel inside the iteration (comm.cu, cg.cu).
#define FEMCY_MAX_RANKS 8
struct P2PView {
  int nranks = 0, rank = 0;
  double* d_of[FEMCY_MAX_RANKS];                 // every rank's CG direction vector `d`
  unsigned long long* win_of[FEMCY_MAX_RANKS];   // every rank's window (layout below, 8-byte words)
};
// window layout in 8-byte words: flagD[8] (halo flags) | A[8][2] (d.Ad partials) | B[8][4] (rMr, max|r| partials).
// A/B carry no separate flag: every double travels as two self-validating 8-byte words
// {32-bit half of the value, 32-bit exchange tag}; an aligned 8-byte store is single-copy atomic, so the
// reader needs no fence -- it polls until both halves carry the current tag (cg.cu: p2p_allgather).
#define P2P_FLAG_D(r) (r)
#define P2P_SLOT_A(r) (FEMCY_MAX_RANKS + 2 * (r))
#define P2P_SLOT_B(r) (3 * FEMCY_MAX_RANKS + 4 * (r))
#define P2P_WINDOW_WORDS (7 * FEMCY_MAX_RANKS)

// tensor map of the C3D4 record array (TMA store of the gather assembly's first pass); a plain pointer under the emulation
#ifdef FEMCY_SIMT_EMU
struct FemcyTmap { double* base = nullptr; int64_t rows = 0; };
#else
#include <cuda.h>
struct alignas(64) FemcyTmap { CUtensorMap m; };
#endif

// device scalar slots in ctx->scal
enum {
  S_RMR = 0, S_DAD = 1, S_ALPHA = 2, S_BETA = 3, S_RMAX = 4, S_R0 = 5, S_EPS = 6, S_DONE = 7, S_ITER = 8,
  S_FIXED = 9, S_RMR_NEW = 10, S_SEQ = 11, S_ERR = 12,   // S_SEQ: monotone exchange counter of the peer-memory path (never reset)
  // phase clock of the persistent kernel (block 0, nanoseconds summed over the iterations of a solve; femcy_cg_phase_ns):
  // SpMV loop | barrier + fold | cross-rank exchange | x/r update | barrier + fold | exchange | d update + push + barrier
  S_PHASE = 52, S_PHASE_COUNT = 7,
  // multi-GPU staging: [16..] local partials, [24..] gathered
  S_SEND = 16, S_GATHER = 24
};

