// Internal definitions shared by the translation units of libfemcy_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/femcy_b200.h"

#include "kernel_types.cuh"

struct CommState;  // comm.cu

// femcy_set_option: switches that used to be environment variables read inside the hot calls
struct FemcyOptions {
  int cg_kernel = 0;      // 0 auto (streaming persistent; NCCL exchange: three-kernel graph), 1 three-kernel graph, 2 persistent with plain loads, 3 streaming persistent
  int cg_sym = 0;         // 1: PCG SpMV over the upper half of the (symmetric) matrix, fp64 atomics for the transposed products
  int cg_profile = 0;     // 1: per-kernel CUDA events on the three-kernel path (femcy_last_time_ms kinds 4-6)
  int cg_stream_cfg = 0;  // ring shape of the streaming kernel (A/B)
  int no_graph = 0;       // 1: plain launches instead of the CUDA graph of the three-kernel path
  int no_p2p = 0;         // 1: NCCL exchange even where NVLink peer memory is available
  int sell_sigma = -1;    // SELL-32-sigma row order of the NEXT femcy_build_pattern: -1 automatic (1024 when natural-order slices would
                          // be > 15 % padding), 0 natural order, else a multiple of 32
  int consistent_tangent = 0;   // 1: femcy_assemble_K builds the exact linearisation of the internal force (material + geometric
                          // stiffness, k_assemble_scatter_ct) instead of the reference's constant-C stiffness (row f2, opt-in)
  int cg_precond = 0;     // 0 Jacobi (the reference's), 1 two-level: Chebyshev-Jacobi + rigid-body-mode coarse space (precond.cu)
};

// Row f4: a mesh of several SECTIONS -- element sets with their own element kind and material over one node set (Abaqus
// *Solid Section; the reference reader rejects such decks, reader/inp_info.py:125-128, main.py:24).  The ctx fields below
// marked "per section" always describe the SELECTED section; the other sections are parked here.  A single-section mesh
// (every deck the reference accepts) keeps `sections` empty and runs exactly the code it ran before.
struct FemcySection {
  int n_en = 0, n_gp = 0;
  int64_t ne = 0;
  int32_t* elems = nullptr;
  ElemTables tab;
  bool have_elem = false, have_mat = false;
  int mat_kind = -1;
  int32_t* elem_slot = nullptr;
  double *vol = nullptr, *dsdx = nullptr, *F = nullptr, *cauchy = nullptr, *mises = nullptr, *strain = nullptr, *energy = nullptr;
};

struct femcy_ctx {
  FemcyOptions opt;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;
  int64_t bytes = 0;
  int64_t launches = 0;

  // mesh
  int dm = 0, n_en = 0, n_gp = 0, n_v = 0;
  int64_t nn = 0, nn_own = 0, ne = 0;
  double* nodes = nullptr;     // [nn,dm]
  int32_t* elems = nullptr;    // [ne,n_en]                                   (per section: n_en, n_gp, ne, elems, tab,
  ElemTables tab;              //                                              have_*, mat_kind, elem_slot, the per-GP arrays)
  bool have_elem = false, have_mat = false;
  int mat_kind = -1;
  std::vector<FemcySection> sections;   // empty: one section (the fields above); else all sections, the selected one stale
  int cur_section = 0;

  // pattern
  BsellPattern P;
  int32_t* elem_slot = nullptr;   // [ne*n_en*n_en] slot of block (a,b) or -1 (row not owned)
  uint32_t* ent_list = nullptr;   // [n_ent] entries e*P+a*n_en+b sorted by block (gather assembly)
  int32_t* slot_ent_beg = nullptr;  // [nslots+1]... begin offset per slot (end = next real)
  int32_t* slot_ent_end = nullptr;
  int64_t n_ent = 0;

  // vectors
  double* vec[FEMCY_VEC_COUNT] = {nullptr};
  // per-GP arrays
  double *vol = nullptr, *dsdx = nullptr, *F = nullptr, *cauchy = nullptr, *mises = nullptr,
         *strain = nullptr, *energy = nullptr;
  // gather assembly, pass 1: node-sector records [ne][n_en][n_gp][4] = (grad N_a, vol_gp)
  double* egeo4 = nullptr;
  FemcyTmap egeo4_tmap; const double* egeo4_tmap_for = nullptr; int64_t egeo4_tmap_ne = -1;   // TMA store of the C3D4 records
  SymPattern U;                  // upper-half copy of the matrix for the PCG SpMV (option cg_sym)
  bool cg_breakdown = false;     // the last solve stopped on NaN / inf (femcy_cg_breakdown)
  void* precond2 = nullptr;      // state of the two-level preconditioner (precond.cu)
  void* topology = nullptr;      // facet tables, boundary facets, loaded-facet staging (topology.cu, row f1)
  void* partition = nullptr;     // result of the last femcy_partition (partition.cu)

  // scratch for reductions / scalars
  double* red_partials = nullptr;  // [red_cap]
  int64_t red_cap = 0;
  unsigned int* red_ticket = nullptr;
  double* scal = nullptr;          // device scalars, see cg.cu
  double* h_scal = nullptr;        // pinned mirror
  int32_t* bc_nodes = nullptr; int32_t* bc_comps = nullptr; double* bc_vals = nullptr; int64_t bc_cap = 0;
  unsigned char* bc_flag = nullptr;  // [nn*dm]
  double* bc_val_full = nullptr;     // [nn*dm]

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;      // scratch pair (pattern build, cg)
  cudaEvent_t evA0 = nullptr, evA1 = nullptr;    // assemble_K pair (resolved lazily)
  double last_ms[4] = {0, 0, 0, 0};
  double cg_phase_ns[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // femcy_cg_phase_ns
  double prof_ms[3] = {0, 0, 0};                 // option cg_profile: in-loop averages spmv / update_xr / update_d

  // cached CUDA graph of `cg_graph_chunk` CG iterations (cg.cu); dropped when the matrix is rebuilt
  cudaGraphExec_t cg_graph_exec = nullptr;
  int cg_graph_chunk = 0;
  int cg_graph_mode = 0;
  int64_t cg_graph_launches = 0;

  CommState* comm = nullptr;
};

int femcy_fail(femcy_ctx* ctx, const char* what, cudaError_t e, const char* file, int line);
int femcy_fail_msg(femcy_ctx* ctx, const std::string& msg);

#define CK(call)                                                          \
  do {                                                                    \
    cudaError_t _e = (call);                                              \
    if (_e != cudaSuccess) return femcy_fail(ctx, #call, _e, __FILE__, __LINE__); \
  } while (0)

#define CK_LAUNCH()                                                       \
  do {                                                                    \
    ctx->launches++;                                                      \
    cudaError_t _e = cudaGetLastError();                                  \
    if (_e != cudaSuccess) return femcy_fail(ctx, "kernel launch", _e, __FILE__, __LINE__); \
  } while (0)

template <typename T>
int femcy_alloc(femcy_ctx* ctx, T** p, int64_t count) {
  if (*p) {
    cudaError_t fe = cudaFree(*p);
    if (fe != cudaSuccess) return femcy_fail(ctx, "cudaFree", fe, __FILE__, __LINE__);   // keep the old pointer
    *p = nullptr;
  }
  if (count <= 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, (size_t)count * sizeof(T));
  if (e != cudaSuccess) return femcy_fail(ctx, "cudaMalloc", e, __FILE__, __LINE__);
  ctx->bytes += count * (int64_t)sizeof(T);
  return 0;
}
template <typename T>
void femcy_free(T** p) {
  if (*p) { cudaFree(*p); *p = nullptr; }
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// implemented in the other translation units
int femcy_pattern_free(femcy_ctx* ctx);
void femcy_section_park(femcy_ctx* ctx);                 // core.cu: ctx fields -> sections[cur_section]
void femcy_section_load(femcy_ctx* ctx, int s);          //          sections[s] -> ctx fields
void femcy_sections_free(femcy_ctx* ctx);
int femcy_alloc_gp_state(femcy_ctx* ctx);
// run f() once per section with that section selected (once, as is, on a single-section mesh); restores the selection
template <class F>
int femcy_for_sections(femcy_ctx* ctx, F f) {
  if (ctx->sections.empty()) return f();
  const int keep = ctx->cur_section;
  int rc = 0;
  for (int s = 0; s < (int)ctx->sections.size() && !rc; ++s) {
    femcy_section_park(ctx);
    femcy_section_load(ctx, s);
    rc = f();
  }
  femcy_section_park(ctx);
  femcy_section_load(ctx, keep);
  return rc;
}
int femcy_build_sym_pattern(femcy_ctx* ctx);
int femcy_sym_extract(femcy_ctx* ctx);
int femcy_alloc_state(femcy_ctx* ctx);       // vectors + gp arrays after mesh+element known
int femcy_ensure_reduction_scratch(femcy_ctx* ctx, int64_t nblocks);
int femcy_cg_comm_allgather(femcy_ctx* ctx, int nvals);  // comm.cu hook used by cg.cu
int femcy_comm_halo(femcy_ctx* ctx, double* v);
bool femcy_p2p_view(femcy_ctx* ctx, P2PView* pv, const unsigned char** bflag, const int32_t** push_ptr,
                    const int32_t** push_peer, const int32_t** push_ridx, const int32_t** bnodes, int64_t* n_bnodes,
                    const int32_t** slice_order, const unsigned char** slice_ghost, const int4** bpush = nullptr);
int femcy_comm_size(femcy_ctx* ctx);
int femcy_comm_rank(femcy_ctx* ctx);
void femcy_comm_free(femcy_ctx* ctx);
void femcy_precond_free(femcy_ctx* ctx);
void femcy_topology_free(femcy_ctx* ctx);
void femcy_partition_free(femcy_ctx* ctx);
void femcy_drop_graph(femcy_ctx* ctx);
