// C entry points that run the product's CUDA kernel source on the CPU SIMT emulation (simt.h).
// TEST INFRASTRUCTURE ONLY: built by tests/simt/build.py with g++ -DFEMCY_SIMT_EMU into
// tests/simt/_build/libfemcy_simt.so and loaded only by tests/test_simt_kernels.py.
// The launch geometry mirrors femcy_b200/csrc/{assembly,cg}.cu (grid caps scaled down: every block of a
// cooperative launch is an OS thread here).
#define FEMCY_SIMT_EMU 1
#include "../../include/femcy_b200.h"
#include "../../femcy_b200/csrc/assembly_kernels.cuh"
#include "../../femcy_b200/csrc/cg_kernels.cuh"
#include "../../femcy_b200/csrc/pattern_kernels.cuh"
#include "../../femcy_b200/csrc/topology_kernels.cuh"

#include <algorithm>
#include <thread>

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct EmuAsm {
  int dm, n_en, n_gp;
  const ElemTables* tab;
  const double* nodes;
  const double* dof;
  const int32_t* elems;
  const int32_t* elem_slot;
  int64_t ne;
  const int32_t* slice_ptr;
  int64_t nslice;
  const int32_t* slot_beg;
  const int32_t* slot_end;
  const uint32_t* ent_list;
  int max_row_blocks;
  double* val;
  int64_t nslots;
  double* vol;    // [ne*n_gp]
  double* dsdx;   // [ne*n_gp*n_en*dm] or null
  double* egeo;   // scratch [ne*(n_en*dm+1)] for the single-Gauss-point gather
  int variant;
  int chunk_warps;  // variant-specific knob (0 = default)
  const int32_t* inc_ptr;      // rows assembly (variant 6)
  const uint32_t* inc_list;
  double* egeo4;               // [ne*n_en*4]
  int64_t nn_own;
  const int32_t* rowof;        // SELL-32-sigma position -> row (null: identity)
  const int32_t* tile_ptr; const uint32_t* tile_elems; const uint32_t* ent_tile;   // (unused since round 2: tile lists)
  int max_tile;
};

template <int DM, int NEN, int NGP>
static int emu_dsdx(const EmuAsm& a) {
  if (a.ne == 0) return 0;
  int grid = (int)cdiv(a.ne, 128);
  const ElemTables tab = *a.tab;
  simt::launch(dim3(grid), dim3(128), false, [&]() {
    k_dsdx_vol<DM, NEN, NGP>(tab, a.nodes, a.dof, a.elems, a.ne, a.dsdx, a.vol);
  });
  return 0;
}

template <int DM, int NEN, int NGP>
static int emu_assemble(const EmuAsm& a) {
  constexpr int DM2 = DM * DM;
  if (a.ne == 0) return 0;
  const ElemTables tab = *a.tab;
  int variant = (a.variant == 0 || a.variant == 3) ? FEMCY_ASSEMBLY_GATHER : a.variant;
  if (variant == FEMCY_ASSEMBLY_SCATTER) {
    memset(a.val, 0, (size_t)(a.nslots * DM2) * sizeof(double));
    if constexpr (NEN >= 8) {
      int64_t blocks = cdiv(a.ne, 4);
      if (blocks > 24) blocks = 24;   // product: 148*64
      simt::launch(dim3((unsigned)blocks), dim3(128), false, [&]() {
        k_assemble_scatter_warp<DM, NEN, NGP>(tab, a.nodes, a.dof, a.elems, a.elem_slot, a.ne, a.val);
      });
    } else {
      simt::launch(dim3((unsigned)cdiv(a.ne, 128)), dim3(128), false, [&]() {
        k_assemble_scatter<DM, NEN, NGP, 1>(tab, a.nodes, a.dof, a.elems, a.elem_slot, a.ne, a.val);
      });
    }
    return 0;
  }
  if (variant == 4) {      // option consistent_tangent (assembly.cu); the knob carries the material kind
    memset(a.val, 0, (size_t)(a.nslots * DM2) * sizeof(double));
    simt::launch(dim3((unsigned)cdiv(a.ne, 128)), dim3(128), false, [&]() {
      k_assemble_scatter_ct<DM, NEN, NGP>(tab, a.chunk_warps, a.nodes, a.dof, a.elems, a.elem_slot, a.ne, a.val);
    });
    return 0;
  }
  if (variant != FEMCY_ASSEMBLY_GATHER) return 2;
  // pass 1 (assembly.cu: the TMA tensor store for 128-byte records, else the staged copy-out)
  bool tma_store = false;
  if constexpr (NEN == 4 && NGP == 1) {
    FemcyTmap tm;
    tm.base = a.egeo4; tm.rows = a.ne;
    simt::launch(dim3((unsigned)cdiv(a.ne, 128)), dim3(128), false, [&]() {
      k_elem_geometry4t<DM, NEN>(tab, a.nodes, a.dof, a.elems, a.ne, tm, a.vol);
    });
    tma_store = true;
  }
  if (!tma_store) {
    using G = Geo4Cfg<NEN, NGP>;
    simt::launch(dim3((unsigned)cdiv(a.ne, G::TPB)), dim3(G::TPB), false, [&]() {
      k_elem_geometry4s<DM, NEN, NGP>(tab, a.nodes, a.dof, a.elems, a.ne, a.egeo4, a.vol);
    });
  }
  if constexpr (NGP == 4) {
    if (a.variant != 3) {       // assembly.cu: a quad of lanes per stored block for four-Gauss-point elements
      if (tangent_is_cubic(tab.C, DM))
        simt::launch(dim3((unsigned)a.nslice), dim3(32, 8), false, [&]() {
          k_assemble_gather_q<DM, NEN, true>(tab, a.slice_ptr, a.slot_beg, a.slot_end, a.ent_list, a.egeo4, a.val, a.nslice);
        });
      else
        simt::launch(dim3((unsigned)a.nslice), dim3(32, 8), false, [&]() {
          k_assemble_gather_q<DM, NEN, false>(tab, a.slice_ptr, a.slot_beg, a.slot_end, a.ent_list, a.egeo4, a.val, a.nslice);
        });
      return 0;
    }
  }
  if constexpr (NGP == 1) {
    if (a.variant != 3) {       // assembly.cu: a pair of lanes per stored block for one-Gauss-point elements
      if (tangent_is_cubic(tab.C, DM))
        simt::launch(dim3((unsigned)a.nslice), dim3(32, 8), false, [&]() {
          k_assemble_gather_h<DM, NEN, true>(tab, a.slice_ptr, a.slot_beg, a.slot_end, a.ent_list, a.egeo4, a.val, a.nslice);
        });
      else
        simt::launch(dim3((unsigned)a.nslice), dim3(32, 8), false, [&]() {
          k_assemble_gather_h<DM, NEN, false>(tab, a.slice_ptr, a.slot_beg, a.slot_end, a.ent_list, a.egeo4, a.val, a.nslice);
        });
      return 0;
    }
  }
  if (tangent_is_cubic(tab.C, DM))
    simt::launch(dim3((unsigned)a.nslice), dim3(32, 8), false, [&]() {
      k_assemble_gather_p<DM, NEN, NGP, true>(tab, a.slice_ptr, a.slot_beg, a.slot_end, a.ent_list, a.egeo4, a.val, a.nslice);
    });
  else
    simt::launch(dim3((unsigned)a.nslice), dim3(32, 8), false, [&]() {
      k_assemble_gather_p<DM, NEN, NGP, false>(tab, a.slice_ptr, a.slot_beg, a.slot_end, a.ent_list, a.egeo4, a.val, a.nslice);
    });
  return 0;
}

#define EMU_DISPATCH(FN, a)                                 \
  switch ((a).dm * 1000 + (a).n_en * 10 + (a).n_gp) {       \
    case 2031: return FN<2, 3, 1>(a);                       \
    case 2063: return FN<2, 6, 3>(a);                       \
    case 2044: return FN<2, 4, 4>(a);                       \
    case 2084: return FN<2, 8, 4>(a);                       \
    case 3041: return FN<3, 4, 1>(a);                       \
    case 3104: return FN<3, 10, 4>(a);                      \
    default: return 3;                                      \
  }

extern "C" int emu_get_dsdx_and_vol(const EmuAsm* a) { EMU_DISPATCH(emu_dsdx, *a); }
extern "C" int emu_assemble_K(const EmuAsm* a) { EMU_DISPATCH(emu_assemble, *a); }
extern "C" int emu_sizeof_tables() { return (int)sizeof(ElemTables); }

// ---- PCG ---------------------------------------------------------------------------------------------
#include <condition_variable>
#include <mutex>

struct HostBarrier {
  std::mutex m; std::condition_variable cv; int n, count = 0, gen = 0;
  explicit HostBarrier(int n_) : n(n_) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    int g = gen;
    if (++count == n) { count = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};

struct EmuCG {
  int dm;
  int64_t nn_own, nn, nslice;
  const int32_t* slice_ptr; const int32_t* colidx; const int32_t* diag_slot; const double* val;
  const double* b;
  double *x, *r, *d, *M, *Ad;        // [nn*dm]
  double* scal;                      // [64]
  double* partials;                  // [>= 4*max grid]
  unsigned int* ticket;              // [8]
  double eps; int64_t max_iter; int check_every; int fixed;
  int rank, nranks;
  double* d_of[FEMCY_MAX_RANKS];
  unsigned long long* win_of[FEMCY_MAX_RANKS];
  const unsigned char* bflag; const int32_t *push_ptr, *push_peer, *push_ridx, *bnodes; int64_t n_bnodes;
  const int32_t* slice_order; const unsigned char* slice_ghost;
  int persistent_grid;               // blocks of the cooperative launch
  int64_t iters_out; double r0_out, rmax_out;
  int variant;                       // CG algorithm variant (0 = reference recurrence)
  const int32_t* rowof;              // SELL-32-sigma position -> row (null: identity)
  int late_fence;                    // removed in round 2 (must be 0)
  int fold_bar;                      // removed in round 2 (must be 0)
  int sym;                           // option cg_sym: upper-half SpMV with transposed scatter (persistent kernel)
};

static inline int emu_vec_grid(int64_t n) {
  int64_t g = cdiv(n, 256 * 4);
  if (g > 6) g = 6;
  if (g < 1) g = 1;
  return (int)g;
}

template <int DM>
static int emu_cg_rank(EmuCG& c, int mode, HostBarrier* hb, EmuCG* all) {
  const int nranks = c.nranks;
  const int multi = nranks > 1 ? 2 : 0;       // the emulation covers the single-GPU and the peer-memory path
  P2PView pv;
  pv.nranks = nranks; pv.rank = c.rank;
  for (int r = 0; r < FEMCY_MAX_RANKS; ++r) { pv.d_of[r] = c.d_of[r]; pv.win_of[r] = c.win_of[r]; }
  const int64_t n = c.nn_own * DM;
  memset(c.ticket, 0, 8 * sizeof(unsigned int));
  simt::launch(dim3(1), dim3(1), false, [&]() { k_set_scalars(c.scal, c.eps, c.fixed ? 1.0 : 0.0); });
  int vg = emu_vec_grid(n);
  simt::launch(dim3(vg), dim3(256), false, [&]() {
    k_cg_init<DM>(c.diag_slot, c.val, c.b, c.x, c.r, c.d, c.M, c.Ad, c.nn_own, c.partials, c.ticket, c.scal, multi);
  });
  if (multi) {
    hb->wait();
    for (int r = 0; r < nranks; ++r) {
      c.scal[S_GATHER + 2 * r] = all[r].scal[S_SEND];
      c.scal[S_GATHER + 2 * r + 1] = all[r].scal[S_SEND + 1];
    }
    hb->wait();
    simt::launch(dim3(1), dim3(1), false, [&]() { k_finish_init(c.scal, nranks); });
  }
  auto update_d = [&]() {
    if (multi == 2) {
      unsigned int* tk = c.ticket + 4;
      simt::launch(dim3(vg), dim3(256), false, [&]() {
        k_update_d_p2p<DM>(c.d, c.r, c.M, (int)n, c.scal, pv, c.bflag, c.push_ptr, c.push_peer, c.push_ridx, c.bnodes,
                           (int)c.n_bnodes, tk);
      });
    } else {
      simt::launch(dim3(vg), dim3(256), false, [&]() { k_update_d(c.d, c.r, c.M, n, c.scal); });
    }
  };
  if (multi == 2) update_d();
  int sgrid = (int)cdiv(c.nslice, 8);
  if (sgrid < 1) sgrid = 1;
  auto iteration = [&]() {
    if (multi == 2)
      simt::launch(dim3(sgrid), dim3(256), false, [&]() {
        k_spmv_dot<DM, true>(c.slice_ptr, c.colidx, c.val, c.d, c.Ad, c.nn_own, c.nslice, c.partials, c.ticket, c.scal, 1,
                             multi, pv, c.slice_order, c.slice_ghost, c.rowof);
      });
    else
      simt::launch(dim3(sgrid), dim3(256), false, [&]() {
        k_spmv_dot<DM, false>(c.slice_ptr, c.colidx, c.val, c.d, c.Ad, c.nn_own, c.nslice, c.partials, c.ticket, c.scal, 1,
                              multi, pv, c.slice_order, c.slice_ghost, c.rowof);
      });
    simt::launch(dim3(vg), dim3(256), false, [&]() {
      k_update_xr(c.x, c.r, c.d, c.Ad, c.M, n, c.partials, c.ticket, c.scal, multi, pv);
    });
    update_d();
  };
  CGPersistArgs pa;
  int pgrid = c.persistent_grid;
  if (pgrid > sgrid) pgrid = sgrid;
  if (pgrid < 1) pgrid = 1;
  pa.slice_ptr = c.slice_ptr; pa.colidx = c.colidx; pa.val = c.val; pa.nrows = c.nn_own; pa.nslice = c.nslice;
  pa.x = c.x; pa.r = c.r; pa.d = c.d; pa.Ad = c.Ad; pa.M = c.M; pa.n = n;
  pa.part1 = c.partials; pa.part2 = c.partials + pgrid;
  pa.scal = c.scal; pa.p2p = (multi == 2) ? 1 : 0;
  pa.pv = pv; pa.bflag = c.bflag; pa.push_ptr = c.push_ptr; pa.push_peer = c.push_peer; pa.push_ridx = c.push_ridx;
  pa.bnodes = c.bnodes; pa.n_bnodes = (int)c.n_bnodes; pa.slice_order = c.slice_order; pa.slice_ghost = c.slice_ghost;
  pa.ticket = c.ticket + 6;
  // comm.cu: one record per boundary node {node, first peer, first remote index, further push entries}
  std::vector<int4> bpush((size_t)(c.n_bnodes > 0 ? c.n_bnodes : 1));
  for (int64_t k = 0; k < c.n_bnodes; ++k) {
    const int32_t nd = c.bnodes[k], o = c.push_ptr[nd];
    bpush[k] = make_int4(nd, c.push_peer[o], c.push_ridx[o], c.push_ptr[nd + 1] - o - 1);
  }
  pa.bpush = bpush.data();
  pa.rowof = c.rowof;
  if (c.variant != 0 || c.late_fence != 0 || c.fold_bar != 0) return 8;   // removed in round 2 (measured slower on hardware)
  // opt-in symmetric half storage (cg.cu: option cg_sym; pattern.cu: femcy_build_sym_pattern / femcy_sym_extract)
  std::vector<int32_t> u_kstart, u_slots, u_sptr, u_col, u_src;
  std::vector<double> u_val;
  if (c.sym) {
    if (mode != 1 && mode != 2) return 7;
    const int64_t ns = c.nslice;
    u_kstart.assign((size_t)ns * 32, 0); u_slots.assign((size_t)ns + 1, 0); u_sptr.assign((size_t)ns + 1, 0);
    int wgrid = (int)cdiv(ns, 8); if (wgrid > 3) wgrid = 3; if (wgrid < 1) wgrid = 1;
    simt::launch(dim3(wgrid), dim3(256), false, [&]() { k_sym_rows(c.slice_ptr, c.colidx, ns, u_kstart.data(), u_slots.data(), c.rowof); });
    for (int64_t q = 0; q < ns; ++q) u_sptr[q + 1] = u_sptr[q] + u_slots[q];      // cub::DeviceScan::ExclusiveSum
    const int64_t nu = u_sptr[ns];
    u_col.assign((size_t)nu + 1, -7); u_src.assign((size_t)nu + 1, -7); u_val.assign((size_t)nu * DM * DM + 1, NAN);
    simt::launch(dim3(wgrid), dim3(256), false, [&]() {
      k_sym_fill(c.slice_ptr, c.colidx, u_kstart.data(), u_sptr.data(), ns, u_col.data(), u_src.data());
    });
    int eg = (int)cdiv(nu, 256); if (eg > 4) eg = 4; if (eg < 1) eg = 1;
    simt::launch(dim3(eg), dim3(256), false, [&]() { k_sym_extract<DM>(u_src.data(), nu, c.val, u_val.data()); });
    memset(c.Ad, 0, (size_t)n * sizeof(double));
    pa.sym = 1; pa.u_slice_ptr = u_sptr.data(); pa.u_colidx = u_col.data(); pa.u_val = u_val.data();
  }
  int64_t it = 0;
  bool done = false;
  while (it < c.max_iter && !done) {
    int64_t chunk = c.check_every;
    if (it + chunk > c.max_iter) chunk = c.max_iter - it;
    if (mode == 2) {
      // streaming persistent kernel: a small ring (4 warps per block, 2 block columns per stage, 2 stages)
      pa.iters = (int)chunk;
      int sg = pgrid;
      int need = (int)cdiv(c.nslice, 4);
      if (sg > need) sg = need < 1 ? 1 : need;
      pa.part2 = c.partials + sg;
      const size_t smem = CGStreamCfg<DM, 4, 2, 2>::SMEM_BYTES;
      if (c.sym) simt::launch(dim3(sg), dim3(128), true, [&]() { k_cg_stream<DM, 4, 2, 2, true>(pa); }, smem);
      else simt::launch(dim3(sg), dim3(128), true, [&]() { k_cg_stream<DM, 4, 2, 2, false>(pa); }, smem);
    } else if (mode == 1) {
      pa.iters = (int)chunk;
      if (c.sym) simt::launch(dim3(pgrid), dim3(256), true, [&]() { k_cg_persistent<DM, 4, true>(pa); });
      else simt::launch(dim3(pgrid), dim3(256), true, [&]() { k_cg_persistent<DM>(pa); });
    } else {
      for (int64_t k = 0; k < chunk; ++k) iteration();
    }
    it += chunk;
    if (c.scal[S_DONE] != 0.0) done = true;
  }
  c.iters_out = (int64_t)c.scal[S_ITER];
  c.r0_out = c.scal[S_R0];
  c.rmax_out = c.scal[S_RMAX];
  return c.scal[S_DONE] == 3.0 ? 5 : 0;
}

// mode 0: three kernels per iteration, 1: persistent cooperative kernel, 2: streaming persistent kernel.  ranks[nranks] run concurrently.
extern "C" int emu_cg_solve(EmuCG* ranks, int nranks, int mode) {
  HostBarrier hb(nranks);
  std::vector<int> rc(nranks, 0);
  auto run = [&](int r) {
    EmuCG& c = ranks[r];
    switch (c.dm) {
      case 1: rc[r] = emu_cg_rank<1>(c, mode, &hb, ranks); break;
      case 2: rc[r] = emu_cg_rank<2>(c, mode, &hb, ranks); break;
      case 3: rc[r] = emu_cg_rank<3>(c, mode, &hb, ranks); break;
      default: rc[r] = 3;
    }
  };
  if (nranks == 1) { run(0); return rc[0]; }
  std::vector<std::thread> th;
  for (int r = 0; r < nranks; ++r) th.emplace_back(run, r);
  for (auto& t : th) t.join();
  for (int r = 0; r < nranks; ++r) if (rc[r]) return rc[r];
  return 0;
}

extern "C" int emu_spmv(int dm, int64_t nn_own, int64_t nslice, const int32_t* slice_ptr, const int32_t* colidx,
                        const double* val, const double* x, double* y, double* partials, unsigned int* ticket, double* scal,
                        const int32_t* rowof) {
  int sgrid = (int)cdiv(nslice, 8);
  if (sgrid < 1) sgrid = 1;
  P2PView pv;
  auto go = [&](auto tag) {
    constexpr int DM = decltype(tag)::value;
    simt::launch(dim3(sgrid), dim3(256), false, [&]() {
      k_spmv_dot<DM, false>(slice_ptr, colidx, val, x, y, nn_own, nslice, partials, ticket, scal, 0, 0, pv, nullptr, nullptr, rowof);
    });
  };
  switch (dm) {
    case 1: go(std::integral_constant<int, 1>()); break;
    case 2: go(std::integral_constant<int, 2>()); break;
    case 3: go(std::integral_constant<int, 3>()); break;
    default: return 3;
  }
  return 0;
}

// ---- Dirichlet (bc.cu: bc_apply) -------------------------------------------------------------------------
#include "../../femcy_b200/csrc/bc_kernels.cuh"

// mode 0: linear equations (target = rhs), 1: Newton (target = residual).  flag / valfull: scratch [nn*dm] (zeroed).
extern "C" int emu_dirichlet(int dm, int64_t nn_own, int64_t nslice, const int32_t* slice_ptr, const int32_t* colidx,
                             double* val, const int32_t* rowof, const int32_t* nodes, const int32_t* comps,
                             const double* vals, int64_t n, unsigned char* flag, double* valfull, double* target, int mode) {
  if (n == 0) return 0;
  int gb = (int)cdiv(n, 256);
  if (gb > 4) gb = 4;
  simt::launch(dim3(gb), dim3(256), false, [&]() { k_bc_mark(nodes, comps, mode == 0 ? vals : nullptr, n, dm, flag, valfull, 1); });
  int grid = (int)cdiv(nslice, 8);
  if (grid < 1) grid = 1;
  auto go = [&](auto tag) {
    constexpr int DM = decltype(tag)::value;
    simt::launch(dim3(grid), dim3(256), false, [&]() {
      k_bc_apply<DM>(slice_ptr, colidx, val, nn_own, nslice, flag, valfull, target, mode, rowof);
    });
  };
  switch (dm) {
    case 1: go(std::integral_constant<int, 1>()); break;
    case 2: go(std::integral_constant<int, 2>()); break;
    case 3: go(std::integral_constant<int, 3>()); break;
    default: return 3;
  }
  simt::launch(dim3(gb), dim3(256), false, [&]() { k_bc_mark(nodes, comps, nullptr, n, dm, flag, valfull, 0); });
  return 0;
}

// ---- stress recovery / internal force / energy (post.cu) ----------------------------------------------------
#include "../../femcy_b200/csrc/post_kernels.cuh"

struct EmuPost {
  int dm, n_en, n_gp, kind;
  const ElemTables* tab;
  const double* nodes; const double* dof; const int32_t* elems; int64_t ne, nn_own;
  double *F, *cauchy, *vol, *dsdx, *force, *out;   // out: strain [ngp*dd] / mises [ngp] / energy [ngp]
  double* partials; unsigned int* ticket; double* total;
};

template <int DM, int NEN, int NGP>
static int emu_defgrad(const EmuPost& a) {
  const ElemTables tab = *a.tab;
  if (a.ne == 0) return 0;
  simt::launch(dim3((unsigned)cdiv(a.ne, 128)), dim3(128), false, [&]() {
    k_defgrad<DM, NEN, NGP>(tab, a.nodes, a.dof, a.elems, a.ne, a.F);
  });
  return 0;
}
template <int DM, int NEN, int NGP>
static int emu_force(const EmuPost& a) {
  const ElemTables tab = *a.tab;
  if (a.ne == 0) return 0;
  simt::launch(dim3((unsigned)cdiv(a.ne, 128)), dim3(128), false, [&]() {
    k_internal_force<DM, NEN, NGP>(tab, a.kind, a.nodes, a.dof, a.elems, a.ne, a.nn_own, a.F, a.cauchy, a.vol, a.dsdx, a.force);
  });
  return 0;
}
extern "C" int emu_deformation_gradient(const EmuPost* a) { EMU_DISPATCH(emu_defgrad, *a); }
extern "C" int emu_internal_force(const EmuPost* a) { EMU_DISPATCH(emu_force, *a); }
// what: 0 constitutive -> cauchy ; 1 strain ; 2 mises (from cauchy) ; 3 energy density (from F) + total = sum(energy*vol)
extern "C" int emu_per_gp(const EmuPost* a, int what, int large) {
  const ElemTables tab = *a->tab;
  int64_t ngp = a->ne * a->n_gp;
  if (ngp == 0) return 0;
  unsigned grid = (unsigned)cdiv(ngp, 256);
  if (a->dm == 2)
    simt::launch(dim3(grid), dim3(256), false, [&]() { k_per_gp<2>(tab, a->kind, large, what, a->F, a->cauchy, a->out, ngp); });
  else
    simt::launch(dim3(grid), dim3(256), false, [&]() { k_per_gp<3>(tab, a->kind, large, what, a->F, a->cauchy, a->out, ngp); });
  if (what == 3) {
    int64_t g64 = cdiv(ngp, 1024);
    unsigned g = (unsigned)(g64 > 6 ? 6 : g64);
    simt::launch(dim3(g), dim3(256), false, [&]() { k_weighted_sum(a->out, a->vol, ngp, a->partials, a->ticket, a->total); });
  }
  return 0;
}

// ---- pattern build (pattern.cu: build_from_keys / femcy_build_incidence) -----------------------------------
// The kernels are the product's; the CUB radix sorts / scans between them are replaced by std::stable_sort and
// host loops (CUB itself is not under test).  Mirrors the order of operations of build_from_keys.
#include <numeric>

struct EmuPattern {
  const int32_t* elems; int64_t ne; int n_en; int64_t nn, nn_own; int sigma;
  int64_t cap_slots;                       // capacity of the per-slot outputs
  int64_t stats[4];                        // nnzb, nslots, nslice, max_row_blocks
  int32_t *blkptr, *slice_ptr, *colidx, *diag_slot, *slot_beg, *slot_end, *elem_slot;
  uint32_t* ent_list; int64_t n_ent;
  int32_t *rowof, *rowpos, *inc_ptr; uint32_t* inc_list;
  int32_t* tile_ptr; uint32_t* tile_elems; uint32_t* ent_tile; int64_t n_tile; int max_tile;   // femcy_build_tiles
  int rb_shift;
  // row f4 (pattern.cu: build_pattern_sections): nsec > 0 -> the keys come from these sections instead of (elems, ne, n_en);
  // elem_slot then holds the sections' slots back to back
  int nsec; const int32_t* elems_s[8]; int64_t ne_s[8]; int n_en_s[8];
};

static unsigned egrid(int64_t n) { int64_t g = cdiv(n, 256); if (g > 6) g = 6; if (g < 1) g = 1; return (unsigned)g; }

extern "C" int emu_build_pattern(EmuPattern* p) {
  const int64_t Pn = (int64_t)p->n_en * p->n_en, nrows = p->nn_own, ncols = p->nn;
  int64_t total = p->ne * Pn;
  if (p->nsec > 0) { total = 0; for (int s = 0; s < p->nsec; ++s) total += p->ne_s[s] * (int64_t)p->n_en_s[s] * p->n_en_s[s]; }
  std::vector<uint64_t> keys(total), keys2(total);
  std::vector<uint32_t> ids(total), ids2(total);
  if (p->nsec > 0) {
    int64_t off = 0;
    for (int s = 0; s < p->nsec; ++s) {
      const int64_t cnt = p->ne_s[s] * (int64_t)p->n_en_s[s] * p->n_en_s[s];
      if (cnt > 0)
        simt::launch(dim3(egrid(cnt)), dim3(256), false, [&]() {
          k_elem_keys(p->elems_s[s], p->ne_s[s], p->n_en_s[s], p->nn, p->nn_own, keys.data() + off, ids.data() + off, (uint32_t)off);
        });
      off += cnt;
    }
  } else {
    simt::launch(dim3(egrid(total)), dim3(256), false, [&]() { k_elem_keys(p->elems, p->ne, p->n_en, p->nn, p->nn_own, keys.data(), ids.data()); });
  }
  std::vector<int64_t> order(total);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return keys[a] < keys[b]; });
  for (int64_t t = 0; t < total; ++t) { keys2[t] = keys[order[t]]; ids2[t] = ids[order[t]]; }
  uint64_t invalid = (uint64_t)nrows * (uint64_t)ncols;
  int64_t n_ent = 0;
  simt::launch(dim3(egrid(total)), dim3(256), false, [&]() { k_count_valid(keys2.data(), total, invalid, &n_ent); });
  p->n_ent = n_ent;
  std::vector<int32_t> head(n_ent > 0 ? n_ent : 1), blk_of(n_ent > 0 ? n_ent : 1);
  simt::launch(dim3(egrid(n_ent)), dim3(256), false, [&]() { k_heads(keys2.data(), n_ent, head.data()); });
  int32_t run = 0;
  for (int64_t t = 0; t < n_ent; ++t) { run += head[t]; blk_of[t] = run; }
  const int64_t nnzb = n_ent > 0 ? blk_of[n_ent - 1] : 0;
  std::vector<int32_t> brow(nnzb + 1), bcol(nnzb + 1), bfirst(nnzb + 1), bslot(nnzb + 1);
  simt::launch(dim3(egrid(n_ent)), dim3(256), false, [&]() {
    k_block_info(keys2.data(), head.data(), blk_of.data(), n_ent, ncols, brow.data(), bcol.data(), bfirst.data());
  });
  simt::launch(dim3(egrid(nrows + 1)), dim3(256), false, [&]() { k_blkptr(brow.data(), nnzb, nrows, p->blkptr); });
  const int64_t nslice = cdiv(nrows, 32);
  const int32_t* rowof = nullptr; const int32_t* rowpos = nullptr;
  if (p->sigma > 0 && nrows > 0) {
    std::vector<uint32_t> k1(nrows); std::vector<int32_t> r1(nrows);
    simt::launch(dim3(egrid(nslice * 32)), dim3(256), false, [&]() { k_fill_i32(p->rowof, -1, nslice * 32); });
    simt::launch(dim3(egrid(nrows)), dim3(256), false, [&]() { k_sigma_keys(p->blkptr, nrows, p->sigma, k1.data(), r1.data()); });
    std::vector<int64_t> o2(nrows);
    std::iota(o2.begin(), o2.end(), 0);
    std::stable_sort(o2.begin(), o2.end(), [&](int64_t a, int64_t b) { return k1[a] < k1[b]; });
    for (int64_t t = 0; t < nrows; ++t) p->rowof[t] = r1[o2[t]];
    simt::launch(dim3(egrid(nrows)), dim3(256), false, [&]() { k_rowpos(p->rowof, nrows, p->rowpos); });
    rowof = p->rowof; rowpos = p->rowpos;
  }
  std::vector<int32_t> sps(nslice + 1, 0);
  int32_t maxw = 0;
  simt::launch(dim3(egrid(nslice)), dim3(256), false, [&]() { k_slice_width(p->blkptr, nrows, nslice, sps.data(), &maxw, rowof); });
  int32_t acc = 0;
  for (int64_t s = 0; s <= nslice; ++s) { p->slice_ptr[s] = acc; acc += sps[s]; }
  const int64_t nslots = p->slice_ptr[nslice];
  if (nslots > p->cap_slots) return 7;
  simt::launch(dim3(egrid(nslots)), dim3(256), false, [&]() { k_fill_i32(p->colidx, -1, nslots); });
  simt::launch(dim3(egrid(nrows)), dim3(256), false, [&]() { k_fill_i32(p->diag_slot, -1, nrows); });
  memset(p->slot_beg, 0, (size_t)nslots * 4);
  memset(p->slot_end, 0, (size_t)nslots * 4);
  simt::launch(dim3(egrid(nnzb)), dim3(256), false, [&]() {
    k_block_slots(brow.data(), bcol.data(), bfirst.data(), p->blkptr, p->slice_ptr, nnzb, n_ent, p->colidx, p->diag_slot,
                  bslot.data(), p->slot_beg, p->slot_end, rowpos);
  });
  simt::launch(dim3(egrid(total)), dim3(256), false, [&]() { k_fill_i32(p->elem_slot, -1, total); });
  simt::launch(dim3(egrid(n_ent)), dim3(256), false, [&]() { k_entry_slots(ids2.data(), blk_of.data(), bslot.data(), n_ent, p->elem_slot); });
  for (int64_t t = 0; t < n_ent; ++t) p->ent_list[t] = ids2[t];
  p->stats[0] = nnzb; p->stats[1] = nslots; p->stats[2] = nslice; p->stats[3] = maxw;
  p->n_tile = 0; p->max_tile = 0;     // (incidence / tile lists: removed with the kernels that used them)
  return 0;
}

extern "C" int emu_gp_sum(const double* a, int64_t n, double* partials, unsigned int* ticket, double* out) {
  int64_t g64 = cdiv(n > 0 ? n : 1, 1024);
  unsigned g = (unsigned)(g64 > 6 ? 6 : g64);
  simt::launch(dim3(g), dim3(256), false, [&]() { k_weighted_sum(a, nullptr, n, partials, ticket, out); });
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// Row f1: topology builders + Neumann vector (femcy_b200/csrc/topology.cu with the CUB sort / scan replaced by host loops)
struct EmuTopo {
  const int32_t* elems; int64_t ne; int n_en; int64_t nn; int dm;
  const double* nodes;
  int nkeys, width, nfp;
  const int32_t* key_nodes; const double* w; const double* normal; const double* N; const double* dN;
  int32_t* b_elem; int32_t* b_kid; int64_t n_boundary;                 // out: boundary facets (capacity ne*nkeys)
  int32_t* ne_ptr; int32_t* ne_list;                                    // out: node -> elements CSR
  int64_t nf; const int32_t* f_elem; const int32_t* f_kid; double traction; int has_dir; double dir[3]; double* rhs;
};

static unsigned tgrid(int64_t n) { int64_t g = cdiv(n > 0 ? n : 1, 256); if (g > 6) g = 6; return (unsigned)g; }

extern "C" int emu_boundary_facets(EmuTopo* p) {
  const int64_t ne = p->ne, total = ne * p->nkeys;
  p->n_boundary = 0;
  if (total == 0) return 0;
  std::vector<uint64_t> keys(total), keys2(total);
  std::vector<uint32_t> ids(total), ids2(total);
  simt::launch(dim3(tgrid(total)), dim3(256), false, [&]() {
    k_facet_keys(p->elems, ne, p->n_en, p->nn, p->key_nodes, p->nkeys, p->width, keys.data(), ids.data());
  });
  std::vector<int64_t> order(total);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return keys[a] < keys[b]; });
  for (int64_t t = 0; t < total; ++t) { keys2[t] = keys[order[t]]; ids2[t] = ids[order[t]]; }
  std::vector<int32_t> flag(total), pos(total);
  simt::launch(dim3(tgrid(total)), dim3(256), false, [&]() {
    k_facet_unique(keys2.data(), ids2.data(), total, p->elems, ne, p->n_en, p->key_nodes, p->width, flag.data());
  });
  int32_t run = 0;
  for (int64_t t = 0; t < total; ++t) { pos[t] = run; run += flag[t]; }
  p->n_boundary = run;
  simt::launch(dim3(tgrid(total)), dim3(256), false, [&]() { k_facet_compact(flag.data(), pos.data(), total, ne, p->b_elem, p->b_kid); });
  return 0;
}

extern "C" int emu_node_elements(EmuTopo* p) {
  const int64_t ne = p->ne, total = ne * p->n_en;
  std::vector<uint64_t> keys(total > 0 ? total : 1);
  if (total > 0) {
    simt::launch(dim3(tgrid(total)), dim3(256), false, [&]() { k_node_elem_keys(p->elems, ne, p->n_en, keys.data()); });
    std::sort(keys.begin(), keys.begin() + total);
  }
  simt::launch(dim3(tgrid(total > p->nn ? total : p->nn + 1)), dim3(256), false, [&]() {
    k_node_elem_csr(keys.data(), total, ne > 0 ? ne : 1, p->nn, p->ne_ptr, p->ne_list);
  });
  return 0;
}

extern "C" int emu_neumann(EmuTopo* p) {
  FacetTables T;
  T.nkeys = p->nkeys; T.width = p->width; T.nfp = p->nfp;
  T.key_nodes = p->key_nodes; T.w = p->w; T.normal = p->normal; T.N = p->N; T.dN = p->dN;
  const int64_t N = p->nn * p->dm;
  for (int64_t i = 0; i < N; ++i) p->rhs[i] = 0.0;
  if (p->nf == 0) return 0;
  int64_t g = cdiv(p->nf, 128);
  if (g > 6) g = 6;
  if (p->dm == 2)
    simt::launch(dim3((unsigned)g), dim3(128), false, [&]() {
      k_neumann<2>(T, p->f_elem, p->f_kid, p->nf, p->elems, p->n_en, p->nodes, p->traction, p->has_dir, p->dir[0], p->dir[1], p->dir[2], p->rhs);
    });
  else
    simt::launch(dim3((unsigned)g), dim3(128), false, [&]() {
      k_neumann<3>(T, p->f_elem, p->f_kid, p->nf, p->elems, p->n_en, p->nodes, p->traction, p->has_dir, p->dir[0], p->dir[1], p->dir[2], p->rhs);
    });
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Device partitioner (femcy_b200/csrc/partition.cu): the SAME orchestration (partition_build) over an emulation backend
struct EmuPartBackend;
#define PART_LAUNCH(be, n, kernel, ...)                                                              \
  do {                                                                                              \
    if ((n) > 0) simt::launch(dim3((be).grid(n)), dim3(256), false, [&]() { kernel(__VA_ARGS__); }); \
  } while (0)
#include "../../femcy_b200/csrc/partition_kernels.cuh"

struct EmuPartBackend {
  std::vector<void*> owned;
  ~EmuPartBackend() { for (void* p : owned) free(p); }
  bool ok() const { return true; }
  unsigned grid(int64_t n) const { int64_t g = cdiv(n > 0 ? n : 1, 256); return (unsigned)(g > 6 ? 6 : g); }
  template <class T> T* alloc(int64_t n) {
    void* p = calloc((size_t)(n > 0 ? n : 1), sizeof(T));
    owned.push_back(p);
    return (T*)p;
  }
  template <class T> T read(const T* p) { return *p; }
  template <class T> void upload(T* dst, const T* src, int64_t n) { if (n > 0) memcpy(dst, src, sizeof(T) * (size_t)n); }
  void sort_pairs(uint64_t* kin, uint64_t* kout, uint32_t* vin, uint32_t* vout, int64_t n) {
    std::vector<int64_t> o((size_t)n);
    std::iota(o.begin(), o.end(), 0);
    std::stable_sort(o.begin(), o.end(), [&](int64_t a, int64_t b) { return kin[a] < kin[b]; });
    for (int64_t t = 0; t < n; ++t) { kout[t] = kin[o[t]]; vout[t] = vin[o[t]]; }
  }
  void sort_keys(uint64_t* kin, uint64_t* kout, int64_t n) {
    for (int64_t t = 0; t < n; ++t) kout[t] = kin[t];
    std::sort(kout, kout + n);
  }
  void exclusive_sum(const int32_t* in, int32_t* out, int64_t n) {
    int32_t run = 0;
    for (int64_t t = 0; t < n; ++t) { out[t] = run; run += in[t]; }
  }
};

struct EmuPartition {
  int dm; int64_t nn; const double* nodes; int64_t ne; int n_en; const int32_t* elems; int rank, nranks, axis; const int64_t* bounds;
  int64_t sizes[6];
  // outputs (capacity: nn / ne / ne*n_en / nn*dm)
  int32_t* owner; int64_t* elem_ids; unsigned char* primary; int64_t* l2g; int32_t* loc_elems; double* loc_nodes;
  int32_t* peers; int64_t* send_ptr; int32_t* send_nodes; int64_t* recv_ptr; int32_t* recv_nodes;
};

extern "C" int emu_partition(EmuPartition* p) {
  EmuPartBackend be;
  PartitionResult R;
  int rc = partition_build(be, p->dm, p->nn, p->nodes, p->ne, p->n_en, p->elems, p->rank, p->nranks, p->axis, p->bounds, R);
  if (rc) return rc;
  p->sizes[0] = R.n_own; p->sizes[1] = R.n_local; p->sizes[2] = R.ne_local; p->sizes[3] = R.npeers;
  p->sizes[4] = R.send_ptr[R.npeers]; p->sizes[5] = R.recv_ptr[R.npeers];
  memcpy(p->owner, R.owner, sizeof(int32_t) * (size_t)R.nn);
  memcpy(p->elem_ids, R.elem_ids, sizeof(int64_t) * (size_t)R.ne_local);
  memcpy(p->primary, R.primary, (size_t)R.ne_local);
  memcpy(p->l2g, R.l2g, sizeof(int64_t) * (size_t)R.n_local);
  memcpy(p->loc_elems, R.loc_elems, sizeof(int32_t) * (size_t)(R.ne_local * p->n_en));
  memcpy(p->loc_nodes, R.loc_nodes, sizeof(double) * (size_t)(R.n_local * p->dm));
  memcpy(p->send_nodes, R.send_nodes, sizeof(int32_t) * (size_t)R.send_ptr[R.npeers]);
  memcpy(p->recv_nodes, R.recv_nodes, sizeof(int32_t) * (size_t)R.recv_ptr[R.npeers]);
  for (int k = 0; k < R.npeers; ++k) p->peers[k] = R.peers[k];
  for (int k = 0; k <= R.npeers; ++k) { p->send_ptr[k] = R.send_ptr[k]; p->recv_ptr[k] = R.recv_ptr[k]; }
  return 0;
}
