// Device code of the assembly path (rows a2 + a3): shape-function gradients / volumes and the stiffness
// assembly kernels.  Kept in a header so that the host launch code (assembly.cu) and the CPU SIMT emulation
// used by the not-gpu kernel-logic tests (tests/simt, test infrastructure only) compile the same source.
//
// Reference kernels replaced:
//   System_of_equations.get_dsdx_and_vol        /root/reference/stiffnessMtrx.py:132-150
//   System_of_equations.assemble_stiffnessMtrx  /root/reference/stiffnessMtrx.py:161-186
//   System_of_equations.sparseMatrix_get_j      /root/reference/stiffnessMtrx.py:414-420  (row scan -> precomputed slot)
#pragma once
#include "device_compat.cuh"
#include "kernel_types.cuh"
#include "elem_math.cuh"

// ---------------------------------------------------------------------------------------------
// geometry at all Gauss points of one element on the configuration X + u
template <int DM, int NEN>
__device__ __forceinline__ void load_current_coords(const double* __restrict__ nodes, const double* __restrict__ dof,
                                                    const int32_t* __restrict__ conn, double (&x)[NEN][DM]) {
#pragma unroll
  for (int a = 0; a < NEN; ++a) {
    int64_t n = conn[a];
#pragma unroll
    for (int i = 0; i < DM; ++i) x[a][i] = nodes[n * DM + i] + dof[n * DM + i];
  }
}

template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(128)
k_dsdx_vol(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes, const double* __restrict__ dof,
           const int32_t* __restrict__ elems, int64_t ne, double* __restrict__ dsdx, double* __restrict__ vol) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int32_t conn[NEN];
#pragma unroll
  for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
  double x[NEN][DM];
  load_current_coords<DM, NEN>(nodes, dof, conn, x);
#pragma unroll 1
  for (int gp = 0; gp < NGP; ++gp) {
    double g[NEN][DM];
    double det = shape_gradients<DM, NEN>(x, &tab.dN[gp * NEN * DM], g);
    vol[e * NGP + gp] = det * tab.w[gp];
    if (dsdx) {
      double* o = dsdx + (e * NGP + gp) * (NEN * DM);
#pragma unroll
      for (int a = 0; a < NEN; ++a)
#pragma unroll
        for (int j = 0; j < DM; ++j) o[a * DM + j] = g[a][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// scatter assembly: thread per element
template <int DM, int NEN, int NGP, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_assemble_scatter(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                   const double* __restrict__ dof, const int32_t* __restrict__ elems,
                   const int32_t* __restrict__ elem_slot, int64_t ne, double* __restrict__ val) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int32_t conn[NEN];
#pragma unroll
  for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
  double x[NEN][DM];
  load_current_coords<DM, NEN>(nodes, dof, conn, x);

  double g[NGP][NEN][DM];
  double vol[NGP];
#pragma unroll
  for (int gp = 0; gp < NGP; ++gp) vol[gp] = shape_gradients<DM, NEN>(x, &tab.dN[gp * NEN * DM], g[gp]) * tab.w[gp];

  const int32_t* slots = elem_slot + e * (NEN * NEN);
  // big elements keep the pair loops rolled (g is then indexed dynamically -> local memory)
  constexpr bool ROLL = (NEN * NGP > 16);
#pragma unroll(ROLL ? 1 : NEN)
  for (int b = 0; b < NEN; ++b) {
    double T[NGP][NV][DM];
#pragma unroll
    for (int gp = 0; gp < NGP; ++gp) C_times_B<DM>(tab.C, g[gp][b], T[gp]);
#pragma unroll(ROLL ? 1 : NEN)
    for (int a = 0; a < NEN; ++a) {
      int32_t slot = slots[a * NEN + b];
      if (slot < 0) continue;  // row node owned by another rank
      double acc[DM][DM];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
#pragma unroll
      for (int gp = 0; gp < NGP; ++gp) Bt_times_T_acc<DM>(g[gp][a], T[gp], vol[gp], acc);
      double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + (slot & 31);
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) atomicAdd(dst + ((i * DM + j) << 5), acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// scatter assembly for big elements (C3D10, CPS8/CPE8): one WARP per element.
//   stage 1  lanes < NEN load the element's nodes (X + u) into shared memory
//   stage 2  lanes < NGP invert the Jacobian of their Gauss point; then the NGP*NEN (gp, node) pairs
//            are spread over the lanes to form grad N -> shared memory
//   stage 3  register tiling of K_e: lane = (column node b, row group a0); per Gauss point the lane
//            forms T = C.B_b once in registers and applies it to its <= APL row nodes a = a0, a0+G, ...
//            (3 + 3*APL shared loads per Gauss point instead of 21 per node pair), then DM*DM atomics per
//            pair into the precomputed slot.
// (The thread-per-element kernel needs NGP*NEN*DM gradients live per thread: 120 doubles for C3D10.)
template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(128, 4)
k_assemble_scatter_warp(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                        const double* __restrict__ dof, const int32_t* __restrict__ elems,
                        const int32_t* __restrict__ elem_slot, int64_t ne, double* __restrict__ val, int64_t chunk) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int WPB = 4;                   // warps per block
  constexpr int G = 32 / NEN;              // row groups per warp (3 for NEN=10, 4 for NEN=8)
  constexpr int APL = (NEN + G - 1) / G;   // row nodes per lane
  __shared__ double xs[WPB][NEN][DM];
  __shared__ double Ji_s[WPB][NGP][DM][DM];
  __shared__ double vol_s[WPB][NGP];
  __shared__ double g_s[WPB][NGP][NEN][DM];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t nwarps = (int64_t)gridDim.x * WPB;
  const int b = lane % NEN, a0 = lane / NEN;
  const bool active = lane < G * NEN;
  // chunk == 0: grid-stride (concurrently running warps work on CONSECUTIVE elements, which share nodes and collide
  // on the same K entries).  chunk > 0 (experimental variant 4): warp g owns the contiguous range
  // [g*chunk, (g+1)*chunk) -- concurrent warps are `chunk` elements apart, consecutive elements of one warp reuse
  // their nodes from L1.
  const int64_t gwarp = blockIdx.x * (int64_t)WPB + w;
  const int64_t e_beg = chunk > 0 ? gwarp * chunk : gwarp;
  const int64_t e_end = chunk > 0 ? ((gwarp + 1) * chunk < ne ? (gwarp + 1) * chunk : ne) : ne;
  const int64_t e_step = chunk > 0 ? 1 : nwarps;
  for (int64_t e = e_beg; e < e_end; e += e_step) {
    if (lane < NEN) {
      int64_t n = elems[e * NEN + lane];
#pragma unroll
      for (int i = 0; i < DM; ++i) xs[w][lane][i] = nodes[n * DM + i] + dof[n * DM + i];
    }
    __syncwarp();
    if (lane < NGP) {
      const double* dN = &tab.dN[lane * NEN * DM];
      double J[DM][DM], Ji[DM][DM];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int k = 0; k < DM; ++k) {
          double sacc = 0.0;
#pragma unroll
          for (int a = 0; a < NEN; ++a) sacc += xs[w][a][i] * dN[a * DM + k];
          J[i][k] = sacc;
        }
      double det = inv_dm<DM>(J, Ji);
      vol_s[w][lane] = det * tab.w[lane];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int k = 0; k < DM; ++k) Ji_s[w][lane][i][k] = Ji[i][k];
    }
    __syncwarp();
    for (int p = lane; p < NGP * NEN; p += 32) {
      int gp = p / NEN, a = p - gp * NEN;
      const double* dN = &tab.dN[(gp * NEN + a) * DM];
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < DM; ++k) sacc += dN[k] * Ji_s[w][gp][k][j];
        g_s[w][gp][a][j] = sacc;
      }
    }
    __syncwarp();
    if (active) {
      double acc[APL][DM][DM];
#pragma unroll
      for (int m = 0; m < APL; ++m)
#pragma unroll
        for (int i = 0; i < DM; ++i)
#pragma unroll
          for (int j = 0; j < DM; ++j) acc[m][i][j] = 0.0;
#pragma unroll 1
      for (int gp = 0; gp < NGP; ++gp) {
        double gb[DM], T[NV][DM];
#pragma unroll
        for (int j = 0; j < DM; ++j) gb[j] = g_s[w][gp][b][j];
        C_times_B<DM>(tab.C, gb, T);
        double v = vol_s[w][gp];
#pragma unroll
        for (int m = 0; m < APL; ++m) {
          int a = a0 + m * G;
          if (a < NEN) {
            double ga[DM];
#pragma unroll
            for (int j = 0; j < DM; ++j) ga[j] = g_s[w][gp][a][j];
            Bt_times_T_acc<DM>(ga, T, v, acc[m]);
          }
        }
      }
      const int32_t* slots = elem_slot + e * (NEN * NEN);
#pragma unroll
      for (int m = 0; m < APL; ++m) {
        int a = a0 + m * G;
        if (a < NEN) {
          int32_t slot = slots[a * NEN + b];
          if (slot >= 0) {
            double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + (slot & 31);
#pragma unroll
            for (int i = 0; i < DM; ++i)
#pragma unroll
              for (int j = 0; j < DM; ++j) atomicAdd(dst + ((i * DM + j) << 5), acc[m][i][j]);
          }
        }
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// scatter assembly for big elements, software-pipelined and symmetric (experimental variant 19; unmeasured).
// Same warp-per-element decomposition as k_assemble_scatter_warp, reworked where that kernel serialises
// (measured r1z: 8.8 ms for 1 M C3D10 = ~37 k clocks per element per warp at 16 warps/SM, against ~1.4 k clocks of
// FP64 pipe time and ~0.5 k clocks of atomic issue per element -- it waits, it does not compute):
//   * persistent warps with a two-deep asynchronous pipeline: while element e is computed, the coordinate / displacement
//     rows and the slot row of element e+stride travel global -> shared memory as cp.async copies (LDGSTS: no
//     registers, no scoreboard wait until the consuming iteration), and the node ids of element e+2*stride are loaded
//     into one register -- no load of an iteration feeds an address or an operand of the same iteration;
//   * the NGP*DM*DM Jacobian entries are spread over the lanes (NEN-term dot products) instead of NGP lanes
//     forming DM*DM entries each; the natural derivatives sit in shared memory (lane-dependent indices into the
//     kernel-parameter bank serialise);
//   * only the NEN*(NEN+1)/2 node pairs a <= b are evaluated (K_e is symmetric for a symmetric tangent -- checked by the
//     host, which otherwise takes variant 1): <= 2 blocks = 18 accumulators per lane instead of 4 blocks = 36;
//     block (b,a) is scattered as the transpose of (a,b).
// Results differ from variant 1 by rounding only (block (b,a) is the exact transpose of (a,b)).
template <int DM, int NEN, int NGP, bool CUBIC>
__global__ void __launch_bounds__(128, CUBIC ? 6 : 4)
k_assemble_scatter_pairs(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                         const double* __restrict__ dof, const int32_t* __restrict__ elems,
                         const int32_t* __restrict__ elem_slot, int64_t ne, double* __restrict__ val) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int WPB = 4;                        // warps per block
  constexpr int NP = NEN * (NEN + 1) / 2;       // node pairs a <= b
  constexpr int PPL = (NP + 31) / 32;           // pairs per lane
  constexpr int NSL = NEN * NEN;                // slots per element
  constexpr int CW = (NSL % 4 == 0) ? 16 : 4;   // bytes per slot-row copy (rows are NSL*4 bytes apart)
  constexpr int NCP = NSL * 4 / CW;             // copies per slot row
  static_assert(NEN <= 32 && NGP <= 32, "one lane per node / Gauss point");
  __shared__ double dN_s[NGP * NEN * DM];
  alignas(16) __shared__ int32_t slot_s[WPB][2][NSL];       // double-buffered pipeline stages
  __shared__ double xn_s[WPB][2][NEN][DM];
  __shared__ double xu_s[WPB][2][NEN][DM];
  __shared__ double xs[WPB][NEN][DM];
  __shared__ double J_s[WPB][NGP][DM][DM];
  __shared__ double vol_s[WPB][NGP];
  __shared__ double g_s[WPB][NGP][NEN][DM];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < NGP * NEN * DM; t += blockDim.x) dN_s[t] = tab.dN[t];
  __syncthreads();
  // this lane's node pairs (row-major upper triangle)
  int pa[PPL], pb[PPL];
#pragma unroll
  for (int q = 0; q < PPL; ++q) {
    int p = lane + 32 * q, a = 0;
    if (p < NP) {
      while (p >= NEN - a) { p -= NEN - a; ++a; }
      pa[q] = a; pb[q] = a + p;
    } else {
      pa[q] = -1; pb[q] = -1;
    }
  }
  double cp = 0.0, cq = 0.0, cr = 0.0;
  if constexpr (CUBIC) { cp = tab.C[0]; cq = tab.C[1]; cr = tab.C[NV * NV - 1]; }

  const int64_t stride = (int64_t)gridDim.x * WPB;
  int64_t e = blockIdx.x * (int64_t)WPB + w;
  // stage the rows of element `el` (node id of this lane: n) into pipeline buffer `buf`
  auto stage = [&](int64_t el, int64_t n, int buf) {
    if (lane < NEN) {
#pragma unroll
      for (int i = 0; i < DM; ++i) {
        femcy_cp_async<8>(&xn_s[w][buf][lane][i], nodes + n * DM + i);
        femcy_cp_async<8>(&xu_s[w][buf][lane][i], dof + n * DM + i);
      }
    }
    const char* srow = reinterpret_cast<const char*>(elem_slot + el * NSL);
    for (int c = lane; c < NCP; c += 32)
      femcy_cp_async<CW>(reinterpret_cast<char*>(&slot_s[w][buf][0]) + c * CW, srow + c * CW);
  };
  int32_t nid_next = 0;
  if (e < ne) {
    int64_t n0 = (lane < NEN) ? elems[e * NEN + lane] : 0;
    stage(e, n0, 0);
    if (lane < NEN && e + stride < ne) nid_next = elems[(e + stride) * NEN + lane];
  }
  femcy_cp_async_commit();
  int buf = 0;
  for (; e < ne; e += stride, buf ^= 1) {
    // ---- next element's rows start travelling; the id after that goes into a register ----
    const int64_t e1 = e + stride, e2 = e + 2 * stride;
    if (e1 < ne) {
      stage(e1, nid_next, buf ^ 1);
      if (lane < NEN && e2 < ne) nid_next = elems[e2 * NEN + lane];
    }
    femcy_cp_async_commit();
    femcy_cp_async_wait<1>();                   // everything but the group just committed has landed: element e is here
    __syncwarp();
    if (lane < NEN) {
#pragma unroll
      for (int i = 0; i < DM; ++i) xs[w][lane][i] = xn_s[w][buf][lane][i] + xu_s[w][buf][lane][i];
    }
    __syncwarp();
    // ---- Jacobians: one (gp, i, k) entry per lane, summed over the nodes in the reference's order ----
    for (int t = lane; t < NGP * DM2; t += 32) {
      const int gp = t / DM2, ik = t - gp * DM2, i = ik / DM, k = ik - i * DM;
      double sacc = 0.0;
#pragma unroll
      for (int a = 0; a < NEN; ++a) sacc += xs[w][a][i] * dN_s[(gp * NEN + a) * DM + k];
      J_s[w][gp][i][k] = sacc;
    }
    __syncwarp();
    if (lane < NGP) {
      double J[DM][DM], Ji[DM][DM];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int k = 0; k < DM; ++k) J[i][k] = J_s[w][lane][i][k];
      double det = inv_dm<DM>(J, Ji);
      vol_s[w][lane] = det * tab.w[lane];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int k = 0; k < DM; ++k) J_s[w][lane][i][k] = Ji[i][k];      // in place: J^-1
    }
    __syncwarp();
    for (int p = lane; p < NGP * NEN; p += 32) {
      const int gp = p / NEN, a = p - gp * NEN;
      const double* dN = &dN_s[(gp * NEN + a) * DM];
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < DM; ++k) sacc += dN[k] * J_s[w][gp][k][j];
        g_s[w][gp][a][j] = sacc;
      }
    }
    __syncwarp();
    // ---- K_e blocks of this lane's pairs ----
    double acc[PPL][DM][DM];
#pragma unroll
    for (int q = 0; q < PPL; ++q)
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) acc[q][i][j] = 0.0;
#pragma unroll 1
    for (int gp = 0; gp < NGP; ++gp) {
      const double v = vol_s[w][gp];
#pragma unroll
      for (int q = 0; q < PPL; ++q) {
        if (pa[q] < 0) continue;
        double ga[DM], gb[DM];
#pragma unroll
        for (int j = 0; j < DM; ++j) { ga[j] = g_s[w][gp][pa[q]][j]; gb[j] = g_s[w][gp][pb[q]][j]; }
        if constexpr (CUBIC) {
          block_cubic_acc<DM>(cp, cq, cr, ga, gb, v, acc[q]);
        } else {
          double T[NV][DM];
          C_times_B<DM>(tab.C, gb, T);
          Bt_times_T_acc<DM>(ga, T, v, acc[q]);
        }
      }
    }
    // ---- scatter: block (a,b) and, off the diagonal, its transpose into (b,a) ----
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
      if (pa[q] < 0) continue;
      const int32_t s_ab = slot_s[w][buf][pa[q] * NEN + pb[q]];
      if (s_ab >= 0) {
        double* dst = val + (((int64_t)(s_ab >> 5) * DM2) << 5) + (s_ab & 31);
#pragma unroll
        for (int i = 0; i < DM; ++i)
#pragma unroll
          for (int j = 0; j < DM; ++j) atomicAdd(dst + ((i * DM + j) << 5), acc[q][i][j]);
      }
      if (pa[q] != pb[q]) {
        const int32_t s_ba = slot_s[w][buf][pb[q] * NEN + pa[q]];
        if (s_ba >= 0) {
          double* dst = val + (((int64_t)(s_ba >> 5) * DM2) << 5) + (s_ba & 31);
#pragma unroll
          for (int i = 0; i < DM; ++i)
#pragma unroll
            for (int j = 0; j < DM; ++j) atomicAdd(dst + ((j * DM + i) << 5), acc[q][i][j]);
        }
      }
    }
    __syncwarp();                                // buffer `buf` is free for the stage call after next
  }
  femcy_cp_async_wait<0>();
}

// host + device: is the row-major [NV][NV] tangent symmetric (K_e symmetric => the pair kernel applies)?
static inline bool tangent_is_symmetric(const double* C, int dm) {
  const int nv = (dm == 2) ? 3 : 6;
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < i; ++j)
      if (C[i * nv + j] != C[j * nv + i]) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------
// gather assembly (single Gauss point): pass 1 = per-element record [g[NEN][DM], vol]
// (measured on B200, profiles/r1_notes.md: padding the record to a 128 B line and walking the element
//  list 2-4 entries at a time raised the register count 46 -> 72-118 and made pass 2 1.8-2x SLOWER;
//  the plain loop below is the fastest of the variants tried.)
template <int DM, int NEN>
struct GeoRec { static constexpr int N = NEN * DM + 1; };

template <int DM, int NEN>
__global__ void __launch_bounds__(256)
k_elem_geometry(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                const double* __restrict__ dof, const int32_t* __restrict__ elems, int64_t ne,
                double* __restrict__ egeo, double* __restrict__ vol_out) {
  constexpr int REC = GeoRec<DM, NEN>::N;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int32_t conn[NEN];
#pragma unroll
  for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
  double x[NEN][DM], g[NEN][DM];
  load_current_coords<DM, NEN>(nodes, dof, conn, x);
  double v = shape_gradients<DM, NEN>(x, tab.dN, g) * tab.w[0];
  double* o = egeo + e * REC;
#pragma unroll
  for (int a = 0; a < NEN; ++a)
#pragma unroll
    for (int j = 0; j < DM; ++j) o[a * DM + j] = g[a][j];
  o[NEN * DM] = v;
  vol_out[e] = v;
}

// pass 1 with coalesced stores (variant 11): same 13-double records, but staged in shared memory (pitch 13 doubles is odd,
// so the 8-byte stores of consecutive threads fall on different banks) and copied out as one contiguous chunk -- the
// thread-per-element version above issues 13 stores of 32 x 8 B at a 104 B stride (ncu r1: 1.3 GB in 0.61 ms = 2.2 TB/s).
template <int DM, int NEN>
__global__ void __launch_bounds__(128)
k_elem_geometry_s(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                  const double* __restrict__ dof, const int32_t* __restrict__ elems, int64_t ne,
                  double* __restrict__ egeo, double* __restrict__ vol_out) {
  constexpr int REC = GeoRec<DM, NEN>::N;
  constexpr int TPB = 128;
  __shared__ double tile[TPB * REC];
  const int t = threadIdx.x;
  const int64_t e0 = blockIdx.x * (int64_t)TPB;
  const int64_t e = e0 + t;
  if (e < ne) {
    int32_t conn[NEN];
#pragma unroll
    for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
    double x[NEN][DM], g[NEN][DM];
    load_current_coords<DM, NEN>(nodes, dof, conn, x);
    double v = shape_gradients<DM, NEN>(x, tab.dN, g) * tab.w[0];
    double* o = tile + t * REC;
#pragma unroll
    for (int a = 0; a < NEN; ++a)
#pragma unroll
      for (int j = 0; j < DM; ++j) o[a * DM + j] = g[a][j];
    o[NEN * DM] = v;
    vol_out[e] = v;
  }
  __syncthreads();
  const int64_t rem = ne - e0;
  const int nel = rem < TPB ? (int)rem : TPB;
  double* out = egeo + e0 * REC;
  for (int i = t; i < nel * REC; i += TPB) out[i] = tile[i];
}

// pass 1 with a bulk copy-out (variant 21): as k_elem_geometry_s, but the block's 128 records -- contiguous in global
// memory -- leave shared memory as ONE asynchronous bulk copy issued by one thread (cp.async.bulk.global.shared::cta, the
// TMA engine without a tensor map): no per-thread global stores at all.  The last block (or an odd record count, whose
// byte size is not a multiple of 16) takes the plain loop.
template <int DM, int NEN>
__global__ void __launch_bounds__(128)
k_elem_geometry_b(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                  const double* __restrict__ dof, const int32_t* __restrict__ elems, int64_t ne,
                  double* __restrict__ egeo, double* __restrict__ vol_out) {
  constexpr int REC = GeoRec<DM, NEN>::N;
  constexpr int TPB = 128;
  static_assert((TPB * REC * 8) % 16 == 0, "a full tile is a whole number of 16-byte units");
  alignas(128) __shared__ double tile[TPB * REC];
  const int t = threadIdx.x;
  const int64_t e0 = blockIdx.x * (int64_t)TPB;
  const int64_t e = e0 + t;
  if (e < ne) {
    int32_t conn[NEN];
#pragma unroll
    for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
    double x[NEN][DM], g[NEN][DM];
    load_current_coords<DM, NEN>(nodes, dof, conn, x);
    double v = shape_gradients<DM, NEN>(x, tab.dN, g) * tab.w[0];
    double* o = tile + t * REC;
#pragma unroll
    for (int a = 0; a < NEN; ++a)
#pragma unroll
      for (int j = 0; j < DM; ++j) o[a * DM + j] = g[a][j];
    o[NEN * DM] = v;
    vol_out[e] = v;
  }
  femcy_fence_async_smem();          // this thread's record -> visible to the TMA engine
  __syncthreads();
  const int64_t rem = ne - e0;
  const int nel = rem < TPB ? (int)rem : TPB;
  double* out = egeo + e0 * REC;
  const unsigned bytes = (unsigned)(nel * REC * 8);
  if ((bytes & 15u) == 0u) {
    if (t == 0) femcy_bulk_store(out, tile, bytes);
  } else {
    for (int i = t; i < nel * REC; i += TPB) out[i] = tile[i];
  }
}

// pass 2: block (32 lanes, KB k-rows): one thread per stored block slot sums its element list
template <int DM, int NEN>
__global__ void __launch_bounds__(256)
k_assemble_gather(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr, int64_t nslice,
                  const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                  const uint32_t* __restrict__ ent_list, const double* __restrict__ egeo, double* __restrict__ val,
                  int kgroups) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int REC = GeoRec<DM, NEN>::N;
  constexpr int P = NEN * NEN;
  // kgroups == 0: 2-D grid (slice, k-group): all slices of k-group 0 run first, then k-group 1, ... -- the mesh is
  // swept once per k-group.  kgroups > 0 (experimental variant 5): 1-D grid, the k-groups of a slice are adjacent in
  // launch order, so the element records of a slice's neighbourhood are fetched from HBM once and re-read from L2.
  int64_t s = kgroups > 0 ? (int64_t)(blockIdx.x / (unsigned)kgroups) : (int64_t)blockIdx.x;
  int lane = threadIdx.x;
  int k = (kgroups > 0 ? (int)(blockIdx.x % (unsigned)kgroups) : (int)blockIdx.y) * blockDim.y + threadIdx.y;
  int base = slice_ptr[s];
  int w = (slice_ptr[s + 1] - base) >> 5;
  if (k >= w) return;
  int slot = base + (k << 5) + lane;
  int beg = slot_beg[slot], end = slot_end[slot];
  double acc[DM][DM];
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
  for (int t = beg; t < end; ++t) {
    uint32_t id = ent_list[t];
    uint32_t e = id / P;
    int p = (int)(id - e * P);
    int a = p / NEN, b = p - a * NEN;
    const double* rec = egeo + (int64_t)e * REC;
    double ga[DM], gb[DM];
#pragma unroll
    for (int j = 0; j < DM; ++j) { ga[j] = rec[a * DM + j]; gb[j] = rec[b * DM + j]; }
    double v = rec[NEN * DM];
    double T[NV][DM];
    C_times_B<DM>(tab.C, gb, T);
    Bt_times_T_acc<DM>(ga, T, v, acc);
  }
  double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + lane;
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int j = 0; j < DM; ++j) dst[(i * DM + j) << 5] = acc[i][j];
}

// gather assembly for elements with several Gauss points (EXPERIMENTAL, opt-in variant 2 on C3D10 etc.; not
// yet measured on hardware -- written for round 2).  Same scheme as k_assemble_gather, but the per-element
// record is the reference's own pair of fields dsdx[e][gp][a][:] and vol[e][gp] (stiffnessMtrx.py:59-61),
// produced by k_dsdx_vol; no atomics, bit-reproducible.
template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(256)
k_assemble_gather_mgp(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr, int64_t nslice,
                      const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                      const uint32_t* __restrict__ ent_list, const double* __restrict__ dsdx,
                      const double* __restrict__ vol, double* __restrict__ val) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int P = NEN * NEN;
  int64_t s = blockIdx.x;
  int lane = threadIdx.x;
  int k = blockIdx.y * blockDim.y + threadIdx.y;
  int base = slice_ptr[s];
  int w = (slice_ptr[s + 1] - base) >> 5;
  if (k >= w) return;
  int slot = base + (k << 5) + lane;
  int beg = slot_beg[slot], end = slot_end[slot];
  double acc[DM][DM];
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
  for (int t = beg; t < end; ++t) {
    uint32_t id = ent_list[t];
    uint32_t e = id / P;
    int p = (int)(id - e * P);
    int a = p / NEN, b = p - a * NEN;
    const double* ge = dsdx + (int64_t)e * (NGP * NEN * DM);
    const double* ve = vol + (int64_t)e * NGP;
#pragma unroll
    for (int gp = 0; gp < NGP; ++gp) {
      double ga[DM], gb[DM];
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        ga[j] = ge[(gp * NEN + a) * DM + j];
        gb[j] = ge[(gp * NEN + b) * DM + j];
      }
      double T[NV][DM];
      C_times_B<DM>(tab.C, gb, T);
      Bt_times_T_acc<DM>(ga, T, ve[gp], acc);
    }
  }
  double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + lane;
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int j = 0; j < DM; ++j) dst[(i * DM + j) << 5] = acc[i][j];
}

// ---------------------------------------------------------------------------------------------
// "rows" assembly (EXPERIMENTAL, opt-in variant 6; verified on the SIMT emulation, not yet measured on hardware):
// owner-computes, no atomics, no zero-fill, bit-reproducible.
//   pass 1  k_elem_geometry4: per-element record, node-major, one 32 B sector per (node, Gauss point):
//           rec[e][a][gp] = (dN_a/dx, dN_a/dy, dN_a/dz | 0, vol_gp)
//   pass 2  k_assemble_rows: one block per R consecutive rows of a 32-row slice (R = 32 for elements with <= 4
//           nodes, 8 for the big ones).  The rows are accumulated in SHARED memory (w*dm2*R doubles: ~35 KB for
//           C3D4 with R = 32, ~42 KB for C3D10 with R = 8) and written once (the slice's values are one contiguous
//           chunk of `val`; R = 32 writes whole 256 B planes).  Work split: NEN lanes per row -- lane (row, b) walks
//           the row's node->element incidence list and for incidence (e, a) forms the block
//           K_e[a][b] = sum_gp vol B_a^T C B_b and adds it to the row's shared-memory block
//           k = (elem_slot[e][a][b] - slice base) / 32.  Lanes of one row handle the NEN different column nodes of
//           the same element in the same iteration => distinct k, no conflict; rows are private to their lanes =>
//           no atomics.  The NEN lanes of a row together read the element's whole record exactly once.
// Against the per-block gather (k_assemble_gather) this moves 4-5x fewer L2->SM bytes (each element record is read
// once per incident row instead of once per stored block) and needs no per-block element lists (647 MB for cfg 4).
template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(128)
k_elem_geometry4(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                 const double* __restrict__ dof, const int32_t* __restrict__ elems, int64_t ne,
                 double* __restrict__ rec, double* __restrict__ vol_out) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int32_t conn[NEN];
#pragma unroll
  for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
  double x[NEN][DM];
  load_current_coords<DM, NEN>(nodes, dof, conn, x);
  double2* o = reinterpret_cast<double2*>(rec + e * (NEN * NGP * 4));
#pragma unroll 1
  for (int gp = 0; gp < NGP; ++gp) {
    double g[NEN][DM];
    double v = shape_gradients<DM, NEN>(x, &tab.dN[gp * NEN * DM], g) * tab.w[gp];
#pragma unroll
    for (int a = 0; a < NEN; ++a) {
      double2 lo, hi;
      lo.x = g[a][0]; lo.y = g[a][1];
      hi.x = (DM == 3) ? g[a][DM - 1] : 0.0; hi.y = v;
      o[(a * NGP + gp) * 2] = lo;
      o[(a * NGP + gp) * 2 + 1] = hi;
    }
    vol_out[e * NGP + gp] = v;
  }
}

// pass 1 with coalesced stores (variants 7-9): the thread-per-element version above writes each 32 B sector of its
// record with its own store instruction at a stride of the record size (ncu r1: k_elem_geometry moves 1.3 GB in
// 0.61 ms = 2.2 TB/s).  Here the records of a block are staged in shared memory (pitch = record + 16 B, which keeps
// the 16-byte stores of consecutive threads on different bank groups) and copied out as one contiguous chunk.
template <int NEN, int NGP>
struct Geo4Cfg {
  static constexpr int CH = NEN * NGP * 2;                 // 16-byte chunks per element record
  static constexpr int TPB = (CH <= 8) ? 128 : 32;         // elements (= threads) per block
};

template <int DM, int NEN, int NGP>
__global__ void __launch_bounds__(Geo4Cfg<NEN, NGP>::TPB)
k_elem_geometry4s(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                  const double* __restrict__ dof, const int32_t* __restrict__ elems, int64_t ne,
                  double* __restrict__ rec, double* __restrict__ vol_out) {
  constexpr int CH = Geo4Cfg<NEN, NGP>::CH, TPB = Geo4Cfg<NEN, NGP>::TPB;
  __shared__ double2 tile[TPB * (CH + 1)];
  const int t = threadIdx.x;
  const int64_t e0 = blockIdx.x * (int64_t)TPB;
  const int64_t e = e0 + t;
  if (e < ne) {
    int32_t conn[NEN];
#pragma unroll
    for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
    double x[NEN][DM];
    load_current_coords<DM, NEN>(nodes, dof, conn, x);
    double2* o = tile + t * (CH + 1);
#pragma unroll 1
    for (int gp = 0; gp < NGP; ++gp) {
      double g[NEN][DM];
      double v = shape_gradients<DM, NEN>(x, &tab.dN[gp * NEN * DM], g) * tab.w[gp];
#pragma unroll
      for (int a = 0; a < NEN; ++a) {
        double2 lo, hi;
        lo.x = g[a][0]; lo.y = g[a][1];
        hi.x = (DM == 3) ? g[a][DM - 1] : 0.0; hi.y = v;
        o[(a * NGP + gp) * 2] = lo;
        o[(a * NGP + gp) * 2 + 1] = hi;
      }
      vol_out[e * NGP + gp] = v;
    }
  }
  __syncthreads();
  const int64_t rem = ne - e0;
  const int nel = rem < TPB ? (int)rem : TPB;
  double2* out = reinterpret_cast<double2*>(rec + e0 * (NEN * NGP * 4));
  for (int gi = t; gi < nel * CH; gi += TPB) {
    int el = gi / CH, c = gi - el * CH;
    out[gi] = tile[el * (CH + 1) + c];
  }
}

// pass 1 with a TMA tensor store (variant 18; C3D4: the record is exactly one 128-byte row).  The block's 128 records
// are written into a 16 KB shared-memory tile in the 128-byte-swizzled layout (16-byte chunk c of row t at chunk
// c ^ (t & 7): conflict-free for the per-thread row writes) and ONE thread hands the whole tile to the TMA unit
// (cp.async.bulk.tensor.2d, CU_TENSOR_MAP_SWIZZLE_128B un-swizzles on the way out; rows past `ne` are clipped by the
// tensor map) -- no copy-out loop, no per-thread global stores.  Under the CPU emulation the store is a plain loop.
#ifdef FEMCY_SIMT_EMU
struct FemcyTmap { double* base; int64_t rows; };
#else
#include <cuda.h>
struct alignas(64) FemcyTmap { CUtensorMap m; };
#endif

template <int DM, int NEN>
__global__ void __launch_bounds__(128)
k_elem_geometry4t(const __grid_constant__ ElemTables tab, const double* __restrict__ nodes,
                  const double* __restrict__ dof, const int32_t* __restrict__ elems, int64_t ne,
                  const __grid_constant__ FemcyTmap tm, double* __restrict__ vol_out) {
  static_assert(NEN * 2 == 8, "k_elem_geometry4t: records of exactly 128 bytes (4 nodes, one Gauss point)");
  constexpr int TPB = 128, CH = 8;
#ifdef FEMCY_SIMT_EMU
  __shared__ double2 tile[TPB * CH];
#else
  __shared__ __align__(1024) double2 tile[TPB * CH];
#endif
  const int t = threadIdx.x;
  const int64_t e0 = blockIdx.x * (int64_t)TPB;
  const int64_t e = e0 + t;
  if (e < ne) {
    int32_t conn[NEN];
#pragma unroll
    for (int a = 0; a < NEN; ++a) conn[a] = elems[e * NEN + a];
    double x[NEN][DM], g[NEN][DM];
    load_current_coords<DM, NEN>(nodes, dof, conn, x);
    double v = shape_gradients<DM, NEN>(x, tab.dN, g) * tab.w[0];
    double2* o = tile + t * CH;
    const int sw = t & 7;
#pragma unroll
    for (int a = 0; a < NEN; ++a) {
      double2 lo, hi;
      lo.x = g[a][0]; lo.y = g[a][1];
      hi.x = (DM == 3) ? g[a][DM - 1] : 0.0; hi.y = v;
      o[(2 * a) ^ sw] = lo;
      o[(2 * a + 1) ^ sw] = hi;
    }
    vol_out[e] = v;
  }
#ifdef FEMCY_SIMT_EMU
  __syncthreads();
  const int64_t rem = tm.rows - e0;
  const int nel = rem < TPB ? (int)rem : TPB;
  double2* out = reinterpret_cast<double2*>(tm.base + e0 * (NEN * 4));
  for (int gi = t; gi < nel * CH; gi += TPB) {
    int el = gi / CH, c = gi - el * CH;
    out[gi] = tile[el * CH + (c ^ (el & 7))];
  }
#else
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of the tile -> visible to the TMA unit
  __syncthreads();
  if (t == 0) {
    const uint64_t tmap = reinterpret_cast<uint64_t>(&tm);
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
    const int c0 = 0, c1 = (int)e0;
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(src) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile must outlive the read
  }
#endif
}

template <int NEN>
struct RowsCfg {
  static constexpr int R = (NEN <= 4) ? 32 : 8;            // rows per block
  static constexpr int RPW = 32 / NEN;                     // rows per warp (NEN lanes each)
  static constexpr int NW = (R + RPW - 1) / RPW;           // warps per block
  static constexpr int PITCH = R + 1;                      // shared-memory pitch of one (k, q) plane (spreads k over the banks)
};

// PF: 0 = plain loop (measured r1z: latency-bound, 24 warps/SM and two dependent loads per incidence);
//     1 / 2 = software prefetch: the incidence ids run 3 steps ahead in registers and the record / slot lines of the
//     incidence two steps ahead are prefetched into L2 (1) or L1 (2) while the current one is computed.
template <int PF>
__device__ __forceinline__ void rows_prefetch(const void* p) {
#ifndef FEMCY_SIMT_EMU
  if constexpr (PF == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  if constexpr (PF == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

//     3 = register double-buffering (single-Gauss-point elements): the record + slot of incidence j+1 are loaded into
//     registers before incidence j is computed (the kernel's residency is capped by its shared-memory accumulator at
//     24 warps/SM, so up to ~85 registers per thread cost no occupancy); CUBIC adds the cubic-form tangent fast path.
template <int DM, int NEN, int NGP, int PF, bool CUBIC = false>
__global__ void __launch_bounds__(RowsCfg<NEN>::NW * 32)
k_assemble_rows(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr, int64_t nrows,
                const int32_t* __restrict__ inc_ptr, const uint32_t* __restrict__ inc_list,
                const int32_t* __restrict__ elem_slot, const double* __restrict__ rec, double* __restrict__ val,
                const int32_t* __restrict__ rowof) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int R = RowsCfg<NEN>::R, RPW = RowsCfg<NEN>::RPW, PITCH = RowsCfg<NEN>::PITCH;
  constexpr int SUB = 32 / R;                              // blocks per slice
#ifdef FEMCY_SIMT_EMU
  double* acc_s = static_cast<double*>(simt::dyn_smem());      // exactly the bytes the launch asked for
#else
  extern __shared__ double acc_s[];
#endif
  const int64_t s = blockIdx.x / SUB;
  const int r0 = (int)(blockIdx.x % SUB) * R;              // first row of this block within the slice
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) >> 5;
  const int nplane = w * DM2;
  for (int i = threadIdx.x; i < nplane * PITCH; i += blockDim.x) acc_s[i] = 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rw = lane / NEN, b = lane - rw * NEN;
  const int row_b = warp * RPW + rw;                       // row within the block
  const int64_t pos = s * 32 + r0 + row_b;                // position in the (sigma-sorted) row order
  const bool active = (rw < RPW) && (row_b < R) && (pos < nrows);
  int beg = 0, end = 0;
  if (active) {
    const int64_t row = rowof ? (int64_t)rowof[pos] : pos;
    beg = inc_ptr[row]; end = inc_ptr[row + 1];
  }
  int nmax = end - beg;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int other = __shfl_xor_sync(0xffffffffu, nmax, o);
    nmax = other > nmax ? other : nmax;
  }
  if constexpr (PF == 3) {
    static_assert(NGP == 1, "register double-buffering is written for single-Gauss-point elements");
    auto fetch = [&](uint32_t id, double2& alo, double2& ahi, double2& blo, double2& bhi, int& k) {
      uint32_t e = id / NEN;
      int a = (int)(id - e * NEN);
      const double2* r2 = reinterpret_cast<const double2*>(rec + (int64_t)e * (NEN * 4));
      alo = r2[a * 2]; ahi = r2[a * 2 + 1];
      blo = r2[b * 2]; bhi = r2[b * 2 + 1];
      k = (elem_slot[((int64_t)e * NEN + a) * NEN + b] - base) >> 5;
    };
    double2 a_lo = {0.0, 0.0}, a_hi = {0.0, 0.0}, b_lo = {0.0, 0.0}, b_hi = {0.0, 0.0};
    int k = 0;
    uint32_t id1 = (beg + 1 < end) ? inc_list[beg + 1] : 0u;
    if (beg < end) fetch(inc_list[beg], a_lo, a_hi, b_lo, b_hi, k);
    for (int j = 0; j < nmax; ++j) {
      const uint32_t id2 = (beg + j + 2 < end) ? inc_list[beg + j + 2] : 0u;
      double2 na_lo = {0.0, 0.0}, na_hi = {0.0, 0.0}, nb_lo = {0.0, 0.0}, nb_hi = {0.0, 0.0};
      int nk = 0;
      if (beg + j + 1 < end) fetch(id1, na_lo, na_hi, nb_lo, nb_hi, nk);      // in flight while incidence j is computed
      if (beg + j < end) {
        double ga[DM], gb[DM];
        ga[0] = a_lo.x; ga[1] = a_lo.y;
        gb[0] = b_lo.x; gb[1] = b_lo.y;
        if constexpr (DM == 3) { ga[2] = a_hi.x; gb[2] = b_hi.x; }
        double blk[DM][DM];
#pragma unroll
        for (int i = 0; i < DM; ++i)
#pragma unroll
          for (int jj = 0; jj < DM; ++jj) blk[i][jj] = 0.0;
        if constexpr (CUBIC) {
          block_cubic_acc<DM>(tab.C[0], tab.C[1], tab.C[NV * NV - 1], ga, gb, a_hi.y, blk);
        } else {
          double T[NV][DM];
          C_times_B<DM>(tab.C, gb, T);
          Bt_times_T_acc<DM>(ga, T, a_hi.y, blk);
        }
        double* dst = acc_s + (k * DM2) * PITCH + row_b;
#pragma unroll
        for (int i = 0; i < DM; ++i)
#pragma unroll
          for (int jj = 0; jj < DM; ++jj) dst[(i * DM + jj) * PITCH] += blk[i][jj];
      }
      a_lo = na_lo; a_hi = na_hi; b_lo = nb_lo; b_hi = nb_hi; k = nk; id1 = id2;
      __syncwarp();
    }
  }
  // id pipeline (PF 1, 2): idq[0] = this step's incidence, idq[1], idq[2] the next two, one more load in flight
  uint32_t idq[3] = {0u, 0u, 0u};
  if constexpr (PF == 1 || PF == 2) {
#pragma unroll
    for (int u = 0; u < 3; ++u)
      if (beg + u < end) idq[u] = inc_list[beg + u];
  }
  for (int j = 0; j < (PF == 3 ? 0 : nmax); ++j) {
    uint32_t id_new = 0u;
    if constexpr (PF == 1 || PF == 2) {
      if (beg + j + 3 < end) id_new = inc_list[beg + j + 3];
      if (beg + j + 2 < end) {
        uint32_t e2 = idq[2] / NEN;
        int a2 = (int)(idq[2] - e2 * NEN);
        const double* rp = rec + (int64_t)e2 * (NEN * NGP * 4);
        rows_prefetch<PF>(rp + (int64_t)b * (NGP * 4));          // this lane's column-node part (NGP sectors)
        if (b == 0) {
          rows_prefetch<PF>(rp + (int64_t)a2 * (NGP * 4));       // the row node's part, once per row
          rows_prefetch<PF>(elem_slot + ((int64_t)e2 * NEN + a2) * NEN);
        }
      }
    }
    if (beg + j < end) {
      uint32_t id = (PF == 1 || PF == 2) ? idq[0] : inc_list[beg + j];
      uint32_t e = id / NEN;
      int a = (int)(id - e * NEN);
      const double2* r2 = reinterpret_cast<const double2*>(rec + (int64_t)e * (NEN * NGP * 4));
      int slot = elem_slot[((int64_t)e * NEN + a) * NEN + b];
      int k = (slot - base) >> 5;
      double blk[DM][DM];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int jj = 0; jj < DM; ++jj) blk[i][jj] = 0.0;
#pragma unroll
      for (int gp = 0; gp < NGP; ++gp) {
        double2 a_lo = r2[(a * NGP + gp) * 2], a_hi = r2[(a * NGP + gp) * 2 + 1];
        double2 b_lo = r2[(b * NGP + gp) * 2], b_hi = r2[(b * NGP + gp) * 2 + 1];
        double ga[DM], gb[DM];
        ga[0] = a_lo.x; ga[1] = a_lo.y;
        gb[0] = b_lo.x; gb[1] = b_lo.y;
        if constexpr (DM == 3) { ga[2] = a_hi.x; gb[2] = b_hi.x; }
        double T[NV][DM];
        C_times_B<DM>(tab.C, gb, T);
        Bt_times_T_acc<DM>(ga, T, a_hi.y, blk);
      }
      double* dst = acc_s + (k * DM2) * PITCH + row_b;
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int jj = 0; jj < DM; ++jj) dst[(i * DM + jj) * PITCH] += blk[i][jj];
    }
    if constexpr (PF == 1 || PF == 2) { idq[0] = idq[1]; idq[1] = idq[2]; idq[2] = id_new; }
    __syncwarp();
  }
  __syncthreads();
  double* out = val + (((int64_t)(base >> 5) * DM2) << 5) + r0;
  for (int i = threadIdx.x; i < nplane * R; i += blockDim.x) {
    int kq = i / R, l = i - kq * R;
    out[(kq << 5) + l] = acc_s[kq * PITCH + l];
  }
}

// per-block gather over the node-sector records of k_elem_geometry4(s) (variant 9; any number of Gauss points).
// Against k_assemble_gather (13-double records: ~124 B of L2->SM sectors per contribution, ncu r1) a contribution
// reads exactly two 32 B sectors per Gauss point (one when a == b) with 16-byte loads.  Launched slice-major.
// V256 (variant 20): each record is fetched with ONE 256-bit load (LDG.E.256, new on sm_100) instead of two 128-bit ones:
// the gather is bound by load-instruction / sector-request issue, not by FP64 or DRAM (ncu r1).
template <int DM, int NEN, int NGP, bool CUBIC, int MINB = 0, bool V256 = false>
__global__ void __launch_bounds__(256, MINB)
k_assemble_gather4(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr,
                   const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                   const uint32_t* __restrict__ ent_list, const double* __restrict__ rec, double* __restrict__ val,
                   int kgroups) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int P = NEN * NEN;
  int64_t s = (int64_t)(blockIdx.x / (unsigned)kgroups);
  int lane = threadIdx.x;
  int k = (int)(blockIdx.x % (unsigned)kgroups) * blockDim.y + threadIdx.y;
  int base = slice_ptr[s];
  int w = (slice_ptr[s + 1] - base) >> 5;
  if (k >= w) return;
  int slot = base + (k << 5) + lane;
  int beg = slot_beg[slot], end = slot_end[slot];
  double acc[DM][DM];
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
  for (int t = beg; t < end; ++t) {
    uint32_t id = ent_list[t];
    uint32_t e = id / P;
    int p = (int)(id - e * P);
    int a = p / NEN, b = p - a * NEN;
    const double2* r2 = reinterpret_cast<const double2*>(rec + (int64_t)e * (NEN * NGP * 4));
#pragma unroll
    for (int gp = 0; gp < NGP; ++gp) {
      double2 a_lo, a_hi, b_lo, b_hi;
      if constexpr (V256) {
        const double* r1 = reinterpret_cast<const double*>(r2);
        femcy_d4 ra = femcy_ld256_nc(r1 + (a * NGP + gp) * 4), rb = femcy_ld256_nc(r1 + (b * NGP + gp) * 4);
        a_lo.x = ra.x; a_lo.y = ra.y; a_hi.x = ra.z; a_hi.y = ra.w;
        b_lo.x = rb.x; b_lo.y = rb.y; b_hi.x = rb.z; b_hi.y = rb.w;
      } else {
        a_lo = r2[(a * NGP + gp) * 2]; a_hi = r2[(a * NGP + gp) * 2 + 1];
        b_lo = r2[(b * NGP + gp) * 2]; b_hi = r2[(b * NGP + gp) * 2 + 1];
      }
      double ga[DM], gb[DM];
      ga[0] = a_lo.x; ga[1] = a_lo.y;
      gb[0] = b_lo.x; gb[1] = b_lo.y;
      if constexpr (DM == 3) { ga[2] = a_hi.x; gb[2] = b_hi.x; }
      if constexpr (CUBIC) {
        // variant 10: tangent of cubic form (checked on the host): ~27 instead of 99 FP64 instructions per block
        block_cubic_acc<DM>(tab.C[0], tab.C[1], tab.C[NV * NV - 1], ga, gb, a_hi.y, acc);
      } else {
        double T[NV][DM];
        C_times_B<DM>(tab.C, gb, T);
        Bt_times_T_acc<DM>(ga, T, a_hi.y, acc);
      }
    }
  }
  double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + lane;
#pragma unroll
  for (int i = 0; i < DM; ++i)
#pragma unroll
    for (int j = 0; j < DM; ++j) dst[(i * DM + j) << 5] = acc[i][j];
}

// Block from the accumulated gradient products  P = sum_{contributions, Gauss points} vol * (grad N_a) (grad N_b)^T :
//   K_ab[i][j] = sum_{k,l} C[voigt(i,k)][voigt(j,l)] * P[k][l]          (B_a^T C B_b is bilinear in grad N_a, grad N_b)
// valid because the tangent C is one constant matrix per run (stiffnessMtrx.py:124-129).  Cubic-form tangents (every
// tangent the reference ships, see tangent_is_cubic):  K_ij = q P_ij + r P_ji (i != j),  K_ii = p P_ii + r (tr P - P_ii).
template <int DM, bool CUBIC>
__device__ __forceinline__ void block_from_products(const double* __restrict__ C, const double (&P)[DM][DM],
                                                    double (&K)[DM][DM]) {
  constexpr int NV = Voigt<DM>::NV;
  if constexpr (CUBIC) {
    const double p = C[0], q = C[1], r = C[NV * NV - 1];
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < DM; ++i) tr += P[i][i];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) K[i][j] = (i == j) ? (p - r) * P[i][i] + r * tr : q * P[i][j] + r * P[j][i];
  } else {
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DM; ++k)
#pragma unroll
          for (int l = 0; l < DM; ++l) {
            // Voigt row of the symmetric pair: 2-D [xx, yy, xy]; 3-D [xx, yy, zz, xy, zx, yz]
            const int vik = (i == k) ? i : (DM == 2 ? 2 : (i + k == 1 ? 3 : (i + k == 2 ? 4 : 5)));
            const int vjl = (j == l) ? j : (DM == 2 ? 2 : (j + l == 1 ? 3 : (j + l == 2 ? 4 : 5)));
            s += C[vik * NV + vjl] * P[k][l];
          }
        K[i][j] = s;
      }
  }
}

// Default assembly, pass 2: per-block gather over the node-sector records with the gradient products accumulated first
// (12 FP64 instructions per contribution and Gauss point instead of 43 / 99; the tangent enters once per stored block).
// One thread per stored block; a block of 8 warps walks the block columns k = ty, ty+8, ... of ONE 32-row slice, so
// all loads of a slice's element records come from the same SM (L1 reuse) and no warp exits idle.
template <int DM, int NEN, int NGP, bool CUBIC>
__global__ void __launch_bounds__(256)
k_assemble_gather_p(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr,
                    const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                    const uint32_t* __restrict__ ent_list, const double* __restrict__ rec, double* __restrict__ val,
                    int64_t nslice) {
  constexpr int DM2 = DM * DM;
  constexpr int P = NEN * NEN;
  const int lane = threadIdx.x;
  for (int64_t s = blockIdx.x; s < nslice; s += gridDim.x) {
    const int base = slice_ptr[s];
    const int w = (slice_ptr[s + 1] - base) >> 5;
    for (int k = threadIdx.y; k < w; k += blockDim.y) {
      const int slot = base + (k << 5) + lane;
      const int beg = slot_beg[slot], end = slot_end[slot];
      double acc[DM][DM];
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
      uint32_t id_next = (beg < end) ? ent_list[beg] : 0u;
      for (int t = beg; t < end; ++t) {
        const uint32_t id = id_next;
        if (t + 1 < end) id_next = ent_list[t + 1];
        const uint32_t e = id / P;
        const int p = (int)(id - e * P);
        const int a = p / NEN, b = p - a * NEN;
        const double* r1 = rec + (int64_t)e * (NEN * NGP * 4);
#pragma unroll
        for (int gp = 0; gp < NGP; ++gp) {
          const femcy_d4 ra = femcy_ld256_nc(r1 + (a * NGP + gp) * 4), rb = femcy_ld256_nc(r1 + (b * NGP + gp) * 4);
          const double sa0 = ra.w * ra.x, sa1 = ra.w * ra.y;
          acc[0][0] += sa0 * rb.x; acc[0][1] += sa0 * rb.y;
          acc[1][0] += sa1 * rb.x; acc[1][1] += sa1 * rb.y;
          if constexpr (DM == 3) {
            const double sa2 = ra.w * ra.z;
            acc[0][2] += sa0 * rb.z; acc[1][2] += sa1 * rb.z;
            acc[2][0] += sa2 * rb.x; acc[2][1] += sa2 * rb.y; acc[2][2] += sa2 * rb.z;
          }
        }
      }
      double K[DM][DM];
      block_from_products<DM, CUBIC>(tab.C, acc, K);
      double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + lane;
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) dst[(i * DM + j) << 5] = K[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// "tile" assembly (EXPERIMENTAL, variant 14; single-Gauss-point elements): the per-block gather with its operands in
// SHARED memory.  One block per 32-row slice first copies the node-sector records of every element that touches the
// slice (tile_elems, ~400 records = 51 KB for the 10 M-element C3D4 mesh) into shared memory with coalesced 16-byte
// loads -- each record leaves L2 once per slice instead of once per stored block -- then every thread sums the
// contributions of its block slots (k = ty, ty+8, ...) out of shared memory: the dependent L2 round trip of the gather's
// inner loop becomes a shared-memory access.  Contributions are visited in the order of the per-block gather, so the
// result is bitwise that of variants 9 / 10.  Chunks of record j are rotated by j so that lanes reading the same node
// of different records spread over the banks.
template <int DM, int NEN, bool CUBIC>
__global__ void __launch_bounds__(256)
k_assemble_tile(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr,
                const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                const uint32_t* __restrict__ ent_tile, const int32_t* __restrict__ tile_ptr,
                const uint32_t* __restrict__ tile_elems, const double* __restrict__ rec, double* __restrict__ val) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int CH = NEN * 2;                 // 16-byte chunks per record (one Gauss point)
#ifdef FEMCY_SIMT_EMU
  double2* tile_s = static_cast<double2*>(simt::dyn_smem());
#else
  extern __shared__ double2 tile_s[];
#endif
  const int64_t s = blockIdx.x;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 32 + lane;
  const int t0 = tile_ptr[s], nt = tile_ptr[s + 1] - t0;
  const double2* rec2 = reinterpret_cast<const double2*>(rec);
  for (int i = tid; i < nt * CH; i += 256) {
    int j = i / CH, c = i - j * CH;
    uint32_t e = tile_elems[t0 + j];
    tile_s[j * CH + (c + j) % CH] = rec2[(int64_t)e * CH + c];
  }
  __syncthreads();
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) >> 5;
  for (int k = ty; k < w; k += 8) {
    int slot = base + (k << 5) + lane;
    int beg = slot_beg[slot], end = slot_end[slot];
    double acc[DM][DM];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
    for (int t = beg; t < end; ++t) {
      uint32_t id = ent_tile[t];
      int j = (int)(id >> 8), p = (int)(id & 255u);
      int a = p / NEN, b = p - a * NEN;
      const double2* r2 = tile_s + j * CH;
      double2 a_lo = r2[(2 * a + j) % CH], a_hi = r2[(2 * a + 1 + j) % CH];
      double2 b_lo = r2[(2 * b + j) % CH], b_hi = r2[(2 * b + 1 + j) % CH];
      double ga[DM], gb[DM];
      ga[0] = a_lo.x; ga[1] = a_lo.y;
      gb[0] = b_lo.x; gb[1] = b_lo.y;
      if constexpr (DM == 3) { ga[2] = a_hi.x; gb[2] = b_hi.x; }
      if constexpr (CUBIC) {
        block_cubic_acc<DM>(tab.C[0], tab.C[1], tab.C[NV * NV - 1], ga, gb, a_hi.y, acc);
      } else {
        double T[NV][DM];
        C_times_B<DM>(tab.C, gb, T);
        Bt_times_T_acc<DM>(ga, T, a_hi.y, acc);
      }
    }
    double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + lane;
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) dst[(i * DM + j) << 5] = acc[i][j];
  }
}

// tile assembly with TMA-engine loads (EXPERIMENTAL, variant 22): k_assemble_tile whose staging loop -- 8 LDG.128 + 8 STS.128
// per record and thread -- is replaced by ONE bulk copy per record (cp.async.bulk.shared::cta.global, 128 bytes for C3D4)
// completing on an mbarrier: thread 0 arrives with the tile's byte count, every thread issues the copies of its records,
// everybody waits on the barrier's phase.  A bulk copy cannot rotate the chunks of a record, so the bank spreading of
// k_assemble_tile comes from the pitch instead: records are PB = record + 16 bytes apart (144 B: record j starts at
// 16-byte group 9j mod 8 = j mod 8).  Same visiting order, so bitwise the result of variants 9 / 10 / 14.
template <int NEN>
struct TileBCfg { static constexpr int RECB = NEN * 32; static constexpr int PB = RECB + 16; };

template <int DM, int NEN, bool CUBIC>
__global__ void __launch_bounds__(256)
k_assemble_tile_b(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr,
                  const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                  const uint32_t* __restrict__ ent_tile, const int32_t* __restrict__ tile_ptr,
                  const uint32_t* __restrict__ tile_elems, const double* __restrict__ rec, double* __restrict__ val) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int RECB = TileBCfg<NEN>::RECB, PB = TileBCfg<NEN>::PB;
#ifdef FEMCY_SIMT_EMU
  char* tile_s = static_cast<char*>(simt::dyn_smem());
#else
  extern __shared__ __align__(128) char tile_bytes_s[];
  char* tile_s = tile_bytes_s;
#endif
  alignas(8) __shared__ unsigned long long mbar;
  const int64_t s = blockIdx.x;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 32 + lane;
  const int t0 = tile_ptr[s], nt = tile_ptr[s + 1] - t0;
  if (tid == 0) femcy_mbar_init(&mbar, 1);
  __syncthreads();
  if (tid == 0) femcy_mbar_arrive_expect_tx(&mbar, (unsigned)(nt * RECB));
  for (int j = tid; j < nt; j += 256) {
    const uint32_t e = tile_elems[t0 + j];
    femcy_bulk_load(tile_s + (size_t)j * PB, reinterpret_cast<const char*>(rec) + (size_t)e * RECB, RECB, &mbar);
  }
  femcy_mbar_wait(&mbar, 0);
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) >> 5;
  for (int k = ty; k < w; k += 8) {
    int slot = base + (k << 5) + lane;
    int beg = slot_beg[slot], end = slot_end[slot];
    double acc[DM][DM];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
    for (int t = beg; t < end; ++t) {
      uint32_t id = ent_tile[t];
      int j = (int)(id >> 8), p = (int)(id & 255u);
      int a = p / NEN, b = p - a * NEN;
      const double2* r2 = reinterpret_cast<const double2*>(tile_s + (size_t)j * PB);
      double2 a_lo = r2[2 * a], a_hi = r2[2 * a + 1];
      double2 b_lo = r2[2 * b], b_hi = r2[2 * b + 1];
      double ga[DM], gb[DM];
      ga[0] = a_lo.x; ga[1] = a_lo.y;
      gb[0] = b_lo.x; gb[1] = b_lo.y;
      if constexpr (DM == 3) { ga[2] = a_hi.x; gb[2] = b_hi.x; }
      if constexpr (CUBIC) {
        block_cubic_acc<DM>(tab.C[0], tab.C[1], tab.C[NV * NV - 1], ga, gb, a_hi.y, acc);
      } else {
        double T[NV][DM];
        C_times_B<DM>(tab.C, gb, T);
        Bt_times_T_acc<DM>(ga, T, a_hi.y, acc);
      }
    }
    double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + lane;
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) dst[(i * DM + j) << 5] = acc[i][j];
  }
}

// tile assembly for elements with several Gauss points / many nodes (EXPERIMENTAL, variant 15; C3D10, CPS6, CPS8, CPS4):
// one block per 8 consecutive row positions of a slice, one thread per (row, k) block slot: 8 x 72 threads (a C3D10
// corner node has 65 blocks per row; rows with more than 72 take another pass over the tile).  The records of the
// elements touching the 8 rows are staged in shared memory ONE GAUSS POINT AT A TIME (C3D10: ~80 elements x 10 nodes x
// 32 B = 26 KB per Gauss point, so several blocks stay resident per SM); the accumulator of a slot lives in registers
// across the Gauss-point loop.  Per Gauss point a contribution costs two 32 B shared-memory reads instead of two
// dependent L2 sector reads; each element record leaves L2 once per 8-row block it touches (~8x) instead of once per
// stored block it contributes to (100x).
#define FEMCY_TILE_RB 8
#ifndef FEMCY_TILE_KT
#define FEMCY_TILE_KT 72      // (the emulation tests also build with 8 to exercise the multi-pass path)
#endif
template <int DM, int NEN, int NGP, bool CUBIC>
__global__ void __launch_bounds__(FEMCY_TILE_RB * FEMCY_TILE_KT, 2)
k_assemble_tile_mgp(const __grid_constant__ ElemTables tab, const int32_t* __restrict__ slice_ptr,
                    const int32_t* __restrict__ slot_beg, const int32_t* __restrict__ slot_end,
                    const uint32_t* __restrict__ ent_tile, const int32_t* __restrict__ tile_ptr,
                    const uint32_t* __restrict__ tile_elems, const double* __restrict__ rec, double* __restrict__ val) {
  constexpr int NV = Voigt<DM>::NV;
  constexpr int DM2 = DM * DM;
  constexpr int CHG = NEN * 2;                // 16-byte chunks per element and Gauss point
  constexpr int RB = FEMCY_TILE_RB, KT = FEMCY_TILE_KT;
#ifdef FEMCY_SIMT_EMU
  double2* tile_s = static_cast<double2*>(simt::dyn_smem());
#else
  extern __shared__ double2 tile_s[];
#endif
  const int64_t blk = blockIdx.x;
  const int64_t s = blk >> 2;
  const int r0 = (int)(blk & 3) * RB;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * RB + tx;
  const int t0 = tile_ptr[blk], nt = tile_ptr[blk + 1] - t0;
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) >> 5;
  const double2* rec2 = reinterpret_cast<const double2*>(rec);
  for (int k0 = 0; k0 < w; k0 += KT) {        // one pass unless a row has more than KT blocks
    const int k = k0 + ty;
    int beg = 0, end = 0, slot = 0;
    if (k < w) {
      slot = base + (k << 5) + r0 + tx;
      beg = slot_beg[slot]; end = slot_end[slot];
    }
    double acc[DM][DM];
#pragma unroll
    for (int i = 0; i < DM; ++i)
#pragma unroll
      for (int j = 0; j < DM; ++j) acc[i][j] = 0.0;
#pragma unroll 1
    for (int gp = 0; gp < NGP; ++gp) {
      for (int i = tid; i < nt * CHG; i += RB * KT) {
        int j = i / CHG, c = i - j * CHG;     // c = node*2 + half
        uint32_t e = tile_elems[t0 + j];
        tile_s[j * CHG + (c + j) % CHG] = rec2[((int64_t)e * NEN * NGP + (c >> 1) * NGP + gp) * 2 + (c & 1)];
      }
      __syncthreads();
      for (int t = beg; t < end; ++t) {
        uint32_t id = ent_tile[t];
        int j = (int)(id >> 8), p = (int)(id & 255u);
        int a = p / NEN, b = p - a * NEN;
        const double2* r2 = tile_s + j * CHG;
        double2 a_lo = r2[(2 * a + j) % CHG], a_hi = r2[(2 * a + 1 + j) % CHG];
        double2 b_lo = r2[(2 * b + j) % CHG], b_hi = r2[(2 * b + 1 + j) % CHG];
        double ga[DM], gb[DM];
        ga[0] = a_lo.x; ga[1] = a_lo.y;
        gb[0] = b_lo.x; gb[1] = b_lo.y;
        if constexpr (DM == 3) { ga[2] = a_hi.x; gb[2] = b_hi.x; }
        if constexpr (CUBIC) {
          block_cubic_acc<DM>(tab.C[0], tab.C[1], tab.C[NV * NV - 1], ga, gb, a_hi.y, acc);
        } else {
          double T[NV][DM];
          C_times_B<DM>(tab.C, gb, T);
          Bt_times_T_acc<DM>(ga, T, a_hi.y, acc);
        }
      }
      __syncthreads();
    }
    if (k < w) {
      double* dst = val + (((int64_t)(slot >> 5) * DM2) << 5) + (slot & 31);
#pragma unroll
      for (int i = 0; i < DM; ++i)
#pragma unroll
        for (int j = 0; j < DM; ++j) dst[(i * DM + j) << 5] = acc[i][j];
    }
  }
}
