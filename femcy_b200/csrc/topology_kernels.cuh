// Device code of row f1: mesh-topology builders and the Neumann load vector -- everything except
// the CUB sort / scan calls that topology.cu makes between these kernels.  Header for the same reason as
// assembly_kernels.cuh (host launch code in topology.cu, CPU SIMT emulation in tests/simt).
//
// Reference code replaced (host Python loops over Python sets / dicts):
//   Body.get_boundary                          /root/reference/body.py:197-234
//   Body.get_nodeEles, nodeEles field          /root/reference/body.py:165-179, stiffnessMtrx.py:70-76
//   System_of_equations.neumannBC              /root/reference/stiffnessMtrx.py:369-411
//   ELE.globalNormal                           e.g. /root/reference/element_zoo/element_linear_tetrahedral.py:98-134
#pragma once
#include "device_compat.cuh"
#include "kernel_types.cuh"
#include "elem_math.cuh"

#define FEMCY_MAX_FACET_KEYS 8      // facet keys per element kind (CPS8: 8 half edges)
#define FEMCY_MAX_FACET_WIDTH 6     // nodes per facet (C3D10: 6-node face)
#define FEMCY_MAX_FACET_POINTS 6    // integration points per facet

// sorted global node ids of facet (element e, key k)
__device__ __forceinline__ void facet_nodes_sorted(const int32_t* __restrict__ elems, int n_en, const int32_t* __restrict__ key_nodes,
                                                   int width, int64_t e, int k, int32_t (&s)[FEMCY_MAX_FACET_WIDTH]) {
  for (int q = 0; q < width; ++q) {
    int32_t v = elems[e * n_en + key_nodes[k * width + q]];
    int p = q;
    while (p > 0 && s[p - 1] > v) { s[p] = s[p - 1]; --p; }     // insertion sort (width <= 6)
    s[p] = v;
  }
}

// facet id t = k*ne + e (the order of Body.boundary_arrays); sort key = the facet's two smallest node ids
__global__ void k_facet_keys(const int32_t* __restrict__ elems, int64_t ne, int n_en, int64_t nn, const int32_t* __restrict__ key_nodes,
                             int nkeys, int width, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids) {
  const int64_t total = ne * nkeys;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t / ne);
    const int64_t e = t - (int64_t)k * ne;
    int32_t s[FEMCY_MAX_FACET_WIDTH];
    facet_nodes_sorted(elems, n_en, key_nodes, width, e, k, s);
    keys[t] = (uint64_t)s[0] * (uint64_t)nn + (uint64_t)s[1];
    ids[t] = (uint32_t)t;
  }
}

// a facet is on the boundary when no other facet has the same node set: facets with the same two smallest nodes are
// neighbours in the sorted order (a handful per run), so each one compares itself with the rest of its run
__global__ void k_facet_unique(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids, int64_t total,
                               const int32_t* __restrict__ elems, int64_t ne, int n_en, const int32_t* __restrict__ key_nodes,
                               int width, int32_t* __restrict__ is_boundary /*[total], by facet id*/) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[t];
    const uint32_t id = ids[t];
    int32_t a[FEMCY_MAX_FACET_WIDTH], b[FEMCY_MAX_FACET_WIDTH];
    facet_nodes_sorted(elems, n_en, key_nodes, width, (int64_t)(id % (uint64_t)ne), (int)(id / (uint64_t)ne), a);
    int twins = 0;
    for (int dir = -1; dir <= 1; dir += 2) {
      for (int64_t j = t + dir; j >= 0 && j < total && keys[j] == key; j += dir) {
        const uint32_t jd = ids[j];
        facet_nodes_sorted(elems, n_en, key_nodes, width, (int64_t)(jd % (uint64_t)ne), (int)(jd / (uint64_t)ne), b);
        bool same = true;
        for (int q = 2; q < width; ++q) same = same && (a[q] == b[q]);
        if (same) ++twins;
      }
    }
    is_boundary[id] = (twins == 0) ? 1 : 0;
  }
}

// pos = exclusive scan of is_boundary: boundary facets in ascending facet id
__global__ void k_facet_compact(const int32_t* __restrict__ is_boundary, const int32_t* __restrict__ pos, int64_t total, int64_t ne,
                                int32_t* __restrict__ b_elem, int32_t* __restrict__ b_kid) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    if (is_boundary[t]) {
      b_elem[pos[t]] = (int32_t)(t % ne);
      b_kid[pos[t]] = (int32_t)(t / ne);
    }
  }
}

// node -> elements: keys node*ne + e of all (element, local node) pairs; after the sort, ptr[i] = first key >= i*ne
__global__ void k_node_elem_keys(const int32_t* __restrict__ elems, int64_t ne, int n_en, uint64_t* __restrict__ keys) {
  const int64_t total = ne * n_en;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    keys[t] = (uint64_t)elems[t] * (uint64_t)ne + (uint64_t)(t / n_en);
}
__global__ void k_node_elem_csr(const uint64_t* __restrict__ keys, int64_t total, int64_t ne, int64_t nn, int32_t* __restrict__ ptr,
                                int32_t* __restrict__ list) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= nn; i += gs) {
    const uint64_t want = (uint64_t)i * (uint64_t)ne;
    int64_t lo = 0, hi = total;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    ptr[i] = (int32_t)lo;
  }
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += gs) list[t] = (int32_t)(keys[t] % (uint64_t)ne);
}

// ---------------------------------------------------------------------------------------------------------------
// Facet tables of one element kind (element_zoo: facet_natural_coos, facet_point_weights, facet_natural_normals,
// shapeFunc / dshape_dnat at the facet points), device arrays owned by the ctx (femcy_set_facet_tables)
struct FacetTables {
  int nkeys = 0, width = 0, nfp = 0;
  const int32_t* key_nodes = nullptr;   // [nkeys][width]  local nodes of the facet, ascending (the reference's sorted key)
  const double* w = nullptr;            // [nkeys][nfp]
  const double* normal = nullptr;       // [nkeys][nfp][dm]      natural-space outward normal
  const double* N = nullptr;            // [nkeys][nfp][width]   shape functions of the facet's nodes at the facet points
  const double* dN = nullptr;           // [nkeys][nfp][n_en][dm]
};

// consistent nodal loads of a traction on a list of (element, facet key) pairs (stiffnessMtrx.py:386-411):
//   rhs[node*dm + i] += t * (n or dir)_i * size * w_p * N_node(xi_p)
// n = unit outward normal (natural normal pushed forward with (dx/dxi)^-1 at xi_p), size = distance of the facet's first two
// nodes (2-D) / area of the triangle of its first three (3-D), on the INITIAL geometry (dead loads, quirk B3).  rhs is
// zero-filled by the caller (rhs.fill(0) at :384: only the last *Dsload of a deck acts).  One thread per facet.
template <int DM>
__global__ void __launch_bounds__(128)
k_neumann(const FacetTables T, const int32_t* __restrict__ f_elem, const int32_t* __restrict__ f_kid, int64_t nf,
          const int32_t* __restrict__ elems, int n_en, const double* __restrict__ nodes, double traction, int has_dir, double d0,
          double d1, double d2, double* __restrict__ rhs) {
  const double dir[3] = {d0, d1, d2};
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = f_elem[f];
    const int k = f_kid[f];
    const int32_t* kn = T.key_nodes + k * T.width;
    const int32_t* conn = elems + e * n_en;
    double X0[DM], X1[DM], X2[DM];
#pragma unroll
    for (int i = 0; i < DM; ++i) {
      X0[i] = nodes[(int64_t)conn[kn[0]] * DM + i];
      X1[i] = nodes[(int64_t)conn[kn[1]] * DM + i];
      X2[i] = (DM == 3) ? nodes[(int64_t)conn[kn[2]] * DM + i] : 0.0;
    }
    double size;
    if constexpr (DM == 2) {
      size = sqrt((X0[0] - X1[0]) * (X0[0] - X1[0]) + (X0[1] - X1[1]) * (X0[1] - X1[1]));
    } else {
      const double a[3] = {X1[0] - X0[0], X1[1] - X0[1], X1[2] - X0[2]}, b[3] = {X2[0] - X0[0], X2[1] - X0[1], X2[2] - X0[2]};
      const double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
      size = 0.5 * sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    }
    for (int p = 0; p < T.nfp; ++p) {
      const int kp = k * T.nfp + p;
      double flux[DM];
      if (has_dir) {
#pragma unroll
        for (int i = 0; i < DM; ++i) flux[i] = traction * dir[i] * (size * T.w[kp]);
      } else {
        // dx/dxi at the facet point from ALL nodes of the element, then n = natural normal . (dx/dxi)^-1, normalised
        const double* dN = T.dN + (int64_t)kp * n_en * DM;
        double J[DM][DM], Ji[DM][DM];
#pragma unroll
        for (int i = 0; i < DM; ++i)
#pragma unroll
          for (int q = 0; q < DM; ++q) J[i][q] = 0.0;
        for (int a = 0; a < n_en; ++a) {
          const int64_t nd = conn[a];
#pragma unroll
          for (int i = 0; i < DM; ++i) {
            const double x = nodes[nd * DM + i];
#pragma unroll
            for (int q = 0; q < DM; ++q) J[i][q] += x * dN[a * DM + q];
          }
        }
        inv_dm<DM>(J, Ji);
        double n[DM], nrm = 0.0;
#pragma unroll
        for (int j = 0; j < DM; ++j) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < DM; ++q) s += T.normal[kp * DM + q] * Ji[q][j];
          n[j] = s;
          nrm += s * s;
        }
        nrm = sqrt(nrm) + 1.0e-30;
#pragma unroll
        for (int i = 0; i < DM; ++i) flux[i] = traction * (n[i] / nrm) * (size * T.w[kp]);
      }
      for (int q = 0; q < T.width; ++q) {
        const double Na = T.N[kp * T.width + q];
        const int64_t nd = conn[kn[q]];
#pragma unroll
        for (int i = 0; i < DM; ++i) atomicAdd(rhs + nd * DM + i, flux[i] * Na);
      }
    }
  }
}
