// Device partitioner (row f1 / section 8e): the piece of a global mesh that one rank works on -- node owners (slabs along one
// axis), the elements touching the rank's nodes, local numbering (owned nodes, then ghosts grouped by owner) and the halo
// plan -- computed on the rank's own GPU instead of redundantly in host NumPy by every rank (femcy_b200/partition.py states
// the scheme; the reference is single-device, so there is no reference counterpart).
// Kernels + ONE orchestration (`partition_build`) shared by the product (partition.cu: CUB sorts / scans on the ctx stream)
// and by the CPU SIMT emulation (tests/simt: std::stable_sort / loops), so the not-gpu suite checks the orchestration too.
#pragma once
#include <string.h>

#include <vector>

#include "device_compat.cuh"
#include "kernel_types.cuh"

// sortable 64-bit image of a coordinate (ascending order of the doubles = ascending order of the keys; -0.0 == +0.0)
__global__ void k_part_coord_keys(const double* __restrict__ nodes, int64_t nn, int dm, int axis, uint64_t* __restrict__ keys,
                                  uint32_t* __restrict__ ids) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    double x = nodes[i * dm + axis];
    if (x == 0.0) x = 0.0;                                   // -0.0 -> +0.0 (they compare equal on the host)
    unsigned long long b;
    memcpy(&b, &x, sizeof(b));
    keys[i] = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    ids[i] = (uint32_t)i;
  }
}

// the i-th node of the sorted order belongs to the rank whose chunk [bounds[r], bounds[r+1]) holds i
__global__ void k_part_owner(const uint32_t* __restrict__ order, int64_t nn, const int64_t* __restrict__ bounds, int nranks,
                             int32_t* __restrict__ owner) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    int r = 0;
    while (r + 1 < nranks && i >= bounds[r + 1]) ++r;
    owner[order[i]] = r;
  }
}

// elements with a node of `rank`; their nodes are `used`; an owned node of such an element is sent to the owners of the
// element's other nodes (bit p of send_mask)
__global__ void k_part_touch(const int32_t* __restrict__ elems, int64_t ne, int n_en, const int32_t* __restrict__ owner, int rank,
                             int32_t* __restrict__ touch, int32_t* __restrict__ used, unsigned int* __restrict__ send_mask) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    unsigned int owners = 0u;
    for (int a = 0; a < n_en; ++a) owners |= 1u << owner[elems[e * n_en + a]];
    const bool mine = (owners >> rank) & 1u;
    touch[e] = mine ? 1 : 0;
    if (!mine) continue;
    const unsigned int others = owners & ~(1u << rank);
    for (int a = 0; a < n_en; ++a) {
      const int32_t nd = elems[e * n_en + a];
      used[nd] = 1;
      if (others && owner[nd] == rank) atomicOr(send_mask + nd, others);
    }
  }
}

__global__ void k_part_compact_elems(const int32_t* __restrict__ touch, const int32_t* __restrict__ pos, int64_t ne,
                                     const int32_t* __restrict__ elems, int n_en, const int32_t* __restrict__ owner, int rank,
                                     int64_t* __restrict__ elem_ids, unsigned char* __restrict__ primary) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    if (!touch[e]) continue;
    elem_ids[pos[e]] = e;
    primary[pos[e]] = owner[elems[e * n_en]] == rank ? 1 : 0;       // reported once: on the rank owning its first node
  }
}

// own_flag: nodes of `rank` (used by an element or not, as on the host); ghost_flag: used nodes of other ranks
__global__ void k_part_node_flags(const int32_t* __restrict__ owner, const int32_t* __restrict__ used, int64_t nn, int rank,
                                  int32_t* __restrict__ own_flag, int32_t* __restrict__ ghost_flag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    const bool own = owner[i] == rank;
    own_flag[i] = own ? 1 : 0;
    ghost_flag[i] = (!own && used[i]) ? 1 : 0;
  }
}

// owned nodes: local id = rank among the owned (ascending global id); ghosts: key owner*nn + id, sorted afterwards
__global__ void k_part_number(const int32_t* __restrict__ own_flag, const int32_t* __restrict__ own_pos,
                              const int32_t* __restrict__ ghost_flag, const int32_t* __restrict__ ghost_pos,
                              const int32_t* __restrict__ owner, int64_t nn, int64_t* __restrict__ l2g, int64_t* __restrict__ g2l,
                              uint64_t* __restrict__ ghost_keys) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    g2l[i] = -1;
    if (own_flag[i]) { l2g[own_pos[i]] = i; g2l[i] = own_pos[i]; }
    if (ghost_flag[i]) ghost_keys[ghost_pos[i]] = (uint64_t)owner[i] * (uint64_t)nn + (uint64_t)i;
  }
}

// sorted ghost keys -> local ids n_own + j; per-owner counts (the receive ranges)
__global__ void k_part_ghosts(const uint64_t* __restrict__ ghost_keys, int64_t n_ghost, int64_t nn, int64_t n_own,
                              int64_t* __restrict__ l2g, int64_t* __restrict__ g2l, int32_t* __restrict__ recv_count) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_ghost; j += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = ghost_keys[j];
    const int64_t g = (int64_t)(k % (uint64_t)nn);
    l2g[n_own + j] = g;
    g2l[g] = n_own + j;
    atomicAdd(recv_count + (int)(k / (uint64_t)nn), 1);
  }
}

__global__ void k_part_local_elems(const int64_t* __restrict__ elem_ids, int64_t ne_loc, const int32_t* __restrict__ elems, int n_en,
                                   const int64_t* __restrict__ g2l, int32_t* __restrict__ loc_elems) {
  const int64_t total = ne_loc * n_en;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / n_en;
    loc_elems[t] = (int32_t)g2l[elems[elem_ids[j] * n_en + (t - j * n_en)]];
  }
}

__global__ void k_part_send_flag(const unsigned int* __restrict__ send_mask, const int32_t* __restrict__ own_flag, int64_t nn, int peer,
                                 int32_t* __restrict__ flag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (own_flag[i] && ((send_mask[i] >> peer) & 1u)) ? 1 : 0;
}
__global__ void k_part_send_list(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int64_t nn,
                                 const int64_t* __restrict__ g2l, int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x)
    if (flag[i]) out[pos[i]] = (int32_t)g2l[i];
}
__global__ void k_part_gather_nodes(const int64_t* __restrict__ l2g, int64_t n_local, int dm, const double* __restrict__ nodes,
                                    double* __restrict__ loc_nodes) {
  const int64_t total = n_local * dm;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / dm;
    loc_nodes[t] = nodes[l2g[j] * dm + (t - j * dm)];
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct PartitionResult {        // device arrays (backend-owned) + sizes; downloaded by the caller
  int64_t nn = 0, ne = 0, n_own = 0, n_local = 0, ne_local = 0;
  int npeers = 0;
  int peers[FEMCY_MAX_RANKS];
  int64_t send_ptr[FEMCY_MAX_RANKS + 1], recv_ptr[FEMCY_MAX_RANKS + 1];
  int32_t* owner = nullptr;            // [nn]
  int64_t* elem_ids = nullptr;         // [ne_local]
  unsigned char* primary = nullptr;    // [ne_local]
  int64_t* l2g = nullptr;              // [n_local]
  int32_t* loc_elems = nullptr;        // [ne_local*n_en]
  double* loc_nodes = nullptr;         // [n_local*dm]
  int32_t* send_nodes = nullptr;       // [send_ptr[npeers]]  local ids
  int32_t* recv_nodes = nullptr;       // [recv_ptr[npeers]]
};

// Backend B provides: alloc<T>(n) -> T* (zero-initialised, freed with the backend), grid(n), sort_pairs(keys_in, keys_out, ids_in,
// ids_out, n), sort_keys(in, out, n), exclusive_sum(in, out, n), read(ptr) -> host copy of one element (synchronises),
// upload(dst, host_src, n), ok() (false after a failed runtime call), and the PART_LAUNCH macro of its translation unit.
template <class B>
int partition_build(B& be, int dm, int64_t nn, const double* nodes, int64_t ne, int n_en, const int32_t* elems, int rank, int nranks,
                    int axis, const int64_t* bounds_host, PartitionResult& R) {
  R = PartitionResult();
  R.nn = nn; R.ne = ne;
  // 1. owners: one stable sort of the coordinates along `axis` (ties by node id), chunks of the sorted order
  uint64_t* ck = be.template alloc<uint64_t>(nn); uint64_t* ck2 = be.template alloc<uint64_t>(nn);
  uint32_t* ci = be.template alloc<uint32_t>(nn); uint32_t* ci2 = be.template alloc<uint32_t>(nn);
  int64_t* bounds = be.template alloc<int64_t>(nranks + 1);
  R.owner = be.template alloc<int32_t>(nn);
  if (!be.ok()) return 1;
  be.upload(bounds, bounds_host, nranks + 1);
  PART_LAUNCH(be, nn, k_part_coord_keys, nodes, nn, dm, axis, ck, ci);
  be.sort_pairs(ck, ck2, ci, ci2, nn);
  PART_LAUNCH(be, nn, k_part_owner, ci2, nn, bounds, nranks, R.owner);
  // 2. local elements
  int32_t* touch = be.template alloc<int32_t>(ne); int32_t* tpos = be.template alloc<int32_t>(ne);
  int32_t* used = be.template alloc<int32_t>(nn);
  unsigned int* send_mask = be.template alloc<unsigned int>(nn);
  if (!be.ok()) return 1;
  PART_LAUNCH(be, ne, k_part_touch, elems, ne, n_en, R.owner, rank, touch, used, send_mask);
  be.exclusive_sum(touch, tpos, ne);
  R.ne_local = ne > 0 ? (int64_t)be.read(tpos + (ne - 1)) + be.read(touch + (ne - 1)) : 0;
  R.elem_ids = be.template alloc<int64_t>(R.ne_local);
  R.primary = be.template alloc<unsigned char>(R.ne_local);
  if (!be.ok()) return 1;
  PART_LAUNCH(be, ne, k_part_compact_elems, touch, tpos, ne, elems, n_en, R.owner, rank, R.elem_ids, R.primary);
  // 3. local node numbering: owned (ascending), then ghosts by (owner, id)
  int32_t* own_flag = be.template alloc<int32_t>(nn); int32_t* own_pos = be.template alloc<int32_t>(nn);
  int32_t* gh_flag = be.template alloc<int32_t>(nn); int32_t* gh_pos = be.template alloc<int32_t>(nn);
  if (!be.ok()) return 1;
  PART_LAUNCH(be, nn, k_part_node_flags, R.owner, used, nn, rank, own_flag, gh_flag);
  be.exclusive_sum(own_flag, own_pos, nn);
  be.exclusive_sum(gh_flag, gh_pos, nn);
  R.n_own = nn > 0 ? (int64_t)be.read(own_pos + (nn - 1)) + be.read(own_flag + (nn - 1)) : 0;
  const int64_t n_ghost = nn > 0 ? (int64_t)be.read(gh_pos + (nn - 1)) + be.read(gh_flag + (nn - 1)) : 0;
  R.n_local = R.n_own + n_ghost;
  R.l2g = be.template alloc<int64_t>(R.n_local);
  int64_t* g2l = be.template alloc<int64_t>(nn);
  uint64_t* gk = be.template alloc<uint64_t>(n_ghost); uint64_t* gk2 = be.template alloc<uint64_t>(n_ghost);
  int32_t* recv_count = be.template alloc<int32_t>(FEMCY_MAX_RANKS);
  if (!be.ok()) return 1;
  PART_LAUNCH(be, nn, k_part_number, own_flag, own_pos, gh_flag, gh_pos, R.owner, nn, R.l2g, g2l, gk);
  be.sort_keys(gk, gk2, n_ghost);
  PART_LAUNCH(be, n_ghost, k_part_ghosts, gk2, n_ghost, nn, R.n_own, R.l2g, g2l, recv_count);
  // 4. local connectivity and coordinates
  R.loc_elems = be.template alloc<int32_t>(R.ne_local * n_en);
  R.loc_nodes = be.template alloc<double>(R.n_local * dm);
  if (!be.ok()) return 1;
  PART_LAUNCH(be, R.ne_local * n_en, k_part_local_elems, R.elem_ids, R.ne_local, elems, n_en, g2l, R.loc_elems);
  PART_LAUNCH(be, R.n_local * dm, k_part_gather_nodes, R.l2g, R.n_local, dm, nodes, R.loc_nodes);
  // 5. halo plan: per peer (ascending rank) the ghosts it owns (one contiguous local range) and my owned nodes it needs
  int64_t n_send[FEMCY_MAX_RANKS], n_recv[FEMCY_MAX_RANKS];
  int32_t* sflag = be.template alloc<int32_t>(nn); int32_t* spos = be.template alloc<int32_t>(nn);
  if (!be.ok()) return 1;
  int64_t send_total = 0, recv_total = 0;
  for (int p = 0; p < nranks; ++p) {
    n_send[p] = n_recv[p] = 0;
    if (p == rank) continue;
    n_recv[p] = be.read(recv_count + p);
    PART_LAUNCH(be, nn, k_part_send_flag, send_mask, own_flag, nn, p, sflag);
    be.exclusive_sum(sflag, spos, nn);
    n_send[p] = nn > 0 ? (int64_t)be.read(spos + (nn - 1)) + be.read(sflag + (nn - 1)) : 0;
    send_total += n_send[p];
    recv_total += n_recv[p];
  }
  R.send_nodes = be.template alloc<int32_t>(send_total);
  R.recv_nodes = be.template alloc<int32_t>(recv_total);
  if (!be.ok()) return 1;
  R.npeers = 0;
  R.send_ptr[0] = R.recv_ptr[0] = 0;
  int64_t ghost_start = R.n_own;
  std::vector<int32_t> recv_host((size_t)recv_total);
  for (int p = 0; p < nranks; ++p) {
    if (p == rank) continue;
    const int64_t first_ghost = ghost_start;
    ghost_start += n_recv[p];
    if (n_send[p] == 0 && n_recv[p] == 0) continue;
    const int k = R.npeers++;
    R.peers[k] = p;
    if (n_send[p] > 0) {
      PART_LAUNCH(be, nn, k_part_send_flag, send_mask, own_flag, nn, p, sflag);
      be.exclusive_sum(sflag, spos, nn);
      PART_LAUNCH(be, nn, k_part_send_list, sflag, spos, nn, g2l, R.send_nodes + R.send_ptr[k]);
    }
    for (int64_t j = 0; j < n_recv[p]; ++j) recv_host[(size_t)(R.recv_ptr[k] + j)] = (int32_t)(first_ghost + j);
    R.send_ptr[k + 1] = R.send_ptr[k] + n_send[p];
    R.recv_ptr[k + 1] = R.recv_ptr[k] + n_recv[p];
  }
  if (recv_total > 0) be.upload(R.recv_nodes, recv_host.data(), recv_total);
  return be.ok() ? 0 : 1;
}
