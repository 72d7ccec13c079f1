// Row f1: device-side mesh topology builders and the Neumann load vector.
//
// Reference code replaced -- Python loops over Python sets / dicts on the host, the dominant wall time of the reference for
// any real mesh:
//   Body.get_boundary                    /root/reference/body.py:197-234          -> femcy_boundary_facets
//   Body.get_nodeEles / nodeEles field   /root/reference/body.py:165-179, stiffnessMtrx.py:70-76   -> femcy_node_elements
//   System_of_equations.neumannBC        /root/reference/stiffnessMtrx.py:369-411  -> femcy_neumann
// Same design as the pattern build (pattern.cu): emit one 64-bit key per item, ONE radix sort (CUB), then short kernels over
// the sorted order.  Device code: topology_kernels.cuh.
#include <cub/cub.cuh>

#include "ctx.cuh"
#include "topology_kernels.cuh"

struct TopologyState {
  FacetTables T;                  // device pointers below
  int32_t* key_nodes = nullptr;
  double *w = nullptr, *normal = nullptr, *N = nullptr, *dN = nullptr;
  int n_en = 0, dm = 0;           // element kind the tables were made for
  int32_t *b_elem = nullptr, *b_kid = nullptr;   // boundary facets of the last femcy_boundary_facets
  int64_t n_boundary = -1;
  int32_t *f_elem = nullptr, *f_kid = nullptr;   // staging of the loaded facets
  int64_t f_cap = 0;
};

static TopologyState* topo(femcy_ctx* ctx) {
  if (!ctx->topology) ctx->topology = new TopologyState();
  return (TopologyState*)ctx->topology;
}

void femcy_topology_free(femcy_ctx* ctx) {
  if (!ctx->topology) return;
  TopologyState* S = (TopologyState*)ctx->topology;
  femcy_free(&S->key_nodes); femcy_free(&S->w); femcy_free(&S->normal); femcy_free(&S->N); femcy_free(&S->dN);
  femcy_free(&S->b_elem); femcy_free(&S->b_kid); femcy_free(&S->f_elem); femcy_free(&S->f_kid);
  delete S;
  ctx->topology = nullptr;
}

static inline int gridt(int64_t n) {
  int64_t g = ceil_div64(n > 0 ? n : 1, 256);
  if (g > 148 * 16) g = 148 * 16;
  return (int)g;
}

extern "C" int femcy_set_facet_tables(femcy_ctx* ctx, int nkeys, int width, int nfp, const int32_t* key_nodes, const double* w,
                                      const double* normals, const double* N, const double* dN) {
  cudaSetDevice(ctx->device);
  if (ctx->dm == 0 || ctx->n_en == 0) return femcy_fail_msg(ctx, "set_mesh first");
  if (nkeys < 1 || nkeys > FEMCY_MAX_FACET_KEYS || width < 2 || width > FEMCY_MAX_FACET_WIDTH || nfp < 1 || nfp > FEMCY_MAX_FACET_POINTS)
    return femcy_fail_msg(ctx, "femcy_set_facet_tables: nkeys <= 8, 2 <= width <= 6, nfp <= 6");
  if (!key_nodes || !w || !normals || !N || !dN) return femcy_fail_msg(ctx, "femcy_set_facet_tables: null table");
  for (int i = 0; i < nkeys * width; ++i)
    if (key_nodes[i] < 0 || key_nodes[i] >= ctx->n_en) return femcy_fail_msg(ctx, "femcy_set_facet_tables: local node out of range");
  TopologyState* S = topo(ctx);
  const int dm = ctx->dm, n_en = ctx->n_en;
  if (femcy_alloc(ctx, &S->key_nodes, nkeys * width) || femcy_alloc(ctx, &S->w, nkeys * nfp) ||
      femcy_alloc(ctx, &S->normal, nkeys * nfp * dm) || femcy_alloc(ctx, &S->N, nkeys * nfp * width) ||
      femcy_alloc(ctx, &S->dN, nkeys * nfp * n_en * dm))
    return 1;
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(S->key_nodes, key_nodes, sizeof(int32_t) * nkeys * width, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(S->w, w, sizeof(double) * nkeys * nfp, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(S->normal, normals, sizeof(double) * nkeys * nfp * dm, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(S->N, N, sizeof(double) * nkeys * nfp * width, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(S->dN, dN, sizeof(double) * nkeys * nfp * n_en * dm, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  S->T.nkeys = nkeys; S->T.width = width; S->T.nfp = nfp;
  S->T.key_nodes = S->key_nodes; S->T.w = S->w; S->T.normal = S->normal; S->T.N = S->N; S->T.dN = S->dN;
  S->n_en = n_en; S->dm = dm;
  S->n_boundary = -1;
  return 0;
}

static int tables_ready(femcy_ctx* ctx, TopologyState** out) {
  TopologyState* S = (TopologyState*)ctx->topology;
  if (!S || S->T.nkeys == 0) return femcy_fail_msg(ctx, "femcy_set_facet_tables first");
  if (S->n_en != ctx->n_en || S->dm != ctx->dm) return femcy_fail_msg(ctx, "the facet tables were made for another element kind");
  if (!ctx->elems) return femcy_fail_msg(ctx, "set_mesh first");
  *out = S;
  return 0;
}

// Body.get_boundary on the device: the facets (element, key) that belong to exactly one element of the selected section,
// in ascending facet id k*ne + e (the order of femcy_b200.body.Body.boundary_arrays)
extern "C" int femcy_boundary_facets(femcy_ctx* ctx, int64_t* count_out) {
  cudaSetDevice(ctx->device);
  TopologyState* S;
  if (tables_ready(ctx, &S)) return 1;
  cudaStream_t st = ctx->stream;
  const int64_t ne = ctx->ne, total = ne * S->T.nkeys;
  if (total >= ((int64_t)1 << 31)) return femcy_fail_msg(ctx, "ne * facets per element exceeds int32 facet ids");
  S->n_boundary = -1;                      // (stays "not built" if anything below fails)
  if (total == 0) { S->n_boundary = 0; if (count_out) *count_out = 0; return 0; }
  uint64_t *keys = nullptr, *keys2 = nullptr; uint32_t *ids = nullptr, *ids2 = nullptr;
  int32_t *flag = nullptr, *pos = nullptr;
  void* tmp = nullptr;
  auto done = [&]() { femcy_free(&keys); femcy_free(&keys2); femcy_free(&ids); femcy_free(&ids2); femcy_free(&flag); femcy_free(&pos); if (tmp) { cudaFree(tmp); tmp = nullptr; } };
#define T_CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { done(); return femcy_fail(ctx, #call, _e, __FILE__, __LINE__); } } while (0)
  if (femcy_alloc(ctx, &keys, total) || femcy_alloc(ctx, &keys2, total) || femcy_alloc(ctx, &ids, total) || femcy_alloc(ctx, &ids2, total) ||
      femcy_alloc(ctx, &flag, total) || femcy_alloc(ctx, &pos, total)) { done(); return 1; }
  k_facet_keys<<<gridt(total), 256, 0, st>>>(ctx->elems, ne, ctx->n_en, ctx->nn, S->T.key_nodes, S->T.nkeys, S->T.width, keys, ids);
  ctx->launches++;
  T_CK(cudaGetLastError());
  int end_bit = 1;
  const uint64_t top = (uint64_t)ctx->nn * (uint64_t)ctx->nn;
  while (end_bit < 64 && (top >> end_bit) != 0) ++end_bit;
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, ids, ids2, total, 0, end_bit, st);
  T_CK(cudaMalloc(&tmp, tb + 16));
  T_CK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, ids, ids2, total, 0, end_bit, st));
  ctx->launches += 8;
  k_facet_unique<<<gridt(total), 256, 0, st>>>(keys2, ids2, total, ctx->elems, ne, ctx->n_en, S->T.key_nodes, S->T.width, flag);
  ctx->launches++;
  T_CK(cudaGetLastError());
  T_CK(cudaStreamSynchronize(st));
  cudaFree(tmp); tmp = nullptr;
  tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, pos, total, st);
  T_CK(cudaMalloc(&tmp, tb + 16));
  T_CK(cub::DeviceScan::ExclusiveSum(tmp, tb, flag, pos, total, st));
  ctx->launches += 2;
  int32_t last_pos = 0, last_flag = 0;
  T_CK(cudaMemcpyAsync(&last_pos, pos + (total - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  T_CK(cudaMemcpyAsync(&last_flag, flag + (total - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  T_CK(cudaStreamSynchronize(st));
  const int64_t nb = (int64_t)last_pos + last_flag;
  if (femcy_alloc(ctx, &S->b_elem, nb) || femcy_alloc(ctx, &S->b_kid, nb)) { done(); return 1; }
  k_facet_compact<<<gridt(total), 256, 0, st>>>(flag, pos, total, ne, S->b_elem, S->b_kid);
  ctx->launches++;
  T_CK(cudaGetLastError());
  T_CK(cudaStreamSynchronize(st));
#undef T_CK
  done();
  S->n_boundary = nb;
  if (count_out) *count_out = nb;
  return 0;
}

extern "C" int femcy_get_boundary_facets(femcy_ctx* ctx, int32_t* elem_out, int32_t* kid_out) {
  cudaSetDevice(ctx->device);
  TopologyState* S = (TopologyState*)ctx->topology;
  if (!S || S->n_boundary < 0) return femcy_fail_msg(ctx, "femcy_boundary_facets first");
  if (S->n_boundary == 0) return 0;
  if (elem_out) CK(cudaMemcpyAsync(elem_out, S->b_elem, sizeof(int32_t) * S->n_boundary, cudaMemcpyDeviceToHost, ctx->stream));
  if (kid_out) CK(cudaMemcpyAsync(kid_out, S->b_kid, sizeof(int32_t) * S->n_boundary, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Body.get_nodeEles on the device: CSR (ptr [nn+1], elements ascending per node) of the selected section
extern "C" int femcy_node_elements(femcy_ctx* ctx, int32_t* ptr_out, int32_t* elems_out) {
  cudaSetDevice(ctx->device);
  if (ctx->dm == 0 || !ctx->elems) return femcy_fail_msg(ctx, "set_mesh first");
  cudaStream_t st = ctx->stream;
  const int64_t ne = ctx->ne, nn = ctx->nn, total = ne * ctx->n_en;
  if (total >= ((int64_t)1 << 31)) return femcy_fail_msg(ctx, "ne * n_en exceeds int32 offsets");
  uint64_t *keys = nullptr, *keys2 = nullptr;
  int32_t *ptr = nullptr, *list = nullptr;
  void* tmp = nullptr;
  auto done = [&]() { femcy_free(&keys); femcy_free(&keys2); femcy_free(&ptr); femcy_free(&list); if (tmp) { cudaFree(tmp); tmp = nullptr; } };
#define T_CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { done(); return femcy_fail(ctx, #call, _e, __FILE__, __LINE__); } } while (0)
  if (femcy_alloc(ctx, &keys, total) || femcy_alloc(ctx, &keys2, total) || femcy_alloc(ctx, &ptr, nn + 1) || femcy_alloc(ctx, &list, total)) { done(); return 1; }
  const uint64_t* sorted = keys2;
  if (total > 0) {
    k_node_elem_keys<<<gridt(total), 256, 0, st>>>(ctx->elems, ne, ctx->n_en, keys);
    ctx->launches++;
    T_CK(cudaGetLastError());
    int end_bit = 1;
    const uint64_t top = (uint64_t)nn * (uint64_t)ne;
    while (end_bit < 64 && (top >> end_bit) != 0) ++end_bit;
    size_t tb = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, keys2, total, 0, end_bit, st);
    T_CK(cudaMalloc(&tmp, tb + 16));
    T_CK(cub::DeviceRadixSort::SortKeys(tmp, tb, keys, keys2, total, 0, end_bit, st));
    ctx->launches += 8;
  }
  k_node_elem_csr<<<gridt(total > nn ? total : nn + 1), 256, 0, st>>>(sorted, total, ne > 0 ? ne : 1, nn, ptr, list);
  ctx->launches++;
  T_CK(cudaGetLastError());
  if (ptr_out) T_CK(cudaMemcpyAsync(ptr_out, ptr, sizeof(int32_t) * (nn + 1), cudaMemcpyDeviceToHost, st));
  if (elems_out && total > 0) T_CK(cudaMemcpyAsync(elems_out, list, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, st));
  T_CK(cudaStreamSynchronize(st));
#undef T_CK
  done();
  return 0;
}

// neumannBC on the device: rhs = consistent nodal loads of `traction` on the listed facets of the selected section
// (zero-filled first, as the reference does at stiffnessMtrx.py:384)
extern "C" int femcy_neumann(femcy_ctx* ctx, int64_t nf, const int32_t* elem, const int32_t* kid, double traction,
                             const double* direction) {
  cudaSetDevice(ctx->device);
  TopologyState* S;
  if (tables_ready(ctx, &S)) return 1;
  if (!ctx->vec[FEMCY_VEC_RHS]) return femcy_fail_msg(ctx, "set_element first");
  if (nf < 0 || (nf > 0 && (!elem || !kid))) return femcy_fail_msg(ctx, "femcy_neumann: bad facet list");
  for (int64_t f = 0; f < nf; ++f)
    if (elem[f] < 0 || elem[f] >= ctx->ne || kid[f] < 0 || kid[f] >= S->T.nkeys)
      return femcy_fail_msg(ctx, "femcy_neumann: facet (element, key) out of range");
  cudaStream_t st = ctx->stream;
  CK(cudaMemsetAsync(ctx->vec[FEMCY_VEC_RHS], 0, (size_t)ctx->nn * ctx->dm * sizeof(double), st));
  if (nf == 0) return 0;
  if (nf > S->f_cap) {
    if (femcy_alloc(ctx, &S->f_elem, nf) || femcy_alloc(ctx, &S->f_kid, nf)) return 1;
    S->f_cap = nf;
  }
  CK(cudaMemcpyAsync(S->f_elem, elem, sizeof(int32_t) * nf, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(S->f_kid, kid, sizeof(int32_t) * nf, cudaMemcpyHostToDevice, st));
  const int has_dir = direction ? 1 : 0;
  const double d0 = direction ? direction[0] : 0.0, d1 = direction ? direction[1] : 0.0;
  const double d2 = (direction && ctx->dm == 3) ? direction[2] : 0.0;
  const int grid = (int)ceil_div64(nf, 128) > 148 * 16 ? 148 * 16 : (int)ceil_div64(nf, 128);
  if (ctx->dm == 2)
    k_neumann<2><<<grid, 128, 0, st>>>(S->T, S->f_elem, S->f_kid, nf, ctx->elems, ctx->n_en, ctx->nodes, traction, has_dir, d0, d1, d2, ctx->vec[FEMCY_VEC_RHS]);
  else
    k_neumann<3><<<grid, 128, 0, st>>>(S->T, S->f_elem, S->f_kid, nf, ctx->elems, ctx->n_en, ctx->nodes, traction, has_dir, d0, d1, d2, ctx->vec[FEMCY_VEC_RHS]);
  CK_LAUNCH();
  CK(cudaStreamSynchronize(st));       // the host lists are borrowed for the duration of the call only
  return 0;
}
