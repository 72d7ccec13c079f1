// Device partitioner: femcy_partition computes, on this rank's GPU, the piece of a global mesh
// the rank works on; femcy_partition_get hands the arrays to the host layer (femcy_b200/partition.py), which feeds them to
// femcy_set_mesh / femcy_set_halo exactly like the arrays of its NumPy statement of the same scheme.  Kernels and the
// orchestration: partition_kernels.cuh (shared with the CPU SIMT emulation); here: the CUB backend and the C-ABI.
#include <cub/cub.cuh>

#include "ctx.cuh"

struct CubBackend;
#define PART_LAUNCH(be, n, kernel, ...)                                                              \
  do {                                                                                              \
    if ((be).ok() && (n) > 0) {                                                                     \
      kernel<<<(be).grid(n), 256, 0, (be).stream>>>(__VA_ARGS__);                                   \
      (be).after_launch();                                                                          \
    }                                                                                               \
  } while (0)
#include "partition_kernels.cuh"

struct CubBackend {
  femcy_ctx* ctx;
  cudaStream_t stream;
  std::vector<void*> owned;
  bool good = true;
  cudaError_t err = cudaSuccess;

  explicit CubBackend(femcy_ctx* c) : ctx(c), stream(c->stream) {}
  ~CubBackend() { for (void* p : owned) cudaFree(p); }
  bool ok() const { return good; }
  void check(cudaError_t e) { if (e != cudaSuccess && good) { good = false; err = e; } }
  int grid(int64_t n) const { int64_t g = ceil_div64(n > 0 ? n : 1, 256); return (int)(g > 148 * 16 ? 148 * 16 : g); }
  void after_launch() { ctx->launches++; check(cudaGetLastError()); }
  template <class T> T* alloc(int64_t n) {
    if (!good) return nullptr;
    void* p = nullptr;
    const size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(T);
    check(cudaMalloc(&p, bytes));
    if (!good) return nullptr;
    owned.push_back(p);
    check(cudaMemsetAsync(p, 0, bytes, stream));
    return (T*)p;
  }
  template <class T> void release(T* p) {        // hand a result array over to the caller (no longer freed by the backend)
    for (auto& q : owned) if (q == (void*)p) q = nullptr;
  }
  template <class T> T read(const T* p) {
    T v = T();
    if (!good) return v;
    check(cudaMemcpyAsync(&v, p, sizeof(T), cudaMemcpyDeviceToHost, stream));
    check(cudaStreamSynchronize(stream));
    return v;
  }
  template <class T> void upload(T* dst, const T* src, int64_t n) {
    if (!good || n <= 0) return;
    check(cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, stream));
    check(cudaStreamSynchronize(stream));
  }
  void sort_pairs(uint64_t* kin, uint64_t* kout, uint32_t* vin, uint32_t* vout, int64_t n) {
    if (!good || n <= 0) return;
    size_t tb = 0;
    void* tmp = nullptr;
    check(cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, kout, vin, vout, n, 0, 64, stream));
    check(cudaMalloc(&tmp, tb + 16));
    if (good) check(cub::DeviceRadixSort::SortPairs(tmp, tb, kin, kout, vin, vout, n, 0, 64, stream));
    check(cudaStreamSynchronize(stream));
    if (tmp) cudaFree(tmp);
    ctx->launches += 16;
  }
  void sort_keys(uint64_t* kin, uint64_t* kout, int64_t n) {
    if (!good || n <= 0) return;
    size_t tb = 0;
    void* tmp = nullptr;
    check(cub::DeviceRadixSort::SortKeys(nullptr, tb, kin, kout, n, 0, 64, stream));
    check(cudaMalloc(&tmp, tb + 16));
    if (good) check(cub::DeviceRadixSort::SortKeys(tmp, tb, kin, kout, n, 0, 64, stream));
    check(cudaStreamSynchronize(stream));
    if (tmp) cudaFree(tmp);
    ctx->launches += 16;
  }
  void exclusive_sum(const int32_t* in, int32_t* out, int64_t n) {
    if (!good || n <= 0) return;
    size_t tb = 0;
    void* tmp = nullptr;
    check(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, stream));
    check(cudaMalloc(&tmp, tb + 16));
    if (good) check(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n, stream));
    check(cudaStreamSynchronize(stream));
    if (tmp) cudaFree(tmp);
    ctx->launches += 2;
  }
};

struct PartitionState {
  PartitionResult R;
  int n_en = 0, dm = 0;
  void free_all() {
    femcy_free(&R.owner); femcy_free(&R.elem_ids); femcy_free(&R.primary); femcy_free(&R.l2g); femcy_free(&R.loc_elems);
    femcy_free(&R.loc_nodes); femcy_free(&R.send_nodes); femcy_free(&R.recv_nodes);
    R = PartitionResult();
  }
};

void femcy_partition_free(femcy_ctx* ctx) {
  if (!ctx->partition) return;
  PartitionState* S = (PartitionState*)ctx->partition;
  S->free_all();
  delete S;
  ctx->partition = nullptr;
}

// sizes_out[6] = n_own, n_local, ne_local, npeers, total send nodes, total recv nodes
extern "C" int femcy_partition(femcy_ctx* ctx, int dm, int64_t nn, const double* nodes, int64_t ne, int n_en,
                               const int32_t* elements, int rank, int nranks, int axis, const int64_t* bounds, int64_t* sizes_out) {
  cudaSetDevice(ctx->device);
  if (nranks < 1 || nranks > FEMCY_MAX_RANKS || rank < 0 || rank >= nranks) return femcy_fail_msg(ctx, "femcy_partition: 1 <= nranks <= 8, 0 <= rank < nranks");
  if (dm < 1 || dm > 3 || axis < 0 || axis >= dm || n_en < 1 || nn < 0 || ne < 0) return femcy_fail_msg(ctx, "femcy_partition: bad mesh description");
  if (!nodes || !elements || !bounds) return femcy_fail_msg(ctx, "femcy_partition: null array");
  if (nn >= ((int64_t)1 << 31) || ne * n_en >= ((int64_t)1 << 40)) return femcy_fail_msg(ctx, "femcy_partition: mesh too large");
  if (bounds[0] != 0 || bounds[nranks] != nn) return femcy_fail_msg(ctx, "femcy_partition: bounds must run from 0 to nn");
  for (int r = 0; r < nranks; ++r) if (bounds[r] > bounds[r + 1]) return femcy_fail_msg(ctx, "femcy_partition: bounds must ascend");
  if (!ctx->partition) ctx->partition = new PartitionState();
  PartitionState* S = (PartitionState*)ctx->partition;
  S->free_all();
  CubBackend be(ctx);
  double* d_nodes = be.alloc<double>(nn * dm);
  int32_t* d_elems = be.alloc<int32_t>(ne * n_en);
  be.upload(d_nodes, nodes, nn * dm);
  be.upload(d_elems, elements, ne * n_en);
  int rc = be.ok() ? partition_build(be, dm, nn, d_nodes, ne, n_en, d_elems, rank, nranks, axis, bounds, S->R) : 1;
  be.check(cudaStreamSynchronize(ctx->stream));
  if (rc || !be.ok()) {
    S->R = PartitionResult();                      // the arrays belong to the backend and die with it
    return femcy_fail(ctx, "femcy_partition", be.err, __FILE__, __LINE__);
  }
  be.release(S->R.owner); be.release(S->R.elem_ids); be.release(S->R.primary); be.release(S->R.l2g); be.release(S->R.loc_elems);
  be.release(S->R.loc_nodes); be.release(S->R.send_nodes); be.release(S->R.recv_nodes);
  S->n_en = n_en; S->dm = dm;
  if (sizes_out) {
    sizes_out[0] = S->R.n_own; sizes_out[1] = S->R.n_local; sizes_out[2] = S->R.ne_local; sizes_out[3] = S->R.npeers;
    sizes_out[4] = S->R.send_ptr[S->R.npeers]; sizes_out[5] = S->R.recv_ptr[S->R.npeers];
  }
  return 0;
}

extern "C" int femcy_partition_get(femcy_ctx* ctx, int32_t* owner, int64_t* elem_ids, unsigned char* elem_primary,
                                   int64_t* local_to_global, int32_t* local_elements, double* local_nodes, int32_t* peers,
                                   int64_t* send_ptr, int32_t* send_nodes, int64_t* recv_ptr, int32_t* recv_nodes) {
  cudaSetDevice(ctx->device);
  PartitionState* S = (PartitionState*)ctx->partition;
  if (!S || !S->R.owner) return femcy_fail_msg(ctx, "femcy_partition first");
  const PartitionResult& R = S->R;
  cudaStream_t st = ctx->stream;
#define GET(dst, src, n) do { if ((dst) && (n) > 0) CK(cudaMemcpyAsync((dst), (src), sizeof(*(dst)) * (size_t)(n), cudaMemcpyDeviceToHost, st)); } while (0)
  GET(owner, R.owner, R.nn);
  GET(elem_ids, R.elem_ids, R.ne_local);
  GET(elem_primary, R.primary, R.ne_local);
  GET(local_to_global, R.l2g, R.n_local);
  GET(local_elements, R.loc_elems, R.ne_local * S->n_en);
  GET(local_nodes, R.loc_nodes, R.n_local * S->dm);
  GET(send_nodes, R.send_nodes, R.send_ptr[R.npeers]);
  GET(recv_nodes, R.recv_nodes, R.recv_ptr[R.npeers]);
#undef GET
  CK(cudaStreamSynchronize(st));
  for (int k = 0; k < R.npeers; ++k) if (peers) peers[k] = R.peers[k];
  for (int k = 0; k <= R.npeers; ++k) { if (send_ptr) send_ptr[k] = R.send_ptr[k]; if (recv_ptr) recv_ptr[k] = R.recv_ptr[k]; }
  return 0;
}
