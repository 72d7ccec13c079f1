// ctx lifetime, mesh / plugin upload, named vectors and small vector kernels.
#include <string.h>

#include "ctx.cuh"
#include "elem_math.cuh"

int femcy_fail(femcy_ctx* ctx, const char* what, cudaError_t e, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  if (ctx) ctx->err = buf;
  return 1;
}
int femcy_fail_msg(femcy_ctx* ctx, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return 1;
}

extern "C" const char* femcy_version(void) { return "femcy_b200 0.1 (sm_100a)"; }

extern "C" int femcy_create(int device, femcy_ctx** out) {
  if (!out) return 1;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return 2;  // no CUDA device: the product path fails loudly
  if (device < 0 || device >= ndev) return 3;
  femcy_ctx* ctx = new femcy_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return 4; }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return 5; }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  cudaEventCreate(&ctx->evA0);
  cudaEventCreate(&ctx->evA1);
  if (cudaMalloc((void**)&ctx->scal, 64 * sizeof(double)) != cudaSuccess) { delete ctx; return 6; }
  cudaMemset(ctx->scal, 0, 64 * sizeof(double));
  cudaMallocHost((void**)&ctx->h_scal, 64 * sizeof(double));
  cudaMalloc((void**)&ctx->red_ticket, 8 * sizeof(unsigned int));
  cudaMemset(ctx->red_ticket, 0, 8 * sizeof(unsigned int));
  memset(&ctx->tab, 0, sizeof(ctx->tab));
  *out = ctx;
  return 0;
}

extern "C" void femcy_destroy(femcy_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  femcy_drop_graph(ctx);
  femcy_comm_free(ctx);
  femcy_precond_free(ctx);
  femcy_topology_free(ctx);
  femcy_partition_free(ctx);
  femcy_pattern_free(ctx);
  femcy_sections_free(ctx);
  femcy_free(&ctx->nodes); femcy_free(&ctx->elems);
  for (int i = 0; i < FEMCY_VEC_COUNT; ++i) femcy_free(&ctx->vec[i]);
  femcy_free(&ctx->vol); femcy_free(&ctx->dsdx); femcy_free(&ctx->F); femcy_free(&ctx->cauchy);
  femcy_free(&ctx->mises); femcy_free(&ctx->strain); femcy_free(&ctx->energy); femcy_free(&ctx->egeo4);
  femcy_free(&ctx->red_partials); femcy_free(&ctx->red_ticket); femcy_free(&ctx->scal);
  femcy_free(&ctx->bc_nodes); femcy_free(&ctx->bc_comps); femcy_free(&ctx->bc_vals);
  femcy_free(&ctx->bc_flag); femcy_free(&ctx->bc_val_full);
  if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->evA0) cudaEventDestroy(ctx->evA0);
  if (ctx->evA1) cudaEventDestroy(ctx->evA1);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* femcy_last_error(femcy_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

extern "C" int femcy_set_stream(femcy_ctx* ctx, void* s) {
  cudaSetDevice(ctx->device);
  if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  if (s == nullptr) {
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  } else {
    ctx->stream = (cudaStream_t)s;
    ctx->own_stream = false;
  }
  return 0;
}

extern "C" int femcy_sync(femcy_ctx* ctx) {
  cudaSetDevice(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int64_t femcy_device_bytes(femcy_ctx* ctx) {
  cudaSetDevice(ctx->device);
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return -1;
  return (int64_t)(tot - fr);  // device-wide bytes in use (this process is the only tenant of the GPU)
}
extern "C" int64_t femcy_launch_count(femcy_ctx* ctx) { return ctx->launches; }

extern "C" int femcy_last_time_ms(femcy_ctx* ctx, int kind, double* ms_out) {
  if (kind >= 4 && kind <= 6) { *ms_out = ctx->prof_ms[kind - 4]; return 0; }
  if (kind < 0 || kind > 3) return femcy_fail_msg(ctx, "bad timing kind");
  if (kind == 0) {
    cudaSetDevice(ctx->device);
    float ms = 0;
    if (cudaEventSynchronize(ctx->evA1) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->evA0, ctx->evA1) == cudaSuccess)
      ctx->last_ms[0] = ms;
  }
  *ms_out = ctx->last_ms[kind];
  return 0;
}

static bool supported_shape(int dm, int n_en) {
  return (dm == 2 && (n_en == 3 || n_en == 4 || n_en == 6 || n_en == 8)) || (dm == 3 && (n_en == 4 || n_en == 10)) ||
         (dm == 1 && n_en == 1);
}

extern "C" int femcy_set_mesh(femcy_ctx* ctx, int dm, int64_t nn, int64_t nn_own, const double* nodes, int64_t ne,
                              int n_en, const int32_t* elements) {
  cudaSetDevice(ctx->device);
  if (!supported_shape(dm, n_en)) return femcy_fail_msg(ctx, "unsupported (dm, n_en) element shape");
  if (nn_own < 0 || nn_own > nn) return femcy_fail_msg(ctx, "nn_own out of range");
  if (nn * dm >= (int64_t)1 << 31) return femcy_fail_msg(ctx, "too many dofs for int32 indexing");
  femcy_pattern_free(ctx);
  femcy_topology_free(ctx);        // facet tables and boundary facets belong to the old mesh
  femcy_sections_free(ctx);        // back to one section; the selected section's arrays stay with the ctx and are re-used below
  ctx->dm = dm; ctx->nn = nn; ctx->nn_own = nn_own; ctx->ne = ne; ctx->n_en = n_en;
  ctx->n_v = (dm == 2) ? 3 : (dm == 3 ? 6 : 1);
  if (femcy_alloc(ctx, &ctx->nodes, nn * dm)) return 1;
  if (femcy_alloc(ctx, &ctx->elems, ne * n_en)) return 1;
  if (nodes) CK(cudaMemcpyAsync(ctx->nodes, nodes, (size_t)nn * dm * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (elements) CK(cudaMemcpyAsync(ctx->elems, elements, (size_t)ne * n_en * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_elem = false;
  femcy_pattern_free(ctx);
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Row f4: sections (FemcySection, ctx.cuh).  The reference takes one element kind and the first material only
// (/root/reference/reader/inp_info.py:125-128, main.py:24).
void femcy_section_park(femcy_ctx* ctx) {
  if (ctx->sections.empty()) return;
  FemcySection& S = ctx->sections[ctx->cur_section];
  S.n_en = ctx->n_en; S.n_gp = ctx->n_gp; S.ne = ctx->ne; S.elems = ctx->elems; S.tab = ctx->tab;
  S.have_elem = ctx->have_elem; S.have_mat = ctx->have_mat; S.mat_kind = ctx->mat_kind; S.elem_slot = ctx->elem_slot;
  S.vol = ctx->vol; S.dsdx = ctx->dsdx; S.F = ctx->F; S.cauchy = ctx->cauchy; S.mises = ctx->mises; S.strain = ctx->strain;
  S.energy = ctx->energy;
}
void femcy_section_load(femcy_ctx* ctx, int s) {
  const FemcySection& S = ctx->sections[s];
  ctx->n_en = S.n_en; ctx->n_gp = S.n_gp; ctx->ne = S.ne; ctx->elems = S.elems; ctx->tab = S.tab;
  ctx->have_elem = S.have_elem; ctx->have_mat = S.have_mat; ctx->mat_kind = S.mat_kind; ctx->elem_slot = S.elem_slot;
  ctx->vol = S.vol; ctx->dsdx = S.dsdx; ctx->F = S.F; ctx->cauchy = S.cauchy; ctx->mises = S.mises; ctx->strain = S.strain;
  ctx->energy = S.energy;
  ctx->cur_section = s;
}
// drop every section but the selected one, whose arrays stay in the ctx fields (single-section state again)
void femcy_sections_free(femcy_ctx* ctx) {
  for (int s = 0; s < (int)ctx->sections.size(); ++s) {
    if (s == ctx->cur_section) continue;
    FemcySection& S = ctx->sections[s];
    femcy_free(&S.elems); femcy_free(&S.elem_slot); femcy_free(&S.vol); femcy_free(&S.dsdx); femcy_free(&S.F);
    femcy_free(&S.cauchy); femcy_free(&S.mises); femcy_free(&S.strain); femcy_free(&S.energy);
  }
  ctx->sections.clear();
  ctx->cur_section = 0;
}

extern "C" int femcy_add_section(femcy_ctx* ctx, int64_t ne, int n_en, const int32_t* elements, int* section_out) {
  cudaSetDevice(ctx->device);
  if (ctx->dm == 0) return femcy_fail_msg(ctx, "set_mesh first");
  if (!supported_shape(ctx->dm, n_en) || ctx->dm == 1) return femcy_fail_msg(ctx, "unsupported (dm, n_en) element shape");
  if (ne < 0) return femcy_fail_msg(ctx, "femcy_add_section: negative element count");
  if (ctx->comm || ctx->nn_own != ctx->nn) return femcy_fail_msg(ctx, "femcy_add_section: a mesh of several sections runs on one GPU (no partition)");
  femcy_pattern_free(ctx);
  if (ctx->sections.empty()) { ctx->sections.emplace_back(); ctx->cur_section = 0; }   // the mesh of femcy_set_mesh is section 0
  femcy_section_park(ctx);
  FemcySection S;
  memset(&S.tab, 0, sizeof(S.tab));
  S.n_en = n_en; S.ne = ne;
  ctx->sections.push_back(S);
  femcy_section_load(ctx, (int)ctx->sections.size() - 1);
  if (femcy_alloc(ctx, &ctx->elems, ne * n_en)) return 1;
  if (elements && ne > 0) {
    CK(cudaMemcpyAsync(ctx->elems, elements, (size_t)ne * n_en * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  if (section_out) *section_out = ctx->cur_section;
  return 0;
}

extern "C" int femcy_select_section(femcy_ctx* ctx, int section) {
  const int n = ctx->sections.empty() ? 1 : (int)ctx->sections.size();
  if (section < 0 || section >= n) return femcy_fail_msg(ctx, "femcy_select_section: no such section");
  if (ctx->sections.empty() || section == ctx->cur_section) return 0;
  femcy_section_park(ctx);
  femcy_section_load(ctx, section);
  return 0;
}

extern "C" int femcy_section_count(femcy_ctx* ctx) { return ctx->sections.empty() ? 1 : (int)ctx->sections.size(); }

extern "C" int femcy_set_element(femcy_ctx* ctx, int n_gp, const double* dNdxi, const double* weights) {
  cudaSetDevice(ctx->device);
  if (ctx->dm == 0) return femcy_fail_msg(ctx, "set_mesh first");
  if (n_gp < 1 || n_gp > FEMCY_MAX_GP) return femcy_fail_msg(ctx, "n_gp out of range");
  ctx->n_gp = n_gp;
  for (int g = 0; g < n_gp; ++g) {
    for (int a = 0; a < ctx->n_en; ++a)
      for (int k = 0; k < ctx->dm; ++k)
        ctx->tab.dN[(g * ctx->n_en + a) * ctx->dm + k] = dNdxi[(g * ctx->n_en + a) * ctx->dm + k];
    ctx->tab.w[g] = weights[g];
  }
  ctx->have_elem = true;
  // the named vectors belong to the mesh: (re)allocated with section 0 (or with whichever section comes first)
  if (ctx->cur_section == 0 || !ctx->vec[0]) return femcy_alloc_state(ctx);
  return femcy_alloc_gp_state(ctx);
}

extern "C" int femcy_set_material(femcy_ctx* ctx, int mat_kind, const double* params, int nparams, const double* C,
                                  int n_v) {
  if (ctx->dm == 0) return femcy_fail_msg(ctx, "set_mesh first");
  if (n_v != ctx->n_v) return femcy_fail_msg(ctx, "C has the wrong Voigt size for this mesh dimension");
  if (mat_kind < 0 || mat_kind > 3) return femcy_fail_msg(ctx, "unknown material kind");
  if ((mat_kind == 0 || mat_kind == 3) && ctx->dm != 3) return femcy_fail_msg(ctx, "3-D material on a 2-D mesh");
  if ((mat_kind == 1 || mat_kind == 2) && ctx->dm != 2) return femcy_fail_msg(ctx, "2-D material on a 3-D mesh");
  for (int i = 0; i < n_v * n_v; ++i) ctx->tab.C[i] = C[i];
  for (int i = 0; i < 4; ++i) ctx->tab.mat[i] = (i < nparams) ? params[i] : 0.0;
  ctx->mat_kind = mat_kind;
  ctx->have_mat = true;
  return 0;
}

void femcy_drop_graph(femcy_ctx* ctx) {
  if (ctx->cg_graph_exec) { cudaGraphExecDestroy(ctx->cg_graph_exec); ctx->cg_graph_exec = nullptr; }
  ctx->cg_graph_chunk = 0;
}

int femcy_alloc_state(femcy_ctx* ctx) {
  femcy_drop_graph(ctx);   // the captured CG graph holds the old vector addresses
  int64_t N = ctx->nn * ctx->dm;
  for (int i = 0; i < FEMCY_VEC_COUNT; ++i) {
    if (femcy_alloc(ctx, &ctx->vec[i], N)) return 1;
    CK(cudaMemsetAsync(ctx->vec[i], 0, (size_t)N * sizeof(double), ctx->stream));
  }
  if (femcy_alloc(ctx, &ctx->bc_flag, N)) return 1;
  if (femcy_alloc(ctx, &ctx->bc_val_full, N)) return 1;
  CK(cudaMemsetAsync(ctx->bc_flag, 0, (size_t)N, ctx->stream));
  return femcy_alloc_gp_state(ctx);
}

// per-Gauss-point arrays of the selected section
int femcy_alloc_gp_state(femcy_ctx* ctx) {
  int64_t ngp = ctx->ne * ctx->n_gp, dd = ctx->dm * ctx->dm;
  if (femcy_alloc(ctx, &ctx->vol, ngp)) return 1;
  if (femcy_alloc(ctx, &ctx->mises, ngp)) return 1;
  if (femcy_alloc(ctx, &ctx->energy, ngp)) return 1;
  if (femcy_alloc(ctx, &ctx->F, ngp * dd)) return 1;
  if (femcy_alloc(ctx, &ctx->cauchy, ngp * dd)) return 1;
  CK(cudaMemsetAsync(ctx->vol, 0, (size_t)ngp * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(ctx->mises, 0, (size_t)ngp * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(ctx->energy, 0, (size_t)ngp * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(ctx->F, 0, (size_t)ngp * dd * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(ctx->cauchy, 0, (size_t)ngp * dd * sizeof(double), ctx->stream));
  // dsdx / strain are allocated lazily (only callers that read them pay for them)
  femcy_free(&ctx->dsdx);
  femcy_free(&ctx->strain);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// vectors

static int vec_check(femcy_ctx* ctx, int which) {
  if (which < 0 || which >= FEMCY_VEC_COUNT || !ctx->vec[which]) return femcy_fail_msg(ctx, "bad vector selector (or state not allocated)");
  return 0;
}

extern "C" void* femcy_vec_devptr(femcy_ctx* ctx, int which) {
  if (which < 0 || which >= FEMCY_VEC_COUNT) return nullptr;
  return ctx->vec[which];
}

extern "C" int femcy_vec_set(femcy_ctx* ctx, int which, const double* host, int64_t n) {
  cudaSetDevice(ctx->device);
  if (vec_check(ctx, which)) return 1;
  if (n > ctx->nn * ctx->dm) return femcy_fail_msg(ctx, "vector too long");
  CK(cudaMemcpyAsync(ctx->vec[which], host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int femcy_vec_get(femcy_ctx* ctx, int which, double* host, int64_t n) {
  cudaSetDevice(ctx->device);
  if (vec_check(ctx, which)) return 1;
  if (n > ctx->nn * ctx->dm) return femcy_fail_msg(ctx, "vector too long");
  CK(cudaMemcpyAsync(host, ctx->vec[which], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

__global__ void k_fill(double* __restrict__ v, double a, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) v[i] = a;
}
__global__ void k_lincomb(double* __restrict__ dst, const double* a, double alpha, const double* b, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) dst[i] = a[i] + alpha * b[i];
}
__global__ void k_scale(double* __restrict__ v, double s, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) v[i] *= s;
}

static inline int grid_for(int64_t n, int block = 256, int cap = 148 * 8) {
  int64_t g = ceil_div64(n, block);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int femcy_vec_fill(femcy_ctx* ctx, int which, double value) {
  cudaSetDevice(ctx->device);
  if (vec_check(ctx, which)) return 1;
  int64_t N = ctx->nn * ctx->dm;
  k_fill<<<grid_for(N), 256, 0, ctx->stream>>>(ctx->vec[which], value, N);
  CK_LAUNCH();
  return 0;
}
extern "C" int femcy_vec_copy(femcy_ctx* ctx, int dst, int src) {
  cudaSetDevice(ctx->device);
  if (vec_check(ctx, dst) || vec_check(ctx, src)) return 1;
  CK(cudaMemcpyAsync(ctx->vec[dst], ctx->vec[src], (size_t)ctx->nn * ctx->dm * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
extern "C" int femcy_vec_lincomb(femcy_ctx* ctx, int dst, int a, double alpha, int b) {
  cudaSetDevice(ctx->device);
  if (vec_check(ctx, dst) || vec_check(ctx, a) || vec_check(ctx, b)) return 1;
  int64_t N = ctx->nn * ctx->dm;
  k_lincomb<<<grid_for(N), 256, 0, ctx->stream>>>(ctx->vec[dst], ctx->vec[a], alpha, ctx->vec[b], N);
  CK_LAUNCH();
  return 0;
}
extern "C" int femcy_vec_scale(femcy_ctx* ctx, int which, double s) {
  cudaSetDevice(ctx->device);
  if (vec_check(ctx, which)) return 1;
  int64_t N = ctx->nn * ctx->dm;
  k_scale<<<grid_for(N), 256, 0, ctx->stream>>>(ctx->vec[which], s, N);
  CK_LAUNCH();
  return 0;
}

int femcy_ensure_reduction_scratch(femcy_ctx* ctx, int64_t nblocks) {
  if (nblocks * 4 <= ctx->red_cap) return 0;
  femcy_drop_graph(ctx);
  int64_t cap = nblocks * 4 + 1024;
  if (femcy_alloc(ctx, &ctx->red_partials, cap)) return 1;
  ctx->red_cap = cap;
  return 0;
}

// RMS / max|.| / sum of squares with the deterministic grid reduction (elem_math.cuh)
__global__ void __launch_bounds__(256)
k_norms(const double* __restrict__ v, int64_t n, double* __restrict__ partials, unsigned int* ticket,
        double* __restrict__ out3, double Ntot) {
  double s = 0.0, m = 0.0, nanflag = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double x = v[i];
    s += x * x;
    m = fmax(m, fabs(x));
    if (x != x) nanflag = 1.0;
  }
  double mine[3] = {s, m, nanflag}, tot[3];
  const bool is_max[3] = {false, true, true};
  if (grid_reduce<3>(mine, partials, ticket, tot, is_max)) {
    out3[0] = sqrt(tot[0] / Ntot);
    out3[1] = (tot[2] != 0.0) ? (0.0 / 0.0) : tot[1];   // NaN anywhere -> NaN, like the reference's reductions
    out3[2] = tot[0];
  }
}

extern "C" int femcy_vec_norms(femcy_ctx* ctx, int which, double* out3) {
  cudaSetDevice(ctx->device);
  if (vec_check(ctx, which)) return 1;
  int64_t n = ctx->nn_own * ctx->dm;
  int g = grid_for(n, 256, 148 * 4);
  if (femcy_ensure_reduction_scratch(ctx, g)) return 1;
  k_norms<<<g, 256, 0, ctx->stream>>>(ctx->vec[which], n, ctx->red_partials, ctx->red_ticket + 1, ctx->scal + 40, (double)n);
  CK_LAUNCH();
  CK(cudaMemcpyAsync(ctx->h_scal + 40, ctx->scal + 40, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  out3[0] = ctx->h_scal[40]; out3[1] = ctx->h_scal[41]; out3[2] = ctx->h_scal[42];
  return 0;
}

// ---------------------------------------------------------------------------------------------
// per-GP arrays

static int gp_ptr(femcy_ctx* ctx, int which, double** p, int64_t* count) {
  int64_t ngp = ctx->ne * ctx->n_gp, dd = ctx->dm * ctx->dm;
  switch (which) {
    case FEMCY_GP_VOL: *p = ctx->vol; *count = ngp; break;
    case FEMCY_GP_DSDX: *p = ctx->dsdx; *count = ngp * ctx->n_en * ctx->dm; break;
    case FEMCY_GP_F: *p = ctx->F; *count = ngp * dd; break;
    case FEMCY_GP_CAUCHY: *p = ctx->cauchy; *count = ngp * dd; break;
    case FEMCY_GP_MISES: *p = ctx->mises; *count = ngp; break;
    case FEMCY_GP_STRAIN: *p = ctx->strain; *count = ngp * dd; break;
    case FEMCY_GP_ENERGY: *p = ctx->energy; *count = ngp; break;
    default: return femcy_fail_msg(ctx, "bad gp array selector");
  }
  if (!*p) return femcy_fail_msg(ctx, "gp array not materialised yet");
  return 0;
}
extern "C" int femcy_gp_get(femcy_ctx* ctx, int which, double* host, int64_t n) {
  cudaSetDevice(ctx->device);
  double* p; int64_t c;
  if (gp_ptr(ctx, which, &p, &c)) return 1;
  if (n > c) return femcy_fail_msg(ctx, "gp array read too long");
  CK(cudaMemcpyAsync(host, p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int femcy_gp_set(femcy_ctx* ctx, int which, const double* host, int64_t n) {
  cudaSetDevice(ctx->device);
  double* p; int64_t c;
  if (gp_ptr(ctx, which, &p, &c)) return 1;
  if (n > c) return femcy_fail_msg(ctx, "gp array write too long");
  CK(cudaMemcpyAsync(p, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Switches of the library (formerly environment variables read inside the hot calls).  Unknown keys fail loudly.
extern "C" int femcy_set_option(femcy_ctx* ctx, const char* key, int value) {
  if (!key) return femcy_fail_msg(ctx, "femcy_set_option: null key");
  std::string k(key);
  if (k == "cg_kernel") { if (value < 0 || value > 3) return femcy_fail_msg(ctx, "cg_kernel: 0..3"); ctx->opt.cg_kernel = value; }
  else if (k == "cg_sym") ctx->opt.cg_sym = value != 0;
  else if (k == "cg_profile") ctx->opt.cg_profile = value != 0;
  else if (k == "cg_stream_cfg") ctx->opt.cg_stream_cfg = value;
  else if (k == "no_graph") ctx->opt.no_graph = value != 0;
  else if (k == "cg_precond") { if (value < 0 || value > 1) return femcy_fail_msg(ctx, "cg_precond: 0 (Jacobi) or 1 (two-level)"); ctx->opt.cg_precond = value; }
  else if (k == "no_p2p") ctx->opt.no_p2p = value != 0;
  else if (k == "consistent_tangent") ctx->opt.consistent_tangent = value != 0;
  else if (k == "sell_sigma") {
    if (value < -1 || (value > 0 && (value % 32) != 0)) return femcy_fail_msg(ctx, "sell_sigma must be -1 (automatic), 0 or a multiple of 32");
    ctx->opt.sell_sigma = value;
  } else return femcy_fail_msg(ctx, "femcy_set_option: unknown option '" + k + "'");
  return 0;
}

// 1 when the last femcy_cg_solve stopped on a NaN / inf residual (singular or indefinite system), else 0
extern "C" int femcy_cg_breakdown(femcy_ctx* ctx) { return ctx->cg_breakdown ? 1 : 0; }
