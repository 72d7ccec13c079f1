/* femcy_b200 -- C-ABI of the B200-native FEMcy hot path (K assembly + Jacobi-PCG).
 *
 * The reference (mo-hanxuan/FEMcy) has no FFI: its boundary is a Python object surface whose
 * hot methods are Taichi kernels.  Each entry point below replaces one of those kernels /
 * methods 1:1 (reference file:line cited per function); the Python modules under femcy_b200/ keep the reference's
 * class / method names and calls these through ctypes.
 *
 * Conventions: every function returns 0 on success, non-zero on failure;
 * femcy_last_error(ctx) gives the message.  An opaque femcy_ctx owns all device memory of one
 * GPU/rank plus (unless femcy_set_stream is used) one CUDA stream.  Host pointers are borrowed
 * for the duration of the call only.  Not thread-safe per ctx.  All reals fp64, indices int32
 * unless stated.  Functions are asynchronous on the ctx stream unless they return data to host
 * memory (those synchronise the stream before returning).
 */
#ifndef FEMCY_B200_H
#define FEMCY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct femcy_ctx femcy_ctx;

/* ---- lifetime ------------------------------------------------------------------------- */
/* replaces ti.init(arch=ti.cuda, default_fp=ti.f64)                       main.py:11       */
int femcy_create(int device, femcy_ctx** out);
void femcy_destroy(femcy_ctx* ctx);
const char* femcy_last_error(femcy_ctx* ctx);
const char* femcy_version(void);
/* borrow an external CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); NULL = own */
int femcy_set_stream(femcy_ctx* ctx, void* cuda_stream);
int femcy_sync(femcy_ctx* ctx);
/* bytes of device memory currently owned by the ctx */
int64_t femcy_device_bytes(femcy_ctx* ctx);

/* ---- mesh / plugins ------------------------------------------------------------------- */
/* Body.__init__ fields nodes/elements                                      body.py:13-19   *
 * nn_own <= nn: rows [0,nn_own) are owned by this rank, nodes [nn_own,nn) are ghosts        *
 * (columns only).  Single GPU: nn_own == nn.                                                */
int femcy_set_mesh(femcy_ctx* ctx, int dm, int64_t nn, int64_t nn_own, const double* nodes /*[nn,dm]*/,
                   int64_t ne, int n_en, const int32_t* elements /*[ne,n_en]*/);
/* element_zoo plugin tables: dshape_dnat at the Gauss points + gaussWeights                 *
 * e.g. element_zoo/element_linear_tetrahedral.py:27-30,74-82                                */
int femcy_set_element(femcy_ctx* ctx, int n_gp, const double* dNdxi /*[n_gp,n_en,dm]*/,
                      const double* weights /*[n_gp]*/);
/* material_zoo plugin: constant tangent C (= ddsdde, stiffnessMtrx.py:124-129) + the          *
 * parameters of the constitutive law used for stress recovery.                              *
 * mat_kind: 0 LinearIsotropic(3d) [E,nu]; 1 PlaneStrain [E,nu]; 2 PlaneStress [E,nu];        *
 *           3 NeoHookean [C1,D1]         material_zoo/{linear_isotropic*,neo_hookean}.py     */
int femcy_set_material(femcy_ctx* ctx, int mat_kind, const double* params, int nparams,
                       const double* C /*[n_v,n_v]*/, int n_v);

/* Row f4: a mesh of several SECTIONS -- element sets with their own element kind and material over one node set (Abaqus
 * *Solid Section).  The reference accepts one element kind and uses the first material only
 * (reader/inp_info.py:125-128, main.py:24), so these calls have no reference counterpart; a mesh that never calls
 * femcy_add_section runs exactly the single-section code.
 * femcy_set_mesh defines section 0.  femcy_add_section appends a section over the same nodes (any supported kind of the same
 * dimension) and SELECTS it: femcy_set_element / femcy_set_material / femcy_gp_get / femcy_gp_set / femcy_gp_sum /
 * femcy_extrapolate address the selected section; femcy_build_pattern (union of all couplings), femcy_assemble_K (one
 * zero-fill, then every section scatter-adds with its own tables and tangent), femcy_get_dsdx_and_vol,
 * femcy_deformation_gradient, femcy_constitutive, femcy_strain, femcy_mises, femcy_internal_force and femcy_elastic_energy
 * (sum) work on ALL sections.  Single GPU. */
int femcy_add_section(femcy_ctx* ctx, int64_t ne, int n_en, const int32_t* elements /*[ne,n_en]*/, int* section_out);
int femcy_select_section(femcy_ctx* ctx, int section);
int femcy_section_count(femcy_ctx* ctx);

/* ---- sparsity pattern (a1) ------------------------------------------------------------ */
/* replaces Body.get_nodeEles/get_coElement_nodes + sparseIJ build                            *
 * body.py:165-194, stiffnessMtrx.py:78-107.  Device sort-based build of the node-block      *
 * SELL-32 pattern; *nnz_out = number of scalar non-zeros (= sum over rows of sparseIJ[:,0]). */
int femcy_build_pattern(femcy_ctx* ctx, int64_t* nnz_out);
/* scalar CSR view (sorted columns) for comparison with sparseIJ / scipy                     */
int femcy_get_csr_pattern(femcy_ctx* ctx, int32_t* rowptr /*[N_own+1]*/, int32_t* colidx /*[nnz]*/);
int femcy_get_K_csr_values(femcy_ctx* ctx, double* vals /*[nnz]*/);
int femcy_set_K_csr_values(femcy_ctx* ctx, const double* vals /*[nnz]*/);
/* pattern statistics: out[0]=nnz blocks, out[1]=stored block slots (with SELL padding),       *
 * out[2]=slices, out[3]=max blocks per row                                                    */
int femcy_pattern_stats(femcy_ctx* ctx, int64_t* out4);

/* ---- named device vectors (length N = nn*dm unless noted) ------------------------------ */
enum femcy_vec {
  FEMCY_VEC_DOF = 0,        /* System_of_equations.dof                  stiffnessMtrx.py:34  */
  FEMCY_VEC_RHS = 1,        /* .rhs                                     :33                  */
  FEMCY_VEC_RESIDUAL = 2,   /* .residual_nodal_force                    :56                  */
  FEMCY_VEC_NODAL_FORCE = 3,/* .nodal_force                             :55                  */
  FEMCY_VEC_DU = 4,         /* .du                                      :95                  */
  FEMCY_VEC_DOF_OLD = 5,    /* .dof_old                                 :113                 */
  FEMCY_VEC_X = 6,          /* CG.x                       conjugateGradientSolver.py:21      */
  FEMCY_VEC_R = 7,          /* CG.r                                     :22                  */
  FEMCY_VEC_D = 8,          /* CG.d                                     :23                  */
  FEMCY_VEC_M = 9,          /* CG.M (inverse diagonal)                  :26                  */
  FEMCY_VEC_AD = 10,        /* CG.Ad                                    :29                  */
  FEMCY_VEC_COUNT = 11
};
int femcy_vec_set(femcy_ctx* ctx, int which, const double* host, int64_t n);
int femcy_vec_get(femcy_ctx* ctx, int which, double* host, int64_t n);
int femcy_vec_fill(femcy_ctx* ctx, int which, double value);
int femcy_vec_copy(femcy_ctx* ctx, int dst, int src);                 /* field.copy_from      */
/* dst = a + alpha*b     tiGadgets.c_equals_a_minus_b :6, a_equals_b_plus_c_mul_d :13          */
int femcy_vec_lincomb(femcy_ctx* ctx, int dst, int a, double alpha, int b);
int femcy_vec_scale(femcy_ctx* ctx, int which, double s);             /* tiGadgets.field_multiply :68 */
/* out[0]=sqrt(sum f^2 / N) (tiGadgets.field_norm :29, an RMS), out[1]=max|f| (field_abs_max :20),
 * out[2]=sum f^2; over the OWNED entries only                                                */
int femcy_vec_norms(femcy_ctx* ctx, int which, double* out3);
void* femcy_vec_devptr(femcy_ctx* ctx, int which);                    /* raw device pointer    */

/* ---- per-Gauss-point device arrays ------------------------------------------------------ */
enum femcy_gp_array {
  FEMCY_GP_VOL = 0,      /* vol[ne,n_gp]                                stiffnessMtrx.py:61  */
  FEMCY_GP_DSDX = 1,     /* dsdx[ne,n_gp,n_en,dm]                       :59                  */
  FEMCY_GP_F = 2,        /* F[ne,n_gp,dm,dm]                            :40                  */
  FEMCY_GP_CAUCHY = 3,   /* cauchy_stress[ne,n_gp,dm,dm]                :44                  */
  FEMCY_GP_MISES = 4,    /* mises_stress[ne,n_gp]                       :48                  */
  FEMCY_GP_STRAIN = 5,   /* strain[ne,n_gp,dm,dm]                       :46                  */
  FEMCY_GP_ENERGY = 6    /* elsEngDens[ne,n_gp]                         :50                  */
};
int femcy_gp_get(femcy_ctx* ctx, int which, double* host, int64_t n);
int femcy_gp_set(femcy_ctx* ctx, int which, const double* host, int64_t n);
/* sum over all Gauss points of a scalar array (FEMCY_GP_VOL: volume of the mesh on the configuration of the last   *
 * geometry pass; FEMCY_GP_ENERGY; FEMCY_GP_MISES): an 8-byte step result, fixed fold order.                           */
int femcy_gp_sum(femcy_ctx* ctx, int which, double* total_out);

/* ---- hot path: geometry + assembly (a2, a3) -------------------------------------------- */
/* get_dsdx_and_vol                                               stiffnessMtrx.py:132-150   *
 * materialises dsdx/vol on the current configuration X+dof (needed by internal force and     *
 * by callers who read the fields); femcy_assemble_K does NOT depend on it (fused).            */
int femcy_get_dsdx_and_vol(femcy_ctx* ctx);
/* assemble_stiffnessMtrx (K.fill(0) + Bt.C.B scatter), fused with the geometry pass           *
 *                                                                stiffnessMtrx.py:161-186   *
 * variant: 0 = default (gather), FEMCY_ASSEMBLY_SCATTER = thread / warp per element + fp64 atomic adds into the      *
 *   precomputed slots (the atomic scatter-add formulation), FEMCY_ASSEMBLY_GATHER = two passes, no atomics, no       *
 *   zero-fill, bit-reproducible: per-(element, node, Gauss point) gradient records (TMA tensor store for C3D4), then  *
 *   one thread per stored block over its element list.  Measurements: DESIGN.md section 4.                           */
#define FEMCY_ASSEMBLY_SCATTER 1
#define FEMCY_ASSEMBLY_GATHER 2
int femcy_assemble_K(femcy_ctx* ctx, int variant);

/* ---- boundary conditions (a4) ----------------------------------------------------------- */
/* dirichletBC_linearEquations                                    stiffnessMtrx.py:279-307   *
 * several (node set, comp, val) entries may be batched in one call when no dof repeats.      */
int femcy_dirichlet_linear(femcy_ctx* ctx, const int32_t* nodes, const int32_t* comps,
                           const double* vals, int64_t n);
/* dirichletBC_forNewtonMethod_kernel                             stiffnessMtrx.py:317-341   */
int femcy_dirichlet_newton(femcy_ctx* ctx, const int32_t* nodes, const int32_t* comps, int64_t n);
/* dirichletBC_val                                                stiffnessMtrx.py:357-366   */
int femcy_dirichlet_val(femcy_ctx* ctx, const int32_t* nodes, const int32_t* comps,
                        const double* vals, int64_t n);

/* ---- row f1: mesh topology and the Neumann vector on the device --------------------------- */
/* Facet tables of the selected section's element kind (element_zoo plugin data: facet_natural_coos, facet_point_weights,
 * facet_natural_normals, shapeFunc / dshape_dnat at the facet points; e.g. element_linear_tetrahedral.py:37-60):
 * key_nodes [nkeys,width] local nodes of every facet key, ascending; w [nkeys,nfp]; normals [nkeys,nfp,dm] in natural space;
 * N [nkeys,nfp,width] shape functions of the facet's own nodes; dN [nkeys,nfp,n_en,dm].  nkeys <= 8, width <= 6, nfp <= 6.  */
int femcy_set_facet_tables(femcy_ctx* ctx, int nkeys, int width, int nfp, const int32_t* key_nodes, const double* w,
                           const double* normals, const double* N, const double* dN);
/* Body.get_boundary                                                  body.py:197-234          *
 * the facets that belong to exactly one element (one sort of the facet keys); *count_out = their number; the pairs        *
 * (element, facet key index) come back in ascending key*ne + element order through femcy_get_boundary_facets.             */
int femcy_boundary_facets(femcy_ctx* ctx, int64_t* count_out);
int femcy_get_boundary_facets(femcy_ctx* ctx, int32_t* elem_out, int32_t* kid_out);
/* Body.get_nodeEles / the nodeEles field                             body.py:165-179, stiffnessMtrx.py:70-76             *
 * CSR of the elements around every node: ptr_out [nn+1], elems_out [ne*n_en] (ascending per node).                        */
int femcy_node_elements(femcy_ctx* ctx, int32_t* ptr_out, int32_t* elems_out);
/* neumannBC                                                          stiffnessMtrx.py:369-411                             *
 * rhs = consistent nodal loads of `traction` on nf facets given as (element, facet key index): along `direction` [dm]     *
 * (TRVEC) or, with direction == NULL, along the outward normal (pressure = negative traction).  rhs is zero-filled first  *
 * (:384: only the last *Dsload of a deck acts); loads act on the initial geometry.                                        */
int femcy_neumann(femcy_ctx* ctx, int64_t nf, const int32_t* elem, const int32_t* kid, double traction,
                  const double* direction);

/* ---- post-processing kernels (a8, a9) --------------------------------------------------- */
/* get_deformation_gradient                                       stiffnessMtrx.py:532-556   */
int femcy_deformation_gradient(femcy_ctx* ctx);
/* material.constitutiveOf{Small,Large}Deform on the stored F -> cauchy_stress                 *
 * material_zoo/linear_isotropic.py:35-76 etc.                                                */
int femcy_constitutive(femcy_ctx* ctx, int large_deform);
/* get_strain_{small,large}Deformation                            stiffnessMtrx.py:559-589   */
int femcy_strain(femcy_ctx* ctx, int large_deform);
/* get_mises_stress_{planeStress,planeStrain,3d}                  stiffnessMtrx.py:457-501   */
int femcy_mises(femcy_ctx* ctx);
/* assemble_nodal_force_GN = F -> sigma(large) -> dsdx,vol -> f_int   stiffnessMtrx.py:609-644 */
int femcy_internal_force(femcy_ctx* ctx);
/* get_elasEng: energy density + total                            stiffnessMtrx.py:592-606   */
int femcy_elastic_energy(femcy_ctx* ctx, double* total_out);

/* Row f3: ELE.extrapolate (element_zoo/element_*.py:202-293: nodal = E . Gauss-point values, E [n_en x n_gp] constant per  *
 * element kind) of component `comp` of a per-Gauss-point field (FEMCY_GP_VOL / MISES / ENERGY: comp 0; CAUCHY / F /    *
 * STRAIN: comp = i*dm + j), on the device.  elem_nodal_out [ne*n_en] (the reference's `nodal_vals` field) and/or           *
 * node_mean_out [nn] (mean over the adjacent elements) may be null.                                                       */
int femcy_extrapolate(femcy_ctx* ctx, int which_gp, int comp, const double* E /*[n_en*n_gp]*/, double* elem_nodal_out,
                      double* node_mean_out);

/* ---- Jacobi-PCG (a6, a7) ---------------------------------------------------------------- */
/* ConjugateGradientSolver_rowMajor.re_init + solve      conjugateGradientSolver.py:32-127   *
 * b_sel: FEMCY_VEC_RHS or FEMCY_VEC_RESIDUAL.  Stops at the first iteration with              *
 * max|r| < eps*max|r0| (same rule, :124) or after max_iter.  The test is evaluated on the     *
 * device every iteration; the host polls every check_every iterations and iterations past     *
 * the stopping point are no-ops, so the result equals an every-iteration host test.           *
 * fixed_iters != 0: run exactly max_iter iterations (benchmark mode, no exit).                */
int femcy_cg_solve(femcy_ctx* ctx, int b_sel, double eps, int64_t max_iter, int check_every,
                   int fixed_iters, int64_t* iters_out, double* rmax0_out, double* rmax_out);
/* one SpMV y = K x on named vectors (compute_Ad :53-58)                                       */
int femcy_spmv(femcy_ctx* ctx, int x_sel, int y_sel);
/* drop-in construction from the reference's own ELL arrays (CG ctor :10-19): builds a scalar  *
 * (1x1-block) SELL-32 matrix in a fresh ctx-less state: dm=1, N rows.                         */
int femcy_cg_from_ell(femcy_ctx* ctx, int64_t N, int W, const double* spm /*[N,W]*/,
                      const int32_t* sparseIJ /*[N,W+1]*/);

/* ---- multi-GPU (section 8e) -------------------------------------------------------------- */
/* halo plan: for each peer rank p, the local indices (owned nodes) to send and the local ghost *
 * node indices to receive.  NCCL (ncclSend/ncclRecv, ncclAllGather of per-rank partials summed in *
 * rank order) is the bootstrap / fallback exchange; with the peer-memory path below installed     *
 * (the default on an NVLink box) the CG loop makes no NCCL call at all.                           */
/* Device partitioner: the piece of the GLOBAL mesh (host arrays, borrowed) that rank `rank` of `nranks` works on, computed on *
 * this rank's GPU -- node owners = chunks [bounds[r], bounds[r+1]) of the nodes sorted by their coordinate along `axis` (ties  *
 * by id), local elements = those touching an owned node, local numbering = owned nodes (ascending), then ghosts grouped by     *
 * owner, halo plan per peer in ascending rank.  sizes_out[6] = n_own, n_local, ne_local, npeers, send nodes, recv nodes.       *
 * femcy_partition_get copies the arrays out (null pointers are skipped); they feed femcy_set_mesh / femcy_set_halo.            *
 * (No reference counterpart: the reference is single-device.)                                                                  */
int femcy_partition(femcy_ctx* ctx, int dm, int64_t nn, const double* nodes /*[nn,dm]*/, int64_t ne, int n_en,
                    const int32_t* elements /*[ne,n_en]*/, int rank, int nranks, int axis, const int64_t* bounds /*[nranks+1]*/,
                    int64_t* sizes_out /*[6]*/);
int femcy_partition_get(femcy_ctx* ctx, int32_t* owner /*[nn]*/, int64_t* elem_ids /*[ne_local]*/,
                        unsigned char* elem_primary /*[ne_local]*/, int64_t* local_to_global /*[n_local]*/,
                        int32_t* local_elements /*[ne_local,n_en]*/, double* local_nodes /*[n_local,dm]*/, int32_t* peers,
                        int64_t* send_ptr /*[npeers+1]*/, int32_t* send_nodes, int64_t* recv_ptr /*[npeers+1]*/,
                        int32_t* recv_nodes);
int femcy_comm_init(femcy_ctx* ctx, int rank, int nranks, const void* nccl_unique_id /*128 B*/,
                    const char* nccl_library_path);
int femcy_comm_unique_id(const char* nccl_library_path, void* id_out /*128 B*/);
int femcy_set_halo(femcy_ctx* ctx, int npeers, const int32_t* peer_ranks,
                   const int64_t* send_ptr /*[npeers+1]*/, const int32_t* send_nodes,
                   const int64_t* recv_ptr /*[npeers+1]*/, const int32_t* recv_nodes);
int femcy_halo_exchange(femcy_ctx* ctx, int which_vec);
/* NVLink peer-memory path for the CG loop (replaces the NCCL calls inside the iteration): every    *
 * rank exports cudaIpc handles of {its flag/partial-sum window, its CG direction vector d},          *
 * the host all-gathers them, every rank imports all of them.  remote_start[k]: first index, in      *
 * the numbering of peer k of femcy_set_halo, of the ghost nodes that peer holds for this rank.      *
 * With the path installed the kernels of femcy_cg_solve store boundary values of d and their        *
 * partial dot products straight into the peers' memory and wait on flags (see csrc/cg.cu).          */
int femcy_p2p_export(femcy_ctx* ctx, void* handles_out /*128 B*/);
int femcy_p2p_import(femcy_ctx* ctx, const void* all_handles /*[nranks][128 B]*/, const int64_t* remote_start);

/* ---- instrumentation --------------------------------------------------------------------- */
/* device time (ms) of the most recent call of the given kind, measured with CUDA events on    *
 * the ctx stream.  kind: 0 assemble_K, 1 cg_solve (loop only), 2 pattern build;                *
 * 4/5/6: in-loop average of k_spmv_dot / k_update_xr / k_update_d of the last femcy_cg_solve    *
 * run with option cg_profile = 1 (three-kernel path, plain launches, one event per kernel).     */
int femcy_last_time_ms(femcy_ctx* ctx, int kind, double* ms_out);
/* Phase clock of the persistent PCG kernel during the last femcy_cg_solve: 7 doubles, nanoseconds on the device clock
 * summed over the iterations: SpMV loop | grid barrier + fold | cross-rank exchange | x/r update | barrier + fold |
 * exchange | d update + halo push + barrier.  (No reference counterpart: the reference prints a host-side wall time per
 * solve, conjugateGradientSolver.py:110-123.) */
int femcy_cg_phase_ns(femcy_ctx* ctx, double* out7);
/* Library switches (they are NOT read from the environment inside the hot calls):
 *   cg_kernel      0 auto | 1 three-kernel CUDA graph | 2 persistent, plain loads | 3 persistent, TMA-staged matrix stream
 *   cg_sym         1: the PCG SpMV streams the upper half of the symmetric matrix (fp64 atomics; not bit-reproducible)
 *   cg_profile     1: per-kernel events on the three-kernel path      cg_stream_cfg  ring shape of kernel 3 (A/B)
 *   no_graph, no_p2p, sell_sigma (row order of the next femcy_build_pattern: -1 automatic [default: sigma = 1024 when natural-
 *                  order slices would be > 15 % padding, i.e. quadratic elements], 0 natural, else a multiple of 32)
 *   cg_precond     0 Jacobi (reference) | 1 two-level (see femcy_set_aggregates)
 *   consistent_tangent  1: femcy_assemble_K (variant 0 / scatter) builds the exact linearisation of the internal force -- material
 *                  + geometric stiffness, the tangent differentiated from the constitutive law itself -- instead of the
 *                  reference's constant-C stiffness (ddsdde is never updated: material_zoo/neo_hookean.py:62-64); opt-in, it
 *                  changes the Newton iterates (fewer loops), not the converged solution
 * Unknown keys fail.  (No reference counterpart.) */
int femcy_set_option(femcy_ctx* ctx, const char* key, int value);
/* Row f2 (opt-in, changes the iteration path; the default stays the reference's Jacobi-PCG): with option cg_precond = 1
 * femcy_cg_solve runs PCG with a two-level additive preconditioner -- 2 Chebyshev steps on the Jacobi-scaled operator plus a
 * coarse correction over aggregates x rigid-body modes (P^T K P inverted densely).  femcy_set_aggregates gives the aggregate
 * (0..nagg-1) of every node; (2 or 3) x nagg <= 16384 coarse unknowns.  Single GPU.  (Reference: only Jacobi,
 * conjugateGradientSolver.py:48-51.) */
int femcy_set_aggregates(femcy_ctx* ctx, int64_t nagg, const int32_t* agg_of_node);
/* 1 when the last femcy_cg_solve stopped on a NaN / inf residual (the reference's loop would carry NaN to its
 * iteration bound, conjugateGradientSolver.py:109-127), else 0 */
int femcy_cg_breakdown(femcy_ctx* ctx);
/* number of kernels launched by this ctx since creation                                       */
int64_t femcy_launch_count(femcy_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FEMCY_B200_H */
