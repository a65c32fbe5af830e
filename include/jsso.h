/* jsso.h -- C ABI of the B200-native JaxSSO hot path (libjsso.so).
 *
 * One handle = one frozen finite-element model (the arrays Model.model_ready()
 * produces, JaxSSO/model.py:221-246) on one GPU.  The handle owns everything that
 * is a function of the connectivity (block-CSR pattern, contributor maps, boundary
 * mask) plus the block-CSR values and the PCG work vectors; no allocation happens
 * on the hot path after jsso_create.
 *
 * What each entry point replaces in the reference (paths relative to the reference
 * checkout):
 *   jsso_quad_ke        vmap(Quad.element_K_quad)       JaxSSO/element.py:1190-1241 (1073-1084)
 *   jsso_beam_ke        vmap(BeamCol.element_K_beamcol) JaxSSO/element.py:228-275 (130-139)
 *   jsso_pattern        element_K_*_indices + sort_indices/sum_duplicates
 *                                                       JaxSSO/element.py:141-149, 1097-1106;
 *                                                       JaxSSO/assemblemodel.py:156-162
 *   jsso_assemble*      K_func + K_aug                  JaxSSO/assemblemodel.py:111-163, 196-229
 *   jsso_pcg/jsso_solve sci_sparse_solve / jax_sparse_solve  JaxSSO/solver.py:102-125, 176-210
 *   jsso_adjoint        jax.vjp(f_Ax_b)(lam) + XLA transpose of K_aug/K_func/vmap(element_K_*)
 *                                                       JaxSSO/solver.py:157-166, 239-248
 *   jsso_forward        body of SSO_model.params_u      JaxSSO/SSO_model.py:243-248
 *   jsso_backward       *_sparse_solve_bwd + transposes JaxSSO/solver.py:138-166, 221-248
 *   jsso_quad_area      Quad.A                          JaxSSO/element.py:471-487
 *   jsso_csr_spmv       the examples' dense hat-filter mat-vec B @ z / sens @ B
 *                                                       Examples/Shells_Mannheim_Multihalle_Shape.ipynb cells 10-13
 *   jsso_create_from_bsr / jsso_set_values_host          the (K_aug, f_aug) -> u_aug solver plugin of
 *                                                       Model.select_solver, JaxSSO/model.py:340-356
 * Without a counterpart in the reference (it has no multi-GPU path, no iterative solver and no profiler
 * hooks): jsso_mg_setup / jsso_mg_aggregate / jsso_mg_pattern_lists (multigrid preconditioner), jsso_mg_set_dist /
 * jsso_mg_set_dist_setup / jsso_mg_p2p_* / jsso_mg_dist_counters (its row-range distribution over several GPUs),
 * jsso_set_halo / jsso_p2p_* / jsso_halo_exchange / jsso_nccl_unique_id (partitioned meshes), jsso_gather_rows,
 * jsso_assembly_tasks, jsso_profile* / jsso_profiler_range / jsso_fp64_peak (tests and measurement), and the memory /
 * stream / event helpers at the end of this file.
 *
 * Conventions
 *   - all floating point is FP64, all indices int32 (model.py:284-313).
 *   - dof order per node [ux,uy,uz,rx,ry,rz], global dof = 6*node + k (model.py:197).
 *   - crds (n_node,3) row-major; prop_q (n_quad,5) = t,E,nu,kx_mod,ky_mod;
 *     prop_b (n_beam,6) = E,G,Iy,Iz,J,A (model.py:315-338).
 *   - pointers named *_d are device pointers on the handle's device, *_h host pointers.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are
 *     enqueued on it; calls that return statistics synchronise it before returning.
 *   - block-CSR values are stored per block COLUMN-major: entry (i,j) of block s is
 *     vals[36*s + 6*j + i].
 *   - every function returns 0 on success or a JSSO_ERR_* code; the message is
 *     available from jsso_last_error().  No exception crosses this boundary.
 *   - a handle is not re-entrant: one host thread at a time.
 */
#ifndef JSSO_H_
#define JSSO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  JSSO_OK = 0,
  JSSO_ERR_ARG = 1,             /* bad argument / inconsistent mesh */
  JSSO_ERR_CUDA = 2,            /* a CUDA runtime call or kernel failed */
  JSSO_ERR_NOCONV = 3,          /* PCG hit maxiter, or stagnated at its attainable accuracy, before reaching rtol;
                                   the best iterate is still returned (u, and the gradients computed from it) */
  JSSO_ERR_NAN = 4,             /* NaN/Inf met in the solve */
  JSSO_ERR_BADJAC = 5,          /* non-positive Jacobian determinant in a quad */
  JSSO_ERR_DEGENERATE_BEAM = 6, /* beam exactly parallel to global Y (element.py:92-94) */
  JSSO_ERR_NOT_SPD = 7,         /* diagonal block not positive definite / kx_mod != ky_mod */
  JSSO_ERR_NCCL = 8,
  JSSO_ERR_STATE = 9            /* call order violated (e.g. solve before assemble) */
};

typedef struct jsso_handle jsso_handle;

/* device ordinal for a symbolic-only handle: jsso_create runs the host symbolic pass,
 * jsso_get_sizes / jsso_pattern work, every compute entry point returns JSSO_ERR_STATE. */
#define JSSO_DEVICE_NONE (-1)

typedef struct {
  int32_t n_node;            /* local nodes (owned first, then ghosts) */
  int32_t n_row;             /* owned nodes = block rows of K kept on this GPU; = n_node on one GPU */
  int32_t n_quad;
  const int32_t* cnct_quads; /* host, (n_quad,4): i,j,m,n */
  int32_t n_beam;
  const int32_t* cnct_beams; /* host, (n_beam,2) */
  int32_t n_known;
  const int32_t* known;      /* host, prescribed (zero) dof ids */
  int32_t device;            /* CUDA device ordinal */
} jsso_mesh_desc;

typedef struct {
  int32_t n_node, n_row, n_quad, n_beam;
  int64_t nnzb;     /* stored 6x6 blocks */
  int64_t n_items;  /* (element, a, b) contributions = 16 n_quad + 4 n_beam on one GPU */
  int32_t n_chunk;  /* assembly work chunks */
} jsso_sizes;

typedef struct {
  int32_t iterations;
  int32_t restarts;      /* residual replacements taken */
  int32_t converged;
  int32_t flags;         /* element flags seen in assembly: bit0 bad Jacobian, bit1 degenerate beam, bit2 kx!=ky */
  double relres;         /* final TRUE relative residual |b - A x| / |b| of the block-Jacobi scaled system */
  double relres_recur;   /* recurrence residual at exit */
} jsso_stats;

typedef struct {
  double rtol;           /* relative residual target (scaled system); default 1e-10 */
  int32_t maxiter;       /* default 200000 */
  int32_t check_every;   /* iterations per convergence poll; default 50 */
  int32_t use_x0;        /* 1: `u`/`x` holds an initial guess */
  int32_t compliance;    /* jsso_backward only: 1 => g == f/2, take lam = u/2 (K symmetric) */
  int32_t precond;       /* 0 auto, 1 block-Jacobi CG, 2 smoothed-aggregation multigrid PCG */
  int32_t cheb_degree;   /* Chebyshev smoother degree of the multigrid V-cycle; default 2 */
} jsso_solve_opts;

/* One coarsening step of the multigrid hierarchy: connectivity-only gather lists built on the
 * host (jaxsso_b200/multigrid.py: build_level).  All arrays are host int32. */
typedef struct {
  int32_t n_f, n_c, nnz_p, nnz_ap, nnz_c;
  const int32_t *agg, *p_rowptr, *p_col, *p_own, *ps_ptr, *ps_a, *ps_j;
  const int32_t *apl_ptr, *apl_a, *apl_p;
  const int32_t *c_rowptr, *c_col, *c_diag, *cl_ptr, *cl_p, *cl_ap;
  const int32_t *pt_rowptr, *pt_col, *pt_src, *mem_ptr, *mem;
} jsso_mg_level_desc;

/* ---- lifetime ---------------------------------------------------------------- */
int jsso_create(const jsso_mesh_desc* desc, jsso_handle** out);
/* Solver-plugin compatibility mode: the reference's solver callables take an assembled matrix
 * (model.py:340-356, solver.py:75/103/177).  Create a handle over a given 6x6 block-CSR pattern
 * (sorted columns, diagonal blocks present), then supply values (blocks column-major, no BCs yet)
 * and use jsso_pcg / jsso_spmv as usual.  No element kernels on such a handle. */
int jsso_create_from_bsr(int32_t n_node, const int32_t* rowptr_h, const int32_t* colidx_h, int32_t n_known,
                         const int32_t* known_h, int32_t device, jsso_handle** out);
int jsso_set_values_host(jsso_handle* h, const double* vals_h, int apply_bc);
void jsso_destroy(jsso_handle* h);
const char* jsso_last_error(const jsso_handle* h); /* h may be NULL: last create error */
int jsso_get_sizes(const jsso_handle* h, jsso_sizes* out);

/* ---- symbolic pass results (host copies) -------------------------------------- */
/* rowptr[n_row+1], colidx[nnzb] (sorted within each row): the bit-exact pattern. */
int jsso_pattern(const jsso_handle* h, int32_t* rowptr_h, int32_t* colidx_h);
/* Greedy aggregation step of the multigrid symbolic setup on the host (no device): agg_h[n] receives the
 * aggregate of every node of the block-CSR graph (rowptr_h, colidx_h), *n_agg the number of aggregates. */
int jsso_mg_aggregate(int32_t n, const int32_t* rowptr_h, const int32_t* colidx_h, int32_t* agg_h, int32_t* n_agg);
/* Gather-list construction of the multigrid symbolic setup on the host: m triples (row, col, left, right),
 * emitted in (left, right) order, -> block pattern sorted by (row, col) and per block its (left, right) pairs in
 * input order; output arrays sized for n_blk = m (ptr: m + 1), *n_blk receives the number of blocks. */
int jsso_mg_pattern_lists(int64_t m, const int32_t* row_h, const int32_t* col_h, const int32_t* left_h,
                          const int32_t* right_h, int32_t n_row, int32_t* rowptr_h, int32_t* ocol_h, int32_t* ptr_h,
                          int32_t* left_out_h, int32_t* right_out_h, int64_t* n_blk);
/* Warp-task lists of the two-kernel numeric assembly (host copies, for tests): counts[3] =
 * {n_task, n_task_els, tasks_ok}; every other pointer may be NULL.  task_meta: 4 ints per task
 * {blk0, item0, el0, n_blk | n_item<<8 | n_el<<16}; item_desc per pair item: local block (bits 0-4),
 * b (5-6), a (7-8), local quad (9-11), beam (12), first item of its block (13); blk_bc per block: row
 * mask | col mask<<6 | diagonal<<12; blk_item_ptr[nnzb+1] / item_code[n_items] = (elem<<4)|(a<<2)|b are
 * the contributor lists in the reference's concatenation order (assemblemodel.py:202-211). */
int jsso_assembly_tasks(const jsso_handle* h, int32_t* counts, int32_t* task_meta_h, int32_t* task_els_h,
                        uint16_t* item_desc_h, uint16_t* blk_bc_h, int32_t* blk_item_ptr_h, int32_t* item_code_h);

/* ---- element stiffness, materialised (tests / ncu) ----------------------------- */
/* ke_d: (n_quad,24,24) row-major, the reference's `data` layout (element.py:1236-1237). */
int jsso_quad_ke(jsso_handle* h, const double* crds_d, const double* prop_q_d, double* ke_d, void* stream);
/* ke_d: (n_beam,12,12) row-major (element.py:270-271). */
int jsso_beam_ke(jsso_handle* h, const double* crds_d, const double* prop_b_d, double* ke_d, void* stream);

/* ---- post-processing and the filters either side of the path ------------ */
/* Surface area of every quad (JaxSSO/element.py:471-487); area_d has n_quad entries. */
int jsso_quad_area(jsso_handle* h, const double* crds_d, double* area_d, void* stream);
/* y = A x for a scalar CSR matrix on the current device: the sparse form of the examples' dense
 * hat-filter matrices B_ij (Examples/Shells_Mannheim_Multihalle_Shape.ipynb cells 10-13). */
int jsso_csr_spmv(int32_t n_row, const int32_t* rowptr_d, const int32_t* colidx_d, const double* vals_d,
                  const double* x_d, double* y_d, void* stream);

/* ---- assembly ------------------------------------------------------------------ */
/* Fused Ke + numeric assembly into the handle-owned block-CSR values (K_e never
 * touches HBM).  apply_bc != 0 imposes the prescribed dofs (rows/cols -> identity). */
int jsso_assemble(jsso_handle* h, const double* crds_d, const double* prop_q_d, const double* prop_b_d,
                  int apply_bc, void* stream);
/* Stand-alone segmented reduction of materialised element matrices. */
/* dst_d[i*width + k] = src_d[idx_d[i]*width + k], i < n: node-row gather between numberings on the current
 * device (the local part of a replicated global vector for a partitioned handle). */
int jsso_gather_rows(const double* src_d, const int32_t* idx_d, int32_t n, int32_t width, double* dst_d, void* stream);
/* Per-kernel timing of jsso_assemble's two kernels (CUDA events on the caller's stream; measurement aid for
 * bench.py): jsso_profile(h, 1), then after any jsso_assemble, ms[0] = quad_geometry_kernel, ms[1] =
 * assemble_tasks_kernel of the last call. */
int jsso_profile(jsso_handle* h, int enable);
int jsso_profile_read(jsso_handle* h, float* ms);
int jsso_assemble_from_ke(jsso_handle* h, const double* ke_q_d, const double* ke_b_d, int apply_bc,
                          void* stream);
/* Copy the handle's values (nnzb*36, column-major blocks) to a device / host buffer. */
int jsso_get_values(jsso_handle* h, double* vals_d, void* stream);
int jsso_get_values_host(jsso_handle* h, double* vals_h);
/* Element flags accumulated by the last assembly (see jsso_stats.flags). */
int jsso_get_flags(jsso_handle* h, int32_t* flags_out);

/* ---- linear algebra on the assembled matrix ------------------------------------- */
/* y = K x with the current values (x has 6*n_node entries, y 6*n_row).  A solve scales the stored matrix in place
 * (block-Jacobi: W K W^T); jsso_spmv and jsso_get_values* keep returning the UNSCALED operator afterwards
 * (K x = W^-1 (A^ (W^-T x)), blocks unscaled on the fly; single-GPU handles). */
int jsso_spmv(jsso_handle* h, const double* x_d, double* y_d, void* stream);
/* Block-Jacobi preconditioned CG on the assembled (BC-imposed) matrix:
 * K x = b with b zeroed at prescribed dofs.  Synchronises `stream`. */
int jsso_pcg(jsso_handle* h, const double* b_d, double* x_d, const jsso_solve_opts* opts,
             jsso_stats* stats, void* stream);

/* Optional smoothed-aggregation multigrid preconditioner (single GPU): upload
 * the symbolic hierarchy once; the numeric hierarchy (rigid-body prolongators, Galerkin operators)
 * is rebuilt inside the first solve after each assembly. */
int jsso_mg_setup(jsso_handle* h, int32_t n_levels, const jsso_mg_level_desc* levels);

/* (No counterpart in the reference: it solves K_aug u_aug = f_aug with one SuperLU factorisation on the host,
 * JaxSSO/solver.py:176-210, selected in model.py:340-356; this is how the same u is reached on several GPUs.)
 * Row-range distribution of the multigrid-preconditioned solve over N GPUs (one process per GPU; SURVEY 8(e):
 * "PCG per iteration: halo exchange ... overlapped", 8(f) rank 1).  The handle holds the WHOLE mesh, renumbered
 * so that every rank's nodes are one contiguous range (jaxsso_b200/dist_multigrid.py), with the hierarchy already
 * uploaded; assembly and the numeric multigrid setup stay replicated, the V-cycle and PCG products are computed
 * by row ranges with the halo exchanges described per level below, levels >= n_dist run replicated.  Collective
 * over the communicator of `nccl_id` (jsso_nccl_unique_id on rank 0, distributed by the caller).
 * bounds_h: (n_dist + 1) x (n_rank + 1) int32, row-range bounds of levels 0..n_dist.
 * halo: n_halo entries -- one per distributed level, plus (n_halo = n_dist + 1) the all-gather of the first
 * replicated level written as an exchange plan ("every rank reads every other rank's whole range"), so that over
 * peer memory it needs no NCCL call either. */
typedef struct {
  int32_t n_peer;
  const int32_t *peer_rank;          /* [n_peer] */
  const int32_t *send_ptr, *send_idx; /* [n_peer + 1], node ids of this level this rank owns and the peer reads */
  const int32_t *recv_ptr, *recv_idx; /* [n_peer + 1], node ids the peer owns and this rank's rows read */
  const int32_t *remote_off;          /* [n_peer] or NULL: start of this rank's block in the peer's receive list
                                       * (needed by the peer-memory path only) */
  int32_t n_ghost;                    /* distributed levels: rows of other ranks that this rank's rows of A_l read */
  const int32_t *ghost_rows;          /* [n_ghost] sorted; the fused V-cycle recomputes its corrected iterate on them
                                       * instead of exchanging it (0 / NULL: exchange) */
} jsso_mg_halo_desc;
int jsso_mg_set_dist(jsso_handle* h, const uint8_t nccl_id[128], int32_t rank, int32_t n_rank, int32_t n_dist,
                     const int32_t* bounds_h, int32_t n_halo, const jsso_mg_halo_desc* halo);
/* Optional peer-memory path of the distributed solve (NVLink, CUDA IPC; the distributed CG's counterpart is
 * jsso_p2p_export / jsso_p2p_connect): halo exchanges become a push kernel storing straight into the peers'
 * receive arenas + a wait/unpack kernel, scalar all-reduces a mailbox kernel -- no library call on the iteration
 * path (the two all-gathers per solve / iteration stay on NCCL).  Collective sequence after jsso_mg_set_dist:
 * jsso_mg_p2p_reserve(max over ranks of the largest per-level receive count), jsso_mg_p2p_export (128 bytes per
 * rank), exchange the bytes, jsso_mg_p2p_connect(all bytes, every rank's reserve value). */
int jsso_mg_p2p_reserve(jsso_handle* h, int32_t max_recv_common);
int jsso_mg_p2p_export(jsso_handle* h, uint8_t out[128]);
int jsso_mg_p2p_connect(jsso_handle* h, const uint8_t* all_handles, const int32_t* max_recv_all);
/* Optional: distribute the NUMERIC SETUP of the distributed levels too (after jsso_mg_set_dist; no reference
 * counterpart, see above).  Without it every rank assembles the whole matrix and builds the whole hierarchy
 * redundantly (24 ms at 1M quads: the Amdahl term of an 8-GPU gradient evaluation).  With it a rank assembles,
 * factors and scales only the rows it reads (row hulls of level 0), computes at every distributed level the
 * Galerkin blocks of its own coarse rows and the P / AP blocks they read -- rows of neighbouring ranks are
 * recomputed as ghost rows, no communication -- and the ranks all-gather the coarse matrices (NCCL, by slot ranges).
 * jsso_get_values* / jsso_spmv of such a handle are valid on this rank's rows only; u is still returned whole.
 * One descriptor per distributed level (jaxsso_b200/dist_multigrid.py::setup_plan). */
typedef struct {
  int32_t n_p_slots, n_ap_slots;
  const int32_t *p_slots, *ap_slots;   /* sorted slot ids of the P and AP blocks this rank computes */
  const int32_t* ac_bounds;            /* [n_rank + 1] slot ranges of the coarse matrix owned by the ranks */
  int32_t p_own_lo, p_own_hi;          /* slot range of this rank's prolongation rows P[fs:fe] (FP32 copy) */
  int32_t pt_own_lo, pt_own_hi;        /* slot range of this rank's restriction rows P^T[cs:ce] (transposed here) */
  int32_t scale_row_lo, scale_row_hi;  /* level 0 only: rows whose scaled blocks are read ... */
  int32_t factor_row_lo, factor_row_hi; /* ... and rows whose block-Jacobi factor is needed (assembly hull) */
} jsso_mg_setup_desc;
int jsso_mg_set_dist_setup(jsso_handle* h, int32_t n_dist, const jsso_mg_setup_desc* desc);
/* out[3]: halo exchanges and scalar all-reduces issued by the distributed solve so far, and whether they run over
 * peer memory (1) or NCCL (0). */
int jsso_mg_dist_counters(const jsso_handle* h, int64_t* out);

/* ---- adjoint sensitivity reduction ---------------------------------------------- */
/* d_crds[n,c] = sum_e sum_ab (-lam_e[a] u_e[b]) dK_e[a,b]/dcrds[n,c], same for the
 * element properties.  Any of the three outputs may be NULL. */
int jsso_adjoint(jsso_handle* h, const double* crds_d, const double* prop_q_d, const double* prop_b_d,
                 const double* u_d, const double* lam_d, double* d_crds_d, double* d_prop_q_d,
                 double* d_prop_b_d, void* stream);

/* ---- the drop-in pair (params_u and its VJP) ------------------------------------- */
/* u = K(crds, props)^-1 f with zero prescribed displacements. */
int jsso_forward(jsso_handle* h, const double* crds_d, const double* prop_q_d, const double* prop_b_d,
                 const double* f_d, double* u_d, const jsso_solve_opts* opts, jsso_stats* stats,
                 void* stream);
/* Given g = dL/du: solve K^T lam = g (lam = 0 at prescribed dofs), then reduce the
 * element sensitivities.  Uses the matrix assembled by the preceding jsso_forward
 * (same crds/props).  lam_d may be NULL. */
int jsso_backward(jsso_handle* h, const double* crds_d, const double* prop_q_d, const double* prop_b_d,
                  const double* u_d, const double* g_d, double* d_crds_d, double* d_prop_q_d,
                  double* d_prop_b_d, double* lam_d, const jsso_solve_opts* opts, jsso_stats* stats,
                  void* stream);

/* ---- host-buffer convenience (end-to-end path: H2D, forward, backward, D2H) ------- */
/* Strain-energy objective 0.5 f.u (SSO_model.py:297-301) and its gradient.
 * All pointers are HOST pointers; d_* outputs may be NULL.  Page-locked buffers (jsso_host_alloc_pinned,
 * cudaHostAlloc, cudaHostRegister) are the source / target of the DMA themselves; pageable ones are staged through the
 * handle's pinned buffers with a threaded memcpy (JSSO_HOST_THREADS, default min(8, cores)).  With opts->use_x0 the
 * content of u_h is the initial guess of the solve.  JSSO_ERR_NOCONV still fills every output (best iterate). */
int jsso_value_and_grad_host(jsso_handle* h, const double* crds_h, const double* prop_q_h,
                             const double* prop_b_h, const double* f_h, double* value_out, double* u_h,
                             double* d_crds_h, double* d_prop_q_h, double* d_prop_b_h,
                             const jsso_solve_opts* opts, jsso_stats* fwd_stats, jsso_stats* bwd_stats);

/* Solve-free part of a gradient evaluation with HOST buffers: H2D (crds, props, u, lam),
 * fused Ke + assembly, adjoint reduction, D2H (gradients).  Outputs may be NULL. */
int jsso_assemble_adjoint_host(jsso_handle* h, const double* crds_h, const double* prop_q_h,
                               const double* prop_b_h, const double* u_h, const double* lam_h,
                               double* d_crds_h, double* d_prop_q_h, double* d_prop_b_h);

/* ---- multi-GPU: halo exchange plan + NCCL communicator (one process per GPU) ------- */
/* 128-byte NCCL unique id, generated by rank 0 and distributed by the caller. */
int jsso_nccl_unique_id(uint8_t id_out[128]);
/* n_peer peers; peer p receives send_idx[send_ptr[p]..send_ptr[p+1]) (owned node ids)
 * and fills ghost nodes [recv_start[p], recv_start[p] + recv_count[p]). */
int jsso_set_halo(jsso_handle* h, const uint8_t nccl_id[128], int32_t rank, int32_t n_rank, int32_t n_peer,
                  const int32_t* peer_rank, const int32_t* send_ptr, const int32_t* send_idx,
                  const int32_t* recv_start, const int32_t* recv_count);
/* Optional NVLink peer-memory path (CUDA IPC, one node): after jsso_set_halo every rank exports
 * 128 bytes, the caller all-gathers them (n_rank x 128, rank order) and every rank connects.
 * remote_start[p] = first ghost slot of my nodes in peer p's local numbering.  The CG then does
 * its halo exchange and scalar all-reduces with stores/loads on peer memory inside its kernels. */
int jsso_p2p_export(jsso_handle* h, uint8_t out[128]);
int jsso_p2p_connect(jsso_handle* h, const uint8_t* all_handles, const int32_t* remote_start);
/* Fill the ghost entries of a 6-dof-per-node vector from their owners. */
int jsso_halo_exchange(jsso_handle* h, double* vec_d, void* stream);

/* ---- thin device-memory utilities for bindings without their own CUDA runtime ------ */
int jsso_set_device(int device);
void* jsso_dev_alloc(size_t bytes);
void jsso_dev_free(void* p);
void* jsso_host_alloc_pinned(size_t bytes);
void jsso_host_free_pinned(void* p);
int jsso_memcpy_h2d(void* dst_d, const void* src_h, size_t bytes, void* stream);
int jsso_memcpy_d2h(void* dst_h, const void* src_d, size_t bytes, void* stream);
int jsso_memset(void* dst_d, int value, size_t bytes, void* stream);
int jsso_stream_sync(void* stream);
int jsso_device_count(void);
/* CUDA-event timing on a caller stream (bench.py): returns an opaque event. */
void* jsso_event_create(void);
int jsso_event_record(void* ev, void* stream);
int jsso_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms_out); /* syncs on ev_stop */
void jsso_event_destroy(void* ev);
/* Measured FP64 peak of the device (DFMA chains, 8 per thread, 4 x 256 threads per SM) sustained for about
 * `seconds`: the denominator of the FP64-bound rooflines (quad adjoint, K_e) in bench.py.  No reference
 * counterpart (measurement infrastructure, SURVEY 8(d) "measure the FP64 DFMA peak yourself"). */
int jsso_fp64_peak(int32_t device, double seconds, double* tflops_out, double* seconds_out);
/* cudaProfilerStart (1) / cudaProfilerStop (0): range markers for `ncu --profile-from-start off`. */
int jsso_profiler_range(int start);
/* Number of kernels this library has launched since load (bench `gpu_launches`). */
int64_t jsso_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* JSSO_H_ */
