// libjsso.so -- C ABI (include/jsso.h) over the sm_100a kernels.
#include "../../include/jsso.h"

#include <cuda_runtime.h>
#ifndef JSSO_EMU
#include <cuda_profiler_api.h>
#endif
#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

#include "jsso_adjoint.cuh"
#include "jsso_assemble.cuh"
#include "jsso_solver.cuh"
#include "jsso_multigrid.cuh"
#include "jsso_symbolic.h"

using namespace jsso;

// NCCL is bound lazily with dlopen/dlsym instead of a link-time dependency: a process that also
// imports PyTorch must end up with ONE libnccl.so.2 (PyTorch bundles a newer one than the system's;
// whichever is loaded first wins, and a link-time dependency here would break `import torch`
// after this library).  RTLD_NOLOAD first => reuse the copy that is already in the process.
namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
bool nccl_load() {
  if (g_nccl.ok) return true;
#ifdef JSSO_EMU   // CPU test harness (tests/emu): rank threads of one process instead of libnccl
  g_nccl.GetUniqueId = emurt::NGetUniqueId; g_nccl.CommInitRank = emurt::NCommInitRank;
  g_nccl.CommDestroy = emurt::NCommDestroy; g_nccl.GroupStart = emurt::NGroupStart; g_nccl.GroupEnd = emurt::NGroupEnd;
  g_nccl.Send = emurt::NSend; g_nccl.Recv = emurt::NRecv; g_nccl.AllReduce = emurt::NAllReduce;
  g_nccl.GetErrorString = emurt::NGetErrorString;
  g_nccl.ok = true;
  return true;
#endif
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return false;
  auto sym = [&](const char* n) { return dlsym(lib, n); };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
  g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.GroupStart &&
              g_nccl.GroupEnd && g_nccl.Send && g_nccl.Recv && g_nccl.AllReduce && g_nccl.GetErrorString;
  return g_nccl.ok;
}
}  // namespace

static std::atomic<long long> g_launches{0};
static thread_local std::string g_create_error;
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

struct HaloPeer {
  int rank;
  int send_off, send_cnt;   // into send_idx (nodes)
  int recv_start, recv_cnt; // ghost node range
};

struct jsso_handle {
  int device = 0;
  Symbolic sym;
  std::string err;
  // device copies of the symbolic data
  int32_t *cnct_q = nullptr, *cnct_b = nullptr;
  int32_t *rowptr = nullptr, *colidx = nullptr, *blk_row = nullptr, *diag_slot = nullptr;
  int32_t *blk_item_ptr = nullptr, *item_code = nullptr;
  uint8_t *item_lel = nullptr, *node_mask = nullptr;
  int32_t *chunk_blk = nullptr, *chunk_el_ptr = nullptr, *chunk_els = nullptr, *blk_perm = nullptr;
  // two-kernel assembly (geometry records + warp tasks); asm_tasks = false: chunked single-kernel path
  int32_t *task_meta = nullptr, *task_els = nullptr;
  uint16_t *item_desc = nullptr, *blk_bc = nullptr;
  double* quad_rec = nullptr;    // n_quad x REC geometry records
  bool asm_tasks = false;
  int task_ctas = 148 * 4;       // persistent grid of assemble_tasks_kernel (resident CTAs)
  int adj_ctas = 148 * 2, adj_ctas_prop = 148 * 2;   // persistent grids of quad_adjoint_kernel<false/true>
  // optional per-kernel timing of jsso_assemble (jsso_profile): events before / between / after its two kernels
  bool prof = false;
  cudaEvent_t ev_prof[3] = {nullptr, nullptr, nullptr};
  int32_t *node_inc_ptr = nullptr, *node_inc = nullptr;
  // numeric state
  double* vals = nullptr;   // nnzb*36, column-major blocks
  double* W = nullptr;      // n_node*36 block-Jacobi factors (ghosts filled by halo exchange)
  double *vb = nullptr, *vx = nullptr, *vr = nullptr, *vp = nullptr, *vq = nullptr;  // PCG vectors
  double *corner_q = nullptr, *corner_b = nullptr;   // per-corner gradient partials
  double *tmp_lam = nullptr, *tmp_g = nullptr;
  CgScalars* sc = nullptr;
  CgScalars* sc_host = nullptr;  // pinned
  double* partials = nullptr;
  unsigned* counters = nullptr;
  int* flags = nullptr;
  int* flags_host = nullptr;     // pinned
  bool assembled = false, assembled_bc = false, scaled = false;
  bool pattern_only = false;     // created by jsso_create_from_bsr: no element kernels
  int red_blocks = 148 * 4;
  int spmv_blocks = 148 * 8;
  int rp_blocks = 148 * 3;       // persistent grid of bsr_spmv_rp_kernel (co-resident CTAs)
  int coop_blocks = 0;          // max co-resident blocks of cg_persistent_kernel (0: unsupported)
  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, n_rank = 1;
  std::vector<HaloPeer> peers;
  int32_t* send_idx = nullptr;
  double *send_buf = nullptr;
  int n_send_nodes = 0;
  // NVLink peer-memory path (CUDA IPC): mailbox + peers' p vectors
  Mailbox* mbox = nullptr;
  P2PCtx* p2p = nullptr;            // device copy of the context; null = NCCL path
  std::vector<void*> ipc_opened;
  unsigned long long halo_seq = 0, red_seq_a = 0, red_seq_b = 0;
  // smoothed-aggregation multigrid (single GPU): symbolic data per coarsening step + numeric state
  struct MgLevel {
    int n_f = 0, n_c = 0, nnz_p = 0, nnz_ap = 0, nnz_c = 0;
    int32_t *agg = nullptr, *p_row = nullptr, *p_rowptr = nullptr, *p_col = nullptr, *p_own = nullptr;
    int32_t *ps_ptr = nullptr, *ps_a = nullptr, *ps_j = nullptr;
    int32_t *apl_ptr = nullptr, *apl_a = nullptr, *apl_p = nullptr;
    int32_t *c_rowptr = nullptr, *c_col = nullptr, *c_diag = nullptr, *cl_ptr = nullptr, *cl_p = nullptr,
            *cl_ap = nullptr;
    int32_t *pt_rowptr = nullptr, *pt_col = nullptr, *pt_src = nullptr, *mem_ptr = nullptr, *mem = nullptr;
    double *P = nullptr, *Pt = nullptr, *AP = nullptr, *Ac = nullptr, *Xc = nullptr, *Dinv = nullptr;
    // every level is kept in block-Jacobi-scaled form (unit diagonal blocks): factors of THIS level's coarse matrix,
    // W_c = L_c^-1 and L_c (row-major 6x6 per coarse node), and the row index of every coarse block
    double *Wc = nullptr, *Lc = nullptr;
    int32_t* c_row = nullptr;
    float *P32 = nullptr, *Pt32 = nullptr, *Ac32 = nullptr;   // single-precision copies used by the V-cycle
    double *b = nullptr, *x = nullptr, *r = nullptr, *d = nullptr;   // level vectors (b, x unused at level 0)
    double lam = 0.0;
  };
  std::vector<MgLevel> mg;
  double* Lfac = nullptr;          // L_i (row-major) of the fine diagonal blocks
  float* vals32 = nullptr;         // single-precision copy of the scaled fine matrix (V-cycle only)
  bool mg_fp32 = true;
  __half* vals16 = nullptr;        // binary16 copy of the scaled fine matrix (V-cycle only; JSSO_MG_FP16=1)
  bool mg_fp16 = false;
  int mg_power_iters = 10;         // steps of the power iteration for lambda_max per level (JSSO_MG_POWER_ITERS)
  double mg_power_safety = 1.2;
  // JSSO_MG_TIMING=2: CUDA events between the steps of ONE PCG iteration (the 4th of a solve), printed afterwards
  struct IterProbe {
    bool armed = false, active = false, done = false;
    std::vector<cudaEvent_t> ev;
    std::vector<std::string> names;
    void mark(const char* name, int l, cudaStream_t st) {
      if (!active) return;
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      cudaEventRecord(e, st);
      ev.push_back(e);
      names.push_back(l >= 0 ? std::string(name) + "[" + std::to_string(l) + "]" : std::string(name));
    }
    void report(int rank) {
      if (ev.size() < 2) return;
      cudaEventSynchronize(ev.back());
      std::string out;
      float tot = 0.f;
      for (size_t i = 1; i < ev.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
        tot += ms;
        char buf[96];
        std::snprintf(buf, sizeof buf, " %s=%.1f", names[i].c_str(), 1e3 * ms);
        out += buf;
      }
      std::fprintf(stderr, "JSSO_MG_ITER rank %d (us, one iteration = %.1f):%s\n", rank, 1e3 * tot, out.c_str());
      for (cudaEvent_t e : ev) cudaEventDestroy(e);
      ev.clear(); names.clear();
    }
  } probe;
  int mg_poll = 8;                 // the PCG scalars stay on the device; the host polls the residual every mg_poll iterations
  // opt-in CUDA graph of the V-cycle's launch-bound part (JSSO_MG_GRAPH=1): the whole V-cycle on one GPU, the
  // replicated coarse levels of the distributed solve.  Captured once per numeric setup (the smoother
  // coefficients are kernel arguments) on a private stream, replayed into the caller's stream.
  bool mg_graph = true;
  cudaStream_t st_cap = nullptr;
  cudaGraphExec_t mg_graph_exec = nullptr;
  int mg_graph_level = -1, mg_graph_deg = 0;
  const double* mg_graph_b = nullptr;
  double* mg_graph_x = nullptr;
  double* mg_dense = nullptr;      // [A | A^-1] of the coarsest level
  double* mg_dense_ws = nullptr;   // blocked inverse: pivot-block inverse, row panel, column-panel copy
  double *mg_cb = nullptr, *mg_cx = nullptr;   // coarsest-level vectors
  double* mg_scal = nullptr;       // device scalars of the host-driven PCG
  double* mg_scal_host = nullptr;  // pinned
  bool mg_ready = false;           // numeric hierarchy matches the current matrix
  const double* last_crds = nullptr;   // = crds_keep after an assembly (handle-owned copy: the caller's array may be gone by the time of the solve)
  double* crds_keep = nullptr;
  // row-range distributed multigrid solve (jsso_mg_set_dist): the hierarchy above stays replicated on every
  // rank, the V-cycle / PCG products are computed by row ranges with halo exchanges between the ranks
  struct MgDistPeer { int rank, send_off, send_cnt, recv_off, recv_cnt; };   // node counts
  struct MgDistLevel {
    std::vector<MgDistPeer> peers;
    int32_t *send_idx = nullptr, *recv_idx = nullptr;   // device: node ids of this level
    int n_send = 0, n_recv = 0;
    int32_t* ghost_rows = nullptr;                      // device: rows of other ranks this rank's rows of A_l read
    int n_ghost = 0;
    std::vector<int32_t> remote_off;                    // where my block starts in each peer's receive list (nodes)
    MgdLevelDev* dev = nullptr;                         // device copy for the peer-memory kernels
    unsigned long long seq = 0;                         // exchange counter of this level
  };
  struct MgDist {
    int rank = 0, n_rank = 1, n_dist = 0;
    ncclComm_t comm = nullptr;
    std::vector<std::vector<int32_t>> bounds;   // [n_dist + 1][n_rank + 1] row ranges per level
    std::vector<MgDistLevel> lv;                // [n_dist]
    double *send_buf = nullptr, *recv_buf = nullptr;
    long long n_exchange = 0, n_allreduce = 0;
    // peer-memory path (jsso_mg_p2p_export / jsso_mg_p2p_connect): mailbox + receive arena mapped into every peer
    MgdMailbox* mbox = nullptr;
    double* arena = nullptr;
    size_t max_recv = 0;
    MgdCtx* ctx = nullptr;              // device copy; null = NCCL send/recv path
    unsigned* push_counter = nullptr;
    unsigned long long* red_seq_dev = nullptr;   // sequence number of the mailbox all-reduces (device resident)
    std::vector<void*> ipc_opened;
    // distributed numeric setup (jsso_mg_set_dist_setup)
    struct SetupLevel {
      int32_t *p_list = nullptr, *ap_list = nullptr;
      int n_p = 0, n_ap = 0;
      std::vector<int32_t> ac_bounds;   // [n_rank + 1] slot ranges of the coarse matrix
      int p_lo = 0, p_hi = 0, pt_lo = 0, pt_hi = 0;   // slot ranges of the own prolongation / restriction rows
    };
    bool setup_on = false;
    std::vector<SetupLevel> setup;
    int sc_lo = 0, sc_hi = 0, w_lo = 0, w_hi = 0;      // level-0 row hulls: scaled blocks, block-Jacobi factors
    int asm_t0 = 0, asm_t1 = 0, asm_q0 = 0, asm_q1 = 0;   // assembly: task range and quad range covering [w_lo, w_hi)
  } mgd;
  // host staging for the host-buffer entry point
  double *h_crds = nullptr, *h_pq = nullptr, *h_pb = nullptr, *h_f = nullptr, *h_u = nullptr;
  double *h_dc = nullptr, *h_dpq = nullptr, *h_dpb = nullptr;
  double* val_dev = nullptr;       // f.u of jsso_value_and_grad_host (device dot product instead of a host loop over 6 n_node entries)
  double* val_host = nullptr;      // pinned
  cudaEvent_t ev_out[4] = {nullptr, nullptr, nullptr, nullptr};   // D2H of u / d_crds / d_prop_q / d_prop_b done
  cudaStream_t st_a = nullptr, st_b = nullptr;   // non-blocking streams of the host-buffer entry point
  cudaEvent_t ev_b = nullptr;
  // chunked pipeline of jsso_assemble_adjoint_host (JSSO_E2E_CHUNKS = K, default 8 from 400 000 quads, else 4; 1 = one launch): u / lam arrive
  // in K node ranges, the quad adjoint runs in K quad ranges as soon as the rows a range reads have arrived, and
  // every range's d_prop_q goes back to the host while the next range is differentiated (measured at 1M quads on a
  // B200, PCIe-bound step: 5.04 ms unchunked, 4.17 / 4.22 ms with K = 4, 4.06 / 4.08 with K = 8, 4.07 / 4.13 with K = 16)
  int e2e_chunks = 8;
  cudaStream_t st_c = nullptr;
  std::vector<cudaEvent_t> ev_up, ev_adj;
  std::vector<int> e2e_qb, e2e_nb, e2e_wait;   // quad bounds, node bounds, upload range each quad range waits for
  // device scratch of the host-buffer entry point
  double *s_crds = nullptr, *s_pq = nullptr, *s_pb = nullptr, *s_f = nullptr, *s_u = nullptr;
  double *s_dc = nullptr, *s_dpq = nullptr, *s_dpb = nullptr;
};

static int fail(jsso_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(h, JSSO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)
#define CKL(name)                                                                             \
  do {                                                                                        \
    LAUNCHED();                                                                               \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess)                                                                    \
      return fail(h, JSSO_ERR_CUDA, std::string(name) + " launch: " + cudaGetErrorString(e_)); \
  } while (0)
#define CKN(call)                                                                             \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess)                                                                    \
      return fail(h, JSSO_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_));  \
  } while (0)

template <class T>
static cudaError_t upload(T** dst, const std::vector<T>& v) {
  const size_t bytes = (v.size() ? v.size() : 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void**)dst, bytes);
  if (e != cudaSuccess) return e;
  if (v.size()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}
template <class T>
static cudaError_t dalloc(T** dst, size_t n) {
  return cudaMalloc((void**)dst, (n ? n : 1) * sizeof(T));
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
#define NEED_GPU()                                                                            \
  do {                                                                                        \
    if (h->device == JSSO_DEVICE_NONE)                                                        \
      return fail(h, JSSO_ERR_STATE, "symbolic-only handle (device = JSSO_DEVICE_NONE): no compute"); \
  } while (0)

static int handle_upload(jsso_handle* h, const std::vector<int32_t>& cq, const std::vector<int32_t>& cb) {
  const Symbolic& S = h->sym;
  CK(cudaSetDevice(h->device));
  CK(upload(&h->cnct_q, cq)); CK(upload(&h->cnct_b, cb));
  CK(upload(&h->rowptr, S.rowptr)); CK(upload(&h->colidx, S.colidx));
  CK(upload(&h->blk_row, S.blk_row)); CK(upload(&h->diag_slot, S.diag_slot));
  CK(upload(&h->blk_item_ptr, S.blk_item_ptr)); CK(upload(&h->item_code, S.item_code));
  CK(upload(&h->item_lel, S.item_lel)); CK(upload(&h->node_mask, S.node_mask));
  CK(upload(&h->chunk_blk, S.chunk_blk)); CK(upload(&h->chunk_el_ptr, S.chunk_el_ptr));
  CK(upload(&h->chunk_els, S.chunk_els)); CK(upload(&h->blk_perm, S.blk_perm));
  {
    const char* e = std::getenv("JSSO_ASM_CHUNKED");   // A/B switch: force the chunked single-kernel assembly
    h->asm_tasks = S.tasks_ok && !(e && e[0] == '1');
  }
  if (h->asm_tasks) {
    CK(upload(&h->task_meta, S.task_meta)); CK(upload(&h->task_els, S.task_els));
    CK(upload(&h->item_desc, S.item_desc)); CK(upload(&h->blk_bc, S.blk_bc));
    CK(dalloc(&h->quad_rec, (size_t)S.n_quad * REC_GLD));
  }
  CK(upload(&h->node_inc_ptr, S.node_inc_ptr)); CK(upload(&h->node_inc, S.node_inc));
  const size_t nd = 6 * (size_t)S.n_node;
  CK(dalloc(&h->vals, (size_t)S.nnzb() * 36));
  CK(dalloc(&h->W, (size_t)S.n_node * 36));
  CK(dalloc(&h->vb, nd)); CK(dalloc(&h->vx, nd)); CK(dalloc(&h->vr, nd));
  CK(dalloc(&h->vp, nd)); CK(dalloc(&h->vq, nd));
  CK(dalloc(&h->tmp_lam, nd)); CK(dalloc(&h->tmp_g, nd));
  CK(cudaMemset(h->vp, 0, nd * sizeof(double)));
  CK(cudaMemset(h->vx, 0, nd * sizeof(double)));
  CK(dalloc(&h->corner_q, (size_t)S.n_quad * 12)); CK(dalloc(&h->corner_b, (size_t)S.n_beam * 6));
  CK(dalloc(&h->sc, 1)); CK(cudaMemset(h->sc, 0, sizeof(CgScalars)));
  CK(cudaMallocHost((void**)&h->sc_host, sizeof(CgScalars)));
  CK(dalloc(&h->partials, 2 * (size_t)RED_MAX_BLOCKS));
  CK(dalloc(&h->counters, 4)); CK(cudaMemset(h->counters, 0, 4 * sizeof(unsigned)));
  CK(dalloc(&h->flags, 1)); CK(cudaMemset(h->flags, 0, sizeof(int)));
  CK(cudaMallocHost((void**)&h->flags_host, sizeof(int)));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  h->red_blocks = std::min(RED_MAX_BLOCKS, prop.multiProcessorCount * 4);
  {
    // persistent SpMV grids: exactly the co-resident CTAs (one wave), the smallest over the variants
    int o1 = 0, o2 = 0, o3 = 0, o4 = 0, o5 = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, bsr_spmv_kernel<1>, RED_BLOCK, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, bsr_spmv_axpby_kernel<2, float>, RED_BLOCK, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, bsr_spmv_axpby_kernel<2, double>, RED_BLOCK, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o4, bsr_spmv_dot_kernel, RED_BLOCK, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o5, bsr_spmv_lin_kernel<float, 1>, RED_BLOCK, 0));
    const int occ = std::max(1, std::min(std::min(o1, o2), std::min(o3, std::min(o4, o5))));
    int o6 = 0, o7 = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o6, bsr_spmv_rp_kernel<float, 1>, RED_BLOCK, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o7, bsr_spmv_rp_kernel<__half, 1>, RED_BLOCK, 0));
    h->rp_blocks = std::min(RED_MAX_BLOCKS, prop.multiProcessorCount * std::max(1, std::min(o6, o7)));
    if (const char* e = std::getenv("JSSO_RP_BLOCKS")) h->rp_blocks = std::max(1, std::min(RED_MAX_BLOCKS, std::atoi(e)));
    h->spmv_blocks = std::min(RED_MAX_BLOCKS, prop.multiProcessorCount * occ);
    if (const char* e = std::getenv("JSSO_SPMV_BLOCKS")) h->spmv_blocks = std::max(1, std::min(RED_MAX_BLOCKS, std::atoi(e)));
  }
  {
    int occ = 0, coop = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cg_persistent_kernel, RED_BLOCK, 0));
    h->coop_blocks = coop ? std::min(RED_MAX_BLOCKS, occ * prop.multiProcessorCount) : 0;
  }
  CK(cudaFuncSetAttribute(assemble_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          FUSED_SMEM_DOUBLES * (int)sizeof(double)));
  CK(cudaFuncSetAttribute(assemble_tasks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          TASK_WARPS * TASK_SMEM_DOUBLES * (int)sizeof(double)));
  CK(cudaFuncSetAttribute(quad_geometry_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          G_QUADS * QS * (int)sizeof(double)));
  // 4 CTAs x 55 KB per SM need the largest shared-memory carve-out
  CK(cudaFuncSetAttribute(assemble_tasks_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                          (int)cudaSharedmemCarveoutMaxShared));
  {
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, assemble_tasks_kernel, 32 * TASK_WARPS,
                                                     TASK_WARPS * TASK_SMEM_DOUBLES * sizeof(double)));
    h->task_ctas = std::max(1, occ) * prop.multiProcessorCount;
    if (const char* e = std::getenv("JSSO_TASK_CTAS")) h->task_ctas = std::max(1, std::atoi(e));
    const size_t adj_smem = ADJ_SMEM_DOUBLES * sizeof(double);
    CK(cudaFuncSetAttribute(quad_adjoint_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adj_smem));
    CK(cudaFuncSetAttribute(quad_adjoint_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adj_smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quad_adjoint_kernel<false>, 4 * ADJ_QUADS, adj_smem));
    h->adj_ctas = std::max(1, occ) * prop.multiProcessorCount;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quad_adjoint_kernel<true>, 4 * ADJ_QUADS, adj_smem));
    h->adj_ctas_prop = std::max(1, occ) * prop.multiProcessorCount;
  }
  return JSSO_OK;
}

extern "C" {

const char* jsso_last_error(const jsso_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int jsso_create(const jsso_mesh_desc* d, jsso_handle** out) {
  jsso_handle* h = nullptr;
  if (!d || !out) return fail(h, JSSO_ERR_ARG, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (d->device != JSSO_DEVICE_NONE) {
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      return fail(h, JSSO_ERR_CUDA, "no CUDA device: the jaxsso_b200 hot path has no CPU fallback");
    if (d->device < 0 || d->device >= ndev) return fail(h, JSSO_ERR_ARG, "bad device ordinal");
  }
  jsso_handle* nh = new jsso_handle();
  nh->device = d->device;
  std::string msg = build_symbolic(d->n_node, d->n_row, d->n_quad, d->cnct_quads, d->n_beam, d->cnct_beams,
                                   d->n_known, d->known, nh->sym);
  if (!msg.empty()) { delete nh; return fail(h, JSSO_ERR_ARG, msg); }
  h = nh;
  if (d->device == JSSO_DEVICE_NONE) { *out = h; return JSSO_OK; }   // symbolic-only handle
  std::vector<int32_t> cq(d->cnct_quads, d->cnct_quads + 4 * (size_t)d->n_quad);
  std::vector<int32_t> cb(d->cnct_beams, d->cnct_beams + 2 * (size_t)d->n_beam);
  { int rc_ = handle_upload(h, cq, cb); if (rc_) return rc_; }
  *out = h;
  return JSSO_OK;
}

// Solver-plugin compatibility mode: a handle over a GIVEN 6x6 block-CSR pattern,
// values supplied by the caller (e.g. the reference's own K after sort_indices + sum_duplicates).
int jsso_create_from_bsr(int32_t n_node, const int32_t* rowptr, const int32_t* colidx, int32_t n_known,
                         const int32_t* known, int32_t device, jsso_handle** out) {
  jsso_handle* h = nullptr;
  if (!rowptr || !colidx || !out) return fail(h, JSSO_ERR_ARG, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, JSSO_ERR_CUDA, "no CUDA device: the jaxsso_b200 hot path has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(h, JSSO_ERR_ARG, "bad device ordinal");
  jsso_handle* nh = new jsso_handle();
  nh->device = device;
  std::string msg = build_symbolic_from_bsr(n_node, rowptr, colidx, n_known, known, nh->sym);
  if (!msg.empty()) { delete nh; return fail(h, JSSO_ERR_ARG, msg); }
  h = nh;
  h->pattern_only = true;
  { int rc_ = handle_upload(h, std::vector<int32_t>(), std::vector<int32_t>()); if (rc_) return rc_; }
  *out = h;
  return JSSO_OK;
}

// vals_h: nnzb*36 doubles, blocks column-major, WITHOUT boundary conditions; apply_bc imposes them.
int jsso_set_values_host(jsso_handle* h, const double* vals_h, int apply_bc) {
  if (!h || !vals_h) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  const long long n_out = (long long)h->sym.nnzb() * 36;
  CK(cudaMemcpy(h->vals, vals_h, n_out * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemset(h->flags, 0, sizeof(int)));
  if (apply_bc && n_out > 0) {
    apply_bc_kernel<<<cdiv(n_out, 256), 256>>>(n_out, h->blk_row, h->colidx, h->node_mask, h->vals);
    CKL("apply_bc_kernel");
  }
  h->assembled = true; h->assembled_bc = apply_bc != 0; h->scaled = false; h->mg_ready = false;
  return JSSO_OK;
}

void jsso_destroy(jsso_handle* h) {
  if (!h) return;
  if (h->device == JSSO_DEVICE_NONE) { delete h; return; }
  cudaSetDevice(h->device);
  void* dev[] = {h->cnct_q, h->cnct_b, h->rowptr, h->colidx, h->blk_row, h->diag_slot, h->blk_item_ptr,
                 h->item_code, h->item_lel, h->node_mask, h->chunk_blk, h->chunk_el_ptr, h->chunk_els, h->blk_perm,
                 h->task_meta, h->task_els, h->item_desc, h->blk_bc, h->quad_rec, h->node_inc_ptr, h->node_inc, h->vals, h->W, h->vb, h->vx, h->vr, h->vp, h->vq,
                 h->corner_q, h->corner_b, h->tmp_lam, h->tmp_g, h->sc, h->partials, h->counters, h->flags,
                 h->send_idx, h->send_buf, h->mbox, h->p2p, h->s_crds, h->s_pq, h->s_pb, h->s_f, h->s_u, h->s_dc, h->s_dpq,
                 h->s_dpb, h->val_dev};
  for (void* p : dev) if (p) cudaFree(p);
  void* hst[] = {h->sc_host, h->flags_host, h->h_crds, h->h_pq, h->h_pb, h->h_f, h->h_u, h->h_dc, h->h_dpq,
                 h->h_dpb, h->val_host};
  for (void* p : hst) if (p) cudaFreeHost(p);
  for (cudaEvent_t e : h->ev_out) if (e) cudaEventDestroy(e);
  if (h->st_a) cudaStreamDestroy(h->st_a);
  if (h->st_b) cudaStreamDestroy(h->st_b);
  if (h->ev_b) cudaEventDestroy(h->ev_b);
  if (h->st_c) cudaStreamDestroy(h->st_c);
  if (h->mg_graph_exec) cudaGraphExecDestroy(h->mg_graph_exec);
  if (h->st_cap) cudaStreamDestroy(h->st_cap);
  for (cudaEvent_t e : h->ev_up) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_adj) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_prof) if (e) cudaEventDestroy(e);
  for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
  for (auto& m : h->mg) {
    void* lv[] = {m.agg, m.p_row, m.p_rowptr, m.p_col, m.p_own, m.ps_ptr, m.ps_a, m.ps_j, m.apl_ptr, m.apl_a, m.apl_p,
                  m.c_rowptr, m.c_col, m.c_diag, m.cl_ptr, m.cl_p, m.cl_ap, m.pt_rowptr, m.pt_col, m.pt_src, m.mem_ptr,
                  m.mem, m.P, m.Pt, m.AP, m.Ac, m.Xc, m.Dinv, m.P32, m.Pt32, m.Ac32, m.b, m.x, m.r, m.d, m.Wc, m.Lc, m.c_row};
    for (void* p : lv) if (p) cudaFree(p);
  }
  {
    void* mgp[] = {h->Lfac, h->vals32, h->vals16, h->mg_dense, h->mg_dense_ws, h->mg_cb, h->mg_cx, h->mg_scal, h->crds_keep};
    for (void* p : mgp) if (p) cudaFree(p);
    if (h->mg_scal_host) cudaFreeHost(h->mg_scal_host);
  }
  for (auto& dl : h->mgd.lv) { if (dl.send_idx) cudaFree(dl.send_idx); if (dl.recv_idx) cudaFree(dl.recv_idx); if (dl.ghost_rows) cudaFree(dl.ghost_rows); }
  for (auto& dl : h->mgd.lv) if (dl.dev) cudaFree(dl.dev);
  for (void* p : h->mgd.ipc_opened) cudaIpcCloseMemHandle(p);
  { void* pp[] = {h->mgd.mbox, h->mgd.arena, h->mgd.ctx, h->mgd.push_counter, h->mgd.red_seq_dev}; for (void* p : pp) if (p) cudaFree(p); }
  for (auto& sl : h->mgd.setup) { if (sl.p_list) cudaFree(sl.p_list); if (sl.ap_list) cudaFree(sl.ap_list); }
  if (h->mgd.send_buf) cudaFree(h->mgd.send_buf);
  if (h->mgd.recv_buf) cudaFree(h->mgd.recv_buf);
  if (h->mgd.comm && g_nccl.ok) g_nccl.CommDestroy(h->mgd.comm);
  delete h;
}

int jsso_get_sizes(const jsso_handle* h, jsso_sizes* o) {
  if (!h || !o) return JSSO_ERR_ARG;
  o->n_node = h->sym.n_node; o->n_row = h->sym.n_row; o->n_quad = h->sym.n_quad; o->n_beam = h->sym.n_beam;
  o->nnzb = h->sym.nnzb(); o->n_items = h->sym.n_items(); o->n_chunk = h->sym.n_chunk();
  return JSSO_OK;
}

int jsso_pattern(const jsso_handle* h, int32_t* rowptr, int32_t* colidx) {
  if (!h || !rowptr || !colidx) return JSSO_ERR_ARG;
  std::memcpy(rowptr, h->sym.rowptr.data(), h->sym.rowptr.size() * sizeof(int32_t));
  std::memcpy(colidx, h->sym.colidx.data(), h->sym.colidx.size() * sizeof(int32_t));
  return JSSO_OK;
}

// host-only helper of the multigrid symbolic setup (no handle, no device)
int jsso_mg_aggregate(int32_t n, const int32_t* rowptr, const int32_t* colidx, int32_t* agg, int32_t* n_agg) {
  if (n < 0 || !rowptr || !agg || !n_agg || (rowptr[n] > 0 && !colidx)) return JSSO_ERR_ARG;
  *n_agg = mg_aggregate(n, rowptr, colidx, agg);
  return JSSO_OK;
}

int jsso_mg_pattern_lists(int64_t m, const int32_t* row, const int32_t* col, const int32_t* left,
                          const int32_t* right, int32_t n_row, int32_t* rowptr, int32_t* ocol, int32_t* ptr,
                          int32_t* left_o, int32_t* right_o, int64_t* n_blk) {
  if (m < 0 || m >= INT32_MAX || n_row < 0 || !rowptr || !ptr || !n_blk ||
      (m > 0 && (!row || !col || !left || !right || !ocol || !left_o || !right_o)))
    return JSSO_ERR_ARG;
  for (int64_t t = 0; t < m; ++t)
    if (row[t] < 0 || row[t] >= n_row) return JSSO_ERR_ARG;
  *n_blk = mg_pattern_lists(m, row, col, left, right, n_row, rowptr, ocol, ptr, left_o, right_o);
  return JSSO_OK;
}

int jsso_assembly_tasks(const jsso_handle* h, int32_t* counts, int32_t* task_meta, int32_t* task_els,
                        uint16_t* item_desc, uint16_t* blk_bc, int32_t* blk_item_ptr, int32_t* item_code) {
  if (!h || !counts) return JSSO_ERR_ARG;
  const Symbolic& S = h->sym;
  counts[0] = S.n_task(); counts[1] = (int32_t)S.task_els.size(); counts[2] = S.tasks_ok ? 1 : 0;
  if (task_meta) std::memcpy(task_meta, S.task_meta.data(), S.task_meta.size() * sizeof(int32_t));
  if (task_els) std::memcpy(task_els, S.task_els.data(), S.task_els.size() * sizeof(int32_t));
  if (item_desc) std::memcpy(item_desc, S.item_desc.data(), S.item_desc.size() * sizeof(uint16_t));
  if (blk_bc) std::memcpy(blk_bc, S.blk_bc.data(), S.blk_bc.size() * sizeof(uint16_t));
  if (blk_item_ptr) std::memcpy(blk_item_ptr, S.blk_item_ptr.data(), S.blk_item_ptr.size() * sizeof(int32_t));
  if (item_code) std::memcpy(item_code, S.item_code.data(), S.item_code.size() * sizeof(int32_t));
  return JSSO_OK;
}

int jsso_quad_ke(jsso_handle* h, const double* crds, const double* prop_q, double* ke, void* stream) {
  if (!h) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  const int n = h->sym.n_quad;
  if (n == 0) return JSSO_OK;
  quad_ke_kernel<<<cdiv(n, 16), 256, 0, (cudaStream_t)stream>>>(n, crds, h->cnct_q, prop_q, ke, h->flags);
  CKL("quad_ke_kernel");
  return JSSO_OK;
}

int jsso_quad_area(jsso_handle* h, const double* crds, double* area, void* stream) {
  if (!h || !crds || !area) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  const int n = h->sym.n_quad;
  if (n == 0) return JSSO_OK;
  quad_area_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, crds, h->cnct_q, area);
  CKL("quad_area_kernel");
  return JSSO_OK;
}

int jsso_csr_spmv(int32_t n_row, const int32_t* rowptr_d, const int32_t* colidx_d, const double* vals_d,
                  const double* x_d, double* y_d, void* stream) {
  jsso_handle* h = nullptr;
  if (n_row < 0 || !rowptr_d || !y_d) return fail(h, JSSO_ERR_ARG, "bad csr_spmv argument");
  if (n_row == 0) return JSSO_OK;
  csr_spmv_kernel<<<cdiv(8LL * n_row, 256), 256, 0, (cudaStream_t)stream>>>(n_row, rowptr_d, colidx_d, vals_d, x_d, y_d);
  CKL("csr_spmv_kernel");
  return JSSO_OK;
}

int jsso_beam_ke(jsso_handle* h, const double* crds, const double* prop_b, double* ke, void* stream) {
  if (!h) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  const int n = h->sym.n_beam;
  if (n == 0) return JSSO_OK;
  beam_ke_kernel<<<cdiv(4LL * n, 256), 256, 0, (cudaStream_t)stream>>>(n, crds, h->cnct_b, prop_b, ke, h->flags);
  CKL("beam_ke_kernel");
  return JSSO_OK;
}

// The two kernels of the warp-task assembly by ranges: geometry records of quads [q0, q1), tasks [t0, t1).  A task
// writes its own run of block slots and reads only the records of the quads it stages, so any split into ranges gives
// bitwise the same matrix (the distributed numeric setup assembles only the ranges a rank reads).
static int assemble_geometry_range(jsso_handle* h, const double* crds, const double* prop_q, int q0, int q1, cudaStream_t st) {
  if (q1 <= q0) return JSSO_OK;
  quad_geometry_kernel<<<cdiv(q1 - q0, G_QUADS), G_THREADS, G_QUADS * QS * sizeof(double), st>>>(
      q1 - q0, crds, h->cnct_q + 4 * (size_t)q0, prop_q + 5 * (size_t)q0, h->quad_rec + (size_t)q0 * REC_GLD, h->flags);
  CKL("quad_geometry_kernel");
  return JSSO_OK;
}
static int assemble_task_range(jsso_handle* h, const double* crds, const double* prop_b, int apply_bc, int t0, int t1,
                               cudaStream_t st) {
  if (t1 <= t0) return JSSO_OK;
  TaskArgs T;
  T.rec = h->quad_rec; T.task_meta = (const int4*)h->task_meta + t0; T.task_els = h->task_els;
  T.item_desc = h->item_desc; T.blk_bc = h->blk_bc; T.item_code = h->item_code;
  T.blk_item_ptr = h->blk_item_ptr;
  T.crds = crds; T.cnct_b = h->cnct_b; T.prop_b = prop_b;
  T.vals = h->vals; T.flags = h->flags; T.n_quad = h->sym.n_quad; T.n_task = t1 - t0; T.apply_bc = apply_bc;
  assemble_tasks_kernel<<<std::min(cdiv(T.n_task, TASK_WARPS), h->task_ctas), 32 * TASK_WARPS,
                          TASK_WARPS * TASK_SMEM_DOUBLES * sizeof(double), st>>>(T);
  CKL("assemble_tasks_kernel");
  return JSSO_OK;
}
// bookkeeping after the last kernel of an assembly
static int assemble_done(jsso_handle* h, const double* crds, int apply_bc, cudaStream_t st) {
  h->assembled = true; h->assembled_bc = apply_bc != 0; h->scaled = false; h->mg_ready = false;
  h->last_crds = nullptr;
  if (!h->mg.empty() && crds) {   // the numeric multigrid setup reads the coordinates at the time of the solve
    if (!h->crds_keep) CK(dalloc(&h->crds_keep, 3 * (size_t)h->sym.n_node));
    CK(cudaMemcpyAsync(h->crds_keep, crds, 3 * (size_t)h->sym.n_node * sizeof(double), cudaMemcpyDeviceToDevice, st));
    h->last_crds = h->crds_keep;
  }
  return JSSO_OK;
}

int jsso_assemble(jsso_handle* h, const double* crds, const double* prop_q, const double* prop_b, int apply_bc,
                  void* stream) {
  if (!h) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(h->flags, 0, sizeof(int), st));
  int rc;
  if (h->sym.nnzb() > 0 && h->asm_tasks) {
    const int nq = h->sym.n_quad;
    // distributed numeric setup (jsso_mg_set_dist_setup): only the tasks that touch the rows this rank reads, and
    // the geometry records of the quads they stage
    const bool part = h->mgd.setup_on;
    const int q0 = part ? h->mgd.asm_q0 : 0, q1 = part ? h->mgd.asm_q1 : nq;
    const int t0 = part ? h->mgd.asm_t0 : 0, t1 = part ? h->mgd.asm_t1 : h->sym.n_task();
    if (h->prof) CK(cudaEventRecord(h->ev_prof[0], st));
    if ((rc = assemble_geometry_range(h, crds, prop_q, q0, q1, st))) return rc;
    if (h->prof) CK(cudaEventRecord(h->ev_prof[1], st));
    if ((rc = assemble_task_range(h, crds, prop_b, apply_bc, t0, t1, st))) return rc;
    if (h->prof) CK(cudaEventRecord(h->ev_prof[2], st));
  } else if (h->sym.nnzb() > 0) {
    AsmArgs A;
    A.crds = crds; A.cnct_q = h->cnct_q; A.prop_q = prop_q; A.cnct_b = h->cnct_b; A.prop_b = prop_b;
    A.chunk_blk = h->chunk_blk; A.chunk_el_ptr = h->chunk_el_ptr; A.chunk_els = h->chunk_els;
    A.blk_perm = h->blk_perm; A.blk_item_ptr = h->blk_item_ptr; A.item_code = h->item_code; A.item_lel = h->item_lel;
    A.blk_row = h->blk_row; A.colidx = h->colidx; A.node_mask = h->node_mask;
    A.vals = h->vals; A.flags = h->flags; A.n_quad = h->sym.n_quad; A.apply_bc = apply_bc;
    assemble_fused_kernel<<<h->sym.n_chunk(), kChunkBlocks, FUSED_SMEM_DOUBLES * sizeof(double), st>>>(A);
    CKL("assemble_fused_kernel");
  }
  return assemble_done(h, crds, apply_bc, st);
}

// dst[i*width + k] = src[idx[i]*width + k]: node-row gather between numberings (e.g. the local part
// u_global[l2g] of a replicated solve for the partitioned adjoint)
int jsso_gather_rows(const double* src_d, const int32_t* idx_d, int32_t n, int32_t width, double* dst_d, void* stream) {
  if (n < 0 || width <= 0 || (n > 0 && (!src_d || !idx_d || !dst_d))) return JSSO_ERR_ARG;
  if (n == 0) return JSSO_OK;
  const long long total = (long long)n * width;
  gather_rows_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(total, width, src_d, idx_d, dst_d);
  LAUNCHED();
  return cudaGetLastError() == cudaSuccess ? JSSO_OK : JSSO_ERR_CUDA;
}

// Per-kernel timing of the two-kernel assembly (bench.py roofline): enable, call jsso_assemble, read.
int jsso_profile(jsso_handle* h, int enable) {
  if (!h) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  if (enable && !h->ev_prof[0])
    for (int i = 0; i < 3; ++i) CK(cudaEventCreate(&h->ev_prof[i]));
  h->prof = enable != 0;
  return JSSO_OK;
}

// ms[0] = quad_geometry_kernel, ms[1] = assemble_tasks_kernel of the LAST jsso_assemble (both 0 on the chunked path)
int jsso_profile_read(jsso_handle* h, float* ms) {
  if (!h || !ms) return JSSO_ERR_ARG;
  NEED_GPU();
  ms[0] = ms[1] = 0.f;
  if (!h->prof || !h->asm_tasks || !h->assembled) return JSSO_OK;
  CK(cudaSetDevice(h->device));
  CK(cudaEventSynchronize(h->ev_prof[2]));
  CK(cudaEventElapsedTime(&ms[0], h->ev_prof[0], h->ev_prof[1]));
  CK(cudaEventElapsedTime(&ms[1], h->ev_prof[1], h->ev_prof[2]));
  return JSSO_OK;
}

int jsso_assemble_from_ke(jsso_handle* h, const double* ke_q, const double* ke_b, int apply_bc, void* stream) {
  if (!h) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  const long long n_out = (long long)h->sym.nnzb() * 36;
  if (n_out > 0) {
    assemble_from_ke_kernel<<<cdiv(n_out, 256), 256, 0, (cudaStream_t)stream>>>(
        n_out, h->sym.n_quad, ke_q, ke_b, h->blk_item_ptr, h->item_code, h->blk_row, h->colidx, h->node_mask,
        h->vals, apply_bc);
    CKL("assemble_from_ke_kernel");
  }
  h->assembled = true; h->assembled_bc = apply_bc != 0; h->scaled = false; h->mg_ready = false;
  return JSSO_OK;
}

int jsso_get_values(jsso_handle* h, double* vals_d, void* stream) {
  if (!h || !vals_d) return JSSO_ERR_ARG;
  NEED_GPU();
  if (!h->assembled) return fail(h, JSSO_ERR_STATE, "no assembled matrix");
  CK(cudaSetDevice(h->device));
  const long long nnzb = h->sym.nnzb();
  if (h->scaled && nnzb > 0) {   // after a solve h->vals holds W K W^T: hand back K
    unscale_blocks_kernel<<<cdiv(nnzb, 128), 128, 0, (cudaStream_t)stream>>>(nnzb, h->blk_row, h->colidx, h->W, h->vals, vals_d);
    CKL("unscale_blocks_kernel");
    return JSSO_OK;
  }
  CK(cudaMemcpyAsync(vals_d, h->vals, (size_t)nnzb * 36 * sizeof(double), cudaMemcpyDeviceToDevice,
                     (cudaStream_t)stream));
  return JSSO_OK;
}

int jsso_get_values_host(jsso_handle* h, double* vals_h) {
  if (!h || !vals_h) return JSSO_ERR_ARG;
  NEED_GPU();
  if (!h->assembled) return fail(h, JSSO_ERR_STATE, "no assembled matrix");
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  const size_t bytes = (size_t)h->sym.nnzb() * 36 * sizeof(double);
  if (h->scaled && bytes > 0) {   // after a solve h->vals holds W K W^T: hand back K (through a scratch copy)
    double* tmp = nullptr;
    CK(cudaMalloc((void**)&tmp, bytes));
    int rc = jsso_get_values(h, tmp, nullptr);
    if (!rc && cudaMemcpy(vals_h, tmp, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(h, JSSO_ERR_CUDA, "copy of the unscaled values failed");
    cudaFree(tmp);
    return rc;
  }
  CK(cudaMemcpy(vals_h, h->vals, bytes, cudaMemcpyDeviceToHost));
  return JSSO_OK;
}

int jsso_get_flags(jsso_handle* h, int32_t* out) {
  if (!h || !out) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, h->flags, sizeof(int), cudaMemcpyDeviceToHost));
  return JSSO_OK;
}

// ---------------------------------------------------------------- halo exchange
int jsso_nccl_unique_id(uint8_t id_out[128]) {
  jsso_handle* h = nullptr;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (!nccl_load()) return fail(h, JSSO_ERR_NCCL, "libnccl.so.2 not found");
  CKN(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, 128);
  return JSSO_OK;
}

int jsso_set_halo(jsso_handle* h, const uint8_t nccl_id[128], int32_t rank, int32_t n_rank, int32_t n_peer,
                  const int32_t* peer_rank, const int32_t* send_ptr, const int32_t* send_idx,
                  const int32_t* recv_start, const int32_t* recv_count) {
  if (!h) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  h->rank = rank; h->n_rank = n_rank;
  h->peers.clear();
  for (int p = 0; p < n_peer; ++p)
    h->peers.push_back(HaloPeer{peer_rank[p], send_ptr[p], send_ptr[p + 1] - send_ptr[p], recv_start[p],
                                recv_count[p]});
  h->n_send_nodes = n_peer ? send_ptr[n_peer] : 0;
  std::vector<int32_t> si(send_idx, send_idx + h->n_send_nodes);
  for (int v : si) if (v < 0 || v >= h->sym.n_row) return fail(h, JSSO_ERR_ARG, "halo send index not owned");
  CK(upload(&h->send_idx, si));
  CK(dalloc(&h->send_buf, (size_t)h->n_send_nodes * 36));
  if (n_rank > 1) {
    if (!nccl_load()) return fail(h, JSSO_ERR_NCCL, "libnccl.so.2 not found");
    ncclUniqueId id;
    std::memcpy(&id, nccl_id, 128);
    CKN(g_nccl.CommInitRank(&h->comm, n_rank, id, rank));
  }
  return JSSO_OK;
}

// Fill the ghost part of a 6-doubles-per-node vector from the owning ranks: pack the
// owned interface nodes, then one grouped ncclSend/ncclRecv per neighbour.
static int halo_exchange_w(jsso_handle* h, double* vec, int width, cudaStream_t st) {
  if (h->n_rank <= 1 || h->peers.empty()) return JSSO_OK;
  if (width != 6) return fail(h, JSSO_ERR_ARG, "unsupported halo width");
  if (h->n_send_nodes > 0) {
    halo_pack_kernel<<<cdiv(6LL * h->n_send_nodes, 256), 256, 0, st>>>(h->n_send_nodes, h->send_idx, vec,
                                                                      h->send_buf);
    CKL("halo_pack_kernel");
  }
  CKN(g_nccl.GroupStart());
  for (const HaloPeer& p : h->peers) {
    if (p.send_cnt)
      CKN(g_nccl.Send(h->send_buf + 6 * (size_t)p.send_off, 6 * (size_t)p.send_cnt, ncclDouble, p.rank, h->comm, st));
    if (p.recv_cnt)
      CKN(g_nccl.Recv(vec + 6 * (size_t)p.recv_start, 6 * (size_t)p.recv_cnt, ncclDouble, p.rank, h->comm, st));
  }
  CKN(g_nccl.GroupEnd());
  return JSSO_OK;
}

int jsso_halo_exchange(jsso_handle* h, double* vec_d, void* stream) {
  if (!h) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  return halo_exchange_w(h, vec_d, 6, (cudaStream_t)stream);
}

// global = sum over ranks of this rank's partial (out of place, so repeating it is harmless)
// ---- peer-memory (CUDA IPC) setup: export my handles, import everybody's
int jsso_p2p_export(jsso_handle* h, uint8_t out[128]) {
  if (!h || !out) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  if (!h->mbox) {
    CK(dalloc(&h->mbox, 1));
    CK(cudaMemset(h->mbox, 0, sizeof(Mailbox)));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  cudaIpcMemHandle_t a, b;
  CK(cudaIpcGetMemHandle(&a, h->vp));
  CK(cudaIpcGetMemHandle(&b, h->mbox));
  std::memcpy(out, &a, 64);
  std::memcpy(out + 64, &b, 64);
  return JSSO_OK;
}

int jsso_p2p_connect(jsso_handle* h, const uint8_t* all_handles, const int32_t* remote_start) {
  if (!h || !all_handles || !remote_start) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  if (h->n_rank <= 1 || h->n_rank > P2P_MAX_RANKS || (int)h->peers.size() > P2P_MAX_RANKS || !h->mbox)
    return fail(h, JSSO_ERR_STATE, "p2p_connect needs set_halo + p2p_export first and <= 16 ranks");
  P2PCtx c;
  std::memset(&c, 0, sizeof c);
  c.rank = h->rank; c.n_rank = h->n_rank; c.n_peer = (int)h->peers.size();
  for (int r = 0; r < h->n_rank; ++r) {
    if (r == h->rank) { c.mbox[r] = h->mbox; continue; }
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, all_handles + 128 * (size_t)r + 64, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(p);
    c.mbox[r] = (Mailbox*)p;
  }
  for (int i = 0; i < c.n_peer; ++i) {
    const HaloPeer& hp = h->peers[i];
    cudaIpcMemHandle_t vh;
    std::memcpy(&vh, all_handles + 128 * (size_t)hp.rank, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, vh, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(p);
    c.peer_rank[i] = hp.rank; c.peer_vec[i] = (double*)p;
    c.send_off[i] = hp.send_off; c.send_cnt[i] = hp.send_cnt; c.remote_start[i] = remote_start[i];
  }
  CK(dalloc(&h->p2p, 1));
  CK(cudaMemcpy(h->p2p, &c, sizeof c, cudaMemcpyHostToDevice));
  return JSSO_OK;
}

static int allreduce_scalar(jsso_handle* h, const double* local, double* global, cudaStream_t st) {
  if (h->n_rank <= 1) return JSSO_OK;
  CKN(g_nccl.AllReduce(local, global, 1, ncclDouble, ncclSum, h->comm, st));
  return JSSO_OK;
}

// ---------------------------------------------------------------- SpMV / PCG
static int spmv_plain(jsso_handle* h, const double* x, double* y, cudaStream_t st) {
  const int n_row = h->sym.n_row;
  if (n_row == 0) return JSSO_OK;
  const int blocks = std::min(h->spmv_blocks, cdiv(n_row, RED_BLOCK / 32));
  bsr_spmv_kernel<0><<<blocks, RED_BLOCK, 0, st>>>(n_row, h->rowptr, h->colidx, h->vals, x, y, h->sc, 0,
                                                   h->partials, h->counters, 1, nullptr, 0, 0);
  CKL("bsr_spmv_kernel<0>");
  return JSSO_OK;
}

int jsso_spmv(jsso_handle* h, const double* x, double* y, void* stream) {
  if (!h) return JSSO_ERR_ARG;
  NEED_GPU();
  if (!h->assembled) return fail(h, JSSO_ERR_STATE, "spmv before assemble");
  CK(cudaSetDevice(h->device));
  if (!h->scaled) return spmv_plain(h, x, y, (cudaStream_t)stream);
  // a solve has replaced the values by the block-Jacobi-scaled A^ = W K W^T: K x = W^-1 (A^ (W^-T x))
  if (h->n_rank > 1) return fail(h, JSSO_ERR_STATE, "jsso_spmv after a solve is single-GPU only (the values are scaled)");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_row = h->sym.n_row;
  if (n_row == 0) return JSSO_OK;
  block_solve_wt_kernel<<<cdiv(n_row, 128), 128, 0, st>>>(n_row, h->W, x, nullptr, h->tmp_lam);
  CKL("block_solve_wt_kernel");
  int rc = spmv_plain(h, h->tmp_lam, h->tmp_g, st);
  if (rc) return rc;
  block_solve_w_kernel<<<cdiv(n_row, 128), 128, 0, st>>>(n_row, h->W, h->tmp_g, y);
  CKL("block_solve_w_kernel");
  return JSSO_OK;
}

// block-Jacobi scaling of the assembled matrix (once per assembly)
static int ensure_scaled(jsso_handle* h, cudaStream_t st) {
  if (h->scaled) return JSSO_OK;
  const int n_row = h->sym.n_row;
  if (h->mgd.setup_on) {
    // distributed numeric setup: factors of the rows [w_lo, w_hi), scaled blocks of the rows [sc_lo, sc_hi)
    const jsso_handle::MgDist& D = h->mgd;
    if (D.w_hi > D.w_lo) {
      diag_factor_kernel<<<cdiv(D.w_hi - D.w_lo, 128), 128, 0, st>>>(D.w_hi - D.w_lo, h->diag_slot + D.w_lo, h->vals,
                                                                    h->W + 36 * (size_t)D.w_lo,
                                                                    h->Lfac ? h->Lfac + 36 * (size_t)D.w_lo : nullptr, h->flags);
      CKL("diag_factor_kernel");
    }
    const long long b0 = h->sym.rowptr[D.sc_lo], b1 = h->sym.rowptr[D.sc_hi];
    if (b1 > b0) {
      scale_blocks_kernel<<<cdiv(b1 - b0, 128), 128, 0, st>>>(b1 - b0, h->blk_row + b0, h->colidx + b0, h->W, h->vals + 36 * b0);
      CKL("scale_blocks_kernel");
    }
    h->scaled = true;
    return JSSO_OK;
  }
  if (n_row > 0) {
    diag_factor_kernel<<<cdiv(n_row, 128), 128, 0, st>>>(n_row, h->diag_slot, h->vals, h->W, h->Lfac, h->flags);
    CKL("diag_factor_kernel");
  }
  if (h->n_rank > 1) {
    // ghost columns need the factor of their owner: W is (n_node, 6, 6) row-major, so row s of
    // every factor is a strided 6-vector; exchange the six rows through a dof-vector scratch
    const int n_node = h->sym.n_node;
    for (int s = 0; s < 6; ++s) {
      CK(cudaMemcpy2DAsync(h->vq, 6 * sizeof(double), h->W + 6 * s, 36 * sizeof(double), 6 * sizeof(double),
                           n_node, cudaMemcpyDeviceToDevice, st));
      int rc = halo_exchange_w(h, h->vq, 6, st);
      if (rc) return rc;
      CK(cudaMemcpy2DAsync(h->W + 6 * s, 36 * sizeof(double), h->vq, 6 * sizeof(double), 6 * sizeof(double),
                           n_node, cudaMemcpyDeviceToDevice, st));
    }
  }
  const long long nnzb = h->sym.nnzb();
  if (nnzb > 0) {
    scale_blocks_kernel<<<cdiv(nnzb, 128), 128, 0, st>>>(nnzb, h->blk_row, h->colidx, h->W, h->vals);
    CKL("scale_blocks_kernel");
  }
  h->scaled = true;
  return JSSO_OK;
}

static void default_opts(const jsso_solve_opts* in, jsso_solve_opts& o) {
  o.rtol = 1e-10; o.maxiter = 200000; o.check_every = 50; o.use_x0 = 0; o.compliance = 0;
  o.precond = 0; o.cheb_degree = 2;
  if (in) {
    if (in->rtol > 0) o.rtol = in->rtol;
    if (in->maxiter > 0) o.maxiter = in->maxiter;
    if (in->check_every > 0) o.check_every = in->check_every;
    o.use_x0 = in->use_x0; o.compliance = in->compliance;
    o.precond = in->precond;
    if (in->cheb_degree > 0) o.cheb_degree = in->cheb_degree;
  }
}

// One CG iteration on the scaled system (3 kernels; + halo exchange and 2 scalar
// all-reduces on several GPUs).
static int cg_iteration(jsso_handle* h, int cur, cudaStream_t st) {
  const int n_row = h->sym.n_row;
  const long long n = 6LL * n_row;
  const int single = h->n_rank <= 1;
  const int sblocks = std::max(1, std::min(h->spmv_blocks, cdiv(n_row, RED_BLOCK / 32)));
  const int vblocks = std::max(1, std::min(h->red_blocks, cdiv(n, RED_BLOCK)));
  int rc;
  if (h->p2p) {
    // peer-memory path: 4 launches, no NCCL call; the collectives live inside the kernels
    const unsigned long long hs = ++h->halo_seq, sa = ++h->red_seq_a, sb = ++h->red_seq_b;
    const int pblocks = std::max(1, std::min(64, cdiv(6LL * h->n_send_nodes, RED_BLOCK)));
    p2p_halo_push_kernel<<<pblocks, RED_BLOCK, 0, st>>>(h->p2p, h->send_idx, h->vp, h->counters + 3, hs);
    CKL("p2p_halo_push_kernel");
    bsr_spmv_kernel<1><<<sblocks, RED_BLOCK, 0, st>>>(n_row, h->rowptr, h->colidx, h->vals, h->vp, h->vq, h->sc,
                                                     cur, h->partials, h->counters, 0, h->p2p, hs, sa);
    CKL("bsr_spmv_kernel<1>");
    cg_update_kernel<<<vblocks, RED_BLOCK, 0, st>>>(n, cur, h->vp, h->vq, h->vx, h->vr, h->sc, h->partials,
                                                    h->counters + 1, 0, h->p2p, sa, sb);
    CKL("cg_update_kernel");
    cg_direction_kernel<<<vblocks, RED_BLOCK, 0, st>>>(n, cur, h->vr, h->vp, h->sc, h->p2p, sb);
    CKL("cg_direction_kernel");
    return JSSO_OK;
  }
  if ((rc = halo_exchange_w(h, h->vp, 6, st))) return rc;
  bsr_spmv_kernel<1><<<sblocks, RED_BLOCK, 0, st>>>(n_row, h->rowptr, h->colidx, h->vals, h->vp, h->vq, h->sc,
                                                   cur, h->partials, h->counters, single, nullptr, 0, 0);
  CKL("bsr_spmv_kernel<1>");
  if ((rc = allreduce_scalar(h, &h->sc->loc[0], &h->sc->pq, st))) return rc;
  cg_update_kernel<<<vblocks, RED_BLOCK, 0, st>>>(n, cur, h->vp, h->vq, h->vx, h->vr, h->sc, h->partials,
                                                  h->counters + 1, single, nullptr, 0, 0);
  CKL("cg_update_kernel");
  if (!single) {
    if ((rc = allreduce_scalar(h, &h->sc->loc[1], &h->sc->rr[cur ^ 1], st))) return rc;
    cg_latch_kernel<<<1, 1, 0, st>>>(cur, h->sc);
    CKL("cg_latch_kernel");
  }
  cg_direction_kernel<<<vblocks, RED_BLOCK, 0, st>>>(n, cur, h->vr, h->vp, h->sc, nullptr, 0);
  CKL("cg_direction_kernel");
  return JSSO_OK;
}

// Solve A^ y = b^ (both already scaled) into h->vx; b^ in h->vb.
static int cg_solve_scaled(jsso_handle* h, const jsso_solve_opts& o, bool use_x0, jsso_stats* stats,
                           cudaStream_t st) {
  const int n_row = h->sym.n_row;
  const long long n = 6LL * n_row;
  const int vblocks = std::max(1, std::min(h->red_blocks, cdiv(n, RED_BLOCK)));
  int restarts = 0, total_iter = 0;
  double relres_true = 0.0, relres_rec = 0.0, prev_true = 1e300;
  bool first = true, converged = false;
  int rc;
  for (;;) {
    // (re)start: r = b - A x
    const double* q = nullptr;
    if (!first || use_x0) {
      if ((rc = halo_exchange_w(h, h->vx, 6, st))) return rc;
      if ((rc = spmv_plain(h, h->vx, h->vq, st))) return rc;
      q = h->vq;
    } else {
      CK(cudaMemsetAsync(h->vx, 0, 6 * (size_t)h->sym.n_node * sizeof(double), st));
    }
    const int single = h->n_rank <= 1;
    if (first) {
      cg_init_kernel<1><<<vblocks, RED_BLOCK, 0, st>>>(n, h->vb, q, h->vr, h->vp, h->sc, h->partials,
                                                       h->counters + 2, o.rtol, single);
      CKL("cg_init_kernel<1>");
      if ((rc = allreduce_scalar(h, &h->sc->loc[2], &h->sc->bb, st))) return rc;
    } else {
      cg_init_kernel<0><<<vblocks, RED_BLOCK, 0, st>>>(n, h->vb, q, h->vr, h->vp, h->sc, h->partials,
                                                       h->counters + 2, o.rtol, single);
      CKL("cg_init_kernel<0>");
    }
    if (!single) {
      if ((rc = allreduce_scalar(h, &h->sc->loc[1], &h->sc->rr[0], st))) return rc;
      cg_copy_rr_kernel<<<1, 1, 0, st>>>(h->sc);
      CKL("cg_copy_rr_kernel");
    }
    first = false;
    int cur = 0, it_local = 0;
    bool stop = false;
    while (!stop) {
      const int batch = std::min(o.check_every, o.maxiter - total_iter - it_local);
      if (batch <= 0) break;
      // single GPU and a matrix that stays L2-resident (latency / launch bound regime): the whole
      // batch in one cooperative launch.  HBM-bound systems keep the 3-kernel loop, whose 64
      // warps/SM stream faster than the 24 warps/SM a co-resident grid allows (0.71 vs 0.97 ms
      // per iteration at 1M quads).
      if (h->n_rank <= 1 && h->coop_blocks > 0 && (size_t)h->sym.nnzb() * 288 <= ((size_t)96 << 20)) {
        int n_row_ = n_row, batch_ = batch;
        // ~2 block rows per warp, at most what fits co-resident: small systems use few blocks
        // so that the grid barriers stay cheap
        const int grid = std::max(1, std::min(h->coop_blocks, cdiv(n_row, 2 * (RED_BLOCK / 32))));
        void* args[] = {&n_row_, &h->rowptr, &h->colidx, &h->vals, &h->vx, &h->vr, &h->vp, &h->vq, &h->sc,
                        &h->partials, &batch_};
        CK(cudaLaunchCooperativeKernel((void*)cg_persistent_kernel, dim3(grid), dim3(RED_BLOCK), args, 0, st));
        LAUNCHED();
        cur = 0;
      } else {
        for (int k = 0; k < batch; ++k) {
          if ((rc = cg_iteration(h, cur, st))) return rc;
          cur ^= 1;
        }
      }
      it_local += batch;
      CK(cudaMemcpyAsync(h->sc_host, h->sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      const CgScalars& s = *h->sc_host;
      const double rr = s.rr[cur];
      if (!(rr == rr) || !(s.bb == s.bb)) {
        if (stats) { stats->iterations = total_iter + s.iter; stats->converged = 0; }
        return fail(h, JSSO_ERR_NAN, "PCG breakdown: NaN or non-positive curvature (matrix not SPD?)");
      }
      if (s.bb == 0.0) { stop = true; converged = true; relres_rec = 0.0; it_local = s.iter; break; }
      relres_rec = std::sqrt(rr / s.bb);
      if (!(rr > s.tol2 * s.bb)) { stop = true; it_local = s.iter; }
    }
    total_iter += it_local;
    // true residual
    if ((rc = halo_exchange_w(h, h->vx, 6, st))) return rc;
    if ((rc = spmv_plain(h, h->vx, h->vq, st))) return rc;
    residual_norm_kernel<<<vblocks, RED_BLOCK, 0, st>>>(n, h->vb, h->vq, h->sc, h->partials, h->counters + 2,
                                                        h->n_rank <= 1);
    CKL("residual_norm_kernel");
    if ((rc = allreduce_scalar(h, &h->sc->loc[3], &h->sc->aux, st))) return rc;
    CK(cudaMemcpyAsync(h->sc_host, h->sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    relres_true = (h->sc_host->bb > 0) ? std::sqrt(h->sc_host->aux / h->sc_host->bb) : 0.0;
    if (relres_true <= o.rtol * 1.5 || h->sc_host->bb == 0.0) { converged = true; break; }
    // the recurrence residual has drifted from the true one: residual replacement, i.e.
    // restart from the current iterate -- unless the last restart no longer gained a
    // factor 2 (attainable accuracy reached)
    if (total_iter >= o.maxiter || restarts >= 20 || relres_true > 0.5 * prev_true) break;
    prev_true = relres_true;
    ++restarts;
  }
  if (stats) {
    stats->iterations = total_iter; stats->restarts = restarts; stats->converged = converged ? 1 : 0;
    stats->relres = relres_true; stats->relres_recur = relres_rec;
  }
  if (!converged) {
    char buf[256];
    std::snprintf(buf, sizeof buf, "PCG did not reach rtol=%.3g: true relres %.3g after %d iterations, %d restarts%s",
                  o.rtol, relres_true, total_iter, restarts,
                  total_iter >= o.maxiter ? " (maxiter)" : " (stagnated: attainable accuracy)");
    return fail(h, JSSO_ERR_NOCONV, buf);
  }
  return JSSO_OK;
}

}  // extern "C" (templates below)

// ---------------------------------------------------------------- multigrid (single GPU)
extern "C" int jsso_mg_setup(jsso_handle* h, int32_t n_levels, const jsso_mg_level_desc* L) {
  if (!h || (n_levels > 0 && !L)) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  if (!h->mg.empty()) return fail(h, JSSO_ERR_STATE, "multigrid hierarchy already set");
  if (h->sym.n_row != h->sym.n_node) return fail(h, JSSO_ERR_STATE, "multigrid: single-GPU handles only");
  int n_prev = h->sym.n_row;
  for (int l = 0; l < n_levels; ++l) {
    const jsso_mg_level_desc& d = L[l];
    if (d.n_f != n_prev) return fail(h, JSSO_ERR_ARG, "multigrid level sizes do not chain");
    jsso_handle::MgLevel m;
    m.n_f = d.n_f; m.n_c = d.n_c; m.nnz_p = d.nnz_p; m.nnz_ap = d.nnz_ap; m.nnz_c = d.nnz_c;
    auto up = [&](int32_t** dst, const int32_t* src, size_t n) -> cudaError_t {
      std::vector<int32_t> v(src, src + n);
      return upload(dst, v);
    };
    std::vector<int32_t> prow(d.nnz_p);
    for (int i = 0; i < d.n_f; ++i)
      for (int s = d.p_rowptr[i]; s < d.p_rowptr[i + 1]; ++s) prow[s] = i;
    CK(upload(&m.p_row, prow));
    CK(up(&m.agg, d.agg, d.n_f)); CK(up(&m.p_rowptr, d.p_rowptr, d.n_f + 1)); CK(up(&m.p_col, d.p_col, d.nnz_p));
    CK(up(&m.p_own, d.p_own, d.nnz_p)); CK(up(&m.ps_ptr, d.ps_ptr, d.nnz_p + 1));
    const size_t n_ps = d.ps_ptr[d.nnz_p], n_apl = d.apl_ptr[d.nnz_ap], n_cl = d.cl_ptr[d.nnz_c];
    CK(up(&m.ps_a, d.ps_a, n_ps)); CK(up(&m.ps_j, d.ps_j, n_ps));
    CK(up(&m.apl_ptr, d.apl_ptr, d.nnz_ap + 1)); CK(up(&m.apl_a, d.apl_a, n_apl)); CK(up(&m.apl_p, d.apl_p, n_apl));
    CK(up(&m.c_rowptr, d.c_rowptr, d.n_c + 1)); CK(up(&m.c_col, d.c_col, d.nnz_c)); CK(up(&m.c_diag, d.c_diag, d.n_c));
    CK(up(&m.cl_ptr, d.cl_ptr, d.nnz_c + 1)); CK(up(&m.cl_p, d.cl_p, n_cl)); CK(up(&m.cl_ap, d.cl_ap, n_cl));
    CK(up(&m.pt_rowptr, d.pt_rowptr, d.n_c + 1)); CK(up(&m.pt_col, d.pt_col, d.nnz_p)); CK(up(&m.pt_src, d.pt_src, d.nnz_p));
    CK(up(&m.mem_ptr, d.mem_ptr, d.n_c + 1)); CK(up(&m.mem, d.mem, d.n_f));
    CK(dalloc(&m.P, 36 * (size_t)d.nnz_p)); CK(dalloc(&m.Pt, 36 * (size_t)d.nnz_p));
    CK(dalloc(&m.AP, 36 * (size_t)d.nnz_ap)); CK(dalloc(&m.Ac, 36 * (size_t)d.nnz_c));
    CK(dalloc(&m.Xc, 3 * (size_t)d.n_c));
    CK(dalloc(&m.P32, 36 * (size_t)d.nnz_p)); CK(dalloc(&m.Pt32, 36 * (size_t)d.nnz_p));
    CK(dalloc(&m.Ac32, 36 * (size_t)d.nnz_c));
    if (l > 0) { CK(dalloc(&m.b, 6 * (size_t)d.n_f)); CK(dalloc(&m.x, 6 * (size_t)d.n_f)); }
    CK(dalloc(&m.Wc, 36 * (size_t)d.n_c)); CK(dalloc(&m.Lc, 36 * (size_t)d.n_c));
    {
      std::vector<int32_t> crow(d.nnz_c);
      for (int i = 0; i < d.n_c; ++i)
        for (int s = d.c_rowptr[i]; s < d.c_rowptr[i + 1]; ++s) crow[s] = i;
      CK(upload(&m.c_row, crow));
    }
    CK(dalloc(&m.r, 6 * (size_t)d.n_f)); CK(dalloc(&m.d, 6 * (size_t)d.n_f));
    h->mg.push_back(m);
    n_prev = d.n_c;
  }
  if (n_levels > 0) {
    const size_t nc = 6 * (size_t)n_prev;
    if (nc > 6000) return fail(h, JSSO_ERR_ARG, "multigrid: coarsest level too large for the dense solve");
    CK(dalloc(&h->mg_dense, 2 * nc * nc)); CK(dalloc(&h->mg_cb, nc)); CK(dalloc(&h->mg_cx, nc));
    CK(dalloc(&h->mg_dense_ws, (size_t)MG_DB * MG_DB + (size_t)MG_DB * 2 * nc + (size_t)MG_DB * nc));
    CK(dalloc(&h->Lfac, 36 * (size_t)h->sym.n_node));
    CK(dalloc(&h->mg_scal, MGS_COUNT));
    CK(cudaMemset(h->mg_scal, 0, MGS_COUNT * sizeof(double)));
    CK(cudaMallocHost((void**)&h->mg_scal_host, MGS_COUNT * sizeof(double)));
  }
  {
    const char* e = std::getenv("JSSO_MG_FP64");   // A/B switch: keep the V-cycle matrices in FP64
    h->mg_fp32 = !(e && e[0] == '1');
    if (const char* ea = std::getenv("JSSO_MG_POLL")) h->mg_poll = std::max(1, std::min(64, std::atoi(ea)));
    if (const char* et = std::getenv("JSSO_MG_TIMING")) h->probe.armed = et[0] == '2';
    if (const char* ep = std::getenv("JSSO_MG_POWER_ITERS")) {   // A/B: 30 restores round 1's estimate (x 1.15)
      h->mg_power_iters = std::max(3, std::min(200, std::atoi(ep)));
      if (h->mg_power_iters >= 30) h->mg_power_safety = 1.15;
    }
    if (const char* eg = std::getenv("JSSO_MG_GRAPH")) h->mg_graph = eg[0] != '0';   // A/B switch (default on)
    // binary16 storage of the fine-level V-cycle matrix (the block-Jacobi-scaled matrix has unit diagonal blocks and
    // |entries| <= 1); JSSO_MG_FP16=0 keeps FP32 (A/B switch)
    const char* e16 = std::getenv("JSSO_MG_FP16");
    h->mg_fp16 = h->mg_fp32 && n_levels > 0 && !(e16 && e16[0] == '0');
    if (h->mg_fp16) CK(dalloc(&h->vals16, 36 * (size_t)h->sym.nnzb()));
    else if (n_levels > 0) CK(dalloc(&h->vals32, 36 * (size_t)h->sym.nnzb()));   // the fine level needs ONE reduced-precision copy
  }
  h->mg_ready = false;
  h->assembled = false;   // the fine factor L is produced by the scaling of the NEXT assembly
  h->scaled = false;
  return JSSO_OK;
}

// Row-range distribution of the multigrid solve over the ranks of one NCCL communicator (see mg_solve_dist).
// The handle must hold the WHOLE mesh, renumbered so that every rank's nodes are one range
// (jaxsso_b200/dist_multigrid.py), with the hierarchy already uploaded by jsso_mg_setup.  Collective: every rank
// calls it (ncclCommInitRank).  bounds: (n_dist + 1) x (n_rank + 1) row-range bounds per level; halo: n_dist entries.
extern "C" int jsso_mg_set_dist(jsso_handle* h, const uint8_t nccl_id[128], int32_t rank, int32_t n_rank, int32_t n_dist,
                                const int32_t* bounds, int32_t n_halo, const jsso_mg_halo_desc* halo) {
  if (!h || !nccl_id || !bounds || (n_dist > 0 && !halo)) return JSSO_ERR_ARG;
  if (n_halo != n_dist && n_halo != n_dist + 1)
    return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist: n_halo is n_dist, or n_dist + 1 with the all-gather plan of the first replicated level");
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  if (h->mg.empty()) return fail(h, JSSO_ERR_STATE, "jsso_mg_set_dist needs the hierarchy (jsso_mg_setup) first");
  if (h->n_rank > 1 || h->sym.n_row != h->sym.n_node)
    return fail(h, JSSO_ERR_STATE, "jsso_mg_set_dist: whole-mesh handles only (the partition is by row ranges)");
  if (h->mgd.comm) return fail(h, JSSO_ERR_STATE, "distributed multigrid already set");
  const int nl = (int)h->mg.size();
  if (n_rank < 2 || rank < 0 || rank >= n_rank || n_dist < 1 || n_dist > nl)
    return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist: need n_rank >= 2, 0 <= rank < n_rank, 1 <= n_dist <= levels");
  jsso_handle::MgDist D;
  D.rank = rank; D.n_rank = n_rank; D.n_dist = n_dist;
  size_t max_send = 0, max_recv = 0;
  for (int l = 0; l <= n_dist; ++l) {
    const int n_l = (l == 0) ? h->sym.n_row : h->mg[l - 1].n_c;
    const int32_t* b = bounds + (size_t)l * (n_rank + 1);
    if (b[0] != 0 || b[n_rank] != n_l) return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist: bounds do not cover a level");
    for (int r = 0; r < n_rank; ++r)
      if (b[r + 1] < b[r]) return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist: bounds not monotone");
    D.bounds.emplace_back(b, b + n_rank + 1);
  }
  for (int l = 0; l < n_halo; ++l) {   // entry n_dist (optional): "halo" = everybody's whole range, i.e. the all-gather
    const jsso_mg_halo_desc& d = halo[l];
    const int n_l = (l == 0) ? h->sym.n_row : h->mg[l - 1].n_c;
    const int lo = D.bounds[l][rank], hi = D.bounds[l][rank + 1];
    jsso_handle::MgDistLevel L;
    if (d.n_peer < 0 || (d.n_peer > 0 && (!d.peer_rank || !d.send_ptr || !d.recv_ptr)))
      return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist: bad halo description");
    for (int i = 0; i < d.n_peer; ++i) {
      if (d.peer_rank[i] < 0 || d.peer_rank[i] >= n_rank || d.peer_rank[i] == rank)
        return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist: bad peer rank");
      L.peers.push_back(jsso_handle::MgDistPeer{d.peer_rank[i], d.send_ptr[i], d.send_ptr[i + 1] - d.send_ptr[i],
                                                d.recv_ptr[i], d.recv_ptr[i + 1] - d.recv_ptr[i]});
    }
    L.n_send = d.n_peer ? d.send_ptr[d.n_peer] : 0;
    L.n_recv = d.n_peer ? d.recv_ptr[d.n_peer] : 0;
    if (d.remote_off) L.remote_off.assign(d.remote_off, d.remote_off + d.n_peer);
    std::vector<int32_t> si(d.send_idx, d.send_idx + L.n_send), ri(d.recv_idx, d.recv_idx + L.n_recv);
    for (int v : si) if (v < lo || v >= hi) return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist: send index not owned");
    for (int v : ri) if (v < 0 || v >= n_l || (v >= lo && v < hi)) return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist: bad receive index");
    CK(upload(&L.send_idx, si)); CK(upload(&L.recv_idx, ri));
    if (l < n_dist && d.n_ghost > 0 && d.ghost_rows) {
      std::vector<int32_t> gr(d.ghost_rows, d.ghost_rows + d.n_ghost);
      for (int v : gr) if (v < 0 || v >= n_l || (v >= lo && v < hi)) return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist: bad ghost row");
      CK(upload(&L.ghost_rows, gr));
      L.n_ghost = d.n_ghost;
    }
    max_send = std::max(max_send, (size_t)L.n_send); max_recv = std::max(max_recv, (size_t)L.n_recv);
    D.lv.push_back(L);
  }
  CK(dalloc(&D.send_buf, 6 * max_send)); CK(dalloc(&D.recv_buf, 6 * max_recv));
  D.max_recv = max_recv;
  if (!nccl_load()) return fail(h, JSSO_ERR_NCCL, "libnccl.so.2 not found");
  ncclUniqueId id;
  std::memcpy(&id, nccl_id, 128);
  CKN(g_nccl.CommInitRank(&D.comm, n_rank, id, rank));
  h->mgd = D;
  return JSSO_OK;
}

// Peer-memory halo exchange / all-reduce of the distributed multigrid solve (optional; after jsso_mg_set_dist).
// export: 128 bytes = CUDA IPC handles of this rank's mailbox and receive arena; connect: everybody's 128 bytes.
extern "C" int jsso_mg_p2p_export(jsso_handle* h, uint8_t out[128]) {
  if (!h || !out) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  jsso_handle::MgDist& D = h->mgd;
  if (D.n_rank < 2 || !D.comm) return fail(h, JSSO_ERR_STATE, "jsso_mg_p2p_export needs jsso_mg_set_dist first");
  if (D.n_rank > P2P_MAX_RANKS || (int)D.lv.size() > MGD_MAX_LEVELS)
    return fail(h, JSSO_ERR_STATE, "peer-memory multigrid: at most 16 ranks and 4 exchange plans (distributed levels + the all-gather)");
  for (const auto& L : D.lv)
    if (L.remote_off.size() != L.peers.size() || L.peers.size() > (size_t)P2P_MAX_RANKS)
      return fail(h, JSSO_ERR_STATE, "peer-memory multigrid: the plan has no remote offsets");
  if (!D.mbox) {
    CK(dalloc(&D.mbox, 1));
    CK(cudaMemset(D.mbox, 0, sizeof(MgdMailbox)));
    CK(dalloc(&D.arena, (size_t)MGD_MAX_LEVELS * 2 * 6 * std::max<size_t>(D.max_recv, 1)));
    CK(dalloc(&D.push_counter, 1));
    CK(cudaMemset(D.push_counter, 0, sizeof(unsigned)));
    CK(dalloc(&D.red_seq_dev, 1));
    CK(cudaMemset(D.red_seq_dev, 0, sizeof(unsigned long long)));
  }
  cudaIpcMemHandle_t a, b;
  CK(cudaIpcGetMemHandle(&a, D.mbox));
  CK(cudaIpcGetMemHandle(&b, D.arena));
  std::memcpy(out, &a, 64);
  std::memcpy(out + 64, &b, 64);
  return JSSO_OK;
}

extern "C" int jsso_mg_p2p_connect(jsso_handle* h, const uint8_t* all_handles, const int32_t* max_recv_all) {
  if (!h || !all_handles || !max_recv_all) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  jsso_handle::MgDist& D = h->mgd;
  if (!D.mbox) return fail(h, JSSO_ERR_STATE, "jsso_mg_p2p_connect needs jsso_mg_p2p_export first");
  // every rank sizes its arena by its own max_recv; the strides a pusher uses are the RECEIVER's, so all ranks
  // agree on one stride: the maximum over the ranks
  size_t mr = 0;
  for (int r = 0; r < D.n_rank; ++r) mr = std::max(mr, (size_t)max_recv_all[r]);
  if (mr > D.max_recv) {   // re-size my arena to the common stride (before anybody maps it: connect is collective)
    return fail(h, JSSO_ERR_ARG, "jsso_mg_p2p_connect: export with the common max_recv (jsso_mg_p2p_reserve) first");
  }
  MgdCtx c;
  std::memset(&c, 0, sizeof c);
  c.rank = D.rank; c.n_rank = D.n_rank;
  c.arena_slot_stride = 6LL * (long long)D.max_recv;
  c.arena_level_stride = 2 * c.arena_slot_stride;
  for (int r = 0; r < D.n_rank; ++r) {
    if (r == D.rank) { c.mbox[r] = D.mbox; c.arena[r] = D.arena; continue; }
    cudaIpcMemHandle_t mh, ah;
    std::memcpy(&mh, all_handles + 128 * (size_t)r, 64);
    std::memcpy(&ah, all_handles + 128 * (size_t)r + 64, 64);
    void *pm = nullptr, *pa = nullptr;
    CK(cudaIpcOpenMemHandle(&pm, mh, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle(&pa, ah, cudaIpcMemLazyEnablePeerAccess));
    D.ipc_opened.push_back(pm); D.ipc_opened.push_back(pa);
    c.mbox[r] = (MgdMailbox*)pm; c.arena[r] = (double*)pa;
  }
  for (auto& L : D.lv) {
    MgdLevelDev ld;
    std::memset(&ld, 0, sizeof ld);
    ld.n_peer = (int)L.peers.size(); ld.n_send = L.n_send; ld.n_recv = L.n_recv;
    for (int i = 0; i < ld.n_peer; ++i) {
      ld.peer_rank[i] = L.peers[i].rank; ld.send_off[i] = L.peers[i].send_off; ld.send_cnt[i] = L.peers[i].send_cnt;
      ld.remote_off[i] = L.remote_off[i];
    }
    CK(dalloc(&L.dev, 1));
    CK(cudaMemcpy(L.dev, &ld, sizeof ld, cudaMemcpyHostToDevice));
  }
  CK(dalloc(&D.ctx, 1));
  CK(cudaMemcpy(D.ctx, &c, sizeof c, cudaMemcpyHostToDevice));
  return JSSO_OK;
}

// the arena stride must be the same on every rank: call with the maximum of max_recv over the ranks BEFORE export
extern "C" int jsso_mg_p2p_reserve(jsso_handle* h, int32_t max_recv_common) {
  if (!h || max_recv_common < 0) return JSSO_ERR_ARG;
  if (h->mgd.mbox) return fail(h, JSSO_ERR_STATE, "jsso_mg_p2p_reserve after export");
  h->mgd.max_recv = std::max(h->mgd.max_recv, (size_t)max_recv_common);
  return JSSO_OK;
}

// statistics of the distributed solve since jsso_mg_set_dist: out[0] = halo exchanges, out[1] = scalar all-reduces,
// out[2] = 1 if they run over peer memory (jsso_mg_p2p_connect), 0 over NCCL
extern "C" int jsso_mg_dist_counters(const jsso_handle* h, int64_t* out) {
  if (!h || !out) return JSSO_ERR_ARG;
  out[0] = h->mgd.n_exchange; out[1] = h->mgd.n_allreduce; out[2] = h->mgd.ctx ? 1 : 0;
  return JSSO_OK;
}

static inline void mgd_range(const jsso_handle* h, int l, int& s, int& n);
extern "C" int jsso_mg_set_dist_setup(jsso_handle* h, int32_t n_dist, const jsso_mg_setup_desc* desc) {
  if (!h || !desc) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  jsso_handle::MgDist& D = h->mgd;
  if (D.n_rank < 2 || !D.comm) return fail(h, JSSO_ERR_STATE, "jsso_mg_set_dist_setup needs jsso_mg_set_dist first");
  if (n_dist != D.n_dist) return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist_setup: one descriptor per distributed level");
  if (!h->asm_tasks && h->sym.n_quad + h->sym.n_beam > 0)
    return fail(h, JSSO_ERR_STATE, "jsso_mg_set_dist_setup: needs the warp-task assembly (partial assembly by task ranges)");
  D.setup.clear();
  for (int l = 0; l < n_dist; ++l) {
    const jsso_mg_setup_desc& d = desc[l];
    const jsso_handle::MgLevel& m = h->mg[l];
    if (d.n_p_slots < 0 || d.n_ap_slots < 0 || !d.ac_bounds || (d.n_p_slots && !d.p_slots) || (d.n_ap_slots && !d.ap_slots))
      return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist_setup: bad descriptor");
    jsso_handle::MgDist::SetupLevel L;
    std::vector<int32_t> pl(d.p_slots, d.p_slots + d.n_p_slots), al(d.ap_slots, d.ap_slots + d.n_ap_slots);
    for (int v : pl) if (v < 0 || v >= m.nnz_p) return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist_setup: P slot out of range");
    for (int v : al) if (v < 0 || v >= m.nnz_ap) return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist_setup: AP slot out of range");
    CK(upload(&L.p_list, pl)); CK(upload(&L.ap_list, al));
    L.n_p = d.n_p_slots; L.n_ap = d.n_ap_slots;
    L.ac_bounds.assign(d.ac_bounds, d.ac_bounds + D.n_rank + 1);
    L.p_lo = d.p_own_lo; L.p_hi = d.p_own_hi; L.pt_lo = d.pt_own_lo; L.pt_hi = d.pt_own_hi;
    if (!(0 <= L.p_lo && L.p_lo <= L.p_hi && L.p_hi <= m.nnz_p && 0 <= L.pt_lo && L.pt_lo <= L.pt_hi && L.pt_hi <= m.nnz_p))
      return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist_setup: bad own slot ranges");
    if (L.ac_bounds[0] != 0 || L.ac_bounds[D.n_rank] != m.nnz_c) return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist_setup: ac_bounds do not cover the coarse matrix");
    D.setup.push_back(L);
  }
  const int n_row = h->sym.n_row;
  D.sc_lo = desc[0].scale_row_lo; D.sc_hi = desc[0].scale_row_hi;
  D.w_lo = desc[0].factor_row_lo; D.w_hi = desc[0].factor_row_hi;
  int s0, n0;
  mgd_range(h, 0, s0, n0);
  if (!(0 <= D.w_lo && D.w_lo <= D.sc_lo && D.sc_lo <= s0 && s0 + n0 <= D.sc_hi && D.sc_hi <= D.w_hi && D.w_hi <= n_row))
    return fail(h, JSSO_ERR_ARG, "jsso_mg_set_dist_setup: row hulls must nest: factor >= scale >= own rows");
  // assembly: the tasks (runs of block slots) that touch the rows [w_lo, w_hi), and the quads they stage
  const Symbolic& S = h->sym;
  const int nt = S.n_task();
  const int slot_lo = S.rowptr[D.w_lo], slot_hi = S.rowptr[D.w_hi];
  int t0 = 0, t1 = nt;
  while (t0 + 1 < nt && S.task_meta[4 * (size_t)(t0 + 1)] <= slot_lo) ++t0;
  while (t1 > t0 + 1 && S.task_meta[4 * (size_t)(t1 - 1)] >= slot_hi) --t1;
  int q0 = S.n_quad, q1 = 0;
  for (int t = t0; t < t1; ++t) {
    const int el0 = S.task_meta[4 * (size_t)t + 2], n_el = (S.task_meta[4 * (size_t)t + 3] >> 16) & 255;
    for (int k = 0; k < n_el; ++k) { const int e = S.task_els[el0 + k]; q0 = std::min(q0, e); q1 = std::max(q1, e + 1); }
  }
  if (q1 < q0) { q0 = 0; q1 = 0; }
  D.asm_t0 = t0; D.asm_t1 = t1; D.asm_q0 = q0; D.asm_q1 = q1;
  D.setup_on = true;
  h->assembled = false; h->scaled = false; h->mg_ready = false;
  return JSSO_OK;
}

struct MgMat { const int32_t* rp; const int32_t* ci; const double* v; int n; const float* v32; const __half* v16; long long nnz; };
static MgMat mg_matrix(jsso_handle* h, int l) {
  if (l == 0) return MgMat{h->rowptr, h->colidx, h->vals, h->sym.n_row, h->vals32, h->mg_fp16 ? h->vals16 : nullptr, h->sym.nnzb()};
  const jsso_handle::MgLevel& p = h->mg[l - 1];
  return MgMat{p.c_rowptr, p.c_col, p.Ac, p.n_c, p.Ac32, nullptr, p.nnz_c};
}
static inline int mg_blocks(jsso_handle* h, int n_row) {
  return std::max(1, std::min(h->spmv_blocks, cdiv(n_row, RED_BLOCK / 32)));
}
template <int MODE>
static int mg_spmv(jsso_handle* h, const int32_t* rp, const int32_t* ci, const double* v, int n_row,
                   const double* x, double* y, const double* b, cudaStream_t st) {
  if (n_row == 0) return JSSO_OK;
  bsr_spmv_axpby_kernel<MODE, double><<<mg_blocks(h, n_row), RED_BLOCK, 0, st>>>(n_row, rp, ci, v, x, y, b);
  CKL("bsr_spmv_axpby_kernel");
  return JSSO_OK;
}
#ifndef JSSO_MG_RP
#define JSSO_MG_RP 1   // FP32 / binary16 products by the row-pair kernel (bsr_spmv_rp_kernel); 0: warp-per-row kernels (A/B builds)
#endif
static inline int rp_grid(jsso_handle* h, int n_row) { return std::max(1, std::min(h->rp_blocks, cdiv(3LL * n_row, RED_BLOCK))); }
// same with single-precision (or binary16) block storage (falls back to the FP64 values when fp32 is off)
template <int MODE>
static int mg_spmv_p(jsso_handle* h, const int32_t* rp, const int32_t* ci, const double* v, const float* v32,
                     int n_row, const double* x, double* y, const double* b, cudaStream_t st,
                     bool short_rows = false, const __half* v16 = nullptr, double blocks_per_row = 0.0) {
  if (!h->mg_fp32 || (!v32 && !v16)) return mg_spmv<MODE>(h, rp, ci, v, n_row, x, y, b, st);
  if (n_row == 0) return JSSO_OK;
#if JSSO_MG_RP
  // thread per (row, row pair) when the rows are short enough for a thread to stream one alone and there are enough
  // of them to fill the GPU (measured at 1M quads: 9-block rows 180 vs 306 us); long rows (restrictions: ~26 blocks
  // per coarse row) and the small coarse levels keep a warp per row
  if (short_rows || v16 || (blocks_per_row > 0.0 && blocks_per_row <= 12.0 && n_row >= 4096)) {
    // MODE 0: y = A x;  2: y = b - A x;  3: y += A x   as  ca * b + cb * y_row + cc * A x
    const double ca = (MODE == 2) ? 1.0 : 0.0, cb = (MODE == 3) ? 1.0 : 0.0, cc = (MODE == 2) ? -1.0 : 1.0;
    const double* bv = (MODE == 2) ? b : nullptr;
    const double* xr = (MODE == 3) ? y : nullptr;
    if (v16 && !short_rows)
      bsr_spmv_rp_kernel<__half, 0><<<rp_grid(h, n_row), RED_BLOCK, 0, st>>>(n_row, rp, ci, v16, x, y, bv, xr, ca, cb, cc, nullptr,
                                                                            h->partials, h->counters + 2, nullptr);
    else
      bsr_spmv_rp_kernel<float, 0><<<rp_grid(h, n_row), RED_BLOCK, 0, st>>>(n_row, rp, ci, v32, x, y, bv, xr, ca, cb, cc, nullptr,
                                                                           h->partials, h->counters + 2, nullptr);
    CKL("bsr_spmv_rp_kernel");
    return JSSO_OK;
  }
#endif
  if (v16 && !short_rows) {   // fine level, binary16 blocks
    bsr_spmv_axpby_kernel<MODE, __half><<<mg_blocks(h, n_row), RED_BLOCK, 0, st>>>(n_row, rp, ci, v16, x, y, b);
    CKL("bsr_spmv_axpby_kernel<half>");
    return JSSO_OK;
  }
  if (short_rows) {   // prolongators: a thread per (row, row pair)
    bsr_spmv_short_kernel<MODE><<<cdiv(3LL * n_row, 256), 256, 0, st>>>(n_row, rp, ci, v32, x, y, b);
    CKL("bsr_spmv_short_kernel");
    return JSSO_OK;
  }
  bsr_spmv_axpby_kernel<MODE, float><<<mg_blocks(h, n_row), RED_BLOCK, 0, st>>>(n_row, rp, ci, v32, x, y, b);
  CKL("bsr_spmv_axpby_kernel<float>");
  return JSSO_OK;
}
static int mg_to_float(jsso_handle* h, long long n, const double* a, float* b, cudaStream_t st) {
  if (n == 0) return JSSO_OK;
  mg_to_float_kernel<<<std::max(1, std::min(1184, cdiv(n, 256))), 256, 0, st>>>(n, a, b);
  CKL("mg_to_float_kernel");
  return JSSO_OK;
}
static int mg_dot(jsso_handle* h, long long n, const double* a, const double* b, int slot, cudaStream_t st) {
  const int blocks = std::max(1, std::min(h->red_blocks, cdiv(n, 256)));
  mg_dot_kernel<<<blocks, 256, 0, st>>>(n, a, b, h->partials, h->counters + 2, h->mg_scal + slot);
  CKL("mg_dot_kernel");
  return JSSO_OK;
}
static int mg_read_scalars(jsso_handle* h, cudaStream_t st) {
  CK(cudaMemcpyAsync(h->mg_scal_host, h->mg_scal, MGS_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return JSSO_OK;
}

static inline void mgd_range(const jsso_handle* h, int l, int& s, int& n);
static int mgd_exchange(jsso_handle* h, int l, double* v, cudaStream_t st);
static int mgd_reduce_read(jsso_handle* h, int slot, int count, cudaStream_t st);
static int mg_coarse_graphed(jsso_handle* h, int l, double* b, double* x, cudaStream_t st);
static int mgd_allgather(jsso_handle* h, int l, double* v, cudaStream_t st);

// The coarse matrix of coarsening step l in block-Jacobi-scaled form: factor its diagonal blocks (W_c = L_c^-1),
// A_c <- W_c A_c W_c^T in place, and the prolongator into the scaled coarse coordinates, P <- P W_c^T (the blocks
// of `p_list`, or all n_p of them).  The scaled matrix has unit diagonal blocks, so the smoother of the next level
// needs no D^-1 and its V-cycle products take the fused form of the fine level (mg_vcycle_fused).
static int mg_scale_coarse(jsso_handle* h, int l, const int32_t* p_list, int n_p, cudaStream_t st) {
  jsso_handle::MgLevel& m = h->mg[l];
  if (m.n_c > 0) {
    diag_factor_kernel<<<cdiv(m.n_c, 128), 128, 0, st>>>(m.n_c, m.c_diag, m.Ac, m.Wc, m.Lc, h->flags, 1);
    CKL("diag_factor_kernel");
  }
  if (m.nnz_c > 0) {
    scale_blocks_kernel<<<cdiv(m.nnz_c, 128), 128, 0, st>>>(m.nnz_c, m.c_row, m.c_col, m.Wc, m.Ac);
    CKL("scale_blocks_kernel");
  }
  if (n_p > 0) {
    mg_scale_cols_kernel<<<cdiv(n_p, 128), 128, 0, st>>>(n_p, m.p_col, m.Wc, m.P, p_list);
    CKL("mg_scale_cols_kernel");
  }
  return JSSO_OK;
}

// JSSO_MG_TIMING=1: synchronise between the phases of the numeric setup and print their wall times (diagnostics)
struct PhaseTimer {
  bool on; cudaStream_t st; double t0; std::string out; int rank;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
  PhaseTimer(cudaStream_t s, int r) : st(s), rank(r) {
    const char* e = std::getenv("JSSO_MG_TIMING");
    on = e && e[0] == '1';
    if (on) { cudaStreamSynchronize(st); t0 = now(); }
  }
  void mark(const char* name, int l = -1) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const double t = now();
    char buf[96];
    if (l >= 0) std::snprintf(buf, sizeof buf, " %s[%d]=%.3f", name, l, 1e3 * (t - t0));
    else std::snprintf(buf, sizeof buf, " %s=%.3f", name, 1e3 * (t - t0));
    out += buf;
    t0 = now();
  }
  ~PhaseTimer() { if (on && rank == 0) std::fprintf(stderr, "JSSO_MG_TIMING (ms):%s\n", out.c_str()); }
};

// numeric hierarchy for the current (block-Jacobi-scaled) matrix
static int mg_numeric_setup(jsso_handle* h, cudaStream_t st) {
  if (h->mg_ready) return JSSO_OK;
  PhaseTimer pt_(st, h->mgd.rank);
  if (h->mg_graph_exec) { cudaGraphExecDestroy(h->mg_graph_exec); h->mg_graph_exec = nullptr; }   // coefficients change
  if (!h->last_crds) return fail(h, JSSO_ERR_STATE, "multigrid needs the coordinates of the last jsso_assemble");
  const int nl = (int)h->mg.size();
  const double* X = h->last_crds;
  int rc;
  {
    // reduced-precision copy of the fine matrix for the V-cycle: all rows, or this rank's own rows (distributed setup).
    // First, because the power iterations below already run on it
    long long v0 = 0, v1 = h->sym.nnzb();
    if (h->mgd.setup_on) {
      int s0, n0;
      mgd_range(h, 0, s0, n0);
      v0 = h->sym.rowptr[s0]; v1 = h->sym.rowptr[s0 + n0];
    }
    if (h->mg_fp32 && nl > 0 && h->vals32) {
      if ((rc = mg_to_float(h, 36LL * (v1 - v0), h->vals + 36 * v0, h->vals32 + 36 * v0, st))) return rc;
    }
    if (h->mg_fp16 && nl > 0 && v1 > v0) {
      const long long n16 = 36LL * (v1 - v0);
      mg_to_half_kernel<<<std::max(1, std::min(1184, cdiv(n16, 256))), 256, 0, st>>>(n16, h->vals + 36 * v0, h->vals16 + 36 * v0);
      CKL("mg_to_half_kernel");
    }
  }
  pt_.mark("fine_copy");
  for (int l = 0; l < nl; ++l) {
    jsso_handle::MgLevel& m = h->mg[l];
    const MgMat A = mg_matrix(h, l);
    const int n = m.n_f;
    // every level matrix is block-Jacobi scaled (unit diagonal blocks): no D^-1 anywhere.  Lf: the Cholesky factors of
    // this level's unscaled diagonal blocks (the rigid-body modes enter the prolongator as L^T T)
    const double* Lf = (l == 0) ? h->Lfac : h->mg[l - 1].Lc;
    // lambda_max(D^-1 A): power iteration from a pseudo-random vector; an underestimate would make the Chebyshev
    // smoother amplify the top modes and the V-cycle indefinite.  The estimate grows monotonically: 0.93-0.95 of the
    // true value after 10 steps, 0.98 after 30 (CPU study on three levels of a 64^2 plate), so 10 steps x 1.2 is as
    // safe as the 30 steps x 1.15 of round 1 at a third of the cost (the 150 SpMVs were 22 of the 40 ms of a numeric
    // setup at 1M quads)
    const long long nd = 6LL * n;
    const int vb = std::max(1, std::min(h->red_blocks, cdiv(nd, 256)));
    mg_hash_fill_kernel<<<vb, 256, 0, st>>>(nd, m.r);
    CKL("mg_hash_fill_kernel");
    double lam = 1.0;
    // on the storage the V-cycle runs on (binary16 / FP32: the smoother's operator, and 0.18 instead of 0.46 ms per
    // product at 1M quads); the estimate needs two digits
    const double bpr = (double)A.nnz / std::max(A.n, 1);
    const bool dist_pow = h->mgd.n_rank > 1 && l < h->mgd.n_dist;
    if (dist_pow) {
      // distributed levels: the power iteration (30 SpMVs, 18 of the 42 ms of this setup at 1M quads) by row
      // ranges; the start vector is a function of the index, so it is whole on every rank and only the
      // normalised iterates are exchanged.  The all-reduced norms make lambda bitwise identical on all ranks.
      int rs, rn;
      mgd_range(h, l, rs, rn);
      const size_t off = 6 * (size_t)rs;
      const long long ndl = 6LL * rn;
      const int vbl = std::max(1, std::min(h->red_blocks, cdiv(ndl, 256)));
      for (int it = 0; it < h->mg_power_iters; ++it) {
        if (it > 0) { if ((rc = mgd_exchange(h, l, m.r, st))) return rc; }
        if ((rc = mg_spmv_p<0>(h, A.rp + rs, A.ci, A.v, A.v32, rn, m.r, m.d + off, nullptr, st, false, A.v16, bpr))) return rc;
        if ((rc = mg_dot(h, ndl, m.d + off, m.d + off, 0, st))) return rc;
        if ((rc = mg_dot(h, ndl, m.r + off, m.r + off, 1, st))) return rc;
        if ((rc = mgd_reduce_read(h, 0, 2, st))) return rc;
        lam = std::sqrt(h->mg_scal_host[0] / h->mg_scal_host[1]);
        if (!(lam > 0.0) || !(lam == lam)) return fail(h, JSSO_ERR_NAN, "multigrid: power iteration broke down");
        mg_axpby_kernel<<<vbl, 256, 0, st>>>(ndl, 1.0 / std::sqrt(h->mg_scal_host[0]), m.d + off, 0.0, m.r + off);
        CKL("mg_axpby_kernel");
      }
    }
    for (int it = 0; it < h->mg_power_iters && !dist_pow; ++it) {
      if ((rc = mg_spmv_p<0>(h, A.rp, A.ci, A.v, A.v32, n, m.r, m.d, nullptr, st, false, A.v16, bpr))) return rc;
      if ((rc = mg_dot(h, nd, m.d, m.d, 0, st))) return rc;
      if ((rc = mg_dot(h, nd, m.r, m.r, 5, st))) return rc;
      if ((rc = mg_read_scalars(h, st))) return rc;
      lam = std::sqrt(h->mg_scal_host[0] / h->mg_scal_host[5]);
      if (!(lam > 0.0) || !(lam == lam)) return fail(h, JSSO_ERR_NAN, "multigrid: power iteration broke down");
      mg_axpby_kernel<<<vb, 256, 0, st>>>(nd, 1.0 / std::sqrt(h->mg_scal_host[0]), m.d, 0.0, m.r);
      CKL("mg_axpby_kernel");
    }
    m.lam = h->mg_power_safety * lam;
    pt_.mark("power", l);
    mg_centroid_kernel<<<cdiv(m.n_c, 128), 128, 0, st>>>(m.n_c, m.mem_ptr, m.mem, X, m.Xc);
    CKL("mg_centroid_kernel");
    if (h->mgd.setup_on && l < h->mgd.n_dist) {
      // distributed numeric setup of this level (jsso_mg_set_dist_setup): the Galerkin blocks of the own coarse rows
      // and the P / AP blocks they read (ghost rows recomputed); the coarse matrix is all-gathered by slot ranges
      const jsso_handle::MgDist::SetupLevel& SL = h->mgd.setup[l];
      const int me = h->mgd.rank;
      if (SL.n_p > 0) {
        mg_smooth_prolongator_kernel<<<cdiv(SL.n_p, 128), 128, 0, st>>>(
            SL.n_p, m.p_row, m.p_col, m.p_own, m.ps_ptr, m.ps_a, m.ps_j, m.agg, A.v, nullptr, Lf,
            l == 0 ? h->node_mask : nullptr, X, m.Xc, 4.0 / (3.0 * m.lam), m.P, SL.p_list);
        CKL("mg_smooth_prolongator_kernel");
      }
      // restriction rows of the own coarse nodes: P^T blocks [pt0, pt1); prolongation rows: P blocks [pr0, pr1)
      const int pt[2] = {SL.pt_lo, SL.pt_hi}, pr[2] = {SL.p_lo, SL.p_hi};
      if (SL.n_ap > 0) {
        mg_block_product_kernel<0><<<cdiv(6LL * (SL.n_ap), MG_PROD_THREADS), MG_PROD_THREADS, 0, st>>>(SL.n_ap, m.apl_ptr, m.apl_a, m.apl_p, A.v, m.P, m.AP,
                                                                       SL.ap_list, 0);
        CKL("mg_block_product_kernel<0>");
      }
      const int a0 = SL.ac_bounds[me], a1 = SL.ac_bounds[me + 1];
      if (a1 > a0) {
        mg_block_product_kernel<1><<<cdiv(6LL * (a1 - a0), MG_PROD_THREADS), MG_PROD_THREADS, 0, st>>>(a1 - a0, m.cl_ptr, m.cl_p, m.cl_ap, m.P, m.AP, m.Ac,
                                                                       nullptr, a0);
        CKL("mg_block_product_kernel<1>");
      }
      pt_.mark("P_AP_Ac", l);
      CKN(g_nccl.GroupStart());
      for (int r = 0; r < h->mgd.n_rank; ++r) {
        if (r == me) continue;
        const size_t mine = 36 * (size_t)(a1 - a0), theirs = 36 * (size_t)(SL.ac_bounds[r + 1] - SL.ac_bounds[r]);
        if (mine) CKN(g_nccl.Send(m.Ac + 36 * (size_t)a0, mine, ncclDouble, r, h->mgd.comm, st));
        if (theirs) CKN(g_nccl.Recv(m.Ac + 36 * (size_t)SL.ac_bounds[r], theirs, ncclDouble, r, h->mgd.comm, st));
      }
      CKN(g_nccl.GroupEnd());
      pt_.mark("allgather_Ac", l);
      // the coarse level in scaled form: factor its diagonal blocks, A_c <- W_c A_c W_c^T, P <- P W_c^T (then P^T)
      if ((rc = mg_scale_coarse(h, l, SL.p_list, SL.n_p, st))) return rc;
      if (pt[1] > pt[0]) {
        mg_transpose_blocks_kernel<<<cdiv(36LL * (pt[1] - pt[0]), 256), 256, 0, st>>>(pt[1] - pt[0], m.pt_src, m.P, m.Pt, pt[0]);
        CKL("mg_transpose_blocks_kernel");
      }
      if (h->mg_fp32) {
        if ((rc = mg_to_float(h, 36LL * (pr[1] - pr[0]), m.P + 36 * (size_t)pr[0], m.P32 + 36 * (size_t)pr[0], st))) return rc;
        if ((rc = mg_to_float(h, 36LL * (pt[1] - pt[0]), m.Pt + 36 * (size_t)pt[0], m.Pt32 + 36 * (size_t)pt[0], st))) return rc;
        if ((rc = mg_to_float(h, 36LL * m.nnz_c, m.Ac, m.Ac32, st))) return rc;
      }
      pt_.mark("scale_transpose_convert", l);
      X = m.Xc;
      continue;
    }
    mg_smooth_prolongator_kernel<<<cdiv(m.nnz_p, 128), 128, 0, st>>>(
        m.nnz_p, m.p_row, m.p_col, m.p_own, m.ps_ptr, m.ps_a, m.ps_j, m.agg, A.v, nullptr, Lf,
        l == 0 ? h->node_mask : nullptr, X, m.Xc, 4.0 / (3.0 * m.lam), m.P);
    CKL("mg_smooth_prolongator_kernel");
    mg_block_product_kernel<0><<<cdiv(6LL * (m.nnz_ap), MG_PROD_THREADS), MG_PROD_THREADS, 0, st>>>(m.nnz_ap, m.apl_ptr, m.apl_a, m.apl_p, A.v, m.P, m.AP);
    CKL("mg_block_product_kernel<0>");
    mg_block_product_kernel<1><<<cdiv(6LL * (m.nnz_c), MG_PROD_THREADS), MG_PROD_THREADS, 0, st>>>(m.nnz_c, m.cl_ptr, m.cl_p, m.cl_ap, m.P, m.AP, m.Ac);
    CKL("mg_block_product_kernel<1>");
    pt_.mark("P_AP_Ac", l);
    if ((rc = mg_scale_coarse(h, l, nullptr, m.nnz_p, st))) return rc;
    mg_transpose_blocks_kernel<<<cdiv(36LL * m.nnz_p, 256), 256, 0, st>>>(m.nnz_p, m.pt_src, m.P, m.Pt);
    CKL("mg_transpose_blocks_kernel");
    if (h->mg_fp32) {
      if ((rc = mg_to_float(h, 36LL * m.nnz_p, m.P, m.P32, st))) return rc;
      if ((rc = mg_to_float(h, 36LL * m.nnz_p, m.Pt, m.Pt32, st))) return rc;
      if ((rc = mg_to_float(h, 36LL * m.nnz_c, m.Ac, m.Ac32, st))) return rc;
    }
    pt_.mark("scale_transpose_convert", l);
    X = m.Xc;
  }
  // coarsest level: dense inverse
  const MgMat C = mg_matrix(h, nl);
  const int nc = 6 * C.n;
  mg_dense_from_bsr_kernel<<<std::max(1, std::min(1184, cdiv(2LL * nc * nc, 256))), 256, 0, st>>>(C.n, C.rp, C.ci, C.v, h->mg_dense);
  CKL("mg_dense_from_bsr_kernel");
  mg_dense_fill_kernel<<<C.n, 64, 0, st>>>(C.n, C.rp, C.ci, C.v, h->mg_dense);
  CKL("mg_dense_fill_kernel");
  static const int dense_single_max = [] { const char* e = std::getenv("JSSO_MG_DENSE_SINGLE_MAX"); return e ? std::atoi(e) : 96; }();
  if (nc <= dense_single_max) {
    mg_dense_invert_kernel<<<1, 1024, 0, st>>>(nc, h->mg_dense);
    CKL("mg_dense_invert_kernel");
  } else {
    // blocked Gauss-Jordan over all SMs (see jsso_multigrid.cuh)
    const int w = 2 * nc;
    double* pinv = h->mg_dense_ws;
    double* rpanel = pinv + MG_DB * MG_DB;
    double* cpanel = rpanel + (size_t)MG_DB * w;
    for (int k0 = 0; k0 < nc; k0 += MG_DB) {
      const int nb = std::min(MG_DB, nc - k0);
      mg_dense_pivot_kernel<<<1, MG_DB * MG_DB, 0, st>>>(w, k0, nb, h->mg_dense, pinv);
      CKL("mg_dense_pivot_kernel");
      mg_dense_panel_kernel<<<cdiv((long long)w + (long long)nc * MG_DB, 256), 256, 0, st>>>(nc, w, k0, nb, h->mg_dense, pinv, rpanel, cpanel);
      CKL("mg_dense_panel_kernel");
      mg_dense_update_kernel<<<cdiv(w, MG_DU_COLS) * cdiv(nc, MG_DU_ROWS), MG_DU_COLS, 0, st>>>(nc, w, k0, nb, h->mg_dense, rpanel, cpanel);
      CKL("mg_dense_update_kernel");
    }
  }
  pt_.mark("dense_inverse");
  h->mg_ready = true;
  return JSSO_OK;
}

// Chebyshev(deg) on D^-1 A with eigenvalues in [lam/4, lam]; x updated in place
static int mg_smooth(jsso_handle* h, int l, const double* b, double* x, bool zero_guess, int deg, cudaStream_t st) {
  jsso_handle::MgLevel& m = h->mg[l];
  const MgMat A = mg_matrix(h, l);
  const int n = m.n_f, nb = cdiv(n, 128);
  const double lmax = m.lam, lmin = m.lam / 4.0;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho = 1.0 / sigma;
  int rc;
  if (zero_guess) {
    mg_cheb_kernel<1><<<nb, 128, 0, st>>>(n, m.Dinv, b, m.d, x, 0.0, 1.0 / theta, 1);
  } else {
    if ((rc = mg_spmv_p<2>(h, A.rp, A.ci, A.v, A.v32, n, x, m.r, b, st, false, A.v16, (double)A.nnz / std::max(A.n, 1)))) return rc;
    mg_cheb_kernel<1><<<nb, 128, 0, st>>>(n, m.Dinv, m.r, m.d, x, 0.0, 1.0 / theta, 0);
  }
  CKL("mg_cheb_kernel<1>");
  for (int k = 1; k < deg; ++k) {
    if ((rc = mg_spmv_p<2>(h, A.rp, A.ci, A.v, A.v32, n, x, m.r, b, st, false, A.v16, (double)A.nnz / std::max(A.n, 1)))) return rc;
    const double rho_new = 1.0 / (2.0 * sigma - rho);
    mg_cheb_kernel<0><<<nb, 128, 0, st>>>(n, m.Dinv, m.r, m.d, x, rho_new * rho, 2.0 * rho_new / delta, 0);
    CKL("mg_cheb_kernel<0>");
    rho = rho_new;
  }
  return JSSO_OK;
}

static int mg_vcycle(jsso_handle* h, int l, const double* b, double* x, int deg, cudaStream_t st) {
  const int nl = (int)h->mg.size();
  int rc;
  if (l == nl) {
    const int nc = 6 * mg_matrix(h, nl).n;
    mg_dense_matvec_kernel<<<cdiv(32LL * nc, 256), 256, 0, st>>>(nc, h->mg_dense, b, x);
    CKL("mg_dense_matvec_kernel");
    return JSSO_OK;
  }
  jsso_handle::MgLevel& m = h->mg[l];
  const MgMat A = mg_matrix(h, l);
  double* bc = (l + 1 < nl) ? h->mg[l + 1].b : h->mg_cb;
  double* xc = (l + 1 < nl) ? h->mg[l + 1].x : h->mg_cx;
  if ((rc = mg_smooth(h, l, b, x, true, deg, st))) return rc;
  if ((rc = mg_spmv_p<2>(h, A.rp, A.ci, A.v, A.v32, m.n_f, x, m.r, b, st, false, A.v16, (double)A.nnz / std::max(A.n, 1)))) return rc;          // r = b - A x
  if ((rc = mg_spmv_p<0>(h, m.pt_rowptr, m.pt_col, m.Pt, m.Pt32, m.n_c, m.r, bc, nullptr, st, false, nullptr, (double)m.nnz_p / std::max(m.n_c, 1)))) return rc;   // b_c = P^T r
  if ((rc = mg_vcycle(h, l + 1, bc, xc, deg, st))) return rc;
  if ((rc = mg_spmv_p<3>(h, m.p_rowptr, m.p_col, m.P, m.P32, m.n_f, xc, x, nullptr, st,
                         m.nnz_p <= 5LL * m.n_f))) return rc;                                        // x += P x_c
  return mg_smooth(h, l, b, x, false, deg, st);
}

// The V-cycle from level l as ONE graph launch (JSSO_MG_GRAPH=1): its ~8 kernels per level are a few microseconds
// each on the coarse levels, i.e. bound by launch latency; a graph replays them back to back.  Nothing in
// mg_vcycle synchronises, allocates or copies, so it can be captured as it is.
static int mg_vcycle_graphed(jsso_handle* h, int l, const double* b, double* x, int deg, cudaStream_t st) {
  if (!h->mg_graph) return mg_vcycle(h, l, b, x, deg, st);
  if (h->mg_graph_exec && (h->mg_graph_level != l || h->mg_graph_b != b || h->mg_graph_x != x || h->mg_graph_deg != deg)) {
    cudaGraphExecDestroy(h->mg_graph_exec);
    h->mg_graph_exec = nullptr;
  }
  if (!h->mg_graph_exec) {
    if (!h->st_cap) CK(cudaStreamCreateWithFlags(&h->st_cap, cudaStreamNonBlocking));
    CK(cudaStreamBeginCapture(h->st_cap, cudaStreamCaptureModeThreadLocal));
    const long long launched = g_launches.load();
    const int rc = mg_vcycle(h, l, b, x, deg, h->st_cap);
    g_launches.store(launched);                       // captured, not launched
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->st_cap, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(h, JSSO_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    const cudaError_t e2 = cudaGraphInstantiate(&h->mg_graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) { h->mg_graph_exec = nullptr; return fail(h, JSSO_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2)); }
    h->mg_graph_level = l; h->mg_graph_b = b; h->mg_graph_x = x; h->mg_graph_deg = deg;
  }
  CK(cudaGraphLaunch(h->mg_graph_exec, st));
  LAUNCHED();
  return JSSO_OK;
}

// ---------------------------------------------------------------- multigrid, row-range distributed solve
// The hierarchy is replicated (every rank assembled the whole renumbered mesh and ran mg_numeric_setup); rank r
// computes rows [bounds[l][r], bounds[l][r+1]) of every product at the distributed levels l < n_dist with the
// same kernels (offset pointers: `rowptr + s`, `y + 6 s`; x stays full length in global numbering).  Before a
// product the entries of x that its rows read and other ranks own are delivered by mgd_exchange (pack ->
// grouped ncclSend/ncclRecv -> unpack at the global positions).  Levels >= n_dist run replicated on the
// all-gathered right-hand side.  The same sequence of range products and exchanges is replayed on the CPU with
// NaN-poisoned ghosts by oracle/multigrid_ref.py::emulate_distributed_pcg (tests/test_dist_multigrid.py).
static inline void mgd_range(const jsso_handle* h, int l, int& s, int& n) {
  const std::vector<int32_t>& b = h->mgd.bounds[l];
  s = b[h->mgd.rank];
  n = b[h->mgd.rank + 1] - s;
}

static int mgd_exchange(jsso_handle* h, int l, double* v, cudaStream_t st) {
  jsso_handle::MgDistLevel& d = h->mgd.lv[l];
  if (d.peers.empty()) return JSSO_OK;
  if (h->mgd.ctx) {   // peer-memory path: two kernels, no library call
    const unsigned long long seq = ++d.seq;
    const long long nmax = 6LL * std::max(d.n_send, d.n_recv);
    if (nmax <= MGD_ONE_BLOCK_MAX) {
      const int threads = (int)std::max(32LL, std::min(1024LL, ((nmax + 31) / 32) * 32));
      mgd_exchange_kernel<<<1, threads, 0, st>>>(h->mgd.ctx, d.dev, l, d.send_idx, d.recv_idx, v, h->mgd.push_counter, seq);
      CKL("mgd_exchange_kernel");
    } else {
      const int pb = std::max(1, std::min(32, cdiv(6LL * d.n_send, RED_BLOCK)));
      mgd_push_kernel<<<pb, RED_BLOCK, 0, st>>>(h->mgd.ctx, d.dev, l, d.send_idx, v, h->mgd.push_counter, seq);
      CKL("mgd_push_kernel");
      const int ub = std::max(1, std::min(32, cdiv(6LL * d.n_recv, RED_BLOCK)));
      mgd_wait_unpack_kernel<<<ub, RED_BLOCK, 0, st>>>(h->mgd.ctx, d.dev, l, d.recv_idx, v, seq);
      CKL("mgd_wait_unpack_kernel");
    }
    ++h->mgd.n_exchange;
    return JSSO_OK;
  }
  if (d.n_send > 0) {
    halo_pack_kernel<<<cdiv(6LL * d.n_send, 256), 256, 0, st>>>(d.n_send, d.send_idx, v, h->mgd.send_buf);
    CKL("halo_pack_kernel");
  }
  CKN(g_nccl.GroupStart());
  for (const jsso_handle::MgDistPeer& p : d.peers) {
    if (p.send_cnt)
      CKN(g_nccl.Send(h->mgd.send_buf + 6 * (size_t)p.send_off, 6 * (size_t)p.send_cnt, ncclDouble, p.rank, h->mgd.comm, st));
    if (p.recv_cnt)
      CKN(g_nccl.Recv(h->mgd.recv_buf + 6 * (size_t)p.recv_off, 6 * (size_t)p.recv_cnt, ncclDouble, p.rank, h->mgd.comm, st));
  }
  CKN(g_nccl.GroupEnd());
  if (d.n_recv > 0) {
    halo_unpack_kernel<<<cdiv(6LL * d.n_recv, 256), 256, 0, st>>>(d.n_recv, d.recv_idx, h->mgd.recv_buf, v);
    CKL("halo_unpack_kernel");
  }
  ++h->mgd.n_exchange;
  return JSSO_OK;
}

// every rank's row range of a level-l vector to every other rank (ranges are contiguous: no packing)
static int mgd_allgather(jsso_handle* h, int l, double* v, cudaStream_t st) {
  // the first replicated level has its own exchange plan (everybody's whole range): over peer memory it is a push +
  // wait/unpack kernel pair like every halo exchange, no NCCL call on the iteration path
  if (h->mgd.ctx && l == h->mgd.n_dist && (int)h->mgd.lv.size() > l) return mgd_exchange(h, l, v, st);
  const std::vector<int32_t>& b = h->mgd.bounds[l];
  const int me = h->mgd.rank;
  const size_t mine = 6 * (size_t)(b[me + 1] - b[me]);
  CKN(g_nccl.GroupStart());
  for (int r = 0; r < h->mgd.n_rank; ++r) {
    if (r == me) continue;
    const size_t theirs = 6 * (size_t)(b[r + 1] - b[r]);
    if (mine) CKN(g_nccl.Send(v + 6 * (size_t)b[me], mine, ncclDouble, r, h->mgd.comm, st));
    if (theirs) CKN(g_nccl.Recv(v + 6 * (size_t)b[r], theirs, ncclDouble, r, h->mgd.comm, st));
  }
  CKN(g_nccl.GroupEnd());
  return JSSO_OK;
}

// sum over the ranks of `count` device scalars starting at slot (in place), then read all scalars back
static int mgd_reduce(jsso_handle* h, int slot, int count, cudaStream_t st) {
  if (h->mgd.ctx) {
    if (count > 2) return fail(h, JSSO_ERR_ARG, "peer-memory all-reduce: at most 2 scalars");
    mgd_allreduce_kernel<<<1, 32, 0, st>>>(h->mgd.ctx, h->mg_scal + slot, h->mg_scal + slot, count, h->mgd.red_seq_dev);
    CKL("mgd_allreduce_kernel");
  } else {
    CKN(g_nccl.AllReduce(h->mg_scal + slot, h->mg_scal + slot, (size_t)count, ncclDouble, ncclSum, h->mgd.comm, st));
  }
  ++h->mgd.n_allreduce;
  return JSSO_OK;
}
static int mgd_reduce_read(jsso_handle* h, int slot, int count, cudaStream_t st) {
  const int rc = mgd_reduce(h, slot, count, st);
  return rc ? rc : mg_read_scalars(h, st);
}

static int mg_smooth_dist(jsso_handle* h, int l, const double* b, double* x, bool zero_guess, int deg, cudaStream_t st) {
  jsso_handle::MgLevel& m = h->mg[l];
  const MgMat A = mg_matrix(h, l);
  int s, n;
  mgd_range(h, l, s, n);
  const size_t off = 6 * (size_t)s;
  const int nb = cdiv(n, 128);
  const double* Dinv = m.Dinv ? m.Dinv + 36 * (size_t)s : nullptr;
  const double lmax = m.lam, lmin = m.lam / 4.0;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho = 1.0 / sigma;
  int rc;
  if (zero_guess) {
    if (n > 0) {
      mg_cheb_kernel<1><<<nb, 128, 0, st>>>(n, Dinv, b + off, m.d + off, x + off, 0.0, 1.0 / theta, 1);
      CKL("mg_cheb_kernel<1>");
    }
  } else {
    if ((rc = mgd_exchange(h, l, x, st))) return rc;
    if ((rc = mg_spmv_p<2>(h, A.rp + s, A.ci, A.v, A.v32, n, x, m.r + off, b + off, st, false, A.v16, (double)A.nnz / std::max(A.n, 1)))) return rc;
    if (n > 0) {
      mg_cheb_kernel<1><<<nb, 128, 0, st>>>(n, Dinv, m.r + off, m.d + off, x + off, 0.0, 1.0 / theta, 0);
      CKL("mg_cheb_kernel<1>");
    }
  }
  for (int k = 1; k < deg; ++k) {
    if ((rc = mgd_exchange(h, l, x, st))) return rc;
    if ((rc = mg_spmv_p<2>(h, A.rp + s, A.ci, A.v, A.v32, n, x, m.r + off, b + off, st, false, A.v16, (double)A.nnz / std::max(A.n, 1)))) return rc;
    const double rho_new = 1.0 / (2.0 * sigma - rho);
    if (n > 0) {
      mg_cheb_kernel<0><<<nb, 128, 0, st>>>(n, Dinv, m.r + off, m.d + off, x + off, rho_new * rho, 2.0 * rho_new / delta, 0);
      CKL("mg_cheb_kernel<0>");
    }
    rho = rho_new;
  }
  return JSSO_OK;
}

// b, x: full-length level-l vectors; on entry b is valid on this rank's range (everywhere at replicated levels),
// on exit x is valid on this rank's range (everywhere at replicated levels)
static int mg_vcycle_dist(jsso_handle* h, int l, double* b, double* x, int deg, cudaStream_t st) {
  const int nl = (int)h->mg.size();
  if (l >= h->mgd.n_dist) return mg_vcycle_graphed(h, l, b, x, deg, st);
  int rc;
  jsso_handle::MgLevel& m = h->mg[l];
  const MgMat A = mg_matrix(h, l);
  int s, n, s1, n1;
  mgd_range(h, l, s, n);
  mgd_range(h, l + 1, s1, n1);
  const size_t off = 6 * (size_t)s, off1 = 6 * (size_t)s1;
  double* bc = (l + 1 < nl) ? h->mg[l + 1].b : h->mg_cb;
  double* xc = (l + 1 < nl) ? h->mg[l + 1].x : h->mg_cx;
  if ((rc = mg_smooth_dist(h, l, b, x, true, deg, st))) return rc;
  if ((rc = mgd_exchange(h, l, x, st))) return rc;
  if ((rc = mg_spmv_p<2>(h, A.rp + s, A.ci, A.v, A.v32, n, x, m.r + off, b + off, st, false, A.v16, (double)A.nnz / std::max(A.n, 1)))) return rc;        // r = b - A x
  if ((rc = mgd_exchange(h, l, m.r, st))) return rc;
  if ((rc = mg_spmv_p<0>(h, m.pt_rowptr + s1, m.pt_col, m.Pt, m.Pt32, n1, m.r, bc + off1, nullptr, st, false, nullptr, (double)m.nnz_p / std::max(m.n_c, 1)))) return rc;   // b_c = P^T r
  if (l + 1 == h->mgd.n_dist) { if ((rc = mgd_allgather(h, l + 1, bc, st))) return rc; }
  if ((rc = mg_vcycle_dist(h, l + 1, bc, xc, deg, st))) return rc;
  if (l + 1 < h->mgd.n_dist) { if ((rc = mgd_exchange(h, l + 1, xc, st))) return rc; }
  if ((rc = mg_spmv_p<3>(h, m.p_rowptr + s, m.p_col, m.P, m.P32, n, xc, x + off, nullptr, st,
                         m.nnz_p <= 5LL * m.n_f))) return rc;                                            // x += P x_c
  return mg_smooth_dist(h, l, b, x, false, deg, st);
}

// ---------------------------------------------------------------- multigrid PCG, fused iteration
// One driver for one GPU and for the row-range distributed solve.  Every scalar of the outer PCG lives on the
// device (slots MGS_*): alpha and beta are formed inside the update kernels, the dot products are fused into the
// kernels that produce their operands, and the host only polls the residual every `mg_poll` iterations -- the
// launches it enqueued past convergence are no-ops (mgs_stopped).  Per iteration on the fine level (Chebyshev-1,
// the default): 4 SpMV-shaped kernels of the V-cycle with the smoother folded into their epilogues
// (mg_vcycle_fused), the direction update, q = A p with p.q, and the x / r update with r.r: 7 launches + the coarse
// levels, against 18 launches and 4 host synchronisations of the first (host-driven) version.
// Where a fused dot product goes, and how it is summed over the ranks:
//   one GPU                    -> straight into the slot;
//   peer-memory distributed    -> the kernel's last block all-reduces through the mailboxes and stores the SUM into the
//                                 slot (mgs_store_dot; no extra launch);
//   NCCL distributed           -> the partial goes to MGS_LOC + slot, mgs_reduce sums it out of place into the slot.
static inline double* mgs_dot_target(jsso_handle* h, int slot) {
  return h->mg_scal + ((h->mgd.n_rank > 1 && !h->mgd.ctx) ? MGS_LOC + slot : slot);
}
static inline const MgdCtx* mgs_ctx(jsso_handle* h) { return h->mgd.n_rank > 1 ? h->mgd.ctx : nullptr; }
static int mgs_reduce(jsso_handle* h, int slot, int count, cudaStream_t st) {
  if (h->mgd.n_rank <= 1 || h->mgd.ctx) return JSSO_OK;
  CKN(g_nccl.AllReduce(h->mg_scal + MGS_LOC + slot, h->mg_scal + slot, (size_t)count, ncclDouble, ncclSum,
                       h->mgd.comm, st));
  h->mgd.n_allreduce += count;
  return JSSO_OK;
}
// a rank whose row range is empty still has to take part in the reduction of a fused dot product
static int mgs_zero_dot(jsso_handle* h, int slot, cudaStream_t st) {
  mgs_zero_dot_kernel<<<1, 1, 0, st>>>(h->mg_scal, mgs_dot_target(h, slot), mgs_ctx(h), h->mgd.red_seq_dev);
  CKL("mgs_zero_dot_kernel");
  return JSSO_OK;
}

// y = ca * bvec + cb * xrow + cc * A_l x on the rows [s, s + n) of the level-l V-cycle matrix (fine level: binary16 /
// FP32 / FP64 storage, whichever the handle runs the V-cycle with; coarse levels: FP32)
template <int DOT>
static int mg_lin_level(jsso_handle* h, int l, int s, int n, const double* x, double* y, const double* bvec,
                        const double* xrow, double ca, double cb, double cc, double* dot_out, cudaStream_t st) {
  const MgdCtx* rc_ = (DOT != 0) ? mgs_ctx(h) : nullptr;
  unsigned long long* rs_ = h->mgd.red_seq_dev;
  if (n <= 0) {
    if (DOT != 0) {
      mgs_zero_dot_kernel<<<1, 1, 0, st>>>(h->mg_scal, dot_out, rc_, rs_);
      CKL("mgs_zero_dot_kernel");
    }
    return JSSO_OK;
  }
  const MgMat A = mg_matrix(h, l);
  const int g = mg_blocks(h, n);
  const double* stop = h->mg_scal;
  // thread per (row, row pair) for short rows on levels large enough to fill the GPU, else a warp per row
  const bool rp = JSSO_MG_RP && ((double)A.nnz <= 12.0 * std::max(A.n, 1)) && (l == 0 || n >= 4096);
  if (h->mg_fp32 && A.v16 && rp) {
    bsr_spmv_rp_kernel<__half, DOT><<<rp_grid(h, n), RED_BLOCK, 0, st>>>(n, A.rp + s, A.ci, A.v16, x, y, bvec, xrow, ca, cb, cc,
                                                                        stop, h->partials, h->counters + 2, dot_out, rc_, rs_);
    CKL("bsr_spmv_rp_kernel<half>");
  } else if (h->mg_fp32 && A.v32 && rp) {
    bsr_spmv_rp_kernel<float, DOT><<<rp_grid(h, n), RED_BLOCK, 0, st>>>(n, A.rp + s, A.ci, A.v32, x, y, bvec, xrow, ca, cb, cc,
                                                                       stop, h->partials, h->counters + 2, dot_out, rc_, rs_);
    CKL("bsr_spmv_rp_kernel<float>");
  } else if (h->mg_fp32 && A.v16) {
    bsr_spmv_lin_kernel<__half, DOT><<<g, RED_BLOCK, 0, st>>>(n, A.rp + s, A.ci, A.v16, x, y, bvec, xrow, ca, cb, cc, stop,
                                                             h->partials, h->counters + 2, dot_out, rc_, rs_);
    CKL("bsr_spmv_lin_kernel<half>");
  } else if (h->mg_fp32 && A.v32) {
    bsr_spmv_lin_kernel<float, DOT><<<g, RED_BLOCK, 0, st>>>(n, A.rp + s, A.ci, A.v32, x, y, bvec, xrow, ca, cb, cc, stop,
                                                            h->partials, h->counters + 2, dot_out, rc_, rs_);
    CKL("bsr_spmv_lin_kernel<float>");
  } else {
    bsr_spmv_lin_kernel<double, DOT><<<g, RED_BLOCK, 0, st>>>(n, A.rp + s, A.ci, A.v, x, y, bvec, xrow, ca, cb, cc, stop,
                                                             h->partials, h->counters + 2, dot_out, rc_, rs_);
    CKL("bsr_spmv_lin_kernel<double>");
  }
  return JSSO_OK;
}

// x = M_l^-1 b: one V-cycle from level l with Chebyshev-1 smoothing folded into the products.  Every level matrix is
// block-Jacobi scaled (unit diagonal blocks, D^-1 = I), so with theta = 5 lam / 8:
//   pre-smoother from a zero guess x0 = b / theta   =>  r0 = b - A x0 = b - (1/theta) A b      (one SpMV of b)
//   x1 = x0 + P x_c                                  =>  x1 = (1/theta) b + P x_c               (prolongation epilogue)
//   post-smoother x = x1 + (1/theta)(b - A x1)       =>  one SpMV of x1 (+ b.x in its epilogue at the fine level)
// i.e. 4 launches and ~9 vector passes per level instead of 8 launches and ~20.  On several GPUs the levels
// l < n_dist work on this rank's rows, with a halo exchange before every product that gathers a vector other ranks
// have just written; b / x are full-length level vectors.
static int mg_vcycle_fused_level(jsso_handle* h, int l, double* b, double* x, bool want_dot, cudaStream_t st) {
  const int nl = (int)h->mg.size();
  int rc;
  if (l == nl) {
    const int nc = 6 * mg_matrix(h, nl).n;
    mg_dense_matvec_kernel<<<cdiv(32LL * nc, 256), 256, 0, st>>>(nc, h->mg_dense, b, x);
    CKL("mg_dense_matvec_kernel");
    return JSSO_OK;
  }
  jsso_handle::MgLevel& m = h->mg[l];
  const bool dist = h->mgd.n_rank > 1 && l < h->mgd.n_dist;
  int s = 0, n = m.n_f, s1 = 0, n1 = m.n_c;
  if (dist) { mgd_range(h, l, s, n); mgd_range(h, l + 1, s1, n1); }
  const size_t off = 6 * (size_t)s, off1 = 6 * (size_t)s1;
  const double it = 1.0 / (0.625 * m.lam);   // 1 / theta, theta = (lam + lam / 4) / 2
  double* bc = (l + 1 < nl) ? h->mg[l + 1].b : h->mg_cb;
  double* xc = (l + 1 < nl) ? h->mg[l + 1].x : h->mg_cx;
  if (dist) { if ((rc = mgd_exchange(h, l, b, st))) return rc; h->probe.mark("xch_b", l, st); }
  if ((rc = mg_lin_level<0>(h, l, s, n, b, m.r + off, b + off, nullptr, 1.0, 0.0, -it, nullptr, st))) return rc;
  h->probe.mark("K1_resid", l, st);
  if (dist) { if ((rc = mgd_exchange(h, l, m.r, st))) return rc; h->probe.mark("xch_r", l, st); }
  if ((rc = mg_spmv_p<0>(h, m.pt_rowptr + s1, m.pt_col, m.Pt, m.Pt32, n1, m.r, bc + off1, nullptr, st, false, nullptr,
                         (double)m.nnz_p / std::max(m.n_c, 1)))) return rc;
  h->probe.mark("K2_restrict", l, st);
  if (dist && l + 1 == h->mgd.n_dist) { if ((rc = mgd_allgather(h, l + 1, bc, st))) return rc; h->probe.mark("allgather", l + 1, st); }
  // the levels below are the same on every rank (replicated tail of the distributed solve / coarse levels of one GPU)
  const bool one_launch = (h->mgd.n_rank > 1) ? (l + 1 == h->mgd.n_dist) : (l == 0);
  if (one_launch) rc = mg_coarse_graphed(h, l + 1, bc, xc, st);
  else rc = mg_vcycle_fused_level(h, l + 1, bc, xc, false, st);
  if (rc) return rc;
  if (one_launch) h->probe.mark("coarse_tail", l + 1, st);
  if (dist && l + 1 < h->mgd.n_dist) { if ((rc = mgd_exchange(h, l + 1, xc, st))) return rc; h->probe.mark("xch_xc", l + 1, st); }
  // x1 = b / theta + P x_c into m.d
  if (n > 0) {
    if (h->mg_fp32 && m.P32 && JSSO_MG_RP && m.nnz_p <= 6LL * m.n_f) {
      bsr_spmv_rp_kernel<float, 0><<<rp_grid(h, n), RED_BLOCK, 0, st>>>(n, m.p_rowptr + s, m.p_col, m.P32, xc, m.d + off, b + off,
                                                                       nullptr, it, 0.0, 1.0, h->mg_scal, h->partials,
                                                                       h->counters + 2, nullptr);
    } else if (h->mg_fp32 && m.P32) {
      bsr_spmv_lin_kernel<float, 0><<<mg_blocks(h, n), RED_BLOCK, 0, st>>>(n, m.p_rowptr + s, m.p_col, m.P32, xc, m.d + off,
                                                                          b + off, nullptr, it, 0.0, 1.0, h->mg_scal,
                                                                          h->partials, h->counters + 2, nullptr);
    } else {
      bsr_spmv_lin_kernel<double, 0><<<mg_blocks(h, n), RED_BLOCK, 0, st>>>(n, m.p_rowptr + s, m.p_col, m.P, xc, m.d + off,
                                                                           b + off, nullptr, it, 0.0, 1.0, h->mg_scal,
                                                                           h->partials, h->counters + 2, nullptr);
    }
    CKL("prolongation");
  }
  // the ghost entries of x1 that the post-smoother gathers: recomputed from local data (b on the ghost rows arrived
  // with the exchange of b, x_c on their coarse columns with the exchange of x_c / the replicated coarse solve) --
  // one synchronisation point fewer than exchanging them
  bool ghosts_done = false;
  if (dist && h->mg_fp32 && m.P32 && JSSO_MG_RP && h->mgd.lv[l].n_ghost > 0 && h->mgd.lv[l].ghost_rows) {
    const jsso_handle::MgDistLevel& DL = h->mgd.lv[l];
    bsr_spmv_rp_kernel<float, 0><<<rp_grid(h, DL.n_ghost), RED_BLOCK, 0, st>>>(DL.n_ghost, m.p_rowptr, m.p_col, m.P32, xc, m.d, b,
                                                                              nullptr, it, 0.0, 1.0, h->mg_scal, h->partials,
                                                                              h->counters + 2, nullptr, nullptr, nullptr,
                                                                              DL.ghost_rows);
    CKL("prolongation (ghost rows)");
    ghosts_done = true;
  }
  h->probe.mark("K3_prolong", l, st);
  if (dist && !ghosts_done) { if ((rc = mgd_exchange(h, l, m.d, st))) return rc; h->probe.mark("xch_x1", l, st); }
  if (want_dot) rc = mg_lin_level<1>(h, l, s, n, m.d, x + off, b + off, m.d + off, it, 1.0, -it, mgs_dot_target(h, MGS_RZ), st);
  else rc = mg_lin_level<0>(h, l, s, n, m.d, x + off, b + off, m.d + off, it, 1.0, -it, nullptr, st);
  h->probe.mark(want_dot ? "K4_post_rz" : "K4_post", l, st);
  return rc;
}

// The part of the fused V-cycle that is the same on every rank and launch-latency bound -- the coarse levels on one
// GPU (l = 1), the replicated levels of the distributed solve (l = n_dist) -- as ONE graph launch: captured once per
// numeric setup on a private stream (the smoother coefficients are kernel arguments), replayed into the caller's.
// (Tried in round 2 and removed: the same products as ONE cooperative kernel walking a device-resident list with grid
// barriers in between -- ~11 us per phase against ~7 us per kernel node of the graph, 1.41 vs 1.37 ms per iteration at 1M
// quads, 0.94 vs 0.89 ms on two GPUs; profiles/r2t_coarse_tail_cooperative_kernel_ab.txt.  And the smallest levels only
// (<= 3000 rows) as one thread-block CLUSTER with the hardware cluster barrier: 7.4 us per phase, no better than a graph
// node -- these levels cost their chains of dependent L2 loads, not launch gaps; profiles/r2y_smallest_levels_cluster_kernel_ab.txt.)
static int mg_coarse_graphed(jsso_handle* h, int l, double* b, double* x, cudaStream_t st) {
  if (!h->mg_graph) return mg_vcycle_fused_level(h, l, b, x, false, st);
  if (h->mg_graph_exec && (h->mg_graph_level != l || h->mg_graph_b != b || h->mg_graph_x != x || h->mg_graph_deg != -1)) {
    cudaGraphExecDestroy(h->mg_graph_exec);
    h->mg_graph_exec = nullptr;
  }
  if (!h->mg_graph_exec) {
    if (!h->st_cap) CK(cudaStreamCreateWithFlags(&h->st_cap, cudaStreamNonBlocking));
    CK(cudaStreamBeginCapture(h->st_cap, cudaStreamCaptureModeThreadLocal));
    const long long launched = g_launches.load();
    const int rc = mg_vcycle_fused_level(h, l, b, x, false, h->st_cap);
    g_launches.store(launched);                       // captured, not launched
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->st_cap, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(h, JSSO_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    const cudaError_t e2 = cudaGraphInstantiate(&h->mg_graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) { h->mg_graph_exec = nullptr; return fail(h, JSSO_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2)); }
    h->mg_graph_level = l; h->mg_graph_b = b; h->mg_graph_x = x; h->mg_graph_deg = -1;   // -1: the fused form
  }
  CK(cudaGraphLaunch(h->mg_graph_exec, st));
  LAUNCHED();
  return JSSO_OK;
}

// z = M^-1 b (one V-cycle from the fine level) on this rank's rows and r.z = b.z into slot MGS_RZ.  Chebyshev degree 1
// (the default) takes the fused form above; other degrees the generic V-cycle.
static int mg_vcycle_fused(jsso_handle* h, double* b, double* z, int deg, cudaStream_t st) {
  const bool dist = h->mgd.n_rank > 1;
  const int nl = (int)h->mg.size();
  if (deg == 1 && nl > 0) return mg_vcycle_fused_level(h, 0, b, z, true, st);
  int s = 0, n = h->sym.n_row;
  if (dist) mgd_range(h, 0, s, n);
  const size_t off = 6 * (size_t)s;
  const int rc = dist ? mg_vcycle_dist(h, 0, b, z, deg, st) : mg_vcycle_graphed(h, 0, b, z, deg, st);
  if (rc) return rc;
  mg_dot_kernel<<<std::max(1, std::min(h->red_blocks, cdiv(6LL * n, 256))), 256, 0, st>>>(
      6LL * n, b + off, z + off, h->partials, h->counters + 2, mgs_dot_target(h, MGS_RZ), mgs_ctx(h), h->mgd.red_seq_dev);
  CKL("mg_dot_kernel");
  return JSSO_OK;
}

static int mg_solve_fused(jsso_handle* h, const jsso_solve_opts& o, bool use_x0, jsso_stats* stats, cudaStream_t st) {
  int rc = mg_numeric_setup(h, st);
  if (rc) return rc;
  const bool dist = h->mgd.n_rank > 1;
  int s = 0, n_row = h->sym.n_row;
  if (dist) mgd_range(h, 0, s, n_row);
  const size_t off = 6 * (size_t)s;
  const long long n = 6LL * n_row;
  const int vb = std::max(1, std::min(h->red_blocks, cdiv(n, 256)));
  double *b = h->vb, *x = h->vx, *r = h->vr, *p = h->vp, *q = h->vq, *z = h->tmp_g;
  double* scal = h->mg_scal;
  const MgMat A = mg_matrix(h, 0);
  if (use_x0) {
    if (dist) { if ((rc = mgd_exchange(h, 0, x, st))) return rc; }   // the rows of the other ranks were scaled there
    if ((rc = mg_spmv<2>(h, A.rp + s, A.ci, A.v, n_row, x, r + off, b + off, st))) return rc;
  } else {
    CK(cudaMemsetAsync(x + off, 0, n * sizeof(double), st));
    CK(cudaMemcpyAsync(r + off, b + off, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  CK(cudaMemsetAsync(scal, 0, MGS_COUNT * sizeof(double), st));
  auto dot = [&](const double* a_, const double* b_, int slot) -> int {
    mg_dot_kernel<<<vb, 256, 0, st>>>(n, a_, b_, h->partials, h->counters + 2, mgs_dot_target(h, slot), mgs_ctx(h),
                                      h->mgd.red_seq_dev);
    CKL("mg_dot_kernel");
    return JSSO_OK;
  };
  if ((rc = dot(b + off, b + off, MGS_BB))) return rc;
  if ((rc = dot(r + off, r + off, MGS_RR))) return rc;
  if ((rc = mgs_reduce(h, MGS_BB, 2, st))) return rc;
  mg_pcg_begin_kernel<<<1, 1, 0, st>>>(scal, o.rtol * o.rtol, 1);
  CKL("mg_pcg_begin_kernel");
  if ((rc = mg_read_scalars(h, st))) return rc;
  const double bb = h->mg_scal_host[MGS_BB];
  double rr = h->mg_scal_host[MGS_RR];
  int it = 0, restarts = 0;
  bool converged = (bb == 0.0) || std::sqrt(rr / bb) <= o.rtol;
  bool first = true, stagnated = false;
  double prev_true = 1e300;
  while (!converged && it < o.maxiter) {
    const int batch = std::max(1, std::min(h->mg_poll, o.maxiter - it));
    for (int k = 0; k < batch; ++k) {
      if (h->probe.armed && !h->probe.done && it + k == 3) { h->probe.active = true; h->probe.mark("start", -1, st); }
      if ((rc = mg_vcycle_fused(h, r, z, o.cheb_degree, st))) return rc;
      if ((rc = mgs_reduce(h, MGS_RZ, 1, st))) return rc;
      mg_pcg_dir_kernel<<<vb, 256, 0, st>>>(n, z + off, p + off, scal, first ? 1 : 0);
      CKL("mg_pcg_dir_kernel");
      first = false;
      h->probe.mark("dir", -1, st);
      if (dist) { if ((rc = mgd_exchange(h, 0, p, st))) return rc; h->probe.mark("xch_p", 0, st); }
      if (n_row > 0) {
        bsr_spmv_dot_kernel<<<mg_blocks(h, n_row), RED_BLOCK, 0, st>>>(n_row, A.rp + s, A.ci, A.v, p, p + off, q + off, scal,
                                                                     h->partials, h->counters + 2, mgs_dot_target(h, MGS_PQ),
                                                                     mgs_ctx(h), h->mgd.red_seq_dev);
        CKL("bsr_spmv_dot_kernel");
      } else {
        if ((rc = mgs_zero_dot(h, MGS_PQ, st))) return rc;
      }
      if ((rc = mgs_reduce(h, MGS_PQ, 1, st))) return rc;
      h->probe.mark("K6_Ap_pq", -1, st);
      mg_pcg_update_kernel<<<vb, 256, 0, st>>>(n, p + off, q + off, x + off, r + off, scal, h->partials, h->counters + 2,
                                              mgs_dot_target(h, MGS_RR), mgs_ctx(h), h->mgd.red_seq_dev);
      CKL("mg_pcg_update_kernel");
      if ((rc = mgs_reduce(h, MGS_RR, 1, st))) return rc;
      if (h->probe.active) { h->probe.mark("K7_update_rr", -1, st); h->probe.active = false; h->probe.done = true; }
    }
    if ((rc = mg_read_scalars(h, st))) return rc;
    if (h->probe.done && !h->probe.ev.empty()) h->probe.report(h->mgd.rank);
    const int it_new = (int)h->mg_scal_host[MGS_ITER];
    rr = h->mg_scal_host[MGS_RR];
    if (!(rr == rr)) {
      char buf[240];
      std::snprintf(buf, sizeof buf, "multigrid PCG breakdown by iteration %d: r.z = %.3e, p.Ap = %.3e (lam0 = %.3f)", it_new,
                    h->mg_scal_host[MGS_RZ], h->mg_scal_host[MGS_PQ], h->mg.empty() ? 0.0 : h->mg[0].lam);
      return fail(h, JSSO_ERR_NAN, buf);
    }
    it = it_new;
    if (std::sqrt(rr / bb) <= o.rtol) {
      // confirm on the true residual; if the recurrence drifted keep iterating from it
      if (dist) { if ((rc = mgd_exchange(h, 0, x, st))) return rc; }
      if ((rc = mg_spmv<2>(h, A.rp + s, A.ci, A.v, n_row, x, r + off, b + off, st))) return rc;
      if ((rc = dot(r + off, r + off, MGS_RR))) return rc;
      if ((rc = mgs_reduce(h, MGS_RR, 1, st))) return rc;
      if ((rc = mg_read_scalars(h, st))) return rc;
      rr = h->mg_scal_host[MGS_RR];
      const double true_relres = std::sqrt(rr / bb);
      if (true_relres <= 1.5 * o.rtol) { converged = true; break; }
      // residual replacement: restart the recurrence from the true residual -- unless the last restart no longer
      // gained a factor 2 (attainable accuracy of FP64 for this matrix: 1.5e-9 at 1M quads)
      if (true_relres > 0.5 * prev_true || ++restarts > 20) { stagnated = true; break; }
      prev_true = true_relres;
      first = true;
    }
  }
  if (dist && !h->mgd.setup_on) { if ((rc = mgd_allgather(h, 0, x, st))) return rc; }   // else: after the unscaling
  if (stats) {
    stats->iterations = it; stats->restarts = restarts; stats->converged = converged ? 1 : 0;
    stats->relres = bb > 0 ? std::sqrt(rr / bb) : 0.0; stats->relres_recur = stats->relres;
  }
  if (!converged) {
    char buf[200];
    std::snprintf(buf, sizeof buf, "multigrid PCG did not reach rtol=%.3g: true relres %.3g after %d iterations, %d restarts%s",
                  o.rtol, bb > 0 ? std::sqrt(rr / bb) : 0.0, it, restarts,
                  stagnated ? " (stagnated: attainable accuracy)" : (it >= o.maxiter ? " (maxiter)" : ""));
    return fail(h, JSSO_ERR_NOCONV, buf);
  }
  return JSSO_OK;
}

// ---------------------------------------------------------------- FP64 peak (bench.py roofline denominator)
// DFMA-chain microbenchmark: every thread runs 8 independent fused multiply-add chains (enough to cover the
// FP64 pipe latency at 8 warps per scheduler), persistent grid of 4 x 256 threads per SM; the result is
// stored so that the compiler keeps the chains.
__global__ void __launch_bounds__(256)
fp64_peak_kernel(double* out, int iters, double b, double c) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
      a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int jsso_fp64_peak(int32_t device, double seconds, double* tflops_out, double* seconds_out) {
  jsso_handle* h = nullptr;
  if (!tflops_out || !(seconds > 0.0)) return fail(h, JSSO_ERR_ARG, "jsso_fp64_peak: bad argument");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 4, iters = 4096;
  double* out = nullptr;
  CK(cudaMalloc((void**)&out, (size_t)blocks * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double flop_per_launch = 2.0 * 64.0 * iters * 256.0 * blocks;
  for (int w = 0; w < 3; ++w) { fp64_peak_kernel<<<blocks, 256>>>(out, iters, 0.999999, 1e-6); LAUNCHED(); }
  CK(cudaDeviceSynchronize());
  // calibrate the launch count for the requested duration, then time that many back-to-back launches
  CK(cudaEventRecord(e0));
  fp64_peak_kernel<<<blocks, 256>>>(out, iters, 0.999999, 1e-6); LAUNCHED();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms1 = 0.f;
  CK(cudaEventElapsedTime(&ms1, e0, e1));
  const int reps = std::max(1, std::min(100000, (int)(seconds * 1e3 / std::max(ms1, 1e-3f))));
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) { fp64_peak_kernel<<<blocks, 256>>>(out, iters, 0.999999, 1e-6); LAUNCHED(); }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops_out = flop_per_launch * reps / (ms * 1e-3) / 1e12;
  if (seconds_out) *seconds_out = ms * 1e-3;
  return JSSO_OK;
}

extern "C" {

// K x = b on the assembled BC-imposed matrix: scale, solve, unscale.
static int solve_system(jsso_handle* h, const double* b, double* x, const jsso_solve_opts& o, jsso_stats* stats,
                        cudaStream_t st) {
  if (!h->assembled || !h->assembled_bc)
    return fail(h, JSSO_ERR_STATE, "solve needs a matrix assembled with apply_bc=1");
  PhaseTimer pts_(st, h->mgd.rank);
  int rc = ensure_scaled(h, st);
  if (rc) return rc;
  pts_.mark("block_jacobi_scaling");
  CK(cudaMemcpyAsync(h->flags_host, h->flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int fl = *h->flags_host;
  if (stats) { std::memset(stats, 0, sizeof(*stats)); stats->flags = fl; }
  if (fl & FLAG_DEGBEAM) return fail(h, JSSO_ERR_DEGENERATE_BEAM, "beam parallel to global Y: K_e is unsymmetric in the reference (element.py:92-94); not solvable by CG");
  if (fl & FLAG_UNSYM) return fail(h, JSSO_ERR_NOT_SPD, "kx_mod != ky_mod makes K unsymmetric (element.py:871-873); PCG path supports kx_mod == ky_mod only");
  if (fl & FLAG_BADJAC) return fail(h, JSSO_ERR_BADJAC, "non-positive Jacobian determinant in a quad");
  if (fl & 8) return fail(h, JSSO_ERR_NOT_SPD, "a diagonal 6x6 block is not positive definite");
  const int n_row = h->sym.n_row;
  if (n_row == 0) return JSSO_OK;
  // b^ = W b (prescribed dofs zeroed); an initial guess x0 maps to y0 = W^-T x0 (x = W^T y)
  int rs = 0, rn = n_row;       // rows this rank scales / unscales: all, or its own range (distributed setup)
  if (h->mgd.setup_on) mgd_range(h, 0, rs, rn);
  const size_t ro = 6 * (size_t)rs;
  if (rn > 0) {
    block_apply_kernel<0><<<cdiv(rn, 128), 128, 0, st>>>(rn, h->W + 6 * ro, b + ro, h->node_mask + rs, h->vb + ro);
    CKL("block_apply_kernel<0>");
    if (o.use_x0) {
      block_solve_wt_kernel<<<cdiv(rn, 128), 128, 0, st>>>(rn, h->W + 6 * ro, x + ro, h->node_mask + rs, h->vx + ro);
      CKL("block_solve_wt_kernel");
    }
  }
  // preconditioner: 0 auto (multigrid when a hierarchy is set, the system is not tiny and the
  // handle is single-GPU), 1 block-Jacobi CG, 2 smoothed-aggregation multigrid
  const bool have_mg = !h->mg.empty() && h->n_rank <= 1;
  const bool use_mg = (o.precond == 2) || (o.precond == 0 && have_mg && h->sym.n_row >= 20000);
  if (use_mg && !have_mg) return fail(h, JSSO_ERR_STATE, "precond = multigrid but no hierarchy (jsso_mg_setup)");
  if (h->mgd.setup_on && !use_mg)
    return fail(h, JSSO_ERR_STATE, "a handle with the distributed numeric setup holds only this rank's rows: multigrid solve only");
  if (stats) stats->flags = fl;
  if (use_mg) rc = mg_solve_fused(h, o, o.use_x0 != 0, stats, st);
  else rc = cg_solve_scaled(h, o, o.use_x0 != 0, stats, st);
  if (stats) stats->flags = fl;
  pts_.mark("setup_and_pcg");
  if (rc && rc != JSSO_ERR_NOCONV) return rc;
  if (rn > 0) {
    block_apply_kernel<1><<<cdiv(rn, 128), 128, 0, st>>>(rn, h->W + 6 * ro, h->vx + ro, nullptr, x + ro);
    CKL("block_apply_kernel<1>");
  }
  if (h->mgd.setup_on) { int r2 = mgd_allgather(h, 0, x, st); if (r2) return r2; }   // every rank unscaled its own rows
  if (h->n_rank > 1) { int r2 = halo_exchange_w(h, x, 6, st); if (r2) return r2; }
  CK(cudaStreamSynchronize(st));
  return rc;
}

int jsso_pcg(jsso_handle* h, const double* b, double* x, const jsso_solve_opts* opts, jsso_stats* stats,
             void* stream) {
  if (!h || !b || !x) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  jsso_solve_opts o; default_opts(opts, o);
  return solve_system(h, b, x, o, stats, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- adjoint
// quad adjoint of the quad range [q0, q0 + nq): per-corner partials into corner_q (want_corner), d_prop_q rows of the range
static int adjoint_quad_range(jsso_handle* h, int q0, int nq, const double* crds, const double* prop_q, const double* u,
                              const double* lam, bool want_corner, double* d_prop_q, cudaStream_t st) {
  if (nq <= 0) return JSSO_OK;
  const int blocks = std::min(cdiv(nq, ADJ_QUADS), d_prop_q ? h->adj_ctas_prop : h->adj_ctas);
  const int32_t* cq = h->cnct_q + 4 * (size_t)q0;
  const double* pq = prop_q + 5 * (size_t)q0;
  double* corner = want_corner ? h->corner_q + 12 * (size_t)q0 : nullptr;
  if (d_prop_q)
    quad_adjoint_kernel<true><<<blocks, 4 * ADJ_QUADS, ADJ_SMEM_DOUBLES * sizeof(double), st>>>(nq, crds, cq, pq, u, lam, corner,
                                                               d_prop_q + 5 * (size_t)q0);
  else
    quad_adjoint_kernel<false><<<blocks, 4 * ADJ_QUADS, ADJ_SMEM_DOUBLES * sizeof(double), st>>>(nq, crds, cq, pq, u, lam, corner, nullptr);
  CKL("quad_adjoint_kernel");
  return JSSO_OK;
}

int jsso_adjoint(jsso_handle* h, const double* crds, const double* prop_q, const double* prop_b, const double* u,
                 const double* lam, double* d_crds, double* d_prop_q, double* d_prop_b, void* stream) {
  if (!h || !u || !lam) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const Symbolic& S = h->sym;
  if (S.n_quad > 0) {
    int rc = adjoint_quad_range(h, 0, S.n_quad, crds, prop_q, u, lam, d_crds != nullptr, d_prop_q, st);
    if (rc) return rc;
  }
  if (S.n_beam > 0) {
    beam_adjoint_kernel<<<cdiv(2LL * S.n_beam, 128), 128, 0, st>>>(S.n_beam, crds, h->cnct_b, prop_b, u, lam,
                                                                  d_crds ? h->corner_b : nullptr, d_prop_b,
                                                                  h->flags);
    CKL("beam_adjoint_kernel");
  }
  if (d_crds) {
    node_gather_kernel<<<cdiv(3LL * S.n_node, 256), 256, 0, st>>>(S.n_node, S.n_quad, h->node_inc_ptr,
                                                                 h->node_inc, h->corner_q, h->corner_b, d_crds);
    CKL("node_gather_kernel");
  }
  return JSSO_OK;
}

// ---------------------------------------------------------------- forward / backward
int jsso_forward(jsso_handle* h, const double* crds, const double* prop_q, const double* prop_b, const double* f,
                 double* u, const jsso_solve_opts* opts, jsso_stats* stats, void* stream) {
  if (!h || !crds || !f || !u) return JSSO_ERR_ARG;
  int rc = jsso_assemble(h, crds, prop_q, prop_b, 1, stream);
  if (rc) return rc;
  jsso_solve_opts o; default_opts(opts, o);
  return solve_system(h, f, u, o, stats, (cudaStream_t)stream);
}

int jsso_backward(jsso_handle* h, const double* crds, const double* prop_q, const double* prop_b, const double* u,
                  const double* g, double* d_crds, double* d_prop_q, double* d_prop_b, double* lam,
                  const jsso_solve_opts* opts, jsso_stats* stats, void* stream) {
  if (!h || !crds || !u || (!g && !(opts && opts->compliance))) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  jsso_solve_opts o; default_opts(opts, o);
  double* l = lam ? lam : h->tmp_lam;
  const long long n = 6LL * h->sym.n_node;
  if (o.compliance) {
    // g = f/2 and K symmetric  =>  lam = u/2  (SSO_model.py:297-301)
    scale_copy_kernel<<<std::max(1, std::min(h->red_blocks, cdiv(n, 256))), 256, 0, st>>>(n, 0.5, u, l);
    CKL("scale_copy_kernel");
    if (stats) std::memset(stats, 0, sizeof(*stats)), stats->converged = 1;
  } else {
    int rc = solve_system(h, g, l, o, stats, st);   // K symmetric => K^T lam = g is the same operator
    if (rc) return rc;
  }
  return jsso_adjoint(h, crds, prop_q, prop_b, u, l, d_crds, d_prop_q, d_prop_b, stream);
}

// memcpy of a large host block on several threads: a single thread moves ~8 GB/s and pays every page fault of a
// freshly allocated destination itself (117 MB of results at 1M quads: 15 ms into touched memory, 45 ms into fresh);
// JSSO_HOST_THREADS=1 switches it off
static void par_memcpy(void* dst, const void* src, size_t bytes) {
  static const unsigned max_threads = [] {
    unsigned n = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = std::getenv("JSSO_HOST_THREADS")) n = (unsigned)std::max(1, std::min(64, std::atoi(e)));
    return n;
  }();
#ifdef JSSO_EMU   // CPU test harness: tiny meshes still take the threaded path
  const size_t min_chunk = 256;
#else
  const size_t min_chunk = (size_t)4 << 20;
#endif
  const unsigned nt = (unsigned)std::min<size_t>(max_threads, bytes / min_chunk);
  if (nt <= 1) { std::memcpy(dst, src, bytes); return; }
  const size_t chunk = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  th.reserve(nt);
  size_t done = 0;                                     // bytes handed to a thread so far
  try {
    for (unsigned t = 0; t < nt && done < bytes; ++t) {
      const size_t o = done, n = std::min(chunk, bytes - o);
      th.emplace_back([=] { std::memcpy((char*)dst + o, (const char*)src + o, n); });
      done += n;
    }
  } catch (...) {
    // no more threads to be had (resource limits): the rest on this one -- never let an exception cross the C ABI
  }
  if (done < bytes) std::memcpy((char*)dst + done, (const char*)src + done, bytes - done);
  for (std::thread& t : th) t.join();
}
// page-locked (cudaHostAlloc / cudaHostRegister) host memory can be the source / target of an asynchronous DMA
static bool host_pinned(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

static int ensure_host_staging(jsso_handle* h) {
  if (h->h_crds) return JSSO_OK;
  const Symbolic& S = h->sym;
  CK(cudaMallocHost((void**)&h->h_crds, sizeof(double) * 3 * (size_t)S.n_node));
  CK(cudaMallocHost((void**)&h->h_pq, sizeof(double) * (5 * (size_t)S.n_quad + 1)));
  CK(cudaMallocHost((void**)&h->h_pb, sizeof(double) * (6 * (size_t)S.n_beam + 1)));
  CK(cudaMallocHost((void**)&h->h_f, sizeof(double) * 6 * (size_t)S.n_node));
  CK(cudaMallocHost((void**)&h->h_u, sizeof(double) * 6 * (size_t)S.n_node));
  CK(cudaMallocHost((void**)&h->h_dc, sizeof(double) * 3 * (size_t)S.n_node));
  CK(cudaMallocHost((void**)&h->h_dpq, sizeof(double) * (5 * (size_t)S.n_quad + 1)));
  CK(cudaMallocHost((void**)&h->h_dpb, sizeof(double) * (6 * (size_t)S.n_beam + 1)));
  CK(cudaMallocHost((void**)&h->val_host, sizeof(double)));
  CK(dalloc(&h->val_dev, 1));
  for (cudaEvent_t& e : h->ev_out) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(dalloc(&h->s_crds, 3 * (size_t)S.n_node)); CK(dalloc(&h->s_pq, 5 * (size_t)S.n_quad));
  CK(dalloc(&h->s_pb, 6 * (size_t)S.n_beam)); CK(dalloc(&h->s_f, 6 * (size_t)S.n_node));
  CK(dalloc(&h->s_u, 6 * (size_t)S.n_node)); CK(dalloc(&h->s_dc, 3 * (size_t)S.n_node));
  CK(dalloc(&h->s_dpq, 5 * (size_t)S.n_quad)); CK(dalloc(&h->s_dpb, 6 * (size_t)S.n_beam));
  CK(cudaStreamCreateWithFlags(&h->st_a, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->st_b, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
  if (const char* e = std::getenv("JSSO_E2E_CHUNKS")) h->e2e_chunks = std::max(1, std::min(64, std::atoi(e)));
  else if (S.n_quad < 400000) h->e2e_chunks = 4;      // smaller systems (the per-rank parts at 4 and 8 GPUs): measured with 4
  if (h->e2e_chunks > 1 && S.n_quad >= h->e2e_chunks && S.n_node >= h->e2e_chunks) {
    const int K = h->e2e_chunks;
    CK(cudaStreamCreateWithFlags(&h->st_c, cudaStreamNonBlocking));
    h->ev_up.assign(K, nullptr); h->ev_adj.assign(K, nullptr);
    for (int c = 0; c < K; ++c) {
      CK(cudaEventCreateWithFlags(&h->ev_up[c], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->ev_adj[c], cudaEventDisableTiming));
    }
    h->e2e_qb.resize(K + 1); h->e2e_nb.resize(K + 1); h->e2e_wait.resize(K);
    for (int c = 0; c <= K; ++c) {
      h->e2e_qb[c] = (int)((long long)S.n_quad * c / K);
      h->e2e_nb[c] = (int)((long long)S.n_node * c / K);
    }
    // a quad range may start once the upload range holding its HIGHEST node has arrived (uploads complete in
    // order); on a mesh whose element order is unrelated to the node order every range waits for the last one
    // and the pipeline degenerates to the unchunked schedule, still correct
    std::vector<int32_t> cq(4 * (size_t)S.n_quad);
    CK(cudaMemcpy(cq.data(), h->cnct_q, cq.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int c = 0; c < K; ++c) {
      int hi = 0;
      for (size_t k = 4 * (size_t)h->e2e_qb[c]; k < 4 * (size_t)h->e2e_qb[c + 1]; ++k) hi = std::max(hi, cq[k]);
      int w = 0;
      while (w + 1 < K && hi >= h->e2e_nb[w + 1]) ++w;
      h->e2e_wait[c] = w;
    }
  } else {
    h->e2e_chunks = 1;
  }
  return JSSO_OK;
}

int jsso_value_and_grad_host(jsso_handle* h, const double* crds_h, const double* pq_h, const double* pb_h,
                             const double* f_h, double* value_out, double* u_h, double* dc_h, double* dpq_h,
                             double* dpb_h, const jsso_solve_opts* opts, jsso_stats* fs, jsso_stats* bs) {
  if (!h || !crds_h || !f_h) return JSSO_ERR_ARG;
  if ((h->sym.n_quad > 0 && !pq_h) || (h->sym.n_beam > 0 && !pb_h)) return fail(h, JSSO_ERR_ARG, "null property array");
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  int rc = ensure_host_staging(h);
  if (rc) return rc;
  const Symbolic& S = h->sym;
  const size_t nc = 3 * (size_t)S.n_node, nq = 5 * (size_t)S.n_quad, nb = 6 * (size_t)S.n_beam,
               nd = 6 * (size_t)S.n_node;
  double *d_crds = h->s_crds, *d_pq = h->s_pq, *d_pb = h->s_pb, *d_f = h->s_f, *d_u = h->s_u, *d_dc = h->s_dc,
         *d_dpq = h->s_dpq, *d_dpb = h->s_dpb;
  cudaStream_t st = 0;
  // pinned caller buffers are DMA sources / targets themselves; pageable ones go through the handle's pinned staging
  // (threaded memcpy; the DMA of one array overlaps the staging of the next)
  auto h2d = [&](double* dst, const double* src, double* stage, size_t n) -> cudaError_t {
    if (!n) return cudaSuccess;
    if (!host_pinned(src)) { par_memcpy(stage, src, n * sizeof(double)); src = stage; }
    return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, st);
  };
  CK(h2d(d_crds, crds_h, h->h_crds, nc));
  CK(h2d(d_pq, pq_h, h->h_pq, nq));
  CK(h2d(d_pb, pb_h, h->h_pb, nb));
  CK(h2d(d_f, f_h, h->h_f, nd));
  jsso_solve_opts o; default_opts(opts, o);
  if (o.use_x0 && u_h) {   // warm start: u_h holds the previous design's displacements
    CK(h2d(d_u, u_h, h->h_u, nd));
  } else {
    o.use_x0 = 0;
  }
  rc = jsso_forward(h, d_crds, d_pq, d_pb, d_f, d_u, &o, fs, st);
  const int rc_solve = rc;   // JSSO_ERR_NOCONV still delivers the best iterate (attainable accuracy): finish, then report it
  if (rc && rc != JSSO_ERR_NOCONV) return rc;
  const std::string solve_msg = h->err;
  o.compliance = 1; o.use_x0 = 0;
  const bool want_grad = dc_h || dpq_h || dpb_h;
  if (want_grad) {
    rc = jsso_backward(h, d_crds, d_pq, d_pb, d_u, nullptr, dc_h ? d_dc : nullptr, (dpq_h && nq) ? d_dpq : nullptr,
                       (dpb_h && nb) ? d_dpb : nullptr, nullptr, &o, bs, st);
    if (rc) return rc;
  }
  // value = f.u / 2 on the device (deterministic grid sum)
  mg_dot_kernel<<<std::max(1, std::min(h->red_blocks, cdiv((long long)nd, 256))), 256, 0, st>>>((long long)nd, d_f, d_u, h->partials,
                                                                                           h->counters + 2, h->val_dev);
  CKL("mg_dot_kernel");
  CK(cudaMemcpyAsync(h->val_host, h->val_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
  // results: straight into pinned caller buffers, else into the staging and from there while the next array is in flight
  struct Out { double* dst; double* stage; const double* src; size_t n; bool direct; };
  Out outs[4] = {{u_h, h->h_u, d_u, nd, false}, {dc_h, h->h_dc, d_dc, nc, false}, {dpq_h, h->h_dpq, d_dpq, nq, false},
                 {dpb_h, h->h_dpb, d_dpb, nb, false}};
  for (int k = 0; k < 4; ++k) {
    Out& O = outs[k];
    if (!O.dst || !O.n) continue;
    O.direct = host_pinned(O.dst);
    CK(cudaMemcpyAsync(O.direct ? O.dst : O.stage, O.src, O.n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(h->ev_out[k], st));
  }
  for (int k = 0; k < 4; ++k) {
    const Out& O = outs[k];
    if (!O.dst || !O.n || O.direct) continue;
    CK(cudaEventSynchronize(h->ev_out[k]));
    par_memcpy(O.dst, O.stage, O.n * sizeof(double));
  }
  CK(cudaStreamSynchronize(st));
  if (value_out) *value_out = 0.5 * h->val_host[0];
  if (rc_solve) return fail(h, rc_solve, solve_msg);
  return JSSO_OK;
}

// Host-buffer entry point for the solve-free part of a gradient evaluation: H2D of
// coordinates, properties, u and lam; fused Ke + assembly (BC imposed); adjoint
// reduction; D2H of the gradients.  This is the end-to-end leg bench.py times.
int jsso_assemble_adjoint_host(jsso_handle* h, const double* crds_h, const double* pq_h, const double* pb_h,
                               const double* u_h, const double* lam_h, double* dc_h, double* dpq_h,
                               double* dpb_h) {
  if (!h || !crds_h || !u_h || !lam_h) return JSSO_ERR_ARG;
  NEED_GPU();
  CK(cudaSetDevice(h->device));
  int rc = ensure_host_staging(h);
  if (rc) return rc;
  const Symbolic& S = h->sym;
  const size_t nc = 3 * (size_t)S.n_node, nq = 5 * (size_t)S.n_quad, nb = 6 * (size_t)S.n_beam,
               nd = 6 * (size_t)S.n_node;
  // two streams: coordinates/properties + fused assembly on st, the upload of u and lam on st_b
  // overlaps the assembly; the adjoint waits for both
  CK(cudaDeviceSynchronize());
  cudaStream_t st = h->st_a, st2 = h->st_b;
  // DMA straight from / to the caller's buffers when they are pinned (cudaHostAlloc / registered);
  // pageable buffers are staged through the handle's pinned scratch so the copies stay asynchronous
  auto pinned = [](const void* p) { return host_pinned(p); };
  auto h2d = [&](double* dst, const double* src, double* stage, size_t n, cudaStream_t s_) -> cudaError_t {
    if (!n) return cudaSuccess;
    if (!pinned(src)) { par_memcpy(stage, src, n * sizeof(double)); src = stage; }
    return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, s_);
  };
  CK(h2d(h->s_crds, crds_h, h->h_crds, nc, st));
  CK(h2d(h->s_pq, pq_h, h->h_pq, nq, st));
  CK(h2d(h->s_pb, pb_h, h->h_pb, nb, st));
  // (Tried and removed: the assembly cut into task ranges that alternate with the adjoint's quad ranges on the stream, so
  // that gradients leave while later uploads still arrive -- 4.50 vs 4.18 ms at 4 chunks, 4.12 ms at 16: the extra ramps and
  // tails of the persistent kernels cost more than the overlap gains; scripts/r2_call21.sh.)
  if ((rc = jsso_assemble(h, h->s_crds, h->s_pq, h->s_pb, 1, st))) return rc;
  if (h->e2e_chunks > 1) {
    // chunked pipeline (opt-in, JSSO_E2E_CHUNKS): see the handle fields
    const int K = h->e2e_chunks;
    const bool want_dq = dpq_h && nq;
    if (!pinned(u_h)) { par_memcpy(h->h_f, u_h, nd * sizeof(double)); u_h = h->h_f; }
    if (!pinned(lam_h)) { par_memcpy(h->h_u, lam_h, nd * sizeof(double)); lam_h = h->h_u; }
    for (int c = 0; c < K; ++c) {
      const size_t o = 6 * (size_t)h->e2e_nb[c], cnt = 6 * (size_t)(h->e2e_nb[c + 1] - h->e2e_nb[c]);
      CK(cudaMemcpyAsync(h->s_f + o, u_h + o, cnt * sizeof(double), cudaMemcpyHostToDevice, st2));
      CK(cudaMemcpyAsync(h->s_u + o, lam_h + o, cnt * sizeof(double), cudaMemcpyHostToDevice, st2));
      CK(cudaEventRecord(h->ev_up[c], st2));
    }
    const bool p_dq = want_dq && pinned(dpq_h);
    double* dq_dst = want_dq ? (p_dq ? dpq_h : h->h_dpq) : nullptr;
    int waited = -1;
    for (int c = 0; c < K; ++c) {
      if (h->e2e_wait[c] > waited) { waited = h->e2e_wait[c]; CK(cudaStreamWaitEvent(st, h->ev_up[waited], 0)); }
      const int q0 = h->e2e_qb[c], nqc = h->e2e_qb[c + 1] - q0;
      if ((rc = adjoint_quad_range(h, q0, nqc, h->s_crds, h->s_pq, h->s_f, h->s_u, dc_h != nullptr,
                                   want_dq ? h->s_dpq : nullptr, st))) return rc;
      if (want_dq) {
        CK(cudaEventRecord(h->ev_adj[c], st));
        CK(cudaStreamWaitEvent(h->st_c, h->ev_adj[c], 0));
        CK(cudaMemcpyAsync(dq_dst + 5 * (size_t)q0, h->s_dpq + 5 * (size_t)q0, 5 * (size_t)nqc * sizeof(double),
                           cudaMemcpyDeviceToHost, h->st_c));
      }
    }
    if (waited < K - 1) CK(cudaStreamWaitEvent(st, h->ev_up[K - 1], 0));   // beams and the node gather read every row
    if (S.n_beam > 0) {
      beam_adjoint_kernel<<<cdiv(2LL * S.n_beam, 128), 128, 0, st>>>(S.n_beam, h->s_crds, h->cnct_b, h->s_pb, h->s_f, h->s_u,
                                                                    dc_h ? h->corner_b : nullptr,
                                                                    (dpb_h && nb) ? h->s_dpb : nullptr, h->flags);
      CKL("beam_adjoint_kernel");
    }
    if (dc_h) {
      node_gather_kernel<<<cdiv(3LL * S.n_node, 256), 256, 0, st>>>(S.n_node, S.n_quad, h->node_inc_ptr, h->node_inc,
                                                                   h->corner_q, h->corner_b, h->s_dc);
      CKL("node_gather_kernel");
    }
    const bool p_dc = dc_h && pinned(dc_h), p_db = dpb_h && pinned(dpb_h);
    if (dc_h) CK(cudaMemcpyAsync(p_dc ? dc_h : h->h_dc, h->s_dc, nc * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (dpb_h && nb) CK(cudaMemcpyAsync(p_db ? dpb_h : h->h_dpb, h->s_dpb, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamSynchronize(h->st_c));
    if (dc_h && !p_dc) par_memcpy(dc_h, h->h_dc, nc * sizeof(double));
    if (want_dq && !p_dq) par_memcpy(dpq_h, h->h_dpq, nq * sizeof(double));
    if (dpb_h && nb && !p_db) par_memcpy(dpb_h, h->h_dpb, nb * sizeof(double));
    return JSSO_OK;
  }
  CK(h2d(h->s_f, u_h, h->h_f, nd, st2));
  CK(h2d(h->s_u, lam_h, h->h_u, nd, st2));
  CK(cudaEventRecord(h->ev_b, st2));
  CK(cudaStreamWaitEvent(st, h->ev_b, 0));
  if ((rc = jsso_adjoint(h, h->s_crds, h->s_pq, h->s_pb, h->s_f, h->s_u, dc_h ? h->s_dc : nullptr,
                         (dpq_h && nq) ? h->s_dpq : nullptr, (dpb_h && nb) ? h->s_dpb : nullptr, st)))
    return rc;
  const bool p_dc = dc_h && pinned(dc_h), p_dq = dpq_h && pinned(dpq_h), p_db = dpb_h && pinned(dpb_h);
  if (dc_h) CK(cudaMemcpyAsync(p_dc ? dc_h : h->h_dc, h->s_dc, nc * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (dpq_h && nq) CK(cudaMemcpyAsync(p_dq ? dpq_h : h->h_dpq, h->s_dpq, nq * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (dpb_h && nb) CK(cudaMemcpyAsync(p_db ? dpb_h : h->h_dpb, h->s_dpb, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (dc_h && !p_dc) par_memcpy(dc_h, h->h_dc, nc * sizeof(double));
  if (dpq_h && nq && !p_dq) par_memcpy(dpq_h, h->h_dpq, nq * sizeof(double));
  if (dpb_h && nb && !p_db) par_memcpy(dpb_h, h->h_dpb, nb * sizeof(double));
  return JSSO_OK;
}

// ---------------------------------------------------------------- utilities
int jsso_set_device(int device) { return cudaSetDevice(device) == cudaSuccess ? JSSO_OK : JSSO_ERR_CUDA; }
void* jsso_dev_alloc(size_t bytes) { void* p = nullptr; return cudaMalloc(&p, bytes ? bytes : 8) == cudaSuccess ? p : nullptr; }
void jsso_dev_free(void* p) { if (p) cudaFree(p); }
void* jsso_host_alloc_pinned(size_t bytes) { void* p = nullptr; return cudaMallocHost(&p, bytes ? bytes : 8) == cudaSuccess ? p : nullptr; }
void jsso_host_free_pinned(void* p) { if (p) cudaFreeHost(p); }
int jsso_memcpy_h2d(void* d, const void* s, size_t n, void* st) {
  return cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, (cudaStream_t)st) == cudaSuccess ? JSSO_OK : JSSO_ERR_CUDA;
}
int jsso_memcpy_d2h(void* d, const void* s, size_t n, void* st) {
  if (cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, (cudaStream_t)st) != cudaSuccess) return JSSO_ERR_CUDA;
  return cudaStreamSynchronize((cudaStream_t)st) == cudaSuccess ? JSSO_OK : JSSO_ERR_CUDA;
}
int jsso_memset(void* d, int v, size_t n, void* st) {
  return cudaMemsetAsync(d, v, n, (cudaStream_t)st) == cudaSuccess ? JSSO_OK : JSSO_ERR_CUDA;
}
int jsso_stream_sync(void* st) { return cudaStreamSynchronize((cudaStream_t)st) == cudaSuccess ? JSSO_OK : JSSO_ERR_CUDA; }
int jsso_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
void* jsso_event_create(void) { cudaEvent_t e; return cudaEventCreate(&e) == cudaSuccess ? (void*)e : nullptr; }
int jsso_event_record(void* ev, void* st) {
  return cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)st) == cudaSuccess ? JSSO_OK : JSSO_ERR_CUDA;
}
int jsso_event_elapsed_ms(void* a, void* b, float* ms) {
  if (cudaEventSynchronize((cudaEvent_t)b) != cudaSuccess) return JSSO_ERR_CUDA;
  return cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b) == cudaSuccess ? JSSO_OK : JSSO_ERR_CUDA;
}
void jsso_event_destroy(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
int64_t jsso_launch_count(void) { return g_launches.load(); }
// range markers for `ncu --profile-from-start off` (scripts/mg_profile.py)
int jsso_profiler_range(int start) {
#ifndef JSSO_EMU
  return (start ? cudaProfilerStart() : cudaProfilerStop()) == cudaSuccess ? JSSO_OK : JSSO_ERR_CUDA;
#else
  (void)start; return JSSO_OK;
#endif
}

}  // extern "C"
