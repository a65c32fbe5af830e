// TEST INFRASTRUCTURE: the product's kernel headers compiled by g++ against the SIMT emulator (cuda_emu.h) and
// exported with a C ABI for the CPU tests (tests/test_emu_kernels.py).  Never linked into libjsso.so.
#include "cuda_emu.h"

#include "../../jaxsso_b200/csrc/jsso_adjoint.cuh"
#include "../../jaxsso_b200/csrc/jsso_assemble.cuh"
#include "../../jaxsso_b200/csrc/jsso_solver.cuh"
#include "../../jaxsso_b200/csrc/jsso_multigrid.cuh"
#include "../../jaxsso_b200/csrc/jsso_symbolic.h"

using namespace jsso;

namespace {
template <int MODE, class VT>
void run_axpby(int n_row, const int32_t* rp, const int32_t* ci, const void* v, const double* x, double* y,
               const double* b, int grid) {
  emu::launch(grid, RED_BLOCK, 0, [&] { bsr_spmv_axpby_kernel<MODE, VT>(n_row, rp, ci, (const VT*)v, x, y, b); });
}
template <class VT>
int run_axpby_mode(int mode, int n_row, const int32_t* rp, const int32_t* ci, const void* v, const double* x,
                   double* y, const double* b, int grid) {
  switch (mode) {
    case 0: run_axpby<0, VT>(n_row, rp, ci, v, x, y, b, grid); return 0;
    case 2: run_axpby<2, VT>(n_row, rp, ci, v, x, y, b, grid); return 0;
    case 3: run_axpby<3, VT>(n_row, rp, ci, v, x, y, b, grid); return 0;
  }
  return 1;
}
}  // namespace

#pragma GCC visibility push(default)
extern "C" {

// y (op)= A x over n_row rows; vt: 0 double (column-major blocks), 1 float, 2 binary16 (row-pair-major blocks)
int emu_spmv_axpby(int mode, int vt, int n_row, const int32_t* rp, const int32_t* ci, const void* v, const double* x,
                   double* y, const double* b, int grid) {
  if (vt == 0) return run_axpby_mode<double>(mode, n_row, rp, ci, v, x, y, b, grid);
  if (vt == 1) return run_axpby_mode<float>(mode, n_row, rp, ci, v, x, y, b, grid);
  if (vt == 2) return run_axpby_mode<__half>(mode, n_row, rp, ci, v, x, y, b, grid);
  return 1;
}

int emu_spmv_short(int mode, int n_row, const int32_t* rp, const int32_t* ci, const float* v, const double* x,
                   double* y, const double* b) {
  const unsigned grid = (unsigned)((3LL * n_row + 255) / 256);
  switch (mode) {
    case 0: emu::launch(grid, 256, 0, [&] { bsr_spmv_short_kernel<0>(n_row, rp, ci, v, x, y, b); }); return 0;
    case 2: emu::launch(grid, 256, 0, [&] { bsr_spmv_short_kernel<2>(n_row, rp, ci, v, x, y, b); }); return 0;
    case 3: emu::launch(grid, 256, 0, [&] { bsr_spmv_short_kernel<3>(n_row, rp, ci, v, x, y, b); }); return 0;
  }
  return 1;
}

// the CG SpMV (MODE 0: y = A x), persistent grid
void emu_spmv_main(int n_row, const int32_t* rp, const int32_t* ci, const double* v, const double* x, double* y, int grid) {
  emu::launch(grid, RED_BLOCK, 0, [&] {
    bsr_spmv_kernel<0>(n_row, rp, ci, v, x, y, nullptr, 0, nullptr, nullptr, 1, nullptr, 0ull, 0ull);
  });
}

void emu_to_float(long long n, const double* a, float* b) {
  emu::launch(2, 256, 0, [&] { mg_to_float_kernel(n, a, b); });
}
void emu_to_half(long long n, const double* a, void* b) {
  emu::launch(2, 256, 0, [&] { mg_to_half_kernel(n, a, (__half*)b); });
}
void emu_halo_pack(int n, const int32_t* idx, const double* v, double* buf) {
  emu::launch((6 * n + 255) / 256, 256, 0, [&] { halo_pack_kernel(n, idx, v, buf); });
}
void emu_halo_unpack(int n, const int32_t* idx, const double* buf, double* v) {
  emu::launch((6 * n + 255) / 256, 256, 0, [&] { halo_unpack_kernel(n, idx, buf, v); });
}
// out = a . b with the deterministic two-stage grid reduction
double emu_dot(long long n, const double* a, const double* b, int grid) {
  std::vector<double> partials(2 * RED_MAX_BLOCKS, 0.0);
  unsigned counter = 0;
  double out = 0.0;
  emu::launch(grid, 256, 0, [&] { mg_dot_kernel(n, a, b, partials.data(), &counter, &out); });
  return out;
}
// Chebyshev smoother step (FIRST = 1: d = c2 r, x = zero_guess ? d : x + d; else d = c1 d + c2 r, x += d)
void emu_cheb(int first, int n, const double* Dinv, const double* r, double* d, double* x, double c1, double c2, int zero_guess) {
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (first) emu::launch(grid, 128, 0, [&] { mg_cheb_kernel<1>(n, Dinv, r, d, x, c1, c2, zero_guess); });
  else emu::launch(grid, 128, 0, [&] { mg_cheb_kernel<0>(n, Dinv, r, d, x, c1, c2, zero_guess); });
}


// ---- Ke + numeric assembly, the launches of jsso_assemble (csrc/jsso_api.cu).  path: 0 = two-kernel warp tasks
// (quad_geometry_kernel + assemble_tasks_kernel), 1 = chunked single kernel (assemble_fused_kernel).
// vals_out: nnzb x 36 (column-major blocks); rowptr/colidx as jsso_pattern.  Returns 0, or 2 if the mesh cannot
// use the task path, or 1 on a symbolic error.  task_ctas: persistent grid of the task kernel (small in tests).
int emu_assemble(int path, int n_node, int n_quad, const int32_t* cq, int n_beam, const int32_t* cb, int n_known,
                 const int32_t* known, const double* crds, const double* prop_q, const double* prop_b, int apply_bc,
                 double* vals_out, int32_t* flags_out, int task_ctas) {
  Symbolic S;
  if (!build_symbolic(n_node, n_node, n_quad, cq, n_beam, cb, n_known, known, S).empty()) return 1;
  int flags = 0;
  if (S.nnzb() == 0) return 0;
  if (path == 0) {
    if (!S.tasks_ok) return 2;
    std::vector<double> rec((size_t)std::max(n_quad, 1) * REC_GLD, 0.0);
    if (n_quad > 0)
      emu::launch((n_quad + G_QUADS - 1) / G_QUADS, G_THREADS, G_QUADS * QS * sizeof(double),
                  [&] { quad_geometry_kernel(n_quad, crds, cq, prop_q, rec.data(), &flags); });
    TaskArgs T;
    T.rec = rec.data(); T.task_meta = (const int4*)S.task_meta.data(); T.task_els = S.task_els.data();
    T.item_desc = S.item_desc.data(); T.blk_bc = S.blk_bc.data(); T.item_code = S.item_code.data();
    T.blk_item_ptr = S.blk_item_ptr.data();
    T.crds = crds; T.cnct_b = cb; T.prop_b = prop_b;
    T.vals = vals_out; T.flags = &flags; T.n_quad = n_quad; T.n_task = S.n_task(); T.apply_bc = apply_bc;
    const int grid = std::min((T.n_task + TASK_WARPS - 1) / TASK_WARPS, std::max(1, task_ctas));
    emu::launch(grid, 32 * TASK_WARPS, TASK_WARPS * TASK_SMEM_DOUBLES * sizeof(double), [&] { assemble_tasks_kernel(T); });
  } else {
    AsmArgs A;
    A.crds = crds; A.cnct_q = cq; A.prop_q = prop_q; A.cnct_b = cb; A.prop_b = prop_b;
    A.chunk_blk = S.chunk_blk.data(); A.chunk_el_ptr = S.chunk_el_ptr.data(); A.chunk_els = S.chunk_els.data();
    A.blk_perm = S.blk_perm.data(); A.blk_item_ptr = S.blk_item_ptr.data(); A.item_code = S.item_code.data();
    A.item_lel = S.item_lel.data(); A.blk_row = S.blk_row.data(); A.colidx = S.colidx.data();
    A.node_mask = S.node_mask.data(); A.vals = vals_out; A.flags = &flags; A.n_quad = n_quad; A.apply_bc = apply_bc;
    emu::launch(S.n_chunk(), kChunkBlocks, FUSED_SMEM_DOUBLES * sizeof(double), [&] { assemble_fused_kernel(A); });
  }
  if (flags_out) *flags_out = flags;
  return 0;
}

// ---- adjoint reduction, the launches of jsso_adjoint: d_crds (n_node x 3), d_prop_q (n_quad x 5), d_prop_b (n_beam x 6);
// any output may be null.  adj_ctas: persistent grid of the quad kernel.
int emu_adjoint(int n_node, int n_quad, const int32_t* cq, int n_beam, const int32_t* cb, const double* crds,
                const double* prop_q, const double* prop_b, const double* u, const double* lam, double* d_crds,
                double* d_prop_q, double* d_prop_b, int adj_ctas) {
  Symbolic S;
  if (!build_symbolic(n_node, n_node, n_quad, cq, n_beam, cb, 0, nullptr, S).empty()) return 1;
  std::vector<double> corner_q((size_t)std::max(n_quad, 1) * 12, 0.0), corner_b((size_t)std::max(n_beam, 1) * 6, 0.0);
  int flags = 0;
  if (n_quad > 0) {
    const int blocks = std::min((n_quad + ADJ_QUADS - 1) / ADJ_QUADS, std::max(1, adj_ctas));
    if (d_prop_q)
      emu::launch(blocks, 4 * ADJ_QUADS, ADJ_SMEM_DOUBLES * sizeof(double), [&] {
        quad_adjoint_kernel<true>(n_quad, crds, cq, prop_q, u, lam, d_crds ? corner_q.data() : nullptr, d_prop_q);
      });
    else
      emu::launch(blocks, 4 * ADJ_QUADS, ADJ_SMEM_DOUBLES * sizeof(double), [&] {
        quad_adjoint_kernel<false>(n_quad, crds, cq, prop_q, u, lam, d_crds ? corner_q.data() : nullptr, nullptr);
      });
  }
  if (n_beam > 0)
    emu::launch((2 * n_beam + 127) / 128, 128, 0, [&] {
      beam_adjoint_kernel(n_beam, crds, cb, prop_b, u, lam, d_crds ? corner_b.data() : nullptr, d_prop_b, &flags);
    });
  if (d_crds)
    emu::launch((3 * n_node + 255) / 256, 256, 0, [&] {
      node_gather_kernel(n_node, n_quad, S.node_inc_ptr.data(), S.node_inc.data(), corner_q.data(), corner_b.data(), d_crds);
    });
  return 0;
}

// pattern of the symbolic pass (sizes first with null outputs)
int emu_pattern(int n_node, int n_quad, const int32_t* cq, int n_beam, const int32_t* cb, int64_t* nnzb, int32_t* rowptr,
                int32_t* colidx) {
  Symbolic S;
  if (!build_symbolic(n_node, n_node, n_quad, cq, n_beam, cb, 0, nullptr, S).empty()) return 1;
  *nnzb = S.nnzb();
  if (rowptr) std::memcpy(rowptr, S.rowptr.data(), sizeof(int32_t) * S.rowptr.size());
  if (colidx) std::memcpy(colidx, S.colidx.data(), sizeof(int32_t) * S.colidx.size());
  return 0;
}

}  // extern "C"
#pragma GCC visibility pop
