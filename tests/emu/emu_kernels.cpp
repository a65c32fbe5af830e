// TEST INFRASTRUCTURE: the product's kernel headers compiled by g++ against the SIMT emulator (cuda_emu.h) and
// exported with a C ABI for the CPU tests (tests/test_emu_kernels.py).  Never linked into libjsso.so.
#include "cuda_emu.h"

#include "../../jaxsso_b200/csrc/jsso_solver.cuh"
#include "../../jaxsso_b200/csrc/jsso_multigrid.cuh"

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

}  // extern "C"
