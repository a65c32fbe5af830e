// Smoothed-aggregation multigrid preconditioner on 6x6 block-CSR (numeric kernels).
//
// Not part of the reference (which solves with SuperLU, JaxSSO/solver.py:195-197): it only
// changes how fast the same u is reached.  The symbolic part (aggregates, patterns, gather
// lists) is built once per model by jaxsso_b200/multigrid.py; everything here is numeric and
// re-run after each assembly.  All sparse products are gathers over precomputed
// (left slot, right slot) lists: no atomics, fixed summation order.  Blocks are column-major
// (entry (i,j) at 6 j + i) like the stiffness values.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace jsso {

// rigid-body block T = [[I, -[r]x], [0, I]] (column-major), rows of prescribed dofs zeroed,
// optionally left-multiplied by Lt (row-major 6x6: the transposed Cholesky factor of the node's
// diagonal block, when the level matrix is the block-Jacobi-scaled one)
__device__ inline void rigid_block(const double* X, const double* Xc, int i, int a, unsigned mask,
                                   const double* LtRow /* L row-major = Lt column-major, or null */,
                                   double* T /* 36, column-major */) {
  const double rx = X[3 * (size_t)i] - Xc[3 * (size_t)a], ry = X[3 * (size_t)i + 1] - Xc[3 * (size_t)a + 1],
               rz = X[3 * (size_t)i + 2] - Xc[3 * (size_t)a + 2];
  double B[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) B[k] = 0.0;
#pragma unroll
  for (int d = 0; d < 6; ++d) B[6 * d + d] = 1.0;
  // translation rows vs rotation columns: -[r]x = [[0, rz, -ry], [-rz, 0, rx], [ry, -rx, 0]]
  B[6 * 4 + 0] = rz;  B[6 * 5 + 0] = -ry;
  B[6 * 3 + 1] = -rz; B[6 * 5 + 1] = rx;
  B[6 * 3 + 2] = ry;  B[6 * 4 + 2] = -rx;
#pragma unroll
  for (int r = 0; r < 6; ++r)
    if ((mask >> r) & 1u) {
#pragma unroll
      for (int c = 0; c < 6; ++c) B[6 * c + r] = 0.0;
    }
  if (!LtRow) {
#pragma unroll
    for (int k = 0; k < 36; ++k) T[k] = B[k];
    return;
  }
  // T = L^T B : T[r][c] = sum_k L[k][r] B[k][c], L lower triangular (row-major LtRow[k*6+r])
#pragma unroll
  for (int c = 0; c < 6; ++c)
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
#pragma unroll
      for (int k = r; k < 6; ++k) s += LtRow[k * 6 + r] * B[6 * c + k];
      T[6 * c + r] = s;
    }
}

// acc (col-major) += A (col-major) * B (col-major)
__device__ inline void blk_mma(const double* __restrict__ A, const double* B, double* acc) {
  double a[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) a[k] = A[k];
#pragma unroll
  for (int c = 0; c < 6; ++c)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double b = B[6 * c + k];
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[6 * c + r] = fma(a[6 * k + r], b, acc[6 * c + r]);
    }
}
// acc += A^T * B
__device__ inline void blk_mma_t(const double* __restrict__ A, const double* __restrict__ B, double* acc) {
  double a[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) a[k] = A[k];
#pragma unroll
  for (int c = 0; c < 6; ++c)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double b = B[6 * c + k];
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[6 * c + r] = fma(a[6 * r + k], b, acc[6 * c + r]);
    }
}

// centroid of every aggregate (members listed in ascending node order)
__global__ void mg_centroid_kernel(int n_c, const int32_t* __restrict__ mem_ptr, const int32_t* __restrict__ mem,
                                   const double* __restrict__ X, double* __restrict__ Xc) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_c) return;
  double s0 = 0, s1 = 0, s2 = 0;
  const int m0 = mem_ptr[a], m1 = mem_ptr[a + 1];
  for (int k = m0; k < m1; ++k) {
    const int i = mem[k];
    s0 += X[3 * (size_t)i]; s1 += X[3 * (size_t)i + 1]; s2 += X[3 * (size_t)i + 2];
  }
  const double inv = 1.0 / (double)(m1 - m0);
  Xc[3 * (size_t)a] = s0 * inv; Xc[3 * (size_t)a + 1] = s1 * inv; Xc[3 * (size_t)a + 2] = s2 * inv;
}

// P[s] = own(s) T~_i - omega * Dinv_i * sum_k A[a_k] T~_{j_k},  T~ = (Lt) T;  thread per P block
__global__ void __launch_bounds__(128)
mg_smooth_prolongator_kernel(int nnz_p, const int32_t* __restrict__ p_row, const int32_t* __restrict__ p_col,
                             const int32_t* __restrict__ p_own, const int32_t* __restrict__ ps_ptr,
                             const int32_t* __restrict__ ps_a, const int32_t* __restrict__ ps_j,
                             const int32_t* __restrict__ agg, const double* __restrict__ A,
                             const double* __restrict__ Dinv /* row-major or null */,
                             const double* __restrict__ Lfac /* row-major L per node or null */,
                             const uint8_t* __restrict__ mask, const double* __restrict__ X,
                             const double* __restrict__ Xc, double omega, double* __restrict__ P,
                             const int32_t* __restrict__ list = nullptr /* nnz_p slots to compute, or all */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz_p) return;
  const int s = list ? list[t] : t;
  const int i = p_row[s];
  double acc[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = 0.0;
  for (int k = ps_ptr[s]; k < ps_ptr[s + 1]; ++k) {
    const int j = ps_j[k];
    double T[36];
    rigid_block(X, Xc, j, agg[j], mask ? mask[j] : 0u, Lfac ? Lfac + (size_t)j * 36 : nullptr, T);
    blk_mma(A + (size_t)ps_a[k] * 36, T, acc);
  }
  double out[36];
  if (p_own[s]) rigid_block(X, Xc, i, agg[i], mask ? mask[i] : 0u, Lfac ? Lfac + (size_t)i * 36 : nullptr, out);
  else {
#pragma unroll
    for (int k = 0; k < 36; ++k) out[k] = 0.0;
  }
  if (Dinv) {
    const double* di = Dinv + (size_t)i * 36;
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) t += di[r * 6 + k] * acc[6 * c + k];
        out[6 * c + r] -= omega * t;
      }
  } else {
#pragma unroll
    for (int k = 0; k < 36; ++k) out[k] -= omega * acc[k];
  }
  double* o = P + (size_t)s * 36;
#pragma unroll
  for (int k = 0; k < 36; ++k) o[k] = out[k];
}

// out[s] = sum_k op(L[l_k]) * R[r_k]   (TRANS: op = transpose).  Six threads per output block, thread c owns column
// c: it reads column c of R (48 contiguous bytes) and the whole L block (the six threads read the same 288 bytes, one
// set of sectors), and the six columns leave as 288 contiguous bytes.  Every entry is summed in the order of the
// thread-per-block form it replaced -- bitwise the same hierarchy.  Measured at 1M quads: 4.4 + 2.4 ms for the two
// Galerkin products of a numeric setup, the same as the thread-per-block form (4.3 + 2.5 ms): coalescing was not the
// limit (profiles/r2x_launches_numeric_setup.txt); ~7 GB of DRAM traffic would be 1.1 ms.
constexpr int MG_PROD_THREADS = 192;   // 32 output blocks per CTA
template <int TRANS>
__global__ void __launch_bounds__(MG_PROD_THREADS)
mg_block_product_kernel(int nnz_out, const int32_t* __restrict__ ptr, const int32_t* __restrict__ li,
                        const int32_t* __restrict__ ri, const double* __restrict__ Lv,
                        const double* __restrict__ Rv, double* __restrict__ out,
                        const int32_t* __restrict__ list = nullptr /* nnz_out slots to compute */, int first = 0) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int t = (int)(g / 6), c = (int)(g - 6LL * t);
  if (t >= nnz_out) return;
  const int s = list ? list[t] : first + t;
  double acc[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) acc[r] = 0.0;
  for (int k = ptr[s]; k < ptr[s + 1]; ++k) {
    const double2* A2 = (const double2*)(Lv + (size_t)li[k] * 36);
    const double2* B2 = (const double2*)(Rv + (size_t)ri[k] * 36 + 6 * c);
    const double2 b01 = B2[0], b23 = B2[1], b45 = B2[2];
    const double b[6] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y};
    double a[36];
#pragma unroll
    for (int q = 0; q < 18; ++q) { const double2 v = A2[q]; a[2 * q] = v.x; a[2 * q + 1] = v.y; }
#pragma unroll
    for (int kk = 0; kk < 6; ++kk)
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[r] = fma(TRANS ? a[6 * r + kk] : a[6 * kk + r], b[kk], acc[r]);
  }
  double2* o = (double2*)(out + (size_t)s * 36 + 6 * c);
  o[0] = make_double2(acc[0], acc[1]);
  o[1] = make_double2(acc[2], acc[3]);
  o[2] = make_double2(acc[4], acc[5]);
}

// P[s] <- P[s] W_c^T, c = p_col[s] (W_c lower triangular, row-major): the prolongator into the block-Jacobi-SCALED
// coordinates of the coarse level (every level of the hierarchy is kept in scaled form, x_l = W_l^T y_l, so that its
// diagonal blocks are the identity and the Chebyshev smoother needs no D^-1); thread per block
__global__ void __launch_bounds__(128)
mg_scale_cols_kernel(int n, const int32_t* __restrict__ p_col, const double* __restrict__ Wc, double* __restrict__ P,
                     const int32_t* __restrict__ list = nullptr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int s = list ? list[t] : t;
  const double* w = Wc + (size_t)p_col[s] * 36;
  double* pb = P + (size_t)s * 36;
  double a[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) a[k] = pb[k];
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k <= j; ++k) v += a[6 * k + i] * w[j * 6 + k];   // (P W^T)[i][j] = sum_k P[i][k] W[j][k]
      pb[6 * j + i] = v;
    }
}

// Pt[s] = P[src[s]]^T
__global__ void mg_transpose_blocks_kernel(int nnz, const int32_t* __restrict__ src, const double* __restrict__ P,
                                           double* __restrict__ Pt, int first = 0 /* slots [first, first + nnz) */) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 36LL * nnz) return;
  const int s = first + (int)(t / 36), k = (int)(t % 36);
  const int r = k % 6, c = k / 6;
  Pt[(size_t)s * 36 + k] = P[(size_t)src[s] * 36 + 6 * r + c];
}

// Chebyshev smoother step on D^-1 A, thread per node:
//   r <- Dinv r (if Dinv);  FIRST: d = r / theta, x = zero_guess ? d : x + d
//                           else : d = c1 d + c2 r, x += d
template <int FIRST>
__global__ void __launch_bounds__(128)
mg_cheb_kernel(int n, const double* __restrict__ Dinv, const double* __restrict__ r, double* __restrict__ d,
               double* __restrict__ x, double c1, double c2, int zero_guess) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rv[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) rv[k] = r[6 * (size_t)i + k];
  if (Dinv) {
    const double* di = Dinv + (size_t)i * 36;
    double t[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += di[a * 6 + k] * rv[k];
      t[a] = s;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) rv[k] = t[k];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const size_t o = 6 * (size_t)i + k;
    double dv;
    if (FIRST) dv = rv[k] * c2;   // c2 = 1 / theta
    else dv = c1 * d[o] + c2 * rv[k];
    d[o] = dv;
    x[o] = (FIRST && zero_guess) ? dv : x[o] + dv;
  }
}

// ---- coarsest level: dense inverse (n <= a few hundred), one CTA, matrix in global memory
__global__ void mg_dense_from_bsr_kernel(int n_node, const int32_t* __restrict__ rowptr,
                                         const int32_t* __restrict__ colidx, const double* __restrict__ vals,
                                         double* __restrict__ M /* (6n) x (12n) row-major: [A | I] */) {
  const int n = 6 * n_node;
  const long long total = (long long)n * 2 * n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / (2 * n)), c = (int)(t % (2 * n));
    M[t] = (c >= n && c - n == r) ? 1.0 : 0.0;
  }
}
__global__ void mg_dense_fill_kernel(int n_node, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                     const double* __restrict__ vals, double* __restrict__ M) {
  const int n = 6 * n_node;
  const int br = blockIdx.x;
  for (int s = rowptr[br]; s < rowptr[br + 1]; ++s) {
    const int bc = colidx[s];
    for (int k = threadIdx.x; k < 36; k += blockDim.x) {
      const int i = k % 6, j = k / 6;
      M[(size_t)(6 * br + i) * 2 * n + 6 * bc + j] = vals[(size_t)s * 36 + k];
    }
  }
}
// Gauss-Jordan on [A | I] without pivoting (A SPD; empty rows -> identity); one CTA
__global__ void __launch_bounds__(1024)
mg_dense_invert_kernel(int n, double* __restrict__ M) {
  __shared__ double piv;
  const int w = 2 * n;
  for (int k = 0; k < n; ++k) {
    if (threadIdx.x == 0) {
      double p = M[(size_t)k * w + k];
      if (!(fabs(p) > 1e-300)) { p = 1.0; M[(size_t)k * w + k] = 1.0; }
      piv = 1.0 / p;
    }
    __syncthreads();
    const double ip = piv;
    for (int j = threadIdx.x; j < w; j += blockDim.x) M[(size_t)k * w + j] *= ip;
    __syncthreads();
    // eliminate column k from all other rows: thread handles (row i, column chunk)
    for (int i = threadIdx.x >> 5; i < n; i += blockDim.x >> 5) {
      if (i == k) continue;
      const double f = M[(size_t)i * w + k];
      if (f == 0.0) continue;
      for (int j = (threadIdx.x & 31); j < w; j += 32)
        if (j != k) M[(size_t)i * w + j] = fma(-f, M[(size_t)k * w + j], M[(size_t)i * w + j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (i != k) M[(size_t)i * w + k] = 0.0;
    __syncthreads();
  }
}
// ---- blocked Gauss-Jordan for a coarsest level of up to a few thousand unknowns ----------------------------------
// [A | I] -> [I | A^-1] in pivot blocks of MG_DB rows, three launches per block, all SMs: (1) invert the pivot block
// W[K,K] in shared memory (no pivoting: A is SPD, so is every pivot block of its Schur complements; an empty row becomes
// an identity row as in the one-CTA kernel above), (2) row panel W[K,:] <- P^-1 W[K,:] (kept in `rpanel`) and a copy of
// the column panel W[:,K] (`cpanel`), (3) W[i,:] -= cpanel[i,:] rpanel for every other row.  2 n^3 multiply-adds in
// n / 32 steps.  Measured at 1M quads: the default coarsest level (20 nodes, n = 120) 0.51 -> 0.16 ms per numeric setup;
// n = 1020 (hierarchy stopped one level earlier, max_coarse_nodes = 256) 1.9 ms, where the one-CTA kernel -- every pivot
// a sweep of the whole matrix through one SM -- would need tens of milliseconds.  Stopping there did NOT pay, which is
// why 64 nodes stays the default: the 8 MB dense solve costs what the four small products it replaces cost, and the
// PCG needed 167 instead of 166 iterations (profiles/r2aa_coarsest_level_ab.txt).
constexpr int MG_DB = 32;
__global__ void __launch_bounds__(MG_DB * MG_DB)
mg_dense_pivot_kernel(int w, int k0, int nb, const double* __restrict__ W, double* __restrict__ pinv) {
  __shared__ double A[MG_DB][MG_DB + 1], B[MG_DB][MG_DB + 1];
  __shared__ double piv;
  const int i = threadIdx.x / MG_DB, j = threadIdx.x % MG_DB;
  A[i][j] = (i < nb && j < nb) ? W[(size_t)(k0 + i) * w + k0 + j] : (i == j ? 1.0 : 0.0);
  B[i][j] = (i == j) ? 1.0 : 0.0;
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    if (i == 0 && j == 0) {
      double p = A[k][k];
      if (!(fabs(p) > 1e-300)) { p = 1.0; A[k][k] = 1.0; }
      piv = 1.0 / p;
    }
    __syncthreads();
    if (i == k) { A[k][j] *= piv; B[k][j] *= piv; }
    __syncthreads();
    const double f = A[i][k];
    __syncthreads();
    if (i != k) { A[i][j] = fma(-f, A[k][j], A[i][j]); B[i][j] = fma(-f, B[k][j], B[i][j]); }
    __syncthreads();
  }
  pinv[i * MG_DB + j] = (i < nb && j < nb) ? B[i][j] : 0.0;
}
// threads [0, w): column j of the row panel; threads [w, w + n * MG_DB): one entry of the column-panel copy
__global__ void __launch_bounds__(256)
mg_dense_panel_kernel(int n, int w, int k0, int nb, double* __restrict__ W, const double* __restrict__ pinv,
                      double* __restrict__ rpanel, double* __restrict__ cpanel) {
  __shared__ double P[MG_DB * MG_DB];
  for (int t = threadIdx.x; t < MG_DB * MG_DB; t += blockDim.x) P[t] = pinv[t];
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < w) {
    const int j = (int)t;
    double old[MG_DB];
#pragma unroll
    for (int k = 0; k < MG_DB; ++k) old[k] = (k < nb) ? W[(size_t)(k0 + k) * w + j] : 0.0;
    for (int i = 0; i < nb; ++i) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < MG_DB; ++k) v = fma(P[i * MG_DB + k], old[k], v);
      W[(size_t)(k0 + i) * w + j] = v;
      rpanel[(size_t)i * w + j] = v;
    }
  } else if (t < (long long)w + (long long)n * MG_DB) {
    const long long q = t - w;
    const int i = (int)(q / MG_DB), k = (int)(q - (long long)i * MG_DB);
    const bool pivot_row = i >= k0 && i < k0 + nb;            // rewritten by the threads above and not used by the update
    cpanel[q] = (k < nb && !pivot_row) ? W[(size_t)i * w + k0 + k] : 0.0;
  }
}
constexpr int MG_DU_ROWS = 16, MG_DU_COLS = 128;
__global__ void __launch_bounds__(MG_DU_COLS)
mg_dense_update_kernel(int n, int w, int k0, int nb, double* __restrict__ W, const double* __restrict__ rpanel,
                       const double* __restrict__ cpanel) {
  __shared__ double R[MG_DB][MG_DU_COLS];
  __shared__ double Cs[MG_DU_ROWS][MG_DB];
  const int n_col_tile = (w + MG_DU_COLS - 1) / MG_DU_COLS;       // 1-D grid: column tiles fastest
  const int j = (blockIdx.x % n_col_tile) * MG_DU_COLS + threadIdx.x, i0 = (blockIdx.x / n_col_tile) * MG_DU_ROWS;
  for (int k = 0; k < MG_DB; ++k) R[k][threadIdx.x] = (k < nb && j < w) ? rpanel[(size_t)k * w + j] : 0.0;
  for (int t = threadIdx.x; t < MG_DU_ROWS * MG_DB; t += MG_DU_COLS) {
    const int r = t / MG_DB, k = t - r * MG_DB;
    Cs[r][k] = (i0 + r < n) ? cpanel[(size_t)(i0 + r) * MG_DB + k] : 0.0;    // zero on the pivot rows themselves
  }
  __syncthreads();
  if (j >= w) return;
#pragma unroll 4
  for (int r = 0; r < MG_DU_ROWS; ++r) {
    const int i = i0 + r;
    if (i >= n || (i >= k0 && i < k0 + nb)) continue;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < MG_DB; ++k) acc = fma(Cs[r][k], R[k][threadIdx.x], acc);
    W[(size_t)i * w + j] -= acc;
  }
}

// x = Ainv b with Ainv = right half of M; one warp per row
__global__ void mg_dense_matvec_kernel(int n, const double* __restrict__ M, const double* __restrict__ b,
                                       double* __restrict__ x) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* a = M + (size_t)row * 2 * n + n;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s = fma(a[j], b[j], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) x[row] = s;
}

// ---- small vector kernels of the host-driven outer PCG
// out[0] = a.b (deterministic two-stage); n entries
__global__ void __launch_bounds__(256)
mg_dot_kernel(long long n, const double* __restrict__ a, const double* __restrict__ b, double* partials,
              unsigned* counter, double* out, const MgdCtx* rctx = nullptr, unsigned long long* rseq = nullptr) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc = fma(a[i], b[i], acc);
  double total;
  if (grid_sum(acc, partials, counter, total)) mgs_store_dot_block(total, out, rctx, rseq);
}
// y = a*x + b*y
__global__ void mg_axpby_kernel(long long n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = a * x[i] + b * y[i];
}
// ---- outer PCG with the scalars kept on the device (slots: MGS_* in jsso_solver.cuh)
// p = z + (rz / rz_old) p   (first: p = z)
__global__ void mg_pcg_dir_kernel(long long n, const double* __restrict__ z, double* __restrict__ p,
                                  const double* __restrict__ scal, int first) {
  if (mgs_stopped(scal)) return;
  const double beta = first ? 0.0 : scal[MGS_RZ] / scal[MGS_RZ_OLD];
  const long long n2 = n >> 1;   // n = 6 x nodes is even and the vectors are 16-byte aligned: 16-byte accesses
  const double2* z2 = (const double2*)z;
  double2* p2 = (double2*)p;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const double2 zv = z2[i];
    if (first) { p2[i] = zv; continue; }
    const double2 pv = p2[i];
    p2[i] = make_double2(fma(beta, pv.x, zv.x), fma(beta, pv.y, zv.y));
  }
}
// alpha = rz / pq;  x += alpha p;  r -= alpha q;  *rr_out = r.r;  the block that finalises the sum also latches
// rz_old = rz and counts the iteration (it runs after every block has read alpha and passed the stop test).
// Non-positive curvature or r.z (matrix or preconditioner not SPD) poisons r.r with NaN, which stops everything.
__global__ void __launch_bounds__(256)
mg_pcg_update_kernel(long long n, const double* __restrict__ p, const double* __restrict__ q,
                     double* __restrict__ x, double* __restrict__ r, double* __restrict__ scal,
                     double* partials, unsigned* counter, double* rr_out, const MgdCtx* rctx = nullptr,
                     unsigned long long* rseq = nullptr) {
  if (mgs_stopped(scal)) return;
  const double rz = scal[MGS_RZ], pq = scal[MGS_PQ];
  const double alpha = rz / pq;
  double acc = 0.0;
  const long long n2 = n >> 1;
  const double2* p2 = (const double2*)p;
  const double2* q2 = (const double2*)q;
  double2* x2 = (double2*)x;
  double2* r2 = (double2*)r;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const double2 pv = p2[i], qv = q2[i], xv = x2[i], rv = r2[i];
    x2[i] = make_double2(fma(alpha, pv.x, xv.x), fma(alpha, pv.y, xv.y));
    const double2 rn = make_double2(fma(-alpha, qv.x, rv.x), fma(-alpha, qv.y, rv.y));
    r2[i] = rn;
    acc = fma(rn.x, rn.x, acc);
    acc = fma(rn.y, rn.y, acc);
  }
  double total;
  if (grid_sum(acc, partials, counter, total)) {
    if (threadIdx.x == 0) {
      scal[MGS_RZ_OLD] = rz;
      scal[MGS_ITER] += 1.0;
      if (!(pq > 0.0 && rz > 0.0)) total = __longlong_as_double(0x7ff8000000000000LL);
    }
    // (after the bookkeeping: in the distributed solve the store waits for the other ranks)
    mgs_store_dot_block(total, rr_out, rctx, rseq);
  }
}
// scal[MGS_TOL] = rtol^2 |b|^2 and the iteration counter (start of a solve; one thread)
__global__ void mg_pcg_begin_kernel(double* scal, double rtol2, int reset_iter) {
  scal[MGS_TOL] = rtol2 * scal[MGS_BB];
  if (reset_iter) scal[MGS_ITER] = 0.0;
}
// FP64 column-major blocks a[36 s + 6 j + i] -> FP32 row-pair-major blocks b[36 s + 12 sub + 2 j + r],
// i = 2 sub + r (the layout bsr_row_product<float> reads with 16-byte loads)
__global__ void mg_to_float_kernel(long long n, const double* __restrict__ a, float* __restrict__ b) {
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
    const long long s = o / 36;
    const int k = (int)(o - 36 * s), sub = k / 12, rem = k - 12 * sub, j = rem >> 1, r = rem & 1;
    b[o] = (float)a[36 * s + 6 * j + 2 * sub + r];
  }
}
// the same layout in binary16 (fine level only: the block-Jacobi-scaled matrix has |entries| <= 1)
__global__ void mg_to_half_kernel(long long n, const double* __restrict__ a, __half* __restrict__ b) {
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
    const long long s = o / 36;
    const int k = (int)(o - 36 * s), sub = k / 12, rem = k - 12 * sub, j = rem >> 1, r = rem & 1;
    b[o] = __double2half(a[36 * s + 6 * j + 2 * sub + r]);
  }
}
// deterministic pseudo-random start vector in (-1, 1) (integer hash of the index): a constant vector
// is nearly orthogonal to the top of the spectrum on large meshes and the power iteration stalls
__global__ void mg_hash_fill_kernel(long long n, double* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = (unsigned long long)i + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    y[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
  }
}

}  // namespace jsso
