// Block-CSR (6x6, column-major blocks) SpMV and the block-Jacobi PCG kernels.
//
// Replaces the reference's host SuperLU / cuSOLVER sparse direct solves
// (JaxSSO/solver.py:102-125, 176-210).  The reference imposes boundary conditions
// with Lagrange multipliers (assemblemodel.py:111-163), which makes the matrix
// indefinite; with zero prescribed displacements (assemblemodel.py:192) the same u
// solves the reduced SPD system, which is what CG runs on (prescribed rows/cols are
// identity rows).
//
// Block-Jacobi is applied as a symmetric scaling  A^ = W A W^T,  W_r = L_r^-1 with
// D_r = L_r L_r^T the diagonal blocks, so the Krylov loop is plain CG on A^ and the
// preconditioner costs no memory traffic per iteration.
//
// All kernels are HBM-bound streaming kernels; the SpMV moves 288 B per block with
// 16-byte loads, 30 of 32 lanes active (3 lanes x 2 rows per block column, 10 block
// columns per warp-load), all loads of a block row issued before the first FMA, and the
// row loop is software-pipelined (rowptr two rows ahead, colidx one row ahead) so that one
// memory latency per row is exposed instead of three: 0.96 of the measured HBM bandwidth.
#pragma once
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#ifndef JSSO_SPMV_PAIRED
#define JSSO_SPMV_PAIRED 1   // FP32 / binary16 level SpMVs: two rows per warp iteration (bsr_rows_paired)
#endif

namespace jsso {

// Device-resident CG state.  There is no stored "done" flag: every kernel derives
// "stop" from the scalars themselves (r.r <= tol^2 |b|^2, or NaN), so that after
// convergence all later launches of a batch are no-ops and, on several GPUs, the
// dot products can be all-reduced in place between kernels.
struct CgScalars {
  double rr[2];     // ping-pong r.r
  double pq;        // p.Ap
  double bb;        // |b^|^2
  double tol2;      // rtol^2
  double aux;       // scratch dot (true residual)
  // this rank's partial sums [p.Ap, r.r, |b|^2, aux]: on several GPUs the kernels write here and an
  // out-of-place ncclAllReduce(loc -> target) fills the globals, which keeps the reduction
  // idempotent when later launches of a batch are no-ops after convergence
  double loc[4];
  int iter;
  int pad;
};

__device__ inline bool cg_stop(const CgScalars* sc, int cur) {
  return !(sc->rr[cur] > sc->tol2 * sc->bb);   // also true for NaN
}

// ---- NVLink peer-memory collectives for the distributed CG ---------------------------------
// One process per GPU; every rank maps its peers' `p` vector and mailbox through CUDA IPC.
// The halo exchange and the two scalar all-reduces of a CG iteration are then done by the CG
// kernels themselves with plain stores / loads on peer pointers (no NCCL call, no extra launch
// on the critical path except the push kernel):
//   * halo: the owner writes its interface entries of p straight into the ghost slots of the
//     neighbours' p vectors, fences system-wide and raises halo[rank] = seq in their mailboxes;
//     the neighbour's SpMV spins on its own mailbox before touching ghost columns;
//   * all-reduce: the block that finalises a dot product stores (value, tag = seq) into slot
//     [channel][rank] of EVERY rank's mailbox; the consumer kernel's blocks spin until all
//     n_rank tags equal seq and sum the values in rank order (bitwise identical on all ranks).
// Tags only grow, so no reset or barrier is needed; see DESIGN.md section 5 for the ordering
// argument (a rank cannot reach reduction k+1 before every rank has read reduction k).
constexpr int P2P_MAX_RANKS = 16;
struct Mailbox {
  double val[2][P2P_MAX_RANKS];
  unsigned long long tag[2][P2P_MAX_RANKS];
  unsigned long long halo[P2P_MAX_RANKS];
};
struct P2PCtx {
  int rank, n_rank, n_peer;
  Mailbox* mbox[P2P_MAX_RANKS];      // by rank (own entry = local mailbox)
  int peer_rank[P2P_MAX_RANKS];      // by peer index
  double* peer_vec[P2P_MAX_RANKS];   // peer's p vector
  int send_off[P2P_MAX_RANKS], send_cnt[P2P_MAX_RANKS], remote_start[P2P_MAX_RANKS];
};

// system-scope release store / acquire load (PTX memory model): the release orders every write
// this thread has performed or observed (through CTA barriers / gpu-scope atomics) before the
// flag; the acquire makes them visible to the loads that follow the successful poll.
#ifdef JSSO_EMU   // CPU test harness (single process: the peer-memory kernels are not exercised there)
__device__ inline void st_release_sys(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ inline unsigned long long ld_acquire_sys(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ inline double ld_relaxed_sys_f64(const double* p) { return *(const volatile double*)p; }
__device__ inline void st_relaxed_sys_f64(double* p, double v) { *(volatile double*)p = v; }
#else
__device__ inline void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ inline unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ inline double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ inline void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
#endif

// called by ONE thread: publish this rank's partial of reduction `seq` on channel ch
__device__ inline void p2p_publish(const P2PCtx* c, int ch, unsigned long long seq, double v) {
  for (int r = 0; r < c->n_rank; ++r) {
    st_relaxed_sys_f64(&c->mbox[r]->val[ch][c->rank], v);
  }
  for (int r = 0; r < c->n_rank; ++r) st_release_sys(&c->mbox[r]->tag[ch][c->rank], seq);
}
// called by ALL threads of a block: wait for reduction `seq` and return the sum (rank order)
__device__ inline double p2p_reduce(const P2PCtx* c, int ch, unsigned long long seq) {
  __shared__ double total;
  if (threadIdx.x == 0) {
    const Mailbox* m = c->mbox[c->rank];
    double s = 0.0;
    for (int r = 0; r < c->n_rank; ++r) {
      while (ld_acquire_sys(&m->tag[ch][r]) < seq) { }
      s += ld_relaxed_sys_f64(&m->val[ch][r]);
    }
    total = s;
  }
  __syncthreads();
  return total;
}
// called by ALL threads of a block: wait until every neighbour has pushed halo `seq`
__device__ inline void p2p_wait_halo(const P2PCtx* c, unsigned long long seq) {
  if ((int)threadIdx.x < c->n_peer) {
    const Mailbox* m = c->mbox[c->rank];
    while (ld_acquire_sys(&m->halo[c->peer_rank[threadIdx.x]]) < seq) { }
  }
  __syncthreads();
}

constexpr int RED_BLOCK = 256;
constexpr int RED_MAX_BLOCKS = 1184;   // 148 SMs x 8

// ---- fused multigrid-PCG iteration (device-resident scalars) -------------------------------------------
// Slots of the scalar block `scal` (doubles): the outer PCG keeps every scalar on the device, so an iteration
// needs no host round trip.  A kernel that gets a `stop` pointer (= scal) returns at once when the recurrence
// residual has reached the tolerance (r.r <= tol2 |b|^2, also true for NaN), so the launches the host enqueued
// past convergence (it polls every few iterations) are no-ops.  On several GPUs every dot product is written to
// its MGS_LOC + slot partial and summed over the ranks OUT OF PLACE into the slot itself (idempotent when the
// producing kernel was a no-op).
constexpr int MGS_BB = 1, MGS_RR = 2, MGS_RZ = 3, MGS_PQ = 4, MGS_RZ_OLD = 6, MGS_TOL = 7, MGS_ITER = 8,
              MGS_LOC = 16, MGS_COUNT = 32;
__device__ inline bool mgs_stopped(const double* stop) {
  return stop && !(stop[MGS_RR] > stop[MGS_TOL]);
}

struct MgdCtx;
// Store a finished dot product: on one GPU straight into *out; in the peer-memory distributed solve (ctx != null) publish
// this rank's partial to every rank's mailbox, wait for all ranks and store the sum (rank order: bitwise identical
// everywhere).  Called by ONE thread of the block that finalised the sum; defined below the mailbox structs.
__device__ inline void mgs_store_dot(double total, double* out, const MgdCtx* ctx, unsigned long long* seq_ctr);
// the same called by ALL threads of that block (total valid in thread 0): lane r of the first warp talks to rank r, so
// the remote stores, the fences and the polls of the ranks run in parallel instead of one after the other
__device__ inline void mgs_store_dot_block(double total, double* out, const MgdCtx* ctx, unsigned long long* seq_ctr);

__device__ inline double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid reduction: per-block partials, the last block to arrive sums
// them in a fixed order.  Returns true in the threads of the last block, with the
// total in `total` (thread 0 only).
__device__ inline bool grid_sum(double v, double* partials, unsigned* counter, double& total) {
  __shared__ double ws[RED_BLOCK / 32];
  __shared__ bool last;
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += ws[i];
    partials[blockIdx.x] = s;
    __threadfence();
    const unsigned t = atomicInc(counter, gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  double s = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += ((volatile double*)partials)[i];
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += ws[i];
    total = t;
  }
  return true;
}

// ---- SpMV ------------------------------------------------------------------------
// One block row times x by one warp: lanes 0..2 return rows (2 lane, 2 lane + 1) in (u0, u1).
// Both row products take the row's block range [b0, b1) and the column indices of its FIRST chunk
// (RowU<VT>::U per lane, -1 = no work) in registers, so that the kernels can load them one row ahead
// (software pipeline: rowptr two rows ahead, colidx one row ahead); longer rows continue with direct loads.
template <class VT> struct RowU;
template <> struct RowU<double> { static constexpr int U = 6; };   // 10 blocks per chunk: 30 lanes x 6 columns
template <> struct RowU<float> { static constexpr int U = 3; };    // 9 blocks per chunk: 27 lanes x 3 blocks
template <> struct RowU<__half> { static constexpr int U = 3; };   // same lane mapping as float, 8-byte loads

__device__ inline void row_colidx(int lane, int b0, int b1, const int32_t* __restrict__ colidx, int (&ci)[6], const double*) {
  const int cg = lane / 3, ncol = 6 * (b1 - b0);
#pragma unroll
  for (int u = 0; u < 6; ++u) {
    const int c = 10 * u + cg;
    ci[u] = ((cg < 10) && (c < ncol)) ? colidx[b0 + c / 6] : -1;
  }
}
__device__ inline void row_colidx(int lane, int b0, int b1, const int32_t* __restrict__ colidx, int (&ci)[3], const float*) {
  const int kb = lane / 9;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int k = b0 + 3 * u + kb;
    ci[u] = ((lane < 27) && (k < b1)) ? colidx[k] : -1;
  }
}

__device__ inline void row_colidx(int lane, int b0, int b1, const int32_t* __restrict__ colidx, int (&ci)[3], const __half*) {
  row_colidx(lane, b0, b1, colidx, ci, (const float*)nullptr);
}

// FP64 blocks, column-major: lane = 3 cg + sub handles rows (2 sub, 2 sub + 1) of columns 10 u + cg;
// lanes 0..2 return rows (2 lane, 2 lane + 1) in (u0, u1).
__device__ inline void bsr_row_product(int b0, int b1, const int (&ci)[6], int lane,
                                       const int32_t* __restrict__ colidx, const double* __restrict__ vals,
                                       const double* __restrict__ x, double& u0, double& u1) {
  const int sub = lane % 3, cg = lane / 3;   // rows (2 sub, 2 sub + 1), column group 0..9 (10 = idle)
  const int ncol = 6 * (b1 - b0);
  const double* base = vals + (size_t)b0 * 36 + 2 * sub;
  double acc0 = 0.0, acc1 = 0.0;
  {
    double2 a[6];
    double xv[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int c = 10 * u + cg;
      const bool ok = ci[u] >= 0;
      a[u] = ok ? __ldg((const double2*)(base + (size_t)c * 6)) : make_double2(0.0, 0.0);
      xv[u] = ok ? x[6 * (size_t)ci[u] + c % 6] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      acc0 = fma(a[u].x, xv[u], acc0);
      acc1 = fma(a[u].y, xv[u], acc1);
    }
  }
  for (int c0 = 60; c0 < ncol; c0 += 60) {   // rows with more than 10 blocks
    double2 a[6];
    double xv[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int c = c0 + 10 * u + cg;
      const bool ok = (cg < 10) && (c < ncol);
      a[u] = ok ? __ldg((const double2*)(base + (size_t)c * 6)) : make_double2(0.0, 0.0);
      xv[u] = ok ? x[6 * (size_t)colidx[b0 + c / 6] + c % 6] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      acc0 = fma(a[u].x, xv[u], acc0);
      acc1 = fma(a[u].y, xv[u], acc1);
    }
  }
  // sum the 10 column groups: lanes l, l+3, ..., l+27 -> lanes 0..2
  const double s0 = acc0 + __shfl_down_sync(0xffffffffu, acc0, 15);
  const double s1 = acc1 + __shfl_down_sync(0xffffffffu, acc1, 15);
  const double t0 = s0 + __shfl_down_sync(0xffffffffu, s0, 6);
  const double t1 = s1 + __shfl_down_sync(0xffffffffu, s1, 6);
  u0 = t0 + __shfl_down_sync(0xffffffffu, t0, 3);
  u1 = t1 + __shfl_down_sync(0xffffffffu, t1, 3);
  u0 += __shfl_down_sync(0xffffffffu, s0, 12);
  u1 += __shfl_down_sync(0xffffffffu, s1, 12);
}

// FP32 block storage of the multigrid preconditioner (half the HBM traffic of the V-cycle; accumulation,
// vectors and the outer PCG stay FP64).  Blocks are stored ROW-PAIR major,
//   f[12 sub + 2 j + r] = A[2 sub + r][j]   (sub = 0..2, j = 0..5, r = 0..1),
// so that the 2 x 2 sub-block (rows 2 sub, 2 sub + 1; columns 2 cp, 2 cp + 1) is one aligned 16-byte load.
// lane = 9 kb + 3 cp + sub handles that sub-block of block kb of the current triple of blocks (27 lanes,
// three 16-byte loads of A and three of x in flight); lanes 0..2 return rows (2 lane, 2 lane + 1).
__device__ inline void bsr_row_product(int b0, int b1, const int (&ci)[3], int lane,
                                       const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                       const double* __restrict__ x, double& u0, double& u1) {
  const int kb = lane / 9, rem = lane - 9 * kb, cp = rem / 3, sub = rem - 3 * cp;
  const float4* base = (const float4*)(vals + (size_t)b0 * 36) + 3 * sub + cp + 9 * kb;   // + 27 per step
  double acc0 = 0.0, acc1 = 0.0;
  {
    float4 a[3];
    double2 xv[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const bool ok = ci[u] >= 0;
      a[u] = ok ? __ldg(base + 27 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
      xv[u] = ok ? *(const double2*)(x + 6 * (size_t)ci[u] + 2 * cp) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      acc0 = fma((double)a[u].x, xv[u].x, acc0); acc0 = fma((double)a[u].z, xv[u].y, acc0);
      acc1 = fma((double)a[u].y, xv[u].x, acc1); acc1 = fma((double)a[u].w, xv[u].y, acc1);
    }
  }
  for (int k0 = b0 + 9; k0 < b1; k0 += 9) {   // rows with more than 9 blocks
    float4 a[3];
    double2 xv[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int k = k0 + 3 * u + kb;
      const bool ok = (lane < 27) && (k < b1);
      a[u] = ok ? __ldg(base + 9 * (size_t)(k0 - b0) + 27 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
      xv[u] = ok ? *(const double2*)(x + 6 * (size_t)colidx[k] + 2 * cp) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      acc0 = fma((double)a[u].x, xv[u].x, acc0); acc0 = fma((double)a[u].z, xv[u].y, acc0);
      acc1 = fma((double)a[u].y, xv[u].x, acc1); acc1 = fma((double)a[u].w, xv[u].y, acc1);
    }
  }
  // sum the 9 lanes of every sub: lanes l, l+3, ..., l+24 -> lanes 0..2
  const double s0 = acc0 + __shfl_down_sync(0xffffffffu, acc0, 12);
  const double s1 = acc1 + __shfl_down_sync(0xffffffffu, acc1, 12);
  const double t0 = s0 + __shfl_down_sync(0xffffffffu, s0, 6);
  const double t1 = s1 + __shfl_down_sync(0xffffffffu, s1, 6);
  u0 = t0 + __shfl_down_sync(0xffffffffu, t0, 3);
  u1 = t1 + __shfl_down_sync(0xffffffffu, t1, 3);
  u0 += __shfl_down_sync(0xffffffffu, acc0, 24);
  u1 += __shfl_down_sync(0xffffffffu, acc1, 24);
}

// FP16 block storage of the FINE level of the multigrid preconditioner (opt-in, JSSO_MG_FP16=1).  The fine matrix
// is the block-Jacobi-scaled one (unit diagonal blocks, |entries| <= 1), so binary16 holds it without overflow;
// a CPU study (same V-cycle, level-0 matrix rounded to binary16, coarse levels binary32) keeps the PCG iteration
// count (96^2: 83 -> 84 at Chebyshev-1, 56 -> 56 at Chebyshev-2) while bfloat16 loses 15-25 %.  Same row-pair-major
// layout and lane mapping as the FP32 path; the 2 x 2 sub-block is one aligned 8-byte load (block = 72 bytes).
__device__ inline void half4_to_float(const uint2 v, float& x, float& y, float& z, float& w) {
  const __half2 lo = *reinterpret_cast<const __half2*>(&v.x), hi = *reinterpret_cast<const __half2*>(&v.y);
  const float2 a = __half22float2(lo), b = __half22float2(hi);
  x = a.x; y = a.y; z = b.x; w = b.y;
}
__device__ inline void bsr_row_product(int b0, int b1, const int (&ci)[3], int lane,
                                       const int32_t* __restrict__ colidx, const __half* __restrict__ vals,
                                       const double* __restrict__ x, double& u0, double& u1) {
  const int kb = lane / 9, rem = lane - 9 * kb, cp = rem / 3, sub = rem - 3 * cp;
  const uint2* base = (const uint2*)(vals + (size_t)b0 * 36) + 3 * sub + cp + 9 * kb;   // + 27 per step
  double acc0 = 0.0, acc1 = 0.0;
  {
    uint2 a[3];
    double2 xv[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const bool ok = ci[u] >= 0;
      a[u] = ok ? __ldg(base + 27 * u) : make_uint2(0u, 0u);
      xv[u] = ok ? *(const double2*)(x + 6 * (size_t)ci[u] + 2 * cp) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      float ax, ay, az, aw;
      half4_to_float(a[u], ax, ay, az, aw);
      acc0 = fma((double)ax, xv[u].x, acc0); acc0 = fma((double)az, xv[u].y, acc0);
      acc1 = fma((double)ay, xv[u].x, acc1); acc1 = fma((double)aw, xv[u].y, acc1);
    }
  }
  for (int k0 = b0 + 9; k0 < b1; k0 += 9) {   // rows with more than 9 blocks
    uint2 a[3];
    double2 xv[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int k = k0 + 3 * u + kb;
      const bool ok = (lane < 27) && (k < b1);
      a[u] = ok ? __ldg(base + 9 * (size_t)(k0 - b0) + 27 * u) : make_uint2(0u, 0u);
      xv[u] = ok ? *(const double2*)(x + 6 * (size_t)colidx[k] + 2 * cp) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      float ax, ay, az, aw;
      half4_to_float(a[u], ax, ay, az, aw);
      acc0 = fma((double)ax, xv[u].x, acc0); acc0 = fma((double)az, xv[u].y, acc0);
      acc1 = fma((double)ay, xv[u].x, acc1); acc1 = fma((double)aw, xv[u].y, acc1);
    }
  }
  const double s0 = acc0 + __shfl_down_sync(0xffffffffu, acc0, 12);
  const double s1 = acc1 + __shfl_down_sync(0xffffffffu, acc1, 12);
  const double t0 = s0 + __shfl_down_sync(0xffffffffu, s0, 6);
  const double t1 = s1 + __shfl_down_sync(0xffffffffu, s1, 6);
  u0 = t0 + __shfl_down_sync(0xffffffffu, t0, 3);
  u1 = t1 + __shfl_down_sync(0xffffffffu, t1, 3);
  u0 += __shfl_down_sync(0xffffffffu, acc0, 24);
  u1 += __shfl_down_sync(0xffffffffu, acc1, 24);
}

// un-pipelined form (one row, everything loaded here)
template <class VT>
__device__ inline void bsr_row_product(int r, int lane, const int32_t* __restrict__ rowptr,
                                       const int32_t* __restrict__ colidx, const VT* __restrict__ vals,
                                       const double* __restrict__ x, double& u0, double& u1) {
  const int b0 = rowptr[r], b1 = rowptr[r + 1];
  int ci[RowU<VT>::U];
  row_colidx(lane, b0, b1, colidx, ci, vals);
  bsr_row_product(b0, b1, ci, lane, colidx, vals, x, u0, u1);
}

// Software-pipelined row loop of a warp: rows r, r + n_warp, ...; `body(r, u0, u1)` consumes a row's
// product (lanes 0..2).  rowptr is loaded two rows ahead and colidx one row ahead, so that per row only
// the matrix / x loads are on the critical path (one memory latency instead of three).
// A body that returns double has its values summed per lane and returned (dot products fused into an epilogue).
template <class VT, class Body>
__device__ inline double bsr_rows_pipelined(int warp, int n_warp, int lane, int n_row,
                                          const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                          const VT* __restrict__ vals, const double* __restrict__ x, Body body) {
  constexpr int U = RowU<VT>::U;
  int r = warp;
  double acc = 0.0;
  if (r >= n_row) return acc;
  int b0 = rowptr[r], b1 = rowptr[r + 1];
  int ci[U];
  row_colidx(lane, b0, b1, colidx, ci, vals);
  int nb0 = 0, nb1 = 0;
  if (r + n_warp < n_row) { nb0 = rowptr[r + n_warp]; nb1 = rowptr[r + n_warp + 1]; }
  for (; r < n_row; r += n_warp) {
    const int r1 = r + n_warp, r2 = r + 2 * n_warp;
    int nci[U];
#pragma unroll
    for (int u = 0; u < U; ++u) nci[u] = -1;
    if (r1 < n_row) row_colidx(lane, nb0, nb1, colidx, nci, vals);   // nb0/nb1 were loaded one iteration ago
    int nnb0 = 0, nnb1 = 0;
    if (r2 < n_row) { nnb0 = rowptr[r2]; nnb1 = rowptr[r2 + 1]; }
    double u0, u1;
    bsr_row_product(b0, b1, ci, lane, colidx, vals, x, u0, u1);
    if constexpr (std::is_void<decltype(body(r, u0, u1))>::value) body(r, u0, u1);
    else acc += body(r, u0, u1);
    b0 = nb0; b1 = nb1; nb0 = nnb0; nb1 = nnb1;
#pragma unroll
    for (int u = 0; u < U; ++u) ci[u] = nci[u];
  }
  return acc;
}

// ---- two rows per warp iteration (FP32 / binary16 block storage) --------------------------------------------
// With half or a quarter of the bytes per row the one-row pipeline above is bound by memory latency, not bandwidth
// (ncu, 1M quads: long-scoreboard stalls 12 warps per issue, 4.4 TB/s): the loads of TWO rows (r and r + n_warp)
// are issued before the first FMA, which doubles the bytes in flight per warp at the same occupancy.
template <class VT> struct RowRaw;
template <> struct RowRaw<float> { float4 a[3]; double2 xv[3]; };
template <> struct RowRaw<__half> { uint2 a[3]; double2 xv[3]; };
__device__ inline void row_issue(int b0, const int (&ci)[3], int lane, const float* __restrict__ vals,
                                 const double* __restrict__ x, RowRaw<float>& q) {
  const int kb = lane / 9, rem = lane - 9 * kb, cp = rem / 3, sub = rem - 3 * cp;
  const float4* base = (const float4*)(vals + (size_t)b0 * 36) + 3 * sub + cp + 9 * kb;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const bool ok = ci[u] >= 0;
    q.a[u] = ok ? __ldg(base + 27 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
    q.xv[u] = ok ? *(const double2*)(x + 6 * (size_t)ci[u] + 2 * cp) : make_double2(0.0, 0.0);
  }
}
__device__ inline void row_issue(int b0, const int (&ci)[3], int lane, const __half* __restrict__ vals,
                                 const double* __restrict__ x, RowRaw<__half>& q) {
  const int kb = lane / 9, rem = lane - 9 * kb, cp = rem / 3, sub = rem - 3 * cp;
  const uint2* base = (const uint2*)(vals + (size_t)b0 * 36) + 3 * sub + cp + 9 * kb;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const bool ok = ci[u] >= 0;
    q.a[u] = ok ? __ldg(base + 27 * u) : make_uint2(0u, 0u);
    q.xv[u] = ok ? *(const double2*)(x + 6 * (size_t)ci[u] + 2 * cp) : make_double2(0.0, 0.0);
  }
}
__device__ inline float4 row_block_f32(const float4 v) { return v; }
__device__ inline float4 row_block_f32(const uint2 v) {
  float4 f;
  half4_to_float(v, f.x, f.y, f.z, f.w);
  return f;
}
template <class VT>
__device__ inline void row_finish(const RowRaw<VT>& q, double& u0, double& u1) {
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const float4 a = row_block_f32(q.a[u]);
    acc0 = fma((double)a.x, q.xv[u].x, acc0); acc0 = fma((double)a.z, q.xv[u].y, acc0);
    acc1 = fma((double)a.y, q.xv[u].x, acc1); acc1 = fma((double)a.w, q.xv[u].y, acc1);
  }
  const double s0 = acc0 + __shfl_down_sync(0xffffffffu, acc0, 12);
  const double s1 = acc1 + __shfl_down_sync(0xffffffffu, acc1, 12);
  const double t0 = s0 + __shfl_down_sync(0xffffffffu, s0, 6);
  const double t1 = s1 + __shfl_down_sync(0xffffffffu, s1, 6);
  u0 = t0 + __shfl_down_sync(0xffffffffu, t0, 3);
  u1 = t1 + __shfl_down_sync(0xffffffffu, t1, 3);
  u0 += __shfl_down_sync(0xffffffffu, acc0, 24);
  u1 += __shfl_down_sync(0xffffffffu, acc1, 24);
}
template <class VT> struct RowPair { static constexpr bool value = false; static constexpr int minb = 4; };
// FP32 blocks: two rows in flight need ~100 registers (144 bytes of spills at 80); the coarse levels that use them are
// 1/9 of the work, so they keep the one-row pipeline.  binary16 (the fine level) fits in 64 registers.
template <> struct RowPair<float> { static constexpr bool value = false; static constexpr int minb = 4; };
template <> struct RowPair<__half> { static constexpr bool value = true; static constexpr int minb = 4; };

// rows r, r + n_warp of a warp per iteration (rows with more than 9 blocks fall back to the one-row product)
template <class VT, class Body>
__device__ inline double bsr_rows_paired(int warp, int n_warp, int lane, int n_row, const int32_t* __restrict__ rowptr,
                                         const int32_t* __restrict__ colidx, const VT* __restrict__ vals,
                                         const double* __restrict__ x, Body body) {
  double acc = 0.0;
  const int step = 2 * n_warp;
  int r = warp;
  if (r >= n_row) return acc;
  int a0 = rowptr[r], a1 = rowptr[r + 1], b0 = 0, b1 = 0;
  if (r + n_warp < n_row) { b0 = rowptr[r + n_warp]; b1 = rowptr[r + n_warp + 1]; }
  int ciA[3], ciB[3];
  row_colidx(lane, a0, a1, colidx, ciA, vals);
  row_colidx(lane, b0, b1, colidx, ciB, vals);
  int na0 = 0, na1 = 0, nb0 = 0, nb1 = 0;
  if (r + step < n_row) { na0 = rowptr[r + step]; na1 = rowptr[r + step + 1]; }
  if (r + step + n_warp < n_row) { nb0 = rowptr[r + step + n_warp]; nb1 = rowptr[r + step + n_warp + 1]; }
  for (; r < n_row; r += step) {
    RowRaw<VT> qa, qb;
    row_issue(a0, ciA, lane, vals, x, qa);
    row_issue(b0, ciB, lane, vals, x, qb);
    int nciA[3], nciB[3];
    row_colidx(lane, na0, na1, colidx, nciA, vals);   // empty range (past the end) gives -1 everywhere
    row_colidx(lane, nb0, nb1, colidx, nciB, vals);
    int nna0 = 0, nna1 = 0, nnb0 = 0, nnb1 = 0;
    const int r2 = r + 2 * step;
    if (r2 < n_row) { nna0 = rowptr[r2]; nna1 = rowptr[r2 + 1]; }
    if (r2 + n_warp < n_row) { nnb0 = rowptr[r2 + n_warp]; nnb1 = rowptr[r2 + n_warp + 1]; }
    double ua0, ua1, ub0, ub1;
    if (a1 - a0 <= 9 && b1 - b0 <= 9) {
      row_finish(qa, ua0, ua1);
      row_finish(qb, ub0, ub1);
    } else {   // long rows (coarse levels): the one-row product re-reads its first chunk
      bsr_row_product(a0, a1, ciA, lane, colidx, vals, x, ua0, ua1);
      bsr_row_product(b0, b1, ciB, lane, colidx, vals, x, ub0, ub1);
    }
    if constexpr (std::is_void<decltype(body(r, ua0, ua1))>::value) {
      body(r, ua0, ua1);
      if (r + n_warp < n_row) body(r + n_warp, ub0, ub1);
    } else {
      acc += body(r, ua0, ua1);
      if (r + n_warp < n_row) acc += body(r + n_warp, ub0, ub1);
    }
    a0 = na0; a1 = na1; b0 = nb0; b1 = nb1; na0 = nna0; na1 = nna1; nb0 = nnb0; nb1 = nnb1;
#pragma unroll
    for (int u = 0; u < 3; ++u) { ciA[u] = nciA[u]; ciB[u] = nciB[u]; }
  }
  return acc;
}

// MODE 0: y = A x.            MODE 1: y = A x and sc->pq = x.y (CG step).
// One warp per block row, persistent grid-stride over rows.
template <int MODE>
__global__ void __launch_bounds__(RED_BLOCK)
bsr_spmv_kernel(int n_row, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                CgScalars* sc, int cur, double* partials, unsigned* counter, int single_gpu,
                const P2PCtx* p2p, unsigned long long halo_seq, unsigned long long red_seq) {
  if (MODE == 1 && cg_stop(sc, cur)) {
    // latch the stop state into the other parity slot (nobody reads it in this kernel)
    if (p2p && blockIdx.x == 0 && threadIdx.x == 0) sc->rr[cur ^ 1] = sc->rr[cur];
    return;
  }
  if (p2p) p2p_wait_halo(p2p, halo_seq);
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warp = (gridDim.x * blockDim.x) >> 5;
  double dot = 0.0;
  bsr_rows_pipelined(warp, n_warp, lane, n_row, rowptr, colidx, vals, x, [&](int r, double u0, double u1) {
    if (lane < 3) {
      *(double2*)(y + 6 * (size_t)r + 2 * lane) = make_double2(u0, u1);
      if (MODE == 1) {
        const double2 xr = *(const double2*)(x + 6 * (size_t)r + 2 * lane);
        dot += u0 * xr.x + u1 * xr.y;
      }
    }
  });
  if (MODE == 1) {
    double total;
    if (grid_sum(dot, partials, counter, total) && threadIdx.x == 0) {
      sc->loc[0] = total;
      if (single_gpu) sc->pq = total;
      if (p2p) p2p_publish(p2p, 0, red_seq, total);
    }
  }
}

// q = A p and *dot_out = p.q on the rows [0, n_row) (outer product of the multigrid PCG; the plain FP64 pipeline of
// bsr_spmv_kernel without the linear epilogue's extra operands: 60 registers, no spills)
__global__ void __launch_bounds__(RED_BLOCK, 4)
bsr_spmv_dot_kernel(int n_row, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                    const double* __restrict__ vals, const double* __restrict__ x, const double* __restrict__ xrow,
                    double* __restrict__ y, const double* stop, double* partials, unsigned* counter, double* dot_out,
                    const MgdCtx* rctx = nullptr, unsigned long long* rseq = nullptr) {
  if (mgs_stopped(stop)) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warp = (gridDim.x * blockDim.x) >> 5;
  // the row's own x entries (for p.q) are loaded BEFORE the row product is reduced, so that their latency hides behind
  // the product instead of extending every row's critical path (the first version of this kernel loaded them in the
  // epilogue: 518 us against 466 us of the plain SpMV at 1M quads)
  constexpr int U = RowU<double>::U;
  double dot = 0.0;
  int r = warp;
  if (r < n_row) {
    int b0 = rowptr[r], b1 = rowptr[r + 1];
    int ci[U];
    row_colidx(lane, b0, b1, colidx, ci, vals);
    int nb0 = 0, nb1 = 0;
    if (r + n_warp < n_row) { nb0 = rowptr[r + n_warp]; nb1 = rowptr[r + n_warp + 1]; }
    for (; r < n_row; r += n_warp) {
      const int r1 = r + n_warp, r2 = r + 2 * n_warp;
      int nci[U];
#pragma unroll
      for (int u = 0; u < U; ++u) nci[u] = -1;
      if (r1 < n_row) row_colidx(lane, nb0, nb1, colidx, nci, vals);
      int nnb0 = 0, nnb1 = 0;
      if (r2 < n_row) { nnb0 = rowptr[r2]; nnb1 = rowptr[r2 + 1]; }
      double2 xr = make_double2(0.0, 0.0);
      if (lane < 3) xr = *(const double2*)(xrow + 6 * (size_t)r + 2 * lane);
      double u0, u1;
      bsr_row_product(b0, b1, ci, lane, colidx, vals, x, u0, u1);
      if (lane < 3) {
        *(double2*)(y + 6 * (size_t)r + 2 * lane) = make_double2(u0, u1);
        dot += u0 * xr.x + u1 * xr.y;
      }
      b0 = nb0; b1 = nb1; nb0 = nnb0; nb1 = nnb1;
#pragma unroll
      for (int u = 0; u < U; ++u) ci[u] = nci[u];
    }
  }
  double total;
  if (grid_sum(dot, partials, counter, total)) mgs_store_dot_block(total, dot_out, rctx, rseq);
}

// MODE 0: y = A x;  2: y = b - A x;  3: y += A x;  5 (short-row kernel only): y = s b + A x
// (any 6x6 block-CSR, also rectangular)
#ifndef JSSO_SPMV_MINB
#define JSSO_SPMV_MINB 4   // resident CTAs per SM the register allocation of the level SpMV aims at
#endif

template <int MODE, class VT>
__global__ void __launch_bounds__(RED_BLOCK, (JSSO_SPMV_PAIRED ? RowPair<VT>::minb : JSSO_SPMV_MINB))
bsr_spmv_axpby_kernel(int n_row, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                      const VT* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                      const double* __restrict__ bvec) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warp = (gridDim.x * blockDim.x) >> 5;
  auto epilogue = [=](int r, double u0, double u1) {
    if (lane < 3) {
      double2* yp = (double2*)(y + 6 * (size_t)r + 2 * lane);
      if (MODE == 2) {
        const double2 bv = *(const double2*)(bvec + 6 * (size_t)r + 2 * lane);
        *yp = make_double2(bv.x - u0, bv.y - u1);
      } else if (MODE == 3) {
        const double2 yv = *yp;
        *yp = make_double2(yv.x + u0, yv.y + u1);
      } else {
        *yp = make_double2(u0, u1);
      }
    }
  };
  if constexpr (RowPair<VT>::value && JSSO_SPMV_PAIRED) bsr_rows_paired(warp, n_warp, lane, n_row, rowptr, colidx, vals, x, epilogue);
  else bsr_rows_pipelined(warp, n_warp, lane, n_row, rowptr, colidx, vals, x, epilogue);
}

// y_r = ca * bvec_r + cb * xrow_r + cc * (A x)_r  for the block rows [0, n_row) of (rowptr, y, bvec, xrow);
// x is the full-length gather vector.  bvec / xrow may be null when their coefficient is 0.
// DOT 1: *dot_out = sum_r bvec_r . y_r;  DOT 2: *dot_out = sum_r xrow_r . (A x)_r   (deterministic grid sum).
// Covers the fused V-cycle steps of the fine level -- r0 = b - (1/theta) A b (pre-smoother from a zero guess
// folded into the residual), z = x + (1/theta)(b - A x) with r.z (post-smoother + the PCG's dot) -- and the outer
// q = A p with p.q.
template <class VT, int DOT>
__global__ void __launch_bounds__(RED_BLOCK, (JSSO_SPMV_PAIRED ? RowPair<VT>::minb : JSSO_SPMV_MINB))
bsr_spmv_lin_kernel(int n_row, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                    const VT* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                    const double* __restrict__ bvec, const double* __restrict__ xrow, double ca, double cb, double cc,
                    const double* stop, double* partials, unsigned* counter, double* dot_out,
                    const MgdCtx* rctx = nullptr, unsigned long long* rseq = nullptr) {
  if (mgs_stopped(stop)) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warp = (gridDim.x * blockDim.x) >> 5;
  auto epilogue = [=](int r, double u0, double u1) -> double {
    double d = 0.0;
    if (lane < 3) {
      const size_t o = 6 * (size_t)r + 2 * lane;
      double2 v = make_double2(cc * u0, cc * u1);
      double2 bv = make_double2(0.0, 0.0), xv = make_double2(0.0, 0.0);
      if (bvec) { bv = *(const double2*)(bvec + o); v.x = fma(ca, bv.x, v.x); v.y = fma(ca, bv.y, v.y); }
      if (xrow) { xv = *(const double2*)(xrow + o); v.x = fma(cb, xv.x, v.x); v.y = fma(cb, xv.y, v.y); }
      *(double2*)(y + o) = v;
      if (DOT == 1) d = bv.x * v.x + bv.y * v.y;
      if (DOT == 2) d = xv.x * u0 + xv.y * u1;
    }
    return d;
  };
  double dot;
  if constexpr (RowPair<VT>::value && JSSO_SPMV_PAIRED) dot = bsr_rows_paired(warp, n_warp, lane, n_row, rowptr, colidx, vals, x, epilogue);
  else dot = bsr_rows_pipelined(warp, n_warp, lane, n_row, rowptr, colidx, vals, x, epilogue);
  if (DOT != 0) {
    double total;
    if (grid_sum(dot, partials, counter, total)) mgs_store_dot_block(total, dot_out, rctx, rseq);
  }
}

// ---- row-pair SpMV for FP32 / binary16 block storage (the V-cycle's fine-level products) ---------------------
// One THREAD per (block row, row pair): the thread streams its 2 x 6 slice of every block of the row (12 values:
// three 16-byte loads in FP32, three 8-byte loads in binary16 -- the three threads of a row read 144 / 72
// contiguous bytes per block) and the block's six x entries, and keeps its two row sums in registers: no shuffles,
// no idle lanes, and ~36 instructions per (thread, block) instead of the ~200 warp instructions per row of the
// warp-per-row kernels above, which at half / a quarter of the FP64 bytes are bound by instruction issue and the LSU
// pipe, not by HBM (ncu at 1M quads, binary16: issue active 62 %, LSU 48 %, 2.5 TB/s).  Two blocks per loop
// iteration are issued before the first FMA.  Same epilogue and dot options as bsr_spmv_lin_kernel; persistent
// grid-stride over the (row, pair) items.
template <class VT> struct RpLoad;
template <> struct RpLoad<float> {
  struct Raw { float4 a, b, c; };
  __device__ static inline Raw ld(const float* vals, size_t blk, int sub) {
    const float4* p = (const float4*)(vals + blk * 36) + 3 * sub;
    return Raw{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
  }
  __device__ static inline void unpack(const Raw& r, float (&f)[12]) {
    f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w; f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
    f[8] = r.c.x; f[9] = r.c.y; f[10] = r.c.z; f[11] = r.c.w;
  }
};
template <> struct RpLoad<__half> {
  struct Raw { uint2 a, b, c; };
  __device__ static inline Raw ld(const __half* vals, size_t blk, int sub) {
    const uint2* p = (const uint2*)(vals + blk * 36) + 3 * sub;
    return Raw{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
  }
  __device__ static inline void unpack(const Raw& r, float (&f)[12]) {
    half4_to_float(r.a, f[0], f[1], f[2], f[3]);
    half4_to_float(r.b, f[4], f[5], f[6], f[7]);
    half4_to_float(r.c, f[8], f[9], f[10], f[11]);
  }
};
// row-pair-major block: f[2 j + r] = A[2 sub + r][j]
__device__ inline void rp_fma(const float (&f)[12], const double2 x0, const double2 x1, const double2 x2, double& acc0,
                              double& acc1) {
  acc0 = fma((double)f[0], x0.x, acc0); acc1 = fma((double)f[1], x0.x, acc1);
  acc0 = fma((double)f[2], x0.y, acc0); acc1 = fma((double)f[3], x0.y, acc1);
  acc0 = fma((double)f[4], x1.x, acc0); acc1 = fma((double)f[5], x1.x, acc1);
  acc0 = fma((double)f[6], x1.y, acc0); acc1 = fma((double)f[7], x1.y, acc1);
  acc0 = fma((double)f[8], x2.x, acc0); acc1 = fma((double)f[9], x2.x, acc1);
  acc0 = fma((double)f[10], x2.y, acc0); acc1 = fma((double)f[11], x2.y, acc1);
}
// rows (2 sub, 2 sub + 1) of block row [b0, b1) times x: two blocks issued per round
template <class VT>
__device__ inline void rp_row_pair(int b0, int b1, int sub, const int32_t* __restrict__ colidx, const VT* __restrict__ vals,
                                   const double* __restrict__ x, double& acc0, double& acc1) {
  int k = b0;
  for (; k + 1 < b1; k += 2) {
    const int c0 = colidx[k], c1 = colidx[k + 1];
    const typename RpLoad<VT>::Raw ra = RpLoad<VT>::ld(vals, (size_t)k, sub), rb = RpLoad<VT>::ld(vals, (size_t)k + 1, sub);
    const double2* xa = (const double2*)(x + 6 * (size_t)c0);
    const double2* xb = (const double2*)(x + 6 * (size_t)c1);
    const double2 xa0 = xa[0], xa1 = xa[1], xa2 = xa[2], xb0 = xb[0], xb1 = xb[1], xb2 = xb[2];
    float f[12];
    RpLoad<VT>::unpack(ra, f);
    rp_fma(f, xa0, xa1, xa2, acc0, acc1);
    RpLoad<VT>::unpack(rb, f);
    rp_fma(f, xb0, xb1, xb2, acc0, acc1);
  }
  if (k < b1) {
    const int c0 = colidx[k];
    const typename RpLoad<VT>::Raw ra = RpLoad<VT>::ld(vals, (size_t)k, sub);
    const double2* xa = (const double2*)(x + 6 * (size_t)c0);
    const double2 xa0 = xa[0], xa1 = xa[1], xa2 = xa[2];
    float f[12];
    RpLoad<VT>::unpack(ra, f);
    rp_fma(f, xa0, xa1, xa2, acc0, acc1);
  }
}

template <class VT, int DOT>
__global__ void __launch_bounds__(RED_BLOCK, 3)
bsr_spmv_rp_kernel(int n_row, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                   const VT* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                   const double* __restrict__ bvec, const double* __restrict__ xrow, double ca, double cb, double cc,
                   const double* stop, double* partials, unsigned* counter, double* dot_out,
                   const MgdCtx* rctx = nullptr, unsigned long long* rseq = nullptr,
                   const int32_t* __restrict__ row_list = nullptr /* n_row row ids (then rowptr, y, bvec, xrow are un-offset) */) {
  if (mgs_stopped(stop)) return;
  double dot = 0.0;
  const long long n_item = 3LL * n_row;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_item; t += (long long)gridDim.x * blockDim.x) {
    const int ri = (int)(t / 3), sub = (int)(t - 3LL * ri);
    const int r = row_list ? row_list[ri] : ri;
    const int b0 = rowptr[r], b1 = rowptr[r + 1];
    const size_t o = 6 * (size_t)r + 2 * sub;
    // the epilogue's operands are requested now, their latency hides behind the block loop
    double2 bv = make_double2(0.0, 0.0), xv = make_double2(0.0, 0.0);
    if (bvec) bv = *(const double2*)(bvec + o);
    if (xrow) xv = *(const double2*)(xrow + o);
    double acc0 = 0.0, acc1 = 0.0;
    rp_row_pair(b0, b1, sub, colidx, vals, x, acc0, acc1);
    double2 v = make_double2(cc * acc0, cc * acc1);
    if (bvec) { v.x = fma(ca, bv.x, v.x); v.y = fma(ca, bv.y, v.y); }
    if (xrow) { v.x = fma(cb, xv.x, v.x); v.y = fma(cb, xv.y, v.y); }
    *(double2*)(y + o) = v;
    if (DOT == 1) dot += bv.x * v.x + bv.y * v.y;
    if (DOT == 2) dot += xv.x * acc0 + xv.y * acc1;
  }
  if (DOT != 0) {
    double total;
    if (grid_sum(dot, partials, counter, total)) mgs_store_dot_block(total, dot_out, rctx, rseq);
  }
}

// Short rows (prolongators: 1-4 blocks per row) in the FP32 row-pair-major layout: one THREAD per
// (block row, row pair), no shuffles -- a warp works on ten rows at once instead of one, which is what
// hides the rowptr -> colidx -> x load chain when rows are this short.  MODE as above.
template <int MODE>
__global__ void __launch_bounds__(256, 4)
bsr_spmv_short_kernel(int n_row, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                      const float* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                      const double* __restrict__ bvec, double s = 1.0, const double* stop = nullptr) {
  if (mgs_stopped(stop)) return;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3LL * n_row) return;
  const int r = (int)(t / 3), sub = (int)(t - 3LL * r);
  const int b0 = rowptr[r], b1 = rowptr[r + 1];
  const float4* base = (const float4*)(vals + (size_t)b0 * 36) + 3 * sub;
  double acc0 = 0.0, acc1 = 0.0;
  for (int k = b0; k < b1; ++k, base += 9) {
    const double2* xp = (const double2*)(x + 6 * (size_t)colidx[k]);
    const float4 a0 = __ldg(base), a1 = __ldg(base + 1), a2 = __ldg(base + 2);
    const double2 x0 = xp[0], x1 = xp[1], x2 = xp[2];
    acc0 = fma((double)a0.x, x0.x, acc0); acc0 = fma((double)a0.z, x0.y, acc0);
    acc1 = fma((double)a0.y, x0.x, acc1); acc1 = fma((double)a0.w, x0.y, acc1);
    acc0 = fma((double)a1.x, x1.x, acc0); acc0 = fma((double)a1.z, x1.y, acc0);
    acc1 = fma((double)a1.y, x1.x, acc1); acc1 = fma((double)a1.w, x1.y, acc1);
    acc0 = fma((double)a2.x, x2.x, acc0); acc0 = fma((double)a2.z, x2.y, acc0);
    acc1 = fma((double)a2.y, x2.x, acc1); acc1 = fma((double)a2.w, x2.y, acc1);
  }
  double2* yp = (double2*)(y + 6 * (size_t)r + 2 * sub);
  if (MODE == 2) {
    const double2 bv = *(const double2*)(bvec + 6 * (size_t)r + 2 * sub);
    *yp = make_double2(bv.x - acc0, bv.y - acc1);
  } else if (MODE == 3) {
    const double2 yv = *yp;
    *yp = make_double2(yv.x + acc0, yv.y + acc1);
  } else if (MODE == 5) {   // y = s * bvec + A x  (prolongation onto the lazily formed pre-smoothed iterate b / theta)
    const double2 bv = *(const double2*)(bvec + 6 * (size_t)r + 2 * sub);
    *yp = make_double2(fma(s, bv.x, acc0), fma(s, bv.y, acc1));
  } else {
    *yp = make_double2(acc0, acc1);
  }
}

// push this rank's interface entries of `v` into the neighbours' ghost slots, then raise the flags
__global__ void __launch_bounds__(RED_BLOCK)
p2p_halo_push_kernel(const P2PCtx* c, const int32_t* __restrict__ send_idx, const double* __restrict__ v,
                     unsigned* counter, unsigned long long seq) {
  __shared__ bool last;
  for (int pi = 0; pi < c->n_peer; ++pi) {
    double* dst = c->peer_vec[pi] + 6 * (size_t)c->remote_start[pi];
    const int32_t* idx = send_idx + c->send_off[pi];
    const int n = 6 * c->send_cnt[pi];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
      dst[t] = v[6 * (size_t)idx[t / 6] + t % 6];
  }
  // block barrier, then a gpu-scope release by thread 0 (fence + atomic): the last block to
  // arrive has observed every block's stores and releases them system-wide with the flags
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && (int)threadIdx.x < c->n_peer) {
    __threadfence();
    st_release_sys(&c->mbox[c->peer_rank[threadIdx.x]]->halo[c->rank], seq);
  }
}

// ---- peer-memory halo exchange and scalar all-reduce of the row-range distributed multigrid solve ----------
// (jsso_mg_p2p_connect; the NCCL send/recv path stays as the reference.)  An exchange of level l is two kernels
// and no library call: mgd_push_kernel gathers this rank's entries and stores them straight into the peers'
// receive arenas over NVLink, then raises halo[l][me] = seq in their mailboxes (release); mgd_wait_unpack_kernel
// spins on its own mailbox (acquire) and scatters the arena into the vector's ghost positions.  The arena is
// double-buffered per level by the parity of the level's exchange counter.  Why that suffices: a rank pushes
// exchange k+2 of a level only after its wait of exchange k+1 saw every peer's flag k+1, and a peer raises flag k+1
// (in its push kernel) only after its unpack of exchange k has finished (stream order) -- peers are symmetric and
// raise the flag even when they have nothing to send.  The all-reduce follows the same pattern all-to-all.
constexpr int MGD_MAX_LEVELS = 4;
struct MgdMailbox {
  unsigned long long halo[MGD_MAX_LEVELS][P2P_MAX_RANKS];
  unsigned long long red_tag[2][P2P_MAX_RANKS];
  double red_val[2][P2P_MAX_RANKS][2];
};
struct MgdCtx {
  int rank, n_rank;
  MgdMailbox* mbox[P2P_MAX_RANKS];     // by rank (own entry = local mailbox)
  double* arena[P2P_MAX_RANKS];        // by rank: receive arena, [level][parity][6 * max_recv]
  long long arena_level_stride, arena_slot_stride;   // doubles
};
struct MgdLevelDev {                   // one distributed level, device copy
  int n_peer, n_send, n_recv;
  int peer_rank[P2P_MAX_RANKS], send_off[P2P_MAX_RANKS], send_cnt[P2P_MAX_RANKS], remote_off[P2P_MAX_RANKS];
};

__global__ void __launch_bounds__(RED_BLOCK)
mgd_push_kernel(const MgdCtx* c, const MgdLevelDev* L, int level, const int32_t* __restrict__ send_idx,
                const double* __restrict__ v, unsigned* counter, unsigned long long seq) {
  __shared__ bool last;
  const long long slot = (long long)level * c->arena_level_stride + (long long)(seq & 1ull) * c->arena_slot_stride;
  for (int pi = 0; pi < L->n_peer; ++pi) {
    double* dst = c->arena[L->peer_rank[pi]] + slot + 6 * (long long)L->remote_off[pi];
    const int32_t* idx = send_idx + L->send_off[pi];
    const int n = 6 * L->send_cnt[pi];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
      dst[t] = v[6 * (size_t)idx[t / 6] + t % 6];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && (int)threadIdx.x < L->n_peer) {
    __threadfence();
    st_release_sys(&c->mbox[L->peer_rank[threadIdx.x]]->halo[level][c->rank], seq);
  }
}

__global__ void __launch_bounds__(RED_BLOCK)
mgd_wait_unpack_kernel(const MgdCtx* c, const MgdLevelDev* L, int level, const int32_t* __restrict__ recv_idx,
                       double* __restrict__ v, unsigned long long seq) {
  if ((int)threadIdx.x < L->n_peer) {
    const MgdMailbox* m = c->mbox[c->rank];
    while (ld_acquire_sys(&m->halo[level][L->peer_rank[threadIdx.x]]) < seq) { }
  }
  __syncthreads();
  const double* src = c->arena[c->rank] + (long long)level * c->arena_level_stride +
                      (long long)(seq & 1ull) * c->arena_slot_stride;
  const int n = 6 * L->n_recv;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
    v[6 * (size_t)recv_idx[t / 6] + t % 6] = ld_relaxed_sys_f64(src + t);
}

// A small exchange as ONE kernel of ONE block: push this rank's entries into the peers' arenas and raise the flags, then
// wait for the peers' flags and scatter the own arena into the ghost positions.  Every rank pushes before it waits,
// so the kernels of the ranks cannot wait for each other in a cycle; a single block keeps that true without relying
// on the co-residency of several blocks (a block that waits never holds back a block that still has to push).
// Halos of a strip partition are a few thousand nodes: one block of 1024 threads moves them in a few microseconds
// and saves a launch per exchange; larger exchanges (the all-gather level) use the two-kernel form above.
constexpr int MGD_ONE_BLOCK_MAX = 49152;   // doubles
// 16-byte pieces (a node's six doubles = three pieces), four independent load / store pairs per thread and round:
// a single block is bound by memory latency, not bandwidth, so the rounds must be few (a strip's halo of ~3 700 nodes
// took 17 + 22 dependent rounds = ~30 us with one 8-byte element per thread and round, measured at 8 GPUs)
__device__ inline void mgd_gather_pieces(double* __restrict__ dst, const double* __restrict__ v,
                                         const int32_t* __restrict__ idx, int n_node, int tid, int nthr) {
  const int n = 3 * n_node;
  for (int t0 = tid; t0 < n; t0 += 4 * nthr) {
    double2 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * nthr;
      if (t < n) r[u] = *(const double2*)(v + 6 * (size_t)idx[t / 3] + 2 * (t % 3));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * nthr;
      if (t < n) *(double2*)(dst + 2 * (size_t)t) = r[u];
    }
  }
}
__device__ inline void mgd_scatter_pieces(double* __restrict__ v, const double* __restrict__ src,
                                          const int32_t* __restrict__ idx, int n_node, int tid, int nthr) {
  const int n = 3 * n_node;
  for (int t0 = tid; t0 < n; t0 += 4 * nthr) {
    double2 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * nthr;
      if (t < n) { r[u].x = ld_relaxed_sys_f64(src + 2 * (size_t)t); r[u].y = ld_relaxed_sys_f64(src + 2 * (size_t)t + 1); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * nthr;
      if (t < n) *(double2*)(v + 6 * (size_t)idx[t / 3] + 2 * (t % 3)) = r[u];
    }
  }
}
__global__ void __launch_bounds__(1024)
mgd_exchange_kernel(const MgdCtx* c, const MgdLevelDev* L, int level, const int32_t* __restrict__ send_idx,
                    const int32_t* __restrict__ recv_idx, double* __restrict__ v, unsigned* counter, unsigned long long seq) {
  // ONE block (see above): the push is complete at the barrier, thread p then releases the flag of peer p
  const long long slot = (long long)level * c->arena_level_stride + (long long)(seq & 1ull) * c->arena_slot_stride;
  for (int pi = 0; pi < L->n_peer; ++pi)
    mgd_gather_pieces(c->arena[L->peer_rank[pi]] + slot + 6 * (long long)L->remote_off[pi], v, send_idx + L->send_off[pi],
                      L->send_cnt[pi], threadIdx.x, blockDim.x);
  // block barrier, then the flag threads fence system-wide and release: the barrier makes every thread's stores
  // observed by them, the fence / release is cumulative over what its thread has observed (PTX memory model) -- the
  // same pattern as mgd_push_kernel, and one fence per peer instead of one per thread on the critical path
  __syncthreads();
  if ((int)threadIdx.x < L->n_peer) {
    __threadfence_system();
    st_release_sys(&c->mbox[L->peer_rank[threadIdx.x]]->halo[level][c->rank], seq);
    const MgdMailbox* m = c->mbox[c->rank];
    while (ld_acquire_sys(&m->halo[level][L->peer_rank[threadIdx.x]]) < seq) { }
  }
  __syncthreads();
  mgd_scatter_pieces(v, c->arena[c->rank] + slot, recv_idx, L->n_recv, threadIdx.x, blockDim.x);
}

// Mailbox all-reduce of `count` <= 2 scalars by ONE thread.  The sequence number lives on the device (*seq_ctr) and
// advances only when a reduction really runs, so that kernels skipped after convergence (mgs_stopped) keep the
// parity double buffer of the mailbox consistent on every rank.
__device__ inline void mgd_allreduce_thread(const MgdCtx* c, const double* src, double* dst, int count,
                                            unsigned long long* seq_ctr) {
  const unsigned long long seq = *seq_ctr + 1ull;
  *seq_ctr = seq;
  const int par = (int)(seq & 1ull);
  for (int r = 0; r < c->n_rank; ++r)
    for (int k = 0; k < count; ++k) st_relaxed_sys_f64(&c->mbox[r]->red_val[par][c->rank][k], src[k]);
  __threadfence_system();
  for (int r = 0; r < c->n_rank; ++r) st_release_sys(&c->mbox[r]->red_tag[par][c->rank], seq);
  const MgdMailbox* m = c->mbox[c->rank];
  double acc[2] = {0.0, 0.0};
  for (int r = 0; r < c->n_rank; ++r) {
    while (ld_acquire_sys(&m->red_tag[par][r]) < seq) { }
    for (int k = 0; k < count; ++k) acc[k] += ld_relaxed_sys_f64(&m->red_val[par][r][k]);
  }
  for (int k = 0; k < count; ++k) dst[k] = acc[k];
}
// dst[0..count) <- sum over the ranks of src[0..count) (rank order: bitwise identical everywhere); one block,
// count <= 2.  dst may be src (in place) or another slot.
__global__ void mgd_allreduce_kernel(const MgdCtx* c, const double* src, double* dst, int count, unsigned long long* seq_ctr) {
  if (threadIdx.x == 0) mgd_allreduce_thread(c, src, dst, count, seq_ctr);
}
__device__ inline void mgs_store_dot(double total, double* out, const MgdCtx* ctx, unsigned long long* seq_ctr) {
  if (!ctx) { *out = total; return; }
  mgd_allreduce_thread(ctx, &total, out, 1, seq_ctr);
}
__device__ inline void mgs_store_dot_block(double total, double* out, const MgdCtx* ctx, unsigned long long* seq_ctr) {
  if (!ctx) { if (threadIdx.x == 0) *out = total; return; }
  __shared__ double sh_total;
  if (threadIdx.x == 0) sh_total = total;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    const unsigned long long seq = *seq_ctr + 1ull;
    const int par = (int)(seq & 1ull);
    __syncwarp();
    double v = 0.0;
    if (r < ctx->n_rank) {
      st_relaxed_sys_f64(&ctx->mbox[r]->red_val[par][ctx->rank][0], sh_total);
      __threadfence_system();
      st_release_sys(&ctx->mbox[r]->red_tag[par][ctx->rank], seq);
      const MgdMailbox* m = ctx->mbox[ctx->rank];
      while (ld_acquire_sys(&m->red_tag[par][r]) < seq) { }
      v = ld_relaxed_sys_f64(&m->red_val[par][r][0]);
    }
    double acc = 0.0;
    for (int k = 0; k < ctx->n_rank; ++k) acc += __shfl_sync(0xffffffffu, v, k);   // rank order: identical on every rank
    if (r == 0) { *out = acc; *seq_ctr = seq; }
  }
}
// a rank with an empty row range still takes part in the reduction of a fused dot product
__global__ void mgs_zero_dot_kernel(const double* stop, double* out, const MgdCtx* ctx, unsigned long long* seq_ctr) {
  if (mgs_stopped(stop)) return;
  mgs_store_dot(0.0, out, ctx, seq_ctr);
}

// ---- persistent CG (single GPU) ------------------------------------------------------
// All iterations of a batch in ONE cooperative launch: the three phases of an iteration are
// separated by grid-wide barriers instead of kernel boundaries (the 3-kernel loop is launch
// bound below ~1e5 dofs: ~17 us / iteration against ~5 us here).  Block b owns a contiguous
// range of block rows for the SpMV and the same slice of the vectors for the updates.  The dot
// products are summed from per-block partials in a fixed order by every block (deterministic,
// and every block takes the same stop decision).
__device__ inline double block_total(const double* partials, int n, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += ((const volatile double*)partials)[i];
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__device__ inline void block_partial(double v, double* partials, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(RED_BLOCK)
cg_persistent_kernel(int n_row, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                     const double* __restrict__ vals, double* __restrict__ x, double* __restrict__ r,
                     double* __restrict__ p, double* __restrict__ q, CgScalars* sc, double* partials,
                     int max_it) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  __shared__ double red[RED_BLOCK / 32];
  const int nb = gridDim.x, lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int r0 = (int)((long long)n_row * blockIdx.x / nb), r1 = (int)((long long)n_row * (blockIdx.x + 1) / nb);
  const long long v0 = 6LL * r0, v1 = 6LL * r1;
  double* part_a = partials;
  double* part_b = partials + RED_MAX_BLOCKS;
  double rr = sc->rr[0];
  const double tol = sc->tol2 * sc->bb;
  int it = 0;
  for (; it < max_it && rr > tol; ++it) {
    double dot = 0.0;
    for (int row = r0 + wib; row < r1; row += nw) {
      double u0, u1;
      bsr_row_product(row, lane, rowptr, colidx, vals, p, u0, u1);
      if (lane < 3) {
        *(double2*)(q + 6 * (size_t)row + 2 * lane) = make_double2(u0, u1);
        const double2 pr = *(const double2*)(p + 6 * (size_t)row + 2 * lane);
        dot += u0 * pr.x + u1 * pr.y;
      }
    }
    block_partial(dot, part_a, red);
    grid.sync();
    const double pq = block_total(part_a, nb, red);
    const double alpha = rr / pq;
    double acc = 0.0;
    for (long long i = v0 + threadIdx.x; i < v1; i += blockDim.x) {
      x[i] = fma(alpha, p[i], x[i]);
      const double ri = fma(-alpha, q[i], r[i]);
      r[i] = ri;
      acc = fma(ri, ri, acc);
    }
    block_partial(acc, part_b, red);
    grid.sync();
    double rr_new = block_total(part_b, nb, red);
    if (!(pq > 0.0)) rr_new = __longlong_as_double(0x7ff8000000000000LL);   // not SPD: poison
    const double beta = rr_new / rr;
    if (rr_new > tol)
      for (long long i = v0 + threadIdx.x; i < v1; i += blockDim.x) p[i] = fma(beta, p[i], r[i]);
    rr = rr_new;
    grid.sync();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->rr[0] = rr; sc->rr[1] = rr; sc->iter += it;
  }
}

// ---- block-Jacobi scaling ----------------------------------------------------------
// W_r = L_r^-1 (lower triangular, stored dense row-major 6x6) from the diagonal blocks.
__global__ void __launch_bounds__(128)
diag_factor_kernel(int n_row, const int32_t* __restrict__ diag_slot, const double* __restrict__ vals,
                   double* __restrict__ W, double* __restrict__ Lout /* optional: L row-major */, int* flags,
                   int tolerant = 0 /* coarse multigrid levels: a (near) zero pivot -> that dof keeps L = 1, no flag */) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_row) return;
  const double* d = vals + (size_t)diag_slot[r] * 36;
  double L[6][6];
  bool ok = true;
  double scale = 0.0;
#pragma unroll
  for (int j = 0; j < 6; ++j) scale = fmax(scale, fabs(d[j * 6 + j]));
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = d[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    // coarse levels: aggregates made of prescribed dofs give empty rows / columns -- those dofs stay unscaled
    const bool dead = tolerant && !(s > 1e-14 * scale);
    if (!(s > 0.0)) { ok = false; s = 1.0; }
    if (dead) s = 1.0;
    const double ljj = sqrt(s);
    L[j][j] = ljj;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = 0.5 * (d[j * 6 + i] + d[i * 6 + j]);   // symmetrised lower entry (i,j)
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
      L[i][j] = dead ? 0.0 : t / ljj;
    }
  }
  if (!ok && !tolerant) atomicOr(flags, 8);
  if (Lout) {
    double* lo = Lout + (size_t)r * 36;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) lo[i * 6 + j] = (j <= i) ? L[i][j] : 0.0;
  }
  // invert L (forward substitution on identity)
  double Wl[6][6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i < j) { Wl[i][j] = 0.0; continue; }
      double t = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) t -= L[i][k] * Wl[k][j];
      Wl[i][j] = t / L[i][i];
    }
  }
  double* w = W + (size_t)r * 36;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) w[i * 6 + j] = Wl[i][j];
}

// A^_rc = W_r A_rc W_c^T, in place; one thread per block.
__global__ void __launch_bounds__(128)
scale_blocks_kernel(long long nnzb, const int32_t* __restrict__ blk_row, const int32_t* __restrict__ colidx,
                    const double* __restrict__ W, double* __restrict__ vals) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nnzb) return;
  const double* wr = W + (size_t)blk_row[s] * 36;
  const double* wc = W + (size_t)colidx[s] * 36;
  double* a = vals + (size_t)s * 36;
  double A[6][6], T[6][6];
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int i = 0; i < 6; ++i) A[i][j] = a[j * 6 + i];
  // T = W_r A  (W_r lower triangular)
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k <= i; ++k) t += wr[i * 6 + k] * A[k][j];
      T[i][j] = t;
    }
  // A^ = T W_c^T : A^[i][j] = sum_k T[i][k] W_c[j][k], k <= j
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k <= j; ++k) t += T[i][k] * wc[j * 6 + k];
      a[j * 6 + i] = t;
    }
}

// out_r = W_r v_r (TRANS = 0) or W_r^T v_r (TRANS = 1); mask != null zeroes prescribed dofs of v first.
template <int TRANS>
__global__ void __launch_bounds__(128)
block_apply_kernel(int n_row, const double* __restrict__ W, const double* __restrict__ v,
                   const uint8_t* __restrict__ mask, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_row) return;
  const double* w = W + (size_t)r * 36;
  double x[6];
  const unsigned m = mask ? mask[r] : 0u;
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = ((m >> i) & 1u) ? 0.0 : v[6 * (size_t)r + i];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) t += (TRANS ? w[k * 6 + i] : w[i * 6 + k]) * x[k];
    out[6 * (size_t)r + i] = t;
  }
}

// y_r = W_r^-T x_r  (back substitution with the upper-triangular W_r^T): maps an initial guess
// of the unscaled system into the scaled one (x = W^T y).
__global__ void __launch_bounds__(128)
block_solve_wt_kernel(int n_row, const double* __restrict__ W, const double* __restrict__ x,
                      const uint8_t* __restrict__ mask, double* __restrict__ y) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_row) return;
  const double* w = W + (size_t)r * 36;
  const unsigned m = mask ? mask[r] : 0u;
  double v[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = ((m >> i) & 1u) ? 0.0 : x[6 * (size_t)r + i];
  // sum_k W[k][i] y[k] = v[i], W lower triangular => solve from i = 5 down
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double t = v[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) t -= w[k * 6 + i] * v[k];
    v[i] = t / w[i * 6 + i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) y[6 * (size_t)r + i] = v[i];
}

// y_r = W_r^-1 x_r  (forward substitution with the lower-triangular W_r): with block_solve_wt_kernel it applies
// the UNSCALED operator from the scaled values, K x = W^-1 (A^ (W^-T x))  (jsso_spmv after a solve)
__global__ void __launch_bounds__(128)
block_solve_w_kernel(int n_row, const double* __restrict__ W, const double* __restrict__ x, double* __restrict__ y) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_row) return;
  const double* w = W + (size_t)r * 36;
  double v[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = x[6 * (size_t)r + i];
#pragma unroll
    for (int k = 0; k < i; ++k) t -= w[i * 6 + k] * v[k];
    v[i] = t / w[i * 6 + i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) y[6 * (size_t)r + i] = v[i];
}

// out_rc = W_r^-1 A^_rc W_c^-T: the unscaled block from the scaled one (jsso_get_values after a solve); thread per block
__device__ inline void tri_inverse_lower(const double* __restrict__ w, double (&L)[6][6]) {
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i < j) { L[i][j] = 0.0; continue; }
      double t = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) t -= w[i * 6 + k] * L[k][j];
      L[i][j] = t / w[i * 6 + i];
    }
}
__global__ void __launch_bounds__(128)
unscale_blocks_kernel(long long nnzb, const int32_t* __restrict__ blk_row, const int32_t* __restrict__ colidx,
                      const double* __restrict__ W, const double* __restrict__ vals, double* __restrict__ out) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nnzb) return;
  double Lr[6][6], Lc[6][6], A[6][6], T[6][6];
  tri_inverse_lower(W + (size_t)blk_row[s] * 36, Lr);
  tri_inverse_lower(W + (size_t)colidx[s] * 36, Lc);
  const double* a = vals + (size_t)s * 36;
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int i = 0; i < 6; ++i) A[i][j] = a[j * 6 + i];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k <= i; ++k) t += Lr[i][k] * A[k][j];
      T[i][j] = t;
    }
  double* o = out + (size_t)s * 36;
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k <= j; ++k) t += T[i][k] * Lc[j][k];
      o[j * 6 + i] = t;
    }
}

// ---- CG vector kernels ---------------------------------------------------------------
// r = b - q (q = A x0 or 0), p = r, rr[0] = bb-candidate.  SETBB: also bb = |b|^2.
template <int SETBB>
__global__ void __launch_bounds__(RED_BLOCK)
cg_init_kernel(long long n, const double* __restrict__ b, const double* __restrict__ q, double* __restrict__ r,
               double* __restrict__ p, CgScalars* sc, double* partials, unsigned* counter, double rtol,
               int single_gpu) {
  double acc = 0.0, accb = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double bi = b[i];
    const double ri = q ? bi - q[i] : bi;
    r[i] = ri; p[i] = ri;
    acc += ri * ri; accb += bi * bi;
  }
  double total;
  if (grid_sum(acc, partials, counter, total) && threadIdx.x == 0) {
    sc->loc[1] = total; sc->iter = 0; sc->tol2 = rtol * rtol;
    if (single_gpu) { sc->rr[0] = total; sc->rr[1] = total; }
  }
  if (SETBB) {
    __syncthreads();
    double tb;
    if (grid_sum(accb, partials + RED_MAX_BLOCKS, counter + 1, tb) && threadIdx.x == 0) {
      sc->loc[2] = tb;
      if (single_gpu) sc->bb = tb;
    }
  }
}

// x += alpha p, r -= alpha q, rr[nxt] = r.r.
__global__ void __launch_bounds__(RED_BLOCK)
cg_update_kernel(long long n, int cur, const double* __restrict__ p, const double* __restrict__ q,
                 double* __restrict__ x, double* __restrict__ r, CgScalars* sc, double* partials,
                 unsigned* counter, int single_gpu, const P2PCtx* p2p, unsigned long long seq_a,
                 unsigned long long seq_b) {
  if (cg_stop(sc, cur)) return;
  const double rr = sc->rr[cur];
  const double pq = p2p ? p2p_reduce(p2p, 0, seq_a) : sc->pq;
  const double alpha = rr / pq;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    acc = fma(ri, ri, acc);
  }
  double total;
  if (grid_sum(acc, partials, counter, total) && threadIdx.x == 0) {
    // non-positive curvature => not SPD: poison the state so that everything stops
    const double nrr = (pq > 0.0) ? total : __longlong_as_double(0x7ff8000000000000LL);
    sc->loc[1] = nrr;
    sc->iter += 1;
    if (p2p) { sc->pq = pq; p2p_publish(p2p, 1, seq_b, nrr); }
    if (single_gpu) {
      sc->rr[cur ^ 1] = nrr;
      // On stop, latch BOTH parity slots: later launches of the batch test rr[cur] with
      // alternating cur.  Safe here: every block read rr[cur] before contributing its partial.
      if (!(nrr > sc->tol2 * sc->bb)) sc->rr[cur] = nrr;
    }
  }
}

// multi-GPU: after the all-reduce of rr[cur^1], latch the stop state in both slots
__global__ void cg_latch_kernel(int cur, CgScalars* sc) {
  if (!(sc->rr[cur ^ 1] > sc->tol2 * sc->bb)) sc->rr[cur] = sc->rr[cur ^ 1];
}
__global__ void cg_copy_rr_kernel(CgScalars* sc) { sc->rr[1] = sc->rr[0]; }

// p = r + beta p
__global__ void __launch_bounds__(RED_BLOCK)
cg_direction_kernel(long long n, int cur, const double* __restrict__ r, double* __restrict__ p,
                    CgScalars* sc, const P2PCtx* p2p, unsigned long long seq_b) {
  double beta;
  if (p2p) {
    if (cg_stop(sc, cur)) return;
    const double rr_new = p2p_reduce(p2p, 1, seq_b);
    // rr[cur^1] is not read by any block of this kernel in p2p mode
    if (blockIdx.x == 0 && threadIdx.x == 0) sc->rr[cur ^ 1] = rr_new;
    if (!(rr_new > sc->tol2 * sc->bb)) return;
    beta = rr_new / sc->rr[cur];
  } else {
    if (cg_stop(sc, cur) || cg_stop(sc, cur ^ 1)) return;
    beta = sc->rr[cur ^ 1] / sc->rr[cur];
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], r[i]);
}

// aux = |b - q|^2 (true residual check)
__global__ void __launch_bounds__(RED_BLOCK)
residual_norm_kernel(long long n, const double* __restrict__ b, const double* __restrict__ q, CgScalars* sc,
                     double* partials, unsigned* counter, int single_gpu) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = b[i] - q[i];
    acc += d * d;
  }
  double total;
  if (grid_sum(acc, partials, counter, total) && threadIdx.x == 0) {
    sc->loc[3] = total;
    if (single_gpu) sc->aux = total;
  }
}

// out = s * in  (owned part)
__global__ void scale_copy_kernel(long long n, double s, const double* __restrict__ in, double* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = s * in[i];
}

// halo pack: buf[k*6+d] = v[6*idx[k]+d]
__global__ void halo_pack_kernel(int n, const int32_t* __restrict__ idx, const double* __restrict__ v,
                                 double* __restrict__ buf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * n) return;
  buf[t] = v[6 * (size_t)idx[t / 6] + t % 6];
}
// halo unpack: v[6*idx[k]+d] = buf[k*6+d]  (receive side of the row-range distributed multigrid solve, where
// ghosts keep their global position instead of sitting in a contiguous tail)
__global__ void halo_unpack_kernel(int n, const int32_t* __restrict__ idx, const double* __restrict__ buf,
                                   double* __restrict__ v) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 6 * n) return;
  v[6 * (size_t)idx[t / 6] + t % 6] = buf[t];
}

}  // namespace jsso
