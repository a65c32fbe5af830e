// Element stiffness generation and numeric assembly kernels (sm_100a, FP64).
//
//  * stage_quad_geometry : per (quad, Gauss point) thread -> shared memory record of the quad
//  * quad_pair_block     : one (quad, node a, node b) pair item -> one 6x6 global block, from the record
//  * quad_ke_kernel / beam_ke_kernel : materialise K_e in the reference's `data`
//    layout (element.py:1236-1237, 270-271) for tests and for the stand-alone
//    assembly path
//  * quad_geometry_kernel + assemble_tasks_kernel : the production assembly.  The geometry record of
//    every quad is written once (496 B); persistent warps then own runs of block slots ("tasks"),
//    one pair item per lane, sums in list order, coalesced stores -- K_e never reaches HBM, no atomics.
//  * assemble_fused_kernel : the chunked single-kernel variant (one CTA per chunk of <= 128 block
//    slots, one thread owns one 6x6 block): fallback for meshes the warp tasks cannot hold, A/B reference.
//  * assemble_from_ke_kernel : stand-alone segmented reduction of materialised K_e.
#pragma once
#include "jsso_elem.cuh"
#include "jsso_symbolic.h"

namespace jsso {

// shared-memory record of one quad (doubles).  Every field group starts on an even offset and the stride is
// even, so that quad_pair_block reads the record with 16-byte loads (lanes working on the same quad read the
// same address: one broadcast wavefront moves twice the data of an 8-byte load); 70 k mod 16 is distinct for 8
// consecutive quads (the per-Gauss-point staging phase).
constexpr int QS = 70;
constexpr int Q_R = 0;      // 9: dirCos rows x^,y^,z^
constexpr int Q_KRZ = 9;    // drilling stiffness                      -> (R, krz) = 5 pairs
constexpr int Q_MAT = 10;   // cm11 cm12 cm21 cm22 cm33 D nu hb ks
constexpr int Q_M = 19;     // m11, m12, m22                           -> (MAT, M) = 6 pairs
constexpr int Q_GRY = 22;   // 2
constexpr int Q_GRX = 24;   // 2
constexpr int Q_GSY = 26;   // 2
constexpr int Q_GSX = 28;   // 2
constexpr int Q_GP = 30;    // 4 x {ji0..3, det, prr, prs, pss}
constexpr int Q_XY = 62;    // x0,y0,x1,y1,x3,y3 local coordinates (node 3 is the origin) -> 68

constexpr int FLAG_BADJAC = 1, FLAG_DEGBEAM = 2, FLAG_UNSYM = 4;

// Stage the geometry of `n_el` quads (ids from `els`, or first_el + i when els is null) into
// `sm`.  Called by all threads of the CTA.  Phase 0: coordinate gather.  Phase A: one thread per quad (frame, projected
// coordinates, gp-independent shear data, material).  Phase B: one thread per (quad, Gauss point)
// (inverse Jacobian, detJ, shear scale factors); the four Gauss points of a quad sit in adjacent
// lanes, so the drilling-stiffness minimum over the summed diagonal (element.py:978) is reduced
// with two shuffles.
__device__ inline void stage_quad_geometry(double* sm, int n_el, const int32_t* els, int first_el,
                                           const double* __restrict__ crds,
                                           const int32_t* __restrict__ cnct,
                                           const double* __restrict__ prop, int* flags) {
  // phase 0: one thread per (quad, node) gathers the coordinates (all global-load chains in
  // parallel); they are parked in the record's Gauss-point area, which phase B overwrites later
  for (int t = threadIdx.x; t < 4 * n_el; t += blockDim.x) {
    const int le = t >> 2, k = t & 3;
    const int e = els ? els[le] : first_el + le;
    const int nd = cnct[4 * e + k];
    double* s = sm + le * QS + Q_GP + 3 * k;
    s[0] = crds[3 * (size_t)nd]; s[1] = crds[3 * (size_t)nd + 1]; s[2] = crds[3 * (size_t)nd + 2];
  }
  __syncthreads();
  for (int le = threadIdx.x; le < n_el; le += blockDim.x) {
    const int e = els ? els[le] : first_el + le;
    double P[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) P[k][c] = sm[le * QS + Q_GP + 3 * k + c];
    QuadFrame<double> f;
    quad_frame(P, f);
    QuadShear<double> sh;
    quad_shear(f, sh);
    QuadMat m;
    quad_mat(prop + 5 * (size_t)e, m);
    double* s = sm + le * QS;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) s[Q_R + 3 * i + j] = f.R[i][j];
    s[Q_GRY] = sh.gry[0]; s[Q_GRY + 1] = sh.gry[1];
    s[Q_GRX] = sh.grx[0]; s[Q_GRX + 1] = sh.grx[1];
    s[Q_GSY] = sh.gsy[0]; s[Q_GSY + 1] = sh.gsy[1];
    s[Q_GSX] = sh.gsx[0]; s[Q_GSX + 1] = sh.gsx[1];
    s[Q_M] = sh.m11; s[Q_M + 1] = sh.m12; s[Q_M + 2] = sh.m22;
    s[Q_MAT] = m.cm11; s[Q_MAT + 1] = m.cm12; s[Q_MAT + 2] = m.cm21; s[Q_MAT + 3] = m.cm22;
    s[Q_MAT + 4] = m.cm33; s[Q_MAT + 5] = m.D; s[Q_MAT + 6] = m.nu; s[Q_MAT + 7] = m.hb;
    s[Q_MAT + 8] = m.ks;
    s[Q_XY] = f.x[0]; s[Q_XY + 1] = f.y[0]; s[Q_XY + 2] = f.x[1]; s[Q_XY + 3] = f.y[1];
    s[Q_XY + 4] = f.x[3]; s[Q_XY + 5] = f.y[3];
    if (prop[5 * (size_t)e + 3] != prop[5 * (size_t)e + 4]) atomicOr(flags, FLAG_UNSYM);
  }
  __syncthreads();
  const int n_task = 4 * n_el;
  for (int base = (threadIdx.x & ~31); base < n_task; base += blockDim.x) {
    int task = base + (threadIdx.x & 31);
    const bool live = task < n_task;
    if (!live) task = n_task - 1;
    const int le = task >> 2, q = task & 3;
    double* s = sm + le * QS;
    QuadFrame<double> f;
    f.x[0] = s[Q_XY]; f.y[0] = s[Q_XY + 1]; f.x[1] = s[Q_XY + 2]; f.y[1] = s[Q_XY + 3];
    f.x[2] = 0.0; f.y[2] = 0.0; f.x[3] = s[Q_XY + 4]; f.y[3] = s[Q_XY + 5];
    QuadShear<double> sh;
    sh.gry[0] = s[Q_GRY]; sh.gry[1] = s[Q_GRY + 1]; sh.grx[0] = s[Q_GRX]; sh.grx[1] = s[Q_GRX + 1];
    sh.gsy[0] = s[Q_GSY]; sh.gsy[1] = s[Q_GSY + 1]; sh.gsx[0] = s[Q_GSX]; sh.gsx[1] = s[Q_GSX + 1];
    sh.m11 = s[Q_M]; sh.m12 = s[Q_M + 1]; sh.m22 = s[Q_M + 2];
    QuadMat m;
    m.D = s[Q_MAT + 5]; m.nu = s[Q_MAT + 6]; m.hb = s[Q_MAT + 7]; m.ks = s[Q_MAT + 8];
    QuadGp<double> g;
    quad_gp(f, q, g);
    double dg[8];
    quad_diag_gp(sh, g, q, m, dg);
    double krz = 1e300;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dg[i] += __shfl_xor_sync(0xffffffffu, dg[i], 1);
      dg[i] += __shfl_xor_sync(0xffffffffu, dg[i], 2);
      krz = fmin(krz, fabs(dg[i]));
    }
    if (live) {
      double* sg = s + Q_GP + 8 * q;
      sg[0] = g.ji[0]; sg[1] = g.ji[1]; sg[2] = g.ji[2]; sg[3] = g.ji[3];
      sg[4] = g.det; sg[5] = g.prr; sg[6] = g.prs; sg[7] = g.pss;
      if (!(g.det > 0.0)) atomicOr(flags, FLAG_BADJAC);
      if (q == 0) s[Q_KRZ] = krz / 1000.0;
    }
  }
}

// G = R^T S R for S = [[s00,s01,0],[s10,s11,0],[0,0,s22]], ADDED into the 6x6
// column-major block `out` at sub-block (br, bc).
__device__ inline void rtsr_diag(const double* R, double s00, double s01, double s10, double s11,
                                 double s22, double* out, int br, int bc) {
  double u[3], v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    u[i] = s00 * R[i] + s10 * R[3 + i];      // sum_p R[p][i] S[p][0]
    v[i] = s01 * R[i] + s11 * R[3 + i];      // sum_p R[p][i] S[p][1]
  }
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i)
      out[(3 * bc + j) * 6 + 3 * br + i] += u[i] * R[j] + v[i] * R[3 + j] + (s22 * R[6 + i]) * R[6 + j];
}

// 6x6 global block (a,b) of a quad from its staged record `s`, ADDED into the column-major
// accumulator out[6*j+i].
__device__ inline void quad_pair_block(const double* s, int a, int b, double* out) {
  const double2* s2 = (const double2*)s;   // 16-byte aligned record (even stride, aligned base)
  const double ra = node_r(a), sa = node_s(a), rb = node_r(b), sb = node_s(b);
  double pxx = 0, pxy = 0, pyx = 0, pyy = 0, crr = 0, crs = 0, csr = 0, css = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double2 j01 = s2[Q_GP / 2 + 4 * q], j23 = s2[Q_GP / 2 + 4 * q + 1];      // J^-1
    const double2 dp = s2[Q_GP / 2 + 4 * q + 2], ps = s2[Q_GP / 2 + 4 * q + 3];    // (det, prr), (prs, pss)
    const double r = JSSO_GP * node_r(q), t = JSSO_GP * node_s(q);
    const double far = 1.0 + t * sa, fas = 1.0 + r * ra, fbr = 1.0 + t * sb, fbs = 1.0 + r * rb;
    const double dra = 0.25 * ra * far, dsa = 0.25 * sa * fas;
    const double drb = 0.25 * rb * fbr, dsb = 0.25 * sb * fbs;
    const double ha0 = j01.x * dra + j01.y * dsa, ha1 = j23.x * dra + j23.y * dsa;
    const double hb0 = j01.x * drb + j01.y * dsb, hb1 = j23.x * drb + j23.y * dsb;
    const double d0 = dp.x * ha0, d1 = dp.x * ha1;
    pxx += d0 * hb0; pxy += d0 * hb1; pyx += d1 * hb0; pyy += d1 * hb1;
    crr += dp.y * (far * fbr); crs += ps.x * (far * fbs);
    csr += ps.x * (fas * fbr); css += ps.y * (fas * fbs);
  }
  // (R, krz): 5 pairs; (MAT, M): 6 pairs; shear edge data: 4 pairs
  double R[10];
#pragma unroll
  for (int k = 0; k < 5; ++k) { const double2 v = s2[k]; R[2 * k] = v.x; R[2 * k + 1] = v.y; }
  double mt[12];
#pragma unroll
  for (int k = 0; k < 6; ++k) { const double2 v = s2[Q_MAT / 2 + k]; mt[2 * k] = v.x; mt[2 * k + 1] = v.y; }
  const double2 gry = s2[Q_GRY / 2], grx = s2[Q_GRX / 2], gsy = s2[Q_GSY / 2], gsx = s2[Q_GSX / 2];
  const double D = mt[5], nu = mt[6], hb = mt[7], ks = mt[8];
  // membrane 2x2 (u,v)
  const double muu = mt[0] * pxx + mt[4] * pyy, muv = mt[1] * pxy + mt[4] * pyx;
  const double mvu = mt[2] * pyx + mt[4] * pxy, mvv = mt[3] * pyy + mt[4] * pxx;
  // plate 3x3 (w, theta_x, theta_y): bending + MITC4 shear
  const bool ira = a >= 2, isa = !(a == 0 || a == 3);
  const bool irb = b >= 2, isb = !(b == 0 || b == 3);
  const double gra[3] = {0.5 * ra, ira ? gry.y : gry.x, ira ? grx.y : grx.x};
  const double gsa[3] = {0.5 * sa, isa ? gsy.y : gsy.x, isa ? gsx.y : gsx.x};
  const double grb[3] = {0.5 * rb, irb ? gry.y : gry.x, irb ? grx.y : grx.x};
  const double gsb[3] = {0.5 * sb, isb ? gsy.y : gsy.x, isb ? gsx.y : gsx.x};
  const double m11 = mt[9], m12 = mt[10], m22 = mt[11];
  double P[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double va = ks * (m11 * crr * gra[i] + m12 * csr * gsa[i]);
    const double wa = ks * (m12 * crs * gra[i] + m22 * css * gsa[i]);
#pragma unroll
    for (int j = 0; j < 3; ++j) P[i][j] = va * grb[j] + wa * gsb[j];
  }
  P[1][1] += D * (pyy + hb * pxx);
  P[1][2] -= D * (nu * pyx + hb * pxy);
  P[2][1] -= D * (nu * pxy + hb * pyx);
  P[2][2] += D * (pxx + hb * pyy);
  const double drill = (a == b) ? R[9] : 0.0;
  // translational-translational and rotational-rotational sub-blocks
  rtsr_diag(R, muu, muv, mvu, mvv, P[0][0], out, 0, 0);
  rtsr_diag(R, P[1][1], P[1][2], P[2][1], P[2][2], drill, out, 1, 1);
  // TR: S = e_z (P_w,thx  P_w,thy  0);  RT: S = (P_thx,w  P_thy,w  0)^T e_z^T
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double cj = P[0][1] * R[j] + P[0][2] * R[3 + j];
    const double rj = P[1][0] * R[j] + P[2][0] * R[3 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      out[(3 + j) * 6 + i] += R[6 + i] * cj;      // rows: translations, cols: rotations
      out[i * 6 + 3 + j] += rj * R[6 + i];        // rows: rotations (j), cols: translations (i)
    }
  }
}

// 6x6 global block (a,b) of a beam-column, K_e = T^-1 K_local T (element.py:107-139), ADDED into out.
__device__ inline void beam_pair_block(const double* __restrict__ crds, const int32_t* __restrict__ cnct,
                                       const double* __restrict__ prop, int e, int a, int b, double* out,
                                       int* flags) {
  double P[2][3];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int nd = cnct[2 * e + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) P[k][c] = crds[3 * (size_t)nd + c];
  }
  double R[3][3], L;
  const bool deg = beam_dircos(P, R, L);
  double Ri[3][3];  // T^-1 block: R^T when orthonormal, explicit inverse on the Cxz == 0 branch
  if (deg) {
    atomicOr(flags, FLAG_DEGBEAM);
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) -
                       R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    const double id = 1.0 / det;
    Ri[0][0] = (R[1][1] * R[2][2] - R[1][2] * R[2][1]) * id;
    Ri[0][1] = (R[0][2] * R[2][1] - R[0][1] * R[2][2]) * id;
    Ri[0][2] = (R[0][1] * R[1][2] - R[0][2] * R[1][1]) * id;
    Ri[1][0] = (R[1][2] * R[2][0] - R[1][0] * R[2][2]) * id;
    Ri[1][1] = (R[0][0] * R[2][2] - R[0][2] * R[2][0]) * id;
    Ri[1][2] = (R[0][2] * R[1][0] - R[0][0] * R[1][2]) * id;
    Ri[2][0] = (R[1][0] * R[2][1] - R[1][1] * R[2][0]) * id;
    Ri[2][1] = (R[0][1] * R[2][0] - R[0][0] * R[2][1]) * id;
    Ri[2][2] = (R[0][0] * R[1][1] - R[0][1] * R[1][0]) * id;
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Ri[i][j] = R[j][i];
  }
  const double* pr = prop + 6 * (size_t)e;
  const double E = pr[0], G = pr[1], Iy = pr[2], Iz = pr[3], J = pr[4], A = pr[5];
  const double L2 = L * L, L3 = L2 * L;
  const double ax = A * E / L, tz = G * J / L;
  const double bz12 = 12 * E * Iz / L3, bz6 = 6 * E * Iz / L2, bz4 = 4 * E * Iz / L, bz2 = 2 * E * Iz / L;
  const double by12 = 12 * E * Iy / L3, by6 = 6 * E * Iy / L2, by4 = 4 * E * Iy / L, by2 = 2 * E * Iy / L;
  const double sg = (a == b) ? 1.0 : -1.0;  // off-diagonal node blocks flip TT and RR(0,0)
  // local 6x6 block in 3x3 pieces (rows node a, cols node b), literal from element.py:115-126
  double TT[3] = {sg * ax, sg * bz12, sg * by12};                // diagonal
  double RR[3] = {sg * tz, (a == b) ? by4 : by2, (a == b) ? bz4 : bz2};
  // TR: (v, thz) and (w, thy);  RT: (thy, w) and (thz, v)
  const double s_a = (a == 0) ? 1.0 : -1.0;  // row node sign for TR
  const double s_b = (b == 0) ? 1.0 : -1.0;  // col node sign for RT
  const double tr_v_thz = s_a * bz6, tr_w_thy = -s_a * by6;
  const double rt_thy_w = -s_b * by6, rt_thz_v = s_b * bz6;
  // G_xy = Ri * S * R  for each 3x3 piece
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double tt = 0, rr = 0;
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        tt += Ri[i][p] * TT[p] * R[p][j];
        rr += Ri[i][p] * RR[p] * R[p][j];
      }
      const double tr = Ri[i][1] * tr_v_thz * R[2][j] + Ri[i][2] * tr_w_thy * R[1][j];
      const double rt = Ri[i][1] * rt_thy_w * R[2][j] + Ri[i][2] * rt_thz_v * R[1][j];
      out[j * 6 + i] += tt;
      out[(3 + j) * 6 + 3 + i] += rr;
      out[(3 + j) * 6 + i] += tr;
      out[j * 6 + 3 + i] += rt;
    }
}

// Materialise quad K_e: 16 threads per quad (one per node pair), 16 quads per CTA.
__global__ void __launch_bounds__(256)
quad_ke_kernel(int n_quad, const double* __restrict__ crds, const int32_t* __restrict__ cnct,
               const double* __restrict__ prop, double* __restrict__ ke, int* flags) {
  __shared__ __align__(16) double sm[16 * QS];
  const int first = blockIdx.x * 16;
  const int n_el = min(16, n_quad - first);
  stage_quad_geometry(sm, n_el, nullptr, first, crds, cnct, prop, flags);
  __syncthreads();
  const int le = threadIdx.x >> 4, a = (threadIdx.x >> 2) & 3, b = threadIdx.x & 3;
  if (le >= n_el) return;
  double out[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) out[k] = 0.0;
  quad_pair_block(sm + le * QS, a, b, out);
  double* dst = ke + (size_t)(first + le) * 576 + (6 * a) * 24 + 6 * b;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) dst[i * 24 + j] = out[j * 6 + i];
}

__global__ void __launch_bounds__(256)
beam_ke_kernel(int n_beam, const double* __restrict__ crds, const int32_t* __restrict__ cnct,
               const double* __restrict__ prop, double* __restrict__ ke, int* flags) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = t >> 2, a = (t >> 1) & 1, b = t & 1;
  if (e >= n_beam) return;
  double out[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) out[k] = 0.0;
  beam_pair_block(crds, cnct, prop, e, a, b, out, flags);
  double* dst = ke + (size_t)e * 144 + (6 * a) * 12 + 6 * b;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) dst[i * 12 + j] = out[j * 6 + i];
}

// Boundary conditions on one block entry: prescribed rows/cols -> identity
// (equivalent to the reference's Lagrange rows with zero prescribed displacement,
// assemblemodel.py:145-163, 192).
__device__ inline double bc_entry(double v, unsigned rmask, unsigned cmask, int i, int j, bool diag_blk) {
  const bool rk = (rmask >> i) & 1u, ck = (cmask >> j) & 1u;
  if (rk || ck) return (diag_blk && i == j) ? 1.0 : 0.0;
  return v;
}

constexpr int OUT_LD = 37;  // padded block stride of the output transposition buffer (odd => few conflicts)
constexpr int FUSED_SMEM_DOUBLES =
    (kChunkQuads * QS > kChunkBlocks * OUT_LD) ? kChunkQuads * QS : kChunkBlocks * OUT_LD;

struct AsmArgs {
  const double* crds; const int32_t* cnct_q; const double* prop_q;
  const int32_t* cnct_b; const double* prop_b;
  const int32_t* chunk_blk; const int32_t* chunk_el_ptr; const int32_t* chunk_els;
  const int32_t* blk_perm; const int32_t* blk_item_ptr; const int32_t* item_code; const uint8_t* item_lel;
  const int32_t* blk_row; const int32_t* colidx; const uint8_t* node_mask;
  double* vals; int* flags; int n_quad; int apply_bc;
};

// One CTA per chunk of consecutive block slots.  After the chunk's quads are staged in shared
// memory, ONE THREAD OWNS ONE 6x6 BLOCK: it loops over the block's contributors (element, a, b)
// in list order and accumulates in registers -- no atomics, a fixed summation order; the chunk's
// blocks are then transposed through shared memory and written with coalesced stores.  Blocks are dealt to threads by decreasing
// contributor count (blk_perm), so the lanes of a warp loop the same number of times.
__global__ void __launch_bounds__(kChunkBlocks, 4)
assemble_fused_kernel(AsmArgs A) {
  JSSO_DYN_SMEM(sm);
  const int c = blockIdx.x;
  const int blk0 = A.chunk_blk[c], blk1 = A.chunk_blk[c + 1];
  const int el0 = A.chunk_el_ptr[c], n_el = A.chunk_el_ptr[c + 1] - el0;
  // this thread's block and contributor range: loaded before the staging phases so that the
  // dependent global-load chain (perm -> item ptr) is hidden behind them
  const bool has_blk = (int)threadIdx.x < blk1 - blk0;
  const int blk = has_blk ? A.blk_perm[blk0 + threadIdx.x] : blk0;
  const int i0 = A.blk_item_ptr[blk], i1 = has_blk ? A.blk_item_ptr[blk + 1] : i0;
  stage_quad_geometry(sm, n_el, A.chunk_els + el0, 0, A.crds, A.cnct_q, A.prop_q, A.flags);
  __syncthreads();
  double acc[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = 0.0;
  for (int it = i0; it < i1; ++it) {
    const int code = A.item_code[it];
    const int el = code >> 4, a = (code >> 2) & 3, b = code & 3;
    if (el < A.n_quad) quad_pair_block(sm + A.item_lel[it] * QS, a, b, acc);
    else beam_pair_block(A.crds, A.cnct_b, A.prop_b, el - A.n_quad, a, b, acc, A.flags);
  }
  if (has_blk && A.apply_bc) {
    const int r = A.blk_row[blk], cc = A.colidx[blk];
    const unsigned rm = A.node_mask[r], cm = A.node_mask[cc];
    if (rm | cm) {
#pragma unroll
      for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[j * 6 + i] = bc_entry(acc[j * 6 + i], rm, cm, i, j, r == cc);
    }
  }
  // Transpose through shared memory (the geometry records are dead now) so that the chunk's
  // contiguous range of vals is written with fully coalesced 16-byte stores: thread-private
  // 288-byte segments would touch 32 different sectors per store instruction.
  __syncthreads();
  if (has_blk) {
    double* dstb = sm + (blk - blk0) * OUT_LD;
#pragma unroll
    for (int k = 0; k < 36; ++k) dstb[k] = acc[k];
  }
  __syncthreads();
  const int n_pair = (blk1 - blk0) * 18;   // 16-byte pairs in this chunk's slice of vals
  double2* out2 = (double2*)(A.vals + (size_t)blk0 * 36);
  for (int o = threadIdx.x; o < n_pair; o += blockDim.x) {
    const int bl = o / 18, k2 = o - bl * 18;
    const double* src = sm + bl * OUT_LD + 2 * k2;
    out2[o] = make_double2(src[0], src[1]);
  }
}

// ---- two-kernel assembly: per-quad geometry records + warp tasks ------------------------------------
// (1) quad_geometry_kernel evaluates the staged record of EVERY quad once (no redundancy between
//     neighbouring chunks, full occupancy) and writes its first REC doubles (everything
//     quad_pair_block reads) to a global buffer, 496 bytes per quad.
// (2) assemble_tasks_kernel: one WARP owns a run of consecutive block slots with <= 32 pair items
//     (symbolic pass, `task_meta`), copies the <= kTaskQuads records it needs from global/L2 into its
//     private slice of shared memory, evaluates ONE pair item per lane (uniform work, no loop, no
//     dependent loads), parks the 36 values of every item in shared memory and then writes the run's
//     contiguous slice of `vals` with coalesced 16-byte stores, each entry summing its contributors
//     in list order (fixed order, no atomics).  Warps never synchronise with each other.
// tuning switches (A/B builds: scripts/variants.py)
#ifndef JSSO_T_DEEP
#define JSSO_T_DEEP 0      // descriptors loaded two tasks ahead (0: one task ahead)
#endif
#ifndef JSSO_T_L2PF
#define JSSO_T_L2PF 0      // L2 prefetch of the next task's records during the compute phase
#endif
#ifndef JSSO_T_UNROLL1
#define JSSO_T_UNROLL1 1   // item loop of the output phase not unrolled
#endif
#ifndef JSSO_T_OUT
#define JSSO_T_OUT 0       // 0: three blocks per step, coalesced; 1: lane l sums and stores block l
#endif
#ifndef JSSO_T_WARPS
#define JSSO_T_WARPS 4
#endif
#ifndef JSSO_T_GLD
#define JSSO_T_GLD 62
#endif
constexpr int REC = 62;        // doubles of the record that quad_pair_block reads (Q_R .. Q_GP + 31), 31 x 16 B
constexpr int REC_GLD = JSSO_T_GLD;    // global stride of a record: 512 B = four aligned 128-byte lines
constexpr int REC_LD = 66;     // shared-memory stride: 16-byte aligned rows; 2*lq + c distinct mod 16 for 8 quads
constexpr int ITEM_LD = 38;    // even stride (16-byte aligned rows), 19 x 16 B: conflict-free 16-byte accesses
constexpr int TASK_SMEM_DOUBLES = kTaskQuads * REC_LD + kTaskItems * ITEM_LD;   // 1744 doubles = 13 952 B per warp
constexpr int TASK_WARPS = JSSO_T_WARPS;  // warps per CTA
static_assert((kTaskQuads * REC_LD) % 2 == 0 && TASK_SMEM_DOUBLES % 2 == 0, "16-byte alignment of the item rows");

#ifndef JSSO_G_QUADS
#define JSSO_G_QUADS 64      // quads per CTA of quad_geometry_kernel (variant sweep: 32 -> 0.223 ms, 64 -> 0.200 ms, 128 -> 0.255 ms)
#endif
#ifndef JSSO_G_THREADS
#define JSSO_G_THREADS 128
#endif
constexpr int G_QUADS = JSSO_G_QUADS, G_THREADS = JSSO_G_THREADS;

__global__ void __launch_bounds__(G_THREADS)
quad_geometry_kernel(int n_quad, const double* __restrict__ crds, const int32_t* __restrict__ cnct,
                     const double* __restrict__ prop, double* __restrict__ rec, int* flags) {
  JSSO_DYN_SMEM(sm);   // G_QUADS * QS doubles
  const int first = blockIdx.x * G_QUADS;
  const int n_el = min(G_QUADS, n_quad - first);
  stage_quad_geometry(sm, n_el, nullptr, first, crds, cnct, prop, flags);
  __syncthreads();
  double* dst = rec + (size_t)first * REC_GLD;
  if (REC_GLD == REC) {
    for (int idx = threadIdx.x; idx < n_el * REC; idx += blockDim.x) {
      const int le = idx / REC, w = idx - le * REC;
      dst[idx] = sm[le * QS + w];
    }
  } else {
    for (int idx = threadIdx.x; idx < n_el * REC_GLD; idx += blockDim.x) {
      const int le = idx / REC_GLD, w = idx % REC_GLD;
      if (w < REC) dst[idx] = sm[le * QS + w];
    }
  }
}

struct TaskArgs {
  const double* rec;            // n_quad x REC_GLD
  const int4* task_meta; const int32_t* task_els; const uint16_t* item_desc; const uint16_t* blk_bc;
  const int32_t* blk_item_ptr;  // item start of every block
  const int32_t* item_code;     // beams: global element id
  const double* crds; const int32_t* cnct_b; const double* prop_b;
  double* vals; int* flags; int n_quad; int n_task; int apply_bc;
};

#ifndef JSSO_T_BULK
#define JSSO_T_BULK 1      // records staged by 1-D bulk async copies (cp.async.bulk + mbarrier: the TMA unit) instead of per-lane cp.async
                           // (A/B at 1M quads, profiles/r2_asm_bulk_copy_ab.txt: 1.202 vs 1.220 ms for the assembly stage)
#endif
#ifdef JSSO_EMU   // CPU test harness: a synchronous copy (a missing wait is not detected there)
__device__ inline void cp_async16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
__device__ inline void prefetch_l2(const void*) {}
__device__ inline void cp_async_commit() {}
__device__ inline void cp_async_wait_all() {}
__device__ inline void mbar_init(unsigned long long*, int) {}
__device__ inline void mbar_expect_tx(unsigned long long*, unsigned) {}
__device__ inline void bulk_copy_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long*) { memcpy(smem_dst, gsrc, bytes); }
__device__ inline void mbar_wait(unsigned long long*, unsigned) {}
__device__ inline void fence_proxy_async() {}
#else
// ---- 1-D bulk asynchronous copies (sm_90+: cp.async.bulk, executed by the TMA unit; SASS UBLKCP) ------------------
// One instruction of ONE lane moves a whole 496-byte record global -> shared; completion is counted in bytes on an
// mbarrier in shared memory, which the warp polls with try_wait.  No per-lane LDGSTS, no LSU wavefronts per lane --
// the LSU data pipe is what bounds assemble_tasks_kernel (85 % of peak with the cp.async staging).
__device__ inline void mbar_init(unsigned long long* bar, int count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {   // one arrival + the bytes to wait for
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ inline void bulk_copy_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(d), "l"(gsrc), "r"(bytes), "r"(a) : "memory");
}
__device__ inline void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "JSSO_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra JSSO_MBAR_DONE;\n"
      "bra JSSO_MBAR_WAIT;\n"
      "JSSO_MBAR_DONE:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
// orders this thread's earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) writes
__device__ inline void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ inline void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ inline void prefetch_l2(const void* g) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(g)); }
__device__ inline void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ inline void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
#endif

// per-lane data of one task that is loaded ahead of time (registers)
struct TaskRegs {
  int blk0, cnt;      // first block slot; n_blk | n_item<<8 | n_el<<16
  unsigned desc;      // item descriptor of item `lane`
  unsigned bc;        // boundary word of local block `lane`
  int st;             // lane index of the first item of local block `lane` (n_item beyond the last block)
  int el;             // quad id of record `lane`
  int beam_el;        // beam id of item `lane` (beam items only)
};

__device__ inline void task_load(const TaskArgs& A, const int4 m, int lane, TaskRegs& t) {
  const int n_blk = m.w & 255, n_item = (m.w >> 8) & 255, n_el = (m.w >> 16) & 255;
  t.blk0 = m.x; t.cnt = m.w;
  t.desc = (lane < n_item) ? A.item_desc[m.y + lane] : 0u;
  t.bc = (A.apply_bc && lane < n_blk) ? A.blk_bc[m.x + lane] : 0u;
  t.st = (lane < n_blk) ? A.blk_item_ptr[m.x + lane] - m.y : n_item;
  t.el = (lane < n_el) ? A.task_els[m.z + lane] : 0;
  t.beam_el = ((t.desc & kDescBeam) && lane < n_item) ? (A.item_code[m.y + lane] >> 4) - A.n_quad : 0;
}

// records of the task's quads -> this warp's record buffer: one bulk copy per record issued by lane `le` (JSSO_T_BULK),
// or 16-byte cp.async by 31 lanes per record
__device__ inline void task_stage_records(const TaskArgs& A, const TaskRegs& t, int lane, double* rec, unsigned long long* bar) {
  const int n_el = (t.cnt >> 16) & 255;
#if JSSO_T_BULK
  // the warp has finished reading the buffer (the caller's __syncwarp); make that visible to the async proxy, arm the
  // barrier with the byte count, then every lane < n_el sends its record
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) mbar_expect_tx(bar, (unsigned)(n_el * REC * sizeof(double)));
  __syncwarp();
  if (lane < n_el) bulk_copy_g2s(rec + lane * REC_LD, A.rec + (size_t)t.el * REC_GLD, (unsigned)(REC * sizeof(double)), bar);
#else
  for (int le = 0; le < n_el; ++le) {
    const int e = __shfl_sync(0xffffffffu, t.el, le);
    if (lane < REC / 2) cp_async16(rec + le * REC_LD + 2 * lane, A.rec + (size_t)e * REC_GLD + 2 * lane);
  }
  cp_async_commit();
#endif
}

// Persistent warps: warp gw handles tasks gw, gw + W, gw + 2W, ...  Software pipeline per warp:
// the meta word of task k+2 and the per-lane descriptors of task k+1 are in registers while task k
// is computed, and the records of task k+1 stream into shared memory (cp.async) during the output
// phase of task k, so no global-load latency is exposed after the prologue.
__global__ void __launch_bounds__(32 * TASK_WARPS, 16 / TASK_WARPS)
assemble_tasks_kernel(TaskArgs A) {
  JSSO_DYN_SMEM(sm);
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int stride = gridDim.x * TASK_WARPS;
  int task = blockIdx.x * TASK_WARPS + w;
  if (task >= A.n_task) return;                       // whole warp; no CTA-wide barrier anywhere
  double* rec = sm + w * TASK_SMEM_DOUBLES;
  double* buf = rec + kTaskQuads * REC_LD;
  __shared__ unsigned long long bars[TASK_WARPS];   // one mbarrier per warp (bulk-copy completion of its record buffer)
  unsigned long long* bar = &bars[w];
  unsigned phase = 0;
#if JSSO_T_BULK
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
#endif
  // lane -> (block of the triple, 16-byte pieces k and k + 9 of its 18) in the output phase
  const int c0 = lane / 9, kp = lane - 9 * c0;
  TaskRegs cur, nxt, nn;
  task_load(A, A.task_meta[task], lane, cur);
  task_stage_records(A, cur, lane, rec, bar);
  nxt = cur; nn = cur;
  if (task + stride < A.n_task) task_load(A, A.task_meta[task + stride], lane, nxt);
  int4 m2 = make_int4(0, 0, 0, 0);                    // meta word of task + 2 stride (one iteration ahead of its use)
  if (task + 2 * stride < A.n_task) m2 = A.task_meta[task + 2 * stride];
  for (; task < A.n_task; task += stride) {
    const bool has_next = task + stride < A.n_task, has_next2 = task + 2 * stride < A.n_task;
#if JSSO_T_DEEP
    if (has_next2) task_load(A, m2, lane, nn);        // descriptors two tasks ahead (m2 arrived long ago)
    if (task + 3 * stride < A.n_task) m2 = A.task_meta[task + 3 * stride];
#else
    if (has_next2) m2 = A.task_meta[task + 2 * stride];
#endif
    const int n_blk = cur.cnt & 255, n_item = (cur.cnt >> 8) & 255;
#if JSSO_T_L2PF
    {
      // pull the next task's records (four 128-byte lines each) into L2 while this task computes
      const int e = __shfl_sync(FULL, nxt.el, lane >> 2);
      if (has_next && (lane >> 2) < ((nxt.cnt >> 16) & 255)) prefetch_l2(A.rec + (size_t)e * REC_GLD + 16 * (lane & 3));
    }
#endif
#if JSSO_T_BULK
    mbar_wait(bar, phase);
    phase ^= 1u;
#else
    cp_async_wait_all();
#endif
    __syncwarp();
    if (lane < n_item) {
      double out[36];
#pragma unroll
      for (int k = 0; k < 36; ++k) out[k] = 0.0;
      const int b = (cur.desc >> 5) & 3, a = (cur.desc >> 7) & 3, lq = (cur.desc >> 9) & 7;
      if (cur.desc & kDescBeam) beam_pair_block(A.crds, A.cnct_b, A.prop_b, cur.beam_el, a, b, out, A.flags);
      else quad_pair_block(rec + lq * REC_LD, a, b, out);
      double2* dst = (double2*)(buf + lane * ITEM_LD);
#pragma unroll
      for (int k2 = 0; k2 < 18; ++k2) dst[k2] = make_double2(out[2 * k2], out[2 * k2 + 1]);
    }
    __syncwarp();                                     // item rows complete; record buffer free
    if (has_next) task_stage_records(A, nxt, lane, rec, bar);
#if !JSSO_T_DEEP
    if (has_next2) task_load(A, m2, lane, nn);
#endif
#if JSSO_T_OUT == 1
    // output: lane l sums block l over its items in list order (registers) and stores its 288 bytes
    {
      int s1 = __shfl_down_sync(FULL, cur.st, 1);
      if (lane + 1 >= n_blk) s1 = n_item;
      if (lane < n_blk) {
        const double2* src = (const double2*)(buf + cur.st * ITEM_LD);
        double2 v[18];
#pragma unroll
        for (int k2 = 0; k2 < 18; ++k2) v[k2] = src[k2];
#pragma unroll 1
        for (int it = cur.st + 1; it < s1; ++it) {
          src += ITEM_LD / 2;
#pragma unroll
          for (int k2 = 0; k2 < 18; ++k2) { const double2 t = src[k2]; v[k2].x += t.x; v[k2].y += t.y; }
        }
        if (cur.bc & 0xfffu) {
          const bool diag = (cur.bc >> 12) & 1u;
          const unsigned rm = cur.bc & 63u, cm = (cur.bc >> 6) & 63u;
#pragma unroll
          for (int k2 = 0; k2 < 18; ++k2) {
            v[k2].x = bc_entry(v[k2].x, rm, cm, 2 * (k2 % 3), k2 / 3, diag);
            v[k2].y = bc_entry(v[k2].y, rm, cm, 2 * (k2 % 3) + 1, k2 / 3, diag);
          }
        }
        double2* o = (double2*)(A.vals + (size_t)(cur.blk0 + lane) * 36);
#pragma unroll
        for (int k2 = 0; k2 < 18; ++k2) o[k2] = v[k2];
      }
    }
#else
    // output: three blocks per step, lane (c0, kp) sums pieces kp and kp + 9 of block 3 j + c0 over
    // the block's items in list order and writes them (27 lanes x 2 x 16 B, contiguous per block)
    double2* out2 = (double2*)(A.vals + (size_t)cur.blk0 * 36);
    for (int b3 = 0; b3 < n_blk; b3 += 3) {
      const int bl = b3 + c0;
      const bool live = (lane < 27) && (bl < n_blk);
      const int s0 = __shfl_sync(FULL, cur.st, bl & 31);
      int s1 = __shfl_sync(FULL, cur.st, (bl + 1) & 31);
      if (bl + 1 >= n_blk) s1 = n_item;
      const unsigned bc = __shfl_sync(FULL, cur.bc, bl & 31);
      if (live) {
        const double2* src = (const double2*)(buf + s0 * ITEM_LD) + kp;
        double2 v0 = src[0], v1 = src[9];
#if JSSO_T_UNROLL1
#pragma unroll 1
#endif
        for (int it = s0 + 1; it < s1; ++it) {
          src += ITEM_LD / 2;
          const double2 t0 = src[0], t1 = src[9];
          v0.x += t0.x; v0.y += t0.y; v1.x += t1.x; v1.y += t1.y;
        }
        if (bc & 0xfffu) {                            // prescribed rows/cols -> identity
          const int j = kp / 3, i = 2 * (kp - 3 * j); // piece kp: entries (i, j), (i+1, j); piece kp+9: column j+3
          const bool diag = (bc >> 12) & 1u;
          const unsigned rm = bc & 63u, cm = (bc >> 6) & 63u;
          v0.x = bc_entry(v0.x, rm, cm, i, j, diag);     v0.y = bc_entry(v0.y, rm, cm, i + 1, j, diag);
          v1.x = bc_entry(v1.x, rm, cm, i, j + 3, diag); v1.y = bc_entry(v1.y, rm, cm, i + 1, j + 3, diag);
        }
        out2[bl * 18 + kp] = v0;
        out2[bl * 18 + kp + 9] = v1;
      }
    }
#endif
    __syncwarp();                                     // item rows read before the next task overwrites them
    cur = nxt;
    nxt = nn;
  }
}

// Post-processing on the same data: surface area of every quad as the two
// triangles (1,2,4) and (3,4,2) (JaxSSO/element.py:471-487, used by the size-optimisation example
// for the material volume sum_e t_e A_e).
__global__ void __launch_bounds__(256)
quad_area_kernel(int n_quad, const double* __restrict__ crds, const int32_t* __restrict__ cnct,
                 double* __restrict__ area) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_quad) return;
  double P[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int nd = cnct[4 * e + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) P[k][c] = crds[3 * (size_t)nd + c];
  }
  double a[3], b[3], c[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    a[i] = P[1][i] - P[0][i]; b[i] = P[3][i] - P[0][i];
    c[i] = P[3][i] - P[2][i]; d[i] = P[1][i] - P[2][i];
  }
  const double x1 = a[1] * b[2] - a[2] * b[1], y1 = a[2] * b[0] - a[0] * b[2], z1 = a[0] * b[1] - a[1] * b[0];
  const double x2 = c[1] * d[2] - c[2] * d[1], y2 = c[2] * d[0] - c[0] * d[2], z2 = c[0] * d[1] - c[1] * d[0];
  area[e] = 0.5 * (sqrt(x1 * x1 + y1 * y1 + z1 * z1) + sqrt(x2 * x2 + y2 * y2 + z2 * z2));
}

// y = A x for a scalar CSR matrix (the sparse radius filters either side of the path, SURVEY 8(f)
// rank 3); 8 lanes per row
__global__ void __launch_bounds__(256)
csr_spmv_kernel(int n_row, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = t >> 3, sub = t & 7;
  double s = 0.0;
  if (row < n_row)
    for (int k = rowptr[row] + sub; k < rowptr[row + 1]; k += 8) s = fma(vals[k], x[colidx[k]], s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (row < n_row && sub == 0) y[row] = s;
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(long long total, int width, const double* __restrict__ src, const int32_t* __restrict__ idx,
                   double* __restrict__ dst) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const long long i = o / width;
  dst[o] = src[(size_t)idx[i] * width + (o - i * width)];
}

// Boundary conditions on an already assembled matrix (values given by the caller)
__global__ void __launch_bounds__(256)
apply_bc_kernel(long long n_out, const int32_t* __restrict__ blk_row, const int32_t* __restrict__ colidx,
                const uint8_t* __restrict__ node_mask, double* __restrict__ vals) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int blk = (int)(o / 36), k = (int)(o - (long long)blk * 36);
  const int r = blk_row[blk], c = colidx[blk];
  vals[o] = bc_entry(vals[o], node_mask[r], node_mask[c], k % 6, k / 6, r == c);
}

// Stand-alone numeric assembly: one thread per stored entry, contributors summed in
// list order from the materialised element matrices.
__global__ void __launch_bounds__(256)
assemble_from_ke_kernel(long long n_out, int n_quad, const double* __restrict__ ke_q,
                        const double* __restrict__ ke_b, const int32_t* __restrict__ blk_item_ptr,
                        const int32_t* __restrict__ item_code, const int32_t* __restrict__ blk_row,
                        const int32_t* __restrict__ colidx, const uint8_t* __restrict__ node_mask,
                        double* __restrict__ vals, int apply_bc) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int blk = (int)(o / 36), k = (int)(o - (long long)blk * 36);
  const int i = k % 6, j = k / 6;
  double s = 0.0;
  for (int it = blk_item_ptr[blk]; it < blk_item_ptr[blk + 1]; ++it) {
    const int code = item_code[it];
    const int el = code >> 4, a = (code >> 2) & 3, b = code & 3;
    if (el < n_quad) s += ke_q[(size_t)el * 576 + (6 * a + i) * 24 + 6 * b + j];
    else s += ke_b[(size_t)(el - n_quad) * 144 + (6 * a + i) * 12 + 6 * b + j];
  }
  if (apply_bc) {
    const int r = blk_row[blk], c = colidx[blk];
    s = bc_entry(s, node_mask[r], node_mask[c], i, j, r == c);
  }
  vals[o] = s;
}

}  // namespace jsso
