// Fused adjoint sensitivity kernels.
//
// The reference obtains dL/dx by (i) the VJP of -(K u - f) w.r.t. every stored COO
// entry at cotangent lam, i.e. W_e[a,b] = -lam[dof_e[a]] u[dof_e[b]]
// (JaxSSO/solver.py:157-166, 239-248), then (ii) XLA reverse mode through
// sum_duplicates / sort / concatenate and vmap(element_K_*) including its batched LU
// solves.  Mathematically that is
//     dL/dx = - d/dx [ lam_e^T K_e(x) u_e ]   summed over elements,
// so these kernels never form K_e, dK_e or the nse-long cotangent: they differentiate the
// scalar bilinear form e(x) = lam_e^T K_e(x) u_e evaluated through the element's strains.
//   * quads: hand-derived reverse mode, thread = (quad, Gauss point) (below);
//   * beam-columns: forward-mode duals seeded on the three coordinates of one node,
//     thread = (beam, node), on the same templated geometry code as the stiffness kernel.
// Property derivatives are closed-form from the split energies.  Nodal gradients are then
// gathered per node from the per-corner partials in a fixed order (no atomics).
#pragma once
#include "jsso_elem.cuh"

namespace jsso {

using D3 = Dual<3>;

// ---------------------------------------------------------------------------------------
// Hand-derived reverse mode of e(X) = lam_e^T K_e(X) u_e for the MITC4 quad.
//
// thread = (quad, Gauss point); the 4 lanes of a quad cooperate through shuffles.  With
// h_k = J^-1 dN_k (physical shape-function gradients), fields phi in {u, v, theta_x, theta_y}
// of both vectors, Phi the membrane+bending energy density and S_phi = dPhi/d(grad phi):
//   d(detJ Phi)/dx_m = detJ (h_m,x Phi - h_m . T_x),   T = sum_phi S_phi (x) grad phi
//   d(detJ Phi)/dphi_k = detJ S_phi . h_k
// (the classic shape derivative: d grad phi = -h_m (grad phi)_x dx_m, d detJ = detJ h_m,x dx_m).
// The MITC4 shear energy is  ks [m11 Prr al au + m12 Prs (al bu + bl au) + m22 Pss bl bu]
// with a(s) = (1+s) A0 + (1-s) A1, b(r) = (1+r) B0 + (1-r) B1 the covariant edge strains and
// Prr = |J row 1|^2/(4 detJ), Pss = |J row 0|^2/(4 detJ), Prs = sqrt(Prr Pss); its adjoint goes
// through (Prr, Prs, Pss) -> J -> local coordinates, through m12 and through the edge vectors.
// The drilling spring min|diag|/1000 * sum_k lam_thz,k u_thz,k differentiates the arg-min
// diagonal entry only (first arg-min, element.py:978).  Local-coordinate and frame adjoints are
// finally pulled back through the normalised cross-product frame (element.py:502-539).
// ---------------------------------------------------------------------------------------
constexpr int ADJ_QUADS = 32;        // quads per batch (128 threads)
constexpr int ADJ_ST = 66;           // doubles per quad of the gather stage: P 0..11 | u 12..35 | lam 36..59 | prop 60..64
                                     // (even: 16-byte aligned rows; 66 k mod 16 distinct for the 8 quads of a warp)
constexpr int ADJ_LD = 87;           // doubles per quad of the local-vector buffer (odd stride): UL 24 | LL 24 | 12 property sums | 6 local xy | 11 shear data | 9 material
constexpr int A_P = 0, A_UG = 12, A_LG = 36, A_PR = 60;   // gather stage: coordinates, global u, global lam, properties
constexpr int A_UL = 0, A_LL = 24, A_PS = 48, A_XY = 60, A_SH = 66, A_MT = 77;  // local-vector buffer: local u, local lam, property sums, local x0 y0 x1 y1 x3 y3
constexpr int ADJ_SMEM_DOUBLES = 2 * ADJ_QUADS * ADJ_ST + ADJ_QUADS * ADJ_LD;   // 56 064 B of dynamic shared memory

__device__ inline double quad4_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

#ifdef JSSO_EMU   // CPU test harness: synchronous copies
__device__ inline void adj_cp8(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 8); }
__device__ inline void adj_cp16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
__device__ inline void adj_cp_commit() {}
__device__ inline void adj_cp_wait_1() {}
__device__ inline void adj_cp_wait_all() {}
#else
__device__ inline void adj_cp8(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ inline void adj_cp16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ inline void adj_cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ inline void adj_cp_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ inline void adj_cp_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
#endif
// lane (quad, q) copies node q's coordinates, u and lam rows into the quad's stage record
__device__ inline void adj_stage(double* rec, int q, int nd, int e, const double* __restrict__ crds,
                                 const double* __restrict__ prop, const double* __restrict__ u,
                                 const double* __restrict__ lam) {
#pragma unroll
  for (int c = 0; c < 3; ++c) adj_cp8(rec + A_P + 3 * q + c, crds + 3 * (size_t)nd + c);
  adj_cp8(rec + A_PR + q, prop + 5 * (size_t)e + q);
  if (q == 0) adj_cp8(rec + A_PR + 4, prop + 5 * (size_t)e + 4);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    adj_cp16(rec + A_UG + 6 * q + 2 * j, u + 6 * (size_t)nd + 2 * j);
    adj_cp16(rec + A_LG + 6 * q + 2 * j, lam + 6 * (size_t)nd + 2 * j);
  }
  adj_cp_commit();
}

// Persistent CTAs: batch b, b + gridDim.x, ... of 32 quads.  The gather of the NEXT batch (node ids one
// more batch ahead, in a register) streams into the other stage buffer with cp.async while this batch is
// differentiated, so the two-deep global-load chain cnct -> (crds, u, lam) is exposed only once per CTA.
// Warps never exchange data with each other (a quad's four lanes share a warp): __syncwarp only.
#ifndef JSSO_ADJ_MINB
#define JSSO_ADJ_MINB 2   // resident CTAs per SM the register allocation aims at (3 -> 168 registers)
#endif
template <bool WANT_PROP>
__global__ void __launch_bounds__(4 * ADJ_QUADS, JSSO_ADJ_MINB)
quad_adjoint_kernel(int n_quad, const double* __restrict__ crds, const int32_t* __restrict__ cnct,
                    const double* __restrict__ prop, const double* __restrict__ u,
                    const double* __restrict__ lam, double* __restrict__ corner, double* __restrict__ d_prop) {
  JSSO_DYN_SMEM(adj_sm);
  double* const sv = adj_sm + 2 * ADJ_QUADS * ADJ_ST;
  const int le = threadIdx.x >> 2, q = threadIdx.x & 3;
  const int n_batch = (n_quad + ADJ_QUADS - 1) / ADJ_QUADS;
  const int G = gridDim.x;
  int batch = blockIdx.x;
  if (batch >= n_batch) return;
  int nd_next = cnct[4 * min(batch * ADJ_QUADS + le, n_quad - 1) + q];
  adj_stage(adj_sm + le * ADJ_ST, q, nd_next, min(batch * ADJ_QUADS + le, n_quad - 1), crds, prop, u, lam);
  nd_next = (batch + G < n_batch) ? cnct[4 * min((batch + G) * ADJ_QUADS + le, n_quad - 1) + q] : 0;
  for (int it = 0; batch < n_batch; batch += G, ++it) {
  const bool has_next = batch + G < n_batch;
  if (has_next)
    adj_stage(adj_sm + ((it + 1) & 1) * (ADJ_QUADS * ADJ_ST) + le * ADJ_ST, q, nd_next, min((batch + G) * ADJ_QUADS + le, n_quad - 1), crds, prop, u, lam);
  if (batch + 2 * G < n_batch) nd_next = cnct[4 * min((batch + 2 * G) * ADJ_QUADS + le, n_quad - 1) + q];
  if (has_next) adj_cp_wait_1();
  else adj_cp_wait_all();
  __syncwarp();
  int e = batch * ADJ_QUADS + le;
  const bool valid = e < n_quad;
  if (!valid) e = n_quad - 1;
  const double* sg = adj_sm + (it & 1) * (ADJ_QUADS * ADJ_ST) + le * ADJ_ST;
  double* sm = sv + le * ADJ_LD;
  const double* UG = sg + A_UG;
  const double* LG = sg + A_LG;
  QuadFrame<double> f;
  {
    double P[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) P[k][c] = sg[A_P + 3 * k + c];
    quad_frame(P, f);
    double ug[6], lg[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) { ug[d] = UG[6 * q + d]; lg[d] = LG[6 * q + d]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      sm[A_UL + 6 * q + i] = f.R[i][0] * ug[0] + f.R[i][1] * ug[1] + f.R[i][2] * ug[2];
      sm[A_UL + 6 * q + 3 + i] = f.R[i][0] * ug[3] + f.R[i][1] * ug[4] + f.R[i][2] * ug[5];
      sm[A_LL + 6 * q + i] = f.R[i][0] * lg[0] + f.R[i][1] * lg[1] + f.R[i][2] * lg[2];
      sm[A_LL + 6 * q + 3 + i] = f.R[i][0] * lg[3] + f.R[i][1] * lg[4] + f.R[i][2] * lg[5];
    }
    // the frame is NOT kept in registers: the pull-back at the end recomputes dirCos from the staged
    // coordinates (same expressions), and the m12 section re-reads the local coordinates from here
    if (q == 0) {
      sm[A_XY] = f.x[0]; sm[A_XY + 1] = f.y[0]; sm[A_XY + 2] = f.x[1]; sm[A_XY + 3] = f.y[1];
      sm[A_XY + 4] = f.x[3]; sm[A_XY + 5] = f.y[3];
    }
  }
  __syncwarp();
  const double* UL = sm + A_UL;   // local u: per node [u, v, w, thx, thy, thz]
  const double* LL = sm + A_LL;
  const double* pr = sg + A_PR;   // t, E, nu, kx, ky (staged)
  QuadMat m;
  quad_mat(pr, m);
  QuadShear<double> sh;
  quad_shear(f, sh);
  // per-quad constants are parked in shared memory and RE-LOADED at the head of the later sections
  // (ADJ_RELOAD), so that they do not occupy 40 registers from here to the end of the kernel
  if (q == 0) {
    double* p = sm + A_SH;
    p[0] = sh.gry[0]; p[1] = sh.gry[1]; p[2] = sh.grx[0]; p[3] = sh.grx[1];
    p[4] = sh.gsy[0]; p[5] = sh.gsy[1]; p[6] = sh.gsx[0]; p[7] = sh.gsx[1];
    p[8] = sh.m11; p[9] = sh.m12; p[10] = sh.m22;
    double* t = sm + A_MT;
    t[0] = m.cm11; t[1] = m.cm12; t[2] = m.cm21; t[3] = m.cm22; t[4] = m.cm33;
    t[5] = m.D; t[6] = m.nu; t[7] = m.hb; t[8] = m.ks;
  }
  __syncwarp();
#define ADJ_RELOAD()                                                                        \
  do {                                                                                      \
    const double* p_ = sm + A_SH;                                                           \
    sh.gry[0] = p_[0]; sh.gry[1] = p_[1]; sh.grx[0] = p_[2]; sh.grx[1] = p_[3];             \
    sh.gsy[0] = p_[4]; sh.gsy[1] = p_[5]; sh.gsx[0] = p_[6]; sh.gsx[1] = p_[7];             \
    sh.m11 = p_[8]; sh.m12 = p_[9]; sh.m22 = p_[10];                                        \
    const double* t_ = sm + A_MT;                                                           \
    m.cm11 = t_[0]; m.cm12 = t_[1]; m.cm21 = t_[2]; m.cm22 = t_[3]; m.cm33 = t_[4];         \
    m.D = t_[5]; m.nu = t_[6]; m.hb = t_[7]; m.ks = t_[8];                                  \
  } while (0)
  QuadGp<double> g;
  quad_gp(f, q, g);
  const double r = JSSO_GP * node_r(q), s = JSSO_GP * node_s(q);

  // ---- pass 1: drilling stiffness = min |sum_gp diag| / 1000, first arg-min
  int im = 0;
  double dsel;
  {
    double dg[8];
    quad_diag_gp(sh, g, q, m, dg);
#pragma unroll
    for (int i = 0; i < 8; ++i) dg[i] = quad4_sum(dg[i]);
    double best = fabs(dg[0]);
    dsel = dg[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const double a = fabs(dg[i]);
      if (a < best) { best = a; im = i; dsel = dg[i]; }
    }
  }
  const double sgn = (dsel < 0.0) ? -1.0 : 1.0;
  const double krz = fabs(dsel) * 1e-3;
  double theta = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) theta += LL[6 * k + 5] * UL[6 * k + 5];
  const double dbar = sgn * theta * 1e-3;   // adjoint of the selected (summed) diagonal entry

  // ---- pass 2: this Gauss point
  double h0[4], h1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double dr = 0.25 * node_r(k) * (1.0 + s * node_s(k));
    const double ds = 0.25 * node_s(k) * (1.0 + r * node_r(k));
    h0[k] = g.ji[0] * dr + g.ji[1] * ds;
    h1[k] = g.ji[2] * dr + g.ji[3] * ds;
  }
  // gradients of the eight fields: a = u, b = v, c = theta_x, d = theta_y (lower: u, upper: lam)
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0, d0 = 0, d1 = 0;
  double A0 = 0, A1 = 0, B0 = 0, B1 = 0, C0 = 0, C1 = 0, D0 = 0, D1 = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double uu = UL[6 * k], uv = UL[6 * k + 1], ux = UL[6 * k + 3], uy = UL[6 * k + 4];
    const double lu = LL[6 * k], lv = LL[6 * k + 1], lx = LL[6 * k + 3], ly = LL[6 * k + 4];
    a0 += h0[k] * uu; a1 += h1[k] * uu; b0 += h0[k] * uv; b1 += h1[k] * uv;
    c0 += h0[k] * ux; c1 += h1[k] * ux; d0 += h0[k] * uy; d1 += h1[k] * uy;
    A0 += h0[k] * lu; A1 += h1[k] * lu; B0 += h0[k] * lv; B1 += h1[k] * lv;
    C0 += h0[k] * lx; C1 += h1[k] * lx; D0 += h0[k] * ly; D1 += h1[k] * ly;
  }
  const double exu = a0, eyu = b1, gxu = a1 + b0, kxu = -d0, kyu = c1, kzu = c0 - d1;
  const double exl = A0, eyl = B1, gxl = A1 + B0, kxl = -D0, kyl = C1, kzl = C0 - D1;
  // stresses: dPhi/d(strain of u) and dPhi/d(strain of lam)
  const double Nxu = m.cm11 * exl + m.cm21 * eyl, Nyu = m.cm12 * exl + m.cm22 * eyl, Nsu = m.cm33 * gxl;
  const double Nxl = m.cm11 * exu + m.cm12 * eyu, Nyl = m.cm21 * exu + m.cm22 * eyu, Nsl = m.cm33 * gxu;
  const double Mxu = m.D * (kxl + m.nu * kyl), Myu = m.D * (kyl + m.nu * kxl), Mzu = m.D * m.hb * kzl;
  const double Mxl = m.D * (kxu + m.nu * kyu), Myl = m.D * (kyu + m.nu * kxu), Mzl = m.D * m.hb * kzu;
  const double Phi = exl * Nxl + eyl * Nyl + gxl * Nsl + kxl * Mxl + kyl * Myl + kzl * Mzl;
  if (WANT_PROP) {
    // membrane / bending value sums of the closed-form property derivatives: summed over the Gauss points
    // now and parked in shared memory, so that the strains do not stay live to the end of the kernel
    const double Mxx = quad4_sum(g.det * exl * exu), Mxy = quad4_sum(g.det * exl * eyu);
    const double Myx = quad4_sum(g.det * eyl * exu), Myy = quad4_sum(g.det * eyl * eyu);
    const double Mss = quad4_sum(g.det * gxl * gxu);
    const double Bsum = quad4_sum(g.det * (kxl * kxu + kyl * kyu));
    const double Bnu = quad4_sum(g.det * (kxl * kyu + kyl * kxu));
    const double Bss = quad4_sum(g.det * kzl * kzu);
    if (q == 0) {
      double* ps = sm + A_PS;
      ps[0] = Mxx; ps[1] = Mxy; ps[2] = Myx; ps[3] = Myy; ps[4] = Mss; ps[5] = Bsum; ps[6] = Bnu; ps[7] = Bss;
    }
  }
  // S_phi = dPhi/d(grad phi) for the eight fields
  //   u: (Nx, Ns)  v: (Ns, Ny)  thx: (Mz, My)  thy: (-Mx, -Mz)
  const double Tx0 = Nxu * a0 + Nsu * b0 + Mzu * c0 - Mxu * d0 + Nxl * A0 + Nsl * B0 + Mzl * C0 - Mxl * D0;
  const double Tx1 = Nsu * a0 + Nyu * b0 + Myu * c0 - Mzu * d0 + Nsl * A0 + Nyl * B0 + Myl * C0 - Mzl * D0;
  const double Ty0 = Nxu * a1 + Nsu * b1 + Mzu * c1 - Mxu * d1 + Nxl * A1 + Nsl * B1 + Mzl * C1 - Mxl * D1;
  const double Ty1 = Nsu * a1 + Nyu * b1 + Myu * c1 - Mzu * d1 + Nsl * A1 + Nyl * B1 + Myl * C1 - Mzl * D1;
  double xb[4] = {0, 0, 0, 0}, yb[4] = {0, 0, 0, 0};   // adjoints of the local coordinates
  double Rb[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; // adjoint of dirCos
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    xb[k] += g.det * (h0[k] * Phi - (h0[k] * Tx0 + h1[k] * Tx1));
    yb[k] += g.det * (h1[k] * Phi - (h0[k] * Ty0 + h1[k] * Ty1));
    // nodal adjoints of the local components (u, v, thx, thy) of both vectors -> dirCos rows 0, 1
    const double ub = g.det * (Nxu * h0[k] + Nsu * h1[k]), vb = g.det * (Nsu * h0[k] + Nyu * h1[k]);
    const double txb = g.det * (Mzu * h0[k] + Myu * h1[k]), tyb = -g.det * (Mxu * h0[k] + Mzu * h1[k]);
    const double Ub = g.det * (Nxl * h0[k] + Nsl * h1[k]), Vb = g.det * (Nsl * h0[k] + Nyl * h1[k]);
    const double Txb = g.det * (Mzl * h0[k] + Myl * h1[k]), Tyb = -g.det * (Mxl * h0[k] + Mzl * h1[k]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Rb[0][c] += ub * UG[6 * k + c] + txb * UG[6 * k + 3 + c] + Ub * LG[6 * k + c] + Txb * LG[6 * k + 3 + c];
      Rb[1][c] += vb * UG[6 * k + c] + tyb * UG[6 * k + 3 + c] + Vb * LG[6 * k + c] + Tyb * LG[6 * k + 3 + c];
    }
  }
  ADJ_RELOAD();
  // ---- MITC4 shear
  // covariant edge strains of u and lam: index 0: A0 (edge 1-2), 1: A1 (edge 4-3), 2: B0 (1-4), 3: B1 (2-3)
  // edge e: nodes (n1, n2), w-coefficient +1/2 on n1 and -1/2 on n2, theta sums weighted by gy/gx
  double Eu[4], El[4];
  {
    const int n1[4] = {0, 3, 0, 1}, n2[4] = {1, 2, 3, 2};
    const double gy[4] = {sh.gry[0], sh.gry[1], sh.gsy[0], sh.gsy[1]};
    const double gx[4] = {sh.grx[0], sh.grx[1], sh.gsx[0], sh.gsx[1]};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      Eu[i] = 0.5 * (UL[6 * n1[i] + 2] - UL[6 * n2[i] + 2]) + gy[i] * (UL[6 * n1[i] + 3] + UL[6 * n2[i] + 3]) +
              gx[i] * (UL[6 * n1[i] + 4] + UL[6 * n2[i] + 4]);
      El[i] = 0.5 * (LL[6 * n1[i] + 2] - LL[6 * n2[i] + 2]) + gy[i] * (LL[6 * n1[i] + 3] + LL[6 * n2[i] + 3]) +
              gx[i] * (LL[6 * n1[i] + 4] + LL[6 * n2[i] + 4]);
    }
  }
  const double au = Eu[0] * (1.0 + s) + Eu[1] * (1.0 - s), al = El[0] * (1.0 + s) + El[1] * (1.0 - s);
  const double bu = Eu[2] * (1.0 + r) + Eu[3] * (1.0 - r), bl = El[2] * (1.0 + r) + El[3] * (1.0 - r);
  const double Crr = al * au, Crs = al * bu + bl * au, Css = bl * bu;
  double m12b = m.ks * g.prs * Crs;
  double Prrb = m.ks * sh.m11 * Crr, Prsb = m.ks * sh.m12 * Crs, Pssb = m.ks * sh.m22 * Css;
  const double aub = m.ks * (sh.m11 * g.prr * al + sh.m12 * g.prs * bl);
  const double alb = m.ks * (sh.m11 * g.prr * au + sh.m12 * g.prs * bu);
  const double bub = m.ks * (sh.m12 * g.prs * al + sh.m22 * g.pss * bl);
  const double blb = m.ks * (sh.m12 * g.prs * au + sh.m22 * g.pss * bu);
  double Eub[4] = {(1.0 + s) * aub, (1.0 - s) * aub, (1.0 + r) * bub, (1.0 - r) * bub};
  double Elb[4] = {(1.0 + s) * alb, (1.0 - s) * alb, (1.0 + r) * blb, (1.0 - r) * blb};
  double gyb[4] = {0, 0, 0, 0}, gxb[4] = {0, 0, 0, 0};   // adjoints of (gry0, gry1, gsy0, gsy1), (grx.., gsx..)

  // ---- drilling: derivative of the selected diagonal entry (this Gauss point's share)
  double kb_a = 0, kb_b = 0, ks_c = 0;
  {
    const int kd = im >> 1, comp = im & 1;
    double hk0 = h0[0], hk1 = h1[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) if (k == kd) { hk0 = h0[k]; hk1 = h1[k]; }
    const double pxx = g.det * hk0 * hk0, pyy = g.det * hk1 * hk1;
    // bending part D (pa + hb pb): comp 0 (theta_x): pa = pyy, pb = pxx; comp 1: pa = pxx, pb = pyy
    const double wxx = dbar * m.D * (comp ? 1.0 : m.hb), wyy = dbar * m.D * (comp ? m.hb : 1.0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // d pxx/dx_m = -det hk0^2 h_m0 ; d pxx/dy_m = det hk0 (h_m1 hk0 - 2 h_m0 hk1)
      // d pyy/dx_m = det hk1 (h_m0 hk1 - 2 h_m1 hk0) ; d pyy/dy_m = -det hk1^2 h_m1
      xb[k] += wxx * (-pxx * h0[k]) + wyy * (g.det * hk1 * (h0[k] * hk1 - 2.0 * h1[k] * hk0));
      yb[k] += wxx * (g.det * hk0 * (h1[k] * hk0 - 2.0 * h0[k] * hk1)) + wyy * (-pyy * h1[k]);
    }
    kb_a = comp ? pxx : pyy; kb_b = comp ? pyy : pxx;
    // shear part ks (m11 crr ax^2 + 2 m12 crs ax bx + m22 css bx^2)
    const double fr = 1.0 + s * node_s(kd), fs = 1.0 + r * node_r(kd);
    const int ir = (kd < 2) ? 0 : 1, is = (kd == 0 || kd == 3) ? 0 : 1;
    const double ax = comp ? sh.grx[ir] : sh.gry[ir], bx = comp ? sh.gsx[is] : sh.gsy[is];
    const double crr = g.prr * fr * fr, crs = g.prs * fr * fs, css = g.pss * fs * fs;
    ks_c = sh.m11 * crr * ax * ax + 2.0 * sh.m12 * crs * ax * bx + sh.m22 * css * bx * bx;
    const double w = dbar * m.ks;
    Prrb += w * sh.m11 * fr * fr * ax * ax;
    Prsb += w * sh.m12 * 2.0 * fr * fs * ax * bx;
    Pssb += w * sh.m22 * fs * fs * bx * bx;
    m12b += w * 2.0 * crs * ax * bx;
    const double axb = w * 2.0 * (sh.m11 * crr * ax + sh.m12 * crs * bx);
    const double bxb = w * 2.0 * (sh.m12 * crs * ax + sh.m22 * css * bx);
    // ax is gr?[ir] (edge index ir), bx is gs?[is] (edge index 2 + is)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (comp == 0) { if (i == ir) gyb[i] += axb; if (i == is) gyb[2 + i] += bxb; }
      else           { if (i == ir) gxb[i] += axb; if (i == is) gxb[2 + i] += bxb; }
    }
  }
  // ---- (Prr, Prs, Pss) -> J -> local coordinates
  {
    // recover J from J^-1: J = det * [[ji3, -ji1], [-ji2, ji0]]
    const double j00 = g.det * g.ji[3], j01 = -g.det * g.ji[1], j10 = -g.det * g.ji[2], j11 = g.det * g.ji[0];
    const double n1 = j10 * j10 + j11 * j11, n0 = j00 * j00 + j01 * j01;
    // one division and one rsqrt for 1/det, sqrt(n1 n0) and 1/sqrt(n1 n0)
    const double idet = 1.0 / g.det, iw = rsqrt(n1 * n0);
    const double qq = 0.25 * idet, w = (n1 * n0) * iw;
    double n1b = Prrb * qq, n0b = Pssb * qq;
    double qb = Prrb * n1 + Pssb * n0 + Prsb * w;
    const double wb = Prsb * qq;
    n1b += wb * n0 * (0.5 * iw);
    n0b += wb * n1 * (0.5 * iw);
    const double detb = -qb * qq * idet;
    const double j00b = 2.0 * j00 * n0b + detb * j11, j01b = 2.0 * j01 * n0b - detb * j10;
    const double j10b = 2.0 * j10 * n1b - detb * j01, j11b = 2.0 * j11 * n1b + detb * j00;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double dr = 0.25 * node_r(k) * (1.0 + s * node_s(k));
      const double ds = 0.25 * node_s(k) * (1.0 + r * node_r(k));
      xb[k] += dr * j00b + ds * j10b;
      yb[k] += dr * j01b + ds * j11b;
    }
  }
  ADJ_RELOAD();
  // ---- edge strains -> local vectors (dirCos rows 2, 0, 1) and edge geometry
  {
    const int n1[4] = {0, 3, 0, 1}, n2[4] = {1, 2, 3, 2};
    const double gy[4] = {sh.gry[0], sh.gry[1], sh.gsy[0], sh.gsy[1]};
    const double gx[4] = {sh.grx[0], sh.grx[1], sh.gsx[0], sh.gsx[1]};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = n1[i], t = n2[i];
      gyb[i] += Eub[i] * (UL[6 * p + 3] + UL[6 * t + 3]) + Elb[i] * (LL[6 * p + 3] + LL[6 * t + 3]);
      gxb[i] += Eub[i] * (UL[6 * p + 4] + UL[6 * t + 4]) + Elb[i] * (LL[6 * p + 4] + LL[6 * t + 4]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // w = e_z . ut  (row 2), theta_x = e_x . ur (row 0), theta_y = e_y . ur (row 1)
        Rb[2][c] += 0.5 * (Eub[i] * (UG[6 * p + c] - UG[6 * t + c]) + Elb[i] * (LG[6 * p + c] - LG[6 * t + c]));
        Rb[0][c] += gy[i] * (Eub[i] * (UG[6 * p + 3 + c] + UG[6 * t + 3 + c]) + Elb[i] * (LG[6 * p + 3 + c] + LG[6 * t + 3 + c]));
        Rb[1][c] += gx[i] * (Eub[i] * (UG[6 * p + 3 + c] + UG[6 * t + 3 + c]) + Elb[i] * (LG[6 * p + 3 + c] + LG[6 * t + 3 + c]));
      }
      // gy = -(y_p - y_t)/4, gx = (x_p - x_t)/4
      yb[p] -= 0.25 * gyb[i]; yb[t] += 0.25 * gyb[i];
      xb[p] += 0.25 * gxb[i]; xb[t] -= 0.25 * gxb[i];
    }
  }
  // ---- m12 = (|ry||sy| - rx sx)/(nr ns) -> local coordinates
  {
    const double x0 = sm[A_XY], y0 = sm[A_XY + 1], x1 = sm[A_XY + 2], y1 = sm[A_XY + 3], x3 = sm[A_XY + 4], y3 = sm[A_XY + 5];
    const double rx = ((x0 + x3) - x1) * 0.5, ry = ((y0 + y3) - y1) * 0.5;   // node 3 (index 2) is the origin
    const double sx = ((x0 + x1) - x3) * 0.5, sy = ((y0 + y1) - y3) * 0.5;
    const double inr = rsqrt(rx * rx + ry * ry), ins = rsqrt(sx * sx + sy * sy);   // 1/|r|, 1/|s|
    const double sig = ((ry < 0.0) != (sy < 0.0)) ? -1.0 : 1.0;
    const double Nb = m12b * (inr * ins), nrb = -m12b * sh.m12 * inr, nsb = -m12b * sh.m12 * ins;
    const double rxb = -Nb * sx + nrb * rx * inr, ryb = Nb * sig * sy + nrb * ry * inr;
    const double sxb = -Nb * rx + nsb * sx * ins, syb = Nb * sig * ry + nsb * sy * ins;
    xb[0] += 0.5 * (rxb + sxb); xb[1] += 0.5 * (sxb - rxb); xb[3] += 0.5 * (rxb - sxb);
    yb[0] += 0.5 * (ryb + syb); yb[1] += 0.5 * (syb - ryb); yb[3] += 0.5 * (ryb - syb);
  }
  // ---- drilling: theta_z = e_z . ur  (row 2), shared by the four lanes -> count it once (lane 0)
  if (q == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        Rb[2][c] += krz * (LL[6 * k + 5] * UG[6 * k + 3 + c] + UL[6 * k + 5] * LG[6 * k + 3 + c]);
  }
  // ---- sum the four Gauss points
  xb[0] = quad4_sum(xb[0]); xb[1] = quad4_sum(xb[1]); xb[3] = quad4_sum(xb[3]);
  yb[0] = quad4_sum(yb[0]); yb[1] = quad4_sum(yb[1]); yb[3] = quad4_sum(yb[3]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) Rb[i][c] = quad4_sum(Rb[i][c]);
  // ---- local coordinates x_k = v3k . e_x, y_k = v3k . e_y  (k = 0, 1, 3; node 3 is the origin)
  double v31[3], v32[3], v34[3], v42[3], vb31[3], vb32[3], vb34[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // coordinates re-read from the gather stage (not kept in registers across the kernel)
    const double p0 = sg[A_P + c], p1 = sg[A_P + 3 + c], p2 = sg[A_P + 6 + c], p3 = sg[A_P + 9 + c];
    v31[c] = p0 - p2; v32[c] = p1 - p2;
    v34[c] = p3 - p2; v42[c] = p1 - p3;
    Rb[0][c] += xb[0] * v31[c] + xb[1] * v32[c] + xb[3] * v34[c];
    Rb[1][c] += yb[0] * v31[c] + yb[1] * v32[c] + yb[3] * v34[c];
  }
  // ---- frame: e_x = v31/|v31|, z_r = v31 x v42, y_r = z_r x v31, e_y = y_r/|y_r|, e_z = z_r/|z_r|
  // (dirCos recomputed here exactly as quad_frame does, instead of living in registers since the start)
  double zr[3] = {v31[1] * v42[2] - v31[2] * v42[1], v31[2] * v42[0] - v31[0] * v42[2],
                  v31[0] * v42[1] - v31[1] * v42[0]};
  double yr[3] = {zr[1] * v31[2] - zr[2] * v31[1], zr[2] * v31[0] - zr[0] * v31[2],
                  zr[0] * v31[1] - zr[1] * v31[0]};
  const double inx = rsqrt(v31[0] * v31[0] + v31[1] * v31[1] + v31[2] * v31[2]);
  const double iny = rsqrt(yr[0] * yr[0] + yr[1] * yr[1] + yr[2] * yr[2]);
  const double inz = rsqrt(zr[0] * zr[0] + zr[1] * zr[1] + zr[2] * zr[2]);
  double R0[3], R1[3], R2[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    R0[c] = v31[c] * inx; R1[c] = yr[c] * iny; R2[c] = zr[c] * inz;
    vb31[c] = xb[0] * R0[c] + yb[0] * R1[c];
    vb32[c] = xb[1] * R0[c] + yb[1] * R1[c];
    vb34[c] = xb[3] * R0[c] + yb[3] * R1[c];
  }
  const double dx = R0[0] * Rb[0][0] + R0[1] * Rb[0][1] + R0[2] * Rb[0][2];
  const double dy = R1[0] * Rb[1][0] + R1[1] * Rb[1][1] + R1[2] * Rb[1][2];
  const double dz = R2[0] * Rb[2][0] + R2[1] * Rb[2][1] + R2[2] * Rb[2][2];
  double yrb[3], zrb[3], vb42[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    vb31[c] += (Rb[0][c] - R0[c] * dx) * inx;
    yrb[c] = (Rb[1][c] - R1[c] * dy) * iny;
    zrb[c] = (Rb[2][c] - R2[c] * dz) * inz;
  }
  // y_r = z_r x v31:  z_r_bar += v31 x y_r_bar,  v31_bar += y_r_bar x z_r
  zrb[0] += v31[1] * yrb[2] - v31[2] * yrb[1];
  zrb[1] += v31[2] * yrb[0] - v31[0] * yrb[2];
  zrb[2] += v31[0] * yrb[1] - v31[1] * yrb[0];
  vb31[0] += yrb[1] * zr[2] - yrb[2] * zr[1];
  vb31[1] += yrb[2] * zr[0] - yrb[0] * zr[2];
  vb31[2] += yrb[0] * zr[1] - yrb[1] * zr[0];
  // z_r = v31 x v42:  v31_bar += v42 x z_r_bar,  v42_bar = z_r_bar x v31
  vb31[0] += v42[1] * zrb[2] - v42[2] * zrb[1];
  vb31[1] += v42[2] * zrb[0] - v42[0] * zrb[2];
  vb31[2] += v42[0] * zrb[1] - v42[1] * zrb[0];
  vb42[0] = zrb[1] * v31[2] - zrb[2] * v31[1];
  vb42[1] = zrb[2] * v31[0] - zrb[0] * v31[2];
  vb42[2] = zrb[0] * v31[1] - zrb[1] * v31[0];
  if (valid && corner) {
    // node 0: v31; node 1: v32 + v42; node 2: -(v31 + v32 + v34); node 3: v34 - v42.  dL = -de.
    double* cdst = corner + 3 * ((size_t)e * 4 + q);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double p0 = vb31[c], p1 = vb32[c] + vb42[c], p3 = vb34[c] - vb42[c];
      const double p2 = -(vb31[c] + vb32[c] + vb34[c]);
      cdst[c] = -((q == 0) ? p0 : (q == 1) ? p1 : (q == 2) ? p2 : p3);
    }
  }
  if (WANT_PROP) {
    ADJ_RELOAD();
    // value accumulators for the closed-form property derivatives (summed over Gauss points)
    const double Sh = quad4_sum(sh.m11 * g.prr * Crr + sh.m12 * g.prs * Crs + sh.m22 * g.pss * Css);
    const double kba = quad4_sum(kb_a), kbb = quad4_sum(kb_b), ksc = quad4_sum(ks_c);
    if (valid && q == 0 && d_prop) {
      const double* ps = sm + A_PS;   // written by this same lane above
      const double Mxx = ps[0], Mxy = ps[1], Myx = ps[2], Myy = ps[3], Mss = ps[4], Bsum = ps[5], Bnu = ps[6], Bss = ps[7];
      const double th = pr[0], E = pr[1], nu = pr[2], kx = pr[3], ky = pr[4];
      const double e_m = Mxx * m.cm11 + Mxy * m.cm12 + Myx * m.cm21 + Myy * m.cm22 + Mss * m.cm33;
      const double bcon = Bsum + Bnu * nu + Bss * m.hb;
      const double e_b = bcon * m.D, e_s = Sh * m.ks;
      const double dsv = m.D * (kba + m.hb * kbb) + m.ks * ksc;   // the selected diagonal entry
      const double sg = sgn * 1e-3 * theta;                       // e_d = sg * dsv
      // the four reciprocals 1/t, 1/E, 1/(1+nu), 1/(1-nu^2) from ONE division
      const double tE = th * E, cd = (1.0 + nu) * (1.0 - nu);
      const double rr = 1.0 / (tE * cd);
      const double pre = rr * tE, inv_tE = rr * cd, ip = pre * (1.0 - nu), dpre = 2.0 * nu * pre * pre;
      const double inv_th = inv_tE * E, inv_E = inv_tE * th;
      const double de_dt = ((e_m + 3.0 * e_b + e_s) + sg * (3.0 * m.D * (kba + m.hb * kbb) + m.ks * ksc)) * inv_th;
      const double de_dE = (e_m + e_b + e_s + sg * dsv) * inv_E;
      const double dem = tE * (dpre * (kx * (Mxx + nu * Mxy) + ky * (nu * Myx + Myy)) + pre * (kx * Mxy + ky * Myx)) -
                         tE * Mss * (0.5 * ip * ip);
      const double dD = m.D * 2.0 * nu * pre, dks = -m.ks * ip;
      const double deb = dD * bcon + m.D * (Bnu - 0.5 * Bss);
      const double ddsel = dD * (kba + m.hb * kbb) - 0.5 * m.D * kbb + dks * ksc;
      double* o = d_prop + 5 * (size_t)e;
      o[0] = -de_dt; o[1] = -de_dE; o[2] = -(dem + deb + dks * Sh + sg * ddsel);
      o[3] = -(tE * pre * (Mxx + nu * Mxy)); o[4] = -(tE * pre * (nu * Myx + Myy));
    }
  }
  __syncwarp();   // this batch's stage and local vectors are dead: the next iterations may overwrite them
  }  // batch loop
#undef ADJ_RELOAD
}

// e = lam_e^T K_e u_e for a beam-column, K_e = T^T K_local T (orthonormal T).
// geo[4] returns the geometry factors multiplying (E A, E Iz, E Iy, G J).
template <class S>
__device__ inline S beam_bilinear(const S P[2][3], const double* prop, const double* ue, const double* le,
                                  double* geo, bool* degenerate) {
  S R[3][3], L;
  *degenerate = beam_dircos(P, R, L);
  S ut[2][3], ur[2][3], lt[2][3], lr[2][3];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      ut[k][i] = R[i][0] * ue[6 * k] + R[i][1] * ue[6 * k + 1] + R[i][2] * ue[6 * k + 2];
      ur[k][i] = R[i][0] * ue[6 * k + 3] + R[i][1] * ue[6 * k + 4] + R[i][2] * ue[6 * k + 5];
      lt[k][i] = R[i][0] * le[6 * k] + R[i][1] * le[6 * k + 1] + R[i][2] * le[6 * k + 2];
      lr[k][i] = R[i][0] * le[6 * k + 3] + R[i][1] * le[6 * k + 4] + R[i][2] * le[6 * k + 5];
    }
  const S iL = 1.0 / L, iL2 = iL * iL, iL3 = iL2 * iL;
  // axial and torsion
  const S g_ax = (lt[0][0] - lt[1][0]) * (ut[0][0] - ut[1][0]) * iL;
  const S g_t = (lr[0][0] - lr[1][0]) * (ur[0][0] - ur[1][0]) * iL;
  // bending about local z (dofs v = t[.][1], theta_z = r[.][2]):
  //   12/L^3 dv dv + 6/L^2 (dv (thz1+thz2) + (thz1+thz2) dv) + 4/L,2/L rotations
  const S dvu = ut[0][1] - ut[1][1], dvl = lt[0][1] - lt[1][1];
  const S szu = ur[0][2] + ur[1][2], szl = lr[0][2] + lr[1][2];
  const S g_z = dvl * dvu * iL3 * 12.0 + (dvl * szu + szl * dvu) * iL2 * 6.0 +
                (lr[0][2] * ur[0][2] + lr[1][2] * ur[1][2]) * iL * 4.0 +
                (lr[0][2] * ur[1][2] + lr[1][2] * ur[0][2]) * iL * 2.0;
  // bending about local y (dofs w = t[.][2], theta_y = r[.][1]); the 6 E Iy / L^2 terms change sign
  const S dwu = ut[0][2] - ut[1][2], dwl = lt[0][2] - lt[1][2];
  const S syu = ur[0][1] + ur[1][1], syl = lr[0][1] + lr[1][1];
  const S g_y = dwl * dwu * iL3 * 12.0 - (dwl * syu + syl * dwu) * iL2 * 6.0 +
                (lr[0][1] * ur[0][1] + lr[1][1] * ur[1][1]) * iL * 4.0 +
                (lr[0][1] * ur[1][1] + lr[1][1] * ur[0][1]) * iL * 2.0;
  const double E = prop[0], G = prop[1], Iy = prop[2], Iz = prop[3], J = prop[4], A = prop[5];
  if (geo) { geo[0] = val(g_ax); geo[1] = val(g_z); geo[2] = val(g_y); geo[3] = val(g_t); }
  return g_ax * (E * A) + g_z * (E * Iz) + g_y * (E * Iy) + g_t * (G * J);
}

// thread = (beam, node a)
__global__ void __launch_bounds__(128)
beam_adjoint_kernel(int n_beam, const double* __restrict__ crds, const int32_t* __restrict__ cnct,
                    const double* __restrict__ prop, const double* __restrict__ u,
                    const double* __restrict__ lam, double* __restrict__ corner, double* __restrict__ d_prop,
                    int* flags) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = t >> 1, a = t & 1;
  if (e >= n_beam) return;
  D3 P[2][3];
  double ue[12], le[12];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int nd = cnct[2 * e + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      P[k][c] = mk<3>(crds[3 * (size_t)nd + c]);
      if (k == a) P[k][c].d[c] = 1.0;
    }
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      ue[6 * k + d] = u[6 * (size_t)nd + d];
      le[6 * k + d] = lam[6 * (size_t)nd + d];
    }
  }
  const double* pr = prop + 6 * (size_t)e;
  double geo[4];
  bool deg;
  const D3 en = beam_bilinear<D3>(P, pr, ue, le, geo, &deg);
  if (deg) atomicOr(flags, 2);
  if (corner) {
    double* c = corner + 3 * (size_t)t;
    c[0] = -en.d[0]; c[1] = -en.d[1]; c[2] = -en.d[2];
  }
  if (a == 0 && d_prop) {
    const double E = pr[0], G = pr[1], Iy = pr[2], Iz = pr[3], J = pr[4], A = pr[5];
    double* o = d_prop + 6 * (size_t)e;
    o[0] = -(A * geo[0] + Iz * geo[1] + Iy * geo[2]);   // d/dE
    o[1] = -(J * geo[3]);                               // d/dG
    o[2] = -(E * geo[2]);                               // d/dIy
    o[3] = -(E * geo[1]);                               // d/dIz
    o[4] = -(G * geo[3]);                               // d/dJ
    o[5] = -(E * geo[0]);                               // d/dA
  }
}

// d_crds[node] = sum over incident corners, in list order (beams first, then quads)
__global__ void __launch_bounds__(256)
node_gather_kernel(int n_node, int n_quad, const int32_t* __restrict__ inc_ptr, const int32_t* __restrict__ inc,
                   const double* __restrict__ corner_q, const double* __restrict__ corner_b,
                   double* __restrict__ d_crds) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n_node) return;
  const int nd = t / 3, c = t - 3 * nd;
  double s = 0.0;
  for (int i = inc_ptr[nd]; i < inc_ptr[nd + 1]; ++i) {
    const int code = inc[i];
    const int el = code >> 2, a = code & 3;
    if (el < n_quad) s += corner_q[3 * ((size_t)el * 4 + a) + c];
    else s += corner_b[3 * ((size_t)(el - n_quad) * 2 + a) + c];
  }
  d_crds[t] = s;
}

}  // namespace jsso
