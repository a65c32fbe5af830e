// Fused adjoint sensitivity kernels.
//
// The reference obtains dL/dx by (i) the VJP of -(K u - f) w.r.t. every stored COO
// entry at cotangent lam, i.e. W_e[a,b] = -lam[dof_e[a]] u[dof_e[b]]
// (JaxSSO/solver.py:157-166, 239-248), then (ii) XLA reverse mode through
// sum_duplicates / sort / concatenate and vmap(element_K_*) including its batched LU
// solves.  Mathematically that is
//     dL/dx = - d/dx [ lam_e^T K_e(x) u_e ]   summed over elements,
// so these kernels never form K_e, dK_e or the nse-long cotangent: each thread
// evaluates the scalar bilinear form e(x) = lam_e^T K_e(x) u_e through the element's
// strains with forward-mode duals seeded on the three coordinates of ONE node of the
// element (thread = (element, node)), using the same templated geometry code as the
// stiffness kernels.  Property derivatives are closed-form from the split energies.
// Nodal gradients are then gathered per node from the per-corner partials in a fixed
// order (no atomics).
#pragma once
#include "jsso_elem.cuh"

namespace jsso {

using D3 = Dual<3>;

struct QuadEnergyParts {      // value parts needed for the closed-form property derivatives
  double Mxx, Mxy, Myx, Myy, Mss;   // membrane contractions (sum_gp detJ ...)
  double Bsum, Bnu, Bss;            // bending: kxx+kyy terms, nu cross terms, twist term
  double Sh;                        // shear contraction
  double theta;                     // sum_k lam_thz,k u_thz,k (local)
  double kb_a, kb_b, ks_c;          // arg-min diagonal: |D (kb_a + hb kb_b) + ks ks_c|
  double sign;                      // sign of that diagonal entry
};

// e = lam_e^T K_e u_e for one quad.  P: nodal coordinates (S), prop: t,E,nu,kx,ky,
// ue/le: the 24 global dofs of u and lam.
template <class S>
__device__ inline S quad_bilinear(const S P[4][3], const double* prop, const double* ue, const double* le,
                                  QuadEnergyParts* parts) {
  QuadFrame<S> f;
  quad_frame(P, f);
  QuadShear<S> sh;
  quad_shear(f, sh);
  QuadMat m;
  quad_mat(prop, m);
  // local vectors  (T u)_k = (R u_k[0:3], R u_k[3:6])
  S ut[4][3], ur[4][3], lt[4][3], lr[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      ut[k][i] = f.R[i][0] * ue[6 * k] + f.R[i][1] * ue[6 * k + 1] + f.R[i][2] * ue[6 * k + 2];
      ur[k][i] = f.R[i][0] * ue[6 * k + 3] + f.R[i][1] * ue[6 * k + 4] + f.R[i][2] * ue[6 * k + 5];
      lt[k][i] = f.R[i][0] * le[6 * k] + f.R[i][1] * le[6 * k + 1] + f.R[i][2] * le[6 * k + 2];
      lr[k][i] = f.R[i][0] * le[6 * k + 3] + f.R[i][1] * le[6 * k + 4] + f.R[i][2] * le[6 * k + 5];
    }
  // covariant edge shear strains (gp independent): A+/A- along r on edges 1-2 / 4-3,
  // B+/B- along s on edges 1-4 / 2-3
  S Au[2], Al[2], Bu[2], Bl[2];
  Au[0] = (ut[0][2] - ut[1][2]) * 0.5 + sh.gry[0] * (ur[0][0] + ur[1][0]) + sh.grx[0] * (ur[0][1] + ur[1][1]);
  Au[1] = (ut[3][2] - ut[2][2]) * 0.5 + sh.gry[1] * (ur[2][0] + ur[3][0]) + sh.grx[1] * (ur[2][1] + ur[3][1]);
  Bu[0] = (ut[0][2] - ut[3][2]) * 0.5 + sh.gsy[0] * (ur[0][0] + ur[3][0]) + sh.gsx[0] * (ur[0][1] + ur[3][1]);
  Bu[1] = (ut[1][2] - ut[2][2]) * 0.5 + sh.gsy[1] * (ur[1][0] + ur[2][0]) + sh.gsx[1] * (ur[1][1] + ur[2][1]);
  Al[0] = (lt[0][2] - lt[1][2]) * 0.5 + sh.gry[0] * (lr[0][0] + lr[1][0]) + sh.grx[0] * (lr[0][1] + lr[1][1]);
  Al[1] = (lt[3][2] - lt[2][2]) * 0.5 + sh.gry[1] * (lr[2][0] + lr[3][0]) + sh.grx[1] * (lr[2][1] + lr[3][1]);
  Bl[0] = (lt[0][2] - lt[3][2]) * 0.5 + sh.gsy[0] * (lr[0][0] + lr[3][0]) + sh.gsx[0] * (lr[0][1] + lr[3][1]);
  Bl[1] = (lt[1][2] - lt[2][2]) * 0.5 + sh.gsy[1] * (lr[1][0] + lr[2][0]) + sh.gsx[1] * (lr[1][1] + lr[2][1]);

  const S zero = Lift<S>::from(0.0);
  S Mxx = zero, Mxy = zero, Myx = zero, Myy = zero, Mss = zero;
  S Bsum = zero, Bnu = zero, Bss = zero, Sh = zero;
  S diag[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) diag[i] = zero;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    QuadGp<S> g;
    quad_gp(f, q, g);
    const double r = JSSO_GP * node_r(q), s = JSSO_GP * node_s(q);
    S exu = zero, eyu = zero, gxu = zero, exl = zero, eyl = zero, gxl = zero;     // membrane
    S kxu = zero, kyu = zero, kzu = zero, kxl = zero, kyl = zero, kzl = zero;     // curvatures
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double dr = 0.25 * node_r(k) * (1.0 + s * node_s(k));
      const double ds = 0.25 * node_s(k) * (1.0 + r * node_r(k));
      const S h0 = g.ji[0] * dr + g.ji[1] * ds, h1 = g.ji[2] * dr + g.ji[3] * ds;
      exu = exu + h0 * ut[k][0]; eyu = eyu + h1 * ut[k][1]; gxu = gxu + h1 * ut[k][0] + h0 * ut[k][1];
      exl = exl + h0 * lt[k][0]; eyl = eyl + h1 * lt[k][1]; gxl = gxl + h1 * lt[k][0] + h0 * lt[k][1];
      kxu = kxu - h0 * ur[k][1]; kyu = kyu + h1 * ur[k][0]; kzu = kzu + h0 * ur[k][0] - h1 * ur[k][1];
      kxl = kxl - h0 * lr[k][1]; kyl = kyl + h1 * lr[k][0]; kzl = kzl + h0 * lr[k][0] - h1 * lr[k][1];
    }
    Mxx = Mxx + g.det * exl * exu; Mxy = Mxy + g.det * exl * eyu;
    Myx = Myx + g.det * eyl * exu; Myy = Myy + g.det * eyl * eyu;
    Mss = Mss + g.det * gxl * gxu;
    Bsum = Bsum + g.det * (kxl * kxu + kyl * kyu);
    Bnu = Bnu + g.det * (kxl * kyu + kyl * kxu);
    Bss = Bss + g.det * kzl * kzu;
    const S au = Au[0] * (1.0 + s) + Au[1] * (1.0 - s), al = Al[0] * (1.0 + s) + Al[1] * (1.0 - s);
    const S bu = Bu[0] * (1.0 + r) + Bu[1] * (1.0 - r), bl = Bl[0] * (1.0 + r) + Bl[1] * (1.0 - r);
    Sh = Sh + sh.m11 * g.prr * al * au + sh.m12 * g.prs * (al * bu + bl * au) + sh.m22 * g.pss * bl * bu;
    S dq[8];
    quad_diag_gp(sh, g, q, m, dq);
#pragma unroll
    for (int i = 0; i < 8; ++i) diag[i] = diag[i] + dq[i];
  }
  // drilling stiffness: first arg-min of |diag| (element.py:978; XLA's min VJP sends the
  // cotangent to the arg-min, and to the first one unless two entries are bitwise equal)
  int im = 0;
  double best = fabs(val(diag[0]));
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    const double a = fabs(val(diag[i]));
    if (a < best) { best = a; im = i; }
  }
  S dsel = diag[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) if (i == im) dsel = diag[i];
  const S krz = sabs(dsel) * 1e-3;
  S theta = zero;
#pragma unroll
  for (int k = 0; k < 4; ++k) theta = theta + lr[k][2] * ur[k][2];

  const S e_m = Mxx * m.cm11 + Mxy * m.cm12 + Myx * m.cm21 + Myy * m.cm22 + Mss * m.cm33;
  const S e_b = (Bsum + Bnu * m.nu + Bss * m.hb) * m.D;
  const S e_s = Sh * m.ks;
  const S e_d = krz * theta;
  if (parts) {
    parts->Mxx = val(Mxx); parts->Mxy = val(Mxy); parts->Myx = val(Myx); parts->Myy = val(Myy);
    parts->Mss = val(Mss); parts->Bsum = val(Bsum); parts->Bnu = val(Bnu); parts->Bss = val(Bss);
    parts->Sh = val(Sh); parts->theta = val(theta);
    parts->sign = (val(dsel) < 0.0) ? -1.0 : 1.0;
    // split the selected diagonal value into its D- and ks-proportional parts: recompute
    // with unit material constants (cheap, value-only)
    QuadMat mb = m; mb.D = 1.0; mb.ks = 0.0;
    QuadMat mh = m; mh.D = 1.0; mh.ks = 0.0; mh.hb = 0.0;
    QuadMat msh = m; msh.D = 0.0; msh.ks = 1.0;
    double db = 0, dh = 0, dsv = 0;
    QuadFrame<double> fv; QuadShear<double> shv;
#pragma unroll
    for (int k = 0; k < 4; ++k) { fv.x[k] = val(f.x[k]); fv.y[k] = val(f.y[k]); }
    quad_shear(fv, shv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      QuadGp<double> gv; quad_gp(fv, q, gv);
      double t1[8], t2[8], t3[8];
      quad_diag_gp(shv, gv, q, mb, t1); quad_diag_gp(shv, gv, q, mh, t2); quad_diag_gp(shv, gv, q, msh, t3);
#pragma unroll
      for (int i = 0; i < 8; ++i) if (i == im) { db += t1[i]; dh += t2[i]; dsv += t3[i]; }
    }
    // diag = D (kb_a + hb kb_b) + ks ks_c  with kb_a = dh, kb_b = (db - dh)/hb
    parts->kb_a = dh; parts->kb_b = (m.hb != 0.0) ? (db - dh) / m.hb : 0.0; parts->ks_c = dsv;
  }
  return e_m + e_b + e_s + e_d;
}

// thread = (quad, node a): writes the three coordinate partials of corner (e,a) and,
// from a == 0, the five property partials.
__global__ void __launch_bounds__(128)
quad_adjoint_kernel(int n_quad, const double* __restrict__ crds, const int32_t* __restrict__ cnct,
                    const double* __restrict__ prop, const double* __restrict__ u,
                    const double* __restrict__ lam, double* __restrict__ corner, double* __restrict__ d_prop) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = t >> 2, a = t & 3;
  if (e >= n_quad) return;
  D3 P[4][3];
  double ue[24], le[24];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int nd = cnct[4 * e + k];
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
  const double* pr = prop + 5 * (size_t)e;
  QuadEnergyParts parts;
  const D3 en = quad_bilinear<D3>(P, pr, ue, le, (a == 0 && d_prop) ? &parts : nullptr);
  if (corner) {
    double* c = corner + 3 * (size_t)t;
    c[0] = -en.d[0]; c[1] = -en.d[1]; c[2] = -en.d[2];
  }
  if (a == 0 && d_prop) {
    const double th = pr[0], E = pr[1], nu = pr[2], kx = pr[3], ky = pr[4];
    QuadMat m; quad_mat(pr, m);
    const double e_m = parts.Mxx * m.cm11 + parts.Mxy * m.cm12 + parts.Myx * m.cm21 + parts.Myy * m.cm22 +
                       parts.Mss * m.cm33;
    const double bcon = parts.Bsum + parts.Bnu * nu + parts.Bss * m.hb;
    const double e_b = bcon * m.D, e_s = parts.Sh * m.ks;
    const double dsel = m.D * (parts.kb_a + m.hb * parts.kb_b) + m.ks * parts.ks_c;   // signed diagonal
    const double sg = parts.sign * 1e-3 * parts.theta;                                 // e_d = sg * dsel
    const double one_m_nu2 = 1.0 - nu * nu;
    // d/dt: membrane ~ t, bending ~ t^3, shear ~ t
    const double de_dt = (e_m + 3.0 * e_b + e_s) / th +
                         sg * (3.0 * m.D * (parts.kb_a + m.hb * parts.kb_b) + m.ks * parts.ks_c) / th;
    const double de_dE = (e_m + e_b + e_s + sg * dsel) / E;
    // d/dnu of the prefactors
    const double pre = 1.0 / one_m_nu2, dpre = 2.0 * nu * pre * pre;
    const double tE = th * E;
    const double dem = tE * (dpre * (kx * (parts.Mxx + nu * parts.Mxy) + ky * (nu * parts.Myx + parts.Myy)) +
                             pre * (kx * parts.Mxy + ky * parts.Myx)) -
                       tE * parts.Mss / (2.0 * (1.0 + nu) * (1.0 + nu));
    const double dD = m.D * 2.0 * nu / one_m_nu2;
    const double deb = dD * bcon + m.D * (parts.Bnu - 0.5 * parts.Bss);
    const double dks = -m.ks / (1.0 + nu);
    const double des = dks * parts.Sh;
    const double ddsel = dD * (parts.kb_a + m.hb * parts.kb_b) - 0.5 * m.D * parts.kb_b + dks * parts.ks_c;
    const double de_dnu = dem + deb + des + sg * ddsel;
    const double de_dkx = tE * pre * (parts.Mxx + nu * parts.Mxy);
    const double de_dky = tE * pre * (nu * parts.Myx + parts.Myy);
    double* o = d_prop + 5 * (size_t)e;
    o[0] = -de_dt; o[1] = -de_dE; o[2] = -de_dnu; o[3] = -de_dkx; o[4] = -de_dky;
  }
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
