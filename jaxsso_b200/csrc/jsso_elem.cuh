// Element mathematics of the JaxSSO hot path, written once as scalar-generic
// device code: S = double for stiffness generation, S = Dual<N> (forward-mode
// tangents) for the adjoint sensitivity kernel.
//
// Reference behaviour restated (not copied): JaxSSO/element.py
//   beam-column  T :66-105, K_local :107-128, T^-1 K T :130-139
//   MITC4 quad   loc_crds :488-539, T :643-696, J :698-709, B_kappa :711-733,
//                B_gamma_MITC4 :735-801, B_m :803-818, Cb/Cs/Cm :820-875,
//                k_b :923-995, k_m :1035-1071, element_K_quad :1073-1084
//
// The quad is evaluated in a factored form that never builds a B matrix:
//   * membrane and bending blocks of a node pair (a,b) only need the four sums
//     p_xy = sum_gp detJ dH_x,a dH_y,b;
//   * the MITC4 shear rows are (1 + s s_k) g^r_k and (1 + r r_k) g^s_k with
//     gp-independent 3-vectors g^r_k, g^s_k, so the pair block needs four more
//     gp sums (c_rr, c_rs, c_sr, c_ss);
//   * gr = |J row 1| / (2 detJ), gs = |J row 0| / (2 detJ) (the reference's
//     (Cx + r Bx, Cy + r By) is 4x the second Jacobian row, element.py:749-771);
//   * T^-1 K T is evaluated as R^T S R on the sparse 3x3 sub-blocks (T is
//     orthonormal, so this equals the reference's LU solve to rounding).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// JSSO_EMU is defined only by the CPU test harness (tests/emu: the kernels compiled by g++ against a SIMT
// emulator to check their logic without a GPU).  The product build (nvcc) never defines it.
#ifdef JSSO_EMU
#define JSSO_DYN_SMEM(name) double* name = (double*)emu::dyn_smem()
#else
#define JSSO_DYN_SMEM(name) extern __shared__ __align__(16) double name[]
#endif

namespace jsso {

// ------------------------------------------------------------------ duals
template <int N>
struct Dual {
  double v;
  double d[N];
};

template <int N> __host__ __device__ inline Dual<N> mk(double v) {
  Dual<N> r; r.v = v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = 0.0;
  return r;
}
template <int N> __host__ __device__ inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N> __host__ __device__ inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N> __host__ __device__ inline Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N> __host__ __device__ inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N> __host__ __device__ inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <int N> __host__ __device__ inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> __host__ __device__ inline Dual<N> operator+(double b, const Dual<N>& a) { Dual<N> r = a; r.v += b; return r; }
template <int N> __host__ __device__ inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> __host__ __device__ inline Dual<N> operator-(double b, const Dual<N>& a) { Dual<N> r = -a; r.v += b; return r; }
template <int N> __host__ __device__ inline Dual<N> operator*(const Dual<N>& a, double b) {
  Dual<N> r; r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
  return r;
}
template <int N> __host__ __device__ inline Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
template <int N> __host__ __device__ inline Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> __host__ __device__ inline Dual<N> operator/(double a, const Dual<N>& b) { return mk<N>(a) / b; }
template <int N> __host__ __device__ inline Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) { a = a + b; return a; }

__host__ __device__ inline double ssqrt(double a) { return sqrt(a); }
template <int N> __host__ __device__ inline Dual<N> ssqrt(const Dual<N>& a) {
  Dual<N> r; r.v = sqrt(a.v); const double h = 0.5 / r.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * h;
  return r;
}
// 1/sqrt(a): one MUFU + Newton step for doubles instead of a sqrt followed by divisions
__host__ __device__ inline double srsqrt(double a) {
#ifdef __CUDA_ARCH__
  return rsqrt(a);
#else
  return 1.0 / sqrt(a);
#endif
}
template <int N> __host__ __device__ inline Dual<N> srsqrt(const Dual<N>& a) { return 1.0 / ssqrt(a); }
__host__ __device__ inline double sabs(double a) { return fabs(a); }
template <int N> __host__ __device__ inline Dual<N> sabs(const Dual<N>& a) { return a.v < 0.0 ? -a : a; }
__host__ __device__ inline double val(double a) { return a; }
template <int N> __host__ __device__ inline double val(const Dual<N>& a) { return a.v; }

template <class S> struct Lift { __host__ __device__ static S from(double v); };
template <> struct Lift<double> { __host__ __device__ static double from(double v) { return v; } };
template <int N> struct Lift<Dual<N>> { __host__ __device__ static Dual<N> from(double v) { return mk<N>(v); } };

// ------------------------------------------------------------------ MITC4 tables
// natural coordinates of nodes 1..4 (i,j,m,n): (+1,+1), (-1,+1), (-1,-1), (+1,-1)
// (the shape-function derivative rows of element.py:721-722); the Gauss points
// use the same sign pattern scaled by 1/sqrt(3), in the order of element.py:939-949.
#define JSSO_GP 0.57735026918962576451
__host__ __device__ inline double node_r(int k) { return (k == 0 || k == 3) ? 1.0 : -1.0; }
__host__ __device__ inline double node_s(int k) { return (k < 2) ? 1.0 : -1.0; }

// Local frame of a quad (element.py:502-521 / 649-671): x^ along the 3->1 diagonal,
// z^ = x_raw x v42, y^ = z_raw x x_raw; 2-D coordinates relative to node 3.
template <class S>
struct QuadFrame {
  S R[3][3];      // rows x^, y^, z^  (dirCos)
  S x[4], y[4];   // projected local coordinates; x[2] = y[2] = 0
};

template <class S>
__host__ __device__ inline void quad_frame(const S P[4][3], QuadFrame<S>& f) {
  S v31[3], v32[3], v34[3], v42[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    v31[c] = P[0][c] - P[2][c];
    v32[c] = P[1][c] - P[2][c];
    v34[c] = P[3][c] - P[2][c];
    v42[c] = P[1][c] - P[3][c];
  }
  S zr[3] = {v31[1] * v42[2] - v31[2] * v42[1], v31[2] * v42[0] - v31[0] * v42[2],
             v31[0] * v42[1] - v31[1] * v42[0]};
  S yr[3] = {zr[1] * v31[2] - zr[2] * v31[1], zr[2] * v31[0] - zr[0] * v31[2],
             zr[0] * v31[1] - zr[1] * v31[0]};
  const S inx = srsqrt(v31[0] * v31[0] + v31[1] * v31[1] + v31[2] * v31[2]);
  const S iny = srsqrt(yr[0] * yr[0] + yr[1] * yr[1] + yr[2] * yr[2]);
  const S inz = srsqrt(zr[0] * zr[0] + zr[1] * zr[1] + zr[2] * zr[2]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    f.R[0][c] = v31[c] * inx;
    f.R[1][c] = yr[c] * iny;
    f.R[2][c] = zr[c] * inz;
  }
  f.x[0] = v31[0] * f.R[0][0] + v31[1] * f.R[0][1] + v31[2] * f.R[0][2];
  f.y[0] = v31[0] * f.R[1][0] + v31[1] * f.R[1][1] + v31[2] * f.R[1][2];
  f.x[1] = v32[0] * f.R[0][0] + v32[1] * f.R[0][1] + v32[2] * f.R[0][2];
  f.y[1] = v32[0] * f.R[1][0] + v32[1] * f.R[1][1] + v32[2] * f.R[1][2];
  f.x[2] = Lift<S>::from(0.0);
  f.y[2] = Lift<S>::from(0.0);
  f.x[3] = v34[0] * f.R[0][0] + v34[1] * f.R[0][1] + v34[2] * f.R[0][2];
  f.y[3] = v34[0] * f.R[1][0] + v34[1] * f.R[1][1] + v34[2] * f.R[1][2];
}

// gp-independent MITC4 shear data (element.py:749-762, 783-798)
template <class S>
struct QuadShear {
  S gry[2], grx[2];  // g^r_k theta_x / theta_y entries: -dy/4, dx/4 of edge 1-2 (s_k=+1) / 4-3 (s_k=-1)
  S gsy[2], gsx[2];  // g^s_k entries of edge 1-4 (r_k=+1) / 2-3 (r_k=-1)
  S m11, m12, m22;   // M^T M, M = [[sin b, -sin a], [-cos b, cos a]]
};

template <class S>
__host__ __device__ inline void quad_shear(const QuadFrame<S>& f, QuadShear<S>& q) {
  const S* x = f.x; const S* y = f.y;
  q.gry[0] = (y[0] - y[1]) * -0.25; q.grx[0] = (x[0] - x[1]) * 0.25;
  q.gry[1] = (y[3] - y[2]) * -0.25; q.grx[1] = (x[3] - x[2]) * 0.25;
  q.gsy[0] = (y[0] - y[3]) * -0.25; q.gsx[0] = (x[0] - x[3]) * 0.25;
  q.gsy[1] = (y[1] - y[2]) * -0.25; q.gsx[1] = (x[1] - x[2]) * 0.25;
  // r-axis ~ (x1+x4-x2-x3, y1+y4-y2-y3)/2, s-axis ~ (x1+x2-x3-x4, ...)/2, normalised
  S rx = ((x[0] + x[3]) - (x[1] + x[2])) * 0.5, ry = ((y[0] + y[3]) - (y[1] + y[2])) * 0.5;
  S sx = ((x[0] + x[1]) - (x[2] + x[3])) * 0.5, sy = ((y[0] + y[1]) - (y[2] + y[3])) * 0.5;
  const S inr = srsqrt(rx * rx + ry * ry), ins = srsqrt(sx * sx + sy * sy);
  const S ca = rx * inr, cb = sx * ins;
  const S sa = -sabs(ry * inr);   // element.py:795: sin_alpha = -|r^ x e_x|
  const S sb = sabs(sy * ins);    // element.py:796
  q.m11 = sb * sb + cb * cb;
  q.m12 = -(sa * sb + ca * cb);
  q.m22 = sa * sa + ca * ca;
}

// per-Gauss-point data: inverse Jacobian, detJ and detJ*{gr^2, gr gs, gs^2}
template <class S>
struct QuadGp {
  S ji[4];   // J^-1 row-major
  S det;
  S prr, prs, pss;
};

template <class S>
__host__ __device__ inline void quad_gp(const QuadFrame<S>& f, int q, QuadGp<S>& g) {
  const double r = JSSO_GP * node_r(q), s = JSSO_GP * node_s(q);
  // J = dN * [x y]   (element.py:708-709), dN_r,k = r_k (1 + s s_k)/4, dN_s,k = s_k (1 + r r_k)/4
  S j00 = Lift<S>::from(0.0), j01 = j00, j10 = j00, j11 = j00;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double dr = 0.25 * node_r(k) * (1.0 + s * node_s(k));
    const double ds = 0.25 * node_s(k) * (1.0 + r * node_r(k));
    j00 = j00 + f.x[k] * dr; j01 = j01 + f.y[k] * dr;
    j10 = j10 + f.x[k] * ds; j11 = j11 + f.y[k] * ds;
  }
  g.det = j00 * j11 - j01 * j10;
  const S id = 1.0 / g.det;
  g.ji[0] = j11 * id; g.ji[1] = -(j01 * id);
  g.ji[2] = -(j10 * id); g.ji[3] = j00 * id;
  const S n1 = j10 * j10 + j11 * j11;   // (Cx + r Bx)^2 + (Cy + r By)^2 = 16 n1
  const S n0 = j00 * j00 + j01 * j01;   // (Ax + s Bx)^2 + (Ay + s By)^2 = 16 n0
  const S q4 = id * 0.25;
  g.prr = n1 * q4;                      // detJ gr^2, gr = sqrt(16 n1)/(8 detJ)
  g.pss = n0 * q4;
  g.prs = ssqrt(n1 * n0) * q4;
}

// material constants of one quad (element.py:820-875); cm* already include t
struct QuadMat {
  double cm11, cm12, cm21, cm22, cm33;  // t * Cm
  double D, nu, hb;                     // Cb = D [[1,nu,0],[nu,1,0],[0,0,hb]], hb = (1-nu)/2
  double ks;                            // Cs = ks I
};

__host__ __device__ inline void quad_mat(const double* prop, QuadMat& m) {
  const double t = prop[0], E = prop[1], nu = prop[2], kx = prop[3], ky = prop[4];
  const double ip = 1.0 / (1.0 + nu);          // two divisions in total
  const double pre = ip / (1.0 - nu);          // 1 / (1 - nu^2)
  const double tE = t * E;
  m.cm11 = pre * (tE * kx);
  m.cm12 = pre * (nu * (tE * kx));
  m.cm21 = pre * (nu * (tE * ky));
  m.cm22 = pre * (tE * ky);
  m.cm33 = 0.5 * tE * ip;                      // pre * (1-nu^2) * G, G = E / (2 (1+nu))
  m.D = tE * t * t * pre * (1.0 / 12.0);
  m.nu = nu;
  m.hb = 0.5 * (1.0 - nu);
  m.ks = tE * (5.0 / 12.0) * ip;
}

// Contribution of Gauss point q to the eight diagonal entries
// k[1,1],k[2,2],k[4,4],k[5,5],... of k1 + k2 (element.py:976-978); summed over q
// they feed k_rz = min|.|/1000.
template <class S>
__host__ __device__ inline void quad_diag_gp(const QuadShear<S>& sh, const QuadGp<S>& g, int q,
                                             const QuadMat& m, S diag[8]) {
  const double r = JSSO_GP * node_r(q), s = JSSO_GP * node_s(q);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double dr = 0.25 * node_r(k) * (1.0 + s * node_s(k));
    const double ds = 0.25 * node_s(k) * (1.0 + r * node_r(k));
    const S h0 = g.ji[0] * dr + g.ji[1] * ds, h1 = g.ji[2] * dr + g.ji[3] * ds;
    const S pxx = g.det * h0 * h0, pyy = g.det * h1 * h1;
    const double fr = 1.0 + s * node_s(k), fs = 1.0 + r * node_r(k);
    const S crr = g.prr * (fr * fr), crs = g.prs * (fr * fs), css = g.pss * (fs * fs);
    const int ir = (k < 2) ? 0 : 1, is = (k == 0 || k == 3) ? 0 : 1;
    const S ax = sh.gry[ir], bx = sh.gsy[is];   // theta_x entries
    const S ay = sh.grx[ir], by = sh.gsx[is];   // theta_y entries
    diag[2 * k] = (pyy + pxx * m.hb) * m.D +
                  (sh.m11 * crr * ax * ax + sh.m12 * crs * ax * bx * 2.0 + sh.m22 * css * bx * bx) * m.ks;
    diag[2 * k + 1] = (pxx + pyy * m.hb) * m.D +
                      (sh.m11 * crr * ay * ay + sh.m12 * crs * ay * by * 2.0 + sh.m22 * css * by * by) * m.ks;
  }
}

// ------------------------------------------------------------------ beam-column
// Direction cosines of element.py:76-97 (sin_alpha = 0, cos_alpha = 1).  Returns
// true when the member is exactly parallel to global Y (Cxz == 0), where the
// reference switches to a non-orthonormal matrix (element.py:92-94).
template <class S>
__host__ __device__ inline bool beam_dircos(const S P[2][3], S R[3][3], S& L) {
  const S dx = P[1][0] - P[0][0], dy = P[1][1] - P[0][1], dz = P[1][2] - P[0][2];
  L = ssqrt(dx * dx + dy * dy + dz * dz);
  const S Cx = dx / L, Cy = dy / L, Cz = dz / L;
  const S Cxz = ssqrt(Cx * Cx + Cz * Cz);
  const S zero = Lift<S>::from(0.0);
  if (val(Cxz) == 0.0) {
    R[0][0] = zero; R[0][1] = Cy;   R[0][2] = zero;
    R[1][0] = -Cy;  R[1][1] = zero; R[1][2] = Lift<S>::from(-1.0);
    R[2][0] = -Cy;  R[2][1] = zero; R[2][2] = zero;
    return true;
  }
  R[0][0] = Cx;               R[0][1] = Cy;   R[0][2] = Cz;
  R[1][0] = Cz / Cxz;         R[1][1] = zero; R[1][2] = -(Cx / Cxz);
  R[2][0] = -(Cx * Cy / Cxz); R[2][1] = Cxz;  R[2][2] = -(Cy * Cz / Cxz);
  return false;
}

}  // namespace jsso
