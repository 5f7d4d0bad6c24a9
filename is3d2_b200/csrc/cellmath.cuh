// Per-cell algebra shared by all kernels: flow normalisation, completion of pi^{mu nu}, Milne basis, boosts to
// the local rest frame.  Restates (does not copy) reference src/cpp/LocalRestFrame.cpp and the per-cell
// prologues of MomentumSpectra.cpp:109-193.
#pragma once

#include "common.cuh"

namespace is3d {

// one freezeout cell as unpacked by the reference (EmissionFunction.cpp:1113-1161)
struct Cell {
  double tau, x, y, eta;
  double dat, dax, day, dan;
  double ux, uy, un;
  double E, T, P;
  double pixx, pixy, pixn, piyy, piyn;
  double bulkPi;
  double muB, nB, Vx, Vy, Vn;
};

struct SurfaceView {
  const double *col[25];
  int64_t n;
};

IS3D_HD Cell load_cell(const SurfaceView &s, int64_t i, bool has_baryon)
{
  Cell c;
  c.tau = s.col[0][i];  c.x = s.col[1][i];  c.y = s.col[2][i];  c.eta = s.col[3][i];
  c.dat = s.col[4][i];  c.dax = s.col[5][i]; c.day = s.col[6][i]; c.dan = s.col[7][i];
  c.ux = s.col[8][i];   c.uy = s.col[9][i];  c.un = s.col[10][i];
  c.E = s.col[11][i];   c.T = s.col[12][i];  c.P = s.col[13][i];
  c.pixx = s.col[14][i]; c.pixy = s.col[15][i]; c.pixn = s.col[16][i]; c.piyy = s.col[17][i]; c.piyn = s.col[18][i];
  c.bulkPi = s.col[19][i];
  if (has_baryon) {
    c.muB = s.col[20][i]; c.nB = s.col[21][i]; c.Vx = s.col[22][i]; c.Vy = s.col[23][i]; c.Vn = s.col[24][i];
  } else {
    c.muB = c.nB = c.Vx = c.Vy = c.Vn = 0.0;
  }
  return c;
}

// full contravariant shear stress in Milne coordinates
struct Shear {
  double tt = 0, tx = 0, ty = 0, tn = 0, xx = 0, xy = 0, xn = 0, yy = 0, yn = 0, nn = 0;
};

// Reconstruct pi^{tau mu}, pi^{eta eta} from the five stored components by orthogonality pi.u = 0 and
// tracelessness (MomentumSpectra.cpp:156-160).
IS3D_HD Shear complete_shear(double pixx, double pixy, double pixn, double piyy, double piyn, double ut, double ux,
                             double uy, double un, double tau2)
{
  Shear p;
  double ux2 = ux * ux, uy2 = uy * uy, ut2 = ut * ut;
  double utperp2 = 1.0 + ux2 + uy2;
  double tau2_un = tau2 * un;
  p.xx = pixx; p.xy = pixy; p.xn = pixn; p.yy = piyy; p.yn = piyn;
  p.nn = (pixx * (ux2 - ut2) + piyy * (uy2 - ut2) + 2.0 * (pixy * ux * uy + tau2_un * (pixn * ux + piyn * uy))) / (tau2 * utperp2);
  p.tn = (pixn * ux + piyn * uy + tau2_un * p.nn) / ut;
  p.ty = (pixy * ux + piyy * uy + tau2_un * piyn) / ut;
  p.tx = (pixx * ux + pixy * uy + tau2_un * pixn) / ut;
  p.tt = (p.tx * ux + p.ty * uy + tau2_un * p.tn) / ut;
  return p;
}

// Milne_Basis, LocalRestFrame.cpp:12-41
struct Basis {
  double Xt, Xx, Xy, Xn, Yx, Yy, Zt, Zn;
};

IS3D_HD Basis milne_basis(double ut, double ux, double uy, double un, double tau)
{
  Basis b;
  double uperp = sqrt(ux * ux + uy * uy);
  double utperp = sqrt(1.0 + ux * ux + uy * uy);
  double sinhL = tau * un / utperp;
  double coshL = ut / utperp;
  b.Xt = uperp * coshL;
  b.Xx = 1.0; b.Xy = 0.0;
  b.Xn = uperp * sinhL / tau;
  b.Yx = 0.0; b.Yy = 1.0;
  b.Zt = sinhL;
  b.Zn = coshL / tau;
  if (uperp > 1.e-5) {
    b.Xx = utperp * ux / uperp;
    b.Xy = utperp * uy / uperp;
    b.Yx = -uy / uperp;
    b.Yy = ux / uperp;
  }
  return b;
}

// pi_ij in the local rest frame, Shear_Stress::boost_pimunu_to_lrf (LocalRestFrame.cpp:133-154)
struct ShearLRF {
  double xx, xy, xz, yy, yz, zz;
};

IS3D_HD ShearLRF boost_shear_to_lrf(const Shear &p, const Basis &b, double tau2)
{
  ShearLRF l;
  double Xt = b.Xt, Xx = b.Xx, Xy = b.Xy, Xn = b.Xn, Yx = b.Yx, Yy = b.Yy, Zt = b.Zt, Zn = b.Zn;
  l.xx = p.tt * Xt * Xt + p.xx * Xx * Xx + p.yy * Xy * Xy + tau2 * tau2 * p.nn * Xn * Xn
       + 2.0 * (-Xt * (p.tx * Xx + p.ty * Xy) + p.xy * Xx * Xy + tau2 * Xn * (p.xn * Xx + p.yn * Xy - p.tn * Xt));
  l.xy = Yx * (-p.tx * Xt + p.xx * Xx + p.xy * Xy + tau2 * p.xn * Xn) + Yy * (-p.ty * Xt + p.xy * Xx + p.yy * Xy + tau2 * p.yn * Xn);
  l.xz = Zt * (p.tt * Xt - p.tx * Xx - p.ty * Xy - tau2 * p.tn * Xn) - tau2 * Zn * (p.tn * Xt - p.xn * Xx - p.yn * Xy - tau2 * p.nn * Xn);
  l.yy = p.xx * Yx * Yx + 2.0 * p.xy * Yx * Yy + p.yy * Yy * Yy;
  l.yz = -Zt * (p.tx * Yx + p.ty * Yy) + tau2 * Zn * (p.xn * Yx + p.yn * Yy);
  l.zz = -(l.xx + l.yy);
  return l;
}

// V_i = -X_i.V, Baryon_Diffusion::boost_Vmu_to_lrf (LocalRestFrame.cpp:173-185)
IS3D_HD void boost_V_to_lrf(double Vt, double Vx, double Vy, double Vn, const Basis &b, double tau2, double *Vx_LRF,
                            double *Vy_LRF, double *Vz_LRF)
{
  *Vx_LRF = -Vt * b.Xt + Vx * b.Xx + Vy * b.Xy + tau2 * Vn * b.Xn;
  *Vy_LRF = Vx * b.Yx + Vy * b.Yy;
  *Vz_LRF = -Vt * b.Zt + tau2 * Vn * b.Zn;
}

// Surface_Element_Vector::boost_dsigma_to_lrf / compute_dsigma_magnitude (LocalRestFrame.cpp:81-98)
struct DsigmaLRF {
  double t, x, y, z, space, magnitude;
};

IS3D_HD DsigmaLRF boost_dsigma_to_lrf(double dat, double dax, double day, double dan, const Basis &b, double ut,
                                      double ux, double uy, double un)
{
  DsigmaLRF d;
  d.t = dat * ut + dax * ux + day * uy + dan * un;
  d.x = -(dat * b.Xt + dax * b.Xx + day * b.Xy + dan * b.Xn);
  d.y = -(dax * b.Yx + day * b.Yy);
  d.z = -(dat * b.Zt + dan * b.Zn);
  d.space = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
  d.magnitude = fabs(d.t) + d.space;
  return d;
}

// 3x3 inverse by cofactors followed by one Newton-Schulz refinement X <- X(2I - AX): the reference inverts with
// GSL LU (MomentumSpectra.cpp:729-747) and then iteratively refines every momentum solve to 1e-16
// (:959-971); refining the inverse once per cell gives the same p' = A^-1 p to rounding.
IS3D_HD void invert3x3(const double A[9], double Ainv[9], double *det_out)
{
  double c00 = A[4] * A[8] - A[5] * A[7];
  double c01 = A[5] * A[6] - A[3] * A[8];
  double c02 = A[3] * A[7] - A[4] * A[6];
  double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  double r = 1.0 / det;
  double X[9];
  X[0] = c00 * r; X[1] = (A[2] * A[7] - A[1] * A[8]) * r; X[2] = (A[1] * A[5] - A[2] * A[4]) * r;
  X[3] = c01 * r; X[4] = (A[0] * A[8] - A[2] * A[6]) * r; X[5] = (A[2] * A[3] - A[0] * A[5]) * r;
  X[6] = c02 * r; X[7] = (A[1] * A[6] - A[0] * A[7]) * r; X[8] = (A[0] * A[4] - A[1] * A[3]) * r;
  // R = I - A X
  double R[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < 3; k++) s = fma(-A[3 * i + k], X[3 * k + j], s);
      R[3 * i + j] = s;
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = X[3 * i + j];
      for (int k = 0; k < 3; k++) s = fma(X[3 * i + k], R[3 * k + j], s);
      Ainv[3 * i + j] = s;
    }
  *det_out = det;
}

}  // namespace is3d
