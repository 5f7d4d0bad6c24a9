// Gauss-Laguerre thermal integrals (host + device).  Restates reference src/cpp/GaussThermal.cpp: the integrands
// carry an explicit exp(pbar) because the tabulated weights already contain pbar^alpha exp(-pbar).
#pragma once

#include "common.cuh"

namespace is3d {

enum ThermalIntegrand { TI_NEQ = 0, TI_J10, TI_J11, TI_J20, TI_J30, TI_J31 };

// one quadrature node; GaussThermal.cpp:19-85
template <int KIND>
IS3D_HD double thermal_integrand(double pbar, double mbar, double alphaB, double baryon, double sign)
{
  double Ebar = sqrt(pbar * pbar + mbar * mbar);
  if (KIND == TI_NEQ) return pbar * exp(pbar) / (exp(Ebar - baryon * alphaB) + sign);
  double qstat = exp(Ebar - baryon * alphaB) + sign;
  double boltz = exp(pbar + Ebar - baryon * alphaB) / (qstat * qstat);
  if (KIND == TI_J10) return pbar * boltz;
  if (KIND == TI_J11) return pbar * pbar * pbar / (Ebar * Ebar) * boltz;
  if (KIND == TI_J20) return Ebar * boltz;
  if (KIND == TI_J30) return Ebar * Ebar / pbar * boltz;
  return pbar * boltz;   // TI_J31
}

// GaussThermal, GaussThermal.cpp:7-15
template <int KIND>
IS3D_HD double gauss_thermal(const double *root, const double *weight, int pts, double mbar, double alphaB, double baryon,
                             double sign)
{
  double s = 0.0;
  for (int k = 0; k < pts; k++) s += weight[k] * thermal_integrand<KIND>(root[k], mbar, alphaB, baryon, sign);
  return s;
}

// PTB integrands E_mod_int / P_mod_int, GaussThermal.cpp:93-116
IS3D_HD double E_mod_int(double pbar, double mbar, double lambda, double sign)
{
  double scale2 = (1.0 + lambda) * (1.0 + lambda);
  double Ebar = sqrt(pbar * pbar + mbar * mbar);
  return sqrt(pbar * pbar * scale2 + mbar * mbar) * exp(pbar) / (exp(Ebar) + sign);
}
IS3D_HD double P_mod_int(double pbar, double mbar, double lambda, double sign)
{
  double scale2 = (1.0 + lambda) * (1.0 + lambda);
  double Ebar = sqrt(pbar * pbar + mbar * mbar);
  return pbar * pbar * scale2 / sqrt(pbar * pbar * scale2 + mbar * mbar) * exp(pbar) / (exp(Ebar) + sign);
}

}  // namespace is3d
