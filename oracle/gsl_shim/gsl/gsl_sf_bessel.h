/* included by the reference but no Bessel function is ever called */
#ifndef IS3D_GSL_SHIM_BESSEL_H
#define IS3D_GSL_SHIM_BESSEL_H
#endif
