/* TEST INFRASTRUCTURE ONLY.  Force-included (after <iomanip>) when the unmodified reference sources are
 * compiled into oracle/_ref/: raises the precision of every text writer from 6/8 to 17 significant digits so
 * that the reference's results/ files (the only place the dN/dX histograms surface, reference
 * SpacetimeDistribution.cpp:448-490) can pin parity below 1e-10.  No arithmetic is changed. */
#ifndef IS3D_REF_PRECISION_H
#define IS3D_REF_PRECISION_H
#define setprecision(n) setprecision(17)
#endif
