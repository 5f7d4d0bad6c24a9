/* TEST INFRASTRUCTURE ONLY -- the CPU oracle of the Cooper-Frye hot path.
 *
 * A plain, loop-for-loop restatement of the reference algorithm (file:line cited at each function in
 * cf_oracle.cpp), written from the reference's formulas but not copied from it, scalar FP64, one thread.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never does.
 *
 * Pinning: checked against the UNMODIFIED reference compiled into oracle/_ref (tests/test_oracle_cpu.py runs both
 * on the same seeded surfaces here, and against the committed tests/golden vectors everywhere).
 * The GSL pieces (natural cubic spline, 3x3 LU) are un-vendored third-party code (GSL 2.x, FindGSL.cmake): the
 * oracle restates their published algorithms; parity at that boundary is pinned only through oracle/_ref, which
 * itself links oracle/gsl_shim instead of the real GSL.
 */
#ifndef CF_ORACLE_H
#define CF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int operation, dimension, df_mode;
  int include_baryon, include_bulk_deltaf, include_shear_deltaf, include_baryondiff_deltaf;
  int regulate_deltaf, outflow;
  double deta_min, mass_pion0;
  int fast;
  double y_cut;
  /* histogram grids (dN/dX) */
  double tau_min, tau_max; int tau_bins;
  double r_min, r_max; int r_bins;
  int phip_bins;
  int famod_chain;   /* 1 = previous cell's solution as initial guess (reference), 0 = always (T,1,1) */
} cf_params;

typedef struct {
  /* surface, structure of arrays, physical units; 25 columns in the order of include/is3d_b200.h */
  long n_cells;
  const double *col[25];
  /* chosen species */
  int n_species;
  const double *mass, *sign, *degeneracy, *baryon;
  const double *equilibrium_density, *bulk_density, *diffusion_density;
  /* whole PDG table (PTMA) */
  int n_pdg;
  const double *pdg_mass, *pdg_sign, *pdg_degeneracy, *pdg_baryon;
  /* momentum tables as in the files (node, weight) */
  int n_pT, n_phi, n_y, n_eta;
  const double *pT, *pT_w, *phi, *phi_w, *y, *y_w, *eta, *eta_w;
  /* Gauss-Laguerre tables [n_alpha][n_gla] */
  int n_alpha, n_gla;
  const double *gla_root, *gla_weight;
  /* df coefficient tables, [n_muB][n_T] row-major */
  int n_T, n_muB;
  const double *T_arr, *muB_arr;
  const double *c0, *c1, *c2, *c3, *c4, *F, *G, *betabulk, *betaV, *betapi;
  /* PTB tables */
  int n_ptb;
  const double *ptb_x, *ptb_lambda2, *ptb_z;
  double ptb_x_max;
  /* surface averages */
  double T_avg, E_avg, P_avg, muB_avg, nB_avg;
} cf_inputs;

typedef struct {
  long cells_skipped, cells_breakdown, cells_pl_negative, reconstruction_failures, newton_iterations;
  long cells_out_of_table;
} cf_stats;

/* dN/pTdpTdphidy, all df modes; out has n_species * n_pT * n_phi * Ny doubles (Ny = n_y in 3+1d, 1 in 2+1d),
 * index iy + Ny*(iphi + n_phi*(ipT + n_pT*is)).  Returns 0, or non-zero when a cell leaves the df tables. */
int cf_oracle_spectra(const cf_params *p, const cf_inputs *in, double *out, cf_stats *st);

/* dN/dX histograms (df modes 1-4), each n_species x bins, unnormalised sums */
int cf_oracle_dndx(const cf_params *p, const cf_inputs *in, double *tau_hist, double *r_hist, double *phi_hist, cf_stats *st);

/* calculate_total_yield and the sampler's per-cell mean yields */
int cf_oracle_total_yield(const cf_params *p, const cf_inputs *in, double *ntotal);
int cf_oracle_cell_yields(const cf_params *p, const cf_inputs *in, double *dn_tot, double *dn_list);

/* calculate_spin_polzn (Polarization.cpp:25-263): w = thermal vorticity columns wtx wty wtn wxy wxn wyn; St..Snorm have
 * n_species * n_pT * n_phi * Ny doubles each in the spectra layout.  chunk_compat = 1 reads the vorticity with the index
 * inside the reference's 10 000-cell chunk (:125-130). */
int cf_oracle_polarization(const cf_params *p, const cf_inputs *in, const double *const w[6], int chunk_compat, double *St,
                           double *Sx, double *Sy, double *Sn, double *Snorm);

#ifdef __cplusplus
}
#endif
#endif
