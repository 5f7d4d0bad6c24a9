/*
 * is3d_b200.h -- C ABI of the B200-native Cooper-Frye particlization hot path.
 *
 * Drop-in boundary: the reference has no FFI; the seam is the C++ member-function boundary
 * between EmissionFunctionArray::calculate_spectra and its eight compute members, taken AFTER the AoS->SoA
 * unpack of the freezeout surface (reference src/cpp/EmissionFunction.cpp:1050-1161, dispatch :1164-1277).
 * Every entry point below names the reference interface it replaces.  Plain C: pointers and sizes only, no
 * C++/torch types, no exceptions.  All floating point is FP64; all arrays are caller-owned unless stated.
 *
 * Conventions
 *   - every function returns an is3d_status (0 = ok); is3d_last_error() gives the message.  The reference's
 *     convention is printf + exit(-1) (e.g. EmissionFunction.cpp:156-157, DeltafData.cpp:430-434) or a GSL
 *     abort; the host layer (is3d_host.h) turns a non-zero status back into that behaviour.
 *   - one context per host thread and per GPU; calls are synchronous unless suffixed _async.
 *   - there is NO CPU fallback: if no sm_100-class device is usable, is3d_create fails.
 */
#ifndef IS3D_B200_H
#define IS3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct is3d_ctx is3d_ctx;

typedef enum {
  IS3D_OK = 0,
  IS3D_ERR_INVALID = 1,      /* bad argument / call order */
  IS3D_ERR_CUDA = 2,         /* CUDA runtime failure (message holds cudaGetErrorString) */
  IS3D_ERR_TABLE_RANGE = 3,  /* a cell's (T, muB, Pi/P) is outside the df coefficient tables: the reference aborts
                                (GSL domain error, DeltafData.cpp:338-377) or exit(-1)s (DeltafData.cpp:430-434) */
  IS3D_ERR_UNSUPPORTED = 4,  /* combination the reference itself rejects (e.g. dN/dX with df_mode 5,
                                EmissionFunction.cpp:1184-1189; PTB with include_baryon, DeltafData.cpp:480-484) */
  IS3D_ERR_NO_DEVICE = 5
} is3d_status;

/* Number of structure-of-arrays surface columns and their order (is3d_set_surface).  This is the column set the
 * reference unpacks in EmissionFunction.cpp:1050-1092 (tau x y eta | dat dax day dan | ux uy un | E T P |
 * pixx pixy pixn piyy piyn | bulkPi | muB nB Vx Vy Vn), in physical units (GeV, fm) as left by the readers. */
#define IS3D_SURFACE_COLUMNS 25
enum {
  IS3D_COL_TAU = 0, IS3D_COL_X, IS3D_COL_Y, IS3D_COL_ETA,
  IS3D_COL_DAT, IS3D_COL_DAX, IS3D_COL_DAY, IS3D_COL_DAN,
  IS3D_COL_UX, IS3D_COL_UY, IS3D_COL_UN,
  IS3D_COL_E, IS3D_COL_T, IS3D_COL_P,
  IS3D_COL_PIXX, IS3D_COL_PIXY, IS3D_COL_PIXN, IS3D_COL_PIYY, IS3D_COL_PIYN,
  IS3D_COL_BULKPI,
  IS3D_COL_MUB, IS3D_COL_NB, IS3D_COL_VX, IS3D_COL_VY, IS3D_COL_VN
};

/* Run-time switches: the members of EmissionFunctionArray that the compute paths read
 * (EmissionFunction.h:34-93, filled from iS3D_parameters.dat in EmissionFunction.cpp:139-247). */
typedef struct {
  int operation;                 /* 0 dN/dX, 1 continuous spectra, 2 sampler */
  int dimension;                 /* 2 = boost-invariant 2+1d, 3 = 3+1d */
  int df_mode;                   /* 1 Grad 14-moment, 2 RTA Chapman-Enskog, 3 PTM, 4 PTB, 5 PTMA */
  int include_baryon;
  int include_bulk_deltaf;
  int include_shear_deltaf;
  int include_baryondiff_deltaf;
  int regulate_deltaf;
  int outflow;
  double deta_min;
  double mass_pion0;
  int fast;                      /* sampler: species densities at (T_avg, muB_avg) */
  double y_cut;
  int64_t sampler_seed;          /* must be >= 0 here; the host layer resolves "< 0 = clock" */
  int test_sampler;
  /* histogram grids for the sampler self-test and dN/dX (EmissionFunction.h:69-93) */
  double pT_min, pT_max;  int pT_bins;
  int y_bins;
  int phip_bins;
  double eta_cut;         int eta_bins;
  double tau_min, tau_max; int tau_bins;
  double r_min, r_max;     int r_bins;
  /* device selection and library-only knobs (no reference counterpart) */
  int device;                    /* CUDA ordinal */
  int famod_chain;               /* df_mode 5 initial guess policy: 0 (default) = every cell starts from (T,1,1), what the
                                    reference does for a cell without a previous solution (MomentumSpectra.cpp:1288-1313):
                                    cells independent, shardable; 1 = the reference's serial chain (previous cell's solution,
                                    :1308-1364) walked by ONE warp -- bit-level parity runs on one GPU only */
  int dndx_bug_compat;           /* 1 = is3d_dndx (host output) reproduces the reference's partial memset
                                    (SpacetimeDistribution.cpp:166-168): histograms accumulate over species above bin
                                    CORES*bins/8; is3d_dndx_device always returns the clean per-species histograms */
  int polzn_chunk_compat;        /* 0 (default) = every cell reads its own thermal vorticity; 1 = the reference's index INSIDE
                                    its 10 000-cell chunk (Polarization.cpp:125-130 use wtx_fo[icell], not [icell_glb]):
                                    identical for surfaces of up to 10 000 cells, unsharded surfaces only */
  double negligible_margin;      /* continuous spectra (all df modes) and dN/dX: (cell, y, phi) items whose every exponent (u.p - b mu_B)/T
                                    in a block of momentum columns exceeds the block row's smallest possible exponent by more
                                    than this margin are dropped before the momentum loop -- their terms are below e^-margin of
                                    the bins' leading terms.  A speed heuristic, not a precision knob: the library sums a
                                    rigorous bound of everything it dropped and compares it with each finished bin (bound <=
                                    1e-13 |bin|); if any bin fails, the call is repeated without the margin
                                    (is3d_stats.prune_reruns).  <= 0: off.  Default 60: with up to 1e7 dropped
                                    terms per pass against ONE leading term the bound test still passes with a margin of ~50 */
} is3d_params;

/* Counters the reference prints (MomentumSpectra.cpp:1039-1040, :1674-1679; ParticleSampler.cpp:1133). */
typedef struct {
  int64_t cells_total;
  int64_t cells_skipped;         /* u.dsigma <= 0 */
  int64_t cells_breakdown;       /* feqmod / famod breakdown */
  int64_t cells_pl_negative;
  int64_t reconstruction_failures;
  int64_t newton_iterations;
  int64_t cells_out_of_table;
  int64_t sampler_proposals;
  int64_t sampler_accepted;
  double  tau_breakdown;
  double  tau_pl_negative;
  double  kernel_ms;             /* device time of the dominant kernel(s) of the last call (CUDA events) */
  int64_t kernel_launches;       /* kernels launched by the last call */
  int64_t evals_executed;        /* class-evaluations the dominant spectra kernel executed in the last call, padding slots and
                                    idle thread columns included (df_mode 1, 2; 0 = not reported): x FP64 instructions per
                                    evaluation (SASS) = the executed FP64-pipe work behind kernel_ms */
  int64_t pair_evals_executed;   /* the part of evals_executed done in charge-conjugate pair slots (a baryon class and its
                                    antibaryon class share x_E and its exponential) */
  int64_t evals_dropped;         /* class-evaluations NOT executed because their (cell, y, phi) item was dropped as negligible for the
                                    whole block of momentum columns (range guard at x >= 680, and is3d_params.negligible_margin) */
  int64_t prune_reruns;          /* 1 = the dropped-term bound test failed and the spectra were recomputed without negligible_margin */
} is3d_stats;

/* One sampled hadron: the reference's Sampled_Particle (SampledParticle.h:32-54), same fields. */
typedef struct {
  int32_t chosen_index;
  int32_t mcid;
  int32_t event;
  int32_t pad_;
  double mass;
  double tau, x, y, eta;
  double t, z;
  double E, px, py, pz;
} is3d_particle;

/* The same hadron as a 64-byte wire record (is3d_sample_compact): everything that is not a function of the other fields.
 * mass and mcid follow from chosen_index (the species arrays of is3d_set_species), E from the mass shell, (t, z) from
 * (tau, eta); is3d_expand_particles restores the full record.  The list is PCIe-bound: 64 instead of 104 bytes per hadron. */
typedef struct {
  int32_t chosen_index;
  int32_t event;
  double tau, x, y, eta;
  double px, py, pz;
} is3d_particle_compact;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
void        is3d_default_params(is3d_params *p);
is3d_status is3d_create(const is3d_params *p, is3d_ctx **out);
void        is3d_destroy(is3d_ctx *ctx);
const char *is3d_last_error(const is3d_ctx *ctx);          /* ctx may be NULL: last create error */
const char *is3d_version(void);
int         is3d_device_count(void);                       /* usable CUDA devices (0 when there is no driver / GPU) */

/* ---- static inputs (reference: arrays built in EmissionFunction.cpp:998-1046) --------------------------------- */
/* chosen species: Mass/Sign/Degeneracy/Baryon/MCID and the fast-mode densities (EmissionFunction.cpp:998-1021) */
is3d_status is3d_set_species(is3d_ctx *ctx, int n, const double *mass, const double *sign, const double *degeneracy,
                             const double *baryon, const int *mcid, const double *equilibrium_density,
                             const double *bulk_density, const double *diffusion_density);
/* whole PDG table for the PTMA reconstruction (EmissionFunction.cpp:1025-1036; first min(320,n) used) */
is3d_status is3d_set_pdg(is3d_ctx *ctx, int n, const double *mass, const double *sign, const double *degeneracy,
                         const double *baryon);
/* momentum / rapidity tables: column 1 = node, column 2 = weight of tables/momentum/{pT,phi,y}_table.dat and
 * tables/spacetime_rapidity/eta_table.dat (iS3D.cpp:254-257).  Lengths are the FILE lengths; the library applies
 * the reference's dimension rule itself (2+1d: y = {0}; 3+1d: eta = cell eta, EmissionFunction.cpp:146-153). */
is3d_status is3d_set_momentum_tables(is3d_ctx *ctx, int npT, const double *pT, const double *pT_weight,
                                     int nphi, const double *phi, const double *phi_weight,
                                     int ny, const double *y, const double *y_weight,
                                     int neta, const double *eta, const double *eta_weight);
/* Gauss-Laguerre (n_alpha x n_points, row-major; tables/gauss/gla_roots_weights.txt) and Gauss-Legendre tables
 * (readindata.cpp:26-92) */
is3d_status is3d_set_gauss_tables(is3d_ctx *ctx, int n_alpha, int n_points, const double *gla_root,
                                  const double *gla_weight, int n_legendre, const double *legendre_root,
                                  const double *legendre_weight);
/* surface-averaged thermodynamics (Plasma, readindata.cpp:104-119) */
is3d_status is3d_set_thermo_averages(is3d_ctx *ctx, double T, double E, double P, double muB, double nB);
/* df coefficient tables, each n_muB x n_T row-major (DeltafData.cpp:65-217).  For include_baryon = 0 pass
 * n_muB = 1 (the reader's rule, DeltafData.cpp:134); the library builds the natural cubic splines in T
 * (DeltafData.cpp:298-321) itself. */
is3d_status is3d_set_df_tables(is3d_ctx *ctx, int n_T, int n_muB, const double *T, const double *muB,
                               const double *c0, const double *c1, const double *c2, const double *c3,
                               const double *c4, const double *F, const double *G, const double *betabulk,
                               const double *betaV, const double *betapi);
/* PTB tables lambda^2(Pi/P), z(Pi/P) (DeltafData.cpp:220-295); n = 301 in the reference */
is3d_status is3d_set_ptb_tables(is3d_ctx *ctx, int n, const double *bulkPi_over_P, const double *lambda_squared,
                                const double *z, double bulkPi_over_P_max);

/* ---- surface -------------------------------------------------------------------------------------------- */
/* cols[k] = host pointer to n doubles (k as IS3D_COL_*); copied to HBM as structure-of-arrays.  Baryon columns
 * may be NULL when include_baryon = 0.  global_offset = index of cols[.][0] in the whole surface when the caller
 * shards cells across GPUs (keys the sampler's counter-based RNG so results do not depend on the sharding). */
is3d_status is3d_set_surface(is3d_ctx *ctx, int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS],
                             int64_t global_offset);
/* same, but cols[k] are DEVICE pointers that stay owned by the caller and must outlive the compute calls */
is3d_status is3d_set_surface_device(is3d_ctx *ctx, int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS],
                                    int64_t global_offset);

/* thermal vorticity of a mode-5 surface (readindata.cpp:299-307): w[k] = host pointer to n doubles, k = wtx wty wtn wxy
 * wxn wyn (FO_surf order, readindata.h:90); n must equal the surface set last.  Copied to HBM. */
is3d_status is3d_set_vorticity(is3d_ctx *ctx, int64_t n, const double *const w[6]);

/* ---- compute (reference: the eight compute members, EmissionFunction.h:147-179) ----------------------------- */
/* number of doubles is3d_spectra writes: Ns * NpT * Nphi * Ny, index iy + Ny*(iphi + Nphi*(ipT + NpT*is))
 * (MomentumSpectra.cpp:252-295) */
int64_t     is3d_spectra_size(const is3d_ctx *ctx);
/* calculate_dN_pTdpTdphidy / _feqmod / _famod (EmissionFunction.h:147-152): out is a HOST buffer */
is3d_status is3d_spectra(is3d_ctx *ctx, double *out, is3d_stats *stats);
/* same, result left in a caller-owned DEVICE buffer (for a cross-GPU reduction by the caller) */
is3d_status is3d_spectra_device(is3d_ctx *ctx, double *out_device, is3d_stats *stats);

/* calculate_dN_dX / calculate_dN_dX_feqmod (EmissionFunction.h:155-161): three histograms per species, each
 * Ns x bins row-major, UNnormalised sums of dN_dy_cell exactly as accumulated in SpacetimeDistribution.cpp:413-440
 * (the writers' normalisation lives in the host layer). */
is3d_status is3d_dndx(is3d_ctx *ctx, double *tau_hist, double *r_hist, double *phi_hist, is3d_stats *stats);
is3d_status is3d_dndx_device(is3d_ctx *ctx, double *tau_hist_dev, double *r_hist_dev, double *phi_hist_dev,
                             is3d_stats *stats);

/* calculate_spin_polzn (EmissionFunction.h, src/cpp/Polarization.cpp:25-263; run by calculate_spectra for mode-5 surfaces,
 * EmissionFunction.cpp:1304-1310): the four components of the polarization vector and its norm, each Ns NpT Nphi Ny doubles in
 * the SPECTRA layout iy + Ny*(iphi + Nphi*(ipT + NpT*is)); host buffers.  (The reference stores them species-fastest and
 * writes them with the spectra index; that file-level mix-up lives in the host layer's writer, not here.) */
is3d_status is3d_polarization(is3d_ctx *ctx, double *St, double *Sx, double *Sy, double *Sn, double *Snorm, is3d_stats *stats);

/* calculate_total_yield (EmissionFunction.h:172) */
is3d_status is3d_total_yield(is3d_ctx *ctx, double *ntotal, is3d_stats *stats);
/* per-cell mean yields of the sampler: dn_tot[cell] (after the 2*y_max*ds_max factor, ParticleSampler.cpp:915)
 * and, if dn_list != NULL, dn_list[cell*Ns + s] (ParticleSampler.cpp:876-911).  Host buffers. */
is3d_status is3d_cell_yields(is3d_ctx *ctx, double *dn_tot, double *dn_list, is3d_stats *stats);

/* sample_dN_pTdpTdphidy / _famod (EmissionFunction.h:175-179), df_mode 1-5, fast = 0 or 1.  Particles of all events
 * are returned in one library-owned, page-locked host array (event index in each record, grouped by event, within an
 * event ordered by (cell, draw) -- independent of the launch geometry); counts[e] = hadrons in event e (caller-owned,
 * nevents entries).  The array stays valid until is3d_free_particles, also across is3d_destroy; released buffers are
 * reused by later calls on the same context. */
is3d_status is3d_sample(is3d_ctx *ctx, int64_t nevents, is3d_particle **particles, int64_t *total,
                        int64_t *counts, is3d_stats *stats);
void        is3d_free_particles(void *particles);   /* any host list handed out by is3d_sample / _compact / is3d_group_sample */
/* same sampling (same hadrons, same order), delivered as 64-byte wire records in a library-owned pinned host array */
is3d_status is3d_sample_compact(is3d_ctx *ctx, int64_t nevents, is3d_particle_compact **particles, int64_t *total,
                                int64_t *counts, is3d_stats *stats);
/* same sampling, the full records left in a library-owned DEVICE array for a device-resident consumer (valid until the next
 * sampler call on this context; never copied to the host).  The context's cells must fit one sampler pass (16 M cells). */
is3d_status is3d_sample_device(is3d_ctx *ctx, int64_t nevents, const is3d_particle **particles_dev, int64_t *total,
                               int64_t *counts, is3d_stats *stats);
/* host-side expansion of n wire records into full records (multi-threaded); E, t, z agree with the device-computed fields of
 * is3d_sample to rounding (a few ulp) */
is3d_status is3d_expand_particles(const is3d_ctx *ctx, const is3d_particle_compact *compact, int64_t n, is3d_particle *out);
/* sampler self-test histograms (BinSampledParticle.cpp): filled on the device during is3d_sample when
 * test_sampler = 1.  Each output is Ns x bins row-major (vn: 7 x Ns x pT_bins); pass NULL to skip one. */
is3d_status is3d_sample_histograms(is3d_ctx *ctx, double *dN_dy, double *dN_deta, double *dN_dphipdy,
                                   double *dN_2pipTdpTdy, double *pT_count, double *vn_real, double *vn_imag,
                                   double *dN_taudtaudy, double *dN_twopirdrdy, double *dN_dphisdy);

/* ---- multi-GPU: cells sharded, ONE all-reduce (no reference counterpart: iS3D.cpp:81-286 is one OpenMP process) --- */
/* Every surface cell is an independent additive contribution to the spectra / dN/dX histograms / total yield and an
 * independent Poisson source of the sampler (MomentumSpectra.cpp:99-375 sums over icell), so each GPU takes a contiguous
 * block of cells (is3d_set_surface with global_offset) and the partial results are combined by one
 * ncclAllReduce(sum, double) over NVLink.  NCCL is resolved at run time (dlopen libnccl.so.2); single-GPU use needs none.
 *
 * (1) One process per GPU (torchrun / MPI): rank 0 calls is3d_comm_unique_id, the caller ships the bytes to all ranks, every
 *     rank calls is3d_comm_attach (collective).  Afterwards is3d_spectra[_device], is3d_dndx[_device], is3d_total_yield and
 *     is3d_polarization return the SUM over all ranks on every rank; is3d_cell_yields and is3d_sample stay per-rank (the
 *     sampler needs no exchange: its random streams are keyed by the global cell index). */
#define IS3D_COMM_ID_BYTES 128
is3d_status is3d_comm_unique_id(char id[IS3D_COMM_ID_BYTES]);
is3d_status is3d_comm_attach(is3d_ctx *ctx, const char id[IS3D_COMM_ID_BYTES], int nranks, int rank);
void        is3d_comm_detach(is3d_ctx *ctx);
int         is3d_comm_size(const is3d_ctx *ctx);                 /* 1 = no communicator attached */
int64_t     is3d_comm_collectives(const is3d_ctx *ctx);          /* all-reduces issued by this context so far */
const char *is3d_comm_last_error(void);                          /* message of a failed is3d_comm_unique_id / is3d_group_create */

/* (2) One process, several GPUs (the drop-in executable, an embedding C++ program): a group owns one context per device
 *     and the communicator over them.  Static inputs are set per context (is3d_group_ctx(g, i), same calls as above); the
 *     group calls split the surface into contiguous cell blocks (sizes differ by at most one cell), run one host thread per
 *     device and return the combined result; stats are summed over the devices (kernel_ms: the slowest device). */
typedef struct is3d_group is3d_group;
is3d_status is3d_group_create(const is3d_params *p, int ndev, const int *devices /* NULL = 0..ndev-1 */, is3d_group **out);
void        is3d_group_destroy(is3d_group *g);
int         is3d_group_size(const is3d_group *g);
is3d_ctx   *is3d_group_ctx(is3d_group *g, int i);
const char *is3d_group_last_error(const is3d_group *g);          /* g may be NULL: last create error */
void        is3d_group_cell_block(const is3d_group *g, int i, int64_t *begin, int64_t *count);
is3d_status is3d_group_set_surface(is3d_group *g, int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS], int64_t global_offset);
is3d_status is3d_group_set_vorticity(is3d_group *g, int64_t n, const double *const w[6]);
is3d_status is3d_group_spectra(is3d_group *g, double *out, is3d_stats *stats);
is3d_status is3d_group_dndx(is3d_group *g, double *tau_hist, double *r_hist, double *phi_hist, is3d_stats *stats);
is3d_status is3d_group_total_yield(is3d_group *g, double *ntotal, is3d_stats *stats);
is3d_status is3d_group_polarization(is3d_group *g, double *St, double *Sx, double *Sy, double *Sn, double *Snorm, is3d_stats *stats);
/* every device samples its cell block; the per-device lists are merged event by event in device order, which is the cell
 * order a single GPU produces (same hadrons, same order).  The merged list is library-owned: is3d_free_particles. */
is3d_status is3d_group_sample(is3d_group *g, int64_t nevents, is3d_particle **particles, int64_t *total, int64_t *counts,
                              is3d_stats *stats);
is3d_status is3d_group_sample_compact(is3d_group *g, int64_t nevents, is3d_particle_compact **particles, int64_t *total,
                                      int64_t *counts, is3d_stats *stats);
is3d_status is3d_group_sample_histograms(is3d_group *g, double *dN_dy, double *dN_deta, double *dN_dphipdy, double *dN_2pipTdpTdy,
                                         double *pT_count, double *vn_real, double *vn_imag, double *dN_taudtaudy,
                                         double *dN_twopirdrdy, double *dN_dphisdy);

/* ---- measurement helpers (no reference counterpart) ---------------------------------------------------------- */
/* sustained DFMA throughput of this GPU in TFLOP/s (2 flops per DFMA), measured with a register-resident
 * micro-kernel and CUDA events: the roofline denominator of the FP64-bound kernels. */
is3d_status is3d_measure_fp64_peak(is3d_ctx *ctx, double *tflops);
/* device-math probe: out_exp[i] = e^x[i], out_rcp[i] = 1/x[i], out_sqrt[i] = sqrt(x[i]) computed by the FP64-pipe
 * approximations the kernels use in place of exp / division / sqrt (csrc/common.cuh); host buffers of n doubles.
 * The parity tests bound their error against libm. */
is3d_status is3d_probe_math(is3d_ctx *ctx, int64_t n, const double *x, double *out_exp, double *out_rcp,
                            double *out_sqrt);
/* same for the angular primitives of the df_mode 5 (PTMA) solve: for x > 0, atan(sqrt x)/sqrt x, atanh(sqrt x)/sqrt x (x < 1;
 * 0 is returned for x >= 1) and ln x, computed by the FP64-pipe approximations of csrc/aniso.cuh */
is3d_status is3d_probe_aniso_math(is3d_ctx *ctx, int64_t n, const double *x, double *out_atan_over_s, double *out_atanh_over_s,
                                  double *out_log);
/* Host-only helper (no context, no GPU): the species classes and thread groups the spectra and dN/dX kernels use.
 * Hadrons with the same (mass, quantum-statistics sign[, baryon number when include_baryon]) form one class (their
 * Cooper-Frye integrands differ only by the degeneracy, which the reference multiplies in front:
 * MomentumSpectra.cpp:399-401); the classes are laid out in groups of slots_per_group slots with ONE baryon number per
 * group (padding slots = -1).  class_of[ns] receives the class of every species, slot_class[capacity] the class of
 * every slot, *nclass the number of classes.  Returns the number of slots written, -1 for bad arguments, -2 if capacity
 * is too small, -3 for a baryon number outside -2 .. 2 (the reference's PDG readers produce hadrons and the deuteron). */
int         is3d_species_groups(int ns, const double *mass, const double *sign, const double *baryon, int include_baryon,
                                int slots_per_group, int *class_of, int *slot_class, int capacity, int *nclass);
/* Host-only helper: the layout the spectra kernels use WITH baryon terms (include_baryon = 1).  A baryon class and its
 * antibaryon class (same mass, same statistics, opposite baryon number) form a charge-conjugate PAIR that is evaluated from one
 * exponential; every other class is a single.  single_slots = groups of slots_per_group class ids with one baryon number
 * per group; pair_slots = (class with b > 0, its partner) per slot, groups of slots_per_group slots with one |b|; padding =
 * -1.  Class ids are those of is3d_species_groups with include_baryon = 1.  Returns the number of classes, -1 / -2 / -3 as
 * is3d_species_groups. */
int         is3d_species_pairs(int ns, const double *mass, const double *sign, const double *baryon, int slots_per_group,
                               int *single_slots, int single_capacity, int *pair_slots, int pair_capacity, int *n_single, int *n_pair);
/* Host-only helper: the launch order of the spectra kernels' thread columns (no GPU).  slots[nslots] = the class ids of the thread
 * groups, ids_per_group each (-1 = padding), as is3d_species_groups / is3d_species_pairs return them; class_mass[nclass].  A column is
 * (group, pT node) = group * NpT + ip.  order[ngroups * NpT] receives the columns sorted by their smallest transverse mass
 * sqrt(min mass of the group^2 + pT^2) (stable), so that the threads_per_block columns of a block span a narrow mT range and agree
 * on which (cell, y) items are negligible (is3d_params.negligible_margin); bin_row[nclass * NpT] the block row
 * (position in order / threads_per_block) whose dropped-term bounds belong to bin (class, ip), -1 for classes not in slots.
 * Returns the number of columns, -1 for bad arguments. */
int         is3d_launch_order(int nslots, const int *slots, int ids_per_group, int nclass, const double *class_mass, int NpT,
                              const double *pT, int threads_per_block, int *order, int *bin_row);
/* device -> host copy on the context's stream (e.g. to inspect the list of is3d_sample_device without a CUDA runtime of one's own) */
is3d_status is3d_copy_from_device(is3d_ctx *ctx, void *host, const void *device, size_t bytes);
/* the CUDA stream all kernels of this context are launched on (as a void* cudaStream_t) */
void       *is3d_stream(is3d_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* IS3D_B200_H */
