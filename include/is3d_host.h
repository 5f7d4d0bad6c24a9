/*
 * is3d_host.h -- C entry points of the host library (libis3d_host.so): the reference's driver sequence
 * IS3D::run_particlization (reference src/cpp/iS3D.cpp:81-286) split into steps so that an embedding program (or
 * a test) can hand over a surface in memory, as JETSCAPE does through IS3D::read_fo_surf_from_memory
 * (reference src/cpp/iS3D.h:80-103), and read results back without going through the text files.
 *
 * All paths are relative to `root` (a directory laid out like the reference's repository root:
 * iS3D_parameters.dat, PDG/, tables/, deltaf_coefficients/, input/, results/).
 * Errors follow the reference's convention: message on stdout and exit(-1).
 */
#ifndef IS3D_HOST_H
#define IS3D_HOST_H

#include "is3d_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct is3d_host is3d_host;

/* read iS3D_parameters.dat (optionally overriding "key = value" pairs, NULL-terminated list may be NULL) */
is3d_host *is3d_host_open(const char *root, const char *const *overrides);
void       is3d_host_close(is3d_host *h);

/* surface: either input/surface.dat through the reader selected by `mode`, or structure-of-arrays columns in
 * physical units (IS3D_COL_* order; baryon columns may be NULL).  Both write the thermodynamic-average side file. */
int64_t    is3d_host_read_surface(is3d_host *h);
int64_t    is3d_host_set_surface(is3d_host *h, int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS]);

/* Sharded surfaces (one cell block per rank): the surface averages behind the fast-mode densities and the PTB tables
 * must be those of the WHOLE surface.  thermo_sums returns this block's six additive sums (T, E, P, muB, nB weighted
 * by ds_max, and sum ds_max; readindata.cpp:330-360); after summing them over the ranks, set_thermo_averages writes
 * avg[k] = sum[k] / sum[5] to the side file the later stages read.  Call both between set/read_surface and prepare. */
void       is3d_host_thermo_sums(is3d_host *h, double sums6[6]);
void       is3d_host_set_thermo_averages(is3d_host *h, const double avg5[5]);

/* PDG + chosen particles + df tables + PTB tables + fast-mode densities + momentum tables, then the CUDA context */
void       is3d_host_prepare(is3d_host *h);
/* the table half of is3d_host_prepare only (no GPU needed): PDG, chosen particles, df/PTB tables, densities */
void       is3d_host_prepare_tables(is3d_host *h);
is3d_ctx  *is3d_host_context(is3d_host *h);      /* first (single-GPU runs: the only) context of the run */
is3d_group *is3d_host_group(is3d_host *h);       /* all contexts of the run (IS3D_DEVICES) and their communicator */

/* EmissionFunctionArray::calculate_spectra, writers included (results/ under root) */
void       is3d_host_run(is3d_host *h);

/* results of the last run kept in memory */
int64_t    is3d_host_spectra(is3d_host *h, const double **data, int64_t dims[4]);   /* Ns, NpT, Nphi, Ny */
int64_t    is3d_host_dndx(is3d_host *h, const double **tau, const double **r, const double **phi);
int64_t    is3d_host_events(is3d_host *h);
int64_t    is3d_host_event_particles(is3d_host *h, int64_t event, double *out13 /* n x 13 or NULL */);
double     is3d_host_seconds(is3d_host *h);
void       is3d_host_stats(is3d_host *h, is3d_stats *out);

/* introspection for tests: PDG table as parsed (n x 8: mcid mass gspin baryon sign neq dn_bulk dn_diff),
 * PTB tables (3 x 301 + max), surface columns as read */
int64_t    is3d_host_pdg(is3d_host *h, double *out8);
int64_t    is3d_host_ptb(is3d_host *h, double *x, double *lambda2, double *z, double *xmax);
int64_t    is3d_host_surface_column(is3d_host *h, int k, const double **data);
int64_t    is3d_host_chosen(is3d_host *h, int *mcid);

#ifdef __cplusplus
}
#endif
#endif
