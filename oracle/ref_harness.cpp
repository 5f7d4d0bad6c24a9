// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Harness `main` for the UNMODIFIED reference sources (/root/reference/src/cpp/*.cpp, compiled where
// they lie by oracle/Makefile into oracle/_ref/).  It drives the same sequence as
// IS3D::run_particlization(1) (reference src/cpp/iS3D.cpp:81-286) from the current working directory
// (which must hold iS3D_parameters.dat, input/surface.dat, PDG/, tables/, deltaf_coefficients/ and the
// results/ tree of clear_results.sh) and additionally dumps, in binary, the in-memory results that the
// reference's text writers truncate to 9 digits:
//   ref_dump/spectra.bin      EmissionFunctionArray::dN_pTdpTdphidy (EmissionFunction.h:114)
//   ref_dump/particles.bin    particle_event_list (EmissionFunction.h:121)
//   ref_dump/surface.bin      the freezeout surface as parsed by the reference's reader (IS3D_REF_SURFACE_ONLY=1 stops there)
//   ref_dump/species.bin      per-PDG-entry (mcid, mass, gspin, baryon, sign, neq, dn_bulk, dn_diff)
//   ref_dump/jonah.bin        PTB tables bulkPi/P, lambda^2, z (DeltafData.h:72-79), include_baryon = 0 only
//   ref_dump/timing.txt       seconds spent inside calculate_spectra (same region as the reference's
//                             own "Spectra calculation took" print, EmissionFunction.cpp:1375-1385)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <chrono>
#include <sys/stat.h>

#define private public
#include "iS3D.h"
#include "readindata.h"
#include "EmissionFunction.h"
#include "ParameterReader.h"
#include "DeltafData.h"
#include "Table.h"
#undef private

static void wr(FILE *f, const void *p, size_t n) { if (fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(1); } }

int main(int argc, char **argv)
{
  bool quiet_exit = true;   // skip destructors that segfault for include_baryon = 1 (DeltafData.cpp:49-63)
  (void)argc; (void)argv;
  mkdir("ref_dump", 0755);

  ParameterReader *paraRdr = new ParameterReader;
  paraRdr->readFromFile("iS3D_parameters.dat");
  int include_baryon = paraRdr->getVal("include_baryon");
  int operation = paraRdr->getVal("operation");

  FO_data_reader freeze_out_data(paraRdr, "input");
  long FO_length = freeze_out_data.get_number_cells();
  FO_surf *surf_ptr = new FO_surf[FO_length];
  freeze_out_data.read_freezeout_surface(surf_ptr);
  printf("Number of freezeout cells = %ld\n", FO_length);
  {
    // the surface exactly as the reference's readers leave it (FO_surf, readindata.h:79-91): 31 doubles per cell
    FILE *f = fopen("ref_dump/surface.bin", "wb");
    wr(f, &FO_length, sizeof(long));
    for(long i = 0; i < FO_length; i++)
    {
      const FO_surf &c = surf_ptr[i];
      double rec[31] = {c.tau, c.x, c.y, c.eta, c.dat, c.dax, c.day, c.dan, c.ux, c.uy, c.un, c.E, c.T, c.P, c.pixx, c.pixy, c.pixn,
                        c.piyy, c.piyn, c.bulkPi, c.muB, c.nB, c.Vx, c.Vy, c.Vn, c.wtx, c.wty, c.wtn, c.wxy, c.wxn, c.wyn};
      wr(f, rec, sizeof(rec));
    }
    fclose(f);
    if(getenv("IS3D_REF_SURFACE_ONLY")) { fflush(stdout); _Exit(0); }
  }

  particle_info *particle_data = new particle_info[Maxparticle];
  PDG_Data pdg(paraRdr);
  int Nparticle = pdg.read_resonances(particle_data);

  Table chosen_particles("PDG/chosen_particles.dat");

  Deltaf_Data *df_data = new Deltaf_Data(paraRdr);
  df_data->load_df_coefficient_data();
  if(!include_baryon)
  {
    df_data->construct_cubic_splines();
    df_data->compute_jonah_coefficients(particle_data, Nparticle);
  }
  df_data->compute_particle_densities(particle_data, Nparticle);
  df_data->test_df_coefficients(-0.1);

  {
    FILE *f = fopen("ref_dump/species.bin", "wb");
    long n = Nparticle; wr(f, &n, sizeof(long));
    for(int i = 0; i < Nparticle; i++)
    {
      double rec[8] = {(double)particle_data[i].mc_id, particle_data[i].mass, (double)particle_data[i].gspin,
                       (double)particle_data[i].baryon, (double)particle_data[i].sign,
                       particle_data[i].equilibrium_density, particle_data[i].bulk_density, particle_data[i].diff_density};
      wr(f, rec, sizeof(rec));
    }
    fclose(f);
    if(!include_baryon)
    {
      f = fopen("ref_dump/jonah.bin", "wb");
      long m = df_data->jonah_points; wr(f, &m, sizeof(long));
      wr(f, df_data->bulkPi_over_Peq_array, m * sizeof(double));
      wr(f, df_data->lambda_squared_array, m * sizeof(double));
      wr(f, df_data->z_array, m * sizeof(double));
      wr(f, &df_data->bulkPi_over_Peq_max, sizeof(double));
      fclose(f);
    }
  }

  Table pT_tab("tables/momentum/pT_table.dat");
  Table phi_tab("tables/momentum/phi_table.dat");
  Table y_tab("tables/momentum/y_table.dat");
  Table eta_tab("tables/spacetime_rapidity/eta_table.dat");

  EmissionFunctionArray efa(paraRdr, &chosen_particles, &pT_tab, &phi_tab, &y_tab, &eta_tab, particle_data, Nparticle, surf_ptr, FO_length, df_data);

  // exact mean total yield: call the reference's own public calculate_total_yield (EmissionFunction.h:172) on the
  // same structure-of-arrays unpack that calculate_spectra builds (EmissionFunction.cpp:998-1161); the reference only
  // prints it truncated to an integer (ParticleSampler.cpp:633)
  if(operation == 2)
  {
    int ns = efa.number_of_chosen_particles;
    std::vector<double> neq(ns), dnb(ns), dnd(ns);
    for(int i = 0; i < ns; i++)
    {
      int k = efa.chosen_particles_sampling_table[i];
      neq[i] = particle_data[k].equilibrium_density; dnb[i] = particle_data[k].bulk_density; dnd[i] = particle_data[k].diff_density;
    }
    std::vector<std::vector<double>> c(25, std::vector<double>(FO_length, 0.0));
    for(long i = 0; i < FO_length; i++)
    {
      FO_surf &f = surf_ptr[i];
      double v[25] = {f.tau, f.x, f.y, f.eta, f.dat, f.dax, f.day, f.dan, f.ux, f.uy, f.un, f.E, f.T, f.P, f.pixx, f.pixy, f.pixn,
                      f.piyy, f.piyn, f.bulkPi, 0, 0, 0, 0, 0};
      if(include_baryon) { v[20] = f.muB; v[21] = f.nB; v[22] = f.Vx; v[23] = f.Vy; v[24] = f.Vn; }
      for(int k = 0; k < 25; k++) c[k][i] = v[k];
    }
    Gauss_Laguerre *gla = new Gauss_Laguerre;
    gla->load_roots_and_weights("tables/gauss/gla_roots_weights.txt");
    double Ntot = efa.calculate_total_yield(neq.data(), dnb.data(), dnd.data(), c[12].data(), c[13].data(), c[11].data(), c[0].data(),
                                            c[8].data(), c[9].data(), c[10].data(), c[4].data(), c[5].data(), c[6].data(), c[7].data(),
                                            c[14].data(), c[15].data(), c[16].data(), c[17].data(), c[18].data(), c[19].data(),
                                            c[20].data(), c[21].data(), c[22].data(), c[23].data(), c[24].data(), df_data, gla);
    FILE *f = fopen("ref_dump/total_yield.bin", "wb");
    wr(f, &Ntot, sizeof(double));
    fclose(f);
    if(getenv("IS3D_REF_YIELD_ONLY")) { fflush(stdout); _Exit(0); }
  }

  std::vector<std::vector<Sampled_Particle>> events;
  auto t0 = std::chrono::steady_clock::now();
  efa.calculate_spectra(events);
  auto t1 = std::chrono::steady_clock::now();
  double secs = std::chrono::duration<double>(t1 - t0).count();
  {
    FILE *f = fopen("ref_dump/timing.txt", "w");
    fprintf(f, "%.9e\n", secs);
    fclose(f);
  }

  if(operation == 1)
  {
    FILE *f = fopen("ref_dump/spectra.bin", "wb");
    long dims[4] = {(long)efa.number_of_chosen_particles, efa.pT_tab_length, efa.phi_tab_length, efa.y_tab_length};
    wr(f, dims, sizeof(dims));
    wr(f, efa.dN_pTdpTdphidy, sizeof(double) * dims[0] * dims[1] * dims[2] * dims[3]);
    fclose(f);
  }
  if((int)paraRdr->getVal("mode") == 5)
  {
    // spin polarization (Polarization.cpp): the five arrays in the reference's STORAGE order, species fastest
    // (iS3D = ipart + npart * (ipT + NpT * (iphip + Nphi * iy)), :226)
    FILE *f = fopen("ref_dump/polarization.bin", "wb");
    long dims[4] = {(long)efa.number_of_chosen_particles, efa.pT_tab_length, efa.phi_tab_length, efa.y_tab_length};
    long ntot = dims[0] * dims[1] * dims[2] * dims[3];
    wr(f, dims, sizeof(dims));
    wr(f, efa.St, sizeof(double) * ntot); wr(f, efa.Sx, sizeof(double) * ntot); wr(f, efa.Sy, sizeof(double) * ntot);
    wr(f, efa.Sn, sizeof(double) * ntot); wr(f, efa.Snorm, sizeof(double) * ntot);
    fclose(f);
  }
  if(operation == 2)
  {
    FILE *g = fopen("ref_dump/nevents.txt", "w");
    fprintf(g, "%ld\n", efa.Nevents);
    fclose(g);
  }
  if(operation == 2 && !(int)paraRdr->getVal("test_sampler"))
  {
    FILE *f = fopen("ref_dump/particles.bin", "wb");
    long nev = (long)events.size(); wr(f, &nev, sizeof(long));
    for(long e = 0; e < nev; e++)
    {
      long np = (long)events[e].size(); wr(f, &np, sizeof(long));
      for(long i = 0; i < np; i++)
      {
        const Sampled_Particle &p = events[e][i];
        double rec[13] = {(double)p.chosen_index, (double)p.mcID, p.mass, p.tau, p.x, p.y, p.eta, p.t, p.z, p.E, p.px, p.py, p.pz};
        wr(f, rec, sizeof(rec));
      }
    }
    fclose(f);
  }
  fflush(stdout);
  if(quiet_exit) _Exit(0);
  return 0;
}
