// Hadron-resonance-gas particle tables.  Behaviour follows reference src/cpp/readindata.cpp:734-1252
// (read_resonances_conventional for UrQMD/SMASH with automatic antibaryons; read_resonances_smash_box with quantum
// numbers decoded from the Monte-Carlo id).  Decay channels are parsed and discarded: resonance decays are out of
// scope (reference iS3D_parameters.dat:86, "not finished").
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "is3d_host.hpp"

namespace is3dhost {

PDG_Data::PDG_Data(ParameterReader *paraRdr) { hrg_eos = paraRdr->getVal("hrg_eos"); }

int PDG_Data::read_resonances(std::vector<particle_info> &particle)
{
  switch (hrg_eos) {
    case 1: return read_resonances_conventional(particle, "PDG/pdg-urqmd_v3.3+.dat");
    case 2: return read_resonances_conventional(particle, "PDG/pdg_smash.dat");
    case 3: return read_resonances_smash_box(particle, "PDG/pdg_box.dat");
    default: fatal("\nread_resonances error: need to set hrg_eos = (1,2,3)");
  }
}

// Entry: mcid name mass width gspin baryon strange charm bottom gisospin charge ndecays, then ndecays rows of
// (mcid npart branching p1..p5).  Every baryon is followed by its antibaryon (readindata.cpp:1011-1030).
// The reference's `while(!eof)` loop parses one phantom entry after the last record and drops it again
// (Nparticle = local_i - 1); a file whose last token is not followed by whitespace therefore loses its final
// particle.  Both behaviours are reproduced.
int PDG_Data::read_resonances_conventional(std::vector<particle_info> &particle, const std::string &pdg_filename)
{
  std::ifstream in(path(pdg_filename).c_str());
  if (!in) fatal("couldn't open " + pdg_filename);
  particle.clear();
  bool ended_inside_record = false;
  while (true) {
    particle_info p;
    in >> p.mc_id;
    if (!in) break;                                     // the reference's phantom entry
    in >> p.name >> p.mass >> p.width >> p.gspin >> p.baryon >> p.strange >> p.charm >> p.bottom >> p.gisospin >> p.charge >> p.decays;
    int first_npart = 0;
    for (int j = 0; j < p.decays; j++) {
      int dummy_int, npart, part[5];
      double branch;
      in >> dummy_int >> npart >> branch >> part[0] >> part[1] >> part[2] >> part[3] >> part[4];
      if (j == 0) first_npart = npart;
    }
    p.stable = (first_npart == 1) ? 1 : 0;
    particle.push_back(p);
    if (p.baryon > 0) {
      particle_info a = p;
      a.mc_id = -p.mc_id;
      a.name = "Anti-baryon-" + p.name;
      a.baryon = -p.baryon; a.strange = -p.strange; a.charm = -p.charm; a.bottom = -p.bottom; a.charge = -p.charge;
      particle.push_back(a);
    }
    if (in.eof()) { ended_inside_record = true; break; }
    if ((int)particle.size() > Maxparticle - 2) fatal("Error: number of particles in file exceeds Maxparticle");
  }
  if (ended_inside_record && !particle.empty()) particle.pop_back();   // local_i - 1 without a phantom entry
  for (auto &p : particle) p.sign = (p.baryon % 2 == 0) ? -1 : 1;
  int meson = 0, baryon = 0, antibaryon = 0;
  for (auto &p : particle) { if (p.baryon == 0) meson++; else if (p.baryon > 0) baryon++; else antibaryon++; }
  if (baryon != antibaryon) printf("Error: (anti)baryons not paired correctly\n");
  printf("\nNumber of resonances = %d\n\n\t%d mesons\n\t%d baryons\n\t%d antibaryons\n\n", (int)particle.size(), meson, baryon, antibaryon);
  return (int)particle.size();
}

namespace {
// quantum numbers from a PDG Monte-Carlo id (hadrons only), cf. read_mcid (readindata.cpp:734-957)
struct McidInfo {
  bool is_deuteron, is_hadron, is_meson, is_baryon, has_antiparticle;
  int baryon, spin, gspin, sign;
};
McidInfo decode_mcid(long mcid)
{
  McidInfo m{};
  if (mcid < 0) printf("Error: should only be particles (not antiparticles) in pdg_test.dat\n");
  int digit[10];
  long x = labs(mcid);
  for (int i = 0; i < 10; i++) { digit[i] = (int)(x % 10); x /= 10; }
  int nJ = digit[0] + digit[7];                          // 8th digit extends the spin field
  int nq3 = digit[1], nq2 = digit[2], nq1 = digit[3];
  m.is_deuteron = (mcid == 1000010020);
  if (m.is_deuteron) printf("Error: there is a deuteron in HRG\n");
  m.is_hadron = (!m.is_deuteron && nq3 != 0 && nq2 != 0);
  m.is_meson = (m.is_hadron && nq1 == 0);
  m.is_baryon = (m.is_hadron && nq1 != 0);
  if (m.is_hadron) m.spin = (nJ == 0) ? 0 : nJ - 1;
  else if (m.is_deuteron) m.spin = 2;
  else { printf("Error: particle is not a deuteron or hadron\n"); m.spin = nq3; }
  if (m.is_hadron && nJ > 0) m.gspin = nJ;
  else if (m.is_deuteron) m.gspin = 3;
  else { printf("Error: particle is not a deuteron or hadron\n"); m.gspin = m.spin + 1; }
  if (m.is_deuteron) m.baryon = 2;
  else if (m.is_hadron) m.baryon = m.is_baryon ? 1 : 0;
  else { printf("Error: particle is not a deuteron or hadron\n"); m.baryon = 0; }
  if (m.is_deuteron) m.sign = -1;
  else if (m.is_hadron) m.sign = m.is_baryon ? 1 : -1;
  else { printf("Error: particle is not a deuteron or hadron\n"); m.sign = m.spin % 2; }
  if (m.is_hadron) m.has_antiparticle = ((m.baryon != 0) || (nq2 != nq3));
  else if (m.is_deuteron) m.has_antiparticle = true;
  else { printf("Error: particle is not a deuteron or hadron\n"); m.has_antiparticle = (nq3 == 1); }
  return m;
}
}  // namespace

// Line: name mass width parity mcid[0..3]; '#' or blank lines skipped; antiparticles appended (readindata.cpp:1098-1214)
int PDG_Data::read_resonances_smash_box(std::vector<particle_info> &particle, const std::string &pdg_filename)
{
  const int mcid_entries = 4;
  std::ifstream in(path(pdg_filename).c_str());
  if (!in) fatal("couldn't open " + pdg_filename);
  particle.clear();
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty() || line.at(0) == '#') continue;
    std::istringstream cur(line);
    std::string name;
    double mass = 0, width = 0;
    char parity = 0;
    long mc_id[mcid_entries] = {0, 0, 0, 0};
    cur >> name >> mass >> width >> parity;
    for (int k = 0; k < mcid_entries; k++) { long v = 0; if (cur >> v) mc_id[k] = v; else break; }
    for (int k = 0; k < mcid_entries; k++) {
      if (mc_id[k] == 0) continue;
      McidInfo info = decode_mcid(mc_id[k]);
      particle_info p;
      p.name = name; p.mass = mass; p.width = width; p.mc_id = mc_id[k];
      p.gspin = info.gspin; p.baryon = info.baryon; p.sign = info.sign;
      particle.push_back(p);
      if (info.has_antiparticle) {
        particle_info a = p;
        a.name = "Anti-" + name; a.mc_id = -mc_id[k]; a.baryon = -info.baryon;
        particle.push_back(a);
      }
    }
    if ((int)particle.size() > Maxparticle - 1) fatal("\nError: number of particles in file exceeds Maxparticle. Exiting...\n");
  }
  int meson = 0, baryon = 0, antibaryon = 0;
  for (auto &p : particle) { if (p.baryon == 0) meson++; else if (p.baryon > 0) baryon++; else antibaryon++; }
  if (baryon != antibaryon) printf("Error: (anti)baryons not paired correctly\n");
  printf("\nNumber of resonances = %d\n\n\t%d mesons\n\t%d baryons\n\t%d antibaryons\n\n", (int)particle.size(), meson, baryon, antibaryon);
  return (int)particle.size();
}

}  // namespace is3dhost
