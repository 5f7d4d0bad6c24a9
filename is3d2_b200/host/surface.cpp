// Freezeout-surface readers (modes 1/5/6/7) into a structure-of-arrays surface, plus the volume-weighted
// thermodynamic averages.  Column contracts and unit conversions follow reference src/cpp/readindata.cpp:167-729.
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "host_parallel.hpp"
#include "is3d_host.hpp"

namespace is3dhost {

void FO_surface::resize(int64_t n, bool with_vorticity)
{
  for (auto &c : col) c.assign((size_t)n, 0.0);
  for (auto &c : vorticity) c.assign(with_vorticity ? (size_t)n : 0, 0.0);
}

// ds_max-weighted averages, identical in all three readers (e.g. readindata.cpp:330-360)
void compute_thermodynamic_averages(const FO_surface &s, double avg[5])
{
  double sums[6];
  compute_thermodynamic_sums(s, sums);
  for (int k = 0; k < 5; k++) avg[k] = sums[k] / sums[5];
}

// the numerators (T, E, P, muB, nB weighted by ds_max) and the denominator (sum of ds_max) of the averages: additive
// over cell blocks, so ranks that each hold a block can all-reduce these six numbers to get the surface averages
void compute_thermodynamic_sums(const FO_surface &s, double sums[6])
{
  double T_avg = 0, E_avg = 0, P_avg = 0, muB_avg = 0, nB_avg = 0, max_volume = 0;
  const int64_t n = s.size();
  for (int64_t i = 0; i < n; i++) {
    double tau = s.col[IS3D_COL_TAU][i], tau2 = tau * tau;
    double ux = s.col[IS3D_COL_UX][i], uy = s.col[IS3D_COL_UY][i], un = s.col[IS3D_COL_UN][i];
    double ut = sqrt(1. + ux * ux + uy * uy + tau2 * un * un);
    double dat = s.col[IS3D_COL_DAT][i], dax = s.col[IS3D_COL_DAX][i], day = s.col[IS3D_COL_DAY][i], dan = s.col[IS3D_COL_DAN][i];
    double uds = ut * dat + ux * dax + uy * day + un * dan;
    double ds_ds = dat * dat - dax * dax - day * day - dan * dan / tau2;
    double ds_max = fabs(uds) + sqrt(fabs(uds * uds - ds_ds));
    max_volume += ds_max;
    E_avg += (s.col[IS3D_COL_E][i] * ds_max);
    T_avg += (s.col[IS3D_COL_T][i] * ds_max);
    P_avg += (s.col[IS3D_COL_P][i] * ds_max);
    muB_avg += (s.col[IS3D_COL_MUB][i] * ds_max);
    nB_avg += (s.col[IS3D_COL_NB][i] * ds_max);
  }
  sums[0] = T_avg; sums[1] = E_avg; sums[2] = P_avg; sums[3] = muB_avg; sums[4] = nB_avg; sums[5] = max_volume;
}

// 15 significant digits, no trailing newline (readindata.cpp:363-366); later stages re-read this file, so the
// rounding to 15 digits is part of the reference's arithmetic
void write_thermodynamic_averages(const double avg[5])
{
  std::ofstream f(path("tables/thermodynamic/average_thermodynamic_quantities.dat").c_str(), std::ios_base::out);
  f << std::setprecision(15) << avg[0] << "\n" << avg[1] << "\n" << avg[2] << "\n" << avg[3] << "\n" << avg[4];
}

FO_data_reader::FO_data_reader(ParameterReader *paraRdr, const std::string &)
{
  mode = paraRdr->getVal("mode");
  dimension = paraRdr->getVal("dimension");
  include_baryon = paraRdr->getVal("include_baryon");
}

// ---- parallel text ingestion -----------------------------------------------------------------------
// The reference reads surface.dat as one flat stream of numbers (`ifstream >> double`, readindata.cpp:222-295) after
// counting newline-terminated rows for the cell count (readindata.cpp:137-146).  At 10^7 cells that is ~2.5e8 numbers and
// 4-5 GB of text, minutes of serial strtod.  Here the file is read once, cut at whitespace into one slice per hardware
// thread, tokens are counted per slice (pass 1), and every slice is parsed into its place of the flat array with
// std::from_chars -- correctly rounded like strtod, several times faster (pass 2).  A token that is not a plain decimal
// number falls back to strtod; the first token that does not parse at all ends the stream, as `>>` would, and everything
// after it stays 0.
namespace {

// IS3D_READER_VERBOSE=1 prints the wall time of each ingestion phase
struct PhaseTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  bool on = getenv("IS3D_READER_VERBOSE") != nullptr;
  void lap(const char *what)
  {
    auto t1 = std::chrono::steady_clock::now();
    if (on) printf("[reader] %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};

inline bool is_space(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// at least ~1 MB of text per thread
int ingest_threads(size_t bytes) { return host_threads(bytes / (1 << 20) + 1, "IS3D_READER_THREADS"); }

// fn(i) for every row i in [0, n), rows split into contiguous blocks over the ingestion threads
template <class Fn>
void parallel_rows(long n, Fn fn)
{
  const int nt = ingest_threads((size_t)n * 256);
  parallel_for(nt, [&](int t) {
    const long b = n * t / nt, e = n * (t + 1) / nt;
    for (long i = b; i < e; i++) fn(i);
  });
}

// parses the token [p, e); returns 0 = whole token consumed, 1 = a number followed by junk (the NEXT extraction fails),
// 2 = not a number at all
int parse_token(const char *p, const char *e, double *out)
{
  const char *q = p;
  if (q < e && *q == '+') q++;                         // from_chars takes no leading '+'
  auto r = std::from_chars(q, e, *out);
  if (r.ec == std::errc() && r.ptr == e) return 0;
  char buf[128];
  size_t len = (size_t)(e - p);
  if (len >= sizeof(buf)) return 2;
  memcpy(buf, p, len);
  buf[len] = '\0';
  char *end = nullptr;
  double x = strtod(buf, &end);
  if (end == buf) return 2;
  *out = x;
  return (size_t)(end - buf) == len ? 0 : 1;
}

}  // namespace

FO_data_reader::~FO_data_reader() { unmap(); }

void FO_data_reader::unmap()
{
  if (map_ && map_size_) munmap(const_cast<char *>(map_), map_size_);
  map_ = nullptr;
  map_size_ = 0;
  mapped_ = false;
}

// number of cells = number of newline-terminated rows of input/surface.dat (Table rule, readindata.cpp:137-146).
// The file is memory-mapped (no copy through a read buffer); the slices below fault their pages in in parallel.
long FO_data_reader::get_number_cells()
{
  unmap();
  int fd = open(path("input/surface.dat").c_str(), O_RDONLY);
  if (fd < 0) fatal("Table::loadTableFromFile error: the data file input/surface.dat cannot be opened.");
  struct stat st;
  if (fstat(fd, &st) != 0) fatal("cannot stat input/surface.dat");
  const long sz = (long)st.st_size;
  if (sz > 0) {
    void *m = mmap(nullptr, (size_t)sz, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) fatal("cannot map input/surface.dat");
    madvise(m, (size_t)sz, MADV_WILLNEED);
    map_ = (const char *)m;
    map_size_ = (size_t)sz;
  }
  close(fd);
  mapped_ = true;
  const int nt = ingest_threads((size_t)sz);
  std::vector<long> part(nt, 0);
  const char *base = map_;
  parallel_for(nt, [&](int t) {
    const char *p = base + (size_t)sz * t / nt, *e = base + (size_t)sz * (t + 1) / nt;
    long rows = 0;
    while (p < e) {
      const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
      if (!q) break;
      rows++;
      p = q + 1;
    }
    part[t] = rows;
  });
  long rows = 0;
  for (long r : part) rows += r;
  number_of_cells = rows;
  return rows;
}

// the readers consume a flat stream of numbers (ifstream >> double), `columns` per cell
std::vector<double> FO_data_reader::slurp(long columns)
{
  if (!mapped_) get_number_cells();
  PhaseTimer timer;
  const size_t total = (size_t)number_of_cells * columns;
  std::vector<double> v(total, 0.0);
  timer.lap("allocate flat array");
  const char *base = map_;
  const size_t sz = map_size_;
  const int nt = ingest_threads(sz);
  // slice boundaries moved forward to the next whitespace so that no token straddles two slices
  std::vector<size_t> cut(nt + 1, sz);
  cut[0] = 0;
  for (int t = 1; t < nt; t++) {
    size_t pos = sz * t / nt;
    if (pos < cut[t - 1]) pos = cut[t - 1];
    while (pos < sz && !is_space(base[pos])) pos++;
    cut[t] = pos;
  }
  auto count_tokens = [&](const char *p, const char *e) {
    size_t n = 0;
    while (p < e) {
      while (p < e && is_space(*p)) p++;
      if (p >= e) break;
      n++;
      while (p < e && !is_space(*p)) p++;
    }
    return n;
  };
  // index of the first token of every slice.  Fast path: a well-formed file has exactly `columns` tokens per
  // newline-terminated row, so first[t] = rows before the slice x columns + tokens of the partial row in front of it --
  // one memchr pass for the newlines and a scan of at most one row per slice.  The parse below verifies the guess.
  std::vector<size_t> first(nt + 1, 0), parsed(nt, 0), fail(nt, (size_t)-1);
  {
    std::vector<size_t> rows(nt, 0);
    parallel_for(nt, [&](int t) {
      const char *p = base + cut[t], *e = base + cut[t + 1];
      size_t r = 0;
      while (p < e) {
        const char *q = (const char *)memchr(p, '\n', (size_t)(e - p));
        if (!q) break;
        r++;
        p = q + 1;
      }
      rows[t] = r;
    });
    size_t rows_before = 0;
    for (int t = 0; t < nt; t++) {
      size_t row_start = cut[t];
      while (row_start > 0 && base[row_start - 1] != '\n') row_start--;
      first[t] = rows_before * (size_t)columns + count_tokens(base + row_start, base + cut[t]);
      rows_before += rows[t];
    }
    first[nt] = (size_t)-1;                              // the last slice may hold an unterminated partial row
  }
  timer.lap("slice offsets (fast path)");
  auto parse_all = [&]() {
    parallel_for(nt, [&](int t) {
      const char *p = base + cut[t], *e = base + cut[t + 1];
      size_t k = first[t], n = 0;
      fail[t] = (size_t)-1;
      while (p < e) {
        while (p < e && is_space(*p)) p++;
        if (p >= e) break;
        const char *q = p;
        while (q < e && !is_space(*q)) q++;
        n++;
        if (k < total && fail[t] == (size_t)-1) {
          double x = 0.0;
          int rc = parse_token(p, q, &x);
          if (rc == 2) fail[t] = k;                      // the stream ends here; keep counting tokens for the check
          else { v[k++] = x; if (rc == 1) fail[t] = k; }
        }
        p = q;
      }
      parsed[t] = n;
    });
  };
  parse_all();
  timer.lap("parse");
  bool consistent = true;
  for (int t = 0; t + 1 < nt; t++) consistent = consistent && (first[t] + parsed[t] == first[t + 1]);
  if (!consistent) {
    // rows of uneven length: exact token offsets from the counts just taken, then parse again
    for (int t = 0; t < nt; t++) first[t + 1] = first[t] + parsed[t];
    std::fill(v.begin(), v.end(), 0.0);
    parse_all();
    timer.lap("re-parse (uneven rows)");
  }
  size_t stop = (size_t)-1;
  for (int t = 0; t < nt; t++) if (fail[t] < stop) stop = fail[t];
  if (stop != (size_t)-1)
    for (size_t k = stop; k < total; k++) v[k] = 0.0;
  unmap();                                             // the text is not needed again
  timer.lap("unmap");
  return v;
}

// ---- optional binary cache of the parsed surface ("parse once, cache as binary SoA") ------------------
// IS3D_SURFACE_CACHE=1 keeps the structure-of-arrays columns, as the reader left them, in input/surface.dat.soa next to the
// text.  The header records the reader settings and the size and modification time of surface.dat; a cache that does not
// match is ignored and rewritten.  Off by default: the reference writes no such file.
namespace {

struct CacheHeader {
  char magic[8];                 // "IS3DSOA1"
  int64_t n;
  int32_t mode, dimension, include_baryon, with_vorticity;
  int64_t source_size, source_mtime_ns;
};

bool source_stamp(int64_t *size, int64_t *mtime_ns)
{
  struct stat st;
  if (stat(path("input/surface.dat").c_str(), &st) != 0) return false;
  *size = (int64_t)st.st_size;
  *mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + (int64_t)st.st_mtim.tv_nsec;
  return true;
}

}  // namespace

bool FO_data_reader::load_cache(FO_surface &s)
{
  CacheHeader want{}, have{};
  if (!source_stamp(&want.source_size, &want.source_mtime_ns)) return false;
  FILE *f = fopen(path("input/surface.dat.soa").c_str(), "rb");
  if (!f) return false;
  bool ok = fread(&have, sizeof(have), 1, f) == 1 && memcmp(have.magic, "IS3DSOA1", 8) == 0 && have.n == number_of_cells &&
            have.mode == mode && have.dimension == dimension && have.include_baryon == include_baryon &&
            have.source_size == want.source_size && have.source_mtime_ns == want.source_mtime_ns;
  if (ok) {
    s.resize(have.n, have.with_vorticity != 0);
    for (int k = 0; k < IS3D_SURFACE_COLUMNS && ok; k++) ok = fread(s.col[k].data(), sizeof(double), (size_t)have.n, f) == (size_t)have.n;
    if (have.with_vorticity)
      for (int k = 0; k < 6 && ok; k++) ok = fread(s.vorticity[k].data(), sizeof(double), (size_t)have.n, f) == (size_t)have.n;
  }
  fclose(f);
  return ok;
}

void FO_data_reader::store_cache(const FO_surface &s)
{
  CacheHeader h{};
  memcpy(h.magic, "IS3DSOA1", 8);
  h.n = s.size(); h.mode = mode; h.dimension = dimension; h.include_baryon = include_baryon;
  h.with_vorticity = s.vorticity[0].empty() ? 0 : 1;
  if (!source_stamp(&h.source_size, &h.source_mtime_ns)) return;
  FILE *f = fopen(path("input/surface.dat.soa").c_str(), "wb");
  if (!f) return;                                        // read-only input directory: run without a cache
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  for (int k = 0; k < IS3D_SURFACE_COLUMNS && ok; k++) ok = fwrite(s.col[k].data(), sizeof(double), (size_t)h.n, f) == (size_t)h.n;
  if (h.with_vorticity)
    for (int k = 0; k < 6 && ok; k++) ok = fwrite(s.vorticity[k].data(), sizeof(double), (size_t)h.n, f) == (size_t)h.n;
  fclose(f);
  if (!ok) remove(path("input/surface.dat.soa").c_str());
}

void FO_data_reader::read_freezeout_surface(FO_surface &surf)
{
  const char *env = getenv("IS3D_SURFACE_CACHE");
  const bool use_cache = env && atoi(env) != 0;
  if (use_cache && mapped_ && load_cache(surf)) {
    unmap();
    printf("Freezeout surface taken from the binary cache input/surface.dat.soa\n");
    double avg[5];
    compute_thermodynamic_averages(surf, avg);
    if (mode == 7) avg[4] = 0.0;
    write_thermodynamic_averages(avg);
    return;
  }
  if (mode == 1 || mode == 5) read_surface_cpu_vh(surf);
  else if (mode == 6) read_surface_music(surf);
  else if (mode == 7) read_surface_hic_eventgen(surf);
  if (use_cache) store_cache(surf);
}

// mode 1: t x y n ds_t ds_x ds_y ds_n u^x u^y u^n E T P pi^xx pi^xy pi^xn pi^yy pi^yn Pi [muB nB V^x V^y V^n] [6 wbar]
void FO_data_reader::read_surface_cpu_vh(FO_surface &s)
{
  const long ncol = 20 + (include_baryon ? 5 : 0) + (mode == 5 ? 6 : 0);
  std::vector<double> v = slurp(ncol);
  const long n = number_of_cells;
  s.resize(n, mode == 5);
  parallel_rows(n, [&](long i) {
    const double *r = &v[(size_t)i * ncol];
    for (int k = 0; k <= IS3D_COL_UN; k++) s.col[k][i] = r[k];          // tau..un unchanged
    s.col[IS3D_COL_E][i] = r[11] * hbarC;
    s.col[IS3D_COL_T][i] = r[12] * hbarC;
    s.col[IS3D_COL_P][i] = r[13] * hbarC;
    s.col[IS3D_COL_PIXX][i] = r[14] * hbarC;
    s.col[IS3D_COL_PIXY][i] = r[15] * hbarC;
    s.col[IS3D_COL_PIXN][i] = r[16] * hbarC;
    s.col[IS3D_COL_PIYY][i] = r[17] * hbarC;
    s.col[IS3D_COL_PIYN][i] = r[18] * hbarC;
    s.col[IS3D_COL_BULKPI][i] = r[19] * hbarC;
    long k = 20;
    if (include_baryon) {
      s.col[IS3D_COL_MUB][i] = r[20] * hbarC;
      s.col[IS3D_COL_NB][i] = r[21];
      s.col[IS3D_COL_VX][i] = r[22];
      s.col[IS3D_COL_VY][i] = r[23];
      s.col[IS3D_COL_VN][i] = r[24];
      k = 25;
    }
    if (mode == 5) for (int w = 0; w < 6; w++) s.vorticity[w][i] = r[k + w];
    if (dimension == 2 && s.col[IS3D_COL_ETA][i] != 0) s.col[IS3D_COL_ETA][i] = 0;   // readindata.cpp:311-319
  });
  double avg[5];
  compute_thermodynamic_averages(s, avg);
  write_thermodynamic_averages(avg);
}

// mode 6 (MUSIC): t x y n ds_t/t ds_x/t ds_y/t ds_n/t u^t u^x u^y t.u^n E T muB muS muC (E+P)/T pi^tt pi^tx pi^ty
// t.pi^tn pi^xx pi^xy t.pi^xn pi^yy t.pi^yn t2.pi^nn Pi [nB V^t V^x V^y t.V^n]
void FO_data_reader::read_surface_music(FO_surface &s)
{
  const long ncol = 29 + (include_baryon ? 5 : 0);
  std::vector<double> v = slurp(ncol);
  const long n = number_of_cells;
  s.resize(n);
  parallel_rows(n, [&](long i) {
    const double *r = &v[(size_t)i * ncol];
    double tau = r[0];
    s.col[IS3D_COL_TAU][i] = tau; s.col[IS3D_COL_X][i] = r[1]; s.col[IS3D_COL_Y][i] = r[2]; s.col[IS3D_COL_ETA][i] = r[3];
    s.col[IS3D_COL_DAT][i] = r[4] * tau; s.col[IS3D_COL_DAX][i] = r[5] * tau;
    s.col[IS3D_COL_DAY][i] = r[6] * tau; s.col[IS3D_COL_DAN][i] = r[7] * tau;
    s.col[IS3D_COL_UX][i] = r[9]; s.col[IS3D_COL_UY][i] = r[10]; s.col[IS3D_COL_UN][i] = r[11] / tau;
    double E = r[12] * hbarC, T = r[13] * hbarC;
    s.col[IS3D_COL_E][i] = E; s.col[IS3D_COL_T][i] = T;
    s.col[IS3D_COL_MUB][i] = r[14] * hbarC;
    s.col[IS3D_COL_P][i] = r[17] * T - E;
    s.col[IS3D_COL_PIXX][i] = r[22] * hbarC;
    s.col[IS3D_COL_PIXY][i] = r[23] * hbarC;
    s.col[IS3D_COL_PIXN][i] = r[24] * hbarC / tau;
    s.col[IS3D_COL_PIYY][i] = r[25] * hbarC;
    s.col[IS3D_COL_PIYN][i] = r[26] * hbarC / tau;
    s.col[IS3D_COL_BULKPI][i] = r[28] * hbarC;
    if (include_baryon) {
      s.col[IS3D_COL_NB][i] = r[29];
      s.col[IS3D_COL_VX][i] = r[31];
      s.col[IS3D_COL_VY][i] = r[32];
      s.col[IS3D_COL_VN][i] = r[33] / tau;
    }
    if (dimension == 2 && s.col[IS3D_COL_ETA][i] != 0) s.col[IS3D_COL_ETA][i] = 0;
  });
  double avg[5];
  compute_thermodynamic_averages(s, avg);
  write_thermodynamic_averages(avg);
}

// mode 7 (HIC-EventGen, 2+1d, GeV units): t x y n ds_t/t ds_x/t ds_y/t ds_n/t v^x v^y t.v^n pi^tt pi^tx pi^ty t.pi^tn
// pi^xx pi^xy t.pi^xn pi^yy t.pi^yn t2.pi^nn Pi T E P muB
void FO_data_reader::read_surface_hic_eventgen(FO_surface &s)
{
  if (dimension != 2) fatal("read_surface_hic_eventgen error: HIC-EventGen surface must be 2+1d (set dimension = 2)");
  if (include_baryon) fatal("read_surface_hic_eventgen error: HIC-EventGen has no baryon chemical potential (set include_baryon = 0)");
  const long ncol = 26;
  std::vector<double> v = slurp(ncol);
  const long n = number_of_cells;
  s.resize(n);
  parallel_rows(n, [&](long i) {
    const double *r = &v[(size_t)i * ncol];
    double tau = r[0];
    s.col[IS3D_COL_TAU][i] = tau; s.col[IS3D_COL_X][i] = r[1]; s.col[IS3D_COL_Y][i] = r[2]; s.col[IS3D_COL_ETA][i] = 0;
    s.col[IS3D_COL_DAT][i] = r[4] * tau; s.col[IS3D_COL_DAX][i] = r[5] * tau; s.col[IS3D_COL_DAY][i] = r[6] * tau;
    s.col[IS3D_COL_DAN][i] = 0;
    double vx = r[8], vy = r[9];
    double ut = 1. / sqrt(fabs(1. - vx * vx - vy * vy));
    s.col[IS3D_COL_UX][i] = ut * vx; s.col[IS3D_COL_UY][i] = ut * vy; s.col[IS3D_COL_UN][i] = 0;
    s.col[IS3D_COL_PIXX][i] = r[15]; s.col[IS3D_COL_PIXY][i] = r[16]; s.col[IS3D_COL_PIXN][i] = 0;
    s.col[IS3D_COL_PIYY][i] = r[18]; s.col[IS3D_COL_PIYN][i] = 0;
    s.col[IS3D_COL_BULKPI][i] = r[21];
    s.col[IS3D_COL_T][i] = r[22]; s.col[IS3D_COL_E][i] = r[23]; s.col[IS3D_COL_P][i] = r[24];
    s.col[IS3D_COL_MUB][i] = r[25];
  });
  // the reference averages with its local ut (from v) and nB = 0; ut from u^x,u^y is the same number up to rounding
  double avg[5];
  compute_thermodynamic_averages(s, avg);
  avg[4] = 0.0;
  write_thermodynamic_averages(avg);
}

}  // namespace is3dhost
