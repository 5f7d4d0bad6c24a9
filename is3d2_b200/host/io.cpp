// ParameterReader, Table, quadrature/average loaders.  Behavioural mirror of reference
// src/cpp/ParameterReader.cpp, Table.cpp + Arsenal.cpp:51-137 and readindata.cpp:20-119, written from scratch.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "is3d_host.hpp"

namespace is3dhost {

static std::string g_root;

void set_root(const std::string &root)
{
  g_root = root;
  if (!g_root.empty() && g_root.back() != '/') g_root += '/';
}
std::string path(const std::string &relative) { return g_root + relative; }

double two_pi2_hbarC3() { return 2.0 * pow(M_PI, 2) * pow(hbarC, 3); }

std::vector<int> parse_device_list(const char *spec, int fallback)
{
  std::vector<int> out;
  if (!spec || !*spec) { out.push_back(fallback); return out; }
  std::string s(spec);
  if (s == "all") {
    const int n = is3d_device_count();
    for (int i = 0; i < (n > 0 ? n : 1); i++) out.push_back(i);
    return out;
  }
  size_t pos = 0;
  while (pos < s.size()) {
    size_t comma = s.find(',', pos);
    std::string tok = s.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
    size_t dash = tok.find('-');
    if (tok.empty()) fatal("IS3D_DEVICES: empty entry in \"" + s + "\"");
    if (dash != std::string::npos && dash > 0) {
      int a = atoi(tok.substr(0, dash).c_str()), b = atoi(tok.substr(dash + 1).c_str());
      if (b < a) fatal("IS3D_DEVICES: bad range \"" + tok + "\"");
      for (int i = a; i <= b; i++) out.push_back(i);
    } else {
      out.push_back(atoi(tok.c_str()));
    }
    if (comma == std::string::npos) break;
    pos = comma + 1;
  }
  if (out.empty()) out.push_back(fallback);
  return out;
}

void fatal(const std::string &message)
{
  printf("%s\n", message.c_str());
  fflush(stdout);
  exit(-1);
}

static std::string trim(const std::string &s)
{
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) a++;
  while (b > a && isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}
static std::string lower(std::string s)
{
  for (auto &c : s) c = (char)tolower((unsigned char)c);
  return s;
}
// first number on the line, 0 if none (stringToDouble, Arsenal.cpp:67-74)
static double first_double(const std::string &s)
{
  std::istringstream in(s + " ");
  double v = 0.0;
  in >> v;
  return v;
}

// ---- ParameterReader ------------------------------------------------------------------------------------------
long ParameterReader::find(const std::string &name) const
{
  std::string key = lower(trim(name));
  for (size_t i = 0; i < names_.size(); i++)
    if (names_[i] == key) return (long)i;
  return -1;
}
bool ParameterReader::exist(const std::string &name) const { return find(name) != -1; }
void ParameterReader::setVal(const std::string &name, double value)
{
  long i = find(name);
  if (i < 0) { names_.push_back(lower(trim(name))); values_.push_back(value); }
  else values_[i] = value;
}
double ParameterReader::getVal(const std::string &name) const
{
  long i = find(name);
  if (i < 0) {
    std::cout << "ParameterReader::getVal error: parameter with name " << name << " not found." << std::endl;
    exit(-1);
  }
  return values_[i];
}
void ParameterReader::phraseOneLine(const std::string &line, const std::string &commentSymbol)
{
  if (trim(line).empty()) return;
  std::string eq = line.substr(0, line.find(commentSymbol));
  if (trim(eq).empty()) return;
  size_t pos = eq.find('=');
  if (pos == std::string::npos) {
    std::cout << "ParameterReader::phraseEquationWithoutComments error: \"=\" symbol not found in equation assignment " << eq << std::endl;
    exit(-1);
  }
  setVal(eq.substr(0, pos), first_double(trim(eq.substr(pos + 1))));
}
void ParameterReader::readFromFile(const std::string &filename, const std::string &commentSymbol)
{
  std::ifstream f(path(filename).c_str());
  if (!f) {
    std::cout << "ParameterReader::readFromFile error: file " << filename << " does not exist." << std::endl;
    exit(-1);
  }
  std::string line;
  while (std::getline(f, line)) phraseOneLine(line, commentSymbol);
}
void ParameterReader::readFromArguments(long argc, char *argv[], const std::string &commentSymbol, long start_from)
{
  for (long i = start_from; i < argc; i++) phraseOneLine(argv[i], commentSymbol);
}
void ParameterReader::echo() const
{
  for (size_t i = 0; i < names_.size(); i++) std::cout << names_[i] << " = " << values_[i] << std::endl;
}

// ---- Table ----------------------------------------------------------------------------------------------------
// Row rule of the reference (readBlockData, Arsenal.cpp:79-125): the column count is fixed by the first line; a
// line becomes a row only if it is newline-terminated (a last line without '\n' is dropped).
void Table::loadTableFromFile(const std::string &filename)
{
  cols_.clear();
  rows_ = 0;
  FILE *f = fopen(path(filename).c_str(), "rb");
  if (!f) {
    std::cout << "Table::loadTableFromFile error: the data file " << filename << " cannot be opened." << std::endl;
    exit(-1);
  }
  std::vector<char> buf;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(sz + 1);
  if (sz > 0 && fread(buf.data(), 1, sz, f) != (size_t)sz) fatal("Table: short read of " + filename);
  fclose(f);
  buf[sz] = '\0';
  const char *p = buf.data(), *end = buf.data() + sz;
  size_t ncol = 0;
  while (p < end) {
    const char *nl = (const char *)memchr(p, '\n', end - p);
    if (!nl) break;                                   // unterminated last line: dropped
    std::vector<double> vals;
    const char *q = p;
    while (q < nl) {
      char *e = nullptr;
      double v = strtod(q, &e);
      if (e == q || e > nl) break;
      vals.push_back(v);
      q = e;
    }
    if (rows_ == 0) {
      if (vals.empty()) {
        std::cout << "readBlockData warning: input stream has empty first row; no data read" << std::endl;
        fatal("Table: " + filename + " has an empty first row");
      }
      ncol = vals.size();
      cols_.assign(ncol, {});
    }
    if (vals.size() < ncol) fatal("Table: " + filename + " has a short or blank row (undefined behaviour in the reference)");
    for (size_t c = 0; c < ncol; c++) cols_[c].push_back(vals[c]);
    rows_++;
    p = nl + 1;
  }
  if (rows_ == 0) fatal("Table: no rows in " + filename);
}

// ---- quadrature tables -----------------------------------------------------------------------------------------
void Gauss_Laguerre::load_roots_and_weights(const std::string &file_name)
{
  FILE *f = fopen(path(file_name).c_str(), "r");
  if (!f) fatal("load_roots_and_weights flag: couldn't open gauss laguerre file " + file_name);
  if (fscanf(f, "%d\t%d", &alpha, &points) != 2) fatal("bad gauss laguerre header");
  root.assign((size_t)alpha * points, 0.0);
  weight.assign((size_t)alpha * points, 0.0);
  int dummy;
  for (int i = 0; i < alpha; i++)
    for (int j = 0; j < points; j++)
      if (fscanf(f, "%d\t%lf\t%lf", &dummy, &root[(size_t)i * points + j], &weight[(size_t)i * points + j]) != 3)
        fatal("bad gauss laguerre row");
  fclose(f);
}
void Gauss_Legendre::load_roots_and_weights(const std::string &file_name)
{
  FILE *f = fopen(path(file_name).c_str(), "r");
  if (!f) fatal("load_roots_and_weights flag: couldn't open gauss legendre file " + file_name);
  if (fscanf(f, "%d", &points) != 1) fatal("bad gauss legendre header");
  root.assign(points, 0.0);
  weight.assign(points, 0.0);
  for (int i = 0; i < points; i++)
    if (fscanf(f, "%lf\t%lf", &root[i], &weight[i]) != 2) fatal("bad gauss legendre row");
  fclose(f);
}
void Plasma::load_thermodynamic_averages()
{
  FILE *f = fopen(path("tables/thermodynamic/average_thermodynamic_quantities.dat").c_str(), "r");
  if (!f) fatal("load_thermodynamic_averages flag: couldn't open average thermodynamic file");
  if (fscanf(f, "%lf\n%lf\n%lf\n%lf\n%lf", &temperature, &energy_density, &pressure, &baryon_chemical_potential,
             &net_baryon_density) != 5)
    fatal("bad average thermodynamic file");
  fclose(f);
}

}  // namespace is3dhost
