// brawl_host.cpp -- see brawl_host.hpp.  Host-side mirror of the BraWl drivers on top of the C ABI.
#include "brawl_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

namespace brawl {

// =====================================================================================================
// list-directed input helpers (Fortran `read(buffer, *)`)
// =====================================================================================================
static std::vector<std::string> ld_tokens(const std::string &s) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == ',' || s[i] == '\r')) i++;
    if (i >= s.size()) break;
    if (s[i] == '\'' || s[i] == '"') {
      char q = s[i++];
      std::string t;
      while (i < s.size() && s[i] != q) t += s[i++];
      i++;
      out.push_back(t);
    } else {
      std::string t;
      while (i < s.size() && s[i] != ' ' && s[i] != '\t' && s[i] != ',' && s[i] != '\r') t += s[i++];
      out.push_back(t);
    }
  }
  return out;
}
static bool ld_logical(const std::string &v, bool dflt) {
  auto t = ld_tokens(v);
  if (t.empty()) return dflt;
  std::string s = t[0];
  size_t k = (s.size() && s[0] == '.') ? 1 : 0;
  if (k < s.size() && (s[k] == 'T' || s[k] == 't')) return true;
  if (k < s.size() && (s[k] == 'F' || s[k] == 'f')) return false;
  return dflt;
}
static long ld_int(const std::string &v, long dflt) { auto t = ld_tokens(v); return t.empty() ? dflt : std::strtol(t[0].c_str(), nullptr, 10); }
static double ld_real(const std::string &v, double dflt) {
  auto t = ld_tokens(v);
  if (t.empty()) return dflt;
  std::string s = t[0];
  for (auto &c : s) if (c == 'd' || c == 'D') c = 'e';
  return std::strtod(s.c_str(), nullptr);
}
static std::string ld_string(const std::string &v, const std::string &dflt) { auto t = ld_tokens(v); return t.empty() ? dflt : t[0]; }

// label = text before the first '=', compared blank-padded: trailing blanks are insignificant, leading
// blanks break the key (io.f90:196-198)
static bool split_line(const std::string &line, std::string &label, std::string &value) {
  size_t pos = line.find('=');
  if (pos == std::string::npos) return false;
  label = line.substr(0, pos);
  while (!label.empty() && (label.back() == ' ' || label.back() == '\t')) label.pop_back();
  value = line.substr(pos + 1);
  return true;
}
static std::vector<std::string> read_lines(const std::string &filename, const char *missing_msg) {
  std::ifstream f(filename);
  if (!f) throw Stop(std::string(missing_msg) + filename);
  std::vector<std::string> lines;
  std::string l;
  while (std::getline(f, l)) lines.push_back(l);
  return lines;
}

int RunParams::lattice_id() const {
  if (lattice == "simple_cubic") return 0;
  if (lattice == "bcc") return 1;
  if (lattice == "fcc") return 2;
  throw Stop("Lattice type not yet implemented!");
}

RunParams read_control_file(const std::string &filename) {
  RunParams p;
  bool check[11] = {false};
  bool cs = false, counts = false;
  auto lines = read_lines(filename, "Could not find input file ");
  for (auto &line : lines) {
    std::string k, v;
    if (!split_line(line, k, v)) continue;
    if (k == "mode") { p.mode = (int)ld_int(v, p.mode); check[0] = true; }
    else if (k == "lattice") { p.lattice = ld_string(v, p.lattice); check[1] = true; }
    else if (k == "lattice_parameter") { p.lattice_parameter = ld_real(v, p.lattice_parameter); check[2] = true; }
    else if (k == "n_1") { p.n_1 = (int)ld_int(v, p.n_1); check[3] = true; }
    else if (k == "n_2") { p.n_2 = (int)ld_int(v, p.n_2); check[4] = true; }
    else if (k == "n_3") { p.n_3 = (int)ld_int(v, p.n_3); check[5] = true; }
    else if (k == "n_species") { p.n_species = (int)ld_int(v, p.n_species); check[6] = true; }
    else if (k == "interaction_file") { p.interaction_file = ld_string(v, p.interaction_file); check[7] = true; }
    else if (k == "interaction_range") { p.interaction_range = (int)ld_int(v, 0); check[8] = true; }
    else if (k == "wc_range") p.wc_range = (int)ld_int(v, p.wc_range);
    else if (k == "static_seed") p.static_seed = ld_logical(v, p.static_seed);
  }
  p.species_names.assign(p.n_species, "");
  p.species_concentrations.assign(p.n_species + 1, 0.0);
  p.species_numbers.assign(p.n_species, 0);
  for (auto &line : lines) {                       // second pass: the species arrays (io.f90:249-289)
    std::string k, v;
    if (!split_line(line, k, v)) continue;
    auto t = ld_tokens(v);
    if (k == "species_names") { for (int i = 0; i < p.n_species && i < (int)t.size(); i++) p.species_names[i] = t[i].substr(0, 2); check[9] = true; }
    else if (k == "species_concentrations") { for (int i = 0; i < p.n_species && i < (int)t.size(); i++) p.species_concentrations[i + 1] = ld_real(t[i], 0.0); cs = true; check[10] = true; }
    else if (k == "species_numbers") { for (int i = 0; i < p.n_species && i < (int)t.size(); i++) p.species_numbers[i] = std::strtoll(t[i].c_str(), nullptr, 10); counts = true; check[10] = true; }
  }
  if (cs && counts) throw Stop("You cannot specify both chemical concentrations and numbers of atoms!");
  if (!cs && !counts) throw Stop("You must specify either chemical concentrations or numbers of atoms.");
  static const char *names[10] = {"mode", "lattice", "lattice_parameter", "n_1", "n_2", "n_3", "n_species", "interaction_file",
                                  "interaction_range", "species_names"};
  for (int i = 0; i < 10; i++) if (!check[i]) throw Stop(std::string("Missing '") + names[i] + "' in system file");
  return p;
}

MetropolisParams read_metropolis_file(const std::string &filename) {
  MetropolisParams m;
  bool c_mode = false, c_n = false, c_s = false, c_T = false;
  for (auto &line : read_lines(filename, "Could not find Metropolis control file: ")) {
    std::string k, v;
    if (!split_line(line, k, v)) continue;
    if (k == "mode") { m.mode = ld_string(v, ""); c_mode = true; }
    else if (k == "n_mc_steps") { m.n_mc_steps = ld_int(v, 0); c_n = true; }
    else if (k == "burn_in_start") m.burn_in_start = ld_logical(v, false);
    else if (k == "burn_in") m.burn_in = ld_logical(v, false);
    else if (k == "n_burn_in_steps") m.n_burn_in_steps = ld_int(v, 0);
    else if (k == "calculate_energies") m.calculate_energies = ld_logical(v, true);
    else if (k == "n_sample_steps") { m.n_sample_steps = ld_int(v, 0); c_s = true; }
    else if (k == "calculate_asro") m.calculate_asro = ld_logical(v, true);
    else if (k == "n_sample_steps_asro") m.n_sample_steps_asro = ld_int(v, 0);
    else if (k == "calculate_alro") m.calculate_alro = ld_logical(v, false);
    else if (k == "n_sample_steps_alro") m.n_sample_steps_alro = ld_int(v, 0);
    else if (k == "n_sample_steps_trajectory") m.n_sample_steps_trajectory = ld_int(v, 0);
    else if (k == "write_trajectory_xyz") m.write_trajectory_xyz = ld_logical(v, false);
    else if (k == "write_trajectory_energy") m.write_trajectory_energy = ld_logical(v, false);
    else if (k == "write_trajectory_asro") m.write_trajectory_asro = ld_logical(v, false);
    else if (k == "write_initial_config_xyz") m.write_initial_config_xyz = ld_logical(v, false);
    else if (k == "write_initial_config_nc") m.write_initial_config_nc = ld_logical(v, false);
    else if (k == "write_final_config_xyz") m.write_final_config_xyz = ld_logical(v, false);
    else if (k == "write_final_config_nc") m.write_final_config_nc = ld_logical(v, false);
    else if (k == "read_start_config_nc") m.read_start_config_nc = ld_logical(v, false);
    else if (k == "start_config_file") m.start_config_file = ld_string(v, "");
    else if (k == "T") { m.T = ld_real(v, 0.0); c_T = true; }
    else if (k == "T_steps") m.T_steps = (int)ld_int(v, 1);
    else if (k == "delta_T") m.delta_T = ld_real(v, 1.0);
    else if (k == "nbr_swap") m.nbr_swap = ld_logical(v, false);
  }
  // io.f90:654-662 (the first assignment is the reference's typo: asro is overwritten when *alro* is 0)
  if (m.n_sample_steps_alro == 0) m.n_sample_steps_asro = m.n_sample_steps;
  if (m.n_sample_steps_alro == 0) m.n_sample_steps_alro = m.n_sample_steps;
  if (m.n_sample_steps_trajectory == 0) m.n_sample_steps_trajectory = m.n_sample_steps;
  if (!c_mode) throw Stop("Missing 'mode' in Metropolis input file");
  if (!c_n) throw Stop("Missing 'n_mc_steps' in Metropolis input file");
  if (!c_s) throw Stop("Missing 'n_sample_steps' in Metropolis input file");
  if (!c_T) throw Stop("Missing 'T' in Metropolis input file");
  if (m.burn_in) m.burn_in_start = true;            // io.f90:681-684
  return m;
}

NSParams read_ns_file(const std::string &filename) {
  NSParams n;
  for (auto &line : read_lines(filename, "Could not find nested sampling control file: ")) {
    std::string k, v;
    if (!split_line(line, k, v)) continue;
    if (k == "n_walkers") n.n_walkers = (int)ld_int(v, 0);
    else if (k == "n_steps") n.n_steps = (int)ld_int(v, 0);
    else if (k == "n_iter") n.n_iter = (int)ld_int(v, 0);
    else if (k == "traj_freq") n.traj_freq = (int)ld_int(v, 100);
    else if (k == "outfile_ener") n.outfile_ener = ld_string(v, "");
    else if (k == "outfile_traj") n.outfile_traj = ld_string(v, "");
  }
  if (n.n_walkers < 1 || n.n_steps < 1 || n.n_iter < 1) throw Stop("Missing parameter in nested sampling input file");
  return n;
}

void initialise_function_pointers(RunParams &s) {    // initialise.F90:153-257
  int lat = s.lattice_id();
  s.n_basis = 1;
  s.n_atoms = (lat == 0 ? 8 : lat == 1 ? 2 : 4) * s.n_1 * s.n_2 * s.n_3;
  int maxs = lat == 0 ? 2 : lat == 1 ? 10 : 6;
  if (s.interaction_range < 1 || s.interaction_range > maxs) throw Stop("Unsupported number of shells");
  double csum = 0.0;
  for (double c : s.species_concentrations) csum += c;
  int64_t nsum = 0;
  for (auto n : s.species_numbers) nsum += n;
  if (std::fabs(csum - 1.0) > 0.001 && nsum != s.n_atoms) throw Stop("Invalid numbers of atoms or concentrations specified");
}

std::vector<double> read_exchange(const RunParams &s) {      // io.f90:389-415: read(16,*) V_ex
  std::ifstream f(s.interaction_file);
  if (!f) throw Stop("Could not find interaction file " + s.interaction_file);
  size_t need = (size_t)s.n_species * s.n_species * s.interaction_range;
  std::vector<double> V;
  std::string tok;
  while (V.size() < need && (f >> tok)) {
    for (auto &c : tok) if (c == 'd' || c == 'D') c = 'e';
    V.push_back(std::strtod(tok.c_str(), nullptr));
  }
  if (V.size() < need) throw Stop("Interaction file " + s.interaction_file + " holds too few values");
  return V;
}

// =====================================================================================================
// MT19937
// =====================================================================================================
void MT19937::init_genrand(uint32_t s) {
  mt[0] = s;
  for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  mti = 624;
}
uint32_t MT19937::genrand_int32() {
  if (mti >= 624) {
    if (mti == 625) init_genrand(5489u);
    for (int k = 0; k < 624; k++) {
      uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    mti = 0;
  }
  uint32_t y = mt[mti++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
uint32_t MT19937::f90_init_genrand(int seedtime, int my_rank, unsigned long job_id) {
  unsigned long seed;
  if (seedtime) { seed = (unsigned long)time(nullptr); seed += 11 * my_rank; seed += job_id; if (seed % 2 == 0) seed += 1; }
  else seed = 110179 + 11 * my_rank;
  init_genrand((uint32_t)seed);
  return (uint32_t)seed;
}
void MT19937::export625(uint32_t *s) const { std::memcpy(s, mt, sizeof mt); s[624] = (uint32_t)mti; }
void MT19937::import625(const uint32_t *s) { std::memcpy(mt, s, sizeof mt); mti = (int)s[624]; }

// =====================================================================================================
// initial configuration and shells
// =====================================================================================================
static inline size_t gidx(const RunParams &s, int x, int y, int z) { return ((size_t)z * 2 * s.n_2 + y) * 2 * s.n_1 + x; }

void initial_setup(RunParams &s, Config &config, MT19937 &rng) {     // initialise.F90:434-617
  const int S = s.n_species, lat = s.lattice_id();
  const int64_t n_sites = (lat == 0 ? 8 : lat == 1 ? 2 : 4) * (int64_t)s.n_1 * s.n_2 * s.n_3;
  std::vector<int64_t> count(S), check(S, 0);
  int64_t nsum = 0;
  for (auto n : s.species_numbers) nsum += n;
  if (nsum == s.n_atoms) {
    for (int i = 0; i < S; i++) {
      count[i] = s.species_numbers[i];
      s.species_concentrations[i + 1] = (double)((float)count[i] / (float)s.n_atoms);   // single precision, :471
    }
  } else {
    int64_t tot = 0;
    for (int i = 0; i < S; i++) { count[i] = (int64_t)std::llround((double)(float)n_sites * s.species_concentrations[i + 1]); tot += count[i]; }
    int64_t round_err = tot - n_sites, inc = round_err > 0 ? -1 : 1, chk = round_err;
    while (chk != 0)
      for (int j = 0; j < S; j++) { if (count[j] == 0) continue; if (chk == 0) continue; count[j] += inc; chk += inc; }
  }
  int64_t tot = 0;
  for (auto c : count) tot += c;
  if (tot != n_sites) throw Stop("Error in initial_setup()");
  std::vector<double> cum(S + 1);
  for (int l = 0; l <= S; l++) { double c = 0.0; for (int i = 0; i <= l; i++) c += s.species_concentrations[i]; cum[l] = c; }
  const int gx = 2 * s.n_1, gy = 2 * s.n_2, gz = 2 * s.n_3;
  config.assign((size_t)gx * gy * gz, 0);
  auto fill = [&](int8_t &cell) {
    while (cell == 0) {
      double r = rng.genrand();
      for (int l = 1; l <= S; l++)
        if (r >= cum[l - 1] && r <= cum[l] && check[l - 1] < count[l - 1]) { cell = (int8_t)l; check[l - 1]++; }
    }
  };
  if (lat == 0) {
    for (int k = 1; k <= gz; k++) for (int j = 1; j <= gy; j++) for (int i = 1; i <= gx; i++) fill(config[gidx(s, i - 1, j - 1, k - 1)]);
  } else if (lat == 1) {
    for (int k = 1; k <= gz; k++) for (int j = 1; j <= gy / 2; j++) for (int i = 1; i <= gx / 2; i++)
      fill(config[gidx(s, 2 * i - (k % 2) - 1, 2 * j - (k % 2) - 1, k - 1)]);
  } else {
    for (int k = 1; k <= gz; k++) for (int j = 1; j <= gy; j++) for (int i = 1; i <= gx / 2; i++) {
      int x1 = 2 * i - (k % 2) * (j % 2) - ((k + 1) % 2) * ((j + 1) % 2);
      fill(config[gidx(s, x1 - 1, j - 1, k - 1)]);
    }
  }
}

std::vector<double> lattice_shells(const RunParams &s, const Config &config) {   // analytics.f90:205-275
  const int gx = 2 * s.n_1, gy = 2 * s.n_2, gz = 2 * s.n_3;
  std::vector<double> all((size_t)gx * gy * gz + 1, 0.0);
  size_t l = 0;
  for (int k = 0; k < gz; k++) for (int j = 0; j < gy; j++) for (int i = 0; i < gx; i++)
    if (config[gidx(s, i, j, k)] != 0) all[l++] = (double)std::sqrt((float)(k * k) + (float)(j * j) + (float)(i * i));
  std::sort(all.begin(), all.end());
  std::vector<double> shells(s.wc_range, 0.0);
  int n = 0;
  for (size_t i = 0; i + 1 < all.size() && n < s.wc_range; i++) {
    if (std::fabs(all[i] - all[i + 1]) < 1e-3) continue;
    shells[n++] = all[i];
  }
  return shells;
}

// =====================================================================================================
// text writers (Fortran edit descriptors of src/metropolis_output.f90)
// =====================================================================================================
void mkdir_p(const std::string &d) { ::mkdir(d.c_str(), 0777); }
static bool file_exists(const std::string &f) { struct stat st; return ::stat(f.c_str(), &st) == 0; }

void energy_trajectory_writer(const std::string &f, int64_t step, double energy) {     // "(I13,x,f20.10,x)"
  bool ex = file_exists(f);
  FILE *fp = std::fopen(f.c_str(), "a");
  if (!fp) throw Stop("cannot open " + f);
  if (!ex) std::fprintf(fp, " # step_number E\n");
  std::fprintf(fp, "%13lld %20.10f\n", (long long)step, energy);
  std::fclose(fp);
}
void asro_trajectory_writer(const std::string &f, int64_t step, const std::vector<double> &asro) {   // I13,x then (f8.5,x)...
  bool ex = file_exists(f);
  FILE *fp = std::fopen(f.c_str(), "a");
  if (!fp) throw Stop("cannot open " + f);
  if (!ex) std::fprintf(fp, " # step_number ASRO\n");
  std::fprintf(fp, "%13lld ", (long long)step);
  for (double v : asro) std::fprintf(fp, "%8.5f ", v);
  std::fprintf(fp, "\n");
  std::fclose(fp);
}
void diagnostics_writer(const std::string &f, const std::vector<double> &T, const std::vector<double> &E,
                        const std::vector<double> &C, const std::vector<double> &acc) {   // '(F8.1,2X,F24.15,2X,F24.15,2X,F6.4)'
  FILE *fp = std::fopen(f.c_str(), "w");
  if (!fp) throw Stop("cannot open " + f);
  std::fprintf(fp, " # T E C acceptance_rate\n");
  for (size_t i = 0; i < E.size(); i++) std::fprintf(fp, "%8.1f  %24.15f  %24.15f  %6.4f\n", T[i], E[i], C[i], acc[i]);
  std::fclose(fp);
}

// =====================================================================================================
// NetCDF-3 classic ("CDF\x01") writer/reader -- only what the reference's files need
// =====================================================================================================
namespace {
struct NcBuf {
  std::string b;
  void u32(uint32_t v) { char c[4] = {(char)(v >> 24), (char)(v >> 16), (char)(v >> 8), (char)v}; b.append(c, 4); }
  void pad() { while (b.size() % 4) b.push_back('\0'); }
  void name(const std::string &n) { u32((uint32_t)n.size()); b += n; pad(); }
  void f64(double d) { uint64_t u; std::memcpy(&u, &d, 8); for (int i = 7; i >= 0; i--) b.push_back((char)(u >> (8 * i))); }
  void i16(int16_t v) { b.push_back((char)((uint16_t)v >> 8)); b.push_back((char)v); }
};
struct NcAtt { std::string name; int type; std::vector<int32_t> iv; std::string sv; std::vector<double> dv; };
struct NcVar { std::string name; std::vector<int> dimids; int type; size_t nelem; };
enum { NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_DOUBLE = 6 };

std::string nc_header(const std::vector<std::pair<std::string, uint32_t>> &dims, const std::vector<NcAtt> &atts,
                      const std::vector<NcVar> &vars, std::vector<uint32_t> &begins) {
  // two passes: header length is needed for the `begin` offsets
  std::string out;
  begins.assign(vars.size(), 0);
  for (int pass = 0; pass < 2; pass++) {
    NcBuf h;
    h.b = "CDF\x01";
    h.u32(0);                                              // numrecs
    h.u32(0x0A); h.u32((uint32_t)dims.size());
    for (auto &d : dims) { h.name(d.first); h.u32(d.second); }
    if (atts.empty()) { h.u32(0); h.u32(0); }              // ABSENT
    else { h.u32(0x0C); h.u32((uint32_t)atts.size()); }
    for (auto &a : atts) {
      h.name(a.name); h.u32((uint32_t)a.type);
      if (a.type == NC_INT) { h.u32((uint32_t)a.iv.size()); for (auto v : a.iv) h.u32((uint32_t)v); }
      else if (a.type == NC_CHAR) { h.u32((uint32_t)a.sv.size()); h.b += a.sv; h.pad(); }
      else { h.u32((uint32_t)a.dv.size()); for (auto v : a.dv) h.f64(v); }
    }
    h.u32(0x0B); h.u32((uint32_t)vars.size());
    for (size_t i = 0; i < vars.size(); i++) {
      auto &v = vars[i];
      h.name(v.name); h.u32((uint32_t)v.dimids.size());
      for (int d : v.dimids) h.u32((uint32_t)d);
      h.u32(0); h.u32(0);                                  // no variable attributes
      h.u32((uint32_t)v.type);
      size_t vs = v.nelem * (v.type == NC_SHORT ? 2 : 8);
      vs = (vs + 3) & ~(size_t)3;
      h.u32((uint32_t)vs); h.u32(begins[i]);
    }
    if (pass == 0) {
      size_t off = h.b.size();
      for (size_t i = 0; i < vars.size(); i++) {
        begins[i] = (uint32_t)off;
        size_t vs = vars[i].nelem * (vars[i].type == NC_SHORT ? 2 : 8);
        off += (vs + 3) & ~(size_t)3;
      }
    } else out = h.b;
  }
  return out;
}
std::string rtrim(std::string s) { while (!s.empty() && s.back() == ' ') s.pop_back(); return s; }
}  // namespace

// =====================================================================================================
// xyz_writer (src/write_xyz.f90:39-116): extended XYZ -- particle count, Lattice="..." line, then one line per atom in the
// reference's loop order (x outermost, then y, then z: configuration(1, j, k, l) with l fastest) with
//   pos = ((j-1) a1 + (k-1) a2 + (l-1) a3 + basis) * lattice_parameter,  a_i = 0.5 e_i for every lattice the drivers support
// (initialise.F90:176-223: the grid is the doubled conventional cell).  The reference writes list-directed (`write(7,*)`);
// the field layout below follows gfortran's list-directed output (integer(4): I11 after the record's leading blank;
// real(8): 17 significant digits in a 25-wide field; real(4) zeros as 0.00000000 in a 16-wide field).  There is no golden
// .xyz in the reference's tests, so the exact whitespace is unpinned; the tokens (what extended-XYZ readers parse) are not.
// =====================================================================================================
static std::string ld_real64(double v) {                 // G25.17E3 as list-directed output uses it for 0.1 <= |v| < 1e17
  char b[64];
  if (v == 0.0) { std::snprintf(b, sizeof b, "%20.16f     ", 0.0); return b; }
  const double a = std::fabs(v);
  if (a >= 0.1 && a < 1e17) {
    const int intdigits = a >= 1.0 ? (int)std::floor(std::log10(a)) + 1 : 0;
    std::snprintf(b, sizeof b, "%20.*f     ", std::max(0, 17 - std::max(intdigits, 1)), v);
  } else std::snprintf(b, sizeof b, "%25.16E", v);
  return b;
}
void xyz_writer(const std::string &f, const Config &config, const RunParams &s, bool trajectory) {
  std::ofstream fp(f, trajectory ? std::ios::app : std::ios::trunc);
  if (!fp) throw Stop("cannot open " + f);
  const int gx = 2 * s.n_1, gy = 2 * s.n_2, gz = 2 * s.n_3;
  long n_particles = 0;
  for (int8_t v : config) n_particles += v != 0;
  char b[64];
  std::snprintf(b, sizeof b, " %11ld\n", n_particles);
  fp << b;
  const std::string z4 = "   0.00000000    ";
  fp << " Lattice=\"" << ld_real64(s.lattice_parameter * s.n_1) << z4 << z4 << z4 << ld_real64(s.lattice_parameter * s.n_2) << z4 << z4 << z4
     << ld_real64(s.lattice_parameter * s.n_3) << "\"\n";
  for (int x = 0; x < gx; x++) for (int y = 0; y < gy; y++) for (int z = 0; z < gz; z++) {
    const int8_t sp = config[((size_t)z * gy + y) * gx + x];
    if (sp == 0) continue;
    std::string name = sp <= (int)s.species_names.size() ? s.species_names[sp - 1] : "X";
    name.resize(2, ' ');                                                      // character(len=2)
    fp << " " << name << ld_real64(0.5 * x * s.lattice_parameter) << ld_real64(0.5 * y * s.lattice_parameter)
       << ld_real64(0.5 * z * s.lattice_parameter) << "\n";
  }
}

void ncdf_writer_1d(const std::string &f, const std::vector<double> &grid_data) {        // netcdf_io.f90:731-806
  std::vector<std::pair<std::string, uint32_t>> dims = {{"x", (uint32_t)grid_data.size()}};
  std::vector<NcVar> vars = {{"grid data", {0}, NC_DOUBLE, grid_data.size()}};
  std::vector<uint32_t> begins;
  NcBuf out;
  out.b = nc_header(dims, {}, vars, begins);
  for (double v : grid_data) out.f64(v);
  std::ofstream fp(f, std::ios::binary);
  if (!fp) throw Stop("cannot open " + f);
  fp.write(out.b.data(), (std::streamsize)out.b.size());
}

void ncdf_grid_state_writer(const std::string &f, const Config &state, const RunParams &s) {
  std::vector<std::pair<std::string, uint32_t>> dims = {{"b", (uint32_t)s.n_basis}, {"x", (uint32_t)(2 * s.n_1)},
                                                        {"y", (uint32_t)(2 * s.n_2)}, {"z", (uint32_t)(2 * s.n_3)}};
  std::vector<NcAtt> atts = {{"N_basis", NC_INT, {s.n_basis}, "", {}}, {"N_1", NC_INT, {s.n_1}, "", {}},
                             {"N_2", NC_INT, {s.n_2}, "", {}}, {"N_3", NC_INT, {s.n_3}, "", {}},
                             {"Number of Species", NC_INT, {s.n_species}, "", {}},
                             {"Lattice Type", NC_CHAR, {}, rtrim(s.lattice), {}},
                             {"Concentrations", NC_DOUBLE, {}, "", s.species_concentrations}};
  std::vector<NcVar> vars = {{"configuration", {3, 2, 1, 0}, NC_SHORT, state.size()}};   // Fortran (b,x,y,z) -> C order (z,y,x,b)
  std::vector<uint32_t> begins;
  NcBuf out;
  out.b = nc_header(dims, atts, vars, begins);
  for (int8_t v : state) out.i16((int16_t)v);
  out.pad();
  std::ofstream fp(f, std::ios::binary);
  if (!fp) throw Stop("cannot open " + f);
  fp.write(out.b.data(), (std::streamsize)out.b.size());
}

void ncdf_radial_density_writer(const std::string &f, const std::vector<double> &rho, const std::vector<double> &r,
                                const std::vector<double> &T, const std::vector<double> &U, const RunParams &s) {
  const uint32_t S = (uint32_t)s.n_species;
  std::vector<std::pair<std::string, uint32_t>> dims = {{"i", S}, {"j", S}, {"r", (uint32_t)s.wc_range}, {"T", (uint32_t)T.size()},
                                                        {"r_i", (uint32_t)r.size()}, {"T_i", (uint32_t)T.size()}, {"U_i", (uint32_t)U.size()}};
  std::vector<NcAtt> atts = {{"N_1", NC_INT, {s.n_1}, "", {}}, {"N_2", NC_INT, {s.n_2}, "", {}}, {"N_3", NC_INT, {s.n_3}, "", {}},
                             {"Number of Species", NC_INT, {s.n_species}, "", {}},
                             {"Lattice Type", NC_CHAR, {}, rtrim(s.lattice), {}},
                             {"Interaction file", NC_CHAR, {}, rtrim(s.interaction_file), {}},
                             {"Concentrations", NC_DOUBLE, {}, "", s.species_concentrations},
                             {"Warren-Cowley Range", NC_INT, {s.wc_range}, "", {}}};
  std::vector<NcVar> vars = {{"rho data", {3, 2, 1, 0}, NC_DOUBLE, rho.size()}, {"r data", {4}, NC_DOUBLE, r.size()},
                             {"T data", {5}, NC_DOUBLE, T.size()}, {"U data", {6}, NC_DOUBLE, U.size()}};
  std::vector<uint32_t> begins;
  NcBuf out;
  out.b = nc_header(dims, atts, vars, begins);
  for (double v : rho) out.f64(v);
  for (double v : r) out.f64(v);
  for (double v : T) out.f64(v);
  for (double v : U) out.f64(v);
  std::ofstream fp(f, std::ios::binary);
  if (!fp) throw Stop("cannot open " + f);
  fp.write(out.b.data(), (std::streamsize)out.b.size());
}

// ncdf_order_writer (netcdf_io.f90:362-480): site occupancies per temperature; the reference names the six dimensions
// of order(species, basis, x, y, z, T) "b","x","y","z","s","t" in that order (labels shifted by one -- kept).
void ncdf_order_writer(const std::string &f, const std::vector<double> &order, const std::vector<double> &T, const RunParams &s) {
  std::vector<std::pair<std::string, uint32_t>> dims = {{"b", (uint32_t)s.n_species}, {"x", (uint32_t)s.n_basis}, {"y", (uint32_t)(2 * s.n_1)},
                                                        {"z", (uint32_t)(2 * s.n_2)}, {"s", (uint32_t)(2 * s.n_3)}, {"t", (uint32_t)T.size()},
                                                        {"temp", (uint32_t)T.size()}};
  std::vector<NcAtt> atts = {{"N_Basis", NC_INT, {s.n_basis}, "", {}}, {"N_1", NC_INT, {s.n_1}, "", {}}, {"N_2", NC_INT, {s.n_2}, "", {}},
                             {"N_3", NC_INT, {s.n_3}, "", {}}, {"Number of Species", NC_INT, {s.n_species}, "", {}},
                             {"Lattice Type", NC_CHAR, {}, rtrim(s.lattice), {}},
                             {"Interaction file", NC_CHAR, {}, rtrim(s.interaction_file), {}},
                             {"Concentrations", NC_DOUBLE, {}, "", s.species_concentrations},
                             {"Warren-Cowley Range", NC_INT, {s.wc_range}, "", {}}};
  std::vector<NcVar> vars = {{"grid data", {5, 4, 3, 2, 1, 0}, NC_DOUBLE, order.size()}, {"temperature data", {6}, NC_DOUBLE, T.size()}};
  std::vector<uint32_t> begins;
  NcBuf out;
  out.b = nc_header(dims, atts, vars, begins);
  for (double v : order) out.f64(v);
  for (double v : T) out.f64(v);
  std::ofstream fp(f, std::ios::binary);
  if (!fp) throw Stop("cannot open " + f);
  fp.write(out.b.data(), (std::streamsize)out.b.size());
}

// Minimal classic reader for files written by ncdf_grid_state_writer (restart, netcdf_io.f90:1368-1429)
void ncdf_config_reader(const std::string &f, Config &config, const RunParams &s) {
  std::ifstream fp(f, std::ios::binary);
  if (!fp) throw Stop("Could not open start configuration file " + f);
  std::string d((std::istreambuf_iterator<char>(fp)), std::istreambuf_iterator<char>());
  if (d.size() < 8 || d.compare(0, 4, "CDF\x01") != 0) throw Stop("start configuration file is not NetCDF classic");
  auto rd = [&](size_t o) { return ((uint32_t)(uint8_t)d[o] << 24) | ((uint32_t)(uint8_t)d[o + 1] << 16) | ((uint32_t)(uint8_t)d[o + 2] << 8) | (uint8_t)d[o + 3]; };
  // locate the variable record by name; `begin` is the last word of its entry
  size_t p = d.find("configuration");
  if (p == std::string::npos) throw Stop("start configuration file holds no 'configuration' variable");
  size_t q = p + 16;                       // name padded to 16
  uint32_t nd = rd(q); q += 4 + 4 * nd;    // dimids
  q += 8;                                  // vatt_list ABSENT
  uint32_t type = rd(q), vsize = rd(q + 4), begin = rd(q + 8);
  size_t n = (size_t)8 * s.n_1 * s.n_2 * s.n_3;
  if (type != NC_SHORT || vsize < 2 * n || begin + 2 * n > d.size()) throw Stop("start configuration does not match the lattice in the input file");
  config.assign(n, 0);
  for (size_t i = 0; i < n; i++) config[i] = (int8_t)(int16_t)(((uint8_t)d[begin + 2 * i] << 8) | (uint8_t)d[begin + 2 * i + 1]);
}

// =====================================================================================================
// GPU handle
// =====================================================================================================
void Gpu::check(int rc) { if (rc) throw Stop(std::string("brawl_cuda: ") + brawl_cuda_last_error()); }
Gpu::Gpu(const RunParams &s, const std::vector<double> &V, int device, int n_replicas) {
  check(brawl_cuda_create(s.lattice_id(), s.n_1, s.n_2, s.n_3, s.n_species, s.interaction_range, V.data(), device, n_replicas, &h));
}
Gpu::~Gpu() { if (h) brawl_cuda_destroy(h); }

// radial_densities (analytics.f90:293-404) from the GPU pair counts; Fortran order r_densities(i,j,l)
static std::vector<double> radial_densities(Gpu &gpu, const RunParams &s, int replica = 0) {
  const int S = s.n_species;
  std::vector<int64_t> cnt((size_t)S * S * s.wc_range), sc(S);
  Gpu::check(brawl_cuda_radial_counts(gpu.h, replica, s.wc_range, cnt.data(), sc.data()));
  std::vector<double> rho(cnt.size());
  for (int l = 0; l < s.wc_range; l++) for (int j = 0; j < S; j++) for (int i = 0; i < S; i++)
    rho[((size_t)l * S + j) * S + i] = (double)cnt[((size_t)l * S + j) * S + i] / (double)sc[i];
  return rho;
}

static std::string temp_tag(double temp) {              // I4.4 , F2.1  ->  e.g. "0300.0"
  char b[32];
  int it = (int)temp;
  int frac = (int)std::lround((temp - it) * 10.0);
  std::snprintf(b, sizeof b, "%04d.%d", it, frac);
  return b;
}
static std::string rank_tag(int r) { char b[16]; std::snprintf(b, sizeof b, "%04d", r); return b; }

// =====================================================================================================
// metropolis_main / metropolis_simulated_annealing (src/metropolis.F90:46-558)
// =====================================================================================================
static void metropolis_decorrelated_samples(RunParams &setup, MetropolisParams &mp, const DriverOptions &opt);
void metropolis_main(RunParams &setup, MetropolisParams &mp, const DriverOptions &opt) {
  if (mp.mode == "decorrelated_samples") { metropolis_decorrelated_samples(setup, mp, opt); return; }   // metropolis.F90:59-69
  if (mp.mode != "simulated_annealing")
    throw Stop("Unrecognised mode '" + mp.mode + "' in Metropolis input file");
  const int S = setup.n_species, T_steps = mp.T_steps, p = opt.ranks;
  const bool replay = opt.rng == "mt19937";
  if (!replay && opt.rng != "philox") throw Stop("rng must be mt19937 or philox");
  if (mp.write_final_config_xyz || mp.write_final_config_nc || mp.write_initial_config_nc) mkdir_p("configs");
  if (mp.calculate_energies) mkdir_p("energies");
  if (mp.calculate_asro) mkdir_p("asro");
  if (mp.calculate_alro) mkdir_p("alro");                                 // :129-131
  if (mp.write_trajectory_energy || mp.write_trajectory_asro || mp.write_trajectory_xyz) mkdir_p("trajectories");   // :132-134
  if (mp.write_initial_config_xyz) mkdir_p("configs");
  std::vector<double> V = read_exchange(setup);
  const size_t nrho = (size_t)S * S * setup.wc_range;
  std::vector<double> av_E(T_steps, 0.0), av_C(T_steps, 0.0), av_acc(T_steps, 0.0), av_rho(nrho * T_steps, 0.0), temperature(T_steps);
  std::vector<double> shells;
  uint64_t offset = 0;
  for (int my_rank = 0; my_rank < p; my_rank++) {
    MT19937 rng;
    rng.f90_init_genrand(setup.static_seed ? 0 : 1, my_rank, 0);        // initialise_prng, initialise.F90:52-100
    Config config;
    if (mp.read_start_config_nc) ncdf_config_reader(mp.start_config_file, config, setup);
    else initial_setup(setup, config, rng);
    shells = lattice_shells(setup, config);
    Gpu gpu(setup, V, opt.device, 1);
    Gpu::check(brawl_cuda_set_config(gpu.h, 0, 1, config.data()));
    const int n_save_energy = (int)std::floor((float)mp.n_mc_steps / (float)mp.n_sample_steps);
    const int n_save_asro = (int)std::floor((float)mp.n_mc_steps / (float)mp.n_sample_steps_asro);
    std::vector<double> energies_of_T(T_steps), C_of_T(T_steps), acceptance_of_T(T_steps), rho_of_T(nrho * T_steps, 0.0);
    // ALRO (:164-165, 185): site occupancies accumulate on the device (brawl_cuda_store_state) and are cleared per
    // temperature (:196) by the reset flag of brawl_cuda_get_order
    const size_t norder = (size_t)setup.n_species * setup.n_basis * 8 * setup.n_1 * setup.n_2 * setup.n_3;
    const int n_save_alro = mp.calculate_alro ? (int)std::floor((float)mp.n_mc_steps / (float)mp.n_sample_steps_alro) : 1;
    std::vector<double> order_of_T(mp.calculate_alro ? norder * T_steps : 0, 0.0), order(mp.calculate_alro ? norder : 0, 0.0);
    uint32_t st[625];
    auto trials = [&](double beta, int64_t n) -> double {     // n x setup%mc_step; returns the acceptance increment
      if (n <= 0) return 0.0;
      if (replay) {
        int64_t acc = 0;
        rng.export625(st);
        Gpu::check(brawl_cuda_metropolis_replay(gpu.h, 0, beta, n, mp.nbr_swap ? 1 : 0, st, &acc));
        rng.import625(st);
        return (double)acc;
      }
      int64_t att = 0, acc = 0; double dE = 0.0;
      Gpu::check(brawl_cuda_metropolis_run(gpu.h, &beta, n, mp.nbr_swap ? 1 : 0, opt.seed + (uint64_t)my_rank, offset, &offset, &att, &acc, &dE));
      return att > 0 ? (double)acc * (double)n / (double)att : 0.0;   // box kernels round attempts up to whole phases
    };
    for (int j = 1; j <= T_steps; j++) {
      std::vector<double> r_densities(nrho, 0.0), asro(nrho, 0.0);
      double step_E = 0.0, step_Esq = 0.0, acceptance = 0.0, current_energy = 0.0;
      const double temp = mp.T + (double)(j - 1) * mp.delta_T, sim_temp = temp * k_b_in_Ry, beta = 1.0 / sim_temp;
      temperature[j - 1] = temp;
      if ((j == 1 && mp.burn_in_start) || (j > 1 && mp.burn_in)) acceptance += trials(beta, mp.n_burn_in_steps);   // :214-238
      const std::string tt = temp_tag(temp), rt = rank_tag(my_rank);
      const std::string efile = "trajectories/proc_" + rt + "_energy_trajectory_at_T_" + tt + ".dat";
      const std::string afile = "trajectories/proc_" + rt + "_asro_trajectory_at_T_" + tt + ".dat";
      if (mp.write_trajectory_energy) std::remove(efile.c_str());
      if (mp.write_trajectory_asro) std::remove(afile.c_str());
      if (mp.calculate_energies) {                                      // :295-300
        Gpu::check(brawl_cuda_total_energy(gpu.h, 0, 1, 1, &current_energy));
        if (mp.write_trajectory_energy) energy_trajectory_writer(efile, 0, current_energy);
      }
      if (mp.calculate_asro) {                                          // :303-308
        asro = radial_densities(gpu, setup);
        if (mp.write_trajectory_asro) asro_trajectory_writer(afile, 0, asro);
      }
      const std::string xyz_traj = "trajectories/proc_" + rt + "_trajectory_at_T_" + tt + ".xyz";            // :254-266
      if (mp.write_trajectory_xyz) {
        std::remove(xyz_traj.c_str());
        Gpu::check(brawl_cuda_get_config(gpu.h, 0, 1, config.data()));
        xyz_writer(xyz_traj, config, setup, true);                                                       // :310-313
      }
      if (mp.write_initial_config_xyz) {                                                                  // :320-326
        Gpu::check(brawl_cuda_get_config(gpu.h, 0, 1, config.data()));
        xyz_writer("configs/proc_" + rt + "_initial_config_at_T_" + tt + ".xyz", config, setup);
      }
      if (mp.write_initial_config_nc) {
        Gpu::check(brawl_cuda_get_config(gpu.h, 0, 1, config.data()));
        ncdf_grid_state_writer("configs/proc_" + rt + "_initial_config_at_T_" + tt + ".nc", config, setup);
      }
      const int64_t n_sweeps = mp.n_mc_steps / mp.n_sample_steps, n_sweep_steps = mp.n_mc_steps / n_sweeps;   // :343-344
      acceptance = 0.0;
      for (int64_t i = 1; i <= n_sweeps; i++) {
        acceptance += trials(beta, n_sweep_steps);                      // the k-loop :350-354
        const int64_t step_n = i * n_sweep_steps;
        if (mp.calculate_energies) {
          Gpu::check(brawl_cuda_total_energy(gpu.h, 0, 1, 1, &current_energy));   // exact order == total_energy
          step_E = step_E + current_energy;
          step_Esq = step_Esq + current_energy * current_energy;
          if (mp.write_trajectory_energy && step_n % mp.n_sample_steps_trajectory == 0) energy_trajectory_writer(efile, step_n, current_energy);
        }
        if (mp.calculate_asro) {
          if (step_n % mp.n_sample_steps_asro == 0) {
            asro = radial_densities(gpu, setup);
            for (size_t q = 0; q < nrho; q++) r_densities[q] = r_densities[q] + asro[q];
          }
          if (mp.write_trajectory_asro && step_n % mp.n_sample_steps_trajectory == 0) asro_trajectory_writer(afile, step_n, asro);
        }
        if (mp.calculate_alro && step_n % mp.n_sample_steps_alro == 0) Gpu::check(brawl_cuda_store_state(gpu.h, 0, 1));   // :394-399
        if (mp.write_trajectory_xyz && step_n % mp.n_sample_steps_trajectory == 0) {                                  // :401-407
          Gpu::check(brawl_cuda_get_config(gpu.h, 0, 1, config.data()));
          xyz_writer(xyz_traj, config, setup, true);
        }
      }
      acceptance_of_T[j - 1] = acceptance / (double)(float)mp.n_mc_steps;           // :412
      if (mp.calculate_energies) {
        energies_of_T[j - 1] = step_E / n_save_energy / setup.n_atoms;              // :416
        double C = (step_Esq / n_save_energy - (step_E / n_save_energy) * (step_E / n_save_energy)) / (sim_temp * temp) / setup.n_atoms;
        if (C < 0.0) C = 0.0;
        if (temp <= 0.0) C = 0.0;
        C_of_T[j - 1] = C;
      }
      if (mp.calculate_asro) for (size_t q = 0; q < nrho; q++) rho_of_T[(size_t)(j - 1) * nrho + q] = r_densities[q] / n_save_asro;
      if (mp.calculate_alro) {                                                      // :444-447
        Gpu::check(brawl_cuda_get_order(gpu.h, 0, order.data(), 1));
        for (size_t q = 0; q < norder; q++) order_of_T[(size_t)(j - 1) * norder + q] = order[q] / (double)(float)n_save_alro;
      }
      if (mp.write_final_config_xyz) {                                                                    // :450-456
        Gpu::check(brawl_cuda_get_config(gpu.h, 0, 1, config.data()));
        xyz_writer("configs/proc_" + rt + "_final_config_at_T_" + tt + ".xyz", config, setup);
      }
      if (mp.write_final_config_nc) {
        Gpu::check(brawl_cuda_get_config(gpu.h, 0, 1, config.data()));
        ncdf_grid_state_writer("configs/proc_" + rt + "_final_config_at_T_" + tt + ".nc", config, setup);
      }
      if (my_rank == 0) {
        std::printf(" Sampling at temperature %7.2f complete on process 0.\n Attempted%10lld trial Monte Carlo moves,\n of which %10lld were accepted,\n"
                    " corresponding to an acceptance rate of %7.2f %%\n Average internal energy was %7.2f meV/atom\n",
                    temp, (long long)mp.n_mc_steps, (long long)acceptance, 100.0 * acceptance / (double)mp.n_mc_steps,
                    13.606 * 1000 * energies_of_T[j - 1]);
      }
    }
    if (mp.calculate_energies) diagnostics_writer("energies/proc_" + rank_tag(my_rank) + "_energy_diagnostics.dat", temperature, energies_of_T, C_of_T, acceptance_of_T);
    if (mp.calculate_asro) ncdf_radial_density_writer("asro/proc_" + rank_tag(my_rank) + "_rho_of_T.nc", rho_of_T, shells, temperature, energies_of_T, setup);
    if (mp.calculate_alro) ncdf_order_writer("alro/proc_" + rank_tag(my_rank) + "_rho_of_T.nc", order_of_T, temperature, setup);   // :506-510
    for (int j = 0; j < T_steps; j++) { av_E[j] += energies_of_T[j]; av_C[j] += C_of_T[j]; av_acc[j] += acceptance_of_T[j]; }
    for (size_t q = 0; q < av_rho.size(); q++) av_rho[q] += rho_of_T[q];
  }
  if (p > 1) {                                             // comms_reduce_metropolis_results, comms.F90:122-160: SUM then /p
    for (int j = 0; j < T_steps; j++) { av_E[j] /= (double)p; av_C[j] /= (double)p; av_acc[j] /= (double)p; }
    for (auto &v : av_rho) v /= (double)p;
    if (mp.calculate_energies) diagnostics_writer("energies/av_energy_diagnostics.dat", temperature, av_E, av_C, av_acc);
    if (mp.calculate_asro) ncdf_radial_density_writer("asro/av_radial_density.nc", av_rho, shells, temperature, av_E, setup);
  }
}

// =====================================================================================================
// nested_sampling_main (src/nested_sampling.f90:45-206)
// =====================================================================================================
// gfortran list-directed real64 output: a 25-character field holding 17 significant digits.
// 0.1 <= |v| < 1e16: F form, right-aligned in 20 characters followed by 5 blanks (where the exponent
// would be); otherwise d.ddddddddddddddddE+ddd right-aligned in 25.
// =====================================================================================================
// metropolis_decorrelated_samples (src/metropolis.F90:572-738): burn in at every temperature of the ladder, then
// n_mc_steps trials at the LAST temperature, dumping the configuration as configs/proc_RRRR_config_NNNN_at_T_TTTT.T.xyz
// every n_sample_steps trials (the file index is i / n_sample_steps_asro, as in the reference).
// =====================================================================================================
static void metropolis_decorrelated_samples(RunParams &setup, MetropolisParams &mp, const DriverOptions &opt) {
  const bool replay = opt.rng == "mt19937";
  if (!replay && opt.rng != "philox") throw Stop("rng must be mt19937 or philox");
  mkdir_p("configs");
  if (mp.calculate_energies) mkdir_p("energies");
  if (mp.calculate_asro) mkdir_p("asro");
  if (mp.calculate_alro) mkdir_p("alro");
  if (mp.write_trajectory_energy || mp.write_trajectory_asro || mp.write_trajectory_xyz) mkdir_p("trajectories");
  std::vector<double> V = read_exchange(setup);
  uint64_t offset = 0;
  for (int my_rank = 0; my_rank < opt.ranks; my_rank++) {
    MT19937 rng;
    rng.f90_init_genrand(setup.static_seed ? 0 : 1, my_rank, 0);
    Config config;
    initial_setup(setup, config, rng);                                      // :588
    Gpu gpu(setup, V, opt.device, 1);
    Gpu::check(brawl_cuda_set_config(gpu.h, 0, 1, config.data()));
    uint32_t st[625];
    auto trials = [&](double beta, int64_t n) -> double {
      if (n <= 0) return 0.0;
      if (replay) {
        int64_t acc = 0;
        rng.export625(st);
        Gpu::check(brawl_cuda_metropolis_replay(gpu.h, 0, beta, n, mp.nbr_swap ? 1 : 0, st, &acc));
        rng.import625(st);
        return (double)acc;
      }
      int64_t att = 0, acc = 0; double dE = 0.0;
      Gpu::check(brawl_cuda_metropolis_run(gpu.h, &beta, n, mp.nbr_swap ? 1 : 0, opt.seed + (uint64_t)my_rank, offset, &offset, &att, &acc, &dE));
      return att > 0 ? (double)acc * (double)n / (double)att : 0.0;
    };
    double temp = mp.T, beta = 1.0 / (temp * k_b_in_Ry);
    for (int j = 1; j <= mp.T_steps; j++) {                                 // :641-671
      temp = mp.T + (double)(j - 1) * mp.delta_T;
      beta = 1.0 / (temp * k_b_in_Ry);
      if (mp.burn_in) {
        const double acceptance = trials(beta, mp.n_burn_in_steps);
        if (my_rank == 0)
          std::printf(" Burn-in complete at temperature %7.2f on process 0.\n Accepted %7d Monte Carlo moves at this temperature,\n"
                      " Corresponding to an acceptance rate of %7.2f %%\n\n", temp, (int)acceptance, 100.0 * acceptance / (float)mp.n_burn_in_steps);
      }
    }
    const int64_t n_samples = mp.n_mc_steps / mp.n_sample_steps;            // the i-loop (:676-703) in blocks of n_sample_steps
    for (int64_t k = 1; k <= n_samples; k++) {
      const double acceptance = trials(beta, mp.n_sample_steps);
      const int64_t i = k * mp.n_sample_steps;
      Gpu::check(brawl_cuda_get_config(gpu.h, 0, 1, config.data()));
      char name[128];
      std::snprintf(name, sizeof name, "configs/proc_%04d_config_%04d_at_T_%s.xyz", my_rank, (int)(i / mp.n_sample_steps_asro), temp_tag(temp).c_str());
      xyz_writer(name, config, setup);
      if (my_rank == 0) std::printf(" Accepted an additional %7d Monte Carlo moves before sample.\n\n", (int)acceptance);
    }
    trials(beta, mp.n_mc_steps - n_samples * mp.n_sample_steps);            // the tail of the loop writes nothing
  }
}

static std::string ld_real17_field(double v) {
  char b[64], out[64];
  double a = std::fabs(v);
  if (a != 0.0 && (a < 0.1 || a >= 1e16)) {
    std::snprintf(b, sizeof b, "%.16E", v);
    std::string s(b);
    size_t e = s.find('E');
    int ex = std::atoi(s.c_str() + e + 1);
    std::snprintf(out, sizeof out, "%25s", (s.substr(0, e) + (ex < 0 ? "E-" : "E+") + (std::abs(ex) < 10 ? "00" : std::abs(ex) < 100 ? "0" : "") + std::to_string(std::abs(ex))).c_str());
    return out;
  }
  int digits_before = a < 1.0 ? 0 : (int)std::floor(std::log10(a)) + 1;
  std::snprintf(b, sizeof b, "%.*f", 17 - digits_before, v);
  std::snprintf(out, sizeof out, "%20s     ", b);
  return out;
}

void nested_sampling_main(RunParams &setup, const DriverOptions &opt) {
  NSParams ns = read_ns_file("ns_input.inp");
  const bool replay = opt.rng == "mt19937";
  std::vector<double> V = read_exchange(setup);
  MT19937 rng;
  rng.f90_init_genrand(setup.static_seed ? 0 : 1, 0, 0);
  const int K = ns.n_walkers;
  Gpu gpu(setup, V, opt.device, K);
  FILE *f35 = std::fopen(ns.outfile_ener.c_str(), "w");
  if (!f35) throw Stop("cannot open " + ns.outfile_ener);
  const int n_at = setup.n_1 * setup.n_2 * setup.n_3 * setup.n_basis * setup.n_species;      // :84 (sic)
  std::fprintf(f35, "%12d%12d%12d False%12d\n", K, 1, 0, n_at);                                  // :87
  std::vector<double> walker_energies(K, 0.0);
  Config config;
  for (int w = 0; w < K; w++) {                                                                    // :78-97
    initial_setup(setup, config, rng);
    Gpu::check(brawl_cuda_set_config(gpu.h, w, 1, config.data()));
    double rnde = rng.genrand(), e = 0.0;
    Gpu::check(brawl_cuda_total_energy(gpu.h, w, 1, 1, &e));
    walker_energies[w] = e + rnde * (double)1e-8f;                                                 // default-real literal
  }
  int extra_steps = 0;
  int64_t n_acc = 0;
  uint32_t st[625];
  for (int it = 1; it <= ns.n_iter; it++) {
    int i_max = 0;
    for (int w = 1; w < K; w++) if (walker_energies[w] > walker_energies[i_max]) i_max = w;       // maxloc: first maximum
    const double ener_limit = walker_energies[i_max];
    std::fprintf(f35, "%12d %s\n", it, ld_real17_field(ener_limit).c_str());                      // write(35,*) i_iter, ener_limit
    if (it % (int)(K / 2.0) == 0)                                                                  // :129-144
      if (((float)n_acc < (float)n_at * 0.05f) && (extra_steps < ns.n_steps * 100)) extra_steps += ns.n_steps;
    const double rnd = rng.genrand();                                                              // :149
    int irnd = (int)std::ceil(rnd * K);
    if (irnd < 1) irnd = 1;
    Gpu::check(brawl_cuda_copy_replica(gpu.h, irnd - 1, i_max));                                   // :151
    walker_energies[i_max] = walker_energies[irnd - 1];
    if (replay) {
      rng.export625(st);
      Gpu::check(brawl_cuda_ns_walk_replay(gpu.h, i_max, &walker_energies[i_max], ener_limit, ns.n_steps + extra_steps, st, &n_acc));
      rng.import625(st);
    } else {
      int32_t id = i_max;
      Gpu::check(brawl_cuda_ns_walk(gpu.h, 1, &id, &walker_energies[i_max], &ener_limit, ns.n_steps + extra_steps, opt.seed, (uint64_t)it, &n_acc));
    }
  }
  std::fclose(f35);
}

}  // namespace brawl
