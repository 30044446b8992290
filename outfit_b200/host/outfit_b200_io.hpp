// outfit_b200_io.hpp -- header-only C++17 host side for the data formats either side of the path
// (SURVEY 8f rows 1 and 4): what the reference delegates to photom / hifitime, restated so a C++ caller
// can go from files to an OutfitObsBatch and from LSQ records to Keplerian elements without Python.
// The same functions exist in Python (outfit_b200/{mpc80,ut1,elements}.py); tests/test_cpp_io.py checks
// the two against each other.
//
//   MPC 80-column optical astrometry  -> outfit::Observation rows        (reference input: tests/data/*.obs)
//   JPL latest_eop2.long              -> UT1 epochs for pvobs             observer_extension.rs:191-192
//   JPL DE binary file                -> Chebyshev table                  horizon_data.rs:239-292,598-684
//   equinoctial LSQ record            -> Keplerian elements + covariance  keplerian_element.rs:185-233,
//                                        equinoctial_element.rs:1049-1140, uncertainty.rs:412-416
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "outfit_b200.hpp"

namespace outfit {
namespace io {

constexpr double kArcsec = 3.14159265358979323846 / 648000.0;
constexpr double kAuKm = 149597870.7;
constexpr double kErau = 6378.137 / kAuKm;  // Earth equatorial radius in AU (observer_extension.rs:159-171)
constexpr double kTwoPi = 6.283185307179586476925286766559;

// Gregorian calendar date (day may carry a fraction) -> MJD
inline double calendar_to_mjd(int year, int month, double day) {
  const int a = (14 - month) / 12;
  const int y = year + 4800 - a, m = month + 12 * a - 3;
  const int d = (int)day;
  const long jdn = d + (153 * m + 2) / 5 + 365L * y + y / 4 - y / 100 + y / 400 - 32045;
  return (double)(jdn - 2400001) + (day - d);
}

// TAI - UTC (s) from the given MJD on (IERS Bulletin C)
inline int tai_minus_utc(double mjd_utc) {
  static const struct { double mjd; int s; } leap[] = {
      {41317.0, 10}, {41499.0, 11}, {41683.0, 12}, {42048.0, 13}, {42413.0, 14}, {42778.0, 15}, {43144.0, 16},
      {43509.0, 17}, {43874.0, 18}, {44239.0, 19}, {44786.0, 20}, {45151.0, 21}, {45516.0, 22}, {46247.0, 23},
      {47161.0, 24}, {47892.0, 25}, {48257.0, 26}, {48804.0, 27}, {49169.0, 28}, {49534.0, 29}, {50083.0, 30},
      {50630.0, 31}, {51179.0, 32}, {53736.0, 33}, {54832.0, 34}, {56109.0, 35}, {57204.0, 36}, {57754.0, 37}};
  int out = 0;
  for (const auto &l : leap)
    if (mjd_utc >= l.mjd) out = l.s;
  return out;
}
inline double utc_to_tt(double mjd_utc) { return mjd_utc + (tai_minus_utc(mjd_utc) + 32.184) / 86400.0; }

struct Mpc80Record {
  std::string number, designation, obscode;
  bool discovery = false;
  double mjd_utc = 0, ra = 0, dec = 0, mag = NAN;
  char band = ' ';
};

// One 80-column record; false for blank / non-optical (satellite, radar, roving) lines.
inline bool parse_mpc80_line(const std::string &line_in, Mpc80Record &r) {
  std::string line = line_in;
  while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
  if (line.size() < 80) return false;
  const char note2 = line[14];
  if (std::string("RrVvSs").find(note2) != std::string::npos) return false;
  auto trim = [](std::string s) {
    const size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(' ');
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  };
  auto num = [&](size_t pos, size_t len, double &out) {
    const std::string s = trim(line.substr(pos, len));
    if (s.empty()) return false;
    char *end = nullptr;
    out = std::strtod(s.c_str(), &end);
    return end && *end == '\0';
  };
  double year, month, day, rh, rm, rs, dd, dm, ds;
  if (!num(15, 4, year) || !num(20, 2, month) || !num(23, 9, day) || !num(32, 2, rh) || !num(35, 2, rm) ||
      !num(38, 6, rs) || !num(45, 2, dd) || !num(48, 2, dm) || !num(51, 5, ds))
    return false;
  const double sign = line[44] == '-' ? -1.0 : 1.0;
  r.number = trim(line.substr(0, 5));
  r.designation = trim(line.substr(5, 7));
  r.discovery = line[12] == '*';
  r.mjd_utc = calendar_to_mjd((int)year, (int)month, day);
  r.ra = (rh + rm / 60.0 + rs / 3600.0) * 15.0 * 3.14159265358979323846 / 180.0;
  r.dec = sign * (dd + dm / 60.0 + ds / 3600.0) * 3.14159265358979323846 / 180.0;
  double mag;
  r.mag = num(65, 5, mag) ? mag : NAN;
  r.band = line[70];
  r.obscode = line.substr(77, 3);
  return true;
}

// All optical records of an 80-column file in file order, grouped by object id (number, else designation)
// -- or, with single_trajectory, as ONE trajectory named after the last record (the reference's quick start).
inline std::vector<std::pair<std::string, std::vector<Mpc80Record>>> parse_mpc80(const std::string &text,
                                                                                 bool single_trajectory = false) {
  std::vector<std::pair<std::string, std::vector<Mpc80Record>>> out;
  std::map<std::string, size_t> index;
  std::istringstream in(text);
  std::string ln;
  while (std::getline(in, ln)) {
    Mpc80Record r;
    if (!parse_mpc80_line(ln, r)) continue;
    const std::string id = r.number.empty() ? r.designation : r.number;
    auto it = index.find(id);
    if (it == index.end()) {
      index[id] = out.size();
      out.push_back({id, {}});
      it = index.find(id);
    }
    out[it->second].second.push_back(r);
  }
  if (single_trajectory && !out.empty()) {
    std::vector<Mpc80Record> all;
    for (auto &kv : out)
      for (auto &r : kv.second) all.push_back(r);
    const Mpc80Record &last = all.back();
    return {{last.number.empty() ? last.designation : last.number, all}};
  }
  return out;
}

// MPC parallax constants (east longitude in degrees, rho cos phi', rho sin phi') -> Earth-fixed position in AU
inline void body_fixed_position(double lon_deg, double rho_cos, double rho_sin, double out[3]) {
  const double lon = lon_deg * 3.14159265358979323846 / 180.0;
  out[0] = kErau * rho_cos * std::cos(lon);
  out[1] = kErau * rho_cos * std::sin(lon);
  out[2] = kErau * rho_sin;
}

// hifitime's Ut1Provider restated (un-vendored; parity with the crate unpinned): JPL EOP2 namelist, step
// lookup "last entry strictly earlier than the epoch", integer nanoseconds.
class Ut1Table {
 public:
  static constexpr int64_t kNsPerDay = 86400000000000LL;
  static Ut1Table from_eop2_text(const std::string &text) {
    Ut1Table t;
    std::istringstream in(text);
    std::string ln;
    bool ignore = true;
    while (std::getline(in, ln)) {
      if (!ln.empty() && ln.back() == '\r') ln.pop_back();
      if (ln == " EOP2=") { ignore = false; continue; }
      if (ln == " $END") break;
      if (ignore) continue;
      std::vector<std::string> cols;
      std::istringstream ls(ln);
      std::string c;
      while (std::getline(ls, c, ',')) cols.push_back(c);
      if (cols.size() < 4) throw Error(OUTFIT_E_INVALID_ARGUMENT, "EOP2 data line with fewer than 4 columns");
      t.epoch_ns_.push_back(days_to_ns(std::strtod(cols[0].c_str(), nullptr)));
      t.offset_ns_.push_back((int64_t)std::llround(std::strtod(cols[3].c_str(), nullptr) * 1e6));
    }
    if (ignore) throw Error(OUTFIT_E_INVALID_ARGUMENT, "no ' EOP2=' marker: not a JPL EOP2 file");
    return t;
  }
  size_t size() const { return epoch_ns_.size(); }
  int64_t offset_ns_at(int64_t tai_ns) const {  // TAI - UT1 of Epoch::ut1_offset
    for (size_t i = epoch_ns_.size(); i-- > 0;)
      if (tai_ns > epoch_ns_[i]) return offset_ns_[i];
    return 0;
  }
  // Epoch::from_mjd_in_time_scale(mjd_tt, TT).to_ut1(self).to_mjd_tai_days()
  double mjd_ut1(double mjd_tt) const {
    const int64_t tai = days_to_ns(mjd_tt) - 32184000000LL;
    const int64_t ut1 = tai - offset_ns_at(tai);
    int64_t whole = ut1 / kNsPerDay, frac = ut1 % kNsPerDay;
    if (frac < 0) { frac += kNsPerDay; --whole; }
    return (double)whole + (double)frac / (double)kNsPerDay;
  }

 private:
  static int64_t days_to_ns(double d) { return (int64_t)std::llround(d * (double)kNsPerDay); }
  std::vector<int64_t> epoch_ns_, offset_ns_;
};

// Records of one trajectory -> Observation rows (TT epochs, constant sigma, UT1 = UTC + dut1 or the table's)
struct Site { double lon_deg, rho_cos, rho_sin; };

// MPC observatory list (ObsCodes.html / obscodes.txt) -> parallax constants by code.  Fixed columns: code 1-3,
// east longitude 5-13, rho cos phi' 14-21, rho sin phi' 22-30; space-based / roving entries (blank constants)
// are skipped.  Same rule as outfit_b200/mpc80.py: parse_obscodes.
inline std::map<std::string, Site> parse_obscodes(const std::string &text) {
  std::map<std::string, Site> out;
  std::istringstream in(text);
  std::string ln;
  auto number = [](std::string f, double &v) {
    f.erase(std::remove(f.begin(), f.end(), ' '), f.end());
    if (f.empty()) return false;
    char *end = nullptr;
    v = std::strtod(f.c_str(), &end);
    return end && *end == '\0';
  };
  while (std::getline(in, ln)) {
    if (ln.size() < 30 || ln[0] == '<' || ln.compare(0, 4, "Code") == 0) continue;
    Site s;
    if (!number(ln.substr(4, 9), s.lon_deg) || !number(ln.substr(13, 8), s.rho_cos) || !number(ln.substr(21, 9), s.rho_sin)) continue;
    out[ln.substr(0, 3)] = s;
  }
  return out;
}
inline std::vector<Observation> to_observations(const std::vector<Mpc80Record> &recs,
                                                const std::map<std::string, Site> &sites, double sigma_arcsec = 0.5,
                                                double dut1_s = 0.0, const Ut1Table *ut1 = nullptr) {
  std::vector<Observation> out;
  for (const Mpc80Record &r : recs) {
    auto it = sites.find(r.obscode);
    if (it == sites.end()) throw Error(OUTFIT_E_INVALID_ARGUMENT, "unknown observatory code " + r.obscode);
    Observation o;
    o.mjd_tt = utc_to_tt(r.mjd_utc);
    o.ra = r.ra; o.dec = r.dec;
    o.sigma_ra = o.sigma_dec = sigma_arcsec * kArcsec;
    body_fixed_position(it->second.lon_deg, it->second.rho_cos, it->second.rho_sin, o.body_fixed);
    o.mjd_ut1 = ut1 ? ut1->mjd_ut1(o.mjd_tt) : r.mjd_utc + dut1_s / 86400.0;
    out.push_back(o);
  }
  return out;
}

// ---- JPL DE binary ephemeris file (little-endian linux_p1550p2650.440 family) -----------------------
// Layout as the reference parses it (HorizonData::read_horizon_file,
// jpl_ephem/horizon/horizon_data.rs:239-251, 270-292, 336-368, 598-684): SS[3] f64 at byte 2652 (start JD,
// end JD, days per record), NCON i32 at 2676, EMRAT f64 at 2688, IPT[12][3] u32 + NUMDE u32 + LPT[3] u32 at
// 2696, IPT[13], IPT[14] after the constant names when NCON > 400; record size = (4 + sum 2 n_coeff n_sub dim)
// * 4 bytes; data records from byte 2 * recsize, IPT offsets 1-based.  Rows used: 2 (EMB), 9 (Moon), 10 (Sun).
struct DeTable {
  std::vector<double> cheb;      // [n_blocks][block_stride]
  size_t n_blocks = 0, block_stride = 0;
  double jd_start = 0, jd_end = 0, block_days = 0, emrat = 0;
  uint32_t ipt[3][3] = {};       // EMB, Moon, Sun: {0-based offset in a block, n_coeff, n_sub}
  uint32_t numde = 0;
};
inline DeTable read_de_binary(const std::string &path) {
  std::FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) throw Error(OUTFIT_E_INVALID_ARGUMENT, "cannot open " + path);
  std::vector<unsigned char> buf;
  unsigned char chunk[1 << 16];
  size_t got;
  while ((got = std::fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
  std::fclose(f);
  if (buf.size() < 4096) throw Error(OUTFIT_E_INVALID_ARGUMENT, "DE file shorter than its header");
  auto f64 = [&](size_t off) { double v; std::memcpy(&v, &buf[off], 8); return v; };
  auto u32 = [&](size_t off) { uint32_t v; std::memcpy(&v, &buf[off], 4); return v; };
  DeTable t;
  t.jd_start = f64(2652); t.jd_end = f64(2660); t.block_days = f64(2668);
  int32_t ncon; std::memcpy(&ncon, &buf[2676], 4);
  t.emrat = f64(2688);
  uint32_t ipt[15][3] = {};
  for (int i = 0; i < 36; ++i) ipt[i / 3][i % 3] = u32(2696 + 4 * (size_t)i);
  t.numde = u32(2696 + 144);
  for (int c = 0; c < 3; ++c) ipt[12][c] = u32(2696 + 148 + 4 * (size_t)c);
  if (t.numde >= 440 && ncon > 400) {
    const size_t off = 2856 + (size_t)(ncon - 400) * 6;
    if (off + 24 > buf.size()) throw Error(OUTFIT_E_INVALID_ARGUMENT, "DE header truncated");
    for (int i = 0; i < 6; ++i) ipt[13 + i / 3][i % 3] = u32(off + 4 * (size_t)i);
  }
  static const int dim[15] = {3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2, 3, 3, 1};
  size_t words = 4;
  for (int i = 0; i < 15; ++i) words += 2 * (size_t)ipt[i][1] * ipt[i][2] * dim[i];
  const size_t recsize = words * 4;
  t.block_stride = recsize / 8;
  if (buf.size() < 2 * recsize + recsize) throw Error(OUTFIT_E_INVALID_ARGUMENT, "no data records");
  t.n_blocks = (buf.size() - 2 * recsize) / recsize;
  t.cheb.resize(t.n_blocks * t.block_stride);
  std::memcpy(t.cheb.data(), &buf[2 * recsize], t.cheb.size() * 8);
  if (std::fabs(t.cheb[0] - t.jd_start) > 1e-6 || std::fabs((t.cheb[1] - t.cheb[0]) - t.block_days) > 1e-6)
    throw Error(OUTFIT_E_INVALID_ARGUMENT, "first data record does not start at SS[0] / span SS[2] days");
  const int rows[3] = {2, 9, 10};
  for (int b = 0; b < 3; ++b) {
    t.ipt[b][0] = ipt[rows[b]][0] - 1; t.ipt[b][1] = ipt[rows[b]][1]; t.ipt[b][2] = ipt[rows[b]][2];
  }
  return t;
}
inline void load_ephemeris(Context &ctx, const DeTable &t) {
  ctx.load_ephemeris(t.cheb.data(), t.n_blocks, t.block_stride, t.jd_start, t.block_days, t.ipt, t.emrat);
}

// ---- Keplerian form of an equinoctial LSQ record ----------------------------------------------
inline double rem_euclid(double x, double m) {
  const double r = std::fmod(x, m);
  return r < 0.0 ? r + std::fabs(m) : r;
}
// (a, h, k, p, q, lambda) -> (a, e, i, Omega, omega, M)   keplerian_element.rs:185-233
inline void equinoctial_to_keplerian(const double eq[6], double kep[6]) {
  const double eps = 1.0e-12;
  const double h = eq[1], k = eq[2], p = eq[3], q = eq[4];
  const double ecc = std::sqrt(h * h + k * k);
  const double dig = ecc < eps ? 0.0 : std::atan2(h, k);
  const double t = std::sqrt(p * p + q * q);
  const double node = t < eps ? 0.0 : std::atan2(p, q);
  kep[0] = eq[0]; kep[1] = ecc; kep[2] = 2.0 * std::atan(t); kep[3] = node;
  kep[4] = rem_euclid(dig - node, kTwoPi);
  kep[5] = rem_euclid(eq[5] - dig, kTwoPi);
}
// d(a, e, i, Omega, omega, M) / d(a, h, k, p, q, lambda), column-major 6x6   equinoctial_element.rs:1049-1140
inline void jacobian_to_keplerian(const double eq[6], double jac[36]) {
  const double eps = 1.0e-12;
  const double h = eq[1], k = eq[2], p = eq[3], q = eq[4];
  const double e = std::sqrt(h * h + k * k), e_sq = e * e;
  double dvh = 0, dvk = 0, dip = 0, diq = 0, dnp = 0, dnq = 0;
  if (!(e < eps)) { dvh = k / e_sq; dvk = -h / e_sq; }
  const double t = std::sqrt(p * p + q * q), t_sq = t * t;
  if (!(t < eps)) {
    const double denom = t * (1.0 + t_sq);
    dip = 2.0 * p / denom; diq = 2.0 * q / denom; dnp = q / t_sq; dnq = -p / t_sq;
  }
  const double em = std::fmax(e, eps);
  for (int i = 0; i < 36; ++i) jac[i] = 0.0;
  auto J = [&](int r, int c) -> double & { return jac[6 * c + r]; };
  J(0, 0) = 1.0;
  J(1, 1) = h / em; J(4, 1) = dvh; J(5, 1) = -dvh;
  J(1, 2) = k / em; J(4, 2) = dvk; J(5, 2) = -dvk;
  J(2, 3) = dip; J(3, 3) = dnp; J(4, 3) = -dnp;
  J(2, 4) = diq; J(3, 4) = dnq; J(4, 4) = -dnq;
  J(5, 5) = 1.0;
}
// J C J^T, all column-major 6x6   uncertainty.rs:412-416
inline void propagate_covariance(const double jac[36], const double cov[36], double out[36]) {
  double jc[36];
  for (int c = 0; c < 6; ++c)
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
      for (int m = 0; m < 6; ++m) s += jac[6 * m + r] * cov[6 * c + m];
      jc[6 * c + r] = s;
    }
  for (int c = 0; c < 6; ++c)
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
      for (int m = 0; m < 6; ++m) s += jc[6 * m + r] * jac[6 * m + c];
      out[6 * c + r] = s;
    }
}
// LsqOrbitResult (corrected) -> Keplerian elements, covariance and 1-sigma
struct KeplerianFit { double elem[6], covariance[36], sigma[6]; };
inline KeplerianFit lsq_to_keplerian(const LsqOrbitResult &r) {
  KeplerianFit k;
  double jac[36];
  equinoctial_to_keplerian(r.elem, k.elem);
  jacobian_to_keplerian(r.elem, jac);
  propagate_covariance(jac, r.covariance, k.covariance);
  for (int j = 0; j < 6; ++j) k.sigma[j] = std::sqrt(k.covariance[7 * j]);
  return k;
}

}  // namespace io
}  // namespace outfit
