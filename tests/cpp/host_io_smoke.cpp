// Prints the C++ host-side I/O results for tests/test_cpp_io.py to compare with the Python modules.
#include "../../outfit_b200/host/outfit_b200_io.hpp"
#include <cstdio>
#include <fstream>
#include <iterator>
int main(int argc, char **argv) {
  using namespace outfit;
  using namespace outfit::io;
  if (argc < 3) return 2;
  std::ifstream f(argv[1]);
  const std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  auto traj = parse_mpc80(text, true);
  std::printf("ntraj %zu id %s n %zu\n", traj.size(), traj[0].first.c_str(), traj[0].second.size());
  std::ifstream g(argv[2]);
  const std::string eop((std::istreambuf_iterator<char>(g)), std::istreambuf_iterator<char>());
  Ut1Table ut1 = Ut1Table::from_eop2_text(eop);
  // the site table from the MPC observatory list's fixed columns (same rule as mpc80.parse_obscodes)
  std::map<std::string, Site> sites = parse_obscodes(
      "Code  Long.   cos      sin    Name\n"
      "500   0.0000 0.00000  0.00000  Geocentric\n"
      "G96 249.211280.845111+0.533614Mt. Lemmon Survey\n"
      "C51                           WISE\n"
      "F51 203.744090.936241+0.351543Pan-STARRS 1, Haleakala\n");
  std::printf("sites %zu\n", sites.size());
  auto obs = to_observations(traj[0].second, sites, 0.5, 0.0, &ut1);
  for (const Observation &o : obs)
    std::printf("obs %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", o.mjd_tt, o.ra, o.dec, o.sigma_ra, o.body_fixed[0],
                o.body_fixed[1], o.body_fixed[2], o.mjd_ut1);
  ObsBatchBuilder b;
  b.add_trajectory(obs);
  OutfitObsBatch ob = b.finish();
  std::printf("batch %llu %llu sorted %d\n", (unsigned long long)ob.n_traj, (unsigned long long)ob.n_obs,
              (int)(ob.mjd_tt[0] <= ob.mjd_tt[ob.n_obs - 1]));
  LsqOrbitResult r{};
  const double eq[6] = {1.8017360713, 0.2693736809404963, 0.08856415260522467, 0.0008089970142830734, 0.10168201110394352, 1.693697008};
  for (int j = 0; j < 6; ++j) r.elem[j] = eq[j];
  for (int i = 0; i < 36; ++i) r.covariance[i] = (i % 7 == 0) ? 1e-8 * (1 + i / 7) : 1e-10 * ((i * 7) % 5);
  for (int c = 0; c < 6; ++c)
    for (int rr = 0; rr < c; ++rr) r.covariance[6 * c + rr] = r.covariance[6 * rr + c];
  KeplerianFit k = lsq_to_keplerian(r);
  std::printf("kep");
  for (int j = 0; j < 6; ++j) std::printf(" %.17g", k.elem[j]);
  std::printf("\ncov");
  for (int i = 0; i < 36; ++i) std::printf(" %.17g", k.covariance[i]);
  std::printf("\nsig");
  for (int j = 0; j < 6; ++j) std::printf(" %.17g", k.sigma[j]);
  std::printf("\n");
  if (argc > 3) {
    DeTable t = read_de_binary(argv[3]);
    double sum = 0.0;
    for (double v : t.cheb) sum += v;
    std::printf("de %zu %zu %.17g %.17g %.17g %.17g %u", t.n_blocks, t.block_stride, t.jd_start, t.jd_end, t.block_days, t.emrat, t.numde);
    for (int b = 0; b < 3; ++b) std::printf(" %u %u %u", t.ipt[b][0], t.ipt[b][1], t.ipt[b][2]);
    std::printf(" %.17g\n", sum);
  }
  return 0;
}
