#include "../../outfit_b200/host/outfit_b200.hpp"
#include <cstdio>
int main() {
  using namespace outfit;
  OutfitIodParams p = IODParamsBuilder().n_noise_realizations(0).max_triplets(10).build();
  bool threw = false;
  try { IODParamsBuilder().aberth_eps(-1.0).build(); } catch (const Error &e) { threw = e.code == OUTFIT_E_INVALID_IOD_PARAMETER; }
  ObsBatchBuilder b;
  std::vector<Observation> tr(3);
  for (int i = 0; i < 3; ++i) { tr[i] = Observation{60000.0 + (2 - i), 1.0, 0.5, 1e-6, 1e-6, {0, 0, 0}, 60000.0 + (2 - i)}; }
  b.add_trajectory(tr);
  OutfitObsBatch ob = b.finish();
  bool sorted = ob.mjd_tt[0] < ob.mjd_tt[1] && ob.mjd_tt[1] < ob.mjd_tt[2] && ob.n_traj == 1 && ob.n_obs == 3 && ob.max_obs_per_traj == 3;
  OutfitLsqConfig lc = Context::default_lsq_config();
  bool lsq_default = lc.max_newton_iterations == 30 && lc.max_outlier_rejection_passes == 10 && lc.free_elements[5] == 1 &&
                     lc.chi2_rejection_threshold == 25.0 && sizeof(OutfitLsqResult) == 8 * 90 && sizeof(OutfitObsFit) == 32;
  bool nodev = false;
  try { Context c(0); } catch (const Error &e) { nodev = e.code == OUTFIT_E_NO_DEVICE; }
  std::printf("params=%u threw=%d sorted=%d nodev_or_ok=%d\n", p.max_triplets, (int)threw, (int)sorted, (int)nodev);
  return (threw && sorted && lsq_default) ? 0 : 1;
}
