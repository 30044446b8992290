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
  // the multi-GPU group: without a device it fails loudly like a single context; the cut is pure host arithmetic
  bool group_nodev = false;
  try { Group g; group_nodev = g.size() > 0; } catch (const Error &e) { group_nodev = e.code == OUTFIT_E_NO_DEVICE; }
  const uint64_t off[6] = {0, 12, 24, 36, 48, 60};
  uint64_t cuts[3] = {9, 9, 9};
  const bool cut_ok = outfit_b200_shard_ranges(5, off, 30, 10, 2, cuts) == OUTFIT_OK && cuts[0] == 0 && cuts[2] == 5 &&
                      (cuts[1] == 2 || cuts[1] == 3);
  OutfitEphemerisConfig ec;
  outfit_b200_ephemeris_config_default(&ec);
  const bool eph_default = ec.propagator == OUTFIT_PROPAGATOR_TWOBODY && ec.aberration == OUTFIT_ABERRATION_FIRST &&
                           sizeof(OutfitIodResult) == 128;
  std::printf("params=%u threw=%d sorted=%d nodev_or_ok=%d group=%d cut=%d\n", p.max_triplets, (int)threw, (int)sorted, (int)nodev,
              (int)group_nodev, (int)cut_ok);
  return (threw && sorted && lsq_default && group_nodev && cut_ok && eph_default) ? 0 : 1;
}
