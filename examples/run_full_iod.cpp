// The reference's quick start (README / examples/run_full_iod.rs) in C++ over the C-ABI:
//   MPC 80-column file + JPL DE binary file [+ JPL latest_eop2.long] -> Gauss IOD -> differential correction
//
//   g++ -std=c++17 examples/run_full_iod.cpp -o run_full_iod -Loutfit_b200 -loutfit_b200 -Wl,-rpath,$PWD/outfit_b200
//   ./run_full_iod FILE.obs DE_FILE [EOP2_FILE] [--dry-run]
//
// Observatory parallax constants are the ones of outfit_b200/mpc80.py (the codes of the reference's
// tests/data/2015AB.obs).  Needs a CUDA device: there is no CPU fallback (--dry-run stops before the GPU).
#include <cstdio>
#include <fstream>
#include <iterator>

#include "../outfit_b200/host/outfit_b200_io.hpp"

int main(int argc, char **argv) {
  using namespace outfit;
  using namespace outfit::io;
  std::vector<std::string> args;
  bool dry = false;
  for (int i = 1; i < argc; ++i) {
    if (std::string(argv[i]) == "--dry-run") dry = true;
    else args.push_back(argv[i]);
  }
  if (args.size() < 2) {
    std::fprintf(stderr, "usage: %s FILE.obs DE_FILE [EOP2_FILE] [--dry-run]\n", argv[0]);
    return 2;
  }
  try {
    std::ifstream f(args[0]);
    const std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto traj = parse_mpc80(text, /*single_trajectory=*/true);
    if (traj.empty()) throw Error(OUTFIT_E_INVALID_ARGUMENT, "no optical records in " + args[0]);
    const DeTable de = read_de_binary(args[1]);
    Ut1Table ut1;
    const bool have_ut1 = args.size() > 2;
    if (have_ut1) {
      std::ifstream g(args[2]);
      ut1 = Ut1Table::from_eop2_text(std::string((std::istreambuf_iterator<char>(g)), std::istreambuf_iterator<char>()));
    }
    const std::map<std::string, Site> sites = {
        {"500", {0.0, 0.0, 0.0}},                 {"204", {8.7700, 0.69740, 0.71440}},
        {"291", {248.4010, 0.84950, 0.52640}},    {"705", {254.17942, 0.841939, 0.538633}},
        {"F51", {203.74409, 0.936241, 0.351543}}, {"G96", {249.21128, 0.845111, 0.533614}}};
    ObsBatchBuilder builder;
    builder.add_trajectory(to_observations(traj[0].second, sites, 0.5, 0.0, have_ut1 ? &ut1 : nullptr));
    builder.set_seeds({42});
    OutfitObsBatch batch = builder.finish();
    std::printf("%s: %llu observations, DE%u table %zu blocks of %.0f days from JD %.1f\n", traj[0].first.c_str(),
                (unsigned long long)batch.n_obs, de.numde, de.n_blocks, de.block_days, de.jd_start);
    if (dry) return 0;

    Context ctx(0);
    load_ephemeris(ctx, de);
    const OutfitIodParams params = IODParamsBuilder().n_noise_realizations(10).noise_scale(1.1).max_triplets(30).build();
    std::vector<OutfitIodResult> raw(batch.n_traj);
    if (int rc = outfit_b200_fit_full_iod(ctx.raw(), &params, &batch, raw.data())) throw Error(rc, outfit_b200_last_error(ctx.raw()));
    const OutfitIodResult &r = raw[0];
    if (r.status != OUTFIT_ST_OK) {
      std::printf("IOD failed: status %d cause %d\n", r.status, r.cause);
      return 1;
    }
    std::printf("%sOrbit %s epoch %.6f rms %.4f\n   ", r.corrected ? "Corrected" : "Prelim",
                r.element_kind == 0 ? "Keplerian" : "Cometary", r.epoch, r.rms);
    for (double x : r.elem) std::printf(" %.10f", x);
    std::printf("\n");
    std::vector<OutfitObsFit> fit;
    auto lsq = ctx.fit_lsq(batch, params, Context::default_lsq_config(), raw.data(), &fit);
    if (lsq[0].corrected) {
      size_t rejected = 0;
      for (const OutfitObsFit &o : fit) rejected += o.selection == 1;
      const KeplerianFit k = lsq_to_keplerian(lsq[0]);
      std::printf("differential correction: %llu Newton steps, normalised rms %.4f, %zu observation(s) rejected\n   ",
                  (unsigned long long)lsq[0].total_newton_iterations, lsq[0].normalised_rms, rejected);
      for (double x : k.elem) std::printf(" %.10f", x);
      std::printf("\n   1-sigma");
      for (double x : k.sigma) std::printf(" %.3e", x);
      std::printf("\n");
    } else if (lsq[0].ok) {
      std::printf("differential correction fell back to the IOD orbit (cause %d)\n", lsq[0].fallback_cause);
    }
    return 0;
  } catch (const Error &e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}
