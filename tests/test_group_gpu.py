"""GPU tests of the multi-GPU group inside the C-ABI (the replacement of `fit_full_iod_parallel`,
obs_dataset_api.rs:175-207), of the single-trajectory entry `fit_iod` (:118-143), of the multi-observer
ephemeris request (ephemeris/request.rs:276-340) and of the pipelined bulk host entries.

On a single-GPU box the group is built from several contexts on device 0: the sharding, the per-shard host
threads, the strided copies of a trajectory range and the placement of the records at their global index are
the same code whatever the devices are."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import OutfitB200, OutfitGroup, synth
    table = synth.make_ephemeris_table()
    ctx = OutfitB200(0)
    ctx.load_ephemeris(table)
    n_dev = torch.cuda.device_count()
    groups = {}
    for n in (2, 3):
        g = OutfitGroup(list(range(n)) if n_dev >= n else [0] * n)
        g.load_ephemeris(table)
        groups[n] = g
    et = oracle.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    return dict(ctx=ctx, groups=groups, table=table, synth=synth, O=oracle, et=et, n_dev=n_dev)


def test_group_full_iod_is_bit_identical_to_one_context(env):
    """Ragged 8-30 observations, host-drawn deviates: 1 context == 2 == 3 contexts, byte for byte, and the cut
    is the work-balanced one."""
    from outfit_b200 import IODParams, shard_ranges
    batch = env["synth"].make_trajectories(6000, (8, 30), seed=501, table=env["table"], max_triplets=10, n_noise=3)
    params = IODParams.builder(n_noise_realizations=3, max_triplets=10, noise_scale=1.1)
    one = env["ctx"].fit_full_iod(batch, params)
    assert (one["status"] == 0).mean() > 0.8
    for n, g in env["groups"].items():
        got = g.fit_full_iod(batch, params)
        assert got.tobytes() == one.tobytes(), f"{n} shards differ from the single context"
        cuts, ms = g.last_shards()
        assert [(int(cuts[i]), int(cuts[i + 1])) for i in range(n)] == shard_ranges(batch["traj_offset"], n, 10, 3)
        assert (ms > 0).all()


def test_group_body_fixed_and_seeded_paths(env):
    """On-device observer geometry (strided body-fixed planes of a range) and on-device deviates (per-trajectory
    seeds travel with their shard)."""
    from outfit_b200 import IODParams
    batch = env["synth"].make_trajectories(3000, 12, seed=502, table=env["table"], max_triplets=10, n_noise=2)
    params = IODParams.builder(n_noise_realizations=2, max_triplets=10)
    seeded = dict(batch)
    seeded["noise_z"] = None
    seeded["traj_seed"] = (np.arange(3000, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(17))
    one = env["ctx"].fit_full_iod(seeded, params, use_body_fixed=True)
    got = env["groups"][3].fit_full_iod(seeded, params, use_body_fixed=True)
    assert got.tobytes() == one.tobytes()


def test_group_with_fewer_trajectories_than_gpus_and_empty_batch(env):
    from outfit_b200 import IODParams
    params = IODParams.builder(n_noise_realizations=0, max_triplets=10)
    batch = env["synth"].make_trajectories(2, 12, seed=503, table=env["table"], max_triplets=10, n_noise=1)
    one = env["ctx"].fit_full_iod(batch, params)
    assert env["groups"][3].fit_full_iod(batch, params).tobytes() == one.tobytes()
    empty = {k: (v[:0] if isinstance(v, np.ndarray) and v.ndim == 1 else v) for k, v in batch.items()}
    empty["traj_offset"] = np.zeros(1, dtype=np.uint64)
    for k in ("helio_equ", "geo_ecl", "body_fixed"):
        empty[k] = np.zeros((3, 0))
    assert len(env["groups"][2].fit_full_iod(empty, params)) == 0


def test_group_lsq_propagation_and_ephemeris(env):
    from outfit_b200 import DifferentialCorrectionConfig, IODParams, SolverType
    synth, ctx, g = env["synth"], env["ctx"], env["groups"][3]
    batch = synth.make_trajectories(2000, (8, 20), seed=504, table=env["table"], max_triplets=10, n_noise=1)
    params = IODParams.builder(n_noise_realizations=1, max_triplets=10)
    iod = ctx.fit_full_iod(batch, params)
    cfg = DifferentialCorrectionConfig.default()
    l1, f1 = ctx.fit_lsq(batch, params, cfg, initial_orbits=iod)
    l3, f3 = g.fit_lsq(batch, params, cfg, initial_orbits=iod)
    assert l3.tobytes() == l1.tobytes() and f3.tobytes() == f1.tobytes()
    l0, f0 = g.fit_lsq(batch, params, cfg)  # initial_orbits = None: every shard runs its IOD first
    assert l0.tobytes() == l1.tobytes() and f0.tobytes() == f1.tobytes()
    # bulk propagation: columns of plane-major arrays, cuts aligned to 128
    rv, t0, t1 = synth.make_propagation_states(50_001, seed=5)
    o1, s1 = ctx.propagate_universal(rv, t0, t1, SolverType(kind=2))
    o3, s3 = g.propagate_universal(rv, t0, t1, SolverType(kind=2))
    assert o3.tobytes() == o1.tobytes() and s3.tobytes() == s1.tobytes()
    # ephemeris request with two observers == the two single-observer calls, stacked in request order
    kind, epoch0, elem = synth.make_ephemeris_orbits(3000, seed=6)
    tt, ut1, bf = synth.make_ephemeris_epochs(7)
    bf2 = np.array([bf[1], bf[2], bf[0]]) * 0.9
    tt2, ut2 = tt[:4] + 0.25, ut1[:4] + 0.25
    a, sa = ctx.ephemeris_twobody(kind, epoch0, elem, tt, ut1, bf)
    b, sb = ctx.ephemeris_twobody(kind, epoch0, elem, tt2, ut2, bf2)
    req = [(bf, tt, ut1), (bf2, tt2, ut2)]
    r1, rs1 = ctx.ephemeris_request(kind, epoch0, elem, req)
    assert r1.tobytes() == np.concatenate([a, b], axis=1).tobytes() and rs1.tobytes() == np.concatenate([sa, sb]).tobytes()
    r3, rs3 = g.ephemeris_request(kind, epoch0, elem, req)
    assert r3.tobytes() == r1.tobytes() and rs3.tobytes() == rs1.tobytes()


def test_fit_iod_single_trajectory_equals_its_record_in_the_batch(env):
    from outfit_b200 import IODParams
    batch = env["synth"].make_trajectories(50, (8, 30), seed=505, table=env["table"], max_triplets=30, n_noise=10)
    params = IODParams.builder(n_noise_realizations=10, max_triplets=30, noise_scale=1.1)
    full = env["ctx"].fit_full_iod(batch, params)
    for t in (0, 17, 49):
        assert env["ctx"].fit_iod(batch, params, t).tobytes() == full[t].tobytes()
    with pytest.raises(Exception):
        env["ctx"].fit_iod(batch, params, 50)


def test_pipelined_host_entries_equal_the_device_entries(env):
    """Chunked H2D / kernel / D2H rings (several chunks, a ragged last one) give the bytes of one device launch."""
    import os
    import torch
    from outfit_b200 import SolverType
    synth, ctx = env["synth"], env["ctx"]
    n = 70_003
    rv, t0, t1 = synth.make_propagation_states(n, seed=9)
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(x).to(dev) for x in (rv, t0, t1)]
    d_o = torch.empty(11 * n, dtype=torch.float64, device=dev)
    d_s = torch.empty(n, dtype=torch.int32, device=dev)
    ctx.propagate_universal_device(n, d[0], d[1], d[2], d_o, d_s, SolverType(kind=2))
    torch.cuda.synchronize()
    os.environ["OUTFIT_B200_PROP_CHUNK"] = "8192"
    os.environ["OUTFIT_B200_EPH_CHUNK"] = "512"
    try:
        o, s = ctx.propagate_universal(rv, t0, t1, SolverType(kind=2))
        assert o.tobytes() == d_o.cpu().numpy().tobytes() and s.tobytes() == d_s.cpu().numpy().tobytes()
        kind, epoch0, elem = synth.make_ephemeris_orbits(2500, seed=10)
        tt, ut1, bf = synth.make_ephemeris_epochs(9)
        dk, de, dl, dt, du = (torch.from_numpy(x).to(dev) for x in (kind, epoch0, elem, tt, ut1))
        d_eo = torch.empty(9 * 9 * 2500, dtype=torch.float64, device=dev)
        d_es = torch.empty(9 * 2500, dtype=torch.int32, device=dev)
        ctx.ephemeris_twobody_device(2500, dk, de, dl, 9, dt, du, bf, d_eo, d_es)
        torch.cuda.synchronize()
        ho, hs = ctx.ephemeris_twobody(kind, epoch0, elem, tt, ut1, bf)
        assert ho.tobytes() == d_eo.cpu().numpy().tobytes() and hs.tobytes() == d_es.cpu().numpy().tobytes()
    finally:
        del os.environ["OUTFIT_B200_PROP_CHUNK"], os.environ["OUTFIT_B200_EPH_CHUNK"]


def test_epoch_outside_the_ephemeris_is_a_per_trajectory_error(env):
    """The reference panics ("Time outside ephemeris range", horizon_data.rs:722); here the trajectory carries
    OUTFIT_ST_EPHEM_OUT_OF_RANGE (17) and every other trajectory is untouched -- IOD and LSQ."""
    from outfit_b200 import DifferentialCorrectionConfig, IODParams
    batch = env["synth"].make_trajectories(300, 12, seed=506, table=env["table"], max_triplets=10, n_noise=1)
    params = IODParams.builder(n_noise_realizations=0, max_triplets=10)
    good = env["ctx"].fit_full_iod(batch, params, use_body_fixed=True)
    bad = dict(batch)
    bad["mjd_tt"] = batch["mjd_tt"].copy()
    off = batch["traj_offset"].astype(np.int64)
    # trajectory 5: its LAST epoch far outside the table (stays time-sorted)
    bad["mjd_tt"][off[6] - 1] = 99000.0
    for use_bf in (True, False):
        got = env["ctx"].fit_full_iod(bad, params, use_body_fixed=use_bf)
        assert got["status"][5] == 17
        keep = np.arange(300) != 5
        ref = good if use_bf else env["ctx"].fit_full_iod(batch, params)
        assert got[keep].tobytes() == ref[keep].tobytes()
    lres, _ = env["ctx"].fit_lsq(bad, params, DifferentialCorrectionConfig.default(), initial_orbits=good)
    assert lres["status"][5] == 17 and lres["kind"][5] == 0


def test_zero_max_triplets_reports_no_feasible_triplets(env):
    from outfit_b200 import IODParams
    batch = env["synth"].make_trajectories(40, 12, seed=507, table=env["table"], max_triplets=10, n_noise=1)
    got = env["ctx"].fit_full_iod(batch, IODParams.builder(n_noise_realizations=0, max_triplets=0))
    assert (got["status"] == 13).all()
    want = env["O"].fit_full_iod(env["O"].from_soa_batch(batch), env["et"],
                                 env["O"].default_iod_params(n_noise_realizations=0, max_triplets=0), n_threads=0)
    assert (want["status"] == 13).all() and np.array_equal(got["span"], want["span"])


def test_thread_triplet_kernel_serves_larger_k(env):
    """max_triplets 150 and 400 take the thread-per-trajectory kernel with 64 / 32-thread blocks (it used to fall
    back to the ~9x slower warp kernel above K = 106): same records as the oracle's."""
    from outfit_b200 import IODParams
    O = env["O"]
    batch = env["synth"].make_trajectories(200, 14, seed=508, table=env["table"], max_triplets=10, n_noise=1)
    for K in (150, 400):
        got = env["ctx"].fit_full_iod(batch, IODParams.builder(n_noise_realizations=0, max_triplets=K))
        want = O.fit_full_iod(O.from_soa_batch(batch), env["et"], O.default_iod_params(n_noise_realizations=0, max_triplets=K),
                              n_threads=0)
        for f in ("status", "triplet_idx", "triplet_rank", "attempts"):
            assert np.array_equal(got[f], want[f]), (K, f)
