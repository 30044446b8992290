"""CPU-side tests (no GPU): the oracle against the committed golden fixture, the C-ABI library
(loads, exports every symbol include/outfit_b200.h declares, parameter helpers), host logic
(synthetic generator, sharding) and the world_size-2 `gloo` path of the result gather."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "outfit_b200", "liboutfit_b200.so")):
        g.build()
    from outfit_b200.api import load_library
    return load_library()


def test_oracle_reproduces_golden_fixture(oracle):
    from outfit_b200 import synth
    g = np.load(os.path.join(GOLD, "iod_golden.npz"))
    meta = json.loads(str(g["meta"]))
    table = synth.make_ephemeris_table()
    et = oracle.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    batch = synth.make_trajectories(meta["T"], meta["n_obs"], seed=meta["seed"], table=table,
                                    max_triplets=meta["K"], n_noise=meta["nn"])
    # the generator is deterministic: same inputs as when the fixture was made
    assert np.array_equal(np.array([batch["ra"].sum(), batch["dec"].sum(), batch["mjd_tt"].sum()]), g["input_digest"])
    op = oracle.default_iod_params(n_noise_realizations=meta["nn"], max_triplets=meta["K"], noise_scale=meta["noise_scale"])
    res = oracle.fit_full_iod(oracle.from_soa_batch(batch), et, op, n_threads=2)
    assert res.tobytes() == g["results"].tobytes()
    # the de-duplicated Earth evaluation (not the reference's behaviour) gives the same bits
    res2 = oracle.fit_full_iod(oracle.from_soa_batch(batch), et, op, n_threads=2, dedup_earth=True)
    assert res2.tobytes() == res.tobytes()


def test_c_abi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "outfit_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(outfit_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    from outfit_b200.api import ABI_SYMBOLS
    assert sorted(ABI_SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.outfit_b200_abi_version() == 6  # v4: + OutfitGroup (multi-GPU), fit_iod, ephemeris_request; v5: + propagate_nbody; v6: + fit_lsq_nbody


def test_struct_layouts_match_the_oracle_and_numpy_views(lib, oracle):
    from outfit_b200 import api
    assert C.sizeof(api.IODParams) == C.sizeof(oracle.IodParams) == 22 * 8
    for name, _ in oracle.IodParams._fields_:
        assert getattr(api.IODParams, name).offset == getattr(oracle.IodParams, name).offset, name
    assert C.sizeof(api.IodResult) == C.sizeof(oracle.IodResult) == api.RESULT_DTYPE.itemsize == 128
    for name, _ in api.IodResult._fields_:
        assert getattr(api.IodResult, name).offset == api.RESULT_DTYPE.fields[name][1]
    assert C.sizeof(api.ObsBatch) == 15 * 8  # ABI v2: + traj_seed
    # ABI v3: FitLSQ records -- same field offsets as the oracle's (which has no explicit padding member)
    assert api.LSQ_RESULT_DTYPE.itemsize == oracle.LSQ_RESULT_DTYPE.itemsize == 720
    for name in oracle.LSQ_RESULT_DTYPE.names:
        assert api.LSQ_RESULT_DTYPE.fields[name][1] == oracle.LSQ_RESULT_DTYPE.fields[name][1], name
    assert C.sizeof(api.DifferentialCorrectionConfig) == C.sizeof(oracle.LsqConfig) == 144
    for name, _ in oracle.LsqConfig._fields_:
        assert getattr(api.DifferentialCorrectionConfig, name).offset == getattr(oracle.LsqConfig, name).offset, name
    c = api.DifferentialCorrectionConfig.default(max_newton_iterations=7, free_elements=(1, 0, 1, 1, 1, 1))
    o = oracle.default_lsq_config(max_newton_iterations=7, free_elements=(1, 0, 1, 1, 1, 1))
    assert bytes(c)[:56] == bytes(o)[:56] and list(c.free_elements) == list(o.free_elements)
    assert (c.chi2_rejection_threshold, c.eccentricity_limit, c.max_apoapsis_distance) == (25.0, 1.2, 1e4)


def test_iod_params_default_and_validation(lib, oracle):
    """IODParams::default() and IODParamsBuilder::build() (mod.rs:308-344, 544-624)."""
    from outfit_b200 import IODParams, OutfitError
    p = IODParams.builder()
    o = oracle.default_iod_params()
    for name, _ in oracle.IodParams._fields_:
        assert getattr(p, name) == getattr(o, name), name
    assert (p.n_noise_realizations, p.max_triplets, p.max_tested_solutions, p.newton_max_it) == (20, 10, 3, 50)
    assert p.kepler_eps == 1e3 * 2.220446049250313e-16 and p.gap_max == 8.0 / 24.0
    bad = [dict(noise_scale=-1.0), dict(dt_min=-0.1), dict(max_ecc=float("nan")), dict(min_rho2_au=0.0),
           dict(aberth_eps=0.0), dict(newton_max_it=0), dict(aberth_max_iter=0), dict(max_tested_solutions=0),
           dict(r2_min_au=10.0, r2_max_au=1.0), dict(root_imag_eps=-1e-9), dict(kepler_eps=float("nan"))]
    for kw in bad:
        with pytest.raises(OutfitError) as e:
            IODParams.builder(**kw)
        assert e.value.code == -5
        assert oracle.lib().oo_iod_params_validate(C.byref(oracle.default_iod_params(**kw))) == 16
    IODParams.builder(noise_scale=0.0, max_ecc=0.0, root_imag_eps=0.0, r2_min_au=1.0, r2_max_au=1.0)


def test_no_device_is_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from outfit_b200 import OutfitB200, OutfitError
    with pytest.raises(OutfitError) as e:
        OutfitB200(0)
    assert e.value.code == -2
    assert b"no CPU fallback" in lib.outfit_b200_strerror(-2)


def test_product_never_references_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "outfit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(dirpath, f)
    so = os.path.join(ROOT, "outfit_b200", "liboutfit_b200.so")
    if os.path.exists(so):
        out = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
        assert "oracle" not in out


def test_synth_generator_is_deterministic_and_sorted():
    from outfit_b200 import synth
    table = synth.make_ephemeris_table()
    a = synth.make_trajectories(50, (3, 30), seed=5, table=table, max_triplets=4, n_noise=2)
    b = synth.make_trajectories(50, (3, 30), seed=5, table=table, max_triplets=4, n_noise=2)
    for k in ("mjd_tt", "ra", "dec", "helio_equ", "geo_ecl", "noise_z"):
        assert np.array_equal(a[k], b[k])
    off = a["traj_offset"].astype(np.int64)
    for t in range(50):
        tt = a["mjd_tt"][off[t]:off[t + 1]]
        assert (np.diff(tt) > 0).all()
    assert a["helio_equ"].shape == (3, off[-1]) and a["noise_z"].shape == (50, 4, 2, 6)
    # the table reproduces its analytic model to ~1e-11 AU
    mjd = np.linspace(58010.0, 61990.0, 500)
    emb, moon, sun = synth._model_bodies(2400000.5 + mjd)
    truth = ((emb - moon / (1 + synth.EMRAT)) - sun) / synth.AU_KM
    assert np.abs(synth.earth_position_np(table, mjd) - truth).max() < 1e-10


def test_earth_ephemeris_oracle_vs_numpy(oracle):
    from outfit_b200 import synth
    table = synth.make_ephemeris_table()
    et = oracle.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    mjd = np.random.default_rng(3).uniform(58001.0, 61999.0, 300)
    ref = synth.earth_position_np(table, mjd)
    pos, vel = oracle.D3(), oracle.D3()
    for i, t in enumerate(mjd):
        assert oracle.lib().oo_earth_ephemeris(C.byref(et), float(t), 1, pos, vel) == 0
        assert np.abs(np.array(list(pos)) - ref[:, i]).max() < 5e-12
        assert 0.015 < np.linalg.norm(list(vel)) < 0.0185  # AU/day
    assert oracle.lib().oo_earth_ephemeris(C.byref(et), 40000.0, 0, pos, vel) == 17  # outside the table


def test_shard_ranges_cover_and_balance():
    from outfit_b200 import shard
    rng = np.random.default_rng(0)
    counts = rng.integers(8, 31, 10_000)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    for ws in (1, 2, 4, 8):
        r = shard.shard_ranges(off, ws, max_triplets=30, n_noise=10)
        assert r[0][0] == 0 and r[-1][1] == 10_000 and all(r[i][1] == r[i + 1][0] for i in range(ws - 1))
        cost = shard.work_estimate(counts, 30, 10)
        per = np.array([cost[b:e].sum() for b, e in r])
        assert per.max() / per.mean() < 1.02
    assert shard.shard_ranges(np.array([0, 5, 9], dtype=np.uint64), 4) [-1][1] == 2
    assert shard.shard_ranges(np.zeros(1, dtype=np.uint64), 2) == [(0, 0), (0, 0)]


def test_o3_baseline_build_of_the_oracle_is_bit_identical_to_the_parity_build(oracle):
    """bench.py's CPU arm times oracle/liboutfit_oracle_o3.so (-O3 -march=x86-64-v3, no FMA contraction): same
    sources, same bits as the -O2 parity build, with and without the de-duplicated Earth evaluation."""
    code = (
        "import sys, hashlib\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from oracle import binding as O\n"
        "from outfit_b200 import synth\n"
        "t = synth.make_ephemeris_table()\n"
        "et = O.make_ephem_table(t['cheb'], t['jd_start'], t['block_days'], t['ipt'], t['emrat'])\n"
        "b = synth.make_trajectories(60, (8, 20), seed=3, table=t, max_triplets=10, n_noise=2)\n"
        "p = O.default_iod_params(n_noise_realizations=2, max_triplets=10, noise_scale=1.1)\n"
        "for dd in (False, True):\n"
        "    r = O.fit_full_iod(O.from_soa_batch(b), et, p, n_threads=2, dedup_earth=dd)\n"
        "    print(hashlib.md5(r.tobytes()).hexdigest())\n")
    outs = []
    for build in ("", "o3"):
        env = dict(os.environ, OUTFIT_ORACLE_BUILD=build)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append(r.stdout.split())
    assert len(outs[0]) == 2 and len(set(outs[0] + outs[1])) == 1, outs


def test_library_cut_equals_the_python_cut(lib):
    """outfit_b200_shard_ranges (the cut OutfitGroup uses inside the C-ABI; pure host arithmetic, no device
    needed) == outfit_b200/shard.py on ragged, uniform, tiny and empty batches."""
    from outfit_b200 import shard, shard_ranges
    rng = np.random.default_rng(3)
    for T, lo, hi, K, nn in ((10_000, 8, 31, 30, 10), (1000, 12, 13, 30, 10), (7, 0, 6, 10, 20), (3, 3, 4, 10, 0),
                             (50_000, 3, 60, 10, 20)):
        off = np.concatenate([[0], np.cumsum(rng.integers(lo, hi, T))]).astype(np.uint64)
        for parts in (1, 2, 3, 4, 8):
            assert shard_ranges(off, parts, K, nn) == shard.shard_ranges(off, parts, K, nn), (T, parts)
    assert shard_ranges(np.zeros(1, dtype=np.uint64), 2) == [(0, 0), (0, 0)]


def test_slice_batch_carries_seeds_and_longest_trajectory():
    from outfit_b200 import shard
    off = np.array([0, 3, 10, 14, 30], dtype=np.uint64)
    n = 30
    batch = {"traj_offset": off, "mjd_tt": np.arange(n, dtype=float), "ra": np.zeros(n), "dec": np.zeros(n),
             "sigma_ra": np.ones(n), "sigma_dec": np.ones(n), "helio_equ": np.zeros((3, n)), "geo_ecl": np.zeros((3, n)),
             "noise_z": None, "traj_seed": np.array([11, 22, 33, 44], dtype=np.uint64), "max_obs_per_traj": 16}
    s = shard.slice_batch(batch, 1, 3)
    assert list(s["traj_seed"]) == [22, 33] and s["max_obs_per_traj"] == 7
    assert list(s["traj_offset"]) == [0, 7, 11] and s["mjd_tt"][0] == 3.0


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from oracle import binding as O
from outfit_b200 import shard, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
table = synth.make_ephemeris_table()
et = O.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
batch = synth.make_trajectories(41, (5, 14), seed=77, table=table, max_triplets=6, n_noise=2)
op = O.default_iod_params(n_noise_realizations=2, max_triplets=6)
ranges = shard.shard_ranges(batch["traj_offset"], world, 6, 2)
b, e = ranges[rank]
local = O.fit_full_iod(O.from_soa_batch(shard.slice_batch(batch, b, e)), et, op, n_threads=1)
allr = shard.gather_results(local, ranges, rank, world, dist=dist)
# the differential correction shards the same way (its records are per trajectory as well)
cfg = O.default_lsq_config()
lloc, lfit = O.fit_lsq(O.from_soa_batch(shard.slice_batch(batch, b, e)), et, cfg, local, n_threads=1)
lall = shard.gather_results(lloc, ranges, rank, world, dist=dist)
if rank == 0:
    whole = O.fit_full_iod(O.from_soa_batch(batch), et, op, n_threads=1)
    assert allr.tobytes() == whole.tobytes(), "sharded + gathered != single process"
    lwhole, _ = O.fit_lsq(O.from_soa_batch(batch), et, cfg, whole, n_threads=1)
    assert lall.tobytes() == lwhole.tobytes(), "sharded + gathered LSQ != single process"
    print("GATHER_OK", len(allr), "LSQ_OK", int((lall["kind"] == 1).sum()))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_sharding_and_gather_gloo(tmp_path, oracle):
    """N > 1 host path on CPU: contiguous trajectory shards, per-rank fit, one gather; the result
    equals the single-process run bit for bit (the per-trajectory noise travels with its shard)."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29653", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK 41 LSQ_OK" in out.stdout


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "small"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "full_iod_trajectories_per_s"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["value"] > 0


def test_cpp_host_header_compiles(tmp_path):
    """outfit_b200/host/outfit_b200.hpp (the C++17 mirror of the reference interface) builds against the
    C-ABI with g++ alone and its host-side logic works without a GPU: builder validation
    (mod.rs:544-624), total_cmp time sort (obs_dataset_api.rs:222-223), SoA flattening, and the
    no-device error (no CPU fallback)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_smoke")
    libdir = os.path.join(root, "outfit_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", os.path.join(root, "tests", "cpp", "host_header_smoke.cpp"),
                           "-o", exe, "-L" + libdir, "-loutfit_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "threw=1 sorted=1" in out.stdout
