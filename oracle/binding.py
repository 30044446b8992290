"""ctypes binding of liboutfit_oracle.so (ORACLE: test infrastructure only; see oo.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)
D3 = C.c_double * 3
D9 = C.c_double * 9


def build():
    """Compile the oracle with gcc (oracle/Makefile)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    """The parity build, or -- OUTFIT_ORACLE_BUILD=o3, set by bench.py's CPU arm only -- the -O3 baseline build of
    the same sources (bit-identical results, see oracle/Makefile)."""
    global _LIB
    if _LIB is None:
        name = "liboutfit_oracle_o3.so" if os.environ.get("OUTFIT_ORACLE_BUILD") == "o3" else "liboutfit_oracle.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _declare(_LIB)
    return _LIB


class IodParams(C.Structure):
    _fields_ = [
        ("n_noise_realizations", C.c_uint64), ("noise_scale", C.c_double), ("extf", C.c_double),
        ("dtmax", C.c_double), ("dt_min", C.c_double), ("dt_max_triplet", C.c_double),
        ("optimal_interval_time", C.c_double), ("max_obs_for_triplets", C.c_uint64),
        ("max_triplets", C.c_uint32), ("gap_max", C.c_double), ("max_ecc", C.c_double),
        ("max_perihelion_au", C.c_double), ("min_rho2_au", C.c_double),
        ("aberth_max_iter", C.c_uint32), ("aberth_eps", C.c_double), ("kepler_eps", C.c_double),
        ("max_tested_solutions", C.c_uint64), ("r2_min_au", C.c_double), ("r2_max_au", C.c_double),
        ("newton_eps", C.c_double), ("newton_max_it", C.c_uint64), ("root_imag_eps", C.c_double),
    ]


class KeplerParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("r0", C.c_double), ("sig0", C.c_double), ("mu", C.c_double),
        ("alpha", C.c_double), ("e0", C.c_double), ("kind", C.c_int), ("convergency", C.c_double),
        ("has_psi_guess", C.c_int), ("psi_guess", C.c_double),
        ("max_iter_prelim_kepuni", C.c_uint64), ("parabolic_method", C.c_int),
    ]


class KeplerSolution(C.Structure):
    _fields_ = [("psi", C.c_double), ("s0", C.c_double), ("s1", C.c_double), ("s2", C.c_double),
                ("s3", C.c_double)]


class Elements(C.Structure):
    _fields_ = [("kind", C.c_int), ("epoch", C.c_double), ("e", C.c_double * 6)]


class GaussObs(C.Structure):
    _fields_ = [("idx", C.c_uint64 * 3), ("ra", D3), ("dec", D3), ("t", D3), ("obs_pos", D9)]


class GaussResult(C.Structure):
    _fields_ = [("corrected", C.c_int), ("orbit", Elements)]


class WeightedTriplet(C.Structure):
    _fields_ = [("weight", C.c_double), ("i", C.c_uint64), ("j", C.c_uint64), ("k", C.c_uint64)]


class EphemTable(C.Structure):
    _fields_ = [("cheb", c_double_p), ("n_blocks", C.c_size_t), ("block_stride", C.c_size_t),
                ("jd_start", C.c_double), ("jd_end", C.c_double), ("block_days", C.c_double),
                ("ipt", (C.c_uint32 * 3) * 3), ("emrat", C.c_double)]


class TrajView(C.Structure):
    _fields_ = [("n", C.c_size_t), ("mjd_tt", c_double_p), ("ra", c_double_p), ("dec", c_double_p),
                ("sigma_ra", c_double_p), ("sigma_dec", c_double_p), ("helio_equ", c_double_p),
                ("geo_ecl", c_double_p), ("scorer_obs_equ", c_double_p)]


class IodResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("cause", C.c_int32), ("cause_value", C.c_double),
                ("attempts", C.c_uint64), ("span", C.c_double), ("corrected", C.c_int32),
                ("element_kind", C.c_int32), ("epoch", C.c_double), ("elem", C.c_double * 6),
                ("rms", C.c_double), ("triplet_idx", C.c_uint32 * 3), ("triplet_rank", C.c_uint32),
                ("realization", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "sfunct_calls", "sfunct_terms", "newton_steps", "prelim_calls", "prelim_steps",
        "kepler_universal_solves", "brent_evals", "aberth_solves", "aberth_sweeps", "gauss_solves",
        "roots_accepted", "fg_iterations", "ecc_controls", "orbits_built", "scorer_evals",
        "scorer_newton", "earth_cheb_evals", "pvobs_evals", "propagate_universal_calls")]


IOD_RESULT_DTYPE = np.dtype([
    ("status", "<i4"), ("cause", "<i4"), ("cause_value", "<f8"), ("attempts", "<u8"),
    ("span", "<f8"), ("corrected", "<i4"), ("element_kind", "<i4"), ("epoch", "<f8"),
    ("elem", "<f8", (6,)), ("rms", "<f8"), ("triplet_idx", "<u4", (3,)), ("triplet_rank", "<u4"),
    ("realization", "<u4")], align=True)
assert IOD_RESULT_DTYPE.itemsize == C.sizeof(IodResult), (IOD_RESULT_DTYPE.itemsize, C.sizeof(IodResult))


def _declare(L):
    L.oo_s_funct.argtypes = [C.c_double, C.c_double, C.c_double * 4]
    L.oo_s_funct.restype = None
    for n in ("oo_prelim_elliptic", "oo_prelim_hyperbolic", "oo_prelim_parabolic"):
        getattr(L, n).argtypes = [C.POINTER(KeplerParams)]
        getattr(L, n).restype = C.c_double
    L.oo_prelim_kepuni.argtypes = [C.POINTER(KeplerParams), c_double_p]
    L.oo_kepler_solve.argtypes = [C.POINTER(KeplerParams), C.POINTER(KeplerSolution)]
    L.oo_kepler_params_default_solver.argtypes = [C.POINTER(KeplerParams)]
    L.oo_kepler_params_default_solver.restype = None
    L.oo_velocity_correction_with_guess.argtypes = [D3, D3, D3, C.c_double, C.c_double, C.c_double,
                                                    C.c_int, C.c_double, C.c_double, D3, c_double_p,
                                                    c_double_p, c_double_p]
    L.oo_propagate_universal.argtypes = [D3, D3, C.c_double, C.c_double, C.c_int, C.c_double,
                                         C.c_double * 11]
    L.oo_propagate_universal_batch.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
    L.oo_propagate_universal_batch.restype = None
    L.oo_rotmt.argtypes = [C.c_double, C.c_int, D9]
    L.oo_rotmt.restype = None
    L.oo_eccentricity_control.argtypes = [D3, D3, C.c_double, C.c_double, C.POINTER(C.c_int),
                                          c_double_p, c_double_p, c_double_p]
    L.oo_ccek1.argtypes = [D3, D3, C.c_double, C.POINTER(Elements)]
    L.oo_ccek1.restype = None
    L.oo_to_equinoctial.argtypes = [C.POINTER(Elements), C.POINTER(Elements)]
    L.oo_equinoctial_solve_kepler.argtypes = [C.POINTER(Elements), C.c_double, C.c_double, c_double_p]
    L.oo_propagate_twobody.argtypes = [C.POINTER(Elements), C.c_double, C.c_double, D3, D3]
    L.oo_gauss_prelim.argtypes = [C.POINTER(GaussObs), c_double_p, c_double_p, D9, D9, D3, D3]
    L.oo_coeff_eight_poly.argtypes = [C.POINTER(GaussObs), D9, D9, D3, D3, D3]
    L.oo_coeff_eight_poly.restype = None
    L.oo_aberth8.argtypes = [C.c_double * 9, C.c_uint32, C.c_double, C.c_double * 8, C.c_double * 8,
                             C.POINTER(C.c_uint32)]
    L.oo_solve_8poly.argtypes = [C.c_double * 9, C.c_uint32, C.c_double, C.c_double, C.c_double * 8,
                                 C.POINTER(C.c_int)]
    L.oo_position_vector_and_reference_epoch.argtypes = [C.POINTER(GaussObs), C.POINTER(IodParams),
                                                         D9, D9, D3, D9, c_double_p]
    L.oo_gibbs_correction.argtypes = [D9, C.c_double, C.c_double, D3]
    L.oo_gibbs_correction.restype = None
    L.oo_pos_and_vel_correction.argtypes = [C.POINTER(GaussObs), C.POINTER(IodParams), D9, D3, D9, D9,
                                            C.c_double, C.c_double, C.c_double, C.c_uint64, D9, D3,
                                            c_double_p]
    L.oo_prelim_orbit.argtypes = [C.POINTER(GaussObs), C.POINTER(IodParams), C.POINTER(GaussResult)]
    L.oo_prelim_orbit_all.argtypes = [C.POINTER(GaussObs), C.POINTER(IodParams),
                                      C.POINTER(GaussResult), C.c_int, C.POINTER(C.c_int)]
    L.oo_iod_params_default.argtypes = [C.POINTER(IodParams)]
    L.oo_iod_params_default.restype = None
    L.oo_iod_params_validate.argtypes = [C.POINTER(IodParams)]
    L.oo_downsample_uniform_with_edges.argtypes = [C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]
    L.oo_downsample_uniform_with_edges.restype = C.c_size_t
    L.oo_enumerate_triplets.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_void_p,
                                        C.c_size_t]
    L.oo_enumerate_triplets.restype = C.c_size_t
    L.oo_triplet_weight_with_inv.argtypes = [C.c_double] * 4
    L.oo_triplet_weight_with_inv.restype = C.c_double
    L.oo_best_k_triplets.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(IodParams),
                                     C.POINTER(WeightedTriplet)]
    L.oo_best_k_triplets.restype = C.c_size_t
    L.oo_earth_ephemeris.argtypes = [C.POINTER(EphemTable), C.c_double, C.c_int, D3, D3]
    L.oo_cheb_record.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_uint32, C.c_double, C.c_int, D3, D3]
    L.oo_cheb_record.restype = None
    for n in ("oo_solve_kepuni_newton", "oo_solve_kepuni_brent"):
        getattr(L, n).argtypes = [C.POINTER(KeplerParams), C.POINTER(KeplerSolution)]
    L.oo_obleq.argtypes = [C.c_double]
    L.oo_obleq.restype = C.c_double
    L.oo_nutn80.argtypes = [C.c_double, c_double_p, c_double_p]
    L.oo_nutn80.restype = None
    L.oo_rnut80.argtypes = [C.c_double, D9]
    L.oo_rnut80.restype = None
    L.oo_equequ.argtypes = [C.c_double]
    L.oo_equequ.restype = C.c_double
    L.oo_prec.argtypes = [C.c_double, D9]
    L.oo_prec.restype = None
    L.oo_gmst.argtypes = [C.c_double]
    L.oo_gmst.restype = C.c_double
    L.oo_rotpn.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, D9]
    L.oo_earth_fixed_position.argtypes = [C.c_double, C.c_double, C.c_double, D3, D3]
    L.oo_earth_fixed_position.restype = None
    L.oo_pvobs.argtypes = [C.c_double, C.c_double, D3, D3, D3, D3]
    L.oo_pvobs.restype = None
    L.oo_helio_position.argtypes = [C.POINTER(EphemTable), C.c_double, D3, D3]
    L.oo_scorer_observer_position.argtypes = [C.POINTER(EphemTable), C.c_double, D3, D3]
    L.oo_ephemeris_error.argtypes = [C.POINTER(TrajView), C.c_size_t, C.POINTER(EphemTable),
                                     C.POINTER(Elements), c_double_p]
    L.oo_compute_apparent_position.argtypes = [C.POINTER(TrajView), C.c_size_t, C.POINTER(EphemTable),
                                               C.POINTER(Elements), c_double_p, c_double_p]
    L.oo_estimate_best_orbit.argtypes = [C.POINTER(TrajView), C.POINTER(EphemTable),
                                         C.POINTER(IodParams), C.c_void_p, C.POINTER(IodResult)]
    L.oo_estimate_best_orbit.restype = None
    L.oo_fit_full_iod.argtypes = [C.c_size_t] + [C.c_void_p] * 8 + [C.POINTER(EphemTable),
                                  C.POINTER(IodParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_int]
    L.oo_fit_full_iod.restype = None
    L.oo_counters_reset.restype = None
    L.oo_draw_noise.argtypes = [C.c_uint64, C.c_size_t, C.c_void_p]
    L.oo_draw_noise.restype = None
    L.oo_splitmix64_next.argtypes = [C.POINTER(C.c_uint64)]
    L.oo_splitmix64_next.restype = C.c_uint64
    L.oo_xoshiro_next.argtypes = [C.c_uint64 * 4]
    L.oo_xoshiro_next.restype = C.c_uint64
    L.oo_xoshiro_seed_from_u64.argtypes = [C.c_uint64, C.c_uint64 * 4]
    L.oo_xoshiro_seed_from_u64.restype = None
    L.oo_ziggurat_tables.argtypes = [C.c_double * 257, C.c_double * 257]
    L.oo_ziggurat_tables.restype = None
    L.oo_ephemeris_twobody_batch.argtypes = [C.POINTER(EphemTable), C.c_size_t, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, D3,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.oo_ephemeris_twobody_batch.restype = None
    L.oo_set_aberration_order.argtypes = [C.c_int]
    L.oo_set_aberration_order.restype = None
    L.oo_ephemeris_observer_pv.argtypes = [C.POINTER(EphemTable), C.c_double, C.c_double, D3, D3, D3, D3]
    L.oo_counters_get.argtypes = [C.POINTER(Counters)]
    L.oo_counters_get.restype = None


# ---- small helpers ----------------------------------------------------------------------------
def d3(v):
    return D3(*[float(x) for x in v])


def d9(v):
    return D9(*[float(x) for x in v])


def default_iod_params(**kw):
    p = IodParams()
    lib().oo_iod_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def make_ephem_table(cheb, jd_start, block_days, ipt, emrat):
    """cheb: float64 array [n_blocks, block_stride]; ipt: 3x3 (0-based offset, n_coeff, n_sub)."""
    cheb = np.ascontiguousarray(cheb, dtype=np.float64)
    t = EphemTable()
    t.cheb = cheb.ctypes.data_as(c_double_p)
    t.n_blocks = cheb.shape[0]
    t.block_stride = cheb.shape[1]
    t.jd_start = jd_start
    t.block_days = block_days
    t.jd_end = jd_start + block_days * cheb.shape[0]
    for b in range(3):
        for j in range(3):
            t.ipt[b][j] = int(ipt[b][j])
    t.emrat = emrat
    t._keep = cheb
    return t


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def from_soa_batch(batch):
    """Re-pack a product-layout batch (plane-major SoA 3-vectors, [T,K,n,6] noise) for fit_full_iod
    below (AoS 3-vectors, flat noise + per-trajectory offsets)."""
    T = len(batch["traj_offset"]) - 1
    out = {k: batch[k] for k in ("traj_offset", "mjd_tt", "ra", "dec", "sigma_ra", "sigma_dec")}
    out["helio_equ"] = np.ascontiguousarray(batch["helio_equ"].T)
    out["geo_ecl"] = np.ascontiguousarray(batch["geo_ecl"].T)
    if batch.get("noise_z") is not None:
        nz = batch["noise_z"]
        stride = nz.shape[1] * nz.shape[2] * 6
        out["noise_z"] = np.ascontiguousarray(nz.reshape(-1))
        out["noise_offset"] = (np.arange(T + 1, dtype=np.uint64) * np.uint64(stride))
    return out


def traj_view(batch, t):
    """oo_traj_view of trajectory t of a from_soa_batch() batch (the arrays must outlive the view)."""
    o, e = int(batch["traj_offset"][t]), int(batch["traj_offset"][t + 1])
    tv = TrajView()
    tv.n = e - o
    for k in ("mjd_tt", "ra", "dec", "sigma_ra", "sigma_dec"):
        setattr(tv, k, C.cast(batch[k].ctypes.data + 8 * o, c_double_p))
    tv.helio_equ = C.cast(batch["helio_equ"].ctypes.data + 24 * o, c_double_p)
    tv.geo_ecl = C.cast(batch["geo_ecl"].ctypes.data + 24 * o, c_double_p)
    tv.scorer_obs_equ = None
    return tv


def fit_full_iod(batch, table, params, n_threads=0, dedup_earth=False):
    """batch: dict of contiguous float64/uint64 numpy arrays (see outfit_b200.synth)."""
    T = len(batch["traj_offset"]) - 1
    out = np.zeros(T, dtype=IOD_RESULT_DTYPE)
    nz = batch.get("noise_z")
    no = batch.get("noise_offset")
    lib().oo_fit_full_iod(T, ptr(batch["traj_offset"]), ptr(batch["mjd_tt"]), ptr(batch["ra"]),
                          ptr(batch["dec"]), ptr(batch["sigma_ra"]), ptr(batch["sigma_dec"]),
                          ptr(batch["helio_equ"]), ptr(batch["geo_ecl"]), C.byref(table),
                          C.byref(params), ptr(nz) if nz is not None else None,
                          ptr(no) if no is not None else None, ptr(out), n_threads,
                          1 if dedup_earth else 0)
    return out


def propagate_universal_batch(rv, t0, t1, kind=2, convergency=100 * 2.220446049250313e-16, n_threads=0):
    n = rv.shape[1]
    out = np.empty((11, n), dtype=np.float64)
    status = np.empty(n, dtype=np.int32)
    lib().oo_propagate_universal_batch(n, ptr(rv), ptr(t0), ptr(t1), kind, convergency, ptr(out),
                                       ptr(status), n_threads)
    return out, status


def ephemeris_twobody_batch(table, kind, epoch, elem, mjd_tt, mjd_ut1, body_fixed, n_threads=0,
                            dedup_observer=False, aberration_order=1):
    """kind (n,) int32, epoch (n,), elem (6, n), mjd_tt/mjd_ut1 (E,) -> out (9, E, n), status (E, n).
    aberration_order: 1 = AberrationOrder::First, 2 = Second (ephemeris/aberration.rs:60-75)."""
    n, E = kind.shape[0], mjd_tt.shape[0]
    out = np.empty((9, E, n), dtype=np.float64)
    status = np.empty((E, n), dtype=np.int32)
    lib().oo_set_aberration_order(int(aberration_order))
    try:
        lib().oo_ephemeris_twobody_batch(C.byref(table), n, ptr(kind), ptr(epoch), ptr(elem), E, ptr(mjd_tt),
                                         ptr(mjd_ut1), d3(body_fixed), ptr(out), ptr(status), n_threads,
                                         1 if dedup_observer else 0)
    finally:
        lib().oo_set_aberration_order(1)
    return out, status


class Perturber(C.Structure):
    _fields_ = [("gm", C.c_double), ("pos", C.c_double * 3)]


def planet_gm(body):
    """planet_gm.rs: 0 Sun, 1 Mercury, 2 Venus, 3 Earth-Moon, 4 Mars, 5 Jupiter, 6 Saturn, 7 Uranus, 8 Neptune, 9 Pluto, 10 Moon."""
    L = lib()
    L.oo_planet_gm.restype = C.c_double
    L.oo_planet_gm.argtypes = [C.c_int]
    return L.oo_planet_gm(int(body))


def _perturbers(gm, pos):
    arr = (Perturber * len(gm))()
    for i, (g, p) in enumerate(zip(gm, pos)):
        arr[i].gm = float(g)
        arr[i].pos[0], arr[i].pos[1], arr[i].pos[2] = (float(x) for x in p)
    return arr


def nbody_rhs(y42, gm, pos):
    L = lib()
    L.oo_nbody_rhs.argtypes = [C.c_void_p, C.POINTER(Perturber), C.c_size_t, C.c_void_p]
    L.oo_nbody_rhs.restype = None
    y = np.ascontiguousarray(y42, dtype=np.float64)
    dy = np.empty(42)
    L.oo_nbody_rhs(y.ctypes.data, _perturbers(gm, pos), len(gm), dy.ctypes.data)
    return dy


def propagate_nbody(kind, epoch, elem, t1, gm, pert_pos, atol=1e-12, rtol=1e-12):
    """EquinoctialElements::propagate_nbody for n orbits: kind (n,), epoch (n,), elem (6, n), t1 (n,), gm (P,),
    pert_pos (P, 3, n) heliocentric ecliptic J2000 at each orbit's epoch -> state (6, n) ecliptic, stm (36, n)
    column-major, status (n,), steps (n,)."""
    L = lib()
    L.oo_propagate_nbody.argtypes = [C.POINTER(Elements), C.c_double, C.POINTER(Perturber), C.c_size_t, C.c_double,
                                     C.c_double, D3, D3, C.c_double * 36, C.POINTER(C.c_uint32)]
    n = len(kind)
    state, stm = np.full((6, n), np.nan), np.full((36, n), np.nan)
    status, steps = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.uint32)
    for i in range(n):
        el = Elements()
        el.kind, el.epoch = int(kind[i]), float(epoch[i])
        for q in range(6):
            el.e[q] = float(elem[q, i])
        eq = Elements()
        rc = L.oo_to_equinoctial(C.byref(el), C.byref(eq)) if el.kind != 1 else 0
        if el.kind == 1:
            eq = el
        if rc != 0:
            status[i] = rc
            continue
        if not (np.hypot(eq.e[1], eq.e[2]) < 1.0):  # the two-body start of propagate_nbody needs e < 1
            status[i] = 10
            continue
        p, v, m, ns = D3(), D3(), (C.c_double * 36)(), C.c_uint32()
        rc = L.oo_propagate_nbody(C.byref(eq), float(t1[i]), _perturbers(gm, pert_pos[:, :, i]), len(gm), atol, rtol, p, v, m,
                                  C.byref(ns))
        status[i], steps[i] = rc, ns.value
        if rc == 0:
            state[0:3, i], state[3:6, i], stm[:, i] = list(p), list(v), list(m)
    return state, stm, status, steps


def ephemeris_nbody_batch(table, kind, epoch, elem, mjd_tt, mjd_ut1, body_fixed, gm, pert_pos, atol=1e-12, rtol=1e-12,
                          aberration_order=1):
    """PropagatorKind::NBody flavour of ephemeris_twobody_batch (one observer): out (9, E, n), status (E, n)."""
    L = lib()
    L.oo_ephemeris_nbody.argtypes = [C.POINTER(EphemTable), C.POINTER(Elements), C.c_size_t, C.c_void_p, C.c_void_p, D3,
                                     C.POINTER(Perturber), C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    L.oo_ephemeris_nbody.restype = None
    n, E = len(kind), len(mjd_tt)
    out = np.empty((9, E, n))
    status = np.empty((E, n), dtype=np.int32)
    tt = np.ascontiguousarray(mjd_tt, dtype=np.float64)
    ut = np.ascontiguousarray(mjd_ut1, dtype=np.float64)
    L.oo_set_aberration_order(int(aberration_order))
    try:
        for i in range(n):
            el = Elements()
            el.kind, el.epoch = int(kind[i]), float(epoch[i])
            for q in range(6):
                el.e[q] = float(elem[q, i])
            o = np.empty((9, E))
            st = np.empty(E, dtype=np.int32)
            L.oo_ephemeris_nbody(C.byref(table), C.byref(el), E, tt.ctypes.data, ut.ctypes.data, d3(body_fixed),
                                 _perturbers(gm, pert_pos[:, :, i]), len(gm), atol, rtol, o.ctypes.data, st.ctypes.data)
            out[:, :, i], status[:, i] = o, st
    finally:
        L.oo_set_aberration_order(1)
    return out, status


def draw_noise(seeds, per_traj):
    """Deviates of SmallRng::seed_from_u64(seed) + StandardNormal for every seed: (len(seeds), per_traj)."""
    out = np.empty((len(seeds), per_traj), dtype=np.float64)
    for t, sd in enumerate(seeds):
        lib().oo_draw_noise(int(sd), per_traj, out[t].ctypes.data)
    return out


def counters():
    c = Counters()
    lib().oo_counters_get(C.byref(c))
    return {n: getattr(c, n) for n, _ in Counters._fields_}


# ---- differential orbit correction (oo_lsq.c) ---------------------------------------------------
class LsqConfig(C.Structure):
    _fields_ = [("max_newton_iterations", C.c_uint64), ("max_outlier_rejection_passes", C.c_uint64),
                ("convergence_threshold", C.c_double), ("convergence_before_rejection_threshold", C.c_double),
                ("rms_stagnation_ratio", C.c_double), ("rms_divergence_ratio", C.c_double),
                ("max_stagnation_iterations", C.c_uint64), ("enable_outlier_rejection", C.c_int32),
                ("chi2_rejection_threshold", C.c_double), ("chi2_recovery_threshold", C.c_double),
                ("eccentricity_limit", C.c_double), ("min_semi_major_axis", C.c_double),
                ("max_semi_major_axis", C.c_double), ("min_periapsis_distance", C.c_double),
                ("max_apoapsis_distance", C.c_double), ("free_elements", C.c_int32 * 6)]


OBS_EQUATION_DTYPE = np.dtype([("d_ra", "<f8", (6,)), ("d_dec", "<f8", (6,)), ("residual_ra", "<f8"),
                               ("residual_dec", "<f8"), ("weight_ra", "<f8"), ("weight_dec", "<f8"),
                               ("weight_cross", "<f8"), ("active", "<i4")], align=True)
OBS_FIT_DTYPE = np.dtype([("sigma_ra", "<f8"), ("sigma_dec", "<f8"), ("bias_ra", "<f8"), ("bias_dec", "<f8"),
                          ("residual_ra", "<f8"), ("residual_dec", "<f8"), ("chi", "<f8"),
                          ("selection", "<i4")], align=True)
LSQ_SOLUTION_DTYPE = np.dtype([("correction", "<f8", (6,)), ("normal_matrix", "<f8", (36,)),
                               ("covariance", "<f8", (36,)), ("normalised_rms", "<f8"),
                               ("num_measurements", "<u8"), ("inversion_succeeded", "<i4")], align=True)
LSQ_RESULT_DTYPE = np.dtype([("status", "<i4"), ("kind", "<i4"), ("fallback_cause", "<i4"), ("epoch", "<f8"),
                             ("elem", "<f8", (6,)), ("sigma", "<f8", (6,)), ("normal_matrix", "<f8", (36,)),
                             ("covariance", "<f8", (36,)), ("normalised_rms", "<f8"),
                             ("total_newton_iterations", "<u8"), ("num_measurements", "<u8")], align=True)


def default_lsq_config(**kw):
    c = LsqConfig()
    lib().oo_lsq_config_default(C.byref(c))
    for k, v in kw.items():
        if k == "free_elements":
            c.free_elements = (C.c_int32 * 6)(*[int(x) for x in v])
        else:
            setattr(c, k, v)
    return c


def solve_weighted_least_squares(eqs, free=(1, 1, 1, 1, 1, 1)):
    out = np.zeros(1, dtype=LSQ_SOLUTION_DTYPE)
    fr = (C.c_int32 * 6)(*[int(x) for x in free])
    lib().oo_solve_weighted_least_squares(C.c_size_t(len(eqs)), ptr(eqs), fr, ptr(out))
    return out[0]


def update_observation_selection(fit, eqs, covariance, chi2_reject=25.0, chi2_recover=9.0):
    fit = fit.copy()
    cov = np.ascontiguousarray(covariance, dtype=np.float64)
    L = lib()
    L.oo_update_observation_selection.restype = C.c_size_t
    n = L.oo_update_observation_selection(C.c_size_t(len(fit)), ptr(fit), ptr(eqs), ptr(cov),
                                          C.c_double(chi2_reject), C.c_double(chi2_recover))
    return fit, int(n)


def fit_lsq(batch, table, cfg, iod, n_threads=0):
    """batch as for fit_full_iod (AoS geo_ecl); iod: IOD_RESULT_DTYPE array -> (results, per-obs fit data)."""
    T = len(batch["traj_offset"]) - 1
    out = np.zeros(T, dtype=LSQ_RESULT_DTYPE)
    fit = np.zeros(len(batch["mjd_tt"]), dtype=OBS_FIT_DTYPE)
    iod = np.ascontiguousarray(iod)
    lib().oo_fit_lsq(C.c_size_t(T), ptr(batch["traj_offset"]), ptr(batch["mjd_tt"]), ptr(batch["ra"]),
                     ptr(batch["dec"]), ptr(batch["sigma_ra"]), ptr(batch["sigma_dec"]), ptr(batch["geo_ecl"]),
                     C.byref(table), C.byref(cfg), ptr(iod), ptr(out), ptr(fit), n_threads)
    return out, fit


def fit_lsq_nbody(batch, table, cfg, iod, gm, pert_pos, atol=1e-12, rtol=1e-12, n_threads=0):
    """fit_lsq with PropagatorKind::NBody: gm (P,), pert_pos (P, 3, T) = the perturbers frozen at each trajectory's IOD
    epoch (heliocentric, ecliptic J2000, AU)."""
    T = len(batch["traj_offset"]) - 1
    gm = np.asarray(gm, dtype=np.float64)
    pert_pos = np.asarray(pert_pos, dtype=np.float64).reshape(len(gm), 3, T)
    pert = np.zeros((T, len(gm)), dtype=np.dtype([("gm", "<f8"), ("pos", "<f8", (3,))]))
    pert["gm"] = gm[None, :]
    pert["pos"] = np.transpose(pert_pos, (2, 0, 1))
    out = np.zeros(T, dtype=LSQ_RESULT_DTYPE)
    fit = np.zeros(len(batch["mjd_tt"]), dtype=OBS_FIT_DTYPE)
    iod = np.ascontiguousarray(iod)
    L = lib()
    L.oo_fit_lsq_nbody.restype = None
    L.oo_fit_lsq_nbody(C.c_size_t(T), ptr(batch["traj_offset"]), ptr(batch["mjd_tt"]), ptr(batch["ra"]), ptr(batch["dec"]),
                       ptr(batch["sigma_ra"]), ptr(batch["sigma_dec"]), ptr(batch["geo_ecl"]), C.byref(table), C.byref(cfg),
                       ptr(iod), ptr(pert), C.c_size_t(len(gm)), C.c_double(atol), C.c_double(rtol), ptr(out), ptr(fit),
                       C.c_int(n_threads))
    return out, fit
