"""Host-side mirror of the reference's operator interface for the hot path, over the C-ABI.

Reference interface mirrored here (names, argument meaning, error behaviour):
  IODParams / IODParamsBuilder::build   src/initial_orbit_determination/mod.rs:225-344, 544-624
  FitIOD::fit_full_iod                  src/initial_orbit_determination/obs_dataset_api.rs:145-207
  kepler::propagate_universal           src/kepler/propagation.rs:114-174
  SolverType / SolverKind               src/kepler/params.rs:24-73
Per-trajectory failures are values (status codes mirroring the OutfitError variant), never
exceptions; argument / device failures raise OutfitError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATUS_NAMES = {
    0: "Ok", 1: "SingularDirectionMatrix", 2: "GaussNoRootsFound", 3: "PolynomialRootFindingFailed",
    4: "SpuriousRootDetected", 5: "VelocityCorrectionError", 6: "NewtonRaphsonKeplerConvergence",
    7: "BrentDekkerKeplerConvergence", 8: "DegenerateState", 9: "InvalidConversion", 10: "InvalidOrbit",
    11: "RootFindingError", 12: "NonFiniteScore", 13: "NoFeasibleTriplets", 14: "NoViableOrbit",
    15: "ObservationNotFound", 17: "EphemerisOutOfRange", 18: "DifferentialCorrectionFailed", 19: "BizarreOrbit",
    20: "DifferentialCorrectionDiverged", 21: "NBodyPropagationFailed",
}


class OutfitError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"outfit_b200 error {code}: {msg}")
        self.code = code


def library_path():
    # OUTFIT_B200_LIB lets a developer A/B a differently built copy of the SAME library
    return os.environ.get("OUTFIT_B200_LIB") or os.path.join(_HERE, "liboutfit_b200.so")


class IODParams(C.Structure):
    """IODParams (mod.rs:225-266); defaults = IODParams::default() (mod.rs:308-344)."""
    _fields_ = [
        ("n_noise_realizations", C.c_uint64), ("noise_scale", C.c_double), ("extf", C.c_double),
        ("dtmax", C.c_double), ("dt_min", C.c_double), ("dt_max_triplet", C.c_double),
        ("optimal_interval_time", C.c_double), ("max_obs_for_triplets", C.c_uint64),
        ("max_triplets", C.c_uint32), ("_pad0", C.c_uint32), ("gap_max", C.c_double),
        ("max_ecc", C.c_double), ("max_perihelion_au", C.c_double), ("min_rho2_au", C.c_double),
        ("aberth_max_iter", C.c_uint32), ("_pad1", C.c_uint32), ("aberth_eps", C.c_double),
        ("kepler_eps", C.c_double), ("max_tested_solutions", C.c_uint64), ("r2_min_au", C.c_double),
        ("r2_max_au", C.c_double), ("newton_eps", C.c_double), ("newton_max_it", C.c_uint64),
        ("root_imag_eps", C.c_double),
    ]

    @classmethod
    def builder(cls, **kw):
        """IODParams::builder()...build(): validated like IODParamsBuilder::build."""
        p = cls()
        load_library().outfit_b200_iod_params_default(C.byref(p))
        for k, v in kw.items():
            if k.startswith("_") or not hasattr(p, k):
                raise AttributeError(f"IODParams has no field {k!r}")
            setattr(p, k, v)
        rc = load_library().outfit_b200_iod_params_validate(C.byref(p))
        if rc != 0:
            raise OutfitError(rc, "InvalidIODParameter")
        return p


class SolverType(C.Structure):
    """SolverType{kind, params} (kepler/params.rs:59-73); kind: 0 Newton, 1 BrentDecker, 2 Auto."""
    _fields_ = [("kind", C.c_int32), ("parabolic_method", C.c_int32), ("convergency", C.c_double),
                ("max_iter_prelim_kepuni", C.c_uint64)]

    def __init__(self, kind=0, convergency=100.0 * 2.220446049250313e-16, max_iter_prelim_kepuni=20,
                 parabolic_method=0):
        super().__init__(kind, parabolic_method, convergency, max_iter_prelim_kepuni)


class ObsBatch(C.Structure):
    _fields_ = [("n_traj", C.c_uint64), ("n_obs", C.c_uint64), ("traj_offset", C.c_void_p),
                ("mjd_tt", C.c_void_p), ("ra", C.c_void_p), ("dec", C.c_void_p),
                ("sigma_ra", C.c_void_p), ("sigma_dec", C.c_void_p), ("obs_helio_equ", C.c_void_p),
                ("obs_geo_ecl", C.c_void_p), ("observer_body_fixed", C.c_void_p),
                ("mjd_ut1", C.c_void_p), ("noise_z", C.c_void_p), ("max_obs_per_traj", C.c_uint64),
                ("traj_seed", C.c_void_p)]


class IodResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("cause", C.c_int32), ("cause_value", C.c_double),
                ("attempts", C.c_uint64), ("span", C.c_double), ("corrected", C.c_int32),
                ("element_kind", C.c_int32), ("epoch", C.c_double), ("elem", C.c_double * 6),
                ("rms", C.c_double), ("triplet_idx", C.c_uint32 * 3), ("triplet_rank", C.c_uint32),
                ("realization", C.c_uint32), ("_pad0", C.c_uint32)]


class IodCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "gauss_solves", "aberth_sweeps", "roots_accepted", "fg_iterations", "kepler_universal_solves",
        "newton_steps", "sfunct_terms", "scorer_evals", "scorer_newton_steps", "candidates", "fg_iterations_skipped")]


class IodPhaseMs(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("observer_ms", "triplets_ms", "roots_ms", "correct_ms", "score_ms",
                                         "select_ms", "total_ms")] + [("n_chunks", C.c_uint32),
                                                                      ("kernel_launches", C.c_uint32)]


RESULT_DTYPE = np.dtype([
    ("status", "<i4"), ("cause", "<i4"), ("cause_value", "<f8"), ("attempts", "<u8"), ("span", "<f8"),
    ("corrected", "<i4"), ("element_kind", "<i4"), ("epoch", "<f8"), ("elem", "<f8", (6,)),
    ("rms", "<f8"), ("triplet_idx", "<u4", (3,)), ("triplet_rank", "<u4"), ("realization", "<u4"), ("_pad0", "<u4")],
    align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(IodResult)

class EphemerisConfig(C.Structure):
    """EphemerisConfig (ephemeris/mod.rs:124-142): propagator 0 TwoBody | 1 NBody (rejected here: the N-body ephemeris needs the
    perturber snapshots and has its own entry, OutfitB200.ephemeris_nbody); aberration 1 First (default) | 2 Second."""
    _fields_ = [("propagator", C.c_int32), ("aberration", C.c_int32)]

    def __init__(self, propagator=0, aberration=1):
        super().__init__(propagator, aberration)


class NBodyConfig(C.Structure):
    """NBodyConfig (propagator/mod.rs:107-150): tolerances of the DOP853 integration and the number of perturbers."""
    _fields_ = [("abs_tol", C.c_double), ("rel_tol", C.c_double), ("n_perturbers", C.c_uint32), ("max_steps", C.c_uint32)]

    def __init__(self, n_perturbers=1, abs_tol=1e-12, rel_tol=1e-12, max_steps=0):
        super().__init__(abs_tol, rel_tol, n_perturbers, max_steps)


def planet_gm(body):
    """GM in AU^3/day^2 (propagator/planet_gm.rs): 0 Sun, 1 Mercury, ... 9 Pluto, 10 Moon."""
    return float(load_library().outfit_b200_planet_gm(int(body)))


class DifferentialCorrectionConfig(C.Structure):
    """DifferentialCorrectionConfig (diff_cor.rs:100-192) + OutlierRejectionConfig + EquinoctialLimits."""
    _fields_ = [("max_newton_iterations", C.c_uint64), ("max_outlier_rejection_passes", C.c_uint64),
                ("convergence_threshold", C.c_double), ("convergence_before_rejection_threshold", C.c_double),
                ("rms_stagnation_ratio", C.c_double), ("rms_divergence_ratio", C.c_double),
                ("max_stagnation_iterations", C.c_uint64), ("enable_outlier_rejection", C.c_int32),
                ("_pad0", C.c_int32), ("chi2_rejection_threshold", C.c_double),
                ("chi2_recovery_threshold", C.c_double), ("eccentricity_limit", C.c_double),
                ("min_semi_major_axis", C.c_double), ("max_semi_major_axis", C.c_double),
                ("min_periapsis_distance", C.c_double), ("max_apoapsis_distance", C.c_double),
                ("free_elements", C.c_int32 * 6)]

    @classmethod
    def default(cls, **kw):
        c = cls()
        load_library().outfit_b200_lsq_config_default(C.byref(c))
        for k, v in kw.items():
            if k.startswith("_") or not hasattr(c, k):
                raise AttributeError(f"DifferentialCorrectionConfig has no field {k!r}")
            if k == "free_elements":
                v = (C.c_int32 * 6)(*[int(x) for x in v])
            setattr(c, k, v)
        return c


LSQ_RESULT_DTYPE = np.dtype([
    ("status", "<i4"), ("kind", "<i4"), ("fallback_cause", "<i4"), ("_pad0", "<i4"), ("epoch", "<f8"),
    ("elem", "<f8", (6,)), ("sigma", "<f8", (6,)), ("normal_matrix", "<f8", (36,)), ("covariance", "<f8", (36,)),
    ("normalised_rms", "<f8"), ("total_newton_iterations", "<u8"), ("num_measurements", "<u8")], align=True)
OBS_FIT_DTYPE = np.dtype([("residual_ra", "<f8"), ("residual_dec", "<f8"), ("chi", "<f8"), ("selection", "<i4"),
                          ("_pad0", "<i4")], align=True)
assert LSQ_RESULT_DTYPE.itemsize == 8 * (2 + 1 + 6 + 6 + 72 + 3) and OBS_FIT_DTYPE.itemsize == 32

ABI_SYMBOLS = [
    "outfit_b200_abi_version", "outfit_b200_strerror", "outfit_b200_last_error",
    "outfit_b200_iod_params_default", "outfit_b200_iod_params_validate",
    "outfit_b200_solver_type_default", "outfit_b200_init", "outfit_b200_destroy",
    "outfit_b200_load_ephemeris", "outfit_b200_fit_full_iod", "outfit_b200_fit_full_iod_device",
    "outfit_b200_observer_cache_device", "outfit_b200_propagate_universal",
    "outfit_b200_propagate_universal_device", "outfit_b200_last_iod_counters", "outfit_b200_last_iod_phase_ms",
    "outfit_b200_set_work_counters", "outfit_b200_set_pass_streams",
    "outfit_b200_measure_fp64_peak", "outfit_b200_selftest_arith", "outfit_b200_ephemeris_twobody", "outfit_b200_ephemeris_twobody_device",
    "outfit_b200_lsq_config_default", "outfit_b200_fit_lsq", "outfit_b200_fit_lsq_device",
    "outfit_b200_fit_iod", "outfit_b200_ephemeris_request", "outfit_b200_ephemeris_request_device",
    "outfit_b200_init_multi", "outfit_b200_group_destroy", "outfit_b200_group_size", "outfit_b200_group_ctx",
    "outfit_b200_group_last_error", "outfit_b200_group_load_ephemeris", "outfit_b200_group_set_pass_streams",
    "outfit_b200_group_fit_full_iod", "outfit_b200_group_fit_lsq", "outfit_b200_group_propagate_universal",
    "outfit_b200_group_ephemeris_request", "outfit_b200_group_last_shards", "outfit_b200_shard_ranges",
    "outfit_b200_host_alloc", "outfit_b200_host_free",
    "outfit_b200_ephemeris_config_default", "outfit_b200_set_ephemeris_config", "outfit_b200_group_set_ephemeris_config",
    "outfit_b200_nbody_config_default", "outfit_b200_planet_gm", "outfit_b200_propagate_nbody", "outfit_b200_propagate_nbody_device",
    "outfit_b200_ephemeris_nbody", "outfit_b200_ephemeris_nbody_device",
    "outfit_b200_fit_lsq_nbody", "outfit_b200_fit_lsq_nbody_device", "outfit_b200_group_fit_lsq_nbody",
]


def load_library():
    """dlopen liboutfit_b200.so; raises (no fallback) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise OutfitError(-100, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; "
                                "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, cp = C.c_void_p, C.c_char_p
    L.outfit_b200_abi_version.restype = C.c_int
    L.outfit_b200_strerror.restype = cp
    L.outfit_b200_strerror.argtypes = [C.c_int]
    L.outfit_b200_last_error.restype = cp
    L.outfit_b200_last_error.argtypes = [vp]
    L.outfit_b200_iod_params_default.argtypes = [C.POINTER(IODParams)]
    L.outfit_b200_iod_params_default.restype = None
    L.outfit_b200_iod_params_validate.argtypes = [C.POINTER(IODParams)]
    L.outfit_b200_solver_type_default.argtypes = [C.POINTER(SolverType)]
    L.outfit_b200_solver_type_default.restype = None
    L.outfit_b200_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.outfit_b200_destroy.argtypes = [vp]
    L.outfit_b200_destroy.restype = None
    L.outfit_b200_load_ephemeris.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_double, C.c_double, vp,
                                             C.c_double]
    L.outfit_b200_fit_full_iod.argtypes = [vp, C.POINTER(IODParams), C.POINTER(ObsBatch), vp]
    L.outfit_b200_fit_full_iod_device.argtypes = [vp, C.POINTER(IODParams), C.POINTER(ObsBatch), vp, vp]
    L.outfit_b200_observer_cache_device.argtypes = [vp, C.c_size_t, vp, vp, vp, vp, vp, vp, vp]
    L.outfit_b200_propagate_universal.argtypes = [vp, C.c_size_t, vp, vp, vp, vp, C.POINTER(SolverType), vp, vp]
    L.outfit_b200_propagate_universal_device.argtypes = [vp, C.c_size_t, vp, vp, vp, vp, C.POINTER(SolverType),
                                                         vp, vp, vp]
    L.outfit_b200_last_iod_counters.argtypes = [vp, C.POINTER(IodCounters)]
    L.outfit_b200_set_work_counters.argtypes = [vp, C.c_int]
    L.outfit_b200_set_pass_streams.argtypes = [vp, C.c_int]
    L.outfit_b200_last_iod_phase_ms.argtypes = [vp, C.POINTER(IodPhaseMs)]
    L.outfit_b200_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.outfit_b200_lsq_config_default.argtypes = [C.POINTER(DifferentialCorrectionConfig)]
    L.outfit_b200_lsq_config_default.restype = None
    L.outfit_b200_fit_lsq.argtypes = [vp, C.POINTER(IODParams), C.POINTER(DifferentialCorrectionConfig),
                                      C.POINTER(ObsBatch), vp, vp, vp]
    L.outfit_b200_fit_lsq_device.argtypes = [vp, C.POINTER(DifferentialCorrectionConfig), C.POINTER(ObsBatch), vp, vp,
                                             vp, vp]
    L.outfit_b200_selftest_arith.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64 * 4]
    L.outfit_b200_ephemeris_twobody.argtypes = [vp, C.c_size_t, vp, vp, vp, C.c_size_t, vp, vp, C.c_double * 3, vp, vp]
    L.outfit_b200_ephemeris_twobody_device.argtypes = [vp, C.c_size_t, vp, vp, vp, C.c_size_t, vp, vp,
                                                       C.c_double * 3, vp, vp, vp]
    L.outfit_b200_fit_iod.argtypes = [vp, C.POINTER(IODParams), C.POINTER(ObsBatch), C.c_uint64, vp]
    L.outfit_b200_ephemeris_request.argtypes = [vp, C.c_size_t, vp, vp, vp, C.c_size_t, vp, vp, vp, vp, vp, vp]
    L.outfit_b200_ephemeris_request_device.argtypes = [vp, C.c_size_t, vp, vp, vp, C.c_size_t, vp, vp, vp, vp, vp, vp]
    L.outfit_b200_init_multi.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.outfit_b200_group_destroy.argtypes = [vp]
    L.outfit_b200_group_destroy.restype = None
    L.outfit_b200_group_size.argtypes = [vp]
    L.outfit_b200_group_ctx.argtypes = [vp, C.c_int]
    L.outfit_b200_group_ctx.restype = vp
    L.outfit_b200_group_last_error.argtypes = [vp]
    L.outfit_b200_group_last_error.restype = cp
    L.outfit_b200_group_load_ephemeris.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_double, C.c_double, vp, C.c_double]
    L.outfit_b200_group_set_pass_streams.argtypes = [vp, C.c_int]
    L.outfit_b200_group_fit_full_iod.argtypes = [vp, C.POINTER(IODParams), C.POINTER(ObsBatch), vp]
    L.outfit_b200_group_fit_lsq.argtypes = [vp, C.POINTER(IODParams), C.POINTER(DifferentialCorrectionConfig),
                                            C.POINTER(ObsBatch), vp, vp, vp]
    L.outfit_b200_group_fit_lsq_nbody.argtypes = [vp, C.POINTER(DifferentialCorrectionConfig), C.POINTER(NBodyConfig), vp, vp,
                                                  C.POINTER(ObsBatch), vp, vp, vp]
    L.outfit_b200_group_propagate_universal.argtypes = [vp, C.c_size_t, vp, vp, vp, vp, C.POINTER(SolverType), vp, vp]
    L.outfit_b200_group_ephemeris_request.argtypes = [vp, C.c_size_t, vp, vp, vp, C.c_size_t, vp, vp, vp, vp, vp, vp]
    L.outfit_b200_group_last_shards.argtypes = [vp, vp, vp]
    L.outfit_b200_shard_ranges.argtypes = [C.c_uint64, vp, C.c_uint32, C.c_uint64, C.c_int, vp]
    L.outfit_b200_host_alloc.argtypes = [C.c_size_t]
    L.outfit_b200_host_alloc.restype = vp
    L.outfit_b200_host_free.argtypes = [vp]
    L.outfit_b200_host_free.restype = None
    L.outfit_b200_ephemeris_config_default.argtypes = [C.POINTER(EphemerisConfig)]
    L.outfit_b200_ephemeris_config_default.restype = None
    L.outfit_b200_set_ephemeris_config.argtypes = [vp, C.POINTER(EphemerisConfig)]
    L.outfit_b200_group_set_ephemeris_config.argtypes = [vp, C.POINTER(EphemerisConfig)]
    L.outfit_b200_nbody_config_default.argtypes = [C.POINTER(NBodyConfig)]
    L.outfit_b200_nbody_config_default.restype = None
    L.outfit_b200_planet_gm.argtypes = [C.c_int]
    L.outfit_b200_planet_gm.restype = C.c_double
    L.outfit_b200_propagate_nbody.argtypes = [vp, C.c_size_t, vp, vp, vp, vp, C.POINTER(NBodyConfig), vp, vp, vp, vp, vp, vp]
    L.outfit_b200_fit_lsq_nbody.argtypes = [vp, C.POINTER(DifferentialCorrectionConfig), C.POINTER(NBodyConfig), vp, vp,
                                            C.POINTER(ObsBatch), vp, vp, vp]
    L.outfit_b200_fit_lsq_nbody_device.argtypes = [vp, C.POINTER(DifferentialCorrectionConfig), C.POINTER(NBodyConfig), vp, vp,
                                                   C.POINTER(ObsBatch), vp, vp, vp, vp]
    L.outfit_b200_ephemeris_nbody.argtypes = [vp, C.c_size_t, vp, vp, vp, C.c_size_t, vp, vp, vp, vp, C.POINTER(NBodyConfig), vp, vp,
                                              vp, vp]
    L.outfit_b200_ephemeris_nbody_device.argtypes = [vp, C.c_size_t, vp, vp, vp, C.c_size_t, vp, vp, vp, C.POINTER(NBodyConfig), vp,
                                                     vp, vp, vp, vp]
    L.outfit_b200_propagate_nbody_device.argtypes = [vp, C.c_size_t, vp, vp, vp, vp, C.POINTER(NBodyConfig), vp, vp, vp, vp, vp,
                                                     vp, vp]
    _LIB = L
    return L


def shard_ranges(traj_offset, n_parts, max_triplets=10, n_noise=20):
    """The library's work-balanced cut (outfit_b200_shard_ranges; pure host arithmetic): [(t_begin, t_end)] * n_parts."""
    off = np.ascontiguousarray(traj_offset, dtype=np.uint64)
    cuts = np.zeros(n_parts + 1, dtype=np.uint64)
    rc = load_library().outfit_b200_shard_ranges(len(off) - 1, off.ctypes.data, int(max_triplets), int(n_noise), int(n_parts),
                                                 cuts.ctypes.data)
    if rc != 0:
        raise OutfitError(rc, "shard_ranges")
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(n_parts)]


def _p(a):
    """Host pointer of a C-contiguous numpy array, or device pointer of a torch tensor, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    return int(a)


class OutfitB200:
    """One context = one GPU (one process per GPU).  Mirrors the objects `fit_full_iod` borrows:
    the ephemeris (`&JPLEphem`) is loaded once; params and the batch are passed per call."""

    def __init__(self, device=-1):
        L = load_library()
        h = C.c_void_p()
        rc = L.outfit_b200_init(device, C.byref(h))
        if rc != 0:
            raise OutfitError(rc, L.outfit_b200_strerror(rc).decode())
        self._h = h
        self._L = L

    def close(self):
        if getattr(self, "_h", None):
            self._L.outfit_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._L.outfit_b200_last_error(self._h).decode() or self._L.outfit_b200_strerror(rc).decode()
            raise OutfitError(rc, msg)

    # -- JPLEphem -----------------------------------------------------------------------------
    def load_ephemeris(self, table):
        cheb = np.ascontiguousarray(table["cheb"], dtype=np.float64)
        ipt = np.ascontiguousarray(table["ipt"], dtype=np.uint32)
        self._check(self._L.outfit_b200_load_ephemeris(self._h, cheb.ctypes.data, cheb.shape[0], cheb.shape[1],
                                                       float(table["jd_start"]), float(table["block_days"]),
                                                       ipt.ctypes.data, float(table["emrat"])))

    # -- FitIOD::fit_full_iod -----------------------------------------------------------------
    @staticmethod
    def _batch_struct(batch, use_body_fixed=False):
        b = ObsBatch()
        b.n_traj = len(batch["traj_offset"]) - 1 if not hasattr(batch["traj_offset"], "data_ptr") else batch["traj_offset"].numel() - 1
        mj = batch["mjd_tt"]
        b.n_obs = mj.shape[0]
        for k in ("traj_offset", "mjd_tt", "ra", "dec", "sigma_ra", "sigma_dec"):
            setattr(b, k, _p(batch[k]))
        if use_body_fixed:
            b.observer_body_fixed = _p(batch["body_fixed"])
            b.mjd_ut1 = _p(batch["mjd_ut1"])
        else:
            b.obs_helio_equ = _p(batch["helio_equ"])
            b.obs_geo_ecl = _p(batch["geo_ecl"])
        b.noise_z = _p(batch.get("noise_z"))
        b.traj_seed = _p(batch.get("traj_seed"))
        b.max_obs_per_traj = int(batch.get("max_obs_per_traj", 0))
        return b

    def fit_full_iod(self, batch, params, use_body_fixed=False, out=None):
        """HOST buffers in, numpy structured array (RESULT_DTYPE) out; H2D/D2H inside the call.
        `out`: optional caller-owned RESULT_DTYPE array of n_traj records to write into (the C-ABI's buffers
        are caller-owned; a page-locked one makes the final D2H copy asynchronous and ~5x faster)."""
        b = self._batch_struct(batch, use_body_fixed)
        if params.n_noise_realizations == 0:
            b.noise_z = None
        if out is None:
            out = np.zeros(int(b.n_traj), dtype=RESULT_DTYPE)
        assert out.dtype == RESULT_DTYPE and out.shape == (int(b.n_traj),) and out.flags["C_CONTIGUOUS"]
        self._check(self._L.outfit_b200_fit_full_iod(self._h, C.byref(params), C.byref(b), out.ctypes.data))
        return out

    def fit_iod(self, batch, params, traj_index, use_body_fixed=False):
        """FitIOD::fit_iod (obs_dataset_api.rs:118-143): one trajectory of the batch -> one RESULT_DTYPE record."""
        b = self._batch_struct(batch, use_body_fixed)
        if params.n_noise_realizations == 0:
            b.noise_z = None
        out = np.zeros(1, dtype=RESULT_DTYPE)
        self._check(self._L.outfit_b200_fit_iod(self._h, C.byref(params), C.byref(b), int(traj_index), out.ctypes.data))
        return out[0]

    def fit_full_iod_device(self, dev_batch, params, out_ptr, stream=0, use_body_fixed=False):
        """DEVICE-resident buffers (torch tensors or raw pointers); enqueues, does not sync."""
        b = self._batch_struct(dev_batch, use_body_fixed)
        if params.n_noise_realizations == 0:
            b.noise_z = None
        self._check(self._L.outfit_b200_fit_full_iod_device(self._h, C.byref(params), C.byref(b), _p(out_ptr), stream))

    # -- FitLSQ::fit_lsq (differential_orbit_correction/obs_dataset_api.rs:113-190) ----------------
    def fit_lsq(self, batch, iod_params, cfg=None, initial_orbits=None, use_body_fixed=False):
        """HOST buffers.  initial_orbits: the RESULT_DTYPE array of fit_full_iod on the same batch, or None
        to run the IOD first (like the reference with `initial_orbits = None`).
        Returns (LSQ_RESULT_DTYPE[n_traj], OBS_FIT_DTYPE[n_obs])."""
        cfg = cfg or DifferentialCorrectionConfig.default()
        b = self._batch_struct(batch, use_body_fixed)
        if iod_params is None or iod_params.n_noise_realizations == 0:
            b.noise_z = None
        out = np.zeros(int(b.n_traj), dtype=LSQ_RESULT_DTYPE)
        fit = np.zeros(int(b.n_obs), dtype=OBS_FIT_DTYPE)
        io = None
        if initial_orbits is not None:
            io = np.ascontiguousarray(initial_orbits, dtype=RESULT_DTYPE)
            assert io.shape[0] == int(b.n_traj)
        self._check(self._L.outfit_b200_fit_lsq(self._h, C.byref(iod_params) if iod_params is not None else None,
                                                C.byref(cfg), C.byref(b), io.ctypes.data if io is not None else None,
                                                out.ctypes.data, fit.ctypes.data))
        return out, fit

    def fit_lsq_nbody(self, batch, initial_orbits, gm, perturber_pos, cfg=None, nbody=None, use_body_fixed=False):
        """FitLSQ with PropagatorKind::NBody (HOST buffers): initial_orbits = the RESULT_DTYPE array of fit_full_iod on the
        same batch, gm (P,), perturber_pos (P, 3, n_traj) = the perturbers at each IOD epoch (heliocentric, ecliptic
        J2000).  Returns (LSQ_RESULT_DTYPE[n_traj], OBS_FIT_DTYPE[n_obs])."""
        cfg = cfg or DifferentialCorrectionConfig.default()
        b = self._batch_struct(batch, use_body_fixed)
        b.noise_z = None
        T, P = int(b.n_traj), int(len(gm))
        nb = nbody or NBodyConfig(n_perturbers=P)
        nb.n_perturbers = P
        gm = np.ascontiguousarray(gm, dtype=np.float64)
        pp = np.ascontiguousarray(perturber_pos, dtype=np.float64)
        assert pp.shape == (P, 3, T)
        io = np.ascontiguousarray(initial_orbits, dtype=RESULT_DTYPE)
        assert io.shape[0] == T
        out = np.zeros(T, dtype=LSQ_RESULT_DTYPE)
        fit = np.zeros(int(b.n_obs), dtype=OBS_FIT_DTYPE)
        self._check(self._L.outfit_b200_fit_lsq_nbody(self._h, C.byref(cfg), C.byref(nb), _p(gm), _p(pp), C.byref(b),
                                                      io.ctypes.data, out.ctypes.data, fit.ctypes.data))
        return out, fit

    def fit_lsq_device(self, dev_batch, cfg, iod_ptr, out_ptr, fit_ptr, stream=0, use_body_fixed=False):
        """DEVICE-resident buffers; enqueues, does not sync."""
        b = self._batch_struct(dev_batch, use_body_fixed)
        b.noise_z = None
        self._check(self._L.outfit_b200_fit_lsq_device(self._h, C.byref(cfg), C.byref(b), _p(iod_ptr), _p(out_ptr),
                                                       _p(fit_ptr), stream))

    def last_iod_counters(self):
        c = IodCounters()
        self._check(self._L.outfit_b200_last_iod_counters(self._h, C.byref(c)))
        return {n: getattr(c, n) for n, _ in IodCounters._fields_}

    def set_work_counters(self, enabled):
        """Select the counting (exact work counters) or the plain instantiation of the kernels."""
        self._check(self._L.outfit_b200_set_work_counters(self._h, 1 if enabled else 0))

    def set_pass_streams(self, n_streams):
        """Passes in flight for large batches (default 8); 1 = single pass, per-phase timings available."""
        self._check(self._L.outfit_b200_set_pass_streams(self._h, int(n_streams)))

    def last_iod_phase_ms(self):
        """CUDA-event durations (ms) of each kernel of the last full-IOD launch (blocks until done)."""
        c = IodPhaseMs()
        self._check(self._L.outfit_b200_last_iod_phase_ms(self._h, C.byref(c)))
        return {n: getattr(c, n) for n, _ in IodPhaseMs._fields_}

    # -- OutfitCache::build ---------------------------------------------------------------------
    def observer_cache_device(self, n, mjd_tt, mjd_ut1, body_fixed, geo_ecl, helio_equ, status=None, stream=0):
        self._check(self._L.outfit_b200_observer_cache_device(self._h, n, _p(mjd_tt), _p(mjd_ut1), _p(body_fixed),
                                                              _p(geo_ecl), _p(helio_equ), _p(status), stream))

    # -- kepler::propagate_universal --------------------------------------------------------------
    def propagate_universal(self, rv, t0, t1, solver=None, psi_guess=None):
        """rv (6, n) plane-major, t0/t1 (n) host arrays -> out (11, n), status (n)."""
        solver = solver or SolverType(kind=2)
        n = rv.shape[1]
        out = np.empty((11, n), dtype=np.float64)
        status = np.empty(n, dtype=np.int32)
        self._check(self._L.outfit_b200_propagate_universal(self._h, n, _p(rv), _p(t0), _p(t1), _p(psi_guess),
                                                            C.byref(solver), out.ctypes.data, status.ctypes.data))
        return out, status

    def propagate_universal_device(self, n, rv, t0, t1, out, status, solver=None, psi_guess=None, stream=0):
        solver = solver or SolverType(kind=2)
        self._check(self._L.outfit_b200_propagate_universal_device(self._h, n, _p(rv), _p(t0), _p(t1), _p(psi_guess),
                                                                   C.byref(solver), _p(out), _p(status), stream))

    # -- OrbitalElements::compute::<Combined>, two-body ------------------------------------------
    EPHEMERIS_FIELDS = ("ra", "dec", "geocentric_dist", "heliocentric_dist", "phase_angle", "solar_elongation",
                        "radial_velocity", "d_ra_dt", "d_dec_dt")

    def ephemeris_twobody(self, kind, epoch, elem, mjd_tt, mjd_ut1, body_fixed):
        """HOST arrays: kind (n,) int32, epoch (n,), elem (6, n), mjd_tt / mjd_ut1 (E,), body_fixed (3,)
        -> out (9, E, n) in the order of EPHEMERIS_FIELDS, status (E, n)."""
        n, E = int(kind.shape[0]), int(mjd_tt.shape[0])
        out = np.empty((9, E, n), dtype=np.float64)
        status = np.empty((E, n), dtype=np.int32)
        bf = (C.c_double * 3)(*[float(x) for x in body_fixed])
        self._check(self._L.outfit_b200_ephemeris_twobody(self._h, n, _p(kind), _p(epoch), _p(elem), E, _p(mjd_tt),
                                                          _p(mjd_ut1), bf, out.ctypes.data, status.ctypes.data))
        return out, status

    def ephemeris_twobody_device(self, n, kind, epoch, elem, n_epochs, mjd_tt, mjd_ut1, body_fixed, out, status, stream=0):
        bf = (C.c_double * 3)(*[float(x) for x in body_fixed])
        self._check(self._L.outfit_b200_ephemeris_twobody_device(self._h, n, _p(kind), _p(epoch), _p(elem), n_epochs,
                                                                 _p(mjd_tt), _p(mjd_ut1), bf, _p(out), _p(status), stream))

    def ephemeris_request(self, kind, epoch, elem, observers, out=None, status=None):
        """EphemerisRequest with several observers (request.rs:276-340): observers = [(body_fixed (3,), mjd_tt (E_o,),
        mjd_ut1 (E_o,)), ...] -> out (9, E_total, n), status (E_total, n), epochs in request order."""
        n = int(kind.shape[0])
        bf, off, tt, ut = _flatten_request(observers)
        E = int(off[-1])
        out = np.empty((9, E, n), dtype=np.float64) if out is None else out
        status = np.empty((E, n), dtype=np.int32) if status is None else status
        self._check(self._L.outfit_b200_ephemeris_request(self._h, n, _p(kind), _p(epoch), _p(elem), len(observers), _p(bf),
                                                          _p(off), _p(tt), _p(ut), out.ctypes.data, status.ctypes.data))
        return out, status

    def compute_ephemerides(self, results, observers, iod_results=None):
        """FullOrbitResultExt::compute_ephemerides (ephemeris/batch.rs:134-183): the ephemeris request for every orbit of
        a fit result array (IOD or LSQ records); failed fits give status 9 (InvalidConversion) entries."""
        return _compute_ephemerides(self, results, observers, iod_results)

    def propagate_nbody(self, kind, epoch, elem, t1, gm, perturber_pos, config=None, with_stm=True):
        """EquinoctialElements::propagate_nbody in bulk (HOST arrays): kind (n,) int32, epoch (n,), elem (6, n), t1 (n,),
        gm (P,), perturber_pos (P, 3, n) heliocentric ecliptic J2000 at each orbit's epoch
        -> state (6, n) ecliptic, stm (36, n) column-major or None, status (n,), steps (n,)."""
        n, P = int(kind.shape[0]), int(len(gm))
        cfg = config or NBodyConfig(n_perturbers=P)
        cfg.n_perturbers = P
        gm = np.ascontiguousarray(gm, dtype=np.float64)
        pp = np.ascontiguousarray(perturber_pos, dtype=np.float64)
        assert pp.shape == (P, 3, n)
        out = np.empty((6, n))
        stm = np.empty((36, n)) if with_stm else None
        status = np.empty(n, dtype=np.int32)
        steps = np.empty(n, dtype=np.uint32)
        self._check(self._L.outfit_b200_propagate_nbody(self._h, n, _p(kind), _p(epoch), _p(elem), _p(t1), C.byref(cfg), _p(gm),
                                                        _p(pp), out.ctypes.data, stm.ctypes.data if with_stm else None,
                                                        status.ctypes.data, steps.ctypes.data))
        return out, stm, status, steps

    def ephemeris_nbody(self, kind, epoch, elem, observers, gm, perturber_pos, config=None):
        """OrbitalElements::compute::<Combined> with PropagatorKind::NBody (HOST arrays): observers as in ephemeris_request,
        gm / perturber_pos as in propagate_nbody -> out (9, E_total, n), status (E_total, n)."""
        n, P = int(kind.shape[0]), int(len(gm))
        cfg = config or NBodyConfig(n_perturbers=P)
        cfg.n_perturbers = P
        bf, off, tt, ut = _flatten_request(observers)
        E = int(off[-1])
        gm = np.ascontiguousarray(gm, dtype=np.float64)
        pp = np.ascontiguousarray(perturber_pos, dtype=np.float64)
        assert pp.shape == (P, 3, n)
        out = np.empty((9, E, n), dtype=np.float64)
        status = np.empty((E, n), dtype=np.int32)
        self._check(self._L.outfit_b200_ephemeris_nbody(self._h, n, _p(kind), _p(epoch), _p(elem), len(observers), _p(bf), _p(off),
                                                        _p(tt), _p(ut), C.byref(cfg), _p(gm), _p(pp), out.ctypes.data,
                                                        status.ctypes.data))
        return out, status

    def set_ephemeris_config(self, config):
        """EphemerisConfig of the ephemeris entries of this context (aberration order; two-body propagator only)."""
        self._check(self._L.outfit_b200_set_ephemeris_config(self._h, C.byref(config)))

    def selftest_arith(self, n, seed=1, exp_range=60):
        """Mismatch counts (rcp, div, sqrt, sincos) of the library's own arithmetic against CUDA's."""
        out = (C.c_uint64 * 4)()
        self._check(self._L.outfit_b200_selftest_arith(self._h, int(n), int(seed), int(exp_range), out))
        return tuple(int(x) for x in out)

    def measure_fp64_peak(self):
        v = C.c_double()
        self._check(self._L.outfit_b200_measure_fp64_peak(self._h, C.byref(v)))
        return v.value


def ephemeris_mode_epochs(mode):
    """EphemerisMode::epochs() (ephemeris/request.rs:246-267) as MJD days: ("single", t) | ("at", [t, ...]) |
    ("range", start, end, step_seconds): start, start + step, ... while <= end; a non-positive step or start > end gives
    no epochs.  The reference steps hifitime Epochs, i.e. exact integer nanoseconds: so does this (no drift of a float sum)."""
    kind = mode[0]
    if kind == "single":
        return np.array([float(mode[1])])
    if kind == "at":
        return np.asarray(mode[1], dtype=np.float64).copy()
    if kind != "range":
        raise ValueError("mode must be ('single', t), ('at', [...]) or ('range', start, end, step_seconds)")
    start, end, step_s = float(mode[1]), float(mode[2]), float(mode[3])
    ns_day = 86400 * 10 ** 9
    a, b, h = round(start * ns_day), round(end * ns_day), round(step_s * 10 ** 9)
    if h <= 0 or a > b:
        return np.zeros(0)
    k = np.arange((b - a) // h + 1, dtype=np.int64)
    return (a + k * h) / ns_day


def _flatten_request(observers):
    bf = np.ascontiguousarray([np.asarray(o[0], dtype=np.float64) for o in observers]).reshape(-1)
    off = np.concatenate([[0], np.cumsum([len(o[1]) for o in observers])]).astype(np.uint64)
    tt = np.ascontiguousarray(np.concatenate([np.asarray(o[1], dtype=np.float64) for o in observers]))
    ut = np.ascontiguousarray(np.concatenate([np.asarray(o[2], dtype=np.float64) for o in observers]))
    return bf, off, tt, ut


def pinned_empty(shape, dtype):
    """numpy array over page-locked host memory from outfit_b200_host_alloc (kept alive by the array's base)."""
    L = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = L.outfit_b200_host_alloc(n)
    if not ptr:
        raise OutfitError(-4, "outfit_b200_host_alloc failed")

    class _Owner:
        def __init__(self, p):
            self.p = p

        def __del__(self):
            try:
                L.outfit_b200_host_free(self.p)
            except Exception:
                pass
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    buf._owner = _Owner(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)



def orbits_of_results(results, iod_results=None):
    """(kind (n,) int32, epoch (n,), elem (6, n), ok (n,) bool) of fit records, for the ephemeris entries: RESULT_DTYPE
    records give the IOD orbit; LSQ_RESULT_DTYPE records give the corrected equinoctial orbit, or -- for an IOD fallback
    (kind 2) -- the IOD orbit, whose element type is in `iod_results`."""
    r = np.asarray(results)
    n = len(r)
    kind = np.zeros(n, dtype=np.int32)
    if "element_kind" in r.dtype.names:
        ok = r["status"] == 0
        kind[:] = r["element_kind"]
    else:
        ok = r["kind"] != 0
        kind[r["kind"] == 1] = 1  # equinoctial
        fb = r["kind"] == 2
        if fb.any():
            if iod_results is None:
                raise ValueError("LSQ records with IOD fallbacks need iod_results for the element type")
            kind[fb] = np.asarray(iod_results)["element_kind"][fb]
    epoch = np.where(ok, r["epoch"], 0.0)
    elem = np.ascontiguousarray(np.where(ok[:, None], r["elem"], 0.0).T)
    return kind, np.ascontiguousarray(epoch), elem, ok


def _compute_ephemerides(engine, results, observers, iod_results=None):
    kind, epoch, elem, ok = orbits_of_results(results, iod_results)
    # a failed fit is Err(InvalidConversion) in the reference's map (ephemeris/batch.rs:141-147): its entries carry
    # status 9 and NaN, whatever the kernel made of the zeroed elements
    out, status = engine.ephemeris_request(kind, epoch, elem, observers)
    out[:, :, ~ok] = np.nan
    status[:, ~ok] = 9
    return out, status


class OutfitGroup:
    """Every GPU of the box behind ONE call: the drop-in for `fit_full_iod_parallel` (obs_dataset_api.rs:175-207).
    `devices`: None = all visible GPUs, an int = the first n, or a list of device ids (an id may repeat:
    several contexts on one GPU, which is how the sharded path is tested on a single-GPU box)."""

    def __init__(self, devices=None):
        L = load_library()
        h = C.c_void_p()
        if devices is None:
            rc = L.outfit_b200_init_multi(0, None, C.byref(h))
        elif isinstance(devices, int):
            rc = L.outfit_b200_init_multi(devices, None, C.byref(h))
        else:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = L.outfit_b200_init_multi(len(devices), ids, C.byref(h))
        if rc != 0:
            raise OutfitError(rc, L.outfit_b200_strerror(rc).decode())
        self._h, self._L = h, L

    def close(self):
        if getattr(self, "_h", None):
            self._L.outfit_b200_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._L.outfit_b200_group_size(self._h))

    def _check(self, rc):
        if rc != 0:
            msg = self._L.outfit_b200_group_last_error(self._h).decode() or self._L.outfit_b200_strerror(rc).decode()
            raise OutfitError(rc, msg)

    def load_ephemeris(self, table):
        cheb = np.ascontiguousarray(table["cheb"], dtype=np.float64)
        ipt = np.ascontiguousarray(table["ipt"], dtype=np.uint32)
        self._check(self._L.outfit_b200_group_load_ephemeris(self._h, cheb.ctypes.data, cheb.shape[0], cheb.shape[1],
                                                             float(table["jd_start"]), float(table["block_days"]),
                                                             ipt.ctypes.data, float(table["emrat"])))

    def set_pass_streams(self, n):
        self._check(self._L.outfit_b200_group_set_pass_streams(self._h, int(n)))

    def set_ephemeris_config(self, config):
        self._check(self._L.outfit_b200_group_set_ephemeris_config(self._h, C.byref(config)))

    def fit_full_iod(self, batch, params, use_body_fixed=False, out=None):
        b = OutfitB200._batch_struct(batch, use_body_fixed)
        if params.n_noise_realizations == 0:
            b.noise_z = None
        if out is None:
            out = np.zeros(int(b.n_traj), dtype=RESULT_DTYPE)
        assert out.dtype == RESULT_DTYPE and out.shape == (int(b.n_traj),) and out.flags["C_CONTIGUOUS"]
        self._check(self._L.outfit_b200_group_fit_full_iod(self._h, C.byref(params), C.byref(b), out.ctypes.data))
        return out

    def fit_lsq(self, batch, iod_params, cfg=None, initial_orbits=None, use_body_fixed=False):
        cfg = cfg or DifferentialCorrectionConfig.default()
        b = OutfitB200._batch_struct(batch, use_body_fixed)
        if iod_params is None or iod_params.n_noise_realizations == 0:
            b.noise_z = None
        out = np.zeros(int(b.n_traj), dtype=LSQ_RESULT_DTYPE)
        fit = np.zeros(int(b.n_obs), dtype=OBS_FIT_DTYPE)
        io = None
        if initial_orbits is not None:
            io = np.ascontiguousarray(initial_orbits, dtype=RESULT_DTYPE)
        self._check(self._L.outfit_b200_group_fit_lsq(self._h, C.byref(iod_params) if iod_params is not None else None,
                                                      C.byref(cfg), C.byref(b), io.ctypes.data if io is not None else None,
                                                      out.ctypes.data, fit.ctypes.data))
        return out, fit

    def fit_lsq_nbody(self, batch, initial_orbits, gm, perturber_pos, cfg=None, nbody=None, use_body_fixed=False):
        cfg = cfg or DifferentialCorrectionConfig.default()
        b = OutfitB200._batch_struct(batch, use_body_fixed)
        b.noise_z = None
        T, P = int(b.n_traj), int(len(gm))
        nb = nbody or NBodyConfig(n_perturbers=P)
        nb.n_perturbers = P
        gm = np.ascontiguousarray(gm, dtype=np.float64)
        pp = np.ascontiguousarray(perturber_pos, dtype=np.float64)
        assert pp.shape == (P, 3, T)
        io = np.ascontiguousarray(initial_orbits, dtype=RESULT_DTYPE)
        out = np.zeros(T, dtype=LSQ_RESULT_DTYPE)
        fit = np.zeros(int(b.n_obs), dtype=OBS_FIT_DTYPE)
        self._check(self._L.outfit_b200_group_fit_lsq_nbody(self._h, C.byref(cfg), C.byref(nb), _p(gm), _p(pp), C.byref(b),
                                                            io.ctypes.data, out.ctypes.data, fit.ctypes.data))
        return out, fit

    def propagate_universal(self, rv, t0, t1, solver=None, psi_guess=None, out=None, status=None):
        solver = solver or SolverType(kind=2)
        n = rv.shape[1]
        out = np.empty((11, n), dtype=np.float64) if out is None else out
        status = np.empty(n, dtype=np.int32) if status is None else status
        self._check(self._L.outfit_b200_group_propagate_universal(self._h, n, _p(rv), _p(t0), _p(t1), _p(psi_guess),
                                                                  C.byref(solver), out.ctypes.data, status.ctypes.data))
        return out, status

    def ephemeris_request(self, kind, epoch, elem, observers, out=None, status=None):
        n = int(kind.shape[0])
        bf, off, tt, ut = _flatten_request(observers)
        E = int(off[-1])
        out = np.empty((9, E, n), dtype=np.float64) if out is None else out
        status = np.empty((E, n), dtype=np.int32) if status is None else status
        self._check(self._L.outfit_b200_group_ephemeris_request(self._h, n, _p(kind), _p(epoch), _p(elem), len(observers),
                                                                _p(bf), _p(off), _p(tt), _p(ut), out.ctypes.data,
                                                                status.ctypes.data))
        return out, status

    def compute_ephemerides(self, results, observers, iod_results=None):
        """FullOrbitResultExt::compute_ephemerides_parallel (ephemeris/batch.rs:149-183) over every GPU of the group."""
        return _compute_ephemerides(self, results, observers, iod_results)

    def last_shards(self):
        """(cuts [n+1], wall ms per shard [n]) of the last group call."""
        n = len(self)
        cuts = np.zeros(n + 1, dtype=np.uint64)
        ms = np.zeros(n, dtype=np.float32)
        self._check(self._L.outfit_b200_group_last_shards(self._h, cuts.ctypes.data, ms.ctypes.data))
        return cuts, ms
