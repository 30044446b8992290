"""Element conversions of the LSQ records (outfit_b200/elements.py) against the oracle's restatement, the
reference's exact KAT (equinoctial_element.rs:1240-1264) and finite differences (the reference's own
Jacobian tests, equinoctial_element.rs:1645-1740, use the same check).  CPU only."""
import ctypes as C

import numpy as np

from oracle import binding as O
from outfit_b200 import elements as E

dbl = C.c_double


def _oracle_kep(eq):
    out = (dbl * 6)()
    O.lib().oo_equinoctial_to_keplerian((dbl * 6)(*eq), out)
    return np.array(out)


def _oracle_jac(eq):
    out = (dbl * 36)()
    O.lib().oo_jacobian_to_keplerian((dbl * 6)(*eq), out)
    return np.array(out).reshape(6, 6).T  # column-major -> [row, col]


def _random_equinoctial(n, seed):
    rng = np.random.default_rng(seed)
    e, varpi = rng.uniform(0.0, 0.9, n), rng.uniform(0, 2 * np.pi, n)
    t, node = np.tan(rng.uniform(0.0, 2.5, n) / 2), rng.uniform(0, 2 * np.pi, n)
    return np.stack([rng.uniform(0.5, 40, n), e * np.sin(varpi), e * np.cos(varpi), t * np.sin(node), t * np.cos(node),
                     rng.uniform(0, 2 * np.pi, n)], axis=1)


def test_reference_kat_equinoctial_to_keplerian_is_exact():
    eq = [1.8017360713, 0.2693736809404963, 0.08856415260522467, 0.0008089970142830734, 0.10168201110394352, 1.693697008]
    want = [1.8017360713, 0.2835591457, 0.20267383289999996, 0.007955979, 1.2451951388, 0.4405458902000001]
    assert list(_oracle_kep(eq)) == want
    assert np.abs(E.equinoctial_to_keplerian(eq)[0] - want).max() <= 4e-16


def test_conversions_match_the_oracle_on_random_orbits():
    eq = _random_equinoctial(500, 1)
    eq[0, 1:3] = 0.0            # circular: varpi undefined -> 0
    eq[1, 3:5] = 0.0            # equatorial: node undefined -> 0
    eq[2, 1:5] = 1e-14          # both below the 1e-12 thresholds
    kep = E.equinoctial_to_keplerian(eq)
    jac = E.jacobian_to_keplerian(eq)
    for t in range(len(eq)):
        wk = _oracle_kep(eq[t])
        d = np.abs(kep[t] - wk)
        d[3:] = np.minimum(d[3:], 2 * np.pi - d[3:])
        assert d.max() <= 1e-14 * max(1.0, np.abs(wk).max()), (t, kep[t], wk)
        wj = _oracle_jac(eq[t])
        assert np.allclose(jac[t], wj, rtol=1e-14, atol=0), t
    assert kep[0, 3] != 0.0 and kep[1, 3] == 0.0 and kep[2, 3] == 0.0


def test_jacobian_against_finite_differences():
    eq = _random_equinoctial(40, 2)
    eq[:, 1:3] *= 0.5
    jac = E.jacobian_to_keplerian(eq)
    step = 1e-7
    for j in range(6):
        hi, lo = eq.copy(), eq.copy()
        hi[:, j] += step
        lo[:, j] -= step
        d = E.equinoctial_to_keplerian(hi) - E.equinoctial_to_keplerian(lo)
        d[:, 3:] = (d[:, 3:] + np.pi) % (2 * np.pi) - np.pi
        fd = d / (2 * step)
        assert np.allclose(jac[:, :, j], fd, rtol=2e-5, atol=2e-6), j


def test_covariance_propagation_matches_the_oracle_and_keeps_symmetry():
    eq = _random_equinoctial(50, 3)
    rng = np.random.default_rng(4)
    a = rng.normal(size=(50, 6, 6)) * 1e-4
    cov = a @ a.transpose(0, 2, 1)
    jac = E.jacobian_to_keplerian(eq)
    out = E.propagate_covariance(cov, jac)
    for t in range(50):
        res = (dbl * 36)()
        O.lib().oo_propagate_covariance((dbl * 36)(*jac[t].T.reshape(-1)), (dbl * 36)(*cov[t].T.reshape(-1)), res)
        want = np.array(res).reshape(6, 6).T
        assert np.allclose(out[t], want, rtol=1e-12, atol=1e-24)
    assert np.allclose(out, out.transpose(0, 2, 1), rtol=1e-10, atol=1e-24)
    assert (E.keplerian_sigmas(out) >= 0).all()


def test_lsq_records_to_keplerian(oracle):
    from outfit_b200 import synth
    table = synth.make_ephemeris_table()
    et = O.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    batch = synth.make_trajectories(150, 12, seed=9, table=table, max_triplets=10, n_noise=1)
    ob = O.from_soa_batch(batch)
    iod = O.fit_full_iod(ob, et, O.default_iod_params(n_noise_realizations=0, max_triplets=10), n_threads=0)
    res, _ = O.fit_lsq(ob, et, O.default_lsq_config(), iod)
    k = E.lsq_to_keplerian(res)
    ok = k["valid"]
    assert ok.sum() > 40 and np.isnan(k["elem"][~ok]).all()
    # same orbit: a kept, e = |(h, k)|, and the fit moved the IOD's Keplerian elements only a little
    assert np.array_equal(k["elem"][ok, 0], res["elem"][ok, 0])
    kep_iod = ok & (iod["element_kind"] == 0)
    assert np.median(np.abs(k["elem"][kep_iod, 1] - iod["elem"][kep_iod, 1])) < 0.05
    # sigma(a) is invariant (row/column of a is the identity in J), the others are finite and positive
    assert np.allclose(k["sigma"][ok, 0], res["sigma"][ok, 0], rtol=1e-12)
    assert np.isfinite(k["sigma"][ok]).all() and (k["sigma"][ok] > 0).all()


# tests/test_orbit_uncertainty_propag.rs:11-170 of the reference: an equinoctial orbit with its covariance, turned into
# Keplerian form by OrbitalElements::to_keplerian (J C J^T); every element, sigma and covariance entry within 1e-10 ABSOLUTE
# (tests/common/mod.rs:5-80, abs_diff_eq).  No ephemeris involved: a golden that runs here.
UNC_EQ = [1.8021517900042052, 0.2694922786015968, 0.08955282358108035, 0.0008974287327937245, 0.10167548786557225,
          1.6921653421358704]
UNC_COV_EQ = [
    [3.651448459073842e-12, -4.87907485491453e-13, 2.321298362132558e-11, -3.7695250201166625e-13, 8.511532638002078e-13, -3.91138523482157e-11],
    [-4.879074854914533e-13, 7.437576190456506e-12, -1.1647669978804286e-11, 9.359797430147383e-13, -2.8577594338429333e-12, 1.853502993770551e-11],
    [2.3212983621325566e-11, -1.164766997880434e-11, 1.577521262959403e-10, -3.47676746499932e-12, 8.610023673871895e-12, -2.644913915663376e-10],
    [-3.7695250201166625e-13, 9.359797430147385e-13, -3.4767674649993202e-12, 3.7739327795249603e-13, -5.048815271306508e-13, 5.7505636344116006e-12],
    [8.511532638002078e-13, -2.857759433842935e-12, 8.610023673871898e-12, -5.048815271306507e-13, 1.3170255261786945e-12, -1.4110008489365913e-11],
    [-3.911385234821569e-11, 1.8535029937705585e-11, -2.6449139156633765e-10, 5.750563634411601e-12, -1.4110008489365913e-11, 4.437117125245391e-10]]
UNC_KEP = [1.8021517900042052, 0.2839820354128493, 0.20266238925780133, 0.008826172835575467, 1.2411480851756391, 0.4421910841246559]
UNC_SIGMA_KEP = [1.910876358918557e-6, 3.926080684435881e-6, 2.2639852329024065e-6, 6.113264876575711e-6, 4.049775340683106e-5,
                 2.2182426229638676e-5]
UNC_COV_KEP = [
    [3.651448459073842e-12, 6.857127156611333e-12, 1.6782354228854548e-12, -3.781001511911568e-12, -7.433110873463038e-11, 3.899825789832625e-11],
    [6.857127156611329e-12, 1.5414109540700513e-11, 2.690953229794561e-15, -2.0474618140821963e-12, -1.2349406349235225e-10, 5.97243215927523e-11],
    [1.6782354228854548e-12, 2.6909532297930087e-15, 5.1256291348001634e-12, -9.989144038881854e-12, -5.3024087432235095e-11, 3.518354634255312e-11],
    [-3.781001511911568e-12, -2.047461814082196e-12, -9.989144038881855e-12, 3.7372007451174244e-11, 8.98813435388229e-11, -6.947495524468516e-11],
    [-7.433110873463033e-11, -1.2349406349235207e-10, -5.302408743223507e-11, 8.988134353882289e-11, 1.6400680310004965e-9, -8.833005679743845e-10],
    [3.8998257898326207e-11, 5.972432159275218e-11, 3.5183546342553095e-11, -6.947495524468513e-11, -8.833005679743845e-10, 4.920600334333619e-10]]


def test_reference_golden_uncertainty_propagation_to_keplerian():
    kep = _oracle_kep(UNC_EQ)
    assert np.abs(kep - UNC_KEP).max() < 1e-10
    assert np.abs(E.equinoctial_to_keplerian(UNC_EQ)[0] - UNC_KEP).max() < 1e-10
    cov = np.array(UNC_COV_EQ)
    # the oracle: column-major in, column-major out
    out = (dbl * 36)()
    jac = (dbl * 36)()
    O.lib().oo_jacobian_to_keplerian((dbl * 6)(*UNC_EQ), jac)
    O.lib().oo_propagate_covariance(jac, (dbl * 36)(*cov.T.reshape(-1)), out)
    ocov = np.array(out).reshape(6, 6).T
    assert np.abs(ocov - np.array(UNC_COV_KEP)).max() < 1e-10
    assert np.abs(np.sqrt(np.diag(ocov)) - UNC_SIGMA_KEP).max() < 1e-10
    # the host-side module, and the record path of lsq_to_keplerian
    hcov = E.propagate_covariance(cov[None], E.jacobian_to_keplerian(UNC_EQ))[0]
    assert np.abs(hcov - np.array(UNC_COV_KEP)).max() < 1e-10
    # tighter than the reference's own bound: the restatement reproduces its numbers to rounding
    assert np.abs(hcov - np.array(UNC_COV_KEP)).max() < 1e-21 and np.abs(kep - UNC_KEP).max() < 1e-15
