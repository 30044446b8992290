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
