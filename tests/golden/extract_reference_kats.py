#!/usr/bin/env python3
"""Extract the reference's own known-answer tests (numbers only) into reference_kats.json.

Runs ONLY in the build container (needs /root/reference, which does not exist on the GPU box);
the JSON it writes is committed and is what tests/test_oracle_kats.py reads.  Every entry cites the
reference file:line range of the #[test] it was taken from.  No reference code is copied: the
script pulls the numeric literals of each test function, in order of appearance, and labels them.
"""
import json
import os
import re
import sys

REF = os.environ.get("OUTFIT_REFERENCE", "/root/reference")
NUM = re.compile(r"(?<![A-Za-z_0-9.])-?\d[\d_]*\.?[\d_]*(?:[eE][-+]?\d+)?(?![A-Za-z_0-9])")


def fn_body(path, fn_name):
    """Return (text, first_line, last_line) of `fn fn_name` (brace matched)."""
    src = open(os.path.join(REF, path)).read()
    m = re.search(r"fn\s+" + re.escape(fn_name) + r"\s*\(", src)
    if not m:
        raise KeyError(f"{path}: fn {fn_name} not found")
    i = src.index("{", m.end())
    depth, j = 0, i
    while True:
        c = src[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    first = src.count("\n", 0, m.start()) + 1
    last = src.count("\n", 0, j) + 1
    return src[i : j + 1], first, last


def strip_comments(text):
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"\[f64;\s*\d+\]", "", text)  # type annotations are not data
    return re.sub(r'"[^"\n]*"', '""', text)


def floats(text):
    out = []
    for m in NUM.finditer(strip_comments(text)):
        s = m.group(0).replace("_", "")
        out.append(float(s))
    return out


def kat(path, fn_name):
    body, a, b = fn_body(path, fn_name)
    return floats(body), f"src/{path}:{a}-{b}" if not path.startswith("src/") else f"{path}:{a}-{b}"


def main():
    K = {}
    G = "src/initial_orbit_determination/gauss.rs"

    # ---- gauss.rs -------------------------------------------------------------------------
    f, cite = kat(G, "test_gauss_prelim")
    # idx(3) ra(3) dec(3) time(3) | tau1 tau3 | unit(9) | inv(9) | a(3) | b(3)
    K["gauss_prelim"] = dict(cite=cite, ra=f[3:6], dec=f[6:9], time=f[9:12], tau1=f[12], tau3=f[13],
                             unit=f[14:23], inv_unit=f[23:32], a=f[32:35], b=f[35:38])
    f, cite = kat(G, "test_coeff_8poly")
    K["coeff_8poly"] = dict(cite=cite, ra=f[3:6], dec=f[6:9], time=f[9:12], obs_pos_rowmajor_T=f[12:21],
                            c6=f[21], c3=f[22], c0=f[23])
    f, cite = kat(G, "test_solving_polynom")
    K["solve_8poly"] = dict(cite=cite, poly=f[3:12], max_iter=int(f[12]), aberth_eps=f[13],
                            root_eps=f[14], roots=f[15:18])
    f, cite = kat(G, "test_asteroid_position")
    # idx ra dec time obs(9) | first_root, powi 3, 1., -1., | second_root, 3, 1., -1. | pos(9) epoch
    K["asteroid_position"] = dict(cite=cite, ra=f[3:6], dec=f[6:9], time=f[9:12],
                                  obs_pos_rowmajor_T=f[12:21], first_root=f[21], second_root=f[29],
                                  pos=f[-10:-1], epoch=f[-1])
    f, cite = kat(G, "test_gibbs_correction")
    K["gibbs"] = dict(cite=cite, ra=f[3:6], dec=f[6:9], time=f[9:12], pos_rowmajor_T=f[12:21],
                      vel=f[21:24])
    f, cite = kat(G, "test_solve_orbit")
    # tol | case1: idx(3) ra dec time obs(9, Matrix3::new row-major) | expected 7 | case2 (array form,
    # rows = columns of matrix since `[[..]].into()` is column-major) ...
    tol = f[0]
    c1 = f[1:]
    case1 = dict(ra=c1[3:6], dec=c1[6:9], time=c1[9:12], obs_pos_rowmajor=c1[12:21], expected=c1[21:28])
    c2 = c1[28:]
    case2 = dict(ra=c2[3:6], dec=c2[6:9], time=c2[9:12], obs_pos_colmajor=c2[12:21], expected=c2[21:28])
    c3 = c2[28:]
    case3 = dict(ra=c3[3:6], dec=c3[6:9], time=c3[9:12], obs_pos_rowmajor=c3[12:21], expected=c3[21:28])
    K["prelim_orbit"] = dict(cite=cite, tol=tol, cases=[case1, case2, case3],
                             expected_fields=["epoch", "a", "e", "i", "node", "argp", "M"])
    f, cite = kat(G, "test_orbit_correction")
    K["pos_and_vel_correction"] = dict(
        cite=cite, ra=f[3:6], dec=f[6:9], time=f[9:12], obs_pos_rowmajor_T=f[12:21],
        pos_rowmajor_T=f[21:30], vel=f[30:33], unit_rowmajor_T=f[33:42], inv_unit_rowmajor_T=f[42:51],
        peri_max=f[51], ecc_max=f[52], err_max=f[53], itmax=int(f[54]),
        new_pos=f[55:64], new_vel=f[64:67], epoch=f[67])

    # ---- kepler ---------------------------------------------------------------------------
    f, cite = kat("src/kepler/velocity.rs", "test_velocity_correction_real_data")
    K["velocity_correction"] = dict(cite=cite, x1=f[0:3], x2=f[3:6], v2=f[6:9], dt=f[9], peri_max=f[10],
                                    ecc_max=f[11], f=f[12], g=f[13], v=f[14:17], kep_eps=1e3 * 2.220446049250313e-16)
    f, cite = kat("src/kepler/stumpff.rs", "test_s_funct_real_data")
    K["s_funct"] = dict(cite=cite, psi=f[0], alpha=f[1], s=f[2:6])
    f, cite = kat("src/kepler/params.rs", "test_prelim_kepuni_real_data")
    K["prelim_kepuni"] = dict(cite=cite, dt=f[0], r0=f[1], sig0=f[2], mu=f[3], alpha=f[4], e0=f[5],
                              psi_elliptic=f[6], alpha_hyp=f[7], psi_hyperbolic=f[8],
                              psi_parabolic=f[10])
    f, cite = kat("src/kepler/params.rs", "test_returns_none_for_alpha_zero")
    K["prelim_kepuni_alpha_zero"] = dict(cite=cite, dt=f[0], r0=f[1], sig0=f[2], alpha=f[3], e0=f[4],
                                         psi=f[5], mu_name="MU", contr_name="CONTR")

    # propagate_universal: regular structure -> parse by regex on the body
    P = "src/kepler/propagation.rs"
    src = open(os.path.join(REF, P)).read()
    cases = []
    for name in ["test_propag", "test_propag2", "test_propag3", "test_propag4"]:
        f, cite = kat(P, name)
        # r(3) v(3) t0 t1 convergency [psi_guess] max_iter(20) | expected r1(3) v1(3) | 1e-9 1e-9
        has_guess = name == "test_propag4"
        o = 9 + (1 if has_guess else 0)
        cases.append(dict(name=name, cite=cite, r=f[0:3], v=f[3:6], t0=f[6], t1=f[7], convergency=f[8],
                          psi_guess=(f[9] if has_guess else None), r1=f[o + 1:o + 4], v1=f[o + 4:o + 7],
                          tol=1e-9, kind="auto"))
    T0 = 60000.0
    edge = [("test_quasi_circular_orbit", 1.0), ("test_high_eccentricity_near_perihelion", 2.0),
            ("test_near_parabolic_elliptic", 5.0), ("test_near_parabolic_hyperbolic", 5.0),
            ("test_hyperbolic_orbit", 10.0), ("test_negative_dt_backward_propagation", -10.0),
            ("test_gap_35_days_ztf_lsst_cadence", 35.0), ("test_gap_45_days_negative", -45.0),
            ("test_gap_150_days_short_period_neo", 150.0), ("test_gap_400_days_multi_revolution", 400.0)]
    for name, dt in edge:
        body, a, b = fn_body(P, name)
        vecs = re.findall(r"Vector3::new\(([^)]*)\)", strip_comments(body))
        vv = [[float(x.replace("_", "")) for x in v.split(",") if x.strip()] for v in vecs]
        tol = float(re.findall(r"assert_vec_close\([^;]*?(1e-\d+)", body)[0])
        m = re.search(r"T0 ([+-]) ([\d.eE-]+)", body)
        dt_src = float(m.group(2)) * (1 if m.group(1) == "+" else -1)
        assert dt_src == dt, (name, dt_src, dt)
        cases.append(dict(name=name, cite=f"{P}:{a}-{b}", r=vv[0], v=vv[1], t0=T0, t1=T0 + dt,
                          convergency=2.220446049250313e-14, psi_guess=None, r1=vv[2], v1=vv[3],
                          tol=tol, kind="auto"))
    K["propagate_universal"] = dict(cases=cases)
    body, a, b = fn_body(P, "test_dt_near_zero_returns_initial_state")
    vecs = re.findall(r"Vector3::new\(([^)]*)\)", strip_comments(body))
    vv = [[float(x.replace("_", "")) for x in v.split(",") if x.strip()] for v in vecs]
    K["propagate_universal_dt_zero"] = dict(cite=f"{P}:{a}-{b}", r=vv[0], v=vv[1], t0=T0, t1=T0 + 1e-8,
                                            tol=1e-8)

    # ---- elements -------------------------------------------------------------------------
    E = "src/orbit_type/equinoctial_element.rs"
    f, cite = kat(E, "test_kepler_equation")
    K["equinoctial_kepler_equation"] = dict(cite=cite, epoch=f[0], equ=f[1:7], lambda_t1=f[7],
                                            lon_peri=f[8], F=f[9])
    f, cite = kat(E, "test_two_body_problem")
    K["equinoctial_two_body"] = dict(cite=cite, epoch=f[0], equ=f[1:7], t0=f[7], t1=f[8], pos=f[9:12],
                                     vel=f[12:15])

    # ---- earth orientation / time ----------------------------------------------------------
    O = "src/earth_orientation.rs"
    f, cite = kat(O, "test_obliquity")
    K["obleq_t2000"] = dict(cite=cite, value=f[0])
    f, cite = kat(O, "test_nutn80")
    K["nutn80_t2000"] = dict(cite=cite, dpsi=f[0], deps=f[1])
    f, cite = kat(O, "test_rnut80")
    K["rnut80_t2000"] = dict(cite=cite, columns=f[0:9])  # [[..];3].into() = column-major
    f, cite = kat("src/time.rs", "test_gmst")
    K["gmst"] = dict(cite=cite, cases=[dict(tut=f[0], gmst=f[1]), dict(tut=51544.5, gmst=f[2])])

    f, cite = kat("src/orb_elem.rs", "test_elem_regression_reference_with_tolerance")
    K["ccek1"] = dict(cite=cite, r=f[0:3], v=f[3:6], epoch=f[6], elem=f[7:13], tol=5e-13)
    f, cite = kat("src/orb_elem.rs", "test_eccentricity_control")
    K["eccentricity_control"] = dict(cite=cite, r=f[0:3], v=f[3:6], peri_max=f[6], ecc_max=f[7],
                                     ecc=f[8], peri=f[9], energy=f[10])
    R = "src/ref_system.rs"
    f, cite = kat(R, "test_rotpn_equm")
    K["rotpn_equm"] = dict(cite=cite, equm_to_eclm_j2000=f[0:9], equm_to_equt_j2000=f[9:18])
    f, cite = kat(R, "test_rotpn_equt_eclm_date")
    K["rotpn_equt_date_eclm_j2000"] = dict(cite=cite, columns=f[0:9], tmjd=f[9], rel_eps=1e-17)
    I = "src/initial_orbit_determination/triplet_generation/index_generator.rs"
    K["downsample"] = dict(cite=I + ":333-400", cases=[dict(n=0, max_keep=10, keep=[]),
                           dict(n=5, max_keep=10, keep=[0, 1, 2, 3, 4]), dict(n=5, max_keep=5, keep=[0, 1, 2, 3, 4]),
                           dict(n=9, max_keep=3, keep=[0, 4, 8])])

    json.dump(K, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json"), "w"),
              indent=1)
    print("wrote", len(K), "KAT groups")


if __name__ == "__main__":
    main()
