#!/usr/bin/env python3
"""Regenerates tests/golden/config1_2015AB.json: the observations of the reference's single-trajectory
quick-start case (tests/data/2015AB.obs, BASELINE configs[0]) parsed by outfit_b200.mpc80 -- numbers
only (UTC MJD, RA, Dec in radians, observatory code).  Run in the build container, where
/root/reference exists; the fixture travels to the GPU box instead of the reference."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from outfit_b200 import mpc80

text = open("/root/reference/tests/data/2015AB.obs").read()
trajs = mpc80.parse(text, single_trajectory=True)
assert list(trajs) == ["K15A00B"] and len(trajs["K15A00B"]) == 37
out = {"source": "reference tests/data/2015AB.obs (MPC 80-column), parsed by outfit_b200.mpc80",
       "designation": "K15A00B",
       "records": [{k: r[k] for k in ("mjd_utc", "ra", "dec", "mag", "band", "obscode")} for r in trajs["K15A00B"]]}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "config1_2015AB.json"), "w"), indent=0)
print("wrote", len(out["records"]), "records")
