"""What the UNPINNED assumptions are worth: the oracle is rebuilt with each assumption
about code outside /root/reference flipped (oracle/ft_oracle.c FTO_ASSUME) and the reference scenes are rendered again.
The fractions printed here (run with -s) are the table in DESIGN.md; the assertions only pin their order of magnitude, so
that a maintainer with a .NET runtime knows which assumption to check first and what a mismatch would look like."""
import json
import os

import numpy as np
import pytest

import fraytracer_b200 as ft
from fraytracer_b200 import scenes
from oracle import oracle as O

from conftest import bits

SWITCHES = {1: "Vector3 / float32 and Normalize multiply by the reciprocal (System.Numerics.Vectors 4.5.0 / .NET Core 2.x)",
            2: "Vector3.Dot reduces as x*x' + (y*y' + z*z')",
            4: "Array.sortInPlaceBy is stable (ties between equal LowerBounds keep their order)"}


def sensitivity(w, X, Y):
    desc = ft.SceneDescription(w.scene)
    cam = lambda lib: O.camera_look_at(scenes.CAMERA_SPEC["Position"], scenes.CAMERA_SPEC["LookAt"], scenes.CAMERA_SPEC["Up"],
                                       float(lib.fto_lens_create(scenes.CAMERA_SPEC["FieldOfView"])))
    base = O.OracleScene(desc).render(cam(O.lib()), X, Y, w.epsilon, w.length)
    rows = {}
    for bit in SWITCHES:
        path = O.assumption_library(bit)
        assert O.lib(path).fto_assumptions() == bit
        r = O.OracleScene(desc, library=path).render(cam(O.lib(path)), X, Y, w.epsilon, w.length)
        hit = base["leaf"] >= 0
        rel = np.abs(r["rgb"] - base["rgb"]).max(-1) / np.maximum(np.abs(base["rgb"]).max(-1), 1e-6)
        rows[bit] = dict(leaf_changed=float((r["leaf"] != base["leaf"]).mean()),
                         colour_bits_changed=float((bits(r["rgb"]) != bits(base["rgb"])).any(-1).mean()),
                         steps_changed=float((r["steps"] != base["steps"]).mean()),
                         shadow_changed=float((r["shadow"] != base["shadow"]).mean()),
                         rel_err_p50=float(np.percentile(rel[hit], 50)), rel_err_p99=float(np.percentile(rel[hit], 99)),
                         rel_err_max=float(rel.max()), frac_above_1e4=float((rel > 1e-4).mean()))
    return rows


@pytest.mark.parametrize("name,make,X,Y", [("config1", lambda: scenes.default_scene(size=256), 256, 256),
                                            ("config2", lambda: scenes.spheres_capsules_scene(width=320, height=180), 320, 180),
                                            ("config5", lambda: scenes.mixed_scene(width=320, height=180), 320, 180)])
def test_assumption_sensitivity(name, make, X, Y):
    rows = sensitivity(make(), X, Y)
    print(f"\n{name} {X}x{Y}: " + json.dumps(rows, indent=1))
    out = os.environ.get("FT_SENSITIVITY_JSON")
    if out:
        prev = json.load(open(out)) if os.path.exists(out) else {}
        prev[name] = rows
        json.dump(prev, open(out, "w"), indent=1)
    for bit, r in rows.items():
        # none of the assumptions moves the hit object or a shadow on more than a few silhouette pixels, nor 99 % of the
        # colours by more than 0.2 % (a flipped shadow moves single pixels by several percent)
        assert r["leaf_changed"] < 2e-3 and r["shadow_changed"] < 2e-3 and r["rel_err_p99"] < 2e-3, (bit, r)
    # arithmetic assumptions move the low bits of many colours (the forward-difference normal divides rounding noise by
    # eps / 8); the sort assumption only matters where two bounding spheres are equidistant from a cell centre
    assert rows[4]["colour_bits_changed"] <= rows[1]["colour_bits_changed"]
