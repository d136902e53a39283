"""NB2_FP=fast: the FMA-contracted twin library (same sources, nvcc's default contraction, libm-grade inverse trig) against the
CPU oracle at the north-star bar - contact counts per substep identical, body_q within 1e-5 relative after 100 substeps
(BASELINE.json north_star).  The strict library (default) is held to bit equality by every other GPU test; this one states how
far the fast mode is allowed to be and checks that it really is a different arithmetic (otherwise the switch tests nothing)."""

import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-5  # relative, body_q after 100 substeps (north_star)


def test_fast_fp_library_meets_the_north_star_tolerance(cuda_lib):
    env = dict(os.environ, NB2_FP="fast")
    env.pop("NB2_LIB", None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fast_fp_worker.py")], env=env, capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("FAST_FP_RESULT ")][-1]
    res = json.loads(line[len("FAST_FP_RESULT "):])
    assert res["lib"] == "libnewton_b200_fast.so"
    for name in ("quadruped_xpbd", "quadruped_featherstone", "box_stacks_xpbd"):
        r = res[name]
        assert r["counts_equal"], f"{name}: contact counts per substep differ from the oracle"
        assert r["body_q_rel"] < TOL, f"{name}: body_q relative error {r['body_q_rel']:.3e}"
        # twists: absolute error against the scene's velocity scale (many components rest at exactly zero)
        assert r["body_qd_abs"] < 1e-3 * max(1.0, r["body_qd_scale"]), f"{name}: body_qd error {r['body_qd_abs']:.3e}"
    assert not all(res[n]["bit_equal"] for n in ("quadruped_xpbd", "quadruped_featherstone", "box_stacks_xpbd")), \
        "the fast library reproduced the oracle bit for bit everywhere: is it really the contracted build?"
