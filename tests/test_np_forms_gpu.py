"""The world narrowphase has two forms (csrc/narrowphase.cu): one kernel (GJK + EPA per pair) and two kernels (GJK with
the intersecting pairs listed, EPA over the list; the default for cube-only worlds).  The library picks one per world;
NANS_NP_SPLIT=0/1 forces either.  Here the world-level parity tests are re-run in a child process under EACH forced
form, so both forms are checked on worlds with and without spheres whatever the default is."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUBSET = ("test_stages_golden or test_demo_trajectory_golden or test_drop_scene_steps_vs_oracle or "
          "test_mixed_world_with_spheres or test_batched_independent_worlds or test_empty_and_tiny_worlds or "
          "test_capacity_overflow_is_reported")


@pytest.mark.parametrize("form", ["0", "1"])
def test_world_parity_under_forced_narrowphase_form(form):
    env = dict(os.environ, NANS_NP_SPLIT=form)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu",
                        "-x", "-q", "-k", SUBSET, "-p", "no:cacheprovider"], env=env, cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, f"NANS_NP_SPLIT={form}:\n{tail}"
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
