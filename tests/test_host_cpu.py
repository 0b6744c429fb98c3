"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol
include/nans_b200.h declares (no compute calls without a GPU), fails loudly without a device,
and the sinf/cosf emulation the vertex kernel runs equals the running libm bit for bit."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    from nans_projekat_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "nans_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nans_[a-z_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = built.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/nans_b200.h but not exported"
    assert set(built.EXPORTS) == declared


def test_no_cpu_fallback(built):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nans_projekat_b200 import scenes, world
    with pytest.raises(built.NansError):
        world.World(scenes.demo_scene())
    with pytest.raises(built.NansError):
        p = scenes.narrowphase_pairs(8)
        world.check_collision(p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"])


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "nans_projekat_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in (r"^\s*(from|import)\s+oracle", r"libnans_oracle", r"nans_oracle\.h",
                            r"oracle/", r"ref_harness"):
                    assert not re.search(pat, src, flags=re.M), f"{f} reaches into oracle/ ({pat})"


def test_sincos_emulation_matches_libm(tmp_path):
    exe = tmp_path / "sincos_check"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", str(exe),
                           os.path.join(ROOT, "tests", "sincos_check.cpp"), "-lm"])
    out = subprocess.check_output([str(exe), "4000000", "1031"]).decode().split()
    bad_s, bad_c, total = map(int, out)
    assert total > 6_000_000
    assert bad_s == 0 and bad_c == 0, f"glibc sinf/cosf emulation differs: {bad_s} sin, {bad_c} cos of {total}"


def test_scene_builders():
    from nans_projekat_b200 import scenes
    s = scenes.demo_scene()
    assert (s.n_cubes, s.n_spheres, s.n_statics) == (4, 1, 1)
    assert np.allclose(s.moi[:4], 1 / 6) and np.isclose(s.moi[4], 0.05)
    assert np.isclose(s.st_moi[0], (1e5 / 12) * 2e4, rtol=1e-6)
    d = scenes.cube_drop(n=1000, dims=(10, 10, 10))
    assert d.n_statics == 5 and d.pos.shape == (1000, 3) and d.pos[:, 1].min() > 0.8
    p = scenes.cube_pile(n_side=10)
    assert p.n_cubes == 1000
    b = scenes.batched_worlds(n_worlds=4, cubes_per=6, spheres_per=2)
    assert b.world_id.max() == 3 and b.n_cubes == 24 and b.n_spheres == 8


def test_headers_are_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: both headers must compile as C11 (no C++/torch types)."""
    for h in ("nans_b200.h", "nans_plugin.h"):
        src = tmp_path / (h + ".c")
        src.write_text(f'#include "{os.path.join(ROOT, "include", h)}"\nint main(void) {{ return 0; }}\n')
        subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", str(src)])


def test_plugin_exports_the_reference_entry_and_fails_loudly_without_a_device(built):
    """The drop-in plugin exports the reference's ONE entry point (code/nans.h:396-397) under its C name, and the
    headless host + plugin abort with the CUDA layer's error text when no device is usable: no CPU fallback."""
    host_dir = os.path.join(ROOT, "nans_projekat_b200", "host")
    plugin = os.path.join(host_dir, "nans.so")
    assert os.path.exists(plugin), "build() did not produce the plugin"
    syms = subprocess.check_output(["nm", "-D", "--defined-only", plugin]).decode()
    exported = {l.split()[-1] for l in syms.splitlines() if " T " in l}
    assert "SimUpdateAndRender" in exported and "NansPluginPeek" in exported
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([os.path.join(host_dir, "nans"), "--frames", "2", "--dt", "0.016666668"], cwd=host_dir,
                       capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert "nans_world_create" in (r.stderr + r.stdout)
