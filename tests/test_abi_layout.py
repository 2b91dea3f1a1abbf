"""Drop-in boundary: the public structs of include/fastpm_b200_api.h have the reference's layout (api/fastpm/*.h).

The same probe (tests/abi/layout_probe.c) is compiled against the reference's own headers and against ours; sizes, field
offsets, enum values and event names must be identical.  Needs /root/reference (skipped on the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _build_and_run(tmp, name, flags):
    exe = os.path.join(tmp, name)
    cmd = ["gcc", "-std=gnu99", "-w", "-o", exe, os.path.join(ROOT, "tests", "abi", "layout_probe.c")] + flags
    env = dict(os.environ)
    env.pop("CC", None)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0, r.stdout
    return subprocess.run([exe], stdout=subprocess.PIPE, text=True, check=True).stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "api", "fastpm", "libfastpm.h")), reason="reference tree not present")
def test_public_struct_layouts_match_reference(tmp_path):
    ref = _build_and_run(str(tmp_path), "probe_ref", ["-DUSE_REFERENCE", "-DFASTPM_FFT_PRECISION=32", "-I" + os.path.join(ROOT, "oracle", "shims", "include"),
                                                      "-I" + os.path.join(REF, "api")])
    ours = _build_and_run(str(tmp_path), "probe_ours", ["-I" + os.path.join(ROOT, "include")])
    ref_lines, our_lines = ref.strip().splitlines(), ours.strip().splitlines()
    assert len(ref_lines) == len(our_lines) and len(ref_lines) > 100
    diff = [(a, b) for a, b in zip(ref_lines, our_lines) if a != b]
    assert not diff, diff
