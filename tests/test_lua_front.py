"""The Lua parameter surface (BASELINE.json north_star): fastpm_b200/lua_front/_build/fastpm_b200_run evaluates
FastPM parameter files with the reference's own Lua runtime (compiled in place by fastpm_b200/lua_front/Makefile) and drives the
force step / integrator on libfastpm_b200.so.

CPU: the parsed configuration of parameter files written for this repository (tests/lua/), the reference schema's validation, the
refusal of options outside the path.  GPU: the run of the reference's tests/lightcone.lua configuration (minus its light-cone and
FOF outputs) must print the lines of the reference's own golden log, tests/run-test-lightcone.check."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "fastpm_b200", "lua_front", "_build", "fastpm_b200_run")


@pytest.fixture(scope="module")
def run_bin():
    if not os.path.exists(RUN):
        pytest.skip("fastpm_b200_run is not built (make -C fastpm_b200/lua_front where /root/reference exists)")
    return RUN


def _run(run_bin, args, cwd):
    return subprocess.run([run_bin] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)


def test_parameter_file_is_evaluated_by_the_reference_schema(run_bin, tmp_path):
    r = _run(run_bin, ["--dump-config", os.path.join(ROOT, "tests", "lua", "c0_standard.lua"), "cola"], str(tmp_path))
    assert r.returncode == 0, r.stdout
    conf = r.stdout
    assert re.search(r"\bnc = 64,", conf) and re.search(r"boxsize = 128", conf) and 'force_mode = "cola"' in conf
    # linspace(0.1, 1, 5) evaluated by the runtime; output_redshifts {1, 0} -> aout {0.5, 1} (lua-runtime-fastpm.lua:18-33)
    ts = re.search(r"time_step = \{(.*?)\}", conf, flags=re.S).group(1)
    np.testing.assert_allclose([float(v) for v in ts.replace("\n", " ").split(",") if v.strip()], np.linspace(0.1, 1, 5), rtol=1e-15)
    aout = re.search(r"aout = \{(.*?)\}", conf, flags=re.S).group(1)
    np.testing.assert_allclose(sorted(float(v) for v in aout.replace("\n", " ").split(",") if v.strip()), [0.5, 1.0])
    # defaults come from the schema: kernel 1_4, CIC painter, no softening
    assert 'kernel_type = "1_4"' in conf and 'painter_type = "cic"' in conf and 'force_softening_type = "none"' in conf
    # the variable mesh table survives as a 2-d array
    assert re.search(r"pm_nc_factor = \{\s*\{", conf)
    # argument keywords reach the file
    r2 = _run(run_bin, ["--dump-config", os.path.join(ROOT, "tests", "lua", "c0_standard.lua"), "pm", "lanczos2"], str(tmp_path))
    assert 'painter_type = "lanczos"' in r2.stdout and "painter_support = 4" in r2.stdout and 'force_mode = "pm"' in r2.stdout


def test_schema_rejects_bad_files(run_bin, tmp_path):
    """the reference schema's own checks (lua-runtime-config.lua:95-135): required keys, enumerated choices; errors of the file itself"""
    head = 'nc = 16\nboxsize = 10\ntime_step = {1.0}\noutput_redshifts = {0}\nOmega_m = 0.3\nh = 0.7\nread_powerspectrum = "powerspec.txt"\n'
    bad = tmp_path / "bad.lua"
    bad.write_text(head + "np_alloc_factor = 2.0\n")
    r = _run(run_bin, ["--dump-config", str(bad)], str(tmp_path))
    assert r.returncode != 0 and "`pm_nc_factor` is required but undefined" in r.stdout, r.stdout
    bad.write_text(head + 'np_alloc_factor = 2.0\npm_nc_factor = 2\nforce_mode = "warp"\n')
    r = _run(run_bin, ["--dump-config", str(bad)], str(tmp_path))
    assert r.returncode != 0 and "value `warp` of key `force_mode` is not one of" in r.stdout, r.stdout
    bad.write_text(head + 'np_alloc_factor = 2.0\npm_nc_factor = 2\nforce_mode = "pm"\n')
    r = _run(run_bin, ["--dump-config", str(bad)], str(tmp_path))
    assert r.returncode == 0 and 'force_mode = "pm"' in r.stdout, r.stdout
    r = _run(run_bin, ["--dump-config", os.path.join(ROOT, "tests", "lua", "c0_standard.lua"), "nonsense"], str(tmp_path))
    assert r.returncode != 0 and "wrong arg" in r.stdout, r.stdout


def test_options_outside_the_path_are_refused_before_any_device_is_touched(run_bin, tmp_path):
    r = _run(run_bin, [os.path.join(ROOT, "tests", "lua", "c0_standard.lua"), "cola", "lightcone"], str(tmp_path))
    assert r.returncode != 0
    assert "lc_write_usmesh" in r.stdout and "outside the force-step path" in r.stdout, r.stdout
    # a sorted sub-sampled snapshot from several ranks needs the distributed sort of non-dense ids: refused up front, not at the
    # first snapshot
    r = _run(run_bin, ["-n", "2", os.path.join(ROOT, "tests", "lua", "small_nc16.lua"), "8", "3", "0.3"], str(tmp_path))
    assert r.returncode != 0 and "sort_snapshot = false" in r.stdout, r.stdout


GOLDEN = [                                     # /root/reference/tests/run-test-lightcone.check:1-5,8,28,42,56,64,72,80,88
    "Found 1769 pairs of values in input spectrum table",
    "Variance of input white noise is 0.99999619, expectation is 0.99999619",
    "dx1  : 5.36177 5.36177 5.36177 5.36177",
    "dx2  : 0.455678 0.44748 0.453293 0.45215",
]
GOLDEN_PLIN = [(0.1, 17305.5, 6.20821), (0.228571, 17200.9, 2.54189), (0.357143, 17110.0, None), (0.485714, 17064.7, None),
               (0.614286, 17043.4, None), (0.742857, 17028.1, None), (0.871429, 17014.2, None), (1.0, 17002.2, 0.682708)]


@pytest.mark.gpu
def test_golden_log_of_the_reference_run(run_bin, tmp_path):
    shutil.copy(os.path.join(ROOT, "tests", "golden", "powerspec.txt"), str(tmp_path / "powerspec.txt"))
    r = _run(run_bin, [os.path.join(ROOT, "tests", "lua", "golden_nc64.lua")], str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:]
    log = r.stdout
    for line in GOLDEN:
        assert line in log, "missing golden line %r\n%s" % (line, log[-3000:])
    # "Input power spectrum sigma8 0.815897" (tests/run-test-nbodykit.sh:14) and the Sigma8 of every step are adaptive integrals the
    # reference asks GSL to carry to a RELATIVE ACCURACY OF 1e-4 only (powerspectrum.c:250-279): their last printed digits belong to
    # QUADPACK's bisection sequence, not to the physics (SURVEY.md section 8c calls them soft goldens).  Ours must agree to that accuracy.
    s8 = float(re.search(r"Input power spectrum sigma8 ([0-9.]+)", log).group(1))
    assert abs(s8 / 0.815897 - 1) < 1e-4, s8
    found = re.findall(r"D\^2\(([0-9.e+-]+), 1\.0\) P\(k<([0-9.e+-]+)\) = ([0-9.e+-]+) Sigma8 = ([0-9.e+-]+)", log)
    assert len(found) == len(GOLDEN_PLIN), found
    for (a, kmax, plin, s8), (ga, gplin, gs8) in zip(found, GOLDEN_PLIN):
        assert a == "%g" % ga and kmax == "0.0490625", (a, kmax)
        assert plin == "%g" % gplin, (a, plin, gplin)                       # the reference's six printed digits
        if gs8 is not None:                                                  # soft golden, see above (sigma^2 to 1e-4 on both sides)
            assert abs(float(s8) / gs8 - 1) < 3e-4, (a, s8, gs8)
    # the KDK state machine prints the reference's transition lines, the run ends with a snapshot at a = 1 and 8 spectra on disk
    assert "==== -> 001 [000 000 000]" in log and "==== -> 005 [002 001 002]" in log
    assert "written at z = 0.0000 a = 1.0000" in log
    assert os.path.isdir(tmp_path / "golden_nc64" / "fastpm_1.0000" / "1" / "Position")
    assert len([f for f in os.listdir(tmp_path / "golden_nc64") if f.startswith("powerspec_") and f.endswith(".txt")]) == 9   # 8 + linear
