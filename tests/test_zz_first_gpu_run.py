"""Runs the GPU cases of tests/first_gpu_run_cases.py (rows N1-N4: device ICs, snapshots / restart, PGD, softening, the other
windows, the C user programs), one pytest process per case.  All twelve were green on a B200 at the end of round 1
(GPUTEST_r01.json: 12 xpassed); since round 2 they are ordinary tests -- a failure fails the suite.  The cases from
test_wrap_and_summary_match_reference on were written in the last session of round 2, after the GPU time was spent: they have run
on the emulated library only (tests/test_cpu_full_emulation.py) and sit at the end of the list on purpose."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    "test_device_gadget_ic_matches_reference",
    "test_device_ic_chain_matches_reference",
    "test_libfastpm_user_program_runs_and_matches_fixture",
    "test_cli_run_loop_program_matches_reference",
    "test_three_component_readout_equals_three_readouts",
    "test_fused_readout_option_gives_the_same_run",
    "test_pgd_correction_matches_reference",
    "test_snapshot_files_and_restart_match_reference",
    "test_snapshots_during_evolve_match_reference",
    "test_force_softening_matches_reference",
    "test_non_cic_painter_matches_reference",
    "test_single_mode_transfers_match_reference",
    "test_shifted_ics_match_reference",
    "test_passive_handler_keeps_the_fused_update",
    "test_store_fill_q_and_rand_columns_match_reference",
    "test_wrap_and_summary_match_reference",
    "test_store_subsample_matches_reference",
    "test_store_copy_take_extend",
    "test_permute_by_dense_id_kernels",
    "test_sorted_snapshot_of_a_shuffled_store",
    "test_command_line_particle_fraction_snapshot",
]


def test_case_list_is_complete():
    """every test of first_gpu_run_cases.py is listed above (CPU check)"""
    import re
    src = open(os.path.join(ROOT, "tests", "first_gpu_run_cases.py")).read()
    assert sorted(re.findall(r"^def (test_\w+)\(", src, flags=re.M)) == sorted(CASES)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_first_gpu_run(case):
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "first_gpu_run_cases.py") + "::" + case],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:]


@pytest.mark.gpu
def test_first_two_gpu_run_of_the_extras():
    """PGD correction on two slabs and a snapshot written by both ranks from their device columns (tests/mp_worker.py: extras)."""
    sys.path.insert(0, ROOT)
    from fastpm_b200 import _lib
    if _lib.load().fpm_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29651", os.path.join(ROOT, "tests", "mp_worker.py"), "gpu"]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, MP_EXTRAS="1"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert "MP_EXTRAS_OK" in r.stdout
    assert "MP_WINDOWS_OK" in r.stdout
