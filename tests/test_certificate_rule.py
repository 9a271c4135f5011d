"""The ungapped-diagonal certificate of k_diag_certify (pecaller_b200/csrc/sw_int16.cuh) against the full three-state
integer DP, on the CPU: small adversarial cases (2-4 letter alphabets, planted reads, tandem-repeat backgrounds).
Whenever the rule fires, score / maxi / maxk, uniqueness of the last-column maximum and the pure-diagonal traceback
must be what the DP (smith_waterman_align, pemapper.c:1694-1748, in units of 1/36) gives."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_certificate_equals_dp_on_random_small_cases():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "certify_bruteforce.py"), "4000", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok:"), r.stdout
