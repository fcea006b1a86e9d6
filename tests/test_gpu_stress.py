"""Race evidence for the hand-rolled mbarrier / TMEM / TMA protocols (compute-sanitizer is closed on this GPU pool):
tools/stress_launches.py fires > 1000 back-to-back launches of mixed-shape tcgen05 GEMMs, whole train steps and whole
generation calls, hashing every result.  Every case must hash identically across its repetitions, and the run with
programmatic dependent launch switched off (B200_NO_PDL=1: kernels strictly serialised) must produce the same hashes
as the default run (kernel prologues overlapping the predecessor's tail)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_launches.py"), "--reps", "6"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, (out.stdout[-500:], out.stderr[-1500:])
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def test_back_to_back_launches_are_deterministic_with_and_without_pdl(cuda_dev):
    a = _run({})
    env = {"B200_NO_PDL": "1"}
    b = _run(env)
    assert a["mismatches"] == 0 and b["mismatches"] == 0
    assert a["pdl"] is True and b["pdl"] is False
    assert a["launches"] >= 1000 and b["launches"] >= 1000, (a, b)
    assert a["hash"] == b["hash"], (a, b)
