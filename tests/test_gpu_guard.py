"""GPU: the memory-safety check that stands in for compute-sanitizer (closed on the GPU pool): every
launch shape, the map builder, the matcher, the outer registration loop and a sharded solve run in a
fresh process with NLO_GUARD=1 -- 64 KB guard bands of 0xFF around every device buffer of the library,
payloads pre-filled with 0xFF (include/nlo_cuda.h, nlo_debug_guard_report) -- and no guard byte may
change; NLO_GUARD=selftest proves that an overrun of one byte is reported."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mode):
    env = dict(os.environ, NLO_GUARD=mode)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sanitize_case.py")], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SANITIZE_CASE_DONE" in out.stdout
    line = [l for l in out.stdout.splitlines() if l.startswith("GUARD ")][-1]
    return json.loads(line[6:]), out.stdout


def test_no_guard_band_is_touched_by_any_launch_shape():
    rep, log = _run("1")
    assert rep["enabled"]
    assert rep["allocations_checked"] >= 60, rep   # every problem, map, scan, ring and workspace buffer
    assert rep["allocations_live"] == 0, rep       # and everything was freed
    assert rep["corrupted_bytes"] == 0, (rep, log[-1500:])
    # a read of never-written or out-of-bounds memory would have produced NaN sums: the solves ran
    assert "nan" not in log.lower()


def test_guard_selftest_reports_the_deliberate_overrun():
    rep, _ = _run("selftest")
    assert rep["enabled"] and rep["corrupted_bytes"] == 1, rep


def test_guard_is_off_by_default(nlo):
    if os.environ.get("NLO_GUARD", "0") not in ("", "0"):
        pytest.skip("suite itself runs under NLO_GUARD")
    rep = nlo.guard_report()
    assert not rep["enabled"] and rep["allocations_checked"] == 0
