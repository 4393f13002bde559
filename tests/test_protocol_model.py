"""CPU: host-thread models of the kernels' polled exchange protocols (tests/protocol_model.cc) -- the
LL all-gather with parity double buffering (PeerAllReduce / ReduceAndExchange / GatherLL) and the
streaming kernel's counter + leader-publishes-state loop, and the TMA / mbarrier tile ring with its
cross-iteration prefetch, early-exit drain and resident mode -- under random scheduling jitter, across
the 32-bit wrap of the tag, and under ThreadSanitizer; with negative controls (no double buffer, no
drain) that must fail.  racecheck being unavailable on the GPU pool, this is where the protocols' logic is
exercised on its own; the kernels themselves are covered by the bitwise-repeatability GPU tests."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "protocol_model.cc")


def _build(tmp_path, name, extra):
    exe = str(tmp_path / name)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-pthread"] + extra + [SRC, "-o", exe], check=True)
    return exe


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    return _build(tmp_path_factory.mktemp("protocol"), "protocol_model", ["-O2"])


@pytest.mark.parametrize("participants,iterations,seed", [(2, 4000, 1), (8, 3000, 2), (33, 400, 3)])
def test_protocols_hold_under_jitter(model, participants, iterations, seed):
    out = subprocess.run([model, str(participants), str(iterations), str(seed)], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr
    assert "PROTOCOL_MODEL_OK" in out.stdout
    assert "TILE_RING_OK runs=336" in out.stdout   # stages 2-4 x tiles 0-23 x iterations x early exit x mode


def test_negative_control_without_the_prefetch_drain_fails(model):
    env = dict(os.environ, PROTOCOL_MODEL_BREAK="2")
    out = subprocess.run([model, "2", "10", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 1
    assert "bulk copies not waited for" in out.stderr


def test_negative_control_producer_ignoring_the_empty_barrier_fails(model):
    env = dict(os.environ, PROTOCOL_MODEL_BREAK="3")
    out = subprocess.run([model, "2", "10", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 1
    assert "model 3" in out.stderr


def test_negative_control_without_the_double_buffer_fails(model):
    env = dict(os.environ, PROTOCOL_MODEL_BREAK="1")
    out = subprocess.run([model, "8", "3000", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 1
    assert "model 1" in out.stderr


def test_protocols_under_thread_sanitizer(tmp_path):
    try:
        exe = _build(tmp_path, "protocol_model_tsan", ["-O1", "-g", "-fsanitize=thread"])
    except subprocess.CalledProcessError:
        pytest.skip("libtsan not installed")
    probe = subprocess.run([exe, "1", "1", "1"], capture_output=True, text=True, timeout=120)
    if probe.returncode != 0 and "ThreadSanitizer" in probe.stderr and "unexpected memory mapping" in probe.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this sandbox")
    out = subprocess.run([exe, "4", "300", "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "WARNING: ThreadSanitizer" not in out.stderr
    assert "PROTOCOL_MODEL_OK" in out.stdout
