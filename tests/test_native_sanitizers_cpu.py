"""ThreadSanitizer and AddressSanitizer / UBSan over the native code that owns threads.  bench/sanitize_native.sh builds
* tests/native/data_loader_stress.cc against csrc/runtime/data_loader.cc with each sanitizer (concurrent loaders, slow consumers,
  an early stop, the worker-exception path), and
* tests/native/planner_threads_stress.cc against the planner core under ThreadSanitizer (ILP_NUM_THREADS worker pool: the plan with
  6 threads must equal the plan with 1)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_native_threads_are_clean_under_tsan_and_asan(tmp_path):
    probe = subprocess.run(["g++", "-fsanitize=thread", "-x", "c++", "-", "-o", str(tmp_path / "probe")], input="int main(){}", text=True,
                           capture_output=True)
    if probe.returncode != 0:
        pytest.skip("this toolchain has no ThreadSanitizer runtime")
    pr = subprocess.run(["bash", os.path.join(ROOT, "bench", "sanitize_native.sh"), str(tmp_path / "logs")], capture_output=True, text=True,
                        timeout=600)
    assert pr.returncode == 0, pr.stdout[-2000:] + pr.stderr[-2000:]
    assert pr.stdout.count("sanitizer reports: 0") == 3 and pr.stdout.count("OK (0 violations)") == 2, pr.stdout
    assert "OK (identical plans)" in pr.stdout, pr.stdout
