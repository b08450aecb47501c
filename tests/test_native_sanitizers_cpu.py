"""ThreadSanitizer and AddressSanitizer / UBSan over the native code that owns threads (the input pipeline's worker ring):
bench/sanitize_native.sh builds tests/native/data_loader_stress.cc against csrc/runtime/data_loader.cc with each sanitizer and
runs concurrent loaders, slow consumers, an early stop and the worker-exception path."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_data_loader_is_clean_under_tsan_and_asan(tmp_path):
    probe = subprocess.run(["g++", "-fsanitize=thread", "-x", "c++", "-", "-o", str(tmp_path / "probe")], input="int main(){}", text=True,
                           capture_output=True)
    if probe.returncode != 0:
        pytest.skip("this toolchain has no ThreadSanitizer runtime")
    pr = subprocess.run(["bash", os.path.join(ROOT, "bench", "sanitize_native.sh"), str(tmp_path / "logs")], capture_output=True, text=True,
                        timeout=600)
    assert pr.returncode == 0, pr.stdout[-2000:] + pr.stderr[-2000:]
    assert pr.stdout.count("sanitizer reports: 0") == 2 and pr.stdout.count("OK (0 violations)") == 2, pr.stdout
