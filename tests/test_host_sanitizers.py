"""Host code under AddressSanitizer + UndefinedBehaviorSanitizer and under ThreadSanitizer ("ASan/UBSan
for host code"): tools/asan_host.cpp drives the parallel reader against the sequential parser on random FASTA /
FASTQ texts, the parallel unitig loader against the sequential one and the 2-bit packer, with several thread
counts; any report fails the test."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bgreat_b200", "csrc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
@pytest.mark.parametrize("flags", [["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"], ["-fsanitize=thread"]],
                         ids=["asan_ubsan", "tsan"])
def test_host_code_is_clean_under_sanitizers(flags, tmp_path):
    exe = str(tmp_path / "asan_host")
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread"] + flags +
                           ["-fno-omit-frame-pointer", "-I" + CSRC, "-o", exe,
                            os.path.join(ROOT, "tools", "asan_host.cpp"), os.path.join(CSRC, "bgr_host.cpp"),
                            os.path.join(CSRC, "bgr_pack.cpp")], capture_output=True, text=True)
    if build.returncode != 0 and ("asan" in build.stderr.lower() or "tsan" in build.stderr.lower()
                                  or "sanitize" in build.stderr.lower()):
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1"))
    assert run.returncode == 0, (run.stdout[-1000:], run.stderr[-3000:])
    assert run.stdout.strip().startswith("ok ")
    assert "runtime error" not in run.stderr and "AddressSanitizer" not in run.stderr
    assert "ThreadSanitizer" not in run.stderr
