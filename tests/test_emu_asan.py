"""The kernel bodies under AddressSanitizer.  compute-sanitizer is closed on the GPU pool, so memory safety of
the kernels is checked where it can be: the CPU emulation harness (tests/emu) is rebuilt with
-fsanitize=address and the emulation tests of the stages that index global buffers (grid, row spectra,
slabs, CSR, spectra, chains) are re-run against it in a subprocess.  An out-of-bounds access aborts the run.
(This found the one-slot over-read of the Bluestein filter table in round 1.)"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

EMU_DIR = os.path.join(ROOT, "tests", "emu")


def test_kernel_bodies_under_address_sanitizer():
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], stdout=subprocess.PIPE, text=True).stdout.strip()
    if not libasan or not os.path.exists(libasan):
        pytest.skip("libasan not available")
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libcurie_emu_asan.so"])
    env = dict(os.environ, CB_EMU_LIB=os.path.join(EMU_DIR, "libcurie_emu_asan.so"), LD_PRELOAD=libasan,
               ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_emu_kernels.py"), "-x", "-q",
                        "-p", "no:cacheprovider", "-k",
                        "spectrum or shared_row or detrend or fit_body or fit_rows or mh_body or radon or tanaka"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=1500)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stdout, r.stdout[-4000:]
