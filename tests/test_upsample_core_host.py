"""Exhaustive host-side check of the packed arithmetic the CUDA fancy upsampler is
built from (go_webp_b200/csrc/b2_upsample_core.cuh): the very header the kernel
includes is compiled for the host by nvcc and compared with the reference's scalar
formulas (pkg/color/yuv/conversion.go:28-48, pkg/libwebp/dsp/upsampling.c.go:69-86)
over all 2^24 (y, u, v) triples and a 64^4 lattice of interpolation inputs.
No GPU is needed; the GPU parity tests then confirm the hardware's VIADDMNMX / PRMT
agree with the host emulation."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_core_matches_reference_exhaustively(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "core_check")
    src = os.path.join(ROOT, "tests", "host", "upsample_core_check.cu")
    subprocess.check_call([nvcc, "-O2", "-Wno-deprecated-gpu-targets", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "OK (0 mismatches)" in out.stdout
