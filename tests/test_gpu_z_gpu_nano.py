"""The reference's training_configs/gpu/nano.yaml at its REAL dimensions (PEER tail with 65 536 experts, 1600 -> 1280
bridging Linear, 36-layer / 1280-wide / 20-head decoder, SNRAdam; 1.14 G parameters): scripts/gpu_nano_large_check.py builds it,
checks the fp32 forward against the CPU oracle (1e-4), the greedy ids against the oracle's cache-less loop (bit-exact), runs four
bf16 SNRAdam steps on the YAML's parameter groups (the loss must fall) and a bf16 generate."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_nano_yaml_real_dimensions():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_nano_large_check.py")], capture_output=True, text=True,
                       timeout=1500, cwd=ROOT)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().endswith("ok")
