"""The reference's OWN unit tests, run against refiners_b200.

Where /root/reference is mounted (the build container), the weight-free unit tests of finegrain-ai/refiners for the
fluxion API and the adapters on the hot path are executed unmodified, with ``import refiners...`` resolved to this
package (tests/_refrun/alias_plugin.py).  That the same test files pass on both code bases is the drop-in claim for the
host side of the boundary: constructor signatures, tree editing, context plumbing, adapter
inject / eject, LoRA algebra, error text, ``repr``, and two identical UNet forwards being bit-identical.

Skipped where the reference is absent (the GPU box)."""

import os
import subprocess
import sys
from pathlib import Path

import pytest

REFERENCE = Path("/root/reference")
ROOT = Path(__file__).resolve().parent.parent

# (reference test file, tests expected to pass)
FILES = [
    ("tests/adapters/test_adapter.py", 4),
    ("tests/adapters/test_lora.py", 13),
    ("tests/adapters/test_adapter_context.py", 1),
    ("tests/adapters/test_range_adapter.py", 1),
    ("tests/adapters/test_control_lora.py", 2),
    ("tests/fluxion/layers/test_chain.py", 20),
    ("tests/fluxion/layers/test_basics.py", 9),
    ("tests/fluxion/layers/test_converter.py", 2),   # + 2 CUDA-only cases the reference itself skips on a CPU host
    ("tests/fluxion/test_module.py", 2),
    ("tests/fluxion/test_utils.py", 11),               # + 1 half-precision blur case the reference skips on a CPU host
    ("tests/adapters/test_self_attention_guidance.py", 2),
    ("tests/adapters/test_style_aligned_adapter.py", 6),
    ("tests/adapters/test_t2i_adapter.py", 2),
    ("tests/adapters/test_ip_adapter.py", 4),
    ("tests/foundationals/latent_diffusion/test_sd15_unet.py", 1),
    ("tests/foundationals/segment_anything/test_utils.py", 5),
]
SLOW = [("tests/adapters/test_controlnet.py", 8)]  # 80 s on 8 cores: RB200_REFERENCE_TESTS=all


@pytest.mark.skipif(not (REFERENCE / "tests").exists(), reason="/root/reference is not mounted here")
def test_reference_unit_tests_pass_on_the_mirror():
    files = FILES + (SLOW if os.environ.get("RB200_REFERENCE_TESTS") == "all" else [])
    env = dict(os.environ, PYTHONPATH=f"{ROOT}:{ROOT / 'tests' / '_refrun'}")
    parallel: list[str] = []
    try:  # four workers of two threads each: 230 s -> 95 s on the 8-core build container
        import xdist  # noqa: F401

        parallel = ["-n", "4"]
        env["OMP_NUM_THREADS"] = "2"
    except ImportError:
        pass
    cmd = [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "--noconftest", "-p", "alias_plugin", "--import-mode=importlib",
           "-q", *parallel, *[f for f, _ in files]]
    res = subprocess.run(cmd, cwd=REFERENCE, env=env, capture_output=True, text=True, timeout=1800)
    tail = res.stdout[-3000:] + res.stderr[-2000:]
    assert res.returncode == 0, tail
    expected = sum(n for _, n in files)
    assert f"{expected} passed" in res.stdout, f"expected {expected} reference tests to pass:\n{tail}"
