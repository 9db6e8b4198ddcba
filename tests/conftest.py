import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    """Asks the product's own library (it does not depend on torch): a box with a GPU but without a
    CUDA-enabled torch must still run the parity tests."""
    try:
        from icikendalltau_b200 import _lib
        return _lib.load().icikt_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    if os.environ.get("ICIKT_REQUIRE_GPU") == "1" and any("gpu" in item.keywords for item in items):
        # tools/gpu_round.sh sets this on the GPU box: a silent skip there would read as green
        raise pytest.UsageError("ICIKT_REQUIRE_GPU=1 but libicikt_b200 reports no usable CUDA device")
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Worst p-value error per parity test (|dp| / (max(1, z^2) p), bar 1e-12) for profiles/."""
    try:
        mod = sys.modules.get("test_gpu_parity") or sys.modules.get("tests.test_gpu_parity")
        worst = getattr(mod, "PVALUE_WORST", None)
        if worst:
            import json
            out = os.path.join(ROOT, "gpurun_out")
            os.makedirs(out, exist_ok=True)
            with open(os.path.join(out, "pvalue_worst.json"), "w") as f:
                json.dump({"bar": 1e-12, "overall": max(worst.values()),
                           "per_test": dict(sorted(worst.items(), key=lambda kv: -kv[1])[:40])}, f, indent=1)
    except Exception:
        pass


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
