import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests" / "golden", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        # a deadlocked kernel must fail ONE test quickly, not hang the whole session: pytest-timeout in
        # "thread" mode kills the process (the only way out of a stuck cudaDeviceSynchronize)
        try:
            import os
            import pytest_timeout  # noqa: F401
            limit = int(os.environ.get("OU_GPU_TEST_TIMEOUT", "240"))     # tools/sanitize.sh raises it
            for item in items:
                if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
                    item.add_marker(pytest.mark.timeout(limit, method="thread"))
        except ImportError:
            pass
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- measured parity errors: tests record into ``parity_log``; written once per session to
# gpurun_out/parity_r2.json (the only directory that travels back from the GPU box)
_PARITY = {}


@pytest.fixture
def parity_log():
    return _PARITY


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        meta = {}
        try:
            import torch
            from open_universe_b200.engine import lib
            meta = {"gpu": torch.cuda.get_device_name(0), "storage_dtype": lib.act_name(),
                    "torch": torch.__version__}
        except Exception:
            pass
        (out / "parity_r2.json").write_text(json.dumps({"meta": meta, "results": _PARITY}, indent=1,
                                                       sort_keys=True))
    except OSError:
        pass
