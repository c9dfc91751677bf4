import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    from oracle import refload
    has_ref = refload.available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not present"))


@pytest.fixture(autouse=True)
def _tag_parity_records(request):
    from tests import util
    util.CURRENT_TEST[0] = request.node.nodeid
    yield


def pytest_sessionfinish(session, exitstatus):
    """$DMH_PARITY_REPORT=<path>: dump every measured parity number of the session (tests/util.py REPORT)."""
    path = os.environ.get("DMH_PARITY_REPORT")
    if not path:
        return
    import json
    from tests import util
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump({"exitstatus": int(exitstatus), "records": util.REPORT}, f, indent=0)
