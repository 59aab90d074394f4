"""Device counterpart of tests/test_golden_features.py: the CUDA path must reproduce the committed fixtures of the
user-event and byzantine-injector scenarios (tests/golden/features.json)."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import run_case  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "features.json")))


@pytest.mark.parametrize("case", GOLDEN, ids=[c["scenario"] + "-" + "-".join(f"{k}{v}" for k, v in c["args"].items()) for c in GOLDEN])
def test_cuda_reproduces_golden(case):
    from serf_b200 import GossipSim
    got = run_case(case["scenario"], case["args"], lambda n, s, **kw: GossipSim(n, s, **kw))
    assert got == {k: v for k, v in case.items() if k not in ("scenario", "args")}
