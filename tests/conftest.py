import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    ans = json.load(open(os.path.join(GOLDEN, "fixture_answers.json")))
    ans["index_path"] = os.path.join(GOLDEN, "faiss_index.bin")
    ans["mapping_path"] = os.path.join(GOLDEN, "faiss_index.bin.mapping")
    ans["perturbed"] = np.asarray(ans["perturbed_queries"], np.float32)
    return ans


@pytest.fixture(scope="session")
def synth_known():
    return json.load(open(os.path.join(GOLDEN, "synth_known.json")))
