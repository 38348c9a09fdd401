import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)


def golden_mask(goldens, name):
    """int32 mask[NY, NX] (x fastest) exactly as Grid hands it to the partitioner."""
    inp = goldens["inputs"][name]
    vals = np.asarray(inp["mask"], dtype=np.int32)
    # Grid.cpp:119-126 reads the variable in file order and Grid.cpp:176-186 indexes it
    # x-fastest, whatever `order` says (quirk Q7): a raw reinterpretation.
    return vals.reshape(inp["ny"], inp["nx"])


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc
