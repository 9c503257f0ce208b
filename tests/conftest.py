"""pytest configuration: `gpu` marker + shared paths/fixtures."""
import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) when selected on a box without CUDA;
    they are simply deselected by `-m "not gpu"` on CPU."""
    return


@pytest.fixture(scope="session")
def gold_dir():
    return GOLD


def load_json(name):
    with open(os.path.join(GOLD, name)) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def demo_records():
    from oracle import topsicle_oracle as orc
    return list(orc.read_fastx(os.path.join(GOLD, "demo.fastq.gz")))


@pytest.fixture(scope="session")
def edge_records():
    from oracle import topsicle_oracle as orc
    return list(orc.read_fastx(os.path.join(GOLD, "edge.fastq")))
