import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lorem() -> bytes:
    with open(os.path.join(GOLDEN, "lorem_ipsum.txt"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def lorem_encoded() -> bytes:
    with open(os.path.join(GOLDEN, "lorem_ipsum_encoded.bin"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def sunflower_pixels() -> bytes:
    # 54-byte BMP header, then 200 rows of 544 bytes (181 px * 3 + 1 pad), SURVEY.md 2
    with open(os.path.join(GOLDEN, "sunflower.bmp"), "rb") as f:
        return f.read()[54:]
