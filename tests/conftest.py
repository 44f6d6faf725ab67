import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native artefacts exist (the driver also calls build() itself)."""
    import __graft_entry__ as g

    g.build()


def make_db(n, seed=0, key_len=(16, 32), val_len=(1, 512)):
    """Synthetic KV database shaped like utils::generate_random_kv_database (chalametpir_common/src/utils.rs:22-45)."""
    rnd = random.Random(seed)
    db = {}
    while len(db) < n:
        db[rnd.randbytes(rnd.randint(*key_len))] = rnd.randbytes(rnd.randint(*val_len))
    return db


def ptn(n):
    """RFC 9861 test pattern."""
    return bytes(i % 251 for i in range(n))
