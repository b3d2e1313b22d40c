import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


_CACHE = {}


@pytest.fixture(scope="session")
def synthetic_checkpoints():
    """Factory: bits -> (generator state_dict, tokenizer state_dict), cached per session."""
    from maskbit_b200.weights import synthetic_conv_vq_state_dict, synthetic_lfq_bert_state_dict

    def get(bits=12):
        if bits not in _CACHE:
            _CACHE[bits] = (synthetic_lfq_bert_state_dict(seed=0, codebook_size=2 ** bits),
                            synthetic_conv_vq_state_dict(seed=0, token_size=bits))
        return _CACHE[bits]

    return get
