import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "latency_kernel: decode through the warp-per-chunk kernel whatever the job size (SEA_B200_DEC_LATENCY=1)")
    config.addinivalue_line("markers", "auto_route: leave the library's own choice between the small-job and the throughput decode kernels")


@pytest.fixture(autouse=True)
def _pin_decode_route(request, monkeypatch):
    """The library sends small decode jobs (which is what tests are) to the warp-per-chunk kernel of decode_latency.cu.  The
    parity suite was written against the throughput kernels (unrolled / VBR / multichannel / staged / generic) on small batches, so
    by default the small-job route is switched off; tests marked `latency_kernel` force it on, `auto_route` leaves the choice."""
    if "latency_kernel" in request.keywords:
        monkeypatch.setenv("SEA_B200_DEC_LATENCY", "1")
    elif "auto_route" in request.keywords:
        monkeypatch.delenv("SEA_B200_DEC_LATENCY", raising=False)
    else:
        monkeypatch.setenv("SEA_B200_DEC_LATENCY", "0")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import sea_oracle

    sea_oracle.build()
    return sea_oracle


@pytest.fixture(scope="session")
def ctx():
    import sea_codec_b200 as S

    c = S.Context(0)
    yield c
    c.close()
