import os
import sys

# the loop-back tests run several ranks on ONE GPU: every rank's streams need a hardware queue of their own, or a rank's
# barrier kernel could sit in front of the kernel it waits for (must be set before CUDA is initialised)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
