"""pytest configuration: markers + import paths.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol/loader checks (no GPU needed).
`-m gpu`:       parity of the CUDA path (through the C ABI) against the oracle, on a B200.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "llama.cpp-quant-gemm_b200"),
          os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
