"""CPU tests of everything ABOVE the C ABI: the host driver's C sources (argument parsing, table sizing, ingest --
sequential, multi-threaded, concurrent files --, read pairing, .ctx reader and colour filters, header arithmetic,
writer, `sort`) linked against tests/emul/abi_shim.c, where the oracle stands in for libmcxgpu.so.  The test
bodies are those of tests/test_gpu_cli.py (which runs them with the real library on a B200); only the binary differs.
This says nothing about the CUDA path."""
import pytest

import test_gpu_cli as G
import test_zy_gpu_cli_new as G2


@pytest.fixture(autouse=True)
def _use_hostcheck(monkeypatch, hostcheck):
    monkeypatch.setattr(G, "_driver", lambda: hostcheck)


# (function objects carry their own parametrize / skipif marks; the module-level gpu mark of test_gpu_cli stays there)
for _mod in (G, G2):
    for _name in dir(_mod):
        if _name.startswith("test_cli_"):
            globals()[_name.replace("test_cli_", "test_host_")] = getattr(_mod, _name)
del _name, _mod
