"""mccortex_b200 -- B200-native `mccortex build` hot path.

The product is native: mccortex_b200/lib/libmcxgpu.so (CUDA kernels for sm_100a
behind the C ABI in include/mcx_gpu.h) and mccortex_b200/bin/mccortex-b200 (C host
driver with the reference's `build` command line).  This Python package is only
the thin ctypes mirror of that ABI that tests/ and bench.py drive; it holds no
algorithm and has no CPU fallback: importing `binding` without the built .so, or
using it without a CUDA device, raises.
"""
from .binding import (  # noqa: F401
    Graph, LoadStats, McxError, lib, lib_path, driver_path, device_count, build_native,
    host_alloc, host_free, key_owner, sort_records, device_alloc, device_free, ipc_export, ipc_open, ipc_close,
    MCX_LAYOUT_LINES, MCX_LAYOUT_OFFSETS, MCX_MEM_HOST, MCX_MEM_DEVICE,
    MCX_GRAPH_INTERSECT, MCX_GRAPH_READSTRT, MCX_LOAD_MUST_EXIST, MCX_LOAD_INTO_ISEC, MCX_LOAD_MASK_ISEC,
)
