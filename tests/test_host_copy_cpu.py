"""gofpid_host_copy2d: the staging threads that move pageable frames into page-locked memory inside the host
entry points (HostCopier, csrc/pipeline.cu), exercised on the CPU: work sharing between the caller and the
workers, strips reported in order, streaming stores with unaligned heads and tails, pitched rows."""
import ctypes as ct

import numpy as np
import pytest

from pygofpid_b200 import _native as nat


def _copy(dst, dpitch, src, spitch, row, nrows, threads):
    L = nat.lib()
    nat.check(L.gofpid_host_copy2d(dst, dpitch, src, spitch, row, nrows, threads))


@pytest.mark.parametrize("threads", [0, 1, 3, 7, -1])
@pytest.mark.parametrize("rows,row,spitch,dpitch,doff", [
    (1080, 1920, 1920, 1920, 0),      # a dense 1080p frame: chunks of whole rows, four strips
    (1080, 1920, 2048, 1920, 0),      # pitched source
    (577, 777, 800, 1024, 3),         # odd row length, both pitched, destination not 16-byte aligned
    (1, 5_000_000, 5_000_000, 5_000_000, 1),  # one long row: a single chunk, the caller alone
    (300, 100, 100, 100, 0),          # rows below the streaming-store threshold when pitched ...
    (300, 100, 128, 100, 5),          # ... and copied row by row
    (5, 7, 7, 7, 0),                  # fewer chunks than strips
])
def test_host_copy2d_matches_numpy(threads, rows, row, spitch, dpitch, doff):
    rng = np.random.default_rng(rows * 31 + row + threads)
    src = rng.integers(0, 256, size=rows * spitch + 64, dtype=np.uint8)
    dst = np.full(rows * dpitch + 64 + doff, 0xA5, np.uint8)
    for rep in range(3 if rows * row > 1_000_000 else 20):  # repeated: a late worker must never touch the next job
        src[: rows * spitch] = rng.integers(0, 256, size=rows * spitch, dtype=np.uint8) if rep else src[: rows * spitch]
        dst[:] = 0xA5
        _copy(dst.ctypes.data + doff, dpitch, src.ctypes.data, spitch, row, rows, threads)
        want = src[: rows * spitch].reshape(rows, spitch)[:, :row]
        got = np.lib.stride_tricks.as_strided(dst[doff:], shape=(rows, row), strides=(dpitch, 1))
        assert np.array_equal(got, want), (rep, threads)
        # nothing outside the rows was written
        assert np.all(dst[:doff] == 0xA5) and np.all(dst[doff + (rows - 1) * dpitch + row:] == 0xA5)
        if dpitch > row and rows > 1:
            gaps = np.lib.stride_tricks.as_strided(dst[doff + row:], shape=(rows - 1, dpitch - row), strides=(dpitch, 1))
            assert np.all(gaps == 0xA5)


def test_host_copy2d_argument_checks():
    L = nat.lib()
    a = np.zeros(100, np.uint8)
    assert L.gofpid_host_copy2d(a.ctypes.data, 10, a.ctypes.data, 10, 0, 10, 0) == 0      # nothing to do
    assert L.gofpid_host_copy2d(None, 10, a.ctypes.data, 10, 10, 5, 0) == nat.ERR_INVALID
    assert L.gofpid_host_copy2d(a.ctypes.data, 5, a.ctypes.data, 10, 10, 5, 0) == nat.ERR_INVALID
    assert "pitch" in nat.last_error()
