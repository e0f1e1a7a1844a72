"""Host side of update mode (gelato_b200/csrc/host_pool.h): the worker pool that scatters the packed
x-dependent Jacobian slots into the caller's buffers.  Compiled into the host emulator for this tier."""
import ctypes
import threading

import numpy as np

import emu_binding

_pi = ctypes.POINTER(ctypes.c_int64)
_pd = ctypes.POINTER(ctypes.c_double)


def _lib():
    emu_binding.build()
    L = ctypes.CDLL(emu_binding.LIB)
    L.emu_scatter_parallel.argtypes = [_pi, ctypes.c_longlong, _pd, _pd, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int]
    L.emu_scatter_parallel.restype = None
    return L


def _scatter(L, idx, packed, vals, s0, s1, threads):
    L.emu_scatter_parallel(idx.ctypes.data_as(_pi), idx.size, packed.ctypes.data_as(_pd), vals.ctypes.data_as(_pd),
                           vals.shape[1], s0, s1, threads)


def test_scatter_matches_numpy_for_any_thread_count():
    L = _lib()
    rng = np.random.default_rng(3)
    n_vals, n_scen = 5000, 37
    idx = np.sort(rng.choice(n_vals, 700, replace=False)).astype(np.int64)
    packed = rng.random((n_scen, idx.size))
    for threads in (0, 1, 2, 3, 8, 64):
        for s0, s1 in ((0, n_scen), (5, 6), (11, 30), (4, 4)):
            vals = np.full((n_scen, n_vals), -1.0)
            want = vals.copy()
            want[s0:s1, idx] = packed[s0:s1]
            _scatter(L, idx, packed, vals, s0, s1, threads)
            assert np.array_equal(vals, want), (threads, s0, s1)


def test_pool_is_reused_and_serialises_concurrent_callers():
    L = _lib()
    rng = np.random.default_rng(4)
    n_vals, n_scen = 3000, 16
    idx = np.sort(rng.choice(n_vals, 400, replace=False)).astype(np.int64)
    packed = [rng.random((n_scen, idx.size)) for _ in range(4)]
    vals = [np.zeros((n_scen, n_vals)) for _ in range(4)]

    def work(k):
        for _ in range(50):
            _scatter(L, idx, packed[k], vals[k], 0, n_scen, 4)

    callers = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for c in callers:
        c.start()
    for c in callers:
        c.join()
    for k in range(4):
        assert np.array_equal(vals[k][:, idx], packed[k])
    assert 3 <= L.emu_pool_workers() <= 63  # grown on demand, never one set of threads per call
