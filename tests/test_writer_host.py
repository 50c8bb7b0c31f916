"""The writer hand-off ring (include/poppy_host.h poppy_host_writer_*, reference src/poppy.hpp:219) over a fake transport:
ordering, back-pressure, the convert pool and flush semantics - no GPU involved."""
import ctypes as C
import threading
import time

import numpy as np
import pytest

from poppy_b200 import _lib


class IO(C.Structure):
    _fields_ = [("user", C.c_void_p),
                ("download", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_uint64))),
                ("wait", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64)),
                ("alloc", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p))),
                ("release", C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)),
                ("write", C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_size_t)),
                ("convert", C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_size_t)),
                ("owned_transport", C.c_void_p)]


@pytest.mark.parametrize("workers", [0, 3])
def test_frames_are_written_in_order_with_backpressure(workers):
    lib = _lib.load()
    w, h, ring, n = 16, 8, 3, 40
    fb = w * h * 3
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    state = {"in_flight": 0, "max_in_flight": 0, "written": [], "lock": threading.Lock(), "waited": []}

    def download(user, slot, dst, step, ticket):
        with state["lock"]:
            state["in_flight"] += 1
            state["max_in_flight"] = max(state["max_in_flight"], state["in_flight"])
        C.memset(dst, slot & 255, fb)                 # "frame" of slot s = bytes s
        ticket[0] = 1000 + slot
        return 0

    def wait(user, ticket):
        state["waited"].append(ticket)
        time.sleep(0.001)
        return 0

    def alloc(user, size, out):
        out[0] = libc.malloc(size)
        return 0

    def release(user, p):
        libc.free(p)

    def convert(user, idx, bgr, ww, hh, step):
        bgr[0] = (bgr[0] + 1) & 255                   # the pool touched the frame before it is written

    def write(user, idx, bgr, ww, hh, step):
        time.sleep(0.0005)
        with state["lock"]:
            state["written"].append((idx, bgr[0], bgr[fb - 1]))
            state["in_flight"] -= 1

    io = IO()
    cbs = dict(download=IO._fields_[1][1](download), wait=IO._fields_[2][1](wait), alloc=IO._fields_[3][1](alloc),
               release=IO._fields_[4][1](release), write=IO._fields_[5][1](write), convert=IO._fields_[6][1](convert))
    io.download, io.wait, io.alloc, io.release, io.write = cbs["download"], cbs["wait"], cbs["alloc"], cbs["release"], cbs["write"]
    if workers:
        io.convert = cbs["convert"]
    wr = C.c_void_p()
    assert lib.poppy_host_writer_create_io(C.byref(wr), C.byref(io), w, h, ring, workers) == 0
    assert lib.poppy_host_writer_submit(wr, 0, 25, 100) == 0        # slots 0..24 -> frames 100..124
    assert lib.poppy_host_writer_submit(wr, 25, n - 25, 125) == 0
    assert lib.poppy_host_writer_flush(wr) == 0
    lib.poppy_host_writer_destroy(wr)
    assert [x[0] for x in state["written"]] == list(range(100, 100 + n))          # strictly in order
    for k, (idx, first, last) in enumerate(state["written"]):
        assert last == k & 255 and first == ((k + 1) & 255 if workers else k & 255)
    assert state["max_in_flight"] <= ring                                         # never more frames in flight than buffers
    assert sorted(state["waited"]) == [1000 + k for k in range(n)]
