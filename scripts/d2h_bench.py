#!/usr/bin/env python
"""Host-side copy costs that decide how results leave the device (config 2 returns 15 GB of CSC
arrays): pinned allocation, pinned / pageable / registered device->host copies."""
import ctypes
import json
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfermion_b200 import _lib  # noqa: E402

vp = ctypes.c_void_p
GB = 1 << 30
size = 4 * GB
d = _lib.DeviceBuffer(size)
_lib.call('ofb_memset', vp(d.ptr), ctypes.c_int(1), ctypes.c_size_t(size), None)
_lib.call('ofb_device_sync')
out = {}


def clock(fn):
    t0 = time.perf_counter()
    fn()
    _lib.call('ofb_device_sync')
    return time.perf_counter() - t0


h = vp()
out['host_alloc_4GB_s'] = clock(lambda: _lib.call('ofb_host_alloc', ctypes.byref(h), ctypes.c_size_t(size)))
out['d2h_pinned_4GB_s'] = clock(lambda: _lib.call('ofb_memcpy_d2h', h, vp(d.ptr), ctypes.c_size_t(size), None))
out['d2h_pinned_again_s'] = clock(lambda: _lib.call('ofb_memcpy_d2h', h, vp(d.ptr), ctypes.c_size_t(size), None))
out['host_free_s'] = clock(lambda: _lib.call('ofb_host_free', h))
fresh = numpy.empty(size, dtype=numpy.uint8)
out['d2h_pageable_fresh_4GB_s'] = clock(lambda: _lib.call('ofb_memcpy_d2h', _lib.ptr(fresh), vp(d.ptr), ctypes.c_size_t(size), None))
out['d2h_pageable_touched_4GB_s'] = clock(lambda: _lib.call('ofb_memcpy_d2h', _lib.ptr(fresh), vp(d.ptr), ctypes.c_size_t(size), None))
out['host_register_touched_4GB_s'] = clock(lambda: _lib.call('ofb_host_register', _lib.ptr(fresh), ctypes.c_size_t(size)))
out['d2h_registered_4GB_s'] = clock(lambda: _lib.call('ofb_memcpy_d2h', _lib.ptr(fresh), vp(d.ptr), ctypes.c_size_t(size), None))
out['host_unregister_s'] = clock(lambda: _lib.call('ofb_host_unregister', _lib.ptr(fresh)))
fresh2 = numpy.empty(size, dtype=numpy.uint8)
out['host_register_fresh_4GB_s'] = clock(lambda: _lib.call('ofb_host_register', _lib.ptr(fresh2), ctypes.c_size_t(size)))
out['host_unregister_fresh_s'] = clock(lambda: _lib.call('ofb_host_unregister', _lib.ptr(fresh2)))
out['h2d_pageable_4GB_s'] = clock(lambda: _lib.call('ofb_memcpy_h2d', vp(d.ptr), _lib.ptr(fresh), ctypes.c_size_t(size), None))
out['numpy_memcpy_4GB_s'] = clock(lambda: numpy.copyto(fresh2, fresh))
out['cpu_count'] = os.cpu_count()
print(json.dumps(out))
