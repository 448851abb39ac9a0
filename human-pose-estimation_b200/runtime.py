"""Host-side runtime objects over the C ABI: Context (one per GPU), DeviceArray
(a typed handle to device memory from smplb_malloc) and pinned host arrays.
numpy is the only dependency; there is no torch/cupy in the product."""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import DEVICE, HOST, check, lib


class DeviceArray(object):
    """fp32/int32 array in device memory owned by a Context."""

    def __init__(self, ctx, shape, dtype=np.float32):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib().smplb_malloc(ctx.handle, C.byref(p), max(self.nbytes, 1)))
        self.ptr = p.value

    def copy_from(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.nbytes == self.nbytes, (host.shape, self.shape)
        check(lib().smplb_memcpy_h2d(self.ctx.handle, self.ptr, host.ctypes.data, self.nbytes))
        if not getattr(host, "_smplb_pinned", False):
            self.ctx.sync()   # pageable source: make the copy safe against reuse of `host`
        return self

    def numpy(self, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        check(lib().smplb_memcpy_d2h(self.ctx.handle, out.ctypes.data, self.ptr, self.nbytes))
        self.ctx.sync()
        return out

    def zero_(self):
        check(lib().smplb_memset(self.ctx.handle, self.ptr, 0, self.nbytes))
        return self

    def free(self):
        if self.ptr and self.ctx.handle:
            lib().smplb_free(self.ctx.handle, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray(np.ndarray):
    _smplb_pinned = True


def pinned_empty(shape, dtype=np.float32):
    """numpy array over page-locked host memory (true async H2D/D2H)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    p = C.c_void_p()
    check(lib().smplb_host_alloc(C.byref(p), max(n, 1)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    # every view of the array keeps `buf` alive; the page-locked block goes back when the last one is gone
    weakref.finalize(buf, _host_free, p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape).view(PinnedArray)


def _host_free(ptr):
    try:
        lib().smplb_host_free(C.c_void_p(ptr))
    except Exception:
        pass


def _as_f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


class Context(object):
    """smplb_ctx: device copies of the model constants, workspace, one stream."""

    def __init__(self, v_template, shapedirs, posedirs, J_regressor, weights, joint_regressor, parents,
                 device=0, max_batch=64):
        self.handle = None
        keep = [_as_f32(v_template), _as_f32(shapedirs), _as_f32(posedirs), _as_f32(J_regressor), _as_f32(weights),
                _as_f32(joint_regressor), np.ascontiguousarray(parents, dtype=np.int32)]
        V = keep[0].shape[0]
        NB = keep[1].shape[0]
        K = keep[5].shape[1]
        assert keep[0].shape == (V, 3) and keep[1].shape == (NB, 3 * V) and keep[2].shape == (207, 3 * V)
        assert keep[3].shape == (V, 24) and keep[4].shape == (V, 24) and keep[5].shape == (V, K) and keep[6].shape == (24,)
        m = _lib.Model(V, NB, K, 0, *[a.ctypes.data for a in keep])
        h = C.c_void_p()
        check(lib().smplb_create(C.byref(h), C.byref(m), int(device), int(max_batch)))
        self.handle = h
        self.V, self.NB, self.K, self.device = V, NB, K, int(device)

    # -- memory / ordering
    def empty(self, shape, dtype=np.float32):
        return DeviceArray(self, shape, dtype)

    def to_device(self, host, dtype=np.float32):
        host = np.ascontiguousarray(host, dtype=dtype)
        return DeviceArray(self, host.shape, dtype).copy_from(host)

    def sync(self):
        check(lib().smplb_sync(self.handle))

    def order_after(self, other):
        """Work enqueued on this context from now on runs after what `other` has enqueued."""
        check(lib().smplb_order_after(self.handle, other.handle))

    def flush_l2(self, nbytes=256 << 20):
        check(lib().smplb_flush_l2(self.handle, int(nbytes)))

    def timer_start(self, slot=0):
        check(lib().smplb_timer_start(self.handle, slot))

    def timer_stop(self, slot=0):
        check(lib().smplb_timer_stop(self.handle, slot))

    def timer_ms(self, slot=0):
        ms = C.c_float()
        check(lib().smplb_timer_elapsed_ms(self.handle, slot, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_int64()
        check(lib().smplb_launch_count(self.handle, C.byref(n)))
        return n.value

    def debug_set(self, key, value):
        check(lib().smplb_debug_set(self.handle, key.encode(), int(value)))

    def profile(self, on):
        """True / 1: per-kernel times (streams serialised); 2: timeline trace (profile_trace)."""
        check(lib().smplb_profile_enable(self.handle, int(on)))

    def profile_trace(self):
        """[(name, start_ms, end_ms)] of every launch since profile(2), relative to a
        process-wide reference event (comparable across the contexts of a device)."""
        buf = C.create_string_buffer(1 << 20)
        check(lib().smplb_profile_read(self.handle, buf, len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            name, t0, t1 = line.split()
            out.append((name, float(t0), float(t1)))
        return out

    def profile_read(self):
        buf = C.create_string_buffer(1 << 16)
        check(lib().smplb_profile_read(self.handle, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split()
            out[name] = (float(ms), int(n))
        return out

    # -- NCCL plumbing (one process per GPU)
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        check(lib().smplb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks, rank, uid):
        check(lib().smplb_comm_init(self.handle, nranks, rank, C.c_char_p(uid)))

    # -- mailbox exchange over peer memory (the default transport of a sharded step)
    def p2p_export(self):
        """64-byte CUDA IPC handle of this context's mailbox (send it to every rank)."""
        buf = C.create_string_buffer(64)
        check(lib().smplb_comm_p2p_export(self.handle, buf))
        return buf.raw

    def p2p_attach(self, nranks, rank, handles):
        """handles: the nranks exported handles in rank order (one process per rank)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * nranks
        check(lib().smplb_comm_p2p_attach(self.handle, nranks, rank, C.c_char_p(blob)))

    def p2p_attach_local(self, rank, contexts):
        """Ranks that are contexts of this process: contexts[rank] is self."""
        arr = (C.c_void_p * len(contexts))(*[x.handle for x in contexts])
        check(lib().smplb_comm_p2p_attach_local(self.handle, len(contexts), rank, arr))

    def p2p_inject(self, kind, epoch, from_rank, v0=0.0, v1=0.0, cnt=0):
        """Test hook: rank `from_rank`'s push of exchange `epoch` written into this context's mailbox."""
        check(lib().smplb_debug_p2p_inject(self.handle, int(kind), int(epoch), int(from_rank), float(v0), float(v1), int(cnt)))

    def comm_status(self):
        st = C.c_int()
        check(lib().smplb_comm_status(self.handle, C.byref(st)))
        return st.value

    def comm_destroy(self):
        check(lib().smplb_comm_destroy(self.handle))

    def allreduce_sum(self, dev_array, count=None):
        n = int(np.prod(dev_array.shape)) if count is None else count
        check(lib().smplb_comm_allreduce_sum(self.handle, dev_array.ptr, n))

    def close(self):
        if self.handle:
            lib().smplb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Args(object):
    """Resolves a call's arrays to raw pointers and the `mem` flag: all numpy
    (SMPLB_HOST) or all DeviceArray (SMPLB_DEVICE), never mixed."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.mem = None
        self.keep = []

    def _kind(self, mem):
        if self.mem is None:
            self.mem = mem
        elif self.mem != mem:
            raise TypeError("a call takes either numpy arrays or DeviceArrays, not a mix")

    def inp(self, x, shape=None, dtype=np.float32):
        if x is None:
            return None
        if isinstance(x, DeviceArray):
            self._kind(DEVICE)
            if shape is not None:
                assert int(np.prod(x.shape)) == int(np.prod(shape)), (x.shape, shape)
            return x.ptr
        self._kind(HOST)
        a = np.ascontiguousarray(x, dtype=dtype)
        if shape is not None:
            a = a.reshape(shape)
        self.keep.append(a)
        return a.ctypes.data

    def out(self, shape, dtype=np.float32, want=True, like_device=None):
        """Allocates an output of the call's kind; returns (object, pointer)."""
        if not want:
            return None, None
        if self.mem is None:
            self.mem = HOST
        if self.mem == DEVICE:
            d = DeviceArray(self.ctx, shape, dtype)
            return d, d.ptr
        a = np.empty(shape, dtype=dtype)
        self.keep.append(a)
        return a, a.ctypes.data
