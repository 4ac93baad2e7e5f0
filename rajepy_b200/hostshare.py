"""
Host arrays shared by the ranks of one node (POSIX shared memory), for the hand-over of
channel-sharded cubes: every rank moves only the planes it integrated, over its own PCIe
link, into ONE (nchan, nx, nz) array that the host ranks return to the caller.

`segment(nbytes, rank)` is collective over the default torch.distributed group: rank 0 picks a
free pooled segment of that size (or creates one) and broadcasts its name; every rank maps it
and page-locks it once (cudaHostRegister), so later hand-overs reuse both the pages and the
registration.  A segment is free again when the array handed to the caller has been garbage
collected (like torch's caching host allocator does for pinned tensors).
"""
import atexit
import os
from multiprocessing import shared_memory

import numpy as np

_SEGS = {}        # name -> Segment (per process)
_SEQ = [0]


class Segment:
    def __init__(self, name, nbytes, create):
        self.name, self.nbytes, self.owner = name, nbytes, create
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=nbytes)
        if not create:
            # the creating rank unlinks; keep Python's resource tracker from doing it too
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:  # noqa: BLE001
                pass
        self.busy = False
        self.registered = {}          # (offset, nbytes) page-locked by THIS process
        self._np = np.ndarray((nbytes,), dtype=np.uint8, buffer=self.shm.buf)

    def _register(self, offset, nbytes):
        """Page-lock [offset, offset + nbytes) for this process's GPU (only the planes the rank
        writes: the pinned total over all ranks stays one cube, whatever their number)."""
        if nbytes == 0 or (offset, nbytes) in self.registered:
            return
        import torch
        page = 4096
        lo = offset // page * page
        hi = -(-(offset + nbytes) // page) * page
        hi = min(hi, -(-self.nbytes // page) * page)
        err = torch.cuda.cudart().cudaHostRegister(self._np.ctypes.data + lo, hi - lo, 0)
        code = getattr(err, "value", err)
        code = code[0] if isinstance(code, tuple) else code
        if int(code) != 0:
            raise RuntimeError(f"cudaHostRegister of {hi - lo} bytes failed ({err})")
        self.registered[(offset, nbytes)] = lo

    def tensor(self, offset, shape, pin=True):
        """float64 torch view of `shape` at byte `offset`; page-locked for this process's GPU
        unless `pin` is False (host threads are the only writers)."""
        import torch
        n = int(np.prod(shape))
        if pin:
            self._register(offset, 8 * n)
        view = self._np[offset: offset + 8 * n].view(np.float64).reshape(shape)
        return torch.from_numpy(view)

    def array(self, shape):
        """The caller's array; the segment is reusable once it AND every view derived from it
        have been collected (numpy collapses the base of a view to the first non-array owner,
        which is the lease object below)."""
        self.busy = True
        return np.asarray(_Lease(self, tuple(int(v) for v in shape)))

    def release(self):
        self.busy = False

    def close(self):
        try:
            import torch
            for lo in self.registered.values():
                torch.cuda.cudart().cudaHostUnregister(self._np.ctypes.data + lo)
        except Exception:  # noqa: BLE001
            pass
        self._np = None
        try:
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except Exception:  # noqa: BLE001
            pass


class _Lease:
    def __init__(self, seg, shape):
        self._seg = seg
        self.__array_interface__ = {"shape": shape, "typestr": "<f8", "version": 3,
                                    "data": (seg._np.ctypes.data, False)}

    def __del__(self):
        self._seg.release()


def segment(nbytes, rank):
    import torch.distributed as dist
    name = None
    if rank == 0:
        for seg in _SEGS.values():
            if seg.owner and not seg.busy and seg.nbytes == nbytes:
                name = seg.name
                break
        if name is None:
            _SEQ[0] += 1
            name = f"rjp_{os.getpid()}_{_SEQ[0]}"
            _SEGS[name] = Segment(name, nbytes, create=True)
    box = [name]
    dist.broadcast_object_list(box, src=0)
    name = box[0]
    if name not in _SEGS:
        _SEGS[name] = Segment(name, nbytes, create=False)
    return _SEGS[name]


@atexit.register
def _cleanup():
    for seg in list(_SEGS.values()):
        seg.close()
    _SEGS.clear()
