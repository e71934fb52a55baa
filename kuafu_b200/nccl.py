"""A raw NCCL communicator for kfrtReduceNccl (include/kf_rt.h), obtained the way a C++ host would:
ncclGetUniqueId on rank 0, the 128-byte id handed to the other ranks, ncclCommInitRank everywhere.
Plumbing only: `torch.distributed` (already initialised by the caller) carries the id; the reduce of
the sample sums itself is issued by the C ABI on the communicator this returns."""
import ctypes as C


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (loads the bundled libnccl.so.2, so the soname below resolves to it)
        lib = C.CDLL("libnccl.so.2")
        lib.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
        lib.ncclGetUniqueId.restype = C.c_int
        lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        lib.ncclCommInitRank.restype = C.c_int
        lib.ncclCommDestroy.argtypes = [C.c_void_p]
        lib.ncclCommDestroy.restype = C.c_int
        lib.ncclGetErrorString.argtypes = [C.c_int]
        lib.ncclGetErrorString.restype = C.c_char_p
        _lib = lib
    return _lib


class Communicator:
    """ncclComm_t over all ranks of the default torch.distributed group; `.handle` goes to
    kfrtReduceNccl.  The caller must have made its GPU current (torch.cuda.set_device)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        lib = _load()
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = _UniqueId()
        if rank == 0:
            self._ck(lib.ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        C.memmove(C.byref(uid), t.cpu().numpy().tobytes(), 128)
        comm = C.c_void_p()
        self._ck(lib.ncclCommInitRank(C.byref(comm), world, uid, rank), "ncclCommInitRank")
        self.handle = comm.value
        self.rank, self.world = rank, world

    @staticmethod
    def _ck(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {_load().ncclGetErrorString(rc).decode()}")

    def close(self):
        if getattr(self, "handle", None):
            _load().ncclCommDestroy(self.handle)
            self.handle = None
