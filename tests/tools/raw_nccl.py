"""A raw NCCL communicator through ctypes (test helper): what a native host program would own and hand to
wol_hist_allreduce.  Uses the NCCL library bundled with torch."""
import ctypes
import glob
import os


class UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_char * 128)]


def load():
    import nvidia.nccl
    path = glob.glob(os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2"))[0]
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(UniqueId)]
    lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
    lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    return lib


def unique_id(lib):
    uid = UniqueId()
    assert lib.ncclGetUniqueId(ctypes.byref(uid)) == 0
    return uid


def comm_init(lib, uid, world, rank):
    comm = ctypes.c_void_p()
    rc = lib.ncclCommInitRank(ctypes.byref(comm), world, uid, rank)
    assert rc == 0, "ncclCommInitRank failed: %d" % rc
    return comm
