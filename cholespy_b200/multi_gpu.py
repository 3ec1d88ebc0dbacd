"""RHS-column sharding of the solve path over several GPUs (one process per GPU, torch.distributed).

The path shards only across right-hand-side columns: the factor image (both sweep slabs + task tables,
see csrc/panel_schedule.h) is replicated once from rank 0 with a broadcast (NCCL over NVLink on the GPU
box), and every rank then solves its own column block with no data-path collective.  New: the
reference is single-GPU (it always uses device 0, reference src/cuda_driver.cpp:124).
"""
import ctypes

from . import _capi


def column_shards(k_total, world):
    """[(first_column, n_columns)] per rank: contiguous blocks, sizes differ by at most one."""
    base, extra = divmod(int(k_total), int(world))
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, n))
        start += n
    return out


class _DevBuf:
    """Exposes a raw device allocation to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def image_tensors(lib, handle, device):
    """Zero-copy uint8 torch views of the solver's replicable device buffers: [(name, tensor)]."""
    import torch
    out = []
    for i in range(lib.chb_solver_image_count(handle)):
        name, ptr, nbytes = ctypes.c_char_p(), ctypes.c_void_p(), ctypes.c_uint64()
        _capi.check(lib.chb_solver_image_buffer(handle, i, ctypes.byref(name), ctypes.byref(ptr), ctypes.byref(nbytes)))
        out.append((name.value.decode(), torch.as_tensor(_DevBuf(ptr.value, nbytes.value), device=device)))
    return out


def replicate_solver(cls, handle, dtype_code, device, rank, world, dist):
    """Rank 0 owns a built solver (`handle`); every other rank gets a replica of the same shape and the
    image is broadcast into it.  Returns (solver, shape_words, bytes_broadcast)."""
    import torch
    lib = _capi.load()
    meta = torch.zeros(_capi.IMAGE_SHAPE_WORDS, dtype=torch.int64, device=device)
    if rank == 0:
        shape = (ctypes.c_int64 * _capi.IMAGE_SHAPE_WORDS)()
        _capi.check(lib.chb_solver_image_shape(handle, shape, _capi.IMAGE_SHAPE_WORDS))
        meta = torch.tensor(list(shape), dtype=torch.int64, device=device)
    if world > 1:
        dist.broadcast(meta, src=0)
    words = [int(v) for v in meta.tolist()]
    if rank != 0:
        _capi.check(lib.chb_init(device.index))
        handle = ctypes.c_void_p()
        shape = (ctypes.c_int64 * _capi.IMAGE_SHAPE_WORDS)(*words)
        _capi.check(lib.chb_solver_create_replica(ctypes.byref(handle), dtype_code, shape, _capi.IMAGE_SHAPE_WORDS, device.index))
    solver = cls._from_handle(handle, words[0], device.index)
    total = 0
    if world > 1:
        for _, tensor in image_tensors(lib, handle, device):
            dist.broadcast(tensor, src=0)
            total += tensor.numel()
    return solver, words, total
