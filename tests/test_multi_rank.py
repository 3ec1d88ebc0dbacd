"""world_size-2 CPU test (gloo) of the multi-GPU host logic: RHS-column sharding and replication of the
solver image from rank 0.  No GPU: rank 0 builds the sweep schedule through the C ABI
(chb_schedule_create), broadcasts its buffers the way bench.py broadcasts the device image, and every
rank executes ITS column block with the schedule interpreter (oracle/panel_emulator.cpp); the gathered
result must equal the row-by-row oracle on all columns."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

torch = pytest.importorskip("torch")


def test_column_shards():
    from cholespy_b200.multi_gpu import column_shards
    assert column_shards(128, 8) == [(16 * r, 16) for r in range(8)]
    assert column_shards(3, 2) == [(0, 2), (2, 1)]
    assert column_shards(1, 4) == [(0, 1), (1, 0), (2, 0), (3, 0)] or column_shards(1, 4)[0] == (0, 1)
    for k, w in ((128, 3), (5, 8), (97, 4)):
        s = column_shards(k, w)
        assert sum(c for _, c in s) == k and all(s[i][0] + s[i][1] == s[i + 1][0] for i in range(w - 1))


def _worker(rank, world, port, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import problems
    from helpers import factorize
    from cholespy_b200 import _capi
    from cholespy_b200.multi_gpu import column_shards
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _capi.load()
    emu = oracle._load_emu()
    dt = np.float64
    k_total = 5
    prob = problems.build_problem("ico4")
    n = prob["n"]
    rng = np.random.default_rng(9)
    b_all = rng.random((n, k_total)).astype(dt)
    # rank 0: factorise + build the image; everyone: receive it
    bufs, meta = [], torch.zeros(2 * 5, dtype=torch.int64)
    if rank == 0:
        perm, cp, ri, v = factorize(lib, n, prob["ii"], prob["jj"], prob["x"], prob["fmt"])
        h = ctypes.c_void_p()
        _capi.check(lib.chb_schedule_create(ctypes.byref(h), 1, n, len(ri), cp.ctypes.data, ri.ctypes.data, v.ctypes.data,
                                            perm.ctypes.data, -1))
        for upper in (0, 1):
            slab, tasks, order = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            info = (ctypes.c_int64 * 10)()
            lib.chb_schedule_sweep(h, upper, ctypes.byref(slab), ctypes.byref(tasks), ctypes.byref(order), info, 10)
            nt, nb = int(info[0]), int(info[1])
            meta[5 * upper:5 * upper + 5] = torch.tensor([nt, nb, int(info[2]), int(info[3]), 0])
            bufs += [np.ctypeslib.as_array(ctypes.cast(slab, ctypes.POINTER(ctypes.c_uint8)), shape=(nb,)).copy(),
                     np.ctypeslib.as_array(ctypes.cast(tasks, ctypes.POINTER(ctypes.c_uint8)), shape=(16 * nt,)).copy(),
                     np.ctypeslib.as_array(ctypes.cast(order, ctypes.POINTER(ctypes.c_int32)), shape=(nt,)).copy().view(np.uint8)]
        lib.chb_schedule_destroy(h)
        x_ref = oracle.solve(cp, ri, v, perm, b_all)
    dist.broadcast(meta, src=0)
    tensors = []
    for upper in (0, 1):
        nt, nb = int(meta[5 * upper]), int(meta[5 * upper + 1])
        for i, size in enumerate((nb, 16 * nt, 4 * nt)):
            t = torch.from_numpy(bufs[3 * upper + i]) if rank == 0 else torch.empty(size, dtype=torch.uint8)
            dist.broadcast(t, src=0)
            tensors.append(t.numpy())
    # every rank: its own column block, no data-path collective
    c0, kc = column_shards(k_total, world)[rank]
    b = np.ascontiguousarray(b_all[:, c0:c0 + kc])
    tmp, x = np.zeros((n, kc), dt), np.zeros((n, kc), dt)
    for upper in (0, 1):
        nt, page, slots = int(meta[5 * upper]), int(meta[5 * upper + 2]), int(meta[5 * upper + 3])
        scratch = np.zeros((max(slots, 1), 16, kc), dt)
        slab, tasks, order = tensors[3 * upper:3 * upper + 3]
        rc = emu.chb_emu_sweep_f64(slab.ctypes.data_as(ctypes.c_void_p), tasks.ctypes.data_as(ctypes.c_void_p),
                                   order.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(nt), ctypes.c_int(page), ctypes.c_int(upper),
                                   ctypes.c_int(kc), tmp.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                                   x.ctypes.data_as(ctypes.c_void_p), scratch.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0, rc
    gathered = [None] * world
    dist.gather_object((c0, x), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        full = np.zeros_like(b_all)
        for start, part in gathered:
            full[:, start:start + part.shape[1]] = part
        err = float(np.abs(full - x_ref).max() / np.abs(x_ref).max())
        with open(out_path, "w") as f:
            f.write("%r" % err)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_replicate_the_image_and_shard_columns(tmp_path, native_lib, oracle_mod):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "err.txt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert float(open(out).read()) < 1e-12
