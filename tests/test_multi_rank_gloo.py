"""world_size-2 gloo test (CPU) of the N>1 path: frame sharding is a partition, and the single
collective (LUT blob broadcast from rank 0) delivers bit-identical tables to every rank."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import uhdr_testlib as T
from libultrahdr_b200.sharding import broadcast_lut_blob, frames_for_rank


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = C.CDLL(T.GPU_SO)
    lib.uhdr_b200_lut_blob_floats.restype = C.c_size_t
    n = lib.uhdr_b200_lut_blob_floats()
    blob = torch.zeros(n, dtype=torch.float32)
    if rank == 0:
        host = np.zeros(n, np.float32)
        assert lib.uhdr_b200_build_lut_blob(host.ctypes.data_as(C.c_void_p)) == 0
        blob.copy_(torch.from_numpy(host))
    broadcast_lut_blob(blob, dist, src=0)
    mine = np.zeros(n, np.float32)
    lib.uhdr_b200_build_lut_blob(mine.ctypes.data_as(C.c_void_p))
    same = bool((blob.numpy().view(np.uint32) == mine.view(np.uint32)).all())
    q.put((rank, same, frames_for_rank(257, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    import __graft_entry__ as g
    g.build()
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    frames = []
    for rank, same, fr in res:
        assert same, rank
        frames += fr
    assert sorted(frames) == list(range(257))


def test_sharding_is_a_partition():
    for n, w in ((256, 8), (7, 8), (100, 3), (1, 1), (0, 4)):
        got = sum((frames_for_rank(n, r, w) for r in range(w)), [])
        assert got == list(range(n))
        sizes = [len(frames_for_rank(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
