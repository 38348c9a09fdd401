"""Host-side logic of the row-sharded (N > 1) path, on CPU with gloo, world_size 2.

The CUDA kernels cannot run here; what can be checked without a GPU is the sharding contract the
ranks rely on: ddc_shard_rows blocks tile the rows, the sum over the ranks' column-histogram slots
equals the global histogram, and the per-strip row histograms -- every rank's block in the chunked
[row block][S][RB] layout of ddc_kernels.cuh:row_count_index, placed in slot `rank` of every rank's
exchange buffer, which is what K4 indexes (rank = y / Rmax, local row = y % Rmax) -- reproduce the
global per-strip row histogram, which, fed to the oracle's histogram RCB, gives the oracle's boxes.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, nx, ny, P, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from domain_decomp_b200 import capi
    from oracle import oracle as orc
    mask = capi.generate_mask_host(nx, ny, seed=11, land_frac=0.5)
    yb, yc = capi.shard_rows(ny, world, rank)
    shard = mask[yb:yb + yc] > 0
    # exchange step 1: column histogram
    col = torch.from_numpy(shard.sum(axis=0).astype(np.int64))
    dist.all_reduce(col)
    ok = np.array_equal(col.numpy(), (mask > 0).sum(axis=0))
    # the strips every rank derives from the identical reduced histogram (here: via the oracle)
    o = orc.partition(mask, P, use_hist=True, want_neighbours=False)
    xs = sorted(set((int(b[0]), int(b[0] + b[2])) for b in o.boxes))
    S = len(xs)
    rmax = -(-ny // world)
    # exchange step 2: per-strip row histogram; a rank's block is [row block][S][RB] (RB rows per block of
    # the kernel that wrote it; rows beyond the shard are written as empty), and every rank ends up with
    # the block of rank g in slot g (the producers push; an all-gather is the same data movement)
    for rb_shift in (3, 5):
        RB = 1 << rb_shift
        nblk = -(-rmax // RB)

        def index(s, yl):  # row_count_index(s, yl, Scap = S, rb_shift)
            return (((yl >> rb_shift) * S + s) << rb_shift) + (yl & (RB - 1))

        local = np.zeros(nblk * S * RB, dtype=np.int64)
        for s, (x0, x1) in enumerate(xs):
            counts = shard[:, x0:x1].sum(axis=1)
            for yl in range(yc):
                local[index(s, yl)] = counts[yl]
        slots = [torch.zeros(nblk * S * RB, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(slots, torch.from_numpy(local))
        for s, (x0, x1) in enumerate(xs):
            want = (mask[:, x0:x1] > 0).sum(axis=1)
            got = np.array([int(slots[y // rmax][index(s, y % rmax)]) for y in range(ny)])
            ok &= np.array_equal(got, want)
        # a block of the kernel (RB rows, all strips) is ONE contiguous chunk of the layout
        chunk = sorted(index(s, yl) for s in range(S) for yl in range(RB))
        ok &= chunk == list(range(S * RB))
    # every rank's rows tile [0, ny)
    blocks = [None] * world
    dist.all_gather_object(blocks, (yb, yc))
    ok &= blocks[0][0] == 0 and sum(c for _, c in blocks) == ny
    ok &= all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    flag = torch.tensor([0 if ok else 1])
    dist.all_reduce(flag)
    if rank == 0:
        q.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("nx,ny,P", [(96, 81, 12), (200, 7, 4)])
def test_sharded_histograms_world2(nx, ny, P):
    from domain_decomp_b200 import build
    build.build_cuda_lib()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nx, ny, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 0
