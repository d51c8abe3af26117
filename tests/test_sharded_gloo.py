"""Host-side logic of the N>1 path, world_size 2 (and 3) over gloo on the CPU: the row
partition, the handle exchange, and the collective round loop with the oracle's row pass
injected as the compute backend (test infrastructure -- the product backend is CUDA)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from eigen_value_b200 import sharded


def test_shard_rows_partition_covers_everything():
    for dim in (3, 8, 1000, 8191, 131072):
        for world in (1, 2, 3, 4, 8):
            if world > dim:
                continue
            blocks = [sharded.shard_rows(dim, g, world) for g in range(world)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == dim
            for (a0, an), (b0, _) in zip(blocks, blocks[1:]):
                assert a0 + an == b0
            assert max(b[1] for b in blocks) - min(b[1] for b in blocks) <= 1
    with pytest.raises(ValueError):
        sharded.shard_rows(4, 4, 4)


class OracleBackend(sharded.RoundBackend):
    """numpy/oracle implementation of the four per-round operations for one row block."""

    def __init__(self, rows_block: np.ndarray, row0: int):
        self.a, self.row0 = rows_block, row0

    def row_pass(self, e, s_slice):
        ev = e.numpy()
        n = self.a.shape[1]
        out = np.empty(self.a.shape[0], dtype=np.float32)
        for i in range(self.a.shape[0]):
            # same order as oracle.c row_dot(..., ORACLE_SUM_LANES16) with the scale vector
            prod = (self.a[i] * ev).astype(np.float32)
            nb = n & ~15
            acc = prod[:nb].reshape(-1, 16)
            lanes = np.zeros(16, dtype=np.float32)
            for row in acc:
                lanes = (lanes + row).astype(np.float32)
            for j, v in enumerate(prod[nb:]):
                lanes[j] = np.float32(lanes[j] + v)
            a8 = (lanes[:8] + lanes[8:]).astype(np.float32)
            a4 = (a8[:4] + a8[4:]).astype(np.float32)
            t = np.float32(np.float32(a4[0] + a4[2]) + np.float32(a4[1] + a4[3]))
            out[i] = t / ev[self.row0 + i]
        s_slice.copy_(torch.from_numpy(out))

    def find_max(self, s):
        return oracle.find_max(s.numpy())

    def stop(self, s, eps):
        return oracle.stop(s.numpy(), eps) == 1

    def update(self, s, m, e):
        ev = e.numpy()
        oracle.compute_eigen_vector(s.numpy(), m, ev)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dim, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. handle exchange: 64 opaque bytes per rank, returned in rank order
        handle = bytes([(rank * 37 + i) & 0xFF for i in range(64)])
        table = sharded.exchange_handles(handle, rank, world)
        assert len(table) == world and table[rank] == handle
        for g in range(world):
            assert table[g] == bytes([(g * 37 + i) & 0xFF for i in range(64)])
        # 2. the collective round loop on this rank's rows
        row0, rows = sharded.shard_rows(dim, rank, world)
        block = oracle.uniform(dim, seed, row0, rows) + np.float32(0.1)
        lam, e, it = sharded.collective_round_loop(OracleBackend(block, row0), dim, rank, world)
        out.put((rank, lam, e.numpy().copy(), it))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dim", [(2, 96), (3, 50)])
def test_collective_round_loop_over_gloo_matches_single_rank_oracle(world, dim):
    seed = 0x5EED0001
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dim, seed, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mat = oracle.uniform(dim, seed) + np.float32(0.1)
    o_val, o_vec, _, o_it = oracle.similarity_transform(mat, form=oracle.FORM_READONLY)
    for rank, lam, e, it in results:
        # every rank holds the same answer, bit for bit equal to the unsharded oracle
        assert it == o_it
        assert np.float32(lam) == o_val
        assert np.array_equal(e, o_vec)
