"""N>1 path on CPU: two gloo ranks shard independent sequences with no data-path collective.

Each rank runs the CPU oracle on its own sequences (the GPU path does the same with the CUDA
emulator); the gathered per-sequence results must equal a single-process run over all
sequences — shard invariance, SURVEY.md §4 item 5 / §8e.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from v2e2v_b200.sharding import all_shards, owner_of, shard_sequences

N_SEQ, PACKS, T, H, W = 5, 2, 6, 12, 16


def _run_sequence(seq: int):
    from oracle.oracle import OracleEmulator
    from tests import gpu_common as gc
    cfg = dict(gc.CFG2, leak_rate_hz=0, shot_noise_rate_hz=0)
    pos, neg, _ = gc.make_planes(100 + seq, H, W, cfg)
    em = OracleEmulator(**cfg)
    em.set_planes(pos, neg)
    nev = 0
    for k, fr, tfr in gc.video(seq, PACKS, T, H, W):
        vox, n = em.forward(fr.numpy(), tfr)
        nev += n
    return np.concatenate([vox.reshape(-1), em.base_log_frame.reshape(-1), [float(nev)]])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = {s: _run_sequence(s) for s in shard_sequences(N_SEQ, rank, world)}
    # the only communication: timing-style barrier and a gather of the results for the check
    dist.barrier()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        out.put(merged)
    dist.destroy_process_group()


def test_shards_partition_the_sequences():
    for n in (0, 1, 5, 64):
        for w in (1, 2, 3, 8):
            shards = all_shards(n, w)
            assert sorted(i for s in shards for i in s) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
            for i in range(n):
                assert i in shards[owner_of(i, n, w)]


def test_two_ranks_match_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    merged = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(merged) == list(range(N_SEQ))
    for s in range(N_SEQ):
        np.testing.assert_array_equal(merged[s], _run_sequence(s))
