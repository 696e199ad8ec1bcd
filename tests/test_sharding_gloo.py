"""world_size-2 (and 3) gloo runs of the multi-GPU host logic on CPU: contiguous channel
shards, per-rank banks, max-over-ranks timing reduction and the PCM gather. The compute
stand-in is the host emulation of the kernels (test harness); the expected result is the
oracle over the unsharded bank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, nbytes, out_path):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import _emu as E
    from rtlsdrdiags_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_total, rank, world)
    # every rank derives the same global tables, then takes its slice (as bench.py does)
    modes = synth.modes_for("mixed", n_total).numpy()
    iq = synth.noise_bank(n_total, nbytes, 7, "cpu").numpy()
    local = np.zeros((hi - lo, nbytes // 64), dtype=np.int16)
    for kind, ms in [(E.KIND_AM, [1]), (E.KIND_FM, [2]), (E.KIND_WBFM, [3]), (E.KIND_SSB, [4, 5])]:
        idx = [c for c in range(lo, hi) if modes[c] in ms]
        if not idx:
            continue
        bank = E.EmuBank(kind, len(idx), G=3, NT=64)
        gain = {E.KIND_AM: 300.0, E.KIND_FM: np.float32(64000 / (2 * np.pi)),
                E.KIND_WBFM: np.float32(256000 / (2 * np.pi)), E.KIND_SSB: 300.0}[kind]
        for i, c in enumerate(idx):
            bank.scale[i] = E.scale_for(kind, gain)
            bank.lsb[i] = 1 if modes[c] == 4 else 0
        pcm = bank.run(iq[idx], E.FMT_U8)
        for i, c in enumerate(idx):
            local[c - lo] = pcm[i]
    # device-timed region stand-in: the job's time is the slowest rank's
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 10.0 + world - 1
    full = sharding.gather_pcm(torch.from_numpy(local), n_total, dst=0)
    if rank == 0:
        np.save(out_path, full.numpy())
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 10), (3, 11)])
def test_sharded_bank_equals_unsharded_oracle(tmp_path, world, n_total):
    import _oracle as O
    from rtlsdrdiags_b200 import sharding, synth
    nbytes = 8192
    out = str(tmp_path / "pcm.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_total, nbytes, out), nprocs=world, join=True)
    got = np.load(out)
    modes = synth.modes_for("mixed", n_total).numpy()
    iq = synth.noise_bank(n_total, nbytes, 7, "cpu").numpy()
    exp, _ = O.oracle_bank(modes, iq, 32768, 1)
    assert np.array_equal(got, exp)
    sizes = sharding.shard_sizes(n_total, world)
    assert sum(sizes) == n_total and max(sizes) - min(sizes) <= 1


def test_shard_ranges_tile_the_bank():
    from rtlsdrdiags_b200 import sharding
    for n in [1, 7, 1024, 65536]:
        for w in [1, 2, 4, 8]:
            edges = [sharding.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
