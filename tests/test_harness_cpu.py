"""CPU: host logic of the multi-GPU harness with world_size-2 gloo processes (sharding + the single all-gather),
and the torch restatement of the inferencer pre/post-processing against the oracle/golden waveform."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fsn_oracle as O


def test_shard_range_partitions():
    from fsnplus_b200.inference import shard_range
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeModel:
    """Stands in for the CUDA model in the CPU harness test: returns the golden reference mask."""

    def __init__(self, crm):
        self.crm = crm

    def __call__(self, mag, real, imag):
        return self.crm[: mag.size(0)]


def test_pre_post_processing_matches_reference_waveform(golden):
    from fsnplus_b200.inference import enhance_batch
    g = golden("plus_default")
    clips = torch.from_numpy(O.synth_clips(1))
    enh = enhance_batch(_FakeModel(torch.from_numpy(g["out"])), clips)
    assert O.rel_l2(enh.numpy(), g["enhanced"]) < 1e-5


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fsnplus_b200.inference import shard_range, all_gather_enhanced
    lo, hi = shard_range(n_items, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 5)
    out = all_gather_enhanced(local, n_items)
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _run(n_items, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_all_gather_even_and_ragged_gloo():
    for n, port in ((8, 29611), (7, 29612)):
        out = _run(n, port)
        assert out.shape == (n, 5)
        assert np.array_equal(out[:, 0], np.arange(n, dtype=np.float32))
