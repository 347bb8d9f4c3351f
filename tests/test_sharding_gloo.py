"""CPU test of the N > 1 data-parallel path with world_size 2 over gloo: contiguous
shards, no collective on the compute path, host-side in-order gather of the transcripts."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions_exactly():
    from nobs_whisper_b200.sharding import shard_range
    for n in (0, 1, 7, 64, 120, 121):
        for world in (1, 2, 3, 4, 8):
            parts = [list(shard_range(n, world, r)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


class FakeEngine:
    """Stands in for WhisperEngine.transcribe_batch on a box without a GPU."""

    def __init__(self):
        self.seen = []

    def transcribe_batch(self, audios, language=None, vocabulary=None, beam_size=0):
        self.seen.extend(len(a) for a in audios)
        return [f"{language}:{len(a)}" for a in audios]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nobs_whisper_b200.sharding import shard_range, transcribe_sharded
    audios = [[0.0] * (100 + i) for i in range(n_items)]
    eng = FakeEngine()
    texts = transcribe_sharded(eng, audios, language="en")
    q.put((rank, texts, eng.seen, list(shard_range(n_items, world, rank))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 2, 1])
def test_two_ranks_transcribe_their_shards_and_gather_in_order(n_items):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [f"en:{100 + i}" for i in range(n_items)]
    for rank, texts, seen, shard in results:
        assert texts == want                                  # every rank holds the full, ordered result
        assert seen == [100 + i for i in shard]               # ... but only computed its own shard
