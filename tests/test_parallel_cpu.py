"""CPU, gloo, world_size 2: the N>1 host logic (sharding + the token gather, the path's only collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from handwritten_math_ocr_api_b200.parallel import gather_tokens, gather_tokens_device, generate_sharded, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


class _FakeModel:
    """Stands in for the engine on CPU: row i decodes to [sos, i, i, ..., eos] with i % 5 + 1 tokens."""
    pad_id, sos_id, eos_id = 0, 1, 2

    def generate(self, images, max_len=None):
        ids = images[:, 0].long()
        steps = int((ids % 5 + 2).max()) if len(ids) else 1
        out = torch.full((len(ids), 1 + steps), self.pad_id, dtype=torch.int64)
        out[:, 0] = self.sos_id
        for r, i in enumerate(ids.tolist()):
            n = i % 5 + 1
            out[r, 1:1 + n] = 100 + i
            out[r, 1 + n] = self.eos_id
        return out, steps, None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        images = torch.arange(n_images, dtype=torch.float32).reshape(n_images, 1)
        full = generate_sharded(_FakeModel(), images)
        # ragged direct gather: rank r holds r+1 rows of r+2 columns
        mine = torch.full((rank + 1, rank + 2), rank + 10, dtype=torch.int64)
        rag = gather_tokens(mine, pad_id=0)
        # stream-ordered variant: equal batches, full-width buffers, per-rank step counters
        tok = torch.full((2, 5), rank + 20, dtype=torch.int64)
        dev_tok, dev_steps = gather_tokens_device(tok, torch.tensor([rank + 3], dtype=torch.int32))
        assert dev_tok.shape == (2 * world, 5) and dev_tok[2 * rank].tolist() == [rank + 20] * 5
        assert dev_steps.tolist() == [3 + r for r in range(world)]
        q.put((rank, full.tolist(), rag.tolist()))       # plain lists: no shared-memory handles in the queue
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [7, 8])
def test_generate_sharded_world2_gloo(n_images):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    [p.join(60) for p in procs]
    single = _FakeModel().generate(torch.arange(n_images, dtype=torch.float32).reshape(n_images, 1))[0]
    for rank, full, rag in res:
        full, rag = torch.tensor(full), torch.tensor(rag)
        assert full.shape[0] == n_images
        w = min(full.shape[1], single.shape[1])
        assert torch.equal(full[:, :w], single[:, :w])                     # N-GPU result == 1-GPU result
        assert (full[:, w:] == 0).all() and (single[:, w:] == 0).all()
        assert rag.shape == (3, 3) and rag[0].tolist() == [10, 10, 0] and rag[2].tolist() == [11, 11, 11]
