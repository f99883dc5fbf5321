"""world_size-2 checks of the multi-GPU host logic over gloo (CPU): sharding covers every clip once,
the max-over-ranks timing rule, and the ensemble/AENS reduce hooks."""
import os
import socket

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


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from i2v_b200 import dist as D
    r, _, w = D.init_from_env(backend="gloo")
    clips = D.clip_shard(7, r, w)
    gathered = [None] * w
    dist.all_gather_object(gathered, clips)
    # ensemble hook: every rank owns some cosine rows and a partial input gradient
    cos = torch.zeros(4, 6)
    cos[r * 2:(r + 1) * 2] = float(r + 1)
    g = torch.full((2, 3, 4, 4), float(r + 1))
    hook = D.ReduceHook(sum_grad=True, sum_cos=True)
    hook.grad(g)
    hook.cos_rows(cos)
    slow = D.max_over_ranks(10.0 + r)
    total = D.sum_over_ranks(100.0 * (r + 1))
    D.barrier()
    if r == 0:
        torch.save(dict(gathered=gathered, cos=cos, g=g, slow=slow, total=total), out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert sorted(sum(res["gathered"], [])) == list(range(7))
    assert res["slow"] == 11.0 and res["total"] == 300.0
    assert torch.equal(res["g"], torch.full((2, 3, 4, 4), 3.0))
    assert torch.equal(res["cos"][:2], torch.full((2, 6), 1.0)) and torch.equal(res["cos"][2:], torch.full((2, 6), 2.0))
