"""World-size-2 gloo tests (CPU) of the data-parallel host logic: gather ordering, the differentiable all-gather whose
backward is a reduce-scatter (passl/distributed/nn/functional.py:100-127), gradient mean all-reduce (sync_utils.py:18-43), and
'sharded == unsharded' for the global-negative losses using the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from passl_b200 import distributed as D
    from oracle import contrastive as OC
    try:
        assert D.get_world_size() == world and D.get_rank() == rank
        # 1. concat_all_gather: rank-major order, no grad
        x = torch.full((3, 4), float(rank)) + torch.arange(3).float()[:, None]
        g = D.concat_all_gather(x)
        assert g.shape == (3 * world, 4) and not g.requires_grad
        for r in range(world):
            assert torch.equal(g[3 * r:3 * r + 3], torch.full((3, 4), float(r)) + torch.arange(3).float()[:, None])
        # 2. differentiable all_gather: backward = reduce-scatter(sum) of the gathered gradient
        z = (torch.arange(6).float().reshape(3, 2) + 10 * rank).requires_grad_(True)
        zg = D.all_gather(z)
        w = torch.arange(zg.numel()).float().reshape(zg.shape) * (rank + 1)
        (zg * w).sum().backward()
        base = torch.arange(3 * world * 2).float().reshape(3 * world, 2)[3 * rank:3 * rank + 3]
        expect = base * sum(r + 1 for r in range(world))
        assert torch.allclose(z.grad, expect), (z.grad, expect)
        # 3. grad_sync: sum all-reduce on the flat buffer, mean folded into the optimizer's grad_scale = 1/world
        class S:
            pass
        st = S()
        st.grad = torch.full((8,), float(rank + 1))
        D.grad_sync(st)
        assert torch.allclose(st.grad, torch.full((8,), float(sum(r + 1 for r in range(world)))))
        # 4. MoCo v3 style global negatives: labels arange(N) + N*rank; mean over ranks == unsharded loss
        rng = np.random.RandomState(0)
        q_all, k_all = rng.randn(4 * world, 16), rng.randn(4 * world, 16)
        kg = D.concat_all_gather(torch.from_numpy(k_all[4 * rank:4 * rank + 4]))
        assert np.allclose(kg.numpy(), k_all)
        loss_r, _, labels = OC.mocov3_contrastive_loss(q_all[4 * rank:4 * rank + 4], kg.numpy(), 0.2, rank=rank)
        assert np.array_equal(labels, np.arange(4) + 4 * rank)
        t = torch.tensor([loss_r])
        dist.all_reduce(t)
        full, _, _ = OC.mocov3_contrastive_loss(q_all, k_all, 0.2, rank=0)
        assert abs(t.item() / world - full) < 1e-12
        # 5. shuffle-BN protocol (moco.py:107-152): one permutation shared by all ranks, un-shuffle restores every rank's own rows
        from passl_b200.modeling.architectures.moco import MoCo
        torch.manual_seed(100 + rank)                      # ranks draw DIFFERENT permutations; rank 0's must win (broadcast)
        xs = (torch.arange(5).float()[:, None] + 100 * rank).repeat(1, 3)
        sh, idx_un = MoCo._batch_shuffle_ddp(None, xs)
        all_sh = D.concat_all_gather(sh)
        assert sorted(all_sh[:, 0].tolist()) == sorted(D.concat_all_gather(xs)[:, 0].tolist())   # a permutation of the global batch
        idx_all = [torch.zeros_like(idx_un) for _ in range(world)]
        dist.all_gather(idx_all, idx_un)
        assert all(torch.equal(i, idx_un) for i in idx_all)
        back = MoCo._batch_unshuffle_ddp(None, sh, idx_un)
        assert torch.equal(back, xs)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        ret[rank] = repr(e)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)
