"""world_size-2 gloo tests (CPU) of the host-side logic of the N>1 path: the halo exchange protocol the
engine implements with NCCL (forward: owner rows -> halo rows in (owner, gid) order; backward: halo adjoints
accumulate into owners) and the unique-id broadcast used by enable_distributed_mode."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distmlip_b200.structures import si_diamond
from oracle import graph_ref as G


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. unique-id broadcast exactly as CHGNet_Dist.enable_distributed_mode does it
        ids = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        assert ids[0] == bytes(range(128))
        # 2. halo protocol on the partition this rank would own
        atoms = si_diamond(4, nz=8, seed=21)
        o = G.GraphOracle(atoms.get_positions(), atoms.get_cell(), atoms.get_pbc().astype(np.int64), world, 5.0, 3.0,
                          True)
        own = o.owned(rank)
        halo = np.concatenate([o.from_list(rank, q_) for q_ in range(world) if q_ != rank])
        local = np.concatenate([own, halo])
        g2l = {g: i for i, g in enumerate(local)}
        feat = torch.zeros(len(local), 4)
        feat[: len(own)] = torch.tensor(own, dtype=torch.float32)[:, None] * torch.tensor([1.0, 2.0, 3.0, 4.0])
        # forward: pack to-lists, exchange with every neighbour
        reqs, recv = [], {}
        for q_ in range(world):
            if q_ == rank:
                continue
            to = o.to_list(rank, q_)
            send = feat[[g2l[g] for g in to]].contiguous()
            recv[q_] = torch.empty(len(o.from_list(rank, q_)), 4)
            reqs.append(dist.isend(send, q_))
            reqs.append(dist.irecv(recv[q_], q_))
        for r in reqs:
            r.wait()
        off = len(own)
        for q_ in range(world):
            if q_ == rank:
                continue
            feat[off: off + len(recv[q_])] = recv[q_]
            off += len(recv[q_])
        expect = torch.tensor(local, dtype=torch.float32)[:, None] * torch.tensor([1.0, 2.0, 3.0, 4.0])
        assert torch.equal(feat, expect)
        # backward: halo adjoints go home and accumulate
        g = torch.ones(len(local), 1)
        reqs, back = [], {}
        off = len(own)
        for q_ in range(world):
            if q_ == rank:
                continue
            nfrom = len(o.from_list(rank, q_))
            back[q_] = torch.empty(len(o.to_list(rank, q_)), 1)
            reqs.append(dist.isend(g[off: off + nfrom].contiguous(), q_))
            reqs.append(dist.irecv(back[q_], q_))
            off += nfrom
        for r in reqs:
            r.wait()
        for q_ in range(world):
            if q_ == rank:
                continue
            g[[g2l[x] for x in o.to_list(rank, q_)]] += back[q_]
        g[len(own):] = 0
        total = g.sum()
        dist.all_reduce(total)
        # every local copy (owned + halo, on all ranks) contributed exactly once
        n_copies = torch.tensor([float(len(local))])
        dist.all_reduce(n_copies)
        assert total.item() == n_copies.item()
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_halo_protocol_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_enable_distributed_mode_from_one_process_asks_for_a_group():
    """the reference's own call (examples/chgnet_example.ipynb cell 1): one process, several GPUs -> a single-process
    group (b2m_create with ndev = len(gpus)); without a GPU that fails loudly in b2m_create, never on the host side"""
    from distmlip_b200 import _lib
    from distmlip_b200.implementations.matgl import CHGNet_Dist
    from tests._util import make_model

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_group.py")
    m = CHGNet_Dist.from_existing(make_model())
    with pytest.raises(_lib.B2MError) as ei:
        m.enable_distributed_mode([0, 1])
    assert "no CUDA device" in str(ei.value)
