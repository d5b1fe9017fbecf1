"""Multi-rank parity check, launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/run_multirank.py
Every rank drives one GPU (slab partition + NCCL halo exchange inside libb200mlip) and must reproduce the
single-partition oracle (the dist path's arithmetic is partition independent)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distmlip_b200.implementations.matgl import CHGNet_Dist, Potential_Dist  # noqa: E402
from distmlip_b200.structures import si_diamond  # noqa: E402
from oracle.chgnet_ref import potential_ref  # noqa: E402
from tests._util import make_model  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for seed, nxy in ((3, 4), (8, 3)):
        atoms = si_diamond(nxy, nz=4 * world, seed=seed)
        if seed == 3:
            dm = CHGNet_Dist.from_existing(make_model())
            dm.enable_distributed_mode(list(range(world)))
            pot = Potential_Dist(model=dm, data_mean=0.5, data_std=1.5, calc_site_wise=True)
        E, F, S, _, site = pot(atoms)
        c = dm._engine.counts()
        print(f"[rank {rank}] atoms {len(atoms)} own {c['n_own']} halo {c['n_halo']} bonds {c['n_bond_own']}+"
              f"{c['n_bond_halo']} angles {c['n_angles']}", flush=True)
        if rank == 0:
            Eo, Fo, So, siteo = potential_ref(make_model(), atoms, data_mean=0.5, data_std=1.5)
            de = abs(E.item() - Eo.item()) / len(atoms)
            df = (F - Fo).abs().max().item()
            ds = (S - So).abs().max().item()
            dsite = (site - siteo).abs().max().item()
            print(f"world {world} natoms {len(atoms)}: dE/atom {de:.2e} dF {df:.2e} dS {ds:.2e} dsite {dsite:.2e}",
                  flush=True)
            ok = ok and de < 2e-7 and df < 3e-6 and ds < 3e-6 and dsite < 5e-6
        # all ranks must hold identical (all-reduced) results
        t = torch.tensor(np.concatenate([[E.item()], F.numpy().ravel()]), device="cuda")
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        ok = ok and float((tmax - tmin).abs().max()) == 0.0
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIRANK", "PASS" if flag.item() == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
