"""CPU, world_size 2 over gloo: the multi-GPU host logic -- LPT sharding through the C ABI's
scheduler, per-rank compaction and the final gather of betas in block order.  The per-block
compute is stood in for by the oracle (a checker; the product has no CPU path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dbslmm_b200 import multigpu, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_fit(bed, n_ref, csr, sigma_s, n_obs):
    from oracle import oracle as O
    bs, bl, _, _ = O.est(bed, n_ref, n_obs, sigma_s, *csr, threads=1, mode=O.MODE_EXACT)
    return bs, bl


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.make_workload(42, [60, 0, 25, 90, 10, 33], 200, frac_large=0.03)
    owner = multigpu.plan_owner(w["s_off"], w["l_off"], w["n_ref"], world)
    sh = multigpu.shard(w, owner, rank)
    bs, bl = _oracle_fit(sh["bed"], w["n_ref"], (sh["s_off"], sh["s_pos"], sh["s_z"], sh["l_off"], sh["l_pos"], sh["l_z"]), 1e-4, 5000)
    full_s, full_l = multigpu.gather_betas(w, owner, rank, world, bs, bl, dist)
    if rank == 0:
        np.savez(out, s=full_s, l=full_l, owner=owner)
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    out = str(tmp_path / "r.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    w = synth.make_workload(42, [60, 0, 25, 90, 10, 33], 200, frac_large=0.03)
    bs, bl = _oracle_fit(w["bed"], 200, (w["s_off"], w["s_pos"], w["s_z"], w["l_off"], w["l_pos"], w["l_z"]), 1e-4, 5000)
    assert set(got["owner"].tolist()) == {0, 1}
    assert np.allclose(got["s"], bs, rtol=0, atol=0) and np.allclose(got["l"], bl, rtol=0, atol=0)
