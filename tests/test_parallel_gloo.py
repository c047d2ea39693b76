"""Host logic of the N>1 path on CPU: world_size 2 and 4, gloo backend, 127.0.0.1.
Local residuals come from the CPU oracle on each rank's block; after P^T (reduce_to_owner) the owned
dofs must equal the oracle residual of the undivided mesh; P (broadcast_from_owner) makes copies agree."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, p, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spec as S
        from mfem_ad_b200 import meshgen as G, parallel as P
        from oracle import oracle as O
        blk = P.cartesian_block(rank, world, n, p)
        px, py = blk["px"], blk["py"]
        # the undivided mesh and the same smooth state on both
        gmesh = G.cartesian_mesh((px * n, py * n), lengths=(float(px), float(py)))
        gspace = G.h1_space(gmesh, p, mode=O.GRAD)
        xc = G.dof_coords(gmesh, gspace)
        xg = np.sin(1.3 * xc[:, 0]) * np.cos(0.7 * xc[:, 1]) + 0.1 * np.random.default_rng(1).uniform(-1, 1, gspace["ndofs"])
        fs = S.minsurf(2, 0.5)
        y_glob = O.OracleForm(gmesh, [gspace], fs.oracle()).mult(xg)
        lspace = dict(blk["space"], mode=O.GRAD)
        y_loc = O.OracleForm(blk["mesh"], [lspace], fs.oracle()).mult(xg[blk["l2g"]])
        ex = P.SharedDofExchange(blk["l2g"], blk["candidates"], "cpu")
        y = torch.from_numpy(y_loc.copy())
        ex.reduce_to_owner(y)
        owned = ex.owned_mask(lspace["ndofs"])
        err = np.max(np.abs(y.numpy()[owned] - y_glob[blk["l2g"]][owned])) / np.max(np.abs(y_glob))
        # every global dof has exactly one owner
        cnt = torch.zeros(gspace["ndofs"], dtype=torch.float64)
        cnt[torch.from_numpy(blk["l2g"][owned])] = 1.0
        dist.all_reduce(cnt)
        ex.broadcast_from_owner(y)
        err2 = np.max(np.abs(y.numpy() - y_glob[blk["l2g"]])) / np.max(np.abs(y_glob))
        out[rank] = (float(err), float(err2), bool(torch.all(cnt == 1.0)), len(ex.peers))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_shared_dof_exchange_matches_undivided_mesh(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 4, 2, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err, err2, one_owner, npeers = out[r]
        assert err <= 1e-13 and err2 <= 1e-13 and one_owner
        assert npeers == (1 if world == 2 else 3)  # 2x2: two edge neighbours + the corner neighbour
