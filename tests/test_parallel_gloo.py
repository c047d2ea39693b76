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
        # the halo route bench.py takes (lowest_rank_owner + halo_lists + HaloExchange, CPU tensors through gloo): same owners,
        # same values after P^T and P
        owner = P.lowest_rank_owner(blk["l2g"], blk["candidates"], lspace["ndofs"], rank, world)
        assert np.array_equal(owner == rank, owned)
        hx = P.HaloExchange(*P.halo_lists(blk["l2g"], owner, rank, world))
        y2 = torch.from_numpy(y_loc.copy())
        hx.reverse(y2)
        hx.forward(y2)
        assert torch.equal(y2, y)
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


def _worker_ghost(rank, world, port, n, out):
    """ex4 block system (H1 p3 x L2 p1, FermiDirac PG functional) on the overlapping partition: after P (halo of the
    state and of psi_k) the rows of OWNED dofs of the local assembly -- residual and Jacobian -- equal those of the
    undivided mesh (this rank's rows of P^T A P, ex4.cpp:136,169,190), with no exchange of matrix data."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        import spec as S
        from mfem_ad_b200 import meshgen as G, parallel as P
        from oracle import oracle as O
        order = 2
        blk = P.cartesian_block_ghost(rank, world, n, order + 1, order - 1)
        px, py = blk["px"], blk["py"]
        gmesh = G.cartesian_mesh((px * n, py * n), lengths=(float(px), float(py)))
        gh1 = G.h1_space(gmesh, order + 1, mode=O.VALUE | O.GRAD)
        gl2 = G.l2_space(gmesh, order - 1, mode=O.VALUE)
        rng = np.random.default_rng(5)
        xg = rng.uniform(-1, 1, gh1["ndofs"] + gl2["ndofs"])
        pk = rng.normal(0, 1, gl2["ndofs"])
        fs = S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.4)
        qo = 3 * order + 3
        gform = O.OracleForm(gmesh, [gh1, gl2], fs.oracle(), quad_order=qo, params=[dict(type=O.PRM_GF, size=1, data=pk, space=gl2)])
        y_glob = gform.mult(xg)
        rp, ci, vg = gform.grad(xg)
        Kg = sp.csr_matrix((vg, ci, rp), shape=(xg.size,) * 2)
        # local vectors: owned values from the global state, copies filled by the exchange
        h1, l2 = dict(blk["h1"], mode=O.VALUE | O.GRAD), dict(blk["l2"], mode=O.VALUE)
        nh, nl = h1["ndofs"], l2["ndofs"]
        l2g = np.concatenate([blk["l2g_h1"], gh1["ndofs"] + blk["l2g_l2"]])
        owner = np.concatenate([blk["owner_h1"], blk["owner_l2"]])
        own, ghost = P.halo_lists(l2g, owner, rank, world)
        ex = P.HaloExchange(own, ghost)
        x = torch.from_numpy(np.where(owner == rank, xg[l2g], -77.0))
        ex.forward(x)
        assert np.array_equal(x.numpy(), xg[l2g])
        ownl, ghostl = P.halo_lists(blk["l2g_l2"], blk["owner_l2"], rank, world)
        exl = P.HaloExchange(ownl, ghostl)
        pkl = torch.from_numpy(np.where(blk["owner_l2"] == rank, pk[blk["l2g_l2"]], -55.0))
        exl.forward(pkl)
        lform = O.OracleForm(blk["mesh"], [h1, l2], fs.oracle(), quad_order=qo,
                             params=[dict(type=O.PRM_GF, size=1, data=pkl.numpy(), space=l2)])
        y = lform.mult(x.numpy())
        rpl, cil, vl = lform.grad(x.numpy())
        Kl = sp.csr_matrix((vl, cil, rpl), shape=(nh + nl,) * 2)
        mine = owner == rank
        err_y = np.max(np.abs(y[mine] - y_glob[l2g][mine])) / np.max(np.abs(y_glob))
        rows = np.nonzero(mine)[0]
        # owned rows, all global columns: scatter the local columns to their global ids
        Kl_g = sp.csr_matrix((Kl[rows].tocoo().data, (Kl[rows].tocoo().row, l2g[Kl[rows].tocoo().col])), shape=(rows.size, xg.size))
        err_k = abs(Kl_g - Kg[l2g[rows]]).max() / np.max(np.abs(vg))
        # P^T of a non-overlapping assembly: contributions of the own elements only, summed on the owners
        cnt = torch.zeros(xg.size, dtype=torch.float64)
        cnt[torch.from_numpy(l2g[mine])] = 1.0
        dist.all_reduce(cnt)
        one = torch.ones(nh + nl, dtype=torch.float64)
        ex.reverse(one)
        out[rank] = (float(err_y), float(err_k), bool(torch.all(cnt == 1.0)), float(one.max()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ghost_layer_partition_gives_complete_owned_rows(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_ghost, args=(world, _free_port(), 3, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err_y, err_k, one_owner, mult = out[r]
        assert err_y <= 1e-13 and err_k <= 1e-13 and one_owner
    # a dof owned by rank 0 at the corner of the 2x2 grid receives copies from 3 ranks: 1 + 3
    assert out[0][3] == (4.0 if world == 4 else 2.0)
