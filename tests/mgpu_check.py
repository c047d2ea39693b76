"""Multi-GPU parity checks, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
(tests/test_gpu_multi.py launches this when the box has at least 2 GPUs.)

case "residual":  config-2 form on a NON-overlapping block partition: every rank assembles the residual of its block
                  with the CUDA path, P^T sums the interface dofs on the owner over NCCL (C ABI exchange, madb_exchange_*),
                  P sends them back; owned dofs against the CPU oracle on the undivided mesh.
case "block":     ex4 LVPP block system (H1 p3 x L2 p1) on the OVERLAPPING partition (one ghost layer): P for the state and
                  psi_k, fused residual + Jacobian on every rank, rows of owned dofs (= this rank's rows of P^T A P,
                  ex4.cpp:136,169,190) against the oracle on the undivided mesh; global L1 sum through madb_comm_allreduce_sum.
Everything to 1e-12 (norm-relative)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def case_residual(rank, world, local, ctx, comm):
    import mfem_ad_b200 as M
    import spec as S
    from mfem_ad_b200 import meshgen as G, parallel as P
    from oracle import oracle as O
    n, p = 12, 2
    blk = P.cartesian_block(rank, world, n, p)
    px, py = blk["px"], blk["py"]
    gmesh = G.cartesian_mesh((px * n, py * n), lengths=(float(px), float(py)))
    gspace = G.h1_space(gmesh, p, mode=O.GRAD)
    xc = G.dof_coords(gmesh, gspace)
    xg = np.sin(1.3 * xc[:, 0]) * np.cos(0.7 * xc[:, 1]) + 0.1 * np.random.default_rng(1).uniform(-1, 1, gspace["ndofs"])
    fs = S.minsurf(2, 0.5)
    y_glob = O.OracleForm(gmesh, [gspace], fs.oracle()).mult(xg)
    lspace = dict(blk["space"], mode=O.GRAD)
    gm = M.Mesh(ctx, blk["mesh"])
    gs = M.Space(ctx, gm, lspace)
    gi = M.Integrator(ctx, [(gs, O.GRAD)], fs.madb(ctx))
    dev = torch.device("cuda", local)
    # owner of a shared dof: the lowest rank holding it (non-overlapping partition)
    allg = [None] * world
    dist.all_gather_object(allg, blk["l2g"][blk["candidates"]])
    owner = np.full(lspace["ndofs"], rank, dtype=np.int64)
    for r in range(rank):
        owner[blk["candidates"][np.isin(blk["l2g"][blk["candidates"]], allg[r])]] = np.minimum(
            owner[blk["candidates"][np.isin(blk["l2g"][blk["candidates"]], allg[r])]], r)
    own, ghost = P.halo_lists(blk["l2g"], owner, rank, world)
    ex = P.HaloExchange(own, ghost, ctx=ctx, comm=comm)
    x = torch.from_numpy(xg[blk["l2g"]]).to(dev)
    y = torch.empty_like(x)
    gi.mult(x, y)
    ex.reverse(y)
    ctx.sync()
    mine = owner == rank
    err = np.max(np.abs(y.cpu().numpy()[mine] - y_glob[blk["l2g"]][mine])) / np.max(np.abs(y_glob))
    ex.forward(y)
    ctx.sync()
    err2 = np.max(np.abs(y.cpu().numpy() - y_glob[blk["l2g"]])) / np.max(np.abs(y_glob))
    # the torch.distributed path of round 1 gives the same bits
    y2 = torch.empty_like(x)
    gi.mult(x, y2)
    ex1 = P.SharedDofExchange(blk["l2g"], blk["candidates"], dev, ctx=ctx)
    ex1.reduce_to_owner(y2)
    ex1.broadcast_from_owner(y2)
    torch.cuda.synchronize()
    same = bool(torch.equal(y2, y))
    return max(err, err2), same


def case_block(rank, world, local, ctx, comm):
    import scipy.sparse as sp
    import mfem_ad_b200 as M
    import spec as S
    from mfem_ad_b200 import meshgen as G, parallel as P
    from oracle import oracle as O
    n, order = 6, 2
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
    h1, l2 = dict(blk["h1"], mode=O.VALUE | O.GRAD), dict(blk["l2"], mode=O.VALUE)
    nh, nl = h1["ndofs"], l2["ndofs"]
    l2g = np.concatenate([blk["l2g_h1"], gh1["ndofs"] + blk["l2g_l2"]])
    owner = np.concatenate([blk["owner_h1"], blk["owner_l2"]])
    dev = torch.device("cuda", local)
    ex = P.HaloExchange(*P.halo_lists(l2g, owner, rank, world), ctx=ctx, comm=comm)
    exl = P.HaloExchange(*P.halo_lists(blk["l2g_l2"], blk["owner_l2"], rank, world), ctx=ctx, comm=comm)
    x = torch.from_numpy(np.where(owner == rank, xg[l2g], -77.0)).to(dev)
    pkl = torch.from_numpy(np.where(blk["owner_l2"] == rank, pk[blk["l2g_l2"]], -55.0)).to(dev)
    gm = M.Mesh(ctx, blk["mesh"])
    gh, gl = M.Space(ctx, gm, h1), M.Space(ctx, gm, l2)
    gi = M.Integrator(ctx, [(gh, O.VALUE | O.GRAD), (gl, O.VALUE), (gl, O.VALUE, M.ROLE_PARAM)], fs.madb(ctx), quad_order=qo)
    y = torch.empty_like(x)
    vals = torch.empty(gi.nnz, dtype=torch.float64, device=dev)
    # both halos in flight together, then the assembly (all on the context stream)
    ex.begin(x)
    ex.end(x)
    exl.begin(pkl)
    exl.end(pkl)
    gi.set_param_field(2, pkl)
    gi.assemble(x, y, vals)
    ctx.sync()
    assert np.array_equal(x.cpu().numpy(), xg[l2g])
    mine = owner == rank
    yh, vh = y.cpu().numpy(), vals.cpu().numpy()
    err_y = np.max(np.abs(yh[mine] - y_glob[l2g][mine])) / np.max(np.abs(y_glob))
    rpl, cil = gi.pattern()
    Kl = sp.csr_matrix((vh, cil, rpl), shape=(nh + nl,) * 2)
    rows = np.nonzero(mine)[0]
    sub = Kl[rows].tocoo()
    Kl_g = sp.csr_matrix((sub.data, (sub.row, l2g[sub.col])), shape=(rows.size, xg.size))
    err_k = abs(Kl_g - Kg[l2g[rows]]).max() / np.max(np.abs(vg))
    # global weighted L1 norm (ex4.cpp:203-206) from the owned parts
    part = np.array([np.sum(np.abs(yh[mine]))])
    tot = comm.allreduce_sum(part)[0]
    err_s = abs(tot - np.sum(np.abs(y_glob))) / np.sum(np.abs(y_glob))
    return max(err_y, err_k, err_s), True


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mfem_ad_b200 as M
    from mfem_ad_b200 import parallel as P
    ctx = M.Context(local)
    comm = P.Comm(ctx)
    ok = True
    for name, fn in (("residual", case_residual), ("block", case_block)):
        err, flag = fn(rank, world, local, ctx, comm)
        t = torch.tensor([err, 0.0 if flag else 1.0], device=torch.device("cuda", local), dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        good = t[0].item() <= 1e-12 and t[1].item() == 0.0
        ok = ok and good
        if rank == 0:
            print("mgpu_check world=%d case=%s rel err %.2e %s" % (world, name, t[0].item(), "OK" if good else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
