"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Each rank assembles the residual of its block with the CUDA path, P^T sums the interface dofs on the
owner over NCCL; owned dofs must match the CPU oracle on the undivided mesh to 1e-12."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
    ctx = M.Context(local)
    lspace = dict(blk["space"], mode=O.GRAD)
    gm = M.Mesh(ctx, blk["mesh"])
    gs = M.Space(ctx, gm, lspace)
    gi = M.Integrator(ctx, [(gs, O.GRAD)], fs.madb(ctx))
    dev = torch.device("cuda", local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    with torch.cuda.stream(stream):
        x = torch.from_numpy(xg[blk["l2g"]]).to(dev)
        y = torch.empty_like(x)
        gi.mult(x, y)
        ex = P.SharedDofExchange(blk["l2g"], blk["candidates"], dev, ctx=ctx)
        ex.reduce_to_owner(y)
        torch.cuda.synchronize()
        owned = ex.owned_mask(lspace["ndofs"])
        err = np.max(np.abs(y.cpu().numpy()[owned] - y_glob[blk["l2g"]][owned])) / np.max(np.abs(y_glob))
        ex.broadcast_from_owner(y)
        torch.cuda.synchronize()
        err2 = np.max(np.abs(y.cpu().numpy() - y_glob[blk["l2g"]])) / np.max(np.abs(y_glob))
    t = torch.tensor([err, err2], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("mgpu_check world=%d  P^T rel err %.2e  P rel err %.2e  %s" % (world, t[0].item(), t[1].item(),
                                                                              "OK" if t.max().item() <= 1e-12 else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.max().item() <= 1e-12 else 1)


if __name__ == "__main__":
    main()
