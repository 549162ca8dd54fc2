#!/usr/bin/env python
"""FWI objective and adjoint-state gradient (src/fdtd/func_grad.jl:1-49) with the B200 engine (needs a CUDA device).

    pa   = SeisForwExpt(FdtdAcoustic{FullWave}(:forward_save); medium, ageom, srcwav, tgrid, rfields=[:vz])
    m    = get_modelvector(pa, [:invK, :rho])
    f    = lossvalue(m, L2DistLoss(), dobs, pa, mparams)
    gradient!(g, m, L2DistLoss(), dobs, pa, mparams)

Run with torchrun (one process per GPU) to shard the supersources and all-reduce the gradient over NCCL:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/fwi_gradient_2d.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geophyinv_jl_b200 as G
from geophyinv_jl_b200.host import dist as D, gallery

dist = D.init_process_group("nccl") if int(os.environ.get("WORLD_SIZE", 1)) > 1 else None
rank, local_rank, world = D.env_ranks()

# model = layered medium, "observed" data from the same medium with a +5 % vp box; 8 supersources, records :vz
kw, true = gallery.c4_fwi2d(nz=120, nx=300, nt=1200, nss=8, nr=64, fq=10.0)
pt = G.SeisForwExpt(G.FdtdAcoustic(), **{**kw, "medium": true}, nworker=world, rank=rank, device=local_rank)
G.update(pt)
dobs = [d.copy() for d in pt["data", 1]]

pa = G.SeisForwExpt(G.FdtdAcoustic("forward_save"), **kw, nworker=world, rank=rank, device=local_rank)
if dist is not None:
    D.attach_nccl(pa, dist)                        # the engine's own communicator: one ncclAllReduce per parameter
m = pa.get_modelvector()                           # log-parameterised [invK; rho] on the un-extended grid
g = np.zeros_like(m)
loss = G.gradient(g, m, dobs, pa)                  # forward_save pass, adjoint pass, imaging, sum over shots (and ranks)
if rank == 0:
    half = g.size // 2
    print(f"loss {loss:.6e}; |g_invK|max {np.abs(g[:half]).max():.3e}, |g_rho|max {np.abs(g[half:]).max():.3e}; "
          f"{pa.last_run_ms:.0f} ms on the device for both passes")
if dist is not None:
    dist.destroy_process_group()
