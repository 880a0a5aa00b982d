#!/usr/bin/env python
"""Spatial sharding of ONE large image over the ranks of a torchrun job (NCCL): every rank enhances its row band
(72-row halos), the pooled statistics are all-reduced inside the forward, rank 0 compares the gathered result with
the unsharded forward.   torchrun --nproc-per-node 2 tools/spatial_shard_check.py [--h 2448 --w 3264]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from imageenhancement_mp_b200 import dist as idist, synth, weights, model_library as ml

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=1024); ap.add_argument("--w", type=int, default=1536)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
rank, world, local = idist.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
params = dict(synth.DEFAULT_PARAMS)
W = weights.init_weights(weights.simplemodel_layers(params), scheme="stress")
model = ml.Simplemodel(params, weights=W, device=dev)
x, _ = synth.make_batch(1, a.h, a.w, params, seed=3)          # every rank generates the same image
sh = idist.spatial_shards(a.h, world)[rank]
s0, s1 = sh["slab"]
xs = x[:, s0:s1].to(dev).contiguous()
out, bas, ob = model.call_spatial_shard(xs, sh, a.h)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
idist.barrier(); e0.record()
for _ in range(a.reps):
    out, bas, ob = model.call_spatial_shard(xs, sh, a.h)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.reps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# gather the bands on rank 0 (bands may differ by one stride in height: broadcast each from its owner)
bands = [torch.empty(1, d["own"][1] - d["own"][0], a.w, 5, device=dev) for d in idist.spatial_shards(a.h, world)]
bands[rank].copy_(out)
if world > 1:
    for i, b in enumerate(bands):
        dist.broadcast(b, src=i)
if rank == 0:
    full = torch.cat(bands, dim=1)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref = model(x.to(dev))[0]
    torch.cuda.synchronize(); t0.record()
    for _ in range(a.reps):
        ref = model(x.to(dev))[0]
    t1.record(); torch.cuda.synchronize()
    print(json.dumps({"image": [a.h, a.w], "ranks": world, "max_abs_diff_vs_unsharded": float((full - ref).abs().max()),
                      "ms_sharded_max_over_ranks": float(ms), "ms_unsharded_one_gpu": t0.elapsed_time(t1) / a.reps,
                      "halo_rows": idist.SPATIAL_HALO}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
