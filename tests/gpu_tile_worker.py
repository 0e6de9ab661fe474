"""One rank of the multi-GPU tile test (launched by tests/test_gpu_tile.py through torch.distributed.run).
Runs buildingsegment_b200.slabs with the CUDA backend on this rank's chunk of a seeded cloud and saves what it got."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--case", default="block")
    ap.add_argument("--points", dest="n", type=int, default=200000)
    ap.add_argument("--halo", type=int, default=400)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # gloo: every rank on GPU 0 (a one-GPU box still exercises the N > 1 path end to end); nccl: one GPU per rank
    device = 0 if a.backend == "gloo" else local
    torch.cuda.set_device(device)
    if a.backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo")
    from buildingsegment_b200 import lib, slabs

    kw = {"n": a.n} if a.case != "grid_plane" else {"nx": 300, "ny": a.n // 300, "order": "shuffled"}
    xyz = getattr(cases, a.case)(**kw)
    w = np.array([1.0 + 0.5 * ((r * 7) % 3) for r in range(world)])
    cuts = np.concatenate([[0], np.round(np.cumsum(w) / w.sum() * len(xyz)).astype(np.int64)])
    cuts[-1] = len(xyz)
    dev = torch.device("cuda", device)
    chunk = torch.from_numpy(np.ascontiguousarray(xyz[cuts[rank]: cuts[rank + 1]])).to(dev)
    ctx = lib.Context(device)
    p = lib.default_params()
    be = slabs.CudaBackend(ctx, p)
    r = slabs.segment_tile(be, chunk, halo=a.halo)
    img = slabs.raster_tile(be, chunk, r, p.bin, p.bin_height, p.count_bias)
    P = r["partition"]
    np.savez(os.path.join(a.out, f"r{rank}.npz"), labels=r["labels"].cpu().numpy(), plane_idx=r["plane_idx"].cpu().numpy(),
             n_planes=r["n_planes"], halo=r["halo"], n_halo=r["n_halo"], c0=cuts[rank], c1=cuts[rank + 1],
             owned_gid=r["owned_gid"].cpu().numpy(), owned_rows=r["owned_rows"].cpu().numpy(),
             owned_normals=r["owned_normals"].cpu().numpy(), cuts=P.cuts, origin=P.origin,
             image=img["image"].cpu().numpy(), a=img["png_a"].cpu().numpy(), b=img["png_b"].cpu().numpy(), x0=img["x0"], W=img["W"],
             H=img["H"], th=img["ground_th"], launches=ctx.timings()["kernel_launches"])
    be.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
