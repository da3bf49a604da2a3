"""In-kernel timeline of one distributed V-cycle (run under torchrun): every SpMV-family and exchange kernel stamps
%globaltimer (famg_ctx_set_option "trace"); the records of each rank go to gpurun_out/trace_<tag>_rank<r>.txt and a
summary (per kernel: start, duration; per exchange: wait for the producer, pack, wait for the neighbours) is printed.

    torchrun --nproc-per-node N scripts/dist_trace.py [GRID=256] [TAG=n2]       (FAMG_OVERLAP selects the mode)
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import faer_amg_b200 as F  # noqa: E402
from faer_amg_b200.distributed import (Comm, DistGeometricPartitioner, DistHierarchy, DistMat, DistMultigrid,  # noqa: E402
                                       fine_plane_splits)


def summarize(path, out):
    recs = {}
    desc = {}
    for line in open(path):
        p = line.rstrip("\n").split(" ", 4)
        i, kind, blk, t = int(p[0]), int(p[1]), int(p[2]), int(p[3])
        recs.setdefault(i, {}).setdefault(kind, []).append(t)
        desc[i] = p[4] if len(p) > 4 else ""
    if not recs:
        return
    t0 = min(min(v) for r in recs.values() for v in r.values())
    rows = []
    for i, r in recs.items():
        b, e = min(r.get(1, [0])), max(r.get(2, [0]))
        row = {"id": i, "start_us": (b - t0) / 1e3, "dur_us": (e - b) / 1e3, "desc": desc[i]}
        if 3 in r:
            row["sig_ok_us"] = (max(r[3]) - b) / 1e3
            row["packed_us"] = (max(r[4]) - b) / 1e3
            row["flag_ok_us"] = (max(r[5]) - b) / 1e3
        rows.append(row)
    rows.sort(key=lambda x: x["start_us"])
    end = max(x["start_us"] + x["dur_us"] for x in rows)
    with open(out, "w") as f:
        f.write(f"# {len(rows)} traced launches, span {end:.1f} us; times in us relative to the first stamp\n")
        f.write("# start  dur   [exchange: sig_ok packed flag_ok after its own start]  description\n")
        for x in rows:
            ex = f" sig {x['sig_ok_us']:7.1f} packed {x['packed_us']:7.1f} flags {x['flag_ok_us']:7.1f}" if "sig_ok_us" in x else ""
            f.write(f"{x['start_us']:9.1f} {x['dur_us']:8.1f}{ex}  {x['desc']}\n")
    return end


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    tag = sys.argv[2] if len(sys.argv) > 2 else f"n{world}"
    dims = (n, n, n)
    ctx = F.Context.default(local)
    comm = Comm.from_torch(ctx)
    rs = fine_plane_splits(dims, world)
    plane = n * n
    slab = F.gallery.poisson7_slab(ctx, n, n, n, int(rs[rank]) // plane, int(rs[rank + 1]) // plane)
    a0 = DistMat.from_slabs(comm, [slab], n ** 3)
    nn = [np.full(int(rs[rank + 1] - rs[rank]), 1.0 / np.sqrt(n ** 3))]
    dh = DistHierarchy(comm, a0, nn, DistGeometricPartitioner(dims), coarsest_dim=1000, replicate_below=4096)
    ctx.set_option("trace", 1)
    dmg = DistMultigrid.from_hierarchy(comm, dh)
    nloc = dmg.nloc
    b, z = F.DeviceMat.from_host(ctx, np.ones(nloc)), F.DeviceMat(ctx, nloc, 1)
    for _ in range(5):
        dmg.apply_dev(z, b)
    ctx.sync(); torch.cuda.synchronize(); dist.barrier()
    ctx.set_option("trace", 1)  # restart the record list
    dmg.apply_dev(z, b)
    ctx.sync()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    raw = os.path.join(ROOT, "gpurun_out", f"trace_{tag}_rank{rank}.txt")
    ctx.trace_dump(raw)
    end = summarize(raw, os.path.join(ROOT, "gpurun_out", f"timeline_{tag}_rank{rank}.txt"))
    print(f"[rank {rank}] traced cycle span {end:.1f} us", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
