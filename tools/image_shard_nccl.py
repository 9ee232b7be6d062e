"""Image-sharded render over N GPUs with one NCCL all_gather, checked bit for bit.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/image_shard_nccl.py [c4] [frames]

Every rank renders the pixels gid with (gid // 64) % N == rank of the same frames, keeps a linear accumulator,
and the local frames are exchanged once with all_gather + un-interleave (oclpathtracer_b200/sharding.py).
Rank 0 also renders the whole image alone and compares: the sharded result must be IDENTICAL.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oclpathtracer_b200 as pt  # noqa: E402
from oclpathtracer_b200 import sharding  # noqa: E402
from bench import WORKLOADS, load_scene, make_params  # noqa: E402


def main():
    wl_name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[wl_name]
    stream = torch.cuda.current_stream()
    dev = pt.Device(local, stream=stream.cuda_stream)
    tris, mats, light = load_scene(pt, wl)
    scene = dev.scene(tris, mats)
    npix, block = wl["width"] * wl["height"], 64
    n_local = sharding.local_pixels(npix, rank, world, block)
    local_t = torch.zeros((n_local, 4), dtype=torch.float32, device="cuda")
    buf = dev.wrap(local_t.data_ptr(), local_t.numel() * 4)

    def prm(**kw):
        return make_params(pt, wl, first_frame=0, n_frames=frames, light_p1=light[0], light_ea=light[1], light_eb=light[2], **kw)

    p = prm(shard_index=rank, shard_count=world, shard_block=block)
    for _ in range(2):  # warm up the kernels and the NCCL communicator (first collective sets up channels)
        dev.render(scene, p, buf)
        sharding.gather_image(local_t, npix, rank, world, block)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(stream)
    ctr = dev.render(scene, p, buf)
    e1.record(stream)
    image = sharding.gather_image(local_t, npix, rank, world, block)
    e2.record(stream)
    torch.cuda.synchronize()
    ctr = dev.render(scene, p, buf, want_counters=True)
    t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2), float(ctr["rays_closest"] + ctr["rays_any"])], dtype=torch.float64, device="cuda")
    tmax, tsum = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    # ---- the same exchange fused into the resolve kernel: stores into every rank's full image over peer memory -------
    fused = None
    if world > 1:
        full = dev.buffer(npix * 16)
        full.clear()
        dev.sync()
        handles = [None] * world
        dist.all_gather_object(handles, full.ipc_export())
        peers = [dev.ipc_import(handles[r], npix * 16) for r in range(world) if r != rank]
        for _ in range(2):
            dev.render_gather(scene, p, full, peers)
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        g0.record(stream)
        dev.render_gather(scene, p, full, peers)
        g1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()  # all ranks' stores have landed everywhere
        wall_ms = (time.perf_counter() - t0) * 1e3
        tg = torch.tensor([g0.elapsed_time(g1), wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        fused_img = torch.from_numpy(full.read(np.float32).reshape(-1, 4)).cuda()
        fused = {"render_and_scatter_ms_max": float(tg[0]), "wall_ms_with_barrier_max": float(tg[1]),
                 "identical_to_all_gather_path": bool(torch.equal(fused_img.view(torch.int32), image.view(torch.int32)))}
        dist.barrier()
        for b in peers:
            b.close()
        dist.barrier()
        full.close()
    if rank == 0:
        full_t = torch.zeros((npix, 4), dtype=torch.float32, device="cuda")
        fbuf = dev.wrap(full_t.data_ptr(), full_t.numel() * 4)
        dev.render(scene, prm(), fbuf)
        torch.cuda.synchronize()
        identical = bool(torch.equal(full_t.view(torch.int32), image.view(torch.int32)))
        fbuf.close()
        print(json.dumps({"workload": wl_name, "n_gpus": world, "frames": frames, "render_ms_max": float(tmax[0]),
                          "all_gather_ms_max": float(tmax[1]), "Mrays_per_s": float(tsum[2]) / (float(tmax[0] + tmax[1]) * 1e-3) / 1e6,
                          "gathered_bytes": int(npix * 16), "bit_identical_to_single_gpu": identical, "scaling": "strong (one image)",
                          "fused_peer_memory_gather": fused}))
        assert identical and (fused is None or fused["identical_to_all_gather_path"])
    buf.close(); scene.close(); dev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
