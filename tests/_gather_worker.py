"""Worker of test_render_gather_over_ipc: one rank of an image-sharded render whose exchange is fused into the resolve
kernel (ptb_render_gather storing into the peers' images through CUDA IPC mappings).  Launched by torchrun with the
gloo backend for the handle exchange and the barrier; PTB_TEST_ONE_GPU=1 puts every rank on cuda:0 (IPC between
processes works on one device too, which is what the single-GPU test box offers)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oclpathtracer_b200 as pt  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if os.environ.get("PTB_TEST_ONE_GPU") == "1" else int(os.environ.get("LOCAL_RANK", 0))
    dist.init_process_group("gloo")
    tris, mats = pt.load_model(os.path.join(ROOT, "data", "cornellbox.bin"))
    dev = pt.Device(local)
    scene = dev.scene(tris, mats)
    w, h, block = 200, 75, 64  # 15000 pixels: the last block is ragged
    nbytes = w * h * 16
    full = dev.buffer(nbytes)
    full.clear()
    dev.sync()
    handles = [None] * world
    dist.all_gather_object(handles, full.ipc_export())
    peers = [dev.ipc_import(handles[r], nbytes) for r in range(world) if r != rank]
    ok = True
    for accum, chunks in ((pt.ACCUM_LINEAR, [(0, 3)]), (pt.ACCUM_REFERENCE, [(0, 2), (2, 3)])):
        full.clear()
        dev.sync()
        dist.barrier()
        for first, n in chunks:  # REFERENCE accumulation resumes from the state stored in the full image
            prm = pt.default_params(width=w, height=h, first_frame=first, n_frames=n, mode=pt.MODE_PATH, accum=accum, max_depth=6,
                                    shard_index=rank, shard_count=world, shard_block=block)
            dev.render_gather(scene, prm, full, peers)
            dev.sync()
            dist.barrier()  # every rank's stores have landed in every image
        got = full.read(np.uint32)
        ref = dev.buffer(nbytes)
        ref.clear()
        for first, n in chunks:
            dev.render(scene, pt.default_params(width=w, height=h, first_frame=first, n_frames=n, mode=pt.MODE_PATH, accum=accum,
                                                max_depth=6), ref)
        want = ref.read(np.uint32)
        ref.close()
        same = bool(np.array_equal(got, want))
        print(f"rank {rank} accum {accum}: identical={same}", flush=True)
        ok = ok and same
        dist.barrier()
    for b in peers:
        b.close()
    dist.barrier()  # nobody frees an image that a peer still has mapped and in use
    full.close(); scene.close(); dev.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
