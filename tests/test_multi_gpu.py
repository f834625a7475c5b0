"""Two-GPU parity (NCCL): the frame-sharded decomposition of one movie equals the single-GPU decomposition of the same
movie with the same random draws.  Skipped on boxes with fewer than two GPUs (the 1-GPU round-end run); run with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    import torch.distributed as dist

    import localmd_b200
    from synth import make_movie

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        T, d1, d2 = 4096, 64, 72
        movie = make_movie(T, d1, d2, n_cells=8, seed=11)  # numpy (T, d1, d2): every rank reads only its own frames
        kw = dict(block_sizes=[16, 16], frame_range=1500, max_components=10, background_rank=3, seed=5, rank_prune=True)
        det = {}
        arr = localmd_b200.localmd_decomposition(movie, device=dev, group=dist.group.WORLD, details=det, **kw)
        if rank == 0:
            det1 = {}
            ref = localmd_b200.localmd_decomposition(movie, device=dev, details=det1, **kw)
            frames = [0, 1023, 1024, 2047, 2048, 4095]
            np.savez(out_path, mean=arr.mean_img, mean1=ref.mean_img, std=arr.var_img, std1=ref.var_img, ranks=det["ranks"],
                     ranks1=det1["ranks"], s=arr.s, s1=ref.s, rec=arr[frames], rec1=ref[frames], k=arr.v.shape[0], k1=ref.v.shape[0],
                     T=arr.v.shape[1])
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_frame_sharded_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp

    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    np.testing.assert_allclose(r["mean"], r["mean1"], rtol=1e-6)
    np.testing.assert_allclose(r["std"], r["std1"], rtol=1e-5)
    assert np.array_equal(r["ranks"], r["ranks1"])
    assert int(r["k"]) == int(r["k1"]) and int(r["T"]) == 4096
    lead = r["s1"] > 0.05 * r["s1"][0]
    np.testing.assert_allclose(r["s"][lead], r["s1"][lead], rtol=1e-4)
    rel = np.linalg.norm(r["rec"] - r["rec1"]) / np.linalg.norm(r["rec1"] - r["mean1"][None])
    assert rel < 1e-4, rel
