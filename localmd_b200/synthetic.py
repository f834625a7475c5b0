"""Synthetic functional-imaging movies generated directly in HBM.

Y = mu + sigma_px * ( A.C + Bg.F + eps ): A Gaussian blobs, C spike trains convolved with exp(-n/15),
Bg.F a smooth low-rank background (2-D cosines x slow random walks), eps ~ N(0,1) i.i.d.,
sigma_px ~ U(0.5, 2), mu ~ U(100, 300).  Seeded per (seed, rank) so frame shards are reproducible."""
import math

import torch


def make_movie(T, d1, d2, n_cells=400, blob_sigma=(3.0, 5.0), bg_rank=2, seed=1234, device="cuda", dtype=torch.float32,
               frame_lo=0, frame_hi=None, chunk=1024):
    """Frames [frame_lo, frame_hi) of the (T, d1, d2) movie as a device tensor.  The spatial footprints,
    traces and per-pixel scales depend on `seed` only; the noise is drawn per 1024-frame chunk from
    seed + chunk index, so any shard of the same movie can be generated independently."""
    frame_hi = T if frame_hi is None else frame_hi
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    yy = torch.arange(d1, device=dev, dtype=torch.float32)[:, None]
    xx = torch.arange(d2, device=dev, dtype=torch.float32)[None, :]
    cy = torch.rand(n_cells, generator=g, device=dev) * d1
    cx = torch.rand(n_cells, generator=g, device=dev) * d2
    sg = blob_sigma[0] + (blob_sigma[1] - blob_sigma[0]) * torch.rand(n_cells, generator=g, device=dev)
    amp = 2.0 + 4.0 * torch.rand(n_cells, generator=g, device=dev)
    a = amp[:, None, None] * torch.exp(
        -((yy[None] - cy[:, None, None]) ** 2 + (xx[None] - cx[:, None, None]) ** 2) / (2 * sg[:, None, None] ** 2)
    )
    a = a.reshape(n_cells, d1 * d2)
    spikes = (torch.rand((n_cells, T), generator=g, device=dev) < 0.01).to(torch.float32)
    kern = torch.exp(-torch.arange(64, device=dev, dtype=torch.float32) / 15.0)
    c = torch.nn.functional.conv1d(spikes[:, None, :], kern.flip(0)[None, None, :], padding=63)[:, 0, :T]
    bgs = []
    walks = torch.cumsum(torch.randn((bg_rank, T), generator=g, device=dev), dim=1) * 0.05
    for b in range(bg_rank):
        img = torch.cos(math.pi * (b + 1) * yy / d1) * torch.cos(math.pi * (b + 0.5) * xx / d2)
        bgs.append(img.reshape(-1))
    bgm = torch.stack(bgs) if bg_rank else torch.zeros((0, d1 * d2), device=dev)
    sigma_px = 0.5 + 1.5 * torch.rand(d1 * d2, generator=g, device=dev)
    mu = 100.0 + 200.0 * torch.rand(d1 * d2, generator=g, device=dev)
    out = torch.empty((frame_hi - frame_lo, d1 * d2), dtype=dtype, device=dev)
    for f0 in range(frame_lo - frame_lo % chunk, frame_hi, chunk):
        f1 = min(f0 + chunk, T)
        gn = torch.Generator(device=dev)
        gn.manual_seed(seed * 1000003 + f0 // chunk + 1)
        y = torch.randn((f1 - f0, d1 * d2), generator=gn, device=dev)
        y += torch.matmul(c[:, f0:f1].t(), a)
        if bg_rank:
            y += torch.matmul(walks[:, f0:f1].t(), bgm)
        y = mu[None] + sigma_px[None] * y
        lo, hi = max(f0, frame_lo), min(f1, frame_hi)
        if lo < hi:
            blk = y[lo - f0 : hi - f0]
            if not dtype.is_floating_point:
                info = torch.iinfo(dtype)
                blk = blk.round().clamp(info.min, info.max)
            out[lo - frame_lo : hi - frame_lo] = blk.to(dtype)
    return out.view(frame_hi - frame_lo, d1, d2)
