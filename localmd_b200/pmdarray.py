"""PMDArray: array-like view of a PMD decomposition  Y_hat = (U R) diag(s) Vt * std + mean.

Mirrors the reference class (pmdarray.py:7-171): same constructor, properties (u CSR, r, s, v, dtype,
shape, ndim, order, mean_img, var_img, row_indices) and `__getitem__` semantics.  Differences, all
behind the same interface:
  * reconstruction runs on the GPU (csrc/reconstruct.cu, a CSR-SpMM fused with the un-normalisation);
  * the dense (R_total x T) product R diag(s) Vt that the reference builds eagerly on the CPU
    (pmdarray.py:50-52) is never materialised: only the requested frames' columns are formed;
  * the 2-key form arr[frames, rows], which raises TypeError in the reference (pmdarray.py:146-148),
    works as evidently intended (all columns).
"""
from typing import Tuple

import numpy as np
import scipy.sparse
import torch

from . import ops


class PMDArray:
    def __init__(self, u, r, s, v, data_shape: Tuple[int, int, int], data_order: str, mean_img, std_img, device=None):
        self.order = data_order
        self.num_frames, self.fov_dim1, self.fov_dim2 = (int(x) for x in data_shape)
        self._u = scipy.sparse.csr_matrix(u)
        self._r = np.asarray(r)
        self._s = np.asarray(s)
        self._v = np.asarray(v)
        self._mean_img = np.asarray(mean_img)
        self._var_img = np.asarray(std_img)
        self._device = device
        self._dev = None
        self._lazy = None

    @classmethod
    def _from_device(cls, csr_order, csr_phys32, rmix, s, vt, data_shape, data_order, mean, std, device):
        """Build from device-resident factors without any device->host copy: the host views (u, r, s, v)
        are materialised on first access through pinned buffers.  csr_order = (indptr, indices, values64)
        with rows numbered in `data_order`; csr_phys32 = the same matrix over physical pixel rows, float32."""
        self = cls.__new__(cls)
        self.order = data_order
        self.num_frames, self.fov_dim1, self.fov_dim2 = (int(x) for x in data_shape)
        d = self.fov_dim1 * self.fov_dim2
        self._u = self._r = self._s = self._v = None
        self._lazy = dict(csr=csr_order, r=rmix, s=s, v=vt, n_cols=int(rmix.shape[0]))
        self._mean_img = self._var_img = None      # host images: materialised on first access, like the factors
        self._device = device
        ip, ix, v32 = csr_phys32
        self._dev = dict(device=torch.device(device), indptr=ip, indices=ix, values=v32,
                         rs=(rmix * s[None, :]).contiguous(), vt=vt.contiguous(), mean=mean.contiguous(), std=std.contiguous())
        return self

    @property
    def mean_img(self):
        if self._mean_img is None:
            self._mean_img = self._to_host(self._dev["mean"]).reshape(self.fov_dim1, self.fov_dim2)
        return self._mean_img

    @property
    def var_img(self):
        if self._var_img is None:
            self._var_img = self._to_host(self._dev["std"]).reshape(self.fov_dim1, self.fov_dim2)
        return self._var_img

    @property
    def row_indices(self):
        """(d1, d2) array: row of U (numbered in `order`) of every pixel (pmdarray.py:37-38)."""
        if getattr(self, "_row_indices", None) is None:
            self._row_indices = np.arange(self.fov_dim1 * self.fov_dim2).reshape((self.fov_dim1, self.fov_dim2), order=self.order)
        return self._row_indices

    @property
    def _phys_indices(self):
        if getattr(self, "_phys", None) is None:
            self._phys = np.arange(self.fov_dim1 * self.fov_dim2).reshape((self.fov_dim1, self.fov_dim2))
        return self._phys

    @staticmethod
    def _to_host(t):
        # through torch's caching pinned-host allocator: after the first decomposition of a process the page-locked
        # blocks are reused, so the copy runs at PCIe rate; the ndarray is a view of the pinned block and keeps it alive
        if t.numel() == 0 or not t.is_cuda:
            return t.cpu().numpy()
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return host.numpy()

    def _materialise(self, what):
        lz = self._lazy
        if what == "u" and self._u is None:
            ip, ix, v = lz["csr"]
            d = self.fov_dim1 * self.fov_dim2
            self._u = scipy.sparse.csr_matrix(
                (self._to_host(v), self._to_host(ix), self._to_host(ip).astype(np.int32)), shape=(d, lz["n_cols"]))
        elif what == "r" and self._r is None:
            self._r = self._to_host(lz["r"])
        elif what == "s" and self._s is None:
            self._s = self._to_host(lz["s"])
        elif what == "v" and self._v is None:
            self._v = self._to_host(lz["v"])

    # ---- reference properties -------------------------------------------------------------------
    @property
    def u(self):
        if self._u is None:
            self._materialise("u")
        return self._u

    @property
    def r(self):
        if self._r is None:
            self._materialise("r")
        return self._r

    @property
    def s(self):
        if self._s is None:
            self._materialise("s")
        return self._s

    @property
    def v(self):
        if self._v is None:
            self._materialise("v")
        return self._v

    @property
    def dtype(self):
        return np.float32

    @property
    def shape(self):
        return (self.num_frames, self.fov_dim1, self.fov_dim2)

    @property
    def ndim(self):
        return 3

    # ---- device state ---------------------------------------------------------------------------
    def _device_state(self):
        if self._dev is None:
            if not torch.cuda.is_available():
                raise RuntimeError("PMDArray reconstruction needs a CUDA device (sm_100a); there is no CPU fallback")
            dev = torch.device(self._device if self._device is not None else "cuda")
            # CSR rows are numbered in `order`; the kernels index physical (row-major) pixels
            perm = self.row_indices.reshape(-1)  # physical pixel p -> row id in `order`
            u_phys = self.u[perm].tocsr()
            u_phys.sort_indices()
            self._dev = dict(
                device=dev,
                indptr=torch.from_numpy(u_phys.indptr.astype(np.int64)).to(dev),
                indices=torch.from_numpy(u_phys.indices.astype(np.int32)).to(dev),
                values=torch.from_numpy(u_phys.data.astype(np.float32)).to(dev),
                rs=torch.from_numpy(np.ascontiguousarray(self.r * self.s[None, :], dtype=np.float32)).to(dev),
                vt=torch.from_numpy(np.ascontiguousarray(self.v, dtype=np.float32)).to(dev),
                mean=torch.from_numpy(np.ascontiguousarray(self.mean_img, dtype=np.float32).reshape(-1)).to(dev),
                std=torch.from_numpy(np.ascontiguousarray(self.var_img, dtype=np.float32).reshape(-1)).to(dev),
            )
        return self._dev

    # ---- indexing -------------------------------------------------------------------------------
    @staticmethod
    def _parse_int_to_list(elt):
        if isinstance(elt, (int, np.integer)):
            return [int(elt)]
        return elt

    def _frames(self, key):
        if key is None:
            raise ValueError("Cannot use None for indexing")
        return np.arange(self.num_frames)[self._parse_int_to_list(key)]

    def __getitem__(self, key) -> np.ndarray:
        """Returns self[key] as float32 with frames first.  Does NOT support dimension expansion."""
        if key is None:
            raise ValueError("Cannot use None for indexing")
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) > 3:
            raise ValueError("Too many values to unpack in __getitem__")
        key = tuple(key) + (slice(None, None, None),) * (3 - len(key))
        if key[1] is None or key[2] is None:
            raise ValueError("Cannot pass in None for indexing")
        frames = np.atleast_1d(self._frames(key[0])).astype(np.int64)
        k1, k2 = self._parse_int_to_list(key[1]), self._parse_int_to_list(key[2])
        used = self._phys_indices[k1, k2]
        implied = used.shape
        out = self.reconstruct(frames, used.reshape(-1))
        out = out.reshape((len(frames),) + implied)
        return out.squeeze().astype(self.dtype, copy=False)

    def reconstruct(self, frames, pix, max_bytes=1 << 30):
        """(len(frames), len(pix)) float32 host array of reconstructed values at physical pixels `pix`."""
        st = self._device_state()
        dev = st["device"]
        pix_t = torch.from_numpy(np.ascontiguousarray(pix, dtype=np.int32)).to(dev)
        frames = np.asarray(frames, dtype=np.int64)
        # the result lands in a page-locked buffer from torch's caching host allocator (copies at PCIe rate, overlapped
        # with the reconstruction of the next chunk); the returned ndarray is a view that keeps the buffer alive
        out_t = torch.empty((len(frames), len(pix)), dtype=torch.float32, pin_memory=True)
        step = max(4, int(max_bytes // max(4 * len(pix), 1)) // 4 * 4)
        for s0 in range(0, len(frames), step):
            fr = torch.from_numpy(frames[s0 : s0 + step]).to(dev)
            c = torch.matmul(st["rs"], st["vt"].index_select(1, fr)).contiguous()  # (R_total, n)
            chunk = ops.reconstruct(st["indptr"], st["indices"], st["values"], c, pix_t, st["std"], st["mean"])
            out_t[s0 : s0 + step].copy_(chunk, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out_t.numpy()
