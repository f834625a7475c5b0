"""Dataset contract of the reference (dataset.py:7-128) and the device-resident movie wrapper.

The effective contract used by the reference's hot path is duck typed: `.shape == (T, d1, d2)` and
`obj[list_of_frame_ids] -> ndarray (n, d1, d2)` (pmd_loader.py:99,143,188; test/test_pmd.py passes a
raw ndarray).  `lazy_data_loader` mirrors the reference ABC so user subclasses keep working.
`TiffArray` (dataset.py:131-181) is host file I/O and is out of scope: tifffile is not in this image.
"""
from abc import ABC, abstractmethod
from typing import Tuple, Union

import numpy as np
import torch

from . import ops


class lazy_data_loader(ABC):
    """Same interface and indexing/error behaviour as the reference ABC (dataset.py:7-128)."""

    @property
    @abstractmethod
    def dtype(self) -> str:
        pass

    @property
    @abstractmethod
    def shape(self) -> Tuple[int, int, int]:
        pass

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def __getitem__(self, item):
        if isinstance(item, tuple):
            if len(item) > len(self.shape):
                raise IndexError(
                    f"Cannot index more dimensions than exist in the array. "
                    f"You have tried to index with <{len(item)}> dimensions, "
                    f"only <{len(self.shape)}> dimensions exist in the array"
                )
            frame_indexer = item[0]
        else:
            frame_indexer = item
        if isinstance(frame_indexer, np.ndarray):
            frame_indexer = frame_indexer.tolist()
        if isinstance(frame_indexer, (list, int)):
            pass
        elif isinstance(frame_indexer, np.integer):
            frame_indexer = frame_indexer.item()
        elif isinstance(frame_indexer, (slice, range)):
            start, stop, step = frame_indexer.start, frame_indexer.stop, frame_indexer.step
            if start is not None and start > self.shape[0]:
                raise IndexError(
                    f"Cannot index beyond `n_frames`.\nDesired frame start index of <{start}> "
                    f"lies beyond `n_frames` <{self.shape[0]}>"
                )
            if stop is not None and stop > self.shape[0]:
                raise IndexError(
                    f"Cannot index beyond `n_frames`.\nDesired frame stop index of <{stop}> "
                    f"lies beyond `n_frames` <{self.shape[0]}>"
                )
            frame_indexer = slice(start, stop, 1 if step is None else step)
        else:
            raise IndexError(f"Invalid indexing method, you have passed a: <{type(item)}>")
        frames = self._compute_at_indices(frame_indexer)
        if len(frames.shape) < len(self.shape):
            frames = np.expand_dims(frames, axis=0)
        if isinstance(item, tuple):
            if len(item) == 2:
                frames = frames[:, item[1]]
            elif len(item) == 3:
                frames = frames[:, item[1], item[2]]
        return frames.squeeze()

    @abstractmethod
    def _compute_at_indices(self, indices: Union[list, int, slice]) -> np.ndarray:
        pass


class TiffArray(lazy_data_loader):
    """Placeholder with the reference's name (dataset.py:131-181).  Multipage-TIFF decoding is host
    file I/O outside the accelerated path and needs `tifffile`, which this image does not ship."""

    def __init__(self, filename):
        try:
            import tifffile  # noqa: F401
        except ImportError as e:  # pragma: no cover
            raise ImportError("TiffArray needs the `tifffile` package, which is not installed") from e
        import tifffile

        self.filename = filename
        self._tf = tifffile
        with tifffile.TiffFile(filename) as tf:
            n = len(tf.pages)
            page = tf.pages[0]
            self._shape = (n, page.shape[0], page.shape[1])
            self._dtype = str(page.dtype)

    @property
    def dtype(self):
        return self._dtype

    @property
    def shape(self):
        return self._shape

    def _compute_at_indices(self, indices):
        if isinstance(indices, int):
            indices = [indices]
        if isinstance(indices, slice):
            indices = list(range(*indices.indices(self._shape[0])))
        return self._tf.imread(self.filename, key=indices).squeeze()


def _index_rows(t, idx):
    """index_select that also works for uint16 storage (moved bit-exact through an int16 view)."""
    if t.dtype == torch.uint16:
        return t.view(torch.int16).index_select(0, idx).view(torch.uint16)
    return t.index_select(0, idx)


# -------------------------------------------------------------------------------------------------
class DeviceMovie:
    """The movie as the kernels see it: frame-major (T, d) on one GPU.

    * a CUDA torch tensor (T, d1, d2) of a supported dtype is used in place (zero copy);
    * anything else honouring the dataset contract is staged through pinned host buffers: kept
      resident in HBM when it fits `resident_fraction` of the free memory, otherwise re-streamed
      batch by batch for each of the two full passes.
    frame_lo/frame_hi restrict the wrapper to a frame shard (multi-GPU)."""

    def __init__(self, dataset_obj, device, batch_frames=2048, frame_lo=0, frame_hi=None, resident_fraction=0.6):
        self.device = torch.device(device)
        self.T_total = int(dataset_obj.shape[0])
        self.d1, self.d2 = int(dataset_obj.shape[1]), int(dataset_obj.shape[2])
        self.d = self.d1 * self.d2
        self.lo = int(frame_lo)
        self.hi = self.T_total if frame_hi is None else int(frame_hi)
        self.batch_frames = max(1024, (int(batch_frames) // 1024) * 1024)
        self.h2d_bytes = 0
        self._src = dataset_obj
        self._resident = None
        if isinstance(dataset_obj, torch.Tensor):
            if not dataset_obj.is_cuda:
                dataset_obj = dataset_obj.numpy()
                self._src = dataset_obj
            else:
                t = dataset_obj
                if t.dtype not in ops.PMD_DTYPES:
                    t = t.to(torch.float32)
                self._resident = t[self.lo : self.hi].contiguous().view(self.hi - self.lo, self.d)
                self.torch_dtype = t.dtype
                return
        probe = np.asarray(self._src[[self.lo]])
        self._np_native = probe.dtype in ops.NUMPY_NATIVE
        self.torch_dtype = ops.NUMPY_NATIVE[probe.dtype] if self._np_native else torch.float32
        nbytes = (self.hi - self.lo) * self.d * torch.empty((), dtype=self.torch_dtype).element_size()
        free, _ = torch.cuda.mem_get_info(self.device)
        if nbytes <= resident_fraction * free:
            buf = torch.empty((self.hi - self.lo, self.d), dtype=self.torch_dtype, device=self.device)
            for f0, f1, chunk in self._host_batches():
                buf[f0 - self.lo : f1 - self.lo].copy_(chunk, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            self._resident = buf

    @property
    def n_local(self):
        return self.hi - self.lo

    @property
    def shape(self):
        return (self.T_total, self.d1, self.d2)

    def _host_frames(self, ids):
        arr = np.asarray(self._src[list(ids)])
        if arr.ndim == 2:
            arr = arr[None]
        if not self._np_native:
            arr = arr.astype(np.float32)
        arr = np.ascontiguousarray(arr).reshape(len(ids), self.d)
        t = torch.from_numpy(arr)
        self.h2d_bytes += t.numel() * t.element_size()
        return t.pin_memory() if torch.cuda.is_available() else t

    def _host_batches(self):
        for f0 in range(self.lo, self.hi, self.batch_frames):
            f1 = min(self.hi, f0 + self.batch_frames)
            yield f0, f1, self._host_frames(range(f0, f1))

    def batches(self):
        """Yield (first local frame index, (n, d) device tensor) covering the shard once."""
        if self._resident is not None:
            step = self.batch_frames * 8
            for s in range(0, self.n_local, step):
                yield s, self._resident[s : s + step]
        else:
            for f0, f1, chunk in self._host_batches():
                yield f0 - self.lo, chunk.to(self.device, non_blocking=True)

    def gather(self, global_frame_ids):
        """(n, d) device tensor (native dtype) of arbitrary global frames (any shard)."""
        ids = [int(i) for i in global_frame_ids]
        if self._resident is not None and all(self.lo <= i < self.hi for i in ids):
            idx = torch.as_tensor(ids, dtype=torch.int64, device=self.device) - self.lo
            return _index_rows(self._resident, idx)
        if isinstance(self._src, torch.Tensor):
            idx = torch.as_tensor(ids, dtype=torch.int64, device=self._src.device)
            return _index_rows(self._src.view(self.T_total, self.d), idx).to(self.device)
        return self._host_frames(ids).to(self.device)
