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


class _PlainTiff:
    """Minimal reader of uncompressed greyscale multi-page TIFF / BigTIFF files (what acquisition software, ImageJ and
    the reference's demo data write): one image file directory per frame, or ImageJ's single directory followed by
    `images=N` contiguous frames.  Used when the `tifffile` package is not installed; anything it does not understand
    (compression, tiles, colour) raises ValueError naming the tag."""

    _TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 8: "h", 9: "i", 11: "f", 12: "d", 16: "Q", 17: "q"}

    def __init__(self, filename):
        import struct

        self.filename = filename
        with open(filename, "rb") as f:
            head = f.read(16)
            if head[:2] == b"II":
                self.bo = "<"
            elif head[:2] == b"MM":
                self.bo = ">"
            else:
                raise ValueError("%s is not a TIFF file" % filename)
            magic = struct.unpack(self.bo + "H", head[2:4])[0]
            if magic == 42:
                self.big = False
                offset = struct.unpack(self.bo + "I", head[4:8])[0]
            elif magic == 43:
                self.big = True
                offset = struct.unpack(self.bo + "Q", head[8:16])[0]
            else:
                raise ValueError("%s is not a TIFF file (magic %d)" % (filename, magic))
            self.pages = []
            seen = set()
            while offset and offset not in seen:
                seen.add(offset)
                tags, offset = self._read_ifd(f, offset, struct)
                self.pages.append(self._page(tags))
        if not self.pages:
            raise ValueError("%s holds no image" % filename)
        first = self.pages[0]
        n_ij = first.pop("imagej_images", 0)
        if len(self.pages) == 1 and n_ij > 1:  # ImageJ stack: frames follow each other after the first strip
            if len(first["offsets"]) != 1 and not first["contiguous"]:
                raise ValueError("ImageJ stack with non-contiguous first frame")
            nbytes = first["h"] * first["w"] * first["dtype"].itemsize
            base = first["offsets"][0]
            self.pages = [dict(first, offsets=[base + i * nbytes], counts=[nbytes], contiguous=True) for i in range(n_ij)]
        self.shape = (len(self.pages), first["h"], first["w"])
        self.dtype = first["dtype"]

    def _read_ifd(self, f, offset, struct):
        bo = self.bo
        f.seek(offset)
        if self.big:
            n = struct.unpack(bo + "Q", f.read(8))[0]
            raw = f.read(20 * n + 8)
            esz, cfmt, vsz = 20, "Q", 8
        else:
            n = struct.unpack(bo + "H", f.read(2))[0]
            raw = f.read(12 * n + 4)
            esz, cfmt, vsz = 12, "I", 4
        tags = {}
        for i in range(n):
            e = raw[esz * i : esz * (i + 1)]
            tag, typ = struct.unpack(bo + "HH", e[:4])
            count = struct.unpack(bo + cfmt, e[4 : 4 + vsz])[0]
            fmt = self._TYPES.get(typ)
            if fmt is None:
                continue
            size = struct.calcsize(bo + fmt) * count
            if size <= vsz:
                data = e[4 + vsz : 4 + vsz + size]
            else:
                pos = struct.unpack(bo + cfmt, e[4 + vsz : 4 + 2 * vsz])[0]
                here = f.tell()
                f.seek(pos)
                data = f.read(size)
                f.seek(here)
            tags[tag] = data if typ == 2 else struct.unpack(bo + fmt * count, data)
        nxt = struct.unpack(bo + cfmt, raw[esz * n : esz * n + vsz])[0]
        return tags, nxt

    def _page(self, tags):
        one = lambda t, default=None: tags[t][0] if t in tags else default  # noqa: E731
        w, h = one(256), one(257)
        if w is None or h is None:
            raise ValueError("TIFF directory without ImageWidth / ImageLength")
        if one(259, 1) != 1:
            raise ValueError("compressed TIFF (Compression tag %d) needs the `tifffile` package" % one(259))
        if one(277, 1) != 1:
            raise ValueError("only greyscale TIFF is supported (SamplesPerPixel %d)" % one(277))
        if 322 in tags or 324 in tags:
            raise ValueError("tiled TIFF needs the `tifffile` package")
        bits, fmt = one(258, 1), one(339, 1)
        kind = {1: "u", 2: "i", 3: "f"}.get(fmt)
        if kind is None or bits not in (8, 16, 32, 64) or (kind == "f" and bits < 32):
            raise ValueError("unsupported TIFF sample type (BitsPerSample %d, SampleFormat %d)" % (bits, fmt))
        dtype = np.dtype("%s%s%d" % (self.bo, kind, bits // 8))
        offsets, counts = list(tags.get(273, ())), list(tags.get(279, ()))
        if not offsets:
            raise ValueError("TIFF directory without StripOffsets")
        rps = one(278, h)
        if not counts:
            counts = [min(rps, h - i * rps) * w * dtype.itemsize for i in range(len(offsets))]
        contiguous = all(offsets[i] + counts[i] == offsets[i + 1] for i in range(len(offsets) - 1))
        page = dict(w=int(w), h=int(h), dtype=dtype, offsets=offsets, counts=counts, contiguous=contiguous)
        desc = tags.get(270)
        if isinstance(desc, bytes) and desc.startswith(b"ImageJ"):
            for line in desc.split(b"\n"):
                if line.startswith(b"images="):
                    page["imagej_images"] = int(line[7:].strip(b"\x00 "))
        return page

    def read(self, indices):
        out = np.empty((len(indices), self.shape[1], self.shape[2]), dtype=self.dtype.newbyteorder("="))
        with open(self.filename, "rb") as f:
            for k, i in enumerate(indices):
                pg = self.pages[i]
                if (pg["h"], pg["w"]) != self.shape[1:] or pg["dtype"] != self.dtype:
                    raise ValueError("TIFF page %d differs in shape or type from page 0" % i)
                flat = out[k].reshape(-1)
                pos = 0
                for off, cnt in zip(pg["offsets"], pg["counts"]):
                    f.seek(off)
                    part = np.frombuffer(f.read(cnt), dtype=self.dtype)
                    flat[pos : pos + part.size] = part
                    pos += part.size
                if pos != flat.size:
                    raise ValueError("TIFF page %d is truncated" % i)
        return out


class TiffArray(lazy_data_loader):
    """Multi-page TIFF movie (dataset.py:131-181): frames are decoded on demand and returned as float32, like the
    reference.  Decoding goes through `tifffile` when it is installed, otherwise through the built-in reader of
    uncompressed greyscale TIFF / BigTIFF / ImageJ stacks above."""

    def __init__(self, filename):
        self.filename = filename
        try:
            import tifffile
        except ImportError:
            tifffile = None
        self._tf = tifffile
        if tifffile is not None:
            with tifffile.TiffFile(filename) as tf:
                n = len(tf.pages)
                page = tf.pages[0]
                self._shape = (n, int(page.shape[0]), int(page.shape[1]))
            self._plain = None
        else:
            self._plain = _PlainTiff(filename)
            self._shape = self._plain.shape

    @property
    def dtype(self):
        return np.float32

    @property
    def shape(self):
        return self._shape

    def _compute_at_indices(self, indices):
        if isinstance(indices, (int, np.integer)):
            indices = [int(indices)]
        elif isinstance(indices, slice):
            indices = list(range(indices.start or 0, indices.stop or self._shape[0], indices.step or 1))
        else:
            indices = [int(i) for i in indices]
        if self._tf is not None:
            data = self._tf.imread(self.filename, key=indices)
        else:
            data = self._plain.read(indices)
        return np.asarray(data).squeeze().astype(self.dtype)


def _index_rows(t, idx):
    """index_select that also works for uint16 storage (moved bit-exact through an int16 view)."""
    if t.dtype == torch.uint16:
        return t.view(torch.int16).index_select(0, idx).view(torch.uint16)
    return t.index_select(0, idx)


# -------------------------------------------------------------------------------------------------
class DeviceMovie:
    """The movie as the kernels see it: frame-major (T, d) on one GPU.

    * a CUDA torch tensor (T, d1, d2) of a supported dtype is used in place (zero copy);
    * anything else honouring the dataset contract is moved host -> device in `upload_frames`-sized
      chunks on a side stream.  When it fits `resident_fraction` of the free memory it is kept in HBM:
      the FIRST call of batches() (the mean/noise pass) performs the upload and hands every chunk to the
      compute stream as soon as it has landed, so the copy overlaps that pass; later passes read HBM.
      Otherwise every pass re-streams the movie chunk by chunk.
    Contiguous ndarrays / CPU tensors of a native dtype are sliced without a host copy (pinned memory is
    copied from directly, pageable memory goes through two reusable pinned staging buffers).
    frame_lo/frame_hi restrict the wrapper to a frame shard (multi-GPU)."""

    def __init__(self, dataset_obj, device, batch_frames=2048, frame_lo=0, frame_hi=None, resident_fraction=0.6,
                 upload_frames=2048):
        self.device = torch.device(device)
        self.T_total = int(dataset_obj.shape[0])
        self.d1, self.d2 = int(dataset_obj.shape[1]), int(dataset_obj.shape[2])
        self.d = self.d1 * self.d2
        self.lo = int(frame_lo)
        self.hi = self.T_total if frame_hi is None else int(frame_hi)
        self.batch_frames = max(1024, (int(batch_frames) // 1024) * 1024)
        self.upload_frames = max(1024, (int(upload_frames) // 1024) * 1024)
        self.h2d_bytes = 0
        self._src = dataset_obj
        self._resident = None
        self._filled = True
        self._host2d = None
        self._host_off = 0
        self._staging = None
        if isinstance(dataset_obj, torch.Tensor):
            if not dataset_obj.is_cuda:
                if dataset_obj.dtype in ops.PMD_DTYPES and dataset_obj.is_contiguous():
                    self._host2d = dataset_obj.view(self.T_total, self.d)
                else:
                    dataset_obj = dataset_obj.numpy()
                    self._src = dataset_obj
            else:
                t = dataset_obj
                if t.dtype not in ops.PMD_DTYPES:
                    t = t.to(torch.float32)
                self._resident = t[self.lo : self.hi].contiguous().view(self.hi - self.lo, self.d)
                self.torch_dtype = t.dtype
                return
        if self._host2d is None and isinstance(self._src, np.ndarray) and self._src.dtype in ops.NUMPY_NATIVE \
                and self._src.flags.c_contiguous:
            self._host2d = torch.from_numpy(self._src).view(self.T_total, self.d)
        if self._host2d is not None:
            self._np_native = True
            self.torch_dtype = self._host2d.dtype
            self._pinned_src = bool(self._host2d.is_pinned())
        else:
            probe = np.asarray(self._src[[self.lo]])
            self._np_native = probe.dtype in ops.NUMPY_NATIVE
            self.torch_dtype = ops.NUMPY_NATIVE[probe.dtype] if self._np_native else torch.float32
            self._pinned_src = False
        self._copy_stream = torch.cuda.Stream(self.device)
        nbytes = (self.hi - self.lo) * self.d * torch.empty((), dtype=self.torch_dtype).element_size()
        free, _ = torch.cuda.mem_get_info(self.device)
        if nbytes <= resident_fraction * free:
            self._resident = torch.empty((self.hi - self.lo, self.d), dtype=self.torch_dtype, device=self.device)
            self._filled = False

    @classmethod
    def from_shard(cls, local, n_frames_total, frame_lo):
        """Wrap this rank's frame shard `local` (n_local, d1, d2), a CUDA tensor of a supported dtype holding frames
        [frame_lo, frame_lo + n_local) of a movie with n_frames_total frames (multi-GPU: one shard per rank)."""
        if not (isinstance(local, torch.Tensor) and local.is_cuda):
            raise TypeError("from_shard expects a CUDA tensor")
        self = cls.__new__(cls)
        self.device = local.device
        self.T_total = int(n_frames_total)
        self.d1, self.d2 = int(local.shape[1]), int(local.shape[2])
        self.d = self.d1 * self.d2
        self.lo, self.hi = int(frame_lo), int(frame_lo) + int(local.shape[0])
        self.batch_frames = self.upload_frames = 2048
        self.h2d_bytes = 0
        self._src = None
        t = local if local.dtype in ops.PMD_DTYPES else local.to(torch.float32)
        self._resident = t.contiguous().view(self.hi - self.lo, self.d)
        self.torch_dtype = t.dtype
        self._filled = True
        self._host2d = None
        self._staging = None
        return self

    @classmethod
    def from_host_shard(cls, local, n_frames_total, frame_lo, device, **kw):
        """This rank's frame shard as a HOST array / CPU tensor (n_local, d1, d2) holding frames
        [frame_lo, frame_lo + n_local) of a movie with n_frames_total frames; uploaded like any host dataset."""
        n = int(local.shape[0])
        self = cls.__new__(cls)
        proxy = torch.from_numpy(local) if isinstance(local, np.ndarray) else local
        if proxy.dtype not in ops.PMD_DTYPES or not proxy.is_contiguous():
            proxy = proxy.to(torch.float32).contiguous()
        cls.__init__(self, proxy, device, frame_lo=0, frame_hi=n, **kw)  # as a stand-alone movie of n frames ...
        # ... then re-labelled with its global frame range
        self.T_total = int(n_frames_total)
        self.lo, self.hi = int(frame_lo), int(frame_lo) + n
        self._host_off = int(frame_lo)
        return self

    @property
    def n_local(self):
        return self.hi - self.lo

    @property
    def shape(self):
        return (self.T_total, self.d1, self.d2)

    # ---- host side ---------------------------------------------------------------------------------
    def _host_frames(self, ids):
        """Pinned (n, d) host tensor of arbitrary frames through the dataset contract."""
        arr = np.asarray(self._src[list(ids)])
        if arr.ndim == 2:
            arr = arr[None]
        if not self._np_native:
            arr = arr.astype(np.float32)
        arr = np.ascontiguousarray(arr).reshape(len(ids), self.d)
        t = torch.from_numpy(arr)
        return t.pin_memory() if torch.cuda.is_available() else t

    def _stage(self, n):
        """One of two reusable pinned staging buffers (its previous async copy is waited for first)."""
        if self._staging is None:
            self._staging = [[torch.empty((self.upload_frames, self.d), dtype=self.torch_dtype, pin_memory=True), None]
                             for _ in range(2)]
            self._stage_turn = 0
        slot = self._staging[self._stage_turn]
        self._stage_turn ^= 1
        if slot[1] is not None:
            slot[1].synchronize()
        return slot, slot[0][:n]

    def _upload(self, f0, f1, dst):
        """Enqueue the host -> device copy of global frames [f0, f1) into dst on the copy stream; returns
        the event that marks its completion."""
        n = f1 - f0
        with torch.cuda.stream(self._copy_stream):
            if self._host2d is not None and self._pinned_src:
                dst.copy_(self._host2d[f0 - self._host_off : f1 - self._host_off], non_blocking=True)
                slot = None
            else:
                slot, buf = self._stage(n)
                if self._host2d is not None:
                    buf.copy_(self._host2d[f0 - self._host_off : f1 - self._host_off])
                else:
                    arr = np.asarray(self._src[list(range(f0, f1))])
                    arr = arr[None] if arr.ndim == 2 else arr
                    buf.copy_(torch.from_numpy(np.ascontiguousarray(arr).reshape(n, self.d)))
                dst.copy_(buf, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
            if slot is not None:
                slot[1] = ev
        self.h2d_bytes += n * self.d * dst.element_size()
        return ev

    def _fill(self):
        for _ in self.batches():
            pass
        torch.cuda.current_stream(self.device).synchronize()

    def batches(self):
        """Yield (first local frame index, (n, d) device tensor) covering the shard once, in order."""
        cur = torch.cuda.current_stream(self.device)
        if self._resident is not None and self._filled:
            step = max(self.batch_frames * 8, 65536)
            for s in range(0, self.n_local, step):
                yield s, self._resident[s : s + step]
        elif self._resident is not None:
            # first pass: upload chunk by chunk, two copies in flight ahead of the consumer
            self._copy_stream.wait_stream(cur)
            spans = [(f0, min(self.hi, f0 + self.upload_frames)) for f0 in range(self.lo, self.hi, self.upload_frames)]
            evs = {}
            ahead = 2
            for i in range(min(ahead, len(spans))):
                evs[i] = self._upload(spans[i][0], spans[i][1], self._resident[spans[i][0] - self.lo : spans[i][1] - self.lo])
            for i, (f0, f1) in enumerate(spans):
                cur.wait_event(evs.pop(i))
                yield f0 - self.lo, self._resident[f0 - self.lo : f1 - self.lo]
                j = i + ahead
                if j < len(spans):
                    evs[j] = self._upload(spans[j][0], spans[j][1], self._resident[spans[j][0] - self.lo : spans[j][1] - self.lo])
            self._filled = True
        else:
            for f0 in range(self.lo, self.hi, self.upload_frames):
                f1 = min(self.hi, f0 + self.upload_frames)
                # the chunk is allocated on the compute stream and written on the copy stream: the copy must not start
                # before the kernels that used this (recycled) block have run, and the allocator must know both streams
                chunk = torch.empty((f1 - f0, self.d), dtype=self.torch_dtype, device=self.device)
                self._copy_stream.wait_stream(cur)
                chunk.record_stream(self._copy_stream)
                cur.wait_event(self._upload(f0, f1, chunk))
                yield f0 - self.lo, chunk

    def frame_source(self, global_frame_ids):
        """(2-D device tensor, int64 device row indices) addressing the requested frames without copying
        them when the shard is resident; otherwise the frames are gathered first."""
        ids = [int(i) for i in global_frame_ids]
        if self._resident is not None and all(self.lo <= i < self.hi for i in ids):
            if not self._filled:
                self._fill()
            return self._resident, ops.h2d(np.asarray(ids, dtype=np.int64) - self.lo, self.device)
        g = self.gather(ids)
        return g, torch.arange(g.shape[0], dtype=torch.int64, device=self.device)

    def gather(self, global_frame_ids):
        """(n, d) device tensor (native dtype) of arbitrary global frames (any shard)."""
        ids = [int(i) for i in global_frame_ids]
        if self._resident is not None and all(self.lo <= i < self.hi for i in ids):
            if not self._filled:
                self._fill()
            idx = torch.as_tensor(ids, dtype=torch.int64, device=self.device) - self.lo
            return _index_rows(self._resident, idx)
        if isinstance(self._src, torch.Tensor) and self._src.is_cuda:
            idx = torch.as_tensor(ids, dtype=torch.int64, device=self._src.device)
            return _index_rows(self._src.view(self.T_total, self.d), idx).to(self.device)
        if self._host2d is not None:
            idx = torch.as_tensor(ids, dtype=torch.int64) - self._host_off
            host = _index_rows(self._host2d, idx)
            self.h2d_bytes += host.numel() * host.element_size()
            return host.to(self.device)
        host = self._host_frames(ids)
        self.h2d_bytes += host.numel() * host.element_size()
        return host.to(self.device)
