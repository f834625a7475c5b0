"""TiffArray (dataset.py:131-181 of the reference) with the built-in reader of uncompressed TIFF files: the files are
written here byte by byte (classic little / big endian with several strips per page, BigTIFF, an ImageJ stack with a
single directory), independent of any TIFF library."""
import struct

import numpy as np
import pytest

from localmd_b200.dataset import TiffArray, _PlainTiff


def write_tiff(path, frames, bo="<", rows_per_strip=None, big=False, imagej=False):
    frames = np.asarray(frames)
    n, h, w = frames.shape
    dt = frames.dtype.newbyteorder(bo)
    kind = {"u": 1, "i": 2, "f": 3}[frames.dtype.kind]
    rps = rows_per_strip or h
    nstrips = -(-h // rps)
    off_fmt, cnt_fmt, esz = ("Q", "Q", 20) if big else ("I", "H", 12)
    out = bytearray()
    out += (b"II" if bo == "<" else b"MM")
    out += struct.pack(bo + "H", 43 if big else 42)
    if big:
        out += struct.pack(bo + "HHQ", 8, 0, 0)
        first_ptr = 8
    else:
        out += struct.pack(bo + "I", 0)
        first_ptr = 4
    pages = 1 if imagej else n
    prev_ptr = first_ptr
    desc = b"ImageJ=1.53\nimages=%d\nslices=%d\n\x00" % (n, n) if imagej else None
    for pi in range(pages):
        data = frames[pi:] if imagej else frames[pi : pi + 1]
        raw = data.astype(dt).tobytes()
        data_off = len(out)
        out += raw
        strip_bytes = [min(rps, h - s * rps) * w * dt.itemsize for s in range(nstrips)]
        strip_offs = [data_off + sum(strip_bytes[:s]) for s in range(nstrips)]
        extra = bytearray()
        entries = []

        def entry(tag, typ, values):
            fmt = {3: "H", 4: "I", 16: "Q", 2: "c"}[typ]
            if typ == 2:
                payload, count = values, len(values)
            else:
                payload, count = struct.pack(bo + fmt * len(values), *values), len(values)
            entries.append((tag, typ, count, payload))

        entry(256, 4, [w]), entry(257, 4, [h]), entry(258, 3, [dt.itemsize * 8]), entry(259, 3, [1]), entry(262, 3, [1])
        if desc:
            entry(270, 2, desc)
        entry(273, 16 if big else 4, strip_offs), entry(277, 3, [1]), entry(278, 4, [rps])
        entry(279, 16 if big else 4, strip_bytes), entry(339, 3, [kind])
        entries.sort(key=lambda e: e[0])
        while len(out) % 2:
            out += b"\x00"
        ifd_off = len(out)
        head = struct.pack(bo + ("Q" if big else "H"), len(entries))
        vsz = 8 if big else 4
        ifd_size = len(head) + esz * len(entries) + vsz
        body = bytearray()
        for tag, typ, count, payload in entries:
            e = struct.pack(bo + "HH", tag, typ) + struct.pack(bo + ("Q" if big else "I"), count)
            if len(payload) <= vsz:
                e += payload + b"\x00" * (vsz - len(payload))
            else:
                e += struct.pack(bo + off_fmt, ifd_off + ifd_size + len(extra))
                extra += payload
                if len(extra) % 2:
                    extra += b"\x00"
            body += e
        out += head + body + struct.pack(bo + off_fmt, 0) + extra
        out[prev_ptr : prev_ptr + vsz] = struct.pack(bo + off_fmt, ifd_off)
        prev_ptr = ifd_off + len(head) + esz * len(entries)
    with open(path, "wb") as f:
        f.write(bytes(out))


@pytest.mark.parametrize(
    "dtype,bo,rps,big,imagej",
    [(np.uint16, "<", None, False, False), (np.uint16, ">", 3, False, False), (np.uint8, "<", 4, False, False),
     (np.float32, "<", None, True, False), (np.int16, ">", 5, True, False), (np.uint16, ">", None, False, True)],
)
def test_plain_tiff_roundtrip(tmp_path, dtype, bo, rps, big, imagej):
    rng = np.random.default_rng(3)
    frames = (rng.uniform(0, 200, size=(7, 11, 9))).astype(dtype)
    path = str(tmp_path / "movie.tif")
    write_tiff(path, frames, bo=bo, rows_per_strip=rps, big=big, imagej=imagej)
    rd = _PlainTiff(path)
    assert rd.shape == frames.shape
    np.testing.assert_array_equal(rd.read([0, 3, 6]), frames[[0, 3, 6]])
    arr = TiffArray(path)
    if arr._tf is not None:
        pytest.skip("tifffile is installed: TiffArray uses it")
    assert arr.shape == (7, 11, 9) and arr.ndim == 3 and arr.dtype == np.float32
    got = arr[[1, 5]]
    assert got.dtype == np.float32 and got.shape == (2, 11, 9)
    np.testing.assert_array_equal(got, frames[[1, 5]].astype(np.float32))
    np.testing.assert_array_equal(arr[2], frames[2].astype(np.float32))
    np.testing.assert_array_equal(arr[1:6:2], frames[1:6:2].astype(np.float32))
    np.testing.assert_array_equal(arr[:, 2:5, 1], frames[:, 2:5, 1].astype(np.float32))


def test_plain_tiff_rejects_what_it_cannot_read(tmp_path):
    path = str(tmp_path / "bad.tif")
    with open(path, "wb") as f:
        f.write(b"not a tiff at all")
    with pytest.raises(ValueError):
        _PlainTiff(path)
    frames = np.zeros((2, 4, 4), np.uint16)
    good = str(tmp_path / "c.tif")
    write_tiff(good, frames)
    raw = bytearray(open(good, "rb").read())
    # flip the Compression tag (259) value of the first directory from 1 to 5 (LZW)
    idx = raw.find(struct.pack("<HHI", 259, 3, 1))
    raw[idx + 8 : idx + 10] = struct.pack("<H", 5)
    open(good, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="compressed"):
        _PlainTiff(good)
