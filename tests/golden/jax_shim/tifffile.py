"""Placeholder so `import tifffile` in the reference's dataset.py succeeds; TiffArray is never used."""


def TiffFile(*a, **k):
    raise RuntimeError("tifffile is not available in this image")


def imread(*a, **k):
    raise RuntimeError("tifffile is not available in this image")
