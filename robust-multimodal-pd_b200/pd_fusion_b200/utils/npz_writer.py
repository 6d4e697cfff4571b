"""`np.savez_compressed` with the deflate work spread over threads.

The MIL builder stores `embeddings [S, L, D] f32` with `np.savez_compressed`
(scripts/build_resnet2d_mil_embeddings.py:160-168): 3.9 GB at 10 000 subjects, deflated by ONE zlib stream on one core --
minutes of wall time behind a GPU path that produces those embeddings in seconds.  The file written here is the same thing to
every reader (`np.load(path, allow_pickle=True)`, data/openneuro_features.py:302): a ZIP archive whose members are `.npy` files
compressed with DEFLATE.  Only the way the DEFLATE stream is produced differs: the member's bytes are cut into chunks, each chunk is
compressed independently by a worker thread (zlib releases the GIL) and closed with a full flush, and the pieces are concatenated --
a valid single DEFLATE stream (the pigz construction, without a shared dictionary across chunks).
"""
from __future__ import annotations

import io
import os
import time
import zipfile
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

_CHUNK = 8 << 20            # bytes of uncompressed data per deflate job
_SMALL = 1 << 20            # members below this (and object arrays) go through numpy's own writer


def _deflate_chunk(args):
    view, level, last = args
    c = zlib.compressobj(level, zlib.DEFLATED, -15)                       # raw deflate, as ZIP members hold it
    out = c.compress(view)
    out += c.flush(zlib.Z_FINISH if last else zlib.Z_FULL_FLUSH)          # full flush: byte-aligned, no final-block bit
    return out


def _npy_header(arr: np.ndarray) -> bytes:
    buf = io.BytesIO()
    np.lib.format.write_array_header_1_0(buf, np.lib.format.header_data_from_array_1_0(arr)) if arr.nbytes < (1 << 31) else \
        np.lib.format.write_array_header_2_0(buf, np.lib.format.header_data_from_array_1_0(arr))
    return buf.getvalue()


def savez_compressed_parallel(path, threads: int | None = None, level: int = 6, **arrays: np.ndarray) -> Path:
    """Drop-in for `np.savez_compressed(path, **arrays)`; returns the path written (".npz" appended like numpy does)."""
    path = Path(path)
    if path.suffix != ".npz":
        path = path.with_name(path.name + ".npz")
    threads = threads or max(1, min(16, os.cpu_count() or 1))
    with zipfile.ZipFile(path, mode="w", compression=zipfile.ZIP_DEFLATED, allowZip64=True) as zf, ThreadPoolExecutor(threads) as pool:
        for name, value in arrays.items():
            arr = np.asanyarray(value)
            member = name + ".npy"
            if arr.dtype.hasobject or arr.nbytes < _SMALL:
                with zf.open(member, "w", force_zip64=True) as fid:                 # numpy's own route (pickles object arrays)
                    np.lib.format.write_array(fid, arr, allow_pickle=True)
                continue
            arr = np.ascontiguousarray(arr)
            payload = memoryview(arr).cast("B")
            header = _npy_header(arr)
            jobs = [(memoryview(header), level, False)]
            n = len(payload)
            jobs += [(payload[o:o + _CHUNK], level, o + _CHUNK >= n) for o in range(0, n, _CHUNK)]
            zinfo = zipfile.ZipInfo(member, date_time=time.localtime(time.time())[:6])
            zinfo.compress_type = zipfile.ZIP_DEFLATED
            zinfo.external_attr = 0o600 << 16
            zinfo.file_size = len(header) + n
            zinfo.compress_size = 0
            zinfo.CRC = 0
            fp = zf.fp
            zinfo.header_offset = fp.tell()
            fp.write(zinfo.FileHeader(zip64=True))                                  # placeholder; rewritten below with the real sizes
            crc, csize = 0, 0
            # workers deflate, this thread checksums the same bytes (zlib.crc32 releases the GIL too) and writes the pieces in order
            for (view, _, _), blob in zip(jobs, pool.map(_deflate_chunk, jobs)):
                crc = zlib.crc32(view, crc)
                fp.write(blob)
                csize += len(blob)
            end = fp.tell()
            zinfo.CRC, zinfo.compress_size = crc, csize
            fp.seek(zinfo.header_offset)
            fp.write(zinfo.FileHeader(zip64=True))                                  # same length: the zip64 extra field is forced
            fp.seek(end)
            zf.filelist.append(zinfo)                                                # what ZipFile.writestr does after its own write
            zf.NameToInfo[zinfo.filename] = zinfo
            zf.start_dir = end
            zf._didModify = True
    return path
