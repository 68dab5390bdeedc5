"""Snapshot-store ingestion without zarr / wntr.

The reference keeps its simulated snapshots in a zarr **v2** store, a directory or a ``ZipStore`` archive with one
array per feature and split (``/{pressure,head,...}/{train,valid,test}``, float arrays ``[S, nodes]`` chunked
``(batch, nodes)``, written at /root/reference/gnn_pressure_estimation/scenegenv7.py:664-725) and turns every row
into a PyG ``Data`` object behind a DataLoader (utils/DataLoader.py:62-183, 206-258).  This module reads the same
store with the standard library + numpy and hands the training / evaluation drivers ONE device-resident ``[S, N]``
tensor plus the template ``edge_index`` in the reference's order, so a step needs no collation and no H2D copy of
inputs at all (x = y = pressures, utils/auxil.py:96-97).

zarr v2 layout restated here (zarr 2.14.2 is what the reference pins, requirements.txt l.79; it is NOT installed in
this image, so the codec paths are exercised against fixtures written by tests/test_snapshot_store.py itself):
  * ``<path>/.zarray``  JSON: shape, chunks, dtype (numpy typestr), order C/F, compressor {id, ...} | null, filters,
    fill_value, dimension_separator ("." default);  ``<path>/.zattrs`` / ``.zgroup`` JSON
  * chunk key ``<path>/i.j``; a missing chunk is all ``fill_value``; edge chunks are stored full-size.
  * compressor ids handled: null, zlib, gzip, bz2, lzma (stdlib) and **blosc** — zarr's default
    (``Blosc(cname='lz4', clevel=5, shuffle=SHUFFLE)``) — through a pure-Python reader of the Blosc-1 container
    (16-byte header, block offsets, per-block split streams, byte-shuffle) with inner codecs lz4 / lz4hc
    (own LZ4 block decoder) and zlib; blosclz, snappy, zstd and bit-shuffle raise with a clear message.
Host-side, one-off per run; nothing here is on the hot path.
"""
from __future__ import annotations

import bz2
import json
import lzma
import os
import struct
import zipfile
import zlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import topology as _topology


class StoreError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------ codecs
def lz4_block_decode(src: bytes, uncompressed_size: int) -> bytes:
    """LZ4 block format: sequences of [token][literal length ext][literals][offset LE16][match length ext]."""
    out = bytearray()
    i, n = 0, len(src)
    while i < n:
        token = src[i]
        i += 1
        lit = token >> 4
        if lit == 15:
            while True:
                b = src[i]
                i += 1
                lit += b
                if b != 255:
                    break
        out += src[i:i + lit]
        i += lit
        if i >= n:                                   # the last sequence has literals only
            break
        offset = src[i] | (src[i + 1] << 8)
        i += 2
        if offset == 0 or offset > len(out):
            raise StoreError("corrupt LZ4 stream (bad match offset)")
        mlen = (token & 15) + 4
        if (token & 15) == 15:
            while True:
                b = src[i]
                i += 1
                mlen += b
                if b != 255:
                    break
        start = len(out) - offset
        if offset >= mlen:
            out += out[start:start + mlen]
        else:                                        # overlapping match: the pattern repeats
            pattern = bytes(out[start:])
            reps = -(-mlen // offset)
            out += (pattern * reps)[:mlen]
    if len(out) != uncompressed_size:
        raise StoreError(f"LZ4 stream decoded to {len(out)} bytes, expected {uncompressed_size}")
    return bytes(out)


_BLOSC_CODECS = {0: "blosclz", 1: "lz4", 2: "snappy", 3: "zlib", 4: "zstd"}


def blosc_decode(buf: bytes) -> bytes:
    """Blosc-1 container (c-blosc 1.x, what numcodecs.Blosc writes)."""
    if len(buf) < 16:
        raise StoreError("blosc chunk shorter than its header")
    version, _versionlz, flags, typesize = buf[0], buf[1], buf[2], buf[3]
    nbytes, blocksize, cbytes = struct.unpack_from("<III", buf, 4)
    if version != 2:
        raise StoreError(f"unsupported blosc format version {version}")
    if cbytes != len(buf):
        raise StoreError(f"blosc header says {cbytes} compressed bytes, chunk has {len(buf)}")
    if flags & 0x2:                                  # memcpyed: raw payload after the header
        if len(buf) - 16 != nbytes:
            raise StoreError("corrupt memcpyed blosc chunk")
        return bytes(buf[16:])
    if flags & 0x4:
        raise StoreError("blosc bit-shuffle is not supported by this reader (re-save with shuffle=SHUFFLE or NOSHUFFLE)")
    codec = _BLOSC_CODECS.get(flags >> 5)
    if codec not in ("lz4", "zlib"):
        raise StoreError(f"blosc inner codec {codec!r} is not supported by this reader (lz4, lz4hc and zlib are)")
    shuffle, dont_split = bool(flags & 0x1), bool(flags & 0x10)
    if nbytes == 0:
        return b""
    if blocksize <= 0:
        raise StoreError("corrupt blosc header (blocksize)")
    nblocks = -(-nbytes // blocksize)
    bstarts = struct.unpack_from(f"<{nblocks}i", buf, 16)
    out = bytearray(nbytes)
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        nsplits = typesize if (not dont_split and not leftover and 1 < typesize <= 16 and bsize // typesize >= 128) else 1
        neblock = bsize // nsplits
        pos = bstarts[b]
        parts: List[bytes] = []
        for _ in range(nsplits):
            (cb,) = struct.unpack_from("<i", buf, pos)
            pos += 4
            if cb < 0 or pos + cb > len(buf):
                raise StoreError("corrupt blosc block (stream length)")
            data = bytes(buf[pos:pos + cb])
            pos += cb
            if cb == neblock:
                parts.append(data)                   # stored
            elif codec == "lz4":
                parts.append(lz4_block_decode(data, neblock))
            else:
                d = zlib.decompress(data)
                if len(d) != neblock:
                    raise StoreError("corrupt blosc block (zlib stream length)")
                parts.append(d)
        block = b"".join(parts)
        if shuffle and typesize > 1:
            nelem = bsize // typesize
            body = np.frombuffer(block, dtype=np.uint8, count=nelem * typesize).reshape(typesize, nelem).T.tobytes()
            block = body + block[nelem * typesize:]
        out[b * blocksize:b * blocksize + bsize] = block
    return bytes(out)


def decode_chunk(buf: bytes, compressor: Optional[dict]) -> bytes:
    if compressor is None:
        return buf
    cid = compressor.get("id")
    if cid == "blosc":
        return blosc_decode(buf)
    if cid == "zlib":
        return zlib.decompress(buf)
    if cid == "gzip":
        return zlib.decompress(buf, 16 + zlib.MAX_WBITS)
    if cid == "bz2":
        return bz2.decompress(buf)
    if cid == "lzma":
        if compressor.get("format", 1) != 1 or compressor.get("filters"):
            raise StoreError("only the default lzma container (FORMAT_XZ, no custom filters) is supported")
        return lzma.decompress(buf)
    raise StoreError(f"unsupported zarr compressor {cid!r}")


# ------------------------------------------------------------------------------------------------- store
class ZarrV2Store:
    """Read-only key/value view of a zarr v2 DirectoryStore or ZipStore."""

    def __init__(self, path: str):
        self.path = path
        self._zip: Optional[zipfile.ZipFile] = None
        if os.path.isdir(path):
            self._names = None
        elif os.path.isfile(path) and zipfile.is_zipfile(path):
            self._zip = zipfile.ZipFile(path, "r")
            self._names = set(self._zip.namelist())
        else:
            raise StoreError(f"{path} is neither a zarr directory store nor a zip store")

    def get(self, key: str) -> Optional[bytes]:
        key = key.strip("/")
        if self._zip is not None:
            return self._zip.read(key) if key in self._names else None
        p = os.path.join(self.path, *key.split("/"))
        if not os.path.isfile(p):
            return None
        with open(p, "rb") as f:
            return f.read()

    def _json(self, key: str) -> Optional[dict]:
        raw = self.get(key)
        return None if raw is None else json.loads(raw.decode("utf-8"))

    def attrs(self, group: str = "") -> dict:
        return self._json(f"{group}/.zattrs" if group else ".zattrs") or {}

    def group_keys(self, group: str = "") -> List[str]:
        """names of the sub-groups / arrays directly under `group`"""
        prefix = group.strip("/") + "/" if group.strip("/") else ""
        if self._zip is not None:
            keys = self._names
        else:
            keys = set()
            for root, _dirs, files in os.walk(self.path):
                rel = os.path.relpath(root, self.path).replace(os.sep, "/")
                for f in files:
                    keys.add(f if rel == "." else f"{rel}/{f}")
        out = set()
        for k in keys:
            if k.startswith(prefix) and (k.endswith("/.zarray") or k.endswith("/.zgroup")):
                rest = k[len(prefix):].split("/")
                if len(rest) == 2:
                    out.add(rest[0])
        return sorted(out)

    def read_array(self, path: str, rows: Optional[int] = None) -> np.ndarray:
        """The whole array (or its first `rows` rows along axis 0) as a C-contiguous numpy array."""
        path = path.strip("/")
        meta = self._json(f"{path}/.zarray")
        if meta is None:
            raise StoreError(f"no array at {path!r} in {self.path}")
        if meta.get("zarr_format") != 2:
            raise StoreError(f"zarr_format {meta.get('zarr_format')} is not supported (v2 only)")
        if meta.get("filters"):
            raise StoreError("zarr filters are not supported by this reader")
        shape, chunks = tuple(meta["shape"]), tuple(meta["chunks"])
        dtype, order = np.dtype(meta["dtype"]), meta.get("order", "C")
        sep = meta.get("dimension_separator", ".")
        fill = meta.get("fill_value")
        fill = 0 if fill is None else (float(fill) if isinstance(fill, str) else fill)      # "NaN", "Infinity" strings
        want = shape if rows is None else (min(rows, shape[0]),) + shape[1:]
        out = np.full(want, fill, dtype=dtype)
        grid = [range(-(-w // c)) for w, c in zip(want, chunks)]
        for idx in np.ndindex(*[len(g) for g in grid]):
            raw = self.get(f"{path}/" + sep.join(str(i) for i in idx))
            if raw is None:
                continue
            data = decode_chunk(raw, meta.get("compressor"))
            if len(data) != int(np.prod(chunks)) * dtype.itemsize:
                raise StoreError(f"chunk {idx} of {path!r} decodes to {len(data)} bytes, expected a full {chunks} chunk")
            chunk = np.frombuffer(data, dtype=dtype).reshape(chunks, order=order)
            sel_out = tuple(slice(i * c, min((i + 1) * c, w)) for i, c, w in zip(idx, chunks, want))
            sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
            out[sel_out] = chunk[sel_in]
        return out


# -------------------------------------------------------------------------------------- dataset on the GPU
def keep_list(wn: _topology.WaterNetwork, removal: str, attrs: dict, feature: str) -> Optional[List[str]]:
    """utils/DataLoader.py:40-58 (get_keep_list)"""
    if removal == "keep_list":
        if "ordered_name_list" in attrs:
            return list(attrs["ordered_name_list"])
        if feature in attrs.get("ordered_names_by_attr", {}):
            return list(attrs["ordered_names_by_attr"][feature])
        return list(wn.junctions)
    if removal == "reservoir":
        return [n for n in wn.node_names if n not in set(wn.reservoirs)] if wn.reservoirs else None
    if removal == "tank":
        return [n for n in wn.node_names if n not in set(wn.tanks)] if wn.tanks else None
    if removal == "keep_junction":
        return list(wn.junctions)
    if removal == "keep_all":
        return None
    raise ValueError(f"Removal only supports keep_list,reservoir,tank,keep_junction,keep_all. Got {removal}")


@dataclass
class SnapshotSet:
    """What `WDNDataset` (utils/DataLoader.py:62-183) holds, as tensors: the normalised snapshots of one network,
    resident on the device, the template edge_index in the reference's order and the normalisation statistics."""
    snapshots: torch.Tensor           # float32 [S, N] (scaled when norm_type is znorm / minmax)
    edge_index: torch.Tensor          # int64 [2, E]
    node_names: List[str]
    norm_type: str
    mean: float
    std: float
    min: float
    max: float

    def __len__(self) -> int:
        return int(self.snapshots.shape[0])

    @property
    def num_nodes(self) -> int:
        return int(self.snapshots.shape[1])

    @staticmethod
    def load(input_path: str, zip_file_path: str, feature: str = "pressure", from_set: str = "train",
             num_records: Optional[int] = None, removal: str = "keep_junction", norm_type: str = "znorm",
             mean=None, std=None, min=None, max=None, device="cuda") -> "SnapshotSet":
        """DataLoader.collect (:206-258) + the statistics / scaling of WDNDataset.__init__ (:142-157)."""
        if from_set not in ("train", "valid", "test"):
            raise ValueError(f"from_set {from_set} is not supported")
        if norm_type not in ("znorm", "minmax", "unused"):
            raise ValueError("norm_type must be znorm, minmax or unused")
        store = ZarrV2Store(zip_file_path)
        if feature not in store.group_keys():
            raise StoreError(f"feature {feature} is unavailable in zarr file {zip_file_path}")
        wn = _topology.parse_inp(input_path)
        keep = keep_list(wn, removal, store.attrs(), feature)
        array = store.read_array(f"{feature}/{from_set}", rows=num_records)
        names = wn.node_names
        if keep is not None:
            kept = set(keep)
            taken = [i for i, name in enumerate(names) if name in kept]          # registry order (:244-250)
            array = np.take(array, taken, axis=-1)
            names = [names[i] for i in taken]
        ei, graph_names = _topology.reference_edge_index_for(wn, keep)
        if graph_names != names:
            raise StoreError("node order of the data columns and of the graph template disagree")
        flat = array.reshape(-1)
        st = dict(mean=float(np.mean(flat)) if mean is None else float(mean), std=float(np.std(flat)) if std is None else float(std),
                  min=float(np.min(flat)) if min is None else float(min), max=float(np.max(flat)) if max is None else float(max))
        if norm_type == "znorm":
            array = (array - st["mean"]) / (st["std"] + 1e-8)                    # utils/auxil.py:37-39
        elif norm_type == "minmax":
            array = (array - st["min"]) / (st["max"] - st["min"])
        snaps = torch.from_numpy(np.ascontiguousarray(array, dtype=np.float32)).to(device)
        return SnapshotSet(snaps, torch.from_numpy(ei), names, norm_type, **st)

    def batches(self, batch_size: int, shuffle: bool = False, generator: Optional[torch.Generator] = None):
        """device-side batches [B*N] (the last one smaller, like the reference loader without drop_last)"""
        S = len(self)
        order = torch.randperm(S, generator=generator).to(self.snapshots.device) if shuffle else None
        for s0 in range(0, S, batch_size):
            rows = self.snapshots[s0:s0 + batch_size] if order is None else self.snapshots[order[s0:s0 + batch_size]]
            yield rows.reshape(-1)
