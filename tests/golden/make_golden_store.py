"""Writes tests/golden/store_liblz4.zip + store_liblz4_expected.npz: a zarr-v2 ZipStore laid out as the reference's
generator writes it (/root/reference/gnn_pressure_estimation/scenegenv7.py:664-725: groups `pressure`/`head`, arrays
`train`/`valid`/`test` of shape [scenarios, nodes], chunks (batch, nodes), root attrs `ordered_names_by_attr`), whose
chunk payloads are zarr's default compressor container (Blosc-1 frames, byte-shuffle, split streams) around
COMPRESSED STREAMS PRODUCED BY THE REAL LIBRARIES: liblz4 (through pyarrow's `lz4_raw` codec, raw LZ4 blocks) and
zlib (Python's binding of the real zlib).

Neither `zarr` nor `numcodecs` / c-blosc exist in this image (no network), so the 16-byte Blosc header, the block-start
table and the per-stream length prefixes are assembled here from the published c-blosc-1 frame format; the entropy-coded
bytes inside — the part a decoder can get wrong in interesting ways (overlapping matches, length extensions, last
literals) — come from liblz4 itself, not from a test-side encoder.  Run once: `python tests/golden/make_golden_store.py`.
"""
import io
import json
import os
import struct
import zipfile
import zlib

import numpy as np
import pyarrow as pa

HERE = os.path.dirname(os.path.abspath(__file__))


def liblz4_block(data: bytes) -> bytes:
    return pa.Codec("lz4_raw").compress(data, asbytes=True)


def blosc_frame(data: bytes, typesize: int, codec: str, blocksize: int) -> bytes:
    """c-blosc 1.x frame: version 2, versionlz 1, flags (bit0 byte-shuffle, bits 5-7 codec), typesize, nbytes,
    blocksize, cbytes; int32 bstarts[nblocks]; per block `typesize` split streams, each int32 length + payload
    (length == uncompressed stream length means stored)."""
    nbytes = len(data)
    flags = 1 | ({"lz4": 1, "zlib": 3}[codec] << 5)
    nblocks = -(-nbytes // blocksize)
    body, bstarts, pos = bytearray(), [], 16 + 4 * nblocks
    for b in range(nblocks):
        blk = data[b * blocksize:(b + 1) * blocksize]
        bsize = len(blk)
        nelem = bsize // typesize
        blk = np.frombuffer(blk, np.uint8, nelem * typesize).reshape(nelem, typesize).T.tobytes() + blk[nelem * typesize:]
        nsplits = typesize if (bsize == blocksize and bsize // typesize >= 128) else 1      # leftover block: one stream
        ne = bsize // nsplits
        bstarts.append(pos + len(body))
        for s in range(nsplits):
            part = blk[s * ne:(s + 1) * ne]
            comp = liblz4_block(part) if codec == "lz4" else zlib.compress(part, 5)
            if len(comp) >= ne:
                comp = part
            body += struct.pack("<i", len(comp)) + comp
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, 16 + 4 * nblocks + len(body))
    return head + struct.pack(f"<{nblocks}i", *bstarts) + bytes(body)


def main():
    rng = np.random.RandomState(20231017)
    nodes, batch = 19, 16
    names = [f"J{i}" for i in range(nodes - 2)] + ["R1", "T1"]
    # pressure-like fields: smooth across nodes, correlated across scenarios (compressible, with long matches),
    # plus a noisy array (mostly stored streams) and a constant one (maximal overlapping matches)
    base = np.cumsum(rng.randn(nodes)) * 3 + 55
    arrays = {
        "pressure/train": (base + np.round(rng.randn(53, nodes), 1)).astype(np.float64),
        "pressure/valid": (base + np.round(rng.randn(16, nodes), 2)).astype(np.float64),
        "pressure/test": rng.randn(7, nodes).astype(np.float64) * 1e3,
        "head/train": np.full((40, nodes), 71.25, dtype=np.float32),
        "head/valid": (np.arange(21 * nodes, dtype=np.float32).reshape(21, nodes) % 11) * 0.5,
    }
    codec_of = {"pressure/train": "lz4", "pressure/valid": "lz4", "pressure/test": "lz4", "head/train": "lz4", "head/valid": "zlib"}
    attrs = {"ordered_names_by_attr": {"pressure": names, "head": names}, "batch_size": batch}
    files = {".zgroup": json.dumps({"zarr_format": 2}), ".zattrs": json.dumps(attrs)}
    for path, arr in arrays.items():
        grp = path.split("/")[0]
        files[f"{grp}/.zgroup"] = json.dumps({"zarr_format": 2})
        ts = arr.dtype.itemsize
        files[f"{path}/.zarray"] = json.dumps({
            "zarr_format": 2, "shape": list(arr.shape), "chunks": [batch, nodes], "dtype": arr.dtype.str, "order": "C",
            "compressor": {"id": "blosc", "cname": codec_of[path], "clevel": 5, "shuffle": 1, "blocksize": 0},
            "filters": None, "fill_value": 0.0})
        for ci in range(-(-arr.shape[0] // batch)):
            chunk = np.zeros((batch, nodes), dtype=arr.dtype)
            rows = arr[ci * batch:(ci + 1) * batch]
            chunk[:rows.shape[0]] = rows
            files[f"{path}/{ci}.0"] = blosc_frame(chunk.tobytes(), ts, codec_of[path], 128 * ts * 2)
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_STORED) as z:          # zarr's ZipStore stores members uncompressed
        for k in sorted(files):
            info = zipfile.ZipInfo(k, date_time=(2023, 10, 17, 0, 0, 0))
            z.writestr(info, files[k] if isinstance(files[k], bytes) else files[k].encode())
    with open(os.path.join(HERE, "store_liblz4.zip"), "wb") as f:
        f.write(buf.getvalue())
    np.savez_compressed(os.path.join(HERE, "store_liblz4_expected.npz"), names=np.array(names),
                        **{k.replace("/", "__"): v for k, v in arrays.items()})
    # raw liblz4 blocks with their plaintexts: known-answer vectors for the LZ4 block decoder on its own
    kat = {}
    for i, data in enumerate((b"", b"a", b"abcabcabcabcabcabcabcabcabcabcabcabc", bytes(300), bytes(range(256)) * 3,
                              np.repeat(rng.randn(200).astype(np.float32), 5).tobytes(),
                              bytes(rng.randint(0, 256, 2000).astype(np.uint8)), b"xy" * 40000)):
        kat[f"plain{i}"] = np.frombuffer(data, np.uint8)
        kat[f"lz4_{i}"] = np.frombuffer(liblz4_block(data), np.uint8)
    np.savez_compressed(os.path.join(HERE, "liblz4_blocks.npz"), **kat)
    print("wrote", os.path.getsize(os.path.join(HERE, "store_liblz4.zip")), "bytes")


if __name__ == "__main__":
    main()
