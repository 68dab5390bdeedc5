"""zarr v2 snapshot-store reader.  zarr / numcodecs are not installed in this image, so the
fixtures are written here: plain zarr v2 metadata + chunks encoded by small test-side encoders (zlib / gzip from the
standard library; a Blosc-1 container writer with a greedy LZ4 block compressor).  That pins the reader to the
published formats' structure, not to bytes produced by zarr itself — stated in DESIGN.md."""
import gzip
import json
import os
import struct
import zipfile
import zlib

import numpy as np
import pytest
import torch

from gnn_pressure_estimation_b200 import snapshot_store as SS
from gnn_pressure_estimation_b200 import topology as T


# ----------------------------------------------------------------------------- test-side encoders
def lz4_block_encode(data: bytes) -> bytes:
    """greedy LZ4 block compressor (hash of 4-byte windows); emits real matches, incl. overlapping ones"""
    out, n, i, anchor, table = bytearray(), len(data), 0, 0, {}

    def emit(lit: bytes, mlen: int, offset: int):
        ll, ml = len(lit), mlen - 4 if mlen else 0
        out.append((min(ll, 15) << 4) | (min(ml, 15) if mlen else 0))
        if ll >= 15:
            r = ll - 15
            while r >= 255:
                out.append(255); r -= 255
            out.append(r)
        out.extend(lit)
        if mlen:
            out.extend(struct.pack("<H", offset))
            if ml >= 15:
                r = ml - 15
                while r >= 255:
                    out.append(255); r -= 255
                out.append(r)

    while i <= n - 12:                                     # LZ4 end rules: no match starts in the last 12 bytes,
                                                           # the last 5 bytes are always literals
        key = data[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 65535:
            m = 4
            while i + m < n - 5 and data[cand + m] == data[i + m]:
                m += 1
            emit(data[anchor:i], m, i - cand)
            i += m
            anchor = i
        else:
            i += 1
    emit(data[anchor:], 0, 0)
    return bytes(out)


def blosc_encode(data: bytes, typesize: int, shuffle: bool, codec: str = "lz4", blocksize: int = 4096,
                 split: bool = True, memcpy: bool = False) -> bytes:
    nbytes = len(data)
    flags = (1 if shuffle else 0) | (0 if split else 0x10) | ({"lz4": 1, "zlib": 3}[codec] << 5)
    if memcpy:
        return struct.pack("<BBBBIII", 2, 1, flags | 0x2, typesize, nbytes, blocksize, 16 + nbytes) + data
    nblocks = -(-nbytes // blocksize)
    body, bstarts, pos = bytearray(), [], 16 + 4 * nblocks
    for b in range(nblocks):
        blk = data[b * blocksize:(b + 1) * blocksize]
        bsize = len(blk)
        if shuffle and typesize > 1:
            nelem = bsize // typesize
            blk = np.frombuffer(blk, np.uint8, nelem * typesize).reshape(nelem, typesize).T.tobytes() + blk[nelem * typesize:]
        nsplits = typesize if (split and bsize == blocksize and 1 < typesize <= 16 and bsize // typesize >= 128) else 1
        ne = bsize // nsplits
        bstarts.append(pos + len(body))
        for sidx in range(nsplits):
            part = blk[sidx * ne:(sidx + 1) * ne]
            comp = lz4_block_encode(part) if codec == "lz4" else zlib.compress(part)
            if len(comp) >= ne:
                comp = part                                                  # stored stream
            body += struct.pack("<i", len(comp)) + comp
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, 16 + 4 * nblocks + len(body))
    return head + struct.pack(f"<{nblocks}i", *bstarts) + bytes(body)


def write_store(root, arrays, attrs, compressor, encode, as_zip, sep="."):
    """arrays: {path: (ndarray, chunks)}"""
    files = {".zgroup": json.dumps({"zarr_format": 2}), ".zattrs": json.dumps(attrs)}
    for path, (arr, chunks) in arrays.items():
        parts = path.split("/")
        for d in range(1, len(parts)):
            files["/".join(parts[:d]) + "/.zgroup"] = json.dumps({"zarr_format": 2})
        meta = {"zarr_format": 2, "shape": list(arr.shape), "chunks": list(chunks), "dtype": arr.dtype.str, "order": "C",
                "compressor": compressor, "filters": None, "fill_value": 0.0}
        if sep != ".":
            meta["dimension_separator"] = sep
        files[f"{path}/.zarray"] = json.dumps(meta)
        grid = [range(-(-s // c)) for s, c in zip(arr.shape, chunks)]
        for idx in np.ndindex(*[len(g) for g in grid]):
            chunk = np.zeros(chunks, dtype=arr.dtype)
            sel = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, arr.shape))
            chunk[tuple(slice(0, s.stop - s.start) for s in sel)] = arr[sel]
            files[f"{path}/" + sep.join(map(str, idx))] = encode(chunk.tobytes())
    if as_zip:
        with zipfile.ZipFile(root, "w") as z:
            for k, v in files.items():
                z.writestr(k, v if isinstance(v, bytes) else v.encode())
    else:
        for k, v in files.items():
            p = os.path.join(root, *k.split("/"))
            os.makedirs(os.path.dirname(p), exist_ok=True)
            with open(p, "wb") as f:
                f.write(v if isinstance(v, bytes) else v.encode())


# ----------------------------------------------------------------------------------------- tests
def test_lz4_known_answers():
    # literals "abc", match offset 3 length 9 (overlapping), then the literal-only last sequence
    assert SS.lz4_block_decode(bytes([0x35]) + b"abc" + b"\x03\x00" + bytes([0x00]), 12) == b"abcabcabcabc"
    # 20 literals (length extension byte), no match
    assert SS.lz4_block_decode(bytes([0xF0, 5]) + bytes(range(20)), 20) == bytes(range(20))
    # long match: 1 literal then offset 1, length 4 + 15 + 255 + 3 = 277 copies of it
    assert SS.lz4_block_decode(bytes([0x1F]) + b"z" + b"\x01\x00" + bytes([255, 3]) + bytes([0x00]), 278) == b"z" * 278
    with pytest.raises(SS.StoreError):
        SS.lz4_block_decode(bytes([0x10]) + b"a" + b"\x05\x00", 5)          # offset beyond the output
    rng = np.random.RandomState(0)
    for data in (b"", b"x", bytes(rng.randint(0, 4, 5000).astype(np.uint8)), bytes(rng.randint(0, 256, 3000).astype(np.uint8)),
                 np.repeat(rng.randn(300).astype(np.float32), 3).tobytes()):
        assert SS.lz4_block_decode(lz4_block_encode(data), len(data)) == data


@pytest.mark.parametrize("typesize,shuffle,codec,split,memcpy", [
    (4, True, "lz4", True, False), (8, True, "lz4", True, False), (4, False, "lz4", True, False),
    (4, True, "lz4", False, False), (8, True, "zlib", False, False), (4, True, "lz4", True, True), (1, True, "lz4", True, False)])
def test_blosc_container_round_trip(typesize, shuffle, codec, split, memcpy):
    rng = np.random.RandomState(typesize)
    base = np.cumsum(rng.randint(0, 3, 5000)).astype({1: np.uint8, 4: np.float32, 8: np.float64}[typesize])
    data = base.tobytes()[: 4096 * 3 + 1000 + (3 if typesize > 1 else 0)]      # 3 full blocks + a leftover, ragged tail
    enc = blosc_encode(data, typesize, shuffle, codec, 4096, split, memcpy)
    assert SS.blosc_decode(enc) == data
    with pytest.raises(SS.StoreError):
        SS.blosc_decode(enc[:-1])
    bad = bytearray(enc); bad[2] |= 0x4
    if not memcpy:
        with pytest.raises(SS.StoreError, match="bit-shuffle"):
            SS.blosc_decode(bytes(bad))


@pytest.mark.parametrize("kind", ["dir-null", "zip-zlib", "zip-gzip", "zip-blosc-f8", "dir-blosc-f4-slash"])
def test_read_array_and_attrs(kind, tmp_path):
    rng = np.random.RandomState(1)
    dt = np.float32 if "f4" in kind else np.float64
    train = (rng.randn(37, 12) * 20 + 60).astype(dt)                        # 37 rows in chunks of 8 -> ragged last chunk
    test = (rng.randn(5, 12) * 20 + 60).astype(dt)
    comp, enc = {"dir-null": (None, lambda b: b), "zip-zlib": ({"id": "zlib", "level": 1}, zlib.compress),
                 "zip-gzip": ({"id": "gzip", "level": 1}, gzip.compress),
                 "zip-blosc-f8": ({"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1, "blocksize": 0},
                                  lambda b: blosc_encode(b, 8, True, "lz4", 256, True)),
                 "dir-blosc-f4-slash": ({"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1, "blocksize": 0},
                                        lambda b: blosc_encode(b, 4, True, "lz4", 128 * 4, True))}[kind]
    root = str(tmp_path / ("store.zip" if kind.startswith("zip") else "store"))
    attrs = {"ordered_names_by_attr": {"pressure": ["J1", "J2"]}, "batch_size": 8}
    write_store(root, {"pressure/train": (train, (8, 12)), "pressure/test": (test, (8, 12))}, attrs, comp, enc,
                kind.startswith("zip"), sep="/" if "slash" in kind else ".")
    st = SS.ZarrV2Store(root)
    assert st.group_keys() == ["pressure"] and st.group_keys("pressure") == ["test", "train"]
    assert st.attrs() == attrs
    assert np.array_equal(st.read_array("pressure/train"), train)
    assert np.array_equal(st.read_array("pressure/test"), test)
    assert np.array_equal(st.read_array("pressure/train", rows=10), train[:10])
    with pytest.raises(SS.StoreError):
        st.read_array("head/train")


def test_snapshot_set_mirrors_wdn_dataset(tmp_path):
    """DataLoader.collect semantics: junction columns in registry order, statistics over the kept data, z-norm with
    eps, template edge_index in the reference's order"""
    wn = T.tiny_network()
    inp = tmp_path / "tiny.inp"
    inp.write_text(T.write_inp(wn))
    names = wn.node_names                                                    # junctions + reservoirs + tanks
    rng = np.random.RandomState(2)
    data = rng.rand(23, len(names)) * 50 + 20
    root = str(tmp_path / "tiny.zip")
    write_store(root, {"pressure/train": (data, (10, len(names)))}, {}, {"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1},
                lambda b: blosc_encode(b, 8, True, "lz4", 1024, True), True)
    ds = SS.SnapshotSet.load(str(inp), root, "pressure", "train", removal="keep_junction", norm_type="znorm", device="cpu")
    cols = [i for i, n in enumerate(names) if n in set(wn.junctions)]
    kept = data[:, cols]
    assert len(ds) == 23 and ds.num_nodes == len(wn.junctions) and ds.node_names == list(wn.junctions)
    assert ds.mean == pytest.approx(kept.mean()) and ds.std == pytest.approx(kept.std())
    assert np.allclose(ds.snapshots.numpy(), ((kept - kept.mean()) / (kept.std() + 1e-8)).astype(np.float32))
    ei, kept_names = T.reference_edge_index(wn, "keep_junction")
    assert np.array_equal(ds.edge_index.numpy(), ei) and kept_names == ds.node_names
    # given statistics (validation / test sets reuse the training statistics), first rows only, all nodes, min-max
    ds2 = SS.SnapshotSet.load(str(inp), root, "pressure", "train", num_records=7, removal="keep_all", norm_type="minmax",
                              min=10.0, max=90.0, device="cpu")
    assert len(ds2) == 7 and ds2.num_nodes == len(names)
    assert np.allclose(ds2.snapshots.numpy(), ((data[:7] - 10.0) / 80.0).astype(np.float32))
    got = list(ds.batches(10))
    assert [b.numel() for b in got] == [10 * ds.num_nodes, 10 * ds.num_nodes, 3 * ds.num_nodes]
    assert torch.equal(torch.cat(got), ds.snapshots.reshape(-1))
    with pytest.raises(SS.StoreError):
        SS.SnapshotSet.load(str(inp), root, "head", "train", device="cpu")


def test_missing_chunks_fill_value_and_fortran_order(tmp_path):
    """a chunk that was never written reads as fill_value; order 'F' chunks are transposed back"""
    root = tmp_path / "s"
    arr = np.arange(6 * 4, dtype=np.float32).reshape(6, 4)
    os.makedirs(root / "a")
    (root / ".zgroup").write_text(json.dumps({"zarr_format": 2}))
    meta = {"zarr_format": 2, "shape": [6, 4], "chunks": [3, 4], "dtype": "<f4", "order": "F", "compressor": None,
            "filters": None, "fill_value": "NaN"}
    (root / "a" / ".zarray").write_text(json.dumps(meta))
    (root / "a" / "0.0").write_bytes(np.asfortranarray(arr[:3]).tobytes(order="F"))       # chunk 1.0 is missing
    got = SS.ZarrV2Store(str(root)).read_array("a")
    assert np.array_equal(got[:3], arr[:3]) and np.isnan(got[3:]).all()
    meta["zarr_format"] = 3
    (root / "a" / ".zarray").write_text(json.dumps(meta))
    with pytest.raises(SS.StoreError, match="v2"):
        SS.ZarrV2Store(str(root)).read_array("a")


# ------------------------------------------------------------- fixtures whose compressed bytes come from real libraries
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_lz4_decoder_on_blocks_written_by_liblz4():
    """known answers from the real liblz4 (pyarrow's lz4_raw codec; tests/golden/make_golden_store.py): the decoder is
    held to bytes it did not produce itself"""
    kat = np.load(os.path.join(GOLDEN, "liblz4_blocks.npz"))
    n = len([k for k in kat.files if k.startswith("plain")])
    assert n >= 8
    for i in range(n):
        plain, comp = kat[f"plain{i}"].tobytes(), kat[f"lz4_{i}"].tobytes()
        assert SS.lz4_block_decode(comp, len(plain)) == plain, i
    try:                                                     # live cross-check when pyarrow is importable
        import pyarrow as pa
    except Exception:
        return
    rng = np.random.RandomState(3)
    for size in (1, 13, 64, 4096, 70000):
        data = bytes(rng.randint(0, 3, size).astype(np.uint8))
        assert SS.lz4_block_decode(pa.Codec("lz4_raw").compress(data, asbytes=True), size) == data
        ours = lz4_block_encode(data)                        # and the test-side encoder is valid LZ4 for liblz4
        assert pa.Codec("lz4_raw").decompress(ours, decompressed_size=size, asbytes=True) == data


def test_store_with_liblz4_and_zlib_streams(tmp_path):
    """the committed ZipStore (reference generator layout, scenegenv7.py:664-725; Blosc frames around liblz4 / zlib
    streams) reads back exactly, through ZarrV2Store and through SnapshotSet.load"""
    want = np.load(os.path.join(GOLDEN, "store_liblz4_expected.npz"))
    st = SS.ZarrV2Store(os.path.join(GOLDEN, "store_liblz4.zip"))
    assert st.group_keys() == ["head", "pressure"] and st.group_keys("pressure") == ["test", "train", "valid"]
    names = [str(s) for s in want["names"]]
    assert st.attrs()["ordered_names_by_attr"]["pressure"] == names and st.attrs()["batch_size"] == 16
    for key in ("pressure/train", "pressure/valid", "pressure/test", "head/train", "head/valid"):
        got = st.read_array(key)
        exp = want[key.replace("/", "__")]
        assert got.dtype == exp.dtype and np.array_equal(got, exp), key
    # the same file through the dataset mirror (utils/DataLoader.py:206-258): junction columns, z-norm
    wn = T.WaterNetwork(junctions=names[:-2], reservoirs=["R1"], tanks=["T1"],
                        links=[(f"P{i}", names[i], names[i + 1]) for i in range(len(names) - 1)])
    inp = tmp_path / "line.inp"
    inp.write_text(T.write_inp(wn))
    ds = SS.SnapshotSet.load(str(inp), os.path.join(GOLDEN, "store_liblz4.zip"), "pressure", "valid", device="cpu")
    kept = want["pressure__valid"][:, :len(names) - 2]
    assert len(ds) == 16 and ds.num_nodes == len(names) - 2
    assert np.allclose(ds.snapshots.numpy(), ((kept - kept.mean()) / (kept.std() + 1e-8)).astype(np.float32))
