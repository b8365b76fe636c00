"""Source-image decoders of the library (bt_image_load: PNG + TIFF, host only) against files written here from the PNG / TIFF 6.0
specifications: every PNG row filter, 16-bit gray and 8-bit RGB / RGBA; TIFF strips and tiles, none / Deflate / LZW,
predictor 1 and 2, little- and big-endian, signed 16-bit samples kept as bytes (src/formats/tiff.rs casts to R16Unorm)."""
import struct
import zlib

import numpy as np
import pytest

from bevy_terrain_b200 import _ffi as F
from bevy_terrain_b200 import api


# ------------------------------------------------------------------------------------------------ writers (test side)
def _png_filter_row(kind, row, prev, bpp):
    row = row.astype(np.int32)
    a = np.concatenate([np.zeros(bpp, np.int32), row[:-bpp]])
    b = prev.astype(np.int32)
    c = np.concatenate([np.zeros(bpp, np.int32), b[:-bpp]])
    if kind == 0:
        pred = 0
    elif kind == 1:
        pred = a
    elif kind == 2:
        pred = b
    elif kind == 3:
        pred = (a + b) // 2
    else:
        p = a + b - c
        pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
        pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
    return ((row - pred) & 0xFF).astype(np.uint8)


def write_png(path, img, idat_chunks=1):
    if img.dtype == np.uint16:
        colour, depth, raw = 0, 16, img.astype(">u2").view(np.uint8).reshape(img.shape[0], -1)
    else:
        colour, depth, raw = (2 if img.shape[2] == 3 else 6), 8, img.reshape(img.shape[0], -1)
    bpp = raw.shape[1] // img.shape[1]
    data = bytearray()
    prev = np.zeros(raw.shape[1], np.uint8)
    for y, row in enumerate(raw):
        kind = y % 5                     # cycle through None, Sub, Up, Average, Paeth
        data.append(kind)
        data += _png_filter_row(kind, row, prev, bpp).tobytes()
        prev = row
    comp = zlib.compress(bytes(data), 6)

    def chunk(tag, payload):
        return struct.pack(">I", len(payload)) + tag + payload + struct.pack(">I", zlib.crc32(tag + payload) & 0xFFFFFFFF)

    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", img.shape[1], img.shape[0], depth, colour, 0, 0, 0))
    out += chunk(b"tEXt", b"Comment\x00synthetic")
    step = (len(comp) + idat_chunks - 1) // idat_chunks
    for i in range(0, len(comp), step):
        out += chunk(b"IDAT", comp[i:i + step])
    out += chunk(b"IEND", b"")
    open(path, "wb").write(out)


def lzw_encode(data: bytes) -> bytes:
    """TIFF 6.0 section 13: MSB-first, ClearCode 256, EOI 257, code width bumped one code early, table reset before it fills"""
    out, acc, nbits = bytearray(), 0, 0

    def put(code, width):
        nonlocal acc, nbits
        acc = (acc << width) | code
        nbits += width
        while nbits >= 8:
            out.append((acc >> (nbits - 8)) & 0xFF)
            nbits -= 8

    table, nxt, width = {bytes([i]): i for i in range(256)}, 258, 9
    put(256, width)
    w = b""
    for byte in data:
        wk = w + bytes([byte])
        if wk in table:
            w = wk
            continue
        put(table[w], width)
        table[wk] = nxt
        nxt += 1
        if nxt == 4094:                  # libtiff: clear before the table is full
            put(256, width)
            table, nxt, width = {bytes([i]): i for i in range(256)}, 258, 9
        elif nxt > (1 << width) - 1:
            width += 1
        w = bytes([byte])
    if w:
        put(table[w], width)
    put(257, width)
    if nbits:
        out.append((acc << (8 - nbits)) & 0xFF)
    return bytes(out)


def write_tiff(path, img, big=False, compression=1, predictor=1, rows_per_strip=None, tile=None, signed=False):
    e = ">" if big else "<"
    h, w = img.shape[:2]
    samples = 1 if img.ndim == 2 else img.shape[2]
    bits = 16 if img.dtype.itemsize == 2 else 8

    def encode_block(block):   # block: (rows, cols[, samples]) in the image dtype
        b = block.astype(np.int64)
        if predictor == 2:
            d = b.copy()
            d[:, 1:] = b[:, 1:] - b[:, :-1]
            b = d
        b = (b & (0xFFFF if bits == 16 else 0xFF)).astype((e + "u2") if bits == 16 else np.uint8)
        raw = b.tobytes()
        if compression in (8, 32946):
            return zlib.compress(raw, 6)
        if compression == 5:
            return lzw_encode(raw)
        return raw

    blocks = []
    if tile:
        tw, th = tile
        for y in range(0, h, th):
            for x in range(0, w, tw):
                pad = np.zeros((th, tw) + img.shape[2:], dtype=img.dtype)
                part = img[y:y + th, x:x + tw]
                pad[:part.shape[0], :part.shape[1]] = part
                blocks.append(encode_block(pad))
    else:
        rps = rows_per_strip or h
        for y in range(0, h, rps):
            blocks.append(encode_block(img[y:y + rps]))
    entries = [(256, 4, [w]), (257, 4, [h]), (258, 3, [bits] * samples), (259, 3, [compression]), (262, 3, [1 if samples == 1 else 2]),
               (277, 3, [samples]), (284, 3, [1]), (317, 3, [predictor]), (339, 3, [2 if signed else 1] * samples)]
    if samples == 4:
        entries.append((338, 3, [2]))
    body = bytearray()
    offsets, base = [], 8
    for b in blocks:
        offsets.append(base + len(body))
        body += b
        if len(body) % 2:
            body += b"\x00"
    if tile:
        entries += [(322, 4, [tile[0]]), (323, 4, [tile[1]]), (324, 4, offsets), (325, 4, [len(b) for b in blocks])]
    else:
        entries += [(278, 4, [rows_per_strip or h]), (273, 4, offsets), (279, 4, [len(b) for b in blocks])]
    entries.sort()
    ifd_at = base + len(body)
    extra_at = ifd_at + 2 + 12 * len(entries) + 4
    ifd, extra = bytearray(struct.pack(e + "H", len(entries))), bytearray()
    for tag, typ, vals in entries:
        fmt = {3: "H", 4: "I"}[typ]
        payload = struct.pack(e + fmt * len(vals), *vals)
        if len(payload) <= 4:
            ifd += struct.pack(e + "HHI", tag, typ, len(vals)) + payload.ljust(4, b"\x00")
        else:
            ifd += struct.pack(e + "HHII", tag, typ, len(vals), extra_at + len(extra))
            extra += payload
            if len(extra) % 2:
                extra += b"\x00"
    ifd += struct.pack(e + "I", 0)
    header = (b"MM" if big else b"II") + struct.pack(e + "HI", 42, ifd_at)
    open(path, "wb").write(header + bytes(body) + bytes(ifd) + bytes(extra))


# ------------------------------------------------------------------------------------------------ tests
def _gray(h, w, seed):
    rng = np.random.default_rng(seed)
    smooth = (np.add.outer(np.arange(h), np.arange(w)) * 37 % 65536).astype(np.uint16)
    noise = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    return np.where(rng.random((h, w)) < 0.7, smooth, noise).astype(np.uint16)     # long runs AND incompressible parts


def _colour(h, w, channels, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, channels), dtype=np.uint8)
    img[h // 4: h // 2] = img[h // 4]                                             # repeated rows: deep LZW strings
    return img


def test_png_gray16_rgb8_rgba8(tmp_path):
    g = _gray(37, 53, 1)
    write_png(tmp_path / "h.png", g, idat_chunks=3)
    assert np.array_equal(api.load_image(str(tmp_path / "h.png")), g)
    rgb = _colour(29, 31, 3, 2)
    write_png(tmp_path / "c.png", rgb)
    got = api.load_image(str(tmp_path / "c.png"))
    assert np.array_equal(got[..., :3], rgb) and (got[..., 3] == 255).all()
    rgba = _colour(16, 45, 4, 3)
    write_png(tmp_path / "a.png", rgba)
    assert np.array_equal(api.load_image(str(tmp_path / "a.png")), rgba)


@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("compression", [1, 8, 5])
@pytest.mark.parametrize("predictor", [1, 2])
def test_tiff_gray16_strips(tmp_path, big, compression, predictor):
    g = _gray(70, 91, 4 + compression)
    write_tiff(tmp_path / "h.tif", g, big=big, compression=compression, predictor=predictor, rows_per_strip=16)
    assert np.array_equal(api.load_image(str(tmp_path / "h.tif")), g)


def test_tiff_tiles_signed_and_colour(tmp_path):
    g = _gray(100, 75, 9)
    write_tiff(tmp_path / "t.tif", g, compression=5, predictor=2, tile=(32, 48))
    assert np.array_equal(api.load_image(str(tmp_path / "t.tif")), g)
    dem = (np.random.default_rng(5).integers(-400, 9000, (40, 33))).astype(np.int16)     # a DEM with signed heights
    write_tiff(tmp_path / "s.tif", dem.view(np.uint16), compression=8, signed=True)
    assert np.array_equal(api.load_image(str(tmp_path / "s.tif")), dem.view(np.uint16))   # bytes kept: tiff.rs casts I16 to R16 texels
    rgb = _colour(33, 47, 3, 6)
    write_tiff(tmp_path / "c.tif", rgb, compression=5, rows_per_strip=8)
    got = api.load_image(str(tmp_path / "c.tif"))
    assert np.array_equal(got[..., :3], rgb) and (got[..., 3] == 255).all()
    rgba = _colour(20, 20, 4, 7)
    write_tiff(tmp_path / "a.tif", rgba, big=True, compression=8, predictor=2, tile=(16, 16))
    assert np.array_equal(api.load_image(str(tmp_path / "a.tif")), rgba)


def test_lzw_table_reset_on_a_large_strip(tmp_path):
    g = _gray(300, 400, 21)                          # 240 kB in one strip: the 4094-entry table is cleared many times
    write_tiff(tmp_path / "big.tif", g, compression=5)
    assert np.array_equal(api.load_image(str(tmp_path / "big.tif")), g)


def test_decoder_errors_are_reported(tmp_path):
    (tmp_path / "x.png").write_bytes(b"\x89PNG\r\n\x1a\n" + b"\x00" * 20)
    with pytest.raises(F.BtError) as e:
        api.load_image(str(tmp_path / "x.png"))
    assert e.value.status == F.BT_ERR_IO and "PNG" in str(e.value)
    with pytest.raises(F.BtError):
        api.load_image(str(tmp_path / "missing.tif"))
    g8 = np.zeros((4, 4, 3), np.uint8)
    write_png(tmp_path / "ok.png", g8)
    data = bytearray((tmp_path / "ok.png").read_bytes())
    data[40] ^= 0xFF                                  # corrupt a chunk: the CRC check must catch it
    (tmp_path / "bad.png").write_bytes(bytes(data))
    with pytest.raises(F.BtError):
        api.load_image(str(tmp_path / "bad.png"))


def test_truncated_and_corrupted_files_fail_cleanly(tmp_path):
    """untrusted input: every truncation / byte flip must end in BT_ERR_IO or in a decoded image, never in a crash"""
    g = _gray(40, 40, 30)
    write_png(tmp_path / "a.png", g)
    write_tiff(tmp_path / "a.tif", g, compression=5, predictor=2, rows_per_strip=7)
    write_tiff(tmp_path / "b.tif", _colour(24, 24, 3, 31), compression=8, tile=(16, 16), big=True)
    rng = np.random.default_rng(99)
    outcomes = {"ok": 0, "error": 0}
    for name in ("a.png", "a.tif", "b.tif"):
        data = (tmp_path / name).read_bytes()
        variants = [data[:n] for n in rng.integers(0, len(data), 40)]
        for _ in range(120):
            b = bytearray(data)
            for at in rng.integers(0, len(data), rng.integers(1, 4)):
                b[at] = rng.integers(0, 256)
            variants.append(bytes(b))
        for i, v in enumerate(variants):
            p = tmp_path / f"v{i}_{name}"
            p.write_bytes(v)
            try:
                img = api.load_image(str(p))
                assert img.ndim in (2, 3)
                outcomes["ok"] += 1
            except F.BtError as e:
                assert e.status == F.BT_ERR_IO
                outcomes["error"] += 1
    assert outcomes["error"] > 100          # truncations and most flips are detected (CRC, sizes, code tables)
