"""Texture-file and cube-map readers of the facade (kuafu_b200/host/src/image_io.cpp):
what the reference gets from stb_image (`stbi_load(..., STBI_rgb_alpha)`, vkCore.hpp:1761-1795) and
libktx (`scene.cpp:294-309`, vkCore.hpp:1853-1971).  Files are written here -- PNGs by a small encoder of
this test (all colour types, bit depths, the five scanline filters, Adam7) and by PIL where it is
installed, KTX1 by hand -- read back through the facade's reader and compared texel for texel; PIL is
the independent decoder.  Malformed files must be refused, not crash."""
import os
import struct
import zlib

import numpy as np
import pytest

from kuafu_b200 import host

REF_PATTERN = "/root/reference/resources/patterns/fakesense_j415.png"


def _chunk(kind, body):
    return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def _filter_rows(rows, bpp):
    """rows: list of bytes objects (packed scanlines); filter type cycles 0..4 down the image."""
    out = bytearray()
    prev = bytes(len(rows[0])) if rows else b""
    for y, cur in enumerate(rows):
        ft = y % 5
        line = bytearray(len(cur))
        for x in range(len(cur)):
            a = cur[x - bpp] if x >= bpp else 0
            b = prev[x]
            c = prev[x - bpp] if x >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[ft]
            line[x] = (cur[x] - pred) & 0xFF
        out.append(ft)
        out += line
        prev = cur
    return bytes(out)


def _pack_rows(samples, bit_depth):
    """samples: (h, w, ch) integer array of sample values -> packed scanlines."""
    h, w, ch = samples.shape
    rows = []
    for y in range(h):
        if bit_depth == 8:
            rows.append(samples[y].astype("u1").tobytes())
        elif bit_depth == 16:
            rows.append(samples[y].astype(">u2").tobytes())
        else:
            bits = np.zeros(((w * bit_depth + 7) // 8) * 8, "u1")
            for x in range(w):
                v = int(samples[y, x, 0])
                for k in range(bit_depth):
                    bits[x * bit_depth + k] = (v >> (bit_depth - 1 - k)) & 1
            rows.append(np.packbits(bits).tobytes())
    return rows


def write_png(path, samples, color_type, bit_depth, interlace=False, palette=None, trns=None, idat_split=1):
    h, w, ch = samples.shape
    bpp = max(1, ch * bit_depth // 8)
    if not interlace:
        raw = _filter_rows(_pack_rows(samples, bit_depth), bpp)
    else:
        raw = b""
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = samples[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += _filter_rows(_pack_rows(sub, bit_depth), bpp)
    z = zlib.compress(raw, 6)
    parts = [z[i * len(z) // idat_split:(i + 1) * len(z) // idat_split] for i in range(idat_split)]
    data = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bit_depth, color_type, 0, 0, int(interlace)))
    if palette is not None:
        data += _chunk(b"PLTE", np.asarray(palette, "u1").tobytes())
    if trns is not None:
        data += _chunk(b"tRNS", np.asarray(trns, "u1").tobytes())
    data += _chunk(b"tEXt", b"Comment\0written by tests/test_cpu_image_io.py")
    for p in parts:
        data += _chunk(b"IDAT", p)
    data += _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(data)


def expected_rgba(samples, color_type, bit_depth, palette=None, trns=None):
    """What stb hands over for STBI_rgb_alpha, 8 bits per channel."""
    h, w, ch = samples.shape
    s = samples.astype(np.int64)
    if bit_depth == 16:
        s = s >> 8
    out = np.full((h, w, 4), 255, "u1")
    if color_type == 0:
        scale = 255 // ((1 << bit_depth) - 1) if bit_depth < 8 else 1
        out[..., 0] = out[..., 1] = out[..., 2] = s[..., 0] * scale
    elif color_type == 2:
        out[..., :3] = s
    elif color_type == 3:
        pal = np.asarray(palette, "u1").reshape(-1, 3)
        out[..., :3] = pal[s[..., 0]]
        if trns is not None:
            t = np.full(len(pal), 255, "u1")
            t[:len(trns)] = trns
            out[..., 3] = t[s[..., 0]]
    elif color_type == 4:
        out[..., 0] = out[..., 1] = out[..., 2] = s[..., 0]
        out[..., 3] = s[..., 1]
    else:
        out[...] = s
    return out


CASES = [  # colour type, bit depth, channels
    (0, 1, 1), (0, 2, 1), (0, 4, 1), (0, 8, 1), (0, 16, 1),
    (2, 8, 3), (2, 16, 3),
    (3, 1, 1), (3, 2, 1), (3, 4, 1), (3, 8, 1),
    (4, 8, 2), (4, 16, 2),
    (6, 8, 4), (6, 16, 4),
]


@pytest.mark.parametrize("color_type,bit_depth,ch", CASES)
@pytest.mark.parametrize("interlace", [False, True])
def test_png_reader_matches_the_written_samples(tmp_path, built, color_type, bit_depth, ch, interlace):
    rng = np.random.default_rng(1000 * color_type + 10 * bit_depth + int(interlace))
    w, h = 37, 23  # odd sizes: partial bytes at low bit depths, ragged Adam7 passes
    hi = 1 << bit_depth
    palette = trns = None
    if color_type == 3:
        n = min(hi, 200)
        palette = rng.integers(0, 256, (n, 3))
        trns = rng.integers(0, 256, n // 2)
        hi = n
    samples = rng.integers(0, hi, (h, w, ch))
    path = tmp_path / f"t{color_type}_{bit_depth}_{int(interlace)}.png"
    write_png(path, samples, color_type, bit_depth, interlace, palette, trns, idat_split=3)
    got = host.read_texture(path)
    want = expected_rgba(samples, color_type, bit_depth, palette, trns)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    try:
        from PIL import Image
    except ImportError:
        return
    if bit_depth == 16 and color_type != 0:
        return  # PIL reduces 16-bit colour its own way; the samples above are the reference
    pil = np.asarray(Image.open(path).convert("RGBA"))
    if bit_depth == 16:
        return
    assert np.array_equal(got, pil), "disagrees with PIL's decoder"


def test_png_written_by_pil_reads_back(tmp_path, built):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (64, 48, 4), dtype=np.uint8)
    for mode, arr in (("RGBA", img), ("RGB", img[..., :3]), ("L", img[..., 0]), ("LA", img[..., :2])):
        p = tmp_path / f"pil_{mode}.png"
        Image.fromarray(arr.squeeze(), mode).save(p, optimize=True)
        got = host.read_texture(p)
        assert np.array_equal(got, np.asarray(Image.open(p).convert("RGBA"))), mode
    pal = Image.fromarray(img[..., :3], "RGB").quantize(17)
    p = tmp_path / "pil_P.png"
    pal.save(p)
    assert np.array_equal(host.read_texture(p), np.asarray(Image.open(p).convert("RGBA")))


def test_pnm_reader(tmp_path, built):
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (5, 9, 3), dtype=np.uint8)
    p6 = tmp_path / "a.ppm"
    p6.write_bytes(b"P6\n# comment\n9 5\n255\n" + rgb.tobytes())
    got = host.read_texture(p6)
    assert np.array_equal(got[..., :3], rgb) and (got[..., 3] == 255).all()
    p5 = tmp_path / "a.pgm"
    p5.write_bytes(b"P5 9 5 255\n" + rgb[..., 0].tobytes())
    got = host.read_texture(p5)
    assert np.array_equal(got[..., 0], rgb[..., 0]) and np.array_equal(got[..., 1], got[..., 2])


def _good_png_bytes(tmp_path):
    rng = np.random.default_rng(11)
    p = tmp_path / "good.png"
    write_png(p, rng.integers(0, 256, (8, 8, 4)), 6, 8)
    return p.read_bytes()


@pytest.mark.parametrize("how", ["short_ihdr", "huge", "zero", "truncated_idat", "truncated_file", "bad_filter",
                                 "bad_depth", "no_idat", "palette_index"])
def test_malformed_png_is_refused(tmp_path, built, how):
    good = _good_png_bytes(tmp_path)
    ihdr = lambda w, h, d=8, c=6: _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, d, c, 0, 0, 0))
    sig = b"\x89PNG\r\n\x1a\n"
    idat_at = good.index(b"IDAT") - 4
    if how == "short_ihdr":
        bad = sig + _chunk(b"IHDR", struct.pack(">II", 8, 8)) + good[idat_at:]
    elif how == "huge":
        bad = sig + ihdr(0x7FFFFFFF, 0x7FFFFFFF) + good[idat_at:]
    elif how == "zero":
        bad = sig + ihdr(0, 8) + good[idat_at:]
    elif how == "truncated_idat":
        raw = zlib.compress(bytes(100))
        bad = sig + ihdr(8, 8) + _chunk(b"IDAT", raw) + _chunk(b"IEND", b"")
    elif how == "truncated_file":
        bad = good[:len(good) // 2]
    elif how == "bad_filter":
        raw = bytearray((8 * 4 + 1) * 8)
        raw[0] = 9
        bad = sig + ihdr(8, 8) + _chunk(b"IDAT", zlib.compress(bytes(raw))) + _chunk(b"IEND", b"")
    elif how == "bad_depth":
        bad = sig + ihdr(8, 8, 3, 6) + good[idat_at:]
    elif how == "no_idat":
        bad = sig + ihdr(8, 8) + _chunk(b"IEND", b"")
    else:  # palette index beyond the palette
        raw = bytes([0] + [200] * 8) * 8
        bad = (sig + ihdr(8, 8, 8, 3) + _chunk(b"PLTE", bytes(3 * 4)) + _chunk(b"IDAT", zlib.compress(raw)) +
               _chunk(b"IEND", b""))
    p = tmp_path / f"{how}.png"
    p.write_bytes(bad)
    with pytest.raises(RuntimeError):
        host.read_texture(p)


def write_ktx1_cube(path, faces, key_values=b""):
    """KTX 1.1, GL_RGBA / GL_UNSIGNED_BYTE, six faces, one mip level (what libktx hands the reference)."""
    size = faces.shape[1]
    kv = key_values + bytes((-len(key_values)) % 4)
    hdr = bytes([0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A])
    hdr += struct.pack("<13I", 0x04030201, 0x1401, 1, 0x1908, 0x8058, 0x1908, size, size, 0, 0, 6, 1, len(kv))
    body = struct.pack("<I", size * size * 4)
    for f in range(6):
        body += faces[f].tobytes()  # RGBA8 faces are always a multiple of 4 bytes: no cube padding
    with open(path, "wb") as fh:
        fh.write(hdr + kv + body)


def test_ktx1_cube_reader(tmp_path, built):
    rng = np.random.default_rng(5)
    faces = rng.integers(0, 256, (6, 16, 16, 4), dtype=np.uint8)
    kv = struct.pack("<I", 27) + b"KTXorientation\0S=r,T=d,R=i\0"
    p = tmp_path / "env.ktx"
    write_ktx1_cube(p, faces, kv)
    assert np.array_equal(host.read_ktx_cube(p), faces)
    p2 = tmp_path / "env_nokv.ktx"
    write_ktx1_cube(p2, faces)
    assert np.array_equal(host.read_ktx_cube(p2), faces)


@pytest.mark.parametrize("how", ["magic", "endianness", "five_faces", "not_square", "truncated", "rgb"])
def test_malformed_ktx_is_refused(tmp_path, built, how):
    faces = np.zeros((6, 8, 8, 4), "u1")
    p = tmp_path / "x.ktx"
    write_ktx1_cube(p, faces)
    b = bytearray(p.read_bytes())
    if how == "magic":
        b[1] = ord("X")
    elif how == "endianness":
        b[12:16] = struct.pack("<I", 0x01020304)
    elif how == "five_faces":
        b[12 + 40:12 + 44] = struct.pack("<I", 5)
    elif how == "not_square":
        b[12 + 28:12 + 32] = struct.pack("<I", 4)
    elif how == "truncated":
        b = b[:len(b) - 100]
    else:
        b[12 + 12:12 + 16] = struct.pack("<I", 0x1907)
    p.write_bytes(bytes(b))
    with pytest.raises(RuntimeError):
        host.read_ktx_cube(p)


def test_scene_with_ktx_environment_packs_the_file_faces(tmp_path, built):
    """Scene::setEnvironmentMap(path) -> Context::pack(): the wire cube is the file's (scene.cpp:294-309)."""
    rng = np.random.default_rng(9)
    faces = rng.integers(0, 256, (6, 32, 32, 4), dtype=np.uint8)
    p = tmp_path / "sky.ktx"
    write_ktx1_cube(p, faces)
    r = host.Renderer(device=None)
    r.load_scene("million", 64, 36, 1, scale=64)
    r.set_environment_map(p)
    ws = r.wire_scene()
    assert ws.env is not None and np.array_equal(np.stack(ws.env), faces)
    r.set_environment_map(tmp_path / "sky.png")
    with pytest.raises(RuntimeError):
        r.wire_scene()  # "cubemap format not supported" (reference scene.cpp:306-308)
    r.close()


@pytest.mark.skipif(not os.path.exists(REF_PATTERN), reason="reference assets are not mounted on this box")
def test_reference_projector_pattern_decodes(built):
    """resources/patterns/fakesense_j415.png, the eActive projector texture of BASELINE config 4."""
    Image = pytest.importorskip("PIL.Image")
    got = host.read_texture(REF_PATTERN)
    want = np.asarray(Image.open(REF_PATTERN).convert("RGBA"))
    assert got.shape == want.shape == (3000, 3000, 4)
    assert np.array_equal(got, want)


# ---- JPEG (baseline), written by PIL, decoded by the facade and by PIL ------------------------------------
def _smooth_image(rng, h, w):
    """Natural-image-like content (smooth gradients + a few edges): what JPEG is made for, so that two
    decoders with different IDCT / upsampling arithmetic stay within a few LSB of each other."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([128 + 100 * np.sin(x / 17.0 + k) * np.cos(y / 23.0 - k) for k in range(3)], -1)
    img[h // 3:h // 2, w // 4:w // 2] += 60
    img += rng.normal(0, 2, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("subsampling,mode,size,restart", [
    (0, "RGB", (64, 48), 0),      # 4:4:4
    (1, "RGB", (67, 45), 0),      # 4:2:2, sizes that are not multiples of the MCU
    (2, "RGB", (71, 53), 0),      # 4:2:0
    (2, "RGB", (160, 120), 4),    # restart markers every 4 MCU rows' worth of blocks
    (0, "L", (50, 37), 0),        # greyscale
])
def test_jpeg_reader_against_pil(tmp_path, built, subsampling, mode, size, restart):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(subsampling * 10 + len(mode))
    w, h = size
    img = _smooth_image(rng, h, w)
    pil = Image.fromarray(img if mode == "RGB" else img[..., 0], mode)
    path = tmp_path / f"t_{mode}_{subsampling}_{restart}.jpg"
    kw = {"quality": 92}
    if mode == "RGB":
        kw["subsampling"] = subsampling
    if restart:
        kw["restart_marker_blocks"] = restart
    try:
        pil.save(path, "JPEG", **kw)
    except TypeError:
        pytest.skip("this PIL cannot write restart markers")
    got = host.read_texture(path)
    want = np.asarray(Image.open(path).convert("RGBA"))
    assert got.shape == want.shape == (h, w, 4) and (got[..., 3] == 255).all()
    diff = np.abs(got[..., :3].astype(int) - want[..., :3].astype(int))
    # different IDCT rounding (float vs libjpeg's integer islow) and chroma upsampling filters: a few LSB
    assert diff.mean() < 1.0 and diff.max() <= (4 if subsampling == 0 else 24), (diff.mean(), diff.max())
    # and both are close to what was encoded
    src = img if mode == "RGB" else np.repeat(img[..., :1], 3, -1)
    assert np.abs(got[..., :3].astype(int) - src.astype(int)).mean() < 4.0


def test_progressive_jpeg_is_refused(tmp_path, built):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(1)
    p = tmp_path / "prog.jpg"
    Image.fromarray(_smooth_image(rng, 32, 32), "RGB").save(p, "JPEG", progressive=True)
    with pytest.raises(RuntimeError):
        host.read_texture(p)


@pytest.mark.parametrize("cut", [2, 20, 200, 600, 0.5, 0.9])
def test_truncated_jpeg_does_not_crash(tmp_path, built, cut):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(2)
    p = tmp_path / "full.jpg"
    Image.fromarray(_smooth_image(rng, 64, 64), "RGB").save(p, "JPEG", quality=90)
    data = p.read_bytes()
    k = int(len(data) * cut) if isinstance(cut, float) else cut
    q = tmp_path / "cut.jpg"
    q.write_bytes(data[:k])
    try:  # headers cut: refused; entropy data cut: the missing bits read as zero (stb does the same), no crash
        out = host.read_texture(q)
        assert out.shape == (64, 64, 4)
    except RuntimeError:
        pass


def test_corrupted_jpeg_bytes_never_crash(tmp_path, built):
    """Random byte damage anywhere in a valid file: decoded to something or refused, never a crash (asset
    files are untrusted input)."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(4)
    p = tmp_path / "ok.jpg"
    Image.fromarray(_smooth_image(rng, 48, 64), "RGB").save(p, "JPEG", quality=85, subsampling=2)
    data = p.read_bytes()
    q = tmp_path / "damaged.jpg"
    for _ in range(150):
        b = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(2, len(b)))] = int(rng.integers(0, 256))
        q.write_bytes(bytes(b))
        try:
            host.read_texture(q)
        except RuntimeError:
            pass
