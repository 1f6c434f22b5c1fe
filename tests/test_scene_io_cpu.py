"""Scene input through the C ABI (flx_scene_load / flx_envmap_load) against the reference's own loader code
(oracle/_ref/scene_tool = its vendored tinyobjloader + the scene.cpp conversion, its PLY reader, its RGBE reader and
importance tables): byte-identical triangles, materials, texture names and tables -- on small committed inputs that poke
at the loader's corners (always), and on the reference's assets (where /root/reference and the blobs exist).
Host code only: runs without a GPU."""
import os
import struct

import numpy as np
import pytest

from fluctus_b200 import EnvMapData, FluctusError
from fluctus_b200.scene_io import envmap_from_rgb, load_envmap, load_model
from fluctus_b200.structs import MATERIAL_DTYPE, TRIANGLE_DTYPE

from conftest import scene_blob

IO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io")
EXPECTED = np.load(os.path.join(IO, "expected.npz"))
REF_ASSETS = os.path.join(os.environ.get("FLX_REFERENCE_DIR", "/root/reference"), "assets")


def same_tables(env, prob, alias, pdf, what):
    assert np.array_equal(env.prob.view(np.uint32), prob.view(np.uint32)), what + ": prob table"
    assert np.array_equal(env.pdf.view(np.uint32), pdf.view(np.uint32)), what + ": pdf table"
    live = prob < 1.0  # the reference never writes alias[i] where prob[i] = 1 (src/envmap.cpp:101-113): uninitialised memory there
    assert np.array_equal(env.alias[live], alias[live]), what + ": alias table"


@pytest.mark.parametrize("name", ["tricky.obj", "tricky.ply", "plain.ply"])
def test_models_match_the_reference_loader_on_committed_inputs(name):
    """tricky.obj: CRLF line ends, a missing first mtllib file, quads and a pentagon (fan triangulation), negative indices,
    v / v/t / v//n / v/t/n corners (face normals when a normal is missing), numbers with exponents / leading '+' / '.5'
    (which the loader reads as 0) / more digits than a float holds, an unknown usemtl (-> default material), a material
    defined twice (first wins), texture statements with options and backslashes, the `shader` key with trailing blanks and
    with a tab.  tricky.ply: normals and a quad.  plain.ply: no normals (face normals)."""
    key = name.replace(".", "_")
    m = load_model(os.path.join(IO, name))
    assert m.tris.tobytes() == EXPECTED[key + "_tris"].tobytes()
    assert m.materials.tobytes() == EXPECTED[key + "_materials"].tobytes()
    assert m.texture_names == [str(s) for s in EXPECTED[key + "_textures"]]


def test_rle_hdr_and_importance_tables_match_the_reference_reader():
    env = load_envmap(os.path.join(IO, "small_rle.hdr"))
    w, h = EXPECTED["small_rle_hdr_size"]
    assert (env.width, env.height) == (w, h)
    assert np.array_equal(env.rgb.reshape(-1).view(np.uint32), EXPECTED["small_rle_hdr_rgb"].view(np.uint32))
    same_tables(env, EXPECTED["small_rle_hdr_prob"], EXPECTED["small_rle_hdr_alias"], EXPECTED["small_rle_hdr_pdf"], "small_rle.hdr")
    # the Python mirror used by the tests for synthetic maps and the C++ tables agree too
    py = EnvMapData.from_rgb(env.rgb)
    same_tables(envmap_from_rgb(env.rgb), py.prob, py.alias, py.pdf, "from_rgb")
    black = envmap_from_rgb(np.zeros((4, 8, 3), np.float32))  # I == 0: uniform pdf (src/envmap.cpp:62-63)
    assert np.all(black.pdf == np.float32(1.0) / np.float32(32))


def test_loader_errors_are_reported_not_fatal():
    """the reference prints and calls waitExit() (src/scene.cpp:210-214, src/envmap.cpp:13-17); the library returns an error"""
    with pytest.raises(FluctusError):
        load_model(os.path.join(IO, "does_not_exist.obj"))
    with pytest.raises(FluctusError):
        load_model(os.path.join(IO, "small_rle.hdr"))  # unsupported ending
    with pytest.raises(FluctusError):
        load_envmap(os.path.join(IO, "tricky.obj"))  # not a Radiance file


@pytest.mark.parametrize("name,rel", [("teapot", "teapot.ply"), ("conference", "conference/conference.obj"), ("luxball", "luxball/luxball.obj"),
                                      ("country_kitchen", "country_kitchen/Country-Kitchen.obj")])
def test_reference_assets_match_the_reference_loader(name, rel):
    path = os.path.join(REF_ASSETS, rel)
    if not os.path.exists(path):
        pytest.skip("reference assets not present")
    buf = open(scene_blob(name), "rb").read()
    magic, nt, ni, nn, nm, ntex = struct.unpack_from("<6I", buf, 0)
    off = 24
    tris = buf[off:off + nt * 160]; off += nt * 160 + ni * 4 + nn * 48
    mats = buf[off:off + nm * 80]; off += nm * 80
    names = []
    for _ in range(ntex):
        (ln,) = struct.unpack_from("<I", buf, off); off += 4
        names.append(buf[off:off + ln].decode()); off += ln
    m = load_model(path)
    assert len(m.tris) == nt and m.tris.tobytes() == tris
    assert m.materials.tobytes() == mats
    assert m.texture_names == names


def test_reference_env_map_matches_the_reference_reader():
    path = os.path.join(REF_ASSETS, "env_maps", "night.hdr")
    blob = os.path.join(os.path.dirname(scene_blob("teapot")), "night.env.bin")
    if not os.path.exists(path) or not os.path.exists(blob):
        pytest.skip("reference assets not present")
    ref = EnvMapData.load_blob(blob)
    env = load_envmap(path)
    assert np.array_equal(env.rgb.view(np.uint32), ref.rgb.view(np.uint32))
    same_tables(env, ref.prob, ref.alias, ref.pdf, "night.hdr")


# ---------------------------------------------------------------------------------------------- image output
def test_png_and_hdr_writers_follow_saveImage(tmp_path):
    """CLContext::saveImage (src/clcontext.cpp:386-465): '*.hdr' = accumulator / sample count as Radiance RGBE (checked by
    reading it back with the RGBE reader, which is pinned to the reference's), otherwise byte = (uchar)(255 * clamp01(c)) of
    the preview as PNG (checked with an independent decoder, Pillow).  Row 0 of the buffer is the bottom row of the picture
    (DevIL origin lower-left, src/main.cpp:69-71).  The reference writes through DevIL, which is absent: file BYTES are
    unpinned, pixel values are what is compared."""
    from PIL import Image
    from fluctus_b200.scene_io import write_image
    rng = np.random.default_rng(21)
    w, h = 37, 19
    preview = rng.uniform(-0.2, 1.3, (h, w, 4)).astype(np.float32)
    preview[0, 0, :3] = (0.0, 1.0, 0.5)
    png = tmp_path / "out.png"
    write_image(png, preview, w, h)
    got = np.asarray(Image.open(png).convert("RGB"))
    want = (np.float32(255) * np.clip(preview[..., :3], 0.0, 1.0)).astype(np.uint8)[::-1]  # C cast truncates; top row first in the file
    assert got.shape == (h, w, 3) and np.array_equal(got, want)
    # a 300 x 300 image needs more than one stored deflate block (65535 bytes each)
    big = rng.uniform(0, 1, (300, 300, 4)).astype(np.float32)
    write_image(tmp_path / "big.png", big, 300, 300)
    assert np.array_equal(np.asarray(Image.open(tmp_path / "big.png").convert("RGB")), (np.float32(255) * big[..., :3]).astype(np.uint8)[::-1])

    acc = rng.uniform(0, 40, (h, w, 4)).astype(np.float32)
    acc[..., 3] = rng.integers(1, 9, (h, w))
    acc[3, 4, :3] = 0.0
    hdr = tmp_path / "out.hdr"
    write_image(hdr, acc, w, h)
    back = load_envmap(hdr).rgb  # (h, w, 3), top row first
    lin = (acc[..., :3] / acc[..., 3:4])[::-1]
    v = lin.max(axis=-1)
    m, e = np.frexp(v)
    scale = np.where(v >= 1e-32, (m.astype(np.float64) * 256.0 / np.maximum(v, 1e-38)).astype(np.float32), 0).astype(np.float32)
    q = np.clip((lin * scale[..., None]).astype(np.int32), 0, 255).astype(np.float32)
    want = q * np.ldexp(np.float32(1.0), e - 8)[..., None].astype(np.float32)
    assert np.array_equal(back, np.where(v[..., None] >= 1e-32, want, 0).astype(np.float32))
    assert np.all(np.abs(back - lin) <= np.maximum(v[..., None] / 128.0, 1e-30))  # RGBE: 8-bit mantissa shared by the pixel


# ---------------------------------------------------------------------------------------------- hierarchy cache file
def test_hierarchy_cache_written_by_the_reference_is_read_correctly():
    """tests/golden/io/teapot_hierarchy.bin was written by the reference's BVH::exportTo (src/bvh.cpp:174-192) for its SBVH
    of teapot.ply; the same build is stored in the teapot golden fixture.  The file's node-count field holds the index
    count (the reference's bug, src/bvh.cpp:185): 3290 instead of 2065 -- the importer must not believe it."""
    from fluctus_b200.scene_io import import_hierarchy
    z = np.load(os.path.join(os.path.dirname(IO), "teapot_c1.npz"))
    nodes, indices = import_hierarchy(os.path.join(IO, "teapot_hierarchy.bin"))
    raw = open(os.path.join(IO, "teapot_hierarchy.bin"), "rb").read()
    (ni,) = struct.unpack_from("<I", raw, 0)
    (field,) = struct.unpack_from("<I", raw, 4 + 4 * ni)
    assert field == ni and field != len(nodes)
    assert np.array_equal(indices, z["scene_indices"])
    assert nodes.tobytes() == z["scene_nodes"].tobytes()


def test_hierarchy_cache_round_trip_and_reference_importer(tmp_path):
    from fluctus_b200.scene_io import export_hierarchy, import_hierarchy
    from fluctus_b200.scene import make_room_scene
    room = make_room_scene(materials="mixed", n_blobs=8)
    path = tmp_path / "room_hierarchy.bin"
    export_hierarchy(path, room.nodes, room.indices)
    nodes, indices = import_hierarchy(path)
    assert np.array_equal(indices, room.indices)
    keep = ["bmin", "bmax", "parent", "link", "nPrims"]
    assert all(np.array_equal(nodes[k], room.nodes[k]) for k in keep)
    with pytest.raises(FluctusError):
        import_hierarchy(os.path.join(IO, "tricky.obj"))
    # the reference's own importer (BVH::importFrom, src/bvh.cpp:102-152) reads what the exporter wrote
    tool = os.path.join(os.path.dirname(os.path.dirname(IO)), "..", "oracle", "_ref", "scene_tool")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/scene_tool not built (needs /root/reference)")
    import subprocess
    out = tmp_path / "imported.bin"
    subprocess.run([tool, "cache-import", str(path), str(out)], check=True, capture_output=True)
    buf = open(out, "rb").read()
    ni, nn = struct.unpack_from("<2I", buf, 0)
    assert (ni, nn) == (len(room.indices), len(room.nodes))
    from fluctus_b200.structs import NODE_DTYPE
    assert np.array_equal(np.frombuffer(buf, np.uint32, ni, 8), room.indices)
    back = np.frombuffer(buf, NODE_DTYPE, nn, 8 + 4 * ni)
    assert all(np.array_equal(back[k], room.nodes[k]) for k in keep)


def test_number_reader_matches_the_reference_loader_on_random_literals(tmp_path):
    """4000 random decimal literals (signs, fractions of 0-19 digits, exponents, leading '+', '.5'-style and other forms the
    loader rejects) go through both loaders as normals / texture coordinates of sane triangles: every float must come out
    with the same bits.  (The loader's reader is not strtod: e.g. it scales fraction digits one by one.)"""
    tool = os.path.join(os.path.dirname(os.path.dirname(IO)), "..", "oracle", "_ref", "scene_tool")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/scene_tool not built (needs /root/reference)")
    import subprocess
    rng = np.random.default_rng(2024)

    def literal():
        kind = rng.integers(0, 10)
        sign = ["", "", "-", "+"][rng.integers(0, 4)]
        ip = str(rng.integers(0, 10 ** int(rng.integers(1, 10))))
        fp = "".join(str(d) for d in rng.integers(0, 10, int(rng.integers(0, 20))))
        ex = "%s%s%d" % ("eE"[rng.integers(0, 2)], ["", "-", "+"][rng.integers(0, 3)], rng.integers(0, 40))
        if kind == 0:
            return sign + ip
        if kind == 1:
            return sign + ip + "." + fp
        if kind == 2:
            return sign + ip + "." + fp + ex
        if kind == 3:
            return sign + ip + ex
        if kind == 4:
            return sign + "." + fp + "5"      # no integer part: the loader yields the default 0
        if kind == 5:
            return sign + ip + "." + fp + "e"  # empty exponent: default 0
        if kind == 6:
            return "%.9g" % rng.normal(0, 10.0 ** rng.integers(-8, 8))
        if kind == 7:
            return repr(float(np.float32(rng.uniform(-1, 1))))
        if kind == 8:
            return sign + "0." + "0" * int(rng.integers(0, 12)) + ip
        return sign + ip + "." + fp + "x"     # trailing junk is ignored
    n = 1000
    path = tmp_path / "numbers.obj"
    with open(path, "w") as f:
        for i in range(n):
            f.write("v %d 0 0\nv %d 1 0\nv %d 0 1\n" % (i, i, i))
        for i in range(n):
            f.write("vn %s %s %s\nvt %s\n" % (literal(), literal(), literal(), literal()))
        for i in range(n):
            f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % ((3 * i + 1, i + 1, i + 1, 3 * i + 2, i + 1, i + 1, 3 * i + 3, i + 1, i + 1)))
    blob = tmp_path / "numbers.bin"
    subprocess.run([tool, "obj", str(path), str(blob)], check=True, capture_output=True)
    buf = open(blob, "rb").read()
    nt = struct.unpack_from("<6I", buf, 0)[1]
    want = np.frombuffer(buf, TRIANGLE_DTYPE, nt, 24)
    got = load_model(path).tris
    assert nt == n and got.tobytes() == want.tobytes()


def test_differential_fuzz_against_the_reference_loader(tmp_path):
    """Mutants of tricky.obj / tricky.mtl (numbers rewritten in other notations, whitespace and line-ending changes, statements
    duplicated / dropped / shuffled where that keeps every index valid, comments, unknown statements) go through both loaders;
    whenever the reference's loader accepts a mutant, flx_scene_load must produce the same bytes."""
    tool = os.path.join(os.path.dirname(os.path.dirname(IO)), "..", "oracle", "_ref", "scene_tool")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/scene_tool not built (needs /root/reference)")
    import re
    import subprocess
    rng = np.random.default_rng(31337)
    obj_lines = open(os.path.join(IO, "tricky.obj"), newline="").read().split("\r\n")
    mtl_lines = open(os.path.join(IO, "tricky.mtl")).read().split("\n")
    number = re.compile(r"(?<![\w./-])[-+]?\d+\.\d+(?:[eE][-+]?\d+)?")

    def renumber(line):
        def alt(m):
            v = float(m.group(0))
            return [m.group(0), "%.10f" % v, "%e" % v, "%.3E" % v, repr(v), "+%s" % m.group(0).lstrip("+") if v >= 0 else m.group(0)][rng.integers(0, 6)]
        return number.sub(alt, line)

    def mutate(lines, protect):
        out = []
        for ln in lines:
            r = rng.random()
            if ln.startswith(protect) or r > 0.45:
                out.append(ln)
            elif r < 0.12:
                out.append(renumber(ln))
            elif r < 0.2:
                out.append(ln.replace(" ", "  ").replace("  ", " \t", 1))
            elif r < 0.26:
                out += ["# a comment", ln, ""]
            elif r < 0.32:
                out += [ln, "zz unknown statement 1 2 3"]
            elif r < 0.38:
                out.append("  " + ln + "   ")
            else:
                out.append(renumber(ln) + " ")
        return out
    agreed = 0
    for k in range(40):
        d = tmp_path / ("m%d" % k)
        d.mkdir()
        eol = ["\n", "\r\n"][rng.integers(0, 2)]
        obj = mutate(obj_lines, ("mtllib",))
        mtl = mutate(mtl_lines, ("newmtl",))
        if rng.random() < 0.3:  # duplicate a face block at the end: indices stay valid
            obj += [ln for ln in obj_lines if ln.startswith(("usemtl", "f "))][:6]
        (d / "tricky.obj").write_text(eol.join(obj) + eol, newline="")
        (d / "tricky.mtl").write_text("\n".join(mtl) + "\n")
        blob = d / "ref.bin"
        r = subprocess.run([tool, "obj", str(d / "tricky.obj"), str(blob)], capture_output=True, timeout=60)
        if r.returncode != 0 or not blob.exists():
            continue  # the reference rejected (or crashed on) this mutant: nothing to compare
        buf = open(blob, "rb").read()
        magic, nt, ni, nn, nm, ntex = struct.unpack_from("<6I", buf, 0)
        off = 24
        tris = buf[off:off + nt * 160]; off += nt * 160 + ni * 4 + nn * 48
        mats = buf[off:off + nm * 80]
        m = load_model(d / "tricky.obj")
        assert m.tris.tobytes() == tris, "mutant %d: triangles differ" % k
        assert m.materials.tobytes() == mats, "mutant %d: materials differ" % k
        agreed += 1
    assert agreed >= 30, agreed


def test_malformed_files_are_rejected_not_crashed_on(tmp_path):
    """truncated / corrupted inputs of every reader: an error comes back (the reference's readers print and exit, or read past
    the end); nothing crashes and nothing half-read is returned"""
    from fluctus_b200.scene_io import import_hierarchy
    hdr = open(os.path.join(IO, "small_rle.hdr"), "rb").read()
    for name, data in (("cut_header.hdr", hdr[:20]), ("cut_pixels.hdr", hdr[:200]), ("bad_width.hdr", hdr.replace(b"+X 16", b"+X 17")),
                       ("no_size.hdr", hdr.replace(b"-Y 8 +X 16", b"nonsense")), ("zero_run.hdr", hdr[:hdr.index(b"+X 16\n") + 10] + bytes([128, 0]) * 40)):
        p = tmp_path / name
        p.write_bytes(data)
        with pytest.raises(FluctusError):
            load_envmap(p)
    cache = open(os.path.join(IO, "teapot_hierarchy.bin"), "rb").read()
    for name, data in (("empty.bin", b""), ("cut_indices.bin", cache[:1000]), ("no_nodes.bin", cache[:4 + 4 * struct.unpack_from("<I", cache, 0)[0] + 4])):
        p = tmp_path / name
        p.write_bytes(data)
        with pytest.raises(FluctusError):
            import_hierarchy(p)
    for name, text in (("index_out_of_range.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n"), ("negative_out_of_range.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf -1 -2 -9\n"),
                       ("no_faces.obj", "v 0 0 0\nv 1 0 0\n"), ("normal_out_of_range.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//5\n")):
        p = tmp_path / name
        p.write_text(text)
        with pytest.raises(FluctusError):
            load_model(p)
    p = tmp_path / "bad_face.ply"
    p.write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 1 9\n")
    with pytest.raises(FluctusError):
        load_model(p)


def test_sizes_claimed_by_a_file_are_checked_against_the_file(tmp_path):
    """Sizes read from a file are claims: a header that announces more than the file can hold is refused before anything is
    allocated from it, short bodies are errors (not silently repeated lines), a binary PLY is not parsed as text, and no C++
    exception crosses the C ABI (this process is still alive afterwards)."""
    import zlib
    from fluctus_b200.scene_io import load_image
    hdr = open(os.path.join(IO, "small_rle.hdr"), "rb").read()
    huge = tmp_path / "huge.hdr"
    huge.write_bytes(hdr.replace(b"-Y 8 +X 16", b"-Y 2000000000 +X 2000000000"))
    with pytest.raises(FluctusError, match="larger than the file"):
        load_envmap(huge)
    big_flat = tmp_path / "big_flat.hdr"  # 4-pixel-wide images are stored flat: 40000 x 4 pixels claimed, 128 bytes present
    big_flat.write_bytes(hdr.replace(b"-Y 8 +X 16", b"-Y 40000 +X 4"))
    with pytest.raises(FluctusError):
        load_envmap(big_flat)
    head = "ply\nformat %s 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n%selement face %d\nproperty list uchar int vertex_indices\nend_header\n"
    body = "0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n"
    cases = {"binary.ply": head % ("binary_little_endian", 3, "", 1) + body,
             "bomb.ply": head % ("ascii", 2000000000, "", 1) + body,
             "negative.ply": head % ("ascii", -3, "", 1) + body,
             "short_body.ply": head % ("ascii", 3, "", 1) + "0 0 0\n1 0 0\n",
             "no_end.ply": (head % ("ascii", 3, "", 1)).replace("end_header\n", ""),
             # two vertex elements, only the second with normals: the normal array is shorter than the position array
             "mixed_normals.ply": "ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\nelement vertex 1\nproperty float x\n"
                                  "property float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\nelement face 1\n"
                                  "property list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0 0 0 1\n3 0 1 2\n"}
    for name, text in cases.items():
        p = tmp_path / name
        p.write_text(text)
        with pytest.raises(FluctusError):
            load_model(p)
    ok = tmp_path / "ok.ply"  # the same header with honest numbers loads
    ok.write_text(head % ("ascii", 3, "", 1) + body)
    assert len(load_model(ok).tris) == 1
    # a PNG whose compressed stream inflates to far more than its header announces (a "zip bomb"): refused, not inflated
    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xffffffff)
    bomb = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(64 << 20), 9)) + chunk(b"IEND", b"")
    p = tmp_path / "bomb.png"
    p.write_bytes(bomb)
    with pytest.raises(FluctusError):
        load_image(p)
    with pytest.raises(FluctusError):
        envmap_from_rgb(np.zeros((1, 1, 3), np.float32)[:0])


# ---------------------------------------------------------------------------------------------- textures
def test_png_decoder_matches_an_independent_decoder(tmp_path):
    """flx_image_load against Pillow on PNGs of every colour type and bit depth Pillow can write (grey 1/8/16 bit, grey+alpha,
    RGB, RGBA, palette with and without transparency), sizes that are not multiples of anything, compression levels that make
    the encoder use stored, fixed and dynamic Huffman blocks, and -- where present -- the reference's own PNG textures.
    PNG is lossless: the RGBA8 bytes must be identical (rows bottom-up, DevIL's lower-left origin in the reference)."""
    from PIL import Image
    from fluctus_b200.scene_io import load_image
    rng = np.random.default_rng(9)

    def check(path):
        want = np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8)[::-1]
        got = load_image(path)
        assert got.shape == want.shape and np.array_equal(got, want), path

    smooth = (np.add.outer(np.arange(37), np.arange(53)) * 3 % 256).astype(np.uint8)
    noise = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    k = 0
    for base in (smooth, noise):
        for mode in ("L", "LA", "RGB", "RGBA", "P", "1"):
            for level in (0, 1, 9):
                if mode == "L":
                    im = Image.fromarray(base, "L")
                elif mode == "LA":
                    im = Image.fromarray(np.stack([base, base[::-1]], axis=-1), "LA")
                elif mode == "RGB":
                    im = Image.fromarray(np.stack([base, base.T[:37, :53] if base.T.shape == base.shape else base[::-1], 255 - base], axis=-1), "RGB")
                elif mode == "RGBA":
                    im = Image.fromarray(np.stack([base, base[::-1], 255 - base, base[:, ::-1]], axis=-1), "RGBA")
                elif mode == "P":
                    im = Image.fromarray(base, "L").quantize(17)
                else:
                    im = Image.fromarray(base, "L").convert("1")
                p = tmp_path / ("t%d.png" % k)
                k += 1
                im.save(p, compress_level=level)
                check(p)
    pal = Image.fromarray(noise % 5, "P")
    pal.putpalette([10, 20, 30, 200, 100, 0, 0, 0, 255, 255, 255, 255, 9, 99, 199])
    pal.save(tmp_path / "trns.png", transparency=bytes([0, 128, 255, 7, 200]))
    check(tmp_path / "trns.png")
    Image.fromarray(noise.astype(np.uint16) * 257).save(tmp_path / "g16.png")  # uint16 -> mode I;16
    got = load_image(tmp_path / "g16.png")  # 16-bit grey: the high byte of every sample
    assert np.array_equal(got[..., 0], noise[::-1]) and np.array_equal(got[..., 3], np.full_like(noise, 255))
    for rel in ("egyptcat/EgyptCat.png", "country_kitchen/textures/Kitchen-carrot-uv.png", "country_kitchen/textures/Kitchen-mushroom-texture.png"):
        if os.path.exists(os.path.join(REF_ASSETS, rel)):
            check(os.path.join(REF_ASSETS, rel))
    with pytest.raises(FluctusError):
        load_image(os.path.join(IO, "tricky.obj"))
    bad = tmp_path / "cut.png"
    bad.write_bytes(open(tmp_path / "t0.png", "rb").read()[:60])
    with pytest.raises(FluctusError):
        load_image(bad)


def _pillow_rgba(path):
    from PIL import Image
    im = Image.open(path)
    assert im.mode in ("RGB", "L"), im.mode
    return np.ascontiguousarray(np.asarray(im.convert("RGBA"))[::-1])


def test_jpeg_decoder_matches_libjpeg_bit_for_bit(tmp_path):
    """flx_image_load on JPEG against Pillow (libjpeg-turbo, the IJG decoder family DevIL uses too): IDENTICAL bytes on files that
    cover every branch of flx_jpeg.cpp -- baseline and progressive, 4:4:4 / 4:2:2 / 4:2:0, grey, qualities 5..100 (the 16-bit-table
    and clamping corners), optimised Huffman tables, restart intervals, sizes that are not multiples of the MCU, 1x1."""
    from PIL import Image
    from fluctus_b200.scene_io import load_image
    rng = np.random.default_rng(1)

    def picture(h, w, c):
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([(np.sin(xx / 7.0 + k) + np.cos(yy / 5.0 * (k + 1))) * 60 + 128 + rng.normal(0, 12, (h, w)) for k in range(c)], -1)
        return np.clip(img, 0, 255).astype(np.uint8)

    cases = 0
    for (w, h) in ((1, 1), (7, 5), (16, 16), (17, 33), (250, 123)):
        for mode in ("RGB", "L"):
            img = picture(h, w, 3 if mode == "RGB" else 1)
            im = Image.fromarray(img if mode == "RGB" else img[..., 0], mode)
            for sub in ((0, 1, 2) if mode == "RGB" else (0,)):
                for prog in (False, True):
                    for q, extra in ((5, {}), (50, {}), (50, {"optimize": True}), (50, {"restart_marker_blocks": 3}), (50, {"restart_marker_rows": 1}), (92, {}), (100, {})):
                        path = tmp_path / "t.jpg"
                        kw = dict(quality=q, progressive=prog, **extra)
                        if mode == "RGB":
                            kw["subsampling"] = sub
                        try:
                            im.save(path, "JPEG", **kw)
                        except (TypeError, OSError, ValueError):
                            continue  # an encoder option this Pillow does not have
                        got, want = load_image(path), _pillow_rgba(path)
                        assert np.array_equal(got, want), "%dx%d %s subsampling %d progressive %s quality %d %r: %d pixels differ" % (
                            w, h, mode, sub, prog, q, extra, int((got != want).any(axis=2).sum()))
                        cases += 1
    assert cases > 200


def _kitchen_dir():
    for root in (os.path.join(os.path.dirname(REF_ASSETS), "assets"), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "assets")):
        p = os.path.join(root, "country_kitchen")
        if os.path.isdir(p):
            return p
    pytest.skip("the Country-Kitchen files are neither under /root/reference/assets nor under oracle/_ref/assets (oracle/make_scenes.py)")


def test_reference_jpeg_textures_decode_like_libjpeg():
    """the reference's own JPEGs (11 in Country Kitchen: 8 baseline with restart intervals, 2 progressive, 1 with 4:2:0 chroma,
    most with an Adobe marker and no JFIF marker) -- identical to Pillow"""
    from fluctus_b200.scene_io import load_image
    import glob
    files = sorted(glob.glob(os.path.join(_kitchen_dir(), "textures", "*.jpg")))
    assert len(files) == 11
    for f in files:
        assert np.array_equal(load_image(f), _pillow_rgba(f)), os.path.basename(f)


def test_country_kitchen_from_files_only_equals_the_blob_the_parity_tests_use():
    """C3 through the library alone -- flx_scene_load (OBJ + MTL), flx_image_load (its JPEG and PNG textures), flx_pack_textures --
    gives exactly the triangles, materials, texture descriptors and texture bytes of the scene blob (made by the reference's loader
    code + Pillow) that the GPU parity tests render.  So what those tests pin holds for the files-only path."""
    from fluctus_b200 import SceneData
    from fluctus_b200.scene_io import load_model_with_textures
    blob = SceneData.load_blob(scene_blob("country_kitchen"))
    model, desc, data = load_model_with_textures(os.path.join(_kitchen_dir(), "Country-Kitchen.obj"))
    assert model.texture_names == blob.texture_names and len(model.texture_names) == 11  # 9 JPEG + 2 PNG files, used by 17 map_Kd / map_Bump entries
    assert model.tris.tobytes() == blob.tris.tobytes() and model.materials.tobytes() == blob.materials.tobytes()
    assert desc.tobytes() == blob.tex_desc.tobytes()
    assert np.array_equal(data, blob.tex_data)


def test_malformed_jpeg_files_are_rejected_or_decoded_never_crashed_on(tmp_path):
    """truncations and byte flips of a baseline and a progressive file: every outcome is an error or a picture of the announced
    size (the IJG decoder also carries on over corrupt entropy data); arithmetic-coded / 12-bit / CMYK files are refused by name"""
    from PIL import Image
    from fluctus_b200.scene_io import load_image
    rng = np.random.default_rng(7)
    img = (rng.uniform(0, 255, (40, 56, 3))).astype(np.uint8)
    outcomes = {"error": 0, "picture": 0}
    for prog in (False, True):
        src = tmp_path / "src.jpg"
        Image.fromarray(img, "RGB").save(src, "JPEG", quality=80, progressive=prog, subsampling=2)
        raw = src.read_bytes()
        mutants = [raw[:k] for k in (0, 1, 2, 3, 10, 100, len(raw) // 2, len(raw) - 2)]
        for _ in range(150):
            b = bytearray(raw)
            for _k in range(int(rng.integers(1, 4))):
                b[int(rng.integers(2, len(b)))] = int(rng.integers(0, 256))
            mutants.append(bytes(b))
        for m in mutants:
            p = tmp_path / "m.jpg"
            p.write_bytes(m)
            try:
                out = load_image(p)
                assert out.ndim == 3 and out.shape[2] == 4 and out.size > 0
                outcomes["picture"] += 1
            except FluctusError:
                outcomes["error"] += 1
    assert outcomes["error"] > 10 and outcomes["picture"] > 10, outcomes
    cmyk = tmp_path / "cmyk.jpg"
    Image.fromarray(np.zeros((8, 8, 4), np.uint8), "CMYK").save(cmyk, "JPEG")
    with pytest.raises(FluctusError, match="three-component"):
        load_image(cmyk)
    soft = bytearray((tmp_path / "src.jpg").read_bytes())
    i = soft.index(b"\xff\xc2")
    soft[i + 1] = 0xCA  # progressive, arithmetic coding
    (tmp_path / "arith.jpg").write_bytes(bytes(soft))
    with pytest.raises(FluctusError, match="arithmetic"):
        load_image(tmp_path / "arith.jpg")
    with pytest.raises(FluctusError, match="unsupported image format"):
        load_image(tmp_path / "picture.bmp")


def test_texture_packing_matches_packTextures():
    """flx_pack_textures = CLContext::packTextures (src/clcontext.cpp:570-611); same descriptors and blob as the Python helper the
    scene blobs were made with, and -- where the reference's assets are present -- the egyptcat texture decoded + packed in C
    equals the blob the tests render with."""
    from fluctus_b200.scene import pack_textures as pack_py
    from fluctus_b200.scene_io import load_image, pack_textures
    rng = np.random.default_rng(4)
    images = [rng.integers(0, 256, (h, w, 4), dtype=np.uint8) for h, w in ((3, 5), (16, 16), (1, 7))]
    desc, blob = pack_textures(images)
    assert [tuple(d) for d in desc] == [(0, 5, 3), (60, 16, 16), (60 + 1024, 7, 1)]
    assert np.array_equal(blob, np.concatenate([im.reshape(-1) for im in images]))
    assert pack_textures([])[1].size == 0
    png = os.path.join(REF_ASSETS, "egyptcat", "EgyptCat.png")
    side = os.path.join(os.path.dirname(scene_blob("teapot")), "egyptcat.tex.npz")
    if os.path.exists(png) and os.path.exists(side):
        z = np.load(side)
        d2, b2 = pack_textures([load_image(png)])
        assert np.array_equal(d2.view(np.uint32).reshape(-1, 3), z["desc"]) and np.array_equal(b2, z["data"])
        d3, b3 = pack_py([png])
        assert np.array_equal(b3, b2)
