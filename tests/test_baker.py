"""The scene baker (rf_bake_gltf = PtFormat(gltfPath)) on the reference's own assets: the anchors the
survey measured with the reference's code, byte stability, the image decoders against an independent decoder, the places
where the baker must follow the reference rather than the obvious thing (4-vector normal normalisation, std::sort mesh
order), and the reference's error messages.  CPU only.  Needs /root/reference/assets (skipped where it is not mounted;
the committed Duck fixture still pins the Duck bytes)."""
import io
from pathlib import Path

import numpy as np
import pytest

import _oracle as O
import rayfinder_b200 as rf
from rayfinder_b200 import baker, capi

REF_ASSETS = Path("/root/reference/assets")
needs_assets = pytest.mark.skipif(not (REF_ASSETS / "Duck.glb").exists(), reason="/root/reference/assets not mounted")


def channels(bgra):
    return ((bgra[..., None] >> np.array([16, 8, 0])) & 0xFF).astype(np.int32)  # r, g, b


@needs_assets
def test_duck_bake_matches_the_committed_fixture_byte_for_byte():
    pt = baker.bake(REF_ASSETS / "Duck.glb")
    assert pt.dumps() == O.duck_pt_bytes()
    # Duck = 4 212 triangles / 8 383 nodes / 1 texture (512x512 palette PNG)
    assert pt.bvh_nodes.size == 8383 and pt.triangle_position_attributes.shape[0] == 4212 and len(pt.base_color_textures) == 1
    assert pt.base_color_textures[0].shape == (512, 512)
    assert baker.bake(REF_ASSETS / "Duck.glb").dumps() == pt.dumps()  # byte-stable across runs


@needs_assets
@pytest.mark.timeout(600)
def test_sponza_bake_anchors_order_and_stability(tmp_path):
    pt = baker.bake(REF_ASSETS / "Sponza.glb")
    # 262 267 triangles, 501 673 nodes, 25 textures (24 x 1024^2 + one 4x4)
    assert pt.triangle_position_attributes.shape[0] == 262267 and pt.bvh_nodes.size == 501673 and len(pt.base_color_textures) == 25
    shapes = sorted(t.shape for t in pt.base_color_textures)
    assert shapes == [(4, 4)] + [(1024, 1024)] * 24
    # meshes come out sorted by base-colour texture index (gltf_model.cpp:462) ...
    order = pt.model_base_color_texture_indices
    assert np.all(np.diff(order.astype(np.int64)) >= 0) and order.max() == 24 and order.size == pt.model_vertex_indices.shape[0]
    # ... slices tile the vertex / index arrays in mesh order (pt_format.cpp:100-148)
    for slices, total in ((pt.model_vertex_positions, pt.vertex_positions.shape[0]), (pt.model_vertex_indices, pt.vertex_indices.size)):
        assert slices[0, 0] == 0 and np.array_equal(slices[1:, 0], np.cumsum(slices[:-1, 1])) and slices[-1].sum() == total
    assert np.array_equal(pt.model_vertex_positions, pt.model_vertex_normals) and np.array_equal(pt.model_vertex_positions, pt.model_vertex_tex_coords)
    # every texture index a triangle carries exists, alpha is forced to 255 (texture.cpp:41-47)
    assert pt.triangle_vertex_attributes["texture_idx"].max() == 24
    assert all(np.all(t >> 24 == 255) for t in pt.base_color_textures)
    # byte stability: two bakes, and a save / load round trip
    raw = pt.dumps()
    assert baker.bake(REF_ASSETS / "Sponza.glb").dumps() == raw
    path = tmp_path / "sponza.pt"
    assert baker.main([str(REF_ASSETS / "Sponza.glb"), str(path)]) == 0 and path.read_bytes() == raw
    # the scene the benchmark uses is this bake
    from rayfinder_b200 import assets as rfa

    if rfa.scene_path("Sponza") is not None:
        assert rfa.load_scene("Sponza").dumps() == raw


@needs_assets
@pytest.mark.timeout(600)
def test_bake_agrees_with_the_independent_numpy_restatement():
    """Everything that does not depend on the three documented differences (JPEG decoder, normal normalisation, sort
    stability) is identical to tests/_py_baker.py: the multiset of (triangle, texture) pairs, PNG texels, and — Duck has a
    single mesh and a PNG texture — the whole Duck file."""
    import _py_baker

    assert _py_baker.bake(REF_ASSETS / "Duck.glb").dumps() == baker.bake(REF_ASSETS / "Duck.glb").dumps()
    a, b = baker.bake(REF_ASSETS / "Sponza.glb"), _py_baker.bake(REF_ASSETS / "Sponza.glb")
    assert a.bvh_nodes.tobytes() == b.bvh_nodes.tobytes()  # binned SAH does not depend on the input order of equal-texture meshes

    def keyed(pt):
        rows = np.concatenate([O.triangles9(pt).view(np.uint32), pt.triangle_vertex_attributes["uv0"].view(np.uint32),
                               pt.triangle_vertex_attributes["uv1"].view(np.uint32), pt.triangle_vertex_attributes["uv2"].view(np.uint32),
                               pt.triangle_vertex_attributes["texture_idx"][:, None]], axis=1)
        return rows[np.lexsort(rows.T[::-1])]

    assert np.array_equal(keyed(a), keyed(b))
    for ta, tb in zip(a.base_color_textures, b.base_color_textures):
        assert ta.shape == tb.shape
        diff = np.abs(channels(ta) - channels(tb))
        # PNG and 1x1 factor textures exact; JPEG: stb_image's IDCT / colour conversion against libjpeg-turbo's
        assert diff.max() <= 4 and diff.mean() < 0.1
    exact = [np.array_equal(ta, tb) for ta, tb in zip(a.base_color_textures, b.base_color_textures)]
    assert sum(exact) >= 4  # the four PNG textures (and any flat JPEG)


@needs_assets
def test_image_decoders_against_pillow():
    Image = pytest.importorskip("PIL.Image")
    import _py_baker

    glb = _py_baker.Glb(REF_ASSETS / "Sponza.glb")
    seen = {"png": 0, "jpeg": 0}
    for image in glb.json["images"][:12] + glb.json["images"][-6:]:
        data = glb.view(image["bufferView"])
        mine = baker.texture_from_memory(data)
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"), dtype=np.int32)
        kind = "jpeg" if data[:2] == b"\xff\xd8" else "png"
        seen[kind] += 1
        diff = np.abs(channels(mine) - ref)
        assert mine.shape == ref.shape[:2]
        if kind == "png":
            assert diff.max() == 0
        else:
            assert diff.max() <= 4 and diff.mean() < 0.1
    assert seen["jpeg"] >= 10
    # synthetic PNGs: every colour type / bit depth / interlacing Pillow can write
    rng = np.random.default_rng(7)
    for mode, interlace in (("L", False), ("RGB", False), ("RGBA", True), ("P", False), ("LA", False), ("1", False), ("I;16", False), ("RGB", True)):
        w, h = 37, 23
        if mode == "1":
            img = Image.fromarray((rng.random((h, w)) > 0.5).astype(np.uint8) * 255).convert("1")
        elif mode == "I;16":
            img = Image.fromarray(rng.integers(0, 65536, (h, w)).astype(np.uint16))
        elif mode == "P":
            img = Image.fromarray(rng.integers(0, 256, (h, w, 3)).astype(np.uint8)).quantize(17)
        else:
            nch = {"L": 1, "LA": 2, "RGB": 3, "RGBA": 4}[mode]
            arr = rng.integers(0, 256, (h, w, nch)).astype(np.uint8)
            img = Image.fromarray(arr[..., 0] if nch == 1 else arr, mode)
        buf = io.BytesIO()
        img.save(buf, format="PNG", **({"interlace": 1} if interlace else {}))
        mine = channels(baker.texture_from_memory(buf.getvalue()))
        if mode == "I;16":
            ref = np.repeat((np.asarray(img, dtype=np.int32) >> 8)[..., None], 3, axis=2)  # stb: high byte of a 16-bit sample
        else:
            ref = np.asarray(img.convert("RGB"), dtype=np.int32)
        assert np.array_equal(mine, ref), mode
    # baseline JPEG with 4:2:0 chroma and a grey one: stb's upsampling filters differ from libjpeg's "fancy" ones by little
    for mode, subsampling in (("RGB", 2), ("RGB", 1), ("L", 0)):
        arr = rng.integers(0, 256, (48, 64, 3 if mode == "RGB" else 1)).astype(np.uint8)
        arr = np.repeat(np.repeat(arr[::4, ::4], 4, axis=0), 4, axis=1)  # blocky: keeps the comparison about arithmetic, not ringing
        img = Image.fromarray(arr[..., 0] if mode == "L" else arr, mode)
        buf = io.BytesIO()
        img.save(buf, format="JPEG", quality=92, **({"subsampling": subsampling} if mode == "RGB" else {}))
        mine = channels(baker.texture_from_memory(buf.getvalue()))
        ref = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"), dtype=np.int32)
        assert mine.shape == ref.shape and np.abs(mine - ref).mean() < 3.0, (mode, subsampling)
    # progressive JPEG: rejected, never approximated
    buf = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, (16, 16, 3)).astype(np.uint8)).save(buf, format="JPEG", progressive=True)
    with pytest.raises(rf.RayfinderError) as e:
        baker.texture_from_memory(buf.getvalue())
    assert "progressive" in str(e.value) and e.value.status == capi.RF_ERROR_FORMAT


def _write_gltf(path, translation, factor):
    """A one-triangle .gltf with a base64 buffer: node with a translation + non-uniform scale, material with a factor."""
    import base64
    import json
    import struct

    positions = struct.pack("<9f", 0, 0, 0, 1, 0, 0, 0, 1, 0)
    normals = struct.pack("<9f", 0, 0, 1, 0, 0, 1, 0, 0, 1)
    uvs = struct.pack("<6f", 0, 0, 1, 0, 0, 1)
    indices = struct.pack("<3H", 0, 1, 2) + b"\0\0"
    blob = positions + normals + uvs + indices
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
        "nodes": [{"mesh": 0, "translation": translation, "scale": [2.0, 3.0, 0.5]}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0}]}],
        "materials": [{"pbrMetallicRoughness": {"baseColorFactor": factor}}],
        "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 36},
                        {"buffer": 0, "byteOffset": 72, "byteLength": 24}, {"buffer": 0, "byteOffset": 96, "byteLength": 6}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC3"},
                      {"bufferView": 2, "componentType": 5126, "count": 3, "type": "VEC2"}, {"bufferView": 3, "componentType": 5123, "count": 3, "type": "SCALAR"}],
    }
    path.write_text(json.dumps(doc))


def test_node_transform_normals_and_factor_texture(tmp_path):
    """A translated, non-uniformly scaled node: positions = T*S*p; the normal is the xyz of normalize(vec4(inverseTranspose(M) *
    (n, 0))) — the w component -(M^-1 t).n takes part in the length (gltf_model.cpp:427-428), so the stored normal is NOT
    unit length; a material without a texture becomes a 1x1 texture of its factor (Texture::fromPixel), one per distinct factor."""
    f32 = np.float32
    path = tmp_path / "tri.gltf"
    _write_gltf(path, [1.0, -2.0, 4.0], [0.5, 0.25, 1.0, 1.0])
    pt = baker.bake(path)
    assert pt.bvh_nodes.size == 1 and pt.triangle_position_attributes.shape[0] == 1
    pos = np.stack([pt.triangle_position_attributes[k][0] for k in ("p0", "p1", "p2")])
    assert np.array_equal(pos, np.array([[1, -2, 4], [3, -2, 4], [1, 1, 4]], dtype=f32))
    n = pt.triangle_vertex_attributes["n0"][0]
    # inverse transpose of T(1,-2,4) S(2,3,.5): xyz = n / s = (0, 0, 2), w = -(t / s) . n = -(4 / 0.5) * 1 = -8 -> / sqrt(4 + 64)
    expected = f32(2.0) * (f32(1.0) / np.sqrt(f32(68.0)))
    assert n[0] == 0 and n[1] == 0 and n[2] == expected and abs(float(n[2]) - 1.0) > 0.5
    assert len(pt.base_color_textures) == 1 and pt.base_color_textures[0].shape == (1, 1)
    assert int(pt.base_color_textures[0][0, 0]) == (255 | (63 << 8) | (127 << 16) | (255 << 24))  # b | g << 8 | r << 16 | a << 24, truncated
    assert pt.model_base_color_texture_indices.tolist() == [0] and pt.vertex_indices.tolist() == [0, 1, 2]
    assert np.array_equal(pt.vertex_positions[:, 3], np.ones(3, f32)) and np.array_equal(pt.vertex_normals[:, 3], np.zeros(3, f32))


def test_baker_error_messages(tmp_path):
    missing = tmp_path / "nope.glb"
    with pytest.raises(rf.RayfinderError) as e:
        baker.bake(missing)
    assert str(e.value) == f"The gltf file {missing} does not exist."  # gltf_model.cpp:270-274
    garbage = tmp_path / "garbage.gltf"
    garbage.write_text("{ this is not json")
    with pytest.raises(rf.RayfinderError) as e:
        baker.bake(garbage)
    assert str(e.value) == f"Failed to parse gltf file {garbage}."  # :282-286
    external = tmp_path / "external.gltf"
    external.write_text('{"asset": {"version": "2.0"}, "scenes": [{"nodes": []}], "buffers": [{"byteLength": 4, "uri": "missing.bin"}]}')
    with pytest.raises(rf.RayfinderError) as e:
        baker.bake(external)
    assert str(e.value) == f"Failed to load gltf buffers for {external}."  # :289-294
    assert baker.main([]) == 0 and baker.main([str(missing), str(tmp_path / "out.pt")]) == 1  # pt-format-tool/main.cpp:16-33
