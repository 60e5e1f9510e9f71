"""TEST INFRASTRUCTURE: an independent numpy restatement of the GLB -> .pt baking steps (round 1's baker), kept to cross-check
the C++ baker (rayfinder_b200/csrc/host_baker.cpp, ``rf_bake_gltf``) on everything the two must agree on: vertex positions,
texture coordinates, indices, texture assignment, PNG texels, BVH.  It deliberately differs from the reference — and
therefore from the C++ baker — in three places: images are decoded with Pillow (libjpeg-turbo), not stb_image; normals are
normalised as 3-vectors with an fp64 inverse (the reference normalises the 4-vector of an fp32 cofactor inverse,
gltf_model.cpp:427-428); meshes are ordered with a stable sort (the reference's std::sort is not stable).
"""
from __future__ import annotations

import io
import json
import struct
from pathlib import Path

import numpy as np

from rayfinder_b200.api import (POSITION_ATTRIBUTE_DTYPE, POSITIONS_DTYPE, VERTEX_ATTRIBUTES_DTYPE, PtFormat, build_bvh, build_bvh_device,
                                reorder_attributes)

f32 = np.float32
_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NUM = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


class Glb:
    def __init__(self, path):
        data = Path(path).read_bytes()
        magic, version, length = struct.unpack_from("<III", data, 0)
        if magic != 0x46546C67 or version != 2:
            raise RuntimeError(f"Failed to parse gltf file {path}.")
        off = 12
        self.json = None
        self.bin = b""
        while off < length:
            clen, ctype = struct.unpack_from("<II", data, off)
            chunk = data[off + 8: off + 8 + clen]
            if ctype == 0x4E4F534A:
                self.json = json.loads(chunk)
            elif ctype == 0x004E4942:
                self.bin = chunk
            off += 8 + clen
        if self.json is None:
            raise RuntimeError(f"Failed to parse gltf file {path}.")

    def view(self, idx: int) -> bytes:
        bv = self.json["bufferViews"][idx]
        start = bv.get("byteOffset", 0)
        return self.bin[start: start + bv["byteLength"]]

    def accessor(self, idx: int) -> np.ndarray:
        acc = self.json["accessors"][idx]
        bv = self.json["bufferViews"][acc["bufferView"]]
        dt = np.dtype(_COMPONENT[acc["componentType"]]).newbyteorder("<")
        n = _NUM[acc["type"]]
        start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * n
        count = acc["count"]
        if stride == dt.itemsize * n:
            arr = np.frombuffer(self.bin, dtype=dt, count=count * n, offset=start).reshape(count, n)
        else:
            arr = np.lib.stride_tricks.as_strided(
                np.frombuffer(self.bin, dtype=dt, offset=start, count=((count - 1) * stride) // dt.itemsize + n),
                shape=(count, n), strides=(stride, dt.itemsize))
        return np.array(arr)


# glm 0.9.9.8 scalar formulas, column-major 4x4 stored as m[col][row] -------------------------------------
def _mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros((4, 4), dtype=f32)
    for c in range(4):  # Result[c] = ((A0*b[c][0] + A1*b[c][1]) + A2*b[c][2]) + A3*b[c][3]
        acc = a[0] * b[c][0]
        acc = (acc + a[1] * b[c][1]).astype(f32)
        acc = (acc + a[2] * b[c][2]).astype(f32)
        out[c] = (acc + a[3] * b[c][3]).astype(f32)
    return out


def _mat_vec(m: np.ndarray, v: np.ndarray) -> np.ndarray:
    """mat4 * vec4 for many vectors v (n, 4): (m0*v0 + m1*v1) + (m2*v2 + m3*v3)."""
    add0 = (m[0][None, :] * v[:, 0:1] + m[1][None, :] * v[:, 1:2]).astype(f32)
    add1 = (m[2][None, :] * v[:, 2:3] + m[3][None, :] * v[:, 3:4]).astype(f32)
    return (add0 + add1).astype(f32)


def _local_matrix(node: dict) -> np.ndarray:
    if "matrix" in node:
        return np.array(node["matrix"], dtype=f32).reshape(4, 4)  # column-major: rows of this array = columns
    s = np.array(node.get("scale", [1, 1, 1]), dtype=f32)
    q = np.array(node.get("rotation", [0, 0, 0, 1]), dtype=f32)  # x, y, z, w
    t = np.array(node.get("translation", [0, 0, 0]), dtype=f32)
    eye = np.eye(4, dtype=f32)
    scale = eye.copy()
    for i in range(3):
        scale[i] = eye[i] * s[i]
    x, y, z, w = q
    rot = eye.copy()  # glm::mat3_cast
    rot[0][0] = f32(1) - f32(2) * (y * y + z * z)
    rot[0][1] = f32(2) * (x * y + w * z)
    rot[0][2] = f32(2) * (x * z - w * y)
    rot[1][0] = f32(2) * (x * y - w * z)
    rot[1][1] = f32(1) - f32(2) * (x * x + z * z)
    rot[1][2] = f32(2) * (y * z + w * x)
    rot[2][0] = f32(2) * (x * z + w * y)
    rot[2][1] = f32(2) * (y * z - w * x)
    rot[2][2] = f32(1) - f32(2) * (x * x + y * y)
    trans = eye.copy()
    trans[3] = ((eye[0] * t[0] + eye[1] * t[1]).astype(f32) + eye[2] * t[2]).astype(f32) + eye[3]
    return _mat_mul(_mat_mul(trans, rot), scale)  # translation * rotation * scale


def _decode_image(data: bytes) -> np.ndarray:
    from PIL import Image

    img = Image.open(io.BytesIO(data)).convert("RGBA")
    px = np.asarray(img, dtype=np.uint32)
    # Texture::fromMemory (common/texture.cpp:36-47): b | g << 8 | r << 16 | 255 << 24
    return (px[..., 2] | (px[..., 1] << 8) | (px[..., 0] << 16) | np.uint32(255 << 24)).astype("<u4")


def _fnv1a(data: bytes) -> int:
    h = 2166136261
    for b in data:
        h = ((h ^ b) * 16777619) & 0xFFFFFFFF
    return h


def load_gltf_model(path):
    """-> (meshes sorted by texture index, textures); mesh = dict(positions, normals, tex_coords, indices, texture)."""
    glb = Glb(path)
    js = glb.json
    if len(js.get("scenes", [])) != 1:
        raise RuntimeError("expected exactly one scene")
    num_meshes = len(js["meshes"])
    transforms = [(np.eye(4, dtype=f32), np.eye(4, dtype=f32)) for _ in range(num_meshes)]

    def traverse(node_idx: int, parent: np.ndarray) -> None:
        node = js["nodes"][node_idx]
        m = _mat_mul(parent, _local_matrix(node))
        # glm::inverseTranspose, in fp64 then rounded; stored like m (array row = matrix column)
        normal = np.linalg.inv(m.astype(np.float64)).T.astype(f32)
        if "mesh" in node:
            transforms[node["mesh"]] = (m, normal)
        for child in node.get("children", []):
            traverse(child, m)

    for root in js["scenes"][js.get("scene", 0)]["nodes"]:
        traverse(root, np.eye(4, dtype=f32))

    textures: list[np.ndarray] = []
    image_lookup: dict[int, int] = {}
    factor_lookup: dict[int, int] = {}
    meshes = []
    for mesh_idx, mesh in enumerate(js["meshes"]):
        m, nm = transforms[mesh_idx]
        for prim in mesh["primitives"]:
            if prim.get("mode", 4) != 4:
                raise RuntimeError("only triangle primitives are supported")
            pbr = js["materials"][prim["material"]].get("pbrMetallicRoughness", {})
            if "baseColorTexture" in pbr:
                tex = js["textures"][pbr["baseColorTexture"]["index"]]
                sampler = js["samplers"][tex["sampler"]]
                assert sampler.get("wrapS", 10497) == 10497 and sampler.get("wrapT", 10497) == 10497
                image_idx = tex["source"]
                if image_idx not in image_lookup:
                    image_lookup[image_idx] = len(textures)
                    textures.append(_decode_image(glb.view(js["images"][image_idx]["bufferView"])))
                tex_idx = image_lookup[image_idx]
            else:
                factor = np.array(pbr.get("baseColorFactor", [1, 1, 1, 1]), dtype=f32)
                h = _fnv1a(factor.tobytes())
                if h not in factor_lookup:
                    factor_lookup[h] = len(textures)
                    c = (factor * f32(255.0)).astype(np.uint32)  # Texture::fromPixel, texture.cpp:56-65
                    textures.append(np.array([[c[2] | (c[1] << 8) | (c[0] << 16) | (c[3] << 24)]], dtype="<u4"))
                tex_idx = factor_lookup[h]

            indices = glb.accessor(prim["indices"]).reshape(-1).astype(np.uint32)
            assert indices.size % 3 == 0
            local_p = glb.accessor(prim["attributes"]["POSITION"]).astype(f32)
            local_n = glb.accessor(prim["attributes"]["NORMAL"]).astype(f32)
            uv = glb.accessor(prim["attributes"]["TEXCOORD_0"]).astype(f32)
            ones = np.ones((local_p.shape[0], 1), dtype=f32)
            positions = _mat_vec(m, np.concatenate([local_p, ones], axis=1))[:, :3]
            n4 = _mat_vec(nm, np.concatenate([local_n, 0 * ones], axis=1))[:, :3]
            d = ((n4[:, 0] * n4[:, 0] + n4[:, 1] * n4[:, 1]).astype(f32) + n4[:, 2] * n4[:, 2]).astype(f32)
            normals = (n4 * (f32(1.0) / np.sqrt(d))[:, None]).astype(f32)
            meshes.append(dict(positions=positions, normals=normals, tex_coords=uv, indices=indices, texture=tex_idx))

    order = sorted(range(len(meshes)), key=lambda i: meshes[i]["texture"])  # stable
    return [meshes[i] for i in order], textures


def bake(gltf_path, bvh_device: "int | None" = None) -> PtFormat:
    """``PtFormat(gltfPath)``.  ``bvh_device``: build the BVH on that CUDA device (``rf_build_bvh_device``, byte-identical
    to the host builder and ~20x faster on Sponza) instead of on the host (``rf_build_bvh``, the default: the baker also
    runs where there is no GPU)."""
    meshes, textures = load_gltf_model(gltf_path)

    # FlattenedModel (flattened_model.cpp:8-46)
    pos = np.concatenate([m["positions"][m["indices"]].reshape(-1, 3, 3) for m in meshes]).astype(f32)
    nrm = np.concatenate([m["normals"][m["indices"]].reshape(-1, 3, 3) for m in meshes]).astype(f32)
    uvs = np.concatenate([m["tex_coords"][m["indices"]].reshape(-1, 3, 2) for m in meshes]).astype(f32)
    tex = np.concatenate([np.full(m["indices"].size // 3, m["texture"], dtype=np.uint32) for m in meshes])

    if bvh_device is None:
        nodes, tri_idx = build_bvh(pos)
    else:
        nodes, tri_idx, _ = build_bvh_device(pos, bvh_device)
    pos, nrm, uvs, tex = (reorder_attributes(a, tri_idx) for a in (pos, nrm, uvs, tex))

    pt = PtFormat()
    pt.bvh_nodes = nodes
    bpa = np.zeros(pos.shape[0], dtype=POSITIONS_DTYPE)
    bpa["v0"], bpa["v1"], bpa["v2"] = pos[:, 0], pos[:, 1], pos[:, 2]
    pt.bvh_position_attributes = bpa
    pa = np.zeros(pos.shape[0], dtype=POSITION_ATTRIBUTE_DTYPE)
    pa["p0"], pa["p1"], pa["p2"] = pos[:, 0], pos[:, 1], pos[:, 2]
    pt.triangle_position_attributes = pa
    va = np.zeros(pos.shape[0], dtype=VERTEX_ATTRIBUTES_DTYPE)
    va["n0"], va["n1"], va["n2"] = nrm[:, 0], nrm[:, 1], nrm[:, 2]
    va["uv0"], va["uv1"], va["uv2"] = uvs[:, 0], uvs[:, 1], uvs[:, 2]
    va["texture_idx"] = tex
    pt.triangle_vertex_attributes = va

    # per-mesh raster data (pt_format.cpp:84-148)
    vp, vn, vt, vi = [], [], [], []
    s_p, s_i = [], []
    v_off = i_off = 0
    for m in meshes:
        n_v, n_i = m["positions"].shape[0], m["indices"].size
        vp.append(np.concatenate([m["positions"], np.ones((n_v, 1), f32)], axis=1))
        vn.append(np.concatenate([m["normals"], np.zeros((n_v, 1), f32)], axis=1))
        vt.append(m["tex_coords"])
        vi.append(m["indices"])
        s_p.append((v_off, n_v))
        s_i.append((i_off, n_i))
        v_off += n_v
        i_off += n_i
    pt.vertex_positions = np.concatenate(vp).astype(f32)
    pt.vertex_normals = np.concatenate(vn).astype(f32)
    pt.vertex_tex_coords = np.concatenate(vt).astype(f32)
    pt.vertex_indices = np.concatenate(vi).astype(np.uint32)
    pt.model_vertex_positions = np.array(s_p, dtype=np.uint64).reshape(-1, 2)
    pt.model_vertex_normals = pt.model_vertex_positions.copy()
    pt.model_vertex_tex_coords = pt.model_vertex_positions.copy()
    pt.model_vertex_indices = np.array(s_i, dtype=np.uint64).reshape(-1, 2)
    pt.model_base_color_texture_indices = np.array([m["texture"] for m in meshes], dtype=np.uint32)
    pt.base_color_textures = textures
    return pt
