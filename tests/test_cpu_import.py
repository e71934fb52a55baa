"""CPU tests of the asset importers behind kuafu::loadScene (reference
src/core/geometry.cpp:45-232 does this through Assimp): COLLADA, STL, Wavefront, and BASELINE config 3
"loaded via the Assimp path" (million_obj: every mesh written to a Wavefront file and read back)."""
import os
import struct

import numpy as np
import pytest

from kuafu_b200 import host

REF_MODELS = "/root/reference/resources/models"


def _load(name, **kw):
    r = host.Renderer(device=None)
    r.load_scene(name, 32, 32, 1, **kw)
    ws = r.wire_scene()
    r.close()
    return ws


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_MODELS, "suzanne.dae")),
                    reason="the reference's COLLADA asset is only present where /root/reference is mounted")
def test_collada_import_of_the_reference_asset(built):
    """suzanne.dae as Assimp hands it to the reference (Triangulate | GenNormals | FlipUVs |
    PreTransformVertices, no vertex joining): 251 904 triangles, one vertex per face corner."""
    ws = _load("file:" + os.path.join(REF_MODELS, "suzanne.dae"))
    assert len(ws.geoms) == 1
    v, idx, mi, opaque, hide = ws.geoms[0]
    assert idx.size == 3 * 251904 and v.size == 3 * 251904
    assert v.nbytes == 36274176                                       # "suzanne: 36.3 MB" vertex buffer
    assert np.array_equal(idx, np.arange(idx.size, dtype=np.uint32))
    n = np.linalg.norm(v["normal"].astype(np.float64), axis=1)
    assert np.all(np.abs(n - 1.0) < 1e-3)
    uv = v["texCoord"]
    assert uv.min() >= -1e-6 and uv.max() <= 1.0 + 1e-6
    pos = v["pos"]
    assert np.all(np.isfinite(pos)) and np.ptp(pos[:, 0]) > 0.5
    assert opaque and not hide and ws.n_tris() == 251904
    # first corner of the file: <p>15582 0 0 ...: position 15582, normal 0, uv 0 with v flipped
    assert np.allclose(v["normal"][0], [0.7706592, -0.6356531, -0.04505252], atol=1e-7)
    assert np.allclose(uv[0], [0.8909584, 1.0 - 0.5869966], atol=1e-7)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_MODELS, "cube.obj")),
                    reason="the reference's Wavefront assets are only present where /root/reference is mounted")
@pytest.mark.parametrize("name,tris", [("cube.obj", 12), ("plane.obj", 2)])
def test_wavefront_import_of_the_reference_assets(built, name, tris):
    ws = _load("file:" + os.path.join(REF_MODELS, name))
    assert ws.n_tris() == tris
    for v, idx, *_ in ws.geoms:
        assert idx.max() < v.size and np.all(np.isfinite(v["pos"]))


def _tetra():
    p = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    faces = [(0, 2, 1), (0, 1, 3), (0, 3, 2), (1, 2, 3)]
    tris = np.array([[p[a], p[b], p[c]] for a, b, c in faces], np.float32)
    nrm = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return tris, nrm.astype(np.float32)


@pytest.mark.parametrize("binary", [True, False])
def test_stl_import(built, tmp_path, binary):
    tris, nrm = _tetra()
    path = str(tmp_path / ("t_bin.stl" if binary else "t_ascii.stl"))
    if binary:
        with open(path, "wb") as f:
            f.write(b"\0" * 80 + struct.pack("<I", len(tris)))
            for t, n in zip(tris, nrm):
                f.write(struct.pack("<12fH", *n, *t.reshape(-1), 0))
    else:
        with open(path, "w") as f:
            f.write("solid t\n")
            for t, n in zip(tris, nrm):
                f.write("facet normal %.9g %.9g %.9g\nouter loop\n" % tuple(n))
                for p in t:
                    f.write("vertex %.9g %.9g %.9g\n" % tuple(p))
                f.write("endloop\nendfacet\n")
            f.write("endsolid t\n")
    ws = _load("file:" + path)
    v, idx, *_ = ws.geoms[0]
    assert idx.size == 12 and np.array_equal(idx, np.arange(12, dtype=np.uint32))
    assert np.array_equal(v["pos"].reshape(4, 3, 3), tris)
    assert np.array_equal(v["normal"].reshape(4, 3, 3), np.repeat(nrm[:, None, :], 3, axis=1))


def test_collada_transforms_polylists_and_materials(built, tmp_path):
    """A hand-written document: a quad as <polylist> under a translated + scaled node, a phong
    material with shininess (-> roughness rule, geometry.cpp:112-123) and a fully transparent
    'transparency' that the reference repairs to opaque (geometry.cpp:91-97)."""
    dae = """<?xml version="1.0"?>
<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">
  <library_effects><effect id="fx"><profile_COMMON><technique sid="common"><phong>
    <diffuse><color>0.25 0.5 0.75 1</color></diffuse>
    <shininess><float>30</float></shininess>
    <index_of_refraction><float>1.33</float></index_of_refraction>
    <transparent opaque="A_ONE"><color>1 1 1 1</color></transparent><transparency><float>0</float></transparency>
  </phong></technique></profile_COMMON></effect></library_effects>
  <library_materials><material id="mat"><instance_effect url="#fx"/></material></library_materials>
  <library_geometries><geometry id="quad"><mesh>
    <source id="pos"><float_array id="pa" count="12">0 0 0  1 0 0  1 1 0  0 1 0</float_array>
      <technique_common><accessor source="#pa" count="4" stride="3"/></technique_common></source>
    <source id="uv"><float_array id="ua" count="8">0 0 1 0 1 0.25 0 0.25</float_array>
      <technique_common><accessor source="#ua" count="4" stride="2"/></technique_common></source>
    <vertices id="vtx"><input semantic="POSITION" source="#pos"/></vertices>
    <polylist material="m0" count="1">
      <input semantic="VERTEX" source="#vtx" offset="0"/><input semantic="TEXCOORD" source="#uv" offset="1" set="0"/>
      <vcount>4</vcount><p>0 0 1 1 2 2 3 3</p>
    </polylist>
  </mesh></geometry></library_geometries>
  <library_visual_scenes><visual_scene id="s"><node id="n">
    <translate>10 20 30</translate><scale>2 2 2</scale>
    <instance_geometry url="#quad"><bind_material><technique_common>
      <instance_material symbol="m0" target="#mat"/></technique_common></bind_material></instance_geometry>
  </node></visual_scene></library_visual_scenes>
  <scene><instance_visual_scene url="#s"/></scene>
</COLLADA>
"""
    path = str(tmp_path / "quad.dae")
    open(path, "w").write(dae)
    ws = _load("file:" + path)
    v, idx, mi, opaque, hide = ws.geoms[0]
    assert idx.size == 6                                             # the quad as a fan: 2 triangles
    pos = v["pos"]
    assert np.array_equal(pos[0], [10, 20, 30]) and np.array_equal(pos[1], [12, 20, 30])
    assert np.array_equal(pos[2], [12, 22, 30]) and np.array_equal(pos[5], [10, 22, 30])
    assert np.allclose(v["normal"], [0, 0, 1], atol=1e-6)           # generated (no NORMAL input)
    assert np.array_equal(v["texCoord"][2], [1.0, 0.75])            # FlipUVs
    m = ws.mats[int(mi[0])]
    assert np.allclose(m["diffuse"][:3], [0.25, 0.5, 0.75])
    assert m["ior"] == np.float32(1.33) and m["alpha"] == 1.0 and opaque
    assert np.isclose(m["roughness"], 1.0 - np.sqrt(30.0 - 5.0) * 0.025)


def _gltf_doc(bin_len, uri=None):
    """A unit quad (4 vertices, 2 indexed triangles, u16 indices) under a node translated by (1, 2, 3)
    inside a parent scaled by 2, with a metallic-roughness material and the ior / transmission
    extensions."""
    buf = {"byteLength": bin_len}
    if uri is not None:
        buf["uri"] = uri
    return {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
        "nodes": [{"scale": [2, 2, 2], "children": [1]}, {"translation": [1, 2, 3], "mesh": 0}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0}]}],
        "materials": [{"pbrMetallicRoughness": {"baseColorFactor": [0.2, 0.4, 0.6, 0.5], "metallicFactor": 0.25, "roughnessFactor": 0.75},
                       "alphaMode": "BLEND", "emissiveFactor": [0, 0, 0],
                       "extensions": {"KHR_materials_ior": {"ior": 1.33}, "KHR_materials_transmission": {"transmissionFactor": 0.5}}}],
        "buffers": [buf],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 48}, {"buffer": 0, "byteOffset": 48, "byteLength": 48},
                        {"buffer": 0, "byteOffset": 96, "byteLength": 32}, {"buffer": 0, "byteOffset": 128, "byteLength": 12}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3", "min": [0, 0, 0], "max": [1, 1, 0]},
                      {"bufferView": 1, "componentType": 5126, "count": 4, "type": "VEC3"},
                      {"bufferView": 2, "componentType": 5126, "count": 4, "type": "VEC2"},
                      {"bufferView": 3, "componentType": 5123, "count": 6, "type": "SCALAR"}],
    }


def _gltf_bin():
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (4, 1))
    uv = np.array([[0, 0], [1, 0], [1, 0.25], [0, 0.25]], np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    return pos.tobytes() + nrm.tobytes() + uv.tobytes() + idx.tobytes()


@pytest.mark.parametrize("container", ["glb", "gltf+bin", "gltf+base64"])
def test_gltf_import(built, tmp_path, container):
    import base64
    import json
    blob = _gltf_bin()
    if container == "glb":
        js = json.dumps(_gltf_doc(len(blob))).encode()
        js += b" " * (-len(js) % 4)
        body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
        path = str(tmp_path / "quad.glb")
        open(path, "wb").write(b"glTF" + struct.pack("<II", 2, 12 + len(body)) + body)
    elif container == "gltf+bin":
        open(str(tmp_path / "quad.bin"), "wb").write(blob)
        path = str(tmp_path / "quad.gltf")
        open(path, "w").write(json.dumps(_gltf_doc(len(blob), "quad.bin")))
    else:
        path = str(tmp_path / "quad64.gltf")
        uri = "data:application/octet-stream;base64," + base64.b64encode(blob).decode()
        open(path, "w").write(json.dumps(_gltf_doc(len(blob), uri)))
    ws = _load("file:" + path)
    assert len(ws.geoms) == 1
    v, idx, mi, opaque, hide = ws.geoms[0]
    assert np.array_equal(idx, [0, 1, 2, 0, 2, 3]) and v.size == 4       # indexed, as the accessors hold them
    assert np.array_equal(v["pos"], [[2, 4, 6], [4, 4, 6], [4, 6, 6], [2, 6, 6]])   # 2 * (p + (1, 2, 3))
    assert np.allclose(v["normal"], [0, 0, 1], atol=1e-6)
    assert np.array_equal(v["texCoord"], [[0, 1], [1, 1], [1, 0.75], [0, 0.75]])    # FlipUVs
    m = ws.mats[int(mi[0])]
    assert np.allclose(m["diffuse"][:3], [0.2, 0.4, 0.6]) and m["alpha"] == np.float32(0.5) and not opaque
    assert m["metallic"] == np.float32(0.25) and m["roughness"] == np.float32(0.75)
    assert m["ior"] == np.float32(1.33) and m["transmission"] == np.float32(0.5)


def test_unsupported_format_fails_loudly(built, tmp_path):
    path = str(tmp_path / "x.fbx")
    open(path, "w").write("{}")
    with pytest.raises(RuntimeError, match="imports .obj, .dae, .stl, .gltf and .glb only"):
        _load("file:" + path)


def test_config3_through_the_import_path_is_the_same_scene(built):
    """million_obj == million: same triangles in the same order (positions and normals bit for bit,
    v within one ulp of the double FlipUVs), same instances, same material parameters."""
    a, b = _load("million", scale=5), _load("million_obj", scale=5)
    assert len(a.geoms) == len(b.geoms) and a.n_tris() == b.n_tris()
    for (va, ia, ma, oa, ha), (vb, ib, mb, ob, hb) in zip(a.geoms, b.geoms):
        assert ia.size == ib.size and oa == ob and ha == hb
        ca, cb = va[ia], vb[ib]                                      # per-corner attributes
        assert np.array_equal(ca["pos"], cb["pos"]) and np.array_equal(ca["normal"], cb["normal"])
        assert np.array_equal(ca["texCoord"][:, 0], cb["texCoord"][:, 0])
        assert np.max(np.abs(ca["texCoord"][:, 1] - cb["texCoord"][:, 1])) <= 1.2e-7
        pa, pb = a.mats[int(ma[0])], b.mats[int(mb[0])]
        for k in ("diffuse", "alpha", "metallic", "specular", "roughness", "ior", "transmission"):
            assert np.array_equal(pa[k], pb[k]), k
    assert np.array_equal(a.insts["transform"], b.insts["transform"])
    assert np.array_equal(a.insts["geometryIndex"], b.insts["geometryIndex"])
