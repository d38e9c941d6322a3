"""Test helper: writes a SceneBuilder scene as glTF 2.0 (.gltf + .bin + PNG files, or one .glb) so that the library's C++ loader can be checked
against the tables rtxpt_b200/scene_builder.py builds from the same data.  Not part of the product."""
import json, os, struct, zlib
import numpy as np


def png_bytes(rgba):
    rgba = np.ascontiguousarray(rgba, np.uint8); h, w = rgba.shape[:2]
    raw = b"".join(b"\x00" + rgba[y].tobytes() for y in range(h))
    def chunk(t, d): return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")


def export(builder, path, camera=None, glb=False):
    """camera: optional dict(position, direction, up, yfov, znear, zfar)."""
    blob = bytearray(); views = []; accessors = []
    def add_view(data, target=None):
        while len(blob) % 4: blob.append(0)
        v = {"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)}
        if target: v["target"] = target
        blob.extend(data); views.append(v); return len(views) - 1
    def add_accessor(arr, ctype, atype, target=None, minmax=False):
        a = {"bufferView": add_view(arr.tobytes(), target), "componentType": ctype, "count": int(arr.shape[0]), "type": atype}
        if minmax: a["min"] = [float(x) for x in arr.min(0)]; a["max"] = [float(x) for x in arr.max(0)]
        accessors.append(a); return len(accessors) - 1
    images = []; textures = []; base = os.path.splitext(os.path.basename(path))[0]; out_dir = os.path.dirname(path)
    for i, (mips, fmt) in enumerate(builder.textures):
        data = png_bytes(mips[0])
        if glb: images.append({"bufferView": add_view(data), "mimeType": "image/png"})
        else:
            name = "%s_tex%d.png" % (base, i); open(os.path.join(out_dir, name), "wb").write(data); images.append({"uri": name})
        textures.append({"source": i})
    materials = []
    for m in builder.materials:
        e = np.asarray(m.emissive, np.float32) * np.float32(m.emissive_intensity)
        g = {"pbrMetallicRoughness": {"baseColorFactor": [float(np.float32(c)) for c in m.base_color] + [float(np.float32(m.opacity))],
                                      "metallicFactor": float(np.float32(m.metalness)), "roughnessFactor": float(np.float32(m.roughness))},
             "emissiveFactor": [float(x) for x in e], "name": "mat%d" % len(materials)}
        if m.base_texture is not None: g["pbrMetallicRoughness"]["baseColorTexture"] = {"index": m.base_texture}
        if m.orm_texture is not None: g["pbrMetallicRoughness"]["metallicRoughnessTexture"] = {"index": m.orm_texture}
        if m.normal_texture is not None: g["normalTexture"] = {"index": m.normal_texture, "scale": float(np.float32(m.normal_scale))}
        if m.emissive_texture is not None: g["emissiveTexture"] = {"index": m.emissive_texture}
        if m.alpha_test: g["alphaMode"] = "MASK"; g["alphaCutoff"] = float(np.float32(m.alpha_cutoff))
        ext = {}
        if m.enable_transmission:
            ext["KHR_materials_transmission"] = {"transmissionFactor": float(np.float32(m.transmission))}
            ext["KHR_materials_ior"] = {"ior": float(np.float32(m.ior))}
            if not m.thin_surface:
                ext["KHR_materials_volume"] = {"thicknessFactor": 1.0, "attenuationColor": [float(np.float32(c)) for c in m.volume_color]}
                if m.volume_distance < 3e38: ext["KHR_materials_volume"]["attenuationDistance"] = float(np.float32(m.volume_distance))
        if ext: g["extensions"] = ext
        if m.nested_priority: g["extras"] = {"nestedPriority": int(m.nested_priority)}
        materials.append(g)
    meshes = []
    for geos in builder.meshes:
        prims = []
        for gd in geos:
            attrs = {"POSITION": add_accessor(np.asarray(gd["positions"], np.float32), 5126, "VEC3", 34962, True),
                     "NORMAL": add_accessor(np.asarray(gd["normals"], np.float32), 5126, "VEC3", 34962)}
            if gd.get("uvs") is not None: attrs["TEXCOORD_0"] = add_accessor(np.asarray(gd["uvs"], np.float32), 5126, "VEC2", 34962)
            if gd.get("tangents") is not None: attrs["TANGENT"] = add_accessor(np.asarray(gd["tangents"], np.float32), 5126, "VEC4", 34962)
            prims.append({"attributes": attrs, "indices": add_accessor(np.asarray(gd["indices"], np.uint32).reshape(-1), 5125, "SCALAR", 34963), "material": gd["material"], "mode": 4})
        meshes.append({"primitives": prims})
    nodes = []
    for mesh, xf in builder.instances:
        m4 = np.eye(4, dtype=np.float64); m4[:3, :] = np.asarray(xf, np.float32).astype(np.float64)
        nodes.append({"mesh": mesh, "matrix": [float(x) for x in m4.T.reshape(-1)]})          # column-major
    doc = {"asset": {"version": "2.0", "generator": "rtxpt_b200 tests"}, "materials": materials, "meshes": meshes}
    ext_used = sorted({k for m in materials for k in m.get("extensions", {})})
    lights = []
    for L in builder.lights:
        l = {"type": "spot" if L["type"] == 2 else "point", "color": [float(np.float32(c)) for c in L["color"]], "intensity": float(np.float32(L["intensity"])),
             "extras": {"radius": float(np.float32(L["radius"]))}}
        if L["type"] == 2: l["spot"] = {"innerConeAngle": float(np.radians(L["inner"])), "outerConeAngle": float(np.radians(L["outer"]))}
        # orientation: the light points down the node's -Z
        d = np.asarray(L["direction"], np.float64); d = d / np.linalg.norm(d); z = -d
        x = np.cross([0.0, 1.0, 0.0] if abs(z[1]) < 0.99 else [1.0, 0.0, 0.0], z); x /= np.linalg.norm(x); y = np.cross(z, x)
        m4 = np.eye(4); m4[:3, 0], m4[:3, 1], m4[:3, 2], m4[:3, 3] = x, y, z, np.asarray(L["position"], np.float64)
        nodes.append({"matrix": [float(v) for v in m4.T.reshape(-1)], "extensions": {"KHR_lights_punctual": {"light": len(lights)}}})
        lights.append(l)
    if lights: doc["extensions"] = {"KHR_lights_punctual": {"lights": lights}}; ext_used.append("KHR_lights_punctual")
    if camera is not None:
        d = np.asarray(camera["direction"], np.float64); d /= np.linalg.norm(d); z = -d
        x = np.cross(np.asarray(camera["up"], np.float64), z); x /= np.linalg.norm(x); y = np.cross(z, x)
        m4 = np.eye(4); m4[:3, 0], m4[:3, 1], m4[:3, 2], m4[:3, 3] = x, y, z, np.asarray(camera["position"], np.float64)
        nodes.append({"camera": 0, "matrix": [float(v) for v in m4.T.reshape(-1)]})
        doc["cameras"] = [{"type": "perspective", "perspective": {"yfov": camera["yfov"], "znear": camera["znear"], "zfar": camera["zfar"]}}]
    doc["nodes"] = nodes; doc["scenes"] = [{"nodes": list(range(len(nodes)))}]; doc["scene"] = 0
    if images: doc["images"] = images; doc["textures"] = textures
    if ext_used: doc["extensionsUsed"] = sorted(set(ext_used))
    doc["accessors"] = accessors; doc["bufferViews"] = views
    if glb:
        doc["buffers"] = [{"byteLength": len(blob)}]
        js = json.dumps(doc).encode(); js += b" " * (-len(js) % 4); bb = bytes(blob) + b"\0" * (-len(blob) % 4)
        with open(path, "wb") as f:
            f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(bb)))
            f.write(struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(bb), 0x004E4942) + bb)
    else:
        bin_name = base + ".bin"; open(os.path.join(out_dir, bin_name), "wb").write(bytes(blob))
        doc["buffers"] = [{"uri": bin_name, "byteLength": len(blob)}]
        json.dump(doc, open(path, "w"), indent=1)
    return path
