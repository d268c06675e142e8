"""Host-side scene preparation for the testrender mirror (harness code: the
caller of the hot path, not the hot path).

Follows the reference's scene handling so both the CUDA integrator and the CPU
oracle consume identical arrays:
  * XML scene grammar + serialized group spec   src/testrender/simpleraytracer.cpp:281-512,
                                                src/liboslexec/shadingsys.cpp:3232-3300
  * sphere / quad tessellation                  src/testrender/scene.cpp:138-262
  * OBJ models (triangulated, v/vt/vn)          src/testrender/scene.cpp:14-136
  * binned-SAH BVH (16 bins, depth 64)          src/testrender/bvh.cpp:43-219
  * mesh areas + light triangle list            src/testrender/simpleraytracer.cpp:1272-1300
"""
import math
import os
import re
import xml.etree.ElementTree as ET

import numpy as np

f32 = np.float32
NUM_BINS = 16
MAX_DEPTH = 64


def _vec(s):
    v = [float(x) for x in re.split(r"[,\s]+", s.strip()) if x]
    return np.array((v + [0, 0, 0])[:3], f32)


def parse_group_spec(text):
    """`color Cs .75 .25 .25; shader matte layer1; connect a.x b.y` ->
    (layers [{shader, name, params}], connections)."""
    layers, conns, pending = [], [], {}
    for stmt in re.split(r"[;,]", text):
        tok = stmt.split()
        if not tok:
            continue
        if tok[0] == "param":
            tok = tok[1:]
        if tok[0] == "shader":
            layers.append(dict(shader=tok[1], name=tok[2] if len(tok) > 2 else tok[1], params=pending))
            pending = {}
        elif tok[0] == "connect":
            (sl, sp), (dl, dp) = tok[1].split(".", 1), tok[2].split(".", 1)
            conns.append((sl, sp, dl, dp))
        elif re.fullmatch(r"int(\[\d*\])?", tok[0]):
            pending[tok[1]] = [int(x) for x in tok[2:]]
        elif re.fullmatch(r"(float|color|point|vector|normal|matrix)(\[\d*\])?", tok[0]):
            pending[tok[1]] = [float(x) for x in tok[2:]]
        elif tok[0] == "string":
            pending[tok[1]] = [" ".join(tok[2:]).strip('"')]
        elif re.fullmatch(r"string\[\d*\]", tok[0]):
            pending[tok[1]] = re.findall(r'"([^"]*)"', stmt)
        else:
            raise ValueError("bad group statement: %r" % stmt)
    return layers, conns


class Scene:
    def __init__(self):
        self.verts, self.normals, self.uvs = [], [], []
        self.triangles, self.n_triangles, self.uv_triangles = [], [], []
        self.shaderids, self.last_index = [], []
        self.materials = []       # [(layers, connections)]
        self.is_light = []
        self.eye = np.zeros(3, f32)
        self.dir = np.array([0, 0, -1], f32)
        self.up = np.array([0, 1, 0], f32)
        self.fov = 90.0
        self.background_shader = -1
        self.background_resolution = 0
        self.options = {}         # <Option name="int N"/> (max_bounces, rr_depth, ...)
        self.displacements = {}   # material index -> (layers, connections) of its displacement group

    # -- geometry (scene.cpp) -------------------------------------------------
    def add_sphere(self, c, r, shader, resolution=64):
        W, H = 2 * resolution, resolution
        NV = 2 + W * H
        base, nbase, tbase = len(self.verts), len(self.normals), len(self.uvs)
        c = c.astype(f32)
        r = f32(r)
        self.verts.append(c + np.array([0, r, 0], f32))
        self.normals.append(np.array([0, 1, 0], f32))
        for y in range(H):
            t = f32(y + 0.5) / f32(H)
            z = f32(math.cos(float(t * f32(math.pi))))
            q = f32(np.sqrt(max(f32(0), f32(1) - z * z)))
            for x in range(W):
                a = f32(2 * math.pi) * f32(x) / f32(W)
                n = np.array([q * f32(-math.sin(float(a))), z, q * f32(-math.cos(float(a)))], f32)
                self.verts.append((c + r * n).astype(f32))
                self.normals.append(n)
        self.verts.append(c - np.array([0, r, 0], f32))
        self.normals.append(np.array([0, -1, 0], f32))
        for y in range(2):
            for x in range(W):
                self.uvs.append((f32(x + 0.5) / f32(W), f32(y)))
        for y in range(H):
            for x in range(W + 1):
                self.uvs.append((f32(x) / f32(W), f32(y + 0.5) / f32(H)))
        x1 = W - 1
        for x0 in range(W):
            self.triangles.append((base, base + 1 + x1, base + 1 + x0))
            self.n_triangles.append((nbase, nbase + 1 + x1, nbase + 1 + x0))
            self.uv_triangles.append((tbase + x1, tbase + 2 * W + x1, tbase + 2 * W + x1 + 1))
            self.shaderids.append(shader)
            for y in range(H - 1):
                i00, i10 = 1 + (x0 + W * y), 1 + (x1 + W * y)
                i11, i01 = 1 + (x1 + W * (y + 1)), 1 + (x0 + W * (y + 1))
                self.triangles += [(base + i00, base + i10, base + i11), (base + i00, base + i11, base + i01)]
                self.n_triangles += [(nbase + i00, nbase + i10, nbase + i11), (nbase + i00, nbase + i11, nbase + i01)]
                t00 = 2 * W + x1 + 1 + (W + 1) * y
                t10 = 2 * W + x1 + (W + 1) * y
                t11 = 2 * W + x1 + (W + 1) * (y + 1)
                t01 = 2 * W + x1 + 1 + (W + 1) * (y + 1)
                self.uv_triangles += [(tbase + t00, tbase + t10, tbase + t11), (tbase + t00, tbase + t11, tbase + t01)]
                self.shaderids += [shader, shader]
            self.triangles.append((base + NV - 1, base + NV - 1 - W + x0, base + NV - 1 - W + x1))
            self.n_triangles.append((nbase + NV - 1, nbase + NV - 1 - W + x0, nbase + NV - 1 - W + x1))
            self.uv_triangles.append((tbase + W + x1, tbase + 2 * W + x1 + 1 + (W + 1) * (H - 1),
                                      tbase + 2 * W + x1 + (W + 1) * (H - 1)))
            self.shaderids.append(shader)
            x1 = x0
        self.last_index.append(len(self.triangles))

    def add_quad(self, p, ex, ey, shader, resolution=1):
        base, tbase = len(self.verts), len(self.uvs)
        p, ex, ey = p.astype(f32), ex.astype(f32), ey.astype(f32)
        R = resolution
        for v in range(R + 1):
            for u in range(R + 1):
                s, t = f32(u) / f32(R), f32(v) / f32(R)
                self.verts.append(((p + s * ex).astype(f32) + t * ey).astype(f32))
                self.uvs.append((s, t))
        for v in range(R):
            for u in range(R):
                i00, i10 = u + v * (R + 1), u + 1 + v * (R + 1)
                i11, i01 = u + 1 + (v + 1) * (R + 1), u + (v + 1) * (R + 1)
                self.triangles += [(base + i00, base + i10, base + i11), (base + i00, base + i11, base + i01)]
                self.n_triangles += [(-1, -1, -1), (-1, -1, -1)]
                self.uv_triangles += [(tbase + i00, tbase + i10, tbase + i11), (tbase + i00, tbase + i11, tbase + i01)]
                self.shaderids += [shader, shader]
        self.last_index.append(len(self.triangles))

    def add_model(self, filename, shader):
        """Wavefront OBJ: v / vt / vn / f (fan-triangulated); one mesh per file."""
        vbase, nbase, tbase = len(self.verts), len(self.normals), len(self.uvs)
        nv = nn = nt = 0
        faces = []
        with open(filename) as f:
            for line in f:
                t = line.split()
                if not t:
                    continue
                if t[0] == "v":
                    self.verts.append(np.array([float(x) for x in t[1:4]], f32))
                    nv += 1
                elif t[0] == "vn":
                    self.normals.append(np.array([float(x) for x in t[1:4]], f32))
                    nn += 1
                elif t[0] == "vt":
                    self.uvs.append((f32(float(t[1])), f32(float(t[2]))))
                    nt += 1
                elif t[0] == "f":
                    faces.append(t[1:])
        for face in faces:
            idx = []
            for c in face:
                p = (c.split("/") + ["", ""])[:3]

                def fix(s, n):
                    if not s:
                        return -1
                    i = int(s)
                    return i - 1 if i > 0 else n + i
                idx.append((fix(p[0], nv), fix(p[1], nt), fix(p[2], nn)))
            for k in range(1, len(idx) - 1):
                a, b, c = idx[0], idx[k], idx[k + 1]
                self.triangles.append((vbase + a[0], vbase + b[0], vbase + c[0]))
                self.n_triangles.append((-1, -1, -1) if a[2] < 0 else (nbase + a[2], nbase + b[2], nbase + c[2]))
                self.uv_triangles.append((-1, -1, -1) if a[1] < 0 else (tbase + a[1], tbase + b[1], tbase + c[1]))
                self.shaderids.append(shader)
        self.last_index.append(len(self.triangles))

    # -- BVH (bvh.cpp) --------------------------------------------------------
    def build_bvh(self, verts=None):
        """The scene BVH, built by the library's native builder (b200_build_bvh); verts: the displaced
        vertices when the scene has displacement groups."""
        import ctypes
        from .. import api
        L = api.lib()
        verts = np.ascontiguousarray(np.array(self.verts if verts is None else verts, f32).reshape(-1, 3))
        tris = np.ascontiguousarray(np.array(self.triangles, np.int32).reshape(-1, 3))
        n = len(tris)
        nodes = np.zeros((max(2 * n - 1, 1), 8), f32)
        indices = np.zeros(n, np.uint32)
        count = ctypes.c_int(0)
        L.b200_build_bvh.restype = ctypes.c_int
        rc = L.b200_build_bvh(ctypes.c_void_p(verts.ctypes.data), ctypes.c_int(len(verts)),
                              ctypes.c_void_p(tris.ctypes.data), ctypes.c_int(n),
                              ctypes.c_void_p(nodes.ctypes.data), ctypes.c_int(len(nodes)),
                              ctypes.c_void_p(indices.ctypes.data), ctypes.byref(count))
        if rc != 0:
            raise ValueError("b200_build_bvh failed (%d): bad triangle indices?" % rc)
        return nodes[:count.value].copy(), indices

    def build_bvh_py(self):
        """The same builder in numpy (slow; kept as the cross-check of the native one in tests)."""
        verts = np.array(self.verts, f32).reshape(-1, 3)
        tris = np.array(self.triangles, np.int32).reshape(-1, 3)
        n = len(tris)
        tv = verts[tris]                                  # [n,3,3]
        bmin, bmax = tv.min(axis=1), tv.max(axis=1)       # triangle bounds
        cen = ((bmin + bmax) * f32(0.5)).astype(f32)      # Box3::center()
        indices = np.arange(n, dtype=np.uint32)
        nodes = [[None] * 8]

        def set_node(i, lo, hi, child=0, nprims=0):
            nodes[i] = [lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], child, nprims]

        def half_area(lo, hi):
            d = (hi - lo).astype(f32)
            return f32(f32(d[0] * d[1]) + f32(d[1] * d[2])) + f32(d[2] * d[0])

        set_node(0, bmin.min(axis=0), bmax.max(axis=0))
        cur = dict(cmin=cen.min(axis=0), cmax=cen.max(axis=0), left=0, right=n, depth=1, node=0)
        stack = []
        while True:
            left, right = cur["left"], cur["right"]
            nprims = right - left
            split = None
            if nprims > 1 and cur["depth"] < MAX_DEPTH:
                prims = indices[left:right]
                ext = (cur["cmax"] - cur["cmin"]).astype(f32)
                binf = np.where(ext > 0, f32(0.999 * NUM_BINS) / np.where(ext > 0, ext, f32(1)), f32(0)).astype(f32)
                binid = ((cen[prims] - cur["cmin"]).astype(f32) * binf).astype(f32).astype(np.int32)
                nb = nodes[cur["node"]]
                node_lo = np.array([nb[0], nb[2], nb[4]], f32)
                node_hi = np.array([nb[1], nb[3], nb[5]], f32)
                inv_area = f32(1) / half_area(node_lo, node_hi)
                best_cost, best = f32(nprims), None
                for axis in range(3):
                    if binf[axis] == 0:
                        continue
                    ids = binid[:, axis]
                    cnt = np.bincount(ids, minlength=NUM_BINS)
                    lo = np.full((NUM_BINS, 3), np.inf, f32)
                    hi = np.full((NUM_BINS, 3), -np.inf, f32)
                    np.minimum.at(lo, ids, bmin[prims])
                    np.maximum.at(hi, ids, bmax[prims])
                    numL = np.cumsum(cnt)
                    accl_lo = np.minimum.accumulate(lo, axis=0)
                    accl_hi = np.maximum.accumulate(hi, axis=0)
                    rlo, rhi = lo[NUM_BINS - 1].copy(), hi[NUM_BINS - 1].copy()
                    for i in range(NUM_BINS - 2, -1, -1):
                        if numL[i] == 0 or numL[i] == nprims:
                            continue
                        areaR = half_area(rlo, rhi)
                        areaL = half_area(accl_lo[i], accl_hi[i])
                        cost = f32(4) + inv_area * (f32(areaL * f32(numL[i])) + f32(areaR * f32(nprims - numL[i])))
                        if cost < best_cost:
                            best_cost, best = cost, (axis, i)
                        rlo = np.minimum(rlo, lo[i])
                        rhi = np.maximum(rhi, hi[i])
                if best is not None:
                    axis, bbin = best
                    # in-place partition exactly like the reference (swap with the shrinking right end)
                    i, r = left, right
                    while i < r:
                        prim = indices[i]
                        b = int(f32(f32(cen[prim, axis] - cur["cmin"][axis]) * binf[axis]))
                        if b <= bbin:
                            i += 1
                        else:
                            r -= 1
                            indices[i], indices[r] = indices[r], indices[i]
                    mid = r
                    L, Rr = indices[left:mid], indices[mid:right]
                    nxt = len(nodes)
                    nodes.append([None] * 8)
                    nodes.append([None] * 8)
                    nodes[cur["node"]][6] = nxt
                    nodes[cur["node"]][7] = 0
                    set_node(nxt, bmin[L].min(axis=0), bmax[L].max(axis=0))
                    set_node(nxt + 1, bmin[Rr].min(axis=0), bmax[Rr].max(axis=0))
                    c0 = dict(cmin=cen[L].min(axis=0), cmax=cen[L].max(axis=0), left=left, right=mid,
                              depth=cur["depth"] + 1, node=nxt)
                    c1 = dict(cmin=cen[Rr].min(axis=0), cmax=cen[Rr].max(axis=0), left=mid, right=right,
                              depth=cur["depth"] + 1, node=nxt + 1)
                    stack.append(c1)
                    cur = c0
                    split = True
            if split:
                continue
            nodes[cur["node"]][6] = left
            nodes[cur["node"]][7] = nprims
            if not stack:
                break
            cur = stack.pop()
        arr = np.zeros((len(nodes), 8), f32)
        for i, nd in enumerate(nodes):
            arr[i, :6] = nd[:6]
            arr[i, 6:] = np.array(nd[6:], np.uint32).view(f32)
        return arr, indices

    # -- finalize ---------------------------------------------------------------
    def displace_geometry(self, verts, normals, tris, displace):
        """SimpleRaytracer::prepare_geometry (simpleraytracer.cpp:1300-1420): every corner of every triangle
        whose material has a displacement group is one shading point (P, N, Ng, u, v, I, surfacearea) of
        that group; the shader's P is handed back, a vertex ends at the average of its corners, and smooth
        normals are rebuilt from the displaced triangles.  `displace(layers, connections, globals, n)`
        runs the group over the n points (the hot path itself: a ShaderGroup over an SoA batch) and returns
        the P the shaders left, float32 [n, 3].  The float32 sums run in the reference's order."""
        ntri = len(tris)
        n_tris = np.array(self.n_triangles, np.int32).reshape(-1, 3)
        uv_tris = np.array(self.uv_triangles, np.int32).reshape(-1, 3)
        uvs = np.array(self.uvs if self.uvs else [[0, 0]], f32).reshape(-1, 2)
        shaderids = np.array(self.shaderids, np.int32)
        p = verts[tris]                                              # [ntri, 3 corners, 3]
        cr = np.cross((p[:, 0] - p[:, 1]).astype(f32), (p[:, 0] - p[:, 2]).astype(f32)).astype(f32)
        ln = np.sqrt((cr[:, 0] * cr[:, 0] + cr[:, 1] * cr[:, 1] + cr[:, 2] * cr[:, 2]).astype(f32)).astype(f32)
        area = (f32(0.5) * ln).astype(f32)
        with np.errstate(invalid="ignore", divide="ignore"):
            Ng = np.where(ln[:, None] > 0, cr / ln[:, None], cr).astype(f32)          # Imath normalize()
        smooth = n_tris[:, 0] >= 0
        N = np.where(smooth[:, None, None], normals[np.maximum(n_tris, 0)], Ng[:, None, :]).astype(f32)
        has_uv = uv_tris[:, 0] >= 0
        uv = np.where(has_uv[:, None, None], uvs[np.maximum(uv_tris, 0)], f32(0)).astype(f32)
        newp = p.copy()
        has_smooth_normals = False
        for mat, (layers, conns) in sorted(self.displacements.items()):
            sel = np.nonzero(shaderids == mat)[0]
            if not len(sel):
                continue
            has_smooth_normals |= bool(smooth[sel].any())
            P = p[sel].reshape(-1, 3)
            d = (P - self.eye.astype(f32)).astype(f32)
            dl = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(f32)).astype(f32)
            with np.errstate(invalid="ignore", divide="ignore"):
                I = np.where(dl[:, None] > 0, d / dl[:, None], d).astype(f32)
            g = dict(P=P, N=N[sel].reshape(-1, 3), Ng=np.repeat(Ng[sel], 3, axis=0), I=I,
                     u=uv[sel].reshape(-1, 2)[:, 0].copy(), v=uv[sel].reshape(-1, 2)[:, 1].copy(),
                     surfacearea=np.repeat(area[sel], 3))
            out = np.asarray(displace(layers, conns, g, len(P)), f32).reshape(-1, 3, 3)
            newp[sel] = out
        disp = np.zeros_like(verts)
        valence = np.zeros(len(verts), np.int32)
        flat = tris.reshape(-1)
        np.add.at(valence, flat, 1)
        for c in range(3):              # float32 sums in corner order a, b, c of ascending triangles
            np.add.at(disp[:, c], flat, newp.reshape(-1, 3)[:, c])
        used = valence > 0
        disp[used] = (disp[used] / valence[used, None].astype(f32)).astype(f32)
        disp[~used] = verts[~used]
        if has_smooth_normals:
            q = disp[tris]
            cr2 = np.cross((q[:, 0] - q[:, 1]).astype(f32), (q[:, 0] - q[:, 2]).astype(f32)).astype(f32)
            acc = np.zeros_like(normals)
            idx = n_tris[smooth].reshape(-1)
            src = np.repeat(cr2[smooth], 3, axis=0)
            for c in range(3):
                np.add.at(acc[:, c], idx, src[:, c])
            l2 = np.sqrt((acc[:, 0] * acc[:, 0] + acc[:, 1] * acc[:, 1] + acc[:, 2] * acc[:, 2]).astype(f32)).astype(f32)
            with np.errstate(invalid="ignore", divide="ignore"):
                normals = np.where(l2[:, None] > 0, acc / l2[:, None], acc).astype(f32)
        return disp.astype(f32), normals

    def prepare(self, displace=None):
        """displace: required when the scene has displacement groups, see displace_geometry()."""
        verts = np.array(self.verts, f32).reshape(-1, 3)
        tris = np.array(self.triangles, np.int32).reshape(-1, 3)
        normals = np.array(self.normals if self.normals else [[0, 0, 0]], f32).reshape(-1, 3)
        if self.displacements:
            if displace is None:
                raise ValueError("the scene has displacement shaders: prepare(displace=...) must run them")
            verts, normals = self.displace_geometry(verts, normals, tris, displace)
        out = dict(
            verts=verts, normals=normals,
            uvs=np.array(self.uvs if self.uvs else [[0, 0]], f32).reshape(-1, 2),
            triangles=tris, n_triangles=np.array(self.n_triangles, np.int32).reshape(-1, 3),
            uv_triangles=np.array(self.uv_triangles, np.int32).reshape(-1, 3),
            shaderids=np.array(self.shaderids, np.int32))
        nodes, indices = self.build_bvh(verts)
        out["bvh_nodes"], out["bvh_indices"] = nodes, indices
        # per-triangle mesh id, per-mesh area (prepare_lights)
        meshids = np.zeros(len(tris), np.int32)
        areas, first = [], 0
        va, vb, vc = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
        cr = np.cross((va - vb).astype(f32), (va - vc).astype(f32)).astype(f32)
        tri_area = (f32(0.5) * np.sqrt((cr * cr).sum(axis=1, dtype=f32))).astype(f32)
        for m, last in enumerate(self.last_index):
            meshids[first:last] = m
            a = f32(0)
            for t in range(first, last):
                a = f32(a + tri_area[t])
            areas.append(a)
            first = last
        out["meshids"], out["mesh_surfacearea"] = meshids, np.array(areas, f32)
        is_light = np.array(self.is_light, np.int32)
        out["shader_is_light"] = is_light
        out["lightprims"] = np.array([t for t in range(len(tris)) if is_light[self.shaderids[t]]], np.uint32)
        return out


def load_scene(xmlfile):
    """Parse a testrender scene file (or XML text) into a Scene."""
    if os.path.exists(xmlfile):
        root = ET.parse(xmlfile).getroot()
        basedir = os.path.dirname(os.path.abspath(xmlfile))
    else:
        root = ET.fromstring(xmlfile)
        basedir = "."
    if root.tag != "World":
        raise ValueError("Error reading scene: Root element <World> is missing")
    sc = Scene()
    sc.basedir = basedir      # texture file names in shader parameters are relative to the scene
    named = {}
    for node in root:
        a = node.attrib
        if node.tag == "Option":
            # simpleraytracer.cpp:302-309: only "int N" values are honoured
            for k, v in a.items():
                v = v.strip()
                if v.startswith("int "):
                    try:
                        sc.options[k] = int(v[4:].split()[0])
                    except (ValueError, IndexError):
                        pass
        elif node.tag == "Camera":
            if "eye" in a:
                sc.eye = _vec(a["eye"])
            if "dir" in a:
                sc.dir = _vec(a["dir"])
            elif "look_at" in a:
                sc.dir = (_vec(a["look_at"]) - sc.eye).astype(f32)
            if "up" in a:
                sc.up = _vec(a["up"])
            if "fov" in a:
                sc.fov = float(a["fov"])
        elif node.tag == "Sphere":
            if float(a.get("radius", 0)) > 0:
                sc.add_sphere(_vec(a["center"]), float(a["radius"]), len(sc.materials) - 1,
                              int(a.get("resolution", 64)))
        elif node.tag == "Quad":
            sc.add_quad(_vec(a["corner"]), _vec(a["edge_x"]), _vec(a["edge_y"]), len(sc.materials) - 1,
                        int(a.get("resolution", 1)))
        elif node.tag == "Model":
            fn = a["filename"]
            for cand in (os.path.join(basedir, fn), os.path.join(basedir, os.path.basename(fn))):
                if os.path.exists(cand):
                    sc.add_model(cand, len(sc.materials) - 1)
                    break
            else:
                raise FileNotFoundError("Unable to find model file %s" % fn)
        elif node.tag == "Background":
            # simpleraytracer.h:125: the importance table defaults to 1024^2 when a
            # background is present; resolution="0" turns importance sampling off
            sc.background_resolution = int(a.get("resolution", 1024))
            sc.background_shader = len(sc.materials) - 1
        elif node.tag == "ShaderGroup":
            text = a.get("commands", node.text or "")
            layers, conns = parse_group_spec(text)
            name = a.get("name")
            if name and name in named:
                # a second group under a known name updates that material: its displacement or its surface
                # (simpleraytracer.cpp:472-490)
                if a.get("type", "surface") == "displacement":
                    sc.displacements[named[name]] = (layers, conns)
                else:
                    sc.materials[named[name]] = (layers, conns)
                continue
            if name:
                named[name] = len(sc.materials)
            sc.materials.append((layers, conns))
            sc.is_light.append(a.get("is_light", "no").lower() in ("yes", "true", "1", "on"))
    if not sc.materials:
        raise ValueError("No shaders in scene")
    if not sc.triangles:
        raise ValueError("No primitives in scene")
    return sc
