#!/usr/bin/env python3
"""Replay the reference's BATCHED testsuite through this repo (the `.b200` variant of
src/cmake/testing.cmake:214-235, which registers every testsuite directory that holds a BATCHED
marker a second time for the batched back end).

For every such directory under /root/reference/testsuite:
  1. run.py is executed with stub command builders to obtain its command lines;
  2. every .osl in the directory is compiled with tools/mini_oslc.py (no oslc exists here);
  3. each `testshade ...` command is parsed (openshadinglanguage_b200.testshade.parse_command) and
     the group is built through the PRODUCT's generator + NVRTC (no GPU needed for that);
  4. the command is replayed on the CPU oracle and its text compared with ref/out.txt, its images
     with the ref images (8-bit / float TIFFs that PIL reads) at the reference's idiff thresholds.

Results go to tests/golden/testsuite_b200/: manifest.json (status of all directories) and one
bundle per directory that passed (commands, .oso texts, expected text, reference images), which
tests/test_testsuite_b200.py replays through the GPU.  Needs /root/reference: run in the build
container, not on the GPU box.
"""
import json
import os
import re
import shlex
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"
TS = os.path.join(REF, "testsuite")
OUT = os.path.join(ROOT, "tests", "golden", "testsuite_b200")


def commands_of(d):
    """run.py with stubbed command builders -> [(tool, argument string)], test settings"""
    cmds = []

    def stub(name):
        def f(*a, **k):
            cmds.append((name, a[0] if a else ""))
            return "\x00%d\x00" % (len(cmds) - 1)     # placeholder: keeps the order inside `command`
        return f
    ns = {k: stub(k) for k in ("testshade", "oslc", "testrender", "oiiotool", "oslinfo", "maketx", "idiff",
                               "testoptix", "diff_command", "oiio_app")}
    ns.update(dict(command="", outputs=["out.txt"], failthresh=0.004, failpercent=0.02, hardfail=0.012,
                   allowfailures=0, os=os, OIIO_TESTSUITE_IMAGEDIR="", splitsymbol=";", platform=None))
    cwd = os.getcwd()
    os.chdir(os.path.join(TS, d))
    try:
        exec(open("run.py").read(), ns)
    finally:
        os.chdir(cwd)
    # run.py may interleave `echo TEXT>> out.txt` lines with the tool commands (error-dupes): they become
    # ("echo", TEXT) entries at their place in the sequence
    seq = []
    for part in re.split(r"(\x00\d+\x00)", ns["command"] if isinstance(ns.get("command"), str) else ""):
        m = re.fullmatch(r"\x00(\d+)\x00", part)
        if m:
            seq.append(cmds[int(m.group(1))])
            continue
        for line in part.split("\n"):
            e = re.match(r"\s*echo\s+(.*?)\s*>>\s*out\.txt", line)
            if e:
                seq.append(("echo", e.group(1).strip("\"'")))
    used = {id(c) for c in seq}
    seq += [c for c in cmds if id(c) not in used]        # commands run.py did not append to `command`
    return seq, {k: ns[k] for k in ("outputs", "failthresh", "failpercent", "hardfail", "allowfailures")}


def oslinfo_verbose(oso):
    """What `oslinfo -v shader` prints (src/oslinfo/oslinfo.cpp:150-260) for the parameters and
    metadata found in an .oso text (OSLQuery's view: name, type, default values, %meta hints)."""
    lines = []
    hint = re.compile(r'%meta\{(\w+),(\w+),("(?:\\.|[^"\\])*"|[^}]*)\}')

    def metas(text):
        out = []
        for t, n, v in hint.findall(text):
            out.append("\t\tmetadata: %s %s = %s" % (t, n, v))
        return out
    for ln in oso.split("\n"):
        m = re.match(r"(shader|surface|displacement|volume)\s+(\S+)(.*)", ln)
        if m and not lines:
            lines.append('%s "%s"' % (m.group(1), m.group(2)))
            lines += metas(m.group(3))
            continue
        m = re.match(r"(param|oparam)\t(\S+(?: color)?)\t(\S+)\t([^\t]*)\t(.*)", ln)
        if not m:
            continue
        kind, typ, name, vals, rest = m.groups()
        if "lockgeom" in rest and not [h for h in hint.findall(rest) if h[1] != "lockgeom"]:
            rest = ""
        lines.append('    "%s" "%s%s"' % (name, "output " if kind == "oparam" else "", typ))
        base = typ.split("[")[0]
        agg = {"color": 3, "point": 3, "vector": 3, "normal": 3, "matrix": 16}.get(base, 1)
        if base == "string":
            sv = re.findall(r'"((?:\\.|[^"\\])*)"', vals)
            if "[" in typ:
                lines.append("\t\tDefault value: [ " + " ".join('"%s"' % v for v in sv) + " ]")
            else:
                lines.append('\t\tDefault value: "%s"' % (sv[0] if sv else ""))
        else:
            toks = vals.split()
            if base != "int":
                toks = ["%g" % float(t) for t in toks]
            body = " " + " ".join(toks)
            lines.append("\t\tDefault value:" + (" [" + body + " ]" if ("[" in typ or agg > 1) else body))
        lines += [l for l in metas(rest) if " lockgeom " not in l]
    return "\n".join(lines)


def read_exr_scanlines(path):
    """Minimal OpenEXR reader for what OpenCV refuses (single-channel / oddly named channels): scan-line files,
    NONE / ZIPS / ZIP compression, half or float channels.  -> float32 [h, w, nchannels], channels in file
    (alphabetical) order."""
    import struct
    import zlib
    d = open(path, "rb").read()
    if d[:4] != b"\x76\x2f\x31\x01" or d[5] & 0x02:
        raise IOError("not a scan-line OpenEXR file: " + path)
    i, attrs = 8, {}
    while d[i] != 0:
        j = d.index(b"\0", i)
        k = d.index(b"\0", j + 1)
        size = struct.unpack_from("<i", d, k + 1)[0]
        attrs[d[i:j].decode()] = d[k + 5:k + 5 + size]
        i = k + 5 + size
    i += 1
    chans, c = [], attrs["channels"]
    p = 0
    while c[p] != 0:
        q = c.index(b"\0", p)
        chans.append((c[p:q].decode(), struct.unpack_from("<i", c, q + 1)[0]))     # name, 1 = half, 2 = float
        p = q + 17
    comp = attrs["compression"][0]
    if comp not in (0, 2, 3) or any(t not in (1, 2) for _, t in chans):
        raise IOError("unsupported OpenEXR flavour: " + path)
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    per = {0: 1, 2: 1, 3: 16}[comp]
    nblocks = (h + per - 1) // per
    offs = struct.unpack_from("<%dQ" % nblocks, d, i)
    img = np.zeros((h, w, len(chans)), np.float32)
    line_bytes = sum(w * (2 if t == 1 else 4) for _, t in chans)
    for o in offs:
        y, size = struct.unpack_from("<ii", d, o)
        raw = d[o + 8:o + 8 + size]
        nl = min(per, y1 - y + 1)
        if comp and size < nl * line_bytes:
            b = np.frombuffer(zlib.decompress(raw), np.uint8).astype(np.int32)
            b = np.cumsum(np.concatenate([b[:1], b[1:] - 128]), dtype=np.int64).astype(np.uint8)   # predictor
            half = (len(b) + 1) // 2
            out = np.empty(len(b), np.uint8)
            out[0::2], out[1::2] = b[:half], b[half:]                                              # de-interleave
            raw = out.tobytes()
        p = 0
        for ln in range(nl):
            for ci, (_, t) in enumerate(chans):
                n = w * (2 if t == 1 else 4)
                img[y - y0 + ln, :, ci] = np.frombuffer(raw[p:p + n], np.float16 if t == 1 else np.float32)
                p += n
    return img


def read_image(path):
    if path.endswith(".exr"):
        os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            return read_exr_scanlines(path), "float"
        if img.ndim == 3:
            img = img[..., ::-1] if img.shape[2] == 3 else img[..., [2, 1, 0, 3]]
        else:
            img = img[..., None]
        return np.ascontiguousarray(img).astype(np.float32), "float"
    from PIL import Image
    img = np.array(Image.open(path))
    if img.ndim == 2:
        img = img[..., None]
    if img.dtype == np.uint8:
        return img.astype(np.float32) / 255.0, "uint8"
    if img.dtype == np.uint16:
        return img.astype(np.float32) / 65535.0, "uint16"
    return img.astype(np.float32), "float"


def compare_image(got, ref, kind, st):
    from openshadinglanguage_b200 import testshade as tsh
    return tsh.compare_image(got, ref, kind, st["failthresh"], st["failpercent"], st["hardfail"])


def main():
    from tools import mini_oslc
    from oracle import oracle
    import helpers
    import openshadinglanguage_b200 as ob
    from openshadinglanguage_b200 import testshade as tsh
    os.makedirs(OUT, exist_ok=True)
    inc = [os.path.join(REF, "src/shaders")]
    only = set(sys.argv[1:])
    dirs = [d for d in sorted(os.listdir(TS)) if os.path.exists(os.path.join(TS, d, "BATCHED"))]
    manifest = {}
    for d in dirs:
        if only and d not in only:
            continue
        entry = dict(status="?", reason="")
        manifest[d] = entry
        try:
            cmds, st = commands_of(d)
        except Exception as e:
            entry.update(status="harness", reason="run.py: %s" % e)
            continue
        tools_used = sorted({t for t, _ in cmds})
        # `oslinfo -v shader` between the commands: emulated from the compiled .oso further down
        info_ok = all(re.fullmatch(r"-v\s+\S+", a.strip()) for t, a in cmds if t == "oslinfo")
        shade = [a if t == "testshade" else ("\x00oslinfo " + a.split()[-1] if t == "oslinfo" else "\x00echo " + a)
                 for t, a in cmds if t in ("testshade", "echo", "oslinfo")]
        if not any(t == "testshade" for t, _ in cmds) or not info_ok \
                or set(tools_used) - {"testshade", "oslc", "echo", "oslinfo"}:
            entry.update(status="harness", reason="uses " + ",".join(tools_used))
            continue
        def inline_group_files(a):
            def sub(m):
                path = os.path.join(TS, d, m.group(2))
                if not os.path.exists(path):
                    path = os.path.join(TS, d, os.path.basename(m.group(2)))     # run.py may name a copy under data/
                return m.group(1) + " " + shlex.quote(open(path).read()) if os.path.exists(path) else m.group(0)
            return re.sub(r"(--?group)\s+(\S+\.oslgroup)", sub, a)
        shade = [inline_group_files(a) for a in shade]
        entry["commands"] = shade
        # 2. compile the directory's shaders
        oso = {}
        failed = None
        for f in sorted(os.listdir(os.path.join(TS, d))):
            if f.endswith(".osl"):
                try:
                    text = mini_oslc.compile_osl(os.path.join(TS, d, f), inc + [os.path.join(TS, d)],
                                                 metadata="oslinfo" in tools_used)
                    # oslc names the .oso after the SHADER, not the source file (compassign-bool's
                    # varying_le.osl defines shader varying_lt and the other way round)
                    m = re.search(r"^(?:shader|surface|displacement|volume)\s+(\S+)", text, re.M)
                    oso[m.group(1) if m else f[:-4]] = text
                except Exception as e:
                    failed = failed or "%s: %s" % (f, str(e).split("\n")[0][:200])
        specs = []
        try:
            # oslinfo output is fixed once the shaders are compiled: from here on it is an echo line
            shade = ["\x00echo " + oslinfo_verbose(oso[a[9:]]) if a.startswith("\x00oslinfo ") and a[9:] in oso else a
                     for a in shade]
            if any(a.startswith("\x00oslinfo ") for a in shade):
                raise ValueError("oslinfo of a shader that did not compile")
            entry["commands"] = shade
            specs = [dict(echo=a[6:], unsupported=[], layers=[]) if a.startswith("\x00echo ") else tsh.parse_command(a)
                     for a in shade]
        except Exception as e:
            entry.update(status="harness", reason="command line: %s" % e)
            continue
        unsup = sorted({u for s in specs for u in s["unsupported"]})
        if unsup:
            entry.update(status="harness", reason="testshade flags not replayed: " + " ".join(unsup))
            continue
        need = {l["shader"] for s in specs for l in s["layers"]}
        # shaders shared between tests live in testsuite/common/shaders (runtest.py compiles them too)
        common = os.path.join(TS, "common", "shaders")
        for name in sorted(need - set(oso)):
            src = os.path.join(common, name + ".osl")
            if os.path.exists(src):
                try:
                    oso[name] = mini_oslc.compile_osl(src, inc + [common])
                except Exception as e:
                    failed = failed or "%s: %s" % (name, str(e).split("\n")[0][:200])
        if failed and (need - set(oso)):
            entry.update(status="oslc", reason=failed)
            continue
        missing = need - set(oso)
        if missing:
            entry.update(status="harness", reason="shader(s) not in the directory: " + ",".join(sorted(missing)))
            continue
        # 3. product generator + NVRTC
        def product_group(layers, conns, outs, spec):
            arena, descs = ob.pack_userdata(helpers.testshade_userdata(
                spec["xres"] * spec["yres"], *ob.grid_globals(spec["xres"], spec["yres"]), extra=spec["userdata"]))
            return ob.ShaderGroup(layers, conns, outs, options="fma=0,journal=1" + (
                ",error_repeats=1" if "error_repeats=1" in spec["options"] else "") + (
                ",colorspace=" + spec["colorspace"] if spec.get("colorspace") else ""), userdata=descs,
                                  name=spec.get("groupname") or "group",
                                  attributes=tsh.harness_attributes(spec["xres"], spec["yres"]))
        try:
            for s in specs:
                if "echo" in s:
                    continue
                layers = [dict(oso=oso[l["shader"]], name=l["name"], params=l["params"]) for l in s["layers"]]
                g = product_group(layers, s["connections"], (), s)
                assert g.cubin[:4] == b"\x7fELF"
        except Exception as e:
            entry.update(status="codegen", reason=str(e).split("\n")[0][:300])
            continue
        # 4. oracle replay vs the reference's golden output
        def oracle_globals(xres, yres, raytype_bit=1, **kw):
            return oracle.testshade_globals(xres, yres, raytype=raytype_bit, **kw)

        class OracleRunner:
            def __init__(self, layers, conns, outs, spec):
                self.g = oracle.OracleGroup(layers, conns, outs, name=spec.get("groupname") or "group", flags=(
                    ('-DOSLO_COLORSPACE="%s"' % spec["colorspace"],) if spec.get("colorspace") else ()),
                    attributes=tsh.harness_attributes(spec["xres"], spec["yres"]))
                self.spec = spec

            def run(self, n, var, uni, arena):
                uni = dict(uni)
                uni["userdata"] = helpers.testshade_userdata(n, var, uni, extra=self.spec["userdata"])
                return self.g.run_capture(n, var, uni, arena, error_repeats="error_repeats=1" in self.spec["options"])
        try:
            texts, images = [], {}
            for s in specs:
                if "echo" in s:
                    texts.append(s["echo"])
                    continue
                r = tsh.run_command(s, lambda name: oso[name], OracleRunner, oracle_globals)
                if r["text"]:
                    texts.append(r["text"])
                for var, (fn, img) in r["images"].items():
                    if fn != "null":
                        images[fn] = img
            got = "\n".join(texts)
        except Exception as e:
            entry.update(status="oracle", reason=(str(e).split("\n")[0] or traceback.format_exc().split("\n")[-2])[:300])
            continue
        want_path = os.path.join(TS, d, "ref", "out.txt")
        want = ""
        if os.path.exists(want_path):
            # not the shader's output: oslc's "Compiled ..." lines and a debug print of the reference's
            # batched code generator ("x is forced llvm bool.")
            want = "\n".join(l for l in open(want_path).read().split("\n")
                             if not l.startswith("Compiled") and not l.endswith("is forced llvm bool."))
        problems = []
        if "out.txt" in st["outputs"] and os.path.exists(want_path) and got.rstrip("\n") != want.rstrip("\n"):
            gl, wl = got.rstrip("\n").split("\n"), want.rstrip("\n").split("\n")
            k = next((i for i in range(min(len(gl), len(wl))) if gl[i] != wl[i]), min(len(gl), len(wl)))
            problems.append("text differs at line %d: got %r want %r" % (
                k + 1, gl[k][:80] if k < len(gl) else "<end>", wl[k][:80] if k < len(wl) else "<end>"))
        ref_images = {}
        for fn in st["outputs"]:
            if fn == "out.txt" or fn == "null":
                continue
            rp = os.path.join(TS, d, "ref", fn)
            if not os.path.exists(rp):
                continue
            if fn not in images:
                problems.append("output image %s not produced" % fn)
                continue
            try:
                ref, kind = read_image(rp)
            except Exception as e:
                problems.append("reference image %s unreadable here (%s)" % (fn, type(e).__name__))
                continue
            p = compare_image(images[fn], ref, kind, st)
            if p:
                problems.append("%s: %s" % (fn, p))
            ref_images[fn] = (ref, kind)
        if problems:
            entry.update(status="mismatch", reason="; ".join(problems)[:400])
            continue
        entry.update(status="pass", reason="")
        bundle = dict(commands=shade, oso=oso, want=want, settings=st,
                      images={fn: dict(kind=kind, shape=list(ref.shape)) for fn, (ref, kind) in ref_images.items()})
        with open(os.path.join(OUT, d + ".json"), "w") as f:
            json.dump(bundle, f)
        if ref_images:
            np.savez_compressed(os.path.join(OUT, d + ".npz"),
                                **{re.sub(r"\W", "_", fn): (ref * (255 if kind == "uint8" else 1)).astype(
                                    np.uint8 if kind == "uint8" else np.float32) for fn, (ref, kind) in ref_images.items()})
        print("%-32s %s %s" % (d, entry["status"], entry["reason"]), flush=True)
    for d, e in manifest.items():
        if e["status"] != "pass":
            print("%-32s %s %s" % (d, e["status"], e["reason"]), flush=True)
    if not only:
        with open(os.path.join(OUT, "manifest.json"), "w") as f:
            json.dump({d: dict(status=e["status"], reason=e["reason"]) for d, e in manifest.items()}, f, indent=1)
    import collections
    c = collections.Counter(e["status"] for e in manifest.values())
    print("BATCHED testsuite directories: %d; %s" % (len(manifest), dict(c)))


if __name__ == "__main__":
    main()
