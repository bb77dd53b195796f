#!/usr/bin/env python3
"""Extract the golden vectors the reference's own tests hold for the MSM hot path.

Run in the authoring container (reads /root/reference, which does NOT exist on the GPU box):

    python tests/golden/extract_golden.py

and commit the resulting tests/golden/*.json.  For every mocha `it("name", ...)` block of the
listed test files it collects each `let|const NAME = <literal>;` whose right-hand side is a
(nested) array / number / BigInt literal, and stores integers as hex strings.  No reference
*code* is copied -- only the literal test vectors.
"""
import ast, json, os, re, sys

REF = os.environ.get("B200MSM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = {
    "batchAffine": "wasmcurves/test/batchAffine.js",
    "glv": "wasmcurves/test/glv.js",
    "utility": "wasmcurves/test/utility.js",
}


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def js_literal(txt):
    t = txt.strip()
    t = re.sub(r"\b(0x[0-9a-fA-F]+|\d+)n\b", r"\1", t)      # BigInt suffix
    t = re.sub(r",\s*([\]\)])", r"\1", t)                   # trailing commas
    try:
        v = ast.literal_eval(t)
    except Exception:
        return None
    return v


def hexify(v):
    if isinstance(v, bool): return v
    if isinstance(v, int): return hex(v)
    if isinstance(v, (list, tuple)): return [hexify(x) for x in v]
    return v


def blocks(src):
    """yield (name, body, line) for each it("...") block, splitting on the next it( / end."""
    its = [(m.start(), m.group(1)) for m in re.finditer(r'\bit\(\s*"([^"]+)"', src)]
    for k, (pos, name) in enumerate(its):
        end = its[k + 1][0] if k + 1 < len(its) else len(src)
        yield name, src[pos:end], src.count("\n", 0, pos) + 1


def extract(path):
    raw = open(path).read()
    out = {}
    for name, body, line in blocks(raw):
        # skip blocks that are entirely commented out
        first = raw.rfind("\n", 0, raw.find(body)) + 1
        if raw[first:raw.find(body)].strip().startswith("//"):
            continue
        b = strip_comments(body)
        vals = {}
        for m in re.finditer(r"\b(?:let|const)\s+(\w+)\s*=\s*", b):
            start = m.end()
            ch = b[start:start + 1]
            if ch == "[":
                depth = 0; i = start
                while i < len(b):
                    if b[i] == "[": depth += 1
                    elif b[i] == "]":
                        depth -= 1
                        if depth == 0: break
                    i += 1
                lit = b[start:i + 1]
            else:
                j = b.find(";", start); lit = b[start:j]
                if not re.fullmatch(r"\s*(0x[0-9a-fA-F]+n?|\d+n?)\s*", lit): continue
            v = js_literal(lit)
            if v is None: continue
            if m.group(1) in vals: continue
            vals[m.group(1)] = hexify(v)
        if vals:
            out[name] = {"line": line, "values": vals}
    return out


def main():
    for key, rel in FILES.items():
        p = os.path.join(REF, rel)
        data = {"source": rel, "tests": extract(p)}
        dst = os.path.join(HERE, key + ".json")
        json.dump(data, open(dst, "w"), indent=1, sort_keys=True)
        print(dst, len(data["tests"]), "tests")


if __name__ == "__main__":
    main()
