#!/usr/bin/env python3
"""build_ref.py -- TEST INFRASTRUCTURE (oracle), not product code.

Recipe that turns the reference's shipped WASM binaries into native shared
libraries under oracle/_ref/ (git-ignored; travels to the GPU box with the
snapshot):

    /root/reference/wasmcurves/build/bls12381.wasm --wasm2c.py--> oracle/_ref/bls12381.c --gcc -O2--> libref_bls12381.so
    /root/reference/wasmcurves/build/bn128.wasm    --wasm2c.py--> oracle/_ref/bn128.c    --gcc -O2--> libref_bn128.so

The reference's own build system (node + wasmbuilder) cannot run here (no node,
no npm packages); the .wasm files are its build *outputs* and contain the upstream
g1m_multiexpAffine algorithm (src/build_multiexp.js).  Nothing is copied into git.

Also records the module's constant pointers (pq, pG1gen, ...) from
build/<curve>_wasm.js (tools/buildwasm_bls12381.js:14-31) in oracle/_ref/<curve>_consts.txt.
Skips silently (returns False) when /root/reference is absent (GPU box: prebuilt files are used).
"""
import os, re, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFROOT = os.environ.get("B200MSM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build(force=False, verbose=True):
    src_dir = os.path.join(REFROOT, "wasmcurves", "build")
    if not os.path.isdir(src_dir):
        return False
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, HERE)
    import wasm2c
    for mod in ("bls12381", "bn128"):
        wasm = os.path.join(src_dir, mod + ".wasm")
        c_out = os.path.join(OUT, mod + ".c")
        so = os.path.join(OUT, "libref_%s.so" % mod)
        stamp = max(os.path.getmtime(wasm), os.path.getmtime(os.path.join(HERE, "wasm2c.py")))
        if not force and os.path.exists(so) and os.path.getmtime(so) >= stamp:
            continue
        code, m = wasm2c.translate(open(wasm, "rb").read(), mod)
        open(c_out, "w").write(code)
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing", "-w", "-o", so, c_out]
        if verbose: print("[oracle/_ref]", " ".join(cmd))
        subprocess.check_call(cmd)
        js = open(os.path.join(src_dir, mod + "_wasm.js")).read()
        with open(os.path.join(OUT, mod + "_consts.txt"), "w") as f:
            for k, v in re.findall(r"exports\.(\w+)\s*=\s*(\d+);", js):
                f.write("%s %s\n" % (k, v))
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "reference not present: nothing built")
