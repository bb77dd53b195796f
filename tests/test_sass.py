"""CPU check of the shipped library's code generation (no GPU needed): the instruction mix of the hot kernels, read from the SASS with
cuobjdump -- the numbers DESIGN.md section 6 derives the instruction bound of k_tree_bwd from (5 field multiplications x 300 limb products
= 1500 IMAD.WIDE per addition) and profiles/r2_sass_counts.txt records.  Guards against a build whose code generation drifted (a kernel
that lost its register budget, spills in the addition loop, a multiplier that no longer fuses into IMAD.WIDE)."""
import os, re, shutil, subprocess, collections
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "zprize-wasm-msm_b200", "b200msm", "libb200msm.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)), reason="needs the built library and cuobjdump")


def _usage():
    res = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    return {m.group(1): m.group(2) for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res)}


def _mix(fun):
    sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", fun, LIB], capture_output=True, text=True, check=True).stdout
    ops = collections.Counter()
    for ln in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m: ops[m.group(1)] += 1
    return ops


def _find(usage, *parts):
    hits = [f for f in usage if all(p in f for p in parts)]
    assert hits, parts
    return hits


def test_tree_kernels_live_in_one_object_each():
    u = _usage()
    for curve in ("9BLS12_381", "5BN254"):
        for first in ("Lb0E", "Lb1E"):
            assert len(_find(u, "10k_tree_bwdINS_" + curve, first)) == 1      # instantiated once (tree.cu), not once per translation unit
            assert len(_find(u, "10k_tree_fwdINS_" + curve, first)) == 1


@pytest.mark.parametrize("first", ["Lb0E", "Lb1E"])
def test_backward_pass_instruction_mix_bls12381(first):
    u = _usage(); f = _find(u, "10k_tree_bwdINS_9BLS12_381", first)[0]
    regs = int(re.search(r"REG:(\d+)", u[f]).group(1))
    assert regs <= 128                                      # 4 CTAs of 128 threads per SM
    ops = _mix(f)
    wide = sum(v for k, v in ops.items() if k.startswith("IMAD.WIDE"))
    assert 1500 <= wide <= 1560, wide                       # 5 multiplications x (2 * 12^2 + 12) limb products (one of them a squaring, the rest full products) + address arithmetic
    assert sum(v for k, v in ops.items() if k.startswith(("LDL", "STL"))) <= 24      # a handful of spilled words outside the multiplier chains, not a spilling loop
    assert sum(ops.values()) <= 3400


def test_backward_pass_bn254_fits_five_ctas():
    u = _usage()
    for first in ("Lb0E", "Lb1E"):
        f = _find(u, "10k_tree_bwdINS_5BN254", first)[0]
        assert int(re.search(r"REG:(\d+)", u[f]).group(1)) <= 102      # 65536 / (5 * 128)
        assert "STACK:0" in u[f]
