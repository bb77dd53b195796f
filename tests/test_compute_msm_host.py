"""The compiled C++ host (tools/compute_msm.cpp): builds against include/b200msm.h + libb200msm.so with g++ alone, fails loudly without a GPU
(CPU test), and on a B200 its `compute_msm(bases, scalars)` output equals the oracle's on the same seeded inputs (GPU test)."""
import os, subprocess, sys
import pytest

HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "zprize-wasm-msm_b200", "b200msm")
EXE = os.path.join(HERE, "_compute_msm")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


@pytest.fixture(scope="module")
def exe():
    src = os.path.join(ROOT, "tools", "compute_msm.cpp")
    deps = [src, os.path.join(ROOT, "include", "b200msm.h"), os.path.join(LIBDIR, "libb200msm.so")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", src, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
                               "-L", LIBDIR, "-lb200msm", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + os.path.join(CUDA, "lib64"), "-o", EXE])
    return EXE


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_host_builds_with_gcc_and_fails_loudly_without_a_gpu(exe):
    assert os.path.exists(exe)
    if _has_gpu(): pytest.skip("a GPU is present")
    p = subprocess.run([exe, "bls12381", "10", "7"], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr and p.stdout == ""


def _splitmix64(x):
    M = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & M
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
    return x ^ (x >> 31)


def _expected(cname, lg, seed):
    import pyref, coracle
    cv = pyref.CURVES[cname]; n = 1 << lg
    bases = coracle.generate_bases(cv.cid, pyref.affine_to_bytes(cv, cv.G), seed, 0, n)
    sc = b"".join(_splitmix64((seed ^ 0x5ca1ab1e) + i).to_bytes(8, "little") for i in range(4 * n))
    xy = coracle.normalize(cv.cid, coracle.multiexp_affine(cv.cid, bases, sc, 32, n))
    return "x=%s y=%s" % (xy[: cv.n8][::-1].hex(), xy[cv.n8:][::-1].hex())


@pytest.mark.gpu
@pytest.mark.parametrize("cname,lg,extra", [("bls12381", 12, []), ("bn128", 12, []), ("bls12381", 14, ["--bases-on-host"]), ("bls12381", 13, ["--devices", "0,0"])])
def test_compute_msm_matches_oracle(exe, cname, lg, extra):
    seed = 0xB2000000 + lg
    p = subprocess.run([exe, cname, str(lg), str(seed)] + extra, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert p.stdout.splitlines()[0] == _expected(cname, lg, seed)
