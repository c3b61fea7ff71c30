"""Build libstylish_b200.so in-tree with nvcc for sm_100a (no JIT cache).

    python -m stylish_tts_b200.csrc.build [--force]

The objects are compiled in parallel; the shared library links cudart
statically so it can be dlopen'ed on a box without a GPU (symbol checks).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["err.cu", "tma.cu", "spectral.cu", "backward.cu", "wgrad_umma.cu", "attention_bwd.cu", "image_ops.cu", "conv1d.cu", "conv1d_umma.cu", "convnext_fused.cu", "gan_ops.cu", "disc_ops.cu", "gemm_split.cu", "norms.cu", "attention.cu", "attention_umma.cu", "attention64.cu", "source_stft.cu", "misc.cu", "seq_ops.cu", "dropout.cu"]
LIB = os.path.join(HERE, "libstylish_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _digest(paths) -> str:
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target: str, deps) -> bool:
    """content-hash staleness (mtimes do not survive a repo snapshot): `target`.sha holds the digest of the
    flags and sources the target was built from"""
    stamp = target + ".sha"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    with open(stamp) as f:
        return f.read().strip() != _digest(deps)


def _stamp(target: str, deps) -> None:
    with open(target + ".sha", "w") as f:
        f.write(_digest(deps))


def _compile(src: str) -> str:
    obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
    deps = [os.path.join(HERE, src), os.path.join(HERE, "common.cuh"), os.path.join(HERE, "umma.cuh"), os.path.join(HERE, "tma.cuh"),
            os.path.join(HERE, "..", "..", "include", "stylish_b200.h")]
    if _stale(obj, deps):
        cmd = [NVCC, *FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        _stamp(obj, deps)
    return obj


def build(force: bool = False) -> str:
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    if force:
        for f in os.listdir(os.path.join(HERE, "build")):
            os.remove(os.path.join(HERE, "build", f))
        if os.path.exists(LIB + ".sha"):
            os.remove(LIB + ".sha")
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(_compile, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-cudart", "static", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        _stamp(LIB, objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
