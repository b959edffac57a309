"""Build libx2gnn.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m x2gnn_b200.build            # or: python x2-gnn_b200/build.py

The .so lands in x2-gnn_b200/lib/ (git-ignored, but shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libx2gnn.so")
SOURCES = ["api.cu", "graph.cu", "basis.cu", "conv.cu", "norm.cu", "readout.cu", "optim.cu", "collate.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/x2gnn.h"]
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libx2gnn.sha256")
    # the stamp that travels WITH the binary (git-ignored like the .so): libx2gnn.sha256 is tracked, so a
    # `git checkout` can restore a stamp that no longer describes the .so on disk
    built = LIB + ".digest"
    digest = _digest()
    if (not force and os.path.exists(LIB) and os.path.exists(built) and open(built).read() == digest
            and os.path.exists(stamp) and open(stamp).read() == digest):
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out.decode()}")
        if verbose and out.strip():
            print(out.decode())
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    for path in (stamp, built):
        with open(path, "w") as f:
            f.write(digest)
    return LIB


def built_digest() -> str | None:
    """Digest of the sources the .so ON DISK was built from (None: unknown build)."""
    for path in (LIB + ".digest", os.path.join(LIBDIR, "libx2gnn.sha256")):
        if os.path.exists(path):
            return open(path).read().strip()
    return None


def sass_summary(path: str = None) -> str:
    """Counts of the Blackwell-specific SASS mnemonics per object (cuobjdump -sass): tcgen05 MMA (UTC*MMA), tensor
    memory loads / stores (LDTM / STTM), tcgen05 commit barriers (UTCBAR), bulk copies (UBLKCP), cp.async (LDGSTS),
    legacy tensor path (HMMA: must be 0).  Written to profiles/sass_summary.txt by `python build.py --sass`."""
    import re
    cuobjdump = os.path.join(os.path.dirname(_nvcc()), "cuobjdump")
    pats = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "HMMA", "MUFU.EX2", "REDUX"]
    lines = [f"# SASS mnemonic counts per object of libx2gnn.so (cuobjdump -sass, sm_100a; digest {_digest()[:16]})",
             "# object      " + "  ".join(f"{p:>8s}" for p in pats)]
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        if not os.path.exists(obj):
            continue
        out = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True).stdout
        cnt = []
        for pat in pats:
            if pat == "HMMA":       # the legacy mma.sync path, not the UTCHMMA substring
                cnt.append(len(re.findall(r"(?<!UTC)HMMA", out)))
            else:
                cnt.append(out.count(pat))
        lines.append(f"{os.path.basename(obj):13s} " + "  ".join(f"{c:8d}" for c in cnt))
    text = "\n".join(lines) + "\n"
    if path:
        with open(path, "w") as f:
            f.write(text)
    return text


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    if "--sass" in sys.argv:
        root = os.path.dirname(HERE) if "HERE" in globals() else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        print(sass_summary(os.path.join(root, "profiles", "sass_summary.txt")))
