"""Source hash / in-tree build of ``libnngp_b200.so``.

The library embeds the hash of the sources it was compiled from (``nngp_build_id()``); ``_lib.load()`` and
``__graft_entry__.build()`` compare it with the hash of the tree they run in and rebuild on a mismatch, so a
prebuilt ``.so`` that travelled to the GPU box can never silently be a binary of some other tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parents[1]          # nngp-src_b200/
ROOT = PKG.parent
LIB = PKG / "libnngp_b200.so"


def source_files():
    files = sorted((PKG / "csrc").glob("*")) + sorted((ROOT / "include").glob("*.h")) + [PKG / "build.sh"]
    return [f for f in files if f.is_file() and f.suffix in (".cu", ".cuh", ".cc", ".h", ".sh")]


def source_hash() -> str:
    h = hashlib.sha256()
    for f in source_files():
        h.update(f.relative_to(ROOT).as_posix().encode())
        h.update(b"\0")
        h.update(f.read_bytes())
        h.update(b"\0")
    return h.hexdigest()[:16]


def nvcc_path():
    cand = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return cand if Path(cand).exists() else shutil.which("nvcc")


def library_id(path=None):
    """Build id embedded in a library file (read without loading it), or None."""
    try:
        data = Path(path or LIB).read_bytes()
    except OSError:
        return None
    i = data.find(b"nngp-build-id:")
    return data[i + 14:i + 30].decode("ascii", "replace") if i >= 0 else None


def build(extra_flags=()) -> str:
    """Compile for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ...). Returns the build id.
    Serialised across processes (8 torchrun ranks loading a stale library must not run 8 compilers into one file):
    whoever gets the lock second finds the library already up to date."""
    if nvcc_path() is None:
        raise ImportError("nngp_b200: libnngp_b200.so must be (re)built but nvcc is not available")
    import fcntl
    with open(PKG / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if extra_flags or library_id() != source_hash():
                subprocess.run(["bash", str(PKG / "build.sh"), *extra_flags], check=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return source_hash()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--hash":
        print(source_hash())
    else:
        print(build(sys.argv[1:]))
