"""Build oracle/hqq_oracle_c.c (TEST INFRASTRUCTURE ONLY) into oracle/_build/libhqq_oracle_c.so with gcc + OpenMP.

    python -m oracle.build_c [--force]

`__graft_entry__.build()` calls `build()`; the .so is git-ignored and travels to the GPU box with the snapshot (gcc is there too:
`build()` rebuilds when the source is newer).  No FMA contraction and no fast-math: the levels depend on the rounding of W*s + z."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hqq_oracle_c.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libhqq_oracle_c.so")
FLAGS = ["-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c11", "-Wall"]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    # compilers to try, in order: the system gcc first ($CC may point at a wrapper without libgomp), then $CC / cc; OpenMP first,
    # then -- rather than no oracle at all -- a single-threaded build (hqq_oc_threads() then reports 1)
    ccs = [c for c in (shutil.which("gcc"), "/usr/bin/gcc", os.environ.get("CC"), shutil.which("cc")) if c and os.path.exists(c)]
    if not ccs:
        raise RuntimeError("oracle/build_c: no C compiler (gcc) found")
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = LIB + f".{os.getpid()}.tmp"
    log = []
    for flags in (FLAGS, [f for f in FLAGS if f != "-fopenmp"] + ["-Wno-unknown-pragmas"]):
        for cc in dict.fromkeys(ccs):
            r = subprocess.run([cc, *flags, SRC, "-o", tmp, "-lm"], capture_output=True, text=True)
            if r.returncode == 0:
                break
            log.append(f"{cc} {' '.join(flags)}:\n{r.stderr[-600:]}")
        else:
            continue
        break
    else:
        raise RuntimeError("oracle/build_c: compilation failed:\n" + "\n".join(log))
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
