"""Builds tests/cuemu/_build/libbendy2d_b200_emu.so: the product's OWN sources (bendy2d_b200/csrc/solver.cu,
kernels.cuh, plan.cpp) compiled with g++ against the lock-step CUDA emulation in this directory.

TEST INFRASTRUCTURE.  The product never loads this library (bendy2d_b200/_lib.py only knows
bendy2d_b200/lib/libbendy2d_b200.so); tests/test_emu_parity.py runs the GPU parity tests against it in a
subprocess, so that kernel logic can be checked in a container without a GPU.

The sources are used as they are, except for four purely syntactic rewrites that g++ needs:
  1. `kernel<<<grid, block, smem, stream>>>(args)`  ->  `cuemu::Launcher("kernel", grid, block, smem, stream).run(kernel, args)`
  2. `extern __shared__ T name[];`                  ->  `T *name = (T *)cuemu::dyn_smem();`
  3. `__shared__ T a, b[N];`                        ->  references into per-CTA storage (poisoned at CTA start)
  4. `asm volatile("griddepcontrol...")`            ->  nothing (programmatic dependent launch is a scheduling hint)
(The kernels no longer contain a software grid barrier; a busy-wait would have to yield to the fiber scheduler
through BENDY_SPIN_HOOK: a cooperative scheduler never pre-empts a spinning thread.)
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "bendy2d_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libbendy2d_b200_emu.so")
NCCL_STUB = os.path.join(OUT, "libcuemu_nccl.so")

_LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)\s*<<<(.+?)>>>\s*\(", re.S)
_EXTERN_SHARED = re.compile(r"extern\s+__shared__\s+([\w ]+?)\s+(\w+)\s*\[\s*\]\s*;")
_SHARED = re.compile(r"^(\s*)__shared__\s+((?:unsigned\s+|volatile\s+)*\w+)\s+([^;\n]+);", re.M)
_ASM = re.compile(r'asm\s+volatile\s*\(\s*"griddepcontrol[^"]*"[^;]*;')


def transform(text: str, name: str) -> str:
    counter = [0]

    def launch(m):
        return f'cuemu::Launcher("{m.group(1)}", {m.group(2)}).run({m.group(1)}, '

    def extern_shared(m):
        t, n = m.group(1), m.group(2)
        return f"{t} *{n} = reinterpret_cast<{t} *>(cuemu::dyn_smem());"

    def shared(m):
        indent, t, decls = m.group(1), m.group(2), m.group(3)
        out = []
        for d in decls.split(","):
            d = d.strip()
            mm = re.fullmatch(r"(\w+)\s*((?:\[[^\]]*\])*)", d)
            if not mm:
                raise SystemExit(f"cuemu/build.py: cannot parse __shared__ declarator {d!r} in {name}")
            var, dims = mm.group(1), mm.group(2)
            counter[0] += 1
            uid = f"{1 if name == 'kernels.cuh' else 2}{counter[0]:03d}"
            out.append(f"typedef {t} _cuemu_t_{var}_{uid}{dims}; _cuemu_t_{var}_{uid} &{var} = "
                       f"*reinterpret_cast<_cuemu_t_{var}_{uid} *>(cuemu::static_smem({uid}, sizeof(_cuemu_t_{var}_{uid})));")
        return indent + " ".join(out)

    text = _EXTERN_SHARED.sub(extern_shared, text)
    text = _SHARED.sub(shared, text)
    text, n_launch = _LAUNCH.subn(launch, text)
    text = _ASM.sub(";", text)

    if any("<<<" in ln and not ln.lstrip().startswith("//") for ln in text.splitlines()):
        raise SystemExit(f"cuemu/build.py: an unconverted <<< >>> launch is left in {name}")
    if "__shared__" in text.replace("// ", ""):
        left = [ln for ln in text.splitlines() if "__shared__" in ln and not ln.lstrip().startswith("//")]
        if left:
            raise SystemExit(f"cuemu/build.py: unconverted __shared__ in {name}: {left[0].strip()}")
    return text


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in ("solver.cu", "kernels.cuh", "plan.cpp", "plan.h")]
    deps = srcs + [os.path.join(HERE, "runtime.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"),
                   os.path.join(HERE, "include", "nccl.h"), os.path.join(HERE, "nccl_stub.cpp"), os.path.abspath(__file__),
                   os.path.join(ROOT, "include", "bendy2d_b200.h")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(map(os.path.getmtime, deps)):
        return LIB
    gen = os.path.join(OUT, "bendy2d_b200", "csrc")  # same relative depth: solver.cu includes ../../include/...
    os.makedirs(gen, exist_ok=True)
    inc = os.path.join(OUT, "include")
    os.makedirs(inc, exist_ok=True)
    with open(os.path.join(ROOT, "include", "bendy2d_b200.h")) as f:
        open(os.path.join(inc, "bendy2d_b200.h"), "w").write(f.read())
    for f in ("solver.cu", "kernels.cuh", "plan.h", "plan.cpp"):
        with open(os.path.join(CSRC, f)) as fh:
            text = fh.read()
        if f in ("solver.cu", "kernels.cuh"):
            text = transform(text, f)
        open(os.path.join(gen, f if f != "solver.cu" else "solver_emu.cpp"), "w").write(text)
    cxx = os.environ.get("CXX", "g++")
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-fno-fast-math",
             "-fno-strict-aliasing", "-DBENDY_SPIN_HOOK()=cuemu::yield_spin()", "-DBENDY_SPIN_LIMIT=150000000ll", "-I", os.path.join(HERE, "include")]
    cmd = [cxx] + flags + ["-o", LIB, os.path.join(gen, "solver_emu.cpp"), os.path.join(gen, "plan.cpp"),
                           os.path.join(HERE, "runtime.cpp"), "-ldl"]
    subprocess.check_call(cmd)
    # the NCCL entry points over FIFOs, for the multi-process strip tests (solver.cu finds it through BENDY_NCCL_LIB)
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-w", "-o", NCCL_STUB,
                           os.path.join(HERE, "nccl_stub.cpp"), "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="-B" in sys.argv))
