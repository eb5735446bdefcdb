"""TEST INFRASTRUCTURE ONLY.  Builds tests/emu/libpgpu_emu.so: the product's CUDA sources (pyrodigal_b200/csrc/*.cu,
unmodified apart from the mechanical rewrite of `kernel<<<grid, block, smem, stream>>>(args);` into a call of the fiber
engine) compiled with g++ against tests/emu/cuda_emu/cuda_runtime.h, so that the real kernels can be run -- slowly, on
small inputs -- against the oracle in a container without a GPU.  The product package never loads this library."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pyrodigal_b200", "csrc")
OUT = os.path.join(HERE, "_emu_build")
SANITIZE = os.environ.get("PGPU_EMU_SANITIZE") == "1"   # AddressSanitizer build (run python under LD_PRELOAD=libasan)
LIB = os.path.join(HERE, "libpgpu_emu_asan.so" if SANITIZE else "libpgpu_emu.so")
if SANITIZE:
    OUT += "_asan"
SOURCES = ["api.cu", "seq_kernels.cu", "score_kernels.cu", "dp_kernels.cu", "train_kernels.cu"]


def _match(text, i, open_ch, close_ch):
    """index just past the bracket that closes the one at text[i]"""
    depth = 0
    while True:
        c = text[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def rewrite_launches(src):
    out, pos = [], 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            return "".join(out)
        # kernel expression: identifier, optionally followed by <template arguments>
        j = k
        if src[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        m = re.search(r"[A-Za-z_][A-Za-z_0-9:]*$", src[:j])
        start = m.start()
        kernel = src[start:k]
        e = src.index(">>>", k)
        cfg = src[k + 3:e]
        a = e + 3
        while src[a].isspace():
            a += 1
        assert src[a] == "(", (kernel, src[a:a + 20])
        b = _match(src, a, "(", ")")
        args = src[a:b]
        c = b
        while src[c].isspace():
            c += 1
        assert src[c] == ";", (kernel, src[c:c + 20])
        out.append(src[pos:start])
        out.append(f"emu::Launch({cfg}) << [&]() {{ {kernel}{args}; }};")
        pos = c + 1


def rewrite(src):
    src = rewrite_launches(src)
    # PTX prefetch hints have no host equivalent
    src = re.sub(r'asm volatile\("prefetch[^;]*;"\s*::[^;]*\);', "(void)0;", src)
    return src


def build(force=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp"))]
    deps += [os.path.join(HERE, "cuda_emu", f) for f in os.listdir(os.path.join(HERE, "cuda_emu"))]
    deps += [os.path.abspath(__file__), os.path.join(ROOT, "include", "pyrodigal_b200.h")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    objs = []
    flags = ["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-pthread", "-w",
             "-I", os.path.join(HERE, "cuda_emu"), "-I", CSRC]
    if SANITIZE:
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer"]
    procs = []
    for f in SOURCES:
        cpp = os.path.join(OUT, f.replace(".cu", ".emu.cpp"))
        with open(os.path.join(CSRC, f)) as fh:
            text = rewrite(fh.read())
        with open(cpp, "w") as fh:
            fh.write(f'#line 1 "{os.path.join(CSRC, f)}"\n' + text)
        obj = cpp[:-4] + ".o"
        objs.append(obj)
        procs.append(subprocess.Popen(flags + ["-c", cpp, "-o", obj]))
    eng = os.path.join(OUT, "emu_engine.o")
    procs.append(subprocess.Popen(flags + ["-c", os.path.join(HERE, "cuda_emu", "emu_engine.cpp"), "-o", eng]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("emulation build failed")
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", LIB] + (["-fsanitize=address"] if SANITIZE else []) + objs + [eng])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
