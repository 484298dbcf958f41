"""Build an experimental variant of libmt_b200.so next to the shipped one (for same-box A/B runs and debugging builds).

    python tools/build_variant.py NAME SRC.cu[,SRC2.cu] -DMACRO=1 [-DOTHER=2 ...]

recompiles only the listed sources with the extra defines, links them with the shipped objects of the other sources and
writes metatransformer_b200/lib/variants/libmt_b200_NAME.so; select it with MT_B200_LIB=<path>.
"""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from metatransformer_b200 import build as B  # noqa: E402


def main():
    name, srcs, defines = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
    B.build()
    out = B.LIBDIR / "variants"
    out.mkdir(exist_ok=True)
    objs = []
    for src in B.SOURCES:
        if src in srcs:
            obj = out / f"{src[:-3]}_{name}.o"
            r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *defines, "-c", str(B.CSRC / src), "-o", str(obj)],
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit(r.stderr)
            print("\n".join(l for l in r.stderr.splitlines() if "spill" in l or "Used" in l), file=sys.stderr)
            objs.append(str(obj))
        else:
            objs.append(str(B.LIBDIR / "obj" / (src[:-3] + ".o")))
    lib = out / f"libmt_b200_{name}.so"
    r = subprocess.run([B._nvcc(), "-shared", "-o", str(lib), *objs, "-cudart", "shared", "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stderr)
    print(lib)


if __name__ == "__main__":
    main()
