"""Extracts the reference's OpenCL kernel strings into oracle/_ref/*.cl.inc so the CPU shim can compile them.

Usage: python extract.py <reference_root> <out_dir>.  Outputs are build artefacts (git-ignored): the reference's
source text is never committed here.  The only edit is syntactic: OpenCL vector literals `(float8)(a, ...)`,
`(float4)(...)`, `(int2)(...)` become `make_float8(a, ...)` etc., because C++ has no such cast syntax.
"""
import os
import re
import sys


def kernel_text(path):
    s = open(path, encoding="utf-8-sig").read()
    m = re.search(r'R"CLC\((.*?)\)CLC"', s, re.S)
    if not m:
        raise SystemExit("no CLC string in " + path)
    t = m.group(1)
    for ty in ("float8", "float4", "int2"):
        t = t.replace("(%s)(" % ty, "make_%s(" % ty)
    return t


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    kdir = os.path.join(ref, "Watermark_GPU", "kernels")
    for name in ("nvf", "me_p3", "scaled_neighbors_p3"):
        with open(os.path.join(out, name + ".cl.inc"), "w") as f:
            f.write(kernel_text(os.path.join(kdir, name + ".hpp")))
    hpp = open(os.path.join(ref, "Watermark_GPU", "Watermark.hpp"), encoding="utf-8-sig").read()
    m = re.search(r"RxMappings\[64\]\s*\{(.*?)\};", hpp, re.S)
    vals = [int(v) for v in re.findall(r"\d+", m.group(1))]
    assert len(vals) == 64
    with open(os.path.join(out, "rxmappings.inc"), "w") as f:
        f.write(", ".join(map(str, vals)) + "\n")


if __name__ == "__main__":
    main()
