#!/usr/bin/env python
"""SASS evidence for the instruction-level claims in DESIGN.md: per kernel, the count and a few verbatim lines of the
mnemonics that matter (tensor-core MMAs, ldmatrix, vector REDs, bulk copies, multimem loads, shared-memory traffic of
the sort).  Reads the objects of the in-tree build: python tools/sass_excerpts.py > profiles/r02_sass_excerpts.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "hidegs_b200", "csrc", "build")
WANT = [  # (object, kernel regex, mnemonic regex, what it shows)
    ("blend_bwd", r"blend_bwd3_kernelILb1ELb1ELb0", r"HMMA|LDSM|REDG?\.|RED\.|MUFU", "3xTF32 mma.sync reduction, ldmatrix A quads, vector REDs"),
    ("blend_bwd", r"blend_bwd3_kernelILb0ELb0ELb1", r"HMMA|LDSM|MUFU\.(LG2|EX2)|CALL", "hierarchy interpolation on the MMA kernel: lg2 / ex2 pow, out-of-line accurate redo"),
    ("blend_fwd", r"blend_fwd2_kernelILb1ELb1ELb0", r"FFMA2|FMUL2|FADD2", "packed FP32 (fma/mul/add.f32x2): the lane's two pixels per instruction, scalar-broadcast (.F32) and pair (.F32x2.HI_LO) operands, FFMA2.RM of the paired expf"),
    ("blend_bwd", r"blend_bwd3_kernelILb1ELb1ELb0", r"FFMA2|FMUL2|FADD2", "packed FP32 in the backward: pair recurrence, g = <features, dL/dpixel> as 9 packed FMAs, TF32 remainders of the flush"),
    ("blend_fwd", r"blend_fwd2_kernelILb1ELb1ELb0", r"LDG\.E\.128|STS\.128|LDS\.128|MUFU|VOTE|ATOM|RED", "register-double-buffered gather, 128-bit staging, ballots"),
    ("blend_fwd", r"blend_fwd3_kernel", r"UBLKCP|SYNCS", "TMA bulk-copy staging experiment (cp.async.bulk + mbarrier; removed from the tree)"),
    ("exchange", r"nvls_allreduce_kernelILi4", r"LDGMC|STG.*MC|MULTIMEM|ST\.E.*MMIO|RED|ATOM|CAS|LDG\.E\.128", "multimem.ld_reduce / multimem.st through NVSwitch; the gather range = plain LDG.128 + multimem.st"),
    ("preprocess_bwd", r"sh_from_factors", r"STS|LDS|STG\.E\.128|LDG", "SH rows rebuilt from the factors: rows leave through the shared-memory tile as STG.128"),
    ("geometry", r"prologue_bwd_kernel", r"MUFU|LDG\.E\.128|STG\.E\.128|EXIT", "fused prologue backward: early exit on radii, one pass"),
    ("binning", r"tile_sort_small", r"MATCH|ATOMS|ATOMG|LDS|STS|SHFL|REDUX|LDG|STG", "per-list radix sort: shared-memory loads/stores only in the element loops"),
    ("binning", r"scatter_instances", r"ATOMG|STG|SHFL", "scatter: one ATOMG + one STG.64 per instance"),
    ("preprocess", r"preprocess_fwd_kernel", r"REDG?\.|RED\.|LDG\.E\.128|STG\.E\.128", "tile counters by RED, 128-bit record stores"),
]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj + ".o")], capture_output=True, text=True).stdout
    cur, body = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            body[cur].append(line.rstrip())
    return body


def main():
    print("# SASS excerpts (cuobjdump -sass of the in-tree sm_100a build)\n")
    cache = {}
    for obj, krx, mrx, what in WANT:
        if obj not in cache:
            cache[obj] = functions(obj)
        hits = [k for k in cache[obj] if re.search(krx, k)]
        if not hits:
            print("## %s / %s — not in this build\n" % (obj, krx))
            continue
        k = hits[0]
        lines = cache[obj][k]
        print("## `%s` (%s.cu) — %s\n" % (k[:90], obj, what))
        print("%d SASS instructions." % len(lines))
        cnt = collections.Counter()
        sample = {}
        for ln in lines:
            m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
            if m and re.search(mrx, m.group(1)):
                cnt[m.group(1)] += 1
                sample.setdefault(m.group(1), re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).strip())
        print("\n| mnemonic | count | first occurrence |\n|---|---:|---|")
        for mn, c in sorted(cnt.items(), key=lambda x: -x[1])[:14]:
            print("| `%s` | %d | `%s` |" % (mn, c, sample[mn][:110]))
        print()


if __name__ == "__main__":
    main()
