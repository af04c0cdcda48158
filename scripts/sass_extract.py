#!/usr/bin/env python
"""SASS summary of the hot kernels of lib/liblbm3d_b200.so for profiles/: per kernel the register
count and shared memory (cuobjdump -res-usage), the number of SASS instructions, and how many of
them are bulk-TMA copies (UBLKCP), mbarrier operations (SYNCS), global loads / stores, shuffles and
block barriers.  sm_100a only; no UTMALDG / UTCMMA is expected: the step is not a contraction and
the table slices are 1-D."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "taichi_lbm3d_b200", "lib", "liblbm3d_b200.so")
HOT = ["lbm_fast::k_dense<0, 0, true>", "lbm_fast::k_dense<1, 0, true>", "lbm_fast::k_dense_aa<0, 0, 1>",
       "lbm_fast::k_dense_aa<0, 0, 2>", "lbm_fast::k_sparse<1, 0, true, 0>", "lbm_fast::k_sparse<0, 0, true, 0>",
       "lbm_fast::k_sparse<1, 0, true, 1>", "lbm_fast::k_sparse<1, 0, true, 2>",
       "lbm_fast::k_dense_peer<0, true>", "lbm_fast::k_dense_grey<1, 0>",
       "lbm2p_fast::k2p_main<true, 0, true>", "lbm2p_fast::k2p_colour<false>", "lbm2p_fast::k2p_colour<true>",
       "lbm2p_fast::k2p_main_sparse<true, 0>", "lbm2p_fast::k2p_colour_sparse"]
OPS = ["UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "FFMA", "FADD", "FMUL", "MUFU", "UTMALDG", "UTCMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names).encode(), stdout=subprocess.PIPE, check=True).stdout.decode()
    return out.splitlines()


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    usage = {}
    for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
        usage[m.group(1)] = {"registers": int(m.group(2)), "stack_bytes": int(m.group(3)), "shared_bytes": int(m.group(4))}
    names = list(usage)
    plain = dict(zip(demangle(names), names))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    blocks = {}
    for chunk in sass.split("Function : ")[1:]:
        fn, _, body = chunk.partition("\n")
        blocks[fn.strip()] = body
    out = {"library": os.path.relpath(LIB, ROOT), "arch": arch, "kernels": []}
    for want in HOT:
        hits = [p for p in plain if want in p.replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", "")]
        if not hits:
            out["kernels"].append({"kernel": want, "missing": True})
            continue
        mangled = plain[hits[0]]
        body = blocks.get(mangled, "")
        ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", body, flags=re.M)
        rec = {"kernel": want, "mangled": mangled, "sass_instructions": len(ins)}
        rec.update(usage[mangled])
        rec.update({op: sum(1 for i in ins if i == op or i.startswith(op + ".")) for op in OPS})
        out["kernels"].append(rec)
    txt = json.dumps(out, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
