#!/usr/bin/env python
"""SASS evidence for the hot kernels of libvmat_b200.so (no GPU needed): per kernel, how many TMA bulk copies
(UBLKCP = cp.async.bulk), mbarrier operations (SYNCS), programmatic-dependent-launch instructions, shared and
global loads, FMAs and atomics the compiled sm_100a code contains - and that no floating-point atomic and no
tensor-core instruction is among them (the path is a 0.25 flop/byte SpMV pair: HBM-bound).
Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vmatwpencode_b200", "libvmat_b200.so")
WANT = [r"k1_dose_penalty<float, float, true, 16, 1>", r"k2_beamlet_grad<float, float, 4, 1>",      # byte deltas: the default from 5 M nonzeros on
        r"k1_dose_penalty<float, float, true, 16, 2>", r"k2_beamlet_grad<float, float, 4, 2>", r"k3_reduce<float>",
        r"k3_reduce_exchange<float>", r"kf_eval<float, float, 2>", r"k1_dose_penalty<float, double, true, 16, 2>",
        r"k2_beamlet_grad<float, double, 4, 2>", r"k5_price_dp", r"k6_kmedoid_cost_tiled<3>"]
GROUPS = [("UBLKCP (cp.async.bulk, TMA)", r"\bUBLKCP"), ("UTMALDG (tensor-map TMA)", r"\bUTMALDG"),
          ("SYNCS (mbarrier)", r"\bSYNCS"), ("ACQBULK / bulk wait", r"\bACQBULK|\bUBLKWAIT|\bDEPBAR"),
          ("PDL (ACQBULK/griddepcontrol: PREEXIT, ACQGRID)", r"\bPREEXIT|\bACQGRID"),
          ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("LDG", r"\bLDG"), ("STG", r"\bSTG"),
          ("FFMA", r"\bFFMA"), ("DFMA", r"\bDFMA"), ("DADD/DMUL", r"\bDADD|\bDMUL"),
          ("integer atomics (ATOMG/REDG/ATOMS)", r"\b(ATOMG|REDG|ATOMS|ATOM)\b[^;]*"),
          ("floating-point atomics", r"\b(ATOMG|REDG|ATOMS|ATOM)[.\w]*\.(F32|F64|F16)"),
          ("tensor core (UTC*MMA / HMMA)", r"\bUTC\w*MMA|\bHMMA|\bIMMA|\bDMMA"), ("MEMBAR / FENCE", r"\bMEMBAR|\bFENCE"),
          ("SHFL", r"\bSHFL")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.search(r"^\s+/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line)
    demangled = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"library: vmatwpencode_b200/libvmat_b200.so   cubins: {sorted(set(arch))}   kernels: {len(funcs)}")
    print("counts of SASS instructions per kernel (cuobjdump -sass)\n")
    for want in WANT:
        for mangled, pretty in zip(funcs, demangled):
            if want in pretty and (want.startswith("k3_reduce<") is False or "exchange" not in pretty):
                body = funcs[mangled]
                print(pretty.split("(")[0].replace("void vmat::", ""))
                print(f"    instructions: {len(body)}")
                for label, pat in GROUPS:
                    n = sum(1 for ln in body if re.search(pat, ln))
                    if n or label.startswith(("floating", "tensor", "UTMALDG")):
                        print(f"    {label:<52s} {n}")
                print()
                break
    tot_fp_atom = sum(1 for b in funcs.values() for ln in b if re.search(GROUPS[13][1], ln))
    tot_tc = sum(1 for b in funcs.values() for ln in b if re.search(GROUPS[14][1], ln))
    print(f"whole library: floating-point atomics {tot_fp_atom}, tensor-core instructions {tot_tc}")


if __name__ == "__main__":
    main()
