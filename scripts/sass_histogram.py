"""SASS opcode histograms of the hot kernels (cuobjdump -sass on the built objects; runs without a GPU):
python scripts/sass_histogram.py > profiles/r02_sass_histograms.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tokamak-zk-evm_b200", "build")
KERNELS = [("msm.o", "k_tree_apply"), ("msm.o", "k_tree_fwd"), ("msm.o", "k_accumulateILb0"), ("msm.o", "k_accumulateILb1"), ("msm.o", "k_bucket_seg"),
           ("ntt.o", "k_ntt_passILb0ELb0"), ("ntt.o", "k_ntt_passILb1ELb0"), ("poly.o", "k_polyexpr"), ("poly.o", "k_lincomb"), ("api.o", "k_microbenchILi6E")]


def ptxas_info(obj, pat):
    log = open(os.path.join(BUILD, obj.replace(".o", ".ptxas.log"))).read().splitlines()
    for i, l in enumerate(log):
        if "Compiling entry function" in l and pat in l:
            return " | ".join(x.strip() for x in log[i + 1:i + 3])
    return "?"


for obj, pat in KERNELS:
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    cur, hist, total = None, collections.Counter(), 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and pat in cur:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", line)
            if m:
                hist[m.group(1)] += 1
                total += 1
    wide = sum(v for k, v in hist.items() if k.startswith("IMAD.WIDE"))
    fma_other = sum(v for k, v in hist.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE"))
    alu = sum(v for k, v in hist.items() if k.split(".")[0] in ("IADD3", "LOP3", "SEL", "SHF", "VIADD", "ISETP", "PRMT", "LEA", "MOV", "PLOP3", "IABS", "FLO", "BREV", "POPC"))
    print(f"== {pat} ({obj})  {total} instructions; ptxas: {ptxas_info(obj, pat)}")
    print(f"   IMAD.WIDE* {wide}  other IMAD (fma pipe) {fma_other}  ALU-pipe integer {alu}  LDG/STG {hist['LDG.E.128.CONSTANT'] + sum(v for k, v in hist.items() if k.startswith('LDG') or k.startswith('STG'))}"
          f"  LDS/STS {sum(v for k, v in hist.items() if k.startswith('LDS') or k.startswith('STS'))}  BAR {sum(v for k, v in hist.items() if k.startswith('BAR'))}"
          f"  tensor/TMA (UTC*MMA, UTMA*, LDTM/STTM) {sum(v for k, v in hist.items() if k.startswith(('UTC', 'UTMA', 'LDTM', 'STTM', 'HMMA')))}")
    print("   " + ", ".join(f"{k} {v}" for k, v in hist.most_common(14)))
