#!/usr/bin/env python
"""Write profiles/sass_evidence.txt: per kernel, counts of the SASS instructions that identify the
hardware path (tcgen05 MMA / TMA / TMEM loads / DMMA / cp.async / fences and atomics of the fused
finalize), plus `ptxas -v` of the hot kernels.  Needs only the built objects (no GPU):

    make -C matfree_b200/csrc && python tools/sass_evidence.py
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "matfree_b200", "csrc", "build")
PREFIXES = ["UTCHMMA", "UTCMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "LDTM", "STTM", "DMMA", "HMMA", "LDGSTS",
            "LDG.E.128", "STG.E.128", "LDS.128", "SHFL", "SYNCS", "MEMBAR", "ATOMG", "REDG", "PREFETCH", "CCTL"]
HOT = ("gemm_tf32x3", "gemm_dmma_kernel<false, 64", "spmm_csr_kernel<float, 4, 256, 5, false, true, false",
       "spmm_tma_kernel<float, 4, 256, 5, 16, true, false", "spmm_walk_kernel<float, 4, 256, 5, 16, true", "spmm_row_thread_kernel<float",
       "spmm_tma_kernel<float, 4, 256, 7, 16, true, true", "spmm_csr_kernel<float, 4, 256, 7, false, true, false, true",
       "lanczos_update_kernel<float, 4, true", "reorth_dots_all_kernel<float, 4", "cgs_update_dots",
       "reorth_update_kernel<float, 4, true", "probe_gen_signs_kernel<float, 4", "halo_push",
       "tridiag_ql_kernel<float, false")


def demangle(name):
    for tool in ("cu++filt", "c++filt"):
        try:
            r = subprocess.run([tool, name], capture_output=True, text=True)
        except FileNotFoundError:
            continue
        if r.returncode == 0 and r.stdout.strip() and r.stdout.strip() != name:
            return r.stdout.strip()
    return name


def short(n):
    n = re.sub(r"^void ", "", n).replace("(anonymous namespace)::", "").replace("mf::", "")
    depth = 0
    for i, ch in enumerate(n):
        depth += ch == "<"
        depth -= ch == ">"
        if ch == "(" and depth == 0:
            return n[:i]
    return n


def main():
    out = ["# SASS / ptxas evidence: `cuobjdump -sass` of matfree_b200/csrc/build/*.o (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)",
           "# per kernel, counts of the instructions that identify the hardware path:",
           "#   UTCHMMA = tcgen05.mma (kind::tf32 here), UTMALDG = TMA tensor load (cp.async.bulk.tensor), UBLKCP = TMA bulk copy (cp.async.bulk), LDTM = tcgen05.ld (TMEM -> registers),",
           "#   UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, DMMA = FP64 tensor-core MMA, LDGSTS = cp.async, MEMBAR/ATOMG = the fused finalize / peer handshakes",
           ""]
    for obj in sorted(glob.glob(os.path.join(BUILD, "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur, counts = None, collections.OrderedDict()
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = short(demangle(m.group(1)))
                counts[cur] = collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if cur is None or not m:
                continue
            for k in PREFIXES:
                if m.group(1).startswith(k):
                    counts[cur][k] += 1
                    break
        out.append(f"== {os.path.basename(obj)}")
        out += [f"  {fn[:100]:100s} " + " ".join(f"{k}={v}" for k, v in c.items()) for fn, c in counts.items() if c]
        out.append("")
    out.append("# ptxas -v (registers / spills / shared memory) of the hot kernels")
    for log in sorted(glob.glob(os.path.join(BUILD, "*.ptxas.log"))):
        txt = open(log).read().splitlines()
        for i, l in enumerate(txt):
            m = re.search(r"Compiling entry function '(\S+)'", l)
            if not m:
                continue
            name = short(demangle(m.group(1)))
            if any(k in name for k in HOT):
                info = " | ".join(x.strip().replace("ptxas info    : ", "") for x in txt[i + 1:i + 4]
                                  if "registers" in x or "spill" in x)
                out.append(f"  {name[:95]:95s} {info}")
    with open(os.path.join(ROOT, "profiles", "sass_evidence.txt"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
