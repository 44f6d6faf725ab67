"""SASS evidence that the hot kernels of libchalamet_b200.so are Blackwell-native (no GPU needed: cuobjdump reads the cubin).

    python tools/sass_evidence.py > profiles/r2_sass_evidence.txt

Per kernel of interest: the count of every tensor-core / TMA / bulk-copy / cluster mnemonic (the list of
/opt/skills/guides/B200_PROFILING.md) and the first occurrence of each with its neighbouring instructions."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "chalametpir_b200", "libchalamet_b200.so")
KERNELS = ["gemm_tc_pair_kernelILi2", "gemm_tc_pair_kernelILi1", "gemm_tc_kernelILi2", "gemm_tc_kernelILi1", "respond_ring_kernelILi9ELi4ELb0", "respond_ring_kernelILi10ELi4ELb0", "pull_slices_kernel",
           "reduce_parts_kernelI5uint4", "solve_columns_kernelILi3", "encode_rows_kernel", "split_transpose_bILi2", "pack_kernelILi9", "vec_x_mat_kernel", "expand_kernel"]
MNEMONICS = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "SYNCS", "UCGABAR", "ATOMG", "REDG", "RED.E",
             "LDG.E.128", "LDS.128", "SHFL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout.splitlines()
    funcs, cur = collections.OrderedDict(), None
    for line in sass:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({len(funcs)} kernels, sm_100a)")
    for want in KERNELS:
        for name, body in funcs.items():
            if want not in name:
                continue
            ins = [l for l in body if re.search(r"/\*[0-9a-f]{4}\*/", l)]
            print(f"\n## {name}\n   {len(ins)} instructions")
            for mn in MNEMONICS:
                hits = [i for i, l in enumerate(ins) if re.search(r"\b" + re.escape(mn), l)]
                if not hits:
                    continue
                variants = collections.Counter(re.search(r"\b(" + re.escape(mn) + r"[.\w]*)", ins[i]).group(1) for i in hits)
                print(f"   {mn:10s} x{len(hits):4d}   " + ", ".join(f"{k} x{v}" for k, v in variants.most_common(6)))
            for mn in ("UTCIMMA", "UTMALDG", "UTCBAR", "LDTM", "UBLKCP", "UCGABAR"):
                hits = [i for i, l in enumerate(ins) if re.search(r"\b" + mn, l)]
                if hits:
                    i = hits[0]
                    print(f"   first {mn}:")
                    for l in ins[max(0, i - 1): i + 2]:
                        print("      " + re.sub(r"\s+", " ", l.strip())[:170])
            break


if __name__ == "__main__":
    sys.exit(main())
