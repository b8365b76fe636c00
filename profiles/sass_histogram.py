"""SASS opcode histogram per kernel of the product library (evidence for TMA / mbarrier / packed-f32 use and for the absence of
what the numerics contract forbids): python profiles/sass_histogram.py > profiles/sass_histograms_r02.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UBLKCP", "SYNCS", "FMUL2", "FFMA2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "I2F", "I2FP", "F2I", "LDS", "STS", "LDG", "STG", "ATOMS", "REDUX",
         "BAR", "IMAD", "IADD3", "LOP3", "PRMT", "SHF", "ISETP", "FSETP", "FSET", "FMNMX", "SEL", "DADD", "DMUL", "DFMA", "HMMA", "UTCHMMA", "TEX", "TLD"]


def main():
    for obj in sorted(glob.glob(os.path.join(ROOT, "bevy_terrain_b200", "csrc", "*.o"))):
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        if "Function :" not in out:
            continue
        name, hist = None, None
        kernels = []
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name, hist = m.group(1), collections.Counter()
                kernels.append((name, hist))
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and hist is not None:
                hist[m.group(1)] += 1
        print(f"== {os.path.basename(obj)}")
        for name, hist in kernels:
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
            total = sum(hist.values())
            watch = " ".join(f"{k}={hist[k]}" for k in WATCH if hist[k])
            print(f"  {demangled[:90]}\n      {total} instructions: {watch}")


if __name__ == "__main__":
    main()
