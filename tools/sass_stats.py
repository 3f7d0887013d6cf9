"""Static SASS statistics per kernel of a .so / executable / cubin (development tool):
instruction counts by class, as a first look before spending GPU time.
usage: python tools/sass_stats.py <file> [name-filter]"""
import collections
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0].strip()
        dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if filt and filt not in dn:
            continue
        ops = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops.append(m.group(1))
        c = collections.Counter(ops)
        fp64 = c["DFMA"] + c["DMUL"] + c["DADD"] + c["DSETP"]
        print("%-100s total %5d fp64 %4d (DFMA %d DMUL %d DADD %d DSETP %d) MUFU %d BSSY %d BRA %d FSEL %d IMAD %d MOV %d "
              "LDS %d STS %d LDG %d STG %d LDL %d STL %d SHFL %d" %
              (dn[:100], len(ops), fp64, c["DFMA"], c["DMUL"], c["DADD"], c["DSETP"], c["MUFU"], c["BSSY"], c["BRA"],
               c["FSEL"], c["IMAD"], c["MOV"], c["LDS"], c["STS"], c["LDG"], c["STG"], c["LDL"], c["STL"], c["SHFL"]))


if __name__ == "__main__":
    main()
