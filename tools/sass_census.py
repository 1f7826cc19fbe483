"""SASS opcode census of libwsi_b200.so: which kernels carry tcgen05 (UTC*MMA), TMEM loads (LDTM), TMA (UTMALDG / UBLKCP),
FP64 ... -> profiles/rNN_sass_census.txt.  Runs without a GPU (cuobjdump only).

    python tools/sass_census.py profiles/r02_sass_census.txt
"""
import collections
import re
import subprocess
import sys

LIB = "wsi_segmentation_pipeline_b200/libwsi_b200.so"
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "LDG", "STG", "LDS", "STS", "DADD", "DFMA", "F2F", "HMMA"]


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void wsi::", "").replace("wsi::", "")
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            base = op.split(".")[0]
            cur[base] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
            if op.startswith("UTMALDG"):
                cur["UTMALDG"] += 0
    tot = collections.Counter()
    with open(out, "w") as f:
        f.write(f"# SASS census of {LIB} (cuobjdump -sass, sm_100a); counts are static instructions per kernel\n")
        f.write(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{k:>12s}" for k in KEYS) + "\n")
        for name, c in per.items():
            f.write(f"{name[:78]:78s} {c['total']:7d} " + " ".join(f"{c.get(k, 0):12d}" for k in KEYS) + "\n")
            for k in KEYS:
                tot[k] += c.get(k, 0)
            tot["total"] += c["total"]
        f.write(f"{'TOTAL (' + str(len(per)) + ' kernels)':78s} {tot['total']:7d} " + " ".join(f"{tot[k]:12d}" for k in KEYS) + "\n")
    print(open(out).read()[-1500:])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_sass_census.txt")
